// emu_zenc — TEST-ONLY: runs naf_b200/csrc/zstd_enc_hd.cuh's block encoder (LZ77 + Huffman + FSE) on the CPU and lays the
// blocks out as one frame exactly like k_zenc_gather does (FHD 0x00, window byte, 3-byte block headers).
//   emu_zenc IN OUT.zst BLOCK_SIZE USE_LZ [HSTRIDE]
#include "../../naf_b200/csrc/zstd_enc_hd.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>
using namespace nafz;
int main(int argc, char **argv)
{
    if (argc < 5) return 2;
    FILE *f = fopen(argv[1], "rb"); if (!f) return 2;
    std::vector<u8> in; u8 buf[65536]; size_t k;
    while ((k = fread(buf, 1, sizeof buf, f)) > 0) in.insert(in.end(), buf, buf + k);
    fclose(f);
    const u32 bs = (u32)atoi(argv[3]); const bool lz = atoi(argv[4]) != 0;
    const u32 hstride = argc > 5 ? (u32)atoi(argv[5]) : 1;
    if (bs < 16 || bs > ZLZ_MAX_BLOCK) return 2;
    const u32 max_seq = bs / 4;
    std::vector<u8> out = { 0x28, 0xB5, 0x2F, 0xFD, 0x00, (u8)((17 - 10) << 3) };
    std::vector<u16> htab((size_t)(1u << ZLZ_HLOG) * hstride), sll(max_seq), sml(max_seq), sov(max_seq), spos(1280);
    std::vector<u8> lit(bs + 16), tsym(512), codes(3 * max_seq), slot(bs + 512);
    const size_t n = in.size(), nblk = n ? (n + bs - 1) / bs : 1;
    size_t n_comp = 0, n_rle = 0, n_seq = 0;
    for (size_t b = 0; b < nblk; b++) {
        const u32 len = (u32)(n - b * bs < bs ? n - b * bs : bs);
        const u8 *src = in.data() + b * bs;
        ZLzSeqs S{sll.data(), sml.data(), sov.data(), 0};
        ZLzWork W{spos.data(), tsym.data(), codes.data()};
        bool rle = false;
        const u32 cs = zlz_encode_block(src, len, lz, htab.data(), hstride, lit.data(), S, max_seq, W, slot.data(), bs + 512, &rle);
        const u32 last = b + 1 == nblk;
        const u32 type = cs ? 2 : (rle ? 1 : 0), size_field = type == 2 ? cs : len;
        const u32 bh = last | (type << 1) | (size_field << 3);
        out.push_back((u8)bh); out.push_back((u8)(bh >> 8)); out.push_back((u8)(bh >> 16));
        if (type == 2) { out.insert(out.end(), slot.begin(), slot.begin() + cs); n_comp++; }
        else if (type == 1) { out.push_back(src[0]); n_rle++; }
        else out.insert(out.end(), src, src + len);
        (void)n_seq;
    }
    FILE *o = fopen(argv[2], "wb"); if (!o) return 2;
    fwrite(out.data(), 1, out.size(), o); fclose(o);
    fprintf(stderr, "in=%zu out=%zu blocks=%zu compressed=%zu rle=%zu\n", n, out.size(), nblk, n_comp, n_rle);
    return 0;
}
