// proto_lzcol — TEST-SIDE PROTOTYPE (not part of the library): a match finder for the text-like NAF streams that has no
// sequential parse at all.  The greedy hash walk of the shipped thread-per-block path (nafz::zlz_find) is 6.5 of the LZ
// kernel's 9.3 ms; on these streams nearly every match it finds is "the same column of the previous record"
// (ids `SRR1.1234567\0`, comments `1234567/1\0`) or "the previous unit" (lengths: 4-byte units).  So candidates need no hash
// table: per byte, the offset is the distance between the starts of its record and the previous one (a max-scan over the
// positions of '\0'), else 4; a byte "matches" when it equals the byte that far back; maximal runs of matching bytes with one
// offset are the matches (a segmented scan), runs long enough become sequences (a compaction).  Every stage is a map or a
// scan over the block -- what a warp or a CTA does in O(log n) steps -- and is written here as plain loops.  The sequences go
// to the SAME literal / sequence coder as the shipped path (nafz::zlz_emit_block), so the frames are checked by the same
// decoders (libzstd, the oracle).
//   proto_lzcol IN OUT.zst BLOCK_SIZE
#include "lzcol.hpp"
#include <cstdio>
#include <cstdlib>
#include <vector>
using namespace nafz;

int main(int argc, char **argv)
{
    if (argc < 4) return 2;
    FILE *f = fopen(argv[1], "rb"); if (!f) return 2;
    std::vector<u8> in; u8 buf[65536]; size_t k;
    while ((k = fread(buf, 1, sizeof buf, f)) > 0) in.insert(in.end(), buf, buf + k);
    fclose(f);
    const u32 bs = (u32)atoi(argv[3]);
    if (bs < 16 || bs > ZLZ_MAX_BLOCK) return 2;
    const u32 max_seq = bs / 4;
    std::vector<u8> out = { 0x28, 0xB5, 0x2F, 0xFD, 0x00, (u8)((17 - 10) << 3) };
    std::vector<u16> sll(max_seq), sml(max_seq), sov(max_seq), spos(1280);
    std::vector<u8> lit(bs + 16), tsym(512), codes(3 * max_seq), slot(bs + 512);
    const size_t n = in.size(), nblk = n ? (n + bs - 1) / bs : 1;
    u64 nseq = 0, matched = 0;
    for (size_t b = 0; b < nblk; b++) {
        const u32 len = (u32)(n - b * bs < bs ? n - b * bs : bs);
        const u8 *src = in.data() + b * bs;
        ZLzSeqs S{sll.data(), sml.data(), sov.data(), 0};
        ZLzWork W{spos.data(), tsym.data(), codes.data()};
        bool rle = false; u32 cs = 0;
        if (len) { u32 i = 1; while (i < len && src[i] == src[0]) i++; rle = i == len; }
        if (!rle && len >= 16) {
            const u32 nlit = find_columns(src, len, lit.data(), S, max_seq);
            nseq += S.n; matched += len - nlit;
            cs = zlz_emit_block(len, lit.data(), nlit, S, max_seq, W, slot.data(), bs + 512);
        }
        const u32 last = b + 1 == nblk, type = cs ? 2 : (rle ? 1 : 0), size_field = type == 2 ? cs : len;
        const u32 bh = last | (type << 1) | (size_field << 3);
        out.push_back((u8)bh); out.push_back((u8)(bh >> 8)); out.push_back((u8)(bh >> 16));
        if (type == 2) out.insert(out.end(), slot.begin(), slot.begin() + cs);
        else if (type == 1) out.push_back(src[0]);
        else out.insert(out.end(), src, src + len);
    }
    FILE *o = fopen(argv[2], "wb"); if (!o) return 2;
    fwrite(out.data(), 1, out.size(), o); fclose(o);
    printf("in=%zu out=%zu blocks=%zu seqs=%llu matched=%llu\n", n, out.size(), nblk, (unsigned long long)nseq, (unsigned long long)matched);
    return 0;
}
