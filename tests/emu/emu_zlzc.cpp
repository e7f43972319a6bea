// emu_zlzc — TEST-ONLY: runs naf_b200/csrc/zstd_lzc_hd.cuh (the column match finder's CTA phases, the per-stream tables, the
// block coder against them) on the CPU, thread after thread and phase after phase in the order k_zlc_find / k_zlc_define /
// k_zlc_finish run them, and lays the blocks out as one frame like k_zenc_gather does.
//   emu_zlzc IN OUT.zst BLOCK_SIZE [SEQS.txt]       (SEQS.txt: every block's sequences, for comparison with tests/emu/lzcol.hpp)
#include "../../naf_b200/csrc/zstd_lzc_bytes_hd.cuh"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
using namespace nafz;
// the finder's phases in one of their two formulations: -DZLC_BYTES = byte loops (nafz::zlcb, what a level selects), else bit masks
#ifdef ZLC_BYTES
namespace fin = nafz::zlcb;
#else
namespace fin = nafz;
#endif

int main(int argc, char **argv)
{
    if (argc < 4) return 2;
    FILE *f = fopen(argv[1], "rb"); if (!f) return 2;
    std::vector<u8> in; u8 buf[65536]; size_t k;
    while ((k = fread(buf, 1, sizeof buf, f)) > 0) in.insert(in.end(), buf, buf + k);
    fclose(f);
    const u32 bs = (u32)atoi(argv[3]);
    if (bs < 16 || bs > ZLC_MAX) return 2;
    FILE *seqs = argc > 4 ? fopen(argv[4], "w") : nullptr;
    const size_t n = in.size(), nblk = n ? (n + bs - 1) / bs : 1;
    const u32 wstride = zlc_work_bytes(bs), sstride = bs + 512;
    std::vector<ZlcBlk> info(nblk);
    std::vector<u8> work(nblk * (size_t)wstride + 64, 0xEE), slots(nblk * (size_t)sstride + 64, 0xEE);
    std::vector<u32> counts(ZLC_NBINS, 0);
    ZlcStreamView V{in.data(), n, bs, (u32)nblk, info.data(), work.data(), wstride, slots.data(), sstride};
    fin::ZlcSh *shp = new fin::ZlcSh; fin::ZlcSh &sh = *shp;

    // ---- k_zlc_find: one CTA per block, thread k = chunk k.  Between two barriers the threads of a CTA run in no particular order:
    // ZLC_ORDER=<seed> in the environment shuffles the order of the threads of every phase (0 / unset: ascending or descending as
    // written below); a phase in which one thread read what another one writes would then give different frames for different seeds.
    const char *ord_env = getenv("ZLC_ORDER");
    unsigned long long ord_state = ord_env ? strtoull(ord_env, nullptr, 10) : 0;
    std::vector<u32> perm(ZLC_NCH);
    auto order = [&](u32 count, bool descending) -> const std::vector<u32> & {
        for (u32 i = 0; i < count; i++) perm[i] = descending ? count - 1 - i : i;
        if (ord_state) for (u32 i = count; i > 1; i--) {
            ord_state = ord_state * 6364136223846793005ull + 1442695040888963407ull;
            const u32 j = (u32)((ord_state >> 33) % i);
            const u32 t = perm[i - 1]; perm[i - 1] = perm[j]; perm[j] = t;
        }
        return perm;
    };
    for (size_t b = 0; b < nblk; b++) {
        const u32 len = V.len((u32)b);
        const u8 *src = in.data() + b * bs;
        memset(&sh, 0xDD, sizeof sh);
        for (u32 i = 0; i < len; i++) sh.src_[fin::zlc_ix(i)] = src[i];
        memset(sh.hist, 0, sizeof sh.hist);
        sh.n = len; sh.nch = (len + ZLC_CH - 1) / ZLC_CH; sh.rle_break = 0; sh.lastend = 0;
        const u32 nch = sh.nch;
        { const std::vector<u32> P = order(nch, false); for (u32 i = 0; i < nch; i++) fin::zlc_zeros(sh, P[i]); }
        ZlcBlk &I = info[b];
        I.nseq = I.nlit = 0; I.parsed = I.rle = I.conv = I.pad = 0;
        if (len == 0 || !sh.rle_break || len < 16) { I.rle = len && !sh.rle_break; continue; }
        { const std::vector<u32> P = order(nch, false); for (u32 i = 0; i < nch; i++) fin::zlc_columns(sh, P[i]); }
        { const std::vector<u32> P = order(nch, false); for (u32 i = 0; i < nch; i++) fin::zlc_breaks(sh, P[i]); }
        { const std::vector<u32> P = order(nch, true); for (u32 i = 0; i < nch; i++) fin::zlc_choose(sh, P[i]); }
        { const std::vector<u32> P = order(nch, false); for (u32 i = 0; i < nch; i++) fin::zlc_breaks_d(sh, P[i]); }
        { const std::vector<u32> P = order(nch, true); for (u32 i = 0; i < nch; i++) fin::zlc_count(sh, P[i]); }
        fin::zlc_scan_serial(sh);
        const bool sampled = b % ZLC_SAMPLE == 0;
        ZlcWork K = zlc_work(work.data() + b * wstride, bs);
        { const std::vector<u32> P = order(nch, true); for (u32 i = 0; i < nch; i++) fin::zlc_emit_seqs(sh, P[i], K.S, K.lit, sampled); }
        for (u32 t = 0; t < 7; t++) fin::zlc_emit_tail(sh, t, 7, K.lit, sampled);
        I.nseq = sh.nseq; I.nlit = len - sh.mltot; I.parsed = 1;
        if (sampled) { fin::zlc_count_offsets(sh); for (u32 i = 0; i < ZLC_NBINS; i++) counts[i] += sh.hist[i]; }
        if (seqs) {
            fprintf(seqs, "block %zu nlit %u nseq %u\n", b, I.nlit, I.nseq);
            for (u32 i = 0; i < I.nseq; i++) fprintf(seqs, "%u %u %u\n", K.S.ll[i], K.S.ml[i], K.S.ov[i]);
            fwrite(K.lit, 1, I.nlit, seqs); fputc('\n', seqs);
        }
    }
    if (seqs) fclose(seqs);
    // ---- k_zlc_define: one thread per stream
    ZlcTables *T = new ZlcTables; u32 def_fail = 0;
    static_assert(sizeof(ZlcTables) % 16 == 0, "k_zlc_finish copies the tables in 16-byte pieces");
    zlc_define(V, counts.data(), *T);
    // ---- k_zlc_finish: one thread per block; then the frame as k_zenc_gather writes it
    std::vector<u8> out = { 0x28, 0xB5, 0x2F, 0xFD, 0x00, (u8)((17 - 10) << 3) };
    size_t n_comp = 0, n_rle = 0;
    std::vector<u32> types(nblk), sizes(nblk);
    for (size_t b = nblk; b-- > 0;) zlc_finish_block(V, (u32)b, *T, &def_fail, &types[b], &sizes[b]);
    for (size_t b = 0; b < nblk; b++) zlc_finish_block_own(V, (u32)b, *T, def_fail, &types[b], &sizes[b]);      // k_zlc_finish_own
    for (size_t b = 0; b < nblk; b++) {
        const u32 len = V.len((u32)b), last = b + 1 == nblk, type = types[b], size_field = type == 1 ? len : sizes[b];
        const u32 bh = last | (type << 1) | (size_field << 3);
        out.push_back((u8)bh); out.push_back((u8)(bh >> 8)); out.push_back((u8)(bh >> 16));
        const u8 *from = type == 0 ? in.data() + b * bs : slots.data() + b * sstride;
        out.insert(out.end(), from, from + sizes[b]);
        n_comp += type == 2; n_rle += type == 1;
    }
    FILE *o = fopen(argv[2], "wb"); if (!o) return 2;
    fwrite(out.data(), 1, out.size(), o); fclose(o);
    printf("in=%zu out=%zu blocks=%zu compressed=%zu rle=%zu shared=%u fdef=%d def_fail=%u\n", n, out.size(), nblk, n_comp, n_rle, T->ok, (int)T->fdef, def_fail);
    return 0;
}
