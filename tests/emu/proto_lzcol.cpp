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
#include "../../naf_b200/csrc/zstd_enc_hd.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>
using namespace nafz;

static u32 find_columns(const u8 *src, u32 n, u8 *lit, ZLzSeqs &S, u32 max_seq)
{
    std::vector<u32> rs(n), d(n), run(n);
    // scan 1: start of the record a byte belongs to (records end with '\0')
    { u32 cur = 0; for (u32 p = 0; p < n; p++) { rs[p] = cur; if (src[p] == 0) cur = p + 1; } }
    // map: two candidate offsets per byte -- the same column of the previous record, and 4 (the previous length unit) -- and
    // whether the byte matches at each
    std::vector<u32> oc(n), of(n), lc(n), lf(n);
    for (u32 p = 0; p < n; p++) {
        oc[p] = 0;
        if (rs[p] > 0) { const u32 dcol = rs[p] - rs[rs[p] - 1]; if (dcol <= p && src[p] == src[p - dcol]) oc[p] = dcol; }
        of[p] = p >= 4 && src[p] == src[p - 4] ? 4u : 0u;
    }
    // two segmented scans (forward: position in the run; backward: the run's length): how long is the run a byte is in, per candidate
    auto run_len = [&](const std::vector<u32> &o, std::vector<u32> &len) {
        std::vector<u32> pos(n);
        for (u32 p = 0; p < n; p++) pos[p] = o[p] ? ((p && o[p - 1] == o[p]) ? pos[p - 1] + 1 : 1) : 0;
        for (u32 p = n; p-- > 0;) len[p] = o[p] ? ((p + 1 < n && o[p + 1] == o[p]) ? len[p + 1] : pos[p]) : 0;
    };
    run_len(oc, lc); run_len(of, lf);
    // map: a byte takes the candidate whose run around it is longer
    for (u32 p = 0; p < n; p++) d[p] = lf[p] > lc[p] ? of[p] : oc[p];
    // segmented scan: position inside a run of bytes that match at one offset (0: no match here)
    for (u32 p = 0; p < n; p++) run[p] = d[p] ? ((p && d[p - 1] == d[p]) ? run[p - 1] + 1 : 1) : 0;
    // compaction: runs that are long enough become matches, in order; literals are what lies between them
    ZLzRep rep; rep.r[0] = rep.r[1] = rep.r[2] = 0; rep.k = 0;
    u32 anchor = 0, nlit = 0; S.n = 0;
    for (u32 p = 0; p < n && S.n < max_seq; p++) {
        const bool run_ends = run[p] && (p + 1 == n || d[p + 1] != d[p]);
        if (!run_ends) continue;
        const u32 ml = run[p], start = p + 1 - ml, off = d[p];
        if (start < anchor) continue;
        const u32 ll = start - anchor;
        const bool is_rep = rep.k && off == rep.r[0] && ll > 0;
        if (ml < 4 || (ml < 5 && !is_rep)) continue;           // same rule as the serial parse: a 4-byte match at a new offset does not pay
        for (u32 i = 0; i < ll; i++) lit[nlit + i] = src[anchor + i];
        nlit += ll;
        S.ll[S.n] = (u16)ll; S.ml[S.n] = (u16)ml; S.ov[S.n] = (u16)rep.code(off, ll); S.n++;
        anchor = p + 1;
    }
    for (u32 i = anchor; i < n; i++) lit[nlit++] = src[i];
    return nlit;
}

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
