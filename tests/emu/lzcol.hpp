// lzcol.hpp — TEST-SIDE: the column match finder of proto_lzcol.cpp (see there), shared with proto_shared.cpp.
#pragma once
#include "../../naf_b200/csrc/zstd_enc_hd.cuh"
#include <vector>
namespace nafz {
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
        if (ml < 5) continue;                                  // a 4-byte match pays only at a repeated offset, which would make taking it depend on the
                                                               // matches before it; dropping them all costs under 0.3 % on the streams this is for
        for (u32 i = 0; i < ll; i++) lit[nlit + i] = src[anchor + i];
        nlit += ll;
        S.ll[S.n] = (u16)ll; S.ml[S.n] = (u16)ml; S.ov[S.n] = (u16)rep.code(off, ll); S.n++;
        anchor = p + 1;
    }
    for (u32 i = anchor; i < n; i++) lit[nlit++] = src[i];
    return nlit;
}

}  // namespace nafz
