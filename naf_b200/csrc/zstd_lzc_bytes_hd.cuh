// zstd_lzc_bytes_hd.cuh — the column match finder of zstd_lzc_hd.cuh in its first formulation: every phase a loop over the chunk's
// 32 bytes in shared memory (per-byte chosen offsets in an array of their own).  Same phases, same summaries, same sequences --
// both formulations are compared byte for byte with the serial restatement (tests/emu/lzcol.hpp) by tests/test_emu_zenc.py.
// This is the one that ran on a B200 (profiles/r4a, r4b: frames byte-identical to the CPU emulation) and therefore the one a
// level >= 2 selects; the bit-mask formulation (zstd_lzc_hd.cuh, NAFGPU_LZ=b) replaced its later phases after the round's last GPU
// call and is measured next to it by bench.py's level2 sub-record.
#pragma once
#include "zstd_lzc_hd.cuh"

namespace nafz { namespace zlcb {

// Thread k walks bytes 32 k .. 32 k + 31: with the arrays laid out plainly the 32 lanes of a warp would sit 32 bytes (8 banks) or
// 32 u16 (16 banks) apart and every access would be an 8- or 16-way bank conflict.  One element of padding per chunk (33-element
// pitch) puts the lanes of a warp into 32 different banks.
static const u32 ZLC_PITCHED = ZLC_MAX + ZLC_NCH;
HD u32 zlc_ix(u32 p) { return p + (p >> 5); }
struct ZlcSh {                        // shared memory of one CTA = one block of at most ZLC_MAX bytes
    u8  src_[ZLC_PITCHED + 24];
    u16 oc_[ZLC_PITCHED];             // column candidate: its offset where the byte matches there, else 0 (later: offsets / literal lengths of the sequences)
    u16 d_[ZLC_PITCHED];              // chosen offset per byte, 0 = literal
    HD u8 src(u32 p) const { return src_[zlc_ix(p)]; }
    HD u16 oc(u32 p) const { return oc_[zlc_ix(p)]; }
    HD u16 d(u32 p) const { return d_[zlc_ix(p)]; }
    u16 z1[ZLC_NCH], z2[ZLC_NCH];     // last / second-last '\0' of a chunk
    u16 lbc[ZLC_NCH], fbc[ZLC_NCH];   // last / first position of a chunk at which a run of the column candidate does not continue
    u16 lbf[ZLC_NCH], fbf[ZLC_NCH];   // same, candidate 4
    u16 fbd[ZLC_NCH];                 // first position of a chunk at which a run of d does not continue
    u16 cnt[ZLC_NCH], mls[ZLC_NCH], lend[ZLC_NCH];    // matches starting in a chunk: how many, their lengths added up, where the last one ends
    u16 ibase[ZLC_NCH], mbase[ZLC_NCH];               // exclusive prefix of cnt / mls over the chunks
    u32 n, nch, rle_break, lastend, nseq, mltot;
    u32 hist[ZLC_NBINS];
};


HD bool zlc_mf(const ZlcSh &sh, u32 p) { return p >= 4 && sh.src(p) == sh.src(p - 4); }
HD bool zlc_contc(const ZlcSh &sh, u32 p) { return sh.oc(p) && p > 0 && sh.oc(p - 1) == sh.oc(p); }
HD bool zlc_contf(const ZlcSh &sh, u32 p) { return p > 0 && zlc_mf(sh, p) && zlc_mf(sh, p - 1); }
HD bool zlc_contd(const ZlcSh &sh, u32 p) { return sh.d(p) && p > 0 && sh.d(p - 1) == sh.d(p); }
HD u32 zlc_lo(u32 k) { return k * ZLC_CH; }
HD u32 zlc_hi(const ZlcSh &sh, u32 k) { const u32 h = k * ZLC_CH + ZLC_CH; return h < sh.n ? h : sh.n; }

// phase 1: where the chunk's last two terminators are; is the block one repeated byte
HD void zlc_zeros(ZlcSh &sh, u32 k)
{
    u32 a = ZLC_NONE, b = ZLC_NONE; bool same = true; const u8 c0 = sh.src(0);
    for (u32 p = zlc_lo(k), hi = zlc_hi(sh, k); p < hi; p++) { const u8 c = sh.src(p); if (c == 0) { b = a; a = p; } if (c != c0) same = false; }
    sh.z1[k] = (u16)a; sh.z2[k] = (u16)b;
    if (!same) sh.rle_break = 1;
}
// phase 2: the column candidate.  The record a byte is in starts behind the last terminator before it; the candidate offset is the
// length of the record before that one.
HD void zlc_columns(ZlcSh &sh, u32 k)
{
    u32 za = ZLC_NONE, zb = ZLC_NONE;
    for (u32 c = k; c-- > 0;) {
        if (sh.z1[c] == ZLC_NONE) continue;
        if (za == ZLC_NONE) { za = sh.z1[c]; if (sh.z2[c] != ZLC_NONE) { zb = sh.z2[c]; break; } }
        else { zb = sh.z1[c]; break; }
    }
    u32 cur = za == ZLC_NONE ? 0 : za + 1, prev = zb == ZLC_NONE ? 0 : zb + 1;
    for (u32 p = zlc_lo(k), hi = zlc_hi(sh, k); p < hi; p++) {
        u32 o = 0;
        if (cur > 0) { const u32 dcol = cur - prev; if (sh.src(p) == sh.src(p - dcol)) o = dcol; }
        sh.oc_[zlc_ix(p)] = (u16)o;
        if (sh.src(p) == 0) { prev = cur; cur = p + 1; }
    }
}
// phase 3: per chunk, where runs of either candidate break (so that a run's far ends are found chunk by chunk)
HD void zlc_breaks(ZlcSh &sh, u32 k)
{
    u32 lc = ZLC_NONE, fc = ZLC_NONE, lf = ZLC_NONE, ff = ZLC_NONE;
    for (u32 p = zlc_lo(k), hi = zlc_hi(sh, k); p < hi; p++) {
        if (!zlc_contc(sh, p)) { if (fc == ZLC_NONE) fc = p; lc = p; }
        if (!zlc_contf(sh, p)) { if (ff == ZLC_NONE) ff = p; lf = p; }
    }
    sh.lbc[k] = (u16)lc; sh.fbc[k] = (u16)fc; sh.lbf[k] = (u16)lf; sh.fbf[k] = (u16)ff;
}
// phase 4: a byte takes the candidate whose run around it is longer (ties: the column)
HD void zlc_choose(ZlcSh &sh, u32 k)
{
    const u32 lo = zlc_lo(k), hi = zlc_hi(sh, k), n = sh.n, nch = sh.nch;
    u32 endc = 0, lenc = 0, endf = 0, lenf = 0;                  // the run p is in, per candidate (valid while p < end)
    for (u32 p = lo; p < hi; p++) {
        u32 lc = 0, lf = 0;
        if (sh.oc(p)) {
            if (p >= endc) {
                u32 q = p; while (q > lo && zlc_contc(sh, q)) q--;
                u32 start = q;
                if (zlc_contc(sh, q)) { u32 c = k; do c--; while (sh.lbc[c] == ZLC_NONE); start = sh.lbc[c]; }     // (q == lo > 0: chunk 0 breaks at 0)
                u32 e = p + 1; while (e < hi && zlc_contc(sh, e)) e++;
                if (e == hi && hi < n && zlc_contc(sh, hi)) { u32 c = k + 1; while (c < nch && sh.fbc[c] == ZLC_NONE) c++; e = c < nch ? sh.fbc[c] : n; }
                endc = e; lenc = e - start;
            }
            lc = lenc;
        }
        if (zlc_mf(sh, p)) {
            if (p >= endf) {
                u32 q = p; while (q > lo && zlc_contf(sh, q)) q--;
                u32 start = q;
                if (zlc_contf(sh, q)) { u32 c = k; do c--; while (sh.lbf[c] == ZLC_NONE); start = sh.lbf[c]; }
                u32 e = p + 1; while (e < hi && zlc_contf(sh, e)) e++;
                if (e == hi && hi < n && zlc_contf(sh, hi)) { u32 c = k + 1; while (c < nch && sh.fbf[c] == ZLC_NONE) c++; e = c < nch ? sh.fbf[c] : n; }
                endf = e; lenf = e - start;
            }
            lf = lenf;
        }
        sh.d_[zlc_ix(p)] = lf > lc ? (u16)4 : sh.oc(p);
    }
}
// phase 5
HD void zlc_breaks_d(ZlcSh &sh, u32 k)
{
    u32 fd = ZLC_NONE;
    for (u32 p = zlc_lo(k), hi = zlc_hi(sh, k); p < hi; p++) if (!zlc_contd(sh, p)) { fd = p; break; }
    sh.fbd[k] = (u16)fd;
}
// the matches that START in chunk k, in order: f(start, end)
template <class F> HD void zlc_each_match(const ZlcSh &sh, u32 k, F f)
{
    const u32 lo = zlc_lo(k), hi = zlc_hi(sh, k);
    u32 p = lo;
    while (p < hi) {
        if (!sh.d(p) || zlc_contd(sh, p)) { p++; continue; }
        u32 q = p + 1; while (q < hi && zlc_contd(sh, q)) q++;
        u32 end = q;
        if (q == hi && hi < sh.n && zlc_contd(sh, hi)) { u32 c = k + 1; while (c < sh.nch && sh.fbd[c] == ZLC_NONE) c++; end = c < sh.nch ? sh.fbd[c] : sh.n; }
        if (end - p >= ZLC_MINML) f(p, end);
        p = q;
    }
}
// phase 6
HD void zlc_count(ZlcSh &sh, u32 k)
{
    u32 c = 0, m = 0, e = 0;
    zlc_each_match(sh, k, [&](u32 start, u32 end) { c++; m += end - start; e = end; });
    sh.cnt[k] = (u16)c; sh.mls[k] = (u16)m; sh.lend[k] = (u16)e;
    if (c) ZLC_MAXU(sh.lastend, e);
}
// phase 7 on the CPU (the kernel: one block-wide scan)
inline void zlc_scan_serial(ZlcSh &sh)
{
    u32 i = 0, m = 0;
    for (u32 k = 0; k < sh.nch; k++) { sh.ibase[k] = (u16)i; sh.mbase[k] = (u16)m; i += sh.cnt[k]; m += sh.mls[k]; }
    sh.nseq = i; sh.mltot = m;
}
// phase 8: my matches become sequences (literal length, match length, offset), the literals in front of each go to lit[]; in a
// sampled block the literal bytes and the LL / ML codes are counted on the way.  so / sl: the offsets and literal lengths once
// more, in shared memory, for the one thread that turns offsets into repeat codes (zlc_count_offsets).
HD void zlc_emit_seqs(ZlcSh &sh, u32 k, const ZLzSeqs &S, u8 *lit, bool sampled)
{
    if (!sh.cnt[k]) return;
    u32 pe = 0;
    for (u32 c = k; c-- > 0;) if (sh.cnt[c]) { pe = sh.lend[c]; break; }
    u32 idx = sh.ibase[k], msum = sh.mbase[k];
    u16 *so = sh.oc_, *sl = sh.oc_ + ZLC_PITCHED / 2;
    zlc_each_match(sh, k, [&](u32 start, u32 end) {
        const u32 ll = start - pe, ml = end - start, off = sh.d(start);
        S.ll[idx] = (u16)ll; S.ml[idx] = (u16)ml; S.ov[idx] = (u16)off;
        u8 *dst = lit + (pe - msum);
        for (u32 i = 0; i < ll; i++) { const u8 c = sh.src(pe + i); dst[i] = c; if (sampled) ZLC_INC(sh.hist[c]); }
        if (sampled) { so[idx] = (u16)off; sl[idx] = (u16)ll; ZLC_INC(sh.hist[ZLC_LL0 + zlz_ll_code(ll)]); ZLC_INC(sh.hist[ZLC_ML0 + zlz_ml_code(ml)]); }
        msum += ml; pe = end; idx++;
    });
}
// phase 8, the literals behind the last match: thread t of nt
HD void zlc_emit_tail(ZlcSh &sh, u32 t, u32 nt, u8 *lit, bool sampled)
{
    const u32 e = sh.lastend, base = e - sh.mltot;
    for (u32 i = t; e + i < sh.n; i += nt) { const u8 c = sh.src(e + i); lit[base + i] = c; if (sampled) ZLC_INC(sh.hist[c]); }
}
// phase 9 (sampled blocks, one thread): Offset_Value codes need the repeat-offset history, which is serial
HD void zlc_count_offsets(ZlcSh &sh)
{
    const u16 *so = sh.oc_, *sl = sh.oc_ + ZLC_PITCHED / 2;
    ZLzRep rep; rep.r[0] = rep.r[1] = rep.r[2] = 0; rep.k = 0;
    for (u32 i = 0; i < sh.nseq; i++) sh.hist[ZLC_OF0 + (u32)hibit(rep.code(so[i], sl[i]))]++;
}

} }  // namespace nafz::zlcb
