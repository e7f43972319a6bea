// zstd_lzc_hd.cuh — the data-parallel LZ stage of the text-like streams (ids, comments, lengths, mask) as HD code:
//   (1) a match finder made of maps, neighbour walks and one scan (no hash table, no serial parse): one CTA per block, one
//       thread per 32-byte chunk.  Candidate offsets per byte: the same column of the previous '\0'-terminated record, and 4
//       (the previous length unit); a byte takes the candidate whose run of matching bytes around it is longer; runs of at
//       least ZLC_MINML bytes become matches.  What compress/zstd_fast.c:186 finds through its hash table on these streams
//       is exactly this: each name repeats its predecessor but for a few digits.
//   (2) ONE Huffman code and ONE set of FSE tables per stream, built from the statistics of every ZLC_SAMPLE-th block: the
//       first block with literals and sequences carries them (Compressed_Literals + FSE_Compressed x 3), every later block
//       says Treeless_Literals + Repeat_Mode x 3 (what the reference's own frames are full of: compress/
//       zstd_compress_literals.c:70, zstd_compress_sequences.c:238 ZSTD_buildCTable).  A block then costs one thread only the
//       serial coding against read-only tables, and a decoder builds each table once per stream.
// Blocks still reference nothing before themselves (matches inside the block, repeat-offset codes only for offsets the block
// itself pushed), so N GPUs can still concatenate their blocks into one frame: each shard's defining block redefines the
// tables for the blocks behind it.
// tests/emu/emu_zlzc.cpp runs these bodies on the CPU thread by thread; libzstd 1.5.0, the oracle and our own decoder decode
// every frame (tests/test_emu_zenc.py), and tests/emu/lzcol.hpp + proto_shared.cpp are the serial restatement they are
// compared with byte for byte.
#pragma once
#include "zstd_enc_hd.cuh"

namespace nafz {

static const u32 ZLC_MAX = 8192, ZLC_CH = 32, ZLC_NCH = ZLC_MAX / ZLC_CH, ZLC_NONE = 0xFFFF, ZLC_MINML = 5, ZLC_SAMPLE = 8;
static const u32 ZLC_LL0 = 256, ZLC_OF0 = 292, ZLC_ML0 = 324, ZLC_NBINS = 377;      // literal bytes, LL / OF / ML codes

#ifdef __CUDA_ARCH__
#define ZLC_INC(x) atomicAdd(&(x), 1u)
#define ZLC_MAXU(x, v) atomicMax(&(x), (v))
#else
#define ZLC_INC(x) ((x)++)
#define ZLC_MAXU(x, v) ((x) = (x) > (v) ? (x) : (v))
#endif

// Thread k walks bytes 32 k .. 32 k + 31: with the arrays laid out plainly the 32 lanes of a warp would sit 32 bytes (8 banks) or
// 32 u16 (16 banks) apart and every access would be an 8- or 16-way bank conflict.  One element of padding per chunk (33-element
// pitch) puts the lanes of a warp into 32 different banks.
// Only the first three phases touch bytes; what they leave per chunk is a handful of 32-bit masks (one bit per byte: where each
// candidate matches, where its run continues from the byte before), and everything after -- run lengths, the choice between the
// candidates, the matches -- is bit scans over those masks plus per-chunk summaries for runs that leave the chunk.
static const u32 ZLC_PITCHED = ZLC_MAX + ZLC_NCH;
HD u32 zlc_ix(u32 p) { return p + (p >> 5); }
struct ZlcSh {                        // shared memory of one CTA = one block of at most ZLC_MAX bytes
    u8  src_[ZLC_PITCHED + 24];
    u16 oc_[ZLC_PITCHED];             // column candidate: its offset where the byte matches there, else 0
    HD u8 src(u32 p) const { return src_[zlc_ix(p)]; }
    HD u16 oc(u32 p) const { return oc_[zlc_ix(p)]; }
    // --- dead after phase 5, then reused for the sequences' offsets and literal lengths of a sampled block (zlc_seq_scratch) ---
    u32 cm[ZLC_NCH], fm[ZLC_NCH];     // bit i: byte 32 k + i matches at the column candidate / at offset 4
    u32 ccm[ZLC_NCH], ffm[ZLC_NCH];   // ... and so does the byte before it, at the same offset: the run continues
    u16 z1[ZLC_NCH], z2[ZLC_NCH];     // last / second-last '\0' of a chunk
    u16 lbc[ZLC_NCH], fbc[ZLC_NCH];   // last / first position of a chunk at which a run of the column candidate does not continue
    u16 lbf[ZLC_NCH], fbf[ZLC_NCH];   // same, candidate 4
    // --- live until the end ---
    u32 d4m[ZLC_NCH], dcm[ZLC_NCH];   // the choice: offset 4 / the column offset (neither: literal)
    u32 ddm[ZLC_NCH];                 // the chosen offset is the same as the byte before's: the match continues
    u16 fbd[ZLC_NCH];                 // first position of a chunk at which a run of the chosen offset does not continue
    u16 cnt[ZLC_NCH], mls[ZLC_NCH], lend[ZLC_NCH];    // matches starting in a chunk: how many, their lengths added up, where the last one ends
    u16 ibase[ZLC_NCH], mbase[ZLC_NCH];               // exclusive prefix of cnt / mls over the chunks
    u32 n, nch, rle_break, lastend, nseq, mltot;
    u32 hist[ZLC_NBINS];
};

static const u32 ZLC_SEQCAP = 7 * ZLC_NCH;                      // u16 entries per array in the reused region (a block has at most ZLC_MAX / 5 matches)
static_assert(ZLC_SEQCAP >= ZLC_MAX / ZLC_MINML + 1 && 2 * ZLC_SEQCAP * 2 <= 4 * ZLC_NCH * 4 + 6 * ZLC_NCH * 2, "sequence scratch does not fit the reused arrays");
HD u16 *zlc_seq_scratch(ZlcSh &sh) { return (u16 *)sh.cm; }    // [0, SEQCAP): offsets, [SEQCAP, 2 SEQCAP): literal lengths

struct ZlcBlk { u32 nseq, nlit; u8 parsed, rle, conv, pad; };      // what the finder leaves per block (conv: offsets already turned into Offset_Values)

HD u32 zlc_lo(u32 k) { return k * ZLC_CH; }
HD u32 zlc_hi(const ZlcSh &sh, u32 k) { const u32 h = k * ZLC_CH + ZLC_CH; return h < sh.n ? h : sh.n; }
HD u32 zlc_below(u32 i) { return i >= 32 ? 0xFFFFFFFFu : ((1u << i) - 1); }          // bits 0 .. i-1
HD u32 zlc_valid(const ZlcSh &sh, u32 k) { return zlc_below(zlc_hi(sh, k) - zlc_lo(k)); }
HD u32 zlc_low(u32 m)                 // index of the lowest set bit (m != 0)
{
#ifdef __CUDA_ARCH__
    return (u32)__ffs((int)m) - 1;
#else
    return (u32)__builtin_ctz(m);
#endif
}
HD u32 zlc_top(u32 m) { return (u32)hibit(m); }                                     // index of the highest set bit (m != 0)

// phase 1: where the chunk's last two terminators are; is the block one repeated byte
HD void zlc_zeros(ZlcSh &sh, u32 k)
{
    u32 a = ZLC_NONE, b = ZLC_NONE; bool same = true; const u8 c0 = sh.src(0);
    for (u32 p = zlc_lo(k), hi = zlc_hi(sh, k); p < hi; p++) { const u8 c = sh.src(p); if (c == 0) { b = a; a = p; } if (c != c0) same = false; }
    sh.z1[k] = (u16)a; sh.z2[k] = (u16)b;
    if (!same) sh.rle_break = 1;
}
// phase 2: the two candidates, byte by byte.  The record a byte is in starts behind the last terminator before it; the column
// candidate's offset is the length of the record before that one.
HD void zlc_columns(ZlcSh &sh, u32 k)
{
    u32 za = ZLC_NONE, zb = ZLC_NONE;
    for (u32 c = k; c-- > 0;) {
        if (sh.z1[c] == ZLC_NONE) continue;
        if (za == ZLC_NONE) { za = sh.z1[c]; if (sh.z2[c] != ZLC_NONE) { zb = sh.z2[c]; break; } }
        else { zb = sh.z1[c]; break; }
    }
    u32 cur = za == ZLC_NONE ? 0 : za + 1, prev = zb == ZLC_NONE ? 0 : zb + 1;
    const u32 lo = zlc_lo(k), hi = zlc_hi(sh, k);
    u32 cm = 0, fm = 0;
    u32 w = 0;                                                  // the four bytes before p, oldest in the top byte
    for (u32 i = lo >= 4 ? 4 : lo; i > 0; i--) w = (w << 8) | sh.src(lo - i);
    for (u32 p = lo; p < hi; p++) {
        const u32 c = sh.src(p), bit = 1u << (p - lo);
        u32 o = 0;
        if (cur > 0) { const u32 dcol = cur - prev; if (c == sh.src(p - dcol)) o = dcol; }
        sh.oc_[zlc_ix(p)] = (u16)o;
        if (o) cm |= bit;
        if (p >= 4 && c == (w >> 24)) fm |= bit;
        w = (w << 8) | c;
        if (c == 0) { prev = cur; cur = p + 1; }
    }
    sh.cm[k] = cm; sh.fm[k] = fm;
}
// phase 3: where runs continue, and per chunk where they break (so that a run's far ends are found chunk by chunk)
HD void zlc_breaks(ZlcSh &sh, u32 k)
{
    const u32 lo = zlc_lo(k), hi = zlc_hi(sh, k), valid = zlc_valid(sh, k);
    u32 ccm = 0, po = lo ? sh.oc(lo - 1) : 0;
    for (u32 p = lo; p < hi; p++) { const u32 o = sh.oc(p); if (o && o == po) ccm |= 1u << (p - lo); po = o; }      // (p == 0: po = 0, no run continues)
    const u32 fm = sh.fm[k], ffm = fm & ((fm << 1) | (k ? sh.fm[k - 1] >> 31 : 0));       // (chunk k - 1 is a full one)
    sh.ccm[k] = ccm; sh.ffm[k] = ffm;
    const u32 bc = ~ccm & valid, bf = ~ffm & valid;             // (never 0 for k == 0: nothing continues at byte 0)
    sh.lbc[k] = (u16)(bc ? lo + zlc_top(bc) : ZLC_NONE); sh.fbc[k] = (u16)(bc ? lo + zlc_low(bc) : ZLC_NONE);
    sh.lbf[k] = (u16)(bf ? lo + zlc_top(bf) : ZLC_NONE); sh.fbf[k] = (u16)(bf ? lo + zlc_low(bf) : ZLC_NONE);
}
// length of the run (of the candidate whose continue-mask is `cont`) that byte lo + a is in, and where it ends
HD u32 zlc_run(const ZlcSh &sh, u32 k, u32 a, u32 cont, const u16 *lastbrk, const u16 *firstbrk, u32 *end)
{
    const u32 lo = zlc_lo(k), valid = zlc_valid(sh, k);
    const u32 brk = ~cont & valid;
    const u32 before = brk & zlc_below(a + 1), after = brk & ~zlc_below(a + 1);
    u32 start, e;
    if (before) start = lo + zlc_top(before);
    else { u32 c = k; do c--; while (lastbrk[c] == ZLC_NONE); start = lastbrk[c]; }          // (k > 0: chunk 0 breaks at byte 0)
    if (after) e = lo + zlc_low(after);
    else if (valid != 0xFFFFFFFFu) e = sh.n;                                                 // the block's last, partial chunk
    else { u32 c = k + 1; while (c < sh.nch && firstbrk[c] == ZLC_NONE) c++; e = c < sh.nch ? firstbrk[c] : sh.n; }
    *end = e;
    return e - start;
}
// phase 4: a byte takes the candidate whose run around it is longer (ties: the column).  Between two breaks of either candidate
// both runs are the same for every byte, so the choice is made once per such piece.
HD void zlc_choose(ZlcSh &sh, u32 k)
{
    const u32 lo = zlc_lo(k), valid = zlc_valid(sh, k), nv = zlc_hi(sh, k) - lo;
    const u32 cm = sh.cm[k], fm = sh.fm[k], ccm = sh.ccm[k], ffm = sh.ffm[k];
    u32 pieces = ((~ccm | ~ffm) & valid) | 1u, d4 = 0, dc = 0;
    u32 endc = 0, lenc = 0, endf = 0, lenf = 0;                 // the run the current piece is in, per candidate (valid while p < end)
    while (pieces) {
        const u32 a = zlc_low(pieces); pieces &= pieces - 1;
        const u32 b = pieces ? zlc_low(pieces) : nv, p = lo + a;
        if (!(((cm | fm) >> a) & 1)) continue;
        u32 lc = 0, lf = 0;
        if ((cm >> a) & 1) { if (p >= endc) lenc = zlc_run(sh, k, a, ccm, sh.lbc, sh.fbc, &endc); lc = lenc; }
        if ((fm >> a) & 1) { if (p >= endf) lenf = zlc_run(sh, k, a, ffm, sh.lbf, sh.fbf, &endf); lf = lenf; }
        const u32 pm = zlc_below(b) & ~zlc_below(a);
        if (lf > lc) d4 |= pm; else if (lc) dc |= pm;
    }
    sh.d4m[k] = d4; sh.dcm[k] = dc;
}
HD u32 zlc_dval(const ZlcSh &sh, u32 p) { const u32 k = p >> 5, bit = 1u << (p & 31); return (sh.d4m[k] & bit) ? 4u : ((sh.dcm[k] & bit) ? (u32)sh.oc(p) : 0u); }
// phase 5: where the chosen offset continues.  Both bytes at offset 4: yes; both at the column offset: where the column run continues;
// one of each: only if the column offset happens to be 4.
HD void zlc_breaks_d(ZlcSh &sh, u32 k)
{
    const u32 lo = zlc_lo(k), valid = zlc_valid(sh, k);
    const u32 d4 = sh.d4m[k], dc = sh.dcm[k];
    const u32 s4 = (d4 << 1) | (k ? sh.d4m[k - 1] >> 31 : 0), sc = (dc << 1) | (k ? sh.dcm[k - 1] >> 31 : 0);
    u32 dd = (d4 & s4) | (dc & sc & sh.ccm[k]);
    u32 mixed = (d4 & sc) | (dc & s4);
    while (mixed) {
        const u32 a = zlc_low(mixed); mixed &= mixed - 1;
        if (zlc_dval(sh, lo + a) == zlc_dval(sh, lo + a - 1)) dd |= 1u << a;
    }
    sh.ddm[k] = dd;
    const u32 bd = ~dd & valid;
    sh.fbd[k] = (u16)(bd ? lo + zlc_low(bd) : ZLC_NONE);
}
// the matches that START in chunk k, in order: f(start, end)
template <class F> HD void zlc_each_match(const ZlcSh &sh, u32 k, F f)
{
    const u32 lo = zlc_lo(k), valid = zlc_valid(sh, k), dd = sh.ddm[k];
    u32 starts = (sh.d4m[k] | sh.dcm[k]) & ~dd & valid;
    while (starts) {
        const u32 a = zlc_low(starts); starts &= starts - 1;
        const u32 after = ~dd & valid & ~zlc_below(a + 1);
        u32 end;
        if (after) end = lo + zlc_low(after);
        else if (valid != 0xFFFFFFFFu) end = sh.n;
        else { u32 c = k + 1; while (c < sh.nch && sh.fbd[c] == ZLC_NONE) c++; end = c < sh.nch ? sh.fbd[c] : sh.n; }
        if (end - (lo + a) >= ZLC_MINML) f(lo + a, end);
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
    u16 *so = zlc_seq_scratch(sh), *sl = so + ZLC_SEQCAP;
    zlc_each_match(sh, k, [&](u32 start, u32 end) {
        const u32 ll = start - pe, ml = end - start, off = zlc_dval(sh, start);
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
    const u16 *so = zlc_seq_scratch(sh), *sl = so + ZLC_SEQCAP;
    ZLzRep rep; rep.r[0] = rep.r[1] = rep.r[2] = 0; rep.k = 0;
    for (u32 i = 0; i < sh.nseq; i++) sh.hist[ZLC_OF0 + (u32)hibit(rep.code(so[i], sl[i]))]++;
}

// ---------------------------------------------------------------- per-stream tables
struct ZlcTables {
    ZEncMeta M;                                    // literal code + tree description
    short nl[36], no[32], nm[53]; u16 cuml[37], cumo[33], cumm[54];
    u16 spos[512 + 256 + 512];
    int logl, logo, logm, nsl, nso, nsm;
    u8 desc[512]; u32 desc_len;                    // the three FSE table descriptions as the defining block writes them
    u32 ok;                                        // 0: this stream's blocks get their own tables (zlz_emit_block)
    u32 fdef;                                      // the block (index in the stream) that carries the tables
    u32 pad[2];                                    // (size: a multiple of 16, the block coder's CTA copies it to shared memory in 16-byte pieces)
};

// Tables from the sampled statistics, smoothed: every sequence code a block of this size can produce keeps a probability
// (count * 16 + 1), so no block ever needs a fallback that would change the decoder's state; literal bytes the sample never saw
// have no code (a block with one stores its literals raw, which leaves the decoder's tree alone).
HDN inline bool zlc_build_tables(const u32 *cnt, u32 bs, ZlcTables &T, u8 *tsym /* 512 */)
{
    T.logl = 9; T.logo = 8; T.logm = 9; T.desc_len = 0;
    u64 nlit = 0, nseq = 0, mx = 0;
    for (int s = 0; s < 256; s++) { nlit += cnt[s]; if (cnt[s] > mx) mx = cnt[s]; }
    for (int s = 0; s < 36; s++) nseq += cnt[ZLC_LL0 + s];
    if (nlit == 0 || nseq == 0) return false;
    u64 scale = 1; while (mx / scale > 60000) scale *= 2;
    u16 h16[256];
    for (int s = 0; s < 256; s++) { h16[s] = (u16)(cnt[s] / scale); if (cnt[s] && !h16[s]) h16[s] = 1; }
    zenc_huf_build(h16, T.M);
    if (T.M.mode != 2) return false;
    u32 cl[36], co[32], cm[53];
    T.nsl = (int)zlz_ll_code(bs) + 1; T.nsm = (int)zlz_ml_code(bs) + 1; T.nso = hibit(bs + 3) + 1;
    u64 sc = 1; while (nseq * 16 / sc > (1u << 26)) sc *= 2;
    u32 tl = 0, to = 0, tm = 0;
    for (int s = 0; s < T.nsl; s++) { cl[s] = (u32)((u64)cnt[ZLC_LL0 + s] * 16 / sc) + 1; tl += cl[s]; }
    for (int s = 0; s < T.nso; s++) { co[s] = (u32)((u64)cnt[ZLC_OF0 + s] * 16 / sc) + 1; to += co[s]; }
    for (int s = 0; s < T.nsm; s++) { cm[s] = (u32)((u64)cnt[ZLC_ML0 + s] * 16 / sc) + 1; tm += cm[s]; }
    if (!fse_normalize(cl, T.nsl, tl, T.logl, T.nl) || !fse_normalize(co, T.nso, to, T.logo, T.no) || !fse_normalize(cm, T.nsm, tm, T.logm, T.nm)) return false;
    fse_build_enc(T.nl, T.nsl, T.logl, T.spos, T.cuml, tsym);
    fse_build_enc(T.no, T.nso, T.logo, T.spos + 512, T.cumo, tsym);
    fse_build_enc(T.nm, T.nsm, T.logm, T.spos + 768, T.cumm, tsym);
    for (u32 i = 0; i < sizeof T.desc; i++) T.desc[i] = 0;
    BitW hw; hw.init(T.desc, (u32)sizeof T.desc);
    fse_write_ncount(hw, T.nl, T.nsl, T.logl); fse_write_ncount(hw, T.no, T.nso, T.logo); fse_write_ncount(hw, T.nm, T.nsm, T.logm);
    if (!hw.ok) return false;
    T.desc_len = hw.pos;
    return true;
}

// offsets -> Offset_Values (repeat codes where the block itself pushed the offset), in place, once
HD void zlc_offset_values(const ZLzSeqs &S, ZlcBlk &I)
{
    if (I.conv) return;
    ZLzRep rep; rep.r[0] = rep.r[1] = rep.r[2] = 0; rep.k = 0;
    u32 i = 0;                                                  // (loads in groups: the history is a serial chain, the memory waits need not be)
    for (; i + 4 <= S.n; i += 4) {
        u16 o[4], l[4]; ld_group<4>(S.ov + i, o); ld_group<4>(S.ll + i, l);
        for (int k = 0; k < 4; k++) S.ov[i + k] = (u16)rep.code(o[k], l[k]);
    }
    for (; i < S.n; i++) S.ov[i] = (u16)rep.code(S.ov[i], S.ll[i]);
    I.conv = 1;
}

// Literals_Section with the stream's code: Compressed_Literals (+ tree) in the defining block, Treeless_Literals later; raw when
// that is smaller or a byte has no code.  0: the defining block cannot be written.
HDN inline u32 zlc_put_literals(const u8 *lit, u32 nlit, const ZlcTables &T, bool first, u8 *out, u32 cap)
{
    auto raw = [&]() -> u32 {
        u32 h;
        if (nlit + 3 > cap) return 0;
        if (nlit < 32) { out[0] = (u8)(nlit << 3); h = 1; }
        else if (nlit < 4096) { const u32 v = (1u << 2) | (nlit << 4); out[0] = (u8)v; out[1] = (u8)(v >> 8); h = 2; }
        else { const u32 v = (3u << 2) | (nlit << 4); out[0] = (u8)v; out[1] = (u8)(v >> 8); out[2] = (u8)(v >> 16); h = 3; }
        u32 i = 0;
        for (; i + 8 <= nlit; i += 8) { u8 t[8]; ld_group<8>(lit + i, t); for (int k = 0; k < 8; k++) out[h + i + k] = t[k]; }
        for (; i < nlit; i++) out[h + i] = lit[i];
        return h + nlit;
    };
    if (nlit == 0) return raw();
    const u16 *ctab = T.M.ctab;
    // one pass: does every byte have a code; bits in total and per quarter
    const u32 seg4 = (nlit + 3) / 4;
    u32 qbits[4] = {0, 0, 0, 0}; bool coded = true;
    for (u32 k = 0; k < 4; k++) {
        const u32 a = k * seg4 < nlit ? k * seg4 : nlit, b = k == 3 ? nlit : ((k + 1) * seg4 < nlit ? (k + 1) * seg4 : nlit);
        u32 bits = 0, i = a;
        for (; i + 8 <= b; i += 8) { u8 t[8]; ld_group<8>(lit + i, t); for (int j = 0; j < 8; j++) { const u32 l = ctab[t[j]] >> 12; if (!l) coded = false; bits += l; } }
        for (; i < b; i++) { const u32 l = ctab[lit[i]] >> 12; if (!l) coded = false; bits += l; }
        qbits[k] = bits;
    }
    if (!coded) return first ? 0 : raw();
    const u32 tree = first ? T.M.tree_len : 0;
    const u32 total_bits = qbits[0] + qbits[1] + qbits[2] + qbits[3];
    const u32 nstreams = (nlit <= 1023 && tree + total_bits / 8 + 1 <= 1023) ? 1 : 4;
    if (nstreams == 4 && nlit < 16) return first ? 0 : raw();
    const u32 seg = nstreams == 4 ? seg4 : nlit;
    u32 sbytes[4] = {0, 0, 0, 0}, payload = tree + (nstreams == 4 ? 6u : 0u);
    for (u32 k = 0; k < nstreams; k++) { sbytes[k] = (nstreams == 4 ? qbits[k] : total_bits) / 8 + 1; payload += sbytes[k]; }
    const u32 lh = nstreams == 1 ? 3 : ((nlit <= 16383 && payload <= 16383) ? 4 : 5);
    const u32 raw_size = nlit + (nlit < 32 ? 1 : (nlit < 4096 ? 2 : 3));
    if (!first && lh + payload >= raw_size) return raw();       // (the defining block must carry the tree whatever it costs)
    if (lh + payload > cap) return 0;
    const u32 type = first ? 2u : 3u;
    if (nstreams == 1) { const u32 v = type | (0 << 2) | (nlit << 4) | (payload << 14); out[0] = (u8)v; out[1] = (u8)(v >> 8); out[2] = (u8)(v >> 16); }
    else if (lh == 4) { const u32 v = type | (2 << 2) | (nlit << 4) | (payload << 18); out[0] = (u8)v; out[1] = (u8)(v >> 8); out[2] = (u8)(v >> 16); out[3] = (u8)(v >> 24); }
    else { const u64 v = type | (3 << 2) | ((u64)nlit << 4) | ((u64)payload << 22); for (int i = 0; i < 5; i++) out[i] = (u8)(v >> (8 * i)); }
    u32 at = lh;
    for (u32 i = 0; i < tree; i++) out[at++] = T.M.tree[i];
    if (nstreams == 4) for (int j = 0; j < 3; j++) { out[at++] = (u8)sbytes[j]; out[at++] = (u8)(sbytes[j] >> 8); }
    for (u32 k = 0; k < nstreams; k++) {
        const u32 a = k * seg, b = k == nstreams - 1 ? nlit : (a + seg < nlit ? a + seg : nlit);
        BitW bw; bw.init(out + at, sbytes[k]);
        u32 i = b;                                            // the last symbol sits at the lowest bits
        for (; i >= a + 8; i -= 8) { u8 t[8]; ld_group<8>(lit + i - 8, t); for (int j = 7; j >= 0; j--) { const u32 e = ctab[t[j]]; bw.put(e & 0xFFF, e >> 12); } }
        for (; i > a; i--) { const u32 e = ctab[lit[i - 1]]; bw.put(e & 0xFFF, e >> 12); }
        bw.finish_with_mark();
        at += sbytes[k];
    }
    return at;
}

// A Compressed_Block coded against the stream's tables (S.ov: Offset_Values).  0: store the block raw (for the defining block:
// it cannot define the tables).
HDN inline u32 zlc_emit_shared(u32 n, const u8 *lit, u32 nlit, const ZLzSeqs &S, const ZlcTables &T, bool first, u8 *out, u32 cap)
{
    u32 at = zlc_put_literals(lit, nlit, T, first, out, cap);
    if (!at) return 0;
    const u32 nseq = S.n;
    if (at + 4 + (first ? T.desc_len : 0) > cap) return 0;
    if (nseq == 0) { if (first) return 0; out[at++] = 0; return at < n ? at : 0; }
    if (nseq < 128) out[at++] = (u8)nseq;
    else if (nseq < 0x7F00) { out[at++] = (u8)((nseq >> 8) + 128); out[at++] = (u8)nseq; }
    else { out[at++] = 255; out[at++] = (u8)(nseq - 0x7F00); out[at++] = (u8)((nseq - 0x7F00) >> 8); }
    out[at++] = first ? (u8)((2 << 6) | (2 << 4) | (2 << 2)) : (u8)((3 << 6) | (3 << 4) | (3 << 2));      // FSE_Compressed x 3 / Repeat_Mode x 3
    if (first) { for (u32 i = 0; i < T.desc_len; i++) out[at + i] = T.desc[i]; at += T.desc_len; }
    BitW bw; bw.init(out + at, cap - at);
    FseEnc EL{T.nl, T.cuml, T.spos, T.logl, 0, 2, 0}, EO{T.no, T.cumo, T.spos + 512, T.logo, 0, 2, 0}, EM{T.nm, T.cumm, T.spos + 768, T.logm, 0, 2, 0};
    {
        const u32 i = nseq - 1, c_l = zlz_ll_code(S.ll[i]), c_o = (u32)hibit(S.ov[i]), c_m = zlz_ml_code(S.ml[i]);
        EM.start(c_m); EO.start(c_o); EL.start(c_l);
        bw.put(S.ll[i] - ll_base_of(c_l), ll_bits_of(c_l));
        bw.put(S.ml[i] - ml_base_of(c_m), ml_bits_of(c_m));
        bw.put(S.ov[i] - (1u << c_o), c_o);
    }
    auto one = [&](u32 ll, u32 ml, u32 ov) {
        const u32 c_l = zlz_ll_code(ll), c_o = (u32)hibit(ov), c_m = zlz_ml_code(ml);
        EO.put(bw, c_o); EM.put(bw, c_m); EL.put(bw, c_l);
        bw.put(ll - ll_base_of(c_l), ll_bits_of(c_l));
        bw.put(ml - ml_base_of(c_m), ml_bits_of(c_m));
        bw.put(ov - (1u << c_o), c_o);
    };
    u32 i = nseq - 1;                                         // sequences i-1 ... 0 remain
    for (; i >= 4; i -= 4) {
        u16 a[4], b[4], c[4];
        ld_group<4>(S.ll + i - 4, a); ld_group<4>(S.ml + i - 4, b); ld_group<4>(S.ov + i - 4, c);
        for (int k = 3; k >= 0; k--) one(a[k], b[k], c[k]);
    }
    for (; i > 0; i--) one(S.ll[i - 1], S.ml[i - 1], S.ov[i - 1]);
    EM.flush(bw); EO.flush(bw); EL.flush(bw);
    bw.finish_with_mark();
    if (!bw.ok) return 0;
    at += bw.pos;
    return (first || at < n) ? at : 0;
}

// What one stream's blocks look like to the table builder and the block coder: block b's bytes, what the finder left, its slot.
struct ZlcStreamView {
    const u8 *src; u64 n; u32 bs, nblk;             // block b = src[b * bs .. min(n, (b + 1) * bs))
    ZlcBlk *info;                                   // [nblk]
    u8 *work; u32 work_stride;                      // block b's scratch (layout: zlc_work)
    u8 *slots; u32 slot_stride;                     // block b's output slot (capacity slot_stride)
    HD u32 len(u32 b) const { const u64 off = (u64)b * bs, left = n > off ? n - off : 0; return (u32)(left < bs ? left : bs); }
};
struct ZlcWork { u8 *lit; ZLzSeqs S; ZLzWork W; };
HD u32 zlc_work_bytes(u32 bs) { const u32 ms = bs / 4; return (((bs + 64 + 15) & ~15u) + 3 * ms * 2 + 1280 * 2 + 512 + 3 * ms + 15) & ~15u; }
HD ZlcWork zlc_work(u8 *w, u32 bs)
{
    const u32 ms = bs / 4; ZlcWork R;
    R.lit = w; w += (bs + 64 + 15) & ~15u;                      // (the u16 arrays behind it stay aligned for any block size)
    R.S.ll = (u16 *)w; w += ms * 2; R.S.ml = (u16 *)w; w += ms * 2; R.S.ov = (u16 *)w; w += ms * 2; R.S.n = 0;
    R.W.spos = (u16 *)w; w += 1280 * 2; R.W.tsym = w; w += 512; R.W.codes = w;
    return R;
}

// One thread per stream, after the finder: which block will carry the tables -- the first one with literals and sequences (blocks
// before it are stored raw or RLE, which leaves a decoder's entropy state alone) -- and the tables, from the sampled counts plus
// that block's own when it is not a sampled one (so that each of its literal bytes has a code).
HDN inline void zlc_define(const ZlcStreamView &V, u32 *cnt, ZlcTables &T)
{
    T.ok = 0; T.fdef = 0xFFFFFFFFu;
    if (V.bs < 64) return;
    u32 f = 0;
    while (f < V.nblk && !(V.info[f].parsed && V.info[f].nlit && V.info[f].nseq)) f++;
    if (f == V.nblk) return;
    if (f % ZLC_SAMPLE) {
        ZlcBlk &I = V.info[f];
        ZlcWork K = zlc_work(V.work + (size_t)f * V.work_stride, V.bs);
        K.S.n = I.nseq;
        zlc_offset_values(K.S, I);
        for (u32 i = 0; i < I.nlit; i++) cnt[K.lit[i]]++;
        for (u32 i = 0; i < I.nseq; i++) { cnt[ZLC_LL0 + zlz_ll_code(K.S.ll[i])]++; cnt[ZLC_OF0 + (u32)hibit(K.S.ov[i])]++; cnt[ZLC_ML0 + zlz_ml_code(K.S.ml[i])]++; }
    }
    u8 tsym[512];
    if (!zlc_build_tables(cnt, V.bs, T, tsym)) return;
    T.ok = 1; T.fdef = f;
}

// One thread per block, after zlc_define: type (0 raw, 1 RLE, 2 compressed) and content size of block b.  Two bodies, so that the
// common one (coding against the stream's tables) does not carry the registers and the stack of a private Huffman / FSE build:
// zlc_finish_block for streams whose tables are defined, zlc_finish_block_own for the others (each returns false when the block
// belongs to the other one).  *def_fail: the defining block could not be written (its literals, coded with the stream's code,
// overflow the slot): zlc_finish_block_own then redoes the whole stream with private tables.
HDN inline bool zlc_finish_block(const ZlcStreamView &V, u32 b, const ZlcTables &T, u32 *def_fail, u32 *type, u32 *csize)
{
    if (!T.ok) return false;
    const u32 n = V.len(b);
    ZlcBlk &I = V.info[b];
    u8 *slot = V.slots + (size_t)b * V.slot_stride;
    if (I.rle) { slot[0] = V.src[(u64)b * V.bs]; *type = 1; *csize = 1; return true; }
    *type = 0; *csize = n;
    if (!I.parsed || b < T.fdef) return true;                  // (before the defining block the tables are not there yet: raw)
    ZlcWork K = zlc_work(V.work + (size_t)b * V.work_stride, V.bs);
    K.S.n = I.nseq;
    zlc_offset_values(K.S, I);
    const u32 cs = zlc_emit_shared(n, K.lit, I.nlit, K.S, T, b == T.fdef, slot, V.slot_stride);
    if (cs) { *type = 2; *csize = cs; }
    else if (b == T.fdef) *def_fail = 1;
    return true;
}
HDN inline bool zlc_finish_block_own(const ZlcStreamView &V, u32 b, const ZlcTables &T, u32 def_fail, u32 *type, u32 *csize)
{
    if (T.ok && !def_fail) return false;
    const u32 n = V.len(b);
    ZlcBlk &I = V.info[b];
    u8 *slot = V.slots + (size_t)b * V.slot_stride;
    if (I.rle) { slot[0] = V.src[(u64)b * V.bs]; *type = 1; *csize = 1; return true; }
    *type = 0; *csize = n;
    if (!I.parsed) return true;
    ZlcWork K = zlc_work(V.work + (size_t)b * V.work_stride, V.bs);
    K.S.n = I.nseq;
    zlc_offset_values(K.S, I);
    const u32 cs = zlz_emit_block(n, K.lit, I.nlit, K.S, V.bs / 4, K.W, slot, V.slot_stride);
    if (cs) { *type = 2; *csize = cs; }
    return true;
}

}  // namespace nafz
