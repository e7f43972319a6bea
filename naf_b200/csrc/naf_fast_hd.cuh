// naf_fast_hd.cuh — per-chunk logic of the *canonical-input* FASTA / FASTQ parser (host + device).
//
// The general parser (naf_parse.cuh) restates ennaf/src/process.c:358,477 as a byte-level finite-state
// machine so that every quirk of non-well-formed input is reproduced.  Almost all real input is much
// simpler: LF line ends, no stray white space or control bytes, nothing "unexpected".  For such input the
// same machine collapses to line bookkeeping:
//   FASTQ  role of a line = (number of '\n' before it) mod 4: header, sequence, '+' line, quality
//   FASTA  a line is a header iff it starts with '>', otherwise its bytes are sequence
//   header = name up to the first ' ', rest of the line is the comment
// The kernels in naf_parse_fast.cuh run exactly that, and *verify while they go* that the input is canonical
// (conditions C1..C6 below).  The first violation raises a flag and the caller redoes the split with the
// general parser, so results are always those of process.c.
//
// Canonical input (each condition is checked; equivalence with the FSM argued in DESIGN.md "Parser"):
//   C1  no byte < 32 other than '\n', no byte 127 or 255 anywhere after the first '>' / '@'
//   C2  FASTQ: every line of role 0 starts with '@', every line of role 2 with '+'
//   C3  FASTQ: no empty line
//   C4  sequence bytes are expected ones for the alphabet (tables.c:72-115); checked where the bytes are
//       consumed (4-bit pack LUT for DNA/RNA, SWAR tests for protein/text)
//   C5  quality bytes are 33..126 (tables.c:137)
//   C6  a header's first space is found within FAST_LOOKBACK bytes of a chunk start (bounded look-back)
// Everything here is plain C++ over a 64-byte chunk so that tests/emu can run it on the CPU against the
// oracle; the shipped library only instantiates it inside kernels.
#pragma once
#include "zstd_hd.cuh"

namespace nafg {

using nafz::u8; using nafz::u16; using nafz::u32; using nafz::u64;

enum { FR_HDR = 0, FR_SEQ = 1, FR_PLUS = 2, FR_QUAL = 3 };
// FASTA: what a stretch of text does to "which kind of line am I in" (composes associatively)
enum { FE_ID = 0, FE_NONE_GT = 1, FE_NONE_OT = 2, FE_HDR = 3, FE_SEQ = 4, FE_LS = 5 };
enum : u32 { FF_BADBYTE = 1, FF_FIRSTCHAR = 2, FF_BLANK = 4, FF_SEQ = 8, FF_QUAL = 16, FF_LOOKBACK = 32 };
static const u32 FAST_LOOKBACK = 4096;
static const u32 FAST_FILL = 0x41414141u;      // 'A': stands in for bytes outside [p0, n) in partial chunks

HD u32 fe_compose(u32 a, u32 b)                 // first a, then b
{
    if (b == FE_ID) return a;
    if (b >= FE_HDR) return b;
    if (a == FE_ID) return b;
    if (a == FE_LS) return b == FE_NONE_GT ? (u32)FE_HDR : (u32)FE_SEQ;
    return a;
}

// nonzero iff some byte of v is zero; the lowest flagged byte is exact
HD u32 swar_haszero(u32 v) { return (v - 0x01010101u) & ~v & 0x80808080u; }

// '\n' positions of a chunk, "some byte violates C1" and (WITH_SP) the positions of ' '
template <bool WITH_SP = false> HD void fast_chunk_scan(const u32 w[16], u64 &nl, u32 &bad, u64 *sp = nullptr)
{
    u32 nlo = 0, nhi = 0, b = 0, slo = 0, shi = 0;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int k = 0; k < 16; k++) {
        const u32 v = w[k];
        b |= swar_haszero((v & 0x7F7F7F7Fu) ^ 0x7F7F7F7Fu);             // 127 or 255
        if (swar_haszero(v & 0xE0E0E0E0u)) {                            // some byte < 32: almost always a '\n'
            u32 m = 0;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
            for (int j = 0; j < 4; j++) {
                const u32 c = (v >> (8 * j)) & 0xFF;
                if (c == '\n') m |= 1u << j; else if (c < 32) b |= 1;
            }
            if (k < 8) nlo |= m << (4 * k); else nhi |= m << (4 * (k - 8));
        }
        if (WITH_SP && swar_haszero(v ^ 0x20202020u)) {                 // spaces only occur in header lines: rare
            u32 m = 0;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
            for (int j = 0; j < 4; j++) if (((v >> (8 * j)) & 0xFF) == ' ') m |= 1u << j;
            if (k < 8) slo |= m << (4 * k); else shi |= m << (4 * (k - 8));
        }
    }
    nl = (u64)nlo | ((u64)nhi << 32);
    bad = b;
    if (WITH_SP) *sp = (u64)slo | ((u64)shi << 32);
}

// shared-memory layout of byte arrays that are accessed both "one 64-byte row per thread" and linearly: one pad word
// after every 16 (row stride 17 words), so that both patterns are bank-conflict free with a two-instruction address
HD u32 fast_pad(u32 a) { return a + ((a >> 6) << 2); }

struct FastState { u32 role, sp, ls; };          // role: FR_*;  sp: the header's first space was seen;  ls: at a line start
struct FastEmit { u32 ids, comm, seq, qual, rec; };
struct FastLine { u64 base, max; u32 mark; };    // FASTA longest-line bookkeeping (same meaning as in naf_parse.cuh walk())

// FASTA element of one chunk: live bytes are [b0, b1)
template <class Row> HD u32 fasta_chunk_element(const Row &row, u64 nl, u32 b0, u32 b1)
{
    if (b0 >= b1) return FE_ID;
    if (nl == 0) return row(b0) == '>' ? (u32)FE_NONE_GT : (u32)FE_NONE_OT;
    u32 last = 63;
    while (!((nl >> last) & 1)) last--;
    if (last + 1 >= b1) return FE_LS;
    return row(last + 1) == '>' ? (u32)FE_HDR : (u32)FE_SEQ;
}

HD u32 ctz64(u64 v)
{
#ifdef __CUDA_ARCH__
    return (u32)(__ffsll((long long)v) - 1);
#else
    return (u32)__builtin_ctzll(v);
#endif
}

// Walk the lines of one chunk.  COUNT mode only counts; SCATTER mode also hands every run of bytes to `sink`:
//   sink.copy(stream, src_pos_in_chunk, len, dst_off_in_my_stream_region)   stream: 0 ids, 1 comments, 2 sequence, 3 quality
//   sink.put(stream, dst_off, byte)
//   sink.rec_end(index, counted_seq, qual_bytes, text_pos)
template <bool FASTQ, bool SCATTER, class Row, class Sink>
HD void fast_walk(const Row &row, u64 nl, u64 sp, u32 b0, u32 b1, FastState &st, u64 lo, FastEmit &n, Sink &sink,
                  u64 o_cnt, u64 o_qual, u64 o_rec, FastLine &ln, u32 &flag)
{
    u32 pos = b0;
    while (pos < b1) {
        if (st.ls) {
            const u32 c = row(pos);
            st.ls = 0; st.sp = 0;
            if (FASTQ) {
                if (c == '\n') flag |= FF_BLANK;
                if (st.role == FR_HDR) { if (c == '@') { pos++; continue; } flag |= FF_FIRSTCHAR; }
                else if (st.role == FR_PLUS && c != '+') flag |= FF_FIRSTCHAR;
            } else {
                if (c == '>') {                                      // process.c:383: a new record starts
                    if (SCATTER) sink.rec_end(o_rec + n.rec, o_cnt + n.seq, 0, lo + pos);
                    n.rec++; st.role = FR_HDR; pos++;
                    continue;
                }
                st.role = FR_SEQ;
            }
        }
        const u64 rest = nl >> pos;                                  // pos < 64
        u32 e = rest ? pos + ctz64(rest) : b1;
        if (e > b1) e = b1;
        const bool has_nl = e < b1;
        const u32 len = e - pos;
        switch (st.role) {
        case FR_HDR: {
            u32 p = pos;
            if (!st.sp) {
                const u64 srest = sp >> pos;                            // first ' ' at or after pos, if before e
                u32 s = srest ? pos + ctz64(srest) : e;
                if (s > e) s = e;
                if (SCATTER && s > pos) sink.copy(0, pos, s - pos, n.ids);
                n.ids += s - pos;
                p = s;
                if (s < e) { if (SCATTER) sink.put(0, n.ids, 0); n.ids++; st.sp = 1; p = s + 1; }
            }
            if (p < e) { if (SCATTER) sink.copy(1, p, e - p, n.comm); n.comm += e - p; }
            if (has_nl) {
                if (!st.sp) { if (SCATTER) sink.put(0, n.ids, 0); n.ids++; }
                if (SCATTER) sink.put(1, n.comm, 0);
                n.comm++;
            }
            break;
        }
        case FR_SEQ:
            if (len) { if (SCATTER) sink.copy(2, pos, len, n.seq); n.seq += len; }
            if (has_nl && !FASTQ) {                                  // line end: longest-line bookkeeping (process.c:389-393)
                const u64 v = o_cnt + n.seq;
                if (v - ln.base > ln.max) ln.max = v - ln.base;
                ln.base = v; ln.mark = n.seq + 1;
            }
            break;
        case FR_PLUS: break;                                         // process.c:516: content ignored
        default:
            if (len) { if (SCATTER) sink.copy(3, pos, len, n.qual); n.qual += len; }
            if (has_nl) {                                            // process.c:531-533: the record is complete
                if (SCATTER) sink.rec_end(o_rec + n.rec, o_cnt + n.seq, o_qual + n.qual, lo + e);
                n.rec++;
            }
            break;
        }
        if (has_nl) { pos = e + 1; st.ls = 1; st.sp = 0; if (FASTQ) st.role = (st.role + 1) & 3; }
        else pos = e;
    }
}

// Has the header line containing text[at-1] already had its first space before `at`?  (bounded backward scan)
HD u32 fast_lookback_space(const u8 *text, u64 p0, u64 at, u32 &flag)
{
    u64 q = at;
    for (u32 k = 0; q > p0; k++) {
        if (k >= FAST_LOOKBACK) { flag |= FF_LOOKBACK; return 0; }
        const u8 c = text[--q];
        if (c == ' ') return 1;
        if (c == '\n') return 0;
    }
    return 0;
}

// SWAR checks of consumed bytes (C4 for protein / text, C5)
HD u32 swar_bad_qual(u32 v) { return (v & 0x80808080u) | swar_haszero(v ^ 0x20202020u); }      // >= 128 or ' ' (C1 excluded the rest)
HD u32 swar_bad_text(u32 v, bool gt_bad) { return swar_haszero(v ^ 0x20202020u) | (gt_bad ? swar_haszero(v ^ 0x3E3E3E3Eu) : 0u); }
HD u32 swar_bad_protein(u32 v)           // tables.c:104: '*', '-', letters of either case
{
    u32 bad = 0;
    for (int j = 0; j < 4; j++) {
        const u32 c = (v >> (8 * j)) & 0xFF, u = c | 0x20;
        if (!((u >= 'a' && u <= 'z') || c == '*' || c == '-')) bad = 1;
    }
    return bad;
}
HD u32 swar_upper(u32 v)                 // toupper() on four bytes (process.c:49)
{
    u32 r = 0;
    for (int j = 0; j < 4; j++) { u32 c = (v >> (8 * j)) & 0xFF; if (c >= 'a' && c <= 'z') c -= 32; r |= c << (8 * j); }
    return r;
}

// ---- staging sink (shared memory on the device, plain arrays in tests/emu) ----
// tile: the text of the tile, linear offset = 64 * thread + i;  stage: one region per stream; both in fast_pad() layout.
//   run word: src (7 bits, offset in my chunk) | len << 7 (7 bits; 0 = one NUL byte) | stage offset << 16
// (Deferring the copies until after the walk, so that all lanes copy together, was measured slower on B200: 9.4 vs 7.7 ms.)
struct FastSmemSink {
    const u8 *tile; u8 *stage;
    u32 src0;                      // linear tile offset of my chunk
    u32 base[4];                   // linear stage offset where MY bytes of each stream start
    u64 *rec_seq_end, *rec_qual_end, *rec_pos; bool fastq;
    u32 nrun;

    HD u32 ldw(u32 word) const { return *(const u32 *)(tile + fast_pad(word << 2)); }
    HD void do_copy(u32 run) const
    {
        u32 len = (run >> 7) & 127, D = run >> 16;
        if (len == 0) { stage[fast_pad(D)] = 0; return; }
        u32 S = src0 + (run & 127);
        while (len && (D & 3)) { stage[fast_pad(D)] = tile[fast_pad(S)]; D++; S++; len--; }
        if (len >= 4) {
            const u32 shift = (S & 3) * 8;
            u32 wi = S >> 2, cur = ldw(wi);
            while (len >= 4) {
                const u32 nxt = ldw(wi + 1);
                const u32 v = shift ? (cur >> shift) | (nxt << (32 - shift)) : cur;
                *(u32 *)(stage + fast_pad(D)) = v;
                cur = nxt; wi++; D += 4; S += 4; len -= 4;
            }
        }
        while (len) { stage[fast_pad(D)] = tile[fast_pad(S)]; D++; S++; len--; }
    }
    HD void push(u32 run) { do_copy(run); nrun++; }
    HD void begin() { nrun = 0; }
    HD void flush() const {}
    HD void copy(u32 stream, u32 src, u32 len, u32 dst_off) { push(src | (len << 7) | ((base[stream] + dst_off) << 16)); }
    HD void put(u32 stream, u32 dst_off, u8) { push((base[stream] + dst_off) << 16); }
    HD void rec_end(u64 r, u64 cnt, u64 q, u64 pos) const { rec_seq_end[r] = cnt; if (fastq) rec_qual_end[r] = q; rec_pos[r] = pos; }
};

struct FastNoSink {
    HD void copy(u32, u32, u32, u32) const {}
    HD void put(u32, u32, u8) const {}
    HD void rec_end(u64, u64, u64, u64) const {}
};

// parser state at the end of the input, as the general FSM would name it (naf_parse.cuh FA_* / FQ_* values)
HD u32 fast_end_state(bool fastq, u32 role, u32 sp, u32 ls)
{
    if (!fastq) return role == FR_HDR && !ls ? (sp ? 1u : 0u) : (ls ? 2u : 3u);         // FA_NAME, FA_COMMENT, FA_SEQ_LS, FA_SEQ_MID
    switch (role) {
    case FR_HDR:  return ls ? 7u : (sp ? 1u : 0u);                                       // FQ_AFTER_QUAL, FQ_COMMENT, FQ_NAME
    case FR_SEQ:  return 2u;                                                             // FQ_SEQ
    case FR_PLUS: return ls ? 3u : 4u;                                                   // FQ_AFTER_SEQ, FQ_PLUS
    default:      return ls ? 5u : 6u;                                                   // FQ_BEFORE_QUAL, FQ_QUAL
    }
}

}  // namespace nafg
