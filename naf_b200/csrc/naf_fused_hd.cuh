// naf_fused_hd.cuh — the single-pass encode transform: text -> ids / comments / lengths / 4-bit sequence + case bits /
// quality, one read of the text, every output byte written once (host + device logic; kernel: naf_fused.cuh).
//
// Replaces, for canonical input, ennaf/src/process.c:358,477 (record split), encoders.c:30 (encode_dna),
// encoders.c:72 (add_length) and the case predicate of encoders.c:134 in ONE kernel.  "Canonical" is what
// naf_fast_hd.cuh defines (LF line ends, no stray white space / control bytes, only expected codes); every condition
// is verified while the data is moved and a violation raises a flag -- the caller then redoes the split with the
// general, process.c-exact parser (naf_parse.cuh), so results are always the reference's.
//
// Work unit: a tile of FT_BYTES of text, one CTA.  Per tile
//   1  the tile is brought to shared memory once (bulk copy); '\n' positions by SIMD-in-register compares -> the tile's
//      line segments (a segment = a line or the part of a line inside the tile)
//   2  look-back #1 (single-pass chained scan over tiles): which kind of line the tile starts in
//      (FASTQ: lines so far mod 4;  FASTA: header / sequence / line start)
//   3  one thread per segment: role, first space of a header, bytes per stream; block scan -> offsets inside the tile
//   4a look-back #2: bytes per stream, records, bases before the tile + what a record / line / byte that straddles
//      the tile boundary needs (bases and qualities since the last record end, bases since the last line end, last base)
//   4  look-back #2 runs in one warp WHILE the other warps gather the tile's bytes per stream into a staging area (8-lane
//      groups, one copy descriptor = one run of bytes of a line, aligned words built from two source words by a funnel shift)
//   5  staging -> global as aligned 16-byte pieces (the expected-byte checks happen here, on full words): names, comments,
//      quality (and protein / text sequence); bases as 32-base pieces -> 16 bytes of 4-bit codes + one word of case bits
//      (a byte shared by two tiles is written by the later one, which knows the earlier base from look-back #2);
//      per record: length unit, quality-length check
// All of it is plain C++ over explicit (thread id, thread count) so that tests/emu/emu_fused.cpp runs the same phases
// on the CPU, thread after thread, against the oracle.
#pragma once
#include "naf_fast_hd.cuh"

namespace nafg {

#ifndef FT_TILE_BYTES
#define FT_TILE_BYTES 32768          // (16 KB tiles: 13 % slower on B200 -- twice the look-backs per byte.)  tests/emu builds a second emulation with tiny tiles, so that small inputs cross many tile boundaries
#endif
static const u32 FT_BYTES = FT_TILE_BYTES, FT_CHUNKS = FT_BYTES / 16, FT_MAXSEG = FT_BYTES / 16, FT_GROUP = 8;
static const u32 FT_STAGE = FT_BYTES + 160;

enum : u32 { FU_BADBYTE = 1, FU_FIRSTCHAR = 2, FU_BLANK = 4, FU_SEQ = 8, FU_QUAL = 16, FU_LOOKBACK = 32,
             FU_LINES = 64,        // more line segments in a tile than the tables hold (lines shorter than 16 bytes on average)
             FU_QLEN = 128,        // a record's quality length differs from its sequence length (process.c:531)
             FU_BIGREC = 256,      // a sequence of 2^32-1 bases or more: continuation length units (encoders.c:78) are the general path's
             FU_TRUNC = 512 };     // FASTQ input ends inside a record

enum { FS_PACK4 = 0, FS_PROTEIN = 1, FS_TEXT = 2, FS_TEXT_GT = 3 };      // what the sequence stream holds / which bytes are expected

// ---- look-back #2 state: what the tiles before a tile contributed.  Composes associatively (older first).
struct F2 {
    u64 ids, comm, seq, qual, rec;    // bytes per stream (seq: bases), records ended
    u64 srec;                         // bases since the last record boundary
    u64 qrec;                         // FASTQ: quality bytes since the last record boundary;  FASTA: bases since the last sequence-line end
    u64 last;                         // bits 0-7 last base (text byte), bit 8 there is one, bit 9 a record boundary was seen, bit 10 a line end was seen
};
static const u32 F2_B = 1u << 8, F2_R = 1u << 9, F2_L = 1u << 10;
static const int F2_WORDS = 8;

HD F2 f2_compose(const F2 &a, const F2 &b, bool fastq)
{
    F2 r;
    r.ids = a.ids + b.ids; r.comm = a.comm + b.comm; r.seq = a.seq + b.seq; r.qual = a.qual + b.qual; r.rec = a.rec + b.rec;
    r.srec = (b.last & F2_R) ? b.srec : a.srec + b.srec;
    r.qrec = (b.last & (fastq ? F2_R : F2_L)) ? b.qrec : a.qrec + b.qrec;
    if (r.srec > 0xFFFFFFFFull) r.srec = 0xFFFFFFFFull;                 // saturate: 2^32 - 1 or more is FU_BIGREC either way, and the
    if (r.qrec > 0xFFFFFFFFull) r.qrec = 0xFFFFFFFFull;                 // look-back records keep 32 bits of these two
    r.last = ((b.last & F2_B) ? (b.last & 0x1FF) : (a.last & 0x1FF)) | ((a.last | b.last) & (F2_R | F2_L));
    return r;
}
HD F2 f2_initial() { F2 r; r.ids = r.comm = r.seq = r.qual = r.rec = r.srec = r.qrec = 0; r.last = F2_R | F2_L; return r; }

HD u32 f1_compose(bool fastq, u32 a, u32 b) { return fastq ? a + b : fe_compose(a, b); }

// ---- SIMD-in-register byte predicates: 0x80 in every byte that satisfies it (exact, no cross-byte carries)
HD u32 sw_eq(u32 v, u32 c4) { const u32 t = v ^ c4; return ~(((t & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | t) & 0x80808080u; }
HD u32 sw_lt7(u32 l, u32 k) { return ~(l + (0x80u - k) * 0x01010101u) & 0x80808080u; }     // l: bytes < 128;  l < k (k <= 128)
HD u32 sw_movemask(u32 z) { return (((z >> 7) * 0x00204081u) >> 21) & 15u; }                // 0x80 flags -> 4 bits
HD u32 sw_bad_qual(u32 v)  { const u32 l = v & 0x7F7F7F7Fu; return (v & 0x80808080u) | sw_lt7(l, 33) | ((l + 0x01010101u) & 0x80808080u); }   // not 33..126 (tables.c:137)
HD u32 sw_bad_text(u32 v)  { const u32 l = v & 0x7F7F7F7Fu; return (sw_lt7(l, 33) & ~v) | ((l + 0x01010101u) & 0x80808080u); }              // <= 32, 127, 255 (tables.c:115)
HD u32 sw_bad_comm(u32 v)  { const u32 l = v & 0x7F7F7F7Fu; return (sw_lt7(l, 32) & ~v) | ((l + 0x01010101u) & 0x80808080u); }              // < 32, 127, 255 (tables.c:126)
HD u32 sw_bad_protein(u32 v)                                                                                                                // tables.c:104
{
    const u32 u = (v & 0x7F7F7F7Fu) | 0x20202020u;
    const u32 alpha = ~sw_lt7(u, 'a') & sw_lt7(u, 'z' + 1) & 0x80808080u;
    return ((v & 0x80808080u) | ~(alpha | sw_eq(v, 0x2A2A2A2Au) | sw_eq(v, 0x2D2D2D2Du))) & 0x80808080u;
}
HD u32 sw_upper(u32 v)                                                                                                                      // toupper() x 4 (process.c:49)
{
    const u32 l = v & 0x7F7F7F7Fu;
    const u32 lower = ~sw_lt7(l, 'a') & sw_lt7(l, 'z' + 1) & ~v & 0x80808080u;
    return v - (lower >> 2);
}
enum { FC_NONE = 0, FC_QUAL = 1, FC_ID = 2, FC_ID_GT = 3, FC_COMM = 4, FC_PROTEIN = 5, FC_TEXT = 6, FC_TEXT_GT = 7 };
HD u32 sw_check(int kind, u32 v)
{
    switch (kind) {
    case FC_QUAL:    return sw_bad_qual(v);
    case FC_ID:      return sw_bad_text(v);
    case FC_ID_GT:   return sw_bad_text(v) | sw_eq(v, 0x3E3E3E3Eu);       // ennaf.c:466: in text FASTA '>' is unexpected in names too
    case FC_COMM:    return sw_bad_comm(v);
    case FC_PROTEIN: return sw_bad_protein(v);
    case FC_TEXT:    return sw_bad_text(v);
    case FC_TEXT_GT: return sw_bad_text(v) | sw_eq(v, 0x3E3E3E3Eu);
    default:         return 0;
    }
}

HD u32 funnel_r(u32 lo, u32 hi, u32 shift_bits)       // (hi:lo) >> shift_bits, shift_bits in {0, 8, 16, 24}
{
#ifdef __CUDA_ARCH__
    return __funnelshift_r(lo, hi, shift_bits);
#else
    return shift_bits ? (lo >> shift_bits) | (hi << (32 - shift_bits)) : lo;
#endif
}

// `len` bytes of the tile (shared memory, any alignment) -> staging (shared memory, any alignment), by the FT_GROUP lanes of a
// group: aligned words of the destination from two aligned words of the tile, head and tail bytes one lane each.
// Names and comments are checked here (check != FC_NONE), because their terminators share the staged region.
template <bool CHECK> HD u32 group_copy(const u8 *tile, u32 src, u32 len, u8 *dst, u32 lane, int check)
{
    u32 head = (u32)((4 - ((uintptr_t)dst & 3)) & 3), bad = 0;
    if (head > len) head = len;
    const u32 nw = (len - head) >> 2, done = head + (nw << 2), tail = len - done;
    if (lane < head || (lane >= 4 && lane - 4 < tail)) {
        const u32 i = lane < head ? lane : done + lane - 4;
        const u32 c = tile[src + i];
        if (CHECK) bad |= sw_check(check, c * 0x01010101u);
        dst[i] = (u8)c;
    }
    const u32 s0 = src + head, sh = (s0 & 3) * 8;
    const u32 *tw = (const u32 *)(tile + (s0 & ~3u));
    u32 *dw = (u32 *)(dst + head);
    for (u32 w = lane; w < nw; w += FT_GROUP) {
        const u32 v = funnel_r(tw[w], tw[w + 1], sh);       // the tile is padded: tw[w + 1] is readable
        if (CHECK) bad |= sw_check(check, v);
        dw[w] = v;
    }
    return bad;
}

// 16 * NQ bytes at any byte offset of a 16-byte aligned shared-memory array: NQ + 1 aligned 16-byte loads (conflict-free when
// neighbouring lanes read neighbouring pieces), then a window of words picked by the offset's word part -- which is the same
// for every piece of a region, so the switch is uniform -- and a funnel shift by its byte part.
template <int NQ> HD void load_unaligned(const u8 *base, u32 off, u32 *out)
{
    const uint4 *p = (const uint4 *)(base + (off & ~15u));
    u32 w[4 * NQ + 4];
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int k = 0; k <= NQ; k++) { const uint4 q = p[k]; w[4 * k] = q.x; w[4 * k + 1] = q.y; w[4 * k + 2] = q.z; w[4 * k + 3] = q.w; }
    const u32 sh = (off & 3) * 8;
    switch ((off >> 2) & 3) {
    case 0:
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
        for (int k = 0; k < 4 * NQ; k++) out[k] = funnel_r(w[k], w[k + 1], sh);
        break;
    case 1:
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
        for (int k = 0; k < 4 * NQ; k++) out[k] = funnel_r(w[k + 1], w[k + 2], sh);
        break;
    case 2:
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
        for (int k = 0; k < 4 * NQ; k++) out[k] = funnel_r(w[k + 2], w[k + 3], sh);
        break;
    default:
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
        for (int k = 0; k < 4 * NQ; k++) out[k] = funnel_r(w[k + 3], k + 4 < 4 * NQ + 4 ? w[k + 4] : 0u, sh);
        break;
    }
}

// ---- per-tile scalars (shared memory)
struct FusedShared {
    u32 tile, live_lo, live_hi;        // live bytes of the tile: [live_lo, live_hi) (tile-relative)
    u32 nseg;                          // line segments: newlines + 1 (the last one may be empty)
    u32 entry1, entry_ls, entry_sp;    // look-back #1 prefix; the tile starts at a line start; the header it starts in has had its space
    u32 t_ids, t_comm, t_seq, t_qual, t_rec, n_hdr, n_seq, n_qual;    // totals of the tile
    u32 abort_;                        // tables exceeded: publish, skip the rest
    u32 flag;                          // FU_* raised by this tile
    u32 s_ids, s_comm, s_seq, s_qual;  // staging offsets of the four regions
    F2 pre;                            // look-back #2 prefix
    u64 maxlen;                        // longest line (FASTA) / read (FASTQ) ended in this tile
};

struct FusedCfg {
    u64 n, p0;                         // text size, first byte after the leading '>' / '@'
    int fastq, seq_mode, upper, want_mask, id_check;
    const u32 *lut;                    // nuc_code | "unexpected" << 16 (4-bit mode)
    u8 *ids, *comm, *seq, *qual;       // destinations (seq: packed codes or bytes)
    u32 *len, *casebits;
};
HD u32 nuc_lut32(u8 l) { return (u32)(l & 15) | ((u32)(l >> 7) << 16); }      // from the u8 table (bit 7 = unexpected)

// what classify() hands to place() about one segment
enum : u32 { SR_ROLE = 3, SR_NL = 4, SR_LS = 8, SR_REC = 16, SR_SKIP1 = 32, SR_SEEN = 64, SR_NONE = 128 };
static const u32 FT_MAXDESC = 2 * FT_MAXSEG;

// Is there a ' ' after the last '\n' of the `np` bytes before the tile?  (the header the tile starts in has had its first space)
HD u32 prev_scan(const u8 *prev, u32 np, bool &resolved)
{
    resolved = true;
    for (u32 i = np; i-- > 0;) { const u8 c = prev[i]; if (c == ' ') return 1; if (c == '\n') return 0; }
    resolved = false;
    return 0;
}

struct FusedTile {
    u8 *text, *stage;
    u16 *seg_end, *d_src, *d_len, *d_dst, *recseq, *recqual;
    FusedShared *sh;

    HD u32 seg_start(u32 j) const { return j ? (u32)seg_end[j - 1] + 1 : sh->live_lo; }
    // Where the content of segment j (starting at s) ends: at its '\n', or at the end of the live range.  FASTA: a '\r' right
    // before the '\n' is part of the line end -- '\r' is an end-of-line byte to process.c (tables.c:28 is_eol_arr) and runs of
    // them collapse (process.c:383-399), so CR LF files split exactly like LF files.  The '\n' may be the first byte of the
    // next tile (text[FT_BYTES] is readable and holds it).  Any other '\r' stays content and fails the checks of what consumes
    // it.  (FASTQ: the reference dies on CR LF input; the '\r' stays content here too and the general parser reports it.)
    HD u32 seg_content_end(bool fastq, u32 j, u32 s, bool has_nl) const
    {
        const u32 e = has_nl ? (u32)seg_end[j] : sh->live_hi;
        if (!fastq && e > s && text[e - 1] == '\r' && (has_nl || (e == FT_BYTES && text[FT_BYTES] == '\n'))) return e - 1;
        return e;
    }

    // phase 1: newline mask of chunk c (16 bytes), restricted to the live range
    HD u32 chunk_mask(u32 c) const
    {
        const u32 *w = (const u32 *)(text + 16 * c);
        u32 m = 0;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
        for (int k = 0; k < 4; k++) m |= sw_movemask(sw_eq(w[k], 0x0A0A0A0Au)) << (4 * k);
        const u32 lo = 16 * c, L = sh->live_lo, H = sh->live_hi;
        if (L != 0 || H != FT_BYTES) {
            if (lo < L) m &= L - lo >= 16 ? 0u : ~0u << (L - lo);
            if (lo + 16 > H) m &= H <= lo ? 0u : ~(~0u << (H - lo));
        }
        return m & 0xFFFFu;
    }
    // phase 2: positions of my newlines -> seg_end[first + k]
    HD void put_lines(u32 c, u32 m, u32 first) const
    {
        while (m) {
#ifdef __CUDA_ARCH__
            const u32 b = __ffs(m) - 1;
#else
            const u32 b = (u32)__builtin_ctz(m);
#endif
            m &= m - 1;
            if (first < FT_MAXSEG) seg_end[first] = (u16)(16 * c + b);
            first++;
        }
    }
    // FASTA element of the tile for look-back #1 (naf_fast_hd.cuh fe_*): T = newlines in the live range
    HD u32 fasta_element(u32 T, u32 last_nl) const
    {
        if (sh->live_lo >= sh->live_hi) return FE_ID;
        if (T == 0) return text[sh->live_lo] == '>' ? (u32)FE_NONE_GT : (u32)FE_NONE_OT;
        if (last_nl + 1 >= sh->live_hi) return FE_LS;
        return text[last_nl + 1] == '>' ? (u32)FE_HDR : (u32)FE_SEQ;
    }

    // phase 3a: classify segment j; returns its contribution to the two packed scans
    //   A: ids | comm << 16 | seq << 32 | qual << 48        B: rec | n_hdr << 16 | n_seq << 32 | n_qual << 48
    // and what place() needs: role byte (SR_*), position of the header's first space (0xFFFF: none)
    HD void classify(const FusedCfg &C, u32 j, u64 &A, u64 &B, u32 &rb, u32 &sp, u32 &flag) const
    {
        const u32 nseg = sh->nseg;
        const bool has_nl = j + 1 < nseg;
        const u32 s = seg_start(j), e = seg_content_end(C.fastq != 0, j, s, has_nl);
        const bool ls = j ? true : sh->entry_ls != 0;
        u32 role, skip1 = 0, rec = 0;
        A = 0; B = 0; sp = 0xFFFF;
        if (!has_nl && s >= e) { rb = SR_NONE; return; }                      // empty tail: nothing
        if (C.fastq) {
            role = (sh->entry1 + j) & 3;
            if (ls) {
                const u32 c = s < e ? text[s] : (u32)'\n';
                if (c == '\n') flag |= FU_BLANK;
                if (role == FR_HDR) { if (c == '@') skip1 = 1; else flag |= FU_FIRSTCHAR; }
                else if (role == FR_PLUS && c != '+') flag |= FU_FIRSTCHAR;
            }
        } else {
            if (ls) {
                if (s < e && text[s] == '>') { role = FR_HDR; skip1 = 1; rec = 1; }      // process.c:383: the previous record ends here
                else role = FR_SEQ;
            } else role = sh->entry1 == FE_HDR ? (u32)FR_HDR : (u32)FR_SEQ;
        }
        const u32 b = s + skip1, len = e - b;
        u32 seen = 0;
        switch (role) {
        case FR_HDR: {
            u32 ids = 0, comm = 0;
            seen = j == 0 && !ls && sh->entry_sp;
            if (!seen) {
                u32 p = b;
                while (p < e && text[p] != ' ') p++;
                ids = p - b;
                if (p < e) { sp = p; ids++; comm = e - p - 1; }
                else if (has_nl) ids++;
            } else comm = len;
            if (has_nl) comm++;
            A = (u64)ids | ((u64)comm << 16);
            B = 1ull << 16;
            break;
        }
        case FR_SEQ:
            A = (u64)len << 32; B = 1ull << 32;
            break;
        case FR_PLUS:                                                         // process.c:516: content ignored -- but it must not hold
            for (u32 p = b; p < e; p++) if (text[p] < 32) flag |= FU_BADBYTE;  // a byte the general machine treats as a line end
            break;
        default:
            A = (u64)len << 48; B = 1ull << 48;
            if (has_nl) rec = 1;                                              // process.c:531-533: the record is complete
            break;
        }
        B |= rec;
        rb = role | (has_nl ? SR_NL : 0) | (ls ? SR_LS : 0) | (rec ? SR_REC : 0) | (skip1 ? SR_SKIP1 : 0) | (seen ? SR_SEEN : 0);
    }

    // staging layout (tile-local: known as soon as the tile's totals are)
    HD void layout() const
    {
        sh->s_ids = 0;
        sh->s_comm = (sh->t_ids + 15) & ~15u;
        sh->s_seq = ((sh->s_comm + sh->t_comm + 15) & ~15u) + 32;          // the first 32-base piece may start up to 31 bytes before the region
        sh->s_qual = (sh->s_seq + sh->t_seq + 15) & ~15u;
    }

    // phase 3b: copy descriptors of segment j (runs of bytes -> staging), terminators, record marks.
    // a, b: exclusive prefixes of the packed scans at this segment.
    HD void place(const FusedCfg &C, u32 j, u64 a, u64 b, u32 rb, u32 sp) const
    {
        if (rb & SR_NONE) return;
        const u32 role = rb & SR_ROLE;
        const u32 o_ids = (u32)(a & 0xFFFF), o_comm = (u32)((a >> 16) & 0xFFFF), o_seq = (u32)((a >> 32) & 0xFFFF), o_qual = (u32)(a >> 48);
        const u32 k_rec = (u32)(b & 0xFFFF), k_hdr = (u32)((b >> 16) & 0xFFFF), k_seq = (u32)((b >> 32) & 0xFFFF), k_qual = (u32)(b >> 48);
        const bool has_nl = rb & SR_NL;
        const u32 s0 = seg_start(j), s = s0 + ((rb & SR_SKIP1) ? 1 : 0), e = seg_content_end(C.fastq != 0, j, s0, has_nl);
        switch (role) {
        case FR_HDR: {
            const u32 kn = sh->n_seq + sh->n_qual + k_hdr, kc = kn + sh->n_hdr;
            u32 nlen = 0, cs = s;
            if (!(rb & SR_SEEN)) {
                nlen = (sp != 0xFFFF ? sp : e) - s;
                if (sp != 0xFFFF || has_nl) stage[sh->s_ids + o_ids + nlen] = 0;
                cs = sp != 0xFFFF ? sp + 1 : e;
            }
            d_src[kn] = (u16)s; d_len[kn] = (u16)nlen; d_dst[kn] = (u16)(sh->s_ids + o_ids);
            d_src[kc] = (u16)cs; d_len[kc] = (u16)(e - cs); d_dst[kc] = (u16)(sh->s_comm + o_comm);
            if (has_nl) stage[sh->s_comm + o_comm + (e - cs)] = 0;
            break;
        }
        case FR_SEQ:  d_src[k_seq] = (u16)s; d_len[k_seq] = (u16)(e - s); d_dst[k_seq] = (u16)(sh->s_seq + o_seq); break;
        case FR_QUAL: { const u32 k = sh->n_seq + k_qual; d_src[k] = (u16)s; d_len[k] = (u16)(e - s); d_dst[k] = (u16)(sh->s_qual + o_qual); break; }
        default: break;
        }
        if (rb & SR_REC) {
            recseq[k_rec] = (u16)o_seq;                                                 // bases of the tile before this boundary
            if (C.fastq) recqual[k_rec] = (u16)(o_qual + (e - s0));                       // quality bytes up to and including this line
        }
    }

    // what the tile hands to its successors (look-back #2 aggregate); call after place() of every segment
    HD F2 aggregate(const FusedCfg &C) const
    {
        F2 g;
        g.ids = sh->t_ids; g.comm = sh->t_comm; g.seq = sh->t_seq; g.qual = sh->t_qual; g.rec = sh->t_rec;
        g.last = 0;
        const u32 nr = sh->t_rec;
        g.srec = nr ? sh->t_seq - recseq[nr - 1] : sh->t_seq;
        if (nr) g.last |= F2_R;
        if (C.fastq) g.qrec = nr ? sh->t_qual - recqual[nr - 1] : sh->t_qual;
        else {
            // bases after the last sequence-line end of the tile (every sequence run but the tile's last segment ends a line)
            g.qrec = sh->t_seq;
            for (u32 k = sh->n_seq; k-- > 0;) {
                if ((u32)d_src[k] + d_len[k] < sh->live_hi) { g.qrec = sh->t_seq - ((u32)d_dst[k] - sh->s_seq + d_len[k]); g.last |= F2_L; break; }
            }
        }
        for (u32 k = sh->n_seq; k-- > 0;)                                                 // last base of the tile
            if (d_len[k]) { g.last |= F2_B | text[(u32)d_src[k] + d_len[k] - 1]; break; }
        return g;
    }

    // phase 4: copy descriptor k, by the lanes of one group
    HD u32 copy_desc(const FusedCfg &C, u32 k, u32 lane) const
    {
        const u32 len = d_len[k], nsq = sh->n_seq + sh->n_qual;
        if (!len) return 0;
        if (k < nsq) { group_copy<false>(text, d_src[k], len, stage + d_dst[k], lane, FC_NONE); return 0; }   // sequence / quality: checked on their way out
        const int check = k < nsq + sh->n_hdr ? C.id_check : (int)FC_COMM;
        return group_copy<true>(text, d_src[k], len, stage + d_dst[k], lane, check) ? (u32)FU_BADBYTE : 0u;
    }

    // phase 5: record k of the tile ends -> its length unit, the quality-length check, the longest read
    HD u64 finish_record(const FusedCfg &C, u32 k, u32 &flag) const
    {
        const F2 &P = sh->pre;
        const u64 sl = (u64)recseq[k] - (k ? recseq[k - 1] : 0) + (k ? 0 : P.srec);
        if (C.fastq) {
            const u64 ql = (u64)recqual[k] - (k ? recqual[k - 1] : 0) + (k ? 0 : P.qrec);
            if (ql != sl) flag |= FU_QLEN;
        }
        if (sl >= 0xFFFFFFFFull) flag |= FU_BIGREC;
        C.len[P.rec + k] = (u32)sl;
        return C.fastq ? sl : 0;
    }
    // FASTA: the sequence line of descriptor k < n_seq ends -> its length (process.c:389-393)
    HD u64 line_length(u32 k) const
    {
        if ((u32)d_src[k] + d_len[k] >= sh->live_hi) return 0;                            // no '\n' after it inside the tile
        u64 L = d_len[k];
        // the first line end of the tile closes whatever the earlier tiles left open
        if (k == 0) L += sh->pre.qrec;
        return L;
    }

    // phase 5: 16-byte unit u of a staged region -> global (dst + 16 u is 16-byte aligned).  Returns 0x80 flags of bytes
    // that fail the check.
    HD u32 out_unit(u32 region_off, u32 u, u8 *dst, int check, bool upper) const
    {
        u32 v[4];
        load_unaligned<1>(stage, region_off + 16 * u, v);
        u32 bad = 0;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
        for (int k = 0; k < 4; k++) { bad |= sw_check(check, v[k]); if (upper) v[k] = sw_upper(v[k]); }
        *(uint4 *)(dst + 16 * u) = make_uint4(v[0], v[1], v[2], v[3]);
        return bad;
    }
    HD u32 out_byte(u32 region_off, u32 i, u8 *dst, int check, bool upper) const
    {
        u32 v = (u32)stage[region_off + i] * 0x01010101u;
        const u32 bad = sw_check(check, v);
        if (upper) v = sw_upper(v);
        dst[i] = (u8)v;
        return bad;
    }

    // phase 5: piece q of the staged bases (32 bases at a multiple of 32 in the file's base numbering) -> 16 bytes of
    // codes + one word of case bits.  Returns FU_SEQ if a base is not an expected code.
    template <class AtomicOr>
    HD u32 pack_piece(const FusedCfg &C, u32 q, AtomicOr atomic_or) const
    {
        const F2 &P = sh->pre;
        const u64 S = P.seq, A = S & 31, g0 = S - A + 32ull * q;         // first base of the piece
        const u32 lo = q == 0 ? (u32)A : 0u;
        const u64 endb = S + sh->t_seq;
        const u32 hi = g0 + 32 <= endb ? 32u : (u32)(endb - g0);
        u32 x[8];
        load_unaligned<2>(stage, sh->s_seq - (u32)A + 32 * q, x);
        u32 out[4] = {0, 0, 0, 0}, cbits = 0, inv_any = 0, r[8];
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
        for (int k = 0; k < 8; k++) {
            const u32 v = x[k];
            r[k] = C.lut[v & 0xFF] + (C.lut[(v >> 8) & 0xFF] << 4) + (C.lut[(v >> 16) & 0xFF] << 8) + (C.lut[v >> 24] << 12);
            out[k >> 1] |= (r[k] & 0xFFFF) << (16 * (k & 1));
            inv_any |= r[k];
            const u32 t = ((v & (v << 1)) >> 6) & 0x01010101u;            // byte >= 96 (encoders.c:134), exact for bytes < 128
            cbits |= (((t * 0x01020408u) >> 24) & 15u) << (4 * k);
        }
        u8 *dst = C.seq + (g0 >> 1);
        if (lo == 0 && hi == 32) {
            *(uint4 *)dst = make_uint4(out[0], out[1], out[2], out[3]);
            if (C.want_mask) C.casebits[g0 >> 5] = cbits;
            return (inv_any >> 16) ? (u32)FU_SEQ : 0u;
        }
        // partial piece (the tile's first / last): only positions [lo, hi) are mine
        const u32 live = (hi >= 32 ? ~0u : ~(~0u << hi)) & (~0u << lo);
        u32 inv = 0;
        for (int k = 0; k < 8; k++) { const u32 f = (r[k] >> 16) & 0x1111u; inv |= (((f * 0x1248u) >> 12) & 15u) << (4 * k); }
        cbits &= live; inv &= live;
        // a byte shared with the tile before me is mine (I know its low nibble from look-back #2); a dangling last
        // low nibble is my successor's (or the finishing step's, at the end of the input)
        for (u32 b = lo >> 1; 2 * b + 1 < hi; b++) {
            u32 v = (out[b >> 2] >> (8 * (b & 3))) & 0xFF;
            if (2 * b < lo) v = (v & 0xF0) | (C.lut[P.last & 0xFF] & 15);
            dst[b] = (u8)v;
        }
        if (C.want_mask && cbits) atomic_or(C.casebits + (g0 >> 5), cbits);
        return inv ? (u32)FU_SEQ : 0u;
    }
};

// The end of the input, from the final states of both look-backs (process.c:417-425, :535-543; ennaf.c:525): pending
// terminators, the last record, the pending last line, the odd last nibble.
struct FusedTotals {
    u64 n_ids, n_comm, n_bases, n_qual, n_rec, longest;
    u32 flag, end_state;
};

HD void fused_finish(const FusedCfg &C, u32 final1, const F2 &fin, u32 flag_in, u64 longest_in, const u8 *gtext, FusedTotals &T)
{
    u32 flag = flag_in;
    u32 role, ls, sp = 0;
    if (C.fastq) { role = final1 & 3; ls = C.n > C.p0 && gtext[C.n - 1] == '\n'; }
    else { role = final1 == FE_HDR ? (u32)FR_HDR : (u32)FR_SEQ; ls = final1 == FE_LS; }
    if (role == FR_HDR && !ls) { u32 f = 0; sp = fast_lookback_space(gtext, C.p0, C.n, f); if (f) flag |= FU_LOOKBACK; }
    const u32 es = fast_end_state(C.fastq != 0, role, sp, ls);
    u64 n_ids = fin.ids, n_comm = fin.comm, n_rec = fin.rec, longest = longest_in;
    if (!C.fastq) {
        if (es == 0) { C.ids[n_ids++] = 0; C.comm[n_comm++] = 0; }                 // FA_NAME
        else if (es == 1) C.comm[n_comm++] = 0;                                     // FA_COMMENT
        if (fin.srec >= 0xFFFFFFFFull) flag |= FU_BIGREC;
        C.len[n_rec++] = (u32)fin.srec;
        if (fin.qrec > longest) longest = fin.qrec;                                 // process.c:417-422: the unterminated last line
    } else if (es == 6) {                                                           // FQ_QUAL: last quality line without '\n'
        if (fin.srec != fin.qrec) flag |= FU_QLEN;
        if (fin.srec >= 0xFFFFFFFFull) flag |= FU_BIGREC;
        C.len[n_rec++] = (u32)fin.srec;
        if (fin.srec > longest) longest = fin.srec;
    } else if (es != 7) flag |= FU_TRUNC;                                           // anything but "after a complete record"
    if (C.seq_mode == FS_PACK4 && (fin.seq & 1)) C.seq[fin.seq >> 1] = (u8)(C.lut[fin.last & 0xFF] & 15);   // ennaf.c:525
    T.n_ids = n_ids; T.n_comm = n_comm; T.n_bases = fin.seq; T.n_qual = fin.qual; T.n_rec = n_rec; T.longest = longest;
    T.flag = flag; T.end_state = es;
}

}  // namespace nafg
