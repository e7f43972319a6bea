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
//   4  look-back #2: bytes per stream, records, bases before the tile + what a record / line / byte that straddles
//      the tile boundary needs (bases and qualities since the last record end, bases since the last line end, last base)
//   5  8-lane groups copy segment after segment: quality (and protein / text sequence) straight to global memory as aligned
//      words, names / comments / bases into a staging area laid out congruent to their destination
//   6  staging -> global: names and comments as 16-byte pieces; bases as 32-base pieces -> 16 bytes of 4-bit codes + one
//      word of case bits (the piece is owned by the tile that holds its bytes: a byte shared by two tiles is written by
//      the later one, which knows the earlier base from look-back #2); per record: length unit, quality-length check
// All of it is plain C++ over explicit (thread id, thread count) so that tests/emu/emu_fused.cpp runs the same phases
// on the CPU, thread after thread, against the oracle.
#pragma once
#include "naf_fast_hd.cuh"

namespace nafg {

#ifndef FT_TILE_BYTES
#define FT_TILE_BYTES 16384          // tests/emu builds a second emulation with tiny tiles, so that small inputs cross many tile boundaries
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

// `len` bytes of the tile (shared memory, any alignment) -> dst (shared or global memory, any alignment), by the FT_GROUP
// lanes of a group: aligned words of dst from two aligned words of the tile, head and tail bytes one lane each.
// Returns 0x80 flags of bytes that fail the check.
template <class Byte>                                 // Byte: u8 (plain) or volatile-free global pointer; same code
HD u32 group_copy(const u8 *tile, u32 src, u32 len, Byte *dst, u32 lane, int check, bool upper)
{
    u32 bad = 0;
    const u32 head = (u32)((4 - ((uintptr_t)dst & 3)) & 3) < len ? (u32)((4 - ((uintptr_t)dst & 3)) & 3) : len;
    const u32 nw = (len - head) >> 2, done = head + (nw << 2), tail = len - done;
    if (lane < head || (lane >= 4 && lane - 4 < tail)) {
        const u32 i = lane < head ? lane : done + (lane - 4);
        u32 v = (u32)tile[src + i] * 0x01010101u;
        bad |= sw_check(check, v);
        if (upper) v = sw_upper(v);
        dst[i] = (u8)v;
    }
    const u32 s0 = src + head, sh = (s0 & 3) * 8;
    const u32 *tw = (const u32 *)(tile + (s0 & ~3u));
    u32 *dw = (u32 *)(dst + head);
    for (u32 w = lane; w < nw; w += FT_GROUP) {
        u32 v = funnel_r(tw[w], tw[w + 1], sh);           // the tile is padded: tw[w + 1] is readable
        bad |= sw_check(check, v);
        if (upper) v = sw_upper(v);
        dw[w] = v;
    }
    return bad;
}

// ---- per-tile scalars (shared memory)
struct FusedShared {
    u32 tile, live_lo, live_hi;        // live bytes of the tile: [live_lo, live_hi) (tile-relative)
    u32 nseg;                          // line segments: newlines + 1 (the last one may be empty)
    u32 entry1, entry_ls, entry_sp;    // look-back #1 prefix; the tile starts at a line start; the header it starts in has had its space
    u32 t_ids, t_comm, t_seq, t_qual, t_rec, n_hdr, n_seq, n_qual;    // totals of the tile
    u32 abort_;                        // tables exceeded: publish, skip the rest
    u32 flag;                          // FU_* raised by this tile
    u32 s_ids, s_comm, s_seq;          // staging offsets of the three staged regions
    F2 pre;                            // look-back #2 prefix
    u64 maxlen;                        // longest line (FASTA) / read (FASTQ) ended in this tile
};

struct FusedCfg {
    u64 n, p0;                         // text size, first byte after the leading '>' / '@'
    int fastq, seq_mode, upper, want_mask, id_check;
    const u8 *lut;                     // nuc_code with bit 7 = unexpected (4-bit mode)
    u8 *ids, *comm, *seq, *qual;       // destinations (seq: packed codes or bytes)
    u32 *len, *casebits;
};

// role byte of a segment
enum : u32 { SR_ROLE = 3, SR_NL = 4, SR_LS = 8, SR_REC = 16, SR_SKIP1 = 32 };      // SR_REC: a record boundary event; SR_SKIP1: first byte is the '@' / '>' marker

struct FusedTile {
    u8 *text, *stage;
    u16 *nlmask, *seg_end, *seg_sp, *seg_off, *seg_offb, *seg_list, *recseq, *recqual;
    u8 *seg_role;
    FusedShared *sh;

    HD u32 seg_start(u32 j) const { return j ? (u32)seg_end[j - 1] + 1 : sh->live_lo; }

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
        if (lo < L) m &= L - lo >= 16 ? 0u : ~0u << (L - lo);
        if (lo + 16 > H) m &= H <= lo ? 0u : ~(~0u << (H - lo));
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
    HD void classify(const FusedCfg &C, u32 j, u64 &A, u64 &B, u32 &flag) const
    {
        const u32 nseg = sh->nseg;
        const bool has_nl = j + 1 < nseg;
        const u32 s = seg_start(j), e = has_nl ? (u32)seg_end[j] : sh->live_hi;
        const bool ls = j ? true : sh->entry_ls != 0;
        u32 role, skip1 = 0, rec = 0;
        A = 0; B = 0;
        if (!has_nl && s >= e) { seg_role[j] = (u8)(FR_PLUS | (ls ? SR_LS : 0)); seg_sp[j] = 0xFFFF; return; }      // empty tail: nothing
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
        u32 sp = 0xFFFF;
        const u32 b = s + skip1, len = e - b;
        switch (role) {
        case FR_HDR: {
            u32 ids = 0, comm = 0;
            const bool seen = j == 0 && !ls && sh->entry_sp;
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
        seg_sp[j] = (u16)sp;
        seg_role[j] = (u8)(role | (has_nl ? SR_NL : 0) | (ls ? SR_LS : 0) | (rec ? SR_REC : 0) | (skip1 ? SR_SKIP1 : 0));
    }

    // phase 3b: offsets of segment j inside the tile's stream regions, its slot in the per-role list, record marks.
    // a, b: exclusive prefixes of the packed scans at this segment.
    HD void place(const FusedCfg &C, u32 j, u64 a, u64 b) const
    {
        const u32 r = seg_role[j], role = r & SR_ROLE;
        const u32 o_ids = (u32)(a & 0xFFFF), o_comm = (u32)((a >> 16) & 0xFFFF), o_seq = (u32)((a >> 32) & 0xFFFF), o_qual = (u32)(a >> 48);
        const u32 k_rec = (u32)(b & 0xFFFF), k_hdr = (u32)((b >> 16) & 0xFFFF), k_seq = (u32)((b >> 32) & 0xFFFF), k_qual = (u32)(b >> 48);
        const bool has_nl = r & SR_NL;
        const u32 s = seg_start(j), e = has_nl ? (u32)seg_end[j] : sh->live_hi;
        if (!has_nl && s >= e) return;
        switch (role) {
        case FR_HDR:  seg_off[j] = (u16)o_ids; seg_offb[j] = (u16)o_comm; seg_list[sh->n_seq + sh->n_qual + k_hdr] = (u16)j; break;
        case FR_SEQ:  seg_off[j] = (u16)o_seq; seg_list[k_seq] = (u16)j; break;
        case FR_QUAL: seg_off[j] = (u16)o_qual; seg_list[sh->n_seq + k_qual] = (u16)j; break;
        default: break;
        }
        if (r & SR_REC) {
            recseq[k_rec] = (u16)o_seq;                                                 // bases of the tile before this boundary
            if (C.fastq) recqual[k_rec] = (u16)(o_qual + (e - s));                        // quality bytes up to and including this line
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
            // bases after the last sequence-line end of the tile
            g.qrec = sh->t_seq;
            for (u32 k = sh->n_seq; k-- > 0;) {
                const u32 j = seg_list[k];
                if (seg_role[j] & SR_NL) {
                    const u32 s = seg_start(j) + ((seg_role[j] & SR_SKIP1) ? 1 : 0);
                    g.qrec = sh->t_seq - (seg_off[j] + ((u32)seg_end[j] - s));
                    g.last |= F2_L;
                    break;
                }
            }
        }
        for (u32 k = sh->n_seq; k-- > 0;) {                                               // last base of the tile
            const u32 j = seg_list[k];
            const u32 s = seg_start(j), e = (seg_role[j] & SR_NL) ? (u32)seg_end[j] : sh->live_hi;
            if (e > s) { g.last |= F2_B | text[e - 1]; break; }
        }
        return g;
    }

    // staging layout, once the global offsets are known: every staged region congruent (mod 16; bases mod 32) to its destination
    HD void layout(const FusedCfg &C) const
    {
        const F2 &P = sh->pre;
        const u32 a_ids = (u32)((uintptr_t)(C.ids + P.ids) & 15), a_comm = (u32)((uintptr_t)(C.comm + P.comm) & 15);
        sh->s_ids = a_ids;
        sh->s_comm = ((a_ids + sh->t_ids + 15) & ~15u) + a_comm;
        sh->s_seq = ((sh->s_comm + sh->t_comm + 31) & ~31u) + (u32)(P.seq & 31);
    }

    // phase 5: the copies of list entry k (a segment), by the lanes of one group
    HD u32 copy_segment(const FusedCfg &C, u32 k, u32 lane) const
    {
        const u32 j = seg_list[k], r = seg_role[j], role = r & SR_ROLE;
        const bool has_nl = r & SR_NL;
        const u32 s = seg_start(j) + ((r & SR_SKIP1) ? 1 : 0), e = has_nl ? (u32)seg_end[j] : sh->live_hi;
        const F2 &P = sh->pre;
        u32 bad = 0;
        if (role == FR_SEQ) {
            if (C.seq_mode == FS_PACK4) bad = group_copy(text, s, e - s, stage + sh->s_seq + seg_off[j], lane, FC_NONE, false);
            else bad = group_copy(text, s, e - s, C.seq + P.seq + seg_off[j], lane, FC_PROTEIN + (C.seq_mode - FS_PROTEIN), C.upper != 0) ? FU_SEQ : 0;
        } else if (role == FR_QUAL) {
            bad = group_copy(text, s, e - s, C.qual + P.qual + seg_off[j], lane, FC_QUAL, false) ? FU_QUAL : 0;
        } else {                                                             // header: name -> ids, the rest -> comments, terminators
            const u32 sp = seg_sp[j];
            const bool seen = j == 0 && !(r & SR_LS) && sh->entry_sp;
            u8 *di = stage + sh->s_ids + seg_off[j], *dc = stage + sh->s_comm + seg_offb[j];
            u32 nlen = 0, cs = s;
            if (!seen) {
                nlen = (sp != 0xFFFF ? sp : e) - s;
                bad |= group_copy(text, s, nlen, di, lane, C.id_check, false);
                if (lane == 7 && (sp != 0xFFFF || has_nl)) di[nlen] = 0;
                cs = sp != 0xFFFF ? sp + 1 : e;
            }
            bad |= group_copy(text, cs, e - cs, dc, lane, FC_COMM, false);
            if (lane == 6 && has_nl) dc[e - cs] = 0;
            bad = bad ? FU_BADBYTE : 0;
        }
        return bad;
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
    // FASTA: sequence line (list entry k < n_seq) ends -> its length (process.c:389-393)
    HD u64 line_length(u32 k) const
    {
        const u32 j = seg_list[k], r = seg_role[j];
        if (!(r & SR_NL)) return 0;
        const u32 s = seg_start(j);
        u64 L = (u64)seg_end[j] - s;
        // the first line end of the tile closes whatever the earlier tiles left open (every sequence segment but the
        // tile's last one ends a line, so that is list entry 0)
        if (k == 0) L = sh->pre.qrec + seg_off[j] + L;
        return L;
    }

    // phase 6: piece q of the staged bases (32 bases at a multiple of 32 in the file's base numbering) -> 16 bytes of
    // codes + one word of case bits.  Returns FU_SEQ if a base is not an expected code.
    template <class AtomicOr>
    HD u32 pack_piece(const FusedCfg &C, u32 q, AtomicOr atomic_or) const
    {
        const F2 &P = sh->pre;
        const u64 S = P.seq, A = S & 31, g0 = S - A + 32ull * q;         // first base of the piece
        const u32 lo = q == 0 ? (u32)A : 0u;
        const u64 endb = S + sh->t_seq;
        const u32 hi = g0 + 32 <= endb ? 32u : (u32)(endb - g0);
        const u32 *w = (const u32 *)(stage + (sh->s_seq - (u32)A) + 32 * q);
        u32 out[4] = {0, 0, 0, 0}, cbits = 0, inv = 0;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
        for (int k = 0; k < 8; k++) {
            const u32 x = w[k];
            const u32 c0 = C.lut[x & 0xFF], c1 = C.lut[(x >> 8) & 0xFF], c2 = C.lut[(x >> 16) & 0xFF], c3 = C.lut[x >> 24];
            const u32 codes = (c0 & 15) | ((c1 & 15) << 4) | ((c2 & 15) << 8) | ((c3 & 15) << 12);
            const u32 invs = ((c0 >> 7) & 1) | ((c1 >> 6) & 2) | ((c2 >> 5) & 4) | ((c3 >> 4) & 8);
            const u32 ge = ((x & 0xFF) >= 96) | ((((x >> 8) & 0xFF) >= 96) << 1) | ((((x >> 16) & 0xFF) >= 96) << 2) | (((x >> 24) >= 96) << 3);
            out[k >> 1] |= codes << (16 * (k & 1));
            cbits |= ge << (4 * k); inv |= invs << (4 * k);
        }
        const u32 live = (hi >= 32 ? ~0u : ~(~0u << hi)) & (~0u << lo);
        cbits &= live; inv &= live;
        u8 *dst = C.seq + (g0 >> 1);
        if (lo == 0 && hi == 32) {
            *(uint4 *)dst = make_uint4(out[0], out[1], out[2], out[3]);
            if (C.want_mask) C.casebits[g0 >> 5] = cbits;
        } else {
            // a byte shared with the tile before me is mine (I know its low nibble from look-back #2); a dangling last
            // low nibble is my successor's (or the finishing step's, at the end of the input)
            for (u32 b = lo >> 1; 2 * b < hi; b++) {
                u32 v = (out[b >> 2] >> (8 * (b & 3))) & 0xFF;
                if (2 * b < lo) v = (v & 0xF0) | (C.lut[P.last & 0xFF] & 15);
                if (2 * b + 1 >= hi) continue;
                dst[b] = (u8)v;
            }
            if (C.want_mask && cbits) atomic_or(C.casebits + (g0 >> 5), cbits);
        }
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
