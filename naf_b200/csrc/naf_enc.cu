// naf_enc.cu — FASTA/FASTQ text -> the six NAF streams -> .naf, on the GPU.
//
// Replaces ennaf's encode path:
//   confirm_input_format                                   ennaf/src/process.c:547
//   process_non_well_formed_fasta / _fastq (+ well-formed) ennaf/src/process.c:358,477,314,430
//   seq_writer_* , extract_mask/add_mask, encode_dna, add_length   process.c:24-58, encoders.c:30-151
//   report_unexpected_input_char_stats (the counters)      process.c:75-96
//   tail flush + container header/sections                 ennaf/src/ennaf.c:511-589
//
// The reference walks the text one byte at a time through a 16 KB fread buffer.  Here the parser is
// restated as a byte-level finite-state machine (4 states for FASTA, 11 for FASTQ; tests check it against a
// sequential CPU restatement of process.c) and run data-parallel:
//   pass 1  every thread folds its 64 bytes into a state->state map; maps compose associatively, so a
//           scan over tiles yields the parser state entering every tile              (k_fsm_reduce, k_fsm_scan)
//   pass 2  with the entry state known, count what each tile emits per stream        (k_fsm_emit<COUNT>)
//   scans   exclusive sums give every tile its offset in ids / comments / sequence / quality / records
//   pass 3  re-walk and scatter: bytes to their streams, record ends to the lengths  (k_fsm_emit<SCATTER>)
//   pack    bases -> 4-bit codes + one case bit per base (k_pack4); case bits -> mask run-length units
#include "common.cuh"
#include "container.hpp"

namespace nafg {

// ------------------------------------------------------------------ byte classes and the two machines
enum : u32 { F_EOL = 1, F_SPACE = 2, F_ID_OK = 4, F_COMM_OK = 8, F_SEQ_OK = 16, F_QUAL_OK = 32, F_START = 64, F_PLUS = 128 };

enum { FA_NAME = 0, FA_COMMENT = 1, FA_SEQ_LS = 2, FA_SEQ_MID = 3, FA_NSTATES = 4 };
enum { FQ_NAME = 0, FQ_COMMENT = 1, FQ_SEQ = 2, FQ_AFTER_SEQ = 3, FQ_PLUS = 4, FQ_BEFORE_QUAL = 5, FQ_QUAL = 6, FQ_AFTER_QUAL = 7,
       FQ_ERR_NOPLUS = 8, FQ_ERR_NOAT = 9, FQ_ERR_NOTWF = 10, FQ_NSTATES = 11 };

struct FsmTables {
    u8  cls[256];          // byte -> F_* flags
    u64 trans[256];        // flags -> packed transition map (4 bits per source state)
    u8  idem[256];         // map is idempotent (applying it twice == once)
};

__host__ __device__ inline u32 fa_next(u32 s, u32 f)
{
    switch (s) {
    case FA_NAME:    return (f & F_ID_OK) ? FA_NAME : ((f & F_SPACE) ? ((f & F_EOL) ? FA_SEQ_LS : FA_COMMENT) : FA_NAME);
    case FA_COMMENT: return (f & F_EOL) ? FA_SEQ_LS : FA_COMMENT;        // COMM_OK bytes are never EOL
    case FA_SEQ_LS:  return (f & F_START) ? FA_NAME : ((f & F_EOL) ? FA_SEQ_LS : FA_SEQ_MID);
    default:         return (f & F_EOL) ? FA_SEQ_LS : FA_SEQ_MID;
    }
}
__host__ __device__ inline u32 fq_next(u32 s, u32 f, bool wf)
{
    switch (s) {
    case FQ_NAME:        return (f & F_ID_OK) ? FQ_NAME : ((f & F_SPACE) ? ((f & F_EOL) ? FQ_SEQ : FQ_COMMENT) : FQ_NAME);
    case FQ_COMMENT:     return (f & F_EOL) ? FQ_SEQ : FQ_COMMENT;
    case FQ_SEQ:         return (f & F_EOL) ? FQ_AFTER_SEQ : FQ_SEQ;
    case FQ_AFTER_SEQ:   if (wf) return (f & F_PLUS) ? FQ_PLUS : FQ_ERR_NOTWF;
                         return (f & F_EOL) ? FQ_AFTER_SEQ : ((f & F_PLUS) ? FQ_PLUS : FQ_ERR_NOPLUS);
    case FQ_PLUS:        if (wf) return (f & F_EOL) ? FQ_BEFORE_QUAL : FQ_ERR_NOTWF;
                         return (f & F_EOL) ? FQ_BEFORE_QUAL : FQ_PLUS;
    case FQ_BEFORE_QUAL: if (wf) return (f & F_EOL) ? FQ_AFTER_QUAL : FQ_QUAL;
                         return (f & F_EOL) ? FQ_BEFORE_QUAL : FQ_QUAL;
    case FQ_QUAL:        return (f & F_EOL) ? FQ_AFTER_QUAL : FQ_QUAL;
    case FQ_AFTER_QUAL:  if (wf) return (f & F_START) ? FQ_NAME : FQ_ERR_NOTWF;
                         return (f & F_EOL) ? FQ_AFTER_QUAL : ((f & F_START) ? FQ_NAME : FQ_ERR_NOAT);
    default:             return s;
    }
}

struct ParseCfg {
    int fastq, wf, seq_type, text_fasta, no_mask, strict, nstates;
    u8 repl;
};

static void build_tables(const ParseCfg &c, FsmTables &t)
{
    auto is_eol = [](int ch) { return ch >= 0x0A && ch <= 0x0D; };
    auto is_space = [&](int ch) { return (ch >= 0x09 && ch <= 0x0D) || ch == 0x20; };
    auto in_set = [](int ch, const char *set) { if (ch >= 'a' && ch <= 'z') ch -= 32; return ch > 0 && strchr(set, ch) != nullptr; };
    for (int ch = 0; ch < 256; ch++) {
        u32 f = 0;
        if (c.wf) {                                                      // tables.c:61 is_well_formed_space
            if (ch == '\n') f |= F_EOL | F_SPACE;
            if (ch == ' ') f |= F_SPACE;
            if (!(f & F_SPACE)) f |= F_ID_OK;
            if (!(f & F_EOL)) f |= F_COMM_OK | F_SEQ_OK | F_QUAL_OK;
        } else {
            if (is_eol(ch)) f |= F_EOL;
            if (is_space(ch)) f |= F_SPACE;
            if (!(ch <= 32 || ch == 127 || ch == 255)) f |= F_ID_OK;     // tables.c:115
            if (!(ch < 32 || ch == 127 || ch == 255)) f |= F_COMM_OK;    // tables.c:126
            if (ch >= 33 && ch <= 126) f |= F_QUAL_OK;                   // tables.c:137
            bool ok;
            switch (c.seq_type) {
            case NAFGPU_DNA:     ok = in_set(ch, "-ABCDGHKMNRSTVWY"); break;           // tables.c:72
            case NAFGPU_RNA:     ok = in_set(ch, "-ABCDGHKMNRSUVWY"); break;           // tables.c:82
            case NAFGPU_PROTEIN: ok = in_set(ch, "*-ABCDEFGHIJKLMNOPQRSTUVWXYZ"); break; // tables.c:104
            default:             ok = !(ch <= 32 || ch == 127 || ch == 255); break;
            }
            if (ok) f |= F_SEQ_OK;
            if (c.text_fasta && ch == '>') f &= ~(F_SEQ_OK | F_ID_OK);   // ennaf.c:466 flips the shared table entry
        }
        if (ch == (c.fastq ? '@' : '>')) f |= F_START;
        if (ch == '+') f |= F_PLUS;
        t.cls[ch] = (u8)f;
    }
    for (int f = 0; f < 256; f++) {
        u64 m = 0;
        for (int s = 0; s < c.nstates; s++) m |= (u64)(c.fastq ? fq_next(s, f, c.wf) : fa_next(s, f)) << (4 * s);
        t.trans[f] = m;
        u64 mm = 0;
        for (int s = 0; s < c.nstates; s++) mm |= ((m >> (4 * ((m >> (4 * s)) & 15))) & 15) << (4 * s);
        t.idem[f] = mm == m;
    }
}

__device__ __forceinline__ u64 map_compose(u64 f, u64 g, int ns)     // first f, then g
{
    u64 h = 0;
    for (int s = 0; s < ns; s++) h |= ((g >> (4 * ((f >> (4 * s)) & 15))) & 15) << (4 * s);
    return h;
}
__device__ __forceinline__ u64 map_identity(int ns) { u64 m = 0; for (int s = 0; s < ns; s++) m |= (u64)s << (4 * s); return m; }

// ------------------------------------------------------------------ pass 1: per-tile state maps
static const int PT = 256, PB = 64, PTILE = PT * PB;       // threads, bytes per thread, bytes per tile (16 KB)

struct ParseArgs {
    const u8 *text; u64 n, p0;                // p0: first byte after the leading '>' / '@'
    ParseCfg cfg;
    const FsmTables *tab;
    u64 *tile_map; u8 *tile_state;            // per tile: map, entry state
    u64 ntiles;
};

__device__ __forceinline__ u64 fold_bytes(const ParseArgs &A, const FsmTables *T, u64 lo, u64 hi, int ns)
{
    u64 f = map_identity(ns);
    u32 prev = 0xFFFFFFFFu;
    for (u64 p = lo; p < hi; p++) {
        u32 fl = T->cls[A.text[p]];
        if (fl == prev && T->idem[fl]) continue;     // same class as the previous byte and idempotent: nothing new
        f = map_compose(f, T->trans[fl], ns);
        prev = fl;
    }
    return f;
}

__global__ void __launch_bounds__(PT) k_fsm_reduce(const ParseArgs A)
{
    __shared__ FsmTables T;
    __shared__ u64 wmap[PT / 32];
    for (int i = threadIdx.x; i < (int)sizeof(FsmTables) / 4; i += PT) ((u32 *)&T)[i] = ((const u32 *)A.tab)[i];
    __syncthreads();
    const int ns = A.cfg.nstates;
    u64 lo = A.p0 + (u64)blockIdx.x * PTILE + (u64)threadIdx.x * PB, hi = lo + PB;
    if (lo > A.n) lo = A.n; if (hi > A.n) hi = A.n;
    u64 f = fold_bytes(A, &T, lo, hi, ns);
    // ordered reduction: lane i absorbs lane i+d
    for (int d = 1; d < 32; d <<= 1) {
        u64 g = __shfl_down_sync(0xFFFFFFFFu, f, d);
        if ((threadIdx.x & 31) + d < 32 && ((threadIdx.x & 31) % (2 * d)) == 0) f = map_compose(f, g, ns);
    }
    if ((threadIdx.x & 31) == 0) wmap[threadIdx.x >> 5] = f;
    __syncthreads();
    if (threadIdx.x == 0) {
        u64 m = wmap[0];
        for (int w = 1; w < PT / 32; w++) m = map_compose(m, wmap[w], ns);
        A.tile_map[blockIdx.x] = m;
    }
}

// one CTA: chunked scan of the tile maps -> entry state of every tile (initial state = NAME)
__global__ void __launch_bounds__(1024) k_fsm_scan(const ParseArgs A)
{
    __shared__ u64 cmap[1024];
    const int ns = A.cfg.nstates;
    u64 per = (A.ntiles + 1023) / 1024;
    u64 lo = (u64)threadIdx.x * per, hi = lo + per;
    if (lo > A.ntiles) lo = A.ntiles; if (hi > A.ntiles) hi = A.ntiles;
    u64 f = map_identity(ns);
    for (u64 t = lo; t < hi; t++) f = map_compose(f, A.tile_map[t], ns);
    cmap[threadIdx.x] = f;
    __syncthreads();
    if (threadIdx.x == 0) {
        u32 s = 0;                                      // NAME for both machines
        for (int c = 0; c < 1024; c++) { u64 m = cmap[c]; cmap[c] = s; s = (u32)((m >> (4 * s)) & 15); }
    }
    __syncthreads();
    u32 s = (u32)cmap[threadIdx.x];
    for (u64 t = lo; t < hi; t++) { A.tile_state[t] = (u8)s; s = (u32)((A.tile_map[t] >> (4 * s)) & 15); }
    if (hi == A.ntiles && lo < hi) A.tile_state[A.ntiles] = (u8)s;      // state at end of input
    if (A.ntiles == 0 && threadIdx.x == 0) A.tile_state[0] = 0;
}

// ------------------------------------------------------------------ pass 2 / 3: emit
struct TileCounts { u64 ids, comm, seq, seq_counted, qual, rec, line_last; u32 has_line; u32 pad; };

struct EmitArgs {
    ParseArgs P;
    TileCounts *tile;                        // COUNT: per-tile totals
    // SCATTER inputs: exclusive prefixes per tile
    const u64 *pre_ids, *pre_comm, *pre_seq, *pre_cnt, *pre_qual, *pre_rec, *pre_line;   // pre_line: global counted-seq value at the last line end before the tile
    u8 *ids, *comm, *bases, *qual;           // outputs
    u64 *rec_seq_end, *rec_qual_end, *rec_pos;   // per record: counted-seq / qual totals at its end, text offset of its end
    unsigned long long *unexpected;          // [4][257]
    unsigned long long *longest;             // atomicMax target (FASTA)
    unsigned long long *first_bad;           // strict / FSM error: min over (pos << 8 | kind)
};

enum { BAD_ID = 1, BAD_COMMENT = 2, BAD_SEQ = 3, BAD_QUAL = 4, BAD_NOPLUS = 5, BAD_NOAT = 6, BAD_NOTWF = 7 };

struct Emit { u32 ids, comm, seq, cnt, qual, rec; };

// Walk bytes [lo, hi) from state s.  COUNT mode tallies; SCATTER mode writes at the running offsets in `o`.
template <bool SCATTER>
__device__ __forceinline__ u32 walk(const EmitArgs &E, const FsmTables *T, u64 lo, u64 hi, u32 s, Emit &n, u64 o_ids, u64 o_comm,
                                    u64 o_seq, u64 o_cnt, u64 o_qual, u64 o_rec, u64 &line_base, bool &line_base_valid,
                                    u64 &line_max, u64 &line_last, bool &has_line)
{
    const ParseCfg &C = E.P.cfg;
    const u8 *text = E.P.text;
    for (u64 p = lo; p < hi; p++) {
        const u32 c = text[p], f = T->cls[c];
        u32 ns;
        int to_ids = -1, to_comm = -1, to_seq = -1, to_qual = -1; bool counted = true, rec_end = false, line_end = false; int bad = 0, badk = -1;
        if (!C.fastq) {
            ns = fa_next(s, f);
            switch (s) {
            case FA_NAME:
                if (f & F_ID_OK) to_ids = c;
                else if (f & F_SPACE) { to_ids = 0; if (f & F_EOL) to_comm = 0; }
                else { bad = BAD_ID; badk = 0; to_seq = '?'; counted = false; }          // process.c:366 (reference quirk, restated)
                break;
            case FA_COMMENT:
                if (f & F_COMM_OK) to_comm = c;
                else if (f & F_EOL) to_comm = 0;
                else { bad = BAD_COMMENT; badk = 1; to_comm = '?'; }
                break;
            default:
                if (s == FA_SEQ_LS && (f & F_START)) { rec_end = true; break; }
                if (f & F_SEQ_OK) to_seq = c;
                else if (f & F_EOL) line_end = true;
                else if (f & F_SPACE) {}
                else if (C.text_fasta && c == '>') to_seq = c;                           // process.c:413
                else { bad = BAD_SEQ; badk = 2; to_seq = C.repl; }
                break;
            }
        } else {
            ns = fq_next(s, f, C.wf);
            switch (s) {
            case FQ_NAME:
                if (f & F_ID_OK) to_ids = c;
                else if (f & F_SPACE) { to_ids = 0; if (f & F_EOL) to_comm = 0; }
                else { bad = BAD_ID; badk = 0; to_seq = '?'; counted = false; }          // process.c:485
                break;
            case FQ_COMMENT:
                if (f & F_COMM_OK) to_comm = c;
                else if (f & F_EOL) to_comm = 0;
                else { bad = BAD_COMMENT; badk = 1; to_comm = '?'; }
                break;
            case FQ_SEQ:
                if (f & F_SEQ_OK) to_seq = c;
                else if (f & F_EOL) {}
                else if (f & F_SPACE) {}
                else { bad = BAD_SEQ; badk = 2; to_seq = C.repl; }
                break;
            case FQ_BEFORE_QUAL:
                if (!(f & F_EOL)) to_qual = c;                                            // process.c:523: unvalidated
                else if (C.wf) rec_end = true;                                            // empty quality line
                break;
            case FQ_QUAL:
                if (f & F_QUAL_OK) to_qual = c;
                else if (f & F_EOL) rec_end = true;
                else if (f & F_SPACE) {}
                else { bad = BAD_QUAL; badk = 3; to_qual = '!'; }
                break;
            default: break;
            }
            if (ns >= FQ_ERR_NOPLUS && s < FQ_ERR_NOPLUS) { bad = ns == FQ_ERR_NOPLUS ? BAD_NOPLUS : (ns == FQ_ERR_NOAT ? BAD_NOAT : BAD_NOTWF); badk = -2; }
        }
        if (SCATTER) {
            if (to_ids >= 0) E.ids[o_ids + n.ids] = (u8)to_ids;
            if (to_comm >= 0) E.comm[o_comm + n.comm] = (u8)to_comm;
            if (to_seq >= 0) {
                u8 b = (u8)to_seq;
                if (C.seq_type >= NAFGPU_PROTEIN && C.no_mask && b >= 'a' && b <= 'z') b -= 32;   // process.c:49
                E.bases[o_seq + n.seq] = b;
            }
            if (to_qual >= 0) E.qual[o_qual + n.qual] = (u8)to_qual;
            if (bad) {
                if (badk >= 0) atomicAdd(&E.unexpected[badk * 257 + c], 1ull);
                if (badk == -2 || C.strict) atomicMin(E.first_bad, (unsigned long long)((p << 8) | (u32)bad));
            }
        }
        n.ids += to_ids >= 0; n.comm += to_comm >= 0; n.qual += to_qual >= 0;
        if (to_seq >= 0) { n.seq++; if (counted) n.cnt++; }
        if (line_end) {
            // sequence bytes since the previous line end (FASTA only; FASTQ takes max read length)
            u64 v = o_cnt + n.cnt;
            if (line_base_valid) { u64 d = v - line_base; if (d > line_max) line_max = d; }
            line_base = v; line_base_valid = true; line_last = v; has_line = true;
        }
        if (rec_end) {
            if (SCATTER) {
                u64 r = o_rec + n.rec;
                E.rec_seq_end[r] = o_cnt + n.cnt;
                if (C.fastq) E.rec_qual_end[r] = o_qual + n.qual;
                E.rec_pos[r] = p;
            }
            n.rec++;
        }
        s = ns;
    }
    return s;
}

template <bool SCATTER>
__global__ void __launch_bounds__(PT) k_fsm_emit(const EmitArgs E)
{
    __shared__ FsmTables T;
    __shared__ u64 sm[33];
    __shared__ u64 wmap[PT / 32];
    __shared__ u32 wstate[PT / 32];
    const ParseArgs &A = E.P;
    for (int i = threadIdx.x; i < (int)sizeof(FsmTables) / 4; i += PT) ((u32 *)&T)[i] = ((const u32 *)A.tab)[i];
    __syncthreads();
    const int ns = A.cfg.nstates;
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u64 lo = A.p0 + (u64)blockIdx.x * PTILE + (u64)threadIdx.x * PB, hi = lo + PB;
    if (lo > A.n) lo = A.n; if (hi > A.n) hi = A.n;
    // entry state of this thread: tile entry state pushed through the maps of the threads before me
    u64 f = fold_bytes(A, &T, lo, hi, ns);
    u64 incl = f;
    for (int d = 1; d < 32; d <<= 1) { u64 g = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= (unsigned)d) incl = map_compose(g, incl, ns); }
    if (lane == 31) wmap[warp] = incl;
    __syncthreads();
    if (threadIdx.x == 0) {
        u32 s = A.tile_state[blockIdx.x];
        for (int w = 0; w < PT / 32; w++) { wstate[w] = s; s = (u32)((wmap[w] >> (4 * s)) & 15); }
    }
    __syncthreads();
    u64 excl = __shfl_up_sync(0xFFFFFFFFu, incl, 1);
    u32 s0 = wstate[warp];
    if (lane > 0) s0 = (u32)((excl >> (4 * s0)) & 15);

    Emit n = {0, 0, 0, 0, 0, 0};
    u64 line_base = 0, line_max = 0, line_last = 0; bool lbv = false, has_line = false;
    if (!SCATTER) {
        walk<false>(E, &T, lo, hi, s0, n, 0, 0, 0, 0, 0, 0, line_base, lbv, line_max, line_last, has_line);
        // tile totals; line_last of the tile = counted-seq offset (tile-relative) at the last line end in the tile
        u64 tot, pre;
        TileCounts tc;
        pre = block_excl_scan(n.ids, &tot, sm); tc.ids = tot;
        pre = block_excl_scan(n.comm, &tot, sm); tc.comm = tot;
        pre = block_excl_scan(n.seq, &tot, sm); tc.seq = tot;
        u64 pre_cnt = block_excl_scan(n.cnt, &tot, sm); tc.seq_counted = tot;
        pre = block_excl_scan(n.qual, &tot, sm); tc.qual = tot;
        pre = block_excl_scan(n.rec, &tot, sm); tc.rec = tot;
        (void)pre;
        // last line end in the tile: max over threads of (pre_cnt + local line_last) among threads that saw one
        u64 v = has_line ? pre_cnt + line_last + 1 : 0;       // +1 so that 0 means "none"
        for (int d = 16; d; d >>= 1) { u64 o = __shfl_xor_sync(0xFFFFFFFFu, v, d); if (o > v) v = o; }
        if (lane == 0) sm[warp] = v;
        __syncthreads();
        if (threadIdx.x == 0) {
            u64 m = 0; for (int w = 0; w < PT / 32; w++) if (sm[w] > m) m = sm[w];
            tc.has_line = m != 0; tc.line_last = m ? m - 1 : 0; tc.pad = 0;
            E.tile[blockIdx.x] = tc;
        }
        return;
    }
    // SCATTER: first count (cheap re-walk) to get this thread's offsets, then write
    walk<false>(E, &T, lo, hi, s0, n, 0, 0, 0, 0, 0, 0, line_base, lbv, line_max, line_last, has_line);
    u64 tot;
    u64 o_ids = block_excl_scan(n.ids, &tot, sm) + E.pre_ids[blockIdx.x];
    u64 o_comm = block_excl_scan(n.comm, &tot, sm) + E.pre_comm[blockIdx.x];
    u64 o_seq = block_excl_scan(n.seq, &tot, sm) + E.pre_seq[blockIdx.x];
    u64 o_cnt = block_excl_scan(n.cnt, &tot, sm) + E.pre_cnt[blockIdx.x];
    u64 o_qual = block_excl_scan(n.qual, &tot, sm) + E.pre_qual[blockIdx.x];
    u64 o_rec = block_excl_scan(n.rec, &tot, sm) + E.pre_rec[blockIdx.x];
    // counted-seq value at the last line end before this thread (exclusive max-scan; values are monotone)
    u64 mine = has_line ? o_cnt + line_last + 1 : 0;
    u64 run = mine;
    for (int d = 1; d < 32; d <<= 1) { u64 g = __shfl_up_sync(0xFFFFFFFFu, run, d); if (lane >= (unsigned)d && g > run) run = g; }
    if (lane == 31) sm[warp] = run;
    __syncthreads();
    u64 before = E.pre_line[blockIdx.x] + 1;                  // tile carry (+1 encoding; >= 1 because line base 0 = start of data)
    for (unsigned w = 0; w < warp; w++) if (sm[w] > before) before = sm[w];
    u64 prev_lane = __shfl_up_sync(0xFFFFFFFFu, run, 1);
    if (lane > 0 && prev_lane > before) before = prev_lane;
    __syncthreads();
    Emit m = {0, 0, 0, 0, 0, 0};
    line_base = before - 1; lbv = true; line_max = 0; line_last = 0; has_line = false;
    walk<true>(E, &T, lo, hi, s0, m, o_ids, o_comm, o_seq, o_cnt, o_qual, o_rec, line_base, lbv, line_max, line_last, has_line);
    if (!A.cfg.fastq) {
        // pending (unterminated) last line of the input: counted bytes after the last line end
        if (hi == A.n && lo < hi) { u64 d = o_cnt + m.cnt - line_base; if (d > line_max) line_max = d; }
        if (line_max) atomicMax(E.longest, (unsigned long long)line_max);
    }
}

// ------------------------------------------------------------------ 4-bit pack + case bits
// encoders.c:30 encode_dna (first base in the low nibble; odd tail has a zero high nibble, ennaf.c:525)
// and the predicate of encoders.c:134 (masked <=> byte >= 96).  One thread: 32 bases -> 16 bytes + 1 word.
__global__ void k_pack4(const u8 *bases, u64 n, u8 *packed, u32 *casebits, int want_mask, const u8 *nuc_code)
{
    __shared__ u8 c_nuc_code[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) c_nuc_code[i] = nuc_code[i];
    __syncthreads();
    u64 g = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u64 b0 = g * 32;
    if (b0 >= n) return;
    u32 out[4] = {0, 0, 0, 0}, mbits = 0;
    if (b0 + 32 <= n) {
        const uint4 *src = (const uint4 *)(bases + b0);
        uint4 v0 = src[0], v1 = src[1];
        u32 w[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
        for (int k = 0; k < 8; k++) {
            u32 x = w[k];
            u32 c0 = c_nuc_code[x & 0xFF], c1 = c_nuc_code[(x >> 8) & 0xFF], c2 = c_nuc_code[(x >> 16) & 0xFF], c3 = c_nuc_code[x >> 24];
            u32 two = c0 | (c1 << 4) | (c2 << 8) | (c3 << 12);
            out[k >> 1] |= two << (16 * (k & 1));
            // case bit: byte >= 96
            u32 ge = ((x & 0xFF) >= 96) | ((((x >> 8) & 0xFF) >= 96) << 1) | ((((x >> 16) & 0xFF) >= 96) << 2) | (((x >> 24) >= 96) << 3);
            mbits |= ge << (4 * k);
        }
        *(uint4 *)(packed + (b0 >> 1)) = make_uint4(out[0], out[1], out[2], out[3]);
    } else {
        for (u64 i = b0; i < n; i += 2) {
            u32 c0 = c_nuc_code[bases[i]], c1 = i + 1 < n ? c_nuc_code[bases[i + 1]] : 0;
            packed[i >> 1] = (u8)(c0 | (c1 << 4));
        }
        for (u64 i = b0; i < n; i++) if (bases[i] >= 96) mbits |= 1u << (i - b0);
    }
    if (want_mask) casebits[g] = mbits;
}

// flips[w] = positions where the case differs from the previous base (case before base 0 = unmasked)
__global__ void k_flip_count(const u32 *casebits, u64 nwords, u64 *tile_counts)
{
    __shared__ u64 sm[33];
    u64 w = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u32 c = 0;
    if (w < nwords) { u32 cur = casebits[w], prev = w ? casebits[w - 1] >> 31 : 0; c = __popc(cur ^ ((cur << 1) | prev)); }
    u64 tot; block_excl_scan(c, &tot, sm);
    if (threadIdx.x == 0) tile_counts[blockIdx.x] = tot;
}
__global__ void k_flip_scatter(const u32 *casebits, u64 nwords, const u64 *tile_prefix, u64 *flip_pos)
{
    __shared__ u64 sm[33];
    u64 w = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u32 fl = 0;
    if (w < nwords) { u32 cur = casebits[w], prev = w ? casebits[w - 1] >> 31 : 0; fl = cur ^ ((cur << 1) | prev); }
    u64 tot; u64 r = block_excl_scan(__popc(fl), &tot, sm) + tile_prefix[blockIdx.x];
    while (fl) { int k = __ffs(fl) - 1; fl &= fl - 1; flip_pos[r++] = (w << 5) + k; }
}

// runs of equal case: run k = [flip[k-1], flip[k]) (run 0 starts at 0, run R ends at n).  casebits beyond n are
// zero, so a masked tail produces one spurious flip at n; clamping to n makes that final run empty.
struct RunCalc {
    const u64 *fp; u64 R, nb;
    __device__ u64 len(size_t k) const
    {
        u64 s = k ? fp[k - 1] : 0, e = k < R ? fp[k] : nb;
        if (s > nb) s = nb;
        if (e > nb) e = nb;
        return e - s;
    }
    __device__ u64 units(size_t k) const                      // encoders.c:98 add_mask: L/255 bytes of 255, then L%255
    {
        u64 L = len(k);
        if (k == R && L == 0) return 0;                         // final run only if > 0 (ennaf.c:511)
        return L / 255 + 1;
    }
};

// ------------------------------------------------------------------ split orchestration

struct SplitDev {
    u8 *ids, *comm, *len, *mask, *seq, *qual;
    u64 n_ids, n_comm, n_len, n_mask, n_seq, n_qual;     // bytes
    u64 n_bases, n_records, longest;
    int format, store_mask, store_qual;
};

static void die_input(const std::string &m) { fail(NAFGPU_E_INPUT, m); }

static SplitDev split_streams(Ctx &ctx, CudaExec &ex, const u8 *d_text, size_t n, const nafgpu_enc_opts &o, nafgpu_enc_info *info)
{
    SplitDev S; memset(&S, 0, sizeof S);
    if (info) memset(info, 0, sizeof *info);
    ParseCfg C; memset(&C, 0, sizeof C);
    C.seq_type = o.seq_type; C.no_mask = o.no_mask; C.strict = o.strict; C.wf = o.well_formed;
    C.repl = o.seq_type == NAFGPU_PROTEIN ? 'X' : (o.seq_type == NAFGPU_TEXT ? '?' : 'N');     // ennaf.c:447-470
    S.store_mask = !(o.no_mask || o.seq_type >= NAFGPU_PROTEIN);                                 // ennaf.c:445

    // ---- confirm_input_format (process.c:547): skip leading white space, first byte decides
    u64 p0 = 0; int fmt = 0;
    {
        // leading white space is short in practice: look at the first 64 KB on the host
        size_t head = n < 65536 ? n : 65536;
        ctx.host_scratch.resize(head + 1);
        if (head) ex.download(ctx.host_scratch.data(), d_text, head);
        const u8 *h = ctx.host_scratch.data();
        auto is_space = [](int ch) { return (ch >= 0x09 && ch <= 0x0D) || ch == 0x20; };
        u32 last = '\n'; size_t i = 0;
        while (i < head && is_space(h[i])) { last = h[i]; i++; }
        if (i == head && head < n) fail(NAFGPU_E_UNSUPPORTED, "more than 64 KB of leading white space\n");
        if (i < head) {
            u32 c = h[i];
            bool at_line_start = last >= 0x0A && last <= 0x0D;
            if (c == '>' && at_line_start) fmt = NAFGPU_FMT_FASTA;
            else if (c == '@' && at_line_start) fmt = NAFGPU_FMT_FASTQ;
            else if (c == '>' || c == '@') die_input(std::string("invalid input - first '") + (char)c + "' is not at the beginning of the line\n");
            else die_input("input data is in unknown format - first non-space character is neither '>' nor '@'\n");
            p0 = i + 1;
        }
        if (o.input_format != NAFGPU_FMT_AUTO && fmt && o.input_format != fmt) die_input("input format is different from format specified in the command line\n");
    }
    S.format = fmt; S.store_qual = fmt == NAFGPU_FMT_FASTQ;
    if (info) info->format = fmt;
    if (fmt == 0) return S;                                   // empty input: zero sequences, empty streams (process.c:589)
    C.fastq = fmt == NAFGPU_FMT_FASTQ;
    C.nstates = C.fastq ? FQ_NSTATES : FA_NSTATES;
    C.text_fasta = o.seq_type == NAFGPU_TEXT && !C.fastq;

    FsmTables ht; build_tables(C, ht);
    FsmTables *d_tab = ex.alloc<FsmTables>(1);
    ex.upload(d_tab, &ht, sizeof ht);

    const u64 body = n - p0, ntiles = (body + PTILE - 1) / PTILE;
    ParseArgs P; P.text = d_text; P.n = n; P.p0 = p0; P.cfg = C; P.tab = d_tab; P.ntiles = ntiles;
    P.tile_map = ex.alloc<u64>(ntiles + 1); P.tile_state = ex.alloc<u8>(ntiles + 2);
    if (ntiles) { KLAUNCH(ex, "k_fsm_reduce", k_fsm_reduce<<<(unsigned)ntiles, PT, 0, ex.stream>>>(P)); }
    KLAUNCH(ex, "k_fsm_scan", k_fsm_scan<<<1, 1024, 0, ex.stream>>>(P));

    EmitArgs E; memset(&E, 0, sizeof E);
    E.P = P;
    E.tile = ex.alloc<TileCounts>(ntiles + 1);
    if (ntiles) { KLAUNCH(ex, "k_fsm_emit", k_fsm_emit<false><<<(unsigned)ntiles, PT, 0, ex.stream>>>(E)); }
    // exclusive sums of the six counters + exclusive max of the line-end marker
    u64 *pre[7];
    for (int k = 0; k < 7; k++) pre[k] = ex.alloc<u64>(ntiles + 2);
    const TileCounts *tc = E.tile;
    exclusive_scan(ex, [tc] __device__ (size_t i) { return tc[i].ids; }, ntiles, pre[0]);
    exclusive_scan(ex, [tc] __device__ (size_t i) { return tc[i].comm; }, ntiles, pre[1]);
    exclusive_scan(ex, [tc] __device__ (size_t i) { return tc[i].seq; }, ntiles, pre[2]);
    exclusive_scan(ex, [tc] __device__ (size_t i) { return tc[i].seq_counted; }, ntiles, pre[3]);
    exclusive_scan(ex, [tc] __device__ (size_t i) { return tc[i].qual; }, ntiles, pre[4]);
    exclusive_scan(ex, [tc] __device__ (size_t i) { return tc[i].rec; }, ntiles, pre[5]);
    // pre_line[t] = counted-seq value at the last line end in tiles < t (0 if none): sequential max over few 100k tiles
    {
        u64 *pl = pre[6]; const u64 *pc = pre[3]; const u64 nt = ntiles;
        ex.for_each(1, [=] __device__ (size_t) { u64 m = 0; for (u64 t = 0; t < nt; t++) { pl[t] = m; if (tc[t].has_line) { u64 v = pc[t] + tc[t].line_last; if (v > m) m = v; } } pl[nt] = m; });
    }
    u64 tot[6]; u8 end_state;
    for (int k = 0; k < 6; k++) ex.download(&tot[k], pre[k] + ntiles, 8);
    ex.download(&end_state, P.tile_state + ntiles, 1);
    if (ntiles == 0) end_state = 0;

    // ---- what the end of input adds (process.c:417-425, :535-543): pending terminators and the last record
    u64 add_ids = 0, add_comm = 0, add_rec = 0;
    if (!C.fastq) {
        if (end_state == FA_NAME) { add_ids = 1; add_comm = 1; }
        else if (end_state == FA_COMMENT) add_comm = 1;
        add_rec = 1;
    } else {
    
        switch (end_state) {
        case FQ_NAME: case FQ_COMMENT: /* decided below, after earlier errors */ break;
        case FQ_QUAL: add_rec = 1; break;
        case FQ_BEFORE_QUAL: if (C.wf) add_rec = 1; break;       // well-formed: empty last quality line without '\n'
        default: break;
        }

    }
    const u64 n_ids = tot[0] + add_ids, n_comm = tot[1] + add_comm, n_seq = tot[2], n_cnt = tot[3], n_qual = tot[4], n_rec = tot[5] + add_rec;

    S.ids = ex.alloc<u8>(n_ids + 64); S.comm = ex.alloc<u8>(n_comm + 64); S.qual = ex.alloc<u8>(n_qual + 64);
    u8 *bases = ex.alloc<u8>(n_seq + 64);
    u64 *rec_seq_end = ex.alloc<u64>(n_rec + 2), *rec_qual_end = ex.alloc<u64>(n_rec + 2), *rec_pos = ex.alloc<u64>(n_rec + 2);
    unsigned long long *d_unexp = ex.alloc<unsigned long long>(4 * 257 + 4);
    ex.zero(d_unexp, (4 * 257 + 4) * 8);
    unsigned long long *d_longest = d_unexp + 4 * 257, *d_first_bad = d_unexp + 4 * 257 + 1;
    ex.fill(d_first_bad, 0xFF, 8);
    E.pre_ids = pre[0]; E.pre_comm = pre[1]; E.pre_seq = pre[2]; E.pre_cnt = pre[3]; E.pre_qual = pre[4]; E.pre_rec = pre[5]; E.pre_line = pre[6];
    E.ids = S.ids; E.comm = S.comm; E.bases = bases; E.qual = S.qual;
    E.rec_seq_end = rec_seq_end; E.rec_qual_end = rec_qual_end; E.rec_pos = rec_pos;
    E.unexpected = d_unexp; E.longest = d_longest; E.first_bad = d_first_bad;
    if (ntiles) { KLAUNCH(ex, "k_fsm_emit", k_fsm_emit<true><<<(unsigned)ntiles, PT, 0, ex.stream>>>(E)); }
    // end-of-input additions
    {
        u8 *ids = S.ids, *comm = S.comm; const u64 a = tot[0], b = tot[1], r = tot[5], cnt = n_cnt, ql = n_qual, nn = n;
        const u64 ai = add_ids, ac = add_comm, ar = add_rec;
        ex.for_each(1, [=] __device__ (size_t) {
            if (ai) ids[a] = 0;
            if (ac) comm[b] = 0;
            if (ar) { rec_seq_end[r] = cnt; rec_qual_end[r] = ql; rec_pos[r] = nn; }
        });
    }
    // ---- errors, in input order (the reference dies at the first one it meets)
    unsigned long long h_tail[2 + 0]; (void)h_tail;
    std::vector<unsigned long long> h_unexp(4 * 257 + 4);
    ex.download(h_unexp.data(), d_unexp, h_unexp.size() * 8);
    const unsigned long long first_bad = h_unexp[4 * 257 + 1];
    u64 bad_pos = ~0ull; int bad_kind = 0;
    if (first_bad != ~0ull) { bad_pos = first_bad >> 8; bad_kind = (int)(first_bad & 0xFF); }
    // quality-length mismatch: first record whose quality length differs (process.c:531)
    u64 mism_rec = ~0ull, mism_pos = ~0ull, mism_q = 0, mism_s = 0;
    if (C.fastq && n_rec) {
        unsigned long long *d_m = ex.alloc<unsigned long long>(1);
        ex.fill(d_m, 0xFF, 8);
        ex.for_each(n_rec, [=] __device__ (size_t i) {
            u64 sl = rec_seq_end[i] - (i ? rec_seq_end[i - 1] : 0), ql = rec_qual_end[i] - (i ? rec_qual_end[i - 1] : 0);
            if (sl != ql) atomicMin(d_m, (unsigned long long)i);
        });
        unsigned long long m; ex.download(&m, d_m, 8);
        if (m != ~0ull) {
            mism_rec = m;
            u64 e[2], q[2], pos;
            ex.download(e, rec_seq_end + (m ? m - 1 : 0), 16); ex.download(q, rec_qual_end + (m ? m - 1 : 0), 16); ex.download(&pos, rec_pos + m, 8);
            mism_s = m ? e[1] - e[0] : e[0]; mism_q = m ? q[1] - q[0] : q[0]; mism_pos = pos;
        }
    }
    auto seq_no_at = [&](u64 pos) -> u64 {          // records completed before text offset `pos` (for "... of sequence N")
        if (n_rec == 0) return 0;
        std::vector<u64> rp(n_rec);                  // rare path: host binary search over record end offsets
        ex.download(rp.data(), rec_pos, n_rec * 8);
        u64 lo = 0, hi = n_rec; while (lo < hi) { u64 mid = (lo + hi) / 2; if (rp[mid] < pos) lo = mid + 1; else hi = mid; }
        return lo;
    };
    if (bad_pos != ~0ull && bad_pos <= mism_pos) {
        u8 ch; ex.download(&ch, d_text + bad_pos, 1);
        u64 k = seq_no_at(bad_pos);
        static const char *tn[4] = { "DNA", "RNA", "protein", "text" };
        char buf[256];
        switch (bad_kind) {
        case BAD_ID: snprintf(buf, sizeof buf, "unexpected character '%c' in ID of sequence %llu\n", ch, (unsigned long long)k + 1); break;
        case BAD_COMMENT: snprintf(buf, sizeof buf, "unexpected character '%c' in comment of sequence %llu\n", ch, (unsigned long long)k + 1); break;
        case BAD_SEQ: snprintf(buf, sizeof buf, "unexpected %s code '%c' in sequence %llu\n", tn[o.seq_type & 3], ch, (unsigned long long)k + 1); break;
        case BAD_QUAL: snprintf(buf, sizeof buf, "unexpected quality code '%c' in sequence %llu\n", ch, (unsigned long long)k + 1); break;
        case BAD_NOPLUS: snprintf(buf, sizeof buf, "invalid FASTQ input: can't find '+' line of sequence %llu\n", (unsigned long long)k + 1); break;
        case BAD_NOAT: snprintf(buf, sizeof buf, "invalid FASTQ input: Can't find '@' after sequence %llu\n", (unsigned long long)k); break;
        default: snprintf(buf, sizeof buf, "not well-formed FASTQ input\n"); break;
        }
        die_input(buf);
    }
    if (mism_rec != ~0ull) {
        char buf[256];
        if (C.wf) snprintf(buf, sizeof buf, "quality length of sequence %llu doesn't match sequence length\n", (unsigned long long)mism_rec + 1);
        else snprintf(buf, sizeof buf, "quality length of sequence %llu (%llu) doesn't match sequence length (%llu)\n",
                      (unsigned long long)mism_rec + 1, (unsigned long long)mism_q, (unsigned long long)mism_s);
        die_input(buf);
    }
    if (C.fastq) {
        if (end_state == FQ_NAME || end_state == FQ_COMMENT) die_input("truncated FASTQ input: last sequence has no sequence data\n");
        if (end_state == FQ_SEQ || end_state == FQ_AFTER_SEQ || end_state == FQ_PLUS || (end_state == FQ_BEFORE_QUAL && !C.wf))
            die_input("truncated FASTQ input: last sequence has no quality\n");
    }
    if (info) for (int k = 0; k < 4; k++) for (int c = 0; c < 257; c++) info->unexpected[k][c] = h_unexp[k * 257 + c];

    // ---- lengths: u32 units with 0xFFFFFFFF continuation (encoders.c:72)
    S.n_records = n_rec; S.n_bases = n_seq;
    {
        u64 *units_pre = ex.alloc<u64>(n_rec + 2);
        const u64 *rse = rec_seq_end;
        exclusive_scan(ex, [rse] __device__ (size_t i) { u64 L = rse[i] - (i ? rse[i - 1] : 0); return L / 0xFFFFFFFFull + 1; }, n_rec, units_pre);
        u64 n_units; ex.download(&n_units, units_pre + n_rec, 8);
        S.n_len = n_units * 4;
        u32 *len = ex.alloc<u32>(n_units + 16);
        S.len = (u8 *)len;
        ex.for_each(n_rec, [=] __device__ (size_t i) {
            u64 L = rse[i] - (i ? rse[i - 1] : 0); u64 at = units_pre[i];
            while (L >= 0xFFFFFFFFull) { len[at++] = 0xFFFFFFFFu; L -= 0xFFFFFFFFull; }
            len[at] = (u32)L;
        });
        // longest line: FASTA tracked line ends; FASTQ = longest read (process.c:495)
        if (C.fastq) {
            unsigned long long *dl = d_longest;
            ex.for_each(n_rec, [=] __device__ (size_t i) { u64 L = rse[i] - (i ? rse[i - 1] : 0); if (L) atomicMax(dl, (unsigned long long)L); });
        }
        unsigned long long lg; ex.download(&lg, d_longest, 8);
        S.longest = lg;
    }
    S.n_ids = n_ids; S.n_comm = n_comm; S.n_qual = n_qual;

    // ---- sequence stream: 4-bit pack (+ mask) for DNA/RNA, bytes as they are for protein/text
    if (o.seq_type < NAFGPU_PROTEIN) {
        u8 lut[256];
        for (int c = 0; c < 256; c++) {                                    // tables.c:189 nuc_code
            int u = (c >= 'a' && c <= 'z') ? c - 32 : c;
            const char *order = "-TGKCYSBAWRDMHV"; const char *q = u ? strchr(order, u) : nullptr;
            lut[c] = u == 'U' ? 1 : (q ? (u8)(q - order) : 15);
        }
        u8 *d_lut = ex.alloc<u8>(256);
        ex.upload(d_lut, lut, 256);
        CUDA_TRY(cudaStreamSynchronize(ex.stream));                        // lut is a stack array
        S.n_seq = (n_seq + 1) / 2;
        S.seq = ex.alloc<u8>(S.n_seq + 64);
        u64 nwords = (n_seq + 31) / 32;
        u32 *casebits = ex.alloc<u32>(nwords + 2);
        if (nwords) { KLAUNCH(ex, "k_pack4", k_pack4<<<(unsigned)((nwords + 255) / 256), 256, 0, ex.stream>>>(bases, n_seq, S.seq, casebits, S.store_mask, d_lut)); }
        if (S.store_mask && n_seq) {
            // case flips -> runs -> units (encoders.c:98-151; final run flushed by ennaf.c:511)
            u64 ft = (nwords + 255) / 256;
            u64 *fcount = ex.alloc<u64>(ft + 1), *fpre = ex.alloc<u64>(ft + 2);
            KLAUNCH(ex, "k_flip_count", k_flip_count<<<(unsigned)ft, 256, 0, ex.stream>>>(casebits, nwords, fcount));
            const u64 *fc = fcount;
            exclusive_scan(ex, [fc] __device__ (size_t i) { return fc[i]; }, ft, fpre);
            u64 R; ex.download(&R, fpre + ft, 8);                          // number of flips; runs = R + 1
            u64 *flip_pos = ex.alloc<u64>(R + 2);
            KLAUNCH(ex, "k_flip_scatter", k_flip_scatter<<<(unsigned)ft, 256, 0, ex.stream>>>(casebits, nwords, fpre, flip_pos));
            // casebits beyond n_seq are zero, so a masked tail produces one spurious flip at n_seq: drop it
            RunCalc rc{flip_pos, R, n_seq};
            u64 *upre = ex.alloc<u64>(R + 3);
            exclusive_scan(ex, [rc] __device__ (size_t k) { return rc.units(k); }, R + 1, upre);
            u64 n_units; ex.download(&n_units, upre + R + 1, 8);
            S.n_mask = n_units;
            S.mask = ex.alloc<u8>(n_units + 64);
            ex.fill(S.mask, 0xFF, n_units);
            u8 *mk = S.mask;
            ex.for_each(R + 1, [=] __device__ (size_t k) { u64 u = rc.units(k); if (u) mk[upre[k] + u - 1] = (u8)(rc.len(k) % 255); });
        }
    } else {
        S.seq = bases; S.n_seq = n_seq;
    }
    if (!S.len) S.len = ex.alloc<u8>(64);
    if (!S.mask) S.mask = ex.alloc<u8>(64);
    if (info) { info->n_sequences = n_rec; info->longest_line = S.longest; info->n_bases = n_seq; }
    return S;
}

SplitOut split_on_device(Ctx &ctx, CudaExec &ex, const u8 *d_text, size_t n, const nafgpu_enc_opts &o, nafgpu_enc_info *info)
{
    SplitDev S = split_streams(ctx, ex, d_text, n, o, info);
    SplitOut r;
    const u8 *p[6] = { S.ids, S.comm, S.len, S.mask, S.seq, S.qual };
    u64 sz[6] = { S.n_ids, S.n_comm, S.n_len, S.store_mask ? S.n_mask : 0, S.n_seq, S.store_qual ? S.n_qual : 0 };
    for (int k = 0; k < 6; k++) { r.d[k] = p[k]; r.size[k] = sz[k]; if (info) info->stream_raw[k] = sz[k]; }
    return r;
}

}  // namespace nafg

#include "zstd_enc.cu"

namespace nafg {

// ennaf.c:538-589: header, then per stream VLE(original size) VLE(compressed size - 4) frame-without-magic
EncodeOut encode_on_device(Ctx &ctx, CudaExec &ex, const u8 *d_text, size_t n, const nafgpu_enc_opts &o, nafgpu_enc_info *info)
{
    nafgpu_enc_info local; if (!info) info = &local;
    SplitDev S = split_streams(ctx, ex, d_text, n, o, info);
    const u8 *sp[6] = { S.ids, S.comm, S.len, S.mask, S.seq, S.qual };
    const u64 ss[6] = { S.n_ids, S.n_comm, S.n_len, S.n_mask, S.n_seq, S.n_qual };
    const bool present[6] = { true, true, true, (bool)S.store_mask, true, (bool)S.store_qual };
    const u64 orig[6] = { S.n_ids, S.n_comm, S.n_len, S.n_mask, S.n_bases, S.n_qual };   // sequence: bases, not bytes (ennaf.c:582)

    ZEncBatch batch;
    int which[6], ns = 0;
    for (int k = 0; k < 6; k++) if (present[k]) { which[ns++] = k; batch.add(sp[k], ss[k], k == 4 ? o.window_log : 0); }
    zstd_compress_batch(ctx, ex, batch);                       // sizes known on the host afterwards

    std::vector<u8> hdr;
    hdr.push_back(0x01); hdr.push_back(0xF9); hdr.push_back(0xEC);
    if (o.seq_type == NAFGPU_DNA) hdr.push_back(1); else { hdr.push_back(2); hdr.push_back((u8)o.seq_type); }
    const bool has_title = o.title != nullptr;
    hdr.push_back((u8)((has_title << 6) | (1 << 5) | (1 << 4) | (1 << 3) | (S.store_mask << 2) | (1 << 1) | S.store_qual));
    hdr.push_back(' ');
    nafc::put_vle(hdr, o.have_line_length ? o.line_length : S.longest);
    nafc::put_vle(hdr, S.n_records);
    if (has_title) { size_t tl = strlen(o.title); nafc::put_vle(hdr, tl); hdr.insert(hdr.end(), o.title, o.title + tl); }
    // layout: [hdr][vle vle payload]...
    std::vector<std::vector<u8>> sec_hdr(ns);
    u64 total = hdr.size();
    std::vector<u64> payload_at(ns);
    for (int j = 0; j < ns; j++) {
        int k = which[j];
        u64 csz = batch.frame_size[j] - 4;                     // magic stripped (compressor.c:158)
        nafc::put_vle(sec_hdr[j], orig[k]); nafc::put_vle(sec_hdr[j], csz);
        total += sec_hdr[j].size();
        payload_at[j] = total; total += csz;
        info->stream_comp[k] = csz; info->stream_raw[k] = ss[k];
    }
    u8 *d_naf = ex.alloc<u8>(total + 64);
    std::vector<u8> small(hdr);
    // upload the header and the tiny per-section headers; gather the frames next to them
    ex.upload(d_naf, hdr.data(), hdr.size());
    for (int j = 0; j < ns; j++) ex.upload(d_naf + payload_at[j] - sec_hdr[j].size(), sec_hdr[j].data(), sec_hdr[j].size());
    CUDA_TRY(cudaStreamSynchronize(ex.stream));                // the small host vectors above must outlive the copies
    for (int j = 0; j < ns; j++) batch.dest[j] = d_naf + payload_at[j] - 4;   // frame byte i lands at dest + i; bytes 0..3 (magic) are skipped
    zstd_gather_frames(ctx, ex, batch, true);
    ex.check();
    return EncodeOut{d_naf, total};
}

EncodeOut zstd_compress_on_device(Ctx &ctx, CudaExec &ex, const u8 *d_src, size_t n, int window_log)
{
    ZEncBatch batch;
    batch.add(d_src, n, window_log);
    zstd_compress_batch(ctx, ex, batch);
    u8 *out = ex.alloc<u8>(batch.frame_size[0] + 64);
    batch.dest[0] = out;
    zstd_gather_frames(ctx, ex, batch, false);
    ex.check();
    return EncodeOut{out, batch.frame_size[0]};
}

}  // namespace nafg
