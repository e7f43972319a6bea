// naf_parse.cuh — the FASTA / FASTQ parser of ennaf as a data-parallel finite-state transducer.
//
// Replaces ennaf/src/process.c:358 process_non_well_formed_fasta, :477 process_non_well_formed_fastq and
// their --well-formed variants (:314, :430), including the in_get_until / str_append_char plumbing
// (:177-311) and the unexpected-character accounting (:99-141).
//
// The reference consumes one byte at a time; what it does with a byte depends only on (a) the byte's
// class and (b) which part of a record it is in.  That is a finite-state machine — 4 states for FASTA,
// 11 for FASTQ (3 of them error sinks) — whose per-byte behaviour is tabulated once on the host:
//     act[state][class] = next state | destination stream | which byte to emit | events
// Almost every byte of real input is "ordinary" (a printable byte that is not '+', '@' / '>'): after the
// first one the machine sits in a stable state and every further ordinary byte just goes to that state's
// stream.  The kernels therefore find the few non-ordinary bytes of a 64-byte chunk with SIMD-in-register
// compares, run only those (and the first byte of each ordinary run) through the action table, and handle the
// runs in between in bulk.  The host proves this shortcut equivalent to the table for the configured
// alphabet (analyse_ordinary); if it is not, every byte simply goes through the table.
// Three passes over the text, each thread owning 64 consecutive bytes:
//   k_fsm_reduce   fold my bytes into a state->state map (4 bits per state); maps compose associatively;
//                  per-thread map saved (8 B per 64 B of text), per-tile map for the scan
//   k_fsm_scan     one CTA: parser state entering every tile
//   k_fsm_count    entry state per thread (warp scan of saved maps) -> walk -> how many bytes I emit to
//                  ids / comments / sequence / quality, how many records end; saved (8 B / thread)
//   (device-wide exclusive sums over tiles)
//   k_fsm_scatter  block scan of the saved counts -> my offsets -> walk again, writing every byte to its
//                  stream and every record end to the per-record arrays
#pragma once
#include "common.cuh"

namespace nafg {

enum : u32 { F_EOL = 1, F_SPACE = 2, F_ID_OK = 4, F_COMM_OK = 8, F_SEQ_OK = 16, F_QUAL_OK = 32, F_START = 64, F_PLUS = 128 };

enum { FA_NAME = 0, FA_COMMENT = 1, FA_SEQ_LS = 2, FA_SEQ_MID = 3, FA_NSTATES = 4 };
enum { FQ_NAME = 0, FQ_COMMENT = 1, FQ_SEQ = 2, FQ_AFTER_SEQ = 3, FQ_PLUS = 4, FQ_BEFORE_QUAL = 5, FQ_QUAL = 6, FQ_AFTER_QUAL = 7,
       FQ_ERR_NOPLUS = 8, FQ_ERR_NOAT = 9, FQ_ERR_NOTWF = 10, FQ_NSTATES = 11 };
static const int MAX_STATES = 11;

// action word
enum : u32 {
    A_NEXT = 0xF,                                  // bits 0-3  next state
    A_DEST_SHIFT = 4, A_DEST = 7u << 4,            // bits 4-6  0 none, 1 ids, 2 comments, 3 sequence, 4 quality
    A_BYTE_SHIFT = 7, A_BYTE = 7u << 7,            // bits 7-9  0 the byte itself, 1 NUL, 2 sequence replacement, 3 '?', 4 '!'
    A_COMM_NUL = 1u << 10,                         // also terminate the comment (name ended by end of line)
    A_UNCOUNTED = 1u << 11,                        // sequence byte that no record length counts (process.c:366 quirk)
    A_REC_END = 1u << 12, A_LINE_END = 1u << 13,
    A_BAD_SHIFT = 14, A_BAD = 3u << 14             // bits 14-15  0 fine, 1 unexpected character, 2 FSM error
};
enum { DEST_IDS = 1, DEST_COMM = 2, DEST_SEQ = 3, DEST_QUAL = 4 };
enum { BAD_ID = 1, BAD_COMMENT = 2, BAD_SEQ = 3, BAD_QUAL = 4, BAD_NOPLUS = 5, BAD_NOAT = 6, BAD_NOTWF = 7 };

struct ParseCfg {
    int fastq, wf, seq_type, text_fasta, no_mask, strict, nstates;
    u8 repl;
};

struct FsmTables {
    u8  cls[256];                    // byte -> F_* class flags
    u8  tid[256];                    // byte -> id of its state map (bytes that move the machine alike share an id)
    u64 trans[32];                   // id -> packed state map (4 bits per source state)
    u8  idem[32];                    // that map is idempotent
    u16 act[MAX_STATES][256];        // (state, class) -> action word
    // ordinary-run shortcut (valid iff bulk_ok)
    u32 bulk_ok;                     // SWAR predicate + per-state bulk rule proven equivalent to the table
    u32 ord_thr4;                    // ordinary <=> byte > thr and byte not in exc[0..4)
    u32 ord_exc4[4];
    u8  stable[MAX_STATES + 5];      // state unchanged by ordinary bytes
    u8  bulk_dest[MAX_STATES + 5];   // stream ordinary bytes go to in that state (0 = dropped)
    u8  bulk_need[MAX_STATES + 5];   // class flag an ordinary byte needs there, else it is "unexpected"
    u8  bulk_repl[MAX_STATES + 5];   // byte kind written instead of an unexpected byte
    u8  bulk_badk[MAX_STATES + 5];   // which unexpected-character counter (0 id, 1 comment, 2 sequence, 3 quality)
};

inline u32 fa_next(u32 s, u32 f)
{
    switch (s) {
    case FA_NAME:    return (f & F_ID_OK) ? FA_NAME : ((f & F_SPACE) ? ((f & F_EOL) ? FA_SEQ_LS : FA_COMMENT) : FA_NAME);
    case FA_COMMENT: return (f & F_EOL) ? FA_SEQ_LS : FA_COMMENT;
    case FA_SEQ_LS:  return (f & F_START) ? FA_NAME : ((f & F_EOL) ? FA_SEQ_LS : FA_SEQ_MID);
    default:         return (f & F_EOL) ? FA_SEQ_LS : FA_SEQ_MID;
    }
}
inline u32 fq_next(u32 s, u32 f, bool wf)
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

// what the reference does with one byte of class f in state s (process.c, restated per state)
inline u32 action_of(const ParseCfg &C, u32 s, u32 f)
{
    u32 dest = 0, byte = 0, flags = 0, bad = 0;
    auto emit = [&](u32 d, u32 b) { dest = d; byte = b; };
    const u32 ns = C.fastq ? fq_next(s, f, C.wf) : fa_next(s, f);
    if (!C.fastq) {
        switch (s) {
        case FA_NAME:
            if (f & F_ID_OK) emit(DEST_IDS, 0);
            else if (f & F_SPACE) { emit(DEST_IDS, 1); if (f & F_EOL) flags |= A_COMM_NUL; }
            else { emit(DEST_SEQ, 3); flags |= A_UNCOUNTED; bad = 1; }              // process.c:366
            break;
        case FA_COMMENT:
            if (f & F_COMM_OK) emit(DEST_COMM, 0);
            else if (f & F_EOL) emit(DEST_COMM, 1);
            else { emit(DEST_COMM, 3); bad = 1; }
            break;
        default:
            if (s == FA_SEQ_LS && (f & F_START)) { flags |= A_REC_END; break; }
            if (f & F_SEQ_OK) emit(DEST_SEQ, 0);
            else if (f & F_EOL) flags |= A_LINE_END;
            else if (f & F_SPACE) {}
            else if (C.text_fasta && (f & F_START)) emit(DEST_SEQ, 0);              // process.c:413
            else { emit(DEST_SEQ, 2); bad = 1; }
            break;
        }
    } else {
        switch (s) {
        case FQ_NAME:
            if (f & F_ID_OK) emit(DEST_IDS, 0);
            else if (f & F_SPACE) { emit(DEST_IDS, 1); if (f & F_EOL) flags |= A_COMM_NUL; }
            else { emit(DEST_SEQ, 3); flags |= A_UNCOUNTED; bad = 1; }              // process.c:485
            break;
        case FQ_COMMENT:
            if (f & F_COMM_OK) emit(DEST_COMM, 0);
            else if (f & F_EOL) emit(DEST_COMM, 1);
            else { emit(DEST_COMM, 3); bad = 1; }
            break;
        case FQ_SEQ:
            if (f & F_SEQ_OK) emit(DEST_SEQ, 0);
            else if ((f & F_EOL) || (f & F_SPACE)) {}
            else { emit(DEST_SEQ, 2); bad = 1; }
            break;
        case FQ_BEFORE_QUAL:
            if (!(f & F_EOL)) emit(DEST_QUAL, 0);                                    // process.c:523: unvalidated
            else if (C.wf) flags |= A_REC_END;                                       // empty quality line
            break;
        case FQ_QUAL:
            if (f & F_QUAL_OK) emit(DEST_QUAL, 0);
            else if (f & F_EOL) flags |= A_REC_END;
            else if (f & F_SPACE) {}
            else { emit(DEST_QUAL, 4); bad = 1; }
            break;
        default: break;
        }
        if (ns >= FQ_ERR_NOPLUS && s < FQ_ERR_NOPLUS) bad = 2;
    }
    return ns | (dest << A_DEST_SHIFT) | (byte << A_BYTE_SHIFT) | flags | (bad << A_BAD_SHIFT);
}

// Derive the ordinary-run shortcut from the tables and verify it against them exhaustively.
inline void analyse_ordinary(const ParseCfg &c, FsmTables &t)
{
    t.bulk_ok = 0;
    // the map shared by most bytes
    int cnt[32] = {0}, ord = 0;
    for (int ch = 0; ch < 256; ch++) cnt[t.tid[ch]]++;
    for (int id = 1; id < 32; id++) if (cnt[id] > cnt[ord]) ord = id;
    if (!t.idem[ord]) return;
    // ordinary set as "byte > thr, minus up to four exceptions"
    int thr = -1;
    for (int ch = 0; ch < 256; ch++) if (t.tid[ch] != ord) thr = ch; else break;      // leading non-ordinary prefix [0, thr]
    int exc[4] = {-1, -1, -1, -1}, ne = 0;
    for (int ch = thr + 1; ch < 256; ch++) if (t.tid[ch] != ord) { if (ne == 4) return; exc[ne++] = ch; }
    if (thr < 0) {            // SWAR compare is "greater than": no threshold needed -> make byte 0 an exception instead
        if (t.tid[0] == ord) { thr = 0; /* byte 0 ordinary: cannot express with > */ return; }
    }
    for (int k = 0; k < 4; k++) { int e = exc[k] < 0 ? (exc[0] < 0 ? (thr >= 0 ? thr : 0) : exc[0]) : exc[k]; t.ord_exc4[k] = (u32)e * 0x01010101u; }
    t.ord_thr4 = (u32)(thr < 0 ? 0 : thr) * 0x01010101u;
    for (int ch = 0; ch < 256; ch++) {
        bool o = ch > (thr < 0 ? -1 : thr);
        for (int k = 0; k < 4; k++) if ((int)(t.ord_exc4[k] & 0xFF) == ch) o = false;
        if (o != (t.tid[ch] == ord)) return;
    }
    // per-state bulk rule, checked against the action table for every ordinary byte
    for (int s = 0; s < c.nstates; s++) {
        const u32 m = (u32)((t.trans[ord] >> (4 * s)) & 15);
        t.stable[s] = m == (u32)s;
        if (!t.stable[s]) continue;
        int dest = -1, need = 0, repl = 0, badk = 0;
        for (int ch = 0; ch < 256; ch++) {
            if (t.tid[ch] != ord) continue;
            const u32 a = t.act[s][t.cls[ch]];
            const int d = (a & A_DEST) >> A_DEST_SHIFT, kind = (a & A_BYTE) >> A_BYTE_SHIFT, bad = (a & A_BAD) >> A_BAD_SHIFT;
            if (dest < 0) dest = d;
            if (d != dest || (a & (A_COMM_NUL | A_REC_END | A_LINE_END | A_UNCOUNTED)) || bad == 2) return;
            if (bad == 1) { repl = kind; badk = d == DEST_QUAL ? 3 : (d == DEST_COMM ? 1 : 2); }
            else if (kind != 0 && d != 0) return;
        }
        // which single class flag separates the good ordinary bytes from the unexpected ones in this state?
        const u32 flags[4] = { 0, F_SEQ_OK, F_QUAL_OK, F_COMM_OK };
        bool found = false;
        for (int k = 0; k < 4 && !found; k++) {
            bool okk = true;
            for (int ch = 0; ch < 256 && okk; ch++) {
                if (t.tid[ch] != ord) continue;
                const u32 a = t.act[s][t.cls[ch]];
                const bool bad = ((a & A_BAD) >> A_BAD_SHIFT) == 1;
                const bool pred_bad = flags[k] ? !(t.cls[ch] & flags[k]) : false;
                if (bad != pred_bad) okk = false;
            }
            if (okk) { need = (int)flags[k]; found = true; }
        }
        if (!found) return;
        t.bulk_dest[s] = (u8)(dest < 0 ? 0 : dest); t.bulk_need[s] = (u8)need; t.bulk_repl[s] = (u8)repl; t.bulk_badk[s] = (u8)badk;
    }
    t.bulk_ok = 1;
}

inline void build_tables(const ParseCfg &c, FsmTables &t)
{
    auto is_eol = [](int ch) { return ch >= 0x0A && ch <= 0x0D; };
    auto is_space = [&](int ch) { return (ch >= 0x09 && ch <= 0x0D) || ch == 0x20; };
    auto in_set = [](int ch, const char *set) { if (ch >= 'a' && ch <= 'z') ch -= 32; return ch > 0 && strchr(set, ch) != nullptr; };
    memset(&t, 0, sizeof t);
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
            case NAFGPU_DNA:     ok = in_set(ch, "-ABCDGHKMNRSTVWY"); break;             // tables.c:72
            case NAFGPU_RNA:     ok = in_set(ch, "-ABCDGHKMNRSUVWY"); break;             // tables.c:82
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
    int nmaps = 0;
    for (int f = 0; f < 256; f++) for (int s = 0; s < c.nstates; s++) t.act[s][f] = (u16)action_of(c, s, f);
    for (int ch = 0; ch < 256; ch++) {
        const int f = t.cls[ch];
        u64 m = 0;
        for (int s = 0; s < c.nstates; s++) m |= (u64)(t.act[s][f] & A_NEXT) << (4 * s);
        int id = 0;
        while (id < nmaps && t.trans[id] != m) id++;
        if (id == nmaps) {                                   // at most a handful: ordinary, EOL, blank, '+', '@' / '>', non-id bytes
            t.trans[nmaps++] = m;
            u64 mm = 0;
            for (int s = 0; s < c.nstates; s++) mm |= ((m >> (4 * ((m >> (4 * s)) & 15))) & 15) << (4 * s);
            t.idem[id] = mm == m;
        }
        t.tid[ch] = (u8)id;
    }
    analyse_ordinary(c, t);
}

__device__ __forceinline__ u64 map_compose(u64 f, u64 g, int ns)     // first f, then g
{
    u64 h = 0;
    for (int s = 0; s < ns; s++) h |= ((g >> (4 * ((f >> (4 * s)) & 15))) & 15) << (4 * s);
    return h;
}
__device__ __forceinline__ u64 map_identity(int ns) { u64 m = 0; for (int s = 0; s < ns; s++) m |= (u64)s << (4 * s); return m; }
__device__ __forceinline__ u32 map_apply(u64 m, u32 s) { return (u32)((m >> (4 * s)) & 15); }

static const int PT = 256, PB = 64, PTILE = PT * PB;       // threads per CTA, bytes per thread, bytes per tile (16 KB)
static const int PBS = PB + 4;                             // shared-memory stride of a thread's chunk (odd word count: no bank conflicts)

// per-thread record saved by k_fsm_count: entry state, emitted byte counts, line-end marker (8 bytes)
struct ThreadInfo { u8 state, ids, comm, seq, cnt, qual, rec, line; };   // line: 0 = no line end, else 1 + counted bytes before my last line end

struct TileCounts { u64 ids, comm, seq, seq_counted, qual, rec, line_last; u32 has_line; u32 pad; };

struct ParseArgs {
    const u8 *text; u64 n, p0;                // bytes before p0 (leading white space + the first '>' / '@') are skipped
    ParseCfg cfg;
    const FsmTables *tab;
    u64 *thread_map; u64 *tile_map; u8 *tile_state;
    ThreadInfo *tinfo;
    TileCounts *tile;
    u64 ntiles;
    // scatter
    const u64 *pre_ids, *pre_comm, *pre_seq, *pre_cnt, *pre_qual, *pre_rec, *pre_line;
    u8 *ids, *comm, *bases, *qual;
    u64 *rec_seq_end, *rec_qual_end, *rec_pos;
    unsigned long long *unexpected, *longest, *first_bad;
};

__device__ __forceinline__ void load_tables(FsmTables *T, const FsmTables *src)
{
    for (int i = threadIdx.x; i < (int)(sizeof(FsmTables) / 4); i += blockDim.x) ((u32 *)T)[i] = ((const u32 *)src)[i];
    __syncthreads();
}

// My 64 bytes -> my private 68-byte row of the shared tile (zero beyond the end of the text), and two masks:
//   live   bytes inside [p0, n)
//   tbl    live bytes that must go through the action table: non-ordinary bytes and the first byte of every
//          ordinary run (the chunk's first live byte included).  Everything else is bulk.
struct Chunk { u64 live, tbl; };

__device__ __forceinline__ Chunk load_chunk(const ParseArgs &A, const FsmTables &T, u64 lo, u8 *row)
{
    u32 w[16];
    if ((((uintptr_t)A.text) & 15) == 0 && lo + PB <= A.n) {
        const uint4 *v = (const uint4 *)(A.text + lo);
#pragma unroll
        for (int k = 0; k < 4; k++) { uint4 x = __ldg(v + k); w[4 * k] = x.x; w[4 * k + 1] = x.y; w[4 * k + 2] = x.z; w[4 * k + 3] = x.w; }
    } else {
#pragma unroll
        for (int k = 0; k < 16; k++) {
            u32 x = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) { u64 p = lo + 4 * k + j; if (p < A.n) x |= (u32)A.text[p] << (8 * j); }
            w[k] = x;
        }
    }
    u64 ord = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        ((u32 *)row)[k] = w[k];
        // ordinary <=> byte > thr and not one of four exceptions; one bit per byte via a multiply-gather
        u32 o = __vcmpgtu4(w[k], T.ord_thr4) & ~(__vcmpeq4(w[k], T.ord_exc4[0]) | __vcmpeq4(w[k], T.ord_exc4[1]) | __vcmpeq4(w[k], T.ord_exc4[2]) | __vcmpeq4(w[k], T.ord_exc4[3]));
        ord |= (u64)(((o & 0x01010101u) * 0x01020408u) >> 24) << (4 * k);
    }
    Chunk c;
    const u64 b0 = lo >= A.p0 ? 0 : (A.p0 - lo >= 64 ? 64 : A.p0 - lo), b1 = lo + 64 <= A.n ? 64 : (lo >= A.n ? 0 : A.n - lo);
    c.live = (b1 >= 64 ? ~0ull : ((1ull << b1) - 1)) & ~(b0 >= 64 ? ~0ull : ((1ull << b0) - 1));
    if (!T.bulk_ok) ord = 0;
    ord &= c.live;
    const u64 prev_ord = (ord << 1);                          // bit i: byte i-1 is ordinary (and live)
    c.tbl = c.live & (~ord | ~prev_ord);
    return c;
}

// ------------------------------------------------------------------ pass 1
__global__ void __launch_bounds__(PT) k_fsm_reduce(const ParseArgs A)
{
    __shared__ FsmTables T;
    __shared__ u64 wmap[PT / 32];
    __shared__ __align__(16) u8 tile[PT * PBS];
    load_tables(&T, A.tab);
    const int ns = A.cfg.nstates;
    const u64 lo = (u64)blockIdx.x * PTILE + (u64)threadIdx.x * PB;
    u8 *row = tile + threadIdx.x * PBS;
    const Chunk c = load_chunk(A, T, lo, row);
    // only table bytes can change the composed map: bulk bytes repeat an idempotent map
    u64 f = map_identity(ns);
    for (u64 m = c.tbl; m;) {
        const int i = __ffsll((long long)m) - 1;
        m &= m - 1;
        f = map_compose(f, T.trans[T.tid[row[i]]], ns);
    }
    A.thread_map[(u64)blockIdx.x * PT + threadIdx.x] = f;
    for (int d = 1; d < 32; d <<= 1) {                       // ordered reduction: lane i absorbs lane i+d
        u64 g = __shfl_down_sync(0xFFFFFFFFu, f, d);
        if (((threadIdx.x & 31) % (2 * d)) == 0) f = map_compose(f, g, ns);
    }
    if ((threadIdx.x & 31) == 0) wmap[threadIdx.x >> 5] = f;
    __syncthreads();
    if (threadIdx.x == 0) {
        u64 m = wmap[0];
        for (int k = 1; k < PT / 32; k++) m = map_compose(m, wmap[k], ns);
        A.tile_map[blockIdx.x] = m;
    }
}

// one CTA: chunked scan of the tile maps -> entry state of every tile (the machine starts in NAME)
__global__ void __launch_bounds__(1024) k_fsm_scan(const ParseArgs A)
{
    __shared__ u64 cmap[1024];
    const int ns = A.cfg.nstates;
    u64 per = (A.ntiles + 1023) / 1024;
    u64 lo = (u64)threadIdx.x * per, hi = lo + per;
    if (lo > A.ntiles) lo = A.ntiles;
    if (hi > A.ntiles) hi = A.ntiles;
    u64 f = map_identity(ns);
    for (u64 t = lo; t < hi; t++) f = map_compose(f, A.tile_map[t], ns);
    cmap[threadIdx.x] = f;
    __syncthreads();
    if (threadIdx.x == 0) {
        u32 s = 0;
        for (int c = 0; c < 1024; c++) { u64 m = cmap[c]; cmap[c] = s; s = map_apply(m, s); }
    }
    __syncthreads();
    u32 s = (u32)cmap[threadIdx.x];
    for (u64 t = lo; t < hi; t++) { A.tile_state[t] = (u8)s; s = map_apply(A.tile_map[t], s); }
    if (hi == A.ntiles && lo < hi) A.tile_state[A.ntiles] = (u8)s;      // state at end of input
    if (A.ntiles == 0 && threadIdx.x == 0) A.tile_state[0] = 0;
}

// ------------------------------------------------------------------ the walk (count or scatter)
struct Emit { u32 ids, comm, seq, cnt, qual, rec; };

template <bool SCATTER>
__device__ __forceinline__ void walk(const ParseArgs &A, const FsmTables &T, const u8 *row, const Chunk ch, u64 lo, u32 s, Emit &n,
                                     u8 *st_ids, u8 *st_comm, u8 *st_seq, u8 *st_qual,      // SCATTER: where my bytes go (shared-memory staging)
                                     u64 o_cnt, u64 o_qual, u64 o_rec,
                                     u64 &line_base, u64 &line_max, u32 &line_mark)
{
    const ParseCfg &C = A.cfg;
    const bool upper = C.seq_type >= NAFGPU_PROTEIN && C.no_mask;      // process.c:49
    u64 live = ch.live;
    while (live) {
        int i = __ffsll((long long)live) - 1;
        if (!((ch.tbl >> i) & 1)) {
            // ---- bulk: ordinary bytes in a stable state, up to the next table byte
            const u64 rest = ch.tbl & ~((2ull << i) - 1);          // table bytes after i  (note: (2<<63) == 0 is fine: rest = tbl & ~(-1) = 0)
            const u64 above = live & ~((1ull << i) - 1);
            int e = rest ? __ffsll((long long)rest) - 1 : 64;
            // the run also ends at the end of the live range
            const int last_live = 63 - __clzll((long long)above);
            if (e > last_live + 1) e = last_live + 1;
            const u32 len = (u32)(e - i);
            const u32 dest = T.bulk_dest[s];
            if (SCATTER && dest) {
                const u32 need = T.bulk_need[s];
                u8 *dst = dest == DEST_SEQ ? st_seq + n.seq : (dest == DEST_QUAL ? st_qual + n.qual : (dest == DEST_IDS ? st_ids + n.ids : st_comm + n.comm));
                for (int k = i; k < e; k++) {
                    u32 c = row[k];
                    if (need && !(T.cls[c] & need)) {               // unexpected character: count it, write the replacement
                        atomicAdd(&A.unexpected[T.bulk_badk[s] * 257 + c], 1ull);
                        if (C.strict) atomicMin(A.first_bad, (unsigned long long)(((lo + k) << 8) | (T.bulk_badk[s] == 3 ? BAD_QUAL : (T.bulk_badk[s] == 1 ? BAD_COMMENT : BAD_SEQ))));
                        const u32 kind = T.bulk_repl[s];
                        c = kind == 2 ? (u32)C.repl : (kind == 3 ? (u32)'?' : (u32)'!');
                    }
                    if (upper && dest == DEST_SEQ && c >= 'a' && c <= 'z') c -= 32;
                    dst[k - i] = (u8)c;
                }
            }
            if (dest == DEST_SEQ) { n.seq += len; n.cnt += len; }
            else if (dest == DEST_QUAL) n.qual += len;
            else if (dest == DEST_IDS) n.ids += len;
            else if (dest == DEST_COMM) n.comm += len;
            live = e >= 64 ? 0 : (live & ~((1ull << e) - 1));
            continue;
        }
        // ---- one byte through the action table
        live &= live - 1;
        const u64 p = lo + i;
        const u32 c = row[i];
        const u32 a = T.act[s][T.cls[c]];
        const u32 dest = (a & A_DEST) >> A_DEST_SHIFT, kind = (a & A_BYTE) >> A_BYTE_SHIFT;
        if (SCATTER && dest) {
            u32 b = kind == 0 ? c : (kind == 1 ? 0u : (kind == 2 ? (u32)C.repl : (kind == 3 ? (u32)'?' : (u32)'!')));
            if (dest == DEST_SEQ) { if (upper && b >= 'a' && b <= 'z') b -= 32; st_seq[n.seq] = (u8)b; }
            else if (dest == DEST_QUAL) st_qual[n.qual] = (u8)b;
            else if (dest == DEST_IDS) st_ids[n.ids] = (u8)b;
            else st_comm[n.comm] = (u8)b;
        }
        n.ids += dest == DEST_IDS; n.comm += dest == DEST_COMM; n.qual += dest == DEST_QUAL;
        n.seq += dest == DEST_SEQ; n.cnt += (dest == DEST_SEQ) & !(a & A_UNCOUNTED);
        if (a & (A_COMM_NUL | A_REC_END | A_LINE_END | A_BAD)) {
            if (a & A_COMM_NUL) { if (SCATTER) st_comm[n.comm] = 0; n.comm++; }
            if (a & A_LINE_END) {
                const u64 v = o_cnt + n.cnt;                    // sequence bytes since the previous line end
                if (v - line_base > line_max) line_max = v - line_base;
                line_base = v; line_mark = n.cnt + 1;
            }
            if (a & A_REC_END) {
                if (SCATTER) {
                    const u64 r = o_rec + n.rec;
                    A.rec_seq_end[r] = o_cnt + n.cnt;
                    if (C.fastq) A.rec_qual_end[r] = o_qual + n.qual;
                    A.rec_pos[r] = p;
                }
                n.rec++;
            }
            if (SCATTER && (a & A_BAD)) {
                const u32 bad = (a & A_BAD) >> A_BAD_SHIFT;
                u32 code;
                if (bad == 1) {
                    const int k = dest == DEST_QUAL ? 3 : (dest == DEST_COMM ? 1 : ((a & A_UNCOUNTED) ? 0 : 2));
                    atomicAdd(&A.unexpected[k * 257 + c], 1ull);
                    code = k == 0 ? BAD_ID : (k == 1 ? BAD_COMMENT : (k == 2 ? BAD_SEQ : BAD_QUAL));
                } else {
                    const u32 ns2 = a & A_NEXT;
                    code = ns2 == FQ_ERR_NOPLUS ? BAD_NOPLUS : (ns2 == FQ_ERR_NOAT ? BAD_NOAT : BAD_NOTWF);
                }
                if (bad == 2 || C.strict) atomicMin(A.first_bad, (unsigned long long)((p << 8) | code));
            }
        }
        s = a & A_NEXT;
    }
}

// ------------------------------------------------------------------ pass 2
__global__ void __launch_bounds__(PT) k_fsm_count(const ParseArgs A)
{
    __shared__ FsmTables T;
    __shared__ u64 sm[33];
    __shared__ u64 wmap[PT / 32];
    __shared__ u32 wstate[PT / 32];
    __shared__ __align__(16) u8 tile[PT * PBS];
    load_tables(&T, A.tab);
    const int ns = A.cfg.nstates;
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u64 gid = (u64)blockIdx.x * PT + threadIdx.x, lo = gid * PB;
    // entry state: tile entry state pushed through the saved maps of the threads before me
    u64 incl = A.thread_map[gid];
    for (int d = 1; d < 32; d <<= 1) { u64 g = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= (unsigned)d) incl = map_compose(g, incl, ns); }
    if (lane == 31) wmap[warp] = incl;
    __syncthreads();
    if (threadIdx.x == 0) {
        u32 s = A.tile_state[blockIdx.x];
        for (int k = 0; k < PT / 32; k++) { wstate[k] = s; s = map_apply(wmap[k], s); }
    }
    __syncthreads();
    const u64 excl = __shfl_up_sync(0xFFFFFFFFu, incl, 1);
    u32 s0 = wstate[warp];
    if (lane > 0) s0 = map_apply(excl, s0);

    u8 *row = tile + threadIdx.x * PBS;
    const Chunk c = load_chunk(A, T, lo, row);
    Emit n = {0, 0, 0, 0, 0, 0};
    u64 line_base = 0, line_max = 0; u32 line_mark = 0;
    walk<false>(A, T, row, c, lo, s0, n, nullptr, nullptr, nullptr, nullptr, 0, 0, 0, line_base, line_max, line_mark);
    ThreadInfo ti; ti.state = (u8)s0; ti.ids = (u8)n.ids; ti.comm = (u8)n.comm; ti.seq = (u8)n.seq; ti.cnt = (u8)n.cnt; ti.qual = (u8)n.qual;
    ti.rec = (u8)n.rec; ti.line = (u8)line_mark;
    A.tinfo[gid] = ti;
    // tile totals
    TileCounts tc; u64 tot;
    block_excl_scan(n.ids, &tot, sm); tc.ids = tot;
    block_excl_scan(n.comm, &tot, sm); tc.comm = tot;
    block_excl_scan(n.seq, &tot, sm); tc.seq = tot;
    const u64 pre_cnt = block_excl_scan(n.cnt, &tot, sm); tc.seq_counted = tot;
    block_excl_scan(n.qual, &tot, sm); tc.qual = tot;
    block_excl_scan(n.rec, &tot, sm); tc.rec = tot;
    // counted-sequence offset (tile-relative) at the last line end in the tile, +1 (0 = none)
    u64 v = line_mark ? pre_cnt + line_mark : 0;
    for (int d = 16; d; d >>= 1) { u64 o = __shfl_xor_sync(0xFFFFFFFFu, v, d); if (o > v) v = o; }
    if (lane == 0) sm[warp] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        u64 m = 0;
        for (int k = 0; k < PT / 32; k++) if (sm[k] > m) m = sm[k];
        tc.has_line = m != 0; tc.line_last = m ? m - 1 : 0; tc.pad = 0;
        A.tile[blockIdx.x] = tc;
    }
}

// ------------------------------------------------------------------ pass 3
// Staged bytes of one stream -> global memory.  The staging region was placed so that it has the same
// alignment modulo 16 as its destination: head bytes, then 128-bit copies, then tail bytes.
__device__ __forceinline__ void tile_copy_out(u8 *dst, const u8 *s, u32 len)
{
    const u32 head = min(len, (u32)((16 - ((uintptr_t)dst & 15)) & 15));
    if (threadIdx.x < head) dst[threadIdx.x] = s[threadIdx.x];
    const u32 nv = (len - head) / 16;
    uint4 *dv = (uint4 *)(dst + head); const uint4 *sv = (const uint4 *)(s + head);
    for (u32 k = threadIdx.x; k < nv; k += blockDim.x) dv[k] = sv[k];
    const u32 done = head + nv * 16;
    if (threadIdx.x < len - done) dst[done + threadIdx.x] = s[done + threadIdx.x];
}

static const int STAGE_BYTES = PTILE + PTILE / 2 + 4 * 32;   // a byte emits at most 1.5 bytes on average; + alignment slack per stream

__global__ void __launch_bounds__(PT) k_fsm_scatter(const ParseArgs A)
{
    extern __shared__ __align__(16) u8 dyn[];                 // [FsmTables][text tile][staging]
    FsmTables &T = *(FsmTables *)dyn;
    u8 *tile = dyn + ((sizeof(FsmTables) + 15) & ~15);
    u8 *stage = tile + PT * PBS;
    __shared__ u64 sm[33];
    load_tables(&T, A.tab);
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u64 gid = (u64)blockIdx.x * PT + threadIdx.x, lo = gid * PB;
    const ThreadInfo ti = A.tinfo[gid];
    u64 t_ids, t_comm, t_seq, t_qual, tot;
    const u32 l_ids = (u32)block_excl_scan(ti.ids, &t_ids, sm);
    const u32 l_comm = (u32)block_excl_scan(ti.comm, &t_comm, sm);
    const u32 l_seq = (u32)block_excl_scan(ti.seq, &t_seq, sm);
    const u64 o_cnt = block_excl_scan(ti.cnt, &tot, sm) + A.pre_cnt[blockIdx.x];
    const u32 l_qual = (u32)block_excl_scan(ti.qual, &t_qual, sm);
    const u64 o_rec = block_excl_scan(ti.rec, &tot, sm) + A.pre_rec[blockIdx.x];
    const u64 o_qual = l_qual + A.pre_qual[blockIdx.x];
    // counted-sequence value at the last line end before this thread: exclusive max-scan (values are monotone)
    const u64 mine = ti.line ? o_cnt + ti.line : 0;            // +1 encoding
    u64 run = mine;
    for (int d = 1; d < 32; d <<= 1) { u64 g = __shfl_up_sync(0xFFFFFFFFu, run, d); if (lane >= (unsigned)d && g > run) run = g; }
    if (lane == 31) sm[warp] = run;
    __syncthreads();
    u64 before = A.pre_line[blockIdx.x] + 1;                   // tile carry; line base 0 = start of the data
    for (unsigned k = 0; k < warp; k++) if (sm[k] > before) before = sm[k];
    const u64 prev_lane = __shfl_up_sync(0xFFFFFFFFu, run, 1);
    if (lane > 0 && prev_lane > before) before = prev_lane;

    // staging layout: ids | comments | sequence | quality, each region congruent mod 16 to its destination
    u8 *g_ids = A.ids + A.pre_ids[blockIdx.x], *g_comm = A.comm + A.pre_comm[blockIdx.x];
    u8 *g_seq = A.bases + A.pre_seq[blockIdx.x], *g_qual = A.qual + A.pre_qual[blockIdx.x];
    u8 *s_ids = stage + ((uintptr_t)g_ids & 15);
    u8 *s_comm = (u8 *)(((uintptr_t)(s_ids + t_ids) + 15) & ~(uintptr_t)15) + ((uintptr_t)g_comm & 15);
    u8 *s_seq = (u8 *)(((uintptr_t)(s_comm + t_comm) + 15) & ~(uintptr_t)15) + ((uintptr_t)g_seq & 15);
    u8 *s_qual = (u8 *)(((uintptr_t)(s_seq + t_seq) + 15) & ~(uintptr_t)15) + ((uintptr_t)g_qual & 15);
    u8 *row = tile + threadIdx.x * PBS;
    const Chunk c = load_chunk(A, T, lo, row);
    Emit m = {0, 0, 0, 0, 0, 0};
    u64 line_base = before - 1, line_max = 0; u32 line_mark = 0;
    walk<true>(A, T, row, c, lo, ti.state, m, s_ids + l_ids, s_comm + l_comm, s_seq + l_seq, s_qual + l_qual, o_cnt, o_qual, o_rec, line_base, line_max, line_mark);
    if (!A.cfg.fastq) {
        // pending (unterminated) last line of the input: counted bytes after the last line end (process.c:417-422)
        if (lo < A.n && lo + PB >= A.n) { const u64 d = o_cnt + m.cnt - line_base; if (d > line_max) line_max = d; }
        // (skip the atomic unless this thread beats the current maximum: lines are mostly equally long, and an atomicMax per
        // thread on one address serialises in L2)
        if (line_max > *(volatile unsigned long long *)A.longest) atomicMax(A.longest, (unsigned long long)line_max);
    }
    __syncthreads();
    tile_copy_out(g_ids, s_ids, (u32)t_ids);
    tile_copy_out(g_comm, s_comm, (u32)t_comm);
    tile_copy_out(g_seq, s_seq, (u32)t_seq);
    tile_copy_out(g_qual, s_qual, (u32)t_qual);
}
static const size_t SCATTER_SMEM = ((sizeof(FsmTables) + 15) & ~15) + PT * PBS + STAGE_BYTES + 64;

}  // namespace nafg
