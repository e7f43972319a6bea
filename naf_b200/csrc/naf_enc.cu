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

#include "naf_parse.cuh"
#include "naf_parse_fast.cuh"
#include "naf_fused.cuh"

namespace nafg {

// ------------------------------------------------------------------ 4-bit pack + case bits
// encoders.c:30 encode_dna (first base in the low nibble; odd tail has a zero high nibble, ennaf.c:525)
// and the predicate of encoders.c:134 (masked <=> byte >= 96).  One thread: 32 bases -> 16 bytes + 1 word.
// nuc_code[] here carries bit 7 = "not an expected code for this alphabet" (tables.c:72,82); the general parser has
// already replaced such bytes, the canonical-input parser has not and learns about them through *flag (condition C4).
__global__ void k_pack4(const u8 *bases, u64 n, u8 *packed, u32 *casebits, int want_mask, const u8 *nuc_code, u32 *flag)
{
    __shared__ u8 c_nuc_code[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) c_nuc_code[i] = nuc_code[i];
    __syncthreads();
    u64 g = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u64 b0 = g * 32;
    if (b0 >= n) return;
    u32 out[4] = {0, 0, 0, 0}, mbits = 0, inv = 0;
    if (b0 + 32 <= n) {
        const uint4 *src = (const uint4 *)(bases + b0);
        uint4 v0 = src[0], v1 = src[1];
        u32 w[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
        for (int k = 0; k < 8; k++) {
            u32 x = w[k];
            u32 c0 = c_nuc_code[x & 0xFF], c1 = c_nuc_code[(x >> 8) & 0xFF], c2 = c_nuc_code[(x >> 16) & 0xFF], c3 = c_nuc_code[x >> 24];
            inv |= c0 | c1 | c2 | c3;
            u32 two = (c0 & 15) | ((c1 & 15) << 4) | ((c2 & 15) << 8) | ((c3 & 15) << 12);
            out[k >> 1] |= two << (16 * (k & 1));
            // case bit: byte >= 96
            u32 ge = ((x & 0xFF) >= 96) | ((((x >> 8) & 0xFF) >= 96) << 1) | ((((x >> 16) & 0xFF) >= 96) << 2) | (((x >> 24) >= 96) << 3);
            mbits |= ge << (4 * k);
        }
        *(uint4 *)(packed + (b0 >> 1)) = make_uint4(out[0], out[1], out[2], out[3]);
    } else {
        for (u64 i = b0; i < n; i += 2) {
            u32 c0 = c_nuc_code[bases[i]], c1 = i + 1 < n ? c_nuc_code[bases[i + 1]] : 0;
            inv |= c0 | c1;
            packed[i >> 1] = (u8)((c0 & 15) | ((c1 & 15) << 4));
        }
        for (u64 i = b0; i < n; i++) if (bases[i] >= 96) mbits |= 1u << (i - b0);
    }
    if (want_mask) casebits[g] = mbits;
    if (inv & 0x80) atomicOr(flag, (u32)FF_SEQ);
}

// flips[w] = positions where the case differs from the previous base (case before base 0 = prev0: unmasked for a whole
// file, encoders.c:126; for a shard of a file, whatever keeps position 0 from counting as a flip)
__global__ void k_flip_count(const u32 *casebits, u64 nwords, u64 *tile_counts, u32 prev0)
{
    __shared__ u64 sm[33];
    u64 w = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u32 c = 0;
    if (w < nwords) { u32 cur = casebits[w], prev = w ? casebits[w - 1] >> 31 : prev0; c = __popc(cur ^ ((cur << 1) | prev)); }
    u64 tot; block_excl_scan(c, &tot, sm);
    if (threadIdx.x == 0) tile_counts[blockIdx.x] = tot;
}
__global__ void k_flip_scatter(const u32 *casebits, u64 nwords, const u64 *tile_prefix, u64 *flip_pos, u32 prev0)
{
    __shared__ u64 sm[33];
    u64 w = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u32 fl = 0;
    if (w < nwords) { u32 cur = casebits[w], prev = w ? casebits[w - 1] >> 31 : prev0; fl = cur ^ ((cur << 1) | prev); }
    u64 tot; u64 r = block_excl_scan(__popc(fl), &tot, sm) + tile_prefix[blockIdx.x];
    while (fl) { int k = __ffs(fl) - 1; fl &= fl - 1; flip_pos[r++] = (w << 5) + k; }
}

// runs of equal case: run k = [flip[k-1], flip[k]) (run 0 starts at 0, run R ends at n).  casebits beyond n are
// zero, so a masked tail produces one spurious flip at n; clamping to n makes that final run empty.
// For one shard of a file: `bnd` = a flip sits exactly at my position 0 (my first base differs in case from the last
// base before me), `carry` = bases since the last flip in the shards before me (they lengthen my first run),
// `emit_final` = the run after my last flip is written by me (last shard) and not carried into the next shard.
struct RunCalc {
    const u64 *fp; u64 R, nb; u64 carry; u32 bnd, emit_final;      // R counts the boundary flip
    __device__ u64 flip(size_t k) const { if (bnd) return k ? fp[k - 1] : 0; return fp[k]; }
    __device__ u64 len(size_t k) const
    {
        u64 s = k ? flip(k - 1) : 0, e = k < R ? flip(k) : nb;
        if (s > nb) s = nb;
        if (e > nb) e = nb;
        return e - s + (k == 0 ? carry : 0);
    }
    __device__ u64 units(size_t k) const                      // encoders.c:98 add_mask: L/255 bytes of 255, then L%255
    {
        u64 L = len(k);
        if (k == R && (L == 0 || !emit_final)) return 0;        // final run only if > 0 (ennaf.c:511)
        return L / 255 + 1;
    }
};

struct SplitDev;
// case flips -> runs -> mask units (encoders.c:98-151; final run flushed by ennaf.c:511)
static void build_mask_units(CudaExec &ex, const u64 *flip_pos, u64 R_interior, u64 n_seq, u32 bnd, u64 carry, u32 emit_final, u8 **mask, u64 *n_mask)
{
    RunCalc rc{flip_pos, R_interior + bnd, n_seq, carry, bnd, emit_final};
    const u64 R = rc.R;
    u64 *upre = ex.alloc<u64>(R + 3);
    exclusive_scan(ex, [rc] __device__ (size_t k) { return rc.units(k); }, R + 1, upre);
    u64 n_units; ex.download(&n_units, upre + R + 1, 8);
    *n_mask = n_units;
    *mask = ex.alloc<u8>(n_units + 64);
    ex.fill(*mask, 0xFF, n_units);
    u8 *mk = *mask;
    ex.for_each(R + 1, [=] __device__ (size_t k) { u64 u = rc.units(k); if (u) mk[upre[k] + u - 1] = (u8)(rc.len(k) % 255); }, "mask_units");
}

// ------------------------------------------------------------------ split orchestration

struct SplitDev {
    u8 *ids, *comm, *len, *mask, *seq, *qual;
    u64 n_ids, n_comm, n_len, n_mask, n_seq, n_qual;     // bytes
    u64 n_bases, n_records, longest;
    int format, store_mask, store_qual;
    // shard mode (nafgpu_shard_begin): the mask is not built yet; what the link step needs instead
    const u64 *flip_pos; u64 n_flips; u32 first_case, last_case, first_code;
};

static void die_input(const std::string &m) { fail(NAFGPU_E_INPUT, m); }

struct FastFallback {};      // thrown inside split_streams_impl when the canonical-input parser meets input it does not cover

static SplitDev split_streams_impl(Ctx &ctx, CudaExec &ex, const u8 *d_text, size_t n, const nafgpu_enc_opts &o, nafgpu_enc_info *info, bool use_fast,
                                   bool shard = false)
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
    C.nstates = C.fastq ? (int)FQ_NSTATES : (int)FA_NSTATES;
    C.text_fasta = o.seq_type == NAFGPU_TEXT && !C.fastq;

    FsmTables ht; build_tables(C, ht);
    FsmTables *d_tab = ex.alloc<FsmTables>(1);
    ex.upload(d_tab, &ht, sizeof ht);

    const u64 ntiles = (n + PTILE - 1) / PTILE;               // tiles cover the text from offset 0; bytes before p0 are skipped
    ParseArgs P; memset(&P, 0, sizeof P);
    P.text = d_text; P.n = n; P.p0 = p0; P.cfg = C; P.tab = d_tab; P.ntiles = ntiles;
    if (!use_fast) { P.thread_map = ex.alloc<u64>(ntiles * PT + 1); P.tile_map = ex.alloc<u64>(ntiles + 1); }
    P.tile_state = ex.alloc<u8>(ntiles + 2);
    P.tinfo = ex.alloc<ThreadInfo>(ntiles * PT + 1); P.tile = ex.alloc<TileCounts>(ntiles + 1);
    // canonical-input parser (naf_parse_fast.cuh): same passes, same records; raises *flag when the input needs the general one
    FastArgs F; memset(&F, 0, sizeof F);
    u32 *d_flag = ex.alloc<u32>(1);
    ex.zero(d_flag, 4);
    auto check_fast_flag = [&]() { u32 f; ex.download(&f, d_flag, 4); if (f) throw FastFallback{}; };
    if (use_fast) {
        F.tile_elem = ex.alloc<u32>(ntiles + 1); F.tile_entry = ex.alloc<u32>(ntiles + 2); F.flag = d_flag;
        F.upper = o.seq_type >= NAFGPU_PROTEIN && o.no_mask;
        F.seq_check = o.seq_type == NAFGPU_PROTEIN ? 1 : (o.seq_type == NAFGPU_TEXT ? (C.text_fasta ? 3 : 2) : 0);
        F.P = P;
        if (C.fastq) {
            if (ntiles) { KLAUNCH(ex, "k_fast_tiles", k_fast_tiles<true><<<(unsigned)ntiles, PT, 0, ex.stream>>>(F)); }
            KLAUNCH(ex, "k_fast_scan", k_fast_scan<true><<<1, 1024, 0, ex.stream>>>(F));
            if (ntiles) { KLAUNCH(ex, "k_fast_count", k_fast_count<true><<<(unsigned)ntiles, PT, 0, ex.stream>>>(F)); }
        } else {
            if (ntiles) { KLAUNCH(ex, "k_fast_tiles", k_fast_tiles<false><<<(unsigned)ntiles, PT, 0, ex.stream>>>(F)); }
            KLAUNCH(ex, "k_fast_scan", k_fast_scan<false><<<1, 1024, 0, ex.stream>>>(F));
            if (ntiles) { KLAUNCH(ex, "k_fast_count", k_fast_count<false><<<(unsigned)ntiles, PT, 0, ex.stream>>>(F)); }
        }
    } else {
        if (ntiles) { KLAUNCH(ex, "k_fsm_reduce", k_fsm_reduce<<<(unsigned)ntiles, PT, 0, ex.stream>>>(P)); }
        KLAUNCH(ex, "k_fsm_scan", k_fsm_scan<<<1, 1024, 0, ex.stream>>>(P));
        if (ntiles) { KLAUNCH(ex, "k_fsm_count", k_fsm_count<<<(unsigned)ntiles, PT, 0, ex.stream>>>(P)); }
    }
    // exclusive sums of the six counters + exclusive max of the line-end marker
    u64 *pre[7];
    for (int k = 0; k < 7; k++) pre[k] = ex.alloc<u64>(ntiles + 2);
    const TileCounts *tc = P.tile;
    exclusive_scan(ex, [tc] __device__ (size_t i) { return tc[i].ids; }, ntiles, pre[0]);
    exclusive_scan(ex, [tc] __device__ (size_t i) { return tc[i].comm; }, ntiles, pre[1]);
    exclusive_scan(ex, [tc] __device__ (size_t i) { return tc[i].seq; }, ntiles, pre[2]);
    exclusive_scan(ex, [tc] __device__ (size_t i) { return tc[i].seq_counted; }, ntiles, pre[3]);
    exclusive_scan(ex, [tc] __device__ (size_t i) { return tc[i].qual; }, ntiles, pre[4]);
    exclusive_scan(ex, [tc] __device__ (size_t i) { return tc[i].rec; }, ntiles, pre[5]);
    // pre_line[t] = counted-seq value at the last line end in tiles < t (0 if none): chunked exclusive max-scan
    {
        u64 *pl = pre[6]; const u64 *pc = pre[3]; const u64 nt = ntiles;
        const u64 nchunks = 1024, per = (nt + nchunks - 1) / nchunks;
        u64 *cmax = ex.alloc<u64>(nchunks + 1);
        ex.for_each(nchunks, [=] __device__ (size_t c) {
            u64 lo = c * per, hi = lo + per; if (lo > nt) lo = nt; if (hi > nt) hi = nt;
            u64 m = 0; for (u64 t = lo; t < hi; t++) if (tc[t].has_line) { u64 v = pc[t] + tc[t].line_last; if (v > m) m = v; }
            cmax[c] = m; }, "line_scan");
        ex.for_each(1, [=] __device__ (size_t) { u64 m = 0; for (u64 c = 0; c < nchunks; c++) { u64 v = cmax[c]; cmax[c] = m; if (v > m) m = v; } }, "line_scan");
        ex.for_each(nchunks, [=] __device__ (size_t c) {
            u64 lo = c * per, hi = lo + per; if (lo > nt) lo = nt; if (hi > nt) hi = nt;
            u64 m = cmax[c];
            for (u64 t = lo; t < hi; t++) { pl[t] = m; if (tc[t].has_line) { u64 v = pc[t] + tc[t].line_last; if (v > m) m = v; } } }, "line_scan");
    }
    u64 tot[6]; u8 end_state;
    for (int k = 0; k < 6; k++) ex.download(&tot[k], pre[k] + ntiles, 8);
    ex.download(&end_state, P.tile_state + ntiles, 1);
    if (ntiles == 0) end_state = 0;
    if (use_fast) check_fast_flag();

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
    P.pre_ids = pre[0]; P.pre_comm = pre[1]; P.pre_seq = pre[2]; P.pre_cnt = pre[3]; P.pre_qual = pre[4]; P.pre_rec = pre[5]; P.pre_line = pre[6];
    P.ids = S.ids; P.comm = S.comm; P.bases = bases; P.qual = S.qual;
    P.rec_seq_end = rec_seq_end; P.rec_qual_end = rec_qual_end; P.rec_pos = rec_pos;
    P.unexpected = d_unexp; P.longest = d_longest; P.first_bad = d_first_bad;
    if (use_fast) {
        F.P = P;
        if (C.fastq) {
            CUDA_TRY(cudaFuncSetAttribute(k_fast_scatter<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FAST_SCATTER_SMEM));
            if (ntiles) { KLAUNCH(ex, "k_fast_scatter", k_fast_scatter<true><<<(unsigned)ntiles, PT, FAST_SCATTER_SMEM, ex.stream>>>(F)); }
        } else {
            CUDA_TRY(cudaFuncSetAttribute(k_fast_scatter<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FAST_SCATTER_SMEM));
            if (ntiles) { KLAUNCH(ex, "k_fast_scatter", k_fast_scatter<false><<<(unsigned)ntiles, PT, FAST_SCATTER_SMEM, ex.stream>>>(F)); }
        }
    } else {
        CUDA_TRY(cudaFuncSetAttribute(k_fsm_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SCATTER_SMEM));
        if (ntiles) { KLAUNCH(ex, "k_fsm_scatter", k_fsm_scatter<<<(unsigned)ntiles, PT, SCATTER_SMEM, ex.stream>>>(P)); }
    }
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
    if (use_fast) check_fast_flag();
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
        }, "qual_len_check");
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
        }, "length_units");
        // longest line: FASTA tracked line ends; FASTQ = longest read (process.c:495)
        if (C.fastq) {
            unsigned long long *dl = d_longest;
            ex.for_each(n_rec, [=] __device__ (size_t i) {
                u64 L = rse[i] - (i ? rse[i - 1] : 0);
                // reads are identical for most inputs: skip the atomic unless this record beats the current maximum
                if (L > *(volatile unsigned long long *)dl) atomicMax(dl, (unsigned long long)L);
            }, "longest_read");
        }
        unsigned long long lg; ex.download(&lg, d_longest, 8);
        S.longest = lg;
    }
    S.n_ids = n_ids; S.n_comm = n_comm; S.n_qual = n_qual;

    // ---- sequence stream: 4-bit pack (+ mask) for DNA/RNA, bytes as they are for protein/text
    if (o.seq_type < NAFGPU_PROTEIN) {
        const u8 *d_lut = ctx.d_nuc_lut + (o.seq_type == NAFGPU_RNA ? 256 : 0);
        S.n_seq = (n_seq + 1) / 2;
        S.seq = ex.alloc<u8>(S.n_seq + 64);
        u64 nwords = (n_seq + 31) / 32;
        u32 *casebits = ex.alloc<u32>(nwords + 2);
        if (nwords) { KLAUNCH(ex, "k_pack4", k_pack4<<<(unsigned)((nwords + 255) / 256), 256, 0, ex.stream>>>(bases, n_seq, S.seq, casebits, S.store_mask, d_lut, d_flag)); }
        if (use_fast) check_fast_flag();
        if (shard && n_seq) {
            u8 fb; ex.download(&fb, S.seq, 1);
            S.first_code = fb & 15;
        }
        if (S.store_mask && n_seq) {
            u32 prev0 = 0;
            if (shard) {                                                   // no flip at my position 0: the link step decides about that one
                u32 cw[2]; ex.download(&cw[0], casebits, 4); ex.download(&cw[1], casebits + (n_seq - 1) / 32, 4);
                S.first_case = cw[0] & 1; S.last_case = (cw[1] >> ((n_seq - 1) & 31)) & 1;
                prev0 = S.first_case;
            }
            u64 ft = (nwords + 255) / 256;
            u64 *fcount = ex.alloc<u64>(ft + 1), *fpre = ex.alloc<u64>(ft + 2);
            KLAUNCH(ex, "k_flip_count", k_flip_count<<<(unsigned)ft, 256, 0, ex.stream>>>(casebits, nwords, fcount, prev0));
            const u64 *fc = fcount;
            exclusive_scan(ex, [fc] __device__ (size_t i) { return fc[i]; }, ft, fpre);
            u64 R; ex.download(&R, fpre + ft, 8);                          // number of flips; runs = R + 1
            u64 *flip_pos = ex.alloc<u64>(R + 2);
            KLAUNCH(ex, "k_flip_scatter", k_flip_scatter<<<(unsigned)ft, 256, 0, ex.stream>>>(casebits, nwords, fpre, flip_pos, prev0));
            if (shard) {
                // the spurious flip at n_seq after a masked tail is not a flip of mine
                if (R) { u64 lastf; ex.download(&lastf, flip_pos + R - 1, 8); if (lastf >= n_seq) R--; }
                S.flip_pos = flip_pos; S.n_flips = R;
            } else build_mask_units(ex, flip_pos, R, n_seq, 0, 0, 1, &S.mask, &S.n_mask);
        }
    } else {
        S.seq = bases; S.n_seq = n_seq;
    }
    if (!S.len) S.len = ex.alloc<u8>(64);
    if (!S.mask) S.mask = ex.alloc<u8>(64);
    if (info) { info->n_sequences = n_rec; info->longest_line = S.longest; info->n_bases = n_seq; }
    return S;
}

// confirm_input_format (process.c:547): skip leading white space, the first byte decides.  -> format (0: empty input), *p0 =
// offset of the first byte after the leading '>' / '@'.  `head` = the first min(n, 64 KB) bytes of the text, on the host.
static int confirm_format(const u8 *h, size_t head, size_t n, const nafgpu_enc_opts &o, u64 *p0)
{
    auto is_space = [](int ch) { return (ch >= 0x09 && ch <= 0x0D) || ch == 0x20; };
    int fmt = 0; *p0 = 0;
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
        *p0 = i + 1;
    }
    if (o.input_format != NAFGPU_FMT_AUTO && fmt && o.input_format != fmt) die_input("input format is different from format specified in the command line\n");
    return fmt;
}

// one case bit per base -> flips -> (whole file) mask units / (shard) what the link step needs
static void mask_from_casebits(CudaExec &ex, SplitDev &S, const u32 *casebits, u64 n_seq, bool shard)
{
    const u64 nwords = (n_seq + 31) / 32;
    u32 prev0 = 0;
    if (shard) {                                                   // no flip at my position 0: the link step decides about that one
        u32 cw[2]; ex.download(&cw[0], casebits, 4); ex.download(&cw[1], casebits + (n_seq - 1) / 32, 4);
        S.first_case = cw[0] & 1; S.last_case = (cw[1] >> ((n_seq - 1) & 31)) & 1;
        prev0 = S.first_case;
    }
    u64 ft = (nwords + 255) / 256;
    u64 *fcount = ex.alloc<u64>(ft + 1), *fpre = ex.alloc<u64>(ft + 2);
    KLAUNCH(ex, "k_flip_count", k_flip_count<<<(unsigned)ft, 256, 0, ex.stream>>>(casebits, nwords, fcount, prev0));
    const u64 *fc = fcount;
    exclusive_scan(ex, [fc] __device__ (size_t i) { return fc[i]; }, ft, fpre);
    u64 R; ex.download(&R, fpre + ft, 8);                          // number of flips; runs = R + 1
    u64 *flip_pos = ex.alloc<u64>(R + 2);
    if (R) { KLAUNCH(ex, "k_flip_scatter", k_flip_scatter<<<(unsigned)ft, 256, 0, ex.stream>>>(casebits, nwords, fpre, flip_pos, prev0)); }
    if (shard) {
        // the spurious flip at n_seq after a masked tail is not a flip of mine
        if (R) { u64 lastf; ex.download(&lastf, flip_pos + R - 1, 8); if (lastf >= n_seq) R--; }
        S.flip_pos = flip_pos; S.n_flips = R;
    } else build_mask_units(ex, flip_pos, R, n_seq, 0, 0, 1, &S.mask, &S.n_mask);
}

// zstd_enc.cu (same translation unit, below): compression of the big streams behind the upload
static void zenc_early_begin(Ctx &ctx, CudaExec &ex, const u8 *seq, u64 seq_max_bytes, const u8 *qual, u64 qual_max_bytes);
static void zenc_early_step(Ctx &ctx, u64 seq_bytes, u64 qual_bytes);
static void zenc_early_abort(Ctx &ctx);
struct EarlyTotals { u64 seq, qual; };
__global__ void k_early_totals(const ulonglong2 *rec, EarlyTotals *out)      // the inclusive look-back #2 record of a chunk's last tile
{
    F2 g; f2_get(rec, g);
    out->seq = g.seq; out->qual = g.qual;
}

// The single-pass transform (naf_fused.cuh) for canonical input: one kernel reads the text once and writes every stream once.
// Throws FastFallback when the input is not canonical (or has an error the general parser must word).
// h_head: the first bytes of the text on the host if the caller has them (host-buffer API), else nullptr.
static SplitDev split_streams_fused(Ctx &ctx, CudaExec &ex, const u8 *d_text, size_t n, const nafgpu_enc_opts &o, nafgpu_enc_info *info, bool shard,
                                    const u8 *h_head = nullptr)
{
    SplitDev S; memset(&S, 0, sizeof S);
    if (info) memset(info, 0, sizeof *info);
    S.store_mask = !(o.no_mask || o.seq_type >= NAFGPU_PROTEIN);                                 // ennaf.c:445
    u64 p0 = 0;
    const size_t head = n < 65536 ? n : 65536;
    const bool piped = ex.pipe && ex.pipe->uploading;           // host-buffer call: the text is still on its way up, chunk by chunk
    if (piped) h_head = ex.pipe->h_in;
    if (!h_head) {
        ctx.host_scratch.resize(head + 1);
        if (head) ex.download(ctx.host_scratch.data(), d_text, head);
        h_head = ctx.host_scratch.data();
    }
    const int fmt = confirm_format(h_head, head, n, o, &p0);
    S.format = fmt; S.store_qual = fmt == NAFGPU_FMT_FASTQ;
    if (info) info->format = fmt;
    if (fmt == 0) {
        S.len = ex.alloc<u8>(64); S.mask = ex.alloc<u8>(64); S.ids = ex.alloc<u8>(64); S.comm = ex.alloc<u8>(64); S.seq = ex.alloc<u8>(64); S.qual = ex.alloc<u8>(64);
        return S;
    }
    const bool fastq = fmt == NAFGPU_FMT_FASTQ, packed = o.seq_type < NAFGPU_PROTEIN;
    const u64 ntiles = (n + FT_BYTES - 1) / FT_BYTES;
    if (ntiles >= (1ull << 31)) throw FastFallback{};

    FusedArgs A; memset(&A, 0, sizeof A);
    FusedCfg &C = A.C;
    C.n = n; C.p0 = p0; C.fastq = fastq;
    C.seq_mode = o.seq_type == NAFGPU_PROTEIN ? FS_PROTEIN : (o.seq_type == NAFGPU_TEXT ? (fastq ? FS_TEXT : FS_TEXT_GT) : FS_PACK4);
    C.upper = o.seq_type >= NAFGPU_PROTEIN && o.no_mask;
    C.want_mask = S.store_mask;
    C.id_check = (o.seq_type == NAFGPU_TEXT && !fastq) ? FC_ID_GT : FC_ID;
    // destinations at their worst-case sizes (the arena is one slab: untouched bytes cost nothing)
    C.ids = ex.alloc<u8>(n + 64); C.comm = ex.alloc<u8>(n + 64);
    C.qual = ex.alloc<u8>(fastq ? n + 64 : 64);
    C.seq = ex.alloc<u8>((packed ? n / 2 : n) + 64);
    C.len = ex.alloc<u32>(ntiles * FT_MAXSEG + 16);
    const u64 cb_words = packed && S.store_mask ? n / 32 + 2 : 1;
    C.casebits = ex.alloc<u32>(cb_words);
    if (packed && S.store_mask) ex.zero(C.casebits, cb_words * 4);
    A.lut8 = ctx.d_nuc_lut + (o.seq_type == NAFGPU_RNA ? 256 : 0);
    // look-back records + scalars
    A.text = d_text; A.ntiles = (u32)ntiles;
    A.st1 = ex.alloc<u64>(ntiles + 1); A.st2 = ex.alloc<ulonglong2>(4 * ntiles + 4);
    ex.zero(A.st1, (ntiles + 1) * 8); ex.zero(A.st2, (4 * ntiles + 4) * 16);
    u64 *scal = ex.alloc<u64>(4 + (sizeof(FusedTotals) + 7) / 8);
    ex.zero(scal, 32 + sizeof(FusedTotals));
    A.ticket = (u32 *)scal; A.flag = (u32 *)scal + 1; A.longest = (unsigned long long *)(scal + 1); A.totals = (FusedTotals *)(scal + 4);

    CUDA_TRY(cudaFuncSetAttribute(k_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FusedSmem::total));
    if (piped) {
        // one launch per uploaded chunk, each right behind its chunk's copy: tiles are handed out by the global ticket and look
        // back only (naf_fused.cuh), so a launch needs nothing beyond the bytes that have arrived.  Every EARLY_GROUP chunks the
        // stream sizes so far go to the host (mailbox + event), which then has the blocks that became complete compressed on the
        // side stream while the next chunks are still on their way up.
        static const bool env_early = !(getenv("NAFGPU_EARLY") && getenv("NAFGPU_EARLY")[0] == '0');
        const size_t EARLY_GROUP = 8, nchunks = ex.pipe->chunks();
        const bool early = env_early && !shard && ctx.side && ex.mail && !(ex.prof && ex.prof->on) && nchunks >= 2 * EARLY_GROUP && nchunks / EARLY_GROUP <= 240;
        std::vector<cudaEvent_t> gev;
        EarlyTotals *slots = early ? (EarlyTotals *)(ex.mail->p + (48u << 10)) : nullptr;
        if (early) zenc_early_begin(ctx, ex, C.seq, (packed ? n / 2 : n) + 64, C.qual, fastq ? n + 64 : 0);
        u64 t0 = 0;
        for (size_t c = 0; c < nchunks; c++) {
            const u64 hi = (c + 1) * ex.pipe->chunk < n ? (c + 1) * ex.pipe->chunk : n;
            // a tile also reads the byte behind it (a CR LF pair may straddle two tiles): the last tile of a chunk waits for the
            // next chunk
            const u64 t1 = c + 1 == nchunks ? ntiles : (hi - 1) / FT_BYTES;
            ex.pipe->wait_input(ex.stream, hi);
            if (t1 > t0) { KLAUNCH(ex, "k_fused", k_fused<<<(unsigned)(t1 - t0), FUSED_NT, FusedSmem::total, ex.stream>>>(A)); }
            t0 = t1;
            if (early && (c + 1) % EARLY_GROUP == 0 && c + 1 < nchunks && t1 > 0) {
                k_early_totals<<<1, 1, 0, ex.stream>>>(A.st2 + 4ull * (t1 - 1), slots + gev.size());
                cudaEvent_t e = ex.pipe->event();
                CUDA_TRY(cudaEventRecord(e, ex.stream));
                gev.push_back(e);
            }
        }
        for (size_t g = 0; g < gev.size(); g++) {
            CUDA_TRY(cudaEventSynchronize(gev[g]));
            const EarlyTotals t = slots[g];
            zenc_early_step(ctx, packed ? t.seq / 2 : t.seq, t.qual);      // (a byte of codes still waiting for its second base is not final)
        }
    } else { KLAUNCH(ex, "k_fused", k_fused<<<(unsigned)ntiles, FUSED_NT, FusedSmem::total, ex.stream>>>(A)); }
    KLAUNCH(ex, "k_fused_finish", k_fused_finish<<<1, 1, 0, ex.stream>>>(A));
    FusedTotals tot; ex.download(&tot, A.totals, sizeof tot);
    if (tot.flag) throw FastFallback{};

    S.ids = C.ids; S.comm = C.comm; S.qual = C.qual; S.len = (u8 *)C.len;
    S.n_ids = tot.n_ids; S.n_comm = tot.n_comm; S.n_qual = tot.n_qual; S.n_len = tot.n_rec * 4;
    S.n_records = tot.n_rec; S.n_bases = tot.n_bases; S.longest = tot.longest;
    S.seq = C.seq; S.n_seq = packed ? (tot.n_bases + 1) / 2 : tot.n_bases;
    if (packed) {
        if (shard && tot.n_bases) { u8 fb; ex.download(&fb, S.seq, 1); S.first_code = fb & 15; }
        if (S.store_mask && tot.n_bases) mask_from_casebits(ex, S, C.casebits, tot.n_bases, shard);
    }
    if (!S.mask) S.mask = ex.alloc<u8>(64);
    if (info) { info->n_sequences = tot.n_rec; info->longest_line = S.longest; info->n_bases = tot.n_bases; }
    return S;
}

// Canonical input goes through the fast parser; anything it does not cover (or any input error, so that the message
// comes from the exact restatement) is redone by the general one.  --well-formed has its own tables: general only.
static SplitDev split_streams(Ctx &ctx, CudaExec &ex, const u8 *d_text, size_t n, const nafgpu_enc_opts &o, nafgpu_enc_info *info)
{
    static const bool env_general = getenv("NAFGPU_GENERAL_PARSER") != nullptr;
    if (!o.well_formed && !env_general && !o.general_parser) {
        const Arena::Mark mk = ex.arena->mark();
        try { return split_streams_fused(ctx, ex, d_text, n, o, info, false); }
        catch (const FastFallback &) {}
        catch (const NafError &) {}
        zenc_early_abort(ctx);
        CUDA_TRY(cudaStreamSynchronize(ex.stream));
        ex.arena->rewind(mk);
        ctx.fast_fallbacks++;
    }
    if (ex.pipe) ex.pipe->wait_all_input(ex.stream);           // the general parser's passes each read the whole text
    return split_streams_impl(ctx, ex, d_text, n, o, info, false);
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

// ennaf -# (ennaf.c:222-223, compressor.c:7).  Level 1 -- the tools' default -- is the fastest parse: every stream entropy-coded
// in independent 32 KB blocks.  Level >= 2 adds LZ77 matches and FSE-coded sequences to the text-like streams (ids, comments,
// lengths) in 8 KB blocks: the file shrinks to what `ennaf -1` writes (0.356 of the text on 150 bp FASTQ against 0.381; ennaf -1:
// 0.358).  Two formulations (zstd_enc.cu): the data-parallel one (column match finder, one Huffman code and one set of FSE tables
// per stream: + 2.1 ms encode, + 3.8 ms decode per million reads on a B200, profiles/r4b) is what a level selects; NAFGPU_LZ=1 in
// the environment selects the first one (one thread per block, private tables: about twice that), NAFGPU_LZ=0 none, NAFGPU_LZ=s
// the data-parallel one at any level.  The mask stream stays entropy-only in the data-parallel formulation: its run lengths
// have no matches to find, and 8 KB blocks coded with a stream-wide Huffman code come out larger than 32 KB blocks with their own.
static bool lz_for_level(int level)
{
    const char *env = getenv("NAFGPU_LZ");
    if (env && (env[0] == '0' || env[0] == '1' || env[0] == 's' || env[0] == 'b')) return env[0] != '0';
    return level >= 2;
}
static int lz_streams() { return zlc_mode() ? 3 : 4; }           // how many of ids, comments, lengths, mask take the LZ stage

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
    const bool lz = lz_for_level(o.level);                     // ids, comments, lengths, mask: LZ77 + FSE-coded sequences
    for (int k = 0; k < 6; k++) if (present[k]) { which[ns++] = k; batch.add(sp[k], ss[k], k == 4 ? o.window_log : 0, lz && k < lz_streams()); }
    zstd_compress_batch(ctx, ex, batch);                       // sizes known on the host afterwards

    // ---- block index: the compressed size of every block of the sequence / quality frames, as a skippable frame behind the
    // lengths frame (a section is read with ZSTD_decompress, unnaf/src/input.c:211, which skips such frames:
    // zstd/lib/decompress/zstd_decompress.c ZSTD_decompressMultiFrame).  Our decoder then finds every block without walking
    // the chain of headers (zstd_dec.cuh: zstd_walk_indexed) -- which is what lets N GPUs start on their parts at once.
    struct IdxStream { u32 section, nblk, regen, first_block; u64 total; };
    std::vector<IdxStream> idx;
    static const bool env_index = !(getenv("NAFGPU_INDEX") && getenv("NAFGPU_INDEX")[0] == '0');
    if (env_index && !o.no_block_index)
        for (int j = 0; j < ns; j++) {
            const int k = which[j];
            const u32 nblk = batch.first_block[j + 1] - batch.first_block[j];
            if ((k == 4 || k == 5) && !batch.lz[j] && nblk >= 64) idx.push_back(IdxStream{(u32)k, nblk, ZBS, batch.first_block[j], ss[k]});
        }
    std::vector<u8> idx_head;
    u64 idx_bytes = 0;
    if (!idx.empty()) {
        u64 payload = 12 + 24 * idx.size();
        for (auto &e : idx) payload += 2ull * e.nblk;
        auto put32 = [&](u32 v) { for (int b = 0; b < 4; b++) idx_head.push_back((u8)(v >> (8 * b))); };
        put32(0x184D2A5Eu); put32((u32)payload);
        for (const char *t = "NAFGIDX1"; *t; t++) idx_head.push_back((u8)*t);
        put32((u32)idx.size());
        for (auto &e : idx) { put32(e.section); put32(e.nblk); put32(e.regen); put32(0); put32((u32)e.total); put32((u32)(e.total >> 32)); }
        idx_bytes = 8 + payload;
    }

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
    u64 total = hdr.size(), idx_at = 0;
    std::vector<u64> payload_at(ns);
    for (int j = 0; j < ns; j++) {
        int k = which[j];
        u64 csz = batch.frame_size[j] - 4;                     // magic stripped (compressor.c:158)
        if (k == 2 && idx_bytes) csz += idx_bytes;
        nafc::put_vle(sec_hdr[j], orig[k]); nafc::put_vle(sec_hdr[j], csz);
        total += sec_hdr[j].size();
        payload_at[j] = total; total += csz;
        if (k == 2 && idx_bytes) idx_at = payload_at[j] + batch.frame_size[j] - 4;
        info->stream_comp[k] = csz; info->stream_raw[k] = ss[k];
    }
    u8 *d_naf = ex.alloc<u8>(total + 64);
    // upload the header and the tiny per-section headers; gather the frames next to them
    ex.upload(d_naf, hdr.data(), hdr.size());
    for (int j = 0; j < ns; j++) ex.upload(d_naf + payload_at[j] - sec_hdr[j].size(), sec_hdr[j].data(), sec_hdr[j].size());
    if (idx_bytes) {
        ex.upload(d_naf + idx_at, idx_head.data(), idx_head.size());
        u64 at = idx_at + idx_head.size();
        for (auto &e : idx) {
            const ZEncBlock *blk = batch.d_blocks + e.first_block; u8 *dst = d_naf + at;
            ex.for_each(e.nblk, [=] __device__ (size_t i) { const u32 c = blk[i].csize; dst[2 * i] = (u8)c; dst[2 * i + 1] = (u8)(c >> 8); }, "zenc_block_index");
            at += 2ull * e.nblk;
        }
    }
    CUDA_TRY(cudaStreamSynchronize(ex.stream));                // the small host vectors above must outlive the copies
    for (int j = 0; j < ns; j++) batch.dest[j] = d_naf + payload_at[j] - 4;   // frame byte i lands at dest + i; bytes 0..3 (magic) are skipped
    zstd_gather_frames(ctx, ex, batch, true);
    ex.check();
    return EncodeOut{d_naf, total};
}

EncodeOut zstd_compress_on_device(Ctx &ctx, CudaExec &ex, const u8 *d_src, size_t n, int window_log, int level)
{
    ZEncBatch batch;
    batch.add(d_src, n, window_log, lz_for_level(level));
    zstd_compress_batch(ctx, ex, batch);
    u8 *out = ex.alloc<u8>(batch.frame_size[0] + 64);
    batch.dest[0] = out;
    zstd_gather_frames(ctx, ex, batch, false);
    ex.check();
    return EncodeOut{out, batch.frame_size[0]};
}


// ------------------------------------------------------------------ record-aligned cut points (shards of one text)
// Where may a text be cut so that every piece starts a record?  FASTA: before a '>' at a line start.  FASTQ: after every 4th
// line counted from the top of the text (a quality line may begin with '@', so only counting is exact) -- 4-line records
// without blank lines, what the canonical-input transform accepts.  The newline ordinals come from a prefix sum over
// per-tile counts; each cut is then found by one CTA: the k-th newline of the text, or the first "\n>" at or after the target.
static const u32 CUT_TILE = 16384, CUT_NT = 256;
__global__ void __launch_bounds__(CUT_NT) k_cut_tiles(const u8 *text, u64 n, u64 *nl_count, u32 *first_gt)
{
    __shared__ u64 sm[33];
    __shared__ u32 s_first;
    const u64 lo = (u64)blockIdx.x * CUT_TILE;
    if (threadIdx.x == 0) s_first = 0xFFFFFFFFu;
    __syncthreads();
    u32 c = 0, first = 0xFFFFFFFFu;
    const u64 b0 = lo + (u64)threadIdx.x * 64;
    for (u32 i = 0; i < 64; i++) {
        const u64 p = b0 + i;
        if (p >= n) break;
        const u8 ch = text[p];
        if (ch == '\n') { c++; if (p + 1 < n && text[p + 1] == '>' && first == 0xFFFFFFFFu) first = (u32)(p - lo); }   // position of the '\n'
    }
    u64 tot; block_excl_scan(c, &tot, sm);
    if (first != 0xFFFFFFFFu) atomicMin(&s_first, first);
    __syncthreads();
    if (threadIdx.x == 0) { nl_count[blockIdx.x] = tot; first_gt[blockIdx.x] = s_first; }
}
// one CTA per cut k = 1 .. pieces - 1
__global__ void __launch_bounds__(CUT_NT) k_cut_find(const u8 *text, u64 n, int fastq, u32 pieces, u64 ntiles, const u64 *nl_pre, const u32 *first_gt, u64 *cuts)
{
    __shared__ u64 sm[33];
    __shared__ u64 s_val;
    const u32 k = blockIdx.x + 1, tid = threadIdx.x;
    const u64 t = (u64)((unsigned __int128)n * k / pieces);
    const u64 T = t / CUT_TILE;
    if (tid == 0) s_val = ~0ull;
    __syncthreads();
    if (!fastq) {
        // first "\n>" whose '\n' is at or after t: inside tile T by position, then the first later tile that has one
        const u64 lo = T * CUT_TILE;
        for (u32 i = 0; i < 64; i++) {
            const u64 p = lo + (u64)tid * 64 + i;
            if (p >= t && p + 1 < n && text[p] == '\n' && text[p + 1] == '>') { atomicMin((unsigned long long *)&s_val, (unsigned long long)p); break; }
        }
        __syncthreads();
        if (s_val == ~0ull) {
            for (u64 base = T + 1; base < ntiles; base += CUT_NT) {
                const u64 tile = base + tid;
                if (tile < ntiles && first_gt[tile] != 0xFFFFFFFFu) atomicMin((unsigned long long *)&s_val, (unsigned long long)(tile * CUT_TILE + first_gt[tile]));
                __syncthreads();
                if (s_val != ~0ull) break;
            }
        }
        __syncthreads();
        if (tid == 0) cuts[k] = s_val == ~0ull ? n : s_val + 1;
        return;
    }
    // FASTQ: newlines strictly before t -> ordinal of the first newline at or after t; the record end at or after it
    u32 c = 0;
    {
        const u64 lo = T * CUT_TILE + (u64)tid * 64;
        for (u32 i = 0; i < 64; i++) { const u64 p = lo + i; if (p < t && p < n && text[p] == '\n') c++; }
    }
    u64 tot; block_excl_scan(c, &tot, sm);
    const u64 before = nl_pre[T] + tot, total = nl_pre[ntiles];
    const u64 J = before + ((3 - (before & 3)) & 3);                       // smallest ordinal >= before with J % 4 == 3
    if (J >= total) { if (tid == 0) cuts[k] = n; return; }
    // tile holding newline J: nl_pre[tile] <= J < nl_pre[tile + 1]
    u64 a = T, b = ntiles;
    while (b - a > 1) { const u64 mid = (a + b) >> 1; if (nl_pre[mid] <= J) a = mid; else b = mid; }
    const u64 want = J - nl_pre[a];                                        // ordinal inside tile a
    u32 mine = 0;
    const u64 lo = a * CUT_TILE + (u64)tid * 64;
    for (u32 i = 0; i < 64; i++) { const u64 p = lo + i; if (p < n && text[p] == '\n') mine++; }
    u64 tot2; const u64 pre = block_excl_scan(mine, &tot2, sm);
    if (want >= pre && want < pre + mine) {
        u64 seen = pre;
        for (u32 i = 0; i < 64; i++) { const u64 p = lo + i; if (p < n && text[p] == '\n') { if (seen == want) { s_val = p; break; } seen++; } }
    }
    __syncthreads();
    if (tid == 0) cuts[k] = s_val == ~0ull ? n : s_val + 1;
}

void record_cuts_on_device(Ctx &ctx, CudaExec &ex, const u8 *d_text, size_t n, int pieces, uint64_t *cuts)
{
    cuts[0] = 0; cuts[pieces] = n;
    if (pieces <= 1 || n == 0) { for (int k = 1; k < pieces; k++) cuts[k] = n; return; }
    const size_t head = n < 65536 ? n : 65536;
    ctx.host_scratch.resize(head + 1);
    ex.download(ctx.host_scratch.data(), d_text, head);
    nafgpu_enc_opts o; memset(&o, 0, sizeof o);
    u64 p0 = 0;
    const int fmt = confirm_format(ctx.host_scratch.data(), head, n, o, &p0);
    if (fmt == 0) { for (int k = 1; k < pieces; k++) cuts[k] = n; return; }
    const u64 ntiles = (n + CUT_TILE - 1) / CUT_TILE;
    u64 *cnt = ex.alloc<u64>(ntiles + 1), *pre = ex.alloc<u64>(ntiles + 2), *d_cuts = ex.alloc<u64>(pieces + 1);
    u32 *fgt = ex.alloc<u32>(ntiles + 1);
    KLAUNCH(ex, "k_cut_tiles", k_cut_tiles<<<(unsigned)ntiles, CUT_NT, 0, ex.stream>>>(d_text, n, cnt, fgt));
    const u64 *c = cnt;
    exclusive_scan(ex, [c] __device__ (size_t i) { return c[i]; }, ntiles, pre);
    KLAUNCH(ex, "k_cut_find", k_cut_find<<<(unsigned)(pieces - 1), CUT_NT, 0, ex.stream>>>(d_text, n, fmt == NAFGPU_FMT_FASTQ, (u32)pieces, ntiles, pre, fgt, d_cuts));
    std::vector<u64> h(pieces + 1);
    ex.download(h.data() + 1, d_cuts + 1, (size_t)(pieces - 1) * 8);
    for (int k = 1; k < pieces; k++) cuts[k] = h[k] > cuts[k - 1] ? (h[k] < n ? h[k] : n) : cuts[k - 1];     // monotone
}

// ------------------------------------------------------------------ shards of one file (multi-GPU encode)

// packed stream of bases[1:] from the packed stream of bases[0:]: out[j] = in[j] >> 4 | in[j+1] << 4
__global__ void k_nibble_shift(const u8 *in, u64 n_out, u8 *out)
{
    const u64 g = (u64)blockIdx.x * blockDim.x + threadIdx.x, j0 = g * 8;
    if (j0 >= n_out) return;
    // 9 input bytes -> 8 output bytes; `in` is padded, bytes past its end are zero
    u64 lo = 0; u8 hi = in[j0 + 8];
    for (int k = 0; k < 8; k++) lo |= (u64)in[j0 + k] << (8 * k);
    const u64 v = (lo >> 4) | ((u64)hi << 60);
    for (int k = 0; k < 8 && j0 + k < n_out; k++) out[j0 + k] = (u8)(v >> (8 * k));
}

void shard_begin_on_device(Ctx &ctx, CudaExec &ex, const u8 *d_text, size_t n, const nafgpu_enc_opts &o, nafgpu_shard_counts *counts, nafgpu_enc_info *info)
{
    nafgpu_enc_info local; if (!info) info = &local;
    SplitDev S;
    bool done = false;
    static const bool env_general = getenv("NAFGPU_GENERAL_PARSER") != nullptr;
    if (!o.well_formed && !env_general && !o.general_parser) {
        const Arena::Mark mk = ex.arena->mark();
        try { S = split_streams_fused(ctx, ex, d_text, n, o, info, true); done = true; }
        catch (const FastFallback &) {}
        catch (const NafError &) {}
        if (!done) { CUDA_TRY(cudaStreamSynchronize(ex.stream)); ex.arena->rewind(mk); ctx.fast_fallbacks++; }
    }
    if (!done) S = split_streams_impl(ctx, ex, d_text, n, o, info, false, true);
    Ctx::Shard &H = ctx.shard;
    H = Ctx::Shard();
    H.opts = o; H.opts.title = nullptr;
    u8 *sp[6] = { S.ids, S.comm, S.len, S.mask, S.seq, S.qual };
    const u64 ss[6] = { S.n_ids, S.n_comm, S.n_len, 0, S.n_seq, S.n_qual };
    for (int k = 0; k < 6; k++) { H.stream[k] = sp[k]; H.raw[k] = ss[k]; }
    H.n_bases = S.n_bases; H.n_records = S.n_records; H.longest = S.longest; H.n_flips = S.n_flips; H.flip_pos = S.flip_pos;
    H.store_mask = S.store_mask; H.store_qual = S.store_qual; H.format = S.format; H.first_case = S.first_case;
    memset(counts, 0, sizeof *counts);
    counts->n_records = S.n_records; counts->n_bases = S.n_bases; counts->longest_line = S.longest;
    counts->n_flips = S.n_flips;
    if (S.n_flips) { u64 lf; ex.download(&lf, S.flip_pos + S.n_flips - 1, 8); counts->last_flip = lf; }
    counts->first_code = (u8)S.first_code; counts->first_case = (u8)S.first_case; counts->last_case = (u8)S.last_case;
    counts->format = (u8)S.format;
    H.active = true; H.finished = false;
}

void shard_finish_on_device(Ctx &ctx, CudaExec &ex, const nafgpu_shard_link &link, uint64_t raw[6], uint64_t body[6])
{
    Ctx::Shard &H = ctx.shard;
    const bool packed = H.opts.seq_type < NAFGPU_PROTEIN;
    // ---- 4-bit stream: global base index of my first base decides who owns the byte it shares with my predecessor
    if (packed && H.n_bases) {
        const bool odd_start = link.bases_before & 1;
        u64 my_bases = H.n_bases;
        if (odd_start) {                                       // my first base completes my predecessor's last byte
            my_bases = H.n_bases - 1;
            const u64 nb = (my_bases + 1) / 2;
            u8 *shifted = ex.alloc<u8>(nb + 64);
            ex.zero(shifted + nb, 64);
            if (nb) { KLAUNCH(ex, "k_nibble_shift", k_nibble_shift<<<(unsigned)((nb + 8 * 256 - 1) / (8 * 256)), 256, 0, ex.stream>>>(H.stream[4], nb, shifted)); }
            H.stream[4] = shifted; H.raw[4] = nb;
        }
        if ((my_bases & 1) && link.next_first_code) {          // my last byte's high nibble is my successor's first base
            u8 *p = H.stream[4] + H.raw[4] - 1; const u8 code = link.next_first_code;
            ex.for_each(1, [=] __device__ (size_t) { *p = (u8)((*p & 15) | (code << 4)); }, "shard_nibble");
        }
    }
    // ---- mask: runs that end inside this shard (+ the trailing run if this is the last shard)
    if (H.store_mask) {
        // a flip sits at my position 0 iff my first base differs in case from the last base before me
        const u32 bnd = H.n_bases ? (u32)(H.first_case != (u32)link.prev_last_case) : 0u;
        u8 *mask = nullptr; u64 n_mask = 0;
        build_mask_units(ex, H.flip_pos, H.n_flips, H.n_bases, bnd, link.run_carry, link.is_last ? 1u : 0u, &mask, &n_mask);
        H.stream[3] = mask; H.raw[3] = n_mask;
    }
    // ---- zstd blocks of every stream; only the last shard closes the frames
    ZEncBatch batch; batch.final_shard = link.is_last != 0;
    // (a shard that is empty, or FASTA-empty among FASTQ shards, still adds its -- empty -- block to every stream of the file)
    const bool present[6] = { true, true, true, (bool)H.store_mask, true, (bool)(H.store_qual || link.store_qual) };
    int which[6], ns = 0;
    const bool lz = lz_for_level(H.opts.level);
    for (int k = 0; k < 6; k++) if (present[k]) { which[ns++] = k; batch.add(H.stream[k], H.stream[k] ? H.raw[k] : 0, 0, lz && k < lz_streams()); }
    zstd_compress_batch(ctx, ex, batch);
    u64 total = 0; std::vector<u64> at(ns);
    for (int j = 0; j < ns; j++) { at[j] = total; total += (batch.frame_size[j] + 63) & ~63ull; }
    u8 *blob = ex.alloc<u8>(total + 64);
    for (int j = 0; j < ns; j++) batch.dest[j] = blob + at[j];
    zstd_gather_frames(ctx, ex, batch, false);
    for (int k = 0; k < 6; k++) { H.body[k] = nullptr; H.body_size[k] = 0; raw[k] = 0; body[k] = 0; }
    for (int j = 0; j < ns; j++) {
        const int k = which[j];
        H.body[k] = blob + at[j] + 6; H.body_size[k] = batch.frame_size[j] - 6;      // blocks only: magic, FHD and window byte belong to the merged frame
        raw[k] = H.raw[k]; body[k] = H.body_size[k];
    }
    ex.check();
    H.finished = true;
}

}  // namespace nafg
