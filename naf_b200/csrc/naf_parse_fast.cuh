// naf_parse_fast.cuh — kernels of the canonical-input parser (logic and conditions: naf_fast_hd.cuh).
//
// Same three passes and the same per-thread / per-tile records as the general parser (naf_parse.cuh), so
// everything downstream of the scatter (lengths, quality-length check, 4-bit pack, mask RLE, end-of-input
// handling, error reporting) is shared.  What differs is the cost per byte: no action tables, no state-map
// algebra — '\n' and ' ' positions from a few SWAR operations per word, a handful of instructions per *line*, and
// word-wise copies between bank-conflict-free (one pad word per 16: fast_pad()) shared-memory tiles.
//   k_fast_tiles    per tile: number of '\n' (FASTQ) / line-kind element (FASTA); C1 check
//   k_fast_scan     one CTA: line index (mod 4) / line kind entering every tile; parser state at end of input
//   k_fast_count    per thread: bytes emitted per stream, records ended (ThreadInfo), per tile TileCounts
//   k_fast_scatter  per thread: runs of bytes -> staged per stream in shared memory -> coalesced stores
#pragma once
#include "naf_parse.cuh"
#include "naf_fast_hd.cuh"

namespace nafg {

struct FastArgs {
    ParseArgs P;
    u32 *tile_elem;          // pass 1 output, one per tile (+1): FASTQ '\n' count, FASTA element
    u32 *tile_entry;         // scan output: FASTQ lines before the tile (mod 2^32), FASTA element entering the tile; [ntiles] = at end of input
    u32 *flag;               // FF_* bits: input is not canonical -> caller falls back to the general parser
    int upper;               // toupper() the sequence (protein / text with --no-mask, process.c:49)
    int seq_check;           // 0 none here (DNA / RNA: the pack LUT checks), 1 protein, 2 text, 3 text where '>' is unexpected
};

static const int FAST_TILE_SMEM = (PTILE + 64) / 64 * 68;     // fast_pad() layout, + one padding row: the word-wise copies read one word ahead

struct FastRow {                                  // my 64-byte row of the padded tile
    const u8 *row;
    __device__ __forceinline__ u32 operator()(u32 i) const { return row[i]; }
};

// my 64 bytes -> registers and my row of the swizzled tile; bytes outside [p0, n) read as 'A'
__device__ __forceinline__ void fast_load(const ParseArgs &A, u64 lo, u32 w[16], u8 *tile, u32 &b0, u32 &b1)
{
    b0 = lo >= A.p0 ? 0u : (A.p0 - lo >= 64 ? 64u : (u32)(A.p0 - lo));
    b1 = lo + 64 <= A.n ? 64u : (lo >= A.n ? 0u : (u32)(A.n - lo));
    if (b0 == 0 && b1 == 64 && (((uintptr_t)A.text) & 15) == 0) {
        const uint4 *v = (const uint4 *)(A.text + lo);
#pragma unroll
        for (int k = 0; k < 4; k++) { const uint4 x = __ldg(v + k); w[4 * k] = x.x; w[4 * k + 1] = x.y; w[4 * k + 2] = x.z; w[4 * k + 3] = x.w; }
    } else {
#pragma unroll
        for (int k = 0; k < 16; k++) {
            u32 x = FAST_FILL;
            if (b0 < b1) {
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const u32 i = 4 * k + j;
                    if (i >= b0 && i < b1) x = (x & ~(0xFFu << (8 * j))) | ((u32)A.text[lo + i] << (8 * j));
                }
            }
            w[k] = x;
        }
    }
    if (tile) {
        u32 *row = (u32 *)(tile + threadIdx.x * 68);
#pragma unroll
        for (int k = 0; k < 16; k++) row[k] = w[k];
    }
}

// ordered (non-commutative) exclusive scan of FASTA elements across the CTA; *total = composition of all
__device__ __forceinline__ u32 block_excl_scan_fe(u32 e, u32 *total, u32 *sm /* >= 34 */)
{
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    u32 incl = e;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const u32 g = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= (unsigned)d) incl = fe_compose(g, incl); }
    u32 excl = __shfl_up_sync(0xFFFFFFFFu, incl, 1);
    if (lane == 0) excl = FE_ID;
    if (lane == 31) sm[warp] = incl;
    __syncthreads();
    if (threadIdx.x == 0) {
        u32 run = FE_ID;
        for (unsigned k = 0; k < nwarps; k++) { const u32 m = sm[k]; sm[k] = run; run = fe_compose(run, m); }
        sm[32] = run;
    }
    __syncthreads();
    const u32 r = fe_compose(sm[warp], excl);
    *total = sm[32];
    __syncthreads();
    return r;
}

// ------------------------------------------------------------------ pass 1
template <bool FASTQ> __global__ void __launch_bounds__(PT) k_fast_tiles(const FastArgs F)
{
    __shared__ __align__(16) u8 tile[FAST_TILE_SMEM];
    __shared__ u64 sm64[33];
    __shared__ u32 sm32[34];
    const ParseArgs &A = F.P;
    const u64 lo = (u64)blockIdx.x * PTILE + (u64)threadIdx.x * PB;
    u32 w[16], b0, b1;
    fast_load(A, lo, w, FASTQ ? nullptr : tile, b0, b1);
    u64 nl; u32 bad;
    fast_chunk_scan<false>(w, nl, bad);
    if (bad) atomicOr(F.flag, (u32)FF_BADBYTE);
    if (FASTQ) {
        u64 total; block_excl_scan((u64)__popcll(nl), &total, sm64);
        if (threadIdx.x == 0) F.tile_elem[blockIdx.x] = (u32)total;
    } else {
        const FastRow row{tile + threadIdx.x * 68};
        u32 total; block_excl_scan_fe(fasta_chunk_element(row, nl, b0, b1), &total, sm32);
        if (threadIdx.x == 0) F.tile_elem[blockIdx.x] = total;
    }
}

// ------------------------------------------------------------------ scan over tiles (one CTA)
template <bool FASTQ> __global__ void __launch_bounds__(1024) k_fast_scan(const FastArgs F)
{
    __shared__ u32 agg[1024];
    const ParseArgs &A = F.P;
    const u64 per = (A.ntiles + 1023) / 1024;
    u64 lo = (u64)threadIdx.x * per, hi = lo + per;
    if (lo > A.ntiles) lo = A.ntiles;
    if (hi > A.ntiles) hi = A.ntiles;
    u32 f = FASTQ ? 0u : (u32)FE_ID;
    for (u64 t = lo; t < hi; t++) f = FASTQ ? f + F.tile_elem[t] : fe_compose(f, F.tile_elem[t]);
    agg[threadIdx.x] = f;
    __syncthreads();
    if (threadIdx.x == 0) {
        u32 run = FASTQ ? 0u : (u32)FE_HDR;                       // the machine starts inside the first header (process.c:589)
        for (int c = 0; c < 1024; c++) { const u32 m = agg[c]; agg[c] = run; run = FASTQ ? run + m : fe_compose(run, m); }
        F.tile_entry[A.ntiles] = run;
        // parser state at the end of the input, in the general machine's numbering (host-side end-of-input logic is shared)
        u32 role, ls, sp = 0, fl = 0;
        if (FASTQ) { role = run & 3; ls = A.n > A.p0 && A.text[A.n - 1] == '\n'; }
        else { role = run == FE_HDR ? (u32)FR_HDR : (u32)FR_SEQ; ls = run == FE_LS; }
        if (role == FR_HDR && !ls) sp = fast_lookback_space(A.text, A.p0, A.n, fl);
        if (fl) atomicOr(F.flag, fl);
        A.tile_state[A.ntiles] = (u8)fast_end_state(FASTQ, role, sp, ls);
    }
    __syncthreads();
    u32 s = agg[threadIdx.x];
    for (u64 t = lo; t < hi; t++) { F.tile_entry[t] = s; s = FASTQ ? s + F.tile_elem[t] : fe_compose(s, F.tile_elem[t]); }
}

// state entering my chunk
template <bool FASTQ>
__device__ __forceinline__ FastState fast_entry(const FastArgs &F, const FastRow &row, u64 nl, u32 b0, u32 b1, u64 lo, u64 *sm64, u32 *sm32, u32 &flag)
{
    const ParseArgs &A = F.P;
    FastState st; st.sp = 0;
    if (FASTQ) {
        u64 total; const u64 before = block_excl_scan((u64)__popcll(nl), &total, sm64);
        st.role = (F.tile_entry[blockIdx.x] + (u32)before) & 3;
        st.ls = lo > A.p0 && b1 > 0 && A.text[lo - 1] == '\n';
    } else {
        u32 total; const u32 e = fe_compose(F.tile_entry[blockIdx.x], block_excl_scan_fe(fasta_chunk_element(row, nl, b0, b1), &total, sm32));
        st.role = e == FE_HDR ? (u32)FR_HDR : (u32)FR_SEQ;
        st.ls = e == FE_LS;
    }
    if (b0 < b1 && st.role == FR_HDR && !st.ls) st.sp = fast_lookback_space(A.text, A.p0, lo + b0, flag);
    return st;
}

// ------------------------------------------------------------------ pass 2
template <bool FASTQ> __global__ void __launch_bounds__(PT) k_fast_count(const FastArgs F)
{
    __shared__ __align__(16) u8 tile[FAST_TILE_SMEM];
    __shared__ u64 sm64[33];
    __shared__ u32 sm32[34];
    const ParseArgs &A = F.P;
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u64 gid = (u64)blockIdx.x * PT + threadIdx.x, lo = gid * PB;
    u32 w[16], b0, b1;
    fast_load(A, lo, w, tile, b0, b1);
    u64 nl, sp; u32 bad, flag = 0;
    fast_chunk_scan<true>(w, nl, bad, &sp);
    const FastRow row{tile + threadIdx.x * 68};
    FastState st = fast_entry<FASTQ>(F, row, nl, b0, b1, lo, sm64, sm32, flag);
    const u32 entry = st.role | (st.sp << 2) | (st.ls << 3);
    FastEmit n = {0, 0, 0, 0, 0};
    FastLine ln = {0, 0, 0};
    FastNoSink sink;
    fast_walk<FASTQ, false>(row, nl, sp, b0, b1, st, lo, n, sink, 0, 0, 0, ln, flag);
    if (flag) atomicOr(F.flag, flag);
    ThreadInfo ti; ti.state = (u8)entry; ti.ids = (u8)n.ids; ti.comm = (u8)n.comm; ti.seq = (u8)n.seq; ti.cnt = (u8)n.seq; ti.qual = (u8)n.qual;
    ti.rec = (u8)n.rec; ti.line = (u8)ln.mark;
    A.tinfo[gid] = ti;
    // tile totals: four 16-bit counters packed in one word (each <= 16385 per tile), records separately
    TileCounts tc; u64 tot;
    const u64 pre = block_excl_scan((u64)n.ids | ((u64)n.comm << 16) | ((u64)n.seq << 32) | ((u64)n.qual << 48), &tot, sm64);
    tc.ids = tot & 0xFFFF; tc.comm = (tot >> 16) & 0xFFFF; tc.seq = tc.seq_counted = (tot >> 32) & 0xFFFF; tc.qual = tot >> 48;
    block_excl_scan((u64)n.rec, &tot, sm64); tc.rec = tot;
    u64 v = ln.mark ? ((pre >> 32) & 0xFFFF) + ln.mark : 0;          // counted bytes (tile-relative) at my last line end, +1
    for (int d = 16; d; d >>= 1) { const u64 o = __shfl_xor_sync(0xFFFFFFFFu, v, d); if (o > v) v = o; }
    if (lane == 0) sm64[warp] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        u64 m = 0;
        for (int k = 0; k < PT / 32; k++) if (sm64[k] > m) m = sm64[k];
        tc.has_line = m != 0; tc.line_last = m ? m - 1 : 0; tc.pad = 0;
        A.tile[blockIdx.x] = tc;
    }
}

// ------------------------------------------------------------------ pass 3
static const int FAST_STAGE_SMEM = (PTILE + 256) / 64 * 68;    // fast_pad() layout

// staged bytes of one stream (linear stage offset s0, congruent mod 4 to dst) -> global memory, one word per thread and step
template <int CHECK>      // 0 none, 1 protein, 2 text, 3 text with '>' unexpected, 4 quality
__device__ __forceinline__ u32 fast_copy_out(u8 *dst, const u8 *stage, u32 s0, u32 len, bool upper)
{
    u32 bad = 0;
    const u32 head = min(len, (u32)((4 - ((uintptr_t)dst & 3)) & 3));
    const u32 nw = (len - head) / 4, done = head + nw * 4;
    // head / tail bytes: checked as single bytes by the first threads
    if (threadIdx.x < head || (threadIdx.x >= 32 && threadIdx.x - 32 < len - done)) {
        const u32 i = threadIdx.x < head ? threadIdx.x : done + (threadIdx.x - 32);
        u32 c = stage[fast_pad(s0 + i)];
        const u32 v = c * 0x01010101u;
        if (CHECK == 1) bad |= swar_bad_protein(v); else if (CHECK == 2) bad |= swar_bad_text(v, false); else if (CHECK == 3) bad |= swar_bad_text(v, true);
        else if (CHECK == 4) bad |= swar_bad_qual(v);
        if (upper && c >= 'a' && c <= 'z') c -= 32;
        dst[i] = (u8)c;
    }
    u32 *dw = (u32 *)(dst + head);
    for (u32 k = threadIdx.x; k < nw; k += blockDim.x) {
        u32 v = *(const u32 *)(stage + fast_pad(s0 + head + 4 * k));
        if (CHECK == 1) bad |= swar_bad_protein(v); else if (CHECK == 2) bad |= swar_bad_text(v, false); else if (CHECK == 3) bad |= swar_bad_text(v, true);
        else if (CHECK == 4) bad |= swar_bad_qual(v);
        if (upper) v = swar_upper(v);
        dw[k] = v;
    }
    return bad;
}

template <bool FASTQ> __global__ void __launch_bounds__(PT) k_fast_scatter(const FastArgs F)
{
    extern __shared__ __align__(16) u8 dyn[];                  // [text tile][staging]
    u8 *tile = dyn, *stage = dyn + FAST_TILE_SMEM;
    __shared__ u64 sm64[33];
    const ParseArgs &A = F.P;
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u64 gid = (u64)blockIdx.x * PT + threadIdx.x, lo = gid * PB;
    const ThreadInfo ti = A.tinfo[gid];
    u64 tot;
    const u64 pre = block_excl_scan((u64)ti.ids | ((u64)ti.comm << 16) | ((u64)ti.seq << 32) | ((u64)ti.qual << 48), &tot, sm64);
    const u32 l_ids = pre & 0xFFFF, l_comm = (pre >> 16) & 0xFFFF, l_seq = (pre >> 32) & 0xFFFF, l_qual = (u32)(pre >> 48);
    const u32 t_ids = tot & 0xFFFF, t_comm = (tot >> 16) & 0xFFFF, t_seq = (tot >> 32) & 0xFFFF, t_qual = (u32)(tot >> 48);
    u64 tot2;
    const u64 o_rec = block_excl_scan((u64)ti.rec, &tot2, sm64) + A.pre_rec[blockIdx.x];
    const u64 o_cnt = l_seq + A.pre_cnt[blockIdx.x], o_qual = l_qual + A.pre_qual[blockIdx.x];
    // counted-sequence value at the last line end before this thread: exclusive max-scan (values are monotone)
    const u64 mine = ti.line ? o_cnt + ti.line : 0;            // +1 encoding
    u64 run = mine;
    for (int d = 1; d < 32; d <<= 1) { const u64 g = __shfl_up_sync(0xFFFFFFFFu, run, d); if (lane >= (unsigned)d && g > run) run = g; }
    if (lane == 31) sm64[warp] = run;
    __syncthreads();
    u64 before = A.pre_line[blockIdx.x] + 1;
    for (unsigned k = 0; k < warp; k++) if (sm64[k] > before) before = sm64[k];
    const u64 prev_lane = __shfl_up_sync(0xFFFFFFFFu, run, 1);
    if (lane > 0 && prev_lane > before) before = prev_lane;

    // staging layout: ids | comments | sequence | quality, each region congruent mod 4 to its destination
    u8 *g_ids = A.ids + A.pre_ids[blockIdx.x], *g_comm = A.comm + A.pre_comm[blockIdx.x];
    u8 *g_seq = A.bases + A.pre_seq[blockIdx.x], *g_qual = A.qual + A.pre_qual[blockIdx.x];
    const u32 s_ids = (u32)((uintptr_t)g_ids & 3);
    const u32 s_comm = ((s_ids + t_ids + 3) & ~3u) + (u32)((uintptr_t)g_comm & 3);
    const u32 s_seq = ((s_comm + t_comm + 3) & ~3u) + (u32)((uintptr_t)g_seq & 3);
    const u32 s_qual = ((s_seq + t_seq + 3) & ~3u) + (u32)((uintptr_t)g_qual & 3);

    u32 w[16], b0, b1;
    fast_load(A, lo, w, tile, b0, b1);
    u64 nl, sp; u32 bad, flag = 0;
    fast_chunk_scan<true>(w, nl, bad, &sp);
    const FastRow row{tile + threadIdx.x * 68};
    FastState st; st.role = ti.state & 3; st.sp = (ti.state >> 2) & 1; st.ls = (ti.state >> 3) & 1;
    FastEmit m = {0, 0, 0, 0, 0};
    FastLine ln = {before - 1, 0, 0};
    FastSmemSink sink;
    sink.tile = tile; sink.stage = stage; sink.src0 = threadIdx.x * 64u;
    sink.base[0] = s_ids + l_ids; sink.base[1] = s_comm + l_comm; sink.base[2] = s_seq + l_seq; sink.base[3] = s_qual + l_qual;
    sink.rec_seq_end = A.rec_seq_end; sink.rec_qual_end = A.rec_qual_end; sink.rec_pos = A.rec_pos; sink.fastq = FASTQ;
    sink.begin();
    fast_walk<FASTQ, true>(row, nl, sp, b0, b1, st, lo, m, sink, o_cnt, o_qual, o_rec, ln, flag);
    sink.flush();                                               // the copies, all lanes together
    if (!FASTQ) {
        // pending (unterminated) last line of the input (process.c:417-422)
        if (lo < A.n && lo + PB >= A.n) { const u64 d = o_cnt + m.seq - ln.base; if (d > ln.max) ln.max = d; }
    }
    __syncthreads();
    if (!FASTQ) {
        // one atomic per CTA at most, and none once the global maximum is at least ours (lines are mostly equally long:
        // an atomicMax per thread on one address serialises in L2 -- 10 ms on a 1 GB FASTA)
        u64 v = ln.max;
        for (int d = 16; d; d >>= 1) { const u64 o = __shfl_xor_sync(0xFFFFFFFFu, v, d); if (o > v) v = o; }
        if (lane == 0) sm64[warp] = v;
        __syncthreads();
        if (threadIdx.x == 0) {
            u64 mx = 0;
            for (int k = 0; k < PT / 32; k++) if (sm64[k] > mx) mx = sm64[k];
            if (mx > *(volatile unsigned long long *)A.longest) atomicMax(A.longest, (unsigned long long)mx);
        }
        __syncthreads();
    }
    fast_copy_out<0>(g_ids, stage, s_ids, t_ids, false);
    fast_copy_out<0>(g_comm, stage, s_comm, t_comm, false);
    u32 sbad;
    switch (F.seq_check) {
    case 1:  sbad = fast_copy_out<1>(g_seq, stage, s_seq, t_seq, F.upper != 0); break;
    case 2:  sbad = fast_copy_out<2>(g_seq, stage, s_seq, t_seq, F.upper != 0); break;
    case 3:  sbad = fast_copy_out<3>(g_seq, stage, s_seq, t_seq, F.upper != 0); break;
    default: sbad = fast_copy_out<0>(g_seq, stage, s_seq, t_seq, false); break;
    }
    if (sbad) flag |= FF_SEQ;
    if (FASTQ && fast_copy_out<4>(g_qual, stage, s_qual, t_qual, false)) flag |= FF_QUAL;
    if (flag) atomicOr(F.flag, flag);
}
static const size_t FAST_SCATTER_SMEM = FAST_TILE_SMEM + FAST_STAGE_SMEM;

}  // namespace nafg
