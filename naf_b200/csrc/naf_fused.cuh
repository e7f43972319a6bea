// naf_fused.cuh — the single-pass encode transform as one sm_100a kernel (logic: naf_fused_hd.cuh).
//
// k_fused: one CTA per 16 KB tile of text.  Tiles are handed out by an atomic ticket, so a tile's predecessors are always
// resident or finished and the two chained look-backs (decoupled look-back: every tile publishes its own aggregate at once
// and its inclusive state as soon as it knows its prefix) cannot deadlock.  The tile reaches shared memory by one bulk
// asynchronous copy (cp.async.bulk + mbarrier; UBLKCP in SASS); everything after that works out of shared memory and every
// output byte is written once.  Algorithmic traffic = text read once + streams written once.
#pragma once
#include "common.cuh"
#include "naf_fused_hd.cuh"

namespace nafg {

static const int FUSED_NT = 256, FUSED_CPT = FT_CHUNKS / FUSED_NT, FUSED_SPT = FT_MAXSEG / FUSED_NT;

struct FusedArgs {
    FusedCfg C;
    const u8 *text;
    u64 *st1;                 // look-back #1: status << 62 | payload, one word per tile
    u32 *st2_status;          // look-back #2: 0 empty, 1 aggregate, 2 inclusive
    u64 *st2_agg, *st2_inc;   // [tile][F2_WORDS]
    u32 *ticket, *flag;
    unsigned long long *longest;
    FusedTotals *totals;
    u32 ntiles;
};

// shared memory carve-up (bytes)
struct FusedSmem {
    static const u32 o_text = 0, o_stage = o_text + FT_BYTES + 16, o_seg = o_stage + ((FT_STAGE + 15) & ~15u);
    static const u32 seg_n = FT_MAXSEG + 8;
    static const u32 o_role = o_seg + 7 * 2 * seg_n, o_lut = (o_role + seg_n + 15) & ~15u, o_sh = o_lut + 256;
    static const u32 o_scan = (o_sh + (u32)sizeof(FusedShared) + 15) & ~15u, o_mbar = o_scan + 34 * 8, total = o_mbar + 16;
};

__device__ __forceinline__ u64 ld_vol64(const u64 *p) { return *(const volatile u64 *)p; }
__device__ __forceinline__ u32 ld_vol32(const u32 *p) { return *(const volatile u32 *)p; }

__device__ __forceinline__ F2 f2_shfl_down(const F2 &v, int d)
{
    F2 r;
    r.ids = __shfl_down_sync(0xFFFFFFFFu, v.ids, d); r.comm = __shfl_down_sync(0xFFFFFFFFu, v.comm, d);
    r.seq = __shfl_down_sync(0xFFFFFFFFu, v.seq, d); r.qual = __shfl_down_sync(0xFFFFFFFFu, v.qual, d);
    r.rec = __shfl_down_sync(0xFFFFFFFFu, v.rec, d); r.srec = __shfl_down_sync(0xFFFFFFFFu, v.srec, d);
    r.qrec = __shfl_down_sync(0xFFFFFFFFu, v.qrec, d); r.last = __shfl_down_sync(0xFFFFFFFFu, v.last, d);
    return r;
}
__device__ __forceinline__ F2 f2_bcast0(const F2 &v)
{
    F2 r;
    r.ids = __shfl_sync(0xFFFFFFFFu, v.ids, 0); r.comm = __shfl_sync(0xFFFFFFFFu, v.comm, 0); r.seq = __shfl_sync(0xFFFFFFFFu, v.seq, 0);
    r.qual = __shfl_sync(0xFFFFFFFFu, v.qual, 0); r.rec = __shfl_sync(0xFFFFFFFFu, v.rec, 0); r.srec = __shfl_sync(0xFFFFFFFFu, v.srec, 0);
    r.qrec = __shfl_sync(0xFFFFFFFFu, v.qrec, 0); r.last = __shfl_sync(0xFFFFFFFFu, v.last, 0);
    return r;
}
__device__ __forceinline__ void f2_store(u64 *p, const F2 &v)
{
    __stcg(p + 0, v.ids); __stcg(p + 1, v.comm); __stcg(p + 2, v.seq); __stcg(p + 3, v.qual);
    __stcg(p + 4, v.rec); __stcg(p + 5, v.srec); __stcg(p + 6, v.qrec); __stcg(p + 7, v.last);
}
__device__ __forceinline__ F2 f2_load(const u64 *p)
{
    F2 v;
    v.ids = __ldcg(p + 0); v.comm = __ldcg(p + 1); v.seq = __ldcg(p + 2); v.qual = __ldcg(p + 3);
    v.rec = __ldcg(p + 4); v.srec = __ldcg(p + 5); v.qrec = __ldcg(p + 6); v.last = __ldcg(p + 7);
    return v;
}

// look-back #1, by one warp: prefix of f1 elements over the tiles before `tile` (older first), publishing mine
__device__ u32 fused_lookback1(u64 *st1, u32 tile, u32 mine, bool fastq)
{
    const unsigned lane = threadIdx.x & 31;
    const u32 init = fastq ? 0u : (u32)FE_HDR;                  // the machine starts inside the first header (process.c:589)
    if (tile == 0) {
        if (lane == 0) *(volatile u64 *)(st1) = (2ull << 62) | f1_compose(fastq, init, mine);
        return init;
    }
    if (lane == 0) *(volatile u64 *)(st1 + tile) = (1ull << 62) | mine;
    u32 prefix = 0; bool have = false;
    for (long long base = tile;; base -= 32) {
        const long long idx = base - 1 - lane;
        u64 d;
        if (idx >= 0) { do { d = ld_vol64(st1 + idx); } while ((d >> 62) == 0); }
        else d = (2ull << 62) | init;
        const unsigned m = __ballot_sync(0xFFFFFFFFu, (d >> 62) == 2);
        const unsigned k = m ? __ffs(m) - 1 : 31;
        u32 v = (u32)d;
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) { const u32 o = __shfl_down_sync(0xFFFFFFFFu, v, s); if (lane + s <= k) v = f1_compose(fastq, o, v); }
        v = __shfl_sync(0xFFFFFFFFu, v, 0);
        prefix = have ? f1_compose(fastq, v, prefix) : v; have = true;
        if (m) break;
    }
    if (lane == 0) *(volatile u64 *)(st1 + tile) = (2ull << 62) | f1_compose(fastq, prefix, mine);
    return prefix;
}

// look-back #2, by one warp
__device__ F2 fused_lookback2(const FusedArgs &A, u32 tile, const F2 &mine, bool fastq)
{
    const unsigned lane = threadIdx.x & 31;
    const F2 init = f2_initial();
    if (tile == 0) {
        if (lane == 0) { f2_store(A.st2_inc, f2_compose(init, mine, fastq)); __threadfence(); *(volatile u32 *)A.st2_status = 2; }
        return init;
    }
    if (lane == 0) { f2_store(A.st2_agg + (u64)tile * F2_WORDS, mine); __threadfence(); *(volatile u32 *)(A.st2_status + tile) = 1; }
    F2 prefix = init; bool have = false;
    for (long long base = tile;; base -= 32) {
        const long long idx = base - 1 - lane;
        u32 st;
        if (idx >= 0) { do { st = ld_vol32(A.st2_status + idx); } while (st == 0); }
        else st = 3;
        __threadfence();
        const unsigned m = __ballot_sync(0xFFFFFFFFu, st >= 2);
        const unsigned k = m ? __ffs(m) - 1 : 31;
        F2 v = init;
        if (lane <= k && idx >= 0) v = f2_load((st == 2 ? A.st2_inc : A.st2_agg) + (u64)idx * F2_WORDS);
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) { const F2 o = f2_shfl_down(v, s); if (lane + s <= k) v = f2_compose(o, v, fastq); }
        v = f2_bcast0(v);
        prefix = have ? f2_compose(v, prefix, fastq) : v; have = true;
        if (m) break;
    }
    if (lane == 0) { f2_store(A.st2_inc + (u64)tile * F2_WORDS, f2_compose(prefix, mine, fastq)); __threadfence(); *(volatile u32 *)(A.st2_status + tile) = 2; }
    return prefix;
}

// staging -> global, both congruent mod 16: 16-byte pieces, head and tail bytes by a few threads
__device__ __forceinline__ void fused_copy_out(u8 *dst, const u8 *stage, u32 len)
{
    const u32 head = min(len, (u32)((16 - ((uintptr_t)dst & 15)) & 15));
    const u32 nu = (len - head) >> 4, done = head + (nu << 4), tail = len - done;
    if (threadIdx.x < head) dst[threadIdx.x] = stage[threadIdx.x];
    else if (threadIdx.x >= 32 && threadIdx.x - 32 < tail) dst[done + threadIdx.x - 32] = stage[done + threadIdx.x - 32];
    for (u32 u = threadIdx.x; u < nu; u += FUSED_NT) *(uint4 *)(dst + head + 16 * u) = *(const uint4 *)(stage + head + 16 * u);
}

__global__ void __launch_bounds__(FUSED_NT, 4) k_fused(const FusedArgs A)
{
    extern __shared__ __align__(128) u8 smem[];
    FusedTile T;
    T.text = smem + FusedSmem::o_text; T.stage = smem + FusedSmem::o_stage;
    u16 *seg = (u16 *)(smem + FusedSmem::o_seg);
    T.nlmask = nullptr;
    T.seg_end = seg; T.seg_sp = seg + FusedSmem::seg_n; T.seg_off = seg + 2 * FusedSmem::seg_n; T.seg_offb = seg + 3 * FusedSmem::seg_n;
    T.seg_list = seg + 4 * FusedSmem::seg_n; T.recseq = seg + 5 * FusedSmem::seg_n; T.recqual = seg + 6 * FusedSmem::seg_n;
    T.seg_role = smem + FusedSmem::o_role;
    u8 *lut = smem + FusedSmem::o_lut;
    FusedShared *sh = (FusedShared *)(smem + FusedSmem::o_sh);
    T.sh = sh;
    u64 *scan = (u64 *)(smem + FusedSmem::o_scan);
    u64 *mbar = (u64 *)(smem + FusedSmem::o_mbar);
    const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    FusedCfg C = A.C;
    const bool fastq = C.fastq != 0;

    if (tid == 0) {
        const u32 tile = atomicAdd(A.ticket, 1u);
        const u64 lo = (u64)tile * FT_BYTES;
        sh->tile = tile;
        u32 l = C.p0 > lo ? (C.p0 - lo >= FT_BYTES ? FT_BYTES : (u32)(C.p0 - lo)) : 0u;
        const u32 h = C.n >= lo + FT_BYTES ? FT_BYTES : (C.n > lo ? (u32)(C.n - lo) : 0u);
        if (l > h) l = h;
        sh->live_lo = l; sh->live_hi = h;
        sh->abort_ = 0; sh->flag = 0; sh->maxlen = 0;
        const u32 bar = (u32)__cvta_generic_to_shared(mbar);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (C.seq_mode == FS_PACK4) lut[tid] = C.lut[tid];
    __syncthreads();
    const u32 tile = sh->tile;
    const u64 lo = (u64)tile * FT_BYTES;
    C.lut = lut;

    // ---- the tile -> shared memory
    const bool bulk = lo + FT_BYTES <= C.n && (((uintptr_t)A.text) & 15) == 0;
    if (bulk) {
        const u32 bar = (u32)__cvta_generic_to_shared(mbar);
        if (tid == 0) {
            const u32 dst = (u32)__cvta_generic_to_shared(T.text);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((u32)FT_BYTES) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(dst), "l"(A.text + lo), "r"((u32)FT_BYTES), "r"(bar) : "memory");
        }
        u32 done = 0;
        while (!done) {
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(bar) : "memory");
        }
    } else {
        for (u32 c = tid; c < FT_CHUNKS; c += FUSED_NT) {
            const u64 at = lo + 16ull * c;
            uint4 v = make_uint4(0, 0, 0, 0);
            if (at + 16 <= C.n && (((uintptr_t)A.text) & 15) == 0) v = __ldg((const uint4 *)(A.text + at));
            else if (at < C.n) {
                u32 w[4] = {0, 0, 0, 0};
                for (u32 i = 0; i < 16 && at + i < C.n; i++) w[i >> 2] |= (u32)A.text[at + i] << (8 * (i & 3));
                v = make_uint4(w[0], w[1], w[2], w[3]);
            }
            *(uint4 *)(T.text + 16 * c) = v;
        }
    }
    if (tid < 4) ((u32 *)(T.text + FT_BYTES))[tid] = 0;             // the word-wise copies read one word past a segment
    __syncthreads();

    // ---- phase 1: newlines -> line segments
    u32 m[FUSED_CPT]; u64 cnt = 0;
#pragma unroll
    for (int k = 0; k < FUSED_CPT; k++) { m[k] = T.chunk_mask(tid + k * FUSED_NT); cnt |= (u64)__popc(m[k]) << (16 * k); }
    u64 tot;
    const u64 pre = block_excl_scan(cnt, &tot, scan);
    u32 nl = 0, first[FUSED_CPT];
#pragma unroll
    for (int k = 0; k < FUSED_CPT; k++) { first[k] = nl + (u32)((pre >> (16 * k)) & 0xFFFF); nl += (u32)((tot >> (16 * k)) & 0xFFFF); }
    const bool aborted = nl + 1 > FT_MAXSEG;
    if (!aborted) {
#pragma unroll
        for (int k = 0; k < FUSED_CPT; k++) T.put_lines(tid + k * FUSED_NT, m[k], first[k]);
    }
    if (tid == 0) { sh->nseg = nl + 1; if (aborted) { sh->abort_ = 1; sh->flag = FU_LINES; } }
    __syncthreads();

    // ---- look-back #1 (warp 0): the kind of line the tile starts in
    if (warp == 0) {
        const u32 last_nl = nl && !aborted ? T.seg_end[nl - 1] : 0;
        const u32 agg1 = fastq ? nl : (aborted ? (u32)FE_ID : T.fasta_element(nl, last_nl));
        const u32 e1 = fused_lookback1(A.st1, tile, agg1, fastq);
        if (lane == 0) {
            sh->entry1 = e1;
            const u64 at = lo + sh->live_lo;
            const u32 els = fastq ? (u32)(at > C.p0 && A.text[at - 1] == '\n') : (u32)(e1 == FE_LS);
            sh->entry_ls = els;
            const u32 role0 = fastq ? (e1 & 3) : (e1 == FE_HDR ? (u32)FR_HDR : (u32)FR_SEQ);
            u32 sp = 0;
            if (sh->live_lo < sh->live_hi && role0 == FR_HDR && !els) { u32 f = 0; sp = fast_lookback_space(A.text, C.p0, at, f); if (f) sh->flag |= FU_LOOKBACK; }
            sh->entry_sp = sp;
        }
    }
    __syncthreads();

    u32 flag = 0;
    u64 maxlen = 0;
    F2 agg2 = f2_initial(); agg2.last = 0;
    if (!aborted) {
        // ---- phase 3: one thread per segment (FUSED_SPT consecutive ones), block scans -> places inside the tile
        const u32 nseg = nl + 1, j0 = tid * FUSED_SPT;
        u64 sa[FUSED_SPT], sb[FUSED_SPT], ta = 0, tb = 0;
#pragma unroll
        for (int k = 0; k < FUSED_SPT; k++) { sa[k] = sb[k] = 0; if (j0 + k < nseg) T.classify(C, j0 + k, sa[k], sb[k], flag); ta += sa[k]; tb += sb[k]; }
        u64 tota, totb;
        u64 pa = block_excl_scan(ta, &tota, scan);
        u64 pb = block_excl_scan(tb, &totb, scan);
        if (tid == 0) {
            sh->t_ids = (u32)(tota & 0xFFFF); sh->t_comm = (u32)((tota >> 16) & 0xFFFF); sh->t_seq = (u32)((tota >> 32) & 0xFFFF); sh->t_qual = (u32)(tota >> 48);
            sh->t_rec = (u32)(totb & 0xFFFF); sh->n_hdr = (u32)((totb >> 16) & 0xFFFF); sh->n_seq = (u32)((totb >> 32) & 0xFFFF); sh->n_qual = (u32)(totb >> 48);
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < FUSED_SPT; k++) { if (j0 + k < nseg) T.place(C, j0 + k, pa, pb); pa += sa[k]; pb += sb[k]; }
        __syncthreads();
        if (warp == 0) agg2 = T.aggregate(C);
    }
    // ---- look-back #2 (warp 0): global offsets, straddling record / line / byte
    if (warp == 0) {
        const F2 p2 = fused_lookback2(A, tile, agg2, fastq);
        if (lane == 0) { sh->pre = p2; if (!aborted) T.layout(C); }
    }
    __syncthreads();

    if (!aborted) {
        // ---- phase 5: copies (one 8-lane group per segment), records, lines
        const u32 nlist = sh->n_seq + sh->n_qual + sh->n_hdr;
        for (u32 k = tid / FT_GROUP; k < nlist; k += FUSED_NT / FT_GROUP) flag |= T.copy_segment(C, k, tid % FT_GROUP);
        for (u32 k = tid; k < sh->t_rec; k += FUSED_NT) { const u64 L = T.finish_record(C, k, flag); if (L > maxlen) maxlen = L; }
        if (!fastq) for (u32 k = tid; k < sh->n_seq; k += FUSED_NT) { const u64 L = T.line_length(k); if (L > maxlen) maxlen = L; }
        __syncthreads();
        // ---- phase 6: staging -> global
        fused_copy_out(C.ids + sh->pre.ids, T.stage + sh->s_ids, sh->t_ids);
        fused_copy_out(C.comm + sh->pre.comm, T.stage + sh->s_comm, sh->t_comm);
        if (C.seq_mode == FS_PACK4 && sh->t_seq) {
            const u32 npieces = ((u32)(sh->pre.seq & 31) + sh->t_seq + 31) / 32;
            for (u32 q = tid; q < npieces; q += FUSED_NT) flag |= T.pack_piece(C, q, [](u32 *p, u32 v) { atomicOr(p, v); });
        }
    }
    if (tid == 0) flag |= sh->flag;
    if (flag) atomicOr(A.flag, flag);
    // one atomic per warp at most, and none once the global maximum is at least ours (lines / reads are mostly equally long)
    for (int d = 16; d; d >>= 1) { const u64 o = __shfl_xor_sync(0xFFFFFFFFu, maxlen, d); if (o > maxlen) maxlen = o; }
    if (lane == 0 && maxlen > *(volatile unsigned long long *)A.longest) atomicMax(A.longest, (unsigned long long)maxlen);
}

__global__ void k_fused_finish(const FusedArgs A)
{
    const bool fastq = A.C.fastq != 0;
    u32 f1 = fastq ? 0u : (u32)FE_HDR; F2 f2 = f2_initial();
    if (A.ntiles) { f1 = (u32)A.st1[A.ntiles - 1]; f2 = f2_load(A.st2_inc + (u64)(A.ntiles - 1) * F2_WORDS); }
    fused_finish(A.C, f1, f2, *A.flag, *A.longest, A.text, *A.totals);
}

}  // namespace nafg
