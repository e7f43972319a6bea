// naf_fused.cuh — the single-pass encode transform as one sm_100a kernel (logic: naf_fused_hd.cuh).
//
// k_fused: one CTA (512 threads) per 32 KB tile of text.  Tiles are handed out by an atomic ticket, so a tile's predecessors are always
// resident or finished and the two chained look-backs (decoupled look-back: every tile publishes its own aggregate at once
// and its inclusive state as soon as it knows its prefix) cannot deadlock.  The tile reaches shared memory by one bulk
// asynchronous copy (cp.async.bulk + mbarrier; UBLKCP in SASS); everything after that works out of shared memory and every
// output byte is written once.  Algorithmic traffic = text read once + streams written once.
#pragma once
#include "common.cuh"
#include "naf_fused_hd.cuh"

namespace nafg {

static const int FUSED_NT = FT_BYTES / 64, FUSED_CPT = FT_CHUNKS / FUSED_NT, FUSED_SPT = FT_MAXSEG / FUSED_NT;      // 64 bytes of text per thread

struct FusedArgs {
    FusedCfg C;
    const u8 *lut8;           // nuc_code, bit 7 = unexpected (Ctx::d_nuc_lut)
    const u8 *text;
    u64 *st1;                 // look-back #1: status << 62 | payload, one word per tile
    ulonglong2 *st2;          // look-back #2: four tagged 16-byte words per tile (aggregate, inclusive x 3)
    u32 *ticket, *flag;
    unsigned long long *longest;
    FusedTotals *totals;
    u32 ntiles;
};

// shared memory carve-up (bytes)
struct FusedSmem {
    static const u32 o_text = 0, o_stage = o_text + FT_BYTES + 16, o_seg = o_stage + ((FT_STAGE + 15) & ~15u);
    static const u32 seg_n = FT_MAXSEG + 8, desc_n = FT_MAXDESC + 8;
    static const u32 o_desc = o_seg + 3 * 2 * seg_n, o_lut = (o_desc + 3 * 2 * desc_n + 15) & ~15u, o_prev = o_lut + 1024, o_sh = o_prev + 64;
    static const u32 o_scan = (o_sh + (u32)sizeof(FusedShared) + 15) & ~15u, o_mbar = o_scan + 34 * 8, total = o_mbar + 16;
};

__device__ __forceinline__ u64 ld_vol64(const u64 *p) { return *(const volatile u64 *)p; }
__device__ __forceinline__ ulonglong2 ld_vol128(const ulonglong2 *p)
{
    ulonglong2 v;
    asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_vol128(ulonglong2 *p, u64 x, u64 y)
{
    asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(x), "l"(y) : "memory");
}

// Look-back #2 records.  Every 16-byte word carries its own "written" tag and is written once, by one 128-bit store, so a
// reader needs neither a fence nor a second round trip: it loads all four words of a tile and uses the inclusive state if
// all three of its words are there, else the aggregate if that is there, else it asks again.
static const u64 F2_TAG = 1ull << 62;
__device__ __forceinline__ void f2_put_aggregate(ulonglong2 *rec, const F2 &g)       // fields of ONE tile: byte counts <= FT_BYTES <= 2^16 - 1, records <= 2^12 - 1
{
    st_vol128(rec, F2_TAG | g.ids | (g.comm << 16) | (g.seq << 32) | (g.last << 48), g.qual | (g.rec << 16) | (g.srec << 28) | (g.qrec << 44));
}
__device__ __forceinline__ void f2_put_inclusive(ulonglong2 *rec, const F2 &g)       // sizes < 2^62, records < 2^40, srec / qrec saturated to 32 bits
{
    st_vol128(rec + 1, F2_TAG | g.ids, g.comm);
    st_vol128(rec + 2, F2_TAG | g.seq, g.qual);
    st_vol128(rec + 3, F2_TAG | g.rec | (g.last << 40), g.srec | (g.qrec << 32));
}
// -> 0 nothing yet, 1 aggregate, 2 inclusive
__device__ __forceinline__ int f2_get(const ulonglong2 *rec, F2 &g)
{
    const ulonglong2 a = ld_vol128(rec), i0 = ld_vol128(rec + 1), i1 = ld_vol128(rec + 2), i2 = ld_vol128(rec + 3);
    if ((i0.x & i1.x & i2.x) & F2_TAG) {
        g.ids = i0.x & (F2_TAG - 1); g.comm = i0.y; g.seq = i1.x & (F2_TAG - 1); g.qual = i1.y;
        g.rec = i2.x & ((1ull << 40) - 1); g.last = (i2.x >> 40) & 0x7FF; g.srec = i2.y & 0xFFFFFFFFull; g.qrec = i2.y >> 32;
        return 2;
    }
    if (a.x & F2_TAG) {
        g.ids = a.x & 0xFFFF; g.comm = (a.x >> 16) & 0xFFFF; g.seq = (a.x >> 32) & 0xFFFF; g.last = (a.x >> 48) & 0x7FF;
        g.qual = a.y & 0xFFFF; g.rec = (a.y >> 16) & 0xFFF; g.srec = (a.y >> 28) & 0xFFFF; g.qrec = (a.y >> 44) & 0xFFFF;
        return 1;
    }
    return 0;
}

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
        if (lane == 0) f2_put_inclusive(A.st2, f2_compose(init, mine, fastq));
        return init;
    }
    if (lane == 0) f2_put_aggregate(A.st2 + 4ull * tile, mine);
    F2 prefix = init; bool have = false;
    for (long long base = tile;; base -= 32) {
        const long long idx = base - 1 - lane;
        F2 v = init; int st = 2;
        if (idx >= 0) { do { st = f2_get(A.st2 + 4ull * idx, v); } while (st == 0); }
        const unsigned m = __ballot_sync(0xFFFFFFFFu, st == 2);
        const unsigned k = m ? __ffs(m) - 1 : 31;
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) { const F2 o = f2_shfl_down(v, s); if (lane + s <= k) v = f2_compose(o, v, fastq); }
        v = f2_bcast0(v);
        prefix = have ? f2_compose(v, prefix, fastq) : v; have = true;
        if (m) break;
    }
    if (lane == 0) f2_put_inclusive(A.st2 + 4ull * tile, f2_compose(prefix, mine, fastq));
    return prefix;
}

// one staged region -> global: head and tail bytes by a few threads, aligned 16-byte units by all of them
__device__ __forceinline__ u32 fused_region_out(const FusedTile &T, u32 off, u32 len, u8 *dst, int check, bool upper)
{
    if (!len) return 0;
    u32 head = (u32)((16 - ((uintptr_t)dst & 15)) & 15), bad = 0;
    if (head > len) head = len;
    const u32 nu = (len - head) >> 4, done = head + (nu << 4), tail = len - done;
    if (threadIdx.x < head) bad |= T.out_byte(off, threadIdx.x, dst, check, upper);
    else if (threadIdx.x >= 32 && threadIdx.x - 32 < tail) bad |= T.out_byte(off, done + threadIdx.x - 32, dst, check, upper);
    for (u32 u = threadIdx.x; u < nu; u += FUSED_NT) bad |= T.out_unit(off + head, u, dst + head, check, upper);
    return bad;
}

__global__ void __launch_bounds__(FUSED_NT, 1024 / FUSED_NT) k_fused(const FusedArgs A)
{
    extern __shared__ __align__(128) u8 smem[];
    FusedTile T;
    T.text = smem + FusedSmem::o_text; T.stage = smem + FusedSmem::o_stage;
    u16 *seg = (u16 *)(smem + FusedSmem::o_seg), *desc = (u16 *)(smem + FusedSmem::o_desc);
    T.seg_end = seg; T.recseq = seg + FusedSmem::seg_n; T.recqual = seg + 2 * FusedSmem::seg_n;
    T.d_src = desc; T.d_len = desc + FusedSmem::desc_n; T.d_dst = desc + 2 * FusedSmem::desc_n;
    u32 *lut = (u32 *)(smem + FusedSmem::o_lut);
    u8 *prev = smem + FusedSmem::o_prev;
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
    if (C.seq_mode == FS_PACK4 && tid < 256) lut[tid] = nuc_lut32(A.lut8[tid]);
    __syncthreads();
    const u32 tile = sh->tile;
    const u64 lo = (u64)tile * FT_BYTES;
    C.lut = lut;

    // ---- the tile -> shared memory (one bulk asynchronous copy), and the 64 bytes before it
    const bool bulk = lo + FT_BYTES <= C.n && (((uintptr_t)A.text) & 15) == 0;
    if (bulk) {
        if (tid == 0) {
            const u32 bar = (u32)__cvta_generic_to_shared(mbar), dst = (u32)__cvta_generic_to_shared(T.text);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((u32)FT_BYTES) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(dst), "l"(A.text + lo), "r"((u32)FT_BYTES), "r"(bar) : "memory");
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
    if (tid >= 64 && tid < 128) { const u32 i = tid - 64; prev[i] = lo + i >= 64 ? A.text[lo + i - 64] : (u8)0; }
    // the word-wise copies read one word past a run; its first byte is the text's next one (a CR LF pair may straddle the tiles)
    if (tid < 4) ((u32 *)(T.text + FT_BYTES))[tid] = tid == 0 && lo + FT_BYTES < C.n ? (u32)A.text[lo + FT_BYTES] : 0u;
    if (bulk) {
        const u32 bar = (u32)__cvta_generic_to_shared(mbar);
        u32 done = 0;
        while (!done) {
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(bar) : "memory");
        }
    }
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
        for (int k = 0; k < FUSED_CPT; k++) {
            if (m[k]) {                                              // mostly one newline per chunk at most
                const u32 b = __ffs(m[k]) - 1;
                T.seg_end[first[k]] = (u16)(16 * (tid + k * FUSED_NT) + b);
                const u32 rest = m[k] & (m[k] - 1);
                if (rest) T.put_lines(tid + k * FUSED_NT, rest, first[k] + 1);
            }
        }
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
            const u32 np = at > C.p0 && sh->live_lo == 0 ? (u32)(at - C.p0 < 64 ? at - C.p0 : 64) : 0u;
            const u8 *pv = prev + 64 - np;
            const u32 els = fastq ? (u32)(np && pv[np - 1] == '\n') : (u32)(e1 == FE_LS);
            sh->entry_ls = els;
            const u32 role0 = fastq ? (e1 & 3) : (e1 == FE_HDR ? (u32)FR_HDR : (u32)FR_SEQ);
            u32 sp = 0;
            if (sh->live_lo < sh->live_hi && role0 == FR_HDR && !els) {
                bool resolved; sp = prev_scan(pv, np, resolved);
                if (!resolved && at - np > C.p0) { u32 f = 0; sp = fast_lookback_space(A.text, C.p0, at - np, f); if (f) sh->flag |= FU_LOOKBACK; }
            }
            sh->entry_sp = sp;
        }
    }
    __syncthreads();

    u32 flag = 0;
    u64 maxlen = 0;
    if (!aborted) {
        // ---- phase 3: one thread per segment (round r: segment r * FUSED_NT + tid), block scans -> copy descriptors
        const u32 nseg = nl + 1;
        u64 sa[FUSED_SPT], sb[FUSED_SPT], pa[FUSED_SPT], pb[FUSED_SPT], tota = 0, totb = 0;
        u32 rb[FUSED_SPT], sp[FUSED_SPT];
#pragma unroll
        for (int r = 0; r < FUSED_SPT; r++) {
            sa[r] = sb[r] = pa[r] = pb[r] = 0; rb[r] = SR_NONE; sp[r] = 0xFFFF;
            if ((u32)r * FUSED_NT < nseg) {                                  // uniform across the CTA
                const u32 j = r * FUSED_NT + tid;
                if (j < nseg) T.classify(C, j, sa[r], sb[r], rb[r], sp[r], flag);
                u64 ta, tb;
                pa[r] = tota + block_excl_scan(sa[r], &ta, scan);
                pb[r] = totb + block_excl_scan(sb[r], &tb, scan);
                tota += ta; totb += tb;
            }
        }
        if (tid == 0) {
            sh->t_ids = (u32)(tota & 0xFFFF); sh->t_comm = (u32)((tota >> 16) & 0xFFFF); sh->t_seq = (u32)((tota >> 32) & 0xFFFF); sh->t_qual = (u32)(tota >> 48);
            sh->t_rec = (u32)(totb & 0xFFFF); sh->n_hdr = (u32)((totb >> 16) & 0xFFFF); sh->n_seq = (u32)((totb >> 32) & 0xFFFF); sh->n_qual = (u32)(totb >> 48);
            T.layout();
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < FUSED_SPT; r++) { const u32 j = r * FUSED_NT + tid; if (j < nseg) T.place(C, j, pa[r], pb[r], rb[r], sp[r]); }
        __syncthreads();
    }
    // ---- phase 4: warp 0 runs look-back #2 (global offsets; the record / line / byte that straddles the tile boundary)
    //      while the other warps gather the tile's bytes per stream into the staging area
    if (warp == 0) {
        F2 agg2 = f2_initial(); agg2.last = 0;
        if (!aborted) agg2 = T.aggregate(C);
        const F2 p2 = fused_lookback2(A, tile, agg2, fastq);
        if (lane == 0) sh->pre = p2;
    } else if (!aborted) {
        const u32 ndesc = sh->n_seq + sh->n_qual + 2 * sh->n_hdr, ngroups = (FUSED_NT - 32) / FT_GROUP;
        for (u32 k = (tid - 32) / FT_GROUP; k < ndesc; k += ngroups) flag |= T.copy_desc(C, k, tid % FT_GROUP);
    }
    __syncthreads();

    if (!aborted) {
        // ---- phase 5: staging -> global; records, lines
        const F2 &P = sh->pre;
        if (fused_region_out(T, sh->s_qual, sh->t_qual, C.qual + P.qual, FC_QUAL, false)) flag |= FU_QUAL;
        if (C.seq_mode == FS_PACK4) {
            if (sh->t_seq) {
                const u32 npieces = ((u32)(P.seq & 31) + sh->t_seq + 31) / 32;
                for (u32 q = tid; q < npieces; q += FUSED_NT) flag |= T.pack_piece(C, q, [](u32 *p, u32 v) { atomicOr(p, v); });
            }
        } else if (fused_region_out(T, sh->s_seq, sh->t_seq, C.seq + P.seq, FC_PROTEIN + (C.seq_mode - FS_PROTEIN), C.upper != 0)) flag |= FU_SEQ;
        fused_region_out(T, sh->s_ids, sh->t_ids, C.ids + P.ids, FC_NONE, false);
        fused_region_out(T, sh->s_comm, sh->t_comm, C.comm + P.comm, FC_NONE, false);
        for (u32 k = tid; k < sh->t_rec; k += FUSED_NT) { const u64 L = T.finish_record(C, k, flag); if (L > maxlen) maxlen = L; }
        if (!fastq) for (u32 k = tid; k < sh->n_seq; k += FUSED_NT) { const u64 L = T.line_length(k); if (L > maxlen) maxlen = L; }
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
    if (A.ntiles) { f1 = (u32)A.st1[A.ntiles - 1]; f2_get(A.st2 + 4ull * (A.ntiles - 1), f2); }
    __shared__ u32 lut[256];
    for (int i = 0; i < 256; i++) lut[i] = nuc_lut32(A.lut8[i]);
    FusedCfg C = A.C; C.lut = lut;
    fused_finish(C, f1, f2, *A.flag, *A.longest, A.text, *A.totals);
}

}  // namespace nafg
