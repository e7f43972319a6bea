// zstd_dec_cuda.cuh — CUDA-only pieces of the block-parallel zstd decoder.
//
// k_literals_smem: Huffman literal decode (replaces decompress/huf_decompress.c:350
// HUF_decompress4X1_usingDTable_internal_body).  One CTA of 4 warps handles 32 blocks = 128 Huffman streams, one
// stream per lane.  The decode tables the CTA needs (2 << maxbits bytes each, shared between blocks that inherit
// a table) are first copied from the HBM table pool into 32 KB of shared memory, so the per-symbol lookup is a bank-parallel LDS instead of 32 different global cache
// lines per instruction; bit-buffer reloads are unconditional and lock-step across lanes (no divergence).
#pragma once
#include "common.cuh"
#include "zstd_dec.cuh"

namespace nafz {

// LIT_WARPS = 4: 128 threads = 32 blocks x 4 streams per CTA, the shape for files with tens of thousands of blocks (their tables
// are 1 KB or less -- our encoder limits code lengths by alphabet size -- so all 32 fit the 32 KB of staging).
// LIT_WARPS = 1: 8 blocks per CTA, for small batches (a range of a multi-GPU decode, the small streams of a file): those are
// bound by the latency of ONE stream, so occupancy is irrelevant, and 8 tables fit even at the format's 11-bit maximum (4 KB
// each) -- the mask stream's run lengths are near-uniform bytes and do get long codes.
static const int LIT_TAB_ENTRIES = 16384;                        // u16 entries of decode tables staged per CTA (32 KB)

template <int LIT_WARPS> __global__ void __launch_bounds__(LIT_WARPS * 32) k_literals_smem(const ZDecArgs a)
{
    constexpr int LIT_BLOCKS = LIT_WARPS * 8;
    __shared__ __align__(16) u16 tabs[LIT_TAB_ENTRIES];
    __shared__ __align__(16) uint4 ring[HUF_RING * LIT_WARPS * 32];   // slot s of thread t: ring[s * 128 + t] (zstd_hd.cuh: BackBitsR)
    __shared__ u32 tab_off[LIT_BLOCKS], tab_words[LIT_BLOCKS];
    __shared__ const u32 *tab_src[LIT_BLOCKS];
    const u32 tid = threadIdx.x, lane = tid & 31;
    const u32 first = blockIdx.x * LIT_BLOCKS;
    // warp 0: which tables this CTA needs and where they go.  Blocks that share a table with their predecessor
    // (treeless literals of reference-made frames) share the staged copy; what does not fit stays in HBM / L1.
    if (tid < 32) {
        const u32 i = first + lane;
        i32 src = -1; u32 need = 0;
        if (i < a.nblk && lane < (u32)LIT_BLOCKS) {
            const ZBlock &b = a.blk[i];
            if (b.type == 2 && b.lit_type >= 2 && !b.skip && b.huf_src >= 0 && a.blk[b.huf_src].huf_bits) { src = b.huf_src; need = 1u << a.blk[b.huf_src].huf_bits; }
        }
        const i32 prev = __shfl_up_sync(0xFFFFFFFFu, src, 1);
        const bool leader = need && (lane == 0 || prev != src);
        u32 incl = leader ? need : 0;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const u32 t = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= (u32)d) incl += t; }
        const u32 off = incl - (leader ? need : 0);
        const bool staged = leader && off + need <= (u32)LIT_TAB_ENTRIES;
        // followers take the offset of the nearest leader before them
        int lead_lane = leader ? (int)lane : -1;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xFFFFFFFFu, lead_lane, d); if (lane >= (u32)d && t > lead_lane) lead_lane = t; }
        const u32 lead_off = __shfl_sync(0xFFFFFFFFu, staged ? off : 0xFFFFFFFFu, lead_lane < 0 ? 0 : lead_lane);
        if (lane < (u32)LIT_BLOCKS) {
            tab_off[lane] = need && lead_lane >= 0 ? lead_off : 0xFFFFFFFFu;
            tab_words[lane] = staged ? (need + 1) / 2 : 0;
            tab_src[lane] = src >= 0 ? (const u32 *)(a.huf_pool + (size_t)a.blk[src].huf_slot * HUF_SLOT_ENTRIES) : nullptr;
        }
    }
    __syncthreads();
    for (int j = 0; j < LIT_BLOCKS; j++) {
        const u32 words = tab_words[j];
        if (!words) continue;
        const u32 *src = tab_src[j];
        u32 *dst = (u32 *)(tabs + tab_off[j]);
        for (u32 k = tid; k < words; k += LIT_WARPS * 32) dst[k] = src[k];
    }
    __syncthreads();
    const u32 i = first + tid / 4;
    if (i < a.nblk) {
        const u32 off = tab_off[tid / 4];
        k_literals(a, i * 4 + (tid & 3), off == 0xFFFFFFFFu ? nullptr : tabs + off, (u32)__cvta_generic_to_shared(ring + tid), (u32)(LIT_WARPS * 32 * sizeof(uint4)));
    }
}

inline void launch_literals(nafg::CudaExec &ex, const ZDecArgs &a)
{
    if (!a.nblk) return;
    ex.prof_begin("zd_literals");
    if (a.nblk <= 8192) k_literals_smem<1><<<(a.nblk + 7) / 8, 32, 0, ex.stream>>>(a);
    else k_literals_smem<4><<<(a.nblk + 31) / 32, 128, 0, ex.stream>>>(a);
    ex.prof_end();
}


// ---- sequences of reference-made frames: the three serial-per-block kernels, with their working sets in shared memory.
// One block of ids / comments can carry tens of thousands of sequences and the kernel takes as long as its longest
// block, so what counts is the latency per sequence: FSE states walk 5 KB of tables (L2-latency loads when left in
// HBM), repeat offsets are a serial recurrence over fields scattered 24 bytes apart.

static const int SD_BLOCKS = 16;                                 // blocks per CTA of k_seq_decode_smem (one lane each)
static const u32 SD_MIN_SEQ = 64;                                // below that the tables are not worth staging

__global__ void __launch_bounds__(32) k_seq_decode_smem(const ZDecArgs a)
{
    extern __shared__ u32 sd_tabs[];                             // SD_BLOCKS x FSE_SLOT_ENTRIES
    __shared__ u8 staged[SD_BLOCKS];
    const u32 lane = threadIdx.x, first = blockIdx.x * SD_BLOCKS;
    for (int j = 0; j < SD_BLOCKS; j++) {
        const u32 i = first + j;
        bool st = false;
        if (i < a.nblk) {
            const ZBlock &b = a.blk[i];
            st = b.type == 2 && b.nseq >= SD_MIN_SEQ && b.ll_src >= 0 && b.of_src >= 0 && b.ml_src >= 0;
            if (st) {
                int l0, l1, l2;
                const u32 *t0 = fse_table_for(a, b.ll_src, 0, &l0), *t1 = fse_table_for(a, b.of_src, 1, &l1), *t2 = fse_table_for(a, b.ml_src, 2, &l2);
                u32 *d = sd_tabs + (size_t)j * FSE_SLOT_ENTRIES;
                for (u32 k = lane; k < (1u << l0); k += 32) d[k] = t0[k];
                for (u32 k = lane; k < (1u << l1); k += 32) d[FSE_OF_AT + k] = t1[k];
                for (u32 k = lane; k < (1u << l2); k += 32) d[FSE_ML_AT + k] = t2[k];
            }
        }
        if (lane == 0) staged[j] = st;
    }
    __syncwarp();
    if (lane < SD_BLOCKS && first + lane < a.nblk) {
        const u32 *d = sd_tabs + (size_t)lane * FSE_SLOT_ENTRIES;
        if (staged[lane]) k_seq_decode(a, first + lane, d, d + FSE_OF_AT, d + FSE_ML_AT);
        else k_seq_decode(a, first + lane);
    }
}

// one warp per block: 32 sequences' fields at a time into shared memory, lane 0 runs the repeat-offset recurrence over
// them, then every lane validates and stores its own sequence
__global__ void __launch_bounds__(128) k_seq_resolve_warp(const ZDecArgs a)
{
    __shared__ u32 s_of[4][32], s_ll[4][32];
    const u32 lane = threadIdx.x & 31, w = threadIdx.x >> 5, i = blockIdx.x * 4 + w;
    if (i >= a.nblk) return;
    const ZBlock &b = a.blk[i];
    if (b.type != 2 || b.nseq == 0) return;
    u32 r0 = b.rep_in[0], r1 = b.rep_in[1], r2 = b.rep_in[2];
    ZSeq *seq = a.seq + b.seq_base;
    const u64 out_off = b.out_off, frame_out = b.frame_out;
    bool bad = false;
    for (u32 k0 = 0; k0 < b.nseq; k0 += 32) {
        const u32 m = b.nseq - k0 < 32 ? b.nseq - k0 : 32;
        u32 ofv = 0, ll = 0, dr = 0;
        if (lane < m) { const ZSeq s = seq[k0 + lane]; ofv = s.of; ll = s.ll; dr = s.dst_rel; }
        s_of[w][lane] = ofv; s_ll[w][lane] = ll;
        __syncwarp();
        if (lane == 0) {
            for (u32 j = 0; j < m; j++) {
                const u32 v = s_of[w][j]; u32 off;
                if (v > 3) { off = v - 3; r2 = r1; r1 = r0; r0 = off; }
                else {
                    const u32 idx = v - 1 + (s_ll[w][j] == 0 ? 1u : 0u);
                    if (idx == 0) off = r0;
                    else {
                        off = idx == 1 ? r1 : (idx == 2 ? r2 : r0 - 1);
                        if (idx != 1) r2 = r1;
                        r1 = r0; r0 = off;
                    }
                }
                s_of[w][j] = off;
            }
        }
        __syncwarp();
        if (lane < m) {
            u32 off = s_of[w][lane];
            const u64 match_pos = out_off + dr + ll;
            if (off == 0 || off > match_pos - frame_out) { bad = true; off = 0; seq[k0 + lane].ml = 0; }
            seq[k0 + lane].of = off;
        }
        __syncwarp();
    }
    if (bad) zerr(a, Z_ERR_OFFSET, i);
}

// one CTA per block: find the long sequences in parallel, then copy each with the whole CTA; tail literals
__global__ void __launch_bounds__(256) k_seq_exec_big_cta(const ZDecArgs a)
{
    __shared__ u32 big[256];
    __shared__ u32 nbig;
    const u32 i = blockIdx.x, tid = threadIdx.x;
    const ZBlock &b = a.blk[i];
    if (b.type != 2 || b.nseq == 0 || b.local) return;           // local: K7b has written the whole block already
    const ZSeq *seq = a.seq + b.seq_base;
    for (u32 base = 0; base < b.nseq; base += 256) {
        if (tid == 0) nbig = 0;
        __syncthreads();
        const u32 k = base + tid;
        if (k < b.nseq && seq[k].ll + seq[k].ml > BIG_SEQ) big[atomicAdd(&nbig, 1u)] = k;
        __syncthreads();
        const u32 n = nbig;
        for (u32 j = 0; j < n; j++) k_seq_exec_one(a, seq[big[j]], tid, 256);
        __syncthreads();
    }
    const ZSeq &last = seq[b.nseq - 1];
    const u32 lit_used = last.lit_rel + last.ll, out_used = last.dst_rel + last.ll + last.ml;
    const u8 *lit = a.lit_scratch + b.lit_off + lit_used; u8 *o = a.out + b.out_off + out_used;
    for (u32 k = tid; k + lit_used < b.lit_regen; k += 256) o[k] = lit[k];
}

inline void launch_seq_decode(nafg::CudaExec &ex, const ZDecArgs &a)
{
    if (!a.nblk) return;
    const int smem = SD_BLOCKS * FSE_SLOT_ENTRIES * 4;
    cudaFuncSetAttribute(k_seq_decode_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);   // per device: set on every launch (cheap)
    ex.prof_begin("zd_seq_decode");
    k_seq_decode_smem<<<(a.nblk + SD_BLOCKS - 1) / SD_BLOCKS, 32, smem, ex.stream>>>(a);
    ex.prof_end();
}
inline void launch_seq_resolve(nafg::CudaExec &ex, const ZDecArgs &a)
{
    if (!a.nblk) return;
    ex.prof_begin("zd_seq_resolve");
    k_seq_resolve_warp<<<(a.nblk + 3) / 4, 128, 0, ex.stream>>>(a);
    ex.prof_end();
}
inline void launch_seq_exec_big(nafg::CudaExec &ex, const ZDecArgs &a)
{
    if (!a.nblk) return;
    ex.prof_begin("zd_seq_exec_big");
    k_seq_exec_big_cta<<<a.nblk, 256, 0, ex.stream>>>(a);
    ex.prof_end();
}

}  // namespace nafz
