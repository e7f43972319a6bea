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

static const int LIT_WARPS = 4, LIT_BLOCKS = LIT_WARPS * 8;     // 128 threads = 32 blocks x 4 streams per CTA
static const int LIT_TAB_ENTRIES = 16384;                        // u16 entries of decode tables staged per CTA (32 KB)

__global__ void __launch_bounds__(LIT_WARPS * 32) k_literals_smem(const ZDecArgs a)
{
    __shared__ __align__(16) u16 tabs[LIT_TAB_ENTRIES];
    __shared__ u32 tab_off[LIT_BLOCKS], tab_words[LIT_BLOCKS];
    __shared__ const u32 *tab_src[LIT_BLOCKS];
    const u32 tid = threadIdx.x, lane = tid & 31;
    const u32 first = blockIdx.x * LIT_BLOCKS;
    // warp 0: which tables this CTA needs and where they go.  Blocks that share a table with their predecessor
    // (treeless literals of reference-made frames) share the staged copy; what does not fit stays in HBM / L1.
    if (tid < 32) {
        const u32 i = first + lane;
        i32 src = -1; u32 need = 0;
        if (i < a.nblk) {
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
        tab_off[lane] = need && lead_lane >= 0 ? lead_off : 0xFFFFFFFFu;
        tab_words[lane] = staged ? (need + 1) / 2 : 0;
        tab_src[lane] = src >= 0 ? (const u32 *)(a.huf_pool + (size_t)a.blk[src].huf_slot * HUF_SLOT_ENTRIES) : nullptr;
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
    if (i < a.nblk) { const u32 off = tab_off[tid / 4]; k_literals(a, i * 4 + (tid & 3), off == 0xFFFFFFFFu ? nullptr : tabs + off); }
}

inline void launch_literals(nafg::CudaExec &ex, const ZDecArgs &a)
{
    if (!a.nblk) return;
    ex.prof_begin("zd_literals");
    k_literals_smem<<<(a.nblk + LIT_BLOCKS - 1) / LIT_BLOCKS, LIT_WARPS * 32, 0, ex.stream>>>(a);
    ex.prof_end();
}

}  // namespace nafz
