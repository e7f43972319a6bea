// zstd_dec_cuda.cuh — CUDA-only pieces of the block-parallel zstd decoder.
//
// k_literals_smem: Huffman literal decode (replaces decompress/huf_decompress.c:350
// HUF_decompress4X1_usingDTable_internal_body).  One warp handles 8 blocks = 32 Huffman streams, one
// stream per lane.  The 8 decode tables (<= 4 KB each) are first copied from the HBM table pool into
// shared memory, so the per-symbol lookup is a bank-parallel LDS instead of 32 different global cache
// lines per instruction; bit-buffer reloads are unconditional and lock-step across lanes (no divergence).
#pragma once
#include "common.cuh"
#include "zstd_dec.cuh"

namespace nafz {

static const int LIT_BLOCKS_PER_CTA = 8;

__global__ void __launch_bounds__(32) k_literals_smem(const ZDecArgs a)
{
    __shared__ __align__(16) u16 tabs[LIT_BLOCKS_PER_CTA][HUF_SLOT_ENTRIES];
    const u32 lane = threadIdx.x;
    const u32 first = blockIdx.x * LIT_BLOCKS_PER_CTA;
    for (int j = 0; j < LIT_BLOCKS_PER_CTA; j++) {
        const u32 i = first + j;
        if (i >= a.nblk) break;
        const ZBlock &b = a.blk[i];
        if (b.type != 2 || b.lit_type < 2 || b.huf_src < 0) continue;
        const ZBlock &hb = a.blk[b.huf_src];
        if (hb.huf_bits == 0) continue;
        const u32 words = (1u << hb.huf_bits) / 2 > 0 ? (1u << hb.huf_bits) / 2 : 1;       // u16 entries -> u32 words
        const u32 *src = (const u32 *)(a.huf_pool + (size_t)hb.huf_slot * HUF_SLOT_ENTRIES);
        u32 *dst = (u32 *)tabs[j];
        for (u32 k = lane; k < words; k += 32) dst[k] = src[k];
    }
    __syncwarp();
    const u32 i = first + lane / 4;
    if (i < a.nblk) k_literals(a, i * 4 + (lane & 3), tabs[lane / 4]);
}

inline void launch_literals(nafg::CudaExec &ex, const ZDecArgs &a)
{
    if (!a.nblk) return;
    ex.prof_begin("zd_literals");
    k_literals_smem<<<(a.nblk + LIT_BLOCKS_PER_CTA - 1) / LIT_BLOCKS_PER_CTA, 32, 0, ex.stream>>>(a);
    ex.prof_end();
}

}  // namespace nafz
