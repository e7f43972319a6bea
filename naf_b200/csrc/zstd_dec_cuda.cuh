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

// ---- K4 as ONE WARP per block (replaces zstd_decompress_block.c:937 ZSTD_decodeSequence's loop).
// The FSE state chain is serial, so a block's sequences are decoded by one lane -- what the other 31 do is keep everything that
// lane touches on chip: the three decode tables (expanded so that an entry also carries the extra-bit count and base value
// of its symbol), the bitstream (a 1 KB shared-memory window that the warp refills 512 bytes at a time, coalesced, at points
// all lanes reach together), and the decoded sequences (staged eight at a time and stored by the whole warp).  One sequence
// is then a table lookup and a window read -- both from shared memory, both issued at once, since the window's position is
// known before the symbols are -- and a few shifts: ~80 cycles instead of the ~1,000 of the thread-per-block version, whose
// bit reader fetched from global memory (ncu, profiles/r2h: 10 ms for the ids / comments of 2 M reads written by ennaf).
static const int SDW_WARPS = 4;
struct SdwWindow { u32 v2, v1, v0; };      // 96 bits, the most significant bit of v2 is the stream's next bit

__device__ __forceinline__ SdwWindow sdw_window(const u32 *ring, uintptr_t bits_addr, int P)
{
    const int bitpos = P - 96;
    const int q = bitpos >> 3;                                   // floor, also when negative
    const u32 s = (u32)(bitpos - 8 * q);
    const uintptr_t A = bits_addr + (intptr_t)q;
    const u32 sh = (u32)(A & 3) * 8 + s;
    const u32 i0 = (u32)(A >> 2);
    const u32 w0 = ring[i0 & 255], w1 = ring[(i0 + 1) & 255], w2 = ring[(i0 + 2) & 255], w3 = ring[(i0 + 3) & 255];
    SdwWindow w;
    w.v0 = __funnelshift_r(w0, w1, sh); w.v1 = __funnelshift_r(w1, w2, sh); w.v2 = __funnelshift_r(w2, w3, sh);
    return w;
}
// the n bits (0 <= n <= 32) that start `skip` bits (0 <= skip <= 64) below the top of the window
__device__ __forceinline__ u32 sdw_field(const SdwWindow &w, u32 skip, u32 n)
{
    if (n == 0) return 0;
    u32 hi, lo;
    if (skip < 32) { hi = w.v2; lo = w.v1; } else if (skip < 64) { hi = w.v1; lo = w.v0; skip -= 32; } else { hi = w.v0; lo = 0; skip -= 64; }
    const u32 top = __funnelshift_l(lo, hi, skip);               // skip < 32
    return top >> (32 - n);
}

__global__ void __launch_bounds__(SDW_WARPS * 32) k_seq_decode_warp(const ZDecArgs a)
{
    __shared__ uint2 tabs[SDW_WARPS][FSE_SLOT_ENTRIES];          // .x = FSE entry, .y = extra bits | value base << 8 | bad symbol << 31
    __shared__ u32 rings[SDW_WARPS][256];                        // bitstream window: word (addr >> 2) & 255 holds global bytes addr .. addr + 3
    __shared__ u32 stage[SDW_WARPS][48];                         // 8 sequences x 6 words
    const u32 lane = threadIdx.x & 31, w = threadIdx.x >> 5, i = blockIdx.x * SDW_WARPS + w;
    if (i >= a.nblk) return;
    ZBlock &b = a.blk[i];
    if (lane == 0) { b.repfn = repfn_identity(); b.match_total = 0; }
    if (b.type != 2 || b.nseq == 0) return;
    if (b.ll_src < 0 || b.of_src < 0 || b.ml_src < 0) { if (lane == 0) b.nseq = 0; return; }
    int ll_log, of_log, ml_log;
    const u32 *tl = fse_table_for(a, b.ll_src, 0, &ll_log), *to = fse_table_for(a, b.of_src, 1, &of_log), *tm = fse_table_for(a, b.ml_src, 2, &ml_log);
    uint2 *T = tabs[w]; u32 *ring = rings[w];
    for (u32 k = lane; k < (1u << ll_log); k += 32) { const u32 e = tl[k], c = fse_sym(e); T[k] = make_uint2(e, c > 35 ? 0x80000000u : (ll_bits_of(c) | (ll_base_of(c) << 8))); }
    for (u32 k = lane; k < (1u << of_log); k += 32) { const u32 e = to[k], c = fse_sym(e); T[FSE_OF_AT + k] = make_uint2(e, c > 31 ? 0x80000000u : c); }
    for (u32 k = lane; k < (1u << ml_log); k += 32) { const u32 e = tm[k], c = fse_sym(e); T[FSE_ML_AT + k] = make_uint2(e, c > 52 ? 0x80000000u : (ml_bits_of(c) | (ml_base_of(c) << 8))); }
    const u8 *bits = a.in + b.src + b.bits_off; const u32 n = b.csize - b.bits_off;
    const u8 last = n ? bits[n - 1] : 0;
    if (n == 0 || last == 0) { if (lane == 0) { zerr(a, Z_ERR_SEQ_STREAM, i); b.nseq = 0; } return; }
    const uintptr_t bits_addr = (uintptr_t)bits, end_addr = bits_addr + n, base_addr = (uintptr_t)a.in;
    uintptr_t glo = ((end_addr - 1) & ~(uintptr_t)511) - 512;    // window = [glo, glo + 1024), the stream's last byte in its upper half
    auto load_words = [&](uintptr_t from, u32 nwords) {          // aligned words; nothing is read below the input buffer or past the stream's last word
        for (u32 k = lane; k < nwords; k += 32) {
            const uintptr_t ad = from + 4 * (uintptr_t)k;
            u32 v = 0;
            if (ad >= base_addr && ad < ((end_addr + 3) & ~(uintptr_t)3)) v = *(const u32 *)ad;
            ring[(ad >> 2) & 255] = v;
        }
    };
    load_words(glo, 256);
    __syncwarp();
    int P = (int)(n * 8) - (8 - hibit(last));                   // payload bits not yet read
    u32 sl = 0, so = 0, sm = 0, bad = 0;
    u64 lit_total = 0, match_total = 0;
    RepFn rf = repfn_identity();
    if (lane == 0) {
        const SdwWindow win = sdw_window(ring, bits_addr, P);
        sl = sdw_field(win, 0, (u32)ll_log); so = sdw_field(win, (u32)ll_log, (u32)of_log); sm = sdw_field(win, (u32)(ll_log + of_log), (u32)ml_log);
        P -= ll_log + of_log + ml_log;
    }
    ZSeq *seq = a.seq + b.seq_base;
    const u32 nseq = b.nseq;
    for (u32 k0 = 0; k0 < nseq; k0 += 8) {
        const u32 m = nseq - k0 < 8 ? nseq - k0 : 8;
        if (lane == 0) {
            for (u32 j = 0; j < m; j++) {
                const SdwWindow win = sdw_window(ring, bits_addr, P);
                const uint2 el = T[sl], eo = T[FSE_OF_AT + so], em = T[FSE_ML_AT + sm];
                bad |= (el.y | eo.y | em.y) & 0x80000000u;
                const u32 oc = eo.y & 31, mlb = em.y & 0xFF, llb = el.y & 0xFF;
                const u32 c1 = oc, c2 = c1 + mlb, c3 = c2 + llb;
                const u32 ofv = (1u << oc) + sdw_field(win, 0, oc);
                const u32 ml = ((em.y >> 8) & 0x7FFFFF) + sdw_field(win, c1, mlb);
                const u32 ll = ((el.y >> 8) & 0x7FFFFF) + sdw_field(win, c2, llb);
                u32 used = c3;
                if (k0 + j + 1 < nseq) {
                    const u32 nl = fse_nb(el.x), nm = fse_nb(em.x), no = fse_nb(eo.x);
                    sl = fse_base(el.x) + sdw_field(win, c3, nl);
                    sm = fse_base(em.x) + sdw_field(win, c3 + nl, nm);
                    so = fse_base(eo.x) + sdw_field(win, c3 + nl + nm, no);
                    used += nl + nm + no;
                    sl &= 511; sm &= 511; so &= 255;                 // (a damaged table cannot send a state outside its table)
                }
                P -= (int)used;
                u32 *st = stage[w] + 6 * j;
                st[0] = ll; st[1] = ml; st[2] = ofv; st[3] = (u32)lit_total; st[4] = (u32)(lit_total + match_total); st[5] = i;
                repfn_step(rf, ofv, ll);
                lit_total += ll; match_total += ml;
            }
        }
        __syncwarp();
        P = __shfl_sync(0xFFFFFFFFu, P, 0);
        {   // eight sequences = 48 words, stored by the warp
            u32 *dst = (u32 *)(seq + k0);
            if (lane < 6 * m) dst[lane] = stage[w][lane];
            if (lane + 32 < 6 * m) dst[lane + 32] = stage[w][lane + 32];
        }
        if (P < 0) break;                                        // ran past the start of the stream: reported below
        // refill: once the cursor is in the lower half, the upper half takes the 512 bytes below the window
        const uintptr_t cur = bits_addr + (uintptr_t)((P > 0 ? P - 1 : 0) >> 3);
        if (cur < glo + 512 && glo + 16 > bits_addr) { load_words(glo - 512, 128); glo -= 512; }
        __syncwarp();
    }
    if (lane == 0) {
        if (bad || P != 0) { zerr(a, Z_ERR_SEQ_STREAM, i); b.nseq = 0; return; }
        if (lit_total > b.lit_regen || lit_total + match_total > 128 * 1024 + 0u) { zerr(a, Z_ERR_SIZE, i); b.nseq = 0; return; }
        b.match_total = (u32)match_total; b.repfn = rf;
    }
}

inline void launch_seq_decode(nafg::CudaExec &ex, const ZDecArgs &a)
{
    if (!a.nblk) return;
    ex.prof_begin("zd_seq_decode");
    k_seq_decode_warp<<<(a.nblk + SDW_WARPS - 1) / SDW_WARPS, SDW_WARPS * 32, 0, ex.stream>>>(a);
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
