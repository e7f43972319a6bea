// common.cuh — context, device arena, CUDA executor and scan primitives shared by the kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <utility>
#include <vector>
#include "../../include/nafgpu.h"
#include "zstd_hd.cuh"

namespace nafz {
// what the host walk of one stream's block headers found (zstd_walk_stream, zstd_dec.cuh)
struct ZWalked { std::vector<ZBlockHead> blocks; std::vector<u32> regen; std::vector<std::pair<u64, u64>> skips; bool simple = false; u64 consumed = 0; int rc = 0; std::string err; };
}

namespace nafg {

using nafz::u8; using nafz::u16; using nafz::u32; using nafz::u64; using nafz::i32; using nafz::i64;

struct CudaError { cudaError_t e; const char *what; const char *file; int line; };

#define CUDA_TRY(expr)                                                                      \
    do {                                                                                    \
        cudaError_t e_ = (expr);                                                            \
        if (e_ != cudaSuccess) throw nafg::CudaError{e_, #expr, __FILE__, __LINE__};        \
    } while (0)

struct NafError { int code; std::string msg; };
[[noreturn]] inline void fail(int code, const std::string &msg) { throw NafError{code, msg}; }

// ------------------------------------------------------------------ device arena
// Bump allocator over cudaMalloc'd slabs.  reset() at the start of every call; if a call needed more
// than one slab they are merged into a single bigger one afterwards, so steady state is one slab and
// zero cudaMalloc calls per encode/decode.
struct Arena {
    struct Slab { u8 *p; size_t cap; size_t used; };
    std::vector<Slab> slabs;
    size_t high_water = 0, cur_total = 0;

    void *alloc_bytes(size_t n)
    {
        n = (n + 255) & ~(size_t)255;
        if (n == 0) n = 256;
        void *r = nullptr;
        for (auto &s : slabs)
            if (s.cap - s.used >= n) { r = s.p + s.used; s.used += n; break; }
        if (!r) {
            size_t cap = n > (size_t)(256u << 20) ? n : (size_t)(256u << 20);
            Slab s; s.cap = cap; s.used = n;
            CUDA_TRY(cudaMalloc(&s.p, cap));                     // a request the device cannot satisfy (a corrupt file claiming
            slabs.push_back(s);                                  // terabytes) fails HERE, before it can enter the high-water mark
            r = s.p;
        }
        cur_total += n;
        if (cur_total > high_water) high_water = cur_total;
        return r;
    }
    void reset()
    {
        if (slabs.size() > 1 || (slabs.size() == 1 && slabs[0].cap < high_water)) {
            for (auto &s : slabs) cudaFree(s.p);
            slabs.clear();
            Slab s; s.cap = high_water + (high_water >> 3) + (1u << 20); s.used = 0;
            if (cudaMalloc(&s.p, s.cap) == cudaSuccess) slabs.push_back(s);
            else { cudaGetLastError(); high_water = 0; }         // no room for the merged slab: start over with on-demand slabs
        }
        for (auto &s : slabs) s.used = 0;
        cur_total = 0;
    }
    void release() { for (auto &s : slabs) cudaFree(s.p); slabs.clear(); }
    // roll back to an earlier allocation state (an abandoned attempt's scratch is reused by the retry)
    struct Mark { std::vector<size_t> used; size_t total; };
    Mark mark() const { Mark m; m.total = cur_total; for (auto &s : slabs) m.used.push_back(s.used); return m; }
    void rewind(const Mark &m)
    {
        for (size_t i = 0; i < slabs.size(); i++) slabs[i].used = i < m.used.size() ? m.used[i] : 0;
        cur_total = m.total;
    }
};

struct PinnedBuf {
    u8 *p = nullptr; size_t cap = 0;
    u8 *ensure(size_t n)
    {
        if (n <= cap) return p;
        if (p) cudaFreeHost(p);
        cap = n + (n >> 3) + 4096; p = nullptr;
        CUDA_TRY(cudaHostAlloc(&p, cap, cudaHostAllocDefault));
        return p;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

// ------------------------------------------------------------------ host <-> device pipe of the host-buffer calls
// The reference streams its input through 16 KB / 1 MB windows (ennaf/src/process.c:227-240, compressor.c:120-143) and its
// output through 128 KB ones (unnaf/src/output.c:640-650).  Here the counterpart is copy / compute overlap: the host
// buffer goes up in chunks on its own stream, each chunk followed by an event the compute stream waits for right before
// the first kernel that reads it, and finished pieces of the result go down on a third stream while the kernels for the
// next piece run.  PCIe is full duplex, so the upload of a call and the download of its result overlap as well.
struct HostPipe {
    cudaStream_t in = nullptr, out = nullptr;
    std::vector<cudaEvent_t> pool; size_t used = 0;
    // input
    u64 n_in = 0, chunk = 0; std::vector<cudaEvent_t> in_ev; bool uploading = false; const u8 *h_in = nullptr;
    // output
    u8 *h_out = nullptr; u64 out_done = 0; bool emitting = false;
    // output delivered to a callback instead of one big host buffer: two page-locked buffers take turns, the callback gets
    // piece k while piece k + 1 comes down
    int (*sink)(void *, const u8 *, size_t) = nullptr; void *sink_user = nullptr; bool sink_failed = false;
    static const u64 ROT = 32ull << 20;
    u8 *rot[2] = {nullptr, nullptr}; cudaEvent_t rot_ev[2] = {nullptr, nullptr}; u64 rot_len[2] = {0, 0}; bool rot_busy[2] = {false, false}; int rot_next = 0;

    void create()
    {
        CUDA_TRY(cudaStreamCreateWithFlags(&in, cudaStreamNonBlocking));
        CUDA_TRY(cudaStreamCreateWithFlags(&out, cudaStreamNonBlocking));
    }
    void rot_create()
    {
        for (int b = 0; b < 2; b++) if (!rot[b]) {
            CUDA_TRY(cudaHostAlloc(&rot[b], ROT, cudaHostAllocDefault));
            CUDA_TRY(cudaEventCreateWithFlags(&rot_ev[b], cudaEventDisableTiming));
        }
    }
    void rot_flush(int b)
    {
        if (!rot_busy[b]) return;
        CUDA_TRY(cudaEventSynchronize(rot_ev[b]));
        rot_busy[b] = false;
        if (sink && !sink_failed && sink(sink_user, rot[b], rot_len[b]) != 0) sink_failed = true;
    }
    void rot_flush_all() { rot_flush(rot_next); rot_flush(rot_next ^ 1); }
    void destroy()
    {
        for (int b = 0; b < 2; b++) { if (rot[b]) cudaFreeHost(rot[b]); if (rot_ev[b]) cudaEventDestroy(rot_ev[b]); rot[b] = nullptr; rot_ev[b] = nullptr; }
        for (auto e : pool) cudaEventDestroy(e);
        pool.clear();
        if (in) cudaStreamDestroy(in);
        if (out) cudaStreamDestroy(out);
        in = out = nullptr;
    }
    void reset()
    {
        used = 0; in_ev.clear(); n_in = 0; uploading = false; h_in = nullptr; h_out = nullptr; out_done = 0; emitting = false;
        deferred = ordered = false; def_d = nullptr; def_compute = nullptr; ord_lo.clear(); ord_hi.clear(); ord_ev.clear();
        sink = nullptr; sink_user = nullptr; sink_failed = false; rot_busy[0] = rot_busy[1] = false; rot_next = 0;
    }
    cudaEvent_t event()
    {
        if (used == pool.size()) { cudaEvent_t e; CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); pool.push_back(e); }
        return pool[used++];
    }
    // enqueue the whole upload now, in chunks; `after`: an event of the compute stream the copies must not overtake (the
    // destination may still be in use by what that stream ran before)
    void upload(u8 *d, const u8 *h, u64 n, u64 chunk_bytes, cudaStream_t compute)
    {
        cudaEvent_t e0 = event();
        CUDA_TRY(cudaEventRecord(e0, compute));
        CUDA_TRY(cudaStreamWaitEvent(in, e0, 0));
        n_in = n; chunk = chunk_bytes; uploading = true; h_in = h;
        for (u64 off = 0; off < n; off += chunk) {
            const u64 len = n - off < chunk ? n - off : chunk;
            CUDA_TRY(cudaMemcpyAsync(d + off, h + off, len, cudaMemcpyHostToDevice, in));
            cudaEvent_t e = event();
            CUDA_TRY(cudaEventRecord(e, in));
            in_ev.push_back(e);
        }
    }
    size_t chunks() const { return in_ev.size(); }
    // An upload whose ORDER is decided later: a decoder that has read the headers may want the two big streams of a FASTQ file
    // piece by piece in turns (sequence piece 0, quality piece 0, sequence piece 1 ...), so that the first records can be
    // written -- and their text sent down -- while most of the file is still on the host.  start_upload(nullptr) = front to back.
    bool deferred = false, ordered = false; u8 *def_d = nullptr; cudaStream_t def_compute = nullptr;
    std::vector<u64> ord_lo, ord_hi; std::vector<cudaEvent_t> ord_ev;           // in the order the copies were queued
    void defer_upload(u8 *d, const u8 *h, u64 n, u64 chunk_bytes, cudaStream_t compute)
    {
        def_d = d; h_in = h; n_in = n; chunk = chunk_bytes; def_compute = compute; deferred = true; uploading = true;
    }
    void start_upload(const std::vector<std::pair<u64, u64>> *order)
    {
        if (!deferred) return;
        deferred = false;
        if (!order) { upload(def_d, h_in, n_in, chunk, def_compute); return; }
        cudaEvent_t e0 = event();
        CUDA_TRY(cudaEventRecord(e0, def_compute));
        CUDA_TRY(cudaStreamWaitEvent(in, e0, 0));
        ordered = true;
        for (auto &r : *order)
            for (u64 off = r.first; off < r.second; off += chunk) {
                const u64 len = r.second - off < chunk ? r.second - off : chunk;
                CUDA_TRY(cudaMemcpyAsync(def_d + off, h_in + off, len, cudaMemcpyHostToDevice, in));
                cudaEvent_t e = event();
                CUDA_TRY(cudaEventRecord(e, in));
                ord_lo.push_back(off); ord_hi.push_back(off + len); ord_ev.push_back(e);
            }
    }
    // the compute stream's next kernels may read input bytes [lo, hi)
    void wait_range(cudaStream_t compute, u64 lo, u64 hi)
    {
        if (deferred) start_upload(nullptr);
        if (!uploading || hi <= lo) return;
        if (!ordered) { wait_input(compute, hi); return; }
        for (size_t i = ord_ev.size(); i-- > 0;)                                  // the last copy queued that touches the range: the earlier ones are done by then
            if (ord_lo[i] < hi && ord_hi[i] > lo) { CUDA_TRY(cudaStreamWaitEvent(compute, ord_ev[i], 0)); return; }
    }
    // the compute stream's next kernels may read input bytes [0, hi)
    void wait_input(cudaStream_t compute, u64 hi)
    {
        if (deferred) start_upload(nullptr);
        if (ordered) { wait_range(compute, 0, hi); return; }
        if (!uploading || in_ev.empty() || hi == 0) return;
        size_t c = (size_t)((hi - 1) / chunk);
        if (c >= in_ev.size()) c = in_ev.size() - 1;
        CUDA_TRY(cudaStreamWaitEvent(compute, in_ev[c], 0));
    }
    void wait_all_input(cudaStream_t compute) { wait_input(compute, n_in); }
    // result bytes [off, off + len) (device pointer d) are final once the compute stream gets here: send them down
    void begin_output(u8 *h) { h_out = h; out_done = 0; emitting = true; }
    void emit(cudaStream_t compute, const u8 *d, u64 off, u64 len)
    {
        if (!len) return;
        cudaEvent_t e = event();
        CUDA_TRY(cudaEventRecord(e, compute));
        CUDA_TRY(cudaStreamWaitEvent(out, e, 0));
        if (sink) {                                              // in order, through the two rotating buffers
            for (u64 at = 0; at < len; at += ROT) {
                const u64 m = len - at < ROT ? len - at : ROT;
                const int b = rot_next;
                rot_flush(b);                                    // the piece this buffer still holds goes to the callback first
                CUDA_TRY(cudaMemcpyAsync(rot[b], d + at, m, cudaMemcpyDeviceToHost, out));
                CUDA_TRY(cudaEventRecord(rot_ev[b], out));
                rot_busy[b] = true; rot_len[b] = m; rot_next ^= 1;
            }
            if (off == out_done) out_done = off + len;
            return;
        }
        CUDA_TRY(cudaMemcpyAsync(h_out + off, d, len, cudaMemcpyDeviceToHost, out));
        if (off == out_done) out_done = off + len;
    }
    void drain() { if (in) cudaStreamSynchronize(in); if (out) cudaStreamSynchronize(out); }
};

// ------------------------------------------------------------------ generic kernels for HD bodies
template <class F> __global__ void k_for_each(size_t n, F f)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) f(i);
}
template <class F> __global__ void k_for_each_group(F f) { f((size_t)blockIdx.x, threadIdx.x, blockDim.x); }

// Small control transfers (a few counters down, a table of a few KB up) between the kernels of a call must not queue behind
// the bulk copies a piped host-buffer call keeps in flight on the copy engines -- each would wait for a whole 32-64 MB DMA.
// They go through a mailbox instead: page-locked host memory the GPU reads / writes directly from a tiny kernel.
__global__ void k_mail_copy(const u8 *src, u8 *dst, size_t n)
{
    const size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    if ((((uintptr_t)src | (uintptr_t)dst) & 15) == 0) {
        const size_t nv = n / 16;
        for (size_t i = i0; i < nv; i += stride) ((uint4 *)dst)[i] = ((const uint4 *)src)[i];
        for (size_t i = nv * 16 + i0; i < n; i += stride) dst[i] = src[i];
    } else for (size_t i = i0; i < n; i += stride) dst[i] = src[i];
}
struct Mailbox {
    static const size_t DOWN = 64 << 10, CAP = (64 << 10) + (1 << 20);
    u8 *p = nullptr; size_t up_used = DOWN;
    void create() { CUDA_TRY(cudaHostAlloc(&p, CAP, cudaHostAllocMapped | cudaHostAllocPortable)); }
    void destroy() { if (p) cudaFreeHost(p); p = nullptr; }
};

// Optional per-launch timing with CUDA events (bench.py's live roofline; off during timed steps).
struct Prof {
    bool on = false;
    struct Rec { const char *name; cudaEvent_t a, b; };
    std::vector<Rec> recs;
    std::vector<cudaEvent_t> pool;
    cudaEvent_t get() { if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; } cudaEvent_t e; cudaEventCreate(&e); return e; }
};

struct CudaExec {
    cudaStream_t stream;
    Arena *arena;
    Prof *prof = nullptr;
    u32 launches = 0;
    HostPipe *pipe = nullptr;                // host-buffer calls: chunked copies overlapped with the kernels (else nullptr)

    void prof_begin(const char *name) { if (prof && prof->on) { Prof::Rec r{name, prof->get(), prof->get()}; cudaEventRecord(r.a, stream); prof->recs.push_back(r); } }
    void prof_end() { if (prof && prof->on) cudaEventRecord(prof->recs.back().b, stream); launches++; }

    Mailbox *mail = nullptr;
    template <class T> T *alloc(size_t count)
    {
        // a count taken from a damaged header must not wrap the byte size into a small allocation
        if (count > ((size_t)1 << 46) / sizeof(T)) fail(NAFGPU_E_FORMAT, "size field exceeds anything this device could hold\n");
        return (T *)arena->alloc_bytes(sizeof(T) * (count ? count : 1));
    }
    void upload(void *dst, const void *src, size_t n)
    {
        if (!n) return;
        if (mail && n <= (256u << 10)) {                       // through the mailbox: the source may be reused as soon as this returns
            const size_t need = (n + 15) & ~(size_t)15;
            if (mail->up_used + need > Mailbox::CAP) { CUDA_TRY(cudaStreamSynchronize(stream)); mail->up_used = Mailbox::DOWN; }
            u8 *slot = mail->p + mail->up_used; mail->up_used += need;
            memcpy(slot, src, n);
            k_mail_copy<<<(unsigned)((n + 4095) / 4096), 256, 0, stream>>>(slot, (u8 *)dst, n);
            return;
        }
        CUDA_TRY(cudaMemcpyAsync(dst, src, n, cudaMemcpyHostToDevice, stream));
    }
    // upload of a bigger host array that lives in pageable memory: through the context's pinned staging so the DMA is
    // asynchronous and the source may be reused as soon as this returns
    PinnedBuf *staging = nullptr;
    void upload_staged(void *dst, const void *src, size_t n)
    {
        if (!n) return;
        if (!staging) { upload(dst, src, n); CUDA_TRY(cudaStreamSynchronize(stream)); return; }
        CUDA_TRY(cudaStreamSynchronize(stream));                 // the staging buffer may still feed an earlier copy
        u8 *p = staging->ensure(n);
        memcpy(p, src, n);
        if (pipe && pipe->uploading) k_mail_copy<<<(unsigned)(n / 65536 + 1), 256, 0, stream>>>(p, (u8 *)dst, n);   // not behind the bulk upload
        else CUDA_TRY(cudaMemcpyAsync(dst, p, n, cudaMemcpyHostToDevice, stream));
    }
    void download(void *dst, const void *src, size_t n)
    {
        if (mail && n && n <= Mailbox::DOWN) {
            k_mail_copy<<<1, 256, 0, stream>>>((const u8 *)src, mail->p, n);
            CUDA_TRY(cudaStreamSynchronize(stream));
            memcpy(dst, mail->p, n);
            mail->up_used = Mailbox::DOWN;                       // the stream is idle: every earlier upload slot is free again
            return;
        }
        if (n) CUDA_TRY(cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToHost, stream));
        CUDA_TRY(cudaStreamSynchronize(stream));
    }
    void zero(void *p, size_t n) { if (n) CUDA_TRY(cudaMemsetAsync(p, 0, n, stream)); }
    void fill(void *p, int v, size_t n) { if (n) CUDA_TRY(cudaMemsetAsync(p, v, n, stream)); }
    // threads: CTA size.  Thread-serial bodies with long per-item loops use small CTAs so that few
    // thousand items still spread over all 148 SMs.
    template <class F> void for_each(size_t n, F f, const char *name = "for_each", int threads = 256)
    {
        if (!n) return;
        prof_begin(name);
        k_for_each<<<(unsigned)((n + threads - 1) / threads), threads, 0, stream>>>(n, f);
        prof_end();
    }
    template <class F> void for_each_group(size_t ngroups, int threads, F f, const char *name = "for_each_group")
    {
        if (!ngroups) return;
        prof_begin(name);
        k_for_each_group<<<(unsigned)ngroups, threads, 0, stream>>>(f);
        prof_end();
    }
    void check() { CUDA_TRY(cudaGetLastError()); }
};

// ------------------------------------------------------------------ device-wide exclusive scan (u64 sums)
// Reduce-then-scan over tiles of SCAN_TILE items: (1) per-tile sums, (2) one CTA scans the tile sums,
// (3) per-tile local scan + tile prefix.  Input is a functor so predicates (byte == 0, unit != 255 ...)
// are scanned without being materialised.  out has n + 1 entries; out[n] = total.
static const int SCAN_THREADS = 256, SCAN_ITEMS = 16, SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ u64 warp_incl_scan(u64 v)
{
    const unsigned lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { u64 t = __shfl_up_sync(0xFFFFFFFFu, v, d); if (lane >= (unsigned)d) v += t; }
    return v;
}
// exclusive scan across the CTA of one value per thread; returns exclusive prefix, *total = CTA sum
__device__ __forceinline__ u64 block_excl_scan(u64 v, u64 *total, u64 *smem /* >= 33 */)
{
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
    u64 incl = warp_incl_scan(v);
    if (lane == 31) smem[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        u64 w = lane < nwarps ? smem[lane] : 0;
        u64 wi = warp_incl_scan(w);
        smem[lane] = wi - w;
        if (lane == 31) smem[32] = wi;
    }
    __syncthreads();
    u64 r = smem[warp] + incl - v;
    *total = smem[32];
    __syncthreads();
    return r;
}

template <class In> __global__ void k_scan_reduce(In in, size_t n, u64 *tile_sums)
{
    __shared__ u64 sm[33];
    size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
    u64 s = 0;
#pragma unroll 4
    for (int k = 0; k < SCAN_ITEMS; k++) if (base + k < n) s += in(base + k);
    u64 total; block_excl_scan(s, &total, sm);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}
__global__ void k_scan_tiles(u64 *tile_sums, size_t ntiles, u64 *grand_total);
template <class In> __global__ void k_scan_apply(In in, size_t n, const u64 *tile_prefix, u64 *out)
{
    __shared__ u64 sm[33];
    size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
    u64 v[SCAN_ITEMS]; u64 s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) { v[k] = base + k < n ? in(base + k) : 0; s += v[k]; }
    u64 total; u64 p = block_excl_scan(s, &total, sm) + tile_prefix[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) { if (base + k < n) out[base + k] = p; p += v[k]; }
    if (base <= n && n < base + SCAN_ITEMS) out[n] = p - 0;      // thread owning position n writes the total
}

// a few thousand items (the records of a genome, the blocks of a range): one CTA, one launch instead of three
template <class In> __global__ void __launch_bounds__(1024) k_scan_small(In in, size_t n, u64 *out)
{
    __shared__ u64 sm[33];
    const size_t per = (n + 1023) / 1024, lo = (size_t)threadIdx.x * per, hi = lo + per < n ? lo + per : n;
    u64 s = 0;
    for (size_t i = lo; i < hi; i++) s += in(i);
    u64 total; u64 p = block_excl_scan(s, &total, sm);
    for (size_t i = lo; i < hi; i++) { out[i] = p; p += in(i); }
    if (threadIdx.x == 0) out[n] = total;
}

template <class In> void exclusive_scan(CudaExec &ex, In in, size_t n, u64 *out)
{
    if (n <= 16384) {
        ex.prof_begin("scan");
        k_scan_small<<<1, 1024, 0, ex.stream>>>(in, n, out);
        ex.prof_end();
        return;
    }
    size_t ntiles = (n + SCAN_TILE) / SCAN_TILE;          // >= 1, and covers index n
    u64 *tiles = ex.alloc<u64>(ntiles + 1);
    ex.prof_begin("scan");
    k_scan_reduce<<<(unsigned)ntiles, SCAN_THREADS, 0, ex.stream>>>(in, n, tiles);
    k_scan_tiles<<<1, 1024, 0, ex.stream>>>(tiles, ntiles, tiles + ntiles);
    k_scan_apply<<<(unsigned)ntiles, SCAN_THREADS, 0, ex.stream>>>(in, n, tiles, out);
    ex.prof_end(); ex.launches += 2;
}

// tables.c:189 nuc_code, with bit 7 = not an expected code for the alphabet (tables.c:72 DNA, :82 RNA)
inline void build_nuc_lut(int seq_type, u8 lut[256])
{
    for (int c = 0; c < 256; c++) {
        int u = (c >= 'a' && c <= 'z') ? c - 32 : c;
        const char *order = "-TGKCYSBAWRDMHV"; const char *q = u ? strchr(order, u) : nullptr;
        lut[c] = u == 'U' ? 1 : (q ? (u8)(q - order) : 15);
        const char *ok = seq_type == NAFGPU_RNA ? "-ABCDGHKMNRSUVWY" : "-ABCDGHKMNRSTVWY";
        if (!(u && strchr(ok, u))) lut[c] |= 0x80;
    }
}

// ------------------------------------------------------------------ context
struct Ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    Arena arena;
    PinnedBuf pinned_out, pinned_aux, pinned_stage;
    HostPipe pipe;
    Mailbox mail;
    // text arriving in pieces (nafgpu_encode_begin .. _end): its device buffer lives outside the per-call arena
    struct Ingest {
        bool active = false; nafgpu_enc_opts opts{}; std::string title; bool has_title = false;
        u8 *d_text = nullptr; size_t cap = 0, n = 0;
        u8 *rot[2] = {nullptr, nullptr}; cudaEvent_t ev[2] = {nullptr, nullptr}; bool busy[2] = {false, false}; int cur = 0;
    } ingest;
    std::vector<nafz::ZBlockHead> zblock_cache;   // keeps the capacity of the decoder's host block list between calls
    nafz::ZWalked zwalk[6];                  // per section: the host walk of its block headers (capacity kept between calls)
    u32 *d_predef = nullptr;                 // predefined FSE tables
    u8 *d_nuc_lut = nullptr;                 // nuc_code + "unexpected" bit: DNA at 0, RNA at 256
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    // side stream (higher priority) for the latency-bound thread-per-block kernels of the text-like streams: they run next to
    // the throughput kernels of the two big streams instead of in front of them
    cudaStream_t side = nullptr; cudaEvent_t side_fork = nullptr, side_join = nullptr;
    // Host-buffer encode: the two big streams (sequence, quality) are compressed WHILE the text is still arriving -- every few
    // uploaded chunks the blocks that have become complete go through the Huffman kernels on the side stream (zstd_enc.cu
    // zenc_early_*), so that only the last blocks and the small streams are left when the upload ends.
    struct EarlyZ {
        bool on = false;
        const u8 *src[2] = {nullptr, nullptr};   // sequence, quality stream
        void *blk[2] = {nullptr, nullptr};       // ZEncBlock[max blocks]
        u8 *slots[2] = {nullptr, nullptr};
        u64 max_blocks[2] = {0, 0}, done[2] = {0, 0};
    } early;
    std::string err;
    nafgpu_timing timing{};
    std::vector<u8> host_scratch;
    Prof prof;
    std::string prof_report;
    u64 fast_fallbacks = 0;                  // encode calls that had to be redone by the general parser
    // a shard between nafgpu_shard_begin and nafgpu_shard_finish / _fetch (device pointers into the arena, which is kept)
    struct Shard {
        bool active = false, finished = false;
        nafgpu_enc_opts opts{};
        u8 *stream[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr}; u64 raw[6] = {0, 0, 0, 0, 0, 0};
        u64 n_bases = 0, n_records = 0, longest = 0, n_flips = 0;
        const u64 *flip_pos = nullptr;
        int store_mask = 0, store_qual = 0, format = 0; u32 first_case = 0;
        const u8 *body[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr}; u64 body_size[6] = {0, 0, 0, 0, 0, 0};
    } shard;
    bool keep_arena = false;                 // next guarded call must not reset the arena (shard in progress)
};

#define KLAUNCH(ex, name, ...) do { (ex).prof_begin(name); __VA_ARGS__; (ex).prof_end(); } while (0)

struct DecodeOut { const u8 *d_text; u64 size; };
struct EncodeOut { const u8 *d_naf; u64 size; };
struct SplitOut { const u8 *d[6]; u64 size[6]; };

}  // namespace nafg

struct nafgpu_ctx : nafg::Ctx {};
