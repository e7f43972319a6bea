// api.cu — the extern "C" boundary declared in include/nafgpu.h.
#include "common.cuh"
#include "container.hpp"
#include "zstd_dec.cuh"
#include "zstd_dec_cuda.cuh"

namespace nafg {
DecodeOut decode_on_device(Ctx &ctx, CudaExec &ex, const u8 *d_naf, const u8 *h_naf, size_t n, const nafgpu_dec_opts &o);
EncodeOut encode_on_device(Ctx &ctx, CudaExec &ex, const u8 *d_text, size_t n, const nafgpu_enc_opts &o, nafgpu_enc_info *info);
SplitOut split_on_device(Ctx &ctx, CudaExec &ex, const u8 *d_text, size_t n, const nafgpu_enc_opts &o, nafgpu_enc_info *info);
EncodeOut zstd_compress_on_device(Ctx &ctx, CudaExec &ex, const u8 *d_src, size_t n, int window_log, int level);
void record_cuts_on_device(Ctx &ctx, CudaExec &ex, const u8 *d_text, size_t n, int pieces, uint64_t *cuts);
void shard_begin_on_device(Ctx &ctx, CudaExec &ex, const u8 *d_text, size_t n, const nafgpu_enc_opts &o, nafgpu_shard_counts *counts, nafgpu_enc_info *info);
void shard_finish_on_device(Ctx &ctx, CudaExec &ex, const nafgpu_shard_link &link, uint64_t raw[6], uint64_t body[6]);
}

using namespace nafg;

static thread_local std::string g_create_error;

template <class F> static int guarded(nafgpu_ctx *ctx, F f)
{
    if (!ctx) return NAFGPU_E_ARG;
    try {
        ctx->err.clear(); ctx->fast_fallbacks = 0;
        CUDA_TRY(cudaSetDevice(ctx->device));
        if (ctx->side) cudaStreamSynchronize(ctx->side);        // nothing of an earlier (failed) call may still be using the arena
        ctx->early = Ctx::EarlyZ();
        if (!ctx->keep_arena) { ctx->arena.reset(); ctx->shard.active = false; }
        ctx->keep_arena = false;
        ctx->pipe.reset(); ctx->mail.up_used = Mailbox::DOWN;
        f();
        return NAFGPU_OK;
    } catch (const NafError &e) {
        ctx->err = e.msg; ctx->pipe.drain(); cudaStreamSynchronize(ctx->stream); ctx->timing.parser_fallback = ctx->fast_fallbacks != 0; return e.code;
    } catch (const CudaError &e) {
        char buf[512];
        snprintf(buf, sizeof buf, "CUDA error: %s (%s) at %s:%d\n", cudaGetErrorString(e.e), e.what, e.file, e.line);
        ctx->err = buf; ctx->pipe.drain(); cudaGetLastError(); return NAFGPU_E_CUDA;
    } catch (const std::exception &e) {
        ctx->err = std::string("internal error: ") + e.what() + "\n"; ctx->pipe.drain(); return NAFGPU_E_CUDA;
    }
}

extern "C" {

const char *nafgpu_version(void) { return NAFGPU_VERSION; }

int nafgpu_create(int device, nafgpu_ctx **out)
{
    if (!out) return NAFGPU_E_ARG;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        g_create_error = std::string("no CUDA device available: ") + cudaGetErrorString(e) + " (libnafgpu has no CPU path)\n";
        cudaGetLastError();
        return NAFGPU_E_CUDA;
    }
    if (device < 0) { if (cudaGetDevice(&device) != cudaSuccess) device = 0; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major < 10) {
        g_create_error = "device is not sm_100 (Blackwell): libnafgpu is built for sm_100a only\n";
        return NAFGPU_E_CUDA;
    }
    nafgpu_ctx *c = new nafgpu_ctx();
    c->device = device;
    try {
        CUDA_TRY(cudaSetDevice(device));
        CUDA_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        c->pipe.create(); c->mail.create();
        for (auto &ev : c->ev) CUDA_TRY(cudaEventCreate(&ev));
        {
            int lo = 0, hi = 0;
            CUDA_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi));
            CUDA_TRY(cudaStreamCreateWithPriority(&c->side, cudaStreamNonBlocking, hi));
            CUDA_TRY(cudaEventCreateWithFlags(&c->side_fork, cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&c->side_join, cudaEventDisableTiming));
        }
        u32 predef[nafz::FSE_SLOT_ENTRIES];
        nafz::zstd_build_predef(predef);
        CUDA_TRY(cudaMalloc(&c->d_predef, sizeof predef));
        CUDA_TRY(cudaMemcpy(c->d_predef, predef, sizeof predef, cudaMemcpyHostToDevice));
        u8 lut[512];
        build_nuc_lut(NAFGPU_DNA, lut); build_nuc_lut(NAFGPU_RNA, lut + 256);
        CUDA_TRY(cudaMalloc(&c->d_nuc_lut, sizeof lut));
        CUDA_TRY(cudaMemcpy(c->d_nuc_lut, lut, sizeof lut, cudaMemcpyHostToDevice));
    } catch (const CudaError &er) {
        g_create_error = std::string("CUDA error during context creation: ") + cudaGetErrorString(er.e) + "\n";
        delete c; return NAFGPU_E_CUDA;
    }
    *out = c;
    return NAFGPU_OK;
}

void nafgpu_destroy(nafgpu_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    c->pipe.drain(); c->pipe.destroy(); c->mail.destroy();
    for (int b = 0; b < 2; b++) { if (c->ingest.rot[b]) cudaFreeHost(c->ingest.rot[b]); if (c->ingest.ev[b]) cudaEventDestroy(c->ingest.ev[b]); }
    if (c->ingest.d_text) cudaFree(c->ingest.d_text);
    c->arena.release(); c->pinned_out.release(); c->pinned_aux.release(); c->pinned_stage.release();
    if (c->d_predef) cudaFree(c->d_predef);
    if (c->d_nuc_lut) cudaFree(c->d_nuc_lut);
    for (auto &ev : c->ev) if (ev) cudaEventDestroy(ev);
    if (c->side_fork) cudaEventDestroy(c->side_fork);
    if (c->side_join) cudaEventDestroy(c->side_join);
    if (c->side) cudaStreamDestroy(c->side);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

const char *nafgpu_last_error(const nafgpu_ctx *c) { return c ? c->err.c_str() : g_create_error.c_str(); }
int nafgpu_get_timing(const nafgpu_ctx *c, nafgpu_timing *t) { if (!c || !t) return NAFGPU_E_ARG; *t = c->timing; return 0; }
void *nafgpu_stream(nafgpu_ctx *c) { return c ? (void *)c->stream : nullptr; }

int nafgpu_profile(nafgpu_ctx *c, int enable) { if (!c) return NAFGPU_E_ARG; c->prof.on = enable != 0; return 0; }
const char *nafgpu_profile_report(const nafgpu_ctx *c) { return c ? c->prof_report.c_str() : ""; }

int nafgpu_host_alloc(size_t n, void **p) { return cudaHostAlloc(p, n ? n : 1, cudaHostAllocDefault) == cudaSuccess ? 0 : NAFGPU_E_CUDA; }
void nafgpu_host_free(void *p) { if (p) cudaFreeHost(p); }

}  // extern "C"

// Host-buffer calls on inputs of at least PIPE_MIN bytes overlap their copies with the kernels (HostPipe, common.cuh).
// NAFGPU_PIPE=0 in the environment turns that off (A/B measurements).
// NAFGPU_PIPE_MIN / NAFGPU_PIPE_CHUNK (bytes; the chunk is rounded to 64 KB) let the tests run the piped paths on small inputs.
static u64 env_bytes(const char *name, u64 dflt) { const char *e = getenv(name); return e && *e ? strtoull(e, nullptr, 10) : dflt; }
static u64 pipe_chunk() { u64 c = env_bytes("NAFGPU_PIPE_CHUNK", 32ull << 20) & ~0xFFFFull; return c ? c : 0x10000; }
static bool use_pipe(u64 n)
{
    const char *env = getenv("NAFGPU_PIPE");
    if (env && env[0] == '0') return false;
    return n >= env_bytes("NAFGPU_PIPE_MIN", 64ull << 20);
}

// host buffer -> device copy with 64 bytes of zero padding after it
static u8 *to_device(Ctx &c, CudaExec &ex, const u8 *h, size_t n)
{
    u8 *d = ex.alloc<u8>(n + 64);
    if (n) CUDA_TRY(cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, c.stream));
    CUDA_TRY(cudaMemsetAsync(d + n, 0, 64, c.stream));
    return d;
}
static const u8 *to_pinned(Ctx &c, const u8 *d, size_t n)
{
    u8 *h = c.pinned_out.ensure(n + 1);
    if (n) CUDA_TRY(cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, c.stream));
    return h;
}
static void finish_timing(Ctx &c, CudaExec &ex)
{
    CUDA_TRY(cudaEventRecord(c.ev[3], c.stream));
    CUDA_TRY(cudaStreamSynchronize(c.stream));
    CUDA_TRY(cudaStreamSynchronize(c.pipe.in));
    CUDA_TRY(cudaStreamSynchronize(c.pipe.out));
    CUDA_TRY(cudaGetLastError());
    cudaEventElapsedTime(&c.timing.h2d_ms, c.ev[0], c.ev[1]);
    cudaEventElapsedTime(&c.timing.kernels_ms, c.ev[1], c.ev[2]);
    cudaEventElapsedTime(&c.timing.d2h_ms, c.ev[2], c.ev[3]);
    cudaEventElapsedTime(&c.timing.total_ms, c.ev[0], c.ev[3]);
    c.timing.kernel_launches = ex.launches;
    c.timing.parser_fallback = c.fast_fallbacks != 0;
    if (c.prof.on) {
        // aggregate per kernel name, in first-launch order
        std::vector<std::string> names; std::vector<double> ms; std::vector<int> cnt;
        for (auto &r : c.prof.recs) {
            float t = 0; cudaEventElapsedTime(&t, r.a, r.b);
            size_t k = 0; while (k < names.size() && names[k] != r.name) k++;
            if (k == names.size()) { names.push_back(r.name); ms.push_back(0); cnt.push_back(0); }
            ms[k] += t; cnt[k]++;
            c.prof.pool.push_back(r.a); c.prof.pool.push_back(r.b);
        }
        c.prof.recs.clear();
        c.prof_report.clear();
        char line[256];
        for (size_t k = 0; k < names.size(); k++) { snprintf(line, sizeof line, "%s\t%d\t%.6f\n", names[k].c_str(), cnt[k], ms[k]); c.prof_report += line; }
    }
}

extern "C" {

int nafgpu_decode(nafgpu_ctx *c, const uint8_t *naf, size_t n, const nafgpu_dec_opts *opts, const uint8_t **text, size_t *text_size)
{
    if (!naf || !opts || !text || !text_size) return NAFGPU_E_ARG;
    return guarded(c, [&] {
        *text = nullptr; *text_size = 0;
        CudaExec ex{c->stream, &c->arena, &c->prof}; ex.staging = &c->pinned_stage; ex.mail = &c->mail;
        CUDA_TRY(cudaEventRecord(c->ev[0], c->stream));
        nafc::Header h; std::string err;
        if (!nafc::read_header(naf, n, h, false, err)) fail(NAFGPU_E_FORMAT, err);      // fail before any transfer
        u8 *d_naf;
        if (use_pipe(n)) {
            // the file goes up in chunks on its own stream; decode_on_device waits for exactly the bytes each part needs, and
            // sends finished pieces of the text down on a third stream while the next ones are produced
            d_naf = ex.alloc<u8>(n + 64);
            CUDA_TRY(cudaMemsetAsync(d_naf + n, 0, 64, c->stream));
            ex.pipe = &c->pipe;
            c->pipe.defer_upload(d_naf, naf, n, pipe_chunk(), c->stream);      // decode_on_device starts it, in the order it wants the bytes
        } else d_naf = to_device(*c, ex, naf, n);
        CUDA_TRY(cudaEventRecord(c->ev[1], c->stream));
        DecodeOut r = decode_on_device(*c, ex, d_naf, naf, n, *opts);
        CUDA_TRY(cudaEventRecord(c->ev[2], c->stream));
        if (c->pipe.emitting && r.d_text) {
            // [0, out_done) is already on its way down; whatever was produced outside the piece loop follows
            u8 *hout = c->pipe.h_out;
            const u64 done = c->pipe.out_done < r.size ? c->pipe.out_done : r.size;
            if (r.size > done) CUDA_TRY(cudaMemcpyAsync(hout + done, r.d_text + done, r.size - done, cudaMemcpyDeviceToHost, c->stream));
            *text = hout; *text_size = r.size;
        } else { *text = to_pinned(*c, r.d_text, r.size); *text_size = r.size; }
        finish_timing(*c, ex);
    });
}

int nafgpu_decode_to(nafgpu_ctx *c, const uint8_t *naf, size_t n, const nafgpu_dec_opts *opts, nafgpu_write_fn write, void *user, size_t *text_size)
{
    if (!naf || !opts || !write) return NAFGPU_E_ARG;
    return guarded(c, [&] {
        if (text_size) *text_size = 0;
        CudaExec ex{c->stream, &c->arena, &c->prof}; ex.staging = &c->pinned_stage; ex.mail = &c->mail;
        CUDA_TRY(cudaEventRecord(c->ev[0], c->stream));
        nafc::Header h; std::string err;
        if (!nafc::read_header(naf, n, h, false, err)) fail(NAFGPU_E_FORMAT, err);
        // the file goes up in chunks (pageable memory is fine: it is the small side), the text comes down through two rotating
        // page-locked buffers and is handed to `write` piece by piece while the next piece is on its way
        u8 *d_naf = ex.alloc<u8>(n + 64);
        CUDA_TRY(cudaMemsetAsync(d_naf + n, 0, 64, c->stream));
        ex.pipe = &c->pipe;
        c->pipe.rot_create();
        c->pipe.sink = write; c->pipe.sink_user = user;
        c->pipe.defer_upload(d_naf, naf, n, pipe_chunk(), c->stream);      // decode_on_device starts it, in the order it wants the bytes
        CUDA_TRY(cudaEventRecord(c->ev[1], c->stream));
        DecodeOut r = decode_on_device(*c, ex, d_naf, naf, n, *opts);
        CUDA_TRY(cudaEventRecord(c->ev[2], c->stream));
        if (r.d_text && r.size) {
            const u64 done = c->pipe.emitting ? (c->pipe.out_done < r.size ? c->pipe.out_done : r.size) : 0;
            if (!c->pipe.emitting) c->pipe.begin_output(nullptr);
            if (r.size > done) c->pipe.emit(c->stream, r.d_text + done, done, r.size - done);
        }
        c->pipe.rot_flush_all();
        if (text_size) *text_size = r.size;
        const bool failed = c->pipe.sink_failed;
        finish_timing(*c, ex);
        if (failed) fail(NAFGPU_E_ARG, "the output callback reported an error\n");
    });
}

/* ---- text arriving in pieces ---- */
static const size_t INGEST_ROT = 32u << 20;

int nafgpu_encode_begin(nafgpu_ctx *c, const nafgpu_enc_opts *opts, size_t size_hint)
{
    if (!c || !opts) return NAFGPU_E_ARG;
    try {
        c->err.clear();
        CUDA_TRY(cudaSetDevice(c->device));
        Ctx::Ingest &g = c->ingest;
        CUDA_TRY(cudaStreamSynchronize(c->pipe.in));
        g.active = true; g.opts = *opts; g.n = 0; g.cur = 0; g.busy[0] = g.busy[1] = false;
        g.has_title = opts->title != nullptr; g.title = opts->title ? opts->title : ""; g.opts.title = nullptr;
        for (int b = 0; b < 2; b++) if (!g.rot[b]) {
            CUDA_TRY(cudaHostAlloc(&g.rot[b], INGEST_ROT, cudaHostAllocDefault));
            CUDA_TRY(cudaEventCreateWithFlags(&g.ev[b], cudaEventDisableTiming));
        }
        const size_t want = (size_hint ? size_hint : (size_t)(256u << 20)) + 64;
        if (g.cap < want) { if (g.d_text) cudaFree(g.d_text); g.d_text = nullptr; g.cap = 0; CUDA_TRY(cudaMalloc(&g.d_text, want)); g.cap = want; }
        return NAFGPU_OK;
    } catch (const CudaError &e) {
        c->err = std::string("CUDA error: ") + cudaGetErrorString(e.e) + "\n"; cudaGetLastError(); c->ingest.active = false; return NAFGPU_E_CUDA;
    }
}

int nafgpu_encode_buffer(nafgpu_ctx *c, void **buf, size_t *cap)
{
    if (!c || !buf || !cap || !c->ingest.active) return NAFGPU_E_ARG;
    Ctx::Ingest &g = c->ingest;
    if (g.busy[g.cur]) { if (cudaEventSynchronize(g.ev[g.cur]) != cudaSuccess) { c->err = "CUDA error while uploading the text\n"; return NAFGPU_E_CUDA; } g.busy[g.cur] = false; }
    *buf = g.rot[g.cur]; *cap = INGEST_ROT;
    return NAFGPU_OK;
}

int nafgpu_encode_feed(nafgpu_ctx *c, size_t n)
{
    if (!c || !c->ingest.active || n > INGEST_ROT) return NAFGPU_E_ARG;
    if (!n) return NAFGPU_OK;
    Ctx::Ingest &g = c->ingest;
    try {
        CUDA_TRY(cudaSetDevice(c->device));
        if (g.n + n + 64 > g.cap) {                               // size unknown in advance (a pipe): grow geometrically, copy on the device
            const size_t ncap = (g.cap > (g.n + n + 64) / 2 ? g.cap * 2 : g.n + n + 64) + (64u << 20);
            u8 *nd = nullptr;
            CUDA_TRY(cudaMalloc(&nd, ncap));
            CUDA_TRY(cudaMemcpyAsync(nd, g.d_text, g.n, cudaMemcpyDeviceToDevice, c->pipe.in));
            CUDA_TRY(cudaStreamSynchronize(c->pipe.in));
            cudaFree(g.d_text); g.d_text = nd; g.cap = ncap;
        }
        CUDA_TRY(cudaMemcpyAsync(g.d_text + g.n, g.rot[g.cur], n, cudaMemcpyHostToDevice, c->pipe.in));
        CUDA_TRY(cudaEventRecord(g.ev[g.cur], c->pipe.in));
        g.busy[g.cur] = true; g.cur ^= 1; g.n += n;
        return NAFGPU_OK;
    } catch (const CudaError &e) {
        c->err = std::string("CUDA error: ") + cudaGetErrorString(e.e) + "\n"; cudaGetLastError(); return NAFGPU_E_CUDA;
    }
}

static int encode_end_impl(nafgpu_ctx *c, const uint8_t **naf, size_t *naf_size, nafgpu_enc_info *info, nafgpu_write_fn write, void *user);

int nafgpu_encode_end(nafgpu_ctx *c, const uint8_t **naf, size_t *naf_size, nafgpu_enc_info *info)
{
    if (!c || !naf || !naf_size || !c->ingest.active) return NAFGPU_E_ARG;
    return encode_end_impl(c, naf, naf_size, info, nullptr, nullptr);
}
int nafgpu_encode_end_to(nafgpu_ctx *c, nafgpu_write_fn write, void *user, size_t *naf_size, nafgpu_enc_info *info)
{
    if (!c || !write || !c->ingest.active) return NAFGPU_E_ARG;
    const uint8_t *unused = nullptr; size_t sz = 0;
    const int rc = encode_end_impl(c, &unused, &sz, info, write, user);
    if (naf_size) *naf_size = sz;
    return rc;
}

static int encode_end_impl(nafgpu_ctx *c, const uint8_t **naf, size_t *naf_size, nafgpu_enc_info *info, nafgpu_write_fn write, void *user)
{
    c->ingest.active = false;
    return guarded(c, [&] {
        *naf = nullptr; *naf_size = 0;
        Ctx::Ingest &g = c->ingest;
        CudaExec ex{c->stream, &c->arena, &c->prof}; ex.staging = &c->pinned_stage; ex.mail = &c->mail;
        CUDA_TRY(cudaEventRecord(c->ev[0], c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->pipe.in));              // every piece has arrived
        g.busy[0] = g.busy[1] = false;
        CUDA_TRY(cudaMemsetAsync(g.d_text + g.n, 0, 64, c->stream));
        CUDA_TRY(cudaEventRecord(c->ev[1], c->stream));
        nafgpu_enc_opts o = g.opts; o.title = g.has_title ? g.title.c_str() : nullptr;
        EncodeOut r = encode_on_device(*c, ex, g.d_text, g.n, o, info);
        CUDA_TRY(cudaEventRecord(c->ev[2], c->stream));
        if (write) {                                              // the file comes down through the two rotating buffers, piece by piece
            c->pipe.rot_create();
            c->pipe.sink = write; c->pipe.sink_user = user;
            c->pipe.begin_output(nullptr);
            c->pipe.emit(c->stream, r.d_naf, 0, r.size);
            c->pipe.rot_flush_all();
            *naf_size = r.size;
            const bool failed = c->pipe.sink_failed;
            finish_timing(*c, ex);
            if (failed) fail(NAFGPU_E_ARG, "the output callback reported an error\n");
            return;
        }
        *naf = to_pinned(*c, r.d_naf, r.size); *naf_size = r.size;
        finish_timing(*c, ex);
    });
}

int nafgpu_decode_device(nafgpu_ctx *c, const uint8_t *d_naf, size_t n, const uint8_t *host_copy, const nafgpu_dec_opts *opts,
                         const uint8_t **d_text, size_t *text_size)
{
    if (!d_naf || !opts || !d_text || !text_size) return NAFGPU_E_ARG;
    return guarded(c, [&] {
        *d_text = nullptr; *text_size = 0;
        CudaExec ex{c->stream, &c->arena, &c->prof}; ex.staging = &c->pinned_stage; ex.mail = &c->mail;
        CUDA_TRY(cudaEventRecord(c->ev[0], c->stream));
        const u8 *h = host_copy;
        if (!h) {                                     // no host mirror: fetch the compressed bytes once for the header walk
            u8 *tmp = c->pinned_aux.ensure(n + 1);
            CUDA_TRY(cudaMemcpyAsync(tmp, d_naf, n, cudaMemcpyDeviceToHost, c->stream));
            CUDA_TRY(cudaStreamSynchronize(c->stream));
            h = tmp;
        }
        CUDA_TRY(cudaEventRecord(c->ev[1], c->stream));
        DecodeOut r = decode_on_device(*c, ex, d_naf, h, n, *opts);
        CUDA_TRY(cudaEventRecord(c->ev[2], c->stream));
        *d_text = r.d_text; *text_size = r.size;
        finish_timing(*c, ex);
    });
}

int nafgpu_zstd_decompress(nafgpu_ctx *c, const uint8_t *src, size_t n, size_t expected_size, int one_frame,
                           const uint8_t **out, size_t *out_size)
{
    if (!src || !out || !out_size) return NAFGPU_E_ARG;
    return guarded(c, [&] {
        *out = nullptr; *out_size = 0;
        CudaExec ex{c->stream, &c->arena, &c->prof}; ex.staging = &c->pinned_stage; ex.mail = &c->mail;
        CUDA_TRY(cudaEventRecord(c->ev[0], c->stream));
        u8 *d_in = to_device(*c, ex, src, n);
        CUDA_TRY(cudaEventRecord(c->ev[1], c->stream));
        // unknown size: bound it by walking the headers is not possible for compressed blocks, so take
        // the format's worst case of 128 KB per block (cheap: the arena is virtual until touched).
        nafz::ZDecPlan plan;
        u64 cap = expected_size;
        if (!cap) {
            std::vector<nafz::ZBlockHead> blocks; std::string werr; u64 used = 0;
            nafz::ZStreamDesc probe{0, n, 0, ~0ull, one_frame, 0};
            if (nafz::zstd_walk_stream(src, probe, 0, blocks, &used, werr)) fail(NAFGPU_E_FORMAT, werr + "\n");
            for (auto &b : blocks) cap += b.type == 2 ? 128 * 1024 : b.rsize;
        }
        u8 *d_out = ex.alloc<u8>(cap + 256);
        plan.streams.push_back(nafz::ZStreamDesc{0, n, 0, cap, one_frame, 0});
        std::string zerr;
        int rc = nafz::zstd_decode_batch(ex, d_in, src, d_out, plan, c->d_predef, zerr);
        if (rc) fail(rc == -2 ? NAFGPU_E_UNSUPPORTED : NAFGPU_E_FORMAT, zerr + "\n");
        if (expected_size && plan.results[0].out_size != expected_size) fail(NAFGPU_E_FORMAT, "zstd stream size mismatch\n");
        CUDA_TRY(cudaEventRecord(c->ev[2], c->stream));
        *out = to_pinned(*c, d_out, plan.results[0].out_size); *out_size = plan.results[0].out_size;
        finish_timing(*c, ex);
    });
}

int nafgpu_encode(nafgpu_ctx *c, const uint8_t *text, size_t n, const nafgpu_enc_opts *opts, const uint8_t **naf, size_t *naf_size,
                  nafgpu_enc_info *info)
{
    if ((!text && n) || !opts || !naf || !naf_size) return NAFGPU_E_ARG;
    return guarded(c, [&] {
        *naf = nullptr; *naf_size = 0;
        CudaExec ex{c->stream, &c->arena, &c->prof}; ex.staging = &c->pinned_stage; ex.mail = &c->mail;
        CUDA_TRY(cudaEventRecord(c->ev[0], c->stream));
        u8 *d_text;
        if (use_pipe(n)) {
            // the text goes up in chunks on its own stream; the transform kernel is launched chunk by chunk right behind it
            d_text = ex.alloc<u8>(n + 64);
            CUDA_TRY(cudaMemsetAsync(d_text + n, 0, 64, c->stream));
            ex.pipe = &c->pipe;
            c->pipe.upload(d_text, text, n, pipe_chunk(), c->stream);
        } else d_text = to_device(*c, ex, text, n);
        CUDA_TRY(cudaEventRecord(c->ev[1], c->stream));
        EncodeOut r = encode_on_device(*c, ex, d_text, n, *opts, info);
        CUDA_TRY(cudaEventRecord(c->ev[2], c->stream));
        *naf = to_pinned(*c, r.d_naf, r.size); *naf_size = r.size;
        finish_timing(*c, ex);
    });
}

int nafgpu_encode_device(nafgpu_ctx *c, const uint8_t *d_text, size_t n, const nafgpu_enc_opts *opts, const uint8_t **d_naf,
                         size_t *naf_size, nafgpu_enc_info *info)
{
    if ((!d_text && n) || !opts || !d_naf || !naf_size) return NAFGPU_E_ARG;
    return guarded(c, [&] {
        *d_naf = nullptr; *naf_size = 0;
        CudaExec ex{c->stream, &c->arena, &c->prof}; ex.staging = &c->pinned_stage; ex.mail = &c->mail;
        CUDA_TRY(cudaEventRecord(c->ev[0], c->stream));
        CUDA_TRY(cudaEventRecord(c->ev[1], c->stream));
        EncodeOut r = encode_on_device(*c, ex, d_text, n, *opts, info);
        CUDA_TRY(cudaEventRecord(c->ev[2], c->stream));
        *d_naf = r.d_naf; *naf_size = r.size;
        finish_timing(*c, ex);
    });
}

int nafgpu_zstd_compress(nafgpu_ctx *c, const uint8_t *src, size_t n, int window_log, const uint8_t **out, size_t *out_size)
{
    return nafgpu_zstd_compress_level(c, src, n, window_log, 0, out, out_size);
}

int nafgpu_zstd_compress_level(nafgpu_ctx *c, const uint8_t *src, size_t n, int window_log, int level, const uint8_t **out, size_t *out_size)
{
    if ((!src && n) || !out || !out_size) return NAFGPU_E_ARG;
    return guarded(c, [&] {
        *out = nullptr; *out_size = 0;
        CudaExec ex{c->stream, &c->arena, &c->prof}; ex.staging = &c->pinned_stage; ex.mail = &c->mail;
        CUDA_TRY(cudaEventRecord(c->ev[0], c->stream));
        u8 *d_in = to_device(*c, ex, src, n);
        CUDA_TRY(cudaEventRecord(c->ev[1], c->stream));
        EncodeOut r = zstd_compress_on_device(*c, ex, d_in, n, window_log, level);
        CUDA_TRY(cudaEventRecord(c->ev[2], c->stream));
        *out = to_pinned(*c, r.d_naf, r.size); *out_size = r.size;
        finish_timing(*c, ex);
    });
}

int nafgpu_split(nafgpu_ctx *c, const uint8_t *text, size_t n, const nafgpu_enc_opts *opts, const uint8_t *streams[6], size_t sizes[6],
                 nafgpu_enc_info *info)
{
    if ((!text && n) || !opts || !streams || !sizes) return NAFGPU_E_ARG;
    return guarded(c, [&] {
        CudaExec ex{c->stream, &c->arena, &c->prof}; ex.staging = &c->pinned_stage; ex.mail = &c->mail;
        CUDA_TRY(cudaEventRecord(c->ev[0], c->stream));
        u8 *d_text = to_device(*c, ex, text, n);
        CUDA_TRY(cudaEventRecord(c->ev[1], c->stream));
        SplitOut r = split_on_device(*c, ex, d_text, n, *opts, info);
        CUDA_TRY(cudaEventRecord(c->ev[2], c->stream));
        u64 total = 0;
        for (int k = 0; k < 6; k++) total += (r.size[k] + 63) & ~63ull;
        u8 *h = c->pinned_out.ensure(total + 64);
        u64 off = 0;
        for (int k = 0; k < 6; k++) {
            streams[k] = h + off; sizes[k] = r.size[k];
            if (r.size[k]) CUDA_TRY(cudaMemcpyAsync(h + off, r.d[k], r.size[k], cudaMemcpyDeviceToHost, c->stream));
            off += (r.size[k] + 63) & ~63ull;
        }
        finish_timing(*c, ex);
    });
}


/* ---- one file from several shards (multi-GPU encode; naf_b200/sharded.py drives the exchange) ---- */

int nafgpu_record_cuts(nafgpu_ctx *c, const uint8_t *text, size_t n, int text_on_device, int pieces, uint64_t *cuts)
{
    if ((!text && n) || !cuts || pieces < 1) return NAFGPU_E_ARG;
    return guarded(c, [&] {
        CudaExec ex{c->stream, &c->arena, &c->prof}; ex.staging = &c->pinned_stage; ex.mail = &c->mail;
        CUDA_TRY(cudaEventRecord(c->ev[0], c->stream));
        const u8 *d_text = text_on_device ? text : to_device(*c, ex, text, n);
        CUDA_TRY(cudaEventRecord(c->ev[1], c->stream));
        record_cuts_on_device(*c, ex, d_text, n, pieces, cuts);
        CUDA_TRY(cudaEventRecord(c->ev[2], c->stream));
        finish_timing(*c, ex);
    });
}

int nafgpu_shard_begin(nafgpu_ctx *c, const uint8_t *text, size_t n, int text_on_device, const nafgpu_enc_opts *opts,
                       nafgpu_shard_counts *counts, nafgpu_enc_info *info)
{
    if ((!text && n) || !opts || !counts) return NAFGPU_E_ARG;
    return guarded(c, [&] {
        CudaExec ex{c->stream, &c->arena, &c->prof}; ex.staging = &c->pinned_stage; ex.mail = &c->mail;
        CUDA_TRY(cudaEventRecord(c->ev[0], c->stream));
        const u8 *d_text = text_on_device ? text : to_device(*c, ex, text, n);
        CUDA_TRY(cudaEventRecord(c->ev[1], c->stream));
        shard_begin_on_device(*c, ex, d_text, n, *opts, counts, info);
        CUDA_TRY(cudaEventRecord(c->ev[2], c->stream));
        finish_timing(*c, ex);
    });
}

int nafgpu_shard_finish(nafgpu_ctx *c, const nafgpu_shard_link *link, uint64_t raw[6], uint64_t body[6])
{
    if (!c || !link || !raw || !body) return NAFGPU_E_ARG;
    if (!c->shard.active) { c->err = "nafgpu_shard_finish without nafgpu_shard_begin\n"; return NAFGPU_E_ARG; }
    c->keep_arena = true;
    return guarded(c, [&] {
        CudaExec ex{c->stream, &c->arena, &c->prof}; ex.staging = &c->pinned_stage; ex.mail = &c->mail;
        CUDA_TRY(cudaEventRecord(c->ev[0], c->stream));
        CUDA_TRY(cudaEventRecord(c->ev[1], c->stream));
        shard_finish_on_device(*c, ex, *link, raw, body);
        CUDA_TRY(cudaEventRecord(c->ev[2], c->stream));
        finish_timing(*c, ex);
    });
}

int nafgpu_shard_fetch(nafgpu_ctx *c, int stream, void *dst)
{
    if (!c || stream < 0 || stream > 5 || !dst) return NAFGPU_E_ARG;
    if (!c->shard.active || !c->shard.finished) { c->err = "nafgpu_shard_fetch without nafgpu_shard_finish\n"; return NAFGPU_E_ARG; }
    c->keep_arena = true;
    return guarded(c, [&] {
        if (c->shard.body_size[stream]) CUDA_TRY(cudaMemcpyAsync(dst, c->shard.body[stream], c->shard.body_size[stream], cudaMemcpyDefault, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
    });
}

}  // extern "C"
