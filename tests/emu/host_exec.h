// host_exec.h — TEST-ONLY executor: runs the decoder's HD kernel bodies in plain loops on the CPU so
// that the thread-serial zstd logic can be checked against the oracle without a GPU.  Never compiled
// into libnafgpu.so (which instantiates CudaExec only); lives under tests/.
#pragma once
#include <cstdlib>
#include <cstring>
#include <vector>
struct HostExec {
    std::vector<void *> owned;
    unsigned launches = 0;
    ~HostExec() { for (void *p : owned) free(p); }
    template <class T> T *alloc(size_t count) { void *p = calloc(count ? count : 1, sizeof(T)); owned.push_back(p); return (T *)p; }
    void upload(void *dst, const void *src, size_t n) { memcpy(dst, src, n); }
    void upload_staged(void *dst, const void *src, size_t n) { memcpy(dst, src, n); }
    void download(void *dst, const void *src, size_t n) { memcpy(dst, src, n); }
    void zero(void *p, size_t n) { memset(p, 0, n); }
    template <class F> void for_each(size_t n, F f, const char * = nullptr, int = 0) { launches++; for (size_t i = 0; i < n; i++) f(i); }
    // group kernels are written as strided loops: emulate once with 1 thread and once more with an
    // awkward thread count to catch stride bugs (results must be idempotent).
    template <class F> void for_each_group(size_t ngroups, int, F f, const char * = nullptr)
    {
        launches++;
        for (size_t g = 0; g < ngroups; g++) for (unsigned t = 0; t < 3; t++) f(g, t, 3u);
    }
};
