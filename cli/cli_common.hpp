// cli_common.hpp — shared host plumbing of the drop-in `ennaf` / `unnaf` command-line tools.
// Everything here is plain file and argument handling (the reference's files.c / utils.c); the data
// path is one call into libnafgpu.so.
#pragma once
#include <errno.h>
#include <fcntl.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <string>
#include <thread>
#include <vector>

#include "nafgpu.h"

static const char *g_tool = "naf";
static char *g_out_path = nullptr;
static bool g_created_output = false, g_success = false;

__attribute__((format(printf, 1, 2))) static void msg(const char *fmt, ...)
{
    va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap);
}
__attribute__((format(printf, 1, 2))) static void warn(const char *fmt, ...)
{
    fprintf(stderr, "%s warning: ", g_tool);
    va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap);
}
__attribute__((format(printf, 1, 2))) static void err(const char *fmt, ...)
{
    fprintf(stderr, "%s error: ", g_tool);
    va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap);
}
__attribute__((noreturn, format(printf, 1, 2))) static void die(const char *fmt, ...)
{
    fprintf(stderr, "%s error: ", g_tool);
    va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap);
    exit(1);
}

// atexit: remove an incomplete output file (ennaf.c:154-157, unnaf.c:189-192)
static void cleanup_output()
{
    if (!g_success && g_created_output && g_out_path) {
        if (remove(g_out_path) != 0) err("can't remove incomplete output file \"%s\"\n", g_out_path);
    }
}

struct Input {
    const uint8_t *data = nullptr; size_t size = 0;
    bool mapped = false; std::vector<uint8_t> owned;
    struct stat st; bool have_stat = false;
};

// whole input in memory: mmap for regular files, read() loop for pipes
static void load_input(const char *path, Input &in)
{
    int fd = path ? open(path, O_RDONLY) : 0;
    if (fd < 0) die("can't open input file\n");
    if (fstat(fd, &in.st) == 0) in.have_stat = path != nullptr;
    if (path && S_ISREG(in.st.st_mode) && in.st.st_size > 0) {
        void *p = mmap(nullptr, (size_t)in.st.st_size, PROT_READ, MAP_PRIVATE | MAP_POPULATE, fd, 0);
        if (p != MAP_FAILED) { in.data = (const uint8_t *)p; in.size = (size_t)in.st.st_size; in.mapped = true; close(fd); return; }
    }
    std::vector<uint8_t> &b = in.owned;
    size_t cap = 1 << 20; b.resize(cap); size_t n = 0;
    for (;;) {
        if (n == cap) { cap *= 2; b.resize(cap); }
        ssize_t k = read(fd, b.data() + n, cap - n);
        if (k < 0) { if (errno == EINTR) continue; die("can't read input\n"); }
        if (k == 0) break;
        n += (size_t)k;
    }
    b.resize(n);
    in.data = b.data(); in.size = n;
    if (path) close(fd);
}

static FILE *open_output(const char *path, bool force_stdout)
{
    if (path && !force_stdout) {
        FILE *f = fopen(path, "wb");
        if (!f) die("can't create output file\n");
        g_created_output = true;
        return f;
    }
    return stdout;
}

// files.c:114-156 close_output_file_and_set_stat
static void close_output(FILE *f, const Input &in, bool transfer_stat)
{
    if (fflush(f) != 0) die("can't write to file - disk full?\n");
    if (transfer_stat && in.have_stat && f != stdout) {
        if (fchmod(fileno(f), in.st.st_mode & (S_IRWXU | S_IRWXG | S_IRWXO)) != 0) err("can't transfer permissions from input to output file\n");
        if (fchown(fileno(f), in.st.st_uid, in.st.st_gid) != 0) err("can't transfer ownership from input to output file\n");
        struct timespec ts[2] = { in.st.st_atim, in.st.st_mtim };
        if (futimens(fileno(f), ts) != 0) err("can't transfer timestamp from input to output file\n");
    }
    if (f != stdout) { if (fclose(f) != 0) die("can't close file - disk full?\n"); }
}

static void write_all(FILE *f, const uint8_t *p, size_t n)
{
    while (n) {
        size_t k = fwrite(p, 1, n > (1u << 30) ? (1u << 30) : n, f);
        if (k == 0) die("can't write to file - disk full?\n");
        p += k; n -= k;
    }
}

// ---- regular files: a piece of a few tens of MB is read / written by a few threads at once (pread / pwrite on disjoint
// ranges).  One thread moves 3-4 GB/s through the page cache, which made file I/O -- not the GPU, not PCIe -- the longest part of
// a run of the tools on a multi-GB file.  Pipes and terminals keep the plain loops.
static const size_t IO_THREADS = 4, IO_PAR_MIN = 8u << 20;
static bool fd_is_regular(int fd) { struct stat st; return fstat(fd, &st) == 0 && S_ISREG(st.st_mode); }

// up to n bytes at file offset off -> buf; returns the bytes read, contiguous from buf (short only at the end of the file)
static size_t par_pread(int fd, uint8_t *buf, size_t n, off_t off)
{
    auto range = [fd](uint8_t *p, size_t len, off_t at) -> size_t {
        size_t got = 0;
        while (got < len) {
            ssize_t k = pread(fd, p + got, len - got, at + (off_t)got);
            if (k < 0) { if (errno == EINTR) continue; die("can't read input\n"); }
            if (k == 0) break;
            got += (size_t)k;
        }
        return got;
    };
    if (n < IO_PAR_MIN) return range(buf, n, off);
    const size_t part = (n / IO_THREADS + 4095) & ~(size_t)4095;
    size_t got[IO_THREADS] = {0}, want[IO_THREADS] = {0};
    std::thread th[IO_THREADS];
    for (size_t t = 0; t < IO_THREADS; t++) {
        const size_t lo = t * part < n ? t * part : n, hi = (t + 1) * part < n ? (t + 1) * part : n;
        want[t] = t + 1 == IO_THREADS ? n - lo : hi - lo;
        if (want[t]) th[t] = std::thread([&, t, lo] { got[t] = range(buf + lo, want[t], off + (off_t)lo); });
    }
    size_t total = 0; bool full = true;
    for (size_t t = 0; t < IO_THREADS; t++) { if (th[t].joinable()) th[t].join(); if (full) total += got[t]; if (got[t] < want[t]) full = false; }
    return total;
}
static void par_pwrite(int fd, const uint8_t *buf, size_t n, off_t off)
{
    auto range = [fd](const uint8_t *p, size_t len, off_t at) {
        size_t done = 0;
        while (done < len) {
            ssize_t k = pwrite(fd, p + done, len - done, at + (off_t)done);
            if (k < 0) { if (errno == EINTR) continue; die("can't write to file - disk full?\n"); }
            if (k == 0) die("can't write to file - disk full?\n");
            done += (size_t)k;
        }
    };
    if (n < IO_PAR_MIN) { range(buf, n, off); return; }
    const size_t part = (n / IO_THREADS + 4095) & ~(size_t)4095;
    std::thread th[IO_THREADS];
    for (size_t t = 0; t < IO_THREADS; t++) {
        const size_t lo = t * part < n ? t * part : n, hi = t + 1 == IO_THREADS ? n : ((t + 1) * part < n ? (t + 1) * part : n);
        if (hi > lo) th[t] = std::thread([=] { range(buf + lo, hi - lo, off + (off_t)lo); });
    }
    for (size_t t = 0; t < IO_THREADS; t++) if (th[t].joinable()) th[t].join();
}
// the pieces a streamed call delivers, in order, into `f`: through pwrite when f is a regular file, else through stdio
struct PieceWriter {
    FILE *f = nullptr; int fd = -1; off_t off = 0; bool direct = false;
    void attach(FILE *file)
    {
        f = file; fd = fileno(file);
        if (fd >= 0 && fd_is_regular(fd) && fflush(file) == 0) { const off_t at = lseek(fd, 0, SEEK_CUR); if (at >= 0) { off = at; direct = true; } }
    }
    void put(const uint8_t *p, size_t n)
    {
        if (!direct) { write_all(f, p, n); return; }
        par_pwrite(fd, p, n, off); off += (off_t)n;
    }
    void finish() { if (direct) lseek(fd, off, SEEK_SET); }          // whatever stdio writes next (nothing, today) goes behind it
};

// The work is done and the output is closed: leave without tearing down the CUDA context.  Freeing a multi-GB arena, the
// page-locked buffers and the context itself costs a few tenths of a second that the next tool of a pipeline waits for; the
// driver reclaims all of it with the process.
__attribute__((noreturn)) static void exit_done()
{
    g_success = true;
    fflush(stdout); fflush(stderr);
    _exit(0);
}

static nafgpu_ctx *make_ctx()
{
    nafgpu_ctx *ctx = nullptr;
    if (nafgpu_create(-1, &ctx) != 0) die("%s", nafgpu_last_error(nullptr));
    return ctx;
}

static bool parse_ull_strict(const char *s, unsigned long long &v)   // ennaf.c:224-238 set_line_length's checks
{
    char *end; long long a = strtoll(s, &end, 10);
    if (*end != '\0' || a < 0) return false;
    char test[21]; int nc = snprintf(test, 21, "%lld", a);
    if (nc < 1 || nc > 20 || strcmp(test, s) != 0) return false;
    v = (unsigned long long)a; return true;
}
