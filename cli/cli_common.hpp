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
