// cli_io_check — TEST-ONLY: the threaded file I/O of the command-line tools (cli/cli_common.hpp par_pread / par_pwrite /
// PieceWriter) on a scratch file: pieces of awkward sizes, a short read at the end of the file, a stdio prefix in front
// of directly written pieces, and the pipe path.   cli_io_check SCRATCH_DIR   -> exit 0 if every byte is where it belongs
#include "../../cli/cli_common.hpp"

static std::vector<uint8_t> pattern(size_t n, uint32_t seed)
{
    std::vector<uint8_t> v(n);
    uint32_t x = seed * 2654435761u + 12345u;
    for (size_t i = 0; i < n; i++) { x = x * 1664525u + 1013904223u; v[i] = (uint8_t)(x >> 24); }
    return v;
}

int main(int argc, char **argv)
{
    if (argc < 2) return 2;
    const std::string path = std::string(argv[1]) + "/cli_io_check.bin";
    const size_t sizes[] = { 0, 1, 4095, 4096, (8u << 20) - 1, 8u << 20, (8u << 20) + 1, (32u << 20) + 12345, 3u << 20 };
    std::vector<uint8_t> all;
    const char prefix[] = "header written through stdio\n";
    {
        FILE *f = fopen(path.c_str(), "wb");
        if (!f) return 3;
        fputs(prefix, f);                                   // buffered: attach() must flush it and continue behind it
        all.insert(all.end(), prefix, prefix + sizeof prefix - 1);
        PieceWriter w; w.attach(f);
        if (!w.direct) { fprintf(stderr, "a regular file was not taken as one\n"); return 4; }
        uint32_t seed = 1;
        for (size_t n : sizes) { const std::vector<uint8_t> p = pattern(n, seed++); w.put(p.data(), p.size()); all.insert(all.end(), p.begin(), p.end()); }
        w.finish();
        fputs("tail", f); all.insert(all.end(), {'t', 'a', 'i', 'l'});       // stdio continues behind the pieces
        if (fclose(f) != 0) return 5;
    }
    {
        const int fd = open(path.c_str(), O_RDONLY);
        if (fd < 0 || !fd_is_regular(fd)) return 6;
        struct stat st; fstat(fd, &st);
        if ((size_t)st.st_size != all.size()) { fprintf(stderr, "size %zu, expected %zu\n", (size_t)st.st_size, all.size()); return 7; }
        for (size_t cap : { (size_t)(32u << 20), (size_t)((8u << 20) + 7), (size_t)(1u << 20), (size_t)(64u << 20) }) {
            std::vector<uint8_t> got, buf(cap);
            off_t off = 0;
            for (;;) { const size_t k = par_pread(fd, buf.data(), cap, off); got.insert(got.end(), buf.begin(), buf.begin() + k); off += (off_t)k; if (k < cap) break; }
            if (got != all) { fprintf(stderr, "read back with pieces of %zu differs\n", cap); return 8; }
        }
        close(fd);
    }
    {   // not a regular file: PieceWriter goes through stdio
        int fds[2]; if (pipe(fds) != 0) return 9;
        FILE *f = fdopen(fds[1], "wb");
        PieceWriter w; w.attach(f);
        if (w.direct) return 10;
        const std::vector<uint8_t> p = pattern(5000, 99);
        w.put(p.data(), p.size()); w.finish(); fclose(f);
        std::vector<uint8_t> got(6000);
        const ssize_t k = read(fds[0], got.data(), got.size());
        if (k != 5000 || memcmp(got.data(), p.data(), 5000) != 0) return 11;
        close(fds[0]);
    }
    remove(path.c_str());
    puts("ok");
    return 0;
}
