// ennaf — drop-in command line of the reference's compressor (ennaf/src/ennaf.c:360-600), host side only:
// argument parsing, file naming, stat transfer, the unexpected-character report.  The work is one call
// to nafgpu_encode() in libnafgpu.so (B200).  Same flags, same defaults, same messages and exit codes.
// Deliberate differences: no temporary files are ever needed, so --temp-dir / --name / --keep-temp-files
// are accepted and ignored and a missing TMPDIR is not an error (reference quirk);
// -# / --level selects between the two parses the GPU encoder has (1: entropy-only, >= 2: + LZ77 on the text-like streams);
// --version names this implementation.
#include "cli_common.hpp"

static bool verbose = false, force_stdout = false, no_mask = false, strict_mode = false, well_formed = false;
static char *in_path = nullptr, *title = nullptr;
static int level = 1, window_log = 0, seq_type = NAFGPU_DNA, fmt_cli = NAFGPU_FMT_AUTO;
static bool have_line_length = false; static unsigned long long line_length = 0;

static void set_in(char *p) { if (in_path) die("can compress only one file at a time\n"); if (!*p) die("empty input file name\n"); in_path = p; }
static void set_out(char *p) { if (g_out_path) die("double --out parameter\n"); if (!*p) die("empty --out parameter\n"); g_out_path = p; }
static void set_level(const char *s)
{
    char *end; long a = strtol(s, &end, 10);
    if (a < -131072 || a > 22 || *end) die("invalid value of --level, should be from %d to %d\n", -131072, 22);
    level = (int)a;
}
static void set_long(const char *s)
{
    unsigned long long a;
    if (!parse_ull_strict(s, a)) die("can't parse the value of --long argument\n");
    if (a < 10) { warn("--long value of is %llu is smaller than the lowest supported value %d, using %d instead\n", a, 10, 10); a = 10; }
    if (a > 31) { warn("--long value of is %llu is larger than the largest supported value %d, using %d instead\n", a, 31, 31); a = 31; }
    window_log = (int)a;
}
static int parse_fmt(const char *s)
{
    if (!strcasecmp(s, "fasta") || !strcasecmp(s, "fa") || !strcasecmp(s, "fna")) return NAFGPU_FMT_FASTA;
    if (!strcasecmp(s, "fastq") || !strcasecmp(s, "fq")) return NAFGPU_FMT_FASTQ;
    return NAFGPU_FMT_AUTO;
}
static void set_fmt(const char *s)
{
    if (fmt_cli != NAFGPU_FMT_AUTO) die("input format specified more than once\n");
    fmt_cli = parse_fmt(s);
    if (fmt_cli == NAFGPU_FMT_AUTO) die("unknown input format specified: \"%s\"\n", s);
}

static void show_help()
{
    msg("Usage: ennaf [OPTIONS] [infile]\n"
        "Options:\n"
        "  -o FILE            - Write compressed output to FILE\n"
        "  -c                 - Write to standard output\n"
        "  -#, --level #      - 1 (default): fastest parse; 2 and above: also LZ77-match names and lengths\n"
        "  --long N           - Accepted for compatibility (window of size 2^N for sequence stream)\n"
        "  --temp-dir DIR     - Accepted for compatibility (no temporary files are used)\n"
        "  --name NAME        - Accepted for compatibility\n"
        "  --title TITLE      - Store TITLE as dataset title\n"
        "  --fasta            - Input is in FASTA format\n"
        "  --fastq            - Input is in FASTQ format\n"
        "  --dna              - Input sequence is DNA (default)\n"
        "  --rna              - Input sequence is RNA\n"
        "  --protein          - Input sequence is protein\n"
        "  --text             - Input sequence is text\n"
        "  --strict           - Fail on unexpected input characters\n"
        "  --line-length N    - Override line length to N\n"
        "  --verbose          - Verbose mode\n"
        "  --keep-temp-files  - Accepted for compatibility\n"
        "  --no-mask          - Don't store mask\n"
        "  -h, --help         - Show help\n"
        "  -V, --version      - Show version\n");
}

static void report(const unsigned long long *n, const char *what)     // process.c:75-87
{
    unsigned long long total = 0;
    for (int i = 0; i < 257; i++) total += n[i];
    if (!total) return;
    msg("input has %llu unexpected %s characters:\n", total, what);
    for (int i = 0; i < 32; i++) if (n[i]) msg("    '\\x%02X': %llu\n", i, n[i]);
    for (int i = 32; i < 127; i++) if (n[i]) msg("    '%c': %llu\n", (unsigned char)i, n[i]);
    for (int i = 127; i < 256; i++) if (n[i]) msg("    '\\x%02X': %llu\n", i, n[i]);
    if (n[256]) msg("    EOF: %llu\n", n[256]);
}

int main(int argc, char **argv)
{
    g_tool = "ennaf";
    atexit(cleanup_output);
    bool print_version = false;
    for (int i = 1; i < argc; i++) {
        char *a = argv[i];
        if (a[0] == '-') {
            if (a[1] == '-') {
                if (i < argc - 1) {
                    if (!strcmp(a, "--temp-dir")) { i++; if (!*argv[i]) die("empty --temp-dir parameter\n"); continue; }
                    if (!strcmp(a, "--name")) { i++; if (!*argv[i]) die("empty --name parameter\n"); continue; }
                    if (!strcmp(a, "--title")) { i++; if (title) die("double --title parameter\n"); if (!*argv[i]) die("empty --title parameter\n"); title = argv[i]; continue; }
                    if (!strcmp(a, "--level")) { i++; set_level(argv[i]); continue; }
                    if (!strcmp(a, "--line-length")) {
                        i++;
                        char *end; long long v = strtoll(argv[i], &end, 10);
                        if (*end) die("can't parse the value of --line-length parameter\n");
                        if (v < 0) die("negative line length specified\n");
                        if (!parse_ull_strict(argv[i], line_length)) die("can't parse the value of --line-length parameter\n");
                        have_line_length = true; continue;
                    }
                    if (!strcmp(a, "--long")) { i++; set_long(argv[i]); continue; }
                    if (!strcmp(a, "--out")) { i++; set_out(argv[i]); continue; }
                    if (!strcmp(a, "--in")) { i++; set_in(argv[i]); continue; }
                    if (!strcmp(a, "--in-format")) { i++; set_fmt(argv[i]); continue; }
                }
                if (!strcmp(a, "--help")) { show_help(); exit(0); }
                if (!strcmp(a, "--version")) { print_version = true; continue; }
                if (!strcmp(a, "--verbose")) { verbose = true; continue; }
                if (!strcmp(a, "--binary-stderr")) { continue; }
                if (!strcmp(a, "--keep-temp-files")) { continue; }
                if (!strcmp(a, "--no-mask")) { no_mask = true; continue; }
                if (!strcmp(a, "--fasta")) { set_fmt("fasta"); continue; }
                if (!strcmp(a, "--fastq")) { set_fmt("fastq"); continue; }
                if (!strcmp(a, "--dna")) { seq_type = NAFGPU_DNA; continue; }
                if (!strcmp(a, "--rna")) { seq_type = NAFGPU_RNA; continue; }
                if (!strcmp(a, "--protein")) { seq_type = NAFGPU_PROTEIN; continue; }
                if (!strcmp(a, "--text")) { seq_type = NAFGPU_TEXT; continue; }
                if (!strcmp(a, "--well-formed")) { well_formed = true; continue; }
                if (!strcmp(a, "--strict")) { strict_mode = true; continue; }
            }
            if (i < argc - 1 && !strcmp(a, "-o")) { i++; set_out(argv[i]); continue; }
            if (!strcmp(a, "-c")) { force_stdout = true; continue; }
            if (a[1] >= '0' && a[1] <= '9') { set_level(a + 1); continue; }
            if (!strcmp(a, "-h")) { show_help(); exit(0); }
            if (!strcmp(a, "-V")) { print_version = true; continue; }
            die("unknown or incomplete argument \"%s\"\n", a);
        }
        set_in(a);
    }
    if (print_version) {
        msg("ennaf - NAF compressor for NVIDIA B200 (naf-b200 %s), writes NAF format of ennaf 1.3.0\n", nafgpu_version());
        exit(0);
    }
    if (force_stdout && g_out_path) die("'-c' and '-o' can't be used together\n");
    if (well_formed && strict_mode) die("'--well-formed' and '--strict' can't be used together\n");
    if (!in_path && isatty(fileno(stdin))) { err("no input specified, use \"ennaf -h\" for help\n"); exit(0); }

    int fmt_ext = NAFGPU_FMT_AUTO;                             // ennaf.c:296-306
    if (in_path) {
        const char *ext = in_path + strlen(in_path);
        while (ext > in_path && ext[-1] != '/' && ext[-1] != '\\' && ext[-1] != '.') ext--;
        if (ext > in_path && ext[-1] == '.') fmt_ext = parse_fmt(ext);
    }
    // The input is read in pieces straight into page-locked buffers the library rotates: piece k is on its way to the device
    // while piece k + 1 is being read (a file or a pipe alike; the reference reads through a 16 KB buffer, process.c:227).
    Input in;
    int in_fd = in_path ? open(in_path, O_RDONLY) : 0;
    if (in_fd < 0) die("can't open input file\n");
    if (fstat(in_fd, &in.st) == 0) in.have_stat = in_path != nullptr;

    static std::string auto_path;
    if (!force_stdout && !g_out_path && isatty(fileno(stdout))) {
        if (!in_path) die("output file is not specified\n");
        auto_path = std::string(in_path) + ".naf";
        g_out_path = &auto_path[0];
    }

    nafgpu_enc_opts o; memset(&o, 0, sizeof o);
    o.seq_type = seq_type; o.input_format = fmt_cli; o.no_mask = no_mask; o.strict = strict_mode; o.well_formed = well_formed;
    o.have_line_length = have_line_length; o.line_length = line_length; o.level = level; o.window_log = window_log; o.title = title;
    nafgpu_ctx *ctx = make_ctx();
    const uint8_t *naf = nullptr; size_t naf_size = 0; nafgpu_enc_info info;
    const size_t hint = in.have_stat && S_ISREG(in.st.st_mode) ? (size_t)in.st.st_size : 0;
    if (nafgpu_encode_begin(ctx, &o, hint) != 0) die("%s", nafgpu_last_error(ctx));
    const bool in_regular = in_path && fd_is_regular(in_fd);           // then the pieces are read by a few threads at once
    off_t in_off = 0;
    for (;;) {
        void *buf = nullptr; size_t cap = 0, got = 0;
        if (nafgpu_encode_buffer(ctx, &buf, &cap) != 0) die("%s", nafgpu_last_error(ctx));
        if (in_regular) { got = par_pread(in_fd, (uint8_t *)buf, cap, in_off); in_off += (off_t)got; }
        else while (got < cap) {
            ssize_t k = read(in_fd, (char *)buf + got, cap - got);
            if (k < 0) { if (errno == EINTR) continue; die("can't read input\n"); }
            if (k == 0) break;
            got += (size_t)k;
        }
        if (got && nafgpu_encode_feed(ctx, got) != 0) die("%s", nafgpu_last_error(ctx));
        if (got < cap) break;
    }
    if (in_path) close(in_fd);
    // the file is written as it comes down (nothing file-sized is allocated on the host); an input the library refuses leaves
    // no output file behind: the callback creates it on its first piece
    struct Sink { FILE *f; bool force_stdout; PieceWriter w; } sink{nullptr, force_stdout, {}};
    auto write_piece = [](void *user, const uint8_t *piece, size_t k) -> int {
        Sink *s = (Sink *)user;
        if (!s->f) { s->f = open_output(g_out_path, s->force_stdout); s->w.attach(s->f); }
        s->w.put(piece, k);
        return 0;
    };
    int rc = nafgpu_encode_end_to(ctx, write_piece, &sink, &naf_size, &info);
    if (rc != 0) die("%s", nafgpu_last_error(ctx));
    if (sink.f) sink.w.finish();
    (void)naf;
    if (fmt_ext != NAFGPU_FMT_AUTO && info.format && fmt_ext != info.format) warn("input file extension does not match its actual format\n");
    if (fmt_ext != NAFGPU_FMT_AUTO && fmt_cli != NAFGPU_FMT_AUTO && fmt_ext != fmt_cli) warn("input file extension does not match format specified in the command line\n");

    FILE *out = sink.f ? sink.f : open_output(g_out_path, force_stdout);
    if (verbose) msg("Output line length: %llu\n", have_line_length ? line_length : (unsigned long long)info.longest_line);
    close_output(out, in, in_path && g_out_path && !force_stdout);
    if (!well_formed) {
        static const char *tn[4] = { "DNA", "RNA", "protein", "text" };
        report((const unsigned long long *)info.unexpected[0], "id");
        report((const unsigned long long *)info.unexpected[1], "comment");
        report((const unsigned long long *)info.unexpected[2], tn[seq_type]);
        report((const unsigned long long *)info.unexpected[3], "quality");
    }
    if (verbose) msg("Processed %llu sequences\n", (unsigned long long)info.n_sequences);
    exit_done();
}
