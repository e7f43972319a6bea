// unnaf — drop-in command line of the reference's decompressor (unnaf/src/unnaf.c:282-447), host side only:
// argument parsing, output naming, stat transfer, and the text formatting of the views whose payload is a
// header field or a tiny stream (output.c:7-260).  Sequence-sized views are one call to nafgpu_decode_to(), which delivers
// the text in pieces as it comes down from the device.
#include "cli_common.hpp"

enum View { UNDECIDED, FORMAT_NAME, PART_LIST, PART_SIZES, NUMBER, TITLE, IDS, NAMES, LENGTHS, TOTAL_LENGTH, MASK, TOTAL_MASK_LENGTH,
            FOUR_BIT, DNA, MASKED_DNA, UNMASKED_DNA, SEQ, SEQUENCES, CHARCOUNT, FASTA, MASKED_FASTA, UNMASKED_FASTA, FASTQ };
static View view = UNDECIDED;
static bool use_mask = true, force_stdout = false, verbose = false;
static char *in_path = nullptr;
static bool have_line_length = false; static unsigned long long line_length = 0;

static void set_view(View v) { if (view != UNDECIDED) die("only one output type should be specified\n"); view = v; }

static bool get_vle(const uint8_t *p, size_t n, size_t &pos, unsigned long long &v)       // utils.c:117 read_number
{
    unsigned long long a = 0;
    if (pos >= n) die("incomplete or truncated input\n");
    uint8_t c = p[pos++];
    if (c == 128) die("invalid input: error parsing variable length encoded number\n");
    while (c & 128) {
        if (a & (127ull << 57)) { fputs("invalid input: overflow reading a variable length encoded number\n", stderr); exit(1); }
        a = (a << 7) | (c & 127);
        if (pos >= n) die("incomplete or truncated input\n");
        c = p[pos++];
    }
    if (a & (127ull << 57)) { fputs("invalid input: overflow reading a variable length encoded number\n", stderr); exit(1); }
    v = (a << 7) | c;
    return true;
}

struct Hdr {
    int version = 1, seq_type = 0; bool title = false, ids = false, names = false, lengths = false, mask = false, data = false, quality = false;
    unsigned long long line_length = 0, N = 0, title_off = 0, title_len = 0;
    struct { unsigned long long orig = 0, comp = 0; bool present = false; } sec[6];
};
static const char *type_name[4] = { "DNA", "RNA", "protein", "text" };

static void read_header(const uint8_t *p, size_t n, Hdr &h, bool sections)               // input.c:31
{
    if (n == 0) die("empty input");
    if (n < 3) die("incomplete or truncated input\n");
    if (p[0] != 0x01 || p[1] != 0xF9 || p[2] != 0xEC) die("not a NAF format\n");
    size_t pos = 3;
    if (pos >= n) die("incomplete or truncated input\n");
    h.version = p[pos++];
    if (h.version < 1 || h.version > 2) die("unknown version (%d) of NAF format\n", h.version);
    if (h.version > 1) {
        if (pos >= n) die("incomplete or truncated input\n");
        int t = p[pos++];
        if (t < 1 || t > 3) die("unknown sequence type (%d) found in NAF file\n", t);
        h.seq_type = t;
    }
    if (pos + 2 > n) die("incomplete or truncated input\n");
    int f = p[pos++];
    h.title = (f >> 6) & 1; h.ids = (f >> 5) & 1; h.names = (f >> 4) & 1; h.lengths = (f >> 3) & 1; h.mask = (f >> 2) & 1; h.data = (f >> 1) & 1; h.quality = f & 1;
    int sep = p[pos++];
    if (sep < 0x20 || sep > 0x7E) die("unsupported name separator character\n");
    if (!sections) return;
    get_vle(p, n, pos, h.line_length); get_vle(p, n, pos, h.N);
    if (h.title) { get_vle(p, n, pos, h.title_len); if (h.title_len > n - pos) die("incomplete or truncated input\n"); h.title_off = pos; pos += h.title_len; }
    const bool present[6] = { h.ids, h.names, h.lengths, h.mask, h.data, h.quality };
    for (int k = 0; k < 6; k++) {
        if (!present[k]) continue;
        if (h.N == 0 && pos >= n) break;
        get_vle(p, n, pos, h.sec[k].orig); get_vle(p, n, pos, h.sec[k].comp);
        if (h.sec[k].comp > n - pos) die("incomplete or truncated input\n");
        h.sec[k].present = true; pos += h.sec[k].comp;
    }
}

static void show_help()
{
    msg("Usage: unnaf [OUTPUT-TYPE] [file.naf]\n"
        "Options for selecting output type:\n"
        "  --format        - File format version\n"
        "  --part-list     - List of parts\n"
        "  --sizes         - Part sizes\n"
        "  --number        - Number of sequences\n"
        "  --title         - Dataset title\n"
        "  --ids           - Sequence ids (accession numbers)\n"
        "  --names         - Full sequence names (including ids)\n"
        "  --lengths       - Sequence lengths\n"
        "  --total-length  - Sum of sequence lengths\n"
        "  --mask          - Masked region lengths\n"
        "  --4bit          - 4bit-encoded nucleotide sequence (binary data)\n"
        "  --seq           - Continuous concatenated sequence\n"
        "  --sequences     - One sequence per line, no names\n"
        "  --fasta         - FASTA-formatted sequences\n"
        "  --fastq         - FASTQ-formatted sequences\n"
        "Other options:\n"
        "  -o FILE         - Decompress into FILE\n"
        "  -c              - Write to standard output\n"
        "  --line-length N - Use lines of width N for FASTA output\n"
        "  --no-mask       - Ignore mask\n"
        "  --binary-stdout - Set stdout stream to binary mode.\n"
        "  --binary-stderr - Set stderr stream to binary mode.\n"
        "  --binary        - Shortcut for \"--binary-stdout --binary-stderr\"\n"
        "  -h, --help      - Show help\n"
        "  -V, --version   - Show version\n");
}

int main(int argc, char **argv)
{
    g_tool = "unnaf";
    atexit(cleanup_output);
    bool print_version = false;
    static const struct { const char *flag; View v; } views[] = {
        {"--format", FORMAT_NAME}, {"--part-list", PART_LIST}, {"--sizes", PART_SIZES}, {"--number", NUMBER}, {"--title", TITLE}, {"--ids", IDS},
        {"--names", NAMES}, {"--lengths", LENGTHS}, {"--total-length", TOTAL_LENGTH}, {"--mask", MASK}, {"--total-mask-length", TOTAL_MASK_LENGTH},
        {"--4bit", FOUR_BIT}, {"--seq", SEQ}, {"--sequences", SEQUENCES}, {"--charcount", CHARCOUNT}, {"--fasta", FASTA}, {"--fastq", FASTQ},
        {"--dna", DNA}, {"--masked-dna", MASKED_DNA}, {"--unmasked-dna", UNMASKED_DNA}, {"--masked-fasta", MASKED_FASTA}, {"--unmasked-fasta", UNMASKED_FASTA} };
    for (int i = 1; i < argc; i++) {
        char *a = argv[i];
        if (a[0] == '-') {
            if (a[1] == '-') {
                if (i < argc - 1 && !strcmp(a, "--line-length")) {
                    i++;
                    char *end; long long v = strtoll(argv[i], &end, 10);
                    if (*end) die("can't parse the value of --line-length parameter\n");
                    if (v < 0) die("negative line length specified\n");
                    if (!parse_ull_strict(argv[i], line_length)) die("can't parse the value of --line-length parameter\n");
                    have_line_length = true; continue;
                }
                bool found = false;
                for (auto &v : views) if (!strcmp(a, v.flag)) { set_view(v.v); found = true; break; }
                if (found) continue;
                if (!strcmp(a, "--no-mask")) { use_mask = false; continue; }
                if (!strcmp(a, "--binary-stdout") || !strcmp(a, "--binary-stderr") || !strcmp(a, "--binary")) continue;
                if (!strcmp(a, "--help")) { show_help(); exit(0); }
                if (!strcmp(a, "--verbose")) { verbose = true; continue; }
                if (!strcmp(a, "--version")) { print_version = true; continue; }
            }
            if (i < argc - 1 && !strcmp(a, "-o")) { i++; if (g_out_path) die("double --out parameter\n"); if (!*argv[i]) die("empty --out parameter\n"); g_out_path = argv[i]; continue; }
            if (!strcmp(a, "-c")) { force_stdout = true; continue; }
            if (!strcmp(a, "-h")) { show_help(); exit(0); }
            if (!strcmp(a, "-V")) { print_version = true; continue; }
            die("unknown or incomplete argument \"%s\"\n", a);
        }
        if (in_path) die("can process only one file at a time\n");
        if (!*a) die("empty input path specified\n");
        in_path = a;
    }
    if (print_version) { msg("unnaf - NAF decompressor for NVIDIA B200 (naf-b200 %s), reads NAF format of unnaf 1.3.0\n", nafgpu_version()); exit(0); }
    if (force_stdout && g_out_path) die("-c and -o arguments can't be used together\n");
    if (!in_path && isatty(fileno(stdin))) { err("no input specified, use \"unnaf -h\" for help\n"); exit(0); }
    (void)verbose;

    Input in; load_input(in_path, in);
    Hdr h; read_header(in.data, in.size, h, false);
    if (view == UNDECIDED) view = h.quality ? FASTQ : FASTA;
    if ((view == DNA || view == MASKED_DNA || view == UNMASKED_DNA) && h.seq_type != 0) die("input has not DNA, but %s data\n", type_name[h.seq_type]);
    if (view == FOUR_BIT && h.seq_type >= 2) die("input has no 4-bit encoded data, but %s sequences\n", type_name[h.seq_type]);

    // output file naming / TTY policy (unnaf/src/files.c:38-87)
    static std::string auto_path;
    const bool to_original = h.quality ? (view == FASTA) : (view == FASTQ);
    const bool large = view == IDS || view == NAMES || view == LENGTHS || view == MASK || view == FOUR_BIT || view == DNA || view == MASKED_DNA ||
                       view == UNMASKED_DNA || view == SEQ || view == FASTA || view == MASKED_FASTA || view == UNMASKED_FASTA || view == FASTQ;
    if (to_original && !force_stdout && in_path && !g_out_path && isatty(fileno(stdout))) {
        size_t len = strlen(in_path);
        if (len > 4 && !strcmp(in_path + len - 4, ".naf") && in_path[len - 5] != '/' && in_path[len - 5] != '\\') { auto_path.assign(in_path, len - 4); g_out_path = &auto_path[0]; }
    }
    FILE *out = open_output(g_out_path, force_stdout);
    if (large && !force_stdout && isatty(fileno(out)))
        die("output file not specified - please either specify output file with '-o' or '>', or use '-c' option to force writing to console\n");

    auto gpu_view = [&](int type, bool mask_on, const uint8_t **p, size_t *n) {
        nafgpu_dec_opts o; memset(&o, 0, sizeof o);
        o.out_type = type; o.no_mask = !mask_on; o.have_line_length = have_line_length; o.line_length = line_length;
        static nafgpu_ctx *ctx = nullptr;
        if (!ctx) ctx = make_ctx();
        if (nafgpu_decode(ctx, in.data, in.size, &o, p, n) != 0) die("%s", nafgpu_last_error(ctx));
    };
    // the text-sized views are written as they come down: nafgpu_decode_to hands over pieces of a few tens of MB in order, each
    // written to the file while the next is on its way from the device (the reference writes through 128 KB buffers, output.c:640)
    auto gpu_stream = [&](int type, bool mask_on) {
        nafgpu_dec_opts o; memset(&o, 0, sizeof o);
        o.out_type = type; o.no_mask = !mask_on; o.have_line_length = have_line_length; o.line_length = line_length;
        nafgpu_ctx *ctx = make_ctx();
        size_t total = 0;
        PieceWriter w; w.attach(out);
        auto sink = [](void *user, const uint8_t *piece, size_t k) -> int { ((PieceWriter *)user)->put(piece, k); return 0; };
        if (nafgpu_decode_to(ctx, in.data, in.size, &o, sink, &w, &total) != 0) die("%s", nafgpu_last_error(ctx));
        w.finish();
    };
    const uint8_t *p = nullptr; size_t n = 0;

    if (view == FORMAT_NAME) fprintf(out, "%s sequences%s in NAF format version %d\n", type_name[h.seq_type], h.quality ? " with qualities" : "", h.version);
    else if (view == PART_LIST) {
        const char *names[7] = { "Title", "IDs", "Names", "Lengths", "Mask", "Data", "Quality" };
        const bool present[7] = { h.title, h.ids, h.names, h.lengths, h.mask, h.data, h.quality };
        int printed = 0;
        for (int k = 0; k < 7; k++) if (present[k]) { fprintf(out, "%s%s", printed ? ", " : "", names[k]); printed++; }
        fprintf(out, "\n");
    } else {
        read_header(in.data, in.size, h, true);
        if (view == NUMBER) fprintf(out, "%llu\n", h.N);
        else if (view == PART_SIZES) {
            const char *names[6] = { "IDs", "Names", "Lengths", "Mask", "Data", "Quality" };
            if (h.title) fprintf(out, "Title: %llu\n", h.title_len);
            for (int k = 0; k < 6; k++) if (h.sec[k].present)
                fprintf(out, "%s: %llu / %llu (%.3f%%)\n", names[k], h.sec[k].comp, h.sec[k].orig, (double)h.sec[k].comp / (double)h.sec[k].orig * 100);
        }
        else if (view == TITLE) { if (h.title) fwrite(in.data + h.title_off, 1, h.title_len, out); fputc('\n', out); }
        else if (h.N != 0) {
            switch (view) {
            case IDS: gpu_stream(NAFGPU_OUT_IDS, true); break;
            case NAMES: gpu_stream(NAFGPU_OUT_NAMES, true); break;
            case LENGTHS: {                                                                  // output.c:180
                if (!h.lengths) break;
                gpu_view(NAFGPU_OUT_LENGTHS, true, &p, &n);
                const uint32_t *u = (const uint32_t *)p; size_t nu = n / 4;
                for (size_t i = 0; i < nu; i++) {
                    unsigned long long len = 0;
                    while (i < nu && u[i] == 4294967295u) { len += 4294967295llu; i++; }
                    if (i < nu) len += u[i];
                    fprintf(out, "%llu\n", len);
                }
                break;
            }
            case TOTAL_LENGTH: if (h.lengths) fprintf(out, "%llu\n", h.sec[4].orig); break;
            case MASK: case TOTAL_MASK_LENGTH: {                                             // output.c:222,246
                unsigned long long total = 0;
                if (h.mask) {
                    gpu_view(NAFGPU_OUT_MASK, true, &p, &n);
                    for (size_t i = 0; i < n; i++) {
                        if (view == TOTAL_MASK_LENGTH) { total += p[i]; continue; }
                        unsigned long long len = 0;
                        while (i < n && p[i] == 255) { len += 255; i++; }
                        if (i < n) len += p[i];
                        fprintf(out, "%llu\n", len);
                    }
                }
                if (view == TOTAL_MASK_LENGTH) fprintf(out, "%llu\n", total);
                break;
            }
            case FOUR_BIT: gpu_stream(NAFGPU_OUT_4BIT, true); break;
            case DNA: case SEQ: case MASKED_DNA: gpu_stream(NAFGPU_OUT_SEQ, use_mask); break;
            case UNMASKED_DNA: gpu_stream(NAFGPU_OUT_SEQ, false); break;
            case SEQUENCES: gpu_stream(NAFGPU_OUT_SEQUENCES, use_mask); break;
            case CHARCOUNT: {                                                                // output.c:596-598
                if (!h.data) break;
                gpu_view(NAFGPU_OUT_CHARCOUNT, use_mask, &p, &n);
                if (n < 256 * 8) break;
                const unsigned long long *c = (const unsigned long long *)p;
                for (unsigned i = 0; i < 33; i++) if (c[i]) fprintf(out, "\\x%02X\t%llu\n", i, c[i]);
                for (unsigned i = 33; i < 127; i++) if (c[i]) fprintf(out, "%c\t%llu\n", (unsigned char)i, c[i]);
                for (unsigned i = 127; i < 256; i++) if (c[i]) fprintf(out, "\\x%02X\t%llu\n", i, c[i]);
                break;
            }
            case FASTA: case MASKED_FASTA: gpu_stream(NAFGPU_OUT_FASTA, use_mask); break;
            case UNMASKED_FASTA: gpu_stream(NAFGPU_OUT_FASTA, false); break;
            case FASTQ:
                if (!h.quality) die("FASTQ output requested, but input has no qualities\n");
                gpu_stream(NAFGPU_OUT_FASTQ, false); break;
            default: die("unknown output requested\n");
            }
        }
    }
    close_output(out, in, in_path && g_out_path && !force_stdout);
    exit_done();
}
