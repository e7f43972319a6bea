/*
 * oracle_cli.c — command-line front end for the CPU ORACLE (test infrastructure only).
 *   naf_oracle ennaf [--dna|--rna|--protein|--text] [--no-mask] [--well-formed] [--strict]
 *                    [--line-length N] [--title T] [--long N] IN OUT      (report -> stderr)
 *   naf_oracle unnaf [--fasta|--fastq|--seq|...] [--no-mask] [--line-length N] IN OUT
 *   naf_oracle zstd-d IN OUT            (multi-frame decode of a raw zstd file)
 * Mirrors the subset of ennaf/unnaf flags that change the bytes produced
 * (ennaf/src/ennaf.c:360-430, unnaf/src/unnaf.c:282-353).
 */
#include "oracle.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static int slurp(const char *path, obuf_t *b)
{
    FILE *f = strcmp(path, "-") ? fopen(path, "rb") : stdin;
    if (!f) { fprintf(stderr, "can't open %s\n", path); return -1; }
    uint8_t buf[1 << 16]; size_t k;
    while ((k = fread(buf, 1, sizeof buf, f)) > 0) obuf_put(b, buf, k);
    if (f != stdin) fclose(f);
    return 0;
}
static int spill(const char *path, const obuf_t *b)
{
    FILE *f = strcmp(path, "-") ? fopen(path, "wb") : stdout;
    if (!f) { fprintf(stderr, "can't create %s\n", path); return -1; }
    fwrite(b->data, 1, b->size, f);
    if (f != stdout) fclose(f);
    return 0;
}

int main(int argc, char **argv)
{
    if (argc < 4) { fprintf(stderr, "usage: naf_oracle ennaf|unnaf|zstd-d [opts] IN OUT\n"); return 2; }
    const char *mode = argv[1], *in = NULL, *outp = NULL;
    onaf_enc_opts eo; memset(&eo, 0, sizeof eo);
    onaf_dec_opts dop; memset(&dop, 0, sizeof dop);
    static const struct { const char *flag; int type; } views[] = {
        {"--format", ONAF_OUT_FORMAT}, {"--part-list", ONAF_OUT_PART_LIST}, {"--sizes", ONAF_OUT_SIZES},
        {"--number", ONAF_OUT_NUMBER}, {"--title", ONAF_OUT_TITLE}, {"--ids", ONAF_OUT_IDS}, {"--names", ONAF_OUT_NAMES},
        {"--lengths", ONAF_OUT_LENGTHS}, {"--total-length", ONAF_OUT_TOTAL_LENGTH}, {"--mask", ONAF_OUT_MASK},
        {"--total-mask-length", ONAF_OUT_TOTAL_MASK_LENGTH}, {"--4bit", ONAF_OUT_4BIT}, {"--seq", ONAF_OUT_SEQ},
        {"--sequences", ONAF_OUT_SEQUENCES}, {"--charcount", ONAF_OUT_CHARCOUNT}, {"--fasta", ONAF_OUT_FASTA},
        {"--fastq", ONAF_OUT_FASTQ} };
    int is_enc = !strcmp(mode, "ennaf");
    for (int i = 2; i < argc; i++) {
        const char *a = argv[i];
        if (!strcmp(a, "--dna")) eo.seq_type = ONAF_DNA;
        else if (!strcmp(a, "--rna")) eo.seq_type = ONAF_RNA;
        else if (!strcmp(a, "--protein")) eo.seq_type = ONAF_PROTEIN;
        else if (!strcmp(a, "--text")) eo.seq_type = ONAF_TEXT;
        else if (!strcmp(a, "--no-mask")) { eo.no_mask = 1; dop.no_mask = 1; }
        else if (!strcmp(a, "--well-formed")) eo.well_formed = 1;
        else if (!strcmp(a, "--strict")) eo.strict = 1;
        else if (!strcmp(a, "--line-length") && i + 1 < argc) { eo.have_line_length = dop.have_line_length = 1; eo.line_length = dop.line_length = strtoull(argv[++i], NULL, 10); }
        else if (is_enc && !strcmp(a, "--title") && i + 1 < argc) eo.title = argv[++i];
        else if (!strcmp(a, "--long") && i + 1 < argc) eo.window_log = atoi(argv[++i]);
        else if (a[0] == '-' && a[1] == '-') {
            int found = 0;
            for (size_t k = 0; k < sizeof views / sizeof views[0]; k++) if (!strcmp(a, views[k].flag)) { dop.out_type = views[k].type; found = 1; }
            if (!found) { fprintf(stderr, "unknown option %s\n", a); return 2; }
        }
        else if (!in) in = a; else outp = a;
    }
    if (!in || !outp) { fprintf(stderr, "need IN and OUT\n"); return 2; }
    obuf_t src, dst, rep; obuf_init(&src); obuf_init(&dst); obuf_init(&rep);
    char err[256] = "";
    if (slurp(in, &src)) return 1;
    int rc;
    if (is_enc) { rc = onaf_encode(src.data, src.size, &eo, &dst, &rep, err); if (rc) fprintf(stderr, "ennaf error: %s", err); fwrite(rep.data, 1, rep.size, stderr); }
    else if (!strcmp(mode, "unnaf")) { rc = onaf_decode(src.data, src.size, &dop, &dst, err); if (rc) fprintf(stderr, "unnaf error: %s", err); }
    else { rc = ozstd_decompress(src.data, src.size, &dst, err); if (rc) fprintf(stderr, "zstd error: %s\n", err); }
    if (!rc) rc = spill(outp, &dst);
    return rc ? 1 : 0;
}
