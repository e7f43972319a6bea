/*
 * naf_oracle.c — CPU ORACLE (test infrastructure only; see oracle.h).
 *
 * Restates, as explicit byte-level state machines over an in-memory buffer, what the reference's
 * streaming parser / encoders / container writer (ennaf) and container reader / text writers
 * (unnaf) compute.  The same state-machine formulation is what the CUDA kernels implement, so a
 * disagreement between this file and oracle/_ref binaries is a bug in the *formulation*.
 * Paths cited are relative to /root/reference.
 */
#include "oracle.h"

#include <ctype.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ character classes */
/* ennaf/src/tables.c:28-137, restated as predicates instead of 257-entry tables.  c may be 256 (EOF). */

static int is_eol(unsigned c)   { return c >= 0x0A && c <= 0x0D; }                    /* tables.c:28 */
static int is_space(unsigned c) { return (c >= 0x09 && c <= 0x0D) || c == 0x20; }     /* tables.c:47 */

static int unexpected_text(unsigned c)    { return c <= 32 || c == 127 || c >= 255; } /* tables.c:115 */
static int unexpected_comment(unsigned c) { return c < 32 || c == 127 || c >= 255; }  /* tables.c:126 */
static int unexpected_qual(unsigned c)    { return c < 33 || c > 126; }               /* tables.c:137 */

static int in_set(unsigned c, const char *set)
{
    if (c == 0 || c > 255) return 0;
    if (c >= 'a' && c <= 'z') c -= 32;
    return strchr(set, (int)c) != NULL;
}
static int unexpected_dna(unsigned c)     { return !in_set(c, "-ABCDGHKMNRSTVWY"); }   /* tables.c:72 */
static int unexpected_rna(unsigned c)     { return !in_set(c, "-ABCDGHKMNRSUVWY"); }   /* tables.c:82 */
static int unexpected_protein(unsigned c) { return !in_set(c, "*-ABCDEFGHIJKLMNOPQRSTUVWXYZ"); } /* tables.c:104 */

/* tables.c:189 nuc_code: '-'0 T/U 1 G2 K3 C4 Y5 S6 B7 A8 W9 R10 D11 M12 H13 V14, everything else 15 */
static uint8_t nuc_code(uint8_t c)
{
    static const char order[] = "-TGKCYSBAWRDMHV";
    if (c >= 'a' && c <= 'z') c -= 32;
    if (c == 'U') return 1;
    const char *p = c ? strchr(order, c) : NULL;
    return p ? (uint8_t)(p - order) : 15;
}

static const char CODE_TO_NUC[16] = { '-','T','G','K','C','Y','S','B','A','W','R','D','M','H','V','N' }; /* unnaf.c:13 */

/* ------------------------------------------------------------------ VLE numbers */

void onaf_put_vle(obuf_t *b, uint64_t v)          /* encoders.c:175 */
{
    uint8_t tmp[10]; int n = 0;
    tmp[n++] = (uint8_t)(v & 127); v >>= 7;
    while (v) { tmp[n++] = (uint8_t)(128 | (v & 127)); v >>= 7; }
    while (n) obuf_putc(b, tmp[--n]);
}

int onaf_get_vle(const uint8_t *p, size_t n, size_t *pos, uint64_t *v, char *err)   /* unnaf utils.c:117 */
{
    uint64_t a = 0;
    if (*pos >= n) { snprintf(err, 256, "incomplete or truncated input\n"); return -1; }
    uint8_t c = p[(*pos)++];
    if (c == 128) { snprintf(err, 256, "invalid input: error parsing variable length encoded number\n"); return -1; }
    while (c & 128) {
        if (a & (127ull << 57)) { snprintf(err, 256, "invalid input: overflow reading a variable length encoded number\n"); return -2; }
        a = (a << 7) | (c & 127);
        if (*pos >= n) { snprintf(err, 256, "incomplete or truncated input\n"); return -1; }
        c = p[(*pos)++];
    }
    if (a & (127ull << 57)) { snprintf(err, 256, "invalid input: overflow reading a variable length encoded number\n"); return -2; }
    *v = (a << 7) | c;
    return 0;
}

/* ------------------------------------------------------------------ stage transforms */

void onaf_pack4(const uint8_t *bases, size_t n, obuf_t *out)   /* encoders.c:30-69 + ennaf.c:525 */
{
    for (size_t i = 0; i + 1 < n; i += 2) obuf_putc(out, (uint8_t)(nuc_code(bases[i]) | (nuc_code(bases[i + 1]) << 4)));
    if (n & 1) obuf_putc(out, nuc_code(bases[n - 1]));
}

void onaf_unpack4(const uint8_t *packed, size_t n_bases, int rna, obuf_t *out)   /* utils.c:74, output.c:445 */
{
    for (size_t i = 0; i < n_bases; i++) {
        uint8_t code = (i & 1) ? packed[i >> 1] >> 4 : packed[i >> 1] & 15;
        char c = CODE_TO_NUC[code];
        if (rna && code == 1) c = 'U';                       /* unnaf.c:369 */
        obuf_putc(out, (uint8_t)c);
    }
}

static void put_run(obuf_t *out, uint64_t len)                 /* encoders.c:98 add_mask */
{
    while (len >= 255) { obuf_putc(out, 255); len -= 255; }
    obuf_putc(out, (uint8_t)len);
}

void onaf_mask_rle(const uint8_t *bases, size_t n, obuf_t *out)   /* encoders.c:126 + ennaf.c:511 */
{
    int on = 0; uint64_t run = 0;
    for (size_t i = 0; i < n; i++) {
        int m = bases[i] >= 96;
        if (m != on) { put_run(out, run); run = 0; on = m; }
        run++;
    }
    if (run > 0) put_run(out, run);
}

void onaf_mask_apply(uint8_t *bases, size_t n, const uint8_t *units, size_t n_units)   /* output.c:295, input.c:236 */
{
    /* two-scan formulation: unit k covers [start_k, start_k+u_k) and is "on" iff an
     * odd number of non-255 units precede it. */
    size_t pos = 0; int on = 0;
    for (size_t k = 0; k < n_units && pos < n; k++) {
        size_t u = units[k];
        size_t end = pos + u > n ? n : pos + u;
        if (on) for (size_t i = pos; i < end; i++) bases[i] = (uint8_t)(bases[i] + 32);
        pos = end;
        if (units[k] != 255) on = !on;
    }
}

static void put_length(obuf_t *out, uint64_t len)              /* encoders.c:72 add_length */
{
    while (len >= 0xFFFFFFFFull) { uint32_t u = 0xFFFFFFFFu; obuf_put(out, &u, 4); len -= 0xFFFFFFFFull; }
    uint32_t u = (uint32_t)len; obuf_put(out, &u, 4);
}

/* ------------------------------------------------------------------ parser */

void onaf_streams_init(onaf_streams *s)
{
    memset(s, 0, sizeof(*s));
    obuf_init(&s->ids); obuf_init(&s->comm); obuf_init(&s->len); obuf_init(&s->mask); obuf_init(&s->seq); obuf_init(&s->qual);
}
void onaf_streams_free(onaf_streams *s)
{
    obuf_free(&s->ids); obuf_free(&s->comm); obuf_free(&s->len); obuf_free(&s->mask); obuf_free(&s->seq); obuf_free(&s->qual);
}

static int perr(char *err, const char *fmt, ...)
{
    va_list ap; va_start(ap, fmt); vsnprintf(err, 256, fmt, ap); va_end(ap);
    return -1;
}

typedef int (*pred_t)(unsigned);

typedef struct {
    const onaf_enc_opts *o;
    onaf_streams *s;
    obuf_t bases;              /* concatenated sequence bytes before pack/mask (seq.writer input) */
    pred_t unexpected_seq;
    uint8_t repl;
    int text_fasta;            /* ennaf.c:466: '>' becomes "unexpected" for --text FASTA */
    char *err;
} pctx;

static int seq_unexpected(const pctx *p, unsigned c)
{
    if (p->text_fasta && c == '>') return 1;
    return p->unexpected_seq(c);
}

/* process.c:358 process_non_well_formed_fasta, as a 4-state machine. `i` is just past the '>' . */
static int parse_fasta(pctx *p, const uint8_t *t, size_t n, size_t i)
{
    enum { NAME, COMMENT, SEQ_LS, SEQ_MID } st = NAME;
    onaf_streams *s = p->s;
    uint64_t rec_len = 0, line = 0;
    const int wf = p->o->well_formed;
    for (; i <= n; i++) {
        unsigned c = i < n ? t[i] : 256;
        if (c == 256) break;
        switch (st) {
        case NAME:
            if (wf) {                                                   /* process.c:318 */
                if (c == '\n') { obuf_putc(&s->ids, 0); obuf_putc(&s->comm, 0); st = SEQ_LS; }
                else if (c == ' ') { obuf_putc(&s->ids, 0); st = COMMENT; }
                else obuf_putc(&s->ids, (uint8_t)c);
                break;
            }
            /* ennaf.c:466 flips '>' in the table the name scan also uses, so for --text FASTA a
             * '>' inside an id is an unexpected id character. */
            if (!unexpected_text(c) && !(p->text_fasta && c == '>')) obuf_putc(&s->ids, (uint8_t)c);
            else if (is_space(c)) {
                obuf_putc(&s->ids, 0);
                if (is_eol(c)) { obuf_putc(&s->comm, 0); st = SEQ_LS; } else st = COMMENT;
            } else {
                /* process.c:366: the '?' replacement goes to the *sequence* buffer and is not
                 * counted in any record length (reference bug) — restated as is. */
                if (p->o->strict) return perr(p->err, "unexpected character '%c' in ID of sequence %llu\n", (unsigned char)c, (unsigned long long)s->n_sequences + 1);
                s->unexpected[0][c]++; obuf_putc(&p->bases, '?');
            }
            break;
        case COMMENT:
            if (wf) {
                if (c == '\n') { obuf_putc(&s->comm, 0); st = SEQ_LS; } else obuf_putc(&s->comm, (uint8_t)c);
                break;
            }
            if (!unexpected_comment(c)) obuf_putc(&s->comm, (uint8_t)c);
            else if (is_eol(c)) { obuf_putc(&s->comm, 0); st = SEQ_LS; }
            else {
                if (p->o->strict) return perr(p->err, "unexpected character '%c' in comment of sequence %llu\n", (unsigned char)c, (unsigned long long)s->n_sequences + 1);
                s->unexpected[1][c]++; obuf_putc(&s->comm, '?');
            }
            break;
        case SEQ_LS:
        case SEQ_MID:
            if (st == SEQ_LS && c == '>') {                             /* record boundary */
                put_length(&s->len, rec_len); s->n_sequences++;
                rec_len = 0; line = 0; st = NAME;
                break;
            }
            if (wf) {                                                   /* process.c:333-343 */
                if (c == '\n') { if (line > s->longest_line) s->longest_line = line; line = 0; st = SEQ_LS; }
                else { obuf_putc(&p->bases, (uint8_t)c); rec_len++; line++; st = SEQ_MID; }
                break;
            }
            if (!seq_unexpected(p, c)) { obuf_putc(&p->bases, (uint8_t)c); rec_len++; line++; st = SEQ_MID; }
            else if (is_eol(c)) { if (line > s->longest_line) s->longest_line = line; line = 0; st = SEQ_LS; }
            else if (is_space(c)) { st = SEQ_MID; }
            else if (c == '>' && p->text_fasta) { obuf_putc(&p->bases, (uint8_t)c); rec_len++; line++; st = SEQ_MID; }  /* process.c:413 */
            else {
                if (p->o->strict) return perr(p->err, "unexpected %s code '%c' in sequence %llu\n",
                    (const char *[]){ "DNA", "RNA", "protein", "text" }[p->o->seq_type], (unsigned char)c, (unsigned long long)s->n_sequences + 1);
                s->unexpected[2][c]++; obuf_putc(&p->bases, p->repl); rec_len++; line++; st = SEQ_MID;
            }
            break;
        }
    }
    /* end of input */
    if (st == NAME) { obuf_putc(&s->ids, 0); obuf_putc(&s->comm, 0); }
    else if (st == COMMENT) obuf_putc(&s->comm, 0);
    else if (line > s->longest_line) s->longest_line = line;           /* process.c:417-422 */
    put_length(&s->len, rec_len); s->n_sequences++;
    return 0;
}

/* process.c:477 process_non_well_formed_fastq / :430 well-formed, as an 8-state machine. */
static int parse_fastq(pctx *p, const uint8_t *t, size_t n, size_t i)
{
    enum { NAME, COMMENT, SEQ, AFTER_SEQ, PLUS, BEFORE_QUAL, QUAL, AFTER_QUAL } st = NAME;
    onaf_streams *s = p->s;
    uint64_t read_len = 0, qual_len = 0;
    const int wf = p->o->well_formed;
    static const char *no_qual = "truncated FASTQ input: last sequence has no quality\n";
    for (;; i++) {
        unsigned c = i < n ? t[i] : 256;
        switch (st) {
        case NAME:
            if (c == 256) return perr(p->err, "truncated FASTQ input: last sequence has no sequence data\n");
            if (wf) {
                if (c == '\n') { obuf_putc(&s->ids, 0); obuf_putc(&s->comm, 0); st = SEQ; read_len = 0; }
                else if (c == ' ') { obuf_putc(&s->ids, 0); st = COMMENT; }
                else obuf_putc(&s->ids, (uint8_t)c);
                break;
            }
            if (!unexpected_text(c)) obuf_putc(&s->ids, (uint8_t)c);
            else if (is_space(c)) {
                obuf_putc(&s->ids, 0);
                if (is_eol(c)) { obuf_putc(&s->comm, 0); st = SEQ; read_len = 0; } else st = COMMENT;
            } else {
                if (p->o->strict) return perr(p->err, "unexpected character '%c' in ID of sequence %llu\n", (unsigned char)c, (unsigned long long)s->n_sequences + 1);
                s->unexpected[0][c]++; obuf_putc(&p->bases, '?');      /* process.c:485, same bug as FASTA */
            }
            break;
        case COMMENT:
            if (c == 256) return perr(p->err, "truncated FASTQ input: last sequence has no sequence data\n");
            if (wf) {
                if (c == '\n') { obuf_putc(&s->comm, 0); st = SEQ; read_len = 0; } else obuf_putc(&s->comm, (uint8_t)c);
                break;
            }
            if (!unexpected_comment(c)) obuf_putc(&s->comm, (uint8_t)c);
            else if (is_eol(c)) { obuf_putc(&s->comm, 0); st = SEQ; read_len = 0; }
            else {
                if (p->o->strict) return perr(p->err, "unexpected character '%c' in comment of sequence %llu\n", (unsigned char)c, (unsigned long long)s->n_sequences + 1);
                s->unexpected[1][c]++; obuf_putc(&s->comm, '?');
            }
            break;
        case SEQ:
            if (c == 256) return perr(p->err, "%s", no_qual);
            if (wf) {
                if (c == '\n') { if (read_len > s->longest_line) s->longest_line = read_len; st = AFTER_SEQ; }
                else { obuf_putc(&p->bases, (uint8_t)c); read_len++; }
                break;
            }
            if (!p->unexpected_seq(c)) { obuf_putc(&p->bases, (uint8_t)c); read_len++; }
            else if (is_eol(c)) { if (read_len > s->longest_line) s->longest_line = read_len; st = AFTER_SEQ; }
            else if (is_space(c)) {}
            else {
                if (p->o->strict) return perr(p->err, "unexpected %s code '%c' in sequence %llu\n",
                    (const char *[]){ "DNA", "RNA", "protein", "text" }[p->o->seq_type], (unsigned char)c, (unsigned long long)s->n_sequences + 1);
                s->unexpected[2][c]++; obuf_putc(&p->bases, p->repl); read_len++;
            }
            break;
        case AFTER_SEQ:
            if (wf) {                                                   /* process.c:449-458 */
                if (c == '+') {
                    unsigned c2 = i + 1 < n ? t[i + 1] : 256;
                    if (c2 != '\n') return perr(p->err, "not well-formed FASTQ input\n");
                    i++; st = BEFORE_QUAL; qual_len = 0;
                    break;
                }
                if (c == 256) return perr(p->err, "%s", no_qual);
                return perr(p->err, "not well-formed FASTQ input\n");
            }
            if (is_eol(c)) break;
            if (c == 256) return perr(p->err, "%s", no_qual);
            if (c != '+') return perr(p->err, "invalid FASTQ input: can't find '+' line of sequence %llu\n", (unsigned long long)s->n_sequences + 1);
            st = PLUS;
            break;
        case PLUS:
            if (c == 256) return perr(p->err, "%s", no_qual);
            if (is_eol(c)) { st = BEFORE_QUAL; qual_len = 0; }
            break;
        case BEFORE_QUAL:
            if (wf) {                                                   /* process.c:460 */
                st = QUAL; i--;                                         /* re-read c as quality */
                break;
            }
            if (is_eol(c)) break;
            if (c == 256) return perr(p->err, "%s", no_qual);
            obuf_putc(&s->qual, (uint8_t)c); qual_len++;               /* process.c:523: unvalidated */
            st = QUAL;
            break;
        case QUAL: {
            int end = 0;
            if (wf) {
                if (c == '\n' || c == 256) end = 1; else { obuf_putc(&s->qual, (uint8_t)c); qual_len++; }
            } else {
                if (c == 256) end = 1;
                else if (!unexpected_qual(c)) { obuf_putc(&s->qual, (uint8_t)c); qual_len++; }
                else if (is_eol(c)) end = 1;
                else if (is_space(c)) {}
                else {
                    if (p->o->strict) return perr(p->err, "unexpected quality code '%c' in sequence %llu\n", (unsigned char)c, (unsigned long long)s->n_sequences + 1);
                    s->unexpected[3][c]++; obuf_putc(&s->qual, '!'); qual_len++;
                }
            }
            if (end) {
                if (qual_len != read_len) {
                    if (wf) return perr(p->err, "quality length of sequence %llu doesn't match sequence length\n", (unsigned long long)s->n_sequences + 1);
                    return perr(p->err, "quality length of sequence %llu (%llu) doesn't match sequence length (%llu)\n",
                                (unsigned long long)s->n_sequences + 1, (unsigned long long)qual_len, (unsigned long long)read_len);
                }
                put_length(&s->len, read_len); s->n_sequences++;
                if (c == 256) return 0;
                st = AFTER_QUAL;
            }
            break;
        }
        case AFTER_QUAL:
            if (c == 256) return 0;
            if (wf) {
                if (c != '@') return perr(p->err, "not well-formed FASTQ input\n");
                st = NAME; break;
            }
            if (is_eol(c)) break;
            if (c != '@') return perr(p->err, "invalid FASTQ input: Can't find '@' after sequence %llu\n", (unsigned long long)s->n_sequences);
            st = NAME;
            break;
        }
    }
}

int onaf_split(const uint8_t *text, size_t n, const onaf_enc_opts *o, onaf_streams *s, char *err)
{
    pctx p; memset(&p, 0, sizeof(p));
    p.o = o; p.s = s; p.err = err; obuf_init(&p.bases);
    err[0] = 0;
    switch (o->seq_type) {                                              /* ennaf.c:447-470 */
    case ONAF_DNA:     p.unexpected_seq = unexpected_dna;     p.repl = 'N'; break;
    case ONAF_RNA:     p.unexpected_seq = unexpected_rna;     p.repl = 'N'; break;
    case ONAF_PROTEIN: p.unexpected_seq = unexpected_protein; p.repl = 'X'; break;
    default:           p.unexpected_seq = unexpected_text;    p.repl = '?'; break;
    }
    s->store_mask = !(o->no_mask || o->seq_type >= ONAF_PROTEIN);       /* ennaf.c:445 */

    /* process.c:547 confirm_input_format */
    size_t i = 0; unsigned last = '\n';
    while (i < n && is_space(text[i])) { last = text[i]; i++; }
    int rc = 0;
    if (i < n) {
        unsigned c = text[i++];
        if (c == '>' && is_eol(last)) s->format = ONAF_FMT_FASTA;
        else if (c == '@' && is_eol(last)) s->format = ONAF_FMT_FASTQ;
        else if (c == '>' || c == '@') rc = perr(err, "invalid input - first '%c' is not at the beginning of the line\n", (unsigned char)c);
        else rc = perr(err, "input data is in unknown format - first non-space character is neither '>' nor '@'\n");
    }
    s->store_qual = s->format == ONAF_FMT_FASTQ;
    p.text_fasta = o->seq_type == ONAF_TEXT && s->format == ONAF_FMT_FASTA;
    if (!rc && s->format == ONAF_FMT_FASTA) rc = parse_fasta(&p, text, n, i);
    if (!rc && s->format == ONAF_FMT_FASTQ) rc = parse_fastq(&p, text, n, i);
    if (!rc) {
        s->seq_size = p.bases.size;
        if (o->seq_type < ONAF_PROTEIN) {                               /* process.c:24-36 */
            if (s->store_mask) onaf_mask_rle(p.bases.data, p.bases.size, &s->mask);
            onaf_pack4(p.bases.data, p.bases.size, &s->seq);
        } else {                                                        /* process.c:39-51 */
            if (o->no_mask) for (size_t k = 0; k < p.bases.size; k++) p.bases.data[k] = (uint8_t)toupper(p.bases.data[k]);
            obuf_put(&s->seq, p.bases.data, p.bases.size);
        }
    }
    obuf_free(&p.bases);
    return rc;
}

void onaf_report_unexpected(const onaf_streams *s, int seq_type, obuf_t *out)   /* process.c:75-96 */
{
    static const char *names[4] = { "id", "comment", NULL, "quality" };
    static const char *types[4] = { "DNA", "RNA", "protein", "text" };
    char line[128];
    for (int k = 0; k < 4; k++) {
        uint64_t total = 0;
        for (int c = 0; c < 257; c++) total += s->unexpected[k][c];
        if (!total) continue;
        int m = snprintf(line, sizeof line, "input has %llu unexpected %s characters:\n", (unsigned long long)total, k == 2 ? types[seq_type] : names[k]);
        obuf_put(out, line, (size_t)m);
        for (int c = 0; c < 256; c++) {
            if (!s->unexpected[k][c]) continue;
            if (c >= 32 && c < 127) m = snprintf(line, sizeof line, "    '%c': %llu\n", c, (unsigned long long)s->unexpected[k][c]);
            else m = snprintf(line, sizeof line, "    '\\x%02X': %llu\n", c, (unsigned long long)s->unexpected[k][c]);
            obuf_put(out, line, (size_t)m);
        }
    }
}

/* ------------------------------------------------------------------ container writer */

static void put_section(obuf_t *naf, uint64_t orig, const obuf_t *stream, int window_log)   /* compressor.c:150 */
{
    obuf_t z; obuf_init(&z);
    ozstd_compress_raw(stream->data, stream->size, window_log, &z);
    onaf_put_vle(naf, orig);
    onaf_put_vle(naf, z.size - 4);
    obuf_put(naf, z.data + 4, z.size - 4);                              /* magic stripped */
    obuf_free(&z);
}

int onaf_encode(const uint8_t *text, size_t n, const onaf_enc_opts *o, obuf_t *naf, obuf_t *report, char *err)
{
    onaf_streams s; onaf_streams_init(&s);
    if (onaf_split(text, n, o, &s, err)) { onaf_streams_free(&s); return -1; }
    static const uint8_t magic[3] = { 0x01, 0xF9, 0xEC };
    obuf_put(naf, magic, 3);                                            /* ennaf.c:538 */
    if (o->seq_type == ONAF_DNA) obuf_putc(naf, 1); else { obuf_putc(naf, 2); obuf_putc(naf, (uint8_t)o->seq_type); }
    int has_title = o->title != NULL;
    obuf_putc(naf, (uint8_t)((has_title << 6) | (1 << 5) | (1 << 4) | (1 << 3) | (s.store_mask << 2) | (1 << 1) | s.store_qual));
    obuf_putc(naf, ' ');
    onaf_put_vle(naf, o->have_line_length ? o->line_length : s.longest_line);
    onaf_put_vle(naf, s.n_sequences);
    if (has_title) { size_t tl = strlen(o->title); onaf_put_vle(naf, tl); obuf_put(naf, o->title, tl); }
    put_section(naf, s.ids.size, &s.ids, 0);
    put_section(naf, s.comm.size, &s.comm, 0);
    put_section(naf, s.len.size, &s.len, 0);
    if (s.store_mask) put_section(naf, s.mask.size, &s.mask, 0);
    put_section(naf, s.seq_size, &s.seq, o->window_log);
    if (s.store_qual) put_section(naf, s.qual.size, &s.qual, 0);
    if (report && !o->well_formed) onaf_report_unexpected(&s, o->seq_type, report);   /* ennaf.c:594 */
    onaf_streams_free(&s);
    return 0;
}

/* ------------------------------------------------------------------ container reader + views */

typedef struct {
    int version, seq_type;
    int has_title, has_ids, has_names, has_lengths, has_mask, has_data, has_quality;
    uint8_t sep;
    uint64_t line_len, N;
    const uint8_t *title; uint64_t title_len;
    /* sections in file order: ids, names(comments), lengths, mask, data, quality */
    struct { uint64_t orig, comp; const uint8_t *p; int present; } sec[6];
} hdr_t;

static const char *TYPE_NAME[4] = { "DNA", "RNA", "protein", "text" };

static int read_header(const uint8_t *p, size_t n, hdr_t *h, int need_sections, char *err)   /* input.c:31 */
{
    memset(h, 0, sizeof(*h));
    if (n == 0) return perr(err, "empty input");
    if (n < 3) return perr(err, "incomplete or truncated input\n");
    if (p[0] != 0x01 || p[1] != 0xF9 || p[2] != 0xEC) return perr(err, "not a NAF format\n");
    size_t pos = 3;
    if (pos >= n) return perr(err, "incomplete or truncated input\n");
    h->version = p[pos++];
    if (h->version < 1 || h->version > 2) return perr(err, "unknown version (%d) of NAF format\n", h->version);
    if (h->version > 1) {
        if (pos >= n) return perr(err, "incomplete or truncated input\n");
        int t = p[pos++];
        if (t < 1 || t > 3) return perr(err, "unknown sequence type (%d) found in NAF file\n", t);
        h->seq_type = t;
    }
    if (pos + 2 > n) return perr(err, "incomplete or truncated input\n");
    int flags = p[pos++];
    h->has_title = (flags >> 6) & 1; h->has_ids = (flags >> 5) & 1; h->has_names = (flags >> 4) & 1;
    h->has_lengths = (flags >> 3) & 1; h->has_mask = (flags >> 2) & 1; h->has_data = (flags >> 1) & 1; h->has_quality = flags & 1;
    h->sep = p[pos++];
    if (h->sep < 0x20 || h->sep > 0x7E) return perr(err, "unsupported name separator character\n");
    if (!need_sections) return 0;
    if (onaf_get_vle(p, n, &pos, &h->line_len, err)) return -1;
    if (onaf_get_vle(p, n, &pos, &h->N, err)) return -1;
    if (h->has_title) {
        if (onaf_get_vle(p, n, &pos, &h->title_len, err)) return -1;
        if (h->title_len > n - pos) return perr(err, "incomplete or truncated input\n");
        h->title = p + pos; pos += h->title_len;
    }
    int present[6] = { h->has_ids, h->has_names, h->has_lengths, h->has_mask, h->has_data, h->has_quality };
    for (int k = 0; k < 6; k++) {
        if (!present[k]) continue;
        if (h->N == 0 && pos >= n) break;   /* tolerated: views that need nothing */
        if (onaf_get_vle(p, n, &pos, &h->sec[k].orig, err)) return -1;
        if (onaf_get_vle(p, n, &pos, &h->sec[k].comp, err)) return -1;
        if (h->sec[k].comp > n - pos) return perr(err, "incomplete or truncated input\n");
        h->sec[k].p = p + pos; h->sec[k].present = 1; pos += h->sec[k].comp;
    }
    return 0;
}

/* put_magic_number (utils.c:144) + ZSTD_decompress (input.c:155) or the one-frame streaming loop */
static int load_section(const hdr_t *h, int k, int one_frame, obuf_t *out, const char *what, char *err)
{
    obuf_t z; obuf_init(&z);
    static const uint8_t magic[4] = { 0x28, 0xB5, 0x2F, 0xFD };
    obuf_put(&z, magic, 4); obuf_put(&z, h->sec[k].p, h->sec[k].comp);
    char zerr[256]; int rc; size_t used = 0;
    if (one_frame) rc = ozstd_decompress_frame(z.data, z.size, out, &used, zerr);
    else rc = ozstd_decompress(z.data, z.size, out, zerr);
    obuf_free(&z);
    if (rc) return perr(err, "can't decompress %s\n", what);
    return 0;
}

static void put_name(obuf_t *out, const hdr_t *h, const char *id, const char *comment)   /* output.c:105 */
{
    if (h->has_ids) obuf_put(out, id, strlen(id));
    if (h->has_names && (!h->has_ids || comment[0] != 0)) {
        if (h->has_ids) obuf_putc(out, h->sep);
        obuf_put(out, comment, strlen(comment));
    }
}

static int split_strings(const obuf_t *b, uint64_t N, const char ***ptrs, const char *what, char *err)   /* input.c:158-170 */
{
    if (b->size == 0 || b->data[b->size - 1] != 0) return perr(err, "corrupted %s - not 0-terminated\n", what);
    const char **v = (const char **)malloc(sizeof(char *) * (N ? N : 1));
    const char *p = (const char *)b->data, *end = p + b->size;
    for (uint64_t i = 0; i < N; i++) {
        if (p >= end) { free(v); return perr(err, "corrupted %s - can't read %llu\n", what, (unsigned long long)i); }
        v[i] = p; p += strlen(p) + 1;
    }
    *ptrs = v;
    return 0;
}

static void outf(obuf_t *out, const char *fmt, ...)
{
    char line[512]; va_list ap; va_start(ap, fmt); int m = vsnprintf(line, sizeof line, fmt, ap); va_end(ap);
    obuf_put(out, line, (size_t)m);
}

/* Full sequence text (all records concatenated) after unpack + mask / uppercase. */
static int load_bases(const hdr_t *h, const onaf_dec_opts *o, int allow_mask, obuf_t *bases, char *err)
{
    obuf_t raw; obuf_init(&raw);
    if (load_section(h, 4, 1, &raw, "sequence", err)) { obuf_free(&raw); return -1; }
    uint64_t total = h->sec[4].orig;
    if (h->seq_type < ONAF_PROTEIN) {
        if (raw.size * 2 < total) total = raw.size * 2;
        onaf_unpack4(raw.data, total, h->seq_type == ONAF_RNA, bases);
        if (allow_mask > 0 && !o->no_mask && h->has_mask) {
            obuf_t m; obuf_init(&m);
            if (load_section(h, 3, 0, &m, "mask", err)) { obuf_free(&m); obuf_free(&raw); return -1; }
            onaf_mask_apply(bases->data, bases->size, m.data, m.size);
            obuf_free(&m);
        }
    } else {
        if (raw.size < total) total = raw.size;
        obuf_put(bases, raw.data, total);
        if (o->no_mask && allow_mask >= 0)                              /* output.c:500,663: !use_mask => toupper */
            for (size_t i = 0; i < bases->size; i++) bases->data[i] = (uint8_t)toupper(bases->data[i]);
    }
    obuf_free(&raw);
    return 0;
}

int onaf_decode(const uint8_t *naf, size_t n, const onaf_dec_opts *o, obuf_t *out, char *err)
{
    hdr_t h; err[0] = 0;
    int type = o->out_type;
    if (read_header(naf, n, &h, 0, err)) return -1;
    if (type == ONAF_OUT_DEFAULT) type = h.has_quality ? ONAF_OUT_FASTQ : ONAF_OUT_FASTA;   /* unnaf.c:372 */
    if (type == ONAF_OUT_4BIT && h.seq_type >= ONAF_PROTEIN)
        return perr(err, "input has no 4-bit encoded data, but %s sequences\n", TYPE_NAME[h.seq_type]);
    if (type == ONAF_OUT_FORMAT) {
        outf(out, "%s sequences%s in NAF format version %d\n", TYPE_NAME[h.seq_type], h.has_quality ? " with qualities" : "", h.version);
        return 0;
    }
    if (type == ONAF_OUT_PART_LIST) {                                   /* output.c:7 */
        const char *names[7] = { "Title", "IDs", "Names", "Lengths", "Mask", "Data", "Quality" };
        int present[7] = { h.has_title, h.has_ids, h.has_names, h.has_lengths, h.has_mask, h.has_data, h.has_quality };
        int printed = 0;
        for (int k = 0; k < 7; k++) if (present[k]) { outf(out, "%s%s", printed ? ", " : "", names[k]); printed++; }
        obuf_putc(out, '\n');
        return 0;
    }
    if (read_header(naf, n, &h, 1, err)) return -1;
    uint64_t W = o->have_line_length ? o->line_length : h.line_len;
    uint64_t N = h.N;
    if (type == ONAF_OUT_NUMBER) { outf(out, "%llu\n", (unsigned long long)N); return 0; }
    if (type == ONAF_OUT_SIZES) {                                       /* output.c:21 */
        const char *names[6] = { "IDs", "Names", "Lengths", "Mask", "Data", "Quality" };
        if (h.has_title) outf(out, "Title: %llu\n", (unsigned long long)h.title_len);
        for (int k = 0; k < 6; k++) if (h.sec[k].present)
            outf(out, "%s: %llu / %llu (%.3f%%)\n", names[k], (unsigned long long)h.sec[k].comp, (unsigned long long)h.sec[k].orig,
                 (double)h.sec[k].comp / (double)h.sec[k].orig * 100);
        return 0;
    }
    if (type == ONAF_OUT_TITLE) { if (h.has_title) obuf_put(out, h.title, h.title_len); obuf_putc(out, '\n'); return 0; }
    if (N == 0) return 0;                                               /* unnaf.c:409 */

    int rc = 0;
    obuf_t ids, comm, len, bases, qual, mask; const char **idv = NULL, **cmv = NULL;
    obuf_init(&ids); obuf_init(&comm); obuf_init(&len); obuf_init(&bases); obuf_init(&qual); obuf_init(&mask);
    const uint32_t *L = NULL; uint64_t nL = 0;

#define LOAD_IDS()   do { if (h.has_ids)   { if (load_section(&h, 0, 0, &ids, "ids", err) || split_strings(&ids, N, &idv, "ids", err)) { rc = -1; goto done; } } } while (0)
#define LOAD_NAMES() do { if (h.has_names) { if (load_section(&h, 1, 0, &comm, "names", err) || split_strings(&comm, N, &cmv, "names", err)) { rc = -1; goto done; } } } while (0)
#define LOAD_LEN()   do { if (load_section(&h, 2, 0, &len, "lengths", err)) { rc = -1; goto done; } L = (const uint32_t *)len.data; nL = len.size / 4; } while (0)

    switch (type) {
    case ONAF_OUT_IDS:
        LOAD_IDS();
        if (h.has_ids) for (uint64_t i = 0; i < N; i++) outf(out, "%s\n", idv[i]);
        break;
    case ONAF_OUT_NAMES:                                                /* output.c:143 */
        LOAD_IDS(); LOAD_NAMES();
        if (h.has_ids || h.has_names) for (uint64_t i = 0; i < N; i++) { put_name(out, &h, idv ? idv[i] : "", cmv ? cmv[i] : ""); obuf_putc(out, '\n'); }
        break;
    case ONAF_OUT_LENGTHS:                                              /* output.c:180 */
        if (!h.has_lengths) break;
        LOAD_LEN();
        for (uint64_t i = 0; i < nL; i++) {
            uint64_t v = 0;
            while (i < nL && L[i] == 0xFFFFFFFFu) { v += 0xFFFFFFFFull; i++; }
            if (i < nL) v += L[i];
            outf(out, "%llu\n", (unsigned long long)v);
        }
        break;
    case ONAF_OUT_TOTAL_LENGTH:
        if (h.has_lengths) outf(out, "%llu\n", (unsigned long long)h.sec[4].orig);
        break;
    case ONAF_OUT_MASK:                                                 /* output.c:222 */
        if (!h.has_mask) break;
        if (load_section(&h, 3, 0, &mask, "mask", err)) { rc = -1; break; }
        for (uint64_t i = 0; i < mask.size; i++) {
            uint64_t v = 0;
            while (i < mask.size && mask.data[i] == 255) { v += 255; i++; }
            if (i < mask.size) v += mask.data[i];
            outf(out, "%llu\n", (unsigned long long)v);
        }
        break;
    case ONAF_OUT_TOTAL_MASK_LENGTH: {
        uint64_t v = 0;
        if (h.has_mask) { if (load_section(&h, 3, 0, &mask, "mask", err)) { rc = -1; break; } for (uint64_t i = 0; i < mask.size; i++) v += mask.data[i]; }
        outf(out, "%llu\n", (unsigned long long)v);
        break;
    }
    case ONAF_OUT_4BIT:
        if (h.has_data && load_section(&h, 4, 1, out, "sequence", err)) rc = -1;
        break;
    case ONAF_OUT_SEQ:                                                  /* output.c:457 print_dna */
        if (h.has_data && load_bases(&h, o, 1, out, err)) rc = -1;
        break;
    case ONAF_OUT_CHARCOUNT: {                                          /* output.c:544 */
        if (!h.has_data) break;
        if (load_bases(&h, o, 1, &bases, err)) { rc = -1; break; }
        uint64_t cnt[256] = {0};
        for (size_t i = 0; i < bases.size; i++) cnt[bases.data[i]]++;
        for (int c = 0; c < 256; c++) if (cnt[c]) {
            if (c >= 33 && c < 127) outf(out, "%c\t%llu\n", c, (unsigned long long)cnt[c]);
            else outf(out, "\\x%02X\t%llu\n", c, (unsigned long long)cnt[c]);
        }
        break;
    }
    case ONAF_OUT_SEQUENCES: {                                          /* output-sequences.c:61 */
        if (!h.has_data) break;
        LOAD_LEN();
        if (load_bases(&h, o, 1, &bases, err)) { rc = -1; break; }
        if (bases.size == 0) break;                                     /* nothing is flushed when there are no bases */
        /* output-sequences.c:7 as it stands: a length unit at a time; data beyond the last unit is printed bare */
        uint64_t n = bases.size < h.sec[4].orig ? bases.size : h.sec[4].orig, rec_rem = nL ? L[0] : 0, idx = 0;
        size_t pos = 0;
        while (nL && n >= rec_rem) {
            if (rec_rem > 0) { obuf_put(out, bases.data + pos, rec_rem); pos += rec_rem; n -= rec_rem; }
            if (L[idx] != 0xFFFFFFFFu) obuf_putc(out, '\n');
            idx++;
            if (idx >= nL) break;
            rec_rem = L[idx];
        }
        if (n > 0) obuf_put(out, bases.data + pos, n);
        break;
    }
    case ONAF_OUT_FASTA: {                                              /* output.c:608 print_fasta + :369 print_dna_buffer_as_fasta */
        if (!h.has_data) break;
        LOAD_IDS(); LOAD_NAMES(); LOAD_LEN();
        if (load_bases(&h, o, 1, &bases, err)) { rc = -1; break; }
        /* The reference's state machine as it stands, over the whole sequence at once: a length unit at a time, the line
         * budget carried across continuation units, and -- output.c:420-427 -- whatever sequence data is left once the
         * length units are used up (ennaf's id-byte bug, SURVEY A.4 #7, makes such files) printed with the line budget the
         * last record left behind and no newline after it. */
        uint64_t n = bases.size < h.sec[4].orig ? bases.size : h.sec[4].orig;      /* total_seq_n_bp_remaining */
        uint64_t idx = 0, seq = 0, line_rem, rec_rem; size_t pos = 0;
#define NAME(i) do { obuf_putc(out, '>'); put_name(out, &h, idv ? idv[i] : "", cmv ? cmv[i] : ""); obuf_putc(out, '\n'); } while (0)
#define SPLIT(size_) do { uint64_t sz_ = (size_);                                                       \
            if (W == 0) { obuf_put(out, bases.data + pos, sz_); pos += sz_; }                             \
            else {                                                                                        \
                while (sz_ > line_rem) { obuf_put(out, bases.data + pos, line_rem); obuf_putc(out, '\n'); pos += line_rem; sz_ -= line_rem; line_rem = W; } \
                obuf_put(out, bases.data + pos, sz_); pos += sz_; line_rem -= sz_;                        \
            } } while (0)
        while (idx < nL && seq < N && L[idx] == 0) { NAME(seq); idx++; seq++; }
        if (seq >= N || idx >= nL) break;
        NAME(seq); line_rem = W; rec_rem = L[idx];
        while (n >= rec_rem) {
            if (rec_rem > 0) { SPLIT(rec_rem); n -= rec_rem; }
            if (L[idx] == 0xFFFFFFFFu) idx++;
            else {
                obuf_putc(out, '\n'); idx++; seq++;
                while (idx < nL && seq < N && L[idx] == 0) { NAME(seq); idx++; seq++; }   /* empty sequences: no empty lines */
                if (seq < N) { NAME(seq); line_rem = W; }
            }
            if (idx >= nL) break;
            rec_rem = L[idx];
        }
        if (n > 0) SPLIT(n);
#undef NAME
#undef SPLIT
        break;
    }
    case ONAF_OUT_FASTQ: {                                              /* output-fastq.c:100; mask never applied (unnaf.c:442) */
        if (!h.has_quality) { rc = perr(err, "FASTQ output requested, but input has no qualities\n"); break; }
        if (!h.has_data) break;
        LOAD_IDS(); LOAD_NAMES(); LOAD_LEN();
        if (load_bases(&h, o, -1, &bases, err)) { rc = -1; break; }
        if (load_section(&h, 5, 1, &qual, "quality", err)) { rc = -1; break; }
        size_t pos = 0; uint64_t k = 0;
        for (uint64_t i = 0; i < N; i++) {
            obuf_putc(out, '@'); put_name(out, &h, idv ? idv[i] : "", cmv ? cmv[i] : ""); obuf_putc(out, '\n');
            uint64_t reclen = 0;
            while (k < nL && L[k] == 0xFFFFFFFFu) { reclen += L[k]; k++; }
            if (k < nL) { reclen += L[k]; k++; }
            uint64_t a = reclen > bases.size - pos ? bases.size - pos : reclen;
            uint64_t b = reclen > qual.size - pos ? (pos < qual.size ? qual.size - pos : 0) : reclen;
            obuf_put(out, bases.data + pos, a); obuf_put(out, "\n+\n", 3);
            obuf_put(out, qual.data + pos, b); obuf_putc(out, '\n');
            pos += reclen;
        }
        break;
    }
    default:
        rc = perr(err, "unknown output requested\n");
    }
done:
    free(idv); free(cmv);
    obuf_free(&ids); obuf_free(&comm); obuf_free(&len); obuf_free(&bases); obuf_free(&qual); obuf_free(&mask);
    return rc;
}
