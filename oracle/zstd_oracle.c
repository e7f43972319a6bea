/*
 * zstd_oracle.c — CPU ORACLE (test infrastructure only; see oracle.h).
 *
 * A sequential zstd decoder written from the format specification
 * (/root/reference/zstd/doc/zstd_compression_format.md, zstd v1.5.0 as vendored by the
 * reference; section line numbers cited below), plus a raw-block encoder.  It restates what
 * ZSTD_decompress (zstd/lib/decompress/zstd_decompress.c:1030) and the ZSTD_decompressStream
 * loop (…:1867) yield for the streams inside a .naf file.  Pinned against the reference's
 * libzstd (oracle/_ref/libzstd.so) on frames from levels -5..22, --long, and decodecorpus
 * (tests/test_oracle.py::test_oracle_zstd_vs_golden_frames over tests/golden/zstd, made by tools/make_golden.py).
 */
#include "oracle.h"

#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ obuf */

void obuf_init(obuf_t *b) { b->data = NULL; b->size = 0; b->cap = 0; }
void obuf_free(obuf_t *b) { free(b->data); b->data = NULL; b->size = b->cap = 0; }
void obuf_reserve(obuf_t *b, size_t extra)
{
    if (b->size + extra <= b->cap) return;
    size_t nc = b->cap ? b->cap : 256;
    while (nc < b->size + extra) nc *= 2;
    b->data = (uint8_t *)realloc(b->data, nc);
    if (!b->data) { fprintf(stderr, "oracle: out of memory\n"); abort(); }
    b->cap = nc;
}
void obuf_put(obuf_t *b, const void *p, size_t n)
{
    if (!n) return;
    obuf_reserve(b, n);
    memcpy(b->data + b->size, p, n);
    b->size += n;
}
void obuf_putc(obuf_t *b, uint8_t c) { obuf_reserve(b, 1); b->data[b->size++] = c; }

/* ------------------------------------------------------------------ errors */

typedef struct { char *msg; int failed; } zerr_t;

static int zfail(zerr_t *e, const char *fmt, ...)
{
    if (!e->failed && e->msg) {
        va_list ap; va_start(ap, fmt); vsnprintf(e->msg, 256, fmt, ap); va_end(ap);
    }
    e->failed = 1;
    return -1;
}

static inline int highbit(uint64_t v) { return 63 - __builtin_clzll(v); }

/* ------------------------------------------------------------------ bit readers */

/* Forward, LSB-first reader (FSE table descriptions; spec :1036-1110). */
typedef struct { const uint8_t *p; size_t n; size_t bitpos; } fbits_t;

static uint32_t fb_peek(const fbits_t *b, int nb)
{
    uint64_t v = 0;
    size_t byte = b->bitpos >> 3;
    for (int i = 0; i < 5; i++) if (byte + i < b->n) v |= (uint64_t)b->p[byte + i] << (8 * i);
    v >>= (b->bitpos & 7);
    return (uint32_t)(v & ((1ull << nb) - 1));
}

/* Backward reader (Huffman and FSE bitstreams; spec :1202-1225, :800-830).
 * `bits` = number of not-yet-consumed bits; reading below zero yields zero bits. */
typedef struct { const uint8_t *p; int64_t bits; } bbits_t;

static int bb_init(bbits_t *b, const uint8_t *p, size_t n, zerr_t *e)
{
    if (n == 0) return zfail(e, "empty backward bitstream");
    if (p[n - 1] == 0) return zfail(e, "backward bitstream: last byte is 0");
    b->p = p;
    b->bits = (int64_t)n * 8 - (8 - highbit(p[n - 1]));
    return 0;
}

/* Peek nb (<=32) bits located just below the current position, as the high bits first. */
static uint32_t bb_peek(const bbits_t *b, int nb)
{
    if (nb == 0) return 0;
    int64_t lo = b->bits - nb;              /* absolute bit index of the lowest wanted bit */
    uint64_t v = 0;
    int64_t start = lo < 0 ? 0 : lo;
    int64_t byte = start >> 3;
    int64_t end_byte = (b->bits + 7) >> 3;  /* exclusive */
    int k = 0;
    for (int64_t i = byte; i < end_byte && k < 8; i++, k++) v |= (uint64_t)b->p[i] << (8 * k);
    v >>= (start & 7);
    if (lo < 0) v <<= (-lo);                /* missing low bits read as zero */
    return (uint32_t)(v & ((1ull << nb) - 1));
}
static uint32_t bb_read(bbits_t *b, int nb) { uint32_t v = bb_peek(b, nb); b->bits -= nb; return v; }

/* ------------------------------------------------------------------ FSE */

#define FSE_MAX_LOG 9
#define FSE_MAX_SYM 256

typedef struct {
    int      log;
    uint8_t  sym[1 << FSE_MAX_LOG];
    uint8_t  nbits[1 << FSE_MAX_LOG];
    uint16_t base[1 << FSE_MAX_LOG];
} fse_t;

/* spec :1112-1190 "From normalized distribution to decoding tables". */
static int fse_build(fse_t *t, const int16_t *norm, int nsym, int log, zerr_t *e)
{
    if (log > FSE_MAX_LOG) return zfail(e, "FSE accuracy log %d too large", log);
    int size = 1 << log, high = size - 1;
    uint16_t next[FSE_MAX_SYM];
    t->log = log;
    for (int s = 0; s < nsym; s++) {
        if (norm[s] == -1) { t->sym[high--] = (uint8_t)s; next[s] = 1; }
        else next[s] = (uint16_t)norm[s];
    }
    int step = (size >> 1) + (size >> 3) + 3, mask = size - 1, pos = 0;
    for (int s = 0; s < nsym; s++) {
        for (int i = 0; i < norm[s]; i++) {
            t->sym[pos] = (uint8_t)s;
            do { pos = (pos + step) & mask; } while (pos > high);
        }
    }
    if (pos != 0) return zfail(e, "FSE spread did not return to 0");
    for (int i = 0; i < size; i++) {
        uint16_t n = next[t->sym[i]]++;
        int nb = log - highbit(n);
        t->nbits[i] = (uint8_t)nb;
        t->base[i] = (uint16_t)((n << nb) - size);
    }
    return 0;
}

/* spec :1036-1110 "FSE Table Description".  Returns bytes consumed or -1. */
static long fse_read_ncount(const uint8_t *p, size_t n, int16_t *norm, int *nsym_out, int max_sym,
                            int max_log, int *log_out, zerr_t *e)
{
    if (n < 1) return zfail(e, "FSE header truncated");
    fbits_t b = { p, n, 0 };
    int log = (int)fb_peek(&b, 4) + 5; b.bitpos += 4;
    if (log > max_log) return zfail(e, "FSE accuracy log %d > max %d", log, max_log);
    int remaining = 1 << log, sym = 0;
    while (remaining > 0 && sym <= max_sym) {
        int bits = highbit((uint64_t)remaining + 1) + 1;
        uint32_t val = fb_peek(&b, bits);
        uint32_t lower_mask = (1u << (bits - 1)) - 1;
        uint32_t threshold = (1u << bits) - 1 - (uint32_t)(remaining + 1);
        if ((val & lower_mask) < threshold) { b.bitpos += bits - 1; val &= lower_mask; }
        else { b.bitpos += bits; if (val > lower_mask) val -= threshold; }
        int proba = (int)val - 1;
        remaining -= proba < 0 ? 1 : proba;
        norm[sym++] = (int16_t)proba;
        if (proba == 0) {
            uint32_t rep;
            do {
                rep = fb_peek(&b, 2); b.bitpos += 2;
                for (uint32_t i = 0; i < rep && sym <= max_sym; i++) norm[sym++] = 0;
            } while (rep == 3);
        }
        if ((b.bitpos >> 3) > n + 4) return zfail(e, "FSE header overruns input");
    }
    if (remaining != 0) return zfail(e, "FSE distribution does not sum to table size");
    if (sym > max_sym + 1) return zfail(e, "FSE: too many symbols");
    size_t used = (b.bitpos + 7) >> 3;
    if (used > n) return zfail(e, "FSE header truncated");
    *nsym_out = sym; *log_out = log;
    return (long)used;
}

/* ------------------------------------------------------------------ Huffman */

#define HUF_MAX_BITS 11

typedef struct {
    int     max_bits;
    uint8_t sym[1 << HUF_MAX_BITS];
    uint8_t nbits[1 << HUF_MAX_BITS];
} huf_t;

/* spec :1300-1400: weights -> prefix codes -> flat decode table. */
static int huf_build(huf_t *h, uint8_t *weights, int nw, zerr_t *e)
{
    uint32_t total = 0;
    for (int i = 0; i < nw; i++) {
        if (weights[i] > HUF_MAX_BITS) return zfail(e, "Huffman weight > 11");
        if (weights[i]) total += 1u << (weights[i] - 1);
    }
    if (total == 0) return zfail(e, "Huffman: all weights zero");
    int max_bits = highbit(total) + 1;
    if (max_bits > HUF_MAX_BITS) return zfail(e, "Huffman max bits %d > 11", max_bits);
    uint32_t left = (1u << max_bits) - total;
    if (left & (left - 1)) return zfail(e, "Huffman: implied last weight not a power of 2");
    weights[nw] = (uint8_t)(highbit(left) + 1);
    nw++;
    /* bits = max_bits + 1 - weight; longer codes get the numerically lower table slots,
     * symbols of equal length in ascending order (spec :1345-1375). */
    uint32_t rank_count[HUF_MAX_BITS + 2] = {0}, rank_idx[HUF_MAX_BITS + 2] = {0};
    for (int i = 0; i < nw; i++) if (weights[i]) rank_count[max_bits + 1 - weights[i]]++;
    rank_idx[max_bits] = 0;
    for (int b = max_bits; b >= 1; b--) rank_idx[b - 1] = rank_idx[b] + rank_count[b] * (1u << (max_bits - b));
    if (rank_idx[0] != (1u << max_bits)) return zfail(e, "Huffman table not full");
    for (int i = 0; i < nw; i++) {
        if (!weights[i]) continue;
        int bits = max_bits + 1 - weights[i];
        uint32_t len = 1u << (max_bits - bits);
        for (uint32_t k = 0; k < len; k++) { h->sym[rank_idx[bits] + k] = (uint8_t)i; h->nbits[rank_idx[bits] + k] = (uint8_t)bits; }
        rank_idx[bits] += len;
    }
    h->max_bits = max_bits;
    return 0;
}

/* spec :1230-1300 "Huffman Tree Description".  Returns bytes consumed or -1. */
static long huf_read_tree(huf_t *h, const uint8_t *p, size_t n, int *fse_weights, zerr_t *e)
{
    if (n < 1) return zfail(e, "Huffman tree truncated");
    uint8_t weights[257];
    int nw = 0;
    int hb = p[0];
    size_t used;
    if (hb >= 128) {
        nw = hb - 127;
        size_t bytes = (size_t)(nw + 1) / 2;
        if (1 + bytes > n) return zfail(e, "Huffman direct weights truncated");
        for (int i = 0; i < nw; i++) weights[i] = (i & 1) ? (p[1 + i / 2] & 15) : (p[1 + i / 2] >> 4);
        used = 1 + bytes;
        *fse_weights = 0;
    } else {
        if (hb == 0) return zfail(e, "Huffman FSE weights: zero size");
        if ((size_t)1 + hb > n) return zfail(e, "Huffman FSE weights truncated");
        int16_t norm[16]; int nsym, log;
        long hdr = fse_read_ncount(p + 1, hb, norm, &nsym, 12, 6, &log, e);
        if (hdr < 0) return -1;
        fse_t t;
        if (fse_build(&t, norm, nsym, log, e)) return -1;
        bbits_t b;
        if (bb_init(&b, p + 1 + hdr, (size_t)hb - hdr, e)) return -1;
        /* two interleaved states; spec :1275-1295 */
        uint32_t s1 = bb_read(&b, log), s2 = bb_read(&b, log);
        if (b.bits < 0) return zfail(e, "Huffman FSE weights: bitstream too short");
        for (;;) {
            if (nw >= 255) return zfail(e, "Huffman: too many weights");
            weights[nw++] = t.sym[s1];
            s1 = t.base[s1] + bb_read(&b, t.nbits[s1]);
            if (b.bits < 0) { weights[nw++] = t.sym[s2]; break; }
            if (nw >= 255) return zfail(e, "Huffman: too many weights");
            weights[nw++] = t.sym[s2];
            s2 = t.base[s2] + bb_read(&b, t.nbits[s2]);
            if (b.bits < 0) { weights[nw++] = t.sym[s1]; break; }
        }
        used = 1 + (size_t)hb;
        *fse_weights = 1;
    }
    if (nw > 255) return zfail(e, "Huffman: too many weights");
    if (huf_build(h, weights, nw, e)) return -1;
    return (long)used;
}

static int huf_decode_stream(const huf_t *h, const uint8_t *p, size_t n, uint8_t *out, size_t nout, zerr_t *e)
{
    bbits_t b;
    if (bb_init(&b, p, n, e)) return -1;
    for (size_t i = 0; i < nout; i++) {
        uint32_t idx = bb_peek(&b, h->max_bits);
        out[i] = h->sym[idx];
        b.bits -= h->nbits[idx];
    }
    if (b.bits != 0) return zfail(e, "Huffman stream not exactly consumed (%lld bits left)", (long long)b.bits);
    return 0;
}

/* ------------------------------------------------------------------ frame state */

static const int16_t LL_DEFAULT[36] = { 4,3,2,2,2,2,2,2,2,2,2,2,2,1,1,1,2,2,2,2,2,2,2,2,2,3,2,1,1,1,1,1,-1,-1,-1,-1 };
static const int16_t ML_DEFAULT[53] = { 1,4,3,2,2,2,2,2,2,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,
                                        1,1,1,1,1,1,1,1,1,1,1,1,1,1,-1,-1,-1,-1,-1,-1,-1 };
static const int16_t OF_DEFAULT[29] = { 1,1,1,1,1,1,2,2,2,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,-1,-1,-1,-1,-1 };

static const uint32_t LL_BASE[36] = { 0,1,2,3,4,5,6,7,8,9,10,11,12,13,14,15,16,18,20,22,24,28,32,40,48,64,128,256,512,
                                      1024,2048,4096,8192,16384,32768,65536 };
static const uint8_t  LL_BITS[36] = { 0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,1,1,1,1,2,2,3,3,4,6,7,8,9,10,11,12,13,14,15,16 };
static const uint32_t ML_BASE[53] = { 3,4,5,6,7,8,9,10,11,12,13,14,15,16,17,18,19,20,21,22,23,24,25,26,27,28,29,30,31,32,
                                      33,34,35,37,39,41,43,47,51,59,67,83,99,131,259,515,1027,2051,4099,8195,16387,32771,65539 };
static const uint8_t  ML_BITS[53] = { 0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,
                                      1,1,1,1,2,2,3,3,4,4,5,7,8,9,10,11,12,13,14,15,16 };

typedef struct {
    huf_t  huf;  int have_huf;
    fse_t  ll, of, ml; int have_ll, have_of, have_ml;
    uint64_t rep[3];
    size_t frame_start;      /* index in out where this frame's content starts */
    ozstd_census_t *census;
} frame_t;

/* spec :433-560 "Literals Section" */
static long decode_literals(frame_t *f, const uint8_t *p, size_t n, uint8_t **lit_out, size_t *nlit, zerr_t *e)
{
    if (n < 1) return zfail(e, "literals section truncated");
    int type = p[0] & 3, sf = (p[0] >> 2) & 3;
    size_t regen, comp = 0, hdr;
    int streams = 1;
    if (type == 0 || type == 1) {
        if ((sf & 1) == 0) { regen = p[0] >> 3; hdr = 1; }
        else if (sf == 1) { if (n < 2) return zfail(e, "literals header truncated"); regen = (p[0] >> 4) | ((size_t)p[1] << 4); hdr = 2; }
        else { if (n < 3) return zfail(e, "literals header truncated"); regen = (p[0] >> 4) | ((size_t)p[1] << 4) | ((size_t)p[2] << 12); hdr = 3; }
    } else {
        if (sf == 0 || sf == 1) {
            if (n < 3) return zfail(e, "literals header truncated");
            uint32_t v = p[0] | (p[1] << 8) | (p[2] << 16);
            regen = (v >> 4) & 0x3FF; comp = (v >> 14) & 0x3FF; hdr = 3; streams = sf == 0 ? 1 : 4;
        } else if (sf == 2) {
            if (n < 4) return zfail(e, "literals header truncated");
            uint32_t v = p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24);
            regen = (v >> 4) & 0x3FFF; comp = (v >> 18) & 0x3FFF; hdr = 4; streams = 4;
        } else {
            if (n < 5) return zfail(e, "literals header truncated");
            uint64_t v = p[0] | (p[1] << 8) | (p[2] << 16) | ((uint64_t)p[3] << 24) | ((uint64_t)p[4] << 32);
            regen = (v >> 4) & 0x3FFFF; comp = (v >> 22) & 0x3FFFF; hdr = 5; streams = 4;
        }
    }
    if (regen > 128 * 1024) return zfail(e, "literals regenerated size too large");
    uint8_t *lit = (uint8_t *)malloc(regen + 1);
    *lit_out = lit; *nlit = regen;
    ozstd_census_t *c = f->census;
    if (type == 0) {
        if (hdr + regen > n) return zfail(e, "raw literals truncated");
        memcpy(lit, p + hdr, regen);
        if (c) c->lit_raw++;
        return (long)(hdr + regen);
    }
    if (type == 1) {
        if (hdr + 1 > n) return zfail(e, "RLE literals truncated");
        memset(lit, p[hdr], regen);
        if (c) c->lit_rle++;
        return (long)(hdr + 1);
    }
    if (hdr + comp > n) return zfail(e, "compressed literals truncated");
    const uint8_t *q = p + hdr; size_t qn = comp;
    if (type == 2) {
        int fsew = 0;
        long used = huf_read_tree(&f->huf, q, qn, &fsew, e);
        if (used < 0) return -1;
        f->have_huf = 1; q += used; qn -= (size_t)used;
        if (c) { c->lit_huf++; if (fsew) c->huf_fse_weights++; else c->huf_direct_weights++; }
    } else {
        if (!f->have_huf) return zfail(e, "treeless literals without a previous Huffman table");
        if (c) c->lit_treeless++;
    }
    if (streams == 1) {
        if (huf_decode_stream(&f->huf, q, qn, lit, regen, e)) return -1;
    } else {
        if (c) c->lit_4streams++;
        if (qn < 6) return zfail(e, "jump table truncated");
        size_t s1 = q[0] | (q[1] << 8), s2 = q[2] | (q[3] << 8), s3 = q[4] | (q[5] << 8);
        if (6 + s1 + s2 + s3 > qn) return zfail(e, "jump table exceeds literals size");
        size_t s4 = qn - 6 - s1 - s2 - s3;
        size_t seg = (regen + 3) / 4;
        if (seg * 3 > regen) return zfail(e, "4-stream literals too short");
        const uint8_t *d = q + 6;
        if (huf_decode_stream(&f->huf, d, s1, lit, seg, e)) return -1;
        if (huf_decode_stream(&f->huf, d + s1, s2, lit + seg, seg, e)) return -1;
        if (huf_decode_stream(&f->huf, d + s1 + s2, s3, lit + 2 * seg, seg, e)) return -1;
        if (huf_decode_stream(&f->huf, d + s1 + s2 + s3, s4, lit + 3 * seg, regen - 3 * seg, e)) return -1;
    }
    return (long)(hdr + comp);
}

/* spec :700-760: one of the three symbol-coding tables of a sequences section */
static long read_seq_table(int mode, fse_t *t, int *have, const uint8_t *p, size_t n, const int16_t *defnorm,
                           int defn, int deflog, int max_sym, int max_log, ozstd_census_t *c, zerr_t *e)
{
    switch (mode) {
    case 0:
        if (fse_build(t, defnorm, defn, deflog, e)) return -1;
        *have = 1; if (c) c->seq_predef++;
        return 0;
    case 1:
        if (n < 1) return zfail(e, "RLE sequence table truncated");
        if (p[0] > max_sym) return zfail(e, "RLE sequence symbol out of range");
        t->log = 0; t->sym[0] = p[0]; t->nbits[0] = 0; t->base[0] = 0;
        *have = 1; if (c) c->seq_rle++;
        return 1;
    case 2: {
        int16_t norm[64]; int nsym, log;
        long used = fse_read_ncount(p, n, norm, &nsym, max_sym, max_log, &log, e);
        if (used < 0) return -1;
        if (fse_build(t, norm, nsym, log, e)) return -1;
        *have = 1; if (c) c->seq_fse++;
        return used;
    }
    default:
        if (!*have) return zfail(e, "Repeat_Mode without a previous table");
        if (c) c->seq_repeat++;
        return 0;
    }
}

/* spec :621-960 "Sequences Section", "Sequence Execution", "Repeat Offsets" */
static int decode_sequences(frame_t *f, const uint8_t *p, size_t n, const uint8_t *lit, size_t nlit, obuf_t *out, zerr_t *e)
{
    if (n < 1) return zfail(e, "sequences section truncated");
    size_t nseq, pos;
    if (p[0] == 0) { nseq = 0; pos = 1; }
    else if (p[0] < 128) { nseq = p[0]; pos = 1; }
    else if (p[0] < 255) { if (n < 2) return zfail(e, "nbSeq truncated"); nseq = ((size_t)(p[0] - 128) << 8) + p[1]; pos = 2; }
    else { if (n < 3) return zfail(e, "nbSeq truncated"); nseq = (size_t)p[1] + ((size_t)p[2] << 8) + 0x7F00; pos = 3; }
    if (nseq == 0) {
        if (pos != n) return zfail(e, "extra bytes after empty sequences section");
        obuf_put(out, lit, nlit);
        return 0;
    }
    if (f->census) f->census->n_sequences += nseq;
    if (pos >= n) return zfail(e, "sequence modes byte missing");
    int modes = p[pos++];
    if (modes & 3) return zfail(e, "reserved bits set in sequence modes");
    long u;
    u = read_seq_table((modes >> 6) & 3, &f->ll, &f->have_ll, p + pos, n - pos, LL_DEFAULT, 36, 6, 35, 9, f->census, e);
    if (u < 0) return -1;
    pos += (size_t)u;
    u = read_seq_table((modes >> 4) & 3, &f->of, &f->have_of, p + pos, n - pos, OF_DEFAULT, 29, 5, 31, 8, f->census, e);
    if (u < 0) return -1;
    pos += (size_t)u;
    u = read_seq_table((modes >> 2) & 3, &f->ml, &f->have_ml, p + pos, n - pos, ML_DEFAULT, 53, 6, 52, 9, f->census, e);
    if (u < 0) return -1;
    pos += (size_t)u;
    if (pos >= n) return zfail(e, "sequence bitstream missing");

    bbits_t b = { NULL, 0 };
    if (bb_init(&b, p + pos, n - pos, e)) return -1;
    uint32_t sl = bb_read(&b, f->ll.log), so = bb_read(&b, f->of.log), sm = bb_read(&b, f->ml.log);
    size_t lp = 0;
    for (size_t i = 0; i < nseq; i++) {
        int oc = f->of.sym[so], mc = f->ml.sym[sm], lc = f->ll.sym[sl];
        if (oc > 31) return zfail(e, "offset code > 31");
        if (mc > 52 || lc > 35) return zfail(e, "length code out of range");
        uint64_t ofv = ((uint64_t)1 << oc) + bb_read(&b, oc);
        uint32_t mlen = ML_BASE[mc] + bb_read(&b, ML_BITS[mc]);
        uint32_t llen = LL_BASE[lc] + bb_read(&b, LL_BITS[lc]);
        if (i + 1 < nseq) {
            sl = f->ll.base[sl] + bb_read(&b, f->ll.nbits[sl]);
            sm = f->ml.base[sm] + bb_read(&b, f->ml.nbits[sm]);
            so = f->of.base[so] + bb_read(&b, f->of.nbits[so]);
        }
        if (b.bits < 0) return zfail(e, "sequence bitstream overrun");
        /* repeat offsets, spec :896-960 */
        uint64_t off;
        if (ofv > 3) { off = ofv - 3; f->rep[2] = f->rep[1]; f->rep[1] = f->rep[0]; f->rep[0] = off; }
        else {
            uint32_t idx = (uint32_t)ofv - 1 + (llen == 0);
            if (idx == 0) off = f->rep[0];
            else {
                off = idx == 3 ? f->rep[0] - 1 : f->rep[idx];
                if (off == 0) return zfail(e, "repeat offset resolves to 0");
                if (idx != 1) f->rep[2] = f->rep[1];
                f->rep[1] = f->rep[0]; f->rep[0] = off;
            }
        }
        if (lp + llen > nlit) return zfail(e, "sequence literal length exceeds literals");
        obuf_put(out, lit + lp, llen); lp += llen;
        if (off > out->size - f->frame_start) return zfail(e, "match offset %llu beyond frame start", (unsigned long long)off);
        obuf_reserve(out, mlen);
        uint8_t *d = out->data + out->size; const uint8_t *s = d - off;
        for (uint32_t k = 0; k < mlen; k++) d[k] = s[k];
        out->size += mlen;
    }
    if (b.bits != 0) return zfail(e, "sequence bitstream not exactly consumed");
    obuf_put(out, lit + lp, nlit - lp);
    return 0;
}

/* spec :141-330 frame header, :333-430 blocks */
static int decode_frame(const uint8_t *src, size_t n, obuf_t *out, size_t *consumed, ozstd_census_t *census, zerr_t *e)
{
    size_t pos = 4;
    if (n < 6) return zfail(e, "frame truncated");
    int fhd = src[pos++];
    int fcs_flag = fhd >> 6, single = (fhd >> 5) & 1, checksum = (fhd >> 2) & 1, did = fhd & 3;
    if (fhd & 8) return zfail(e, "reserved bit set in frame header descriptor");
    uint64_t window = 0;
    if (!single) {
        int wd = src[pos++];
        int wl = 10 + (wd >> 3);
        window = (1ull << wl) + ((1ull << wl) / 8) * (wd & 7);
        if (census) census->window_log = (uint32_t)wl;
    }
    static const int did_bytes[4] = { 0, 1, 2, 4 };
    if (did) return zfail(e, "dictionary frames are not supported (NAF never uses them)");
    pos += did_bytes[did];
    int fcs_bytes = fcs_flag == 0 ? single : (fcs_flag == 1 ? 2 : fcs_flag == 2 ? 4 : 8);
    if (pos + fcs_bytes > n) return zfail(e, "frame header truncated");
    uint64_t fcs = 0;
    for (int i = 0; i < fcs_bytes; i++) fcs |= (uint64_t)src[pos + i] << (8 * i);
    if (fcs_bytes == 2) fcs += 256;
    pos += fcs_bytes;
    if (single) window = fcs;
    uint64_t block_max = window < 128 * 1024 ? window : 128 * 1024;

    frame_t *f = (frame_t *)calloc(1, sizeof(frame_t));
    f->rep[0] = 1; f->rep[1] = 4; f->rep[2] = 8;
    f->frame_start = out->size;
    f->census = census;
    int rc = 0;
    for (;;) {
        if (pos + 3 > n) { rc = zfail(e, "block header truncated"); break; }
        uint32_t bh = src[pos] | (src[pos + 1] << 8) | ((uint32_t)src[pos + 2] << 16);
        pos += 3;
        int last = bh & 1, type = (bh >> 1) & 3; size_t bsize = bh >> 3;
        if (census) census->n_blocks++;
        if (type == 0) {
            if (pos + bsize > n) { rc = zfail(e, "raw block truncated"); break; }
            obuf_put(out, src + pos, bsize); pos += bsize;
            if (census) census->raw_blocks++;
        } else if (type == 1) {
            if (pos + 1 > n) { rc = zfail(e, "RLE block truncated"); break; }
            obuf_reserve(out, bsize); memset(out->data + out->size, src[pos], bsize); out->size += bsize; pos += 1;
            if (census) census->rle_blocks++;
        } else if (type == 2) {
            if (pos + bsize > n) { rc = zfail(e, "compressed block truncated"); break; }
            if (bsize > 128 * 1024) { rc = zfail(e, "compressed block too large"); break; }
            (void)block_max;
            if (census) census->compressed_blocks++;
            uint8_t *lit = NULL; size_t nlit = 0;
            long used = decode_literals(f, src + pos, bsize, &lit, &nlit, e);
            if (used >= 0) rc = decode_sequences(f, src + pos + used, bsize - (size_t)used, lit, nlit, out, e);
            else rc = -1;
            free(lit);
            if (rc) break;
            pos += bsize;
        } else { rc = zfail(e, "reserved block type"); break; }
        if (last) break;
    }
    if (!rc && checksum) { if (pos + 4 > n) rc = zfail(e, "checksum truncated"); else pos += 4; }
    if (!rc && fcs_bytes && out->size - f->frame_start != fcs) rc = zfail(e, "frame content size mismatch");
    free(f);
    *consumed = pos;
    return rc;
}

static int decode_any(const uint8_t *src, size_t n, obuf_t *out, int one_frame, size_t *consumed_out,
                      ozstd_census_t *census, char *err)
{
    zerr_t e = { err, 0 };
    if (err) err[0] = 0;
    size_t pos = 0; int frames = 0;
    while (pos < n) {
        if (n - pos < 4) return zfail(&e, "trailing garbage after frame");
        uint32_t magic = src[pos] | (src[pos + 1] << 8) | (src[pos + 2] << 16) | ((uint32_t)src[pos + 3] << 24);
        if ((magic & 0xFFFFFFF0u) == 0x184D2A50u) {
            if (n - pos < 8) return zfail(&e, "skippable frame truncated");
            uint32_t sz = src[pos + 4] | (src[pos + 5] << 8) | (src[pos + 6] << 16) | ((uint32_t)src[pos + 7] << 24);
            if (n - pos - 8 < sz) return zfail(&e, "skippable frame truncated");
            pos += 8 + (size_t)sz;
            continue;
        }
        if (magic != 0xFD2FB528u) return zfail(&e, "bad zstd magic %08x", magic);
        size_t used = 0;
        if (decode_frame(src + pos, n - pos, out, &used, census, &e)) return -1;
        pos += used; frames++;
        if (one_frame) break;
    }
    if (consumed_out) *consumed_out = pos;
    if (one_frame && frames == 0) return zfail(&e, "no frame found");
    return 0;
}

int ozstd_decompress(const uint8_t *src, size_t n, obuf_t *out, char *err)
{
    return decode_any(src, n, out, 0, NULL, NULL, err);
}

int ozstd_decompress_frame(const uint8_t *src, size_t n, obuf_t *out, size_t *consumed, char *err)
{
    return decode_any(src, n, out, 1, consumed, NULL, err);
}

int ozstd_census(const uint8_t *src, size_t n, ozstd_census_t *c, char *err)
{
    obuf_t tmp; obuf_init(&tmp);
    memset(c, 0, sizeof(*c));
    int rc = decode_any(src, n, &tmp, 0, NULL, c, err);
    obuf_free(&tmp);
    return rc;
}

/* ------------------------------------------------------------------ raw-block encoder */

void ozstd_compress_raw(const uint8_t *src, size_t n, int window_log, obuf_t *out)
{
    static const uint8_t magic[4] = { 0x28, 0xB5, 0x2F, 0xFD };
    if (window_log < 10) window_log = 19;       /* what ennaf -1 declares */
    obuf_put(out, magic, 4);
    obuf_putc(out, 0x00);                        /* FHD: no FCS, no checksum, no dict */
    obuf_putc(out, (uint8_t)((window_log - 10) << 3));
    size_t pos = 0;
    do {
        size_t bs = n - pos > 128 * 1024 ? 128 * 1024 : n - pos;
        int last = pos + bs == n;
        uint32_t bh = (uint32_t)last | (0u << 1) | ((uint32_t)bs << 3);
        obuf_putc(out, bh & 0xFF); obuf_putc(out, (bh >> 8) & 0xFF); obuf_putc(out, (bh >> 16) & 0xFF);
        obuf_put(out, src + pos, bs);
        pos += bs;
    } while (pos < n);
}
