/*
 * nafgpu.h — C ABI of libnafgpu.so, the B200 (sm_100a) implementation of the NAF encode/decode
 * hot path.  Plain pointers and sizes only; no C++ or torch types cross this boundary.
 *
 * The reference (KirillKryukov/naf v1.3.0) has no plugin/FFI interface: ennaf and unnaf are two
 * single-translation-unit C programs.  The seams this library replaces are the in-process ones the
 * reference already has (all paths relative to /root/reference):
 *
 *   encode   process()                       ennaf/src/process.c:586   text  -> name/comment/seq/qual chunks
 *            seq_writer_masked_4bit          ennaf/src/process.c:24    chunk -> extract_mask + encode_dna
 *            compress()/compressor_end_stream ennaf/src/compressor.c:120,64   stream bytes -> zstd frame
 *            container writer                ennaf/src/ennaf.c:538-589 header + VLE + magic-stripped frames
 *   decode   read_header / load_*            unnaf/src/input.c:31,145-246     container -> streams
 *            ZSTD_decompress / ZSTD_decompressStream loops  unnaf/src/input.c:155, output.c:640-650
 *            write_4bit_as_fasta + print_dna_buffer_as_fasta unnaf/src/output.c:445,369
 *            print_fastq                     unnaf/src/output-fastq.c:100
 *
 * nafgpu_encode() / nafgpu_decode() are what ennaf's main() (ennaf.c:433) and unnaf's main()
 * (unnaf.c:356) call once the command line is parsed and the input is mapped; INTEGRATION.md
 * shows the binding.  Stage entry points below exist for parity tests and profiling.
 *
 * Conventions: every call returns 0 on success or a negative NAFGPU_E_* code; the message that the
 * reference would have passed to die() (without its "ennaf error: " / "unnaf error: " prefix) is
 * available from nafgpu_last_error().  A context owns one CUDA stream, a device arena and pinned
 * host staging; it is not thread-safe — use one context per host thread / GPU.  Output pointers
 * returned by a call stay valid until the next call on the same context.
 * There is NO CPU fallback: every entry point fails with NAFGPU_E_CUDA if no sm_100 device is usable.
 */
#ifndef NAFGPU_H
#define NAFGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NAFGPU_VERSION "0.1.0"

enum {
    NAFGPU_OK          =  0,
    NAFGPU_E_CUDA      = -1,   /* CUDA runtime / no device */
    NAFGPU_E_INPUT     = -2,   /* the reference would die() on this input; see nafgpu_last_error */
    NAFGPU_E_FORMAT    = -3,   /* corrupt .naf container or zstd stream */
    NAFGPU_E_UNSUPPORTED = -4, /* valid, but outside what this build handles (stated in the message) */
    NAFGPU_E_ARG       = -5
};

/* sequence types: ennaf --dna/--rna/--protein/--text (ennaf.c:54), NAF v2 type byte (input.c:44) */
enum { NAFGPU_DNA = 0, NAFGPU_RNA = 1, NAFGPU_PROTEIN = 2, NAFGPU_TEXT = 3 };
enum { NAFGPU_FMT_AUTO = 0, NAFGPU_FMT_FASTA = 1, NAFGPU_FMT_FASTQ = 2 };

/* output views of unnaf (unnaf.c:16-23).  Views that only read the header or tiny sections are
 * produced on the host by the CLI; the ones listed here run on the GPU. */
enum {
    NAFGPU_OUT_DEFAULT = 0,     /* FASTQ if the file has qualities, else FASTA (unnaf.c:372) */
    NAFGPU_OUT_FASTA   = 1,     /* print_fasta            output.c:608 */
    NAFGPU_OUT_FASTQ   = 2,     /* print_fastq            output-fastq.c:100 */
    NAFGPU_OUT_SEQ     = 3,     /* print_dna (--seq)      output.c:457 */
    NAFGPU_OUT_SEQUENCES = 4,   /* print_sequences        output-sequences.c:61 */
    NAFGPU_OUT_4BIT    = 5,     /* print_4bit             output.c:266 */
    NAFGPU_OUT_IDS     = 6,     /* the decompressed id stream, '\0' -> '\n'      output.c:95 */
    NAFGPU_OUT_NAMES   = 7,     /* id[ sep comment]\n per record                 output.c:143 */
    NAFGPU_OUT_LENGTHS = 8,     /* raw u32 LE length units (CLI formats them)    output.c:180 */
    NAFGPU_OUT_MASK    = 9,     /* raw u8 mask units (CLI formats them)          output.c:222 */
    NAFGPU_OUT_CHARCOUNT = 10   /* 256 x u64 LE counts (CLI formats them)        output.c:544 */
};

typedef struct nafgpu_ctx nafgpu_ctx;

typedef struct {
    int32_t  seq_type;          /* NAFGPU_DNA.. */
    int32_t  input_format;      /* NAFGPU_FMT_AUTO: detect like confirm_input_format (process.c:547) */
    int32_t  no_mask;           /* ennaf --no-mask */
    int32_t  strict;            /* ennaf --strict: fail on the first unexpected character */
    int32_t  well_formed;       /* ennaf --well-formed */
    int32_t  have_line_length;  /* ennaf --line-length N */
    uint64_t line_length;
    int32_t  level;             /* ennaf -# (default 1): <= 1 is the fastest parse (every stream entropy-coded); >= 2 adds LZ77
                                   matches + FSE-coded sequences to the ids / comments / lengths streams (file size of
                                   `ennaf -1`, slower).  Sequence, quality and mask are Huffman-coded at every level. */
    int32_t  window_log;        /* ennaf --long N: declared window of the SEQ frame (0 = default) */
    const char *title;          /* ennaf --title, NULL = none */
    int32_t  general_parser;    /* 1: skip the canonical-input fast parser and use the general (process.c-exact FSM) one;
                                   results are identical either way -- for tests and profiling */
    int32_t  no_block_index;    /* 1: do not append the block index -- a zstd skippable frame behind the lengths frame that lists the
                                   compressed size of every block of the sequence / quality frames.  The reference reads sections with
                                   ZSTD_decompress (unnaf/src/input.c:211), which skips it; our decoder uses it to find all blocks
                                   without walking the chain of block headers (multi-GPU decode starts everywhere at once). */
} nafgpu_enc_opts;

typedef struct {
    int32_t  out_type;          /* NAFGPU_OUT_* */
    int32_t  no_mask;           /* unnaf --no-mask */
    int32_t  have_line_length;  /* unnaf --line-length N */
    uint64_t line_length;
    /* record-range decode (one rank of a multi-GPU decode): with n_records != 0 the call returns only the text of
     * records [first_record, first_record + n_records) -- exactly the bytes they occupy in the full output, so the ranks'
     * pieces concatenate to it.  Applies to the per-record views (FASTA, FASTQ, --sequences, --ids, --names). */
    uint64_t first_record, n_records;
} nafgpu_dec_opts;

/* Facts ennaf prints or needs after encoding (ennaf.c:556,594-596). */
typedef struct {
    uint64_t n_sequences;
    uint64_t longest_line;
    uint64_t n_bases;                 /* seq_size_original */
    int32_t  format;                  /* NAFGPU_FMT_FASTA/FASTQ, 0 for empty input */
    int32_t  reserved;
    uint64_t stream_raw[6];           /* ids, comments, lengths, mask, sequence, quality: uncompressed bytes */
    uint64_t stream_comp[6];          /* compressed bytes as stored (frame minus 4-byte magic) */
    uint64_t unexpected[4][257];      /* id, comment, sequence, quality (process.c:75 report) */
} nafgpu_enc_info;

/* Timing of the last call, CUDA events on the context's stream (milliseconds). */
typedef struct {
    float h2d_ms, kernels_ms, d2h_ms, total_ms;
    uint32_t kernel_launches;         /* launches of this library's own kernels in the last call */
    uint32_t parser_fallback;         /* encode / split: 1 if the canonical-input parser declined the input and the general
                                         (process.c-exact) parser redid the split; 0 otherwise */
} nafgpu_timing;

/* ---- context ---- */
int  nafgpu_create(int device, nafgpu_ctx **ctx);     /* -1 = current device */
void nafgpu_destroy(nafgpu_ctx *ctx);
const char *nafgpu_last_error(const nafgpu_ctx *ctx); /* ctx may be NULL: error of a failed nafgpu_create */
const char *nafgpu_version(void);
int  nafgpu_get_timing(const nafgpu_ctx *ctx, nafgpu_timing *t);
void *nafgpu_stream(nafgpu_ctx *ctx);                 /* the cudaStream_t all work is launched on */

/* Per-kernel timing with CUDA events on ctx's stream (for roofline reports; adds an event pair per
 * launch, so leave it off in timed runs).  The report of the last call is "name\tlaunches\tms\n" lines. */
int  nafgpu_profile(nafgpu_ctx *ctx, int enable);
const char *nafgpu_profile_report(const nafgpu_ctx *ctx);

/* Pinned host memory for zero-staging transfers (optional; pageable pointers are accepted too). */
int  nafgpu_host_alloc(size_t n, void **p);
void nafgpu_host_free(void *p);

/* ---- the hot path, host buffers (what ennaf / unnaf call) ---- */

/* FASTA/FASTQ text -> .naf bytes.  Replaces process() + compress() + the container writer.
 * *naf points into pinned memory owned by ctx. */
int nafgpu_encode(nafgpu_ctx *ctx, const uint8_t *text, size_t n, const nafgpu_enc_opts *opts,
                  const uint8_t **naf, size_t *naf_size, nafgpu_enc_info *info);

/* .naf bytes -> text of the requested view.  Replaces load_*(), the ZSTD_decompress* loops and
 * print_fasta()/print_fastq()/print_dna()/print_sequences()/print_4bit(). */
int nafgpu_decode(nafgpu_ctx *ctx, const uint8_t *naf, size_t n, const nafgpu_dec_opts *opts,
                  const uint8_t **text, size_t *text_size);

/* ---- the hot path, streamed (what the command-line tools call: pipes, files read and written in pieces) ----
 * The reference reads its input through a 16 KB buffer (ennaf/src/process.c:227-240) and writes its output through 128 KB
 * ones (unnaf/src/output.c:640-650); it never holds a file in memory.  These calls keep that shape at the boundary -- the
 * caller hands over / receives the text in pieces of a few tens of MB, in page-locked buffers the library rotates, so that
 * reading the next piece (or writing the previous one) overlaps the PCIe copy of the current one -- while the device works on
 * the whole text at once.
 *
 *   nafgpu_encode_begin    start a text; size_hint = its size if known (a regular file), else 0
 *   nafgpu_encode_buffer   a page-locked buffer to put the next piece of text into (*cap bytes at most)
 *   nafgpu_encode_feed     the buffer now holds n bytes: they go up while the caller fills the other buffer
 *   nafgpu_encode_end      no more text: transform + compress, result as from nafgpu_encode
 *   nafgpu_encode_end_to   the same, with the .naf delivered in order, piece by piece, to `write` (no file-sized host buffer)
 *
 *   nafgpu_decode_to       nafgpu_decode whose text is delivered in order, piece by piece, to `write` (which returns 0 to go
 *                          on; anything else stops the call with NAFGPU_E_ARG); *text_size = total bytes delivered */
int nafgpu_encode_begin(nafgpu_ctx *ctx, const nafgpu_enc_opts *opts, size_t size_hint);
int nafgpu_encode_buffer(nafgpu_ctx *ctx, void **buf, size_t *cap);
int nafgpu_encode_feed(nafgpu_ctx *ctx, size_t n);
int nafgpu_encode_end(nafgpu_ctx *ctx, const uint8_t **naf, size_t *naf_size, nafgpu_enc_info *info);
typedef int (*nafgpu_write_fn)(void *user, const uint8_t *piece, size_t n);
int nafgpu_encode_end_to(nafgpu_ctx *ctx, nafgpu_write_fn write, void *user, size_t *naf_size, nafgpu_enc_info *info);
int nafgpu_decode_to(nafgpu_ctx *ctx, const uint8_t *naf, size_t n, const nafgpu_dec_opts *opts,
                     nafgpu_write_fn write, void *user, size_t *text_size);

/* ---- the hot path, device-resident (bench `value`, pipelines that keep data in HBM) ----
 * d_text / d_naf are device pointers on ctx's device; outputs are device pointers into ctx's arena.
 * `host_copy` (may be NULL) is a host mirror of the compressed input used only to walk zstd block
 * headers without a device round trip; if NULL the headers are fetched from the device. */
int nafgpu_encode_device(nafgpu_ctx *ctx, const uint8_t *d_text, size_t n, const nafgpu_enc_opts *opts,
                         const uint8_t **d_naf, size_t *naf_size, nafgpu_enc_info *info);
int nafgpu_decode_device(nafgpu_ctx *ctx, const uint8_t *d_naf, size_t n, const uint8_t *host_copy,
                         const nafgpu_dec_opts *opts, const uint8_t **d_text, size_t *text_size);

/* ---- stage entry points (parity tests, ncu) — host buffers in, pinned ctx-owned buffers out ---- */

/* zstd frame(s) -> bytes.  Replaces ZSTD_decompress (input.c:155; multi-frame) when one_frame = 0 and
 * the single-frame ZSTD_decompressStream loop (output.c:640) when one_frame = 1.  `src` starts with
 * the 4-byte zstd magic.  expected_size = regenerated size if known, else 0 (then it is derived). */
int nafgpu_zstd_decompress(nafgpu_ctx *ctx, const uint8_t *src, size_t n, size_t expected_size, int one_frame,
                           const uint8_t **out, size_t *out_size);

/* bytes -> one zstd frame (with magic).  Replaces ZSTD_initCStream/compressStream/endStream
 * (compressor.c:7-20,120,64). */
int nafgpu_zstd_compress(nafgpu_ctx *ctx, const uint8_t *src, size_t n, int window_log,
                         const uint8_t **out, size_t *out_size);
/* The same with ennaf's -# (ZSTD_initCStream's level, compressor.c:17): level >= 2 parses the bytes with LZ77 matches and
 * FSE-coded sequences in independent 8 KB blocks (what nafgpu_encode does for ids / comments / lengths at those
 * levels: csrc/zstd_lzc_hd.cuh); level <= 1 is nafgpu_zstd_compress. */
int nafgpu_zstd_compress_level(nafgpu_ctx *ctx, const uint8_t *src, size_t n, int window_log, int level,
                               const uint8_t **out, size_t *out_size);

/* text -> the six uncompressed streams (no zstd).  streams[k]/sizes[k] in container order
 * ids, comments, lengths, mask, sequence, quality.  Replaces process() + encoders.c. */
int nafgpu_split(nafgpu_ctx *ctx, const uint8_t *text, size_t n, const nafgpu_enc_opts *opts,
                 const uint8_t *streams[6], size_t sizes[6], nafgpu_enc_info *info);

/* ---- one .naf from several shards: the multi-GPU form of nafgpu_encode ----
 * Every rank encodes a record-aligned piece of the text; one small all-gather of nafgpu_shard_counts in between
 * lets each rank finish on its own (4-bit nibble parity, mask runs that cross shard boundaries, Last_Block), and
 * rank 0 concatenates the zstd *blocks* of all ranks into ONE frame per stream -- the reference unnaf stops after
 * the first frame of the sequence / quality streams (zstd_decompress.c:2129-2140, output.c:640).
 *   nafgpu_shard_begin    parse + split + pack this shard (text: host pointer, or device pointer if text_on_device)
 *   (caller)              all-gather the counts, derive this shard's link record, see naf_b200/sharded.py link_for
 *   nafgpu_shard_finish   nibble shift, boundary mask runs, zstd blocks; raw[k] / body[k] = uncompressed / compressed
 *                         bytes of stream k as this shard contributes them (body = blocks only: no frame header)
 *   nafgpu_shard_fetch    copy the blocks of stream k to dst (host or device memory, cudaMemcpyDefault)
 *   (caller)              rank 0 writes header + per stream VLE sizes, frame header, all ranks' blocks in rank order */
typedef struct {
    uint64_t n_records, n_bases, longest_line;
    uint64_t n_flips;           /* case changes strictly inside the shard's concatenated sequence */
    uint64_t last_flip;         /* position (in the shard's bases) of the last of them, if n_flips > 0 */
    uint8_t  first_code;        /* 4-bit code of the shard's first base */
    uint8_t  first_case, last_case;   /* 1 = masked (byte >= 96, encoders.c:134) */
    uint8_t  format;            /* NAFGPU_FMT_FASTA / FASTQ, 0 for an empty shard */
    uint8_t  pad[4];
} nafgpu_shard_counts;

typedef struct {
    uint64_t bases_before;      /* bases in the shards before this one */
    uint64_t run_carry;         /* bases since the last case change before this shard (= bases_before if there is none) */
    uint8_t  prev_last_case;    /* case of the last base before this shard (0 if there is none) */
    uint8_t  next_first_code;   /* 4-bit code of the first base after this shard (0 if there is none) */
    uint8_t  is_last;           /* last shard: sets Last_Block and writes the trailing mask run */
    uint8_t  store_qual;        /* some shard is FASTQ: the file has a quality stream, and every shard (an empty one too) contributes to it */
    uint8_t  pad[4];
} nafgpu_shard_link;

/* Where to cut a text into `pieces` record-aligned shards: cuts[0] = 0, cuts[pieces] = n, cuts[k] = offset of the first record
 * start at or after k * n / pieces (FASTA: a '>' at a line start; FASTQ: after every 4th line counted from the top of the
 * text -- exact, a quality line may begin with '@').  Found on the GPU: newline ordinals by a prefix sum over per-tile counts
 * (the record-boundary scan of process.c:358,477 restated data-parallel).  text: host pointer, or device pointer if text_on_device. */
int nafgpu_record_cuts(nafgpu_ctx *ctx, const uint8_t *text, size_t n, int text_on_device, int pieces, uint64_t *cuts /* pieces + 1 */);

int nafgpu_shard_begin(nafgpu_ctx *ctx, const uint8_t *text, size_t n, int text_on_device, const nafgpu_enc_opts *opts,
                       nafgpu_shard_counts *counts, nafgpu_enc_info *info);
int nafgpu_shard_finish(nafgpu_ctx *ctx, const nafgpu_shard_link *link, uint64_t raw[6], uint64_t body[6]);
int nafgpu_shard_fetch(nafgpu_ctx *ctx, int stream, void *dst);

#ifdef __cplusplus
}
#endif
#endif
