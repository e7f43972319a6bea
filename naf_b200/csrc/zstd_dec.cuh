// zstd_dec.cuh — block-parallel zstd decoder, orchestration + kernel bodies.
//
// Replaces, for the streams inside a .naf file, what unnaf does through ZSTD_decompress
// (unnaf/src/input.c:155 -> zstd/lib/decompress/zstd_decompress.c:1030) and its
// ZSTD_decompressStream loops (unnaf/src/output.c:640-650, input.c:352-440).
//
// libzstd decodes one block after another on one core.  Here ALL blocks of ALL streams of a file
// are decoded together, one GPU thread per independent item:
//
//   host   walk frame + block headers (3 bytes per block; the only inherently serial chain)
//   K1     one thread / block          literals-section + sequences-section headers
//   S1     chunked scan over blocks    who provides my Huffman / FSE tables (treeless, Repeat_Mode),
//                                      slots in the table pools, offsets in the literal scratch
//   K2     one thread / Huffman table  weights (direct or FSE-coded) -> 2^maxbits-entry table in HBM
//   K3     one thread / block w/ seqs  FSE table descriptions -> tables
//   K4     one thread / block w/ seqs  FSE-decode (ll, ml, offset_value); symbolic repeat-offset map
//   S2     chunked scan over blocks    output offsets; repeat-offset history entering each block
//   K5     one thread / Huffman stream (4 per block) decode literals straight to their final place
//          when the block has no sequences, else to a literal scratch
//   K6     one CTA / block             raw / RLE blocks and raw / RLE literals
//   K7     one thread / block w/ seqs  resolve offsets, per-sequence positions
//   K8     one thread / sequence       place literals, write one back-pointer per match byte
//   K9..   pointer jumping             out[p] = out[link[p]] until every match byte is resolved
//
// All kernel bodies are HD lambdas run through an executor; the library instantiates CudaExec only.
#pragma once
#include "zstd_hd.cuh"
#include <string>
#include <thread>
#include <utility>
#include <vector>
#include <string.h>

namespace nafz {

// ------------------------------------------------------------------ per-block record
struct ZBlock {
    // --- host walk ---
    u64 src;            // offset of the block content in the input buffer
    u64 out_base;       // first_in_stream: arena offset where this stream's output starts
    u32 csize;          // content bytes (raw: size, RLE: 1, compressed: Block_Size)
    u32 rsize;          // raw / RLE: regenerated size
    u32 frame_first_blk;// index of the first block of my frame
    u8  type;           // 0 raw, 1 RLE, 2 compressed
    u8  first_in_frame;
    u8  first_in_stream;
    u8  stream;
    u8  skip;           // output not wanted (record-range decode): K5 / K6 leave the block alone
    u8  local;          // K7b executed this block's sequences itself (every match starts inside the block): K8 / K9 leave it alone
    // --- K1 ---
    u32 lit_regen, lit_csize;
    u8  lit_type, lit_streams, lit_hdr, modes;
    u32 nseq;
    u32 seq_off;        // offset (in block) of the byte following the sequences header
    // --- S1 ---
    i32 huf_src, ll_src, of_src, ml_src;
    u32 huf_slot, fse_slot;
    u64 lit_off;        // offset in the literal scratch (blocks with sequences)
    u64 seq_base;       // index of my first sequence
    u64 seq_cum;        // sequences in blocks [0, me]
    // --- K2/K3 ---
    u32 tree_len;       // bytes of Huffman tree description
    u8  huf_bits, ll_log, of_log, ml_log;
    u32 bits_off;       // offset (in block) of the sequence bitstream
    // --- K4 ---
    u32 match_total;
    RepFn repfn;
    // --- S2 ---
    u64 out_off;        // arena offset of my output
    u64 frame_out;      // arena offset where my frame's output starts (offset validity)
    u32 rep_in[3];
};

struct ZSeq { u32 ll, ml, of, lit_rel, dst_rel, blk; };

struct ZStreamDesc {
    u64 src_off, src_len;       // compressed bytes (starting with the zstd magic) in the input buffer
    u64 out_off, out_size;      // region of the output arena; out_size = expected regenerated size
    int one_frame;              // 1: stop after the first frame like the reference's streaming loops
    int no_magic;               // 1: the first frame's 4-byte magic was stripped (.naf sections, compressor.c:158)
    u64 need_lo = 0, need_hi = ~0ull;   // only bytes [need_lo, need_hi) of the regenerated stream are wanted: a stream WITHOUT
                                // sequences skips the literal decode of every block outside the range
};

struct ZStreamResult { u64 out_size; u64 nseq; u64 consumed; };

struct ZNeedTab { u64 lo[8], hi[8]; u32 on[8]; };

static const int HUF_SLOT_ENTRIES = 2048;     // u16 each
static const int FSE_SLOT_ENTRIES = 1280;     // u32 each: LL 512 | OF 256 | ML 512
static const int FSE_OF_AT = 512, FSE_ML_AT = 768;

// ------------------------------------------------------------------ host: frame / block walk
// spec "Frame_Header" / "Block_Header"; replaces zstd_decompress.c:819 ZSTD_decompressFrame's header handling.
// `regen` / `simple` (optional): the regenerated size of every block, as far as the host can tell without decoding, and whether
// that was possible for all of them -- a SIMPLE stream is one frame whose blocks are all self-contained: raw, RLE, or compressed
// with their own Huffman table (or raw / RLE literals) and no sequences.  That is what our encoder writes; such a stream can
// be cut at any block boundary and the pieces decoded on their own, at known output offsets (piecewise decode behind a
// chunked upload, record-range decode).
inline int zstd_walk_stream(const u8 *h, const ZStreamDesc &sd, int stream_idx, std::vector<ZBlockHead> &blocks,
                            u64 *consumed, std::string &err, std::vector<u32> *regen = nullptr, bool *simple = nullptr,
                            std::vector<std::pair<u64, u64>> *skippable = nullptr)
{
    bool all_simple = true;
    const u8 *p = h + sd.src_off; const u64 n = sd.src_len;
    u64 pos = 0; int frames = 0; bool first_in_stream = true;
    bool skip_magic = sd.no_magic != 0;
    while (pos < n) {
        if (!skip_magic) {
            if (n - pos < 4) { err = "trailing bytes after zstd frame"; return -1; }
            u32 magic = p[pos] | (p[pos + 1] << 8) | (p[pos + 2] << 16) | ((u32)p[pos + 3] << 24);
            if ((magic & 0xFFFFFFF0u) == 0x184D2A50u) {
                if (n - pos < 8) { err = "skippable frame truncated"; return -1; }
                u64 sz = p[pos + 4] | (p[pos + 5] << 8) | (p[pos + 6] << 16) | ((u64)p[pos + 7] << 24);
                if (n - pos - 8 < sz) { err = "skippable frame truncated"; return -1; }
                if (skippable) skippable->push_back(std::make_pair(sd.src_off + pos + 8, sz));     // (offset of its content in h, size)
                pos += 8 + sz; continue;
            }
            if (magic != 0xFD2FB528u) { err = "bad zstd magic"; return -1; }
            pos += 4;
        }
        skip_magic = false;
        if (n - pos < 2) { err = "zstd frame header truncated"; return -1; }
        u32 fhd = p[pos++];
        u32 fcs_flag = fhd >> 6, single = (fhd >> 5) & 1, checksum = (fhd >> 2) & 1, did = fhd & 3;
        if (fhd & 8) { err = "reserved bit set in zstd frame header"; return -1; }
        if (did) { err = "zstd dictionaries are not supported"; return -1; }
        if (!single) pos++;                                   // window descriptor: irrelevant, whole stream is resident
        u32 fcs_bytes = fcs_flag == 0 ? single : (fcs_flag == 1 ? 2 : fcs_flag == 2 ? 4 : 8);
        if (pos + fcs_bytes > n) { err = "zstd frame header truncated"; return -1; }
        pos += fcs_bytes;
        u32 frame_first = (u32)blocks.size(); bool first_in_frame = true;
        for (;;) {
            if (pos + 3 > n) { err = "zstd block header truncated"; return -1; }
            u32 bh = p[pos] | (p[pos + 1] << 8) | ((u32)p[pos + 2] << 16);
            pos += 3;
            u32 last = bh & 1, type = (bh >> 1) & 3, bsize = bh >> 3;
            if (type == 3) { err = "reserved zstd block type"; return -1; }
            ZBlockHead b;
            b.src = sd.src_off + pos; b.type = (u8)type; b.stream = (u8)stream_idx;
            b.first_in_frame = first_in_frame; b.first_in_stream = first_in_stream; b.frame_first_blk = frame_first;
            b.out_base = sd.out_off;
            if (type == 1) { b.csize = 1; b.rsize = bsize; }
            else { b.csize = bsize; b.rsize = type == 0 ? bsize : 0; }
            if (bsize > 128 * 1024 && type != 1) { err = "zstd block larger than 128 KB"; return -1; }
            if (pos + b.csize > n) { err = "zstd block truncated"; return -1; }
            if (regen) {
                u32 r = b.rsize;
                if (type == 2) {
                    LitHeader lh;
                    if (all_simple && lit_header_parse(p + pos, b.csize, lh) == Z_OK && lh.type != 3 && (u64)lh.hdr + lh.csize + 1 == b.csize) r = lh.regen;
                    else all_simple = false;
                }
                regen->push_back(r);
            }
            pos += b.csize;
            // The walk is a chain of dependent reads ~10 KB apart: every header is a cache miss in host memory.  Blocks of
            // one stream have nearly equal sizes, so the headers after the next one are where "same size again" puts
            // them, give or take a few lines: request those lines now and the misses overlap instead of queueing.
            {
                const u64 step = (u64)b.csize + 3;
                for (u64 k = 1; k <= 3; k++) {
                    const u64 guess = pos + k * step;
                    if (guess + 64 * k + 64 >= n) break;
                    for (u64 d = 0; d <= 2 * k; d++) __builtin_prefetch(p + guess - 64 * k + 64 * d, 0, 0);
                }
            }
            blocks.push_back(b);
            first_in_frame = false; first_in_stream = false;
            if (last) break;
        }
        if (checksum) { if (pos + 4 > n) { err = "zstd checksum truncated"; return -1; } pos += 4; }
        frames++;
        if (sd.one_frame) break;
    }
    if (sd.one_frame && frames == 0) { err = "no zstd frame"; return -1; }
    if (simple) *simple = all_simple && frames == 1 && regen != nullptr;
    *consumed = pos;
    return 0;
}

// ------------------------------------------------------------------ host: block list from a block index
// Our encoder appends a skippable frame to the lengths section (zstd_enc.cu: "block index") that lists the compressed size of
// every block of the sequence and quality streams; all their blocks but the last regenerate the same number of bytes.  With
// it the headers are no chain of dependent reads (3.5 ms for the 46 k blocks of a 3 Gbp sequence stream, and on the critical
// path of every rank of a multi-GPU decode) but independent ones at known places, prefetched ahead.  Every header is checked
// against the index; any disagreement -> false, and the caller walks the stream the usual way.
struct ZIndexEntry { u32 section, nblk, regen, reserved; u64 total; const u8 *csize; };    // csize: nblk little-endian u16

// read_headers = false: the file is not touched at all -- the blocks get type 0xFF ("see the header"), and the device, which
// reads every block anyway, checks the header against the index (zstd_decode_blocks, simple plans only).
inline bool zstd_walk_indexed(const u8 *h, const ZStreamDesc &sd, const ZIndexEntry &e, std::vector<ZBlockHead> &blocks, std::vector<u32> &regen,
                              u64 *consumed, bool read_headers = true)
{
    blocks.clear(); regen.clear();
    if (!sd.no_magic || sd.src_len < 2 || e.nblk == 0 || e.regen == 0 || e.regen > 128 * 1024) return false;
    const u8 *p = h + sd.src_off; const u64 n = sd.src_len;
    if (p[0] != 0x00) return false;                           // FHD of our frames: window byte follows, no content size / checksum / dictionary
    if ((u64)(e.nblk - 1) * e.regen >= e.total && e.total) return false;
    if (e.total > (u64)e.nblk * e.regen) return false;
    std::vector<u64> at(e.nblk + 1);
    u64 pos = 2;
    for (u32 i = 0; i < e.nblk; i++) { at[i] = pos; pos += 3 + (u64)(e.csize[2 * i] | (e.csize[2 * i + 1] << 8)); }
    at[e.nblk] = pos;
    if (pos != n) return false;
    blocks.reserve(e.nblk); regen.reserve(e.nblk);
    if (!read_headers) {
        for (u32 i = 0; i < e.nblk; i++) {
            const u32 want = i + 1 < e.nblk ? e.regen : (u32)(e.total - (u64)(e.nblk - 1) * e.regen);
            ZBlockHead b;
            b.src = sd.src_off + at[i] + 3; b.type = 0xFF; b.stream = 0; b.first_in_frame = b.first_in_stream = i == 0; b.frame_first_blk = 0;
            b.out_base = sd.out_off; b.csize = (u32)(at[i + 1] - at[i] - 3); b.rsize = want;
            blocks.push_back(b); regen.push_back(want);
        }
        *consumed = n;
        return true;
    }
    for (u32 i = 0; i < e.nblk; i++) {
        if (i + 24 < e.nblk) __builtin_prefetch(p + at[i + 24], 0, 0);
        const u64 q = at[i];
        const u32 cs = (u32)(at[i + 1] - q - 3);
        const u32 bh = p[q] | (p[q + 1] << 8) | ((u32)p[q + 2] << 16);
        const u32 last = bh & 1, type = (bh >> 1) & 3, bsize = bh >> 3;
        const u32 want = i + 1 < e.nblk ? e.regen : (u32)(e.total - (u64)(e.nblk - 1) * e.regen);
        if (type == 3 || last != (u32)(i + 1 == e.nblk)) return false;
        ZBlockHead b;
        b.src = sd.src_off + q + 3; b.type = (u8)type; b.stream = 0; b.first_in_frame = b.first_in_stream = i == 0; b.frame_first_blk = 0;
        b.out_base = sd.out_off;
        if (type == 1) { if (cs != 1 || bsize != want) return false; b.csize = 1; b.rsize = bsize; }
        else if (type == 0) { if (bsize != cs || bsize != want) return false; b.csize = bsize; b.rsize = bsize; }
        else {
            if (bsize != cs) return false;
            LitHeader lh;
            if (lit_header_parse(p + q + 3, cs, lh) != Z_OK || lh.type == 3 || (u64)lh.hdr + lh.csize + 1 != cs || lh.regen != want) return false;
            b.csize = bsize; b.rsize = 0;
        }
        blocks.push_back(b); regen.push_back(want);
    }
    *consumed = n;
    return true;
}

// ------------------------------------------------------------------ kernel bodies

struct ZDecArgs {
    const u8 *in;           // compressed input (device)
    u8 *out;                // output arena (device)
    ZBlock *blk; u32 nblk;
    u16 *huf_pool; u32 *fse_pool;
    const u32 *predef;      // predefined LL|OF|ML tables in FSE slot layout
    u8 *lit_scratch;
    ZSeq *seq;
    u32 *link; u32 *bitmap; u64 span_lo;     // pointer-jumping state over [span_lo, span_hi)
    u32 *status;            // [0] first error code, [1] block index of it, [2] "still unresolved" flag
    u32 nchunks;            // scan chunking
    struct ScanA *scan_a; struct ScanB *scan_b;
};

HD void zerr(const ZDecArgs &a, int code, u32 blk)
{
    if (a.status[0] == 0) { a.status[0] = (u32)code; a.status[1] = blk; }   // benign race: any first error will do
}

// K1 — literals + sequences section headers of one compressed block
HD void k_block_headers(const ZDecArgs &a, u32 i)
{
    ZBlock &b = a.blk[i];
    if (b.type != 2) return;
    const u8 *p = a.in + b.src;
    LitHeader lh;
    int rc = lit_header_parse(p, b.csize, lh);
    if (rc) { zerr(a, rc, i); b.type = 0; b.rsize = 0; b.csize = 0; return; }
    b.lit_type = (u8)lh.type; b.lit_streams = (u8)lh.streams; b.lit_hdr = (u8)lh.hdr;
    b.lit_regen = lh.regen; b.lit_csize = lh.csize;
    size_t so = (size_t)lh.hdr + lh.csize;
    u32 nseq = 0, modes = 0;
    size_t used = seq_header_parse(p + so, b.csize - so, &nseq, &modes);
    if (used == 0 || (modes & 3)) { zerr(a, Z_ERR_SEQ_HEADER, i); nseq = 0; modes = 0; used = b.csize - so; }
    if (nseq == 0 && so + used != b.csize) zerr(a, Z_ERR_SEQ_HEADER, i);
    b.nseq = nseq; b.modes = (u8)modes; b.seq_off = (u32)(so + used);
}

// S1 — table provenance + slot / offset prefix sums.  INHERIT = no definition seen, NONE = frame start seen.
static const i32 SRC_INHERIT = -3, SRC_NONE = -1;
struct ScanA { i32 huf, ll, of, ml; u32 huf_slots, fse_slots; u64 lit, seqs; };

HD ScanA scana_identity() { ScanA s; s.huf = s.ll = s.of = s.ml = SRC_INHERIT; s.huf_slots = s.fse_slots = 0; s.lit = s.seqs = 0; return s; }
HD i32 src_then(i32 a, i32 b) { return b == SRC_INHERIT ? a : b; }
HD ScanA scana_combine(const ScanA &a, const ScanA &b)
{
    ScanA r;
    r.huf = src_then(a.huf, b.huf); r.ll = src_then(a.ll, b.ll); r.of = src_then(a.of, b.of); r.ml = src_then(a.ml, b.ml);
    r.huf_slots = a.huf_slots + b.huf_slots; r.fse_slots = a.fse_slots + b.fse_slots; r.lit = a.lit + b.lit; r.seqs = a.seqs + b.seqs;
    return r;
}
HD bool blk_needs_fse_slot(const ZBlock &b) { return b.type == 2 && b.nseq > 0 && (((b.modes >> 6) & 3) == 1 || ((b.modes >> 6) & 3) == 2 || ((b.modes >> 4) & 3) == 1 || ((b.modes >> 4) & 3) == 2 || ((b.modes >> 2) & 3) == 1 || ((b.modes >> 2) & 3) == 2); }
// element contribution of block i, applied on top of running state `s` (inclusive)
HD void scana_push(ScanA &s, const ZBlock &b, u32 i)
{
    if (b.first_in_frame) { s.huf = s.ll = s.of = s.ml = SRC_NONE; }
    if (b.type != 2) return;
    if (b.lit_type == 2) { s.huf = (i32)i; s.huf_slots++; }
    if (b.nseq > 0) {
        if (((b.modes >> 6) & 3) != 3) s.ll = (i32)i;
        if (((b.modes >> 4) & 3) != 3) s.of = (i32)i;
        if (((b.modes >> 2) & 3) != 3) s.ml = (i32)i;
        if (blk_needs_fse_slot(b)) s.fse_slots++;
        s.lit += b.lit_regen; s.seqs += b.nseq;
    }
}
HD void chunk_range(u32 n, u32 nchunks, u32 c, u32 &lo, u32 &hi)
{
    u32 per = (n + nchunks - 1) / nchunks;
    lo = c * per; hi = lo + per; if (lo > n) lo = n; if (hi > n) hi = n;
}
HD void k_scan1_phase1(const ZDecArgs &a, u32 c)
{
    u32 lo, hi; chunk_range(a.nblk, a.nchunks, c, lo, hi);
    ScanA s = scana_identity();
    for (u32 i = lo; i < hi; i++) scana_push(s, a.blk[i], i);
    a.scan_a[c] = s;
}
HD void k_scan1_phase2(const ZDecArgs &a)
{
    ScanA run = scana_identity();
    for (u32 c = 0; c < a.nchunks; c++) { ScanA mine = a.scan_a[c]; a.scan_a[c] = run; run = scana_combine(run, mine); }
    a.scan_a[a.nchunks] = run;          // grand totals for the host
}
HD void k_scan1_phase3(const ZDecArgs &a, u32 c)
{
    u32 lo, hi; chunk_range(a.nblk, a.nchunks, c, lo, hi);
    ScanA s = a.scan_a[c];
    for (u32 i = lo; i < hi; i++) {
        ZBlock &b = a.blk[i];
        ScanA before = s;
        scana_push(s, b, i);
        b.seq_cum = s.seqs;
        if (b.type != 2) continue;
        b.huf_src = s.huf == SRC_INHERIT ? SRC_NONE : s.huf;
        b.huf_slot = before.huf_slots;
        if (b.nseq > 0) {
            b.ll_src = s.ll == SRC_INHERIT ? SRC_NONE : s.ll;
            b.of_src = s.of == SRC_INHERIT ? SRC_NONE : s.of;
            b.ml_src = s.ml == SRC_INHERIT ? SRC_NONE : s.ml;
            b.fse_slot = before.fse_slots; b.lit_off = before.lit; b.seq_base = before.seqs;
            if (b.ll_src < 0 || b.of_src < 0 || b.ml_src < 0) zerr(a, Z_ERR_NO_TABLE, i);
        }
        if (b.lit_type >= 2 && b.huf_src < 0) zerr(a, Z_ERR_NO_TABLE, i);
    }
}

// K2 — Huffman table of one block that carries a tree description
HD void k_huf_table(const ZDecArgs &a, u32 i)
{
    ZBlock &b = a.blk[i];
    if (b.type != 2 || b.lit_type != 2) return;
    u8 weights[260]; int nw = 0, max_bits = 0;
    size_t used = huf_read_weights(a.in + b.src + b.lit_hdr, b.lit_csize, weights, &nw, &max_bits);
    if (used == 0) { zerr(a, Z_ERR_HUF_TREE, i); b.tree_len = 0; b.huf_bits = 0; return; }
    b.tree_len = (u32)used; b.huf_bits = (u8)max_bits;
    if (!huf_build_table(a.huf_pool + (size_t)b.huf_slot * HUF_SLOT_ENTRIES, weights, nw, max_bits)) { zerr(a, Z_ERR_HUF_TREE, i); b.huf_bits = 0; }
}

// K3 — FSE tables (RLE_Mode / FSE_Compressed_Mode) of one block with sequences
HD void k_fse_tables(const ZDecArgs &a, u32 i)
{
    ZBlock &b = a.blk[i];
    if (b.type != 2 || b.nseq == 0) return;
    const u8 *p = a.in + b.src; size_t pos = b.seq_off, end = b.csize;
    u32 *slot = a.fse_pool + (size_t)b.fse_slot * FSE_SLOT_ENTRIES;
    const int max_sym[3] = { 35, 31, 52 }, max_log[3] = { 9, 8, 9 }, at[3] = { 0, FSE_OF_AT, FSE_ML_AT }, shift[3] = { 6, 4, 2 };
    u8 logs[3] = { 6, 5, 6 };
    for (int k = 0; k < 3; k++) {
        u32 mode = (b.modes >> shift[k]) & 3;
        if (mode == 1) {
            if (pos >= end || p[pos] > max_sym[k]) { zerr(a, Z_ERR_FSE_HEADER, i); b.nseq = 0; return; }
            slot[at[k]] = p[pos]; logs[k] = 0; pos++;
        } else if (mode == 2) {
            short norm[64]; u16 next[64]; int nsym, log;
            size_t used = fse_read_ncount(p + pos, end - pos, norm, &nsym, max_sym[k], max_log[k], &log);
            if (used == 0 || !fse_build_table(slot + at[k], norm, nsym, log, next)) { zerr(a, Z_ERR_FSE_HEADER, i); b.nseq = 0; return; }
            logs[k] = (u8)log; pos += used;
        }
    }
    b.ll_log = logs[0]; b.of_log = logs[1]; b.ml_log = logs[2];
    b.bits_off = (u32)pos;
    if (pos >= end) { zerr(a, Z_ERR_SEQ_STREAM, i); b.nseq = 0; }
}

HD const u32 *fse_table_for(const ZDecArgs &a, i32 src, int kind, int *log)
{
    const ZBlock &sb = a.blk[src];
    const int shift = kind == 0 ? 6 : (kind == 1 ? 4 : 2), at = kind == 0 ? 0 : (kind == 1 ? FSE_OF_AT : FSE_ML_AT);
    u32 mode = (sb.modes >> shift) & 3;
    if (mode == 0) { *log = kind == 1 ? 5 : 6; return a.predef + at; }
    *log = kind == 0 ? sb.ll_log : (kind == 1 ? sb.of_log : sb.ml_log);
    return a.fse_pool + (size_t)sb.fse_slot * FSE_SLOT_ENTRIES + at;
}

// K4 — decode the sequences of one block (spec "Sequences_Section": bitstream, state update order)
// Replaces zstd_decompress_block.c:937 ZSTD_decodeSequence, minus offset resolution (done in K7).
// `stl/sto/stm`: the block's three decode tables already copied next to the thread (shared memory on the GPU), or nullptr
HD void k_seq_decode(const ZDecArgs &a, u32 i, const u32 *stl = nullptr, const u32 *sto = nullptr, const u32 *stm = nullptr)
{
    ZBlock &b = a.blk[i];
    b.repfn = repfn_identity(); b.match_total = 0;
    if (b.type != 2 || b.nseq == 0) return;
    if (b.ll_src < 0 || b.of_src < 0 || b.ml_src < 0) { b.nseq = 0; return; }
    int ll_log, of_log, ml_log;
    const u32 *tl = fse_table_for(a, b.ll_src, 0, &ll_log), *to = fse_table_for(a, b.of_src, 1, &of_log), *tm = fse_table_for(a, b.ml_src, 2, &ml_log);
    if (stl) { tl = stl; to = sto; tm = stm; }
    BackBits bs;
    if (!bs.init(a.in + b.src + b.bits_off, b.csize - b.bits_off)) { zerr(a, Z_ERR_SEQ_STREAM, i); b.nseq = 0; return; }
    u32 sl = bs.read(ll_log), so = bs.read(of_log), sm = bs.read(ml_log);
    ZSeq *seq = a.seq + b.seq_base;
    RepFn rf = repfn_identity();
    u64 lit_total = 0, match_total = 0;
    for (u32 k = 0; k < b.nseq; k++) {
        u32 el = tl[sl], eo = to[so], em = tm[sm];
        u32 lc = fse_sym(el), oc = fse_sym(eo), mc = fse_sym(em);
        if (oc > 31 || lc > 35 || mc > 52) { zerr(a, Z_ERR_SEQ_STREAM, i); b.nseq = 0; return; }
        u32 ofv = (1u << oc) + bs.read((int)oc);
        u32 ml = ml_base_of(mc) + bs.read((int)ml_bits_of(mc));
        u32 ll = ll_base_of(lc) + bs.read((int)ll_bits_of(lc));
        if (k + 1 < b.nseq) {
            sl = fse_base(el) + bs.read((int)fse_nb(el));
            sm = fse_base(em) + bs.read((int)fse_nb(em));
            so = fse_base(eo) + bs.read((int)fse_nb(eo));
        }
        ZSeq s; s.ll = ll; s.ml = ml; s.of = ofv; s.lit_rel = (u32)lit_total; s.dst_rel = (u32)(lit_total + match_total); s.blk = i;
        seq[k] = s;
        repfn_step(rf, ofv, ll);
        lit_total += ll; match_total += ml;
    }
    if (bs.overrun() || !bs.exact()) { zerr(a, Z_ERR_SEQ_STREAM, i); b.nseq = 0; return; }
    if (lit_total > b.lit_regen || lit_total + match_total > 128 * 1024 + 0u) { zerr(a, Z_ERR_SIZE, i); b.nseq = 0; return; }
    b.match_total = (u32)match_total; b.repfn = rf;
}

// S2 — output offsets and repeat-offset history
struct ScanB { u64 out; u64 out_base; u64 frame_out; RepFn rf; u8 has_stream_start, has_frame_start; };

HD RepFn repfn_compose(const RepFn &f, const RepFn &g)      // first f, then g
{
    RepFn h;
    for (int k = 0; k < 3; k++) {
        if (g.s[k].src < 0) h.s[k] = g.s[k];
        else { h.s[k] = f.s[g.s[k].src]; if (h.s[k].src < 0) h.s[k].value += (u32)g.s[k].delta; else h.s[k].delta += g.s[k].delta; }
    }
    return h;
}
HD RepFn repfn_start()
{
    RepFn f; f.s[0] = repslot_concrete(1); f.s[1] = repslot_concrete(4); f.s[2] = repslot_concrete(8); return f;
}
HD u32 blk_out_size(const ZBlock &b) { return b.type == 2 ? b.lit_regen + b.match_total : b.rsize; }
// running state: out = arena offset of the next byte; frame_out = arena offset of current frame start
HD void scanb_push(ScanB &s, const ZBlock &b)
{
    if (b.first_in_stream) { s.out = b.out_base; s.has_stream_start = 1; }
    if (b.first_in_frame) { s.frame_out = s.out; s.rf = repfn_start(); s.has_frame_start = 1; }
    s.out += blk_out_size(b);
    if (b.type == 2 && b.nseq > 0) s.rf = repfn_compose(s.rf, b.repfn);
}
HD void k_scan2_phase1(const ZDecArgs &a, u32 c)
{
    u32 lo, hi; chunk_range(a.nblk, a.nchunks, c, lo, hi);
    ScanB s; s.out = 0; s.out_base = 0; s.frame_out = 0; s.rf = repfn_identity(); s.has_stream_start = 0; s.has_frame_start = 0;
    for (u32 i = lo; i < hi; i++) scanb_push(s, a.blk[i]);
    a.scan_b[c] = s;
}
HD void k_scan2_phase2(const ZDecArgs &a)
{
    // carry: absolute `out`, frame_out, concrete reps (as a constant RepFn)
    ScanB run; run.out = 0; run.out_base = 0; run.frame_out = 0; run.rf = repfn_start(); run.has_stream_start = run.has_frame_start = 0;
    for (u32 c = 0; c < a.nchunks; c++) {
        ScanB mine = a.scan_b[c];
        a.scan_b[c] = run;
        // chunk aggregate `mine.out` is relative (sum) unless the chunk saw a stream start (then absolute)
        ScanB nxt;
        nxt.out = mine.has_stream_start ? mine.out : run.out + mine.out;
        // frame_out inside the chunk was computed against a relative `out`; phase 3 recomputes it, here we
        // only need the reps: a frame start inside the chunk makes the aggregate a constant function.
        nxt.rf = repfn_compose(run.rf, mine.rf);
        nxt.frame_out = 0; nxt.out_base = 0; nxt.has_stream_start = nxt.has_frame_start = 0;
        run = nxt;
    }
}
HD void k_scan2_phase3(const ZDecArgs &a, u32 c)
{
    u32 lo, hi; chunk_range(a.nblk, a.nchunks, c, lo, hi);
    ScanB s = a.scan_b[c];
    // frame_out carry: recompute by looking at my frame's first block if it lies in an earlier chunk.
    for (u32 i = lo; i < hi; i++) {
        ZBlock &b = a.blk[i];
        if (b.first_in_stream) s.out = b.out_base;
        if (b.first_in_frame) { s.frame_out = s.out; s.rf = repfn_start(); }
        b.out_off = s.out;
        b.frame_out = s.frame_out;          // fixed up below when the frame started in an earlier chunk
        const u32 none[3] = { 0, 0, 0 };
        repfn_apply(s.rf, none, b.rep_in);  // s.rf is always constant here (carry is concrete)
        s.out += blk_out_size(b);
        if (b.type == 2 && b.nseq > 0) s.rf = repfn_compose(s.rf, b.repfn);
    }
}
// frame_out fix-up: every block copies out_off of the first block of its frame (runs after phase 3)
HD void k_frame_out(const ZDecArgs &a, u32 i) { a.blk[i].frame_out = a.blk[a.blk[i].frame_first_blk].out_off; }

// per-stream totals for the host (one thread per stream; walks only chunk boundaries cheaply = last block)
HD void k_stream_totals(const ZDecArgs &a, u32 i, ZStreamResult *res)
{
    // i = block index; the last block of each stream reports
    const ZBlock &b = a.blk[i];
    bool last = (i + 1 == a.nblk) || a.blk[i + 1].first_in_stream;
    if (last) { res[b.stream].out_size = b.out_off + blk_out_size(b) - b.out_base; res[b.stream].nseq = b.seq_cum; }
}

// K5 — one Huffman stream (thread t = 4*block + k)
// `staged`: the block's decode table already copied next to the thread (shared memory on the GPU), or nullptr
HD void k_literals(const ZDecArgs &a, u32 t, const u16 *staged = nullptr, u32 ring_slot0 = 0, u32 ring_stride = 0)
{
    u32 i = t >> 2, k = t & 3;
    const ZBlock &b = a.blk[i];
    if (b.type != 2 || b.lit_type < 2 || b.skip) return;
    if (b.lit_streams == 1 && k) return;
    if (b.huf_src < 0) return;
    const ZBlock &hb = a.blk[b.huf_src];
    if (hb.huf_bits == 0) return;
    const u16 *table = staged ? staged : a.huf_pool + (size_t)hb.huf_slot * HUF_SLOT_ENTRIES;
    u32 tree = b.lit_type == 2 ? b.tree_len : 0;
    const u8 *p = a.in + b.src + b.lit_hdr + tree;
    if (tree > b.lit_csize) { zerr(a, Z_ERR_HUF_STREAM, i); return; }
    u32 n = b.lit_csize - tree;
    u8 *dst = b.nseq == 0 ? a.out + b.out_off : a.lit_scratch + b.lit_off;
    if (b.lit_streams == 1) {
#ifdef __CUDA_ARCH__
        if (ring_stride) { if (!huf_decode_stream_ring(table, hb.huf_bits, p, n, dst, b.lit_regen, ring_slot0, ring_stride)) zerr(a, Z_ERR_HUF_STREAM, i); return; }
#endif
        if (!huf_decode_stream(table, hb.huf_bits, p, n, dst, b.lit_regen)) zerr(a, Z_ERR_HUF_STREAM, i);
        return;
    }
    if (n < 6) { if (k == 0) zerr(a, Z_ERR_HUF_STREAM, i); return; }
    u32 s1 = p[0] | (p[1] << 8), s2 = p[2] | (p[3] << 8), s3 = p[4] | (p[5] << 8);
    if (6 + s1 + s2 + s3 > n) { if (k == 0) zerr(a, Z_ERR_HUF_STREAM, i); return; }
    u32 seg = (b.lit_regen + 3) / 4;
    if (seg * 3 > b.lit_regen) { if (k == 0) zerr(a, Z_ERR_HUF_STREAM, i); return; }
    u32 off = 6, len = s1, cnt = seg;
    if (k >= 1) { off += s1; len = s2; }
    if (k >= 2) { off += s2; len = s3; }
    if (k == 3) { off += s3; len = n - 6 - s1 - s2 - s3; cnt = b.lit_regen - 3 * seg; }
#ifdef __CUDA_ARCH__
    if (ring_stride) { if (!huf_decode_stream_ring(table, hb.huf_bits, p + off, len, dst + (size_t)k * seg, cnt, ring_slot0, ring_stride)) zerr(a, Z_ERR_HUF_STREAM, i); return; }
#endif
    if (!huf_decode_stream(table, hb.huf_bits, p + off, len, dst + (size_t)k * seg, cnt)) zerr(a, Z_ERR_HUF_STREAM, i);
}

// K6 — raw / RLE blocks and raw / RLE literal sections (one thread group per block)
// word-wise copy / fill shared by a thread group (source may be unaligned: two aligned loads + funnel shift;
// input buffers carry >= 8 bytes of padding so the last aligned load stays in bounds)
HD void copy_span(u8 *d, const u8 *s, u32 n, u32 tid, u32 nt)
{
    u32 head = (u32)((4 - ((uintptr_t)d & 3)) & 3); if (head > n) head = n;
    for (u32 k = tid; k < head; k += nt) d[k] = s[k];
    const u32 nw = (n - head) / 4;
    const u8 *s2 = s + head; u32 *dw = (u32 *)(d + head);
    const u32 sh = (u32)((uintptr_t)s2 & 3) * 8; const u32 *sa = (const u32 *)((uintptr_t)s2 & ~(uintptr_t)3);
    for (u32 k = tid; k < nw; k += nt) { u32 lo = sa[k]; dw[k] = sh ? (lo >> sh) | (sa[k + 1] << (32 - sh)) : lo; }
    for (u32 k = head + nw * 4 + tid; k < n; k += nt) d[k] = s[k];
}
HD void fill_span(u8 *d, u8 v, u32 n, u32 tid, u32 nt)
{
    u32 head = (u32)((4 - ((uintptr_t)d & 3)) & 3); if (head > n) head = n;
    for (u32 k = tid; k < head; k += nt) d[k] = v;
    const u32 nw = (n - head) / 4; u32 *dw = (u32 *)(d + head); const u32 vv = v * 0x01010101u;
    for (u32 k = tid; k < nw; k += nt) dw[k] = vv;
    for (u32 k = head + nw * 4 + tid; k < n; k += nt) d[k] = v;
}
HD void k_copy_block(const ZDecArgs &a, u32 i, u32 tid, u32 nthreads)
{
    const ZBlock &b = a.blk[i];
    if (b.skip) return;
    if (b.type == 0) { copy_span(a.out + b.out_off, a.in + b.src, b.rsize, tid, nthreads); return; }
    if (b.type == 1) { fill_span(a.out + b.out_off, a.in[b.src], b.rsize, tid, nthreads); return; }
    if (b.lit_type >= 2) return;
    u8 *d = b.nseq == 0 ? a.out + b.out_off : a.lit_scratch + b.lit_off;
    const u8 *s = a.in + b.src + b.lit_hdr;
    if (b.lit_type == 0) copy_span(d, s, b.lit_regen, tid, nthreads);
    else fill_span(d, s[0], b.lit_regen, tid, nthreads);
}

// K7 — resolve repeat offsets of one block and validate them (spec "Repeat Offsets")
// Replaces the offset part of zstd_decompress_block.c:937 ZSTD_decodeSequence.
HD void k_seq_resolve(const ZDecArgs &a, u32 i)
{
    const ZBlock &b = a.blk[i];
    if (b.type != 2 || b.nseq == 0) return;
    u32 r0 = b.rep_in[0], r1 = b.rep_in[1], r2 = b.rep_in[2];
    ZSeq *seq = a.seq + b.seq_base;
    // batches of 8: all loads of a batch are issued before the (serial) history update needs them
    for (u32 k0 = 0; k0 < b.nseq; k0 += 8) {
        const u32 m = b.nseq - k0 < 8 ? b.nseq - k0 : 8;
        u32 ofv[8], ll[8], dr[8];
        for (u32 j = 0; j < 8; j++) if (j < m) { ofv[j] = seq[k0 + j].of; ll[j] = seq[k0 + j].ll; dr[j] = seq[k0 + j].dst_rel; }
        for (u32 j = 0; j < 8; j++) if (j < m) {
            u32 off;
            if (ofv[j] > 3) { off = ofv[j] - 3; r2 = r1; r1 = r0; r0 = off; }
            else {
                u32 idx = ofv[j] - 1 + (ll[j] == 0 ? 1u : 0u);
                if (idx == 0) off = r0;
                else {
                    off = idx == 1 ? r1 : (idx == 2 ? r2 : r0 - 1);
                    if (idx != 1) r2 = r1;
                    r1 = r0; r0 = off;
                }
            }
            u64 match_pos = b.out_off + dr[j] + ll[j];
            if (off == 0 || off > match_pos - b.frame_out) { zerr(a, Z_ERR_OFFSET, i); off = 0; seq[k0 + j].ml = 0; }
            ofv[j] = off;
        }
        for (u32 j = 0; j < 8; j++) if (j < m) seq[k0 + j].of = ofv[j];
    }
}

// K7b — a block whose matches all start inside the block itself needs neither back-pointers nor pointer jumping: one
// thread runs its sequences in order (zstd_decompress_block.c:804 ZSTD_execSequence as it stands).  That is every block
// our own encoder writes (blocks are kept independent of each other; 8 KB when they carry sequences), so a
// file of ours never enters K9.  Bigger blocks (a reference-made 128 KB first block qualifies too) stay on the
// parallel path: a serial walk over them would take longer than the jumps it saves.
static const u32 LOCAL_MAX_OUT = 16 * 1024;
// Forward copy by one thread, up to 16 bytes per step: all loads of a step are issued before its stores, so the thread
// waits for memory once per step instead of once per byte (a byte loop costs one L2 round trip per byte: measured
// 3.8 ms for 0.25 GB of ids on B200).  `dist` = distance from source to destination when they overlap (16 = apart).
HD void copy_fwd16(u8 *dst, const u8 *src, u32 n, u32 dist)
{
    const u32 step = dist >= 16 ? 16 : (dist >= 8 ? 8 : (dist >= 4 ? 4 : 1));
    u32 i = 0;
    if (dist < 16 && (dist & (dist - 1)) == 0 && n >= 32) {
        // a long match at offset 1, 2, 4 or 8 repeats those bytes (a block of equal length units is ONE such match of 8 KB): with the
        // period in registers the steps are stores only, instead of 2,047 store -> load round trips through memory (measured: the
        // 3.1 ms of zd_block_local at 1 M reads were this one thread's)
        u8 t[16];
        for (u32 k = 0; k < 16; k++) t[k] = src[k & (dist - 1)];
        for (; i + 16 <= n; i += 16) for (int k = 0; k < 16; k++) dst[i + k] = t[k];
        for (u32 k = 0; i < n; i++, k++) dst[i] = t[k];
        return;
    }
    if (step == 16) for (; i + 16 <= n; i += 16) { u8 t[16]; for (int k = 0; k < 16; k++) t[k] = src[i + k]; for (int k = 0; k < 16; k++) dst[i + k] = t[k]; }
    else if (step == 8) for (; i + 8 <= n; i += 8) { u8 t[8]; for (int k = 0; k < 8; k++) t[k] = src[i + k]; for (int k = 0; k < 8; k++) dst[i + k] = t[k]; }
    else if (step == 4) for (; i + 4 <= n; i += 4) { u8 t[4]; for (int k = 0; k < 4; k++) t[k] = src[i + k]; for (int k = 0; k < 4; k++) dst[i + k] = t[k]; }
    for (; i < n; i++) dst[i] = src[i];
}
HD void k_block_local(const ZDecArgs &a, u32 i)
{
    ZBlock &b = a.blk[i];
    if (b.type != 2 || b.nseq == 0 || b.skip || b.lit_regen + b.match_total > LOCAL_MAX_OUT) return;
    const ZSeq *seq = a.seq + b.seq_base;
    for (u32 k = 0; k < b.nseq; k++) if (seq[k].ml && seq[k].of > seq[k].dst_rel + seq[k].ll) return;
    u8 *o = a.out + b.out_off; const u8 *lit = a.lit_scratch + b.lit_off;
    u32 d = 0, l = 0;
    for (u32 k = 0; k < b.nseq; k++) {
        const u32 ll = seq[k].ll, ml = seq[k].ml, of = seq[k].of;
        copy_fwd16(o + d, lit + l, ll, 16);
        d += ll; l += ll;
        copy_fwd16(o + d, o + d - of, ml, of);                  // a match may overlap its own output (offset < length)
        d += ml;
    }
    copy_fwd16(o + d, lit + l, b.lit_regen - l, 16);
    b.local = 1;
}

HD void set_bits(u32 *bitmap, u64 lo, u64 hi)     // [lo, hi) in bit coordinates
{
    while (lo < hi) {
        u64 w = lo >> 5; u32 b0 = (u32)(lo & 31); u64 wend = (w + 1) << 5; u64 e = hi < wend ? hi : wend;
        u32 nb = (u32)(e - lo);
        u32 m = nb == 32 ? 0xFFFFFFFFu : (((1u << nb) - 1) << b0);
#ifdef __CUDA_ARCH__
        atomicOr(bitmap + w, m);
#else
        bitmap[w] |= m;
#endif
        lo = e;
    }
}

// K8 — execute one sequence: literals to their place, one back-pointer per match byte.
// Replaces zstd_decompress_block.c:804 ZSTD_execSequence; the serial dependency between overlapping
// matches is removed by K9's pointer jumping.  (tid, nthreads) stride lets a whole CTA share a long one.
HD void k_seq_exec_one(const ZDecArgs &a, const ZSeq &s, u32 tid, u32 nthreads)
{
    const ZBlock &b = a.blk[s.blk];
    u64 dst = b.out_off + s.dst_rel;
    const u8 *lit = a.lit_scratch + b.lit_off + s.lit_rel;
    u8 *o = a.out + dst;
    for (u32 k = tid; k < s.ll; k += nthreads) o[k] = lit[k];
    if (s.ml == 0) return;
    u64 m = dst + s.ll;                      // first match byte
    u64 src0 = m - s.of;
    u32 *link = a.link + (m - a.span_lo);
    u32 rel = (u32)(src0 - a.span_lo);
    if (s.of >= s.ml) for (u32 k = tid; k < s.ml; k += nthreads) link[k] = rel + k;
    else for (u32 k = tid; k < s.ml; k += nthreads) link[k] = rel + k % s.of;     // periodic match: point into the first period
    if (tid == 0) set_bits(a.bitmap, m - a.span_lo, m - a.span_lo + s.ml);
}
static const u32 BIG_SEQ = 512;
HD void k_seq_exec_small(const ZDecArgs &a, u64 j)
{
    const ZSeq s = a.seq[j];
    if (s.ll + s.ml > BIG_SEQ || a.blk[s.blk].local) return;
    k_seq_exec_one(a, s, 0, 1);
}
// one thread group per block: long sequences + the literals after the last sequence
HD void k_seq_exec_big(const ZDecArgs &a, u32 i, u32 tid, u32 nthreads)
{
    const ZBlock &b = a.blk[i];
    if (b.type != 2 || b.nseq == 0 || b.local) return;
    const ZSeq *seq = a.seq + b.seq_base;
    for (u32 k = 0; k < b.nseq; k++) if (seq[k].ll + seq[k].ml > BIG_SEQ) k_seq_exec_one(a, seq[k], tid, nthreads);
    const ZSeq &last = seq[b.nseq - 1];
    u32 lit_used = last.lit_rel + last.ll, out_used = last.dst_rel + last.ll + last.ml;
    const u8 *lit = a.lit_scratch + b.lit_off + lit_used; u8 *o = a.out + b.out_off + out_used;
    for (u32 k = tid; k + lit_used < b.lit_regen; k += nthreads) o[k] = lit[k];
}

// K9 — one pointer-jumping step over 32 positions (one bitmap word)
HD void k_jump(const ZDecArgs &a, u64 w)
{
    u32 bits = a.bitmap[w];
    if (bits == 0) return;
    u32 clear = 0;
    for (u32 rest = bits; rest;) {
#ifdef __CUDA_ARCH__
        int k = __ffs((int)rest) - 1;
#else
        int k = __builtin_ctz(rest);
#endif
        rest &= rest - 1;
        u64 p = (w << 5) + (u32)k;
        u32 s = a.link[p];
#ifdef __CUDA_ARCH__
        u32 sw = *((volatile u32 *)(a.bitmap + (s >> 5)));
#else
        u32 sw = a.bitmap[s >> 5];
#endif
        if (((sw >> (s & 31)) & 1) == 0) {
#ifdef __CUDA_ARCH__
            __threadfence();                 // pairs with the fence before the writer's bit clear
            a.out[a.span_lo + p] = *((volatile u8 *)(a.out + a.span_lo + s));
#else
            a.out[a.span_lo + p] = a.out[a.span_lo + s];
#endif
            clear |= 1u << k;
        } else {
#ifdef __CUDA_ARCH__
            a.link[p] = *((volatile u32 *)(a.link + s));
#else
            a.link[p] = a.link[s];
#endif
        }
    }
    if (clear) {
#ifdef __CUDA_ARCH__
        __threadfence();
        atomicAnd(a.bitmap + w, ~clear);
#else
        a.bitmap[w] &= ~clear;
#endif
    }
    if (clear != bits) a.status[2] = 1;
}

// K4 / K7 / K8-big launches: generic executors run the HD bodies as they are; the CUDA executor overloads them
// (zstd_dec_cuda.cuh) with kernels that stage tables / sequence fields in shared memory.
template <class Exec> void launch_seq_decode(Exec &ex, const ZDecArgs &a)
{
    ex.for_each(a.nblk, [=] HDN (size_t i) { k_seq_decode(a, (u32)i); }, "zd_seq_decode", 32);
}
template <class Exec> void launch_seq_resolve(Exec &ex, const ZDecArgs &a)
{
    ex.for_each(a.nblk, [=] HDN (size_t i) { k_seq_resolve(a, (u32)i); }, "zd_seq_resolve", 32);
}
template <class Exec> void launch_seq_exec_big(Exec &ex, const ZDecArgs &a)
{
    ex.for_each_group(a.nblk, 256, [=] HDN (size_t i, unsigned tid, unsigned nt) { k_seq_exec_big(a, (u32)i, tid, nt); }, "zd_seq_exec_big");
}

// K5 launch: generic executors run one thread per Huffman stream straight from the table pool; the CUDA
// executor overloads this (zstd_dec_cuda.cuh) with a kernel that first stages the tables in shared memory.
template <class Exec> void launch_literals(Exec &ex, const ZDecArgs &a)
{
    ex.for_each((size_t)a.nblk * 4, [=] HDN (size_t t) { k_literals(a, (u32)t); }, "zd_literals", 64);
}

// ------------------------------------------------------------------ orchestration (templated on the executor)
//
// Exec provides:
//   T*   alloc<T>(count)                       scratch from the arena (lifetime: this call)
//   void upload(dst, src_host, bytes) / download(dst_host, src, bytes) [synchronises] / zero(ptr, bytes)
//   void for_each(n, F(index))                 grid of n threads
//   void for_each_group(ngroups, threads, F(group, tid, nthreads))
struct ZDecPlan {
    // simple = every block is self-contained (raw, RLE, or compressed with its own Huffman table and no sequences) AND the
    // caller knows where each one's output goes: blocks[i].out_base is block i's own arena offset and, for a compressed block,
    // blocks[i].rsize the bytes it must regenerate.  Then nothing depends on another block and the two block scans (table
    // provenance, output offsets, repeat-offset history) are skipped.
    bool simple = false;
    std::vector<ZBlockHead> blocks;
    std::vector<ZStreamDesc> streams;
    std::vector<ZStreamResult> results;
};

template <class Exec>
int zstd_decode_blocks(Exec &ex, const u8 *d_in, u8 *d_out, ZDecPlan &plan, const u32 *d_predef, std::string &err);

// host walk of every stream of the plan -> plan.blocks
inline int zstd_walk_plan(const u8 *h_in, ZDecPlan &plan, std::string &err)
{
    plan.blocks.clear();
    plan.results.assign(plan.streams.size(), ZStreamResult{0, 0, 0});
    {
        // one host thread per big stream (the walks are independent chains of cache misses), the small ones inline
        const size_t ns = plan.streams.size();
        std::vector<std::vector<ZBlockHead>> part(ns);
        std::vector<std::string> errs(ns);
        std::vector<int> rcs(ns, 0);
        std::vector<u64> used(ns, 0);
        std::vector<std::thread> workers;
        for (size_t s = 0; s < ns; s++) if (plan.streams[s].src_len == 0) { err = "empty zstd stream"; return -1; }
        auto walk = [&](size_t s) { rcs[s] = zstd_walk_stream(h_in, plan.streams[s], (int)s, part[s], &used[s], errs[s]); };
        for (size_t s = 0; s < ns; s++) if (ns > 1 && plan.streams[s].src_len > (8u << 20)) workers.emplace_back(walk, s);
        for (size_t s = 0; s < ns; s++) if (!(ns > 1 && plan.streams[s].src_len > (8u << 20))) walk(s);
        for (auto &w : workers) w.join();
        for (size_t s = 0; s < ns; s++) {
            if (rcs[s]) { err = errs[s]; return -1; }
            const u32 base = (u32)plan.blocks.size();
            for (auto &b : part[s]) { ZBlockHead h = b; h.frame_first_blk += base; plan.blocks.push_back(h); }
            plan.results[s].consumed = used[s];
        }
    }
    return 0;
}

template <class Exec>
int zstd_decode_batch(Exec &ex, const u8 *d_in, const u8 *h_in, u8 *d_out, ZDecPlan &plan, const u32 *d_predef, std::string &err)
{
    if (int rc = zstd_walk_plan(h_in, plan, err)) return rc;
    return zstd_decode_blocks(ex, d_in, d_out, plan, d_predef, err);
}

// device side: plan.blocks (from zstd_walk_plan, or assembled by the caller from walks it ran itself) -> bytes
template <class Exec>
int zstd_decode_blocks(Exec &ex, const u8 *d_in, u8 *d_out, ZDecPlan &plan, const u32 *d_predef, std::string &err)
{
    if (plan.results.size() != plan.streams.size()) plan.results.assign(plan.streams.size(), ZStreamResult{0, 0, 0});
    const u32 nblk = (u32)plan.blocks.size();
    if (nblk == 0) return 0;
    u32 n_comp = 0;
    for (auto &b : plan.blocks) n_comp += b.type == 2 || b.type == 0xFF;

    ZDecArgs a; memset(&a, 0, sizeof a);
    a.in = d_in; a.out = d_out; a.nblk = nblk; a.predef = d_predef;
    a.blk = ex.template alloc<ZBlock>(nblk);
    a.status = ex.template alloc<u32>(4);
    const ZBlockHead *plan_heads = nullptr;
    {
        ZBlockHead *heads = ex.template alloc<ZBlockHead>(nblk);
        plan_heads = heads;
        ex.upload_staged(heads, plan.blocks.data(), sizeof(ZBlockHead) * nblk);
        ZBlock *blk = a.blk;
        ex.for_each(nblk, [=] HDN (size_t i) {
            const ZBlockHead h = heads[i];
            ZBlock b; memset(&b, 0, sizeof b);
            b.src = h.src; b.out_base = h.out_base; b.csize = h.csize; b.rsize = h.rsize; b.frame_first_blk = h.frame_first_blk;
            b.type = h.type; b.first_in_frame = h.first_in_frame; b.first_in_stream = h.first_in_stream; b.stream = h.stream;
            b.huf_src = b.ll_src = b.of_src = b.ml_src = -1;
            blk[i] = b;
        }, "zd_block_headers");
    }
    ex.zero(a.status, 16);
    if (plan.simple) {
        a.huf_pool = ex.template alloc<u16>((size_t)(n_comp ? nblk : 1) * HUF_SLOT_ENTRIES);
        a.fse_pool = nullptr; a.lit_scratch = nullptr; a.seq = nullptr;
        const ZBlockHead *heads = plan_heads;
        ex.for_each(nblk, [=] HDN (size_t i) {
            ZBlock &b = a.blk[i];
            if (b.type == 0xFF) {                                 // from a block index: the header decides, and must agree with it
                const u8 *hp = a.in + b.src - 3;
                const u32 bh = hp[0] | (hp[1] << 8) | ((u32)hp[2] << 16), type = (bh >> 1) & 3, bsize = bh >> 3, want = heads[i].rsize;
                bool ok = type != 3;
                if (type == 1) { ok = ok && b.csize == 1 && bsize == want; b.rsize = bsize; }
                else if (type == 0) { ok = ok && bsize == b.csize && bsize == want; b.rsize = bsize; }
                else { ok = ok && bsize == b.csize; b.rsize = 0; }
                if (!ok) { zerr(a, Z_ERR_SIZE, (u32)i); b.type = 0; b.rsize = 0; b.csize = 0; return; }
                b.type = (u8)type;
            }
            k_block_headers(a, (u32)i);
            b.out_off = heads[i].out_base; b.frame_out = b.out_off; b.seq_cum = 0;
            if (b.type == 2) {
                if (b.nseq != 0 || b.lit_type == 3 || b.lit_regen != heads[i].rsize) { zerr(a, Z_ERR_SIZE, (u32)i); b.type = 0; b.rsize = 0; b.csize = 0; return; }
                b.huf_src = b.lit_type == 2 ? (i32)i : -1; b.huf_slot = (u32)i;
            }
        }, "zd_block_headers");
        if (n_comp) ex.for_each(nblk, [=] HDN (size_t i) { k_huf_table(a, (u32)i); }, "zd_huf_table", 64);
        if (n_comp) launch_literals(ex, a);
        ex.for_each_group(nblk, 256, [=] HDN (size_t i, unsigned tid, unsigned nt) { k_copy_block(a, (u32)i, tid, nt); }, "zd_copy_block");
        u32 st[4];
        ex.download(st, a.status, 16);
        if (st[0]) { err = "corrupt zstd data (code " + std::to_string(st[0]) + ", block " + std::to_string(st[1]) + ")"; return -1; }
        // every compressed block regenerated exactly what the caller expected of it (checked above), raw / RLE sizes were read
        // by the host: the streams' sizes are the sums of those
        for (auto &r : plan.results) { r.out_size = 0; r.nseq = 0; }
        for (auto &b : plan.blocks) plan.results[b.stream].out_size += b.rsize;
        return 0;
    }
    // Two block scans in three phases: per-chunk serial (~1.1 us per block, measured on B200), ONE thread over the chunk
    // aggregates (~0.66 us per chunk), per-chunk serial again.  2 * 1.1 * nblk / c + 0.66 * c is least at c = sqrt(3.3 nblk).
    {
        u32 c = 1; while ((u64)c * c < (u64)nblk * 33 / 10) c++;
        a.nchunks = nblk < 256 ? (nblk + 7) / 8 : (c > 1024 ? 1024 : c); if (a.nchunks == 0) a.nchunks = 1;
    }
    a.scan_a = ex.template alloc<ScanA>(a.nchunks + 1);
    a.scan_b = ex.template alloc<ScanB>(a.nchunks);
    ZStreamResult *d_res = ex.template alloc<ZStreamResult>(plan.streams.size());
    ex.zero(d_res, sizeof(ZStreamResult) * plan.streams.size());

    if (n_comp) {
        ex.for_each(nblk, [=] HDN (size_t i) { k_block_headers(a, (u32)i); }, "zd_block_headers");
        ex.for_each(a.nchunks, [=] HDN (size_t c) { k_scan1_phase1(a, (u32)c); }, "zd_scan1");
        ex.for_each(1, [=] HDN (size_t) { k_scan1_phase2(a); }, "zd_scan1");
        ex.for_each(a.nchunks, [=] HDN (size_t c) { k_scan1_phase3(a, (u32)c); }, "zd_scan1");
    }
    // Sizes of the pools / scratch are bounded without a device round trip:
    //   Huffman tables <= compressed blocks, FSE slots <= compressed blocks, literal scratch and number of
    //   sequences are only known on the device -> fetch the totals (one small synchronising copy).
    u64 tot_lit = 0, tot_seq = 0; u32 tot_huf = 0, tot_fse = 0;
    if (n_comp) {
        ScanA tot; ex.download(&tot, a.scan_a + a.nchunks, sizeof tot);
        tot_huf = tot.huf_slots; tot_fse = tot.fse_slots; tot_lit = tot.lit; tot_seq = tot.seqs;
        u32 st[4]; ex.download(st, a.status, 16);
        if (st[0]) { err = "corrupt zstd block headers (code " + std::to_string(st[0]) + ", block " + std::to_string(st[1]) + ")"; return -1; }
    }
    a.huf_pool = ex.template alloc<u16>((size_t)(tot_huf ? tot_huf : 1) * HUF_SLOT_ENTRIES);
    a.fse_pool = ex.template alloc<u32>((size_t)(tot_fse ? tot_fse : 1) * FSE_SLOT_ENTRIES);
    a.lit_scratch = ex.template alloc<u8>(tot_lit + 16);
    a.seq = ex.template alloc<ZSeq>(tot_seq + 1);

    if (tot_huf) ex.for_each(nblk, [=] HDN (size_t i) { k_huf_table(a, (u32)i); }, "zd_huf_table", 64);
    if (tot_seq) {
        ex.for_each(nblk, [=] HDN (size_t i) { k_fse_tables(a, (u32)i); }, "zd_fse_tables", 32);
        launch_seq_decode(ex, a);
    }
    ex.for_each(a.nchunks, [=] HDN (size_t c) { k_scan2_phase1(a, (u32)c); }, "zd_scan2");
    ex.for_each(1, [=] HDN (size_t) { k_scan2_phase2(a); }, "zd_scan2");
    ex.for_each(a.nchunks, [=] HDN (size_t c) { k_scan2_phase3(a, (u32)c); }, "zd_scan2");
    ex.for_each(nblk, [=] HDN (size_t i) { k_frame_out(a, (u32)i); }, "zd_scan2");
    ex.for_each(nblk, [=] HDN (size_t i) { k_stream_totals(a, (u32)i, d_res); }, "zd_scan2");

    // The regenerated sizes must match what the container promised before anything is written.
    {
        std::vector<ZStreamResult> r(plan.streams.size());
        ex.download(r.data(), d_res, sizeof(ZStreamResult) * r.size());
        u32 st[4]; ex.download(st, a.status, 16);
        if (st[0]) { err = "corrupt zstd stream (code " + std::to_string(st[0]) + ", block " + std::to_string(st[1]) + ")"; return -1; }
        for (size_t s = 0; s < r.size(); s++) {
            plan.results[s].out_size = r[s].out_size;
            plan.results[s].nseq = r[s].nseq - (s ? r[s - 1].nseq : 0);     // r[].nseq is cumulative over blocks
            if (r[s].out_size > plan.streams[s].out_size) { err = "zstd stream regenerates more bytes than expected"; return -1; }
        }
    }

    // record-range decode: blocks whose output lies outside the wanted bytes of a sequence-free stream are left alone
    {
        ZNeedTab nt; memset(&nt, 0, sizeof nt); bool any = false;
        for (size_t s = 0; s < plan.streams.size() && s < 8; s++) {
            const ZStreamDesc &sd = plan.streams[s];
            nt.on[s] = plan.results[s].nseq == 0 && (sd.need_lo > 0 || sd.need_hi < plan.results[s].out_size);
            nt.lo[s] = sd.out_off + sd.need_lo; nt.hi[s] = sd.need_hi == ~0ull ? ~0ull : sd.out_off + sd.need_hi;
            any = any || nt.on[s];
        }
        if (any && plan.streams.size() <= 8) {
            const ZNeedTab t = nt;
            ex.for_each(nblk, [=] HDN (size_t i) {
                ZBlock &b = a.blk[i];
                const u64 lo = b.out_off, hi = lo + blk_out_size(b);
                if (t.on[b.stream] && (hi <= t.lo[b.stream] || lo >= t.hi[b.stream])) b.skip = 1;
            }, "zd_need");
        }
    }
    if (n_comp) launch_literals(ex, a);
    ex.for_each_group(nblk, 256, [=] HDN (size_t i, unsigned tid, unsigned nt) { k_copy_block(a, (u32)i, tid, nt); }, "zd_copy_block");

    if (tot_seq) {
        // span of the arena that pointer jumping may touch: streams that contain sequences
        u64 lo = ~0ull, hi = 0;
        for (size_t s = 0; s < plan.streams.size(); s++) {
            if (!plan.results[s].nseq) continue;
            if (plan.streams[s].out_off < lo) lo = plan.streams[s].out_off;
            if (plan.streams[s].out_off + plan.results[s].out_size > hi) hi = plan.streams[s].out_off + plan.results[s].out_size;
        }
        lo &= ~31ull;
        if (hi - lo >= 0xFFFFFFFFull) { err = "zstd streams with matches span more than 4 GiB; not supported by this build"; return -2; }
        u64 span = hi - lo, words = (span + 31) / 32;
        a.span_lo = lo;
        a.link = ex.template alloc<u32>(span + 1);
        a.bitmap = ex.template alloc<u32>(words + 1);
        ex.zero(a.bitmap, (words + 1) * 4);
        launch_seq_resolve(ex, a);
        ex.for_each(nblk, [=] HDN (size_t i) { k_block_local(a, (u32)i); }, "zd_block_local", 32);
        ex.for_each(tot_seq, [=] HDN (size_t j) { k_seq_exec_small(a, j); }, "zd_seq_exec_small");
        launch_seq_exec_big(ex, a);
        for (int round = 0; round < 40; round++) {
            ex.zero(a.status + 2, 4);
            for (int k = 0; k < 4; k++) ex.for_each(words, [=] HDN (size_t w) { k_jump(a, w); }, "zd_jump");
            u32 st[4]; ex.download(st, a.status, 16);
            if (st[0]) { err = "corrupt zstd sequences (code " + std::to_string(st[0]) + ", block " + std::to_string(st[1]) + ")"; return -1; }
            if (!st[2]) break;
            if (round == 39) { err = "zstd match resolution did not converge"; return -1; }
        }
    }
    u32 st[4]; ex.download(st, a.status, 16);
    if (st[0]) { err = "corrupt zstd data (code " + std::to_string(st[0]) + ", block " + std::to_string(st[1]) + ")"; return -1; }
    return 0;
}

// predefined LL | OF | ML tables in slot layout (host builds once, uploads)
inline void zstd_build_predef(u32 *slot)
{
    SeqConsts c; seq_consts_init(c);
    u16 next[64];
    fse_build_table(slot, c.ll_norm, 36, 6, next);
    fse_build_table(slot + FSE_OF_AT, c.of_norm, 29, 5, next);
    fse_build_table(slot + FSE_ML_AT, c.ml_norm, 53, 6, next);
}

}  // namespace nafz
