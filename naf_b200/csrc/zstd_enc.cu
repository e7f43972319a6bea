// zstd_enc.cu — block-parallel zstd *encoder* (included by naf_enc.cu; same translation unit).
//
// Replaces what ennaf gets from ZSTD_initCStream / ZSTD_compressStream / ZSTD_endStream
// (ennaf/src/compressor.c:7-20,120,64) for every NAF stream.  The reference's encoder is free-parse:
// parity on this side is "the reference unnaf decodes our frame back to the same bytes"
// (SURVEY §8a row 19), so the parse is chosen for the GPU:
//   * one frame per stream (unnaf's SEQ/QUAL loops stop after the first frame: SURVEY A.2), FHD 0x00,
//     declared window 128 KB, no content size / checksum / dictionary — what ennaf's frames look like
//   * 32 KB blocks, each independent of every other: own Huffman table, no sequences, so both our
//     decoder and libzstd can start anywhere; RLE block when a block is one repeated byte, raw block
//     when Huffman would not shrink it
//   * literals: canonical length-limited (<= 11 bit) Huffman, 4 streams (1 stream below 1 KB), tree
//     description as FSE-compressed weights when smaller / required (> 128 listed weights)
// Per block: histogram -> code lengths -> tree description -> bit-exact sizes -> encode into a fixed slot
// (k_zenc_hist / k_zenc_tables / k_zenc_encode); a scan over block sizes then lets k_zenc_gather lay the blocks
// out as frames.
//   * the small, text-like streams (ids, comments, lengths) at level >= 2: 8 KB blocks with LZ77 matches
//     (compress/zstd_fast.c:186 ZSTD_compressBlock_fast's role) and FSE-coded sequences (compress/zstd_compress_sequences.c:418).
//     Data-parallel stage (k_zlc_find / k_zlc_define / k_zlc_finish; bodies: zstd_lzc_hd.cuh, zstd_lzc_bytes_hd.cuh): a CTA per
//     block finds the matches with maps and scans, the stream gets ONE Huffman code and ONE set of FSE tables, a thread per block
//     codes against them (Treeless_Literals / Repeat_Mode).  NAFGPU_LZ=1: the first formulation, one thread per block doing
//     everything with private tables (k_zenc_lz; body in zstd_enc_hd.cuh).  All of it is HD code validated on the CPU against
//     libzstd.  Blocks stay independent in what they reference -- matches inside the block, repeat-offset codes only for offsets
//     the block itself pushed -- and each shard of a multi-GPU encode defines its tables anew
//
// Format: zstd/doc/zstd_compression_format.md ("Huffman Tree Description", "Huffman-coded streams",
// "FSE Table Description"); reference counterparts: compress/huf_compress.c:513 HUF_buildCTable_wksp,
// :116 HUF_writeCTable_wksp, compress/fse_compress.c:437 FSE_normalizeCount, :292 FSE_writeNCount,
// compress/zstd_compress_literals.c:70 ZSTD_compressLiterals, zstd_compress.c:3967 ZSTD_writeFrameHeader.

#include "zstd_enc_hd.cuh"
#include "zstd_lzc_hd.cuh"
#include "zstd_lzc_bytes_hd.cuh"

namespace nafg {

using nafz::ZEncMeta; using nafz::BitW;

static const u32 ZBS = 32 * 1024;            // uncompressed bytes per block (Huffman-only streams)
static const u32 ZSLOT = ZBS + 512;          // bytes reserved per block for its compressed content
static const u32 ZLB_MAX = 8 * 1024;         // uncompressed bytes per block of an LZ stream (one thread encodes a block)
static u32 zlb_bytes()                       // NAFGPU_ZLB=1024..8192 (A/B measurements); a thread's latency is proportional to it
{
    static const u32 v = [] { const char *e = getenv("NAFGPU_ZLB"); const u32 x = e ? (u32)atoi(e) : 0u; return x >= 256 && x <= ZLB_MAX ? (x & ~63u) : ZLB_MAX; }();
    return v;
}
static const int ZWINDOW_LOG = 17;

struct ZEncBlock {
    const u8 *src; u32 n; u32 stream; u32 last;
    u32 type;      // 0 raw (content = src bytes), 1 RLE, 2 compressed (content in slot)
    u32 csize;     // content bytes
    u32 lz;        // 1: block of an LZ stream (k_zenc_lz does everything), 2: compressed earlier, behind the upload; the three Huffman kernels pass
    u64 slot_off;  // where my slot starts in the slot pool
};

struct ZEncBatch {
    std::vector<const u8 *> src; std::vector<u64> n; std::vector<int> wlog; std::vector<int> lz;
    std::vector<u64> frame_size; std::vector<u8 *> dest;
    std::vector<u32> first_block;            // per stream (+1 sentinel)
    ZEncBlock *d_blocks = nullptr; u32 nblocks = 0; u8 *d_slots = nullptr; u64 *d_off = nullptr;
    bool final_shard = true;                 // false: these frames continue in another shard, no block carries Last_Block
    void add(const u8 *p, u64 bytes, int window_log, bool with_lz = false) { src.push_back(p); n.push_back(bytes); wlog.push_back(window_log); lz.push_back(with_lz ? 1 : 0); }
};

struct ZEncArgs { ZEncBlock *blk; u8 *slots; };      // blk: the first block the three Huffman kernels look at (LZ streams lie before it)
struct ZEncStreamTab { const u8 *src[8]; u64 n[8]; u64 slot_base[8]; u32 first[9]; u32 bs[8]; u32 lz[8]; u32 ns;
                       const ZEncBlock *early[8]; const u8 *early_slots[8]; u32 early_done[8]; };
struct ZEncFirstBlocks { u32 v[9]; };

// Three kernels per batch of 32 KB blocks.  Shared-memory atomics cost 32-64 cycles per warp instruction on this
// part and a single thread building a Huffman code stalls its whole CTA, so neither appears here:
//   k_zenc_hist    one CTA per block: every thread counts its own 128 bytes into a PRIVATE column of byte
//                  counters (256 bins x 256 threads = 64 KB, bank = lane: conflict-free LDS/ADD/STS), then thread b
//                  sums row b with dp4a -> 256 x u16 counts per block in HBM
//   k_zenc_tables  one THREAD per block (a hundred thousand of them at once hide each other's latency): rank the
//                  symbols, two-queue Huffman, limit to 11 bits, canonical codes, tree description
//   k_zenc_encode  one CTA per block: exact stream sizes, then the bits are assembled in shared memory — words a
//                  thread covers completely are plain stores, only its first / last partial word is an atomicOr —
//                  and the finished block goes to its slot with 128-bit stores
static const u32 ZET = 256, ZEPER = ZBS / ZET;                    // threads per CTA, bytes per thread in a full block
static const u32 ZHIST_SMEM = 65536;
static const u32 ZENC_SMEM = ZBS + 1024;


// A full block's thread range is 128 aligned bytes: read with 128-bit loads; ragged tail blocks go byte by byte.
template <class F> __device__ __forceinline__ void zenc_each(bool full, const u8 *src, u32 t0, u32 t1, F f)
{
    if (full) {
        const uint4 *v = (const uint4 *)(src + t0);
#pragma unroll 2
        for (int i = 0; i < (int)(ZEPER / 16); i++) {
            const uint4 x = __ldg(v + i);
            const u32 w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
            for (int k = 0; k < 4; k++) { f(w[k] & 0xFF); f((w[k] >> 8) & 0xFF); f((w[k] >> 16) & 0xFF); f(w[k] >> 24); }
        }
    } else for (u32 i = t0; i < t1; i++) f((u32)src[i]);
}
// backwards, two symbols per call (the bit assembly checks for a full word once per pair); an odd one goes to f1
template <class F2, class F1> __device__ __forceinline__ void zenc_each_rev2(bool full, const u8 *src, u32 t0, u32 t1, F2 f2, F1 f1)
{
    if (full) {
        const uint4 *v = (const uint4 *)(src + t0);
#pragma unroll 2
        for (int i = (int)(ZEPER / 16) - 1; i >= 0; i--) {
            const uint4 x = __ldg(v + i);
            const u32 w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
            for (int k = 3; k >= 0; k--) { f2(w[k] >> 24, (w[k] >> 16) & 0xFF); f2((w[k] >> 8) & 0xFF, w[k] & 0xFF); }
        }
    } else {
        u32 i = t1;
        for (; i >= t0 + 2; i -= 2) f2((u32)src[i - 1], (u32)src[i - 2]);
        if (i > t0) f1((u32)src[i - 1]);
    }
}

__global__ void __launch_bounds__(256, 3) k_zenc_hist(const ZEncArgs A, u16 *hists)
{
    extern __shared__ __align__(16) u8 R[];
    const ZEncBlock &B = A.blk[blockIdx.x];
    if (B.lz) return;
    const u8 *src = B.src; const u32 n = B.n;
    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    {
        uint4 *z = (uint4 *)R;
#pragma unroll
        for (int i = 0; i < 16; i++) z[tid + 256 * i] = make_uint4(0, 0, 0, 0);
    }
    __syncthreads();
    {
        const bool full = n == ZBS && (((uintptr_t)src) & 15) == 0;
        const u32 per = (n + ZET - 1) / ZET;                      // <= 128: a byte counter cannot overflow
        u32 t0 = tid * per, t1 = t0 + per; if (t0 > n) t0 = n; if (t1 > n) t1 = n;
        u8 *col = R + lane * 4 + (warp >> 2) * 128 + (warp & 3);  // word = lane (+32), byte = warp & 3: bank == lane
        zenc_each(full, src, t0, t1, [&](u32 c) { col[c << 8]++; });
    }
    __syncthreads();
    u32 h = 0;
    const u32 *row = (const u32 *)(R + tid * 256);
#pragma unroll 8
    for (u32 j = 0; j < 64; j++) h = __dp4a(row[(j + tid) & 63], 0x01010101u, h);
    hists[(size_t)blockIdx.x * 256 + tid] = (u16)h;               // <= 32768
}

// one thread per block
__global__ void __launch_bounds__(64) k_zenc_tables(const ZEncArgs A, u32 nblocks, const u16 *hists, ZEncMeta *metas)
{
    const u32 b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblocks) return;
    if (A.blk[b].lz) return;
    ZEncMeta &M = metas[b];
    if (A.blk[b].n == 0) { M.mode = 0; M.tree_len = 0; M.rle_sym = 0; return; }
    nafz::zenc_huf_build(hists + (size_t)b * 256, M);
}

__global__ void __launch_bounds__(256) k_zenc_encode(const ZEncArgs A, const ZEncMeta *metas)
{
    extern __shared__ __align__(16) u8 R[];   // the block being assembled
    __shared__ u32 ctab[256];                  // code | len << 16
    __shared__ u32 pre_bits[257];
    __shared__ u64 smscan[33];
    __shared__ u32 s_mode, s_lit_hdr, s_payload0;
    __shared__ u32 s_stream_bytes[4];

    ZEncBlock &B = A.blk[blockIdx.x];
    if (B.lz) return;                          // k_zenc_lz's block
    const ZEncMeta &M = metas[blockIdx.x];
    const u8 *src = B.src; const u32 n = B.n;
    u8 *slot = A.slots + B.slot_off;
    const u32 tid = threadIdx.x;

    // (blocks of at most 512 bytes stay raw: nothing to gain, and a decoder can read a tiny stream -- the lengths of a genome's
    // two dozen records -- on the host without entropy decoding, naf_dec.cu)
    if (n == 0 || M.mode == 0 || (n <= 512 && M.mode != 1)) { if (tid == 0) { B.type = 0; B.csize = n; } return; }
    if (M.mode == 1) { if (tid == 0) { slot[0] = M.rle_sym; B.type = 1; B.csize = 1; } return; }
    { const u32 e = M.ctab[tid]; ctab[tid] = (e & 0xFFF) | ((e >> 12) << 16); }
    if (tid == 0) s_mode = 2;
    const u32 tree_len = M.tree_len;

    // my symbols: stream k = [k*seg, ...), split evenly over the stream's threads (same split in both passes)
    const u32 nstreams = n <= 1023 ? 1 : 4, tps = ZET / nstreams;
    const u32 seg = nstreams == 4 ? (n + 3) / 4 : n;
    const u32 k = tid / tps, q = tid % tps;
    const u32 sbeg = k * seg, send = (k == nstreams - 1) ? n : (sbeg + seg < n ? sbeg + seg : n);
    const u32 slen = send > sbeg ? send - sbeg : 0;
    const u32 per = (slen + tps - 1) / tps;
    u32 t0 = sbeg + q * per, t1 = t0 + per; if (t0 > send) t0 = send; if (t1 > send) t1 = send;
    const bool full = n == ZBS && (((uintptr_t)src) & 15) == 0;    // then t0 = 128 * tid, t1 = t0 + 128
    __syncthreads();
    // ---- exact size of every stream
    u32 bits = 0;
    zenc_each(full, src, t0, t1, [&](u32 c) { bits += ctab[c] >> 16; });
    u64 total_bits;
    const u64 pre = block_excl_scan(bits, &total_bits, smscan);
    pre_bits[tid] = (u32)pre; if (tid == 255) pre_bits[256] = (u32)total_bits;
    __syncthreads();
    if (tid < nstreams) { const u32 b = pre_bits[(tid + 1) * tps] - pre_bits[tid * tps]; s_stream_bytes[tid] = b / 8 + 1; }
    __syncthreads();
    if (tid == 0) {
        u32 payload = tree_len + (nstreams == 4 ? 6 : 0);
        for (u32 j = 0; j < nstreams; j++) payload += s_stream_bytes[j];
        const u32 lh = nstreams == 1 ? 3 : ((n <= 16383 && payload <= 16383) ? 4 : 5);
        if (nstreams == 1 && payload > 1023) s_mode = 0;
        if (lh + payload + 1 >= n) s_mode = 0;                // no gain: raw block
        if (n < 4 * 4 && nstreams == 4) s_mode = 0;
        s_lit_hdr = lh; s_payload0 = payload;
    }
    __syncthreads();
    if (s_mode == 0) { if (tid == 0) { B.type = 0; B.csize = n; } return; }
    const u32 lit_hdr = s_lit_hdr, payload = s_payload0;
    const u32 content = lit_hdr + payload + 1;
    // ---- assemble the block in shared memory
    u32 *outw = (u32 *)R;
    for (u32 i = tid; i < content / 4 + 2; i += ZET) outw[i] = 0;
    __syncthreads();
    if (tid == 0) {
        // Literals_Section_Header: type 2 (compressed), size format by stream count / sizes
        if (nstreams == 1) { u32 v = 2 | (0 << 2) | (n << 4) | (payload << 14); R[0] = (u8)v; R[1] = (u8)(v >> 8); R[2] = (u8)(v >> 16); }
        else if (lit_hdr == 4) { u32 v = 2 | (2 << 2) | (n << 4) | (payload << 18); R[0] = (u8)v; R[1] = (u8)(v >> 8); R[2] = (u8)(v >> 16); R[3] = (u8)(v >> 24); }
        else { u64 v = 2 | (3 << 2) | ((u64)n << 4) | ((u64)payload << 22); for (int i = 0; i < 5; i++) R[i] = (u8)(v >> (8 * i)); }
        if (nstreams == 4) {
            u8 *jt = R + lit_hdr + tree_len;
            for (int j = 0; j < 3; j++) { jt[2 * j] = (u8)s_stream_bytes[j]; jt[2 * j + 1] = (u8)(s_stream_bytes[j] >> 8); }
        }
        R[content - 1] = 0;                                   // Sequences_Section_Header: 0 sequences
    }
    if (tid < tree_len) R[lit_hdr + tid] = M.tree[tid];
    __syncthreads();
    {
        u32 byte_base = lit_hdr + tree_len + (nstreams == 4 ? 6 : 0);
        for (u32 j = 0; j < k; j++) byte_base += s_stream_bytes[j];
        // symbols later in the stream sit at lower bit positions: my first bit = bits of all threads after me in my stream
        const u64 ab = (u64)byte_base * 8 + (pre_bits[(k + 1) * tps] - (pre_bits[tid] + bits));     // absolute bit position of my first bit
        // The accumulator is kept word-aligned: it starts with ab % 32 zero bits below my first code, so every flush is one
        // whole word -- an atomicOr for my first word (its low bits belong to the thread after me), plain stores after that.
        // A full word is looked for once per two symbols (codes are at most 11 bits: 31 + 22 < 64), at a point all lanes reach
        // together, instead of after every symbol at whichever symbol a lane happens to fill up.
        u32 wi = (u32)(ab >> 5), fill = (u32)(ab & 31);
        u64 acc = 0;
        bool first = true;
        auto flush32 = [&]() {
            if (first) { atomicOr(&outw[wi], (u32)acc); first = false; } else outw[wi] = (u32)acc;
            acc >>= 32; fill -= 32; wi++;
        };
        zenc_each_rev2(full, src, t0, t1,
            [&](u32 c1, u32 c0) {
                const u32 e1 = ctab[c1], e0 = ctab[c0];
                acc |= (u64)(e1 & 0xFFFF) << fill; fill += e1 >> 16;
                acc |= (u64)(e0 & 0xFFFF) << fill; fill += e0 >> 16;
                if (fill >= 32) flush32();
            },
            [&](u32 c) { const u32 e = ctab[c]; acc |= (u64)(e & 0xFFFF) << fill; fill += e >> 16; if (fill >= 32) flush32(); });
        if (q == 0) { acc |= 1ull << fill; fill++; }          // end mark right above the first symbol's code
        while (fill >= 32) flush32();
        if (fill) atomicOr(&outw[wi], (u32)acc);              // my last, partial word: its high bits belong to the thread before me
    }
    __syncthreads();
    {
        uint4 *d = (uint4 *)slot; const uint4 *sv = (const uint4 *)R;
        for (u32 i = tid; i < (content + 15) / 16; i += ZET) d[i] = sv[i];
        if (tid == 0) { B.type = 2; B.csize = content; }
    }
}

// lay blocks out as frames: [magic][FHD][WD] then per block a 3-byte header + content
struct ZGatherArgs { const ZEncBlock *blk; const u8 *slots; const u64 *off; u8 *const *dest; const u32 *first_block; const int *wlog; int skip_magic; };

__global__ void __launch_bounds__(256) k_zenc_gather(const ZGatherArgs A)
{
    const ZEncBlock &B = A.blk[blockIdx.x];
    const u32 s = B.stream, fb = A.first_block[s];
    u8 *frame = A.dest[s];
    u8 *dst = frame + 6 + (A.off[blockIdx.x] - A.off[fb]);
    if (blockIdx.x == fb && threadIdx.x == 0) {
        if (!A.skip_magic) { frame[0] = 0x28; frame[1] = 0xB5; frame[2] = 0x2F; frame[3] = 0xFD; }
        frame[4] = 0x00;                                       // FHD: no content size, no checksum, no dictionary
        // window descriptor: ennaf --long N declares 2^N for the sequence stream (compressor.c:12-16); never below our own
        // block span, never above what unnaf accepts (ZSTD_d_windowLogMax 31, input.c:270)
        int wl = A.wlog[s];
        wl = wl < ZWINDOW_LOG ? ZWINDOW_LOG : (wl > 31 ? 31 : wl);
        frame[5] = (u8)((wl - 10) << 3);
    }
    if (threadIdx.x == 0) {
        u32 size_field = B.type == 1 ? B.n : B.csize;          // RLE: regenerated size
        u32 bh = (B.last & 1) | (B.type << 1) | (size_field << 3);
        dst[0] = (u8)bh; dst[1] = (u8)(bh >> 8); dst[2] = (u8)(bh >> 16);
    }
    // content: 128-bit stores aligned to the destination, each fed by one unaligned 16-byte read (load_bytes16, naf_dec.cu)
    const u8 *from = B.type == 0 ? B.src : A.slots + B.slot_off;
    u8 *d0 = dst + 3; const u32 nbytes = B.csize;
    const u32 head = min(nbytes, (u32)((16 - ((uintptr_t)d0 & 15)) & 15));
    if (threadIdx.x < head) d0[threadIdx.x] = from[threadIdx.x];
    const u32 nv = (nbytes - head) / 16, done = head + nv * 16;
    for (u32 i = threadIdx.x; i < nv; i += 256) *(uint4 *)(d0 + head + 16 * i) = load_bytes16(from + head + 16 * i);
    if (threadIdx.x < nbytes - done) d0[done + threadIdx.x] = from[done + threadIdx.x];
}

// ---- LZ streams: one thread per 8 KB block runs nafz::zlz_encode_block (match finder, literal Huffman, FSE-coded sequences).
// The 32 hash tables of a CTA live in shared memory, interleaved so that entry e of lane l sits in bank l; the rest of a
// block's scratch (literals, sequence arrays, symbol codes, FSE state tables) is a private slice of one HBM workspace.
static const u32 ZLZ_WORK_SPOS = 1280 * 2, ZLZ_WORK_TSYM = 512;
static const u32 ZLZ_SMEM = 32 * (1u << nafz::ZLZ_HLOG) * 2;
struct ZLzArgs { ZEncBlock *blk; u8 *slots; u8 *work; u32 first[9]; u32 lzfirst[9]; u32 ns; u32 nlz; u32 zlb; };
HD u32 zlz_work_bytes(u32 zlb) { return ((zlb + 64) + 3 * (zlb / 4) * 2 + ZLZ_WORK_SPOS + ZLZ_WORK_TSYM + 3 * (zlb / 4) + 15) & ~15u; }

__global__ void __launch_bounds__(32) k_zenc_lz(const ZLzArgs A)
{
    extern __shared__ __align__(16) u16 htabs[];               // 32 interleaved tables of 1 << ZLZ_HLOG entries
    const u32 j = blockIdx.x * 32 + threadIdx.x;               // j-th LZ block of the batch
    if (j >= A.nlz) return;
    u32 s = 0;
    while (s + 1 < A.ns && j >= A.lzfirst[s + 1]) s++;
    ZEncBlock &B = A.blk[A.first[s] + (j - A.lzfirst[s])];
    const u32 zlb = A.zlb, maxseq = zlb / 4;
    u8 *w = A.work + (size_t)j * zlz_work_bytes(zlb);
    u8 *lit = w; w += zlb + 64;
    nafz::ZLzSeqs S; S.ll = (u16 *)w; w += maxseq * 2; S.ml = (u16 *)w; w += maxseq * 2; S.ov = (u16 *)w; w += maxseq * 2; S.n = 0;
    nafz::ZLzWork W; W.spos = (u16 *)w; w += ZLZ_WORK_SPOS; W.tsym = w; w += ZLZ_WORK_TSYM; W.codes = w;
    u8 *slot = A.slots + B.slot_off;
    bool rle = false;
    const u32 cs = nafz::zlz_encode_block(B.src, B.n, true, htabs + threadIdx.x, 32, lit, S, maxseq, W, slot, zlb + 512, &rle);
    if (cs) { B.type = 2; B.csize = cs; }
    else if (rle) { slot[0] = B.src[0]; B.type = 1; B.csize = 1; }
    else { B.type = 0; B.csize = B.n; }
}

// ---- LZ streams, data-parallel (level >= 2; bodies: zstd_lzc_hd.cuh, emulated on the CPU by tests/emu/emu_zlzc.cpp):
//   k_zlc_find    one CTA per block, thread = 32-byte chunk: column match finder (maps, neighbour walks over per-chunk summaries,
//                 one block scan), sequences + compacted literals to the block's scratch; every 8th block of a stream also counts
//                 its literal bytes and sequence codes into the stream's statistics
//   k_zlc_define  one thread per stream: Huffman code + three FSE tables from the statistics; which block will carry them
//   k_zlc_finish  one thread per block: serial coding against the stream's tables (the defining block: Compressed_Literals +
//                 FSE_Compressed; the ones behind it: Treeless_Literals + Repeat_Mode)
//   k_zlc_finish_own  streams that got no tables (no sequences at all, one literal symbol ...): every block builds its own
// NAFGPU_LZ=1 brings back the first formulation (k_zenc_lz: one thread per block does everything, private tables per block) for A/B
// measurements; the switch is read per call, so one process can compare the two.
static bool zlc_mode() { const char *e = getenv("NAFGPU_LZ"); return !(e && e[0] == '1'); }
// NAFGPU_LZ=b: the finder's bit-mask formulation (same frames; written after the round's last GPU call, so a level does not select it yet)
static bool zlc_bits() { const char *e = getenv("NAFGPU_LZ"); return e && e[0] == 'b'; }
struct ZlcArgs {
    ZEncBlock *blk; nafz::ZlcBlk *info; u8 *slots; u8 *work; u32 *counts; nafz::ZlcTables *tables; u32 *def_fail;
    const u8 *src[8]; u64 n[8]; u64 slot_base[8];
    u32 first[9]; u32 lzfirst[9]; u32 is_lz[8]; u32 ns; u32 nlz; u32 zlb;
};
__device__ __forceinline__ u32 zlc_stream_of(const ZlcArgs &A, u32 j) { u32 s = 0; while (s + 1 < A.ns && j >= A.lzfirst[s + 1]) s++; return s; }
__device__ __forceinline__ nafz::ZlcStreamView zlc_view(const ZlcArgs &A, u32 s)
{
    nafz::ZlcStreamView V;
    const u32 ws = nafz::zlc_work_bytes(A.zlb);
    V.src = A.src[s]; V.n = A.n[s]; V.bs = A.zlb; V.nblk = A.first[s + 1] - A.first[s];
    V.info = A.info + A.lzfirst[s]; V.work = A.work + (size_t)A.lzfirst[s] * ws; V.work_stride = ws;
    V.slots = A.slots + A.slot_base[s]; V.slot_stride = A.zlb + 512;
    return V;
}

// The finder kernel, once per formulation of its phases (FN: nafz::zlcb = byte loops, nafz = bit masks; same barriers, same outputs)
#define ZLC_FIND_KERNEL(NAME, FN) \
__global__ void __launch_bounds__(256) NAME(const ZlcArgs A) \
{ \
    extern __shared__ __align__(16) u8 zlc_smem[]; \
    FN::ZlcSh &sh = *reinterpret_cast<FN::ZlcSh *>(zlc_smem); \
    __shared__ u64 smscan[33]; \
    const u32 j = blockIdx.x, k = threadIdx.x; \
    const u32 s = zlc_stream_of(A, j), bidx = j - A.lzfirst[s]; \
    const ZEncBlock &B = A.blk[A.first[s] + bidx]; \
    const u8 *src = B.src; const u32 n = B.n; \
    for (u32 i = k; i < n; i += 256) sh.src_[FN::zlc_ix(i)] = src[i]; \
    for (u32 i = k; i < nafz::ZLC_NBINS; i += 256) sh.hist[i] = 0; \
    if (k == 0) { sh.n = n; sh.nch = (n + nafz::ZLC_CH - 1) / nafz::ZLC_CH; sh.rle_break = 0; sh.lastend = 0; } \
    __syncthreads(); \
    const bool act = k < sh.nch; \
    if (act) FN::zlc_zeros(sh, k); \
    __syncthreads(); \
    nafz::ZlcBlk &I = A.info[j]; \
    if (n == 0 || !sh.rle_break || n < 16) {                   /* empty, one repeated byte, or too small to parse: RLE / raw block */ \
        if (k == 0) { I.nseq = 0; I.nlit = 0; I.parsed = 0; I.rle = (n && !sh.rle_break) ? 1 : 0; I.conv = 0; I.pad = 0; } \
        return; \
    } \
    if (act) FN::zlc_columns(sh, k); \
    __syncthreads(); \
    if (act) FN::zlc_breaks(sh, k); \
    __syncthreads(); \
    if (act) FN::zlc_choose(sh, k); \
    __syncthreads(); \
    if (act) FN::zlc_breaks_d(sh, k); \
    __syncthreads(); \
    if (act) FN::zlc_count(sh, k); \
    __syncthreads(); \
    { \
        const u64 v = act ? ((u64)sh.cnt[k] | ((u64)sh.mls[k] << 32)) : 0; \
        u64 total; const u64 pre = block_excl_scan(v, &total, smscan); \
        if (act) { sh.ibase[k] = (u16)pre; sh.mbase[k] = (u16)(pre >> 32); } \
        if (k == 0) { sh.nseq = (u32)total; sh.mltot = (u32)(total >> 32); } \
    } \
    __syncthreads(); \
    const bool sampled = bidx % nafz::ZLC_SAMPLE == 0; \
    nafz::ZlcWork K = nafz::zlc_work(A.work + (size_t)j * nafz::zlc_work_bytes(A.zlb), A.zlb); \
    if (act) FN::zlc_emit_seqs(sh, k, K.S, K.lit, sampled); \
    FN::zlc_emit_tail(sh, k, 256, K.lit, sampled); \
    if (k == 0) { I.nseq = sh.nseq; I.nlit = n - sh.mltot; I.parsed = 1; I.rle = 0; I.conv = 0; I.pad = 0; } \
    if (!sampled) return; \
    __syncthreads(); \
    if (k == 0) FN::zlc_count_offsets(sh); \
    __syncthreads(); \
    u32 *acc = A.counts + s * nafz::ZLC_NBINS; \
    for (u32 i = k; i < nafz::ZLC_NBINS; i += 256) if (sh.hist[i]) atomicAdd(&acc[i], sh.hist[i]); \
}
ZLC_FIND_KERNEL(k_zlc_find, nafz::zlcb)
ZLC_FIND_KERNEL(k_zlc_find_bits, nafz)

__global__ void __launch_bounds__(32) k_zlc_define(const ZlcArgs A)
{
    const u32 s = blockIdx.x;
    if (threadIdx.x || s >= A.ns || !A.is_lz[s]) return;
    const nafz::ZlcStreamView V = zlc_view(A, s);
    nafz::zlc_define(V, A.counts + s * nafz::ZLC_NBINS, A.tables[s]);
}

__global__ void __launch_bounds__(64) k_zlc_finish(const ZlcArgs A)
{
    // every symbol coded is two or three dependent table reads: the tables of the stream this CTA's first block belongs to are
    // staged in shared memory (a CTA that straddles two streams reads the second one's from HBM / L1)
    __shared__ __align__(16) nafz::ZlcTables Ts;
    const u32 j0 = blockIdx.x * blockDim.x, j = j0 + threadIdx.x;
    const u32 s0 = zlc_stream_of(A, j0);
    {
        const uint4 *g = (const uint4 *)(A.tables + s0); uint4 *d = (uint4 *)&Ts;
        for (u32 i = threadIdx.x; i < sizeof(nafz::ZlcTables) / 16; i += blockDim.x) d[i] = g[i];
    }
    __syncthreads();
    if (j >= A.nlz) return;
    const u32 s = zlc_stream_of(A, j), bidx = j - A.lzfirst[s];
    const nafz::ZlcStreamView V = zlc_view(A, s);
    ZEncBlock &B = A.blk[A.first[s] + bidx];
    u32 type = 0, csize = 0;
    if (nafz::zlc_finish_block(V, bidx, s == s0 ? Ts : A.tables[s], A.def_fail + s, &type, &csize)) { B.type = type; B.csize = csize; }
}
// streams without tables of their own (no sequences at all, one literal symbol ...): every block builds its own, like k_zenc_lz
__global__ void __launch_bounds__(64) k_zlc_finish_own(const ZlcArgs A)
{
    const u32 j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= A.nlz) return;
    const u32 s = zlc_stream_of(A, j), bidx = j - A.lzfirst[s];
    const u32 def_fail = A.def_fail[s];
    if (A.tables[s].ok && !def_fail) return;
    const nafz::ZlcStreamView V = zlc_view(A, s);
    ZEncBlock &B = A.blk[A.first[s] + bidx];
    u32 type = 0, csize = 0;
    if (nafz::zlc_finish_block_own(V, bidx, A.tables[s], def_fail, &type, &csize)) { B.type = type; B.csize = csize; }
}

// ---- host-buffer encode: the big streams compressed behind the upload (Ctx::EarlyZ; called from split_streams_fused)
static void zenc_early_begin(Ctx &ctx, CudaExec &ex, const u8 *seq, u64 seq_max_bytes, const u8 *qual, u64 qual_max_bytes)
{
    Ctx::EarlyZ &E = ctx.early;
    E = Ctx::EarlyZ();
    const u8 *src[2] = {seq, qual}; const u64 mx[2] = {seq_max_bytes, qual_max_bytes};
    for (int k = 0; k < 2; k++) {
        E.src[k] = src[k]; E.max_blocks[k] = mx[k] / ZBS;
        if (!E.max_blocks[k]) continue;
        E.blk[k] = ex.alloc<ZEncBlock>(E.max_blocks[k]);
        E.slots[k] = ex.alloc<u8>(E.max_blocks[k] * ZSLOT + 64);
    }
    static bool attr_done[64] = {};
    int dev = 0; cudaGetDevice(&dev);
    if (dev < 64 && !attr_done[dev]) {
        CUDA_TRY(cudaFuncSetAttribute(k_zenc_hist, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ZHIST_SMEM));
        CUDA_TRY(cudaFuncSetAttribute(k_zenc_encode, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ZENC_SMEM));
        attr_done[dev] = true;
    }
    E.on = true;
    // the side stream must not start before what the compute stream did to the arena so far (memsets of the look-back records)
    CUDA_TRY(cudaEventRecord(ctx.side_fork, ex.stream));
    CUDA_TRY(cudaStreamWaitEvent(ctx.side, ctx.side_fork, 0));
}
// the first seq_bytes / qual_bytes of the two streams are final: compress the blocks that have become complete.  The caller has
// synchronised with the kernels that wrote them (event), so these launches need no further ordering.
static void zenc_early_step(Ctx &ctx, u64 seq_bytes, u64 qual_bytes)
{
    Ctx::EarlyZ &E = ctx.early;
    if (!E.on) return;
    const u64 avail[2] = {seq_bytes, qual_bytes};
    for (int k = 0; k < 2; k++) {
        u64 upto = avail[k] / ZBS; if (upto > E.max_blocks[k]) upto = E.max_blocks[k];
        if (upto <= E.done[k]) continue;
        const u64 d0 = E.done[k], cnt = upto - d0;
        ZEncBlock *blk = (ZEncBlock *)E.blk[k] + d0;
        const u8 *src = E.src[k] + d0 * ZBS;
        k_for_each<<<(unsigned)((cnt + 255) / 256), 256, 0, ctx.side>>>((size_t)cnt, [=] __device__ (size_t i) {
            ZEncBlock e; e.src = src + i * ZBS; e.n = ZBS; e.stream = 0; e.last = 0; e.type = 0; e.csize = 0; e.lz = 0; e.slot_off = (d0 + i) * ZSLOT;
            blk[i] = e;
        });
        // scratch (histograms, codes) lives until the side stream has run these three kernels: taken from the arena, which is
        // only reset by the next call (after the side stream has drained)
        u16 *hists = (u16 *)ctx.arena.alloc_bytes((size_t)cnt * 512);
        ZEncMeta *metas = (ZEncMeta *)ctx.arena.alloc_bytes((size_t)cnt * sizeof(ZEncMeta));
        ZEncArgs A{blk, E.slots[k]};
        k_zenc_hist<<<(unsigned)cnt, 256, ZHIST_SMEM, ctx.side>>>(A, hists);
        k_zenc_tables<<<(unsigned)((cnt + 63) / 64), 64, 0, ctx.side>>>(A, (u32)cnt, hists, metas);
        k_zenc_encode<<<(unsigned)cnt, 256, ZENC_SMEM, ctx.side>>>(A, metas);
        CUDA_TRY(cudaGetLastError());
        E.done[k] = upto;
    }
}
static void zenc_early_abort(Ctx &ctx)
{
    if (ctx.early.on && ctx.side) cudaStreamSynchronize(ctx.side);
    ctx.early = Ctx::EarlyZ();
}

static void zstd_compress_batch(Ctx &ctx, CudaExec &ex, ZEncBatch &b)
{
    const size_t ns = b.src.size();
    if (ns > 8) fail(NAFGPU_E_ARG, "too many streams in one compression batch\n");
    // block list: stream s contributes ceil(n / block size) blocks (an empty stream: one empty raw last block); built on the device
    ZEncStreamTab tab; memset(&tab, 0, sizeof tab);
    ZLzArgs L; memset(&L, 0, sizeof L);
    b.first_block.assign(ns + 1, 0);
    u64 slot_total = 0; u32 nlz = 0; bool side = false, any_early = false;
    const u32 ZLB = zlb_bytes(), ZLSLOT = ZLB + 512;
    L.zlb = ZLB;
    for (size_t s = 0; s < ns; s++) {
        const u32 bs = b.lz[s] ? ZLB : ZBS, slot = b.lz[s] ? ZLSLOT : ZSLOT;
        const u32 nb = (u32)(b.n[s] ? (b.n[s] + bs - 1) / bs : 1);
        b.first_block[s + 1] = b.first_block[s] + nb;
        tab.src[s] = b.src[s]; tab.n[s] = b.n[s]; tab.first[s] = b.first_block[s]; tab.bs[s] = bs; tab.lz[s] = (u32)b.lz[s];
        tab.slot_base[s] = slot_total; slot_total += (u64)nb * slot;
        for (int k = 0; k < 2; k++)
            if (ctx.early.on && !b.lz[s] && ctx.early.done[k] && b.src[s] == ctx.early.src[k]) {
                tab.early[s] = (const ZEncBlock *)ctx.early.blk[k]; tab.early_slots[s] = ctx.early.slots[k];
                tab.early_done[s] = (u32)(ctx.early.done[k] < nb ? ctx.early.done[k] : nb); any_early = true;
            }
        L.first[s] = b.first_block[s]; L.lzfirst[s] = nlz; if (b.lz[s]) nlz += nb;
    }
    b.nblocks = b.first_block[ns];
    tab.first[ns] = b.nblocks; tab.ns = (u32)ns;
    L.first[ns] = b.nblocks; L.lzfirst[ns] = nlz; L.ns = (u32)ns; L.nlz = nlz;
    b.d_blocks = ex.alloc<ZEncBlock>(b.nblocks);
    b.d_slots = ex.alloc<u8>(slot_total + 64);
    if (any_early) {                                           // what the side stream compressed behind the upload is part of this batch
        CUDA_TRY(cudaEventRecord(ctx.side_join, ctx.side));
        CUDA_TRY(cudaStreamWaitEvent(ex.stream, ctx.side_join, 0));
    }
    {
        ZEncBlock *blk = b.d_blocks;
        const u32 fin = b.final_shard ? 1u : 0u;
        const u8 *slots0 = b.d_slots;
        ex.for_each(b.nblocks, [=] __device__ (size_t i) {
            u32 s = 0;
            while (s + 1 < tab.ns && (u32)i >= tab.first[s + 1]) s++;
            const u32 bs = tab.bs[s];
            const u64 k = i - tab.first[s], off = k * bs, left = tab.n[s] - (tab.n[s] < off ? tab.n[s] : off);
            ZEncBlock e; e.src = tab.src[s] + off; e.n = (u32)(left < bs ? left : bs); e.stream = s; e.last = ((u32)i + 1 == tab.first[s + 1]) & fin;
            e.type = 0; e.csize = 0; e.lz = tab.lz[s]; e.slot_off = tab.slot_base[s] + k * (tab.lz[s] ? ZLSLOT : ZSLOT);
            if (k < tab.early_done[s]) {                          // done already: take its result, its slot is in the early pool
                const ZEncBlock d = tab.early[s][k];
                e.type = d.type; e.csize = d.csize; e.lz = 2; e.slot_off = (u64)((tab.early_slots[s] + d.slot_off) - slots0);
            }
            blk[i] = e;
        }, "zenc_init_blocks");
    }
    if (nlz && zlc_mode()) {
        ZlcArgs Z; memset(&Z, 0, sizeof Z);
        Z.blk = b.d_blocks; Z.slots = b.d_slots; Z.ns = (u32)ns; Z.nlz = nlz; Z.zlb = ZLB;
        for (size_t s = 0; s < ns; s++) { Z.src[s] = b.src[s]; Z.n[s] = b.n[s]; Z.slot_base[s] = tab.slot_base[s]; Z.is_lz[s] = (u32)b.lz[s]; }
        for (size_t s = 0; s <= ns; s++) { Z.first[s] = L.first[s]; Z.lzfirst[s] = L.lzfirst[s]; }
        Z.info = ex.alloc<nafz::ZlcBlk>(nlz);
        Z.work = ex.alloc<u8>((size_t)nlz * nafz::zlc_work_bytes(ZLB));
        Z.counts = ex.alloc<u32>(8 * nafz::ZLC_NBINS + 8); ex.zero(Z.counts, (8 * nafz::ZLC_NBINS + 8) * 4);
        Z.def_fail = Z.counts + 8 * nafz::ZLC_NBINS;
        static_assert(sizeof(nafz::ZlcTables) % 16 == 0, "k_zlc_finish copies the tables in 16-byte pieces");
        Z.tables = (nafz::ZlcTables *)ex.alloc<uint4>(8 * sizeof(nafz::ZlcTables) / 16);
        const bool bits = zlc_bits();
        const size_t fsm = bits ? sizeof(nafz::ZlcSh) : sizeof(nafz::zlcb::ZlcSh);
        CUDA_TRY(cudaFuncSetAttribute(k_zlc_find, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(nafz::zlcb::ZlcSh)));
        CUDA_TRY(cudaFuncSetAttribute(k_zlc_find_bits, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(nafz::ZlcSh)));
        static const bool env_side = !(getenv("NAFGPU_SIDE") && getenv("NAFGPU_SIDE")[0] == '0');
        side = env_side && ctx.side && !(ex.prof && ex.prof->on);
        if (side) {
            CUDA_TRY(cudaEventRecord(ctx.side_fork, ex.stream));
            CUDA_TRY(cudaStreamWaitEvent(ctx.side, ctx.side_fork, 0));
            if (bits) k_zlc_find_bits<<<nlz, 256, fsm, ctx.side>>>(Z); else k_zlc_find<<<nlz, 256, fsm, ctx.side>>>(Z);
            k_zlc_define<<<(unsigned)ns, 32, 0, ctx.side>>>(Z);
            k_zlc_finish<<<(nlz + 63) / 64, 64, 0, ctx.side>>>(Z);
            k_zlc_finish_own<<<(nlz + 63) / 64, 64, 0, ctx.side>>>(Z);
            CUDA_TRY(cudaEventRecord(ctx.side_join, ctx.side));
            ex.launches += 4;
        } else {
            if (bits) KLAUNCH(ex, "k_zlc_find_bits", k_zlc_find_bits<<<nlz, 256, fsm, ex.stream>>>(Z));
            else KLAUNCH(ex, "k_zlc_find", k_zlc_find<<<nlz, 256, fsm, ex.stream>>>(Z));
            KLAUNCH(ex, "k_zlc_define", k_zlc_define<<<(unsigned)ns, 32, 0, ex.stream>>>(Z));
            KLAUNCH(ex, "k_zlc_finish", k_zlc_finish<<<(nlz + 63) / 64, 64, 0, ex.stream>>>(Z));
            KLAUNCH(ex, "k_zlc_finish_own", k_zlc_finish_own<<<(nlz + 63) / 64, 64, 0, ex.stream>>>(Z));
        }
        CUDA_TRY(cudaGetLastError());
    } else if (nlz) {
        L.blk = b.d_blocks; L.slots = b.d_slots; L.work = ex.alloc<u8>((size_t)nlz * zlz_work_bytes(ZLB));
        CUDA_TRY(cudaFuncSetAttribute(k_zenc_lz, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ZLZ_SMEM));
        // next to the Huffman kernels of the big streams, not in front of them (when profiling: in line, so that its time is its own)
        static const bool env_side = !(getenv("NAFGPU_SIDE") && getenv("NAFGPU_SIDE")[0] == '0');
        side = env_side && ctx.side && !(ex.prof && ex.prof->on);
        if (side) {
            CUDA_TRY(cudaEventRecord(ctx.side_fork, ex.stream));
            CUDA_TRY(cudaStreamWaitEvent(ctx.side, ctx.side_fork, 0));
            k_zenc_lz<<<(nlz + 31) / 32, 32, ZLZ_SMEM, ctx.side>>>(L);
            CUDA_TRY(cudaEventRecord(ctx.side_join, ctx.side));
            ex.launches++;
        } else KLAUNCH(ex, "k_zenc_lz", k_zenc_lz<<<(nlz + 31) / 32, 32, ZLZ_SMEM, ex.stream>>>(L));
    }
    // the Huffman kernels skip LZ blocks; when the LZ streams come first (they do in a .naf) they are not even launched for them
    u32 base = 0;
    { size_t s = 0; while (s < ns && b.lz[s]) s++; bool tail_plain = true; for (size_t t = s; t < ns; t++) if (b.lz[t]) tail_plain = false; if (tail_plain) base = b.first_block[s]; }
    const u32 nhuf = b.nblocks - base;
    ZEncArgs A{b.d_blocks + base, b.d_slots};
    u16 *d_hists = ex.alloc<u16>((size_t)nhuf * 256);
    ZEncMeta *d_metas = ex.alloc<ZEncMeta>(nhuf);
    CUDA_TRY(cudaFuncSetAttribute(k_zenc_hist, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ZHIST_SMEM));
    CUDA_TRY(cudaFuncSetAttribute(k_zenc_encode, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ZENC_SMEM));
    if (nhuf) {
        KLAUNCH(ex, "k_zenc_hist", k_zenc_hist<<<nhuf, 256, ZHIST_SMEM, ex.stream>>>(A, d_hists));
        KLAUNCH(ex, "k_zenc_tables", k_zenc_tables<<<(nhuf + 63) / 64, 64, 0, ex.stream>>>(A, nhuf, d_hists, d_metas));
        KLAUNCH(ex, "k_zenc_encode", k_zenc_encode<<<nhuf, 256, ZENC_SMEM, ex.stream>>>(A, d_metas));
    }
    if (side) CUDA_TRY(cudaStreamWaitEvent(ex.stream, ctx.side_join, 0));
    b.d_off = ex.alloc<u64>(b.nblocks + 2);
    const ZEncBlock *db = b.d_blocks;
    exclusive_scan(ex, [db] __device__ (size_t i) { return (u64)db[i].csize + 3; }, b.nblocks, b.d_off);
    // frame sizes: the offsets at the stream boundaries, one small download
    u64 *d_fs = ex.alloc<u64>(16);
    {
        u32 fb[9]; for (size_t s = 0; s <= ns; s++) fb[s] = b.first_block[s];
        ZEncFirstBlocks fbs; memcpy(fbs.v, fb, sizeof fb);
        const u64 *off = b.d_off; const u32 nss = (u32)ns;
        ex.for_each(ns, [=] __device__ (size_t s) { if (s < nss) d_fs[s] = 6 + off[fbs.v[s + 1]] - off[fbs.v[s]]; }, "zenc_frame_sizes");
    }
    u64 h_fs[8];
    ex.download(h_fs, d_fs, ns * 8);
    b.frame_size.assign(ns, 0); b.dest.assign(ns, nullptr);
    for (size_t s = 0; s < ns; s++) b.frame_size[s] = h_fs[s];
}

static void zstd_gather_frames(Ctx &ctx, CudaExec &ex, ZEncBatch &b, bool skip_magic)
{
    const size_t ns = b.src.size();
    u8 **d_dest = ex.alloc<u8 *>(ns); u32 *d_first = ex.alloc<u32>(ns + 1); int *d_wlog = ex.alloc<int>(ns);
    ex.upload(d_dest, b.dest.data(), ns * sizeof(u8 *)); ex.upload(d_first, b.first_block.data(), (ns + 1) * 4); ex.upload(d_wlog, b.wlog.data(), ns * 4);
    ZGatherArgs G{b.d_blocks, b.d_slots, b.d_off, d_dest, d_first, d_wlog, skip_magic ? 1 : 0};
    KLAUNCH(ex, "k_zenc_gather", k_zenc_gather<<<b.nblocks, 256, 0, ex.stream>>>(G));
    CUDA_TRY(cudaStreamSynchronize(ex.stream));                 // host vectors above
    (void)ctx;
}

}  // namespace nafg
