// zstd_hd.cuh — zstd *format* building blocks as __host__ __device__ functions.
//
// Everything here is thread-serial logic executed by ONE GPU thread per work item (a block header,
// a Huffman table, a Huffman stream, the sequences of one block).  The parallelism of the decoder
// comes from running tens of thousands of those items at once (zstd_dec.cuh), not from inside them.
// Because the functions are HD, tests/emu runs the very same code on the CPU against the oracle
// before any GPU time is spent; the shipped library only ever instantiates them in kernels.
//
// Written from zstd/doc/zstd_compression_format.md (v1.5.0 as vendored by the reference).  The
// reference functions whose behaviour each piece replaces are cited per function
// (paths relative to /root/reference/zstd/lib).
#pragma once
#include <stdint.h>
#include <stddef.h>

#ifdef __CUDACC__
#define HD __host__ __device__ __forceinline__
#define HDN __host__ __device__
#else
#define HD inline
#define HDN
#endif

namespace nafz {

typedef uint8_t u8; typedef uint16_t u16; typedef uint32_t u32; typedef uint64_t u64;
typedef int32_t i32; typedef int64_t i64;

// What the host walk knows about a block (32 bytes: a file of 100 k blocks uploads 3 MB, not the full records)
struct ZBlockHead {
    u64 src;            // offset of the block content in the input buffer
    u64 out_base;       // first_in_stream: arena offset where this stream's output starts
    u32 csize;          // content bytes (raw: size, RLE: 1, compressed: Block_Size)
    u32 rsize;          // raw / RLE: regenerated size
    u32 frame_first_blk;// index of the first block of my frame
    u8  type;           // 0 raw, 1 RLE, 2 compressed
    u8  first_in_frame;
    u8  first_in_stream;
    u8  stream;
};

enum : int {
    Z_OK = 0,
    Z_ERR_TRUNCATED = 1, Z_ERR_LIT_HEADER = 2, Z_ERR_HUF_TREE = 3, Z_ERR_HUF_STREAM = 4, Z_ERR_FSE_HEADER = 5,
    Z_ERR_SEQ_HEADER = 6, Z_ERR_SEQ_STREAM = 7, Z_ERR_NO_TABLE = 8, Z_ERR_OFFSET = 9, Z_ERR_SIZE = 10,
    Z_ERR_RESERVED = 11, Z_ERR_TOO_LARGE = 12
};

HD int hibit(u32 v)
{
#ifdef __CUDA_ARCH__
    return 31 - __clz((int)v);
#else
    return 31 - __builtin_clz(v);
#endif
}

// ------------------------------------------------------------------ backward bit reader
// Huffman and FSE bitstreams are written forward and read from their last byte (spec "Huffman
// Coding" / "FSE").  `cont` holds the next unread bits left-aligned; bytes are pulled from
// decreasing addresses.  Reading past the start of the stream yields zero bits and is detected by
// `consumed() > total`.  Replaces common/bitstream.h BIT_initDStream/BIT_reloadDStream.
struct BackBits {
    const u8 *p;      // stream start
    i64 pos;          // byte offset (may be negative near the start) of the 8 bytes held in `cont`
    u64 cont;         // bytes p[pos .. pos+8), little endian: the stream's next bit is bit (63 - consumed)
    u32 consumed;     // bits of `cont` already used, counted from its top
    i64 left;         // payload bits not yet consumed (excludes padding + end mark); < 0 = read past the start

    // 8 bytes at any alignment: two aligned loads + funnel shift.  May touch up to 7 bytes on either side of
    // [q, q+8): callers keep >= 8 readable bytes before the first stream and after the last one.
    static HD u64 load64(const u8 *q)
    {
        const u64 *al = (const u64 *)((uintptr_t)q & ~(uintptr_t)7);
        const u32 sh = (u32)((uintptr_t)q & 7) * 8;
        const u64 lo = al[0];
        return sh ? (lo >> sh) | (al[1] << (64 - sh)) : lo;
    }
    HD bool init(const u8 *src, size_t n)
    {
        p = src; pos = (i64)n - 8; cont = 0; consumed = 0; left = 0;
        if (n == 0) return false;
        const u8 last = src[n - 1];
        if (last == 0) return false;
        const int pad = 8 - hibit(last);          // zero padding bits + the end-mark bit
        cont = load64(src + pos);
        consumed = (u32)pad; left = (i64)n * 8 - pad;
        return true;
    }
    // drop whole consumed bytes and fetch 8 fresh ones: afterwards at least 57 bits are available.
    // Bits below the start of the stream are whatever precedes it in memory (the format only ever needs
    // them as "don't care": Huffman prefixes are already decided, FSE overruns are detected by `left`).
    HD void reload()
    {
        pos -= (i64)(consumed >> 3); consumed &= 7;
        if (pos < -8) pos = -8;
        cont = load64(p + pos);
    }
    HD u32 peek(int n) const { return n ? (u32)((cont << consumed) >> (64 - n)) : 0u; }
    HD void skip(int n) { consumed += (u32)n; left -= n; }
    HD u32 read(int n) { if (consumed + (u32)n > 64) reload(); u32 v = peek(n); skip(n); return v; }   // n <= 32
    HD bool overrun() const { return left < 0; }
    HD bool exact() const { return left == 0; }
};

// ------------------------------------------------------------------ FSE
// Table entry: symbol | nbBits << 8 | baseline << 16.  Replaces decompress/zstd_decompress_block.c:508
// ZSTD_buildFSETable and common/fse_decompress.c FSE_buildDTable_internal.
HD u32 fse_sym(u32 e) { return e & 0xFF; }
HD u32 fse_nb(u32 e) { return (e >> 8) & 0xFF; }
HD u32 fse_base(u32 e) { return e >> 16; }

// norm[s] in {-1, 0, 1..}; returns false on an inconsistent distribution.  `tmp` = nsym u16 scratch.
HD bool fse_build_table(u32 *table, const short *norm, int nsym, int log, u16 *next)
{
    const int size = 1 << log;
    int high = size - 1;
    for (int s = 0; s < nsym; s++) {
        if (norm[s] == -1) { table[high--] = (u32)s; next[s] = 1; }
        else next[s] = (u16)norm[s];
    }
    const int step = (size >> 1) + (size >> 3) + 3, mask = size - 1;
    int pos = 0;
    for (int s = 0; s < nsym; s++) {
        for (int i = 0; i < norm[s]; i++) {
            table[pos] = (u32)s;
            do { pos = (pos + step) & mask; } while (pos > high);
        }
    }
    if (pos != 0) return false;
    for (int i = 0; i < size; i++) {
        u32 s = table[i];
        u32 n = next[s]++;
        int nb = log - hibit(n);
        table[i] = s | ((u32)nb << 8) | ((((n << nb) - (u32)size) & 0xFFFF) << 16);
    }
    return true;
}

// Forward LSB-first bit peek used by the FSE table description (spec "FSE Table Description").
HD u32 fwd_peek(const u8 *p, size_t n, size_t bitpos, int nb)
{
    u64 v = 0; size_t byte = bitpos >> 3;
    for (int i = 0; i < 5; i++) if (byte + i < n) v |= (u64)p[byte + i] << (8 * i);
    v >>= (bitpos & 7);
    return (u32)(v & ((1ull << nb) - 1));
}

// Replaces common/entropy_common.c:64 FSE_readNCount_body.  Returns bytes used, 0 on error.
HD size_t fse_read_ncount(const u8 *p, size_t n, short *norm, int *nsym, int max_sym, int max_log, int *log_out)
{
    if (n < 1) return 0;
    size_t bit = 0;
    int log = (int)fwd_peek(p, n, bit, 4) + 5; bit += 4;
    if (log > max_log) return 0;
    int remaining = 1 << log, sym = 0;
    while (remaining > 0 && sym <= max_sym) {
        int bits = hibit((u32)remaining + 1) + 1;
        u32 val = fwd_peek(p, n, bit, bits);
        u32 lower = (1u << (bits - 1)) - 1;
        u32 thresh = (1u << bits) - 1 - (u32)(remaining + 1);
        if ((val & lower) < thresh) { bit += bits - 1; val &= lower; }
        else { bit += bits; if (val > lower) val -= thresh; }
        int proba = (int)val - 1;
        remaining -= proba < 0 ? 1 : proba;
        norm[sym++] = (short)proba;
        if (proba == 0) {
            u32 rep;
            do {
                rep = fwd_peek(p, n, bit, 2); bit += 2;
                for (u32 i = 0; i < rep && sym <= max_sym; i++) norm[sym++] = 0;
            } while (rep == 3);
        }
        if ((bit >> 3) > n + 4) return 0;
    }
    if (remaining != 0 || sym > max_sym + 1) return 0;
    size_t used = (bit + 7) >> 3;
    if (used > n) return 0;
    *nsym = sym; *log_out = log;
    return used;
}

// ------------------------------------------------------------------ predefined distributions / code tables
// spec "Default Distributions", "Literals length codes", "Match length codes"
// (common/zstd_internal.h:185-242).
#ifdef __CUDA_ARCH__
#define ZCONST __constant__
#else
#define ZCONST static const
#endif

struct SeqConsts {
    short ll_norm[36]; short ml_norm[53]; short of_norm[29];
    u32 ll_base[36]; u8 ll_bits[36]; u32 ml_base[53]; u8 ml_bits[53];
};

HD void seq_consts_init(SeqConsts &c)
{
    const short ll[36] = { 4,3,2,2,2,2,2,2,2,2,2,2,2,1,1,1,2,2,2,2,2,2,2,2,2,3,2,1,1,1,1,1,-1,-1,-1,-1 };
    const short ml[53] = { 1,4,3,2,2,2,2,2,2,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,
                           1,1,1,1,1,1,1,1,1,1,1,1,1,1,-1,-1,-1,-1,-1,-1,-1 };
    const short of[29] = { 1,1,1,1,1,1,2,2,2,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,-1,-1,-1,-1,-1 };
    for (int i = 0; i < 36; i++) c.ll_norm[i] = ll[i];
    for (int i = 0; i < 53; i++) c.ml_norm[i] = ml[i];
    for (int i = 0; i < 29; i++) c.of_norm[i] = of[i];
}

HD u32 ll_base_of(u32 c) { return c < 16 ? c : (c < 20 ? 16 + ((c - 16) << 1) : (c < 22 ? 24 + ((c - 20) << 2) : (c < 24 ? 32 + ((c - 22) << 3) : (c == 24 ? 48u : (1u << (c - 19)))))); }
HD u32 ll_bits_of(u32 c) { return c < 16 ? 0 : (c < 20 ? 1 : (c < 22 ? 2 : (c < 24 ? 3 : (c == 24 ? 4 : c - 19)))); }
HD u32 ml_base_of(u32 c)
{
    if (c < 32) return c + 3;
    if (c < 36) return 35 + ((c - 32) << 1);
    if (c < 38) return 43 + ((c - 36) << 2);
    if (c < 40) return 51 + ((c - 38) << 3);
    if (c < 42) return 67 + ((c - 40) << 4);
    if (c == 42) return 99;
    return (1u << (c - 36)) + 3;      // 43 -> 131, 44 -> 259, ... 52 -> 65539
}
HD u32 ml_bits_of(u32 c)
{
    if (c < 32) return 0;
    if (c < 36) return 1;
    if (c < 38) return 2;
    if (c < 40) return 3;
    if (c < 42) return 4;
    if (c == 42) return 5;
    return c - 36;                    // 43 -> 7 ... 52 -> 16
}

// ------------------------------------------------------------------ Huffman
// Decode table entry (u16): symbol | nbBits << 8; 1 << max_bits entries.
// Replaces decompress/huf_decompress.c:142 HUF_readDTableX1_wksp + common/entropy_common.c:265 HUF_readStats_body.

// Reads the tree description into weights[0..*nw) (implied last weight appended).  Returns bytes used, 0 on error.
HD size_t huf_read_weights(const u8 *p, size_t n, u8 *weights, int *nw_out, int *max_bits_out)
{
    if (n < 1) return 0;
    int hb = p[0], nw = 0; size_t used;
    if (hb >= 128) {
        nw = hb - 127;
        size_t bytes = (size_t)(nw + 1) / 2;
        if (1 + bytes > n) return 0;
        for (int i = 0; i < nw; i++) weights[i] = (i & 1) ? (p[1 + i / 2] & 15) : (p[1 + i / 2] >> 4);
        used = 1 + bytes;
    } else {
        if (hb == 0 || (size_t)1 + hb > n) return 0;
        short norm[16]; int nsym, log; u16 next[16]; u32 table[64];
        size_t hdr = fse_read_ncount(p + 1, hb, norm, &nsym, 12, 6, &log);
        if (hdr == 0 || hdr >= (size_t)hb) return 0;
        if (!fse_build_table(table, norm, nsym, log, next)) return 0;
        BackBits b;
        if (!b.init(p + 1 + hdr, (size_t)hb - hdr)) return 0;
        u32 s1 = b.read(log), s2 = b.read(log);
        if (b.overrun()) return 0;
        for (;;) {
            if (nw >= 254) return 0;
            u32 e = table[s1]; weights[nw++] = (u8)fse_sym(e);
            s1 = fse_base(e) + b.read((int)fse_nb(e));
            if (b.overrun()) { weights[nw++] = (u8)fse_sym(table[s2]); break; }
            if (nw >= 254) return 0;
            e = table[s2]; weights[nw++] = (u8)fse_sym(e);
            s2 = fse_base(e) + b.read((int)fse_nb(e));
            if (b.overrun()) { weights[nw++] = (u8)fse_sym(table[s1]); break; }
        }
        used = 1 + (size_t)hb;
    }
    if (nw > 255) return 0;
    u32 total = 0;
    for (int i = 0; i < nw; i++) { if (weights[i] > 11) return 0; if (weights[i]) total += 1u << (weights[i] - 1); }
    if (total == 0) return 0;
    int max_bits = hibit(total) + 1;
    if (max_bits > 11) return 0;
    u32 left = (1u << max_bits) - total;
    if (left & (left - 1)) return 0;
    weights[nw++] = (u8)(hibit(left) + 1);
    *nw_out = nw; *max_bits_out = max_bits;
    return used;
}

// Fills table[0 .. 1<<max_bits).  Longer codes take the numerically lower slots, equal lengths in
// symbol order (spec "Huffman Tree Description": conversion of weights into prefix codes).
HD bool huf_build_table(u16 *table, const u8 *weights, int nw, int max_bits)
{
    u32 rank_count[13], rank_idx[13];
    for (int i = 0; i < 13; i++) rank_count[i] = 0;
    for (int i = 0; i < nw; i++) if (weights[i]) rank_count[max_bits + 1 - weights[i]]++;
    rank_idx[max_bits] = 0;
    for (int b = max_bits; b >= 1; b--) rank_idx[b - 1] = rank_idx[b] + rank_count[b] * (1u << (max_bits - b));
    if (rank_idx[0] != (1u << max_bits)) return false;
    for (int i = 0; i < nw; i++) {
        if (!weights[i]) continue;
        int bits = max_bits + 1 - weights[i];
        u32 len = 1u << (max_bits - bits), at = rank_idx[bits];
        u16 e = (u16)((u32)i | ((u32)bits << 8));
        for (u32 k = 0; k < len; k++) table[at + k] = e;
        rank_idx[bits] = at + len;
    }
    return true;
}

// Backward bit reader for the Huffman streams, built so that every byte of the stream crosses the memory system
// once: the compressed bytes are fetched as ALIGNED 16-byte chunks into a four-register queue and enter a 64-bit
// bit buffer 32 bits at a time.  (BackBits above re-reads two aligned 8-byte words around its cursor on every
// reload: with thousands of streams in flight per SM those lines do not survive in L1 and each reload becomes two
// 32-byte sector requests to L2 for two or three useful bytes.)
struct BackBitsQ {
    u64 bb; int nb;              // next bit = MSB of bb; nb valid bits (32 < nb <= 64 after refill())
    i64 left;                    // payload bits not yet consumed; < 0 = read past the start
    const u8 *src;               // stream start: chunks entirely below it are not fetched (read as zero)
    const u32 *cp;               // lowest 16-byte chunk requested so far
    u32 w0, w1, w2, w3; int qn;  // queued aligned words, the next one is w3
    u32 n0, n1, n2, n3;          // the chunk below the queue, requested one queue ahead: its load latency is covered by
                                 // the decoding of the four queued words instead of stalling the warp
    u32 hi, sh;                  // aligned word holding the cursor, cursor misalignment in bits
#ifdef __CUDA_ARCH__
    // On the GPU a stream is decoded by one thread and what bounds it is the chain of dependent instructions from one
    // symbol to the next (ncu: ~23 dependent instructions per symbol with the 64-bit buffer, warps mostly waiting on their own
    // previous instruction).  So the buffer is two 32-bit registers moved by funnel shifts: peek = one shift of `bh`,
    // skip = one funnel shift, and the bit accounting (`left`) is derived at the end from the number of words taken.
    u32 bh, bl; int npop, loaded0;
#endif

    HD void request()
    {
        cp -= 4;
        if ((uintptr_t)(cp + 4) > (uintptr_t)src) {
#ifdef __CUDA_ARCH__
            const uint4 v = *(const uint4 *)cp; n0 = v.x; n1 = v.y; n2 = v.z; n3 = v.w;
#else
            n0 = cp[0]; n1 = cp[1]; n2 = cp[2]; n3 = cp[3];
#endif
        } else n0 = n1 = n2 = n3 = 0;
    }
    HD void fetch() { w0 = n0; w1 = n1; w2 = n2; w3 = n3; qn = 4; request(); }
    HD u32 pop() { if (qn == 0) fetch(); const u32 x = w3; w3 = w2; w2 = w1; w1 = w0; qn--; return x; }
    HD bool init(const u8 *s, size_t n)
    {
        src = s; bb = 0; nb = 0; left = 0; qn = 0; hi = 0; sh = 0; w0 = w1 = w2 = w3 = 0; n0 = n1 = n2 = n3 = 0; cp = nullptr;
        if (n == 0) return false;
        const u8 last = s[n - 1];
        if (last == 0) return false;
        const int pad = 8 - hibit(last);          // zero padding bits + the end-mark bit
        const u8 *p = s + n - 8;                  // the top 8 bytes (may start below s for tiny streams: don't-care bits)
        bb = BackBits::load64(p) << pad; nb = 64 - pad; left = (i64)n * 8 - pad;
#ifdef __CUDA_ARCH__
        bh = (u32)(bb >> 32); bl = (u32)bb; npop = 0; loaded0 = 64 - pad;
#endif
        const u32 a = (u32)((uintptr_t)p & 3);
        sh = a * 8;
        const u8 *wa = p - a;                     // aligned word holding the cursor
        if (a) hi = *(const u32 *)wa;
        const u8 *na = wa - 4;                    // next aligned word below the cursor
        const u32 j = (u32)(((uintptr_t)na >> 2) & 3);
        cp = (const u32 *)((uintptr_t)na & ~(uintptr_t)15) + 4;
        request(); fetch();                       // chunk holding `na` (its words above index j are not part of the queue) + the one below
        for (u32 k = j; k < 3; k++) { w3 = w2; w2 = w1; w1 = w0; }
        qn = (int)j + 1;
        return true;
    }
#ifdef __CUDA_ARCH__
    HD void refill()
    {
        if (nb <= 32) {                                         // then bl == 0: every valid bit is in bh
            const u32 lw = pop();
            const u32 v = __funnelshift_r(lw, hi, sh);          // low word of (hi : lw) >> sh; sh is 0, 8, 16 or 24
            hi = lw;
            bh |= __funnelshift_rc(v, 0u, (u32)nb);             // v >> nb (0 when nb == 32)
            bl = __funnelshift_lc(0u, v, (u32)(32 - nb));       // v << (32 - nb) (0 when nb == 0)
            nb += 32; npop++;
        }
    }
    HD u32 peek(int n) const { return bh >> (32 - n); }         // 1 <= n <= 32
    HD void skip(int n) { bh = __funnelshift_l(bl, bh, (u32)n); bl <<= n; nb -= n; }      // n <= 11 (Huffman code lengths)
    HD bool exact() const { return (i64)loaded0 + 32ll * npop - nb == left; }
#else
    HD void refill()
    {
        if (nb <= 32) {
            const u32 lo = pop();
            const u32 v = sh ? (lo >> sh) | (hi << (32 - sh)) : lo;
            hi = lo;
            bb |= (u64)v << (32 - nb); nb += 32;
        }
    }
    HD u32 peek(int n) const { return (u32)(bb >> (64 - n)); }      // 1 <= n <= 32
    HD void skip(int n) { bb <<= n; nb -= n; left -= n; }
    HD bool exact() const { return left == 0; }
#endif
};

// One Huffman-coded stream -> nout symbols.  Replaces decompress/huf_decompress.c:350
// HUF_decompress4X1_usingDTable_internal_body's per-stream loop (and the 1X1 variant :285).
HD bool huf_decode_stream(const u16 *table, int max_bits, const u8 *src, size_t n, u8 *dst, size_t nout)
{
    BackBitsQ b;
    if (!b.init(src, n)) return false;
    size_t i = 0;
    // head: byte stores until dst is 16-byte aligned
    while (i < nout && (((uintptr_t)(dst + i)) & 15)) {
        b.refill();
        const u32 e = table[b.peek(max_bits)];
        dst[i++] = (u8)e; b.skip((int)(e >> 8));
    }
    // body: 16 symbols per 128-bit store; one refill per 2 symbols (2 x 11 bits <= the 33 bits refill() guarantees)
    for (; i + 16 <= nout; i += 16) {
        u32 w[4];
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
        for (int k = 0; k < 4; k++) {
            u32 v = 0;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
            for (int j = 0; j < 4; j++) {
                if (!(j & 1)) b.refill();
                const u32 e = table[b.peek(max_bits)];
                v |= (e & 0xFF) << (8 * j); b.skip((int)(e >> 8));
            }
            w[k] = v;
        }
#ifdef __CUDA_ARCH__
        *(uint4 *)(dst + i) = make_uint4(w[0], w[1], w[2], w[3]);
#else
        for (int k = 0; k < 4; k++) for (int j = 0; j < 4; j++) dst[i + 4 * k + j] = (u8)(w[k] >> (8 * j));
#endif
    }
    for (; i < nout; i++) {
        b.refill();
        const u32 e = table[b.peek(max_bits)];
        dst[i] = (u8)e; b.skip((int)(e >> 8));
    }
    return b.exact();
}

static const int HUF_RING = 4;                                 // slots per lane of the shared-memory ring below
#ifdef __CUDA_ARCH__
// The same stream decoded on the GPU with its compressed bytes staged through shared memory.
//
// Why: a warp's 32 lanes decode 32 different streams and run out of queued words at different moments.  With the
// register-only reader above every such moment is a 16-byte global load into the SAME architectural registers for whichever
// lane needs it, and a register's scoreboard belongs to the warp, not to a lane: the lane that refills next waits for the
// load another lane issued a few cycles ago (ncu: a quarter of all stall samples sit on the moves out of those registers).
// Here the global -> on-chip step is cp.async into a per-lane ring of four 16-byte slots, topped up at a point all lanes reach
// together (once per 16 symbols) and awaited one iteration later; a lane that runs dry only does a shared-memory load.
struct BackBitsR {
    u32 bh, bl; int nb, npop, loaded0; i64 left;
    u32 hw, sh;
    const uint4 *gp; const u8 *src;
    u32 sbase, stride, issued;                                 // shared-memory address of slot 0, bytes between my slots
    u32 c;                                                     // words taken out of the ring so far (+ the words of the first chunk above my first one)

    // Words are read straight out of the ring, one 32-bit shared-memory load per refill, addressed by a counter: chunk c / 4
    // sits in slot (c / 4) % HUF_RING, and inside a chunk the words are used from the highest address down.  (The first
    // version of this reader moved whole chunks into a four-word register queue -- a nested branch inside the refill branch, at
    // a different symbol for every lane: ncu showed 21 of 32 lanes active over the whole kernel.)
    __device__ __forceinline__ void top_up()
    {
#pragma unroll
        for (int k = 0; k < 2; k++) {
            if (issued - (c >> 2) < (u32)HUF_RING) {
                const u32 slot = sbase + (issued & (HUF_RING - 1)) * stride;
                if ((uintptr_t)(gp + 1) > (uintptr_t)src) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(slot), "l"(gp) : "memory");
                else asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(slot), "r"(0u) : "memory");
                gp--; issued++;
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    __device__ __forceinline__ u32 pop()
    {
        const u32 at = sbase + ((c >> 2) & (HUF_RING - 1)) * stride + ((3u - (c & 3u)) << 2);
        u32 x;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(x) : "r"(at) : "memory");
        c++;
        return x;
    }
    __device__ __forceinline__ bool init(const u8 *s, size_t n, u32 ring_slot0, u32 ring_stride)
    {
        src = s; sbase = ring_slot0; stride = ring_stride; issued = 0; c = 0; npop = 0; hw = 0;
        if (n == 0) return false;
        const u8 last = s[n - 1];
        if (last == 0) return false;
        const int pad = 8 - hibit(last);
        const u8 *p = s + n - 8;
        const u64 bb = BackBits::load64(p) << pad;
        bh = (u32)(bb >> 32); bl = (u32)bb; nb = 64 - pad; loaded0 = nb; left = (i64)n * 8 - pad;
        const u32 a = (u32)((uintptr_t)p & 3);
        sh = a * 8;
        const u8 *wa = p - a;
        if (a) hw = *(const u32 *)wa;
        const u8 *na = wa - 4;
        const u32 j = (u32)(((uintptr_t)na >> 2) & 3);
        gp = (const uint4 *)((uintptr_t)na & ~(uintptr_t)15);
        top_up(); top_up();                                    // four chunks: the one holding `na` and the three below it
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        c = 3 - j;                                             // the first word I use is word j of the first chunk
        return true;
    }
    __device__ __forceinline__ void refill()
    {
        if (nb <= 32) {
            const u32 lw = pop();
            const u32 v = __funnelshift_r(lw, hw, sh);
            hw = lw;
            bh |= __funnelshift_rc(v, 0u, (u32)nb);
            bl = __funnelshift_lc(0u, v, (u32)(32 - nb));
            nb += 32; npop++;
        }
    }
    __device__ __forceinline__ u32 peek(int n) const { return bh >> (32 - n); }
    __device__ __forceinline__ void skip(int n) { bh = __funnelshift_l(bl, bh, (u32)n); bl <<= n; nb -= n; }
    __device__ __forceinline__ bool exact() const { return (i64)loaded0 + 32ll * npop - nb == left; }
};

// ring_slot0: shared-memory address (cvta'd) of this thread's first 16-byte slot; ring_stride: bytes to its next slot
__device__ __forceinline__ bool huf_decode_stream_ring(const u16 *table, int max_bits, const u8 *src, size_t n, u8 *dst, size_t nout,
                                                       u32 ring_slot0, u32 ring_stride)
{
    BackBitsR b;
    if (!b.init(src, n, ring_slot0, ring_stride)) return false;
    size_t i = 0;
    while (i < nout && (((uintptr_t)(dst + i)) & 15)) {       // <= 15 symbols: at most two slots
        b.refill();
        const u32 e = table[b.peek(max_bits)];
        dst[i++] = (u8)e; b.skip((int)(e >> 8));
    }
    for (; i + 16 <= nout; i += 16) {
        b.top_up();                                            // what the last iteration took out goes back in ...
        asm volatile("cp.async.wait_group 1;" ::: "memory");   // ... and what was requested an iteration ago has arrived
        u32 w[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            u32 v = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                if (!(j & 1)) b.refill();
                const u32 e = table[b.peek(max_bits)];
                v |= (e & 0xFF) << (8 * j); b.skip((int)(e >> 8));
            }
            w[k] = v;
        }
        *(uint4 *)(dst + i) = make_uint4(w[0], w[1], w[2], w[3]);
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    for (; i < nout; i++) {
        b.refill();
        const u32 e = table[b.peek(max_bits)];
        dst[i] = (u8)e; b.skip((int)(e >> 8));
    }
    return b.exact();
}
#endif

// ------------------------------------------------------------------ literals section header
// spec "Literals_Section_Header"; replaces decompress/zstd_decompress_block.c:79 ZSTD_decodeLiteralsBlock's parsing.
struct LitHeader {
    u32 type;        // 0 raw, 1 RLE, 2 compressed, 3 treeless
    u32 streams;     // 1 or 4
    u32 regen;       // regenerated size
    u32 csize;       // compressed payload size (tree + jump table + streams); raw: regen; RLE: 1
    u32 hdr;         // header bytes
};

HD int lit_header_parse(const u8 *p, size_t n, LitHeader &h)
{
    if (n < 1) return Z_ERR_TRUNCATED;
    u32 type = p[0] & 3, sf = (p[0] >> 2) & 3;
    h.type = type; h.streams = 1;
    if (type < 2) {
        if ((sf & 1) == 0) { h.regen = p[0] >> 3; h.hdr = 1; }
        else if (sf == 1) { if (n < 2) return Z_ERR_TRUNCATED; h.regen = (p[0] >> 4) | ((u32)p[1] << 4); h.hdr = 2; }
        else { if (n < 3) return Z_ERR_TRUNCATED; h.regen = (p[0] >> 4) | ((u32)p[1] << 4) | ((u32)p[2] << 12); h.hdr = 3; }
        h.csize = type == 0 ? h.regen : 1;
    } else {
        if (sf < 2) {
            if (n < 3) return Z_ERR_TRUNCATED;
            u32 v = p[0] | ((u32)p[1] << 8) | ((u32)p[2] << 16);
            h.regen = (v >> 4) & 0x3FF; h.csize = (v >> 14) & 0x3FF; h.hdr = 3; h.streams = sf == 0 ? 1 : 4;
        } else if (sf == 2) {
            if (n < 4) return Z_ERR_TRUNCATED;
            u32 v = p[0] | ((u32)p[1] << 8) | ((u32)p[2] << 16) | ((u32)p[3] << 24);
            h.regen = (v >> 4) & 0x3FFF; h.csize = (v >> 18) & 0x3FFF; h.hdr = 4; h.streams = 4;
        } else {
            if (n < 5) return Z_ERR_TRUNCATED;
            u64 v = p[0] | ((u64)p[1] << 8) | ((u64)p[2] << 16) | ((u64)p[3] << 24) | ((u64)p[4] << 32);
            h.regen = (u32)((v >> 4) & 0x3FFFF); h.csize = (u32)((v >> 22) & 0x3FFFF); h.hdr = 5; h.streams = 4;
        }
    }
    if (h.regen > 128 * 1024) return Z_ERR_LIT_HEADER;
    if ((size_t)h.hdr + h.csize > n) return Z_ERR_TRUNCATED;
    return Z_OK;
}

// ------------------------------------------------------------------ sequences section header
// spec "Sequences_Section_Header"; replaces decompress/zstd_decompress_block.c:577 ZSTD_decodeSeqHeaders' first part.
// Returns bytes used (nbSeq field + modes byte if nbSeq > 0), 0 on error.
HD size_t seq_header_parse(const u8 *p, size_t n, u32 *nseq, u32 *modes)
{
    if (n < 1) return 0;
    size_t pos;
    if (p[0] < 128) { *nseq = p[0]; pos = 1; }
    else if (p[0] < 255) { if (n < 2) return 0; *nseq = ((u32)(p[0] - 128) << 8) + p[1]; pos = 2; }
    else { if (n < 3) return 0; *nseq = (u32)p[1] + ((u32)p[2] << 8) + 0x7F00; pos = 3; }
    *modes = 0;
    if (*nseq == 0) return pos;
    if (pos >= n) return 0;
    *modes = p[pos++];
    return pos;
}

// ------------------------------------------------------------------ repeat offsets as transfer functions
// spec "Repeat Offsets".  To decode blocks in parallel, each block first computes how it maps the
// repeat-offset history it receives to the history it leaves (RepFn); a scan composes them.
// A slot value is either a concrete offset (src < 0) or "incoming slot src, plus delta" (delta <= 0).
struct RepSlot { i32 src; i32 delta; u32 value; };
struct RepFn { RepSlot s[3]; };

HD RepFn repfn_identity()
{
    RepFn f;
    for (int i = 0; i < 3; i++) { f.s[i].src = i; f.s[i].delta = 0; f.s[i].value = 0; }
    return f;
}
HD RepSlot repslot_concrete(u32 v) { RepSlot r; r.src = -1; r.delta = 0; r.value = v; return r; }

// Apply one sequence's offset_value / literal length to a symbolic history.  Returns the slot
// describing the actual offset used by this sequence.
HD RepSlot repfn_step(RepFn &f, u32 ofv, u32 ll)
{
    // (no f.s[idx] with a run-time idx: that would put the whole history in local memory on the GPU, and this runs once
    // per sequence in a single thread per block)
    RepSlot used;
    if (ofv > 3) {
        used = repslot_concrete(ofv - 3);
        f.s[2] = f.s[1]; f.s[1] = f.s[0]; f.s[0] = used;
        return used;
    }
    const u32 idx = ofv - 1 + (ll == 0 ? 1u : 0u);
    if (idx == 0) return f.s[0];
    if (idx == 1) { used = f.s[1]; f.s[1] = f.s[0]; f.s[0] = used; return used; }
    if (idx == 2) used = f.s[2];
    else { used = f.s[0]; if (used.src < 0) used.value -= 1; else used.delta -= 1; }
    f.s[2] = f.s[1]; f.s[1] = f.s[0]; f.s[0] = used;
    return used;
}

HD u32 repslot_eval(const RepSlot &s, const u32 in[3]) { return s.src < 0 ? s.value : in[s.src] + (u32)s.delta; }

// history after applying f to `in`
HD void repfn_apply(const RepFn &f, const u32 in[3], u32 out[3])
{
    u32 t0 = repslot_eval(f.s[0], in), t1 = repslot_eval(f.s[1], in), t2 = repslot_eval(f.s[2], in);
    out[0] = t0; out[1] = t1; out[2] = t2;
}

}  // namespace nafz
