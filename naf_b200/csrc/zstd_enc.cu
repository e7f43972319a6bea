// zstd_enc.cu — block-parallel zstd *encoder* (included by naf_enc.cu; same translation unit).
//
// Replaces what ennaf gets from ZSTD_initCStream / ZSTD_compressStream / ZSTD_endStream
// (ennaf/src/compressor.c:7-20,120,64) for every NAF stream.  The reference's encoder is free-parse:
// parity on this side is "the reference unnaf decodes our frame back to the same bytes"
// (SURVEY §8a row 19), so the parse is chosen for the GPU:
//   * one frame per stream (unnaf's SEQ/QUAL loops stop after the first frame: SURVEY A.2), FHD 0x00,
//     declared window 128 KB, no content size / checksum / dictionary — what ennaf's frames look like
//   * 64 KB blocks, each independent of every other: own Huffman table, no sequences, so both our
//     decoder and libzstd can start anywhere; RLE block when a block is one repeated byte, raw block
//     when Huffman would not shrink it
//   * literals: canonical length-limited (<= 11 bit) Huffman, 4 streams (1 stream below 1 KB), tree
//     description as FSE-compressed weights when smaller / required (> 128 listed weights)
// One CTA per block does histogram -> code lengths -> tree description -> bit-exact sizes -> encode into
// a fixed slot; a scan over block sizes then lets k_zenc_gather lay the blocks out as frames.
//
// Format: zstd/doc/zstd_compression_format.md ("Huffman Tree Description", "Huffman-coded streams",
// "FSE Table Description"); reference counterparts: compress/huf_compress.c:513 HUF_buildCTable_wksp,
// :116 HUF_writeCTable_wksp, compress/fse_compress.c:437 FSE_normalizeCount, :292 FSE_writeNCount,
// compress/zstd_compress_literals.c:70 ZSTD_compressLiterals, zstd_compress.c:3967 ZSTD_writeFrameHeader.

namespace nafg {

static const u32 ZBS = 64 * 1024;            // uncompressed bytes per block
static const u32 ZSLOT = ZBS + 512;          // bytes reserved per block for its compressed content
static const int ZWINDOW_LOG = 17;

struct ZEncBlock {
    const u8 *src; u32 n; u32 stream; u32 last;
    u32 type;      // 0 raw (content = src bytes), 1 RLE, 2 compressed (content in slot)
    u32 csize;     // content bytes
};

struct ZEncBatch {
    std::vector<const u8 *> src; std::vector<u64> n; std::vector<int> wlog;
    std::vector<u64> frame_size; std::vector<u8 *> dest;
    std::vector<u32> first_block;            // per stream (+1 sentinel)
    ZEncBlock *d_blocks = nullptr; u32 nblocks = 0; u8 *d_slots = nullptr; u64 *d_off = nullptr;
    void add(const u8 *p, u64 bytes, int window_log) { src.push_back(p); n.push_back(bytes); wlog.push_back(window_log); }
};

// ---- forward LSB-first bit writer into a small local buffer (FSE weights)
struct BitW {
    u8 *p; u32 cap; u32 pos; u64 acc; u32 fill; bool ok;
    __device__ void init(u8 *dst, u32 c) { p = dst; cap = c; pos = 0; acc = 0; fill = 0; ok = true; }
    __device__ void put(u32 v, u32 nb)
    {
        acc |= (u64)(v & ((1u << nb) - 1)) << fill; fill += nb;
        while (fill >= 8) { if (pos < cap) p[pos++] = (u8)acc; else ok = false; acc >>= 8; fill -= 8; }
    }
    __device__ u32 finish_with_mark() { put(1, 1); if (fill) { if (pos < cap) p[pos++] = (u8)acc; else ok = false; fill = 0; acc = 0; } return pos; }
    __device__ u32 finish_aligned() { if (fill) { if (pos < cap) p[pos++] = (u8)acc; else ok = false; fill = 0; acc = 0; } return pos; }
};

// FSE-compress the Huffman weights w[0..n).  Returns the number of bytes written (table description +
// bitstream), or 0 when not representable / not worthwhile.  Mirrors compress/huf_compress.c:76 HUF_compressWeights.
__device__ u32 fse_compress_weights(const u8 *w, int n, u8 *dst, u32 cap)
{
    const int LOG = 6, SIZE = 64;
    if (n <= 1) return 0;
    int count[13]; for (int i = 0; i < 13; i++) count[i] = 0;
    int maxw = 0, maxc = 0;
    for (int i = 0; i < n; i++) { count[w[i]]++; if (w[i] > maxw) maxw = w[i]; }
    for (int s = 0; s <= maxw; s++) if (count[s] > maxc) maxc = count[s];
    if (maxc == n || maxc == 1) return 0;                 // one symbol only / all distinct: not compressible
    // normalise to SIZE slots, every present symbol >= 1
    int norm[13], sum = 0;
    for (int s = 0; s <= maxw; s++) { norm[s] = count[s] ? (count[s] * SIZE + n / 2) / n : 0; if (count[s] && norm[s] < 1) norm[s] = 1; sum += norm[s]; }
    while (sum != SIZE) {
        int best = -1;
        for (int s = 0; s <= maxw; s++) if (norm[s] > (sum > SIZE ? 1 : 0) && (best < 0 || norm[s] > norm[best])) best = s;
        if (best < 0) return 0;
        if (sum > SIZE) { norm[best]--; sum--; } else { norm[best]++; sum++; }
    }
    // table description (spec "FSE Table Description")
    BitW bw; bw.init(dst, cap);
    bw.put(LOG - 5, 4);
    int remaining = SIZE, s = 0;
    while (remaining > 0 && s <= maxw) {
        int bits = nafz::hibit((u32)remaining + 1) + 1;
        u32 lower = (1u << (bits - 1)) - 1, thresh = (1u << bits) - 1 - (u32)(remaining + 1);
        u32 v = (u32)(norm[s] + 1);
        if (v < thresh) bw.put(v, bits - 1);
        else bw.put(v > lower ? v + thresh : v, bits);
        remaining -= norm[s];
        bool zero = norm[s] == 0;
        s++;
        if (zero) {
            int run = 0;
            while (s <= maxw && norm[s] == 0 && remaining > 0) { run++; s++; }
            while (run >= 3) { bw.put(3, 2); run -= 3; }
            bw.put((u32)run, 2);
        }
    }
    if (remaining != 0) return 0;
    u32 hdr = bw.finish_aligned();
    if (!bw.ok) return 0;
    // state table: positions of every symbol in increasing order (the decoder's spread, spec "From normalized distribution...")
    u8 tsym[SIZE], spos[SIZE]; int cum[14];
    cum[0] = 0; for (int k = 0; k <= maxw; k++) cum[k + 1] = cum[k] + norm[k];
    {
        int pos = 0; const int step = (SIZE >> 1) + (SIZE >> 3) + 3, mask = SIZE - 1;
        for (int k = 0; k <= maxw; k++) for (int i = 0; i < norm[k]; i++) { tsym[pos] = (u8)k; pos = (pos + step) & mask; }
        int occ[13]; for (int k = 0; k < 13; k++) occ[k] = 0;
        for (int p = 0; p < SIZE; p++) { int k = tsym[p]; spos[cum[k] + occ[k]++] = (u8)p; }
    }
    BitW bs; bs.init(dst + hdr, cap - hdr);
    int last = n - 1, prev = n - 2;
    u32 st[2];                                               // st[parity of the weight index]
    st[last & 1] = spos[cum[w[last]]];
    st[prev & 1] = spos[cum[w[prev]]];
    for (int i = n - 3; i >= 0; i--) {
        int sym = w[i], p = norm[sym];
        u32 y = st[i & 1] + SIZE;
        int nb = LOG - nafz::hibit((u32)p);
        u32 nn = y >> nb;
        if (nn < (u32)p) { nb--; nn = y >> nb; }
        bs.put(y, nb);                                       // low nb bits of y
        st[i & 1] = spos[cum[sym] + (nn - p)];
    }
    bs.put(st[1], LOG); bs.put(st[0], LOG);                  // decoder reads state1 (even chain) first
    u32 body = bs.finish_with_mark();
    if (!bs.ok) return 0;
    return hdr + body;
}

struct ZEncArgs { ZEncBlock *blk; u8 *slots; };

// The block's bytes are staged in shared memory once (coalesced 128-bit loads during the histogram pass) and
// read from there by the two later passes.  4 bytes of padding per 256 keep the per-thread 256-byte ranges of
// those passes on different banks.
__device__ __forceinline__ u32 zpad(u32 i) { return i + ((i >> 8) << 2); }
static const u32 ZSTAGE_BYTES = ZBS + (ZBS >> 8) * 4 + 16;

// One CTA (256 threads) per block.
__global__ void __launch_bounds__(256) k_zenc_block(const ZEncArgs A)
{
    __shared__ u32 hist_w[8][256];
    __shared__ u32 hist[256];
    __shared__ u32 ctab[256];                  // code | len << 16
    __shared__ u8  sorted[256], len_of[256], weight[257];
    __shared__ u32 pre_bits[257];
    __shared__ u64 smscan[33];
    __shared__ u32 s_nsym, s_maxbits, s_mode, s_tree_len, s_lit_hdr, s_nstreams, s_payload0;
    __shared__ u32 s_stream_bytes[4];

    extern __shared__ __align__(16) u8 sb[];      // staged block bytes, zpad() layout
    ZEncBlock &B = A.blk[blockIdx.x];
    const u8 *src = B.src; const u32 n = B.n;
    u8 *slot = A.slots + (size_t)blockIdx.x * ZSLOT;
    const u32 tid = threadIdx.x, warp = tid >> 5;

    if (n == 0) { if (tid == 0) { B.type = 0; B.csize = 0; } return; }

    // ---- 1. histogram
    for (int i = tid; i < 8 * 256; i += 256) (&hist_w[0][0])[i] = 0;
    __syncthreads();
    {
        const bool aligned = (((uintptr_t)src) & 15) == 0;
        u32 nvec = aligned ? n / 16 : 0;
        const uint4 *v = (const uint4 *)src;
        for (u32 i = tid; i < nvec; i += 256) {
            uint4 x = v[i];
            u32 w[4] = {x.x, x.y, x.z, x.w};
            u32 *st = (u32 *)(sb + zpad(i * 16));
            st[0] = w[0]; st[1] = w[1]; st[2] = w[2]; st[3] = w[3];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                atomicAdd(&hist_w[warp][w[k] & 0xFF], 1u); atomicAdd(&hist_w[warp][(w[k] >> 8) & 0xFF], 1u);
                atomicAdd(&hist_w[warp][(w[k] >> 16) & 0xFF], 1u); atomicAdd(&hist_w[warp][w[k] >> 24], 1u);
            }
        }
        for (u32 i = nvec * 16 + tid; i < n; i += 256) { u8 c = src[i]; sb[zpad(i)] = c; atomicAdd(&hist_w[warp][c], 1u); }
    }
    __syncthreads();
    { u32 h = 0; for (int w = 0; w < 8; w++) h += hist_w[w][tid]; hist[tid] = h; len_of[tid] = 0; weight[tid] = 0; ctab[tid] = 0; }
    if (tid == 0) { s_nsym = 0; s_mode = 2; weight[256] = 0; }
    __syncthreads();
    // ---- 2. rank symbols by (count, symbol): sorted[0] = rarest present symbol
    if (hist[tid]) {
        u32 mine = hist[tid], r = 0;
        for (int j = 0; j < 256; j++) { u32 h = hist[j]; if (h && (h < mine || (h == mine && j < (int)tid))) r++; }
        sorted[r] = (u8)tid;
        atomicAdd(&s_nsym, 1u);
    }
    __syncthreads();
    const u32 nsym = s_nsym;
    if (nsym == 1) {                                          // RLE block
        if (tid == 0) { slot[0] = sorted[0]; B.type = 1; B.csize = 1; }
        return;
    }
    // ---- 3. code lengths (thread 0): two-queue Huffman over the sorted counts, then limit to 11 bits
    if (tid == 0) {
        // reuse hist_w as scratch: [0]=leaf weight, [1]=internal weight, [2]=leaf parent, [3]=internal parent, [4]=internal depth
        u32 *lw = hist_w[0], *iw = hist_w[1], *lp = hist_w[2], *ip = hist_w[3], *idp = hist_w[4];
        for (u32 i = 0; i < nsym; i++) lw[i] = hist[sorted[i]];
        u32 li = 0, ii = 0;
        for (u32 k = 0; k + 1 < nsym; k++) {
            u32 wsum = 0;
            for (int t = 0; t < 2; t++) {
                bool take_leaf = li < nsym && (ii >= k || lw[li] <= iw[ii]);
                if (take_leaf) { wsum += lw[li]; lp[li++] = k; } else { wsum += iw[ii]; ip[ii++] = k; }
            }
            iw[k] = wsum;
        }
        const u32 root = nsym - 2;
        idp[root] = 0;
        for (int k = (int)root - 1; k >= 0; k--) idp[k] = idp[ip[k]] + 1;
        u32 num[40]; for (int i = 0; i < 40; i++) num[i] = 0;
        for (u32 i = 0; i < nsym; i++) { u32 d = idp[lp[i]] + 1; if (d > 39) d = 39; num[d]++; }
        const u32 MAXB = 11;
        for (u32 i = MAXB + 1; i < 40; i++) { num[MAXB] += num[i]; num[i] = 0; }
        u32 total = 0;
        for (u32 i = 1; i <= MAXB; i++) total += num[i] << (MAXB - i);
        while (total != (1u << MAXB)) {
            num[MAXB]--;
            for (u32 i = MAXB - 1; i > 0; i--) if (num[i]) { num[i]--; num[i + 1] += 2; break; }
            total--;
        }
        u32 idx = 0, maxbits = 0;
        for (u32 l = MAXB; l >= 1; l--) { if (num[l] && !maxbits) maxbits = l; for (u32 c = 0; c < num[l]; c++) len_of[sorted[idx++]] = (u8)l; }
        s_maxbits = maxbits;
    }
    __syncthreads();
    const u32 maxbits = s_maxbits;
    // ---- 4. canonical codes exactly as the decoder rebuilds them (longer codes first, symbols ascending)
    if (len_of[tid]) {
        u32 l = len_of[tid], start = 0;
        for (int j = 0; j < 256; j++) { u32 lj = len_of[j]; if (lj > l || (lj == l && j < (int)tid)) start += 1u << (maxbits - lj); }
        ctab[tid] = (start >> (maxbits - l)) | (l << 16);
        weight[tid] = (u8)(maxbits + 1 - l);
    }
    __syncthreads();
    // ---- 5. tree description into a scratch area at the end of the slot (thread 0)
    u8 *tree_tmp = slot + ZBS + 256;                          // <= 130 bytes
    if (tid == 0) {
        int last_sym = 255; while (last_sym > 0 && !weight[last_sym]) last_sym--;
        int nlisted = last_sym;                               // weights of symbols 0 .. last_sym-1; the last one is implied
        u32 fse = fse_compress_weights(weight, nlisted, tree_tmp + 1, 127);
        u32 direct = nlisted <= 128 ? 1 + (nlisted + 1) / 2 : 0xFFFFFFFFu;
        if (fse && fse < 128 && 1 + fse < direct) { tree_tmp[0] = (u8)fse; s_tree_len = 1 + fse; }
        else if (direct != 0xFFFFFFFFu) {
            tree_tmp[0] = (u8)(127 + nlisted);
            for (int i = 0; i < nlisted; i += 2) tree_tmp[1 + i / 2] = (u8)((weight[i] << 4) | (i + 1 < nlisted ? weight[i + 1] : 0));
            s_tree_len = direct;
        } else s_mode = 0;                                    // cannot describe the tree: raw block
        s_nstreams = n <= 1023 ? 1 : 4;
    }
    __syncthreads();
    // ---- 6. exact size of every stream
    const u32 nstreams = s_nstreams, tps = 256 / nstreams;    // threads per stream
    const u32 seg = nstreams == 4 ? (n + 3) / 4 : n;
    const u32 k = tid / tps, q = tid % tps;
    const u32 sbeg = k * seg, send = (k == nstreams - 1) ? n : (sbeg + seg < n ? sbeg + seg : n);
    const u32 slen = send > sbeg ? send - sbeg : 0;
    const u32 per = (slen + tps - 1) / tps;
    u32 t0 = sbeg + q * per, t1 = t0 + per; if (t0 > send) t0 = send; if (t1 > send) t1 = send;
    u32 bits = 0;
    for (u32 i = t0; i < t1;) {
        if (!(i & 3) && i + 4 <= t1) {                        // aligned word of four symbols
            const u32 x = *(const u32 *)(sb + zpad(i));
            bits += (ctab[x & 0xFF] >> 16) + (ctab[(x >> 8) & 0xFF] >> 16) + (ctab[(x >> 16) & 0xFF] >> 16) + (ctab[x >> 24] >> 16);
            i += 4;
        } else { bits += ctab[sb[zpad(i)]] >> 16; i++; }
    }
    u64 total_bits;
    u64 pre = block_excl_scan(bits, &total_bits, smscan);
    pre_bits[tid] = (u32)pre; if (tid == 255) pre_bits[256] = (u32)total_bits;
    __syncthreads();
    if (tid < nstreams) { u32 b = pre_bits[(tid + 1) * tps] - pre_bits[tid * tps]; s_stream_bytes[tid] = b / 8 + 1; }
    __syncthreads();
    if (tid == 0 && s_mode == 2) {
        u32 payload = s_tree_len + (nstreams == 4 ? 6 : 0);
        for (u32 j = 0; j < nstreams; j++) payload += s_stream_bytes[j];
        u32 lh = nstreams == 1 ? 3 : ((n <= 16383 && payload <= 16383) ? 4 : 5);
        if (nstreams == 1 && payload > 1023) s_mode = 0;
        if (lh + payload + 1 >= n) s_mode = 0;                // no gain: raw block
        if (n < 4 * 4 && nstreams == 4) s_mode = 0;
        s_lit_hdr = lh; s_payload0 = payload;
    }
    __syncthreads();
    if (s_mode == 0) { if (tid == 0) { B.type = 0; B.csize = n; } return; }
    const u32 lit_hdr = s_lit_hdr, tree_len = s_tree_len, payload = s_payload0;
    const u32 content = lit_hdr + payload + 1;
    // ---- 7. zero the words we will OR into, write the headers, then every thread ORs its codes in
    u32 *slotw = (u32 *)slot;
    for (u32 i = tid; i < (content + 3) / 4 + 1; i += 256) slotw[i] = 0;
    __syncthreads();
    if (tid == 0) {
        // Literals_Section_Header: type 2 (compressed), size format by stream count / sizes
        if (nstreams == 1) { u32 v = 2 | (0 << 2) | (n << 4) | (payload << 14); slot[0] = (u8)v; slot[1] = (u8)(v >> 8); slot[2] = (u8)(v >> 16); }
        else if (lit_hdr == 4) { u32 v = 2 | (2 << 2) | (n << 4) | (payload << 18); slot[0] = (u8)v; slot[1] = (u8)(v >> 8); slot[2] = (u8)(v >> 16); slot[3] = (u8)(v >> 24); }
        else { u64 v = 2 | (3 << 2) | ((u64)n << 4) | ((u64)payload << 22); for (int i = 0; i < 5; i++) slot[i] = (u8)(v >> (8 * i)); }
        for (u32 i = 0; i < tree_len; i++) slot[lit_hdr + i] = tree_tmp[i];
        if (nstreams == 4) {
            u8 *jt = slot + lit_hdr + tree_len;
            for (int j = 0; j < 3; j++) { jt[2 * j] = (u8)s_stream_bytes[j]; jt[2 * j + 1] = (u8)(s_stream_bytes[j] >> 8); }
        }
        slot[content - 1] = 0;                                // Sequences_Section_Header: 0 sequences
        B.type = 2; B.csize = content;
    }
    __syncthreads();
    {
        u32 byte_base = lit_hdr + tree_len + (nstreams == 4 ? 6 : 0);
        for (u32 j = 0; j < k; j++) byte_base += s_stream_bytes[j];
        // symbols later in the stream sit at lower bit positions: my first bit = bits of all threads after me in my stream
        const u32 stream_bits = pre_bits[(k + 1) * tps] - pre_bits[k * tps];
        u32 bitpos = pre_bits[(k + 1) * tps] - (pre_bits[tid] + bits);
        const u64 abs0 = (u64)byte_base * 8;
        u64 acc = 0; u32 fill = 0;
        for (u32 i = t1; i > t0; i--) {
            u32 e = ctab[sb[zpad(i - 1)]];
            acc |= (u64)(e & 0xFFFF) << fill; fill += e >> 16;
            if (fill >= 32) {
                u64 ab = abs0 + bitpos; u32 wi = (u32)(ab >> 5), sh = (u32)(ab & 31); u32 v = (u32)acc;
                atomicOr(&slotw[wi], v << sh); if (sh) atomicOr(&slotw[wi + 1], v >> (32 - sh));
                acc >>= 32; fill -= 32; bitpos += 32;
            }
        }
        if (q == 0) { acc |= 1ull << fill; fill++; }          // end mark right above the first symbol's code
        (void)stream_bits;
        while (fill) {
            u32 take = fill > 32 ? 32 : fill;
            u64 ab = abs0 + bitpos; u32 wi = (u32)(ab >> 5), sh = (u32)(ab & 31); u32 v = (u32)(acc & (take == 32 ? 0xFFFFFFFFull : ((1ull << take) - 1)));
            if (v) { atomicOr(&slotw[wi], v << sh); if (sh && (v >> (32 - sh))) atomicOr(&slotw[wi + 1], v >> (32 - sh)); }
            acc >>= take; fill -= take; bitpos += take;
        }
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
        frame[5] = (u8)((ZWINDOW_LOG - 10) << 3);
    }
    if (threadIdx.x == 0) {
        u32 size_field = B.type == 1 ? B.n : B.csize;          // RLE: regenerated size
        u32 bh = (B.last & 1) | (B.type << 1) | (size_field << 3);
        dst[0] = (u8)bh; dst[1] = (u8)(bh >> 8); dst[2] = (u8)(bh >> 16);
    }
    const u8 *from = B.type == 0 ? B.src : A.slots + (size_t)blockIdx.x * ZSLOT;
    for (u32 i = threadIdx.x; i < B.csize; i += 256) dst[3 + i] = from[i];
}

static void zstd_compress_batch(Ctx &ctx, CudaExec &ex, ZEncBatch &b)
{
    const size_t ns = b.src.size();
    std::vector<ZEncBlock> blocks;
    b.first_block.assign(ns + 1, 0);
    for (size_t s = 0; s < ns; s++) {
        b.first_block[s] = (u32)blocks.size();
        u64 nb = b.n[s] ? (b.n[s] + ZBS - 1) / ZBS : 1;         // an empty stream is one empty raw last block
        for (u64 k = 0; k < nb; k++) {
            ZEncBlock e; memset(&e, 0, sizeof e);
            e.src = b.src[s] + k * ZBS; e.n = (u32)(b.n[s] - k * ZBS < ZBS ? b.n[s] - k * ZBS : ZBS); e.stream = (u32)s; e.last = k + 1 == nb;
            blocks.push_back(e);
        }
    }
    b.first_block[ns] = (u32)blocks.size();
    b.nblocks = (u32)blocks.size();
    b.d_blocks = ex.alloc<ZEncBlock>(b.nblocks);
    b.d_slots = ex.alloc<u8>((size_t)b.nblocks * ZSLOT);
    ex.upload(b.d_blocks, blocks.data(), sizeof(ZEncBlock) * b.nblocks);
    CUDA_TRY(cudaStreamSynchronize(ex.stream));                 // `blocks` is a local vector
    ZEncArgs A{b.d_blocks, b.d_slots};
    CUDA_TRY(cudaFuncSetAttribute(k_zenc_block, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ZSTAGE_BYTES));
    KLAUNCH(ex, "k_zenc_block", k_zenc_block<<<b.nblocks, 256, ZSTAGE_BYTES, ex.stream>>>(A));
    b.d_off = ex.alloc<u64>(b.nblocks + 2);
    const ZEncBlock *db = b.d_blocks;
    exclusive_scan(ex, [db] __device__ (size_t i) { return (u64)db[i].csize + 3; }, b.nblocks, b.d_off);
    // frame sizes: one small download of the offsets at the stream boundaries
    std::vector<u64> h_off(b.nblocks + 1);
    if (ns <= 8) {
        for (size_t s = 0; s <= ns; s++) ex.download(&h_off[b.first_block[s]], b.d_off + b.first_block[s], 8);
    } else ex.download(h_off.data(), b.d_off, (b.nblocks + 1) * 8);
    b.frame_size.assign(ns, 0); b.dest.assign(ns, nullptr);
    for (size_t s = 0; s < ns; s++) b.frame_size[s] = 6 + h_off[b.first_block[s + 1]] - h_off[b.first_block[s]];
    (void)ctx;
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
