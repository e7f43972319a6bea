// zstd_enc_hd.cuh — thread-serial pieces of the block-parallel zstd *encoder* as __host__ __device__ functions.
//
// Like zstd_hd.cuh on the decode side: every function here is executed by ONE GPU thread per work item (the Huffman
// code of one block; the whole LZ77 + entropy stage of one small block), the parallelism comes from running tens of
// thousands of items at once, and because the bodies are HD, tests/emu/emu_zenc.cpp runs the very same code on the
// CPU — the frames it produces are decoded by the oracle's from-spec decoder, by the unmodified libzstd 1.5.0 and by
// our own decoder — before any GPU time is spent.
//
// What is here:
//   * BitW, fse_compress_weights, zenc_huf_build: literal Huffman code of a block (used by k_zenc_tables for the
//     Huffman-only 32 KB blocks of the sequence / quality streams, and by the LZ path below)
//   * fse_normalize / fse_write_ncount / fse_build_enc / fse_put: a from-spec FSE encoder (spec "FSE", "FSE Table
//     Description"); reference counterparts compress/fse_compress.c:437 FSE_normalizeCount, :292 FSE_writeNCount,
//     :69 FSE_buildCTable_wksp, fse.h:531 FSE_encodeSymbol
//   * zlz_find: greedy single-hash LZ77 match finder with repeat-offset check and backward extension over one block
//     (the counterpart of compress/zstd_fast.c:186 ZSTD_compressBlock_fast, confined to the block)
//   * zlz_encode_block: a complete Compressed_Block — literals section (raw / RLE / Huffman, 1 or 4 streams) +
//     sequences section (per table: predefined / RLE / FSE_Compressed by estimated cost; compress/
//     zstd_compress_sequences.c:418 ZSTD_encodeSequences order of states and extra bits)
//
// Blocks stay independent of each other (that is what lets N GPUs concatenate their blocks into one frame):
// matches never reach before the block, no Repeat_Mode / treeless tables, and repeat-offset codes only ever name
// offsets this block itself has pushed — the history a block inherits is treated as unknown.
#pragma once
#include "zstd_hd.cuh"

namespace nafz {

// ---- forward LSB-first bit writer into a byte buffer
struct BitW {
    u8 *p; u32 cap; u32 pos; u64 acc; u32 fill; bool ok;
    HD void init(u8 *dst, u32 c) { p = dst; cap = c; pos = 0; acc = 0; fill = 0; ok = true; }
    HD void put(u32 v, u32 nb)                    // nb <= 31
    {
        acc |= (u64)(v & ((1u << nb) - 1)) << fill; fill += nb;
        while (fill >= 8) { if (pos < cap) p[pos++] = (u8)acc; else ok = false; acc >>= 8; fill -= 8; }
    }
    HD u32 finish_with_mark() { put(1, 1); if (fill) { if (pos < cap) p[pos++] = (u8)acc; else ok = false; fill = 0; acc = 0; } return pos; }
    HD u32 finish_aligned() { if (fill) { if (pos < cap) p[pos++] = (u8)acc; else ok = false; fill = 0; acc = 0; } return pos; }
};

// K consecutive elements into registers: the loads are independent of each other, so a thread waits for memory once per
// group instead of once per element (the thread-per-block kernels are bound by exactly that wait)
template <int K, class T> HD void ld_group(const T *p, T *t)
{
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int k = 0; k < K; k++) t[k] = p[k];
}

// FSE-compress the Huffman weights w[0..n).  Returns the number of bytes written (table description +
// bitstream), or 0 when not representable / not worthwhile.  Mirrors compress/huf_compress.c:76 HUF_compressWeights.
HDN inline u32 fse_compress_weights(const u8 *w, int n, u8 *dst, u32 cap)
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
        int bits = hibit((u32)remaining + 1) + 1;
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
        int nb = LOG - hibit((u32)p);
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

struct ZEncMeta {            // Huffman code of one block's literals
    u16 ctab[256];           // code | len << 12 (len 0 = symbol absent)
    u8  tree[132];           // Huffman tree description
    u8  mode;                // 0 raw (tree not describable / empty), 1 RLE (one distinct symbol), 2 Huffman
    u8  rle_sym;
    u16 tree_len;
};

// Length-limited canonical Huffman code + tree description from a 256-bin histogram (counts <= 65535).
// Counterpart of compress/huf_compress.c:513 HUF_buildCTable_wksp + :116 HUF_writeCTable_wksp.
template <class H> HDN inline void zenc_huf_build(const H *hist, ZEncMeta &M)
{
    M.mode = 2; M.tree_len = 0; M.rle_sym = 0;
    // symbols that occur, sorted by (count, symbol): shell sort of the keys count << 8 | symbol (a mask or length stream can
    // have all 256 symbols, and this thread is alone with its block)
    u8 sorted[256]; u16 cnt[256]; u32 nsym = 0;
    {
        u32 key[256];
        for (u32 s = 0; s < 256; s++) { const u32 h = hist[s]; if (h) key[nsym++] = (h << 8) | s; }
        const int gaps[6] = { 132, 57, 23, 10, 4, 1 };
        for (int gi = 0; gi < 6; gi++) {
            const u32 gap = (u32)gaps[gi];
            for (u32 i = gap; i < nsym; i++) {
                const u32 v = key[i]; u32 j = i;
                while (j >= gap && key[j - gap] > v) { key[j] = key[j - gap]; j -= gap; }
                key[j] = v;
            }
        }
        for (u32 i = 0; i < nsym; i++) { cnt[i] = (u16)(key[i] >> 8); sorted[i] = (u8)key[i]; }
    }
    for (u32 s = 0; s < 256; s++) M.ctab[s] = 0;
    if (nsym == 0) { M.mode = 0; return; }
    if (nsym == 1) { M.mode = 1; M.rle_sym = sorted[0]; return; }
    // code lengths: two-queue Huffman over the sorted counts, then limit to 11 bits
    u8 len_of[256], weight[257];
    for (u32 s = 0; s < 256; s++) { len_of[s] = 0; weight[s] = 0; }
    weight[256] = 0;
    u32 maxbits = 0;
    {
        u32 iw[256]; u16 lp[256], ip[256]; u8 idp[256];
        u32 li = 0, ii = 0;
        for (u32 m = 0; m + 1 < nsym; m++) {
            u32 wsum = 0;
            for (int t = 0; t < 2; t++) {
                const bool take_leaf = li < nsym && (ii >= m || cnt[li] <= iw[ii]);
                if (take_leaf) { wsum += cnt[li]; lp[li++] = (u16)m; } else { wsum += iw[ii]; ip[ii++] = (u16)m; }
            }
            iw[m] = wsum;
        }
        const u32 root = nsym - 2;
        idp[root] = 0;
        for (int m = (int)root - 1; m >= 0; m--) { const u32 d = idp[ip[m]] + 1u; idp[m] = (u8)(d > 60 ? 60 : d); }
        u32 num[40]; for (int i = 0; i < 40; i++) num[i] = 0;
        for (u32 i = 0; i < nsym; i++) { u32 d = idp[lp[i]] + 1u; if (d > 39) d = 39; num[d]++; }
        // Length limit: the format allows 11 bits; our decoder stages one 2^maxbits-entry table per block in shared memory, and 32
        // blocks share 32 KB (zstd_dec_cuda.cuh), so 9-bit codes are what keeps every lookup of a CTA on chip.  With at most 64
        // distinct symbols (4-bit sequence, qualities, ids) the limit costs well under 0.1 % of the block; beyond that 10 bits (half
        // the tables staged).
        const u32 MAXB = nsym <= 64 ? 9 : 10;
        for (u32 i = MAXB + 1; i < 40; i++) { num[MAXB] += num[i]; num[i] = 0; }
        u32 total = 0;
        for (u32 i = 1; i <= MAXB; i++) total += num[i] << (MAXB - i);
        while (total != (1u << MAXB)) {
            num[MAXB]--;
            for (u32 i = MAXB - 1; i > 0; i--) if (num[i]) { num[i]--; num[i + 1] += 2; break; }
            total--;
        }
        u32 idx = 0;
        for (u32 l = MAXB; l >= 1; l--) { if (num[l] && !maxbits) maxbits = l; for (u32 c = 0; c < num[l]; c++) len_of[sorted[idx++]] = (u8)l; }
    }
    // canonical codes exactly as the decoder rebuilds them: longer codes take the numerically lower values, equal
    // lengths in symbol order
    {
        u32 next[13];                                         // next code value (at full maxbits resolution) per length
        u32 acc = 0;
        u32 count_len[13]; for (int i = 0; i < 13; i++) count_len[i] = 0;
        for (u32 s = 0; s < 256; s++) count_len[len_of[s]]++;
        for (u32 l = maxbits; l >= 1; l--) { next[l] = acc; acc += count_len[l] << (maxbits - l); }
        for (u32 s = 0; s < 256; s++) {
            const u32 l = len_of[s];
            if (!l) continue;
            M.ctab[s] = (u16)((next[l] >> (maxbits - l)) | (l << 12));
            next[l] += 1u << (maxbits - l);
            weight[s] = (u8)(maxbits + 1 - l);
        }
    }
    // tree description
    int last_sym = 255; while (last_sym > 0 && !weight[last_sym]) last_sym--;
    const int nlisted = last_sym;                             // weights of symbols 0 .. last_sym-1; the last one is implied
    u8 tmp[132];
    const u32 fse = fse_compress_weights(weight, nlisted, tmp + 1, 127);
    const u32 direct = nlisted <= 128 ? 1 + (nlisted + 1) / 2 : 0xFFFFFFFFu;
    if (fse && fse < 128 && 1 + fse < direct) { tmp[0] = (u8)fse; M.tree_len = (u16)(1 + fse); }
    else if (direct != 0xFFFFFFFFu) {
        tmp[0] = (u8)(127 + nlisted);
        for (int i = 0; i < nlisted; i += 2) tmp[1 + i / 2] = (u8)((weight[i] << 4) | (i + 1 < nlisted ? weight[i + 1] : 0));
        M.tree_len = (u16)direct;
    } else { M.mode = 0; return; }                            // cannot describe the tree: raw
    for (u32 i = 0; i < M.tree_len; i++) M.tree[i] = tmp[i];
}

// ------------------------------------------------------------------ FSE encoder (sequence symbols)
// norm[s] >= 1 for every symbol that occurs, sum = 1 << log ("less than 1" probabilities are not produced).
HDN inline bool fse_normalize(const u32 *count, int nsym, u32 total, int log, short *norm)
{
    const int size = 1 << log;
    int sum = 0, npresent = 0;
    for (int s = 0; s < nsym; s++) {
        int v = 0;
        if (count[s]) { v = (int)(((u64)count[s] * (u32)size + total / 2) / total); if (v < 1) v = 1; npresent++; }
        norm[s] = (short)v; sum += v;
    }
    if (npresent < 2 || npresent > size) return false;
    // settle the rounding difference on the most probable symbols (largest first; nothing drops below 1)
    while (sum != size) {
        int best = -1;
        for (int s = 0; s < nsym; s++) if (norm[s] > (sum > size ? 1 : 0) && (best < 0 || norm[s] > norm[best])) best = s;
        if (best < 0) return false;
        int d = sum > size ? sum - size : size - sum;
        if (sum > size) { if (d > norm[best] - 1) d = norm[best] - 1; if (d > 1) d = (d + 1) / 2; norm[best] = (short)(norm[best] - d); sum -= d; }
        else { norm[best] = (short)(norm[best] + d); sum += d; }
    }
    return true;
}

// spec "FSE Table Description"; nsym = last symbol with a non-zero probability + 1
HDN inline void fse_write_ncount(BitW &bw, const short *norm, int nsym, int log)
{
    bw.put((u32)(log - 5), 4);
    int remaining = 1 << log, s = 0;
    while (remaining > 0 && s < nsym) {
        const int bits = hibit((u32)remaining + 1) + 1;
        const u32 lower = (1u << (bits - 1)) - 1, thresh = (1u << bits) - 1 - (u32)(remaining + 1);
        const u32 v = (u32)(norm[s] + 1);
        if (v < thresh) bw.put(v, (u32)bits - 1);
        else bw.put(v > lower ? v + thresh : v, (u32)bits);
        remaining -= norm[s];
        const bool zero = norm[s] == 0;
        s++;
        if (zero) {
            int run = 0;
            while (s < nsym && norm[s] == 0) { run++; s++; }
            while (run >= 3) { bw.put(3, 2); run -= 3; }
            bw.put((u32)run, 2);
        }
    }
    bw.finish_aligned();
}

// Encoding view of the decoder's table: spos[cum[s] + j] = position of the j-th cell (in position order) that decodes
// to symbol s.  tsym: 1 << log bytes of scratch.
HDN inline void fse_build_enc(const short *norm, int nsym, int log, u16 *spos, u16 *cum, u8 *tsym)
{
    const int size = 1 << log, step = (size >> 1) + (size >> 3) + 3, mask = size - 1;
    cum[0] = 0; for (int s = 0; s < nsym; s++) cum[s + 1] = (u16)(cum[s] + norm[s]);
    int pos = 0;
    for (int s = 0; s < nsym; s++) for (int i = 0; i < norm[s]; i++) { tsym[pos] = (u8)s; pos = (pos + step) & mask; }
    u16 occ[64]; for (int s = 0; s < 64; s++) occ[s] = 0;
    for (int p = 0; p < size; p++) { const int s = tsym[p]; spos[cum[s] + occ[s]++] = (u16)p; }
}

struct FseEnc {              // one symbol type of one block
    const short *norm; const u16 *cum; const u16 *spos; int log; u32 st; int mode; u32 rle_sym;
    HD void start(u32 sym) { if (mode != 1) st = spos[cum[sym]]; }
    HD void put(BitW &bw, u32 sym)
    {
        if (mode == 1) return;                                // RLE mode: no state bits
        const u32 p = (u32)norm[sym], y = st + (1u << log);
        int nb = log - hibit(p);
        u32 nn = y >> nb;
        if (nn < p) { nb--; nn = y >> nb; }
        bw.put(y, (u32)nb);
        st = spos[cum[sym] + (nn - p)];
    }
    HD void flush(BitW &bw) { if (mode != 1) bw.put(st, (u32)log); }
};

// ~ -256 * log2(p / 2^log): cost of one symbol in 1/256 bit (linear interpolation between powers of two)
HD u32 fse_cost256(u32 p, int log)
{
    const int hb = hibit(p);
    const u32 frac = ((p << 8) >> hb) - 256;                 // 0 .. 255: p = 2^hb * (1 + frac / 256)
    return (u32)((log - hb) << 8) - (frac + ((88 * frac * (256 - frac)) >> 16));      // log2(1 + x) ~ x + 0.344 x (1 - x)
}

// ------------------------------------------------------------------ LZ77 over one block
// 128 entries: on the streams this parse is for (ids, comments, lengths, mask in 8 KB blocks) a 128-, 256- or 1024-entry table
// give the same sizes within 1 % (emulation) -- candidates are recent -- and 256 B per thread instead of 2 KB
// lets 16 CTAs share an SM instead of 3.
static const u32 ZLZ_HLOG = 7, ZLZ_EMPTY = 0xFFFF;
static const u32 ZLZ_MAX_BLOCK = 32768;                       // positions are kept in u16

HD u32 zlz_read32(const u8 *p) { return (u32)p[0] | ((u32)p[1] << 8) | ((u32)p[2] << 16) | ((u32)p[3] << 24); }
HD u32 zlz_hash(u32 v) { return (v * 2654435761u) >> (32 - ZLZ_HLOG); }

struct ZLzSeqs { u16 *ll, *ml, *ov; u32 n; };                 // ov: Offset_Value (1..3 repeat codes, offset + 3 otherwise)

// repeat-offset history as the decoder will hold it, as far as this block knows it (k = number of known entries)
struct ZLzRep {
    u32 r[3]; u32 k;
    HD u32 code(u32 off, u32 ll)                              // Offset_Value for `off`; updates the history like the decoder
    {
        if (ll) {
            if (k >= 1 && off == r[0]) return 1;
            if (k >= 2 && off == r[1]) { r[1] = r[0]; r[0] = off; return 2; }
            if (k >= 3 && off == r[2]) { r[2] = r[1]; r[1] = r[0]; r[0] = off; return 3; }
        } else {
            if (k >= 2 && off == r[1]) { r[1] = r[0]; r[0] = off; return 1; }
            if (k >= 3 && off == r[2]) { r[2] = r[1]; r[1] = r[0]; r[0] = off; return 2; }
            if (k >= 1 && r[0] > 1 && off == r[0] - 1) { r[2] = r[1]; r[1] = r[0]; r[0] = off; if (k < 3) k++; return 3; }
        }
        r[2] = r[1]; r[1] = r[0]; r[0] = off; if (k < 3) k++;
        return off + 3;
    }
};

// Greedy parse.  htab: 1 << ZLZ_HLOG entries, entry e at htab[e * hstride] (the kernel interleaves the tables of a
// warp's lanes in shared memory: stride 32, bank = lane).  Literals go to lit[], sequences to S; the literals after
// the last match are appended to lit[] without a sequence.  Returns the number of literals.
HDN inline u32 zlz_find(const u8 *src, u32 n, u16 *htab, u32 hstride, u8 *lit, ZLzSeqs &S, u32 max_seq)
{
    for (u32 e = 0; e < (1u << ZLZ_HLOG); e++) htab[e * hstride] = (u16)ZLZ_EMPTY;
    ZLzRep rep; rep.r[0] = rep.r[1] = rep.r[2] = 0; rep.k = 0;
    u32 p = 0, anchor = 0, nlit = 0; S.n = 0;
    while (p + 4 <= n && S.n < max_seq) {
        const u32 v = zlz_read32(src + p);
        const u32 h = zlz_hash(v) * hstride;
        const u32 cand = htab[h];
        htab[h] = (u16)p;
        u32 off = 0;
        if (rep.k && p >= rep.r[0] && zlz_read32(src + p - rep.r[0]) == v) off = rep.r[0];
        else if (cand != ZLZ_EMPTY && zlz_read32(src + cand) == v) off = p - cand;
        if (!off) { p += 1 + ((p - anchor) >> 6); continue; }
        u32 ml = 4;
        while (p + ml + 4 <= n && zlz_read32(src + p + ml) == zlz_read32(src + p + ml - off)) ml += 4;
        while (p + ml < n && src[p + ml] == src[p + ml - off]) ml++;
        while (p > anchor && p > off && src[p - 1] == src[p - 1 - off]) { p--; ml++; }
        if (ml < 5 && !(rep.k && off == rep.r[0] && p > anchor)) { p += 1; continue; }   // a 4-byte match at a new offset costs more than its literals
        const u32 ll = p - anchor;
        {   // literal run: 8 bytes per step (loads before stores: one memory wait per step, not per byte)
            u32 i = 0;
            for (; i + 8 <= ll; i += 8) { u8 t[8]; for (int k = 0; k < 8; k++) t[k] = src[anchor + i + k]; for (int k = 0; k < 8; k++) lit[nlit + i + k] = t[k]; }
            for (; i < ll; i++) lit[nlit + i] = src[anchor + i];
        }
        nlit += ll;
        S.ll[S.n] = (u16)ll; S.ml[S.n] = (u16)ml; S.ov[S.n] = (u16)rep.code(off, ll); S.n++;
        p += ml; anchor = p;
        if (p >= 2 && p + 2 <= n) { const u32 q = p - 2; htab[zlz_hash(zlz_read32(src + q)) * hstride] = (u16)q; }
    }
    for (u32 i = anchor; i < n; i++) lit[nlit++] = src[i];
    return nlit;
}

HD u32 zlz_ll_code(u32 ll) { if (ll < 16) return ll; u32 c = 16; while (c < 35 && ll_base_of(c + 1) <= ll) c++; return c; }
HD u32 zlz_ml_code(u32 ml) { if (ml < 35) return ml - 3; u32 c = 32; while (c < 52 && ml_base_of(c + 1) <= ml) c++; return c; }

// scratch one block needs besides lit[] and the sequence arrays
struct ZLzWork {
    u16 *spos;               // 512 (LL) + 256 (OF) + 512 (ML) entries
    u8  *tsym;               // 512 bytes
    u8  *codes;              // 3 * max_seq bytes: LL, OF, ML code of every sequence
};

// Literals_Section for lit[0..nlit) at out; returns its size (never fails: raw literals always fit in nlit + 3).
HDN inline u32 zlz_put_literals(const u8 *lit, u32 nlit, u8 *out, u32 cap)
{
    auto raw_or_rle = [&](u32 type, u32 body) -> u32 {
        u32 h;
        if (nlit < 32) { out[0] = (u8)(type | (nlit << 3)); h = 1; }
        else if (nlit < 4096) { const u32 v = type | (1u << 2) | (nlit << 4); out[0] = (u8)v; out[1] = (u8)(v >> 8); h = 2; }
        else { const u32 v = type | (3u << 2) | (nlit << 4); out[0] = (u8)v; out[1] = (u8)(v >> 8); out[2] = (u8)(v >> 16); h = 3; }
        if (type == 1) out[h] = lit[0]; else for (u32 i = 0; i < body; i++) out[h + i] = lit[i];
        return h + body;
    };
    if (nlit == 0) return raw_or_rle(0, 0);
    u16 hist[256]; for (int s = 0; s < 256; s++) hist[s] = 0;
    {
        u32 i = 0;
        for (; i + 8 <= nlit; i += 8) { u8 t[8]; ld_group<8>(lit + i, t); for (int k = 0; k < 8; k++) hist[t[k]]++; }
        for (; i < nlit; i++) hist[lit[i]]++;
    }
    ZEncMeta M; zenc_huf_build(hist, M);
    if (M.mode == 1) return raw_or_rle(1, 1);
    if (M.mode != 2 || nlit < 8) return raw_or_rle(0, nlit);
    // exact stream sizes
    const u32 nstreams = nlit <= 1023 ? 1 : 4, seg = nstreams == 4 ? (nlit + 3) / 4 : nlit;
    u32 sbytes[4] = {0, 0, 0, 0}, payload = M.tree_len + (nstreams == 4 ? 6u : 0u);
    for (u32 k = 0; k < nstreams; k++) {
        const u32 a = k * seg, b = k == nstreams - 1 ? nlit : (a + seg < nlit ? a + seg : nlit);
        u32 bits = 0, i = a;
        for (; i + 8 <= b; i += 8) { u8 t[8]; ld_group<8>(lit + i, t); for (int k = 0; k < 8; k++) bits += M.ctab[t[k]] >> 12; }
        for (; i < b; i++) bits += M.ctab[lit[i]] >> 12;
        sbytes[k] = bits / 8 + 1; payload += sbytes[k];
    }
    const u32 lh = nstreams == 1 ? 3 : ((nlit <= 16383 && payload <= 16383) ? 4 : 5);
    const u32 raw_size = nlit + (nlit < 32 ? 1 : (nlit < 4096 ? 2 : 3));
    if ((nstreams == 1 && payload > 1023) || lh + payload >= raw_size || lh + payload > cap || (nstreams == 4 && nlit < 16)) return raw_or_rle(0, nlit);
    if (nstreams == 1) { const u32 v = 2 | (0 << 2) | (nlit << 4) | (payload << 14); out[0] = (u8)v; out[1] = (u8)(v >> 8); out[2] = (u8)(v >> 16); }
    else if (lh == 4) { const u32 v = 2 | (2 << 2) | (nlit << 4) | (payload << 18); out[0] = (u8)v; out[1] = (u8)(v >> 8); out[2] = (u8)(v >> 16); out[3] = (u8)(v >> 24); }
    else { const u64 v = 2 | (3 << 2) | ((u64)nlit << 4) | ((u64)payload << 22); for (int i = 0; i < 5; i++) out[i] = (u8)(v >> (8 * i)); }
    u32 at = lh;
    for (u32 i = 0; i < M.tree_len; i++) out[at++] = M.tree[i];
    if (nstreams == 4) { for (int j = 0; j < 3; j++) { out[at++] = (u8)sbytes[j]; out[at++] = (u8)(sbytes[j] >> 8); } }
    for (u32 k = 0; k < nstreams; k++) {
        const u32 a = k * seg, b = k == nstreams - 1 ? nlit : (a + seg < nlit ? a + seg : nlit);
        BitW bw; bw.init(out + at, sbytes[k]);
        u32 i = b;                                            // the last symbol sits at the lowest bits
        for (; i >= a + 8; i -= 8) { u8 t[8]; ld_group<8>(lit + i - 8, t); for (int k = 7; k >= 0; k--) { const u32 e = M.ctab[t[k]]; bw.put(e & 0xFFF, e >> 12); } }
        for (; i > a; i--) { const u32 e = M.ctab[lit[i - 1]]; bw.put(e & 0xFFF, e >> 12); }
        bw.finish_with_mark();
        at += sbytes[k];
    }
    return at;
}

// One symbol type of the sequences section: choose predefined / RLE / FSE_Compressed, write the table description,
// leave the encoder in E.  codes[0..n): the symbol of every sequence.  Returns the mode (0, 1, 2) or -1 on overflow.
HDN inline int zlz_put_table(const u8 *codes, u32 n, int nsym_max, int max_log, const short *predef, int predef_log,
                             short *norm, u16 *spos, u16 *cum, u8 *tsym, BitW &hw, FseEnc &E)
{
    u32 count[64]; for (int s = 0; s < 64; s++) count[s] = 0;
    int top = 0, npresent = 0;
    {
        u32 i = 0;
        for (; i + 8 <= n; i += 8) { u8 t[8]; ld_group<8>(codes + i, t); for (int k = 0; k < 8; k++) count[t[k]]++; }
        for (; i < n; i++) count[codes[i]]++;
    }
    for (int s = 0; s < nsym_max; s++) if (count[s]) { top = s; npresent++; }
    E.norm = norm; E.cum = cum; E.spos = spos; E.st = 0; E.rle_sym = 0;
    if (npresent == 1) {                                      // RLE mode: one byte, no state bits at all
        E.mode = 1; E.rle_sym = (u32)top; E.log = 0;
        hw.put((u32)top, 8);
        return 1;
    }
    // cost of the predefined distribution (usable when it gives every occurring symbol a probability)
    u32 cost_pre = 0; bool pre_ok = true;
    for (int s = 0; s <= top; s++) if (count[s]) {
        const short q = predef[s];
        if (q == 0) { pre_ok = false; break; }
        cost_pre += count[s] * fse_cost256(q < 0 ? 1u : (u32)q, predef_log);
    }
    // own table
    int log = hibit(n) - 1; { int need = hibit((u32)npresent - 1) + 2; if (log < need) log = need; }
    if (log < 5) log = 5;
    if (log > max_log) log = max_log;
    bool own_ok = fse_normalize(count, top + 1, n, log, norm);
    u32 cost_own = 0;
    if (own_ok) {
        for (int s = 0; s <= top; s++) if (count[s]) cost_own += count[s] * fse_cost256((u32)norm[s], log);
        cost_own += (u32)(4 + (top + 1) * (log / 2 + 2)) << 8;               // rough size of the table description
    }
    if (pre_ok && (!own_ok || cost_pre <= cost_own)) {
        for (int s = 0; s < nsym_max; s++) norm[s] = predef[s] < 0 ? (short)1 : predef[s];
        // the predefined tables have "less than 1" symbols, which sit at the END of the table: build the encoding view
        // from the decoder's own table builder so that both sides agree cell by cell
        E.mode = 0; E.log = predef_log;
        u32 table[64]; u16 next[64];
        fse_build_table(table, predef, nsym_max, predef_log, next);
        cum[0] = 0; for (int s = 0; s < nsym_max; s++) cum[s + 1] = (u16)(cum[s] + norm[s]);
        u16 occ[64]; for (int s = 0; s < 64; s++) occ[s] = 0;
        for (int p = 0; p < (1 << predef_log); p++) { const u32 s = table[p] & 0xFF; spos[cum[s] + occ[s]++] = (u16)p; }
        return 0;
    }
    if (!own_ok) return -1;
    E.mode = 2; E.log = log;
    fse_write_ncount(hw, norm, top + 1, log);
    fse_build_enc(norm, top + 1, log, spos, cum, tsym);
    return 2;
}

// Literals section + sequences section for a block of n bytes that was parsed into lit[0..nlit) and S (any match finder).
// Returns the Compressed_Block's size, or 0 when the block should be stored raw (no gain).
HDN inline u32 zlz_emit_block(u32 n, const u8 *lit, u32 nlit, const ZLzSeqs &S, u32 max_seq, ZLzWork W, u8 *out, u32 cap)
{
    u32 at = zlz_put_literals(lit, nlit, out, cap);
    if (at + 4 >= n) return 0;
    const u32 nseq = S.n;
    if (nseq == 0) { out[at++] = 0; return at < n ? at : 0; }
    if (nseq < 128) out[at++] = (u8)nseq;
    else if (nseq < 0x7F00) { out[at++] = (u8)((nseq >> 8) + 128); out[at++] = (u8)nseq; }
    else { out[at++] = 255; out[at++] = (u8)(nseq - 0x7F00); out[at++] = (u8)((nseq - 0x7F00) >> 8); }
    const u32 modes_at = at++;
    // codes
    u8 *cl = W.codes, *co = W.codes + max_seq, *cm = W.codes + 2 * max_seq;
    {
        u32 i = 0;
        for (; i + 4 <= nseq; i += 4) {
            u16 a[4], b[4], c[4]; ld_group<4>(S.ll + i, a); ld_group<4>(S.ov + i, b); ld_group<4>(S.ml + i, c);
            for (int k = 0; k < 4; k++) { cl[i + k] = (u8)zlz_ll_code(a[k]); co[i + k] = (u8)hibit(b[k]); cm[i + k] = (u8)zlz_ml_code(c[k]); }
        }
        for (; i < nseq; i++) { cl[i] = (u8)zlz_ll_code(S.ll[i]); co[i] = (u8)hibit(S.ov[i]); cm[i] = (u8)zlz_ml_code(S.ml[i]); }
    }
    SeqConsts C; seq_consts_init(C);
    short nl[36], no[32], nm[53]; u16 cuml[37], cumo[33], cumm[54];
    FseEnc EL, EO, EM;
    BitW hw; hw.init(out + at, cap - at);
    const int ml_ = zlz_put_table(cl, nseq, 36, 9, C.ll_norm, 6, nl, W.spos, cuml, W.tsym, hw, EL);
    const int mo_ = zlz_put_table(co, nseq, 29, 8, C.of_norm, 5, no, W.spos + 512, cumo, W.tsym, hw, EO);
    const int mm_ = zlz_put_table(cm, nseq, 53, 9, C.ml_norm, 6, nm, W.spos + 768, cumm, W.tsym, hw, EM);
    if (ml_ < 0 || mo_ < 0 || mm_ < 0 || !hw.ok) return 0;
    out[modes_at] = (u8)((ml_ << 6) | (mo_ << 4) | (mm_ << 2));
    at += hw.pos;
    // bitstream: sequences from the last to the first (the decoder reads it backwards)
    BitW bw; bw.init(out + at, cap - at);
    {
        const u32 i = nseq - 1;
        EM.start(cm[i]); EO.start(co[i]); EL.start(cl[i]);
        bw.put(S.ll[i] - ll_base_of(cl[i]), ll_bits_of(cl[i]));
        bw.put(S.ml[i] - ml_base_of(cm[i]), ml_bits_of(cm[i]));
        bw.put(S.ov[i] - (1u << co[i]), co[i]);
    }
    auto one = [&](u32 c_l, u32 c_o, u32 c_m, u32 ll, u32 ml, u32 ov) {
        EO.put(bw, c_o); EM.put(bw, c_m); EL.put(bw, c_l);
        bw.put(ll - ll_base_of(c_l), ll_bits_of(c_l));
        bw.put(ml - ml_base_of(c_m), ml_bits_of(c_m));
        bw.put(ov - (1u << c_o), c_o);
    };
    u32 i = nseq - 1;                                         // sequences i-1, i-2, ... 0 remain
    for (; i >= 4; i -= 4) {
        u8 x[4], y[4], z[4]; u16 a[4], b[4], c[4];
        ld_group<4>(cl + i - 4, x); ld_group<4>(co + i - 4, y); ld_group<4>(cm + i - 4, z);
        ld_group<4>(S.ll + i - 4, a); ld_group<4>(S.ml + i - 4, b); ld_group<4>(S.ov + i - 4, c);
        for (int k = 3; k >= 0; k--) one(x[k], y[k], z[k], a[k], b[k], c[k]);
    }
    for (; i > 0; i--) one(cl[i - 1], co[i - 1], cm[i - 1], S.ll[i - 1], S.ml[i - 1], S.ov[i - 1]);
    EM.flush(bw); EO.flush(bw); EL.flush(bw);
    bw.finish_with_mark();
    if (!bw.ok) return 0;
    at += bw.pos;
    return at < n ? at : 0;
}

// A complete Compressed_Block for src[0..n) at out (capacity cap >= n).  Returns its size, or 0 when the block should
// be stored raw (no gain) — the caller emits a Raw_Block (or an RLE_Block when *rle is set).
HDN inline u32 zlz_encode_block(const u8 *src, u32 n, bool use_lz, u16 *htab, u32 hstride, u8 *lit, ZLzSeqs S, u32 max_seq,
                                ZLzWork W, u8 *out, u32 cap, bool *rle)
{
    *rle = false;
    if (n == 0) return 0;
    { u32 i = 1; while (i < n && src[i] == src[0]) i++; if (i == n) { *rle = true; return 0; } }
    if (n < 16 || n > ZLZ_MAX_BLOCK) return 0;
    u32 nlit;
    if (use_lz) nlit = zlz_find(src, n, htab, hstride, lit, S, max_seq);
    else { S.n = 0; nlit = n; for (u32 i = 0; i < n; i++) lit[i] = src[i]; }
    return zlz_emit_block(n, lit, nlit, S, max_seq, W, out, cap);
}

}  // namespace nafz
