// naf_dec.cu — .naf -> text on the GPU.
//
// Replaces unnaf's decode path after the header is parsed:
//   load_ids / load_names / load_lengths / load_mask      unnaf/src/input.c:145-246
//   the ZSTD_decompressStream loops                       unnaf/src/output.c:640-650, input.c:352-440
//   init_tables + write_4bit_as_fasta (4-bit -> ASCII)    unnaf/src/utils.c:74, output.c:445
//   mask_dna_buffer (+32 over masked runs)                unnaf/src/output.c:295
//   print_dna_buffer_as_fasta / print_dna_split_into_lines / print_name   output.c:369,339,105
//   print_fastq                                           unnaf/src/output-fastq.c:100
//   print_dna (--seq), print_sequences, print_4bit, print_ids, print_names, print_charcount
//
// The reference streams 128 KB at a time through one core.  Here the whole file is resident in HBM:
// all streams are entropy-decoded together (zstd_dec.cuh), prefix sums give every record its place in
// the output text, and ONE kernel (k_write_text) then materialises the text: each thread produces 16
// consecutive output bytes, wherever they fall — header, wrapped sequence line, '+' line or quality.
#include <chrono>
#include "common.cuh"
#include "container.hpp"
#include "zstd_dec.cuh"
#include "zstd_dec_cuda.cuh"
#include "duo_plan.hpp"

namespace nafg {

// ------------------------------------------------------------------ small kernels

__global__ void k_scan_tiles(u64 *tile_sums, size_t ntiles, u64 *grand_total)
{
    __shared__ u64 sm[33];
    u64 carry = 0;
    for (size_t base = 0; base < ntiles; base += blockDim.x) {
        size_t i = base + threadIdx.x;
        u64 v = i < ntiles ? tile_sums[i] : 0, total;
        u64 p = block_excl_scan(v, &total, sm);
        if (i < ntiles) tile_sums[i] = carry + p;
        carry += total;
    }
    if (threadIdx.x == 0) *grand_total = carry;
}

// positions of the '\0' terminators of a string stream (ids / comments): end[r] = index of r-th zero
static const int ZT = 4096;
__global__ void k_zero_count(const u8 *s, u64 n, u64 *tile_counts)
{
    __shared__ u64 sm[33];
    u64 base = (u64)blockIdx.x * ZT + (u64)threadIdx.x * 16;
    u64 c = 0;
    for (int k = 0; k < 16; k++) if (base + k < n && s[base + k] == 0) c++;
    u64 total; block_excl_scan(c, &total, sm);
    if (threadIdx.x == 0) tile_counts[blockIdx.x] = total;
}
__global__ void k_zero_scatter(const u8 *s, u64 n, const u64 *tile_prefix, u32 *end, u64 max_records)
{
    __shared__ u64 sm[33];
    u64 base = (u64)blockIdx.x * ZT + (u64)threadIdx.x * 16;
    u64 c = 0;
    for (int k = 0; k < 16; k++) if (base + k < n && s[base + k] == 0) c++;
    u64 total; u64 r = block_excl_scan(c, &total, sm) + tile_prefix[blockIdx.x];
    for (int k = 0; k < 16; k++) if (base + k < n && s[base + k] == 0) { if (r < max_records) end[r] = (u32)(base + k); r++; }
}

// ------------------------------------------------------------------ text layout

struct TextArgs {
    // record structure
    u8  prefix;           // '>' / '@' / 0
    u8  with_name;        // name bytes present
    u8  name_nl;          // '\n' after the name part
    u8  has_ids, has_names, sep;
    u8  seq_present;      // sequence area present
    u8  seq_nl;           // 0: no newline after the sequence, 1: only when L > 0 (FASTA), 2: always
    u8  with_qual;        // "+\n" qual "\n"
    u8  packed;           // 4-bit sequence stream
    u8  upper;            // toupper() on raw sequence bytes (protein/text with --no-mask)
    u64 W;                // line width inside the sequence area, 0 = unlimited
    // streams
    const u8 *ids, *comm, *seq, *qual;
    const u32 *id_end, *cm_end;      // per record: index of its '\0'
    const u64 *L, *seq_start;        // per record: bases, first base index
    const u32 *maskbits;             // 1 bit per base or nullptr
    const u64 *out_start;            // per record: first output byte; [N] = total
    u64 N, total, total_bases;
    u64 rec0;                        // global index of local record 0 (record-range decode): per-record arrays other than out_start are global
    u64 tile_base;                   // k_write_text: CTA b writes tile tile_base + b (the text is written piece by piece behind an upload)
    u32 lut[4];                      // code_to_nuc as 16 bytes
    u8 *out;
};

struct RecInfo { u64 a, b, c, d, e; u64 L, sbase; u32 id_s, id_len, cm_s, cm_len; };

__device__ __forceinline__ u64 seq_area_len(const TextArgs &A, u64 L)
{
    if (!A.seq_present) return 0;
    u64 nl = A.seq_nl == 0 ? 0 : (A.seq_nl == 2 ? 1 : (L > 0));
    if (A.W > 0 && A.seq_nl == 1) nl = (L + A.W - 1) / A.W;
    return L + nl;
}
__device__ __forceinline__ u32 name_len_of(const TextArgs &A, u32 id_len, u32 cm_len)
{
    if (!A.with_name) return 0;
    if (A.has_ids && A.has_names) return id_len + (cm_len ? 1 + cm_len : 0);
    return A.has_ids ? id_len : cm_len;
}
__device__ __forceinline__ void load_rec(const TextArgs &A, u64 i, RecInfo &R)
{
    i += A.rec0;
    R.id_s = R.id_len = R.cm_s = R.cm_len = 0;
    if (A.with_name) {
        if (A.has_ids) { R.id_s = i ? A.id_end[i - 1] + 1 : 0; R.id_len = A.id_end[i] - R.id_s; }
        if (A.has_names) { R.cm_s = i ? A.cm_end[i - 1] + 1 : 0; R.cm_len = A.cm_end[i] - R.cm_s; }
    }
    R.L = A.seq_present ? A.L[i] : 0; R.sbase = A.seq_present ? A.seq_start[i] : 0;
    R.a = A.prefix ? 1 : 0;
    R.b = R.a + name_len_of(A, R.id_len, R.cm_len);
    R.c = R.b + A.name_nl;
    R.d = R.c + seq_area_len(A, R.L);
    R.e = R.d + (A.with_qual ? R.L + 3 : 0);
}
__host__ __device__ __forceinline__ u64 rec_text_size(u8 prefix, u32 name_len, u8 name_nl, u8 seq_present, u8 seq_nl, u8 with_qual, u64 W, u64 L)
{
    u64 s = (prefix ? 1 : 0) + name_len + name_nl;
    if (seq_present) {
        u64 nl = seq_nl == 0 ? 0 : (seq_nl == 2 ? 1 : (L > 0));
        if (W > 0 && seq_nl == 1) nl = (L + W - 1) / W;
        s += L + nl;
    }
    if (with_qual) s += L + 3;
    return s;
}

// 16 consecutive nibbles starting at base index `bi` of the packed stream -> u64 (low nibble first)
__device__ __forceinline__ u64 load_nibbles16(const u8 *seq, u64 bi)
{
    u64 off = bi >> 1;
    const u64 *al = (const u64 *)((uintptr_t)(seq + off) & ~(uintptr_t)7);
    u32 sh = (u32)((uintptr_t)(seq + off) & 7) * 8 + (u32)(bi & 1) * 4;
    u64 w0 = __ldg(al), w1 = __ldg(al + 1);
    return sh ? (w0 >> sh) | (w1 << (64 - sh)) : w0;
}
// 4 nibbles (16 bits) -> 4 ASCII bytes through the 16-entry table held in 4 registers
__device__ __forceinline__ u32 nib4_to_ascii(u32 x, const u32 lut[4])
{
    u32 sel = x & 0x7777;
    u32 lo = __byte_perm(lut[0], lut[1], sel), hi = __byte_perm(lut[2], lut[3], sel);
    u32 m = ((x >> 3) & 1) | (((x >> 7) & 1) << 8) | (((x >> 11) & 1) << 16) | (((x >> 15) & 1) << 24);
    m *= 0xFF;
    return (lo & ~m) | (hi & m);
}
__device__ __forceinline__ u32 bits4_to_case(u32 b) { return ((b & 1) | ((b & 2) << 7) | ((b & 4) << 14) | ((b & 8) << 21)) * 0x20; }
// 16 mask bits starting at base index bi
__device__ __forceinline__ u32 load_maskbits16(const u32 *mb, u64 bi)
{
    u64 w = bi >> 5; u32 sh = (u32)(bi & 31);
    u64 v = (u64)__ldg(mb + w) | ((u64)__ldg(mb + w + 1) << 32);
    return (u32)(v >> sh) & 0xFFFF;
}
// 16 raw bytes starting at p (any alignment); buffers are padded so the over-read is safe
__device__ __forceinline__ uint4 load_bytes16(const u8 *p)
{
    const u64 *al = (const u64 *)((uintptr_t)p & ~(uintptr_t)7);
    u32 sh = (u32)((uintptr_t)p & 7) * 8;
    u64 w0 = __ldg(al), w1 = __ldg(al + 1), w2 = __ldg(al + 2);
    u64 lo = sh ? (w0 >> sh) | (w1 << (64 - sh)) : w0, hi = sh ? (w1 >> sh) | (w2 << (64 - sh)) : w1;
    return make_uint4((u32)lo, (u32)(lo >> 32), (u32)hi, (u32)(hi >> 32));
}
__device__ __forceinline__ u32 upper4(u32 w)
{
    // per byte: if 'a' <= c <= 'z' then c - 32
    u32 r = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) { u32 c = (w >> (8 * k)) & 0xFF; if (c >= 'a' && c <= 'z') c -= 32; r |= c << (8 * k); }
    return r;
}
__device__ __forceinline__ u8 base_at(const TextArgs &A, u64 bi)
{
    u8 c;
    if (A.packed) {
        u8 byte = __ldg(A.seq + (bi >> 1));
        u32 code = (bi & 1) ? byte >> 4 : byte & 15;
        c = (u8)(A.lut[code >> 2] >> (8 * (code & 3)));
        if (A.maskbits && ((__ldg(A.maskbits + (bi >> 5)) >> (bi & 31)) & 1)) c += 32;
    } else {
        c = __ldg(A.seq + bi);
        if (A.upper && c >= 'a' && c <= 'z') c -= 32;
    }
    return c;
}

static const int WT_THREADS = 256, WT_ITERS = 4, WT_TILE = WT_THREADS * 16 * WT_ITERS;   // 16 KB of text per CTA: four 16-byte chunks per thread
static const int WT_MAXREC = 256;     // records staged in shared memory per tile (more than that: read them from HBM)

struct RecS { u64 out0, L, sbase; u32 id_s, id_len, cm_s, cm_len; };

__device__ __forceinline__ void rec_bounds(const TextArgs &A, const RecS &s, RecInfo &R)
{
    R.id_s = s.id_s; R.id_len = s.id_len; R.cm_s = s.cm_s; R.cm_len = s.cm_len; R.L = s.L; R.sbase = s.sbase;
    R.a = A.prefix ? 1 : 0;
    R.b = R.a + name_len_of(A, s.id_len, s.cm_len);
    R.c = R.b + A.name_nl;
    R.d = R.c + seq_area_len(A, s.L);
    R.e = R.d + (A.with_qual ? s.L + 3 : 0);
}
__device__ __forceinline__ void rec_fetch(const TextArgs &A, u64 i, RecS &s)
{
    s.out0 = A.out_start[i];
    i += A.rec0;
    s.id_s = s.id_len = s.cm_s = s.cm_len = 0;
    if (A.with_name) {
        if (A.has_ids) { s.id_s = i ? A.id_end[i - 1] + 1 : 0; s.id_len = A.id_end[i] - s.id_s; }
        if (A.has_names) { s.cm_s = i ? A.cm_end[i - 1] + 1 : 0; s.cm_len = A.cm_end[i] - s.cm_s; }
    }
    s.L = A.seq_present ? A.L[i] : 0; s.sbase = A.seq_present ? A.seq_start[i] : 0;
}

// The generic composer of one 16-byte chunk of text starting at q0: the chunk is assembled from at most a handful of
// segments (prefix char, id, separator, comment, newline, a run of bases, '+', a run of qualities ...).  Each segment
// contributes one unaligned 16-byte read (or one constant byte) shifted into place: no per-byte loop.
__device__ __noinline__ void wt_chunk_generic(const TextArgs &A, const RecS *recs, u64 first, u32 nrec, u32 nrec_total, u64 q0, u64 tile1)
{
    // record containing q0: last staged record with out0 <= q0; if that is the final staged one, more may follow in HBM
    u32 lo = 0, hi = nrec;                   // invariant: recs[lo].out0 <= q0
    while (hi - lo > 1) { u32 mid = (lo + hi) >> 1; if (recs[mid].out0 <= q0) lo = mid; else hi = mid; }
    u64 rec = first + lo;
    RecS cur = recs[lo];
    if (lo + 1 == nrec && rec + 1 < A.N && nrec_total >= (u32)WT_MAXREC) {
        // tile holds more records than staged: finish the search in global memory
        u64 glo = rec, ghi = A.N;
        while (ghi - glo > 1) { u64 mid = (glo + ghi) >> 1; if (A.out_start[mid] <= q0) glo = mid; else ghi = mid; }
        if (glo != rec) { rec = glo; rec_fetch(A, rec, cur); }
    }
    RecInfo R; rec_bounds(A, cur, R);
    u64 r = q0 - cur.out0;
    const int nvalid = tile1 - q0 >= 16 ? 16 : (int)(tile1 - q0);
    u64 olo = 0, ohi = 0;
    int j = 0;
    u32 li = lo;
    while (j < nvalid) {
        while (r >= R.e && rec + 1 < A.N) {
            rec++; li++;
            if (rec == first + li && li < nrec) cur = recs[li]; else rec_fetch(A, rec, cur);
            rec_bounds(A, cur, R); r = 0;
        }
        u64 slo = 0, shi = 0; u32 take = 1;             // segment bytes (first byte in the low end) and how many
        if (r < R.a) slo = A.prefix;
        else if (r < R.b) {
            const u32 k = (u32)(r - R.a);
            const u8 *src; u32 len;
            if (A.has_ids && k < R.id_len) { src = A.ids + R.id_s + k; len = R.id_len - k; }
            else if (A.has_ids && A.has_names) {
                if (k == R.id_len) { src = nullptr; len = 1; slo = A.sep; }
                else { const u32 kk = k - R.id_len - 1; src = A.comm + R.cm_s + kk; len = R.cm_len - kk; }
            } else { src = A.comm + R.cm_s + k; len = R.cm_len - k; }
            if (src) { const uint4 v = load_bytes16(src); slo = (u64)v.x | ((u64)v.y << 32); shi = (u64)v.z | ((u64)v.w << 32); }
            take = len;
        }
        else if (r < R.c) slo = '\n';
        else if (r < R.d) {
            const u64 s = r - R.c;
            u64 bi; u64 run;
            if (A.W > 0 && A.seq_nl == 1) {
                const u64 line = s / (A.W + 1), col = s - line * (A.W + 1);
                bi = line * A.W + col;
                run = (col == A.W || r + 1 == R.d) ? 0 : min(A.W - col, R.L - bi);
            } else { bi = s; run = s < R.L ? R.L - s : 0; }
            if (run == 0) slo = '\n';
            else {
                bi += R.sbase;
                if (A.packed) {
                    const u64 nib = load_nibbles16(A.seq, bi);
                    u32 w0 = nib4_to_ascii((u32)nib & 0xFFFF, A.lut), w1 = nib4_to_ascii((u32)(nib >> 16) & 0xFFFF, A.lut);
                    u32 w2 = nib4_to_ascii((u32)(nib >> 32) & 0xFFFF, A.lut), w3 = nib4_to_ascii((u32)(nib >> 48) & 0xFFFF, A.lut);
                    if (A.maskbits) {
                        const u32 mb = load_maskbits16(A.maskbits, bi);
                        w0 += bits4_to_case(mb & 15); w1 += bits4_to_case((mb >> 4) & 15); w2 += bits4_to_case((mb >> 8) & 15); w3 += bits4_to_case((mb >> 12) & 15);
                    }
                    slo = (u64)w0 | ((u64)w1 << 32); shi = (u64)w2 | ((u64)w3 << 32);
                } else {
                    uint4 v = load_bytes16(A.seq + bi);
                    if (A.upper) { v.x = upper4(v.x); v.y = upper4(v.y); v.z = upper4(v.z); v.w = upper4(v.w); }
                    slo = (u64)v.x | ((u64)v.y << 32); shi = (u64)v.z | ((u64)v.w << 32);
                }
                take = run > 16 ? 16u : (u32)run;
            }
        }
        else {
            const u64 t = r - R.d;
            if (t == 0) slo = '+';
            else if (t == 1 || t >= 2 + R.L) slo = '\n';
            else {
                const uint4 v = load_bytes16(A.qual + R.sbase + (t - 2));
                slo = (u64)v.x | ((u64)v.y << 32); shi = (u64)v.z | ((u64)v.w << 32);
                const u64 run = R.L - (t - 2);
                take = run > 16 ? 16u : (u32)run;
            }
        }
        if (take > (u32)(16 - j)) take = 16 - j;
        // keep `take` bytes, shift them to byte position j, merge
        if (take < 16) {
            if (take >= 8) shi &= take == 8 ? 0ull : ((1ull << (8 * (take - 8))) - 1);
            else { shi = 0; slo &= (1ull << (8 * take)) - 1; }
        }
        if (j) {
            if (j < 8) { shi = (shi << (8 * j)) | (slo >> (64 - 8 * j)); slo <<= 8 * j; }
            else { shi = slo << (8 * (j - 8)); slo = 0; }
        }
        olo |= slo; ohi |= shi;
        j += (int)take; r += take;
    }
    if (nvalid == 16) *(uint4 *)(A.out + q0) = make_uint4((u32)olo, (u32)(olo >> 32), (u32)ohi, (u32)(ohi >> 32));
    else for (int k = 0; k < nvalid; k++) A.out[q0 + k] = (u8)((k < 8 ? olo >> (8 * k) : ohi >> (8 * (k - 8))));
}

// Each CTA: find the record containing its first byte (32-ary search by warp 0), stage the descriptors of the records
// that intersect the tile in shared memory, then
//   pass 1  every thread looks at its 16-byte chunks: a chunk that lies inside one run of bases or one run of
//           qualities (about three quarters of them) is produced on the spot; the others are only listed
//   pass 2  the listed chunks are shared out again over all threads, so that the long generic composer runs in
//           full warps instead of a few lanes per warp
// record holding the first byte of every tile (+ one past the end): one binary search per tile, all tiles at once
__global__ void k_tile_first(const u64 *out_start, u64 N, u64 ntiles, u32 *tile_first)
{
    const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t > ntiles) return;
    const u64 q = t * WT_TILE;
    u64 lo = 0, hi = N;                          // invariant: out_start[lo] <= q < out_start[hi]   (out_start[N] = total > q unless t == ntiles)
    while (hi - lo > 1) { const u64 mid = (lo + hi) >> 1; if (out_start[mid] <= q) lo = mid; else hi = mid; }
    tile_first[t] = (u32)lo;
}

__global__ void __launch_bounds__(WT_THREADS, 5) k_write_text(const __grid_constant__ TextArgs A, const u32 *tile_first)
{
    __shared__ RecS recs[WT_MAXREC];
    __shared__ u32 s_nslow;
    __shared__ u16 slow[WT_THREADS * WT_ITERS];
    const u64 tile = A.tile_base + blockIdx.x;
    const u64 tile0 = tile * WT_TILE;
    const u64 tile1 = tile0 + WT_TILE < A.total ? tile0 + WT_TILE : A.total;
    // records first .. last intersect this tile (last = the record holding the first byte of the next tile)
    const u64 first = tile_first[tile];
    const u32 s_nrec = tile_first[tile + 1] - (u32)first + 1;
    if (threadIdx.x == 0) s_nslow = 0;
    if (threadIdx.x < s_nrec && threadIdx.x < WT_MAXREC) rec_fetch(A, first + threadIdx.x, recs[threadIdx.x]);
    __syncthreads();
    const u32 nrec_total = s_nrec;
    const u32 nrec = nrec_total < (u32)WT_MAXREC ? nrec_total : (u32)WT_MAXREC;     // staged records: first .. first+nrec-1
    const bool wrapped = A.W > 0 && A.seq_nl == 1;
    const u32 W32 = (u32)A.W;

    for (int it = 0; it < WT_ITERS; it++) {
        const u64 q0 = tile0 + (u64)it * (WT_THREADS * 16) + (u64)threadIdx.x * 16;
        bool is_slow = false;
        if (q0 < tile1) {
            is_slow = true;
            u32 lo = 0, hi = nrec;
            while (hi - lo > 1) { u32 mid = (lo + hi) >> 1; if (recs[mid].out0 <= q0) lo = mid; else hi = mid; }
            // (when the tile holds more records than were staged, chunks that map to the last staged record go the generic way)
            if (q0 + 16 <= tile1 && !(lo + 1 == nrec && nrec_total >= (u32)WT_MAXREC)) {
                const RecS *c = &recs[lo];
                const u64 L = c->L, r = q0 - c->out0;
                const u64 cpos = (A.prefix ? 1 : 0) + name_len_of(A, c->id_len, c->cm_len) + A.name_nl;     // start of the sequence area
                if (A.seq_present && r >= cpos) {
                    const u64 s = r - cpos;
                    u64 bi = ~0ull;                                // first base of the chunk if it is one run of 16 bases
                    if (!wrapped) { if (s + 16 <= L) bi = s; }
                    else if (s < 0xFFFFFFFFull && A.W < 0xFFFFFFFFull) {
                        const u32 line = (u32)s / (W32 + 1), col = (u32)s - line * (W32 + 1);
                        const u64 b = (u64)line * W32 + col;
                        if (col + 16 <= W32 && b + 16 <= L) bi = b;
                    }
                    if (bi != ~0ull) {
                        bi += c->sbase;
                        uint4 v;
                        if (A.packed) {
                            const u64 nib = load_nibbles16(A.seq, bi);
                            v.x = nib4_to_ascii((u32)nib & 0xFFFF, A.lut); v.y = nib4_to_ascii((u32)(nib >> 16) & 0xFFFF, A.lut);
                            v.z = nib4_to_ascii((u32)(nib >> 32) & 0xFFFF, A.lut); v.w = nib4_to_ascii((u32)(nib >> 48) & 0xFFFF, A.lut);
                            if (A.maskbits) {
                                const u32 mb = load_maskbits16(A.maskbits, bi);
                                v.x += bits4_to_case(mb & 15); v.y += bits4_to_case((mb >> 4) & 15); v.z += bits4_to_case((mb >> 8) & 15); v.w += bits4_to_case((mb >> 12) & 15);
                            }
                        } else {
                            v = load_bytes16(A.seq + bi);
                            if (A.upper) { v.x = upper4(v.x); v.y = upper4(v.y); v.z = upper4(v.z); v.w = upper4(v.w); }
                        }
                        *(uint4 *)(A.out + q0) = v;
                        is_slow = false;
                    } else if (A.with_qual) {
                        const u64 d = cpos + seq_area_len(A, L);       // start of the "+\n" qual "\n" area
                        if (r >= d + 2 && r - d - 2 + 16 <= L) {
                            *(uint4 *)(A.out + q0) = load_bytes16(A.qual + c->sbase + (r - d - 2));
                            is_slow = false;
                        }
                    }
                }
            }
        }
        // list the chunks left for pass 2 (one shared-memory atomic per warp)
        const unsigned m = __ballot_sync(0xFFFFFFFFu, is_slow);
        if (m) {
            const unsigned lane = threadIdx.x & 31;
            u32 base = 0;
            if (lane == 0) base = atomicAdd(&s_nslow, (u32)__popc(m));
            base = __shfl_sync(0xFFFFFFFFu, base, 0);
            if (is_slow) slow[base + __popc(m & ((1u << lane) - 1))] = (u16)(it * WT_THREADS + threadIdx.x);
        }
    }
    __syncthreads();
    const u32 nslow = s_nslow;
    for (u32 k = threadIdx.x; k < nslow; k += WT_THREADS)
        wt_chunk_generic(A, recs, first, nrec, nrec_total, tile0 + (u64)slow[k] * 16, tile1);
}

// ------------------------------------------------------------------ the text writer, second formulation: compose in shared memory
// k_write_text above decides per 16-byte chunk of OUTPUT what it holds; a third of the chunks straddle a boundary (header,
// "+", line ends) and go through a long generic composer at a few lanes per warp (ncu, profiles/r2l: 1.47 warp instructions
// per byte of text, 16 of 32 lanes active).  Here the unit of work is a PIECE of the record layout instead -- at most 64
// bytes of a header, of one sequence line, of a quality string -- and one LANE owns a piece: it finds its record once, then
// copies 16 bytes at a time into a shared-memory image of the tile (aligned words of the image; bases through the 4-bit
// table, 16 per step; a few edge bytes one by one).  Pieces are enumerated kind by kind, so the lanes of a warp mostly do
// the same thing to different pieces.  The image then goes to HBM as aligned 16-byte stores.
// Layout rules as in rec_text_size / rec_bounds (output.c:369-406, output-fastq.c:100).
static const u32 CT_P = 64;                 // bytes per piece
static const u32 CT_PH = 16;                // bytes per piece of a header area (composed byte by byte: kept short, they set the time of the slowest lane)
struct CtRec { u64 s_lo; u32 n_h, h_lo, n_s, q_lo, n_q, ppl; };   // a record's pieces that intersect the tile: header, sequence, quality area

// image bytes [lo, hi) (tile-relative) <- src[0 .. hi - lo), src a global pointer of any alignment; one lane
__device__ __forceinline__ void ct_lane_bytes(u8 *img, u32 lo, u32 hi, const u8 *src, bool upper)
{
    while (lo < hi && (lo & 3)) { u8 c = __ldg(src++); if (upper && c >= 'a' && c <= 'z') c -= 32; img[lo++] = c; }
    u32 *dw = (u32 *)(img + lo);
    u32 nw = (hi - lo) >> 2;
    while (nw >= 4) {
        uint4 v = load_bytes16(src);
        if (upper) { v.x = upper4(v.x); v.y = upper4(v.y); v.z = upper4(v.z); v.w = upper4(v.w); }
        dw[0] = v.x; dw[1] = v.y; dw[2] = v.z; dw[3] = v.w;
        dw += 4; src += 16; lo += 16; nw -= 4;
    }
    while (lo < hi) { u8 c = __ldg(src++); if (upper && c >= 'a' && c <= 'z') c -= 32; img[lo++] = c; }
}
// image bytes [lo, hi) <- bases b0 .. b0 + (hi - lo); one lane
__device__ __forceinline__ void ct_lane_bases(const TextArgs &A, u8 *img, u32 lo, u32 hi, u64 b0)
{
    if (!A.packed) { ct_lane_bytes(img, lo, hi, A.seq + b0, A.upper != 0); return; }
    while (lo < hi && (lo & 3)) img[lo++] = base_at(A, b0++);
    u32 *dw = (u32 *)(img + lo);
    u32 nw = (hi - lo) >> 2;
    while (nw >= 4) {
        const u64 nib = load_nibbles16(A.seq, b0);
        u32 w0 = nib4_to_ascii((u32)nib & 0xFFFF, A.lut), w1 = nib4_to_ascii((u32)(nib >> 16) & 0xFFFF, A.lut);
        u32 w2 = nib4_to_ascii((u32)(nib >> 32) & 0xFFFF, A.lut), w3 = nib4_to_ascii((u32)(nib >> 48) & 0xFFFF, A.lut);
        if (A.maskbits) {
            const u32 mb = load_maskbits16(A.maskbits, b0);
            w0 += bits4_to_case(mb & 15); w1 += bits4_to_case((mb >> 4) & 15); w2 += bits4_to_case((mb >> 8) & 15); w3 += bits4_to_case((mb >> 12) & 15);
        }
        dw[0] = w0; dw[1] = w1; dw[2] = w2; dw[3] = w3;
        dw += 4; b0 += 16; lo += 16; nw -= 4;
    }
    while (lo < hi) img[lo++] = base_at(A, b0++);
}
// one byte of a record's header area [0, c): prefix, id, separator, comment, newline
__device__ __forceinline__ u8 ct_header_byte(const TextArgs &A, const RecInfo &R, u32 r)
{
    if (r < R.a) return A.prefix;
    if (r >= R.b) return '\n';
    const u32 k = r - (u32)R.a;
    if (A.has_ids && k < R.id_len) return __ldg(A.ids + R.id_s + k);
    if (A.has_ids && A.has_names) { if (k == R.id_len) return A.sep; return __ldg(A.comm + R.cm_s + (k - R.id_len - 1)); }
    return __ldg(A.comm + R.cm_s + k);
}

__global__ void __launch_bounds__(WT_THREADS, 5) k_compose_text(const __grid_constant__ TextArgs A, const u32 *tile_first)
{
    __shared__ __align__(16) u8 img[WT_TILE + 16];
    __shared__ RecS recs[WT_MAXREC];
    __shared__ CtRec cr[WT_MAXREC];
    __shared__ u32 pre_h[WT_MAXREC + 1], pre_s[WT_MAXREC + 1], pre_q[WT_MAXREC + 1];
    __shared__ u64 sm[33];
    const u32 tid = threadIdx.x;
    const u64 tile = A.tile_base + blockIdx.x;
    const u64 tile0 = tile * WT_TILE;
    const u64 tile1 = tile0 + WT_TILE < A.total ? tile0 + WT_TILE : A.total;
    const u64 first = tile_first[tile];
    const u32 nrec_total = tile_first[tile + 1] - (u32)first + 1;
    if (nrec_total > (u32)WT_MAXREC) {
        // more records in 16 KB than the tables hold (records of a few dozen bytes): chunk by chunk through the generic composer
        if (tid < WT_MAXREC) rec_fetch(A, first + tid, recs[tid]);
        __syncthreads();
        for (u64 q0 = tile0 + 16ull * tid; q0 < tile1; q0 += 16ull * WT_THREADS) wt_chunk_generic(A, recs, first, WT_MAXREC, nrec_total, q0, tile1);
        return;
    }
    const u32 nrec = nrec_total;
    const bool wrapped = A.W > 0 && A.seq_nl == 1;
    u32 ch = 0, cs = 0, cq = 0;
    if (tid < nrec) {
        rec_fetch(A, first + tid, recs[tid]);
        RecInfo R; rec_bounds(A, recs[tid], R);
        const u64 o = recs[tid].out0;
        CtRec c; c.s_lo = 0; c.n_h = c.h_lo = c.n_s = c.q_lo = c.n_q = 0; c.ppl = 1;
        // header area [0, c): pieces of CT_PH bytes
        if (R.c > 0 && o < tile1 && o + R.c > tile0) {
            const u64 lo = tile0 > o ? (tile0 - o) / CT_PH : 0;
            u64 hi = (R.c + CT_PH - 1) / CT_PH;
            if (o + R.c > tile1) hi = (tile1 - o + CT_PH - 1) / CT_PH;
            c.h_lo = (u32)lo; c.n_h = (u32)(hi - lo);
        }
        // sequence area [c, d).  wrapped: lines of W bases, each followed by '\n' (stride W + 1), a line cut into ppl pieces;
        // else ONE line of L bases, the area's single trailing '\n' (if any) written by its last piece
        if (A.seq_present && R.d > R.c) {
            const u64 a0 = o + R.c, a1 = o + R.d;
            if (a0 < tile1 && a1 > tile0) {
                if (wrapped) {
                    const u64 stride = A.W + 1, nline = (R.L + A.W - 1) / A.W, ppl = (A.W + CT_P - 1) / CT_P;
                    u64 lo = tile0 > a0 ? (tile0 - a0) / stride : 0, hi = (tile1 - a0 + stride - 1) / stride;
                    if (hi > nline) hi = nline;
                    if (lo >= nline) lo = nline - 1;
                    if (ppl > 0xFFFFu || (hi - lo) * ppl > 0x7FFFFFFFull) { c.ppl = 0; }      // absurd widths: generic composer (below)
                    else { c.ppl = (u32)ppl; c.s_lo = lo * ppl; c.n_s = (u32)((hi > lo ? hi - lo : 1) * ppl); }
                } else {
                    const u64 np = R.L ? (R.L + CT_P - 1) / CT_P : 1;     // L == 0 with a newline: one piece that is only the '\n'
                    u64 lo = tile0 > a0 ? (tile0 - a0) / CT_P : 0, hi = (tile1 - a0 + CT_P - 1) / CT_P;
                    if (hi > np) hi = np;
                    if (lo >= np) lo = np - 1;
                    c.s_lo = lo; c.n_s = (u32)(hi > lo ? hi - lo : 1);
                }
            }
        }
        // quality area [d, e): "+\n" (piece 0), then pieces of CT_P qualities, the last one followed by '\n'
        if (A.with_qual) {
            const u64 a0 = o + R.d, a1 = o + R.e;
            if (a0 < tile1 && a1 > tile0) {
                const u64 np = 1 + (R.L ? (R.L + CT_P - 1) / CT_P : 1);
                u64 lo = 0, hi = np;
                if (tile0 > a0 + 2) lo = 1 + (tile0 - a0 - 2) / CT_P;
                if (tile1 < a1) { hi = tile1 <= a0 + 2 ? 1 : 1 + (tile1 - a0 - 2 + CT_P - 1) / CT_P; if (hi > np) hi = np; }
                if (lo >= np) lo = np - 1;
                c.q_lo = (u32)lo; c.n_q = (u32)(hi > lo ? hi - lo : 1);
            }
        }
        cr[tid] = c;
        ch = c.n_h; cs = c.n_s; cq = c.n_q;
    }
    // three scans packed into one: header pieces (bits 0-20), sequence pieces (21-41), quality pieces (42-62) -- a tile holds
    // at most a few hundred of each (a record's pieces are clipped to the tile; absurdly narrow lines fall back below)
    const bool fits = ch < (1u << 13) && cs < (1u << 13) && cq < (1u << 13);
    const u32 bad = __syncthreads_or(!fits || (tid < nrec && cr[tid].ppl == 0));
    if (bad) {
        for (u64 q0 = tile0 + 16ull * tid; q0 < tile1; q0 += 16ull * WT_THREADS) wt_chunk_generic(A, recs, first, nrec, nrec_total, q0, tile1);
        return;
    }
    u64 total; const u64 pre = block_excl_scan((u64)ch | ((u64)cs << 21) | ((u64)cq << 42), &total, sm);
    if (tid < nrec) { pre_h[tid] = (u32)(pre & 0x1FFFFF); pre_s[tid] = (u32)((pre >> 21) & 0x1FFFFF); pre_q[tid] = (u32)(pre >> 42); }
    if (tid == 0) { pre_h[nrec] = (u32)(total & 0x1FFFFF); pre_s[nrec] = (u32)((total >> 21) & 0x1FFFFF); pre_q[nrec] = (u32)(total >> 42); }
    __syncthreads();
    const u32 nh = pre_h[nrec], ns = pre_s[nrec], nq = pre_q[nrec];
    auto find = [&](const u32 *p, u32 x) { u32 lo = 0, hi = nrec; while (hi - lo > 1) { const u32 mid = (lo + hi) >> 1; if (p[mid] <= x) lo = mid; else hi = mid; } return lo; };
    // The three kinds are enumerated one after the other and dealt out round-robin ACROSS the kinds: the first quality piece goes
    // to the thread after the one that took the last sequence piece, and so on -- every loop starting at thread 0 left the upper
    // warps of a CTA idle behind a barrier (a tile has about 120 + 140 + 60 pieces for 256 threads; ncu: a third of all stall
    // samples sat on that barrier).
    // ---- sequence pieces (the bulk)
    for (u32 g = tid; g < ns; g += WT_THREADS) {
        const u32 r = find(pre_s, g);
        const RecS &rs = recs[r]; const CtRec c = cr[r];
        RecInfo R; rec_bounds(A, rs, R);
        const u64 idx = c.s_lo + (g - pre_s[r]);
        u64 x0, b0, nb; bool nl;
        if (wrapped) {
            const u64 line = idx / c.ppl, sub = idx - line * c.ppl;
            const u64 lb = line * A.W, ll = R.L - lb < A.W ? R.L - lb : A.W;             // the line's bases
            const u64 s0 = sub * CT_P;
            nb = ll > s0 ? (ll - s0 < CT_P ? ll - s0 : CT_P) : 0;
            b0 = lb + s0; x0 = rs.out0 + R.c + line * (A.W + 1) + s0;
            nl = nb > 0 && s0 + nb == ll;                        // the piece that holds the line's last base writes its '\n' (every line has a base)
        } else {
            b0 = idx * CT_P; nb = R.L > b0 ? (R.L - b0 < CT_P ? R.L - b0 : CT_P) : 0;
            x0 = rs.out0 + R.c + b0;
            nl = b0 + nb >= R.L && R.d - R.c > R.L;
        }
        const u64 x1 = x0 + nb;
        const u64 y0 = x0 > tile0 ? x0 : tile0, y1 = x1 < tile1 ? x1 : tile1;
        if (y1 > y0) ct_lane_bases(A, img, (u32)(y0 - tile0), (u32)(y1 - tile0), rs.sbase + b0 + (y0 - x0));
        if (nl && x1 >= tile0 && x1 < tile1) img[x1 - tile0] = '\n';
    }
    // ---- quality-area pieces
    for (u32 g = (tid + WT_THREADS - ns % WT_THREADS) % WT_THREADS; g < nq; g += WT_THREADS) {
        const u32 r = find(pre_q, g);
        const RecS &rs = recs[r]; const CtRec c = cr[r];
        RecInfo R; rec_bounds(A, rs, R);
        const u64 a0 = rs.out0 + R.d, idx = c.q_lo + (g - pre_q[r]);
        if (idx == 0) {
            if (a0 >= tile0 && a0 < tile1) img[a0 - tile0] = '+';
            if (a0 + 1 >= tile0 && a0 + 1 < tile1) img[a0 + 1 - tile0] = '\n';
            continue;
        }
        const u64 b0 = (idx - 1) * CT_P, nb = R.L > b0 ? (R.L - b0 < CT_P ? R.L - b0 : CT_P) : 0;
        const u64 x0 = a0 + 2 + b0, x1 = x0 + nb;
        const u64 y0 = x0 > tile0 ? x0 : tile0, y1 = x1 < tile1 ? x1 : tile1;
        if (y1 > y0) ct_lane_bytes(img, (u32)(y0 - tile0), (u32)(y1 - tile0), A.qual + rs.sbase + b0 + (y0 - x0), false);
        if (b0 + nb >= R.L && x1 >= tile0 && x1 < tile1) img[x1 - tile0] = '\n';
    }
    // ---- header pieces
    for (u32 g = (tid + WT_THREADS - (ns + nq) % WT_THREADS) % WT_THREADS; g < nh; g += WT_THREADS) {
        const u32 r = find(pre_h, g);
        const RecS &rs = recs[r]; const CtRec c = cr[r];
        RecInfo R; rec_bounds(A, rs, R);
        const u64 idx = c.h_lo + (g - pre_h[r]);
        const u64 x0 = rs.out0 + idx * CT_PH, x1 = x0 + CT_PH < rs.out0 + R.c ? x0 + CT_PH : rs.out0 + R.c;
        const u64 y0 = x0 > tile0 ? x0 : tile0, y1 = x1 < tile1 ? x1 : tile1;
        for (u64 q = y0; q < y1; q++) img[q - tile0] = ct_header_byte(A, R, (u32)(q - rs.out0));
    }
    __syncthreads();
    // the image -> HBM (A.out + tile0 is 16-byte aligned: tiles are 16 KB and the arena hands out 256-byte aligned memory)
    const u32 T = (u32)(tile1 - tile0);
    const uint4 *iv = (const uint4 *)img; uint4 *ov = (uint4 *)(A.out + tile0);
    for (u32 k = tid; k < T / 16; k += WT_THREADS) ov[k] = iv[k];
    for (u32 k = (T & ~15u) + tid; k < T; k += WT_THREADS) A.out[tile0 + k] = img[k];
}

// histogram of the produced text (unnaf --charcount, output.c:544)
__global__ void k_charcount(const u8 *text, u64 n, unsigned long long *counts)
{
    __shared__ u32 h[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) h[i] = 0;
    __syncthreads();
    u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) atomicAdd(&h[text[i]], 1u);
    __syncthreads();
    for (int i = threadIdx.x; i < 256; i += blockDim.x) if (h[i]) atomicAdd(&counts[i], (unsigned long long)h[i]);
}

// ------------------------------------------------------------------ decode orchestration

static inline u64 align256(u64 v) { return (v + 255) & ~255ull; }

// d_naf: the whole file on the device; h_naf: the same bytes on the host (header + block walk)
static DecodeOut decode_impl(Ctx &ctx, CudaExec &ex, const u8 *d_naf, const u8 *h_naf, size_t n, const nafgpu_dec_opts &o, bool allow_index, bool *used_index);

// A file whose streams are fine but whose block index is damaged must still decode (the reference never looks at the index):
// if a call that relied on the index fails, it is repeated once with the header walk.
DecodeOut decode_on_device(Ctx &ctx, CudaExec &ex, const u8 *d_naf, const u8 *h_naf, size_t n, const nafgpu_dec_opts &o)
{
    bool used_index = false;
    const Arena::Mark mk = ex.arena->mark();
    try { return decode_impl(ctx, ex, d_naf, h_naf, n, o, true, &used_index); }
    catch (const NafError &e) {
        if (!used_index || e.code != NAFGPU_E_FORMAT || (ex.pipe && ex.pipe->emitting && ex.pipe->out_done)) throw;
    }
    CUDA_TRY(cudaStreamSynchronize(ex.stream));
    ex.arena->rewind(mk);
    if (ex.pipe) { ex.pipe->wait_all_input(ex.stream); ex.pipe->emitting = false; }
    return decode_impl(ctx, ex, d_naf, h_naf, n, o, false, &used_index);
}

static DecodeOut decode_impl(Ctx &ctx, CudaExec &ex, const u8 *d_naf, const u8 *h_naf, size_t n, const nafgpu_dec_opts &o, bool allow_index, bool *used_index)
{
    using namespace nafc;
    Header h; std::string err;
    if (!read_header(h_naf, n, h, true, err)) fail(NAFGPU_E_FORMAT, err);
    int view = o.out_type;
    if (view == NAFGPU_OUT_DEFAULT) view = h.has_quality ? NAFGPU_OUT_FASTQ : NAFGPU_OUT_FASTA;
    if (view == NAFGPU_OUT_4BIT && h.seq_type >= NAFGPU_PROTEIN)
        fail(NAFGPU_E_INPUT, std::string("input has no 4-bit encoded data, but ") + (h.seq_type == 2 ? "protein" : "text") + " sequences\n");
    if (view == NAFGPU_OUT_FASTQ && !h.has_quality && h.n_sequences > 0) fail(NAFGPU_E_INPUT, "FASTQ output requested, but input has no qualities\n");
    const u64 N = h.n_sequences;
    DecodeOut none{nullptr, 0};
    if (N == 0) return none;                                              // unnaf.c:409
    const bool packed = h.seq_type < NAFGPU_PROTEIN;
    const u64 W = o.have_line_length ? o.line_length : h.line_length;

    // which sections does this view need?
    bool need[6] = {false, false, false, false, false, false};
    switch (view) {
    case NAFGPU_OUT_FASTA:     need[SEC_IDS] = need[SEC_NAMES] = need[SEC_LEN] = need[SEC_DATA] = true; need[SEC_MASK] = !o.no_mask; break;
    case NAFGPU_OUT_FASTQ:     need[SEC_IDS] = need[SEC_NAMES] = need[SEC_LEN] = need[SEC_DATA] = need[SEC_QUAL] = true; break;
    case NAFGPU_OUT_SEQ: case NAFGPU_OUT_CHARCOUNT: need[SEC_DATA] = true; need[SEC_MASK] = !o.no_mask; break;
    case NAFGPU_OUT_SEQUENCES: need[SEC_LEN] = need[SEC_DATA] = true; need[SEC_MASK] = !o.no_mask; break;
    case NAFGPU_OUT_4BIT:      need[SEC_DATA] = true; break;
    case NAFGPU_OUT_IDS:       need[SEC_IDS] = true; break;
    case NAFGPU_OUT_NAMES:     need[SEC_IDS] = need[SEC_NAMES] = true; break;
    case NAFGPU_OUT_LENGTHS:   need[SEC_LEN] = true; break;
    case NAFGPU_OUT_MASK:      need[SEC_MASK] = true; break;
    default: fail(NAFGPU_E_ARG, "unknown output requested\n");
    }
    if (!packed) need[SEC_MASK] = need[SEC_MASK] && false;                // protein/text never carry a mask
    for (int k = 0; k < 6; k++) need[k] = need[k] && h.sec[k].present;
    if ((view == NAFGPU_OUT_FASTA || view == NAFGPU_OUT_FASTQ || view == NAFGPU_OUT_SEQ || view == NAFGPU_OUT_SEQUENCES ||
         view == NAFGPU_OUT_CHARCOUNT || view == NAFGPU_OUT_4BIT) && !h.has_data) return none;
    if (view == NAFGPU_OUT_LENGTHS && !h.has_lengths) return none;
    if (view == NAFGPU_OUT_MASK && !h.has_mask) return none;

    // ---- entropy stage.  The host walks of the block headers (one serial chain per stream) start right away, the long ones
    // on their own threads; the device work comes in up to three parts:
    //   (1) the streams the per-record scans need (ids, names, lengths, mask) -- and, unless it is part (3), the sequence
    //   (2) [the per-record scans below]
    //   (3) the last big stream (quality for FASTQ, else sequence): after (2) when only a range of records is wanted (one
    //       rank of a multi-GPU decode: blocks outside the range are skipped), or when the file is still on its way up from
    //       the host -- then piece by piece behind the upload, each piece's text written and sent down while the next one
    //       arrives (DecodePieces below).  Otherwise (1) and (3) are one batch.
    const bool ranged = o.n_records != 0 && (view == NAFGPU_OUT_FASTA || view == NAFGPU_OUT_FASTQ || view == NAFGPU_OUT_SEQUENCES ||
                                            view == NAFGPU_OUT_IDS || view == NAFGPU_OUT_NAMES);
    nafz::ZDecPlan plan;
    plan.blocks.swap(ctx.zblock_cache);                                   // reuse last call's capacity
    struct GiveBack { nafz::ZDecPlan &p; Ctx &c; ~GiveBack() { p.blocks.swap(c.zblock_cache); } } give_back{plan, ctx};
    static const char *what[6] = { "ids", "names", "lengths", "mask", "sequence", "quality" };
    u64 sbytes[6]; u64 soff[6]; u64 arena_sz = 0;
    nafz::ZStreamDesc sdesc[6];
    for (int k = 0; k < 6; k++) {
        sbytes[k] = 0; soff[k] = 0;
        if (!need[k]) continue;
        u64 expect = h.sec[k].orig;
        if (k == SEC_DATA && packed) expect = (h.sec[k].orig + 1) / 2;
        // no zstd frame regenerates more than 128 KB from the 4 bytes of an RLE block: a header that claims more is
        // damaged, and must fail here rather than as a failed multi-terabyte allocation
        if (expect > (h.sec[k].comp + 4) * 32768 + (128u << 10)) fail(NAFGPU_E_FORMAT, std::string("can't decompress ") + what[k] + "\n");
        if (h.sec[k].comp == 0) fail(NAFGPU_E_FORMAT, std::string("can't decompress: empty zstd stream\n"));
        sbytes[k] = expect; soff[k] = arena_sz;
        nafz::ZStreamDesc sd; sd.src_off = h.sec[k].off; sd.src_len = h.sec[k].comp; sd.out_off = arena_sz; sd.out_size = expect;
        sd.one_frame = (k == SEC_DATA || k == SEC_QUAL) ? 1 : 0; sd.no_magic = 1; sd.need_lo = 0; sd.need_hi = ~0ull;
        sdesc[k] = sd;
        arena_sz += align256(expect + 64);
    }
    // host walks
    struct Walks {
        nafz::ZWalked *w; std::thread th[6]; bool pending[6] = {false, false, false, false, false, false};
        void join(int k) { if (pending[k]) { th[k].join(); pending[k] = false; } }
        ~Walks() { for (int k = 0; k < 6; k++) join(k); }
    } walks{ctx.zwalk};
    const int last_big = need[SEC_QUAL] ? SEC_QUAL : (need[SEC_DATA] ? SEC_DATA : -1);
    static const bool trace = getenv("NAFGPU_TRACE") != nullptr;
    const auto t_begin = std::chrono::steady_clock::now();
    auto mark = [&](const char *what_) {
        if (!trace) return;
        cudaStreamSynchronize(ex.stream);
        fprintf(stderr, "nafgpu trace: %8.3f ms  %s\n", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count(), what_);
    };
    // block index (zstd_walk_indexed): a skippable frame behind the lengths frame of files we wrote.  The lengths section is
    // small -- walk it first (even if this view does not print lengths) and look.
    nafz::ZIndexEntry index[6]; bool indexed[6] = {false, false, false, false, false, false};
    static const bool env_index = !(getenv("NAFGPU_INDEX") && getenv("NAFGPU_INDEX")[0] == '0');
    const bool use_index = env_index && allow_index;
    if (use_index && h.sec[SEC_LEN].present && h.sec[SEC_LEN].comp >= 2 && (need[SEC_DATA] || need[SEC_QUAL])) {
        nafz::ZWalked &w = ctx.zwalk[SEC_LEN];
        w.blocks.clear(); w.regen.clear(); w.skips.clear(); w.simple = false; w.consumed = 0; w.rc = 0; w.err.clear();
        nafz::ZStreamDesc sd = sdesc[SEC_LEN];
        if (!need[SEC_LEN]) { sd.src_off = h.sec[SEC_LEN].off; sd.src_len = h.sec[SEC_LEN].comp; sd.out_off = 0; sd.out_size = 0; sd.one_frame = 0; sd.no_magic = 1; }
        w.rc = nafz::zstd_walk_stream(h_naf, sd, 0, w.blocks, &w.consumed, w.err, &w.regen, &w.simple, &w.skips);
        for (auto &sk : w.skips) {
            const u8 *q = h_naf + sk.first; const u64 sz = sk.second;
            if (sz < 12 || memcmp(q, "NAFGIDX1", 8) != 0) continue;
            const u32 ns = q[8] | (q[9] << 8) | (q[10] << 16) | ((u32)q[11] << 24);
            if (ns > 2 || sz < 12 + (u64)ns * 24) continue;
            u64 arr = 12 + (u64)ns * 24;
            for (u32 j = 0; j < ns; j++) {
                const u8 *e = q + 12 + 24 * j;
                auto rd32 = [&](int o) { return (u32)e[o] | ((u32)e[o + 1] << 8) | ((u32)e[o + 2] << 16) | ((u32)e[o + 3] << 24); };
                nafz::ZIndexEntry ie; ie.section = rd32(0); ie.nblk = rd32(4); ie.regen = rd32(8); ie.reserved = rd32(12);
                ie.total = (u64)rd32(16) | ((u64)rd32(20) << 32); ie.csize = q + arr;
                if (arr + 2ull * ie.nblk > sz) break;
                arr += 2ull * ie.nblk;
                if ((ie.section == SEC_DATA || ie.section == SEC_QUAL) && need[ie.section] && ie.total == sbytes[ie.section]) { index[ie.section] = ie; indexed[ie.section] = true; }
            }
        }
    }
    for (int k = 0; k < 6; k++) {
        if (!need[k]) continue;
        nafz::ZWalked &w = ctx.zwalk[k];
        if (k == SEC_LEN && use_index && (need[SEC_DATA] || need[SEC_QUAL])) continue;          // walked above
        w.blocks.clear(); w.regen.clear(); w.skips.clear(); w.simple = false; w.consumed = 0; w.rc = 0; w.err.clear();
        const nafz::ZStreamDesc sd = sdesc[k];
        if (indexed[k]) {
            const auto t0 = std::chrono::steady_clock::now();
            const bool ok = nafz::zstd_walk_indexed(h_naf, sd, index[k], w.blocks, w.regen, &w.consumed, false);
            if (trace) fprintf(stderr, "nafgpu trace: block index of section %d: %zu blocks, %s, %.3f ms\n", k, w.blocks.size(), ok ? "used" : "REJECTED",
                               std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
            if (ok) { w.simple = true; *used_index = true; continue; }
            w.blocks.clear(); w.regen.clear();
        }
        const bool want_regen = true;        // cheap (the literals header sits next to the block header) and it decides the fast path
        auto body = [&w, sd, h_naf, want_regen, k]() {
            const auto t0 = std::chrono::steady_clock::now();
            w.rc = nafz::zstd_walk_stream(h_naf, sd, 0, w.blocks, &w.consumed, w.err, want_regen ? &w.regen : nullptr, want_regen ? &w.simple : nullptr);
            if (trace) fprintf(stderr, "nafgpu trace: host walk of section %d: %zu blocks, %.3f ms\n", k, w.blocks.size(),
                               std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
        };
        if (sd.src_len > (8u << 20)) { walks.th[k] = std::thread(body); walks.pending[k] = true; }
        else body();
    }
    u8 *d_streams = ex.alloc<u8>(arena_sz + 256);
    // padding bytes between streams are read by the 16-byte loaders: keep them defined (the streams themselves are
    // written by the decoder; bytes of blocks a record-range decode skips are never used)
    for (int k = 0; k < 6; k++) if (need[k]) ex.zero(d_streams + soff[k] + sbytes[k], align256(sbytes[k] + 64) - sbytes[k]);
    ex.zero(d_streams + arena_sz, 256);
    u64 data_out_size = 0;
    // decode the streams of `mask` (bit k = section k) as one batch
    auto run_batch = [&](u32 mask) {
        plan.streams.clear(); plan.blocks.clear();
        int in_batch[6], nb = 0; u64 in_hi = 0; u64 slice_regen[6];
        bool all_simple = true;
        for (int k = 0; k < 6; k++) {
            if (!need[k] || !((mask >> k) & 1)) continue;
            walks.join(k);
            nafz::ZWalked &w = ctx.zwalk[k];
            if (w.rc) fail(NAFGPU_E_FORMAT, std::string("can't decompress: ") + w.err + "\n");
            if (!(w.simple && w.regen.size() == w.blocks.size())) all_simple = false;
        }
        plan.simple = all_simple;
        for (int k = 0; k < 6; k++) {
            if (!need[k] || !((mask >> k) & 1)) continue;
            nafz::ZWalked &w = ctx.zwalk[k];
            if (!all_simple && indexed[k] && !w.blocks.empty() && w.blocks[0].type == 0xFF) {
                // the batch takes the general path (another stream of it has dependent blocks): that one wants the block types
                // from the host, so read the headers after all -- or walk the stream if they disagree with the index
                if (!nafz::zstd_walk_indexed(h_naf, sdesc[k], index[k], w.blocks, w.regen, &w.consumed, true)) {
                    w.blocks.clear(); w.regen.clear();
                    w.rc = nafz::zstd_walk_stream(h_naf, sdesc[k], 0, w.blocks, &w.consumed, w.err, &w.regen, &w.simple);
                    if (w.rc) fail(NAFGPU_E_FORMAT, std::string("can't decompress: ") + w.err + "\n");
                }
            }
            const u32 base = (u32)plan.blocks.size(), si = (u32)plan.streams.size();
            nafz::ZStreamDesc sd = sdesc[k];
            slice_regen[nb] = ~0ull;
            const bool simple_stream = w.simple && w.regen.size() == w.blocks.size();
            const bool sliced = simple_stream && (sd.need_lo > 0 || sd.need_hi < sbytes[k]);
            if (simple_stream && (sliced || all_simple)) {
                // a stream of self-contained blocks: only the blocks that hold bytes [need_lo, need_hi) reach the device, and
                // (when the whole batch is like that) each block is told where its output goes
                u64 off = 0, off0 = 0, sum = 0; size_t i0 = w.blocks.size(), i1 = 0;
                for (size_t i = 0; i < w.blocks.size(); i++) {
                    const u64 lo = off, hi = off + w.regen[i];
                    if ((hi > sd.need_lo && lo < sd.need_hi) || (!sliced && w.regen[i] == 0)) { if (i0 == w.blocks.size()) { i0 = i; off0 = lo; } i1 = i + 1; sum += w.regen[i]; }
                    off = hi;
                }
                const bool exact = !(k == SEC_DATA || k == SEC_QUAL);
                if (exact ? off != sbytes[k] : off < sbytes[k]) fail(NAFGPU_E_FORMAT, std::string("can't decompress ") + what[k] + "\n");
                if (k == SEC_DATA) data_out_size = off;
                if (i1 <= i0) continue;                                  // nothing of this stream is wanted
                u64 at = off0;
                for (size_t i = i0; i < i1; i++) {
                    nafz::ZBlockHead hb = w.blocks[i];
                    hb.frame_first_blk = base; hb.stream = (u8)si;
                    hb.first_in_frame = hb.first_in_stream = i == i0;
                    if (all_simple) { hb.out_base = sd.out_off + at; hb.rsize = w.regen[i]; }
                    else hb.out_base = sd.out_off + off0;
                    at += w.regen[i];
                    plan.blocks.push_back(hb);
                    const u64 end = hb.src + hb.csize; if (end > in_hi) in_hi = end;
                }
                sd.out_off += off0; sd.out_size = sum; sd.need_lo = 0; sd.need_hi = ~0ull;
                slice_regen[nb] = sum;
            } else {
                for (auto &b : w.blocks) { nafz::ZBlockHead hb = b; hb.frame_first_blk += base; hb.stream = (u8)si; plan.blocks.push_back(hb); }
                if (h.sec[k].off + h.sec[k].comp > in_hi) in_hi = h.sec[k].off + h.sec[k].comp;
            }
            plan.streams.push_back(sd); in_batch[nb++] = k;
        }
        if (!nb) return;
        plan.results.assign(plan.streams.size(), nafz::ZStreamResult{0, 0, 0});
        if (ex.pipe) ex.pipe->wait_input(ex.stream, in_hi);
        std::string zerr;
        int rc = nafz::zstd_decode_blocks(ex, d_naf, d_streams, plan, ctx.d_predef, zerr);
        if (rc) fail(rc == -2 ? NAFGPU_E_UNSUPPORTED : NAFGPU_E_FORMAT, std::string("can't decompress: ") + zerr + "\n");
        for (int j = 0; j < nb; j++) {
            const int k = in_batch[j];
            u64 got = plan.results[j].out_size;
            if (slice_regen[j] != ~0ull) {                                // a slice of a simple stream: exactly what the walk promised, no sequences
                if (got != slice_regen[j] || plan.results[j].nseq != 0) fail(NAFGPU_E_FORMAT, std::string("can't decompress ") + what[k] + "\n");
                continue;
            }
            bool exact = !(k == SEC_DATA || k == SEC_QUAL);
            if (exact ? got != sbytes[k] : got < sbytes[k]) fail(NAFGPU_E_FORMAT, std::string("can't decompress ") + what[k] + "\n");
            if (k == SEC_DATA) data_out_size = got;
        }
    };
    const bool rec_text_view = view == NAFGPU_OUT_FASTA || view == NAFGPU_OUT_FASTQ || view == NAFGPU_OUT_SEQUENCES;
    // piecewise decode of the last big stream behind the upload: decided once its walk says the stream can be cut
    bool piecewise = false;
    const u32 all_mask = 0x3F, big_mask = (1u << SEC_DATA) | (1u << SEC_QUAL);
    u32 later_mask = 0;                                                   // streams decoded after the per-record scans
    if (ranged) later_mask = big_mask;
    // A record range of a file with few records (a genome: one length per chromosome): when the lengths frame is a handful of
    // raw / RLE blocks the host reads it where it lies, knows which bases the range covers before any device work, and the
    // sequence / quality blocks of the range join the first batch -- one pass through the decoder's latency instead of two.
    if (ranged && need[SEC_LEN] && (need[SEC_DATA] || need[SEC_QUAL]) && sbytes[SEC_LEN] <= (1u << 20) && sbytes[SEC_LEN] % 4 == 0) {
        const nafz::ZWalked &w = ctx.zwalk[SEC_LEN];
        bool plain = w.rc == 0 && !w.blocks.empty();
        u64 regen = 0;
        for (auto &b : w.blocks) { plain = plain && b.type < 2; regen += b.rsize; }
        if (plain && regen == sbytes[SEC_LEN]) {
            std::vector<u8> hl; hl.reserve(regen);
            for (auto &b : w.blocks) { if (b.type == 0) hl.insert(hl.end(), h_naf + b.src, h_naf + b.src + b.rsize); else hl.insert(hl.end(), b.rsize, h_naf[b.src]); }
            // lengths -> bases before record r (continuation units merged: output.c:390-393)
            const u64 nunits = regen / 4, r0 = o.first_record, r1 = o.first_record + o.n_records < o.first_record ? ~0ull : o.first_record + o.n_records;
            u64 rec = 0, bases = 0, b0 = 0, b1 = 0; bool have0 = false, have1 = false;
            for (u64 k = 0; k < nunits; k++) {
                if (rec == r0 && !have0) { b0 = bases; have0 = true; }
                if (rec == r1 && !have1) { b1 = bases; have1 = true; }
                const u32 v = (u32)hl[4 * k] | ((u32)hl[4 * k + 1] << 8) | ((u32)hl[4 * k + 2] << 16) | ((u32)hl[4 * k + 3] << 24);
                bases += v;
                if (v != 0xFFFFFFFFu) rec++;
            }
            if (!have0) b0 = bases;
            if (!have1) b1 = bases;
            if (bases <= h.sec[SEC_DATA].orig && b0 <= b1) {
                if (need[SEC_DATA]) { sdesc[SEC_DATA].need_lo = packed ? b0 / 2 : b0; sdesc[SEC_DATA].need_hi = packed ? (b1 + 1) / 2 : b1; }
                if (need[SEC_QUAL]) { sdesc[SEC_QUAL].need_lo = b0; sdesc[SEC_QUAL].need_hi = b1; }
                later_mask = 0;
            }
        }
    }
    else if (last_big >= 0 && rec_text_view && (walks.pending[last_big] || (ex.pipe && ex.pipe->uploading))) later_mask = 1u << last_big;
    // FASTQ from a host buffer whose two big streams can both be cut at block boundaries (files we wrote: the block index says
    // so before a byte has gone up): the file is uploaded small streams first, then sequence piece 0, quality piece 0, sequence
    // piece 1 ... and each pair of pieces is decoded, its records written and their text sent down while the rest is still on
    // the host.  Front to back, nothing can be written before the whole sequence stream AND the first quality piece are up.
    std::vector<DuoPiece> duo;                                            // (duo_plan.hpp)
    static const bool env_duo = !(getenv("NAFGPU_DUO") && getenv("NAFGPU_DUO")[0] == '0');
    if (env_duo && ex.pipe && ex.pipe->deferred && !ranged && view == NAFGPU_OUT_FASTQ && need[SEC_DATA] && need[SEC_QUAL] &&
        !walks.pending[SEC_DATA] && !walks.pending[SEC_QUAL] && h.sec[SEC_DATA].off < h.sec[SEC_QUAL].off) {
        const nafz::ZWalked &ws = ctx.zwalk[SEC_DATA], &wq = ctx.zwalk[SEC_QUAL];
        const char *env_piece0 = getenv("NAFGPU_PIPE_PIECE");
        const u64 PIECE = env_piece0 && *env_piece0 ? strtoull(env_piece0, nullptr, 10) : (48ull << 20);
        const bool cuttable = ws.rc == 0 && wq.rc == 0 && ws.simple && wq.simple && ws.regen.size() == ws.blocks.size() && wq.regen.size() == wq.blocks.size() &&
                              h.sec[SEC_QUAL].comp >= 2 * PIECE;
        if (cuttable) {
            std::vector<DuoBlock> sb(ws.blocks.size()), qb(wq.blocks.size());
            for (size_t i = 0; i < sb.size(); i++) sb[i] = DuoBlock{ws.blocks[i].src, ws.blocks[i].csize, (u32)ws.regen[i]};
            for (size_t i = 0; i < qb.size(); i++) qb[i] = DuoBlock{wq.blocks[i].src, wq.blocks[i].csize, (u32)wq.regen[i]};
            std::vector<std::pair<u64, u64>> order;
            if (duo_plan(sb, qb, packed, PIECE, n, duo, order)) { ex.pipe->start_upload(&order); later_mask = big_mask; }
        }
    }
    if (ex.pipe) ex.pipe->start_upload(nullptr);                          // (unless it has just been started in that order)
    mark("walks started / index read");
    run_batch(all_mask & ~later_mask);
    mark("first batch decoded");
    const u8 *d_ids = d_streams + soff[SEC_IDS], *d_comm = d_streams + soff[SEC_NAMES], *d_mask = d_streams + soff[SEC_MASK];
    const u8 *d_seq = d_streams + soff[SEC_DATA], *d_qual = d_streams + soff[SEC_QUAL];
    const u32 *d_len = (const u32 *)(d_streams + soff[SEC_LEN]);
    const u64 nL = sbytes[SEC_LEN] / 4, nM = sbytes[SEC_MASK], total_bases = h.sec[SEC_DATA].orig;

    // raw views
    if (view == NAFGPU_OUT_4BIT) return DecodeOut{d_seq, data_out_size};
    if (view == NAFGPU_OUT_LENGTHS) return DecodeOut{(const u8 *)d_len, sbytes[SEC_LEN]};
    if (view == NAFGPU_OUT_MASK) return DecodeOut{d_mask, nM};

    TextArgs A; memset(&A, 0, sizeof A);
    A.ids = d_ids; A.comm = d_comm; A.seq = d_seq; A.qual = d_qual;
    A.has_ids = need[SEC_IDS]; A.has_names = need[SEC_NAMES]; A.sep = h.sep; A.packed = packed; A.W = W;
    A.total_bases = total_bases;
    {
        char lut[17] = "-TGKCYSBAWRDMHVN";
        if (h.seq_type == NAFGPU_RNA) lut[1] = 'U';
        memcpy(A.lut, lut, 16);
    }
    A.upper = (!packed && o.no_mask && view != NAFGPU_OUT_FASTQ) ? 1 : 0;   // output.c:500,663; FASTQ path never uppercases

    // ---- ids / comments: terminator positions; the reference insists on a final '\0' (input.c:157,185)
    u32 *d_id_end = nullptr, *d_cm_end = nullptr;
    auto find_ends = [&](const u8 *s, u64 bytes, const char *name) -> u32 * {
        if (bytes == 0) fail(NAFGPU_E_FORMAT, std::string("corrupted ") + name + " - not 0-terminated\n");
        // terminator offsets are kept as u32: a string stream of 4 GiB or more would wrap them silently
        if (bytes >= (1ull << 32)) fail(NAFGPU_E_UNSUPPORTED, std::string(name) + " stream of 4 GiB or more is not supported by this build\n");
        u64 ntiles = (bytes + ZT - 1) / ZT;
        u64 *counts = ex.alloc<u64>(ntiles + 1), *prefix = ex.alloc<u64>(ntiles + 2);
        KLAUNCH(ex, "k_zero_count", k_zero_count<<<(unsigned)ntiles, 256, 0, ex.stream>>>(s, bytes, counts));
        const u64 *c = counts;
        exclusive_scan(ex, [c] __device__ (size_t i) { return c[i]; }, ntiles, prefix);
        // the header's N is not trusted: count the terminators before anything is sized by it
        u64 nzero; u8 last;
        ex.download(&nzero, prefix + ntiles, 8);
        ex.download(&last, s + bytes - 1, 1);
        if (last != 0) fail(NAFGPU_E_FORMAT, std::string("corrupted ") + name + " - not 0-terminated\n");
        if (nzero < N || N > bytes) fail(NAFGPU_E_FORMAT, std::string("corrupted ") + name + " - can't read all records\n");
        u32 *end = ex.alloc<u32>(N + 1);
        KLAUNCH(ex, "k_zero_scatter", k_zero_scatter<<<(unsigned)ntiles, 256, 0, ex.stream>>>(s, bytes, prefix, end, N));
        return end;
    };
    if (A.has_ids) d_id_end = find_ends(d_ids, sbytes[SEC_IDS], "ids");
    if (A.has_names) d_cm_end = find_ends(d_comm, sbytes[SEC_NAMES], "names");
    A.id_end = d_id_end; A.cm_end = d_cm_end;

    // ---- lengths: merge 0xFFFFFFFF continuation units (output.c:390-393), prefix-sum into base offsets
    u64 *d_L = nullptr, *d_seq_start = nullptr;
    u64 NR = N;      // records in the text
    u64 len_sum = 0; // bases the length units add up to
    const bool rec_views = view == NAFGPU_OUT_FASTA || view == NAFGPU_OUT_FASTQ || view == NAFGPU_OUT_SEQUENCES;
    if (rec_views) {
        if (nL == 0) fail(NAFGPU_E_FORMAT, "can't decompress lengths\n");
        u64 *rank = ex.alloc<u64>(nL + 1);
        const u32 *len = d_len;
        exclusive_scan(ex, [len] __device__ (size_t k) { return (u64)(len[k] != 0xFFFFFFFFu); }, nL, rank);
        u64 nrec_len; ex.download(&nrec_len, rank + nL, 8);
        if (view == NAFGPU_OUT_SEQUENCES) NR = nrec_len;                   // print_sequences walks length units, not N
        else if (nrec_len < N) NR = nrec_len;                              // output.c:413 stops at n_lengths
        d_L = ex.alloc<u64>(NR + 1); d_seq_start = ex.alloc<u64>(NR + 2);
        ex.zero(d_L, (NR + 1) * 8);
        u64 *L = d_L; const u64 nr = NR;
        ex.for_each(nL, [=] __device__ (size_t k) { u64 r = rank[k]; if (r < nr) atomicAdd((unsigned long long *)(L + r), (unsigned long long)len[k]); });
        exclusive_scan(ex, [L] __device__ (size_t i) { return L[i]; }, NR, d_seq_start);
        // clamp to the bases actually present (print_dna_buffer_as_fasta never prints past total_seq_length)
        u64 sum; ex.download(&sum, d_seq_start + NR, 8);
        if (sum > total_bases) fail(NAFGPU_E_FORMAT, "corrupted lengths - sum exceeds the sequence size\n");
        // FASTQ: record i's qualities are quality[seq_start[i] .. +L[i]) -- the stream must hold all of them (the reference never
        // returns on such a file: refill_quality_buffer_from_file, input.c:409, has nothing left to read)
        if (view == NAFGPU_OUT_FASTQ && sum > sbytes[SEC_QUAL]) fail(NAFGPU_E_FORMAT, "corrupted quality - shorter than the sum of the sequence lengths\n");
        len_sum = sum;
    }
    A.L = d_L; A.seq_start = d_seq_start;

    // ---- record layout per view
    switch (view) {
    case NAFGPU_OUT_FASTA:     A.prefix = '>'; A.with_name = 1; A.name_nl = 1; A.seq_present = 1; A.seq_nl = 1; break;
    case NAFGPU_OUT_FASTQ:     A.prefix = '@'; A.with_name = 1; A.name_nl = 1; A.seq_present = 1; A.seq_nl = 2; A.with_qual = 1; A.W = 0; A.maskbits = nullptr; break;
    case NAFGPU_OUT_SEQUENCES: A.seq_present = 1; A.seq_nl = 2; A.W = 0; break;
    case NAFGPU_OUT_IDS:       A.with_name = 1; A.name_nl = 1; A.has_names = 0; break;
    case NAFGPU_OUT_NAMES:     A.with_name = 1; A.name_nl = 1; break;
    default: break;            // SEQ / CHARCOUNT: handled below as one pseudo-record
    }
    if (view == NAFGPU_OUT_IDS && !A.has_ids) return none;
    if (view == NAFGPU_OUT_NAMES && !A.has_ids && !A.has_names) return none;
    if (view == NAFGPU_OUT_SEQUENCES && total_bases == 0) return none;    // output-sequences.c: nothing is flushed without bases

    u64 *d_out_start;
    if (view == NAFGPU_OUT_SEQ || view == NAFGPU_OUT_CHARCOUNT) {
        // one pseudo-record holding every base, no newline
        A.seq_present = 1; A.seq_nl = 0; A.W = 0; NR = 1;
        u64 hl[2] = { total_bases, 0 }, hs[3] = { 0, total_bases, 0 };
        d_L = ex.alloc<u64>(2); d_seq_start = ex.alloc<u64>(3); d_out_start = ex.alloc<u64>(3);
        ex.upload(d_L, hl, 16); ex.upload(d_seq_start, hs, 24); ex.upload(d_out_start, hs, 24);
        A.L = d_L; A.seq_start = d_seq_start;
    } else {
        d_out_start = ex.alloc<u64>(NR + 2);
        const TextArgs B = A;
        exclusive_scan(ex, [B] __device__ (size_t i) {
            u32 idl = 0, cml = 0;
            if (B.with_name) {
                if (B.has_ids) { u32 s = i ? B.id_end[i - 1] + 1 : 0; idl = B.id_end[i] - s; }
                if (B.has_names) { u32 s = i ? B.cm_end[i - 1] + 1 : 0; cml = B.cm_end[i] - s; }
            }
            u32 nl = !B.with_name ? 0 : ((B.has_ids && B.has_names) ? idl + (cml ? 1 + cml : 0) : (B.has_ids ? idl : cml));
            return rec_text_size(B.prefix, nl, B.name_nl, B.seq_present, B.seq_nl, B.with_qual, B.W, B.seq_present ? B.L[i] : 0);
        }, NR, d_out_start);
    }
    A.N = NR; A.out_start = d_out_start; A.rec0 = 0;
    u64 total; ex.download(&total, d_out_start + NR, 8);
    mark("per-record scans done");
    const u64 NR_all = NR;
    bool range_has_tail = true;                                           // the range ends with the file's last record
    u64 range_b0 = 0, range_b1 = total_bases;                             // bases of the records of this call
    if (ranged) {
        // records [r0, r1) only: their text is the byte range [o0, o1) of the whole output
        const u64 r0 = o.first_record < NR ? o.first_record : NR, r1 = (o.n_records < NR - r0) ? r0 + o.n_records : NR;
        u64 o01[2] = { total, total };
        const bool want_b = need[SEC_DATA] || need[SEC_QUAL];
        {   // the four range bounds in one read-back (each download is a synchronisation)
            u64 *d4 = ex.alloc<u64>(4);
            const u64 *os = d_out_start, *ss = d_seq_start;
            ex.for_each(1, [=] __device__ (size_t) { d4[0] = os[r0]; d4[1] = os[r1]; d4[2] = want_b ? ss[r0] : 0; d4[3] = want_b ? ss[r1] : 0; }, "range_bounds");
            u64 h4[4]; ex.download(h4, d4, 32);
            o01[0] = h4[0]; o01[1] = h4[1];
            if (want_b) { range_b0 = h4[2]; range_b1 = h4[3]; }
        }
        if (want_b) {
            const u64 b01[2] = { range_b0, range_b1 };
            if (later_mask) {
                if (need[SEC_DATA]) { sdesc[SEC_DATA].need_lo = packed ? b01[0] / 2 : b01[0]; sdesc[SEC_DATA].need_hi = packed ? (b01[1] + 1) / 2 : b01[1]; }
                if (need[SEC_QUAL]) { sdesc[SEC_QUAL].need_lo = b01[0]; sdesc[SEC_QUAL].need_hi = b01[1]; }
                run_batch(big_mask);
            }
        }
        u64 *local = ex.alloc<u64>(r1 - r0 + 2);
        const u64 *src = d_out_start + r0; const u64 base = o01[0];
        ex.for_each(r1 - r0 + 1, [=] __device__ (size_t i) { local[i] = src[i] - base; }, "range_out_start");
        A.N = NR = r1 - r0; A.out_start = d_out_start = local; A.rec0 = r0;
        total = o01[1] - o01[0];
        range_has_tail = r1 == NR_all && r1 > r0;
    }
    A.total = total;
    if (total == 0 || NR == 0) return none;
    mark("range streams decoded");
    // ---- mask: run-length units -> one bit per base (output.c:295 semantics, two-scan formulation); only over the bases this
    // call prints
    if (need[SEC_MASK] && packed && nM > 0 && view != NAFGPU_OUT_FASTQ) {
        u64 words = (total_bases + 31) / 32 + 2;
        u32 *bits = ex.alloc<u32>(words);
        const u64 b0 = range_b0, b1 = range_has_tail ? total_bases : range_b1;
        const u64 w0 = b0 / 32, w1 = (b1 + 31) / 32 + 2 < words ? (b1 + 31) / 32 + 2 : words;
        if (w1 > w0) ex.zero(bits + w0, (w1 - w0) * 4);
        u64 *ustart = ex.alloc<u64>(nM + 1), *utog = ex.alloc<u64>(nM + 1);
        const u8 *m = d_mask;
        exclusive_scan(ex, [m] __device__ (size_t k) { return (u64)m[k]; }, nM, ustart);
        exclusive_scan(ex, [m] __device__ (size_t k) { return (u64)(m[k] != 255); }, nM, utog);
        const u64 tb = total_bases < b1 ? total_bases : b1;
        ex.for_each(nM, [=] __device__ (size_t k) {
            if (!(utog[k] & 1)) return;
            u64 lo = ustart[k], hi = lo + m[k];
            if (lo < b0) lo = b0;
            if (hi > tb) hi = tb;
            if (lo < hi) nafz::set_bits(bits, lo, hi);
        });
        A.maskbits = bits;
    }
    // ---- FASTA: sequence data beyond what the length units add up to.  ennaf's id-byte bug (an unexpected byte in an id puts
    // its '?' into the SEQUENCE buffer, process.c:366,485) writes such files, and print_dna_buffer_as_fasta
    // (output.c:420-427) prints the surplus after the last record: first into whatever is left of that record's last line,
    // then in lines of W, with no newline at the end.  Rare and tiny, so it is laid out on the host: the surplus bases are
    // produced by the text kernel as one bare pseudo-record, fetched, wrapped, and put behind the text.
    u64 surplus = 0, surplus_text = 0, line_rem = 0;
    if (view == NAFGPU_OUT_FASTA && range_has_tail && len_sum < total_bases) {
        surplus = total_bases - len_sum;
        if (surplus > (64u << 20)) fail(NAFGPU_E_UNSUPPORTED, "more than 64 MB of sequence beyond the recorded lengths is not supported by this build\n");
        // length of the last non-empty record: its last line decides the budget (empty records do not touch it)
        u64 last_len = 0;
        for (u64 win = 4096; last_len == 0; win *= 16) {
            const u64 cnt = win < NR_all ? win : NR_all;
            std::vector<u64> tail(cnt);
            ex.download(tail.data(), d_L + (NR_all - cnt), cnt * 8);
            for (u64 i = cnt; i > 0 && last_len == 0; i--) last_len = tail[i - 1];
            if (cnt == NR_all) break;
        }
        if (last_len == 0) surplus = 0;                                  // no record has bases: print_fasta returns before any sequence (output.c:629)
        else if (W == 0) surplus_text = surplus;
        else {
            line_rem = W - ((last_len - 1) % W + 1);
            u64 sz = surplus, lr = line_rem;
            while (sz > lr) { surplus_text += lr + 1; sz -= lr; lr = W; }
            surplus_text += sz;
        }
    }
    mark("mask bits built");
    u8 *d_text = ex.alloc<u8>(total + surplus_text + 64);
    A.out = d_text;
    if (NR >= 0xFFFFFFFFull) fail(NAFGPU_E_UNSUPPORTED, "more than 2^32 - 1 records in one file are not supported by this build\n");
    const u64 ntiles = (total + WT_TILE - 1) / WT_TILE;
    u32 *tile_first = ex.alloc<u32>(ntiles + 2);
    KLAUNCH(ex, "k_tile_first", k_tile_first<<<(unsigned)((ntiles + 1 + 255) / 256), 256, 0, ex.stream>>>(d_out_start, NR, ntiles, tile_first));
    // host-buffer call: finished tiles of the text go down while the next ones are being written
    const bool sink = ex.pipe && rec_text_view && view != NAFGPU_OUT_CHARCOUNT;
    if (sink) ex.pipe->begin_output(ex.pipe->sink ? nullptr : ctx.pinned_out.ensure(total + surplus_text + 1));   // (a callback sink brings its own two buffers)
    auto write_tiles = [&](u64 t0, u64 t1) {
        const u64 group = sink ? (64ull << 20) / WT_TILE : ntiles;      // tiles per launch when each launch is followed by its copy
        for (u64 a = t0; a < t1; a += group) {
            const u64 b = a + group < t1 ? a + group : t1;
            TextArgs B = A; B.tile_base = a;
            static const bool old_writer = getenv("NAFGPU_OLD_WRITER") != nullptr;      // A/B: the per-chunk formulation
            if (old_writer) { KLAUNCH(ex, "k_write_text", k_write_text<<<(unsigned)(b - a), WT_THREADS, 0, ex.stream>>>(B, tile_first)); }
            else { KLAUNCH(ex, "k_compose_text", k_compose_text<<<(unsigned)(b - a), WT_THREADS, 0, ex.stream>>>(B, tile_first)); }
            if (sink) { const u64 lo = a * WT_TILE, hi = b * WT_TILE < total ? b * WT_TILE : total; ex.pipe->emit(ex.stream, d_text + lo, lo, hi - lo); }
        }
    };
    if (!duo.empty() && surplus != 0) { run_batch(big_mask); write_tiles(0, ntiles); }       // (a file with bases beyond its lengths: whole, as before)
    else if (!duo.empty()) {
        const nafz::ZWalked *wk[2] = { &ctx.zwalk[SEC_DATA], &ctx.zwalk[SEC_QUAL] };
        const int sec[2] = { SEC_DATA, SEC_QUAL };
        u64 *d_upto = ex.alloc<u64>(2);
        u64 out_off[2] = {0, 0}, t_prev = 0;
        for (size_t p = 0; p < duo.size(); p++) {
            const size_t i0[2] = { duo[p].s0, duo[p].q0 }, i1[2] = { duo[p].s1, duo[p].q1 };
            plan.blocks.clear(); plan.streams.clear();
            u64 regen[2] = {0, 0}; int slot[2] = {-1, -1};
            for (int s = 0; s < 2; s++) {
                if (i1[s] <= i0[s]) continue;
                const int k = sec[s]; const nafz::ZWalked &w = *wk[s];
                const u32 base = (u32)plan.blocks.size();
                u64 pre = 0;
                for (size_t i = i0[s]; i < i1[s]; i++) {
                    nafz::ZBlockHead hb = w.blocks[i];
                    hb.frame_first_blk = base; hb.stream = (u8)plan.streams.size(); hb.out_base = soff[k] + out_off[s] + pre; hb.rsize = w.regen[i];
                    hb.first_in_frame = hb.first_in_stream = i == i0[s];
                    pre += w.regen[i];
                    plan.blocks.push_back(hb);
                }
                nafz::ZStreamDesc sd = sdesc[k]; sd.out_off = soff[k] + out_off[s]; sd.out_size = pre;
                slot[s] = (int)plan.streams.size();
                plan.streams.push_back(sd);
                regen[s] = pre;
                ex.pipe->wait_range(ex.stream, w.blocks[i0[s]].src - 3, w.blocks[i1[s] - 1].src + w.blocks[i1[s] - 1].csize);
            }
            plan.simple = true;
            plan.results.assign(plan.streams.size(), nafz::ZStreamResult{0, 0, 0});
            std::string zerr;
            int rc = nafz::zstd_decode_blocks(ex, d_naf, d_streams, plan, ctx.d_predef, zerr);
            if (rc) fail(rc == -2 ? NAFGPU_E_UNSUPPORTED : NAFGPU_E_FORMAT, std::string("can't decompress: ") + zerr + "\n");
            for (int s = 0; s < 2; s++) {
                if (slot[s] < 0) continue;
                if (plan.results[slot[s]].out_size != regen[s] || plan.results[slot[s]].nseq != 0) fail(NAFGPU_E_FORMAT, std::string("can't decompress ") + what[sec[s]] + "\n");
                out_off[s] += regen[s];
            }
            u64 t_hi = ntiles;
            if (p + 1 < duo.size()) {
                const u64 bases_up = packed ? out_off[0] * 2 : out_off[0];
                const u64 done = bases_up < out_off[1] ? bases_up : out_off[1];      // bases AND qualities available so far
                const u64 *ss = d_seq_start, *os = d_out_start; const u64 nr = NR;
                ex.for_each(1, [=] __device__ (size_t) {
                    u64 lo = 0, hi = nr + 1;                         // largest r in [0, nr] with seq_start[r] <= done: records < r are complete
                    while (hi - lo > 1) { const u64 mid = (lo + hi) >> 1; if (ss[mid] <= done) lo = mid; else hi = mid; }
                    d_upto[0] = lo; d_upto[1] = os[lo];
                }, "piece_upto");
                u64 upto[2]; ex.download(upto, d_upto, 16);
                t_hi = upto[1] / WT_TILE;
                if (t_hi > ntiles) t_hi = ntiles;
            } else if (out_off[0] < sbytes[SEC_DATA] || out_off[1] < sbytes[SEC_QUAL]) fail(NAFGPU_E_FORMAT, "can't decompress sequence\n");
            if (t_hi > t_prev) { write_tiles(t_prev, t_hi); t_prev = t_hi; }
        }
        data_out_size = out_off[0];
        piecewise = true;
    }
    else if (!ranged && later_mask) {
        const int k = last_big;
        const char *env_piece0 = getenv("NAFGPU_PIPE_PIECE");
        const u64 PIECE_MIN = env_piece0 && *env_piece0 ? strtoull(env_piece0, nullptr, 10) : (48ull << 20);
        walks.join(k);
        nafz::ZWalked &w = ctx.zwalk[k];
        if (w.rc) fail(NAFGPU_E_FORMAT, std::string("can't decompress: ") + w.err + "\n");
        piecewise = w.simple && ex.pipe && ex.pipe->uploading && surplus == 0 && h.sec[k].comp >= 2 * PIECE_MIN && w.regen.size() == w.blocks.size();
        if (!piecewise) { run_batch(1u << k); write_tiles(0, ntiles); }
        else {
            // pieces of ~PIECE compressed bytes, cut at block boundaries; piece p is decoded as soon as its bytes have arrived,
            // the records it completes are written and their text sent down while piece p + 1 is still on its way up
            const u64 PIECE = PIECE_MIN, nblk = w.blocks.size();
            u64 *d_upto = ex.alloc<u64>(2);
            u64 i0 = 0, out_off = 0, t_prev = 0;
            while (i0 < nblk) {
                u64 i1 = i0, cbytes = 0, regen = 0;
                while (i1 < nblk && cbytes < PIECE) { cbytes += (u64)w.blocks[i1].csize + 3; regen += w.regen[i1]; i1++; }
                if (nblk - i1 < 64) while (i1 < nblk) { regen += w.regen[i1]; i1++; }      // no tiny last piece
                plan.blocks.clear(); plan.streams.clear();
                u64 pre = 0;
                for (u64 i = i0; i < i1; i++) {
                    nafz::ZBlockHead hb = w.blocks[i];
                    hb.frame_first_blk = 0; hb.stream = 0; hb.out_base = soff[k] + out_off + pre; hb.rsize = w.regen[i];
                    hb.first_in_frame = hb.first_in_stream = i == i0;
                    pre += w.regen[i];
                    plan.blocks.push_back(hb);
                }
                plan.simple = true;
                nafz::ZStreamDesc sd = sdesc[k]; sd.out_off = soff[k] + out_off; sd.out_size = regen;
                plan.streams.push_back(sd);
                plan.results.assign(1, nafz::ZStreamResult{0, 0, 0});
                ex.pipe->wait_input(ex.stream, w.blocks[i1 - 1].src + w.blocks[i1 - 1].csize);
                std::string zerr;
                int rc = nafz::zstd_decode_blocks(ex, d_naf, d_streams, plan, ctx.d_predef, zerr);
                if (rc) fail(rc == -2 ? NAFGPU_E_UNSUPPORTED : NAFGPU_E_FORMAT, std::string("can't decompress: ") + zerr + "\n");
                if (plan.results[0].out_size != regen || plan.results[0].nseq != 0) fail(NAFGPU_E_FORMAT, std::string("can't decompress ") + what[k] + "\n");
                out_off += regen;
                if (out_off > sbytes[k] && !(k == SEC_DATA || k == SEC_QUAL)) fail(NAFGPU_E_FORMAT, std::string("can't decompress ") + what[k] + "\n");
                u64 t_hi = ntiles;
                if (i1 < nblk) {
                    const u64 done = (k == SEC_DATA && packed) ? out_off * 2 : out_off;       // bases (qualities) available so far
                    const u64 *ss = d_seq_start, *os = d_out_start; const u64 nr = NR;
                    ex.for_each(1, [=] __device__ (size_t) {
                        u64 lo = 0, hi = nr + 1;                     // largest r in [0, nr] with seq_start[r] <= done: records < r are complete
                        while (hi - lo > 1) { const u64 mid = (lo + hi) >> 1; if (ss[mid] <= done) lo = mid; else hi = mid; }
                        d_upto[0] = lo; d_upto[1] = os[lo];
                    }, "piece_upto");
                    u64 upto[2]; ex.download(upto, d_upto, 16);
                    t_hi = upto[1] / WT_TILE;
                    if (t_hi > ntiles) t_hi = ntiles;
                } else if (out_off < sbytes[k]) fail(NAFGPU_E_FORMAT, std::string("can't decompress ") + what[k] + "\n");
                if (t_hi > t_prev) { write_tiles(t_prev, t_hi); t_prev = t_hi; }
                i0 = i1;
            }
            if (k == SEC_DATA) data_out_size = out_off;
        }
    } else write_tiles(0, ntiles);
    mark("text written");
    if (surplus) {
        TextArgs S = A;
        S.prefix = 0; S.with_name = 0; S.name_nl = 0; S.seq_present = 1; S.seq_nl = 0; S.with_qual = 0; S.W = 0; S.rec0 = 0;
        u64 hl[2] = { surplus, 0 }, hs[3] = { len_sum, total_bases, 0 }, ho[3] = { 0, surplus, 0 };
        u64 *sL = ex.alloc<u64>(2), *sS = ex.alloc<u64>(3), *sO = ex.alloc<u64>(3);
        ex.upload(sL, hl, 16); ex.upload(sS, hs, 24); ex.upload(sO, ho, 24);
        u8 *d_sur = ex.alloc<u8>(surplus + 64);
        S.L = sL; S.seq_start = sS; S.out_start = sO; S.N = 1; S.total = surplus; S.out = d_sur;
        const u64 ntiles = (surplus + WT_TILE - 1) / WT_TILE;
        u32 *tile_first = ex.alloc<u32>(ntiles + 2);
        KLAUNCH(ex, "k_tile_first", k_tile_first<<<(unsigned)((ntiles + 1 + 255) / 256), 256, 0, ex.stream>>>(sO, 1, ntiles, tile_first));
        KLAUNCH(ex, "k_write_text", k_write_text<<<(unsigned)ntiles, WT_THREADS, 0, ex.stream>>>(S, tile_first));
        std::vector<u8> raw(surplus), wrapped;
        ex.download(raw.data(), d_sur, surplus);
        wrapped.reserve(surplus_text);
        if (W == 0) wrapped = raw;
        else {
            u64 pos = 0, sz = surplus, lr = line_rem;
            while (sz > lr) { wrapped.insert(wrapped.end(), raw.begin() + pos, raw.begin() + pos + lr); wrapped.push_back('\n'); pos += lr; sz -= lr; lr = W; }
            wrapped.insert(wrapped.end(), raw.begin() + pos, raw.begin() + pos + sz);
        }
        if (wrapped.size() != surplus_text) fail(NAFGPU_E_CUDA, "internal error: surplus layout\n");
        ex.upload(d_text + total, wrapped.data(), wrapped.size());
        CUDA_TRY(cudaStreamSynchronize(ex.stream));                      // `wrapped` is on this frame
        total += surplus_text;
    }
    if (view == NAFGPU_OUT_CHARCOUNT) {
        unsigned long long *counts = ex.alloc<unsigned long long>(256);
        ex.zero(counts, 256 * 8);
        KLAUNCH(ex, "k_charcount", k_charcount<<<148 * 8, 256, 0, ex.stream>>>(d_text, total, counts));
        return DecodeOut{(const u8 *)counts, 256 * 8};
    }
    ex.check();
    return DecodeOut{d_text, total};
}

}  // namespace nafg
