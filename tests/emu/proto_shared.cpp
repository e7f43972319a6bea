// proto_shared — TEST-SIDE PROTOTYPE (not part of the library): LZ blocks that share ONE Huffman code and ONE set of FSE
// tables per stream.  The shipped thread-per-block coder (nafz::zlz_emit_block) gives every 8 KB block its own literal
// histogram, sort, Huffman code and three normalised FSE tables -- private working sets in HBM for tens of thousands of
// threads.  The format offers the alternative the reference's own files are full of: the first block of a frame carries
// the tables (Compressed_Literals + FSE_Compressed), every later block says Treeless_Literals / Repeat_Mode.  Then a block
// costs only serial coding against read-only tables (which fit in shared memory, once per CTA), and a decoder builds each
// table once per stream.  The tables come from the statistics of a sample of the stream's blocks (every 8th: the first block
// alone is not representative -- ids start with one digit).  Sequence codes are smoothed so that every code a block can
// produce has a probability (count * 16 + 1); a block with a literal byte the sample never saw stores its literals raw.
// Neither fallback changes the decoder's entropy state.
// Blocks stay self-contained in what they reference (matches inside the block, repeat offsets the block pushed itself).
//   proto_shared IN OUT.zst BLOCK_SIZE [col]     (col: sequences from the column match finder of proto_lzcol.cpp instead of the
//   serial hash parse -- the two prototypes together; prints sizes; frames are checked by libzstd and the oracle in the test)
#include "lzcol.hpp"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
using namespace nafz;

struct Shared {
    ZEncMeta M;                                   // literal code + tree description
    short nl[36], no[32], nm[53]; u16 cuml[37], cumo[33], cumm[54];
    u16 spos[512 + 256 + 512]; u8 tsym[512];
    int logl = 9, logo = 8, logm = 9, nsl = 0, nso = 0, nsm = 0;
    FseEnc EL, EO, EM;
    std::vector<u8> desc;                         // the three table descriptions, as the first block writes them
};

struct Counts { u32 lit[256] = {0}, ll[36] = {0}, of[32] = {0}, ml[53] = {0}; u64 nlit = 0, nseq = 0; };
static void count_block(Counts &C, const u8 *lit, u32 nlit, const ZLzSeqs &S)
{
    for (u32 i = 0; i < nlit; i++) C.lit[lit[i]]++;
    for (u32 i = 0; i < S.n; i++) { C.ll[zlz_ll_code(S.ll[i])]++; C.of[hibit(S.ov[i])]++; C.ml[zlz_ml_code(S.ml[i])]++; }
    C.nlit += nlit; C.nseq += S.n;
}
// tables from the statistics of a sample of the stream's blocks, smoothed: every byte and every code a block of this size can
// produce keeps a code, so no block ever needs a fallback that would change the decoder's state
static bool build_shared(const Counts &C, u32 bs, Shared &T)
{
    if (C.nlit == 0 || C.nseq == 0) return false;
    // literals: the bytes the sample saw (a block with a byte it did not see stores its literals raw, which leaves the decoder's
    // tree alone; giving all 256 bytes a code would push 246 ten-bit codes into the tree and the digits down by a bit or two)
    u64 mx = 0;
    for (int s = 0; s < 256; s++) if (C.lit[s] > mx) mx = C.lit[s];
    u64 scale = 1; while (mx / scale > 60000) scale *= 2;
    u16 h16[256]; for (int s = 0; s < 256; s++) { h16[s] = (u16)(C.lit[s] / scale); if (C.lit[s] && !h16[s]) h16[s] = 1; }
    zenc_huf_build(h16, T.M);
    if (T.M.mode != 2) return false;
    u32 cl[36] = {0}, co[32] = {0}, cm[53] = {0};
    T.nsl = (int)zlz_ll_code(bs) + 1; T.nsm = (int)zlz_ml_code(bs) + 1; T.nso = hibit(bs + 3) + 1;
    u64 sc = 1; while (C.nseq * 16 / sc > (1u << 26)) sc *= 2;
    u32 tl = 0, to = 0, tm = 0;
    for (int s = 0; s < T.nsl; s++) { cl[s] = (u32)((u64)C.ll[s] * 16 / sc) + 1; tl += cl[s]; }
    for (int s = 0; s < T.nso; s++) { co[s] = (u32)((u64)C.of[s] * 16 / sc) + 1; to += co[s]; }
    for (int s = 0; s < T.nsm; s++) { cm[s] = (u32)((u64)C.ml[s] * 16 / sc) + 1; tm += cm[s]; }
    if (!fse_normalize(cl, T.nsl, tl, T.logl, T.nl) || !fse_normalize(co, T.nso, to, T.logo, T.no) || !fse_normalize(cm, T.nsm, tm, T.logm, T.nm)) return false;
    fse_build_enc(T.nl, T.nsl, T.logl, T.spos, T.cuml, T.tsym);
    fse_build_enc(T.no, T.nso, T.logo, T.spos + 512, T.cumo, T.tsym);
    fse_build_enc(T.nm, T.nsm, T.logm, T.spos + 768, T.cumm, T.tsym);
    T.EL = FseEnc{T.nl, T.cuml, T.spos, T.logl, 0, 2, 0};
    T.EO = FseEnc{T.no, T.cumo, T.spos + 512, T.logo, 0, 2, 0};
    T.EM = FseEnc{T.nm, T.cumm, T.spos + 768, T.logm, 0, 2, 0};
    T.desc.assign(512, 0);
    BitW hw; hw.init(T.desc.data(), (u32)T.desc.size());
    fse_write_ncount(hw, T.nl, T.nsl, T.logl); fse_write_ncount(hw, T.no, T.nso, T.logo); fse_write_ncount(hw, T.nm, T.nsm, T.logm);
    if (!hw.ok) return false;
    T.desc.resize(hw.pos);
    return true;
}

// Literals_Section with the shared code: type 2 (+ tree) in the first block, type 3 (treeless) later; raw when that is smaller
static u32 put_literals_shared(const u8 *lit, u32 nlit, const Shared &T, bool first, u8 *out, u32 cap)
{
    auto raw = [&]() -> u32 {
        u32 h;
        if (nlit < 32) { out[0] = (u8)(nlit << 3); h = 1; }
        else if (nlit < 4096) { const u32 v = (1u << 2) | (nlit << 4); out[0] = (u8)v; out[1] = (u8)(v >> 8); h = 2; }
        else { const u32 v = (3u << 2) | (nlit << 4); out[0] = (u8)v; out[1] = (u8)(v >> 8); out[2] = (u8)(v >> 16); h = 3; }
        memcpy(out + h, lit, nlit);
        return h + nlit;
    };
    if (nlit == 0) return raw();
    const ZEncMeta &M = T.M;
    for (u32 i = 0; i < nlit; i++) if (!(M.ctab[lit[i]] >> 12)) { if (first) return 0; return raw(); }      // a byte without a code
    const u32 tree = first ? M.tree_len : 0;
    u32 total_bits = 0; for (u32 i = 0; i < nlit; i++) total_bits += M.ctab[lit[i]] >> 12;
    u32 nstreams = (nlit <= 1023 && tree + total_bits / 8 + 1 <= 1023) ? 1 : 4;
    if (nstreams == 4 && nlit < 16) { if (first) return 0; return raw(); }
    const u32 seg = nstreams == 4 ? (nlit + 3) / 4 : nlit;
    u32 sbytes[4] = {0, 0, 0, 0}, payload = tree + (nstreams == 4 ? 6u : 0u);
    for (u32 k = 0; k < nstreams; k++) {
        const u32 a = k * seg, b = k == nstreams - 1 ? nlit : (a + seg < nlit ? a + seg : nlit);
        u32 bits = 0; for (u32 i = a; i < b; i++) bits += M.ctab[lit[i]] >> 12;
        sbytes[k] = bits / 8 + 1; payload += sbytes[k];
    }
    const u32 lh = nstreams == 1 ? 3 : ((nlit <= 16383 && payload <= 16383) ? 4 : 5);
    const u32 raw_size = nlit + (nlit < 32 ? 1 : (nlit < 4096 ? 2 : 3));
    if (!first && lh + payload >= raw_size) return raw();               // (the first block must carry the tree whatever it costs)
    if (lh + payload > cap) return 0;
    const u32 type = first ? 2u : 3u;
    if (nstreams == 1) { const u32 v = type | (0 << 2) | (nlit << 4) | (payload << 14); out[0] = (u8)v; out[1] = (u8)(v >> 8); out[2] = (u8)(v >> 16); }
    else if (lh == 4) { const u32 v = type | (2 << 2) | (nlit << 4) | (payload << 18); out[0] = (u8)v; out[1] = (u8)(v >> 8); out[2] = (u8)(v >> 16); out[3] = (u8)(v >> 24); }
    else { const u64 v = type | (3 << 2) | ((u64)nlit << 4) | ((u64)payload << 22); for (int i = 0; i < 5; i++) out[i] = (u8)(v >> (8 * i)); }
    u32 at = lh;
    for (u32 i = 0; i < tree; i++) out[at++] = M.tree[i];
    if (nstreams == 4) for (int j = 0; j < 3; j++) { out[at++] = (u8)sbytes[j]; out[at++] = (u8)(sbytes[j] >> 8); }
    for (u32 k = 0; k < nstreams; k++) {
        const u32 a = k * seg, b = k == nstreams - 1 ? nlit : (a + seg < nlit ? a + seg : nlit);
        BitW bw; bw.init(out + at, sbytes[k]);
        for (u32 i = b; i > a; i--) { const u32 e = M.ctab[lit[i - 1]]; bw.put(e & 0xFFF, e >> 12); }
        bw.finish_with_mark();
        at += sbytes[k];
    }
    return at;
}

// a Compressed_Block coded against the stream's tables; 0: store the block raw (never for the first block: then the caller gives up sharing)
static u64 g_lit_bytes = 0, g_seq_bytes = 0, g_nlit = 0, g_nseq = 0;
static u32 emit_shared(u32 n, const u8 *lit, u32 nlit, const ZLzSeqs &S, Shared T, bool first, u8 *out, u32 cap)
{
    u32 at = put_literals_shared(lit, nlit, T, first, out, cap);
    if (!at) return 0;
    const u32 lit_at = at;
    const u32 nseq = S.n;
    if (nseq == 0) { if (first) return 0; out[at++] = 0; return at < n ? at : 0; }
    if (nseq < 128) out[at++] = (u8)nseq;
    else if (nseq < 0x7F00) { out[at++] = (u8)((nseq >> 8) + 128); out[at++] = (u8)nseq; }
    else { out[at++] = 255; out[at++] = (u8)(nseq - 0x7F00); out[at++] = (u8)((nseq - 0x7F00) >> 8); }
    out[at++] = first ? (u8)((2 << 6) | (2 << 4) | (2 << 2)) : (u8)((3 << 6) | (3 << 4) | (3 << 2));      // FSE_Compressed x 3 / Repeat_Mode x 3
    if (first) { memcpy(out + at, T.desc.data(), T.desc.size()); at += (u32)T.desc.size(); }
    BitW bw; bw.init(out + at, cap - at);
    FseEnc &EL = T.EL, &EO = T.EO, &EM = T.EM;
    auto codes = [&](u32 i, u32 &c_l, u32 &c_o, u32 &c_m) { c_l = zlz_ll_code(S.ll[i]); c_o = (u32)hibit(S.ov[i]); c_m = zlz_ml_code(S.ml[i]); };
    u32 c_l, c_o, c_m;
    codes(nseq - 1, c_l, c_o, c_m);
    EM.start(c_m); EO.start(c_o); EL.start(c_l);
    bw.put(S.ll[nseq - 1] - ll_base_of(c_l), ll_bits_of(c_l));
    bw.put(S.ml[nseq - 1] - ml_base_of(c_m), ml_bits_of(c_m));
    bw.put(S.ov[nseq - 1] - (1u << c_o), c_o);
    for (u32 i = nseq - 1; i-- > 0;) {
        codes(i, c_l, c_o, c_m);
        EO.put(bw, c_o); EM.put(bw, c_m); EL.put(bw, c_l);
        bw.put(S.ll[i] - ll_base_of(c_l), ll_bits_of(c_l));
        bw.put(S.ml[i] - ml_base_of(c_m), ml_bits_of(c_m));
        bw.put(S.ov[i] - (1u << c_o), c_o);
    }
    EM.flush(bw); EO.flush(bw); EL.flush(bw);
    bw.finish_with_mark();
    if (!bw.ok) return 0;
    at += bw.pos;
    g_lit_bytes += lit_at; g_seq_bytes += at - lit_at; g_nlit += nlit; g_nseq += nseq;
    return (first || at < n) ? at : 0;
}

int main(int argc, char **argv)
{
    if (argc < 4) return 2;
    FILE *f = fopen(argv[1], "rb"); if (!f) return 2;
    std::vector<u8> in; u8 buf[65536]; size_t k;
    while ((k = fread(buf, 1, sizeof buf, f)) > 0) in.insert(in.end(), buf, buf + k);
    fclose(f);
    const u32 bs = (u32)atoi(argv[3]);
    if (bs < 64 || bs > ZLZ_MAX_BLOCK) return 2;
    const bool use_col = argc > 4 && !strcmp(argv[4], "col");
    const u32 max_seq = bs / 4;
    std::vector<u8> out = { 0x28, 0xB5, 0x2F, 0xFD, 0x00, (u8)((17 - 10) << 3) };
    std::vector<u16> htab(1u << ZLZ_HLOG), sll(max_seq), sml(max_seq), sov(max_seq), spos(1280);
    std::vector<u8> lit(bs + 16), tsym(512), codes(3 * max_seq), slot(bs + 1024);
    const size_t n = in.size(), nblk = n ? (n + bs - 1) / bs : 1;
    // pass 1: parse every block (kept), statistics from every SAMPLE-th block that has sequences
    const size_t SAMPLE = 8;
    struct Parsed { std::vector<u8> lit; std::vector<u16> ll, ml, ov; bool rle = false, parsed = false; };
    std::vector<Parsed> P(nblk);
    Counts C;
    for (size_t b = 0; b < nblk; b++) {
        const u32 len = (u32)(n - b * bs < bs ? n - b * bs : bs);
        const u8 *src = in.data() + b * bs;
        if (len) { u32 i = 1; while (i < len && src[i] == src[0]) i++; P[b].rle = i == len; }
        if (P[b].rle || len < 16) continue;
        ZLzSeqs S{sll.data(), sml.data(), sov.data(), 0};
        const u32 nlit = use_col ? find_columns(src, len, lit.data(), S, max_seq) : zlz_find(src, len, htab.data(), 1, lit.data(), S, max_seq);
        P[b].lit.assign(lit.begin(), lit.begin() + nlit);
        P[b].ll.assign(sll.begin(), sll.begin() + S.n); P[b].ml.assign(sml.begin(), sml.begin() + S.n); P[b].ov.assign(sov.begin(), sov.begin() + S.n);
        P[b].parsed = true;
        if (b % SAMPLE == 0) count_block(C, lit.data(), nlit, S);
    }
    // the block that will carry the tables: the first one with literals and sequences; its own statistics count too when it is
    // not a sampled block (so that each of its literal bytes has a code)
    size_t fdef = 0;
    while (fdef < nblk && !(P[fdef].parsed && !P[fdef].lit.empty() && !P[fdef].ll.empty())) fdef++;
    if (fdef < nblk && fdef % SAMPLE) {
        ZLzSeqs S{P[fdef].ll.data(), P[fdef].ml.data(), P[fdef].ov.data(), (u32)P[fdef].ll.size()};
        count_block(C, P[fdef].lit.data(), (u32)P[fdef].lit.size(), S);
    }
    Shared T; bool shared = bs >= 64 && fdef < nblk && build_shared(C, bs, T);
    size_t n_shared = 0, n_own = 0, n_raw = 0;
    // pass 2: raw / RLE blocks before the defining block do not touch the decoder's entropy state, a compressed block before it would
    // -- so until the tables are defined blocks are stored raw.  Should the defining block not fit its slot when coded with the
    // stream's tables, the whole stream gets per-block tables instead.
    std::vector<std::vector<u8>> body(nblk); std::vector<u32> csz(nblk, 0);
    for (int attempt = 0; attempt < 2; attempt++) {
        bool def_fail = false;
        n_shared = n_own = n_raw = 0;
        for (size_t b = 0; b < nblk; b++) {
            const u32 len = (u32)(n - b * bs < bs ? n - b * bs : bs);
            u32 cs = 0;
            if (P[b].parsed) {
                ZLzSeqs S{P[b].ll.data(), P[b].ml.data(), P[b].ov.data(), (u32)P[b].ll.size()};
                ZLzWork W{spos.data(), tsym.data(), codes.data()};
                const u32 nlit = (u32)P[b].lit.size();
                if (shared) {
                    if (b >= fdef) { cs = emit_shared(len, P[b].lit.data(), nlit, S, T, b == fdef, slot.data(), bs + 512); if (!cs && b == fdef) def_fail = true; }
                    if (cs) n_shared++;
                } else { cs = zlz_emit_block(len, P[b].lit.data(), nlit, S, max_seq, W, slot.data(), bs + 512); if (cs) n_own++; }
            }
            if (!cs) n_raw++;
            csz[b] = cs; body[b].assign(slot.begin(), slot.begin() + cs);
        }
        if (!def_fail) break;
        shared = false;
    }
    for (size_t b = 0; b < nblk; b++) {
        const u32 len = (u32)(n - b * bs < bs ? n - b * bs : bs);
        const u8 *src = in.data() + b * bs;
        const bool rle = P[b].rle; const u32 cs = csz[b];
        const u32 last = b + 1 == nblk, type = cs ? 2 : (rle ? 1 : 0), size_field = type == 2 ? cs : len;
        const u32 bh = last | (type << 1) | (size_field << 3);
        out.push_back((u8)bh); out.push_back((u8)(bh >> 8)); out.push_back((u8)(bh >> 16));
        if (type == 2) out.insert(out.end(), body[b].begin(), body[b].end());
        else if (type == 1) out.push_back(src[0]);
        else out.insert(out.end(), src, src + len);
    }
    FILE *o = fopen(argv[2], "wb"); if (!o) return 2;
    fwrite(out.data(), 1, out.size(), o); fclose(o);
    printf("in=%zu out=%zu blocks=%zu shared=%zu own=%zu raw_or_rle=%zu lit_bytes=%llu nlit=%llu seq_bytes=%llu nseq=%llu\n", n, out.size(), nblk, n_shared, n_own, n_raw,
           (unsigned long long)g_lit_bytes, (unsigned long long)g_nlit, (unsigned long long)g_seq_bytes, (unsigned long long)g_nseq);
    return 0;
}
