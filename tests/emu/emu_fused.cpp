// emu_fused — TEST-ONLY: runs the phases of the single-pass encode transform (naf_b200/csrc/naf_fused_hd.cuh) on the
// CPU, thread after thread and tile after tile (the kernel's block scans and look-backs restated as serial loops), and
// writes the streams for comparison with the oracle.
//   emu_fused IN OUTPREFIX seq_type(0..3) no_mask(0/1) [threads]
// exit 0: ok (streams written), 3: input is not canonical (the library falls back to the general parser), 2: usage
#include <cstdint>
struct uint4 { uint32_t x, y, z, w; };
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
#include "../../naf_b200/csrc/naf_fused_hd.cuh"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
using namespace nafg;

static void dump(const std::string &path, const void *p, size_t n) { FILE *f = fopen(path.c_str(), "wb"); if (n) fwrite(p, 1, n, f); fclose(f); }

static int run_fused(const std::vector<u8> &text_in, u64 p0, bool fastq, int seq_type, int no_mask, u32 NT, const std::string &prefix)
{
    const u64 n = text_in.size(), ntiles = (n + FT_BYTES - 1) / FT_BYTES;
    std::vector<u8> gtext(ntiles * FT_BYTES + 64, 0);
    memcpy(gtext.data(), text_in.data(), n);

    // nuc_code + "unexpected" bit, as naf_enc.cu builds it (tables.c:189, :72, :82)
    u8 lut[256];
    for (int c = 0; c < 256; c++) {
        int u = (c >= 'a' && c <= 'z') ? c - 32 : c;
        const char *order = "-TGKCYSBAWRDMHV"; const char *q = u ? strchr(order, u) : nullptr;
        lut[c] = u == 'U' ? 1 : (q ? (u8)(q - order) : 15);
        const char *ok = seq_type == 1 ? "-ABCDGHKMNRSUVWY" : "-ABCDGHKMNRSTVWY";
        if (!(u && strchr(ok, u))) lut[c] |= 0x80;
    }
    // worst-case destinations, 32-byte aligned
    auto alloc = [](size_t nbytes) { void *p = nullptr; if (posix_memalign(&p, 64, nbytes + 256)) abort(); memset(p, 0xEE, nbytes + 256); return (u8 *)p; };
    FusedCfg C; memset(&C, 0, sizeof C);
    C.n = n; C.p0 = p0; C.fastq = fastq;
    C.seq_mode = seq_type == 2 ? FS_PROTEIN : (seq_type == 3 ? (fastq ? FS_TEXT : FS_TEXT_GT) : FS_PACK4);
    C.upper = seq_type >= 2 && no_mask; C.want_mask = seq_type < 2 && !no_mask;
    C.id_check = (seq_type == 3 && !fastq) ? FC_ID_GT : FC_ID;
    u32 lut32[256];
    for (int c = 0; c < 256; c++) lut32[c] = nuc_lut32(lut[c]);
    C.lut = lut32;
    C.ids = alloc(n + 2) + 3; C.comm = alloc(n + 2) + 5;                      // odd alignments on purpose
    C.seq = alloc(n + 2); C.qual = alloc(n + 2) + 1;
    C.len = (u32 *)alloc((ntiles * FT_MAXSEG + 2) * 4);
    C.casebits = (u32 *)alloc(n / 8 + 64); memset(C.casebits, 0, n / 8 + 64);

    // "shared memory"
    u8 *tile = alloc(FT_BYTES + 16), *stage = alloc(FT_STAGE);
    std::vector<u16> seg_end(FT_MAXSEG + 8), d_src(FT_MAXDESC + 8), d_len(FT_MAXDESC + 8), d_dst(FT_MAXDESC + 8), recseq(FT_MAXSEG + 8), recqual(FT_MAXSEG + 8);
    FusedShared sh;
    FusedTile T; T.text = tile; T.stage = stage; T.seg_end = seg_end.data(); T.d_src = d_src.data(); T.d_len = d_len.data(); T.d_dst = d_dst.data();
    T.recseq = recseq.data(); T.recqual = recqual.data(); T.sh = &sh;

    u32 run1 = fastq ? 0u : (u32)FE_HDR;           // look-back #1 inclusive state of the tiles so far
    F2 run2 = f2_initial();
    u32 gflag = 0; u64 glongest = 0;
    auto atomic_or = [](u32 *p, u32 v) { *p |= v; };

    for (u64 t = 0; t < ntiles; t++) {
        const u64 lo = t * FT_BYTES;
        memset(&sh, 0, sizeof sh);
        memset(stage, 0xDD, FT_STAGE);
        memcpy(tile, gtext.data() + lo, FT_BYTES); memset(tile + FT_BYTES, 0, 16);
        if (lo + FT_BYTES < n) tile[FT_BYTES] = gtext[lo + FT_BYTES];
        sh.tile = (u32)t;
        sh.live_lo = p0 > lo ? (p0 - lo >= FT_BYTES ? FT_BYTES : (u32)(p0 - lo)) : 0;
        sh.live_hi = n >= lo + FT_BYTES ? FT_BYTES : (n > lo ? (u32)(n - lo) : 0);
        if (sh.live_lo > sh.live_hi) sh.live_lo = sh.live_hi;
        // phase 1: newline masks; thread th owns chunks th, th + NT, ...
        std::vector<u32> cnt(FT_CHUNKS), mask(FT_CHUNKS);
        for (u32 th = 0; th < NT; th++) for (u32 c = th; c < FT_CHUNKS; c += NT) { mask[c] = T.chunk_mask(c); cnt[c] = (u32)__builtin_popcount(mask[c]); }
        u32 Tn = 0; std::vector<u32> first(FT_CHUNKS);
        for (u32 c = 0; c < FT_CHUNKS; c++) { first[c] = Tn; Tn += cnt[c]; }
        sh.nseg = Tn + 1;
        if (sh.nseg > FT_MAXSEG) { sh.abort_ = 1; sh.flag |= FU_LINES; }
        u32 last_nl = 0;
        if (!sh.abort_) for (u32 th = 0; th < NT; th++) for (u32 c = th; c < FT_CHUNKS; c += NT) T.put_lines(c, mask[c], first[c]);
        if (Tn && !sh.abort_) last_nl = seg_end[Tn - 1];
        // look-back #1
        const u32 agg1 = fastq ? Tn : (sh.abort_ ? (u32)FE_ID : T.fasta_element(Tn, last_nl));
        sh.entry1 = run1;
        run1 = f1_compose(fastq, run1, agg1);
        {
            // the up to 64 bytes before the tile (the kernel prefetches them): line start?  header already past its first space?
            const u64 at = lo + sh.live_lo;
            const u32 np = at > p0 ? (u32)(at - p0 < 64 ? at - p0 : 64) : 0;
            const u8 *prev = gtext.data() + at - np;
            if (fastq) sh.entry_ls = np && prev[np - 1] == '\n';
            else sh.entry_ls = sh.entry1 == FE_LS;
            sh.entry_sp = 0;
            const u32 role0 = fastq ? (sh.entry1 & 3) : (sh.entry1 == FE_HDR ? (u32)FR_HDR : (u32)FR_SEQ);
            if (sh.live_lo < sh.live_hi && role0 == FR_HDR && !sh.entry_ls) {
                bool resolved; sh.entry_sp = prev_scan(prev, np, resolved);
                if (!resolved && at - np > p0) { u32 f = 0; sh.entry_sp = fast_lookback_space(gtext.data(), p0, at - np, f); if (f) sh.flag |= FU_LOOKBACK; }
            }
        }
        F2 agg2; memset(&agg2, 0, sizeof agg2);
        if (!sh.abort_) {
            // phase 3: thread th owns segments th, th + NT, ...
            std::vector<u64> sa(sh.nseg), sb(sh.nseg); std::vector<u32> rbv(sh.nseg), spv(sh.nseg);
            for (u32 th = 0; th < NT; th++) for (u32 j = th; j < sh.nseg; j += NT) T.classify(C, j, sa[j], sb[j], rbv[j], spv[j], sh.flag);
            u64 ra = 0, rb = 0; std::vector<u64> pa(sh.nseg), pb(sh.nseg);
            for (u32 j = 0; j < sh.nseg; j++) { pa[j] = ra; pb[j] = rb; ra += sa[j]; rb += sb[j]; }
            sh.t_ids = ra & 0xFFFF; sh.t_comm = (ra >> 16) & 0xFFFF; sh.t_seq = (ra >> 32) & 0xFFFF; sh.t_qual = (u32)(ra >> 48);
            sh.t_rec = rb & 0xFFFF; sh.n_hdr = (rb >> 16) & 0xFFFF; sh.n_seq = (rb >> 32) & 0xFFFF; sh.n_qual = (u32)(rb >> 48);
            T.layout();
            for (u32 th = 0; th < NT; th++) for (u32 j = th; j < sh.nseg; j += NT) T.place(C, j, pa[j], pb[j], rbv[j], spv[j]);
            agg2 = T.aggregate(C);
        }
        // look-back #2 (the kernel: one warp, while the others copy)
        sh.pre = run2;
        run2 = f2_compose(run2, agg2, fastq);
        if (!sh.abort_) {
            const u32 ndesc = sh.n_seq + sh.n_qual + 2 * sh.n_hdr, ngroups = NT / FT_GROUP ? NT / FT_GROUP : 1;
            for (u32 g = 0; g < ngroups; g++) for (u32 k = g; k < ndesc; k += ngroups) for (u32 lane = 0; lane < FT_GROUP; lane++) sh.flag |= T.copy_desc(C, k, lane);
            for (u32 th = 0; th < NT; th++) for (u32 k = th; k < sh.t_rec; k += NT) { const u64 L = T.finish_record(C, k, sh.flag); if (L > sh.maxlen) sh.maxlen = L; }
            if (!fastq) for (u32 th = 0; th < NT; th++) for (u32 k = th; k < sh.n_seq; k += NT) { const u64 L = T.line_length(k); if (L > sh.maxlen) sh.maxlen = L; }
            // phase 5: staging -> global, region by region
            struct Region { u32 off, len; u8 *dst; int check; bool upper; u32 fl; };
            Region regs[4] = { { sh.s_ids, sh.t_ids, C.ids + sh.pre.ids, FC_NONE, false, FU_BADBYTE },
                               { sh.s_comm, sh.t_comm, C.comm + sh.pre.comm, FC_NONE, false, FU_BADBYTE },
                               { sh.s_qual, sh.t_qual, C.qual + sh.pre.qual, FC_QUAL, false, FU_QUAL },
                               { sh.s_seq, C.seq_mode == FS_PACK4 ? 0u : sh.t_seq, C.seq + sh.pre.seq, FC_PROTEIN + (C.seq_mode - FS_PROTEIN), C.upper != 0, FU_SEQ } };
            for (auto &R : regs) {
                u32 head = (u32)((16 - ((uintptr_t)R.dst & 15)) & 15); if (head > R.len) head = R.len;
                const u32 nu = (R.len - head) >> 4, done = head + (nu << 4);
                for (u32 i = 0; i < head; i++) if (T.out_byte(R.off, i, R.dst, R.check, R.upper)) sh.flag |= R.fl;
                for (u32 i = done; i < R.len; i++) if (T.out_byte(R.off, i, R.dst, R.check, R.upper)) sh.flag |= R.fl;
                for (u32 th = 0; th < NT; th++) for (u32 u = th; u < nu; u += NT) if (T.out_unit(R.off + head, u, R.dst + head, R.check, R.upper)) sh.flag |= R.fl;
            }
            if (C.seq_mode == FS_PACK4 && sh.t_seq) {
                const u32 A = (u32)(sh.pre.seq & 31), npieces = (A + sh.t_seq + 31) / 32;
                for (u32 th = 0; th < NT; th++) for (u32 q = th; q < npieces; q += NT) sh.flag |= T.pack_piece(C, q, atomic_or);
            }
        }
        gflag |= sh.flag;
        if (sh.maxlen > glongest) glongest = sh.maxlen;
    }
    FusedTotals tot;
    fused_finish(C, run1, run2, gflag, glongest, gtext.data(), tot);
    if (tot.flag) { fprintf(stderr, "declined: flag %u\n", tot.flag); return 3; }
    dump(prefix + ".ids", C.ids, tot.n_ids); dump(prefix + ".comm", C.comm, tot.n_comm);
    dump(prefix + ".seq", C.seq, C.seq_mode == FS_PACK4 ? (tot.n_bases + 1) / 2 : tot.n_bases);
    dump(prefix + ".qual", C.qual, tot.n_qual);
    dump(prefix + ".len", C.len, tot.n_rec * 4);
    dump(prefix + ".casebits", C.casebits, ((tot.n_bases + 31) / 32) * 4);
    FILE *f = fopen((prefix + ".info").c_str(), "w");
    fprintf(f, "%llu %llu %u %llu\n", (unsigned long long)tot.n_rec, (unsigned long long)tot.longest, tot.end_state, (unsigned long long)tot.n_bases);
    fclose(f);
    return 0;
}

int main(int argc, char **argv)
{
    if (argc < 5) return 2;
    FILE *f = fopen(argv[1], "rb"); if (!f) return 2;
    std::vector<u8> text; u8 buf[65536]; size_t k;
    while ((k = fread(buf, 1, sizeof buf, f)) > 0) text.insert(text.end(), buf, buf + k);
    fclose(f);
    const int seq_type = atoi(argv[3]), no_mask = atoi(argv[4]);
    const u32 NT = argc > 5 ? (u32)atoi(argv[5]) : 64;
    auto is_space = [](int ch) { return (ch >= 0x09 && ch <= 0x0D) || ch == 0x20; };
    size_t i = 0; u32 last = '\n';
    while (i < text.size() && is_space(text[i])) { last = text[i]; i++; }
    if (i == text.size()) return 3;
    const bool at_ls = last >= 0x0A && last <= 0x0D;
    if (!at_ls || (text[i] != '>' && text[i] != '@')) return 3;
    return run_fused(text, i + 1, text[i] == '@', seq_type, no_mask, NT, argv[2]);
}
