// emu_parse — TEST-ONLY: runs the per-chunk logic of the canonical-input parser (naf_b200/csrc/naf_fast_hd.cuh:
// chunk scan, FASTA element algebra, line walk, padded staging sink, SWAR byte checks) on the CPU, with the
// kernels' tile / thread plumbing restated as serial loops, and writes the raw streams for comparison with the oracle.
//   emu_parse IN OUTPREFIX seq_type(0..3) no_mask(0/1)
// exit 0: ok (streams written), 3: input is not canonical (the library would fall back to the general parser), 2: usage
#include "../../naf_b200/csrc/naf_fast_hd.cuh"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
using namespace nafg;

static const int PT = 256, PB = 64, PTILE = PT * PB;

struct HostRow { const u8 *tile; u32 src0; u32 operator()(u32 i) const { return tile[fast_pad(src0 + i)]; } };

static void load_words(const std::vector<u8> &text, u64 p0, u64 lo, u32 w[16], u32 &b0, u32 &b1)
{
    const u64 n = text.size();
    b0 = lo >= p0 ? 0u : (p0 - lo >= 64 ? 64u : (u32)(p0 - lo));
    b1 = lo + 64 <= n ? 64u : (lo >= n ? 0u : (u32)(n - lo));
    for (int k = 0; k < 16; k++) {
        u32 x = FAST_FILL;
        for (int j = 0; j < 4; j++) { u32 i = 4 * k + j; if (i >= b0 && i < b1) x = (x & ~(0xFFu << (8 * j))) | ((u32)text[lo + i] << (8 * j)); }
        w[k] = x;
    }
}

static void dump(const std::string &path, const void *p, size_t n) { FILE *f = fopen(path.c_str(), "wb"); if (n) fwrite(p, 1, n, f); fclose(f); }

template <bool FASTQ> static int run(const std::vector<u8> &text, u64 p0, int seq_type, int no_mask, const std::string &prefix)
{
    const u64 n = text.size(), ntiles = (n + PTILE - 1) / PTILE, nthreads = ntiles * PT;
    u32 flag = 0;
    std::vector<u64> nls(nthreads), sps(nthreads); std::vector<u32> elem(nthreads), b0s(nthreads), b1s(nthreads);
    std::vector<u8> tile((PTILE + 64) / 64 * 68), stage((PTILE + 256) / 64 * 68);
    // pass 1
    for (u64 t = 0; t < nthreads; t++) {
        u32 w[16], bad; load_words(text, p0, t * PB, w, b0s[t], b1s[t]);
        fast_chunk_scan<true>(w, nls[t], bad, &sps[t]);
        if (bad) flag |= FF_BADBYTE;
        const u32 src0 = (u32)(t % PT) * 64;
        for (int k = 0; k < 16; k++) *(u32 *)(tile.data() + fast_pad(src0 + 4 * k)) = w[k];
        HostRow row{tile.data(), src0};
        elem[t] = FASTQ ? (u32)__builtin_popcountll(nls[t]) : fasta_chunk_element(row, nls[t], b0s[t], b1s[t]);
    }
    // entry states + pass 2
    struct TI { FastState st; FastEmit n; u32 mark; };
    std::vector<TI> ti(nthreads);
    u32 run_state = FASTQ ? 0u : (u32)FE_HDR;
    for (u64 t = 0; t < nthreads; t++) {
        const u64 lo = t * PB;
        u32 w[16], b0, b1; load_words(text, p0, lo, w, b0, b1);
        const u32 src0 = (u32)(t % PT) * 64;
        for (int k = 0; k < 16; k++) *(u32 *)(tile.data() + fast_pad(src0 + 4 * k)) = w[k];
        HostRow row{tile.data(), src0};
        FastState st; st.sp = 0;
        if (FASTQ) { st.role = run_state & 3; st.ls = lo > p0 && b1 > 0 && text[lo - 1] == '\n'; }
        else { st.role = run_state == FE_HDR ? (u32)FR_HDR : (u32)FR_SEQ; st.ls = run_state == FE_LS; }
        if (b0 < b1 && st.role == FR_HDR && !st.ls) st.sp = fast_lookback_space(text.data(), p0, lo + b0, flag);
        ti[t].st = st;
        FastEmit e = {0, 0, 0, 0, 0}; FastLine ln = {0, 0, 0}; FastNoSink sink;
        fast_walk<FASTQ, false>(row, nls[t], sps[t], b0, b1, st, lo, e, sink, 0, 0, 0, ln, flag);
        ti[t].n = e; ti[t].mark = ln.mark;
        run_state = FASTQ ? run_state + elem[t] : fe_compose(run_state, elem[t]);
    }
    // end state
    u32 end_state;
    {
        u32 role, ls, sp = 0;
        if (FASTQ) { role = run_state & 3; ls = n > p0 && text[n - 1] == '\n'; }
        else { role = run_state == FE_HDR ? (u32)FR_HDR : (u32)FR_SEQ; ls = run_state == FE_LS; }
        if (role == FR_HDR && !ls) sp = fast_lookback_space(text.data(), p0, n, flag);
        end_state = fast_end_state(FASTQ, role, sp, ls);
    }
    if (flag) return 3;
    // totals
    u64 t_ids = 0, t_comm = 0, t_seq = 0, t_qual = 0, t_rec = 0;
    for (auto &x : ti) { t_ids += x.n.ids; t_comm += x.n.comm; t_seq += x.n.seq; t_qual += x.n.qual; t_rec += x.n.rec; }
    std::vector<u8> ids(t_ids + 8), comm(t_comm + 8), bases(t_seq + 8), qual(t_qual + 8);
    std::vector<u64> rec_seq_end(t_rec + 2), rec_qual_end(t_rec + 2), rec_pos(t_rec + 2);
    u64 longest = 0;
    // pass 3, tile by tile
    u64 o_ids = 0, o_comm = 0, o_seq = 0, o_qual = 0, o_rec = 0, line_base = 0;
    const int seq_check = seq_type == 2 ? 1 : (seq_type == 3 ? (FASTQ ? 2 : 3) : 0);
    const bool upper = seq_type >= 2 && no_mask;
    for (u64 tl = 0; tl < ntiles; tl++) {
        u64 tt[4] = {0, 0, 0, 0};
        for (int k = 0; k < PT; k++) { auto &x = ti[tl * PT + k]; tt[0] += x.n.ids; tt[1] += x.n.comm; tt[2] += x.n.seq; tt[3] += x.n.qual; }
        u8 *g[4] = { ids.data() + o_ids, comm.data() + o_comm, bases.data() + o_seq, qual.data() + o_qual };
        u32 s0[4]; u32 at = 0;
        for (int s = 0; s < 4; s++) { s0[s] = ((at + 3) & ~3u) + (u32)((uintptr_t)g[s] & 3); if (s == 0) s0[s] = (u32)((uintptr_t)g[s] & 3); at = s0[s] + (u32)tt[s]; }
        for (int k = 0; k < PT; k++) {            // all rows of the tile first (the kernel's threads load concurrently)
            u32 w[16], b0, b1; load_words(text, p0, (tl * PT + k) * PB, w, b0, b1);
            for (int q = 0; q < 16; q++) *(u32 *)(tile.data() + fast_pad(k * 64 + 4 * q)) = w[q];
        }
        u64 l[4] = {0, 0, 0, 0}; u64 l_rec = 0;
        for (int k = 0; k < PT; k++) {
            const u64 t = tl * PT + k, lo = t * PB;
            HostRow row{tile.data(), (u32)k * 64};
            FastSmemSink sink; sink.tile = tile.data(); sink.stage = stage.data(); sink.src0 = k * 64;
            for (int s = 0; s < 4; s++) sink.base[s] = s0[s] + (u32)l[s];
            sink.rec_seq_end = rec_seq_end.data(); sink.rec_qual_end = rec_qual_end.data(); sink.rec_pos = rec_pos.data(); sink.fastq = FASTQ;
            FastState st = ti[t].st; FastEmit m = {0, 0, 0, 0, 0}; FastLine ln = {line_base, 0, 0};
            sink.begin();
            fast_walk<FASTQ, true>(row, nls[t], sps[t], b0s[t], b1s[t], st, lo, m, sink, o_seq + l[2], o_qual + l[3], o_rec + l_rec, ln, flag);
            sink.flush();
            if (m.ids != ti[t].n.ids || m.comm != ti[t].n.comm || m.seq != ti[t].n.seq || m.qual != ti[t].n.qual || m.rec != ti[t].n.rec) { fprintf(stderr, "count/scatter mismatch\n"); return 1; }
            if (!FASTQ) {
                if (lo < n && lo + PB >= n) { const u64 d = o_seq + l[2] + m.seq - ln.base; if (d > ln.max) ln.max = d; }
                if (ln.max > longest) longest = ln.max;
            }
            line_base = ln.base;
            l[0] += m.ids; l[1] += m.comm; l[2] += m.seq; l[3] += m.qual; l_rec += m.rec;
        }
        for (int s = 0; s < 4; s++) {
            for (u64 i = 0; i < tt[s]; i++) {
                u32 c = stage[fast_pad(s0[s] + (u32)i)];
                const u32 v = c * 0x01010101u;
                if (s == 2) {
                    if ((seq_check == 1 && swar_bad_protein(v)) || (seq_check == 2 && swar_bad_text(v, false)) || (seq_check == 3 && swar_bad_text(v, true))) flag |= FF_SEQ;
                    if (upper) c = swar_upper(v) & 0xFF;
                }
                if (s == 3 && swar_bad_qual(v)) flag |= FF_QUAL;
                g[s][i] = (u8)c;
            }
        }
        o_ids += tt[0]; o_comm += tt[1]; o_seq += tt[2]; o_qual += tt[3]; o_rec += l_rec;
    }
    if (flag) return 3;
    // DNA / RNA: the pack LUT's validity bit (naf_enc.cu) -- restated here
    if (seq_type < 2) {
        const char *ok = seq_type == 1 ? "-ABCDGHKMNRSUVWY" : "-ABCDGHKMNRSTVWY";
        for (u64 i = 0; i < t_seq; i++) { int u = bases[i]; if (u >= 'a' && u <= 'z') u -= 32; if (!(u && strchr(ok, u))) return 3; }
    }
    // end-of-input additions (naf_enc.cu split_streams_impl)
    u64 n_rec = t_rec;
    if (!FASTQ) {
        if (end_state == 0) { ids[t_ids++] = 0; comm[t_comm++] = 0; } else if (end_state == 1) comm[t_comm++] = 0;
        rec_seq_end[n_rec] = t_seq; n_rec++;
    } else if (end_state == 6) { rec_seq_end[n_rec] = t_seq; rec_qual_end[n_rec] = t_qual; n_rec++; }
    else if (end_state != 7) return 3;                  // truncated FASTQ: the general path words the error
    if (FASTQ) for (u64 r = 0; r < n_rec; r++) {
        u64 sl = rec_seq_end[r] - (r ? rec_seq_end[r - 1] : 0), ql = rec_qual_end[r] - (r ? rec_qual_end[r - 1] : 0);
        if (sl != ql) return 3;
        if (sl > longest) longest = sl;
    }
    dump(prefix + ".ids", ids.data(), t_ids); dump(prefix + ".comm", comm.data(), t_comm);
    dump(prefix + ".bases", bases.data(), t_seq); dump(prefix + ".qual", qual.data(), t_qual);
    dump(prefix + ".recend", rec_seq_end.data(), n_rec * 8);
    FILE *f = fopen((prefix + ".info").c_str(), "w");
    fprintf(f, "%llu %llu %u\n", (unsigned long long)n_rec, (unsigned long long)longest, end_state);
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
    // confirm_input_format (process.c:547), as naf_enc.cu does it on the host
    auto is_space = [](int ch) { return (ch >= 0x09 && ch <= 0x0D) || ch == 0x20; };
    size_t i = 0; u32 last = '\n';
    while (i < text.size() && is_space(text[i])) { last = text[i]; i++; }
    if (i == text.size()) return 3;
    const bool at_ls = last >= 0x0A && last <= 0x0D;
    if (!at_ls || (text[i] != '>' && text[i] != '@')) return 3;
    return text[i] == '@' ? run<true>(text, i + 1, seq_type, no_mask, argv[2]) : run<false>(text, i + 1, seq_type, no_mask, argv[2]);
}
