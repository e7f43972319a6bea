// proto_lzw — TEST-SIDE PROTOTYPE (not part of the library): the warp-per-block match finder planned for the next round,
// written as the window-at-a-time algorithm a warp would run (32 positions probed per step: sequential lookup-then-insert
// of the window into the hash table, parallel verification, first acceptable lane taken, cooperative match length), run
// here sequentially.  It feeds the SAME literal / sequence coder as the shipped thread-per-block path
// (nafz::zlz_emit_block), so its frames are checked by the same decoders; it reports the number of warp steps per block,
// which is what the kernel's time will be proportional to.
//   proto_lzw IN OUT.zst BLOCK_SIZE
#include "../../naf_b200/csrc/zstd_enc_hd.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>
using namespace nafz;

static u32 find_warp(const u8 *src, u32 n, u16 *htab, u8 *lit, ZLzSeqs &S, u32 max_seq, u64 *steps, u64 *substeps)
{
    for (u32 e = 0; e < (1u << ZLZ_HLOG); e++) htab[e] = (u16)ZLZ_EMPTY;
    ZLzRep rep; rep.r[0] = rep.r[1] = rep.r[2] = 0; rep.k = 0;
    u32 ip = 0, anchor = 0, nlit = 0; S.n = 0;
    while (ip + 4 <= n && S.n < max_seq) {
        (*steps)++;
        const u32 W = n - 3 - ip < 32 ? n - 3 - ip : 32;
        u32 v[32], cand[32], off[32], hh[32];
        for (u32 l = 0; l < W; l++) {                        // lane 0: lookup-then-insert in position order (what a serial parse sees)
            v[l] = zlz_read32(src + ip + l);
            hh[l] = zlz_hash(v[l]);
            cand[l] = htab[hh[l]]; htab[hh[l]] = (u16)(ip + l);
        }
        for (u32 l = 0; l < W; l++) {                        // all lanes
            const u32 p = ip + l;
            off[l] = 0;
            if (rep.k && p >= rep.r[0] && zlz_read32(src + p - rep.r[0]) == v[l]) off[l] = rep.r[0];
            else if (cand[l] != ZLZ_EMPTY && zlz_read32(src + cand[l]) == v[l]) off[l] = p - cand[l];
        }
        bool taken = false;
        for (u32 l = 0; l < W && !taken; l++) {              // ballot + first set lane; a refused lane costs one more sub-step
            if (!off[l]) continue;
            (*substeps)++;
            u32 p = ip + l, ml = 4; const u32 o = off[l];
            while (p + ml < n && src[p + ml] == src[p + ml - o]) ml++;             // cooperative: 32 bytes per ballot
            while (p > anchor && p > o && src[p - 1] == src[p - 1 - o]) { p--; ml++; }
            if (ml < 5 && !(rep.k && o == rep.r[0] && p > anchor)) continue;
            const u32 ll = p - anchor;
            for (u32 i = 0; i < ll; i++) lit[nlit + i] = src[anchor + i];
            nlit += ll;
            S.ll[S.n] = (u16)ll; S.ml[S.n] = (u16)ml; S.ov[S.n] = (u16)rep.code(o, ll); S.n++;
            // positions of this window that the next step will visit again must not be in the table yet: undo their
            // inserts, last first (each insert remembered what it replaced)
            const u32 w0 = ip;
            ip = p + ml; anchor = ip; taken = true;
            for (u32 k = W; k-- > 0 && w0 + k >= ip;) htab[hh[k]] = (u16)cand[k];
        }
        if (!taken) ip += W;
    }
    for (u32 i = anchor; i < n; i++) lit[nlit++] = src[i];
    return nlit;
}

int main(int argc, char **argv)
{
    if (argc < 4) return 2;
    FILE *f = fopen(argv[1], "rb"); if (!f) return 2;
    std::vector<u8> in; u8 buf[65536]; size_t k;
    while ((k = fread(buf, 1, sizeof buf, f)) > 0) in.insert(in.end(), buf, buf + k);
    fclose(f);
    const u32 bs = (u32)atoi(argv[3]);
    if (bs < 16 || bs > ZLZ_MAX_BLOCK) return 2;
    const u32 max_seq = bs / 4;
    std::vector<u8> out = { 0x28, 0xB5, 0x2F, 0xFD, 0x00, (u8)((17 - 10) << 3) };
    std::vector<u16> htab(1u << ZLZ_HLOG), sll(max_seq), sml(max_seq), sov(max_seq), spos(1280);
    std::vector<u8> lit(bs + 16), tsym(512), codes(3 * max_seq), slot(bs + 512);
    const size_t n = in.size(), nblk = n ? (n + bs - 1) / bs : 1;
    u64 steps = 0, substeps = 0, nseq = 0;
    for (size_t b = 0; b < nblk; b++) {
        const u32 len = (u32)(n - b * bs < bs ? n - b * bs : bs);
        const u8 *src = in.data() + b * bs;
        ZLzSeqs S{sll.data(), sml.data(), sov.data(), 0};
        ZLzWork W{spos.data(), tsym.data(), codes.data()};
        bool rle = false; u32 cs = 0;
        if (len) { u32 i = 1; while (i < len && src[i] == src[0]) i++; rle = i == len; }
        if (!rle && len >= 16) {
            const u32 nlit = find_warp(src, len, htab.data(), lit.data(), S, max_seq, &steps, &substeps);
            nseq += S.n;
            cs = zlz_emit_block(len, lit.data(), nlit, S, max_seq, W, slot.data(), bs + 512);
        }
        const u32 last = b + 1 == nblk, type = cs ? 2 : (rle ? 1 : 0), size_field = type == 2 ? cs : len;
        const u32 bh = last | (type << 1) | (size_field << 3);
        out.push_back((u8)bh); out.push_back((u8)(bh >> 8)); out.push_back((u8)(bh >> 16));
        if (type == 2) out.insert(out.end(), slot.begin(), slot.begin() + cs);
        else if (type == 1) out.push_back(src[0]);
        else out.insert(out.end(), src, src + len);
    }
    FILE *o = fopen(argv[2], "wb"); if (!o) return 2;
    fwrite(out.data(), 1, out.size(), o); fclose(o);
    printf("in=%zu out=%zu blocks=%zu seqs=%llu steps=%llu substeps=%llu steps_per_block=%.1f\n", n, out.size(), nblk,
           (unsigned long long)nseq, (unsigned long long)steps, (unsigned long long)substeps, nblk ? (double)steps / nblk : 0.0);
    return 0;
}
