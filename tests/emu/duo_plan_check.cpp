// duo_plan_check — TEST-ONLY: invariants of naf_b200/csrc/duo_plan.hpp on random block lists.
//   every byte of the file is uploaded exactly once, in ranges that follow each other's ends where they should;
//   pieces partition both block lists in order; every block of a piece lies inside the piece's byte range;
//   after pair p the sequence blocks so far hold at least the bases of the quality blocks so far (or all of them);
//   no pair but the last is tiny; frames out of file order are refused.
#include "../../naf_b200/csrc/duo_plan.hpp"
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <random>
using namespace nafg;

static int fail(const char *what, int it) { fprintf(stderr, "iteration %d: %s\n", it, what); return 1; }

int main()
{
    std::mt19937_64 rng(12345);
    for (int it = 0; it < 3000; it++) {
        const bool packed = rng() & 1;
        const size_t nq = 1 + rng() % 3000, nsb = 1 + rng() % 2000;
        const uint32_t zbs = 1u << (10 + rng() % 6);
        std::vector<DuoBlock> seq, qual;
        uint64_t at = 200 + rng() % 100000;                                   // header + small streams
        at += 2;                                                              // frame header of the sequence frame
        uint64_t bases = 0;
        for (size_t i = 0; i < nsb; i++) {
            const uint32_t regen = i + 1 == nsb ? 1 + (uint32_t)(rng() % zbs) : zbs, cs = 1 + (uint32_t)(rng() % (regen + 16));
            seq.push_back(DuoBlock{at + 3, cs, regen}); at += 3 + cs; bases += packed ? 2ull * regen : regen;
        }
        at += 5 + rng() % 12;                                                 // quality section header + frame header
        uint64_t quals = 0;
        for (size_t i = 0; i < nq; i++) {
            const uint32_t regen = i + 1 == nq ? 1 + (uint32_t)(rng() % zbs) : zbs, cs = 1 + (uint32_t)(rng() % (regen + 16));
            qual.push_back(DuoBlock{at + 3, cs, regen}); at += 3 + cs; quals += regen;
        }
        const uint64_t n = at + rng() % 3;
        const uint64_t piece = 1 + rng() % (4u << 20);
        std::vector<DuoPiece> pieces; std::vector<std::pair<uint64_t, uint64_t>> order;
        if (!duo_plan(seq, qual, packed, piece, n, pieces, order)) return fail("a well-formed file was refused", it);
        // 1. the upload covers [0, n) exactly once
        std::vector<std::pair<uint64_t, uint64_t>> sorted;
        for (auto &r : order) { if (r.first > r.second) return fail("range runs backwards", it); if (r.first < r.second) sorted.push_back(r); }
        std::sort(sorted.begin(), sorted.end());
        uint64_t pos = 0;
        for (auto &r : sorted) { if (r.first != pos) return fail("gap or overlap in the upload", it); pos = r.second; }
        if (pos != n) return fail("upload does not end at the end of the file", it);
        if (order.size() != 1 + 2 * pieces.size()) return fail("one range for the front, two per pair", it);
        // 2. pieces partition the block lists, blocks lie inside their piece's range, bases cover qualities
        size_t s = 0, q = 0; uint64_t sreg = 0, qreg = 0;
        for (size_t p = 0; p < pieces.size(); p++) {
            const DuoPiece &pc = pieces[p];
            if (pc.s0 != s || pc.q0 != q || pc.s1 < pc.s0 || pc.q1 <= pc.q0) return fail("pieces are not a partition in order", it);
            const auto &rs = order[1 + 2 * p], &rq = order[2 + 2 * p];
            for (size_t i = pc.s0; i < pc.s1; i++) { if (seq[i].src - 3 < rs.first || seq[i].src + seq[i].csize > rs.second) return fail("sequence block outside its piece", it); sreg += seq[i].regen; }
            for (size_t i = pc.q0; i < pc.q1; i++) { if (qual[i].src - 3 < rq.first || qual[i].src + qual[i].csize > rq.second) return fail("quality block outside its piece", it); qreg += qual[i].regen; }
            s = pc.s1; q = pc.q1;
            if (s < seq.size() && (packed ? 2 * sreg : sreg) < qreg) return fail("qualities without their bases", it);
            if (p + 1 < pieces.size() && qual.size() - pc.q1 < 64) return fail("a tiny last pair was left", it);
        }
        if (s != seq.size() || q != qual.size()) return fail("blocks left over", it);
        (void)bases; (void)quals;
        // 3. frames out of order are refused
        if (it % 50 == 0) {
            std::vector<DuoBlock> bad = qual; std::swap(bad[0].src, bad.back().src);
            if (bad.size() > 1 && duo_plan(seq, bad, packed, piece, n, pieces, order)) return fail("blocks out of order were accepted", it);
            if (duo_plan(qual, seq, packed, piece, n, pieces, order)) return fail("quality before sequence was accepted", it);
        }
    }
    puts("ok");
    return 0;
}
