// emu_zstd — TEST-ONLY: runs naf_b200/csrc/zstd_dec.cuh's decoder through HostExec.
//   emu_zstd IN.zst OUT [one_frame]
#include "../../naf_b200/csrc/zstd_dec.cuh"
#include "host_exec.h"
#include <cstdio>
using namespace nafz;
int main(int argc, char **argv)
{
    if (argc < 3) return 2;
    FILE *f = fopen(argv[1], "rb"); if (!f) return 2;
    std::vector<u8> in; u8 buf[65536]; size_t k;
    while ((k = fread(buf, 1, sizeof buf, f)) > 0) in.insert(in.end(), buf, buf + k);
    fclose(f);
    const size_t in_size = in.size();
    in.resize(in_size + 64);           // the device input buffer carries the same padding (word-wise loaders over-read)
    int one = argc > 3 ? atoi(argv[3]) : 0;
    // expected size unknown here: give a generous arena (the real caller knows it from the container)
    size_t cap = in_size * 300 + (64u << 20);
    if (cap > (1ull << 31)) cap = 1ull << 31;
    std::vector<u8> out(cap);
    HostExec ex;
    ZDecPlan plan;
    plan.streams.push_back(ZStreamDesc{0, in_size, 64, cap - 64, one, 0});
    u32 predef[FSE_SLOT_ENTRIES]; zstd_build_predef(predef);
    std::string err;
    int rc = zstd_decode_batch(ex, in.data(), in.data(), out.data(), plan, predef, err);
    if (rc) { fprintf(stderr, "emu_zstd: %s\n", err.c_str()); return 1; }
    FILE *o = fopen(argv[2], "wb"); fwrite(out.data() + 64, 1, plan.results[0].out_size, o); fclose(o);
    fprintf(stderr, "ok out=%llu nseq=%llu consumed=%llu launches=%u\n", (unsigned long long)plan.results[0].out_size,
            (unsigned long long)plan.results[0].nseq, (unsigned long long)plan.results[0].consumed, ex.launches);
    return 0;
}
