#include "common.cuh"
namespace nafg {
EncodeOut encode_on_device(Ctx &, CudaExec &, const u8 *, size_t, const nafgpu_enc_opts &, nafgpu_enc_info *) { fail(NAFGPU_E_UNSUPPORTED, "encoder not built yet\n"); }
SplitOut split_on_device(Ctx &, CudaExec &, const u8 *, size_t, const nafgpu_enc_opts &, nafgpu_enc_info *) { fail(NAFGPU_E_UNSUPPORTED, "encoder not built yet\n"); }
EncodeOut zstd_compress_on_device(Ctx &, CudaExec &, const u8 *, size_t, int) { fail(NAFGPU_E_UNSUPPORTED, "encoder not built yet\n"); }
}
