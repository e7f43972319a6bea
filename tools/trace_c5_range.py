"""NAFGPU_TRACE=1 python tools/trace_c5_range.py [gbp] [parts]: phase times of decoding 1/parts of the records of a config-5 file"""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, naf_b200
from naf_b200 import api
import bench
gbp = float(sys.argv[1]) if len(sys.argv) > 1 else 3.0
parts = int(sys.argv[2]) if len(sys.argv) > 2 else 8
d, n, nrec = bench.make_c5_device(torch, int(gbp * 1e9))
ctx = naf_b200.NafGpu(0)
cudart = C.CDLL("libcudart.so"); cudart.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
addr, size, info = ctx.encode_device(d.data_ptr(), n, api.make_enc_opts())
dn = torch.zeros(size + 64, dtype=torch.uint8, device="cuda"); cudart.cudaMemcpy(dn.data_ptr(), addr, size, 3); hn = dn[:size].cpu().pin_memory()
for part in (0, parts - 1):
    first = nrec * part // parts; cnt = nrec * (part + 1) // parts - first
    o = api.make_dec_opts(first_record=first, n_records=cnt)
    for rep in range(4):
        ctx.profile(rep == 3)
        torch.cuda.synchronize(); t = time.perf_counter()
        ctx.decode_device(dn.data_ptr(), size, (hn.data_ptr(), size), o)
        torch.cuda.synchronize(); print(f"part {part}/{parts} rep {rep}: {(time.perf_counter() - t) * 1e3:.3f} ms", file=sys.stderr)
    for name, cnt, ms in ctx.profile_report():
        print(f"    {name:22s} x{cnt:3d} {ms:8.3f} ms", file=sys.stderr)
