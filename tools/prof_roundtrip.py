"""One encode + decode on the device (for ncu): python tools/prof_roundtrip.py RECORDS [ITERS]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, naf_b200
from naf_b200 import api, synth
n = int(sys.argv[1]); iters = int(sys.argv[2]) if len(sys.argv) > 2 else 1
text = torch.from_numpy(synth.fastq_array(n, 150, seed=1))
d_text = torch.zeros(text.numel() + 64, dtype=torch.uint8, device="cuda"); d_text[:text.numel()] = text.cuda()
ctx = naf_b200.NafGpu(0)
cudart = C.CDLL("libcudart.so"); cudart.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
for _ in range(iters):
    addr, size, info = ctx.encode_device(d_text.data_ptr(), text.numel(), api.make_enc_opts())
    d_naf = torch.zeros(size + 64, dtype=torch.uint8, device="cuda"); cudart.cudaMemcpy(d_naf.data_ptr(), addr, size, 3)
    h_naf = d_naf[:size].cpu()
    taddr, tsize = ctx.decode_device(d_naf.data_ptr(), size, (h_naf.data_ptr(), size), api.make_dec_opts())
    print("naf", size, "text", tsize, "enc+dec launches", ctx.timing().kernel_launches)
