"""Per-kernel times of a level-2 encode of 10 M reads (NAFGPU_ZLB / NAFGPU_LIB select block size and build): where k_zenc_lz's time goes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, naf_b200
from naf_b200 import api, synth
n = 10_000_000
text = torch.from_numpy(synth.fastq_array(n, 150, seed=42))
d = torch.zeros(text.numel() + 64, dtype=torch.uint8, device="cuda"); d[:text.numel()] = text.cuda()
ctx = naf_b200.NafGpu(0)
for rep in range(3):
    ctx.profile(rep == 2)
    ctx.encode_device(d.data_ptr(), text.numel(), api.make_enc_opts(level=2))
for name, cnt, ms in sorted(ctx.profile_report(), key=lambda x: -x[2])[:5]:
    print(os.environ.get("NAFGPU_LIB", "default")[-20:], os.environ.get("NAFGPU_ZLB", "8192"), name, round(ms, 3))
