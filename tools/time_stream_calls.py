"""Where the fixed cost of the streamed encode goes: wall clock of every call in a fresh process."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
t0 = time.perf_counter()
import naf_b200
from naf_b200 import api, synth
text = synth.fastq(1000, 150, seed=7)
t1 = time.perf_counter(); print(f"import {t1 - t0:.3f}")
ctx = naf_b200.NafGpu(0)
t2 = time.perf_counter(); print(f"create {t2 - t1:.3f}")
opts = api.make_enc_opts()
lib, h = ctx.lib, ctx.h
def tm(name, f):
    t = time.perf_counter(); r = f(); print(f"{name} {time.perf_counter() - t:.3f}"); return r
tm("begin", lambda: lib.nafgpu_encode_begin(h, C.byref(opts), len(text)))
buf, cap = C.c_void_p(), C.c_size_t()
tm("buffer", lambda: lib.nafgpu_encode_buffer(h, C.byref(buf), C.byref(cap)))
C.memmove(buf.value, text, len(text))
tm("feed", lambda: lib.nafgpu_encode_feed(h, len(text)))
out, size, info = C.c_void_p(), C.c_size_t(), api.EncInfo()
tm("end", lambda: lib.nafgpu_encode_end(h, C.byref(out), C.byref(size), C.byref(info)))
tm("encode again (one shot)", lambda: ctx.encode(text))
tm("decode", lambda: ctx.decode(ctx.encode(text)))
