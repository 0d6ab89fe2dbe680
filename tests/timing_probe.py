import os, sys, time
sys.path.insert(0, '/root/repo')
import torch
from comprox_b200 import api, synth
import ctypes
n = 100 << 20
raw = synth.markov_text(n, seed=42)
host_in = torch.empty(n, dtype=torch.uint8).pin_memory(); host_in.numpy()[:] = memoryview(raw)
L = api.load(); L.crgpu_compress_bound.restype = ctypes.c_uint64
cap = int(L.crgpu_compress_bound(ctypes.c_uint64(n), ctypes.c_uint32(16 << 20)))
host_out = torch.empty(cap, dtype=torch.uint8).pin_memory()
h = api.Handle(api.ROLZ, device=0, stream=torch.cuda.current_stream().cuda_stream)
cfg = api.Config(16 << 20, 0, 0, 0, 0); out_n = ctypes.c_uint64()
for i in range(4):
    if i == 3: os.environ["CRGPU_TIMING"] = "1"
    t0 = time.perf_counter()
    rc = L.crgpu_compress(h.h, ctypes.byref(cfg), ctypes.c_void_p(host_in.data_ptr()), ctypes.c_uint64(n), ctypes.c_void_p(host_out.data_ptr()), ctypes.c_uint64(cap), ctypes.byref(out_n))
    print("call", i, rc, out_n.value, round((time.perf_counter() - t0) * 1e3, 2), "ms", flush=True)
