"""Where the time goes when several handles share one GPU (not a test, not a benchmark): `--shards` 64 MiB shards of the mixed corpus
through crgpu_compress_batch with `--handles` handles, profiling on: wall clock, and per-stage CUDA-event time summed over the handles
(a stage that takes longer here than alone is waiting for SMs or for the host)."""
import argparse
import ctypes
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import bench  # noqa: E402
from comprox_b200 import api  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--shards", type=int, default=24)
ap.add_argument("--handles", default="1,8,16")
ap.add_argument("--block-mb", type=int, default=16)
ap.add_argument("--serial", type=int, default=-1)
ap.add_argument("--kinds", default="0,1,2", help="shard indices modulo 3 to keep: 0 text, 1 x86, 2 bmp")
a = ap.parse_args()
MiB = 1 << 20
SB = bench.SHARD_BYTES
L = api.load()
L.crgpu_compress_bound.restype = ctypes.c_uint64
keep = [int(k) for k in a.kinds.split(",")]
idx = [i for i in range(a.shards * 3) if i % 3 in keep][:a.shards]
host_in = torch.empty(len(idx) * SB, dtype=torch.uint8).pin_memory()
for k, i in enumerate(idx):
    host_in.numpy()[k * SB:(k + 1) * SB] = memoryview(bench.corpus_shard(i))
block = a.block_mb * MiB
cap = int(L.crgpu_compress_bound(ctypes.c_uint64(SB), ctypes.c_uint32(block)))
out = torch.empty(len(idx) * cap, dtype=torch.uint8).pin_memory()
m = len(idx)
for K in [int(x) for x in a.handles.split(",")]:
    streams = [torch.cuda.Stream() for _ in range(K)]
    handles = [api.Handle(api.ROLZ, device=0, stream=s.cuda_stream) for s in streams]
    for h in handles:
        h.set_option("rc_serial", a.serial)
    hs = (ctypes.c_void_p * K)(*[h.h for h in handles])
    cfg = api.Config(block, 1, 0, 0, 0)
    ins = (ctypes.c_void_p * m)(*[host_in.data_ptr() + k * SB for k in range(m)])
    in_lens = (ctypes.c_uint64 * m)(*([SB] * m))
    outs = (ctypes.c_void_p * m)(*[out.data_ptr() + k * cap for k in range(m)])
    caps = (ctypes.c_uint64 * m)(*([cap] * m))
    lens = (ctypes.c_uint64 * m)()
    for rep in range(2):                      # first pass: allocations
        for h in handles:
            h.profile(False); h.profile(rep == 1)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        rc = L.crgpu_compress_batch(hs, ctypes.c_uint32(K), ctypes.byref(cfg), ctypes.c_uint32(m), ins, in_lens, outs, caps, lens)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        assert rc == 0, rc
    stages = {}
    for h in handles:
        for k, v in h.profile_report().items():
            if not k.startswith("#"):
                stages[k] = stages.get(k, 0.0) + v
    print(json.dumps({"handles": K, "shards": m, "kinds": keep, "block_mb": a.block_mb, "serial": a.serial, "wall_s": round(dt, 3), "mibs": round(m * SB / MiB / dt, 1),
                      "stage_ms_sum": {k: round(v, 1) for k, v in stages.items()}, "stage_total_ms": round(sum(stages.values()), 1)}), flush=True)
    for h in handles:
        h.close()
