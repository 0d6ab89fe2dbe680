"""Decode shard mode: many containers in flight on ONE GPU.  One container is one serial model
chain (one warp), so throughput comes from decoding independent containers side by side: one handle (private stream,
private model + matcher tables) and one host thread per container in flight."""
import argparse
import json
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from comprox_b200 import api, synth  # noqa: E402

MiB = 1 << 20
ap = argparse.ArgumentParser()
ap.add_argument("--shard-mb", type=int, default=8)
ap.add_argument("--per-handle", type=int, default=2)
ap.add_argument("--workers", type=int, nargs="+", default=[1, 8, 32, 64])
ap.add_argument("--variant", type=int, default=0)
ap.add_argument("--batch", type=int, nargs="*", default=[], help="also: crgpu_decompress_batch with this many containers in flight")
a = ap.parse_args()
n = a.shard_mb * MiB
base = synth.markov_text(n, seed=46)
maxw = max(a.workers + [(b + a.per_handle - 1) // a.per_handle for b in a.batch] + [1])
nshard_all = maxw * a.per_handle
shards = [base[k * 4096:] + base[:k * 4096] for k in range(nshard_all)]
with api.Handle(a.variant) as hc:
    conts = [hc.compress(s, 16 * MiB) for s in shards]
os.makedirs("gpurun_out", exist_ok=True)
for nw in a.workers:
    ns = nw * a.per_handle
    handles = [api.Handle(a.variant, stream=api.OWN_STREAM) for _ in range(nw)]
    for h in handles:
        h.decompress(conts[0], n + 64)                                            # warm-up: all device allocations happen here
    out = [None] * ns
    nxt = [0]
    lock = threading.Lock()

    def work(h):
        while True:
            with lock:
                i = nxt[0]; nxt[0] += 1
            if i >= ns:
                return
            out[i] = h.decompress(conts[i], n + 64)

    t0 = time.time()
    th = [threading.Thread(target=work, args=(h,)) for h in handles]
    [t.start() for t in th]; [t.join() for t in th]
    dt = time.time() - t0
    for h in handles:
        h.close()
    rec = {"mode": "decompress", "variant": a.variant, "containers": ns, "container_mib": a.shard_mb, "handles": nw, "seconds": round(dt, 3),
           "mib_per_s": round(ns * a.shard_mb / dt, 1), "roundtrip_ok": all(out[i] == shards[i] for i in range(ns))}
    print(json.dumps(rec), flush=True)
    with open("gpurun_out/decode_shards.jsonl", "a") as f:
        f.write(json.dumps(rec) + "\n")

for nb in a.batch:
    handles = [api.Handle(a.variant) for _ in range(nb)]                          # default stream: one launch decodes all of them
    api.decompress_batch(handles, conts[:nb], [n + 64] * nb)                      # warm-up / allocations
    t0 = time.time()
    out = api.decompress_batch(handles, conts[:nb], [n + 64] * nb)
    dt = time.time() - t0
    for h in handles:
        h.close()
    rec = {"mode": "decompress_batch", "variant": a.variant, "containers": nb, "container_mib": a.shard_mb, "seconds": round(dt, 3),
           "mib_per_s": round(nb * a.shard_mb / dt, 1), "roundtrip_ok": all(out[i] == shards[i] for i in range(nb))}
    print(json.dumps(rec), flush=True)
    with open("gpurun_out/decode_shards.jsonl", "a") as f:
        f.write(json.dumps(rec) + "\n")
