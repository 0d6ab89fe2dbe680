"""Shard mode: independent 64 MiB containers compressed side by side on ONE GPU by several handles
(one host thread + one private stream each).  The serial range chains of different containers overlap."""
import argparse
import json
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from comprox_b200 import api, synth  # noqa: E402

MiB = 1 << 20
ap = argparse.ArgumentParser()
ap.add_argument("--shards", type=int, default=16)
ap.add_argument("--shard-mb", type=int, default=64)
ap.add_argument("--workers", type=int, nargs="+", default=[1, 2, 4, 8])
ap.add_argument("--check", action="store_true", help="compare every container with the oracle (slow)")
ap.add_argument("--batch", action="store_true", help="one crgpu_compress_batch call (host threads inside the C ABI) instead of Python threads")
a = ap.parse_args()
base = synth.markov_text(a.shard_mb * MiB, seed=45)
shards = [base[k * 4096:] + base[:k * 4096] for k in range(a.shards)]          # distinct rotations: cheap to make, same statistics
os.makedirs("gpurun_out", exist_ok=True)
ref = None
for nw in a.workers:
    handles = [api.Handle(api.ROLZ, stream=api.OWN_STREAM) for _ in range(nw)]
    for h in handles:
        h.compress(shards[0], 16 * MiB)                                          # warm-up / allocations
    out = [None] * a.shards
    nxt = [0]
    lock = threading.Lock()

    def work(h):
        while True:
            with lock:
                i = nxt[0]; nxt[0] += 1
            if i >= a.shards:
                return
            out[i] = h.compress(shards[i], 16 * MiB)

    t0 = time.time()
    if a.batch:
        out = api.compress_batch(handles, shards, 16 * MiB)
        dt = handles[0].L and time.time() - t0
    else:
        th = [threading.Thread(target=work, args=(h,)) for h in handles]
        [t.start() for t in th]; [t.join() for t in th]
        dt = time.time() - t0
    for h in handles:
        h.close()
    if ref is None:
        ref = out
    rec = {"shards": a.shards, "shard_mib": a.shard_mb, "handles": nw, "seconds": round(dt, 3), "mib_per_s": round(a.shards * a.shard_mb / dt, 1),
           "same_as_single_handle": out == ref, "mode": "crgpu_compress_batch" if a.batch else "python threads"}
    if a.check:
        import oracle_ffi as O
        rec["oracle_identical"] = all(out[i] == O.compress(shards[i], 0, 16 * MiB) for i in range(min(2, a.shards)))
    print(json.dumps(rec), flush=True)
    with open("gpurun_out/shards.jsonl", "a") as f:
        f.write(json.dumps(rec) + "\n")
