"""Block-size sweep (-b1, -b4, -b16, -b64) on one text file: GPU container vs the unmodified reference CLI."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import oracle_ffi as O  # noqa: E402
from comprox_b200 import api, synth  # noqa: E402

MiB = 1 << 20
mb = int(sys.argv[1]) if len(sys.argv) > 1 else 96
data = synth.markov_text(mb * MiB, seed=42)
os.makedirs("gpurun_out", exist_ok=True)
for b in (1, 4, 16, 64):
    with api.Handle(api.ROLZ) as h:
        h.compress(data, b * MiB)
        t0 = time.time(); out = h.compress(data, b * MiB); dt = time.time() - t0
    t0 = time.time(); ref = O.ref_compress(data, "comprolz", ["-b%d" % b], tmpdir="/dev/shm"); dr = time.time() - t0
    rec = {"block_mib": b, "bytes": len(data), "gpu_s": round(dt, 3), "gpu_mibs": round(mb / dt, 1), "ref_s": round(dr, 1), "ref_mibs": round(mb / dr, 2),
           "container": len(out), "identical": out == ref}
    print(json.dumps(rec), flush=True)
    with open("gpurun_out/sweep.jsonl", "a") as f:
        f.write(json.dumps(rec) + "\n")
