"""Ad-hoc GPU probe (not a test): stage timings of text-32M / text-100M with the range-chain variants and job sizes."""
import json
import sys
import time

sys.path.insert(0, ".")
from comprox_b200 import api, synth

MiB = 1 << 20
size = int(sys.argv[1]) if len(sys.argv) > 1 else 32
data = synth.markov_text(size * MiB, seed=42)
ref = None
for rcv, T, cfg in ((7, 0, 0), (4, 0, 0), (8, 0, 0), (8, 0, 1), (8, 16384, 0), (8, 32768, 0), (8, 32768, 1), (8, 49152, 0), (8, 65536, 0), (8, 65536, 1)):
    with api.Handle(api.ROLZ) as h:
        h.set_option("rc_variant", rcv); h.set_option("rc_late_cfg", cfg)
        if T:
            h.set_option("rc_job_symbols", T)
        h.compress(data, 16 * MiB)
        h.profile(True)
        t0 = time.perf_counter(); out = h.compress(data, 16 * MiB); dt = time.perf_counter() - t0
        p = h.profile_report()
    if ref is None:
        ref = out
    print(json.dumps({"rc_variant": rcv, "job_symbols": T, "late_cfg": cfg, "identical": out == ref, "wall_ms": round(dt * 1e3, 1), "range_chain_ms": p.get("range_chain"),
                      "rcp": {k: v for k, v in p.items() if k.startswith("#rcp")}, "triples": p.get("#triples")}), flush=True)
