"""Ad-hoc GPU probe (not a test): range-chain time of text-100M (and x86-64M) with the layouts of the tracking kernels (rc_late_cfg)."""
import json
import sys
import time

sys.path.insert(0, ".")
from comprox_b200 import api, synth

MiB = 1 << 20
for name, data in (("text-100M", synth.markov_text(100 * MiB, seed=42)), ("x86-64M", synth.x86_corpus(64 * MiB, seed=43))):
    ref = None
    for cfg in (0, 1, 2, 4):
        with api.Handle(api.ROLZ) as h:
            h.set_option("rc_late_cfg", cfg)
            h.compress(data, 16 * MiB, filt=True)
            h.profile(True)
            t0 = time.perf_counter(); out = h.compress(data, 16 * MiB, filt=True); dt = time.perf_counter() - t0
            p = h.profile_report()
        if ref is None:
            ref = out
        print(json.dumps({"input": name, "late_cfg": cfg, "identical": out == ref, "wall_ms": round(dt * 1e3, 1), "range_chain_ms": p.get("range_chain")}), flush=True)
