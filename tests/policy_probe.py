"""Ad-hoc GPU probe (not a test): which way the pilot of the cut range chain decides, and what it costs, per input kind and size."""
import json
import sys
import time

sys.path.insert(0, ".")
from comprox_b200 import api, synth

MiB = 1 << 20
cases = [("text-100M", api.ROLZ, lambda: synth.markov_text(100 * MiB, seed=42), False),
         ("x86-64M", api.ROLZ, lambda: synth.x86_corpus(64 * MiB, seed=43), True),
         ("x86-256M", api.ROLZ, lambda: synth.x86_corpus(256 * MiB, seed=43), True),
         ("bmp-64M-rolz", api.ROLZ, lambda: synth.bmp_corpus(64 * MiB, seed=44), True),
         ("bmp-256M-lzp", api.LZP, lambda: synth.bmp_corpus(256 * MiB, seed=44), True)]
for name, variant, make, filt in cases:
    data = make()
    ref = None
    for serial in (-1, 0, 1):
        with api.Handle(variant) as h:
            h.set_option("rc_serial", serial)
            h.compress(data, 16 * MiB, filt=filt)
            h.profile(True)
            t0 = time.perf_counter(); out = h.compress(data, 16 * MiB, filt=filt); dt = time.perf_counter() - t0
            p = h.profile_report()
        if ref is None:
            ref = out
        print(json.dumps({"input": name, "rc_serial": serial, "identical": out == ref, "wall_ms": round(dt * 1e3, 1), "range_chain_ms": round(p.get("range_chain", 0), 2),
                          "decided_serial": p.get("#rcp_decided_serial"), "pilot_jobs": p.get("#rcp_pilot_jobs"), "est_Gsteps": round(p.get("#rcp_est_steps", 0) / 1e9, 1),
                          "serial_equiv_Gsteps": round(p.get("#rcp_serial_equiv", 0) / 1e9, 1), "state_Gsteps": round(p.get("#rcp_state_steps", 0) / 1e9, 1)}), flush=True)
