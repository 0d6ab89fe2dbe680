"""Stage times and counters per input kind (not a test, not a benchmark): one warm compression of --mb MiB per kind and block size."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from comprox_b200 import api, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--mb", type=int, default=64)
ap.add_argument("--kinds", default="text,x86,bmp")
ap.add_argument("--blocks", default="16")
ap.add_argument("--variant", type=int, default=0)
ap.add_argument("--opt", action="append", default=[])
a = ap.parse_args()
n = a.mb << 20
mk = {"text": lambda: synth.markov_text(n, seed=42), "x86": lambda: synth.x86_corpus(n, seed=43), "bmp": lambda: synth.bmp_corpus(n, seed=44)}
with api.Handle(a.variant) as h:
    for o in a.opt:
        k, v = o.split('=')
        h.set_option(k, int(v))
    for kind in a.kinds.split(','):
        data = mk[kind]()
        for b in [int(x) for x in a.blocks.split(',')]:
            h.compress(data, b << 20, filt=True)
            h.profile(True)
            t0 = time.time(); out = h.compress(data, b << 20, filt=True); dt = time.time() - t0
            rep = h.profile_report()
            print(json.dumps({"kind": kind, "mb": a.mb, "b": b, "s": round(dt, 4), "mibs": round(a.mb / dt, 1), "out": len(out),
                              "stages": {k: round(v, 2) for k, v in rep.items() if not k.startswith('#')},
                              "counters": {k: v for k, v in rep.items() if k.startswith('#')}}), flush=True)
