"""Short driver for ncu captures (not a test, not a benchmark): N compressions of a text file of --mb MiB."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from comprox_b200 import api, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--mb", type=int, default=32)
ap.add_argument("--runs", type=int, default=2)
ap.add_argument("--variant", type=int, default=0)
ap.add_argument("--kind", default="text", choices=["text", "x86", "bmp"])
ap.add_argument("--filt", type=int, default=0)
ap.add_argument("--opt", action="append", default=[], help="name=value for crgpu_set_option")
a = ap.parse_args()
n = a.mb << 20
data = {"text": lambda: synth.markov_text(n, seed=42), "x86": lambda: synth.x86_corpus(n), "bmp": lambda: synth.bmp_corpus(n)}[a.kind]()
with api.Handle(a.variant) as h:
    for o in a.opt:
        k, v = o.split('=')
        h.set_option(k, int(v))
    for i in range(a.runs):
        h.profile(True)
        out = h.compress(data, 16 << 20, filt=bool(a.filt))
        rep = h.profile_report()
        print("run", i, len(data), "->", len(out), " ".join("%s=%.1f" % kv for kv in rep.items() if not kv[0].startswith("#")), flush=True)
