"""GPU probe (not a test): range-chain time of every k_range_chain formulation on the same inputs, outputs compared."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from comprox_b200 import api, synth  # noqa: E402

MiB = 1 << 20
cases = [("text32", api.ROLZ, synth.markov_text(32 * MiB, seed=42), False), ("bmp32", api.LZP, synth.bmp_corpus(32 * MiB), True),
         ("text16-lz77", api.LZ77, synth.markov_text(16 * MiB, seed=42), False)]
for name, variant, data, filt in cases:
    base = None
    for rcv in (4, 5, 6):
        with api.Handle(variant) as h:
            h.set_option("rc_variant", rcv)
            best = None
            for i in range(3):
                h.profile(True)
                out = h.compress(data, 16 * MiB, filt=filt)
                rep = h.profile_report()
                t = rep.get("range_chain", 0.0)
                best = t if best is None else min(best, t)
            if base is None:
                base = out
            print("%s rc_variant=%d range_chain=%.2f ms triples=%d identical=%s" % (name, rcv, best, rep.get("#triples", 0), out == base), flush=True)
