"""Ad-hoc timing of crgpu_decompress (not a test)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from comprox_b200 import api, synth
mb = int(sys.argv[1]) if len(sys.argv) > 1 else 8
variant = int(sys.argv[2]) if len(sys.argv) > 2 else 0
data = synth.markov_text(mb << 20, seed=42)
with api.Handle(variant) as h:
    c = h.compress(data, 16 << 20)
    for i in range(2):
        t = time.time(); out = h.decompress(c, len(data) + 64); dt = time.time() - t
        print("decompress %d MiB variant %d: %.2fs = %.2f MiB/s ok=%s" % (mb, variant, dt, mb / dt, out == data), flush=True)
