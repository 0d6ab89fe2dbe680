"""Driver for an ncu capture of the warp decoder (not a test): compress one 4 MiB text container, decode it twice."""
import sys
sys.path.insert(0, ".")
from comprox_b200 import api, synth
data = synth.markov_text(4 << 20, seed=42)
with api.Handle(api.ROLZ) as h:
    c = h.compress(data, 16 << 20)
    for _ in range(2):
        back = h.decompress(c, len(data) + 64)
    print(len(c), back == data)
