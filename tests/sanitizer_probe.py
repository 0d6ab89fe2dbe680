"""Driver for compute-sanitizer (not a test): one small container of every kind and front-end, compared with the oracle."""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import oracle_ffi as O
from comprox_b200 import api, synth
MiB = 1 << 20
cases = [(api.ROLZ, synth.markov_text(6 * MiB, seed=5), 2 * MiB, False), (api.ROLZ, synth.x86_corpus(3 * MiB, seed=6), MiB, True),
         (api.LZP, synth.bmp_corpus(3 * MiB, seed=7), MiB, True), (api.LZ77, synth.markov_text(3 * MiB, seed=8), MiB, False)]
for variant, data, bs, filt in cases:
    with api.Handle(variant) as h:
        got = h.compress(data, bs, filt=filt)
        back = h.decompress(got, len(data) + 64)
    print(variant, len(data), len(got), got == O.compress(data, variant, bs, int(filt)), back == data, flush=True)
