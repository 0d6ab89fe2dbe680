"""Ad-hoc GPU probe (not a test): one text compression with rc_variant 8 for an ncu launch list."""
import sys
sys.path.insert(0, ".")
from comprox_b200 import api, synth
MiB = 1 << 20
size = int(sys.argv[1]) if len(sys.argv) > 1 else 32
T = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
data = synth.markov_text(size * MiB, seed=42)
with api.Handle(api.ROLZ) as h:
    h.set_option("rc_variant", 8); h.set_option("rc_job_symbols", T)
    out = h.compress(data, 16 * MiB)
    print(len(out))
