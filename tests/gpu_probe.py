"""Ad-hoc GPU timing probe (not a test): stage timings of crgpu_lzencode on dictionary-coded text."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import cases
from comprox_b200 import api
import oracle_ffi as O

mb = int(sys.argv[1]) if len(sys.argv) > 1 else 32
bs = int(sys.argv[2]) if len(sys.argv) > 2 else 16
variant = int(sys.argv[3]) if len(sys.argv) > 3 else 0
t = time.time(); blocks = cases.dict_coded_text(mb << 20, bs << 20, seed=42, variant=variant); print("prep", time.time() - t, [len(b) for b in blocks], flush=True)
with api.Handle(variant) as h:
    for it in range(3):
        h.reset_models(); h.profile(True)
        t = time.time(); out = h.lzencode(blocks); dt = time.time() - t
        print("iter", it, "lzencode %.3fs -> %.1f MiB/s raw-equivalent" % (dt, mb / dt), [len(o) for o in out][:4], flush=True)
        print('   ', ' '.join('%s=%.1f' % kv for kv in h.profile_report().items()), flush=True)
orc = O.Oracle(variant)
t = time.time(); want = [orc.lzencode(b) for b in blocks]; print("oracle %.3fs" % (time.time() - t), want == out)
