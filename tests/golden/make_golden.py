#!/usr/bin/env python
"""Regenerates tests/golden/kat.json by running the UNMODIFIED reference build (oracle/_ref, produced by
`make -C oracle ref` from /root/reference) on deterministic inputs.  Run in the build container only."""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle_ffi as O  # noqa: E402
from golden_inputs import INPUTS  # noqa: E402

out = {}
for name, (make, runs) in INPUTS.items():
    data = make()
    for binary, flags in runs:
        c = O.ref_compress(data, binary, flags)
        assert c is not None, "reference binaries missing: make -C oracle ref"
        assert O.ref_decompress(c, binary) == data, (name, binary, flags)
        out["%s|%s|%s" % (name, binary, " ".join(flags))] = {"input_sha256": hashlib.sha256(data).hexdigest(), "input_bytes": len(data),
                                                            "container_bytes": len(c), "container_sha256": hashlib.sha256(c).hexdigest()}
json.dump(out, open(os.path.join(HERE, "kat.json"), "w"), indent=1, sort_keys=True)
print("wrote", len(out), "vectors")
