"""Damaged containers must be refused (CRGPU_ERR_CORRUPT / CRGPU_ERR_ARG) or decoded to some bytes, never crash or leave the decoder's
buffers: a bounded run of tests/fuzz_decode.py on the kernel-logic simulation, in a subprocess (a stray access would kill it)."""
import os
import subprocess
import sys

import pytest

import oracle_ffi as O
from comprox_b200 import api, synth

HERE = os.path.dirname(os.path.abspath(__file__))


def test_sim_decoder_survives_damaged_containers(simlib):
    r = subprocess.run([sys.executable, os.path.join(HERE, "fuzz_decode.py"), "--trials", "120", "--batch", "--seed", "7"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "fuzz_decode: 120 trials" in r.stdout


def test_sim_midchain_stored_block_is_refused_not_crashed(simlib):
    """text | noise | text at -b1 is byte-identical to the reference's container, and -- like the reference's -- cannot be decoded behind
    the stored block.  The reference decoder crashes on it; this one returns CRGPU_ERR_CORRUPT (or garbage-free bytes)."""
    import numpy as np
    MiB = 1 << 20
    rng = np.random.default_rng(5)
    data = synth.markov_text(MiB // 2, seed=85) + rng.integers(0, 256, MiB // 2, dtype=np.uint8).tobytes() + synth.markov_text(MiB // 2, seed=86)
    for variant in (api.ROLZ, api.LZP, api.LZ77):
        cont = O.compress(data, variant, MiB // 2)
        with api.Handle(variant, lib=simlib) as h:
            try:
                out = h.decompress(cont, len(data) * 2)
                assert out[:MiB // 2] == data[:MiB // 2]          # whatever follows the stored block, the first block is right
            except api.CrgpuError as e:
                assert e.code == -9
            good = O.compress(data[:MiB // 2], variant, MiB // 2)
            assert h.decompress(good, MiB) == data[:MiB // 2]     # and the handle is still usable
