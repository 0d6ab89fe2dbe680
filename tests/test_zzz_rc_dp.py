"""The double-precision form of the range recurrence (rc_dp_record / rc_dp_step in cr_rc.cuh, used by k_range_chain<7>, opt-in through
crgpu_set_option("rc_variant", 7)): walked on the host through crgpu_debug_rc_dp with the very functions the kernel calls, against the
integer recurrence of src/cr-rangecoder.c:60-70 (range /= sum; range *= frq; while (range < 2^24) range <<= 8).  No GPU needed; the
kernel itself is compared with the oracle in the gpu test below."""
import ctypes

import numpy as np
import pytest

import oracle_ffi as O
from comprox_b200 import api, synth

MiB = 1 << 20


def _walk_int(frq, sum_):
    r = 0xFFFFFFFF
    q_out, s_out = [], []
    for f, s in zip(frq.tolist(), sum_.tolist()):
        q = r // s
        r = q * f
        k = 0
        while r < (1 << 24):
            r <<= 8
            k += 1
        q_out.append(q); s_out.append(k)
    return np.array(q_out, dtype=np.uint32), np.array(s_out, dtype=np.uint32)


def _run(lib, frq, sum_):
    n = len(frq)
    q = np.zeros(n, dtype=np.uint32); sh = np.zeros(n, dtype=np.uint32)
    rc = lib.crgpu_debug_rc_dp(frq.ctypes.data_as(ctypes.c_void_p), sum_.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint64(n),
                               q.ctypes.data_as(ctypes.c_void_p), sh.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0
    return q, sh


@pytest.mark.parametrize("kind", ["models", "powers_of_two", "tiny_frq", "wide_sums"])
def test_dp_recurrence_matches_integer_recurrence(simlib, kind):
    rng = np.random.default_rng({"models": 1, "powers_of_two": 2, "tiny_frq": 3, "wide_sums": 4}[kind])
    n = 200000
    if kind == "models":                                   # sums as the PPM / order-0 models produce them
        sum_ = rng.integers(2, 65281, n, dtype=np.uint32)
        frq = (rng.random(n) * sum_).astype(np.uint32).clip(1, None)
    elif kind == "powers_of_two":                          # 1/sum exactly representable: the reciprocal must still be strictly above it
        sum_ = (1 << rng.integers(0, 17, n)).astype(np.uint32)
        frq = np.maximum(1, (rng.random(n) * sum_).astype(np.uint32))
    elif kind == "tiny_frq":                               # two- and three-byte renormalisations on almost every symbol
        sum_ = rng.integers(30000, 65281, n, dtype=np.uint32)
        frq = rng.integers(1, 4, n, dtype=np.uint32)
    else:                                                  # up to the 23 bits the triple format allows
        sum_ = rng.integers(1, 1 << 23, n, dtype=np.uint32)
        frq = np.maximum(1, (rng.random(n) ** 3 * sum_).astype(np.uint32))
    frq = np.minimum(frq, sum_).astype(np.uint32)
    want_q, want_s = _walk_int(frq, sum_)
    got_q, got_s = _run(simlib, np.ascontiguousarray(frq), np.ascontiguousarray(sum_))
    assert np.array_equal(got_q, want_q)
    assert np.array_equal(got_s, want_s)


@pytest.mark.gpu
def test_zz_gpu_dp_range_chain_matches_oracle(gpulib):
    """k_range_chain<7> on the GPU (opt-in; the default stays the integer chain until this has passed on a B200)."""
    rng = np.random.default_rng(7)
    skew = rng.choice(np.arange(256, dtype=np.uint8), size=3 * MiB, p=np.r_[0.97, np.full(255, 0.03 / 255)]).tobytes()
    for variant, data in ((api.ROLZ, synth.markov_text(3 * MiB, seed=5)), (api.LZP, skew), (api.LZ77, skew[:MiB] + synth.markov_text(MiB, seed=6))):
        with api.Handle(variant, lib=gpulib) as h:
            h.set_option("rc_variant", 7)
            got = h.compress(data, MiB)
        assert got == O.compress(data, variant, MiB), "rc_variant 7 differs from the oracle"
