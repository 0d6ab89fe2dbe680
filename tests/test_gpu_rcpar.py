"""GPU tests of the parallel range chain (cr_rcpar.cuh, rc_variant 8): the recurrence cut into jobs must give, symbol for symbol, the
quotients and renormalisation shifts of the serial walk -- on synthetic symbol statistics that hit every path (seeded jobs, merged
jobs, seeds that have to be retried, streams that fall back to the serial walk) and through whole containers against the oracle."""
import ctypes

import numpy as np
import pytest

import oracle_ffi as O
from comprox_b200 import api, synth

pytestmark = pytest.mark.gpu
MiB = 1 << 20


def _serial(L, frq, sum_, lens):
    """the serial recurrence on the host (the library's own step function, checked against the integer form in test_zzz_rc_dp.py)"""
    q = np.empty(len(frq), dtype=np.uint32); sh = np.empty(len(frq), dtype=np.uint32)
    at = 0
    for n in lens:
        if n:
            f = np.ascontiguousarray(frq[at:at + n]); s = np.ascontiguousarray(sum_[at:at + n])
            qo = np.empty(n, dtype=np.uint32); so = np.empty(n, dtype=np.uint32)
            rc = L.crgpu_debug_rc_dp(f.ctypes.data_as(ctypes.c_void_p), s.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint64(n),
                                     qo.ctypes.data_as(ctypes.c_void_p), so.ctypes.data_as(ctypes.c_void_p))
            assert rc == 0
            q[at:at + n] = qo; sh[at:at + n] = so
        at += n
    return q, sh


def _parallel(L, h, frq, sum_, lens, job_symbols):
    n = len(frq)
    q = np.empty(n, dtype=np.uint32); sh = np.empty(n, dtype=np.uint32); stats = np.zeros(8, dtype=np.uint64)
    ln = np.asarray(lens, dtype=np.uint64)
    rc = L.crgpu_debug_rc_parallel(h.h, frq.ctypes.data_as(ctypes.c_void_p), sum_.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint64(n),
                                   ln.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint32(len(lens)), ctypes.c_uint32(job_symbols),
                                   q.ctypes.data_as(ctypes.c_void_p), sh.ctypes.data_as(ctypes.c_void_p), stats.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0, "crgpu_debug_rc_parallel failed: %d" % rc
    return q, sh, dict(zip(("state_steps", "live", "merged", "retries", "demoted", "flagged", "max_e"), (int(x) for x in stats[:7])))


def _symbols(rng, n, kind):
    if kind == "textlike":         # PPM-like: sums spread over 2^6 .. 2^18.5, skewed frequencies
        s = np.exp(rng.uniform(np.log(64), np.log(370000), n)).astype(np.uint32)
    elif kind == "small":          # BMP / LZP-like: no sum above 2^14 -> every job merged, the stream runs serially
        s = rng.integers(300, 15000, n, dtype=np.uint32)
    elif kind == "order0":         # order-0 side models: sums 17000 .. 32000
        s = rng.integers(17000, 32001, n, dtype=np.uint32)
    elif kind == "rare_big":       # small sums with a large one every few thousand symbols
        s = rng.integers(100, 3000, n, dtype=np.uint32)
        s[rng.integers(0, n, max(1, n // 1500))] = rng.integers(60000, 400000, max(1, n // 1500), dtype=np.uint32)
    elif kind == "huge":           # sums up to 2^23 - 1 (three-byte renormalisations)
        s = np.exp(rng.uniform(np.log(2), np.log((1 << 23) - 1), n)).astype(np.uint32)
    else:
        raise ValueError(kind)
    s = np.maximum(s, 2).astype(np.uint32)
    u = rng.random(n)
    f = np.where(u < 0.5, s * rng.uniform(0.3, 0.98, n), s * rng.uniform(0.0005, 0.3, n)).astype(np.uint32)
    f = np.clip(f, 1, s).astype(np.uint32)
    return f, s


@pytest.mark.parametrize("kind,expect_live", [("textlike", True), ("order0", True), ("rare_big", True), ("huge", True), ("small", False)])
def test_gpu_rcpar_matches_serial_walk(gpulib, kind, expect_live):
    rng = np.random.default_rng(hash(kind) & 0xFFFF)
    lens = [700000, 0, 1, 2, 5000, 300001, 65536]
    f, s = _symbols(rng, sum(lens), kind)
    want_q, want_sh = _serial(gpulib, f, s, lens)
    with api.Handle(api.ROLZ, lib=gpulib) as h:
        for T in (4096, 16384, 49152):
            q, sh, st = _parallel(gpulib, h, f, s, lens, T)
            bad = np.flatnonzero((q != want_q) | (sh != want_sh))
            assert len(bad) == 0, "%s T=%d: first difference at symbol %d of %d (%r)" % (kind, T, bad[0], len(q), st)
            assert st["flagged"] == 0, "streams fell back to the serial walk: %r" % st
            assert (st["live"] > 0) == expect_live, st


def test_gpu_rcpar_seeds_that_do_not_shrink_are_merged(gpulib):
    """One large sum followed by symbols that never merge (frq == sum: the range does not move): the seed cannot shrink its state
    list, so the job is merged into the one in front of it -- still exact, and without handing the stream to the serial fallback."""
    n = 200000
    s = np.full(n, 1000, dtype=np.uint32); f = np.full(n, 1000, dtype=np.uint32)
    s[::4096] = 40000; f[::4096] = 39999
    want_q, want_sh = _serial(gpulib, f, s, [n])
    with api.Handle(api.ROLZ, lib=gpulib) as h:
        q, sh, st = _parallel(gpulib, h, f, s, [n], 8192)
    assert np.array_equal(q, want_q) and np.array_equal(sh, want_sh), st
    assert st["demoted"] > 0 and st["flagged"] == 0, st


def test_gpu_rcpar_serial_mode_and_pruned_streams(gpulib):
    """rc_serial = 1 (what crgpu_compress_batch picks when handles share a device): one job per stream, same numbers.  A stream with
    only a couple of usable seeds is pruned to one serial job instead of tracking sets around a long serial span."""
    rng = np.random.default_rng(77)
    lens = [300000, 3, 200000]
    f, s = _symbols(rng, sum(lens), "textlike")
    want_q, want_sh = _serial(gpulib, f, s, lens)
    with api.Handle(api.ROLZ, lib=gpulib) as h:
        h.set_option("rc_serial", 1)
        q, sh, st = _parallel(gpulib, h, f, s, lens, 8192)
        assert np.array_equal(q, want_q) and np.array_equal(sh, want_sh), st
        assert st["live"] == 0 and st["flagged"] == 0, st
        h.set_option("rc_serial", -1)
        q, sh, st = _parallel(gpulib, h, f, s, lens, 8192)
        assert np.array_equal(q, want_q) and np.array_equal(sh, want_sh) and st["live"] > 0, st
    n = 400000
    f, s = _symbols(rng, n, "small")
    s[150000:150040] = 200000; f[150000:150040] = 150000          # the only place a seed can stand
    want_q, want_sh = _serial(gpulib, f, s, [n])
    with api.Handle(api.ROLZ, lib=gpulib) as h:
        q, sh, st = _parallel(gpulib, h, f, s, [n], 8192)
    assert np.array_equal(q, want_q) and np.array_equal(sh, want_sh), st
    assert st["demoted"] >= 1 and st["flagged"] == 0, st


def test_gpu_rcpar_on_real_triples(gpulib):
    """(frq, sum) of a real ROLZ main stream and side stream (oracle trace of 6 MiB of text)."""
    data = synth.markov_text(6 * MiB, seed=31)
    o = O.Oracle(api.ROLZ)
    o.dictionary_load(O.dicpick(data))
    D = o.dictionary_encode(data)
    o.reset_models(); o.trace(True)
    o.lzencode(D)
    t = o.triples()
    main, side = t[t[:, 3] == 0], t[t[:, 3] == 1]
    f = np.ascontiguousarray(np.concatenate([main[:, 1], side[:, 1]]).astype(np.uint32))
    s = np.ascontiguousarray(np.concatenate([main[:, 2], side[:, 2]]).astype(np.uint32))
    lens = [len(main), len(side)]
    want_q, want_sh = _serial(gpulib, f, s, lens)
    with api.Handle(api.ROLZ, lib=gpulib) as h:
        for T in (8192, 49152):
            q, sh, st = _parallel(gpulib, h, f, s, lens, T)
            assert np.array_equal(q, want_q) and np.array_equal(sh, want_sh), st
            assert st["flagged"] == 0 and st["live"] > 0, st


@pytest.mark.parametrize("variant", [api.ROLZ, api.LZP, api.LZ77])
def test_gpu_rcpar_containers_match_oracle(gpulib, variant):
    """Whole containers with rc_variant 8 and small jobs (many job borders per stream), text and skewed binary data."""
    rng = np.random.default_rng(11)
    skew = rng.choice(np.arange(256, dtype=np.uint8), size=2 * MiB, p=np.r_[0.97, np.full(255, 0.03 / 255)]).tobytes()
    for data, bs in ((synth.markov_text(5 * MiB, seed=61), 2 * MiB), (skew + synth.markov_text(MiB, seed=62), MiB)):
        with api.Handle(variant, lib=gpulib) as h:
            h.set_option("rc_variant", 8); h.set_option("rc_job_symbols", 8192)
            got = h.compress(data, bs)
        assert got == O.compress(data, variant, bs), "rc_variant 8 differs from the oracle"
