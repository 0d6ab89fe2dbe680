"""Regression tests (kernel-logic simulation, no GPU) for API-level findings of the round-1 review."""
import ctypes

import numpy as np
import pytest

import oracle_ffi as O
from comprox_b200 import api, synth

MiB = 1 << 20


def test_sim_dicpick_does_not_arm_the_staged_input_path(simlib):
    """crgpu_dicpick stages its input in HBM; a following crgpu_compress with the same pointer and length must still copy the
    (possibly different) bytes.  Only crgpu_stage_input arms the 'already resident' path, and only for one call."""
    a = synth.markov_text(300000, seed=81)
    b = synth.markov_text(300000, seed=82)
    buf = ctypes.create_string_buffer(a, len(a))
    with api.Handle(api.ROLZ, lib=simlib) as h:
        out = ctypes.create_string_buffer(1 << 20); n = ctypes.c_uint64()
        assert simlib.crgpu_dicpick(h.h, buf, ctypes.c_uint64(len(a)), out, ctypes.c_uint64(1 << 20), ctypes.byref(n)) == 0
        ctypes.memmove(buf, b, len(b))                                   # the caller reuses its read buffer
        simlib.crgpu_compress_bound.restype = ctypes.c_uint64
        cap = simlib.crgpu_compress_bound(ctypes.c_uint64(len(b)), ctypes.c_uint32(MiB))
        cout = ctypes.create_string_buffer(cap); cn = ctypes.c_uint64()
        cfg = api.Config(MiB, 0, 0, 0, 0)
        assert simlib.crgpu_compress(h.h, ctypes.byref(cfg), buf, ctypes.c_uint64(len(b)), cout, ctypes.c_uint64(cap), ctypes.byref(cn)) == 0
        assert cout.raw[:cn.value] == O.compress(b, api.ROLZ, MiB)
        # stage_input is one-shot: the second compress copies again
        assert simlib.crgpu_stage_input(h.h, buf, ctypes.c_uint64(len(b))) == 0
        assert simlib.crgpu_compress(h.h, ctypes.byref(cfg), buf, ctypes.c_uint64(len(b)), cout, ctypes.c_uint64(cap), ctypes.byref(cn)) == 0
        ctypes.memmove(buf, a, len(a))
        assert simlib.crgpu_compress(h.h, ctypes.byref(cfg), buf, ctypes.c_uint64(len(a)), cout, ctypes.c_uint64(cap), ctypes.byref(cn)) == 0
        assert cout.raw[:cn.value] == O.compress(a, api.ROLZ, MiB)


def test_sim_resume_point_is_dropped_by_stage_calls(simlib):
    """A too-small crgpu_decompress leaves a resume point; a stage call on the same handle in between reuses the buffers, so the
    second crgpu_decompress must decode from scratch instead of returning the stage call's bytes."""
    data = synth.markov_text(200000, seed=83)
    cont = O.compress(data, api.ROLZ, MiB)
    o = O.Oracle(api.ROLZ)
    dic = O.dicpick(data)
    o.dictionary_load(dic)
    coded = o.dictionary_encode(data[:3000])
    with api.Handle(api.ROLZ, lib=simlib) as h:
        out = ctypes.create_string_buffer(16); n = ctypes.c_uint64(0)
        assert simlib.crgpu_decompress(h.h, cont, ctypes.c_uint64(len(cont)), out, ctypes.c_uint64(16), ctypes.byref(n)) == -3
        assert n.value == len(data)
        h.dictionary_load(dic, 0)
        assert h.dictionary_decode(coded, 4096) == data[:3000]
        assert h.decompress(cont, len(data)) == data


def test_sim_lcp_encode_rejects_text_without_line_structure(simlib):
    out = ctypes.create_string_buffer(64); n = ctypes.c_uint64()
    for bad in (b"abc\0", b"\0", b"ab\ncd\0", b"a\0b\n\0"):
        assert simlib.crgpu_dic_lcp_encode(bad, ctypes.c_uint64(len(bad)), out, ctypes.c_uint64(64), ctypes.byref(n)) == -3
    good = b"  \nhttp://www.\nabc\nabd\n\0"
    assert simlib.crgpu_dic_lcp_encode(good, ctypes.c_uint64(len(good)), out, ctypes.c_uint64(64), ctypes.byref(n)) == 0
    assert out.raw[:n.value] == O.lcp_encode(good[:-1])


def test_sim_lzencode_checks_capacity_before_touching_the_models(simlib):
    d1 = synth.markov_text(50000, seed=84)
    with api.Handle(api.ROLZ, lib=simlib) as h, api.Handle(api.ROLZ, lib=simlib) as g:
        sizes = (ctypes.c_uint32 * 1)(len(d1)); osz = (ctypes.c_uint32 * 1)()
        small = ctypes.create_string_buffer(100)
        assert simlib.crgpu_lzencode(h.h, d1, sizes, 1, 0, small, ctypes.c_uint64(100), osz) == -3
        assert h.lzencode([d1], chain_ends=False) == g.lzencode([d1], chain_ends=False)      # the failed call left no trace in the models


def test_sim_cut_blocks_are_reported(simlib):
    """text | noise | text at -b1: byte-identical to the reference, but undecodable behind the stored block; the
    library says so through crgpu_get_stat."""
    rng = np.random.default_rng(5)
    data = synth.markov_text(MiB, seed=85) + rng.integers(0, 256, MiB, dtype=np.uint8).tobytes() + synth.markov_text(MiB, seed=86)
    with api.Handle(api.ROLZ, lib=simlib) as h:
        got = h.compress(data, MiB)
        assert got == O.compress(data, api.ROLZ, MiB)
        assert h.get_stat("last_cut_blocks") == 1 and h.get_stat("cut_blocks") == 1
        h.compress(data[:MiB], MiB)
        assert h.get_stat("last_cut_blocks") == 0 and h.get_stat("cut_blocks") == 1
        with pytest.raises(api.CrgpuError):
            h.get_stat("nope")
