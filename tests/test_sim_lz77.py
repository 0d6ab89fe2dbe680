"""CPU pre-flight (kernel-logic simulation) of the LZ77 front-end (`comprox`, src/roxmain) against the oracle:
chain sort + match kernels, the last-match fixed point and its serial fallback, the coder's own last_match, the
eight order-0 models and the 32-byte block header."""
import pytest

import oracle_ffi as O
from cases import lz_cases
from comprox_b200 import api, synth

MiB = 1 << 20


@pytest.mark.parametrize("name", sorted(lz_cases().keys()))
def test_sim_lz77_lzencode_matches_oracle(simlib, name):
    blocks = lz_cases()[name]
    orc = O.Oracle(api.LZ77)
    want = [orc.lzencode(b) for b in blocks]
    with api.Handle(api.LZ77, lib=simlib) as h:
        assert h.lzencode(blocks) == want


def _containers():
    text = synth.markov_text(MiB + 12345, seed=7)
    return {
        "text_b256k": (text, MiB // 4, 0, 0, 0, 0),
        "text_exact_multiple": (text[:MiB], MiB // 2, 0, 0, 0, 0),
        "empty": (b"", MiB, 0, 0, 0, 0),
        "single_byte": (b"A", MiB, 0, 0, 0, 0),
        "text_prec": (text[:MiB // 2 + 5], MiB, 0, 1, 0, 0),
        "text_flexible": (text[:MiB // 2], MiB, 0, 0, 1, 0),
        "text_m8": (text[:MiB // 2], MiB, 0, 0, 0, 8),
        "x86_filtered": (synth.x86_corpus(MiB, elf_bytes=MiB // 4 + 77, pe_min=MiB // 8, pe_max=MiB // 4), MiB // 2, 1, 0, 0, 0),
        "bmp_filtered": (synth.bmp_corpus(MiB, wmin=201, wmax=500, hmin=60, hmax=300), MiB // 2, 1, 0, 0, 0),
    }


@pytest.mark.parametrize("name", sorted(_containers().keys()))
def test_sim_lz77_container_matches_oracle(simlib, name):
    data, bs, filt, prec, flex, ml = _containers()[name]
    want = O.compress(data, api.LZ77, bs, filt, prec, flex, ml)
    with api.Handle(api.LZ77, lib=simlib) as h:
        if ml:
            h.set_option("match_limit", ml)
        got = h.compress(data, bs, filt=bool(filt), prec=bool(prec), flexible=bool(flex))
    assert got == want
    assert O.decompress(got, api.LZ77) == data


@pytest.mark.parametrize("cap", [0, 1, 2])
def test_sim_lz77_iteration_cap_falls_back_to_serial_parse(simlib, cap):
    """The fixed point over m_last_match may be cut short: the serial parse kernel must give the same tokens."""
    data = synth.markov_text(300000, seed=3) + (b"abcabcabd" * 5000) + synth.markov_text(100000, seed=3)
    want = O.compress(data, api.LZ77, MiB)
    with api.Handle(api.LZ77, lib=simlib) as h:
        h.set_option("lz77_max_iter", cap)
        assert h.compress(data, MiB) == want
