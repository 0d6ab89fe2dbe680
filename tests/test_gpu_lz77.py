"""GPU parity of the LZ77 front-end (`comprox`) through the C ABI: oracle, reference CLI and reference decoder."""
import pytest

import oracle_ffi as O
from cases import lz_cases
from comprox_b200 import api, synth

pytestmark = pytest.mark.gpu
MiB = 1 << 20


@pytest.mark.parametrize("scalar", [0, 1])
@pytest.mark.parametrize("name", sorted(lz_cases(2).keys()))
def test_gpu_lz77_lzencode_matches_oracle(gpulib, name, scalar):
    blocks = lz_cases(2)[name]
    orc = O.Oracle(api.LZ77)
    want = [orc.lzencode(b) for b in blocks]
    with api.Handle(api.LZ77, lib=gpulib) as h:
        h.set_option("scalar_models", scalar)
        assert h.lzencode(blocks) == want


CONTAINERS = {
    "text_8M_b1": (lambda: synth.markov_text(8 * MiB + 4321, seed=42), MiB, [], 0, 0, 0),
    "text_6M_b16": (lambda: synth.markov_text(6 * MiB, seed=43), 16 * MiB, [], 0, 0, 0),
    "text_3M_f": (lambda: synth.markov_text(3 * MiB, seed=44), MiB, ["-f"], 0, 1, 0),
    "text_3M_m8": (lambda: synth.markov_text(3 * MiB, seed=45), MiB, ["-m8"], 0, 0, 8),
    "x86_4M_F": (lambda: synth.x86_corpus(4 * MiB, elf_bytes=MiB + 12345, pe_min=MiB // 2, pe_max=MiB), MiB, ["-F"], 1, 0, 0),
    "bmp_4M_F": (lambda: synth.bmp_corpus(4 * MiB, wmin=301, wmax=900, hmin=100, hmax=500), MiB, ["-F"], 1, 0, 0),
    "empty": (lambda: b"", 16 * MiB, [], 0, 0, 0),
    "zeros_3M": (lambda: bytes(3 * MiB), MiB, [], 0, 0, 0),
}


@pytest.mark.parametrize("name", sorted(CONTAINERS.keys()))
def test_gpu_lz77_container_identical_to_reference(gpulib, name):
    make, bs, flags, filt, flex, ml = CONTAINERS[name]
    data = make()
    want = O.compress(data, api.LZ77, bs, filt, 0, flex, ml)
    with api.Handle(api.LZ77, lib=gpulib) as h:
        if ml:
            h.set_option("match_limit", ml)
        got = h.compress(data, bs, filt=bool(filt), flexible=bool(flex))
    assert got == want
    ref = O.ref_compress(data, "comprox", ["-b%d" % (bs // MiB), *flags])
    if ref is not None:
        assert got == ref
        if not filt:       # the reference decoder cannot undo filters on dictionary-compressible blocks
            assert O.ref_decompress(got, "comprox") == data


def test_gpu_lz77_block_larger_than_16MiB_uses_match_min_11(gpulib):
    data = synth.markov_text(17 * MiB + 99, seed=8)
    want = O.compress(data, api.LZ77, 32 * MiB)
    with api.Handle(api.LZ77, lib=gpulib) as h:
        assert h.compress(data, 32 * MiB) == want


@pytest.mark.parametrize("cap", [0, 1])
def test_gpu_lz77_serial_parse_fallback(gpulib, cap):
    data = synth.markov_text(2 * MiB, seed=3) + (b"abcabcabd" * 50000) + synth.markov_text(MiB, seed=3)
    want = O.compress(data, api.LZ77, MiB)
    with api.Handle(api.LZ77, lib=gpulib) as h:
        h.set_option("lz77_max_iter", cap)
        assert h.compress(data, MiB) == want
