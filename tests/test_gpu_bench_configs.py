"""GPU parity at the sizes bench.py and BASELINE.json's configs run at: several FULL 16 MiB blocks in one container (4-byte-context
hash of blocks >= 4 MiB, multi-block windows, multi-million-symbol range chains, models carried across blocks), each against the
unmodified reference CLI's bytes, the reference DECODER's round trip and the CPU oracle.  48 MiB = 3 blocks + the trailing empty
block the reference writes when the size is a multiple of the block size."""
import hashlib

import pytest

import oracle_ffi as O
from comprox_b200 import api, synth

pytestmark = pytest.mark.gpu
MiB = 1 << 20


def _check_vs_reference(gpulib, variant, binary, data, flags, **kw):
    with api.Handle(variant, lib=gpulib) as h:
        got = h.compress(data, 16 * MiB, **kw)
        assert h.get_stat("last_cut_blocks") == 0
    ref = O.ref_compress(data, binary, ["-b16", *flags])
    assert ref is not None, "oracle/_ref did not travel to this box"
    assert hashlib.sha256(got).hexdigest() == hashlib.sha256(ref).hexdigest(), "container differs from the reference CLI (-b16)"
    assert O.ref_decompress(got, binary) == data, "the reference decompressor does not round-trip our container"
    want = O.compress(data, variant, 16 * MiB, int(kw.get("filt", False)))
    assert got == want, "container differs from the oracle"
    return got


def test_gpu_text_48m_comprolz_b16(gpulib):
    """configs[1] shape: word-Markov text, comprolz default -b16, three full blocks."""
    _check_vs_reference(gpulib, api.ROLZ, "comprolz", synth.markov_text(48 * MiB, seed=42), [])


def test_gpu_x86_48m_comprolz_b16_filtered(gpulib):
    """configs[2] shape: one ELF image followed by PE images, comprolz -b16 -F (E8/E9 rewrite across block borders)."""
    data = synth.x86_corpus(48 * MiB, seed=43, elf_bytes=20 * MiB, pe_min=4 * MiB, pe_max=12 * MiB)
    _check_vs_reference(gpulib, api.ROLZ, "comprolz", data, ["-F"], filt=True)


def test_gpu_bmp_48m_comprop_b16_filtered(gpulib):
    """configs[3] shape: 24-bit BMP images, comprop -b16 -F (images straddle the 16 MiB borders)."""
    data = synth.bmp_corpus(48 * MiB, seed=44)
    _check_vs_reference(gpulib, api.LZP, "comprop", data, ["-F"], filt=True)


def test_gpu_text_40m_comprolz_b16_ragged_tail_matches_its_decoder(gpulib):
    """2.5 blocks (no trailing empty block) and the GPU decoder on the result."""
    data = synth.markov_text(40 * MiB + 12345, seed=52)
    got = _check_vs_reference(gpulib, api.ROLZ, "comprolz", data, [])
    with api.Handle(api.ROLZ, lib=gpulib) as h:
        assert h.decompress(got, len(data) + 64) == data
