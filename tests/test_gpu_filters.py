"""GPU parity of the -F path (PE/ELF E8E9 rewriting, BMP delta) against the oracle and the reference CLI."""
import pytest

import oracle_ffi as O
from comprox_b200 import api, synth
from test_gpu_compress import _check

pytestmark = pytest.mark.gpu
MiB = 1 << 20


@pytest.mark.parametrize("bs", [MiB, 16 * MiB])
def test_gpu_x86_filter_rolz(gpulib, bs):
    """One ELF image followed by PE images; with 1 MiB blocks the images straddle blocks (continuation state, F3)."""
    _check(gpulib, api.ROLZ, synth.x86_corpus(6 * MiB, elf_bytes=2 * MiB + 12345, pe_min=MiB, pe_max=2 * MiB), bs, flags=["-F"], filt=True)


@pytest.mark.parametrize("bs", [MiB, 16 * MiB])
def test_gpu_bmp_filter_lzp(gpulib, bs):
    """24-bpp BMPs with padded rows; rows broken by block boundaries are skipped exactly like filter_bmp.c:189-199."""
    _check(gpulib, api.LZP, synth.bmp_corpus(9 * MiB, wmin=301, wmax=1200, hmin=100, hmax=700), bs, flags=["-F"], filt=True)


def test_gpu_bmp_filter_rolz(gpulib):
    _check(gpulib, api.ROLZ, synth.bmp_corpus(3 * MiB, wmin=301, wmax=900, hmin=100, hmax=500), MiB, flags=["-F"], filt=True)


def test_gpu_filter_on_text_is_a_noop(gpulib):
    _check(gpulib, api.ROLZ, synth.markov_text(2 * MiB + 77, seed=3), MiB, flags=["-F"], filt=True)


def test_gpu_filter_small_windows(gpulib):
    """Filter continuation state must survive window boundaries."""
    data = synth.x86_corpus(5 * MiB, elf_bytes=MiB + 999, pe_min=MiB, pe_max=2 * MiB)
    with api.Handle(api.ROLZ, lib=gpulib) as h:
        got = h.compress(data, MiB, filt=True, window_bytes=2 * MiB)
    assert got == O.compress(data, api.ROLZ, MiB, filt=1)
