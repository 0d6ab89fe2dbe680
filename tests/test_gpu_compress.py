"""GPU parity of crgpu_compress (whole container through the C ABI) against the CPU oracle and, where the
reference binaries travelled with the repo (oracle/_ref), against the unmodified reference CLI itself,
including a round trip through the REFERENCE decompressor."""
import hashlib

import pytest

import oracle_ffi as O
from comprox_b200 import api, synth

pytestmark = pytest.mark.gpu
MiB = 1 << 20
BIN = {api.ROLZ: "comprolz", api.LZP: "comprop"}


def _check(gpulib, variant, data, bs, flags=(), **kw):
    with api.Handle(variant, lib=gpulib) as h:
        got = h.compress(data, bs, **kw)
    want = O.compress(data, variant, bs, int(kw.get("filt", False)), int(kw.get("prec", False)), int(kw.get("flexible", False)))
    assert len(got) == len(want), "container size differs from oracle"
    assert got == want, "container differs from oracle"
    ref = O.ref_compress(data, BIN[variant], ["-b%d" % (bs // MiB), *flags]) if bs % MiB == 0 else None
    if ref is not None:
        assert got == ref, "container differs from the reference CLI"
        assert O.ref_decompress(got, BIN[variant]) == data, "reference decompressor does not round-trip our container"
    return got


@pytest.mark.parametrize("variant", [api.ROLZ, api.LZP])
def test_gpu_compress_text_8m_b1(gpulib, variant):
    _check(gpulib, variant, synth.markov_text(8 * MiB + 4321, seed=42), MiB)


@pytest.mark.parametrize("variant", [api.ROLZ, api.LZP])
def test_gpu_compress_text_8m_b16(gpulib, variant):
    _check(gpulib, variant, synth.markov_text(8 * MiB, seed=43), 16 * MiB)


@pytest.mark.parametrize("variant", [api.ROLZ, api.LZP])
def test_gpu_compress_trailing_empty_block(gpulib, variant):
    _check(gpulib, variant, synth.markov_text(2 * MiB, seed=44), MiB)


@pytest.mark.parametrize("variant", [api.ROLZ, api.LZP])
@pytest.mark.parametrize("data", [b"", b"A", bytes(MiB), (b"The quick brown fox jumps over the lazy dog. " * 30000)[:1300000]],
                         ids=["empty", "A", "zeros", "fox"])
def test_gpu_compress_known_answers(gpulib, variant, data):
    _check(gpulib, variant, data, 16 * MiB)


def test_gpu_compress_survey_digests(gpulib):
    """SURVEY.md App. E known answers of the reference build (sha256[:16] of the comprolz container)."""
    kat = {b"": "87c0b28fce506b04", b"A": "a7b3209b72de7a3c", bytes(MiB): "8a20dfe0637804bc",
           (b"The quick brown fox jumps over the lazy dog. " * 30000)[:1300000]: "67dc0831e1ad8791"}
    with api.Handle(api.ROLZ, lib=gpulib) as h:
        for data, digest in kat.items():
            assert hashlib.sha256(h.compress(data)).hexdigest()[:16] == digest


@pytest.mark.parametrize("variant", [api.ROLZ, api.LZP])
def test_gpu_compress_prec(gpulib, variant):
    _check(gpulib, variant, synth.markov_text(2 * MiB + 5, seed=45), MiB, flags=["-p"], prec=True)


def test_gpu_compress_windows_carry_models(gpulib):
    """Small windows: model state and PPM context must carry across windows exactly as across blocks."""
    data = synth.markov_text(6 * MiB, seed=46)
    with api.Handle(api.ROLZ, lib=gpulib) as h:
        a = h.compress(data, MiB)
    with api.Handle(api.ROLZ, lib=gpulib) as h:
        b = h.compress(data, MiB, window_bytes=2 * MiB)
    assert a == b == O.compress(data, api.ROLZ, MiB)


def test_gpu_compress_binary_no_filter(gpulib):
    _check(gpulib, api.ROLZ, synth.x86_corpus(4 * MiB, elf_bytes=MiB, pe_min=MiB // 2, pe_max=MiB), MiB)


def test_gpu_compress_flexible_parsing(gpulib):
    data = synth.markov_text(2 * MiB, seed=47) * 2 + synth.x86_corpus(2 * MiB, elf_bytes=0, pe_min=MiB, pe_max=2 * MiB)
    _check(gpulib, api.ROLZ, data, MiB, flags=["-f"], flexible=True)


def test_gpu_dicpick_vocabulary_overflow(gpulib):
    from vocab_overflow_input import overflow_text
    data = overflow_text()
    with api.Handle(api.LZP, lib=gpulib) as h:
        assert h.dicpick(data) == O.dicpick(data)
        assert h.compress(data, 16 * MiB) == O.compress(data, api.LZP, 16 * MiB)


@pytest.mark.parametrize("rcv", [1, 2, 3, 4, 5, 6])
def test_gpu_range_chain_variants_identical(gpulib, rcv):
    """Every formulation of the serial range chain (k_range_chain<1..6>, cr_warp.cuh) yields the oracle's bytes: skewed
    inputs make 2- and 3-byte renormalisations (the rare path of variants 5/6) frequent."""
    import numpy as np
    rng = np.random.default_rng(7)
    skew = rng.choice(np.arange(256, dtype=np.uint8), size=3 * MiB, p=np.r_[0.97, np.full(255, 0.03 / 255)]).tobytes()
    for variant, data in ((api.ROLZ, synth.markov_text(3 * MiB, seed=5)), (api.LZP, skew), (api.ROLZ, skew[:MiB] + synth.markov_text(MiB, seed=6))):
        with api.Handle(variant, lib=gpulib) as h:
            h.set_option("rc_variant", rcv)
            got = h.compress(data, MiB)
        assert got == O.compress(data, variant, MiB), "rc_variant %d differs from the oracle" % rcv


@pytest.mark.parametrize("variant", [api.ROLZ, api.LZP])
def test_gpu_model_pass_variants_agree(gpulib, variant):
    """The hot-context passes in all their forms (o2: split chain / single kernel / first form, 256 / 512 / 1024 events per step; o1:
    split / single kernel), the split o2 pass when its table of step records is far too small (contexts are then redone by the single
    kernel), the range chain forced serial / forced cut, the dictionary payload on the main chain, the first form of the match search:
    same container, on text (hit-heavy) and on skewed binary data."""
    import numpy as np
    rng = np.random.default_rng(5)
    skew = rng.choice(np.arange(256, dtype=np.uint8), size=6 * MiB, p=np.r_[0.5, 0.2, np.full(254, 0.3 / 254)]).tobytes()
    for data in (synth.markov_text(6 * MiB, seed=91), skew):
        want = O.compress(data, variant, 4 * MiB)
        for opts in ({}, {"o2_hot_variant": 2}, {"o2_hot_variant": 1, "o1_hot_variant": 1}, {"o2_rec_cap_test": 37}, {"o2_rec_cap_test": 1},
                     {"o2_width": 256}, {"o2_width": 512}, {"o2_width": 1024}, {"rc_serial": 1}, {"rc_serial": 0}, {"dict_mode": 0},
                     {"rolz_match_variant": 1}, {"dp_tiles": 1}, {"dc_listed": 0}):
            with api.Handle(variant, lib=gpulib) as h:
                for k, v in opts.items():
                    h.set_option(k, v)
                assert h.compress(data, 4 * MiB) == want, "options %r: container differs from the oracle" % (opts,)
