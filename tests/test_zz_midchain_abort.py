"""Exact replay of a mid-chain "cannot compress" (LzChain::encode_blocks).  When the main stream of a block grows as
large as the block, the reference's coder loop gives up and stores the block raw (src/rolzmain/cr-coder.c:231-233,253-263,
src/ropmain/cr-coder.c:204, src/roxmain/cr-coder.c:273) -- but its models, side models and PPM context keep what the tokens before that
point did to them, and the next blocks are coded on top of that state.  The containers below (incompressible blocks between, before and
after text) must equal the oracle's / the reference CLI's byte for byte.  (The reference's own DECODER cannot read such containers --
it skips stored blocks without touching its models -- so there is no round trip to check; all-incompressible files do round-trip.)
The file sorts last on purpose: it exercises the re-run path, everything before it the single-pass path."""
import numpy as np
import pytest

import oracle_ffi as O
from comprox_b200 import api, synth

BACKENDS = [pytest.param("sim", id="sim"), pytest.param("gpu", id="gpu", marks=pytest.mark.gpu)]
VARIANTS = [pytest.param(api.ROLZ, id="comprolz"), pytest.param(api.LZP, id="comprop"), pytest.param(api.LZ77, id="comprox")]


def _lib(request, which):
    return request.getfixturevalue("simlib" if which == "sim" else "gpulib")


def _noise(n, seed):
    return np.random.default_rng(seed).integers(0, 256, n, dtype=np.uint8).tobytes()


def _cases(scale):
    t = synth.markov_text(300000 * scale, seed=2)
    r = _noise(200000 * scale, 1)
    return {"text_noise_text": t + r + t, "noise_text": r + t, "noise_noise_text": r + r + t[:100000 * scale], "text_noise": t + r,
            "all_noise": r + r}


@pytest.mark.parametrize("which", BACKENDS)
@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("name", ["text_noise_text", "noise_text", "noise_noise_text", "text_noise", "all_noise"])
def test_container_with_incompressible_blocks(request, which, variant, name):
    if which == "sim" and name in ("noise_text", "text_noise", "noise_noise_text"):
        pytest.skip("CPU pre-flight runs two of the shapes (time); the GPU run takes all five")
    lib = _lib(request, which)
    data = _cases(1)[name]
    for bs in ((100000,) if which == "sim" else (65536, 100000)):
        want = O.compress(data, variant, bs)
        with api.Handle(variant, lib=lib) as h:
            assert h.compress(data, bs) == want, (name, bs)


@pytest.mark.parametrize("which", BACKENDS)
def test_window_boundaries_do_not_matter(request, which):
    """The re-run splits a window at the stored block; windows of 1, 3 and all blocks must give the same container."""
    lib = _lib(request, which)
    data = _cases(1)["text_noise_text"]
    want = O.compress(data, api.ROLZ, 65536)
    for wb in ((65536, 0) if which == "sim" else (65536, 3 * 65536, 0)):
        with api.Handle(api.ROLZ, lib=lib) as h:
            assert h.compress(data, 65536, window_bytes=wb) == want


@pytest.mark.parametrize("which", BACKENDS)
@pytest.mark.parametrize("variant", VARIANTS)
def test_lzencode_per_block_leaves_the_models_where_the_reference_does(request, which, variant):
    """Stage level: one lzencode call per block with chain_ends = 0, as host/cr_shim.c issues them."""
    lib = _lib(request, which)
    blocks = [synth.markov_text(120000, seed=5), _noise(90000, 6), b"abc", synth.markov_text(120000, seed=7), _noise(50000, 8), synth.markov_text(60000, seed=9)]
    orc = O.Oracle(variant)
    want = [orc.lzencode(b) for b in blocks]
    with api.Handle(variant, lib=lib) as h:
        got = [h.lzencode([b], chain_ends=False)[0] for b in blocks]
        assert got == want
        h.reset_models()
        assert h.lzencode(blocks, chain_ends=True) == want          # and the same in one batched call


@pytest.mark.gpu
@pytest.mark.parametrize("binary,variant", [("comprolz", api.ROLZ), ("comprop", api.LZP), ("comprox", api.LZ77)])
def test_gpu_matches_reference_cli_on_mixed_file(gpulib, binary, variant):
    t = synth.markov_text(1300000, seed=2)
    data = t + _noise(1500000, 1) + t
    ref = O.ref_compress(data, binary, ["-b1"])
    if ref is None:
        pytest.skip("oracle/_ref/%s not present" % binary)
    with api.Handle(variant, lib=gpulib) as h:
        assert h.compress(data, 1 << 20) == ref
    with api.Handle(variant, lib=gpulib) as h:
        h.set_option("scalar_models", 1)
        assert h.compress(data, 1 << 20) == ref


@pytest.mark.gpu
def test_gpu_all_incompressible_file_round_trips_through_the_reference_decoder(gpulib):
    data = _noise(3 << 20, 11)
    with api.Handle(api.ROLZ, lib=gpulib) as h:
        got = h.compress(data, 1 << 20)
    assert got == O.compress(data, api.ROLZ, 1 << 20)
    back = O.ref_decompress(got, "comprolz")
    if back is not None:
        assert back == data
