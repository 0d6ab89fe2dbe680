"""GPU parity of crgpu_decompress: containers written by the unmodified reference CLI (oracle/_ref, when present),
by the oracle, and by our own crgpu_compress must decode to the original bytes."""
import pytest

import oracle_ffi as O
from comprox_b200 import api, synth

pytestmark = pytest.mark.gpu
MiB = 1 << 20
BIN = {api.ROLZ: "comprolz", api.LZP: "comprop", api.LZ77: "comprox"}


def _sources(variant, data, bs, flags, filt):
    yield "oracle", O.compress(data, variant, bs, filt, 0)
    ref = O.ref_compress(data, BIN[variant], ["-b%d" % (bs // MiB), *flags])
    if ref is not None:
        yield "reference", ref


@pytest.mark.parametrize("variant", [api.ROLZ, api.LZP, api.LZ77])
def test_gpu_decompress_text(gpulib, variant):
    data = synth.markov_text(3 * MiB + 4321, seed=42)
    for who, container in _sources(variant, data, MiB, [], 0):
        with api.Handle(variant, lib=gpulib) as h:
            assert h.decompress(container, len(data) + 64) == data, who


@pytest.mark.parametrize("variant", [api.ROLZ, api.LZP, api.LZ77])
@pytest.mark.parametrize("data", [b"", b"A", bytes(MiB), (b"The quick brown fox jumps over the lazy dog. " * 30000)[:1300000]],
                         ids=["empty", "A", "zeros", "fox"])
def test_gpu_decompress_known_answers(gpulib, variant, data):
    for who, container in _sources(variant, data, 16 * MiB, [], 0):
        with api.Handle(variant, lib=gpulib) as h:
            assert h.decompress(container, len(data) + 64) == data, who


def test_gpu_decompress_filtered_x86(gpulib):
    data = synth.x86_corpus(3 * MiB, elf_bytes=MiB + 12345, pe_min=MiB // 2, pe_max=MiB)
    for who, container in _sources(api.ROLZ, data, MiB, ["-F"], 1):
        with api.Handle(api.ROLZ, lib=gpulib) as h:
            assert h.decompress(container, len(data) + 64) == data, who


def test_gpu_decompress_filtered_bmp(gpulib):
    data = synth.bmp_corpus(3 * MiB, wmin=301, wmax=900, hmin=100, hmax=500)
    for who, container in _sources(api.LZP, data, MiB, ["-F"], 1):
        with api.Handle(api.LZP, lib=gpulib) as h:
            assert h.decompress(container, len(data) + 64) == data, who


@pytest.mark.parametrize("variant", [api.ROLZ, api.LZP, api.LZ77])
def test_gpu_roundtrip_own_container(gpulib, variant):
    data = synth.markov_text(2 * MiB, seed=5) + synth.x86_corpus(MiB, elf_bytes=0, pe_min=MiB // 2, pe_max=MiB)
    with api.Handle(variant, lib=gpulib) as h:
        container = h.compress(data, MiB)
    with api.Handle(variant, lib=gpulib) as h:
        assert h.decompress(container, len(data) + 64) == data


@pytest.mark.parametrize("variant", [api.ROLZ, api.LZP, api.LZ77])
def test_gpu_decompress_scalar_kernel(gpulib, variant):
    """The single-thread decoder (the one the CPU simulation checks) agrees with the warp decoder on the GPU."""
    data = synth.markov_text(300000, seed=8)
    container = O.compress(data, variant, MiB)
    with api.Handle(variant, lib=gpulib) as h:
        h.set_option("scalar_models", 1)
        assert h.decompress(container, len(data) + 64) == data


@pytest.mark.parametrize("variant", [api.ROLZ, api.LZP, api.LZ77])
def test_gpu_decompress_batch(gpulib, variant):
    """Many containers in flight: one launch decodes all of them (one warp each); ragged, empty, filtered and -p members."""
    inputs = [
        (synth.markov_text(MiB + 777, seed=11), MiB, 0, 0),
        (b"", 16 * MiB, 0, 0),
        (synth.markov_text(2 * MiB, seed=12), MiB // 2, 0, 0),
        (synth.x86_corpus(2 * MiB, elf_bytes=MiB, pe_min=MiB // 2, pe_max=MiB) if variant == api.ROLZ
         else synth.bmp_corpus(2 * MiB, wmin=301, wmax=900, hmin=100, hmax=500), MiB, 1, 0),
        (b"A", 16 * MiB, 0, 0),
        (synth.markov_text(MiB, seed=13), MiB, 0, 1),
        (bytes(300000), MiB, 0, 0),
    ]
    containers = [O.compress(d, variant, bs, filt, prec) for d, bs, filt, prec in inputs]
    handles = [api.Handle(variant, lib=gpulib) for _ in inputs]
    try:
        for _ in range(2):                       # the second call reuses tables and model memory of the first
            outs = api.decompress_batch(handles, containers, [len(d) + 64 for d, *_ in inputs])
            for i, (d, *_r) in enumerate(inputs):
                assert outs[i] == d, i
        # a batch is the same as one call per container
        assert handles[0].decompress(containers[2], 2 * MiB + 64) == inputs[2][0]
    finally:
        for h in handles:
            h.close()


def test_gpu_decompress_batch_rejects_mixed_handles(gpulib):
    a, b = api.Handle(api.ROLZ, lib=gpulib), api.Handle(api.LZP, lib=gpulib)
    c = O.compress(b"abc", api.ROLZ, MiB, 0, 0)
    with pytest.raises(api.CrgpuError):
        api.decompress_batch([a, b], [c, c], [64, 64])
    with pytest.raises(api.CrgpuError):
        api.decompress_batch([a, a], [c, c], [64, 64])
    a.close(); b.close()
