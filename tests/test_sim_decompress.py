"""CPU pre-flight (kernel-logic simulation) of crgpu_decompress: decodes containers written by the ORACLE."""
import pytest

import oracle_ffi as O
from comprox_b200 import api, synth

MiB = 1 << 20


def _cases():
    text = synth.markov_text(MiB + 12345, seed=7)
    return {
        "text": (text, MiB // 2, 0, 0),
        "empty": (b"", MiB, 0, 0),
        "single_byte": (b"A", MiB, 0, 0),
        "fox": ((b"The quick brown fox jumps over the lazy dog. " * 30000)[:1300000], MiB, 0, 0),
        "prec": (text[:MiB // 2], MiB, 0, 1),
        "x86_filtered": (synth.x86_corpus(MiB, elf_bytes=MiB // 4 + 77, pe_min=MiB // 8, pe_max=MiB // 4), MiB // 2, 1, 0),
        "bmp_filtered": (synth.bmp_corpus(MiB, wmin=201, wmax=500, hmin=60, hmax=300), MiB // 2, 1, 0),
    }


@pytest.mark.parametrize("variant", [api.ROLZ, api.LZP, api.LZ77])
@pytest.mark.parametrize("name", sorted(_cases().keys()))
def test_sim_decompress_oracle_container(simlib, variant, name):
    data, bs, filt, prec = _cases()[name]
    container = O.compress(data, variant, bs, filt, prec)
    with api.Handle(variant, lib=simlib) as h:
        assert h.decompress(container, len(data) + 64) == data


def test_sim_decompress_rejects_wrong_magic(simlib):
    container = O.compress(b"hello hello hello", api.ROLZ)
    with api.Handle(api.LZP, lib=simlib) as h:
        with pytest.raises(api.CrgpuError):
            h.decompress(container, 1024)


def test_sim_decompress_batch_host_flow(simlib):
    """crgpu_decompress_batch: argument checks and the per-container phases (begin / middle / finish) on the simulation."""
    datas = [synth.markov_text(200000 + 999 * i, seed=20 + i) for i in range(3)] + [b""]
    conts = [O.compress(d, api.ROLZ, MiB // 4, 0, 0) for d in datas]
    hs = [api.Handle(api.ROLZ, lib=simlib) for _ in datas]
    try:
        outs = api.decompress_batch(hs, conts, [len(d) + 64 for d in datas])
        assert outs == datas
        with pytest.raises(api.CrgpuError):
            api.decompress_batch([hs[0], hs[0]], conts[:2], [1 << 20, 1 << 20])
    finally:
        for h in hs:
            h.close()


@pytest.mark.parametrize("variant", [api.ROLZ, api.LZP, api.LZ77])
def test_sim_decompress_reports_needed_size_and_resumes(simlib, variant):
    """The container does not store its decoded size.  With out_cap too small crgpu_decompress returns CRGPU_ERR_ARG and the size needed;
    the second call with the same container resumes behind the entropy stage (the models are NOT decoded twice) and gives the same bytes."""
    import ctypes
    data = synth.markov_text(400000, seed=71)
    cont = O.compress(data, variant, 150000)
    with api.Handle(variant, lib=simlib) as h:
        out = ctypes.create_string_buffer(1000)
        n = ctypes.c_uint64(0)
        rc = simlib.crgpu_decompress(h.h, cont, ctypes.c_uint64(len(cont)), out, ctypes.c_uint64(1000), ctypes.byref(n))
        assert rc == -3 and n.value == len(data)
        assert h.decompress(cont, n.value) == data                    # resumed
        assert h.decompress(cont, len(data) + 64) == data             # and a fresh full decode afterwards is unaffected
        assert h.decompress(cont, 10, grow=True) == data
        other = O.compress(data[::-1], variant, 150000)
        with pytest.raises(api.CrgpuError):
            h.decompress(cont, 10)                                     # leaves a resume point for `cont` ...
        assert h.decompress(other, len(data) + 64) == data[::-1]      # ... which a different container must not pick up
