"""crgpu_compress_batch (shard mode): independent containers side by side, one host thread per handle inside the
C call, containers dealt to whichever handle is idle.  Every container must be exactly what crgpu_compress / the oracle gives for that
input, whichever handle took it."""
import pytest

import oracle_ffi as O
from comprox_b200 import api, synth

BACKENDS = [pytest.param("sim", id="sim"), pytest.param("gpu", id="gpu", marks=pytest.mark.gpu)]


def _lib(request, which):
    return request.getfixturevalue("simlib" if which == "sim" else "gpulib")


@pytest.mark.parametrize("which", BACKENDS)
@pytest.mark.parametrize("variant", [api.ROLZ, api.LZP])
def test_batch_matches_oracle(request, which, variant):
    lib = _lib(request, which)
    scale = 1 if which == "sim" else 8
    datas = [synth.markov_text(150000 * scale + 1000 * i, seed=60 + i) for i in range(5)] + [b"", b"tiny", synth.bmp_corpus(200000, seed=3, wmin=40, wmax=200, hmin=20, hmax=90)]
    want = [O.compress(d, variant, 65536 * scale, filt=1) for d in datas]
    stream = api.OWN_STREAM if which == "gpu" else None
    hs = [api.Handle(variant, stream=stream, lib=lib) for _ in range(3)]
    try:
        assert api.compress_batch(hs, datas, 65536 * scale, filt=True) == want
        assert api.compress_batch(hs[:1], datas[:2], 65536 * scale, filt=True) == want[:2]      # fewer handles than containers, and reuse
        assert api.compress_batch(hs, [], 65536) == []
    finally:
        for h in hs:
            h.close()


@pytest.mark.parametrize("which", BACKENDS)
def test_batch_rejects_bad_handle_sets(request, which):
    lib = _lib(request, which)
    with api.Handle(api.ROLZ, lib=lib) as a, api.Handle(api.LZP, lib=lib) as b:
        with pytest.raises(api.CrgpuError):
            api.compress_batch([a, b], [b"x"])            # mixed variants
        with pytest.raises(api.CrgpuError):
            api.compress_batch([a, a], [b"x"])            # the same handle twice
    if which == "gpu":
        with api.Handle(api.ROLZ, lib=lib) as a, api.Handle(api.ROLZ, lib=lib) as b:
            with pytest.raises(api.CrgpuError):
                api.compress_batch([a, b], [b"x"])        # two handles on one stream would serialise: private streams are required
