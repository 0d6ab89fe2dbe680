"""Stage-level C ABI (include/crgpu.h: crgpu_filter_inplace, crgpu_dic_lcp_*, crgpu_dictionary_load/encode/decode,
crgpu_lzencode, crgpu_lzdecode): one block per call, like the reference's cr-* functions, checked against the
oracle's stage functions block by block.  The same cases run on the kernel-logic simulation (CPU pre-flight) and, marked
gpu, on the real library."""
import numpy as np
import pytest

import oracle_ffi as O
from comprox_b200 import api, synth

MiB = 1 << 20


def _lib(request, which):
    return request.getfixturevalue("simlib" if which == "sim" else "gpulib")


BACKENDS = [pytest.param("sim", id="sim"), pytest.param("gpu", id="gpu", marks=pytest.mark.gpu)]


def _blocks(data, bs):
    return [data[i:i + bs] for i in range(0, len(data), bs)] or [b""]


@pytest.mark.parametrize("which", BACKENDS)
def test_filter_inplace_per_block_matches_oracle(request, which):
    """Images straddle calls exactly as they straddle blocks; FILTER_DEC on a second handle restores the input."""
    lib = _lib(request, which)
    data = (synth.x86_corpus(700000, elf_bytes=200000, pe_min=150000, pe_max=250000) + b"plain text in between " * 50
            + synth.bmp_corpus(600000, seed=7, wmin=30, wmax=300, hmin=10, hmax=120))
    orc = O.Oracle(api.ROLZ)
    with api.Handle(api.ROLZ, lib=lib) as enc, api.Handle(api.ROLZ, lib=lib) as dec:
        any_fired = 0
        for blk in _blocks(data, 262144 + 13):
            want_f, want = orc.filter_inplace(blk, 0)
            got_f, got = enc.filter_inplace(blk, 0)
            assert (got_f, got) == (want_f, want)
            any_fired |= got_f
            back_f, back = dec.filter_inplace(got, 1)
            assert back == blk
        assert any_fired == 1
        assert enc.filter_inplace(b"", 0) == (0, b"")


@pytest.mark.parametrize("which", BACKENDS)
def test_dictionary_stage_matches_oracle(request, which):
    lib = _lib(request, which)
    data = synth.markov_text(2 * MiB + 777, seed=31)
    text = O.dicpick(data)
    assert api.lcp_encode(text, lib=lib) == O.lcp_encode(text)
    assert api.lcp_decode(O.lcp_encode(text), lib=lib) == text
    orc = O.Oracle(api.ROLZ)
    with api.Handle(api.ROLZ, lib=lib) as h, api.Handle(api.ROLZ, lib=lib) as d:
        assert h.dicpick(data) == text
        assert h.dictionary_load(text, 1) == orc.dictionary_load(text, 1)
        assert d.dictionary_load(text, 0) == h.dictionary_load(text, 1)
        rng = np.random.default_rng(3)
        cases = _blocks(data, MiB + 5) + [b"", b"x", rng.integers(0, 256, 70000, dtype=np.uint8).tobytes()]
        for blk in cases:
            want = orc.dictionary_encode(blk)
            got = h.dictionary_encode(blk)
            assert got == want
            assert d.dictionary_decode(got, len(blk) + 64) == blk == orc.dictionary_decode(want)
        with pytest.raises(api.CrgpuError):
            d.dictionary_decode(h.dictionary_encode(cases[0]), 10)       # out_cap too small is an error, not a truncation


@pytest.mark.parametrize("which", BACKENDS)
def test_dictionary_encode_needs_a_dictionary(request, which):
    with api.Handle(api.ROLZ, lib=_lib(request, which)) as h:
        with pytest.raises(api.CrgpuError):
            h.dictionary_encode(b"hello world")


@pytest.mark.parametrize("which", BACKENDS)
@pytest.mark.parametrize("variant", [api.ROLZ, api.LZP, api.LZ77])
def test_lzencode_lzdecode_per_block(request, which, variant):
    """The reference's call order (src/main.c:163-165,194 / 253-256,277): dictionary payload, reset_models, then one call per
    block with the models carried from block to block -- on the encoder and, mirrored, on the decoder."""
    lib = _lib(request, which)
    data = synth.markov_text(600000, seed=33)
    text = O.dicpick(data)
    orc = O.Oracle(variant)
    orc.dictionary_load(text, 1)
    coded = [orc.dictionary_encode(b) for b in _blocks(data, 200000)] + [b"\x00"]        # + the trailing empty block (F8)
    lcp = O.lcp_encode(text)
    orc.reset_models()
    want_dic = orc.lzencode(lcp)
    orc.reset_models()
    want = [orc.lzencode(b) for b in coded]
    with api.Handle(variant, lib=lib) as enc, api.Handle(variant, lib=lib) as dec:
        enc.reset_models()
        got_dic = enc.lzencode([lcp], chain_ends=True)[0]
        assert got_dic == want_dic
        enc.reset_models()
        got = [enc.lzencode([b], chain_ends=(i == len(coded) - 1))[0] for i, b in enumerate(coded)]
        assert got == want
        dec.reset_models()
        assert dec.lzdecode(got_dic) == lcp
        dec.reset_models()
        for p, b in zip(got, coded):
            assert dec.lzdecode(p) == b
    orc2 = O.Oracle(variant)
    orc2.reset_models()
    assert orc2.lzdecode(want_dic) == lcp


@pytest.mark.parametrize("which", BACKENDS)
def test_lzdecode_rejects_truncated_payload(request, which):
    with api.Handle(api.ROLZ, lib=_lib(request, which)) as h:
        with pytest.raises(api.CrgpuError):
            h.lzdecode(b"\x00\x01\x02")


@pytest.mark.parametrize("which", BACKENDS)
def test_dictionary_with_more_candidates_than_places(request, which):
    """More than 25 000 words with a count above 5 (ties at every count): the dictionary keeps the first 24 998 of the order by
    (count, word) and re-sorts all but the first ~155 by word -- done here by selection instead of two full sorts
    (hd_dictionary_text), so the text, the trie built from it and a coded block must still equal the oracle's."""
    lib = _lib(request, which)
    rng = np.random.default_rng(17)
    vocab = 32000
    lens = rng.integers(3, 9, vocab)
    letters = rng.integers(97, 123, (vocab, 8)).astype(np.uint8)
    words = [letters[w, :lens[w]].tobytes() for w in range(vocab)]
    occ = np.concatenate([np.repeat(np.arange(vocab), rng.integers(6, 10, vocab)), rng.integers(0, 400, 60000)])
    rng.shuffle(occ)
    data = b" ".join(words[w] for w in occ) + b" "
    text = O.dicpick(data)
    assert text.count(b"\n") > 24000
    orc = O.Oracle(api.ROLZ)
    with api.Handle(api.ROLZ, lib=lib) as h:
        assert h.dicpick(data) == text
        assert h.dictionary_load(text, 1) == orc.dictionary_load(text, 1)
        blk = data[:300000]
        assert h.dictionary_encode(blk) == orc.dictionary_encode(blk)
