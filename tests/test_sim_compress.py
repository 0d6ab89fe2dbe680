"""CPU pre-flight (kernel-logic simulation, see cr_common.cuh) of the whole-container path against the oracle."""
import pytest

import oracle_ffi as O
from comprox_b200 import api, synth

MiB = 1 << 20


def _cases():
    text = synth.markov_text(MiB + MiB // 2 + 12345, seed=7)
    return {
        "text_b1": (text, MiB, 0),
        "text_exact_multiple": (text[:MiB], MiB // 2, 0),        # trailing empty block
        "empty": (b"", MiB, 0),
        "single_byte": (b"A", MiB, 0),
        "fox": ((b"The quick brown fox jumps over the lazy dog. " * 30000)[:1300000], MiB, 0),
        "text_prec": (text[:MiB + 5], MiB, 1),
    }


@pytest.mark.parametrize("variant", [api.ROLZ, api.LZP])
@pytest.mark.parametrize("name", sorted(_cases().keys()))
def test_sim_compress_matches_oracle(simlib, variant, name):
    data, bs, prec = _cases()[name]
    want = O.compress(data, variant, bs, 0, prec)
    with api.Handle(variant, lib=simlib) as h:
        got = h.compress(data, bs, prec=bool(prec))
    assert len(got) == len(want)
    assert got == want


def test_sim_flexible_parsing(simlib):
    """-f (comprolz only): shortened matches priced against the table as of the parse point."""
    data = synth.markov_text(MiB // 2, seed=3) * 2 + synth.x86_corpus(MiB // 2, elf_bytes=0, pe_min=MiB // 4, pe_max=MiB // 2)
    want = O.compress(data, api.ROLZ, MiB // 2, 0, 0, 1)
    assert want != O.compress(data, api.ROLZ, MiB // 2)
    with api.Handle(api.ROLZ, lib=simlib) as h:
        assert h.compress(data, MiB // 2, flexible=True) == want
    with api.Handle(api.LZP, lib=simlib) as h:
        with pytest.raises(api.CrgpuError):
            h.compress(data, MiB // 2, flexible=True)


def test_sim_more_blocks_than_one_window(simlib):
    """ROLZ windows hold at most 64 blocks (sort-key layout); models must carry across the split."""
    data = synth.markov_text(67 * 4096, seed=12)
    want = O.compress(data, api.ROLZ, 4096)
    with api.Handle(api.ROLZ, lib=simlib) as h:
        assert h.compress(data, 4096) == want
    blocks = [data[i:i + 3000] for i in range(0, 66 * 3000, 3000)]
    orc = O.Oracle(api.ROLZ)
    want_payloads = [orc.lzencode(b) for b in blocks]
    with api.Handle(api.ROLZ, lib=simlib) as h:
        assert h.lzencode(blocks) == want_payloads


def test_sim_dicpick_vocabulary_overflow(simlib):
    """> 325 000 distinct words: the prune epochs are replayed exactly (the oracle's prune is pinned to the reference
    CLI in tests/test_oracle_vs_ref.py)."""
    from vocab_overflow_input import overflow_text
    data = overflow_text()
    want = O.dicpick(data)
    with api.Handle(api.ROLZ, lib=simlib) as h:
        assert h.dicpick(data) == want
    text = synth.markov_text(300000, seed=2)
    with api.Handle(api.LZP, lib=simlib) as h:
        assert h.dicpick(text) == O.dicpick(text)
