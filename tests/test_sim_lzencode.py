"""CPU (no GPU) check of the kernels' integer logic: the CUDA sources compiled with -DCRGPU_SIM run every
independent-thread kernel sequentially.  This is a pre-flight for the GPU parity tests, not a product path."""
import pytest

import cases
import oracle_ffi as O
from comprox_b200 import api


@pytest.mark.parametrize("variant", [api.ROLZ, api.LZP])
@pytest.mark.parametrize("name", ["zeros", "fox", "rawtext", "periodic", "short_tail", "one_byte", "sub16", "exact16"])
def test_sim_lzencode_matches_oracle(simlib, variant, name):
    blocks = cases.lz_cases()[name]
    orc = O.Oracle(variant)
    want = [orc.lzencode(b) for b in blocks]
    with api.Handle(variant, lib=simlib) as h:
        got = h.lzencode(blocks)
    assert [len(g) for g in got] == [len(w) for w in want]
    assert got == want


@pytest.mark.parametrize("variant", [api.ROLZ, api.LZP])
def test_sim_chain_across_calls(simlib, variant):
    """Models carry over between calls exactly as between blocks."""
    blocks = cases.dict_coded_text(1 << 20, 1 << 18, seed=5, variant=variant)
    orc = O.Oracle(variant)
    want = [orc.lzencode(b) for b in blocks]
    with api.Handle(variant, lib=simlib) as h:
        got = h.lzencode(blocks[:1], chain_ends=False) + h.lzencode(blocks[1:3], chain_ends=False) + h.lzencode(blocks[3:])
    assert got == want
    # and reset_models() really resets
    orc.reset_models()
    want2 = orc.lzencode(blocks[0])
    with api.Handle(variant, lib=simlib) as h:
        h.lzencode(blocks[1:2], chain_ends=False)
        h.reset_models()
        assert h.lzencode(blocks[:1])[0] == want2


def test_sim_midchain_abort_is_loud(simlib):
    """With the exact replay switched off a mid-chain "cannot compress" is a loud error (never a silently wrong stream); with it on
    (the default) the same call reproduces the reference, see tests/test_zz_midchain_abort.py."""
    with api.Handle(api.ROLZ, lib=simlib) as h:
        h.set_option("exact_aborts", 0)
        with pytest.raises(api.CrgpuError) as e:
            h.lzencode([b"abc", b"hello hello hello hello"])
        assert e.value.code == -6
