"""Pins the CPU oracle (oracle/cr_oracle.c) to the unmodified reference:
  * against tests/golden/kat.json (digests of containers written by the reference build, committed), and
  * differentially against oracle/_ref/{comprolz,comprop} when those binaries are present (build container and,
    because oracle/_ref travels with the snapshot, the GPU box)."""
import hashlib
import json
import os

import pytest

import oracle_ffi as O
from golden_inputs import INPUTS, parse_flags, parse_match_limit
from comprox_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))
KAT = json.load(open(os.path.join(HERE, "golden", "kat.json")))
VARIANT = {"comprolz": 0, "comprop": 1, "comprox": 2}
_cache = {}


def _input(name):
    if name not in _cache:
        _cache[name] = INPUTS[name][0]()
    return _cache[name]


@pytest.mark.parametrize("key", sorted(KAT.keys()))
def test_oracle_matches_golden(key):
    name, binary, flags = key.split("|")
    data = _input(name)
    assert hashlib.sha256(data).hexdigest() == KAT[key]["input_sha256"], "generator drifted: regenerate tests/golden/kat.json"
    bs, filt, prec, flex = parse_flags(flags.split())
    c = O.compress(data, VARIANT[binary], bs, filt, prec, flex, parse_match_limit(flags.split()))
    assert len(c) == KAT[key]["container_bytes"]
    assert hashlib.sha256(c).hexdigest() == KAT[key]["container_sha256"]
    assert O.decompress(c, VARIANT[binary]) == data


@pytest.mark.skipif(O.ref_binary("comprolz") is None, reason="oracle/_ref not built")
@pytest.mark.parametrize("binary", ["comprolz", "comprop", "comprox"])
def test_oracle_matches_reference_cli_fresh_input(binary):
    """An input that is NOT in the golden set, straight against the reference binary."""
    data = synth.markov_text((1 << 20) + 333, seed=1234) + synth.x86_corpus(1 << 20, seed=5, elf_bytes=0, pe_min=1 << 19, pe_max=1 << 20)
    for flags in (["-b1"], ["-b1", "-F"]):
        bs, filt, prec, flex = parse_flags(flags)
        want = O.ref_compress(data, binary, flags)
        assert O.compress(data, VARIANT[binary], bs, filt, prec, flex) == want
        if not filt:     # F4: the reference decoder itself cannot undo filters on dictionary-compressible blocks
            assert O.ref_decompress(want, binary) == data


def test_oracle_stage_traces_are_consistent():
    """tokens -> events -> triples of the oracle agree with each other (used by the GPU diagnosis helper)."""
    data = synth.markov_text(300000, seed=9)
    orc = O.Oracle(0)
    orc.trace(True)
    payload = orc.lzencode(data)
    toks, ev, tr = orc.tokens(), orc.events(), orc.triples()
    assert sum(t[1] for t in toks) == len(data) - 1
    assert len(ev) == len(toks)
    assert (tr[:, 3] == 0).sum() >= len(ev)
    assert payload[1] == 1
    assert toks == O.rolz_parse(data)


@pytest.mark.skipif(O.ref_binary("comprop") is None, reason="oracle/_ref not built")
def test_oracle_prune_matches_reference_cli():
    """More than 325 000 distinct words: the oracle's restatement of the order-dependent prune (cr-dicpick.c:115-144)."""
    from vocab_overflow_input import overflow_text
    data = overflow_text()
    assert O.compress(data, 1, 16 << 20) == O.ref_compress(data, "comprop", [])
