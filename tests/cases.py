"""Seeded inputs shared by the simulation (CPU) and GPU parity tests."""
import numpy as np

from comprox_b200 import synth


def dict_coded_text(nbytes, block, seed, variant):
    """Dictionary-coded blocks of Markov text, produced by the ORACLE's dictionary stage (test input only)."""
    import oracle_ffi as O
    text = synth.markov_text(nbytes, seed=seed)
    orc = O.Oracle(variant)
    orc.dictionary_load(O.dicpick(text))
    return [orc.dictionary_encode(text[i:i + block]) for i in range(0, len(text), block)]


def lz_cases(scale=1):
    """name -> list of consecutive blocks of one chain (only the last block may be incompressible)."""
    s = scale
    return {
        "zeros": [bytes(300000 * s)],
        "ramp": [bytes(((i * 7 + (i >> 8)) & 255) for i in range(200000))],
        "fox": [(b"The quick brown fox jumps over the lazy dog. " * 8000)[:300000]] * 2,
        "rawtext": [synth.markov_text(600000 * s, seed=3)[i:i + 200000 * s] for i in range(0, 600000 * s, 200000 * s)],
        "x86": [synth.x86_corpus(s << 20, elf_bytes=300000, pe_min=200000, pe_max=400000)[i:i + (s << 19)] for i in (0, s << 19)],
        "bmp_raw": [synth.bmp_corpus(1 << 20, wmin=100, wmax=300, hmin=50, hmax=200)],
        "periodic": [bytes((i % 3) for i in range(100000)), bytes((i % 17) for i in range(100000))],
        "repeat_text": [synth.markov_text(100000, seed=9), synth.markov_text(100000, seed=9)],
        "short_tail": [synth.markov_text(150000, seed=11), b"tail"],
        "one_byte": [b"\x00"],
        "sub16": [b"0123456789abcde"],
        "exact16": [b"0123456789abcdef"],
    }
