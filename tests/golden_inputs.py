"""Deterministic inputs of the golden vectors (tests/golden/kat.json): name -> (maker, [(reference binary, flags)])."""
from comprox_b200 import synth

MiB = 1 << 20
BOTH = lambda *flags: [("comprolz", list(flags)), ("comprop", list(flags))]  # noqa: E731

INPUTS = {
    # SURVEY.md App. E generator-free known answers
    "empty": (lambda: b"", BOTH()),
    "single_A": (lambda: b"A", BOTH()),
    "zeros_1MiB": (lambda: bytes(MiB), BOTH()),
    "ramp_1MiB": (lambda: bytes(((i * 7 + (i >> 8)) & 255) for i in range(MiB)), BOTH()),
    "fox": (lambda: (b"The quick brown fox jumps over the lazy dog. " * 30000)[:1300000], BOTH() + [("comprolz", ["-b1"])]),
    # seeded synthetic corpora (comprox_b200/synth.py)
    "text_3MiB": (lambda: synth.markov_text(3 * MiB + 12345, seed=7), BOTH("-b1") + BOTH() + [("comprolz", ["-b1", "-p"]), ("comprolz", ["-b1", "-f"])]),
    "text_2MiB_exact": (lambda: synth.markov_text(2 * MiB, seed=44), BOTH("-b1")),
    "x86_3MiB": (lambda: synth.x86_corpus(3 * MiB, elf_bytes=MiB + 12345, pe_min=MiB // 2, pe_max=MiB), BOTH("-b1", "-F") + [("comprolz", ["-F"]), ("comprolz", ["-b1"])]),
    "bmp_3MiB": (lambda: synth.bmp_corpus(3 * MiB, wmin=301, wmax=900, hmin=100, hmax=500), BOTH("-b1", "-F") + [("comprop", ["-F"])]),
}


def parse_flags(flags):
    bs, filt, prec, flex = 16 * MiB, 0, 0, 0
    for f in flags:
        if f.startswith("-b"):
            bs = int(f[2:]) * MiB
        elif f == "-F":
            filt = 1
        elif f == "-p":
            prec = 1
        elif f == "-f":
            flex = 1
    return bs, filt, prec, flex
