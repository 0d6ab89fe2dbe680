"""Deterministic inputs of the golden vectors (tests/golden/kat.json): name -> (maker, [(reference binary, flags)])."""
from comprox_b200 import synth

MiB = 1 << 20
BOTH = lambda *flags: [("comprolz", list(flags)), ("comprop", list(flags))]  # noqa: E731

ROX = lambda *flags: [("comprox", list(flags))]  # noqa: E731   (the LZ77 front-end)

INPUTS = {
    # SURVEY.md App. E generator-free known answers
    "empty": (lambda: b"", BOTH() + ROX()),
    "single_A": (lambda: b"A", BOTH() + ROX()),
    "zeros_1MiB": (lambda: bytes(MiB), BOTH() + ROX()),
    "ramp_1MiB": (lambda: bytes(((i * 7 + (i >> 8)) & 255) for i in range(MiB)), BOTH() + ROX()),
    "fox": (lambda: (b"The quick brown fox jumps over the lazy dog. " * 30000)[:1300000], BOTH() + [("comprolz", ["-b1"])] + ROX()),
    # seeded synthetic corpora (comprox_b200/synth.py)
    "text_3MiB": (lambda: synth.markov_text(3 * MiB + 12345, seed=7), BOTH("-b1") + BOTH() + [("comprolz", ["-b1", "-p"]), ("comprolz", ["-b1", "-f"])]
                  + ROX("-b1") + ROX() + ROX("-b1", "-f") + ROX("-b1", "-m8") + ROX("-b1", "-p")),
    "text_2MiB_exact": (lambda: synth.markov_text(2 * MiB, seed=44), BOTH("-b1") + ROX("-b1")),
    "x86_3MiB": (lambda: synth.x86_corpus(3 * MiB, elf_bytes=MiB + 12345, pe_min=MiB // 2, pe_max=MiB), BOTH("-b1", "-F") + [("comprolz", ["-F"]), ("comprolz", ["-b1"])] + ROX("-b1", "-F")),
    "bmp_3MiB": (lambda: synth.bmp_corpus(3 * MiB, wmin=301, wmax=900, hmin=100, hmax=500), BOTH("-b1", "-F") + [("comprop", ["-F"])] + ROX("-b1", "-F")),
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


def parse_match_limit(flags):
    """-mN of the comprox (LZ77) front-end; 0 = default."""
    for f in flags:
        if f.startswith("-m"):
            return int(f[2:])
    return 0
