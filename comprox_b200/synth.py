"""Deterministic synthetic corpora for the comprox hot path.

All generators are pure numpy, seeded, and fast enough to run inside bench.py on the GPU box
(no dataset can be downloaded there).  They stay inside the reference's own input envelope
(SURVEY.md App. D): fixed vocabulary (< 325 001 distinct words, F10), at most one ELF image (F3),
homogeneous blocks for filtered data (F4), compressible everywhere (F11).
"""
from __future__ import annotations

import numpy as np

MiB = 1 << 20


# --------------------------------------------------------------------------- text
def _vocabulary(rng: np.random.Generator, nwords: int):
    """`nwords` distinct lowercase words, lengths 2..14, as (flat bytes, offsets, lengths)."""
    letters = np.frombuffer(b"etaoinshrdlcumwfgypbvkjxqz", dtype=np.uint8)
    lw = 1.0 / np.arange(1, 27) ** 0.8
    lw /= lw.sum()
    seen, words = set(), []
    while len(words) < nwords:
        n = int(rng.integers(2, 15)) if len(words) > 50 else int(rng.integers(2, 5))
        w = bytes(letters[rng.choice(26, size=n, p=lw)])
        if w not in seen and w != b"http":
            seen.add(w)
            words.append(w)
    lens = np.array([len(w) for w in words], dtype=np.int64)
    offs = np.concatenate([[0], np.cumsum(lens)[:-1]])
    flat = np.frombuffer(b"".join(words), dtype=np.uint8)
    return flat, offs, lens


def markov_text(nbytes: int, seed: int = 42, vocab: int = 20000, chains: int = 4096) -> bytes:
    """English-like text from a word-level first-order Markov chain over a fixed vocabulary.

    Zipf(1.0) unigram prior; each word has 16 preferred successors (followed with p=0.85);
    separators " " 80 %, ", " 8 %, ". " 8 %, "; " 2 %, ": " 1 %, ".\\n" 1 %; sentence starts capitalised.
    `chains` independent chains are advanced in lock-step (vectorised) and laid out one after another.
    """
    if nbytes <= 0:
        return b""
    rng = np.random.default_rng(seed)
    flat, offs, lens = _vocabulary(rng, vocab)
    zipf = 1.0 / np.arange(1, vocab + 1)
    zipf /= zipf.sum()
    zcdf = np.cumsum(zipf)
    succ = np.searchsorted(zcdf, rng.random((vocab, 16))).astype(np.int32).clip(0, vocab - 1)
    kw = 1.0 / np.arange(1, 17)
    kcdf = np.cumsum(kw / kw.sum())

    avg = float((lens * zipf).sum()) + 1.3
    total_words = int(nbytes / avg * 1.15) + 64
    chains = max(1, min(chains, total_words // 64 or 1))
    steps = -(-total_words // chains)
    state = np.searchsorted(zcdf, rng.random(chains)).astype(np.int32).clip(0, vocab - 1)
    seq = np.empty((steps, chains), dtype=np.int32)
    for t in range(steps):
        seq[t] = state
        r = rng.random(chains)
        k = np.searchsorted(kcdf, rng.random(chains)).clip(0, 15)
        fresh = np.searchsorted(zcdf, rng.random(chains)).astype(np.int32).clip(0, vocab - 1)
        state = np.where(r < 0.85, succ[state, k], fresh)
    words = seq.T.reshape(-1)  # chain-major: consecutive words follow the chain
    nw = words.size

    seps = [b" ", b", ", b". ", b"; ", b": ", b".\n"]
    sep_flat = np.frombuffer(b"".join(seps), dtype=np.uint8)
    sep_len = np.array([len(s) for s in seps], dtype=np.int64)
    sep_off = np.concatenate([[0], np.cumsum(sep_len)[:-1]])
    sid = np.searchsorted(np.cumsum([0.80, 0.08, 0.08, 0.02, 0.01, 0.01]), rng.random(nw)).clip(0, 5)
    cap = np.empty(nw, dtype=bool)
    cap[0] = True
    cap[1:] = (sid[:-1] == 2) | (sid[:-1] == 5)

    out = np.empty(nbytes + 64, dtype=np.uint8)
    pos, CH = 0, 1 << 20
    for a in range(0, nw, CH):
        w = words[a:a + CH]
        wl = lens[w]
        sl = sep_len[sid[a:a + CH]]
        tl = wl + sl
        starts = np.cumsum(tl) - tl
        n = int(tl.sum())
        tok = np.repeat(np.arange(w.size), tl)
        within = np.arange(n) - starts[tok]
        isw = within < wl[tok]
        src = np.where(isw, offs[w][tok] + within, 0)
        chunk = np.where(isw, flat[src], sep_flat[(sep_off[sid[a:a + CH]][tok] + within - wl[tok]).clip(0, sep_flat.size - 1)])
        chunk = chunk.astype(np.uint8)
        first = starts[cap[a:a + CH]]
        chunk[first] -= 32
        take = min(n, out.size - pos)
        out[pos:pos + take] = chunk[:take]
        pos += take
        if pos >= nbytes:
            break
    assert pos >= nbytes, "generator under-produced"
    return out[:nbytes].tobytes()


# --------------------------------------------------------------------------- x86 (ELF + PE)
_X86_TEMPLATES = [
    b"\x55", b"\x89\xe5", b"\x5d", b"\xc3", b"\x90", b"\x8b\x45\x08", b"\x8b\x55\x0c", b"\x01\xd0",
    b"\x83\xec\x18", b"\x83\xc4\x10", b"\x89\x04\x24", b"\x85\xc0", b"\x74\x0a", b"\x75\xf2", b"\x31\xc0",
    b"\x8d\x4c\x24\x04", b"\xc7\x45\xf4\x00\x00\x00\x00", b"\x0f\xb6\x00", b"\x50", b"\x53", b"\x5b",
    b"\x89\x44\x24\x04", b"\x8b\x5d\xfc", b"\xc9", b"\x39\xc2", b"\x7e\x05", b"\xeb\x10", b"\x66\x90",
]


def _x86_code(rng: np.random.Generator, nbytes: int, call_frac: float = 0.18, ntargets: int = 4096) -> np.ndarray:
    """x86-like code: template instructions plus `call_frac` E8/E9 rel32 transfers to hot targets."""
    if nbytes <= 0:
        return np.zeros(0, dtype=np.uint8)
    tl = np.array([len(t) for t in _X86_TEMPLATES], dtype=np.int64)
    toff = np.concatenate([[0], np.cumsum(tl)[:-1]])
    tflat = np.frombuffer(b"".join(_X86_TEMPLATES), dtype=np.uint8)
    ninstr = int(nbytes / (tl.mean() * (1 - call_frac) + 5 * call_frac)) + 16
    is_call = rng.random(ninstr) < call_frac
    tid = rng.integers(0, len(_X86_TEMPLATES), ninstr)
    ilen = np.where(is_call, 5, tl[tid])
    start = np.cumsum(ilen) - ilen
    total = int(ilen.sum())
    targets = np.sort(rng.integers(0, max(nbytes - 16, 1), ntargets))
    zipf = 1.0 / np.arange(1, ntargets + 1)
    zcdf = np.cumsum(zipf / zipf.sum())
    code = np.empty(total, dtype=np.uint8)
    ins = np.repeat(np.arange(ninstr), ilen)
    within = np.arange(total) - start[ins]
    code[:] = tflat[(toff[tid][ins] + within).clip(0, tflat.size - 1)]
    cs = start[is_call]
    tgt = targets[rng.permutation(ntargets)[np.searchsorted(zcdf, rng.random(cs.size)).clip(0, ntargets - 1)]]
    rel = (tgt - (cs + 5)).astype(np.int64) & 0xFFFFFFFF
    code[cs] = np.where(rng.random(cs.size) < 0.9, 0xE8, 0xE9)
    for k in range(4):
        code[cs + 1 + k] = (rel >> (8 * k)) & 0xFF
    return code[:nbytes] if total >= nbytes else np.concatenate([code, np.full(nbytes - total, 0x90, np.uint8)])


def elf_image(rng: np.random.Generator, nbytes: int) -> bytes:
    """ELF32/i386 image of exactly `nbytes`; e_shoff marks the end of code (filter_x86_elf.c:106-129)."""
    hdr = bytearray(52)
    hdr[0:4] = b"\x7fELF"
    hdr[4:7] = b"\x01\x01\x01"
    hdr[16:18] = (2).to_bytes(2, "little")
    hdr[18:20] = (3).to_bytes(2, "little")          # EM_386
    hdr[20:24] = (1).to_bytes(4, "little")
    hdr[32:36] = int(nbytes).to_bytes(4, "little")  # e_shoff
    hdr[40:42] = (52).to_bytes(2, "little")
    return bytes(hdr) + _x86_code(rng, nbytes - 52).tobytes()


def pe_image(rng: np.random.Generator, nbytes: int) -> bytes:
    """PE32/i386 image of exactly `nbytes`: DOS stub, e_lfanew=0x80, COFF machine 0x14c, 2 sections."""
    opt, nsec = 224, 2
    size_hdr = 24 + opt + nsec * 40           # measured from the COFF header (filter_x86_pe.c:89-91)
    head = bytearray(0x80 + size_hdr)
    head[0:2] = b"MZ"
    head[0x3C:0x40] = (0x80).to_bytes(4, "little")
    head[0x80:0x84] = b"PE\0\0"
    head[0x84:0x86] = (0x14C).to_bytes(2, "little")
    head[0x86:0x88] = nsec.to_bytes(2, "little")
    head[0x94:0x96] = opt.to_bytes(2, "little")
    head[0x96:0x98] = (0x0102).to_bytes(2, "little")
    # filter_x86_pe.c:148-151 places the body at buf+size_hdr (e_lfanew not added) and consumes size+size_hdr
    body = nbytes - size_hdr

    def clean(v):   # no E8/E9 byte: the reference rewrites "operands" inside the section table too (the body is taken to
        return not any(((v >> (8 * k)) & 0xFE) == 0xE8 for k in range(4))   # start at buf+size_hdr), and its decoder re-reads it

    s0 = body // 2
    while not (clean(s0) and clean(body - s0)):
        s0 -= 4099
    for i, (nm, sz) in enumerate(((b".text\0\0\0", s0), (b".data\0\0\0", body - s0))):
        o = 0x80 + 24 + opt + i * 40
        head[o:o + 8] = nm
        head[o + 16:o + 20] = int(sz).to_bytes(4, "little")
    return bytes(head) + _x86_code(rng, nbytes - len(head)).tobytes()


def x86_corpus(nbytes: int, seed: int = 43, elf_bytes: int | None = None, pe_min: int = 4 * MiB, pe_max: int = 32 * MiB) -> bytes:
    """One ELF image followed by PE images (at most one ELF per stream: SURVEY.md F3)."""
    rng = np.random.default_rng(seed)
    if elf_bytes is None:
        elf_bytes = min(64 * MiB, nbytes // 4)
    elf_bytes = max(min(elf_bytes, nbytes), 0)
    parts, left = [], nbytes
    if elf_bytes >= 4096:
        parts.append(elf_image(rng, elf_bytes))
        left -= elf_bytes
    while left > 0:
        n = int(rng.integers(pe_min, pe_max + 1)) if left > pe_max else left
        if left - n < 4096:
            n = left
        if n < 4096:
            parts.append(bytes(n))
        else:
            parts.append(pe_image(rng, n))
        left -= n
    return b"".join(parts)


# --------------------------------------------------------------------------- BMP
def bmp_image(rng: np.random.Generator, width: int, height: int) -> bytes:
    """24-bpp BMP (BITMAPINFOHEADER, image offset 54): smooth 2-D gradient + 2-bit noise, padded rows."""
    row = (24 * width + 31) // 32 * 4
    y = np.arange(height, dtype=np.float64)[:, None]
    x = np.arange(width, dtype=np.float64)[None, :]
    px = np.empty((height, width, 3), dtype=np.uint8)
    for ch in range(3):
        a, b, c = rng.uniform(0.02, 0.3, 3)
        ph = rng.uniform(0, 6.28)
        v = 128 + 60 * np.sin(a * x / 8 + ph) + 50 * np.cos(b * y / 8) + c * (x + y) / 8
        px[:, :, ch] = (v.astype(np.int64) + rng.integers(0, 4, (height, width))) & 0xFF
    data = np.zeros((height, row), dtype=np.uint8)
    data[:, :width * 3] = px.reshape(height, width * 3)
    hdr = bytearray(54)
    hdr[0:2] = b"BM"
    hdr[2:6] = (54 + row * height).to_bytes(4, "little")
    hdr[10:14] = (54).to_bytes(4, "little")
    hdr[14:18] = (40).to_bytes(4, "little")
    hdr[18:22] = width.to_bytes(4, "little")
    hdr[22:26] = height.to_bytes(4, "little")
    hdr[26:28] = (1).to_bytes(2, "little")
    hdr[28:30] = (24).to_bytes(2, "little")
    hdr[34:38] = (row * height).to_bytes(4, "little")
    return bytes(hdr) + data.tobytes()


def bmp_corpus(nbytes: int, seed: int = 44, wmin: int = 1021, wmax: int = 4099, hmin: int = 512, hmax: int = 2048) -> bytes:
    """Concatenated 24-bpp BMPs (widths include non-multiples of 4 -> row padding); zero-padded tail."""
    rng = np.random.default_rng(seed)
    parts, left = [], nbytes
    while left > 0:
        w = int(rng.integers(wmin, wmax + 1))
        h = int(rng.integers(hmin, hmax + 1))
        row = (24 * w + 31) // 32 * 4
        if 54 + row * h > left:
            h = (left - 54) // row
            if h < 4:
                parts.append(bytes(left))
                break
        img = bmp_image(rng, w, h)
        parts.append(img)
        left -= len(img)
    return b"".join(parts)


def mixed_corpus(nbytes: int, seed: int = 45, segment: int = 64 * MiB) -> bytes:
    """Segments of text / x86 / BMP, each a multiple of `segment` bytes so that every block of any swept
    size dividing `segment` is homogeneous."""
    parts, left, k = [], nbytes, 0
    first_x86 = True
    while left > 0:
        n = min(segment, left)
        kind = k % 3
        if kind == 0:
            parts.append(markov_text(n, seed + 100 * k))
        elif kind == 1:
            parts.append(x86_corpus(n, seed + 100 * k, elf_bytes=None if first_x86 else 0))
            first_x86 = False
        else:
            parts.append(bmp_corpus(n, seed + 100 * k))
        left -= n
        k += 1
    return b"".join(parts)
