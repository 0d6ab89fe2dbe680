"""CPU pre-flight (kernel-logic simulation) of the -F pre-filters against the oracle: the 16-bytes-per-thread scan and BMP
kernels (cr_filter.cuh) on padded rows, 32-bpp pixels, tiny images and block sizes that leave windows unaligned."""
import struct

import numpy as np
import pytest

import oracle_ffi as O
from comprox_b200 import api, synth

MiB = 1 << 20


def _bmp32(rng, w, h):
    """A 32-bpp BMP the reference accepts (src/filter_bmp.c:149-170): BITMAPINFOHEADER, offset 54, no compression."""
    row = 4 * w
    pix = rng.integers(0, 256, size=row * h, dtype=np.uint8)
    pix = (np.cumsum(pix.astype(np.uint32) & 3) & 255).astype(np.uint8)       # smooth-ish
    hdr = struct.pack("<2sIHHIIiiHHIIiiII", b"BM", 54 + row * h, 0, 0, 54, 40, w, h, 1, 32, 0, row * h, 2835, 2835, 0, 0)
    return hdr + pix.tobytes()


def _corpus():
    rng = np.random.default_rng(5)
    parts = [synth.bmp_corpus(700000, seed=3, wmin=5, wmax=40, hmin=4, hmax=60),         # rows shorter than a 16-byte chunk
             _bmp32(rng, 37, 50), b"xyz", _bmp32(rng, 4, 4), _bmp32(rng, 301, 40), b"q", _bmp32(rng, 130, 33), b"qq", _bmp32(rng, 63, 21), b"BM", b"MZ", b"\x7fELF",
             synth.bmp_corpus(900000, seed=4, wmin=301, wmax=700, hmin=40, hmax=200)]
    return b"".join(parts)


@pytest.mark.parametrize("bs", [MiB, 300007, 65536 + 8])
@pytest.mark.parametrize("variant", [api.LZP, api.ROLZ])
def test_sim_bmp_filter_matches_oracle(simlib, variant, bs):
    data = _corpus()
    with api.Handle(variant, lib=simlib) as h:
        got = h.compress(data, bs, filt=True, window_bytes=4 * bs)
    assert got == O.compress(data, variant, bs, filt=1)


def test_sim_x86_filter_matches_oracle(simlib):
    data = synth.x86_corpus(MiB + 4321, elf_bytes=300000, pe_min=200000, pe_max=400000)
    with api.Handle(api.ROLZ, lib=simlib) as h:
        got = h.compress(data, 400003, filt=True, window_bytes=800006)
    assert got == O.compress(data, api.ROLZ, 400003, filt=1)


def test_sim_bmp_rows_broken_at_block_ends(simlib):
    """Tiny images: a block often ends with less than one row left, bmp_transform returns 0 with its flag still set and the
    sub-filter loop re-enters it at the same position (src/filter_bmp.c:188-203, src/cr-filter.c:60-68)."""
    data = synth.bmp_corpus(3 * MiB, seed=3, wmin=5, wmax=40, hmin=4, hmax=60)
    want = O.compress(data, api.LZP, MiB, filt=1)
    ref = O.ref_compress(data, "comprop", ["-b1", "-F"])
    if ref is not None:
        assert want == ref, "oracle differs from the reference CLI"
    with api.Handle(api.LZP, lib=simlib) as h:
        assert h.compress(data, MiB, filt=True) == want


def test_sim_pe_header_cut_by_the_block_end(simlib):
    """Found by tests/fuzz_sim.py: a PE header near the end of a block.  The reference parses the section table wherever e_lfanew points
    without checking that it ends inside the block, i.e. it reads heap memory behind the block (src/filter_x86_pe.c:75-126).  Oracle and
    CUDA path both define those bytes as zero; stage call and container must agree, whatever follows the block in memory."""
    rng = np.random.default_rng(77)
    code = bytearray(rng.integers(0, 256, 120, dtype=np.uint8).tobytes() * 50)       # compressible: keep "cannot compress" out of this test
    for i in range(0, len(code) - 5, 37):
        code[i] = 0xE8
        code[i + 1:i + 5] = struct.pack("<i", int(rng.integers(-2000, 2000)))
    mz = bytearray(200)
    mz[0:2] = b"MZ"
    mz[0x3C:0x40] = struct.pack("<I", 150)
    mz[150:154] = b"PE\0\0"
    mz[154:156] = struct.pack("<H", 0x14c)          # machine
    mz[156:158] = struct.pack("<H", 2)              # two sections: the second header lies behind the end of a 300-byte block
    mz[170:172] = struct.pack("<H", 0)              # optional header size
    mz[190:194] = struct.pack("<I", 5000)           # SizeOfRawData of section 0 (inside the block)
    data = b"p" * 100 + bytes(mz) + bytes(code)
    # second shape: the section table itself ENDS behind the block (size_hdr = 24 + 100 + 2 * 40 = 204 > 200 bytes left): the reference's
    # `len - size_hdr` wraps and it rewrites heap memory behind the block; here nothing is transformed, the flag is still raised
    mz2 = bytearray(mz)
    mz2[170:172] = struct.pack("<H", 100)
    tail = b"p" * 100 + bytes(mz2)
    orc = O.Oracle(api.ROLZ)
    with api.Handle(api.ROLZ, lib=simlib) as h:
        assert h.filter_inplace(tail, 0) == orc.filter_inplace(tail, 0) == (1, tail)
        assert h.filter_inplace(bytes(code), 0) == orc.filter_inplace(bytes(code), 0)         # and the state both are left in agrees
    for bs in (300, 4096):
        want = O.compress(data, api.ROLZ, bs, filt=1)
        with api.Handle(api.ROLZ, lib=simlib) as h:
            assert h.compress(data, bs, filt=True) == want, bs
        orc = O.Oracle(api.ROLZ)
        with api.Handle(api.ROLZ, lib=simlib) as h:
            for i in range(0, len(data), bs):
                assert h.filter_inplace(data[i:i + bs], 0) == orc.filter_inplace(data[i:i + bs], 0), (bs, i)
