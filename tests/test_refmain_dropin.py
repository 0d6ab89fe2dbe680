"""The literal drop-in: the UNMODIFIED reference driver -- /root/reference/src/main.c + src/<variant>/main.c,
compiled by oracle/Makefile into oracle/_ref/refmain_* -- linked against comprox_b200/host/cr_shim.c, which provides the reference's
own symbols (filter_inplace, dicpick, dic_lcp_*, dictionary_*, reset_models, lzencode, lzdecode, data_block_*) over the C ABI of
libcrgpu.so.  Its containers must equal the reference CLI's byte for byte and decode back, block loop and all.  On CPU the shim is
pointed at the kernel-logic simulation (pre-flight), on the GPU box at the real library.  The refmain_* binaries are built where
/root/reference exists and travel with oracle/_ref/; without them the tests skip."""
import os
import subprocess

import pytest

import oracle_ffi as O
from comprox_b200 import api, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MiB = 1 << 20
SIM = os.path.join(ROOT, "tests", "sim", "libcrgpu_sim.so")
VARIANT = {"comprolz": api.ROLZ, "comprop": api.LZP, "comprox": api.LZ77}
BACKENDS = [pytest.param("sim", id="sim"), pytest.param("gpu", id="gpu", marks=pytest.mark.gpu)]


def _run(binary, args, lib, tmp_path, data, tag):
    exe = os.path.join(O.REF_DIR, "refmain_" + binary)
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/refmain_%s not built (needs /root/reference)" % binary)
    src, dst = tmp_path / (tag + ".in"), tmp_path / (tag + ".out")
    src.write_bytes(data)
    env = dict(os.environ, CRGPU_LIB=lib)
    r = subprocess.run([exe, *args, str(src), str(dst)], env=env, capture_output=True)
    assert r.returncode == 0, r.stderr[-2000:]
    return dst.read_bytes()


def _lib(request, which):
    request.getfixturevalue("simlib" if which == "sim" else "gpulib")
    return SIM if which == "sim" else api.LIB_PATH


@pytest.mark.parametrize("which", BACKENDS)
@pytest.mark.parametrize("binary,flags,kw", [
    ("comprolz", ["-b1"], dict(block_size=MiB)),
    ("comprolz", ["-b1", "-f"], dict(block_size=MiB, flexible=1)),
    ("comprop", ["-b1"], dict(block_size=MiB)),
    ("comprox", ["-b1", "-m12"], dict(block_size=MiB, match_limit=12)),
    ("comprop", ["-b1", "-p"], dict(block_size=MiB, prec=1)),
])
def test_reference_driver_over_shims_text(request, tmp_path, which, binary, flags, kw):
    if which == "sim" and (flags == ["-b1", "-f"] or (binary, flags) == ("comprop", ["-b1"])):
        pytest.skip("CPU pre-flight runs three of the five switch sets (time); the GPU run takes all")
    lib = _lib(request, which)
    data = synth.markov_text(2 * MiB + 4321, seed=51)
    want = O.compress(data, VARIANT[binary], **kw)
    ref = O.ref_compress(data, binary, flags)
    if ref is not None:
        assert ref == want
    got = _run(binary, ["-q", *flags, "e"], lib, tmp_path, data, "e")
    assert got == want
    assert _run(binary, ["-q", "d"], lib, tmp_path, got, "d") == data


@pytest.mark.parametrize("which", BACKENDS)
def test_reference_driver_over_shims_filters_and_empty_tail(request, tmp_path, which):
    """-F on BMP data (stored by the dictionary stage, so the reference decoder can inverse-filter it) with a size that is a
    multiple of the block size: the driver's loop then emits the trailing empty block (F8)."""
    lib = _lib(request, which)
    data = synth.bmp_corpus(2 * MiB, seed=9, wmin=200, wmax=500, hmin=60, hmax=200)
    assert len(data) == 2 * MiB
    want = O.compress(data, api.LZP, MiB, filt=1)
    got = _run("comprop", ["-q", "-b1", "-F", "e"], lib, tmp_path, data, "e")
    assert got == want
    assert _run("comprop", ["-q", "d"], lib, tmp_path, got, "d") == data
    back = O.ref_decompress(got, "comprop")
    if back is not None:
        assert back == data


def test_shim_fails_loudly_without_the_cuda_library(tmp_path):
    exe = os.path.join(O.REF_DIR, "refmain_comprolz")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/refmain_comprolz not built")
    src = tmp_path / "a"
    src.write_bytes(b"hello hello hello")
    r = subprocess.run([exe, "-q", "e", str(src), str(tmp_path / "b")], env=dict(os.environ, CRGPU_LIB="/nonexistent/libcrgpu.so"), capture_output=True)
    assert r.returncode != 0 and b"no CPU fallback" in r.stderr
