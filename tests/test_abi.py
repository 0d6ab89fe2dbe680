"""No-GPU checks of the drop-in boundary: the shared library loads, exports every symbol that include/crgpu.h
declares, and refuses to work without a CUDA device (there is no CPU path to fall back to)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "crgpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(crgpu_[a-z_0-9]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    so = os.path.join(ROOT, "comprox_b200", "libcrgpu.so")
    if not os.path.exists(so):
        subprocess.run(["make", "-C", os.path.join(ROOT, "comprox_b200", "csrc")], check=True, capture_output=True)
    return ctypes.CDLL(so)


def test_header_declares_the_expected_entry_points():
    syms = _declared_symbols()
    for s in ("crgpu_create", "crgpu_destroy", "crgpu_reset_models", "crgpu_lzencode", "crgpu_compress", "crgpu_compress_bound"):
        assert s in syms


def test_library_exports_every_declared_symbol(lib):
    for s in _declared_symbols():
        assert hasattr(lib, s), "libcrgpu.so does not export " + s


def test_no_cpu_fallback(lib):
    """Without a device crgpu_create must fail with CRGPU_ERR_NO_DEVICE; with one this test is skipped."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    h = ctypes.c_void_p()
    rc = lib.crgpu_create(ctypes.byref(h), 0, 0, None)
    assert rc == -1 and not h.value
    lib.crgpu_strerror.restype = ctypes.c_char_p
    assert b"no CPU" in lib.crgpu_strerror(rc)


def test_product_library_does_not_link_the_oracle(lib):
    out = subprocess.run(["nm", "-D", os.path.join(ROOT, "comprox_b200", "libcrgpu.so")], capture_output=True, text=True).stdout
    assert "cro_" not in out
    d = os.path.join(ROOT, "comprox_b200", "csrc")
    src = "".join(open(os.path.join(d, f)).read() for f in os.listdir(d) if os.path.isfile(os.path.join(d, f)))
    assert "cr_oracle" not in src and "liboracle" not in src


@pytest.mark.parametrize("variant,extra", [("rolz", ["flexible_parsing"]), ("rop", []), ("rox", ["flexible_parsing", "match_limit"])])
def test_shim_archives_define_the_reference_symbols(variant, extra):
    """bin/libcrshim_*.a (comprox_b200/host/cr_shim.c) must define exactly what the reference's src/main.c and src/<variant>/main.c link
    against (/root/reference/src/main.c:33-59, src/cr-datablock.h:43-46)."""
    subprocess.run(["make", "-C", os.path.join(ROOT, "comprox_b200", "host")], check=True, capture_output=True)
    out = subprocess.run(["nm", "--defined-only", os.path.join(ROOT, "bin", "libcrshim_%s.a" % variant)], capture_output=True, text=True).stdout
    defined = {ln.split()[-1] for ln in out.splitlines() if len(ln.split()) == 3 and ln.split()[1] in "TDB"}
    want = ["data_block_reserve", "data_block_resize", "data_block_add", "data_block_destroy", "filter_inplace", "dicpick", "dic_lcp_encode",
            "dic_lcp_decode", "dictionary_load", "dictionary_encode", "dictionary_decode", "reset_models", "lzencode", "lzdecode", *extra]
    for s in want:
        assert s in defined, "libcrshim_%s.a does not define %s" % (variant, s)
    undefined = subprocess.run(["nm", "-u", os.path.join(ROOT, "bin", "libcrshim_%s.a" % variant)], capture_output=True, text=True).stdout
    assert "cro_" not in undefined and "cuda" not in undefined.lower()          # the shim reaches CUDA only through dlopen(libcrgpu.so)
