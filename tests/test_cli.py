"""The C front-ends (comprox_b200/host/cr_main.c -> bin/comprolz, bin/comprop, bin/comprox): same switches and exit
behaviour as the reference command lines.  On CPU they are pointed at the kernel-logic simulation through CRGPU_LIB (test
infrastructure); on the GPU box they load the real libcrgpu.so and are compared with the reference CLI."""
import os
import subprocess

import pytest

import oracle_ffi as O
from comprox_b200 import api, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MiB = 1 << 20
VARIANT = {"comprolz": api.ROLZ, "comprop": api.LZP, "comprox": api.LZ77}


def _cli(binary, args, env_lib, tmp_path, data):
    exe = os.path.join(ROOT, "bin", binary)
    if not os.path.exists(exe):
        subprocess.run(["make", "-C", os.path.join(ROOT, "comprox_b200", "host")], check=True, capture_output=True)
    src, dst = tmp_path / "in.bin", tmp_path / "out.bin"
    src.write_bytes(data)
    env = dict(os.environ)
    if env_lib:
        env["CRGPU_LIB"] = env_lib
    r = subprocess.run([exe, *args, str(src), str(dst)], env=env, capture_output=True)
    return r, (dst.read_bytes() if dst.exists() else None)


@pytest.mark.parametrize("binary,flags,kw", [
    ("comprolz", ["-q", "-b1"], dict(block_size=MiB)),
    ("comprolz", ["-q", "-b1", "-f"], dict(block_size=MiB, flexible=1)),
    ("comprop", ["-q", "-b1", "-p"], dict(block_size=MiB, prec=1)),
    ("comprox", ["-q", "-b1", "-m8"], dict(block_size=MiB, match_limit=8)),
    ("comprox", ["-q", "-F"], dict(filt=1)),
])
def test_cli_on_simulation_matches_oracle(simlib, tmp_path, binary, flags, kw):
    data = synth.markov_text(MiB + 999, seed=21)
    sim = os.path.join(ROOT, "tests", "sim", "libcrgpu_sim.so")
    r, out = _cli(binary, [*flags, "e"], sim, tmp_path, data)
    assert r.returncode == 0, r.stderr
    assert out == O.compress(data, VARIANT[binary], **kw)


def test_cli_rejects_bad_switches(simlib, tmp_path):
    sim = os.path.join(ROOT, "tests", "sim", "libcrgpu_sim.so")
    for binary, flags in [("comprop", ["-f"]), ("comprolz", ["-m8"]), ("comprox", ["-m0"]), ("comprolz", ["-b0"]), ("comprolz", ["-x"])]:
        r, out = _cli(binary, [*flags, "e"], sim, tmp_path, b"abc")
        assert r.returncode != 0 and b"invalid switch" in r.stderr, (binary, flags)


def test_cli_fails_loudly_without_the_cuda_library(tmp_path):
    r, out = _cli("comprolz", ["-q", "e"], "/nonexistent/libcrgpu.so", tmp_path, b"abc")
    assert r.returncode != 0 and b"no CPU fallback" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("binary,flags", [("comprolz", ["-b1"]), ("comprop", ["-b1", "-F"]), ("comprox", ["-b1"]), ("comprox", ["-b1", "-f", "-m16"])])
def test_cli_on_gpu_matches_reference_cli(gpulib, tmp_path, binary, flags):
    data = synth.markov_text(3 * MiB + 17, seed=22)
    if "-F" in flags:                # unfiltered image noise would hit "cannot compress" mid-chain (undecodable in the reference too)
        data += synth.bmp_corpus(MiB, wmin=201, wmax=500, hmin=60, hmax=300)
    r, out = _cli(binary, ["-q", *flags, "e"], None, tmp_path, data)
    assert r.returncode == 0, r.stderr
    ref = O.ref_compress(data, binary, flags)
    if ref is not None:
        assert out == ref
    (tmp_path / "c.bin").write_bytes(out)
    exe = os.path.join(ROOT, "bin", binary)
    r2 = subprocess.run([exe, "-q", "d", str(tmp_path / "c.bin"), str(tmp_path / "back.bin")], capture_output=True)
    assert r2.returncode == 0, r2.stderr
    if "-F" not in flags:            # filtered + dictionary-compressible blocks do not round-trip in the reference either
        assert (tmp_path / "back.bin").read_bytes() == data


@pytest.mark.parametrize("binary", ["comprolz", "comprop", "comprox"])
def test_cli_decode_on_simulation(simlib, tmp_path, binary):
    """`d` branch: the container does not store its decoded size, so the front-end's first buffer (4 x container + 1 MiB) is too small
    for well-compressible input; it must come back with the size needed and finish on the second call."""
    data = (b"the same line again and again. " * 40000)[:1200000] + synth.markov_text(100000, seed=23)
    cont = O.compress(data, VARIANT[binary], MiB)
    assert len(cont) * 4 + MiB < len(data)
    sim = os.path.join(ROOT, "tests", "sim", "libcrgpu_sim.so")
    r, out = _cli(binary, ["-q", "d"], sim, tmp_path, cont)
    assert r.returncode == 0, r.stderr
    assert out == data
