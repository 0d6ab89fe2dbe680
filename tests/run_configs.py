"""Runs the larger BASELINE.json configurations end to end on the GPU and against the unmodified reference CLI
(byte identity + reference-decoder round trip + timings).  Not part of pytest: too large for a unit test."""
import argparse
import hashlib
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import oracle_ffi as O  # noqa: E402
from comprox_b200 import api, synth  # noqa: E402

MiB = 1 << 20
CONFIGS = {
    "text-100M": dict(make=lambda n: synth.markov_text(n, seed=42), n=100 * MiB, variant=api.ROLZ, binary="comprolz", flags=[], filt=False),
    "x86-256M": dict(make=lambda n: synth.x86_corpus(n, seed=43), n=256 * MiB, variant=api.ROLZ, binary="comprolz", flags=["-F"], filt=True),
    "mixed-192M": dict(make=lambda n: synth.mixed_corpus(n, seed=45, segment=64 * MiB), n=192 * MiB, variant=api.ROLZ, binary="comprolz", flags=["-F"], filt=True),
    "mixed-1G": dict(make=lambda n: synth.mixed_corpus(n, seed=45, segment=64 * MiB), n=1024 * MiB, variant=api.ROLZ, binary="comprolz", flags=["-F"], filt=True),
    "mixed-4G": dict(make=lambda n: synth.mixed_corpus(n, seed=45, segment=64 * MiB), n=4096 * MiB, variant=api.ROLZ, binary="comprolz", flags=["-F"], filt=True),
    "bmp-512M": dict(make=lambda n: synth.bmp_corpus(n, seed=44), n=512 * MiB, variant=api.LZP, binary="comprop", flags=["-F"], filt=True),
}
ap = argparse.ArgumentParser()
ap.add_argument("names", nargs="*", default=["text-100M", "x86-256M", "mixed-192M", "bmp-512M"])
ap.add_argument("--scale", type=float, default=1.0, help="shrink the configs (1.0 = BASELINE sizes)")
ap.add_argument("--no-ref", action="store_true")
ap.add_argument("--no-ref-decode", action="store_true", help="compare with the reference compressor only")
ap.add_argument("--full-warmup", action="store_true", help="warm up with the full input so that the timed run does no allocation")
ap.add_argument("--window-check-mb", type=int, default=0, help="also compress with this window size and require the same container")
ap.add_argument("--out", default="gpurun_out/configs.jsonl")
ap.add_argument("--window-mb", type=int, default=0, help="raw bytes per window (0 = library default)")
a = ap.parse_args()
os.makedirs(os.path.dirname(a.out), exist_ok=True)
for name in a.names:
    c = CONFIGS[name]
    n = int(c["n"] * a.scale) // MiB * MiB
    t0 = time.time(); data = c["make"](n); tgen = time.time() - t0
    rec = {"config": name, "bytes": len(data), "gen_s": round(tgen, 1)}
    with api.Handle(c["variant"]) as h:
        h.compress(data if a.full_warmup else data[:4 * MiB], 16 * MiB, filt=c["filt"])   # warm-up (allocations, module load)
        h.profile(True)
        t0 = time.time(); out = h.compress(data, 16 * MiB, filt=c["filt"], window_bytes=a.window_mb * MiB); dt = time.time() - t0
        rec.update(gpu_s=round(dt, 3), gpu_mibs=round(len(data) / MiB / dt, 1), abi_s=round(h.last_call_s, 3), abi_mibs=round(len(data) / MiB / h.last_call_s, 1), container=len(out), sha=hashlib.sha256(out).hexdigest()[:16],
                   stages_ms={k: round(v, 1) for k, v in h.profile_report().items()})
        if a.window_check_mb:
            out2 = h.compress(data, 16 * MiB, filt=c["filt"], window_bytes=a.window_check_mb * MiB)
            rec["same_with_%dM_windows" % a.window_check_mb] = out2 == out
            del out2
    if not a.no_ref and O.ref_binary(c["binary"]):
        t0 = time.time(); ref = O.ref_compress(data, c["binary"], ["-b16", *c["flags"]], tmpdir="/dev/shm"); dr = time.time() - t0
        rec.update(ref_s=round(dr, 1), ref_mibs=round(len(data) / MiB / dr, 2), identical=(ref == out))
        if not a.no_ref_decode:
            t0 = time.time(); back = O.ref_decompress(out, c["binary"], tmpdir="/dev/shm"); dd = time.time() - t0
            rec.update(ref_decode_s=round(dd, 1), roundtrip=(back == data))
    print(json.dumps(rec), flush=True)
    with open(a.out, "a") as f:
        f.write(json.dumps(rec) + "\n")
