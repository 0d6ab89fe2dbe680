"""Decoder robustness fuzzer (kernel-logic simulation by default, --gpu for the CUDA library): valid containers of all three
front-ends are damaged (bit flips, byte runs, truncation, header fields) and handed to crgpu_decompress / crgpu_decompress_batch.
The decoder must return CRGPU_OK (with any bytes) or an error code -- never touch memory outside its buffers (on the simulation a
stray access is a crash of this process, under compute-sanitizer on the GPU a reported error).  Not collected by pytest;
tests/test_sim_decode_corrupt.py runs a bounded number of trials in a subprocess."""
import argparse
import ctypes
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import oracle_ffi as O  # noqa: E402
from comprox_b200 import api, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--trials", type=int, default=200)
ap.add_argument("--seed", type=int, default=1)
ap.add_argument("--gpu", action="store_true")
ap.add_argument("--batch", action="store_true", help="also decode damaged and intact containers together through crgpu_decompress_batch")
a = ap.parse_args()
rnd = random.Random(a.seed)
L = api.load() if a.gpu else api.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "sim", "libcrgpu_sim.so"))

KiB = 1 << 10
inputs = {
    "text": (synth.markov_text(96 * KiB, seed=3), 0),
    "x86": (synth.x86_corpus(128 * KiB, elf_bytes=40 * KiB, pe_min=16 * KiB, pe_max=40 * KiB), 1),
    "bmp": (synth.bmp_corpus(96 * KiB, wmin=60, wmax=200, hmin=20, hmax=100), 1),
    "mixed": (synth.markov_text(40 * KiB, seed=4) + bytes(rnd.getrandbits(8) for _ in range(20 * KiB)) + synth.markov_text(30 * KiB, seed=5), 0),
}
valid = []
for variant in (api.ROLZ, api.LZP, api.LZ77):
    for name, (data, filt) in inputs.items():
        for bs in (32 * KiB, 1 << 20):
            valid.append((variant, name, data, O.compress(data, variant, bs, filt)))


def damage(c):
    c = bytearray(c)
    kind = rnd.randrange(7)
    if kind == 0:                                       # single bit flips
        for _ in range(rnd.randrange(1, 4)):
            p = rnd.randrange(len(c)); c[p] ^= 1 << rnd.randrange(8)
    elif kind == 1:                                     # a run of random bytes
        p = rnd.randrange(len(c)); n = rnd.randrange(1, 64)
        c[p:p + n] = bytes(rnd.getrandbits(8) for _ in range(min(n, len(c) - p)))
    elif kind == 2:                                     # truncation
        c = c[:rnd.randrange(1, len(c))]
    elif kind == 3:                                     # damage close behind the magic: dictionary length, first block header, inner header
        p = rnd.randrange(20, min(len(c), 90)); c[p] = rnd.getrandbits(8)
    elif kind == 4:                                     # a 32-bit field set to an extreme value
        p = rnd.randrange(20, len(c) - 4); c[p:p + 4] = rnd.choice([b"\xff\xff\xff\xff", b"\x00\x00\x00\x00", b"\x00\x00\x00\x80", b"\xff\xff\xff\x7f"])
    elif kind == 5:                                     # bytes dropped from the middle
        p = rnd.randrange(len(c)); del c[p:p + rnd.randrange(1, 32)]
    else:                                               # bytes inserted
        p = rnd.randrange(len(c)); c[p:p] = bytes(rnd.getrandbits(8) for _ in range(rnd.randrange(1, 32)))
    return bytes(c)


handles = {v: api.Handle(v, lib=L) for v in (api.ROLZ, api.LZP, api.LZ77)}
# containers with a stored block in mid-chain (the "mixed" input at small block sizes) are byte-identical to the reference's but not
# decodable by any decoder: they must be REFUSED cleanly, and are left out of the "intact containers decode" checks
decodable = []
for variant, name, data, cont in valid:
    try:
        decodable.append(handles[variant].decompress(cont, len(data) + 64) == data)
    except api.CrgpuError as e:
        assert e.code == -9 and name == "mixed", (name, e)
        decodable.append(False)
assert sum(decodable) >= len(valid) - 6
valid = [v + (d,) for v, d in zip(valid, decodable)]
stats = {"ok_same": 0, "ok_other": 0, "error": 0}
for t in range(a.trials):
    variant, name, data, cont, good = rnd.choice(valid)
    bad = damage(cont)
    cap = len(data) * 4 + 4096
    out = ctypes.create_string_buffer(cap); n = ctypes.c_uint64(0)
    rc = L.crgpu_decompress(handles[variant].h, bad, ctypes.c_uint64(len(bad)), out, ctypes.c_uint64(cap), ctypes.byref(n))
    if rc == 0:
        stats["ok_same" if out.raw[:n.value] == data else "ok_other"] += 1
    else:
        assert rc < 0, rc
        stats["error"] += 1
    if t % 10 == 0 and good:                            # the handle must still decode an intact container afterwards
        assert handles[variant].decompress(cont, len(data) + 64) == data, "handle unusable after a damaged container (trial %d)" % t
if a.batch:
    for t in range(max(1, a.trials // 20)):
        variant = rnd.choice((api.ROLZ, api.LZP, api.LZ77))
        pool = [v for v in valid if v[0] == variant and v[4]]
        picks = [rnd.choice(pool) for _ in range(4)]
        conts = [damage(p[3]) if k % 2 else p[3] for k, p in enumerate(picks)]
        hs = [api.Handle(variant, lib=L) for _ in picks]
        try:
            k = len(hs)
            outs = [ctypes.create_string_buffer(len(p[2]) * 4 + 4096) for p in picks]
            lens = (ctypes.c_uint64 * k)()
            rc = L.crgpu_decompress_batch((ctypes.c_void_p * k)(*[h.h for h in hs]), ctypes.c_uint32(k), (ctypes.c_char_p * k)(*conts),
                                          (ctypes.c_uint64 * k)(*[len(c) for c in conts]), (ctypes.c_void_p * k)(*[ctypes.addressof(o) for o in outs]),
                                          (ctypes.c_uint64 * k)(*[len(o) for o in outs]), lens)
            for i in (0, 2):                            # the intact containers of the batch are decoded whatever happened to the others
                assert outs[i].raw[:lens[i]] == picks[i][2], "intact container %d of a batch with damaged ones was not decoded (rc %d)" % (i, rc)
        finally:
            for h in hs:
                h.close()
for h in handles.values():
    h.close()
print("fuzz_decode:", a.trials, "trials", stats)
