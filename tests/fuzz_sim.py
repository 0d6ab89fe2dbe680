"""Differential fuzzer (TEST INFRASTRUCTURE, CPU only): random small inputs through the kernel-logic simulation of the CUDA sources
(tests/sim/libcrgpu_sim.so) against the oracle -- whole containers for all three front-ends with random switches, decode round trips,
and the stage-level calls with random block cuts.  Not collected by pytest (it runs for as long as you let it):

    python tests/fuzz_sim.py --seconds 300 [--seed N]

Every failing case is written to gpurun_out/fuzz_fail_<seed>.bin with its parameters on stdout.  So far (~7000 cases): no difference in
the CUDA sources; one finding in the reference itself -- a PE header near the end of a block makes it parse heap memory behind the
block (undefined behaviour, restated too faithfully by the oracle, which crashed): both sides now define those bytes as zero."""
import argparse
import json
import os
import struct
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import oracle_ffi as O  # noqa: E402
from comprox_b200 import api, synth  # noqa: E402


def bmp(rng, w, h, bpp=24, broken=False):
    row = (bpp * w + 31) // 32 * 4
    pix = (np.cumsum(rng.integers(0, 4, row * h, dtype=np.uint8).astype(np.uint32)) & 255).astype(np.uint8).tobytes()
    size = 54 + row * h
    hdr = struct.pack("<2sIHHIIiiHHIIiiII", b"BM", size, 0, 0, 54, 40, w, h if rng.random() < 0.8 else -h, 1, bpp, 0, row * h if rng.random() < 0.7 else 0, 2835, 2835, 0, 0)
    out = hdr + pix
    if broken:
        out = out[:int(rng.integers(1, len(out)))]
    return out


def words(rng, n):
    vocab = [bytes(rng.integers(97, 123, int(rng.integers(2, 12)), dtype=np.uint8)) for _ in range(int(rng.integers(5, 400)))]
    seps = [b" ", b", ", b". ", b"; ", b": ", b".\n", b"  ", b"-"]
    out = bytearray()
    while len(out) < n:
        w = vocab[int(rng.integers(0, len(vocab)))]
        if rng.random() < 0.1:
            w = w.capitalize()
        if rng.random() < 0.03:
            w = w.upper()
        out += w + seps[int(rng.integers(0, len(seps)))]
    return bytes(out[:n])


def piece(rng, n):
    k = int(rng.integers(0, 9))
    if k == 0:
        return rng.integers(0, 256, n, dtype=np.uint8).tobytes()
    if k == 1:
        return rng.integers(0, int(rng.integers(2, 20)), n, dtype=np.uint8).tobytes()
    if k == 2:
        p = rng.integers(0, 256, int(rng.integers(1, 40)), dtype=np.uint8).tobytes()
        return (p * (n // len(p) + 1))[:n]
    if k == 3:
        return words(rng, n)
    if k == 4:
        return synth.markov_text(max(n, 64), seed=int(rng.integers(0, 1 << 30)))[:n]
    if k == 5:
        return bmp(rng, int(rng.integers(4, 90)), int(rng.integers(4, 60)), 24 if rng.random() < 0.7 else 32, broken=rng.random() < 0.3)
    if k == 6:
        img = synth.pe_image(rng, max(n, 4096)) if rng.random() < 0.6 else synth.elf_image(rng, max(n, 4096))
        return img[:int(rng.integers(64, len(img)))] if rng.random() < 0.3 else img
    if k == 7:
        return bytes(n)
    return rng.integers(0, 256, n, dtype=np.uint8).tobytes()[:int(rng.integers(0, 8))] + b"MZ" + b"BM" + b"\x7fELF" + b"\xe8\xe9" * 5


def make_input(rng):
    parts = [piece(rng, int(rng.integers(1, 60000))) for _ in range(int(rng.integers(1, 6)))]
    data = b"".join(parts)
    if rng.random() < 0.1:
        data = data[:int(rng.integers(0, 40))]
    return data


def one_case(rng, sim, idx):
    data = make_input(rng)
    variant = int(rng.integers(0, 3))
    bs = int(rng.choice([4096, 20000, 65536, 100003, 1 << 20]))
    filt = int(rng.random() < 0.5)
    prec = int(rng.random() < 0.1)
    flex = int(variant != api.LZP and rng.random() < 0.25)
    ml = int(rng.integers(1, 60)) if variant == api.LZ77 and rng.random() < 0.3 else 0
    params = dict(variant=variant, bs=bs, filt=filt, prec=prec, flex=flex, match_limit=ml, n=len(data))
    kw = dict(block_size=bs, filt=filt, prec=prec, flexible=flex)
    want = O.compress(data, variant, match_limit=ml, **kw)
    mode = int(rng.integers(0, 3))
    try:
        with api.Handle(variant, lib=sim) as h:
            if ml:
                h.set_option("match_limit", ml)
            got = h.compress(data, bs, filt=bool(filt), prec=bool(prec), flexible=bool(flex), window_bytes=int(rng.choice([0, bs, 3 * bs])))
    except api.CrgpuError as e:
        return "error %d" % e.code, params        # includes -6: mid-chain "cannot compress" blocks are replayed exactly (LzChain::encode_blocks)
    if got != want:
        return "container differs", params
    if mode == 0 and len(data) <= bs:
        # decode our own container (single-block inputs only: after a stored block in mid-chain the reference's decoder -- and the
        # oracle's -- reads garbage); with -F the reference's decoder is only right for stored blocks (F4), so compare
        # with the oracle's decode
        try:
            back_o = O.decompress(want, variant)
        except Exception:
            back_o = None
        with api.Handle(variant, lib=sim) as h:
            back = h.decompress(got, len(data) + 64)
        if back_o is not None and back != back_o:
            return "decode differs from the oracle's", params
        if not filt and back != data:
            return "round trip", params
    if mode == 1 and filt:
        # stage-level filter calls with the same block cuts
        orc = O.Oracle(variant)
        with api.Handle(variant, lib=sim) as h:
            for i in range(0, len(data), bs):
                blk = data[i:i + bs]
                if h.filter_inplace(blk, 0) != orc.filter_inplace(blk, 0):
                    return "filter_inplace block %d" % (i // bs), params
    if mode == 2 and not prec:
        # stage level, one call per block as host/cr_shim.c issues them: dictionary stage (no dictionary -> stored or escaped), lzencode with
        # chain_ends = 0, then the decoder mirror as long as no block of the chain was stored (after one the reference's decoder is lost, F11)
        orc = O.Oracle(variant)
        text = O.dicpick(data)
        orc.dictionary_load(text, 1)
        orc.reset_models()
        with api.Handle(variant, lib=sim) as h, api.Handle(variant, lib=sim) as d:
            if flex:
                h.set_option("flexible", 1)
            h.dictionary_load(text, 1); d.dictionary_load(text, 0)
            h.reset_models(); d.reset_models()
            if flex:
                O.lib().cro_set_flexible(orc.c, 1)
            if ml:
                h.set_option("match_limit", ml); O.lib().cro_set_match_limit(orc.c, ml)
            lost = False
            for i in range(0, len(data), bs):
                blk = data[i:i + bs]
                want_d = orc.dictionary_encode(blk)
                got_d = h.dictionary_encode(blk)
                if got_d != want_d:
                    return "dictionary_encode block %d" % (i // bs), params
                if d.dictionary_decode(got_d, len(blk) + 64) != blk:
                    return "dictionary_decode block %d" % (i // bs), params
                want_p = orc.lzencode(want_d)
                got_p = h.lzencode([got_d], chain_ends=False)[0]
                if got_p != want_p:
                    return "lzencode block %d" % (i // bs), params
                stored = not (got_p[1] if variant == api.ROLZ else got_p[0])
                if not lost and d.lzdecode(got_p) != got_d:
                    return "lzdecode block %d" % (i // bs), params
                lost = lost or (stored and not (variant == api.LZP and len(got_d) < 16))
    return None, params


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=120)
    ap.add_argument("--seed", type=int, default=int(time.time()))
    a = ap.parse_args()
    subprocess.run(["make", "-C", os.path.join(ROOT, "comprox_b200", "csrc"), "sim"], check=True, capture_output=True)
    sim = api.load(os.path.join(ROOT, "tests", "sim", "libcrgpu_sim.so"))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    t0, n, fails, aborts = time.time(), 0, 0, 0
    while time.time() - t0 < a.seconds:
        seed = a.seed * 1000003 + n
        rng = np.random.default_rng(seed)
        res, params = one_case(rng, sim, n)
        if res == "abort":
            aborts += 1
        elif res:
            fails += 1
            rng = np.random.default_rng(seed)
            data = make_input(rng)
            path = os.path.join(ROOT, "gpurun_out", "fuzz_fail_%d.bin" % seed)
            open(path, "wb").write(data)
            print(json.dumps({"fail": res, "seed": seed, "file": path, **params}), flush=True)
        n += 1
    print(json.dumps({"cases": n, "fails": fails, "midchain_aborts": aborts, "seconds": round(time.time() - t0, 1), "seed": a.seed}))


if __name__ == "__main__":
    main()
