#!/usr/bin/env python
"""bench.py -- headline benchmark of comprox-b200 (contract: see the task description / DESIGN.md section 7).

Workload of `value` (BASELINE.json configs[1]): `comprolz` default settings (ROLZ, 16 MiB blocks) on a 100 MiB synthetic
word-level-Markov English-like text file, one container per GPU.  A "step" is one whole-container compression.
  value : MiB/s with the input already resident in HBM (crgpu_stage_input) when the timed region starts
  e2e   : MiB/s through the public C ABI with pinned HOST buffers (H2D of the input and D2H of the container inside)
  parity: the first 32 MiB through the GPU and through the UNMODIFIED reference CLI: sha256 of both containers (N = 1)
N > 1 (torchrun): every rank compresses its own 100 MiB container (independent shards, no data-path collective;
SURVEY.md section 8e) -- weak scaling; time = max over ranks, per-rank times listed.

Further legs on the same JSON line:
  corpus     : BASELINE.json configs[4] -- a 4 GiB mixed corpus (text / x86 / BMP) pre-split into 64 MiB shards = independent
               containers (`comprolz -b16 -F` each), dealt round-robin to the N GPUs (STRONG scaling), several handles per GPU
               (crgpu_compress_batch), the containers delivered IN SHARD ORDER into one host buffer inside the timed region;
               sha256 over the ordered containers is compared with the digest of the N = 1 run (tests/golden/corpus_digest.json)
               and with the digests the reference arm left behind; plus the block-size sweep -b1 / -b4 / -b16 / -b64.
  decompress : many containers in flight through crgpu_decompress_batch (one container = one serial model chain), swept over the
               number in flight, with (C + B) / t against the HBM peak.
--impl reference : the UNMODIFIED reference CLI (oracle/_ref/comprolz, built by oracle/Makefile) on the host cores: `value` = N
               processes side by side on the text sample (one container cannot use more than one core);
               `corpus` = one process per host core over shards of the same corpus (xargs -P nproc); `decompress` = `comprolz d`.
"""
import argparse
import hashlib
import json
import os
import resource
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# handles that share a GPU need their own hardware work queues (crgpu_api.cu: cr_more_work_queues); the variable only counts before the
# CUDA context exists, and torch creates that before libcrgpu.so is loaded
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

MiB = 1 << 20
WORKLOAD_BYTES = 100 * MiB
BLOCK = 16 * MiB
REF_SAMPLE = 32 * MiB          # bounded sample for the CPU arms (2 blocks; models still chain across them)
CORPUS_BYTES = 4096 * MiB      # configs[4]
SHARD_BYTES = int(os.environ.get("CRBENCH_SHARD_MIB", "64")) * MiB      # (the override exists for dry runs on small machines)
SHARD_HANDLES = 8              # handles (private streams + host threads) per GPU in the corpus leg
SWEEP_SHARDS = 12              # shards of the block-size sweep (4 of each kind)
DEC_BYTES = 1 * MiB            # container size of the decompress leg
REF_SHA_FILE = os.path.join(tempfile.gettempdir(), "crbench_reference_corpus_sha.json")


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if not self.proc:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in open(self.path).read().splitlines():
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if sm:
            sm.sort()
            out = {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}
        return out


def host_cores():
    n = os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except AttributeError:
        pass
    return max(1, n)


def shm_dir(need=0):
    """/dev/shm when it has room for `need` bytes (plus slack), else the temp directory."""
    try:
        if os.path.isdir("/dev/shm"):
            st = os.statvfs("/dev/shm")
            if st.f_bavail * st.f_frsize > need + (256 << 20):
                return "/dev/shm"
    except OSError:
        pass
    return tempfile.gettempdir()


def sha(b):
    return hashlib.sha256(b).hexdigest()


# ------------------------------------------------------------------ the mixed corpus (configs[4]), one shard at a time
def corpus_shard(i, nbytes=SHARD_BYTES):
    """Shard i of the mixed corpus: text / x86 / BMP in turn (synth.mixed_corpus's segments; every shard is its own container,
    so every x86 shard may start with its one ELF image)."""
    from comprox_b200 import synth
    kind, seed = i % 3, 45 + 100 * i
    if kind == 0:
        return synth.markov_text(nbytes, seed)
    if kind == 1:
        return synth.x86_corpus(nbytes, seed)
    return synth.bmp_corpus(nbytes, seed)


def _gen_into(args):
    path, slot, i, nbytes = args
    import numpy as np
    data = corpus_shard(i, nbytes)
    m = np.memmap(path, dtype=np.uint8, mode="r+", offset=slot * nbytes, shape=(nbytes,))
    m[:] = np.frombuffer(data, dtype=np.uint8)
    m.flush()
    return i


def generate_shards(indices, nbytes, workers):
    """The shards `indices` back to back in a /dev/shm file, generated by `workers` processes.  Returns the path."""
    import multiprocessing as mp
    path = os.path.join(shm_dir(len(indices) * nbytes), "crbench_corpus_%d.bin" % os.getpid())
    with open(path, "wb") as f:
        f.truncate(len(indices) * nbytes)
    jobs = [(path, slot, i, nbytes) for slot, i in enumerate(indices)]
    if workers <= 1 or len(jobs) <= 1:
        for j in jobs:
            _gen_into(j)
    else:
        with mp.get_context("spawn").Pool(min(workers, len(jobs))) as pool:      # spawn: the parent may hold a CUDA context
            pool.map(_gen_into, jobs, chunksize=1)
    return path


# ------------------------------------------------------------------ reference arm helpers
def reference_binary(name="comprolz"):
    p = os.path.join(ROOT, "oracle", "_ref", name)
    return p if os.path.exists(p) else None


def run_reference_procs(files, flags, mode="e"):
    """One unmodified reference CLI process per file, all side by side; returns (wall seconds, cpu seconds, output paths)."""
    exe = reference_binary()
    outs = [f + (".out" if mode == "e" else ".dec") for f in files]
    r0 = resource.getrusage(resource.RUSAGE_CHILDREN)
    t0 = time.perf_counter()
    ps = [subprocess.Popen([exe, "-q", *flags, mode, f, o]) for f, o in zip(files, outs)]
    ok = all(p.wait() == 0 for p in ps)
    dt = time.perf_counter() - t0
    r1 = resource.getrusage(resource.RUSAGE_CHILDREN)
    if not ok:
        raise RuntimeError("reference CLI failed")
    return dt, (r1.ru_utime - r0.ru_utime) + (r1.ru_stime - r0.ru_stime), outs


def reference_arm(args, rank, config):
    if rank != 0:
        return
    from comprox_b200 import synth
    exe = reference_binary()
    steps, warmup = max(args.steps, 1), min(args.warmup, 1)
    nproc = max(1, min(args.gpus, host_cores()))
    tmp = shm_dir()
    files = []
    line = {"impl": "reference", "metric": "compress_throughput", "unit": "MiB/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": config}
    try:
        # ---- value: N independent containers (the text sample, seed 42 + k) side by side, one process each
        for k in range(nproc):
            p = os.path.join(tmp, "crbench_%d_%d.in" % (os.getpid(), k))
            with open(p, "wb") as f:
                f.write(synth.markov_text(min(REF_SAMPLE, args.bytes), seed=42 + k))
            files.append(p)
        nbytes = os.path.getsize(files[0])
        times, cpus = [], []
        if exe:
            for i in range(warmup + steps):
                dt, cpu, outs = run_reference_procs(files, [])
                if i >= warmup:
                    times.append(dt); cpus.append(cpu)
            kind = "reference"
        else:   # reference binaries did not travel: time the oracle port instead (one process)
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import oracle_ffi as O
            data = open(files[0], "rb").read()
            nproc = 1
            for i in range(warmup + steps):
                t0 = time.perf_counter(); O.compress(data, 0, BLOCK); dt = time.perf_counter() - t0
                if i >= warmup:
                    times.append(dt); cpus.append(dt)
            kind, outs = "port", []
        sec = sum(times) / len(times)
        v = nproc * nbytes / MiB / sec
        cores = max(1, round(sum(cpus) / sum(times)))
        sample = ("%d process(es) side by side, each the first %d MiB (2 blocks) of a text-100M container (seed 42+k), unmodified reference CLI, files in %s"
                  % (nproc, nbytes // MiB, tmp))
        line.update({"value": round(v, 3), "ms_per_step": round(sec * 1e3, 1),
                     "cpu_baseline": {"value": round(v, 3), "unit": "MiB/s", "cores": cores, "kind": kind, "sample": sample,
                                      "note": "one container = one serial model chain: the reference's threads are intra-block helpers (SURVEY.md F1), so a "
                                              "container keeps about one core busy; N containers use N processes"},
                     "e2e": {"value": round(v, 3), "unit": "MiB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        if exe and outs:
            line["parity"] = {"sha_ref": sha(open(outs[0], "rb").read()), "sample": "container of the first %d MiB of the rank-0 workload" % (nbytes // MiB)}
            # ---- decompress: `comprolz d` on the containers just written
            dt, cpu, decs = run_reference_procs(outs, [], mode="d")
            ok = open(decs[0], "rb").read() == open(files[0], "rb").read()
            line["decompress"] = {"value": round(nproc * nbytes / MiB / dt, 3), "unit": "MiB/s", "processes": nproc, "roundtrip_ok": ok,
                                  "sample": "%d reference process(es) decoding the %d MiB text containers side by side" % (nproc, nbytes // MiB)}
            files += outs + decs
        # ---- corpus: all host cores, one process per shard of the mixed corpus (xargs -P nproc over the shard files)
        if exe and not args.no_corpus:
            P = min(host_cores(), 64, CORPUS_BYTES // SHARD_BYTES)
            path = generate_shards(list(range(P)), SHARD_BYTES, host_cores())
            files.append(path)
            shard_files = []
            with open(path, "rb") as f:
                for k in range(P):
                    sp = os.path.join(tmp, "crbench_%d_shard%d" % (os.getpid(), k))
                    with open(sp, "wb") as g:
                        g.write(f.read(SHARD_BYTES))
                    shard_files.append(sp)
            files += shard_files
            dt, cpu, outs = run_reference_procs(shard_files, ["-F"])
            files += outs
            shas = {str(k): sha(open(o, "rb").read()) for k, o in enumerate(outs)}
            with open(REF_SHA_FILE, "w") as f:
                json.dump({"shard_bytes": SHARD_BYTES, "flags": "-b16 -F", "sha": shas}, f)
            cbytes = sum(os.path.getsize(o) for o in outs)
            line["corpus"] = {"value": round(P * SHARD_BYTES / MiB / dt, 2), "unit": "MiB/s", "processes": P, "host_cores": host_cores(),
                              "cores_busy": round(cpu / dt, 1), "seconds": round(dt, 2), "ratio": round(cbytes / (P * SHARD_BYTES), 4),
                              "sample": "shards 0..%d of the 4 GiB mixed corpus (64 MiB each: text / x86 / BMP in turn), `comprolz -b16 -F`, one unmodified "
                                        "reference process per shard, all side by side (xargs -P %d)" % (P - 1, P),
                              "shard_sha_file": REF_SHA_FILE}
            dt, cpu, decs = run_reference_procs(outs, [], mode="d")
            files += decs
            line["corpus"]["decompress"] = {"value": round(P * SHARD_BYTES / MiB / dt, 2), "unit": "MiB/s", "processes": P,
                                            "roundtrip_ok": all(open(d, "rb").read() == open(s, "rb").read() for d, s in zip(decs[:3], shard_files[:3]))}
    finally:
        for p in files:
            if os.path.exists(p):
                os.unlink(p)
    print(json.dumps(line))


# ------------------------------------------------------------------ GPU arm legs
def corpus_leg(args, L, api, torch, dist, rank, local_rank, world, peak):
    """configs[4] in shard mode, strong scaling: see the module docstring."""
    import ctypes
    import numpy as np
    nshards = args.corpus_bytes // SHARD_BYTES
    mine = [i for i in range(nshards) if i % world == rank]
    t_gen = time.perf_counter()
    path = generate_shards(mine, SHARD_BYTES, max(1, host_cores() // world))
    host_in = torch.empty(len(mine) * SHARD_BYTES, dtype=torch.uint8).pin_memory()
    host_in.numpy()[:] = np.memmap(path, dtype=np.uint8, mode="r")
    os.unlink(path)
    t_gen = time.perf_counter() - t_gen
    K = max(1, min(SHARD_HANDLES, len(mine)))
    streams = [torch.cuda.Stream() for _ in range(K)]
    handles = [api.Handle(api.ROLZ, device=local_rank, stream=s.cuda_stream) for s in streams]
    hs = (ctypes.c_void_p * K)(*[h.h for h in handles])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run(block, shard_slots, gather):
        """One pass over the shards in `shard_slots` (indices into `mine`).  Returns (wall s incl. ordered gather, device ms, containers)."""
        m = len(shard_slots)
        cap = int(L.crgpu_compress_bound(ctypes.c_uint64(SHARD_BYTES), ctypes.c_uint32(block)))
        out = torch.empty(max(m, 1) * cap, dtype=torch.uint8).pin_memory()
        cfg = api.Config(block, 1, 0, 0, 0)
        ins = (ctypes.c_void_p * max(m, 1))(*[host_in.data_ptr() + s * SHARD_BYTES for s in shard_slots])
        in_lens = (ctypes.c_uint64 * max(m, 1))(*([SHARD_BYTES] * m))
        outs = (ctypes.c_void_p * max(m, 1))(*[out.data_ptr() + k * cap for k in range(m)])
        caps = (ctypes.c_uint64 * max(m, 1))(*([cap] * m))
        lens = (ctypes.c_uint64 * max(m, 1))()
        barrier()
        e0 = torch.cuda.Event(enable_timing=True); e0.record(streams[0])
        for s in streams[1:]:
            s.wait_event(e0)
        t0 = time.perf_counter()
        rc = L.crgpu_compress_batch(hs, ctypes.c_uint32(K), ctypes.byref(cfg), ctypes.c_uint32(m), ins, in_lens, outs, caps, lens)
        ends = []
        for s in streams:
            e = torch.cuda.Event(enable_timing=True); e.record(s); ends.append(e)
        torch.cuda.synchronize()
        dev_ms = max(e0.elapsed_time(e) for e in ends)
        if rc != 0:
            raise RuntimeError("crgpu_compress_batch failed: %d" % rc)
        sizes = [int(lens[k]) for k in range(m)]
        conts = [out.numpy()[k * cap:k * cap + sizes[k]] for k in range(m)]
        ordered = None
        if gather:
            # ---- ordered delivery: every rank copies its containers to their final offsets in ONE buffer shared by all ranks
            # (a /dev/shm mapping: host memory, so each GPU's containers go device -> pinned host -> final place without crossing
            # another GPU); the only exchange is the table of container sizes
            table = torch.zeros(nshards, dtype=torch.int64, device="cuda")
            for k, s in enumerate(shard_slots):
                table[mine[s]] = sizes[k]
            if world > 1:
                dist.all_reduce(table)
            offs = np.concatenate([[0], np.cumsum(table.cpu().numpy())])
            total = int(offs[-1])
            spath = os.path.join(shm_dir(total), "crbench_gather_%s.bin" % os.environ.get("MASTER_PORT", str(os.getpid())))
            if rank == 0:
                with open(spath, "wb") as f:
                    f.truncate(max(total, 1))
            if world > 1:
                dist.barrier()
            dst = np.memmap(spath, dtype=np.uint8, mode="r+", shape=(max(total, 1),))
            # one memmove per container, on a few host threads (ctypes releases the GIL): 1.3 GB through one core was 0.4 s of a 2.9 s leg
            from concurrent.futures import ThreadPoolExecutor
            def put(k):
                ctypes.memmove(dst.ctypes.data + int(offs[mine[shard_slots[k]]]), conts[k].ctypes.data, sizes[k])
            with ThreadPoolExecutor(max_workers=max(1, min(8, host_cores() // world))) as ex:
                list(ex.map(put, range(m)))
            dst.flush()
            if world > 1:
                dist.barrier()
            ordered = (spath, total, offs, dst)
        wall = time.perf_counter() - t0
        return wall, dev_ms, conts, sizes, ordered

    res = {}
    try:
        slots = list(range(len(mine)))
        run(BLOCK, slots[:K], False)                                   # allocations, first touch
        wall, dev_ms, conts, sizes, ordered = run(BLOCK, slots, True)
        t = torch.tensor([wall, dev_ms / 1e3], device="cuda", dtype=torch.float64)
        per_rank = [t.clone() for _ in range(world)]
        if world > 1:
            dist.all_gather(per_rank, t)
        else:
            per_rank = [t]
        walls = [float(x[0]) for x in per_rank]; devs = [float(x[1]) for x in per_rank]
        spath, total, offs, dst = ordered
        if rank == 0:
            whole = hashlib.sha256(dst[:total]).hexdigest()
            shard_sha = {str(i): sha(bytes(dst[offs[i]:offs[i + 1]])) for i in range(min(nshards, 64))}
            digest = sha("".join(shard_sha[str(i)] for i in range(min(nshards, 64))).encode())
            key = "mixed-%dMiB/shard-%dMiB/-b16 -F" % (args.corpus_bytes // MiB, SHARD_BYTES // MiB)
            golden = None
            try:
                golden = json.load(open(os.path.join(ROOT, "tests", "golden", "corpus_digest.json"))).get(key)
            except Exception:
                pass
            refcheck = None
            try:
                r = json.load(open(REF_SHA_FILE))
                if r.get("shard_bytes") == SHARD_BYTES:
                    common = [k for k in r["sha"] if k in shard_sha]
                    refcheck = {"shards_checked": len(common), "identical": all(r["sha"][k] == shard_sha[k] for k in common)}
            except Exception:
                pass
            nbytes = nshards * SHARD_BYTES
            res = {"value": round(nbytes / MiB / max(walls), 1), "unit": "MiB/s", "scaling": "strong", "bytes": nbytes, "shards": nshards,
                   "shard_bytes": SHARD_BYTES, "handles_per_gpu": K, "flags": "comprolz -b16 -F", "container_bytes": total,
                   "seconds": round(max(walls), 3), "per_rank_s": [round(x, 3) for x in walls], "per_rank_device_s": [round(x, 3) for x in devs],
                   "limiter": "rank %d (%.3f s; fastest %.3f s)" % (walls.index(max(walls)), max(walls), min(walls)),
                   "gather": "inside the timed region: sizes all-reduced, every rank writes its containers at their final offsets of one shared host buffer (shard order)",
                   "sha256_ordered": whole, "digest_of_shard_shas": digest, "digest_key": key,
                   "identical_to_n1": None if golden is None else golden == digest,
                   "parity_vs_reference_arm": refcheck,
                   "pipeline_roofline": {"achieved_gbs": round((nbytes + total) / max(walls) / 1e9, 3), "frac": round((nbytes + total) / max(walls) / 1e9 / peak / world, 6),
                                         "note": "(B + C) / t against N x the measured HBM peak"},
                   "generation_s": round(t_gen, 1)}
        del dst
        if world > 1:
            dist.barrier()
        if rank == 0 and os.path.exists(spath):
            os.unlink(spath)
        # ---- block-size sweep on the first SWEEP_SHARDS shards (whatever rank owns them)
        sweep_slots = [s for s in slots if mine[s] < SWEEP_SHARDS]
        sweep = {}
        for b in (1, 4, 16, 64):
            run(b * MiB, sweep_slots[:1], False)
            wall, dev_ms, conts, sizes, _ = run(b * MiB, sweep_slots, False)
            t = torch.tensor([wall, float(sum(sizes))], device="cuda", dtype=torch.float64)
            if world > 1:
                tm = t.clone(); dist.all_reduce(tm, op=dist.ReduceOp.MAX)
                ts = t.clone(); dist.all_reduce(ts, op=dist.ReduceOp.SUM)
                wall, csum = float(tm[0]), float(ts[1])
            else:
                csum = float(t[1])
            n_sw = min(SWEEP_SHARDS, nshards)
            sweep["-b%d" % b] = {"MiB/s": round(n_sw * SHARD_BYTES / MiB / wall, 1), "ratio": round(csum / (n_sw * SHARD_BYTES), 4)}
        if rank == 0:
            res["block_size_sweep"] = {"shards": min(SWEEP_SHARDS, nshards), "note": "first %d shards (4 of each kind), -F, same handles" % min(SWEEP_SHARDS, nshards), **sweep}
    finally:
        for h in handles:
            h.close()
    return res


def decompress_leg(args, L, api, torch, raw, local_rank, peak):
    """Decode throughput comes from containers in flight (one container = one serial chain): sweep their number."""
    import ctypes
    res = {}
    stream = torch.cuda.current_stream()
    per = DEC_BYTES
    kmax = args.dec_containers
    parts = [raw[(i * 4099 * 251) % (len(raw) - per):][:per] for i in range(kmax)]
    with api.Handle(api.ROLZ, device=local_rank, stream=stream.cuda_stream) as hc:
        conts = [hc.compress(p, BLOCK) for p in parts]
        # one container alone: the per-container latency figure
        hc.decompress(conts[0], per + 64)
        t0 = time.perf_counter(); back = hc.decompress(conts[0], per + 64); dt = time.perf_counter() - t0
        res["single_container"] = {"value": round(per / MiB / dt, 2), "unit": "MiB/s", "roundtrip_ok": back == parts[0]}
    cbytes = sum(len(c) for c in conts)
    hs = [api.Handle(api.ROLZ, device=local_rank, stream=stream.cuda_stream) for _ in range(kmax)]
    try:
        sweep, best = {}, None
        k = min(74, kmax)
        ok = True
        while True:
            api.decompress_batch(hs[:k], conts[:k], [per + 64] * k)                 # tables, first touch
            torch.cuda.synchronize()
            backs = api.decompress_batch(hs[:k], conts[:k], [per + 64] * k); torch.cuda.synchronize(); dt = api.last_batch_call_s
            ok = ok and backs == parts[:k]
            v = k * per / MiB / dt
            sweep[str(k)] = round(v, 1)
            if best is None or v > best[1]:
                best = (k, v, dt, sum(len(c) for c in conts[:k]))
            if k >= kmax:
                break
            k = min(k * 2, kmax)
        k, v, dt, cb = best
        res.update({"value": round(v, 1), "unit": "MiB/s", "containers_in_flight": k, "container_bytes_each": per, "roundtrip_ok": ok,
                    "in_flight_sweep": sweep,
                    "sample": "text containers of %d MiB (`comprolz -b16`) through crgpu_decompress_batch, host buffers in and out, wall clock around the C call" % (per // MiB),
                    "pipeline_roofline": {"achieved_gbs": round((k * per + cb) / dt / 1e9, 4), "frac": round((k * per + cb) / dt / 1e9 / peak, 7), "note": "(C + B) / t"}})
    finally:
        for hh in hs:
            hh.close()
    return res


# algorithmic bytes of a stage per step, from the step's counters; the kernel that dominates the stage
STAGE_MODEL = {
    "range_chain": ("k_rcp_track / k_rcp_seed / k_rcp_emit (cr_rcpar.cuh)", lambda c, n, cb: 12.0 * c["triples"] + cb),
    "o2": ("k_o2_skel / k_o2_eval / k_o2_pass_warp (+ 16-bit radix sort of the events)", lambda c, n, cb: 20.0 * c["events"]),
    "o1": ("k_o1_skel / k_o1_eval / k_o1_pass_warp (+ 64-bit radix sorts of the escapes)", lambda c, n, cb: 20.0 * c["escapes"]),
    "o3": ("k_o3_hot_spec / k_o3_pass_sorted (+ 22-bit radix sort)", lambda c, n, cb: 20.0 * c["events"]),
    "match": ("k_rolz_match_main2 (+ two radix sorts of the positions)", lambda c, n, cb: 0.36 * n + 8.0 * 0.36 * n),
    "diccode": ("k_dc_spans + skip-chain walk", lambda c, n, cb: 1.36 * n),
    "dp_kernels": ("k_dp_count / k_dp_verify", lambda c, n, cb: 2.0 * n),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--bytes", type=int, default=WORKLOAD_BYTES, help="workload size (default: the 100 MiB the metric is quoted on)")
    ap.add_argument("--corpus-bytes", type=int, default=CORPUS_BYTES, help="size of the mixed corpus of the corpus leg (default 4 GiB)")
    ap.add_argument("--dec-containers", type=int, default=1184, help="most containers in flight in the decompress leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-corpus", action="store_true", help="skip the corpus leg")
    ap.add_argument("--no-decompress", action="store_true", help="skip the decompress leg")
    ap.add_argument("--no-shard-leg", action="store_true", help="(older name) skip the corpus leg")
    args = ap.parse_args()
    args.no_corpus = args.no_corpus or args.no_shard_leg
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    steps, warmup = args.steps, max(args.warmup, 3 if args.impl == "ours" else 0)

    from comprox_b200 import synth
    config = {"workload": "text-100M: comprolz (ROLZ) default -b16 on %d bytes of synthetic word-Markov text (seed 42+rank), one container per GPU" % args.bytes,
              "block_size": BLOCK, "containers_per_gpu": 1, "l2": "256 MiB device buffer rewritten between timed steps"}

    if args.impl == "reference":
        reference_arm(args, rank, config)
        return

    # ------------------------------------------------------------------ our arm (GPU)
    import ctypes
    import torch
    import torch.distributed as dist
    from comprox_b200 import api
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    L = api.load()
    L.crgpu_launch_count.restype = ctypes.c_uint64
    L.crgpu_compress_bound.restype = ctypes.c_uint64

    raw = synth.markov_text(args.bytes, seed=42 + rank)
    n = len(raw)
    host_in = torch.empty(n, dtype=torch.uint8).pin_memory()
    host_in.numpy()[:] = memoryview(raw)
    cap = int(L.crgpu_compress_bound(ctypes.c_uint64(n), ctypes.c_uint32(BLOCK)))
    host_out = torch.empty(cap, dtype=torch.uint8).pin_memory()
    flush = torch.empty(256 * MiB, dtype=torch.uint8, device="cuda")
    stream = torch.cuda.current_stream()
    h = api.Handle(api.ROLZ, device=local_rank, stream=stream.cuda_stream)
    cfg = api.Config(BLOCK, 0, 0, 0, 0)
    out_n = ctypes.c_uint64()
    in_ptr, out_ptr = ctypes.c_void_p(host_in.data_ptr()), ctypes.c_void_p(host_out.data_ptr())

    def compress(nbytes=n):
        rc = L.crgpu_compress(h.h, ctypes.byref(cfg), in_ptr, ctypes.c_uint64(nbytes), out_ptr, ctypes.c_uint64(cap), ctypes.byref(out_n))
        if rc != 0:
            raise RuntimeError("crgpu_compress failed: %d" % rc)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(staged, k):
        total_ms, launches = 0.0, 0
        for _ in range(k):
            flush.fill_(1)                                   # evict L2 between iterations
            if staged:
                L.crgpu_stage_input(h.h, in_ptr, ctypes.c_uint64(n))
            barrier()
            l0 = L.crgpu_launch_count()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            compress()
            e1.record(stream)
            barrier()
            total_ms += e0.elapsed_time(e1)
            launches += L.crgpu_launch_count() - l0
        return total_ms, launches

    timed(False, warmup)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    h.profile(True)
    ms_res, launches = timed(True, steps)                    # value: input resident in HBM
    prof = h.profile_report()
    h.profile(False)
    ms_e2e, _ = timed(False, steps)                          # e2e: host buffers, copies inside
    clocks = sampler.stop() if rank == 0 else None
    container_bytes = out_n.value
    container_sha = sha(bytes(host_out.numpy()[:container_bytes]))

    t = torch.tensor([ms_res, ms_e2e], device="cuda", dtype=torch.float64)
    per_rank = [t.clone() for _ in range(world)]
    if world > 1:
        dist.all_gather(per_rank, t)
        ms_res, ms_e2e = max(float(x[0]) for x in per_rank), max(float(x[1]) for x in per_rank)
    else:
        per_rank = [t]
    peak, peak_src = peaks()

    line = None
    if rank == 0:
        value = world * n * steps / MiB / (ms_res / 1e3)
        e2e = world * n * steps / MiB / (ms_e2e / 1e3)
        stage_ms = {k: v / steps for k, v in prof.items() if not k.startswith("#")}
        counters = {k[1:]: v / steps for k, v in prof.items() if k.startswith("#")}
        # dominant stage of the step (CUDA events on the handle's stream around the stage's launches, live in this run)
        dom = max((k for k in stage_ms if k in STAGE_MODEL), key=lambda k: stage_ms[k])
        kern, model = STAGE_MODEL[dom]
        alg = float(model(counters, n, container_bytes))
        achieved = alg / (stage_ms[dom] / 1e3) / 1e9
        traffic = None
        try:   # per-launch DRAM bytes of the dominant kernels from the committed `ncu --set full` captures (profiles/round2_ncu_traffic.json)
            tj = json.load(open(os.path.join(ROOT, "profiles", "round2_ncu_traffic.json")))
            if dom in tj:
                traffic = int(tj[dom]["dram_bytes_per_unit"] * counters[tj[dom]["unit"]])
        except Exception:
            pass
        pipe_gbs = (n + container_bytes) / (ms_res / steps / 1e3) / 1e9
        line = {
            "metric": "compress_throughput", "value": round(value, 2), "unit": "MiB/s", "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": round(ms_res / steps, 2), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic", "config": config,
            "e2e": {"value": round(e2e, 2), "unit": "MiB/s", "h2d_bytes_per_step": n, "d2h_bytes_per_step": int(container_bytes)},
            "per_rank_ms": {"value": [round(float(x[0]) / steps, 2) for x in per_rank], "e2e": [round(float(x[1]) / steps, 2) for x in per_rank]},
            "gpu_launches": int(launches), "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": kern, "stage": dom, "achieved": round(achieved, 3), "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                         "frac": round(achieved / peak, 6), "traffic": traffic,
                         "traffic_source": None if traffic is None else "ncu (dram__bytes_read.sum + dram__bytes_write.sum of every kernel of the stage, one launch each), bytes per unit x units of this step (profiles/round2_ncu_traffic.json)",
                         "algorithmic_bytes": int(alg), "launch_ms": round(stage_ms[dom], 2),
                         "note": "the stage of the step with the largest CUDA-event time; bound by FP64 issue (tracking) and DFMA latency (serial emits), not by HBM: see DESIGN.md section 5"},
            "pipeline_roofline": {"achieved_gbs": round(pipe_gbs, 3), "frac": round(pipe_gbs / peak, 6), "algorithmic_bytes": "raw + container (SURVEY.md 8d)"},
            "stage_ms_per_step": {k: round(v, 2) for k, v in stage_ms.items()},
            "counters_per_step": {k: int(v) for k, v in counters.items()},
            "container_bytes": int(container_bytes), "container_sha256": container_sha,
        }
    # ---- parity + CPU baseline (N = 1): the first 32 MiB through the GPU and through the unmodified reference CLI
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sample_n = min(REF_SAMPLE, n)
        compress(sample_n)
        gpu_sha = sha(bytes(host_out.numpy()[:out_n.value]))
        exe = reference_binary()
        p = os.path.join(shm_dir(), "crbench_%d.in" % os.getpid())
        try:
            with open(p, "wb") as f:
                f.write(raw[:sample_n])
            if exe:
                dt, cpu, outs = run_reference_procs([p], [])
                ref_sha = sha(open(outs[0], "rb").read())
                os.unlink(outs[0])
                kind, cores = "reference", max(1, round(cpu / dt))
            else:
                sys.path.insert(0, os.path.join(ROOT, "tests"))
                import oracle_ffi as O
                t0 = time.perf_counter(); refc = O.compress(raw[:sample_n], 0, BLOCK); dt = time.perf_counter() - t0
                ref_sha, kind, cores = sha(refc), "port", 1
        finally:
            if os.path.exists(p):
                os.unlink(p)
        line["cpu_baseline"] = {"value": round(sample_n / MiB / dt, 3), "unit": "MiB/s", "cores": cores, "kind": kind,
                                "sample": "first %d MiB of the workload (2 blocks) through the unmodified reference CLI, one run" % (sample_n // MiB)}
        line["parity"] = {"sha_gpu": gpu_sha, "sha_ref": ref_sha, "identical": gpu_sha == ref_sha,
                          "sample": "container of the first %d MiB of the workload, `comprolz -b16`" % (sample_n // MiB)}
        if gpu_sha != ref_sha:                                   # a fast wrong answer is not a result
            line["value"] = None; line["e2e"]["value"] = None
            line["error"] = "GPU container differs from the reference CLI's"
    h.close()
    del flush
    # ---- decompression: many containers in flight on every GPU (replicas only: a container is one serial chain); the figure of
    # the job is the sum over the ranks' containers divided by the slowest rank's time
    if not args.no_decompress:
        try:
            d = decompress_leg(args, L, api, torch, raw, local_rank, peak)
            if world > 1:
                k = d["containers_in_flight"]; mine = torch.tensor([float(k * DEC_BYTES), k * DEC_BYTES / MiB / d["value"]], device="cuda", dtype=torch.float64)
                allr = [mine.clone() for _ in range(world)]
                dist.all_gather(allr, mine)
                tot = sum(float(x[0]) for x in allr); tmax = max(float(x[1]) for x in allr)
                d["per_gpu_value"] = d["value"]
                d["value"] = round(tot / MiB / tmax, 1)
                d["per_rank_s"] = [round(float(x[1]), 4) for x in allr]
                d["pipeline_roofline"]["note"] += "; rank 0's GPU"
                d["n_gpus"] = world
            if rank == 0:
                line["decompress"] = d
        except Exception as e:  # keep the compress line even if a leg fails
            if rank == 0:
                line["decompress"] = {"error": repr(e)}
    # ---- the mixed corpus in shard mode, strong scaling (all ranks)
    if not args.no_corpus:
        try:
            c = corpus_leg(args, L, api, torch, dist, rank, local_rank, world, peak)
            if rank == 0:
                line["corpus"] = c
        except Exception as e:
            if rank == 0:
                line["corpus"] = {"error": repr(e)}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
