#!/usr/bin/env python
"""bench.py -- headline benchmark of comprox-b200 (contract: see the task description / DESIGN.md section 6).

Workload (BASELINE.json configs[1]): `comprolz` default settings (ROLZ, 16 MiB blocks) on a 100 MiB synthetic
word-level-Markov English-like text file, one container per GPU.  A "step" is one whole-container compression.
  value : MiB/s with the input already resident in HBM (crgpu_stage_input) when the timed region starts
  e2e   : MiB/s through the public C ABI with pinned HOST buffers (H2D of the input and D2H of the container inside)
  --impl reference : the UNMODIFIED reference CLI (oracle/_ref/comprolz, built by oracle/Makefile) on the host cores
N > 1 (torchrun): every rank compresses its own 100 MiB container (independent shards, no data-path collective;
SURVEY.md section 8e) -- weak scaling; time = max over ranks.
"""
import argparse
import json
import os
import resource
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MiB = 1 << 20
WORKLOAD_BYTES = 100 * MiB
BLOCK = 16 * MiB
REF_SAMPLE = 32 * MiB          # bounded sample for the CPU arms (2 blocks; models still chain across them)
SHARD_SAMPLE = 16 * MiB        # container size of the reference's supplementary shard-mode leg (one process per host core)
SHARD_BYTES = 64 * MiB         # container size of our shard-mode leg (profiles/round1_shards_one_gpu.jsonl uses the same)
SHARD_HANDLES = 8


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


def reference_binary():
    p = os.path.join(ROOT, "oracle", "_ref", "comprolz")
    return p if os.path.exists(p) else None


def time_reference(data, steps, warmup):
    """Wall time of the unmodified reference CLI on `data` (files in /dev/shm, -q).  Returns (seconds per step, cores)."""
    exe = reference_binary()
    tmp = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
    src = os.path.join(tmp, "crbench_%d.in" % os.getpid())
    with open(src, "wb") as f:
        f.write(data)
    times, cpu = [], []
    try:
        for i in range(warmup + steps):
            r0 = resource.getrusage(resource.RUSAGE_CHILDREN)
            t0 = time.perf_counter()
            if exe:
                subprocess.run([exe, "-q", "e", src, src + ".out"], check=True)
            else:   # reference binaries did not travel: time the oracle port instead
                sys.path.insert(0, os.path.join(ROOT, "tests"))
                import oracle_ffi as O
                O.compress(data, 0, BLOCK)
            dt = time.perf_counter() - t0
            r1 = resource.getrusage(resource.RUSAGE_CHILDREN)
            if i >= warmup:
                times.append(dt)
                cpu.append((r1.ru_utime - r0.ru_utime) + (r1.ru_stime - r0.ru_stime))
    finally:
        for p in (src, src + ".out"):
            if os.path.exists(p):
                os.unlink(p)
    sec = sum(times) / len(times)
    cores = max(1, round(sum(cpu) / sum(times))) if exe else 1
    return sec, cores, ("reference" if exe else "port")


def time_reference_shards(data, nproc):
    """Shard mode of the reference: `nproc` unmodified reference CLI processes side by side, one independent
    container each (the same sample), wall time around all of them.  Returns aggregate MiB/s or None without the binary."""
    exe = reference_binary()
    if not exe or nproc < 2:
        return None
    tmp = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
    src = os.path.join(tmp, "crbench_%d.sin" % os.getpid())
    with open(src, "wb") as f:
        f.write(data)
    outs = ["%s.%d.out" % (src, i) for i in range(nproc)]
    try:
        t0 = time.perf_counter()
        ps = [subprocess.Popen([exe, "-q", "e", src, o]) for o in outs]
        ok = all(p.wait() == 0 for p in ps)
        dt = time.perf_counter() - t0
    finally:
        for p in [src] + outs:
            if os.path.exists(p):
                os.unlink(p)
    return nproc * len(data) / MiB / dt if ok else None


def shard_procs():
    """How many reference processes the shard-mode leg runs: every host core, bounded by memory (~0.4 GB per process) and by 64."""
    n = os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except AttributeError:
        pass
    return max(1, min(n, 64))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--bytes", type=int, default=WORKLOAD_BYTES, help="workload size (default: the 100 MiB the metric is quoted on)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-shard-leg", action="store_true", help="skip the supplementary shard-mode legs")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    steps, warmup = args.steps, max(args.warmup, 3 if args.impl == "ours" else 0)

    from comprox_b200 import synth
    config = {"workload": "text-100M: comprolz (ROLZ) default -b16 on %d bytes of synthetic word-Markov text (seed 42+rank), one container per GPU" % args.bytes,
              "block_size": BLOCK, "containers_per_gpu": 1, "l2": "256 MiB device buffer rewritten between timed steps"}

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        data = synth.markov_text(min(REF_SAMPLE, args.bytes), seed=42)
        sec, cores, kind = time_reference(data, max(steps, 1), min(warmup, 1))
        v = len(data) / MiB / sec
        sample = "first %d MiB of the workload (2 blocks), unmodified reference CLI, files in /dev/shm" % (len(data) // MiB)
        line = {"impl": "reference", "metric": "compress_throughput", "value": round(v, 3), "unit": "MiB/s", "n_gpus": args.gpus,
                "steps": steps, "warmup": warmup, "ms_per_step": round(sec * 1e3, 1), "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": round(v, 3), "unit": "MiB/s", "cores": cores, "kind": kind, "sample": sample,
                                 "note": "one container = one serial model chain: the reference's threads are intra-block helpers (SURVEY.md F1), "
                                         "so this workload cannot use more host cores than this"},
                "e2e": {"value": round(v, 3), "unit": "MiB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        if not args.no_shard_leg:
            # supplementary: what ALL host cores give when the job is many independent containers (not this line's workload)
            P = shard_procs()
            sv = time_reference_shards(data[:SHARD_SAMPLE], P)
            if sv is not None:
                line["shard_mode"] = {"value": round(sv, 2), "unit": "MiB/s", "processes": P, "host_cores": os.cpu_count(),
                                      "sample": "%d independent containers of %d MiB, one unmodified reference CLI process each, side by side" % (P, SHARD_SAMPLE // MiB)}
        print(json.dumps(line))
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

    def compress():
        rc = L.crgpu_compress(h.h, ctypes.byref(cfg), in_ptr, ctypes.c_uint64(n), out_ptr, ctypes.c_uint64(cap), ctypes.byref(out_n))
        if rc != 0:
            raise RuntimeError("crgpu_compress failed: %d" % rc)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(staged, k, profile=False):
        total_ms, launches = 0.0, 0
        for _ in range(k):
            flush.fill_(1)                                   # evict L2 between iterations
            if staged:
                L.crgpu_stage_input(h.h, in_ptr, ctypes.c_uint64(n))
            barrier()
            if profile:
                h.profile(True)
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

    if world > 1:
        t = torch.tensor([ms_res, ms_e2e], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_res, ms_e2e = t.tolist()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    value = world * n * steps / MiB / (ms_res / 1e3)
    e2e = world * n * steps / MiB / (ms_e2e / 1e3)
    peak, peak_src = peaks()
    # dominant kernel = the serial range chain, one per (block, stream).  Algorithmic bytes: 12 B per triple in,
    # and its share of the coded bytes out is produced by the parallel low stage (reported under stage_ms_per_step.range_coder).
    rc_ms = prof.get("range_chain", 0.0) / steps
    triples = prof.get("#triples", 0.0) / steps
    rc_bytes = 12.0 * triples + container_bytes
    achieved = rc_bytes / (rc_ms / 1e3) / 1e9 if rc_ms > 0 else 0.0
    # DRAM traffic of the same kernel from the committed `ncu --set full` capture (profiles/round1b_ncu_range_chain_v4.txt:
    # 174.08 MB read + 55.39 MB written by one launch over 10.81 M triples = 21.2 B per triple: the 16-B padded triple in,
    # 4 B quotient + 4 B shift record out, minus what stays in L2), scaled to the triples of this launch.
    NCU_TRAFFIC_PER_TRIPLE = (174.081536e6 + 55.391488e6) / 10.809709e6
    traffic = NCU_TRAFFIC_PER_TRIPLE * triples if triples else None
    pipe_gbs = (n + container_bytes) / (ms_res / steps / 1e3) / 1e9
    line = {
        "metric": "compress_throughput", "value": round(value, 2), "unit": "MiB/s", "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": round(ms_res / steps, 2), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
        "data": "synthetic", "config": config,
        "e2e": {"value": round(e2e, 2), "unit": "MiB/s", "h2d_bytes_per_step": n, "d2h_bytes_per_step": int(container_bytes)},
        "gpu_launches": int(launches), "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "k_range_chain", "achieved": round(achieved, 3), "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                     "frac": round(achieved / peak, 6), "traffic": None if traffic is None else int(traffic),
                     "traffic_source": "ncu --set full capture of this kernel (profiles/round1b_ncu_range_chain_v4.txt), bytes per triple x triples of this launch",
                     "algorithmic_bytes": int(rc_bytes), "launch_ms": round(rc_ms, 2),
                     "note": "serial recurrence per (block, stream): latency bound by construction, see DESIGN.md section 5"},
        "pipeline_roofline": {"achieved_gbs": round(pipe_gbs, 3), "frac": round(pipe_gbs / peak, 6), "algorithmic_bytes": "raw + container (SURVEY.md 8d)"},
        "stage_ms_per_step": {k: round(v / steps, 2) for k, v in prof.items() if not k.startswith("#")},
        "counters_per_step": {k[1:]: int(v / steps) for k, v in prof.items() if k.startswith("#")},
        "container_bytes": int(container_bytes),
    }
    # decompression: one container is one serial chain (replicas only), so this is a per-container latency figure
    try:
        dsample = raw[:4 * MiB]
        with api.Handle(api.ROLZ, device=local_rank, stream=stream.cuda_stream) as hd:
            small = hd.compress(dsample, BLOCK)
            hd.decompress(small, len(dsample) + 64)
            t0 = time.perf_counter(); back = hd.decompress(small, len(dsample) + 64); dt = time.perf_counter() - t0
        line["decompress"] = {"value": round(len(dsample) / MiB / dt, 2), "unit": "MiB/s", "sample": "4 MiB text container, one warp (one container = one serial model chain)",
                              "roundtrip_ok": back == dsample}
    except Exception as e:  # keep the compress line even if the decode leg fails
        line["decompress"] = {"error": str(e)}
    # the scalable decode mode: many independent containers in flight, one warp each, one launch per phase
    try:
        K, per = 296, MiB
        parts = [raw[i * 4096:i * 4096 + per] for i in range(K)]
        with api.Handle(api.ROLZ, device=local_rank, stream=stream.cuda_stream) as hc:
            conts = [hc.compress(p, BLOCK) for p in parts]
        hs = [api.Handle(api.ROLZ, device=local_rank, stream=stream.cuda_stream) for _ in range(K)]
        try:
            api.decompress_batch(hs, conts, [per + 64] * K)
            t0 = time.perf_counter(); backs = api.decompress_batch(hs, conts, [per + 64] * K); dt = time.perf_counter() - t0
        finally:
            for hh in hs:
                hh.close()
        line["decompress_batch"] = {"value": round(K * per / MiB / dt, 1), "unit": "MiB/s", "containers_in_flight": K,
                                    "sample": "%d text containers of 1 MiB through crgpu_decompress_batch, host buffers in and out" % K,
                                    "roundtrip_ok": backs == parts}
    except Exception as e:
        line["decompress_batch"] = {"error": str(e)}
    # supplementary: shard mode on ONE GPU -- independent containers compressed side by side by several handles with
    # private streams, one host thread each; the serial range chains of different containers overlap on different SMs.  Wall clock
    # around the C ABI calls (pinned host buffers in and out, copies inside), not part of `value`.
    if not args.no_shard_leg:
        try:
            K, per, rounds = SHARD_HANDLES, SHARD_BYTES, 2
            hs = [api.Handle(api.ROLZ, device=local_rank, stream=api.OWN_STREAM) for _ in range(K)]
            scap = int(L.crgpu_compress_bound(ctypes.c_uint64(per), ctypes.c_uint32(BLOCK)))
            ins, outs, lens = [], [], [ctypes.c_uint64() for _ in range(K)]
            for j in range(K):                              # distinct containers: the workload's text from different offsets
                t = torch.empty(per, dtype=torch.uint8).pin_memory()
                off = (j * 4099 * 1021) % (n - per)
                t.numpy()[:] = memoryview(raw)[off:off + per]
                ins.append(t); outs.append(torch.empty(scap, dtype=torch.uint8).pin_memory())
            errs = []

            def work(j):
                for _ in range(rounds):
                    rc = L.crgpu_compress(hs[j].h, ctypes.byref(cfg), ctypes.c_void_p(ins[j].data_ptr()), ctypes.c_uint64(per),
                                          ctypes.c_void_p(outs[j].data_ptr()), ctypes.c_uint64(scap), ctypes.byref(lens[j]))
                    if rc != 0:
                        errs.append(rc)

            def run_all():
                th = [threading.Thread(target=work, args=(j,)) for j in range(K)]
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                [t.start() for t in th]; [t.join() for t in th]
                torch.cuda.synchronize()
                return time.perf_counter() - t0
            try:
                run_all()                                   # allocations, first touch
                dt = run_all()
            finally:
                for hh in hs:
                    hh.close()
            if errs:
                raise RuntimeError("crgpu_compress failed in shard mode: %s" % errs[:3])
            with api.Handle(api.ROLZ, device=local_rank, stream=stream.cuda_stream) as h1:
                same = all(bytes(outs[j].numpy()[:lens[j].value]) == h1.compress(bytes(ins[j].numpy()), BLOCK) for j in (0, K - 1))
            line["shard_mode"] = {"value": round(K * rounds * per / MiB / dt, 1), "unit": "MiB/s", "handles": K, "containers": K * rounds,
                                  "sample": "%d independent containers of %d MiB, %d handles with private streams on %d host threads, pinned host buffers, wall clock"
                                            % (K * rounds, per // MiB, K, K),
                                  "identical_to_single_handle": same}
        except Exception as e:
            line["shard_mode"] = {"error": str(e)}
    if not args.no_cpu_baseline and world == 1:
        sample = raw[:REF_SAMPLE]
        sec, cores, kind = time_reference(sample, 1, 0)
        line["cpu_baseline"] = {"value": round(len(sample) / MiB / sec, 3), "unit": "MiB/s", "cores": cores, "kind": kind,
                                "sample": "first %d MiB of the workload (2 blocks) through the unmodified reference CLI, one run" % (len(sample) // MiB)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
