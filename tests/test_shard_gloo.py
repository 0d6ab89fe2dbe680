"""The N > 1 path on CPU: two gloo ranks compress disjoint shards (through the kernel-logic simulation library, since
this container has no GPU) and rank 0 gathers the containers in order.  The result must equal a single-process run."""
import os
import sys

import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NSHARDS = 5


def _shard(i):
    from comprox_b200 import synth
    return synth.markov_text(120000 + 7000 * i, seed=100 + i)


def _worker(rank, world, port, simso, outdir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from comprox_b200 import api, shard
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    lib = api.load(simso)
    mine = {}
    for i in shard.shards_of(rank, world, NSHARDS):
        with api.Handle(api.ROLZ, lib=lib) as h:
            mine[i] = h.compress(_shard(i), 1 << 20)
    slowest = shard.max_over_ranks(float(rank + 1))
    assert slowest == float(world)
    allc = shard.gather_containers(mine, NSHARDS)
    if rank == 0:
        with open(os.path.join(outdir, "gathered.bin"), "wb") as f:
            for c in allc:
                f.write(len(c).to_bytes(8, "little") + c)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_compress_and_gather(simlib, tmp_path):
    import oracle_ffi as O
    from comprox_b200 import shard
    assert shard.shards_of(0, 2, 5) == [0, 2, 4] and shard.shards_of(1, 2, 5) == [1, 3]
    simso = os.path.join(ROOT, "tests", "sim", "libcrgpu_sim.so")
    port = 29600 + os.getpid() % 200
    mp.spawn(_worker, args=(2, port, simso, str(tmp_path)), nprocs=2, join=True)
    blob = open(tmp_path / "gathered.bin", "rb").read()
    pos, got = 0, []
    while pos < len(blob):
        n = int.from_bytes(blob[pos:pos + 8], "little")
        got.append(blob[pos + 8:pos + 8 + n])
        pos += 8 + n
    assert len(got) == NSHARDS
    for i, c in enumerate(got):
        assert c == O.compress(_shard(i), 0, 1 << 20), "shard %d differs from the single-process oracle container" % i
