"""N > 1 host-side logic on the CPU with the gloo backend (world_size 2): the sharding rule
of the C ABI (gpc_shard_range: pure host arithmetic) partitions the patches exactly, the
per-rank pieces of an oracle fit concatenate to the single-rank result, and bench.py's
reference arm behaves under a multi-rank launch (rank 0 prints, the others exit 0)."""
import json
import os
import subprocess
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gp_compressor_b200 import binding
    from oracle import oracle as O
    rng = np.random.default_rng(3)  # same data on every rank, as in the sharded compress
    sizes = rng.integers(0, 90, 61)
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    n = int(off[-1])
    x1 = rng.uniform(-0.05, 0.05, n)
    x2 = rng.uniform(-0.05, 0.05, n)
    y = 0.02 * np.sin(40 * x1) + rng.normal(0, 0.003, n)
    lo, hi = binding.shard_range(off, rank, world)
    # every rank fits only its own patch range; the rand stream offset of the range is the
    # number of draws the earlier patches consume (2 * (n_q - 1) per non-empty patch)
    draws_before = int(sum(2 * (s - 1) for s in sizes[:lo] if s > 0))
    o = O.Oracle(capacity=10, sigmaf_sq=1.0, l_sq=(0.1 / 12) ** 2, s0=1e-4)
    o.set_rand_offset(draws_before)
    sub_off = off[lo:hi + 1] - off[lo]
    r = o.fit_patches(sub_off, x1[off[lo]:off[hi]], x2[off[lo]:off[hi]], y[off[lo]:off[hi]])
    rng_t = torch.tensor([lo, hi], dtype=torch.int64)
    allr = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(allr, rng_t)
    nb = torch.zeros(len(sizes), dtype=torch.int64)
    nb[lo:hi] = torch.from_numpy(r["nbv"].astype(np.int64))
    dist.all_reduce(nb)
    asum = torch.tensor([float(r["alpha"].sum())], dtype=torch.float64)
    parts = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(parts, asum)
    if rank == 0:
        full = O.Oracle(capacity=10, sigmaf_sq=1.0, l_sq=(0.1 / 12) ** 2, s0=1e-4).fit_patches(off, x1, x2, y)
        q.put(dict(ranges=[t.tolist() for t in allr], nbv_ok=bool(np.array_equal(nb.numpy(), full["nbv"])),
                   n=len(sizes), alpha_parts=[float(p) for p in parts],
                   alpha_ranges=[float(full["alpha"][full["bv_off"][a]:full["bv_off"][b]].sum()) for a, b in [t.tolist() for t in allr]]))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_partition_and_match_single_rank():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, q), nprocs=world, join=True)
    res = q.get()
    ranges = res["ranges"]
    assert ranges[0][0] == 0 and ranges[-1][1] == res["n"]
    for a, b in zip(ranges[:-1], ranges[1:]):
        assert a[1] == b[0]
    assert res["nbv_ok"]
    assert res["alpha_parts"] == res["alpha_ranges"]  # bit-identical pieces


def test_shard_range_properties():
    sys.path.insert(0, ROOT)
    from gp_compressor_b200 import binding
    rng = np.random.default_rng(0)
    for trial in range(20):
        P = int(rng.integers(0, 50))
        off = np.concatenate([[0], np.cumsum(rng.integers(0, 30, P))]).astype(np.int64)
        for world in (1, 2, 3, 8):
            prev = 0
            for r in range(world):
                lo, hi = binding.shard_range(off, r, world)
                assert lo == prev and hi >= lo
                prev = hi
            assert prev == P


def test_bench_reference_arm_under_torchrun():
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(29700 + os.getpid() % 200), os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
           "--steps", "1", "--warmup", "0", "--workload", "c1", "--points", "20000"]
    out = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["value"] > 0 and d["cpu_baseline"]["kind"] == "port"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["n_gpus"] == 2


def _prefix_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gp_compressor_b200 import binding
    # what gpc_compress_shard_begin returns on each rank (owned patches, owned rand() draws); the one exchange of the path
    mine = torch.tensor([[1000 + 37 * rank, 2 * (5000 + 11 * rank)]], dtype=torch.int64)
    allc = [torch.zeros(1, 2, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(allc, mine)
    pre = binding.shard_prefix(torch.cat(allc).numpy(), rank)
    outs = [None] * world
    dist.all_gather_object(outs, pre)
    if rank == 0:
        q.put(outs)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_binning_exchange_prefixes():
    """Host logic of the sharded-binning compress at world_size 2 (gloo): after one all-gather of two integers every rank
    knows the global index of its first patch, its window of the rand() stream and the totals."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 400) + 17
    ps = [ctx.Process(target=_prefix_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    outs = q.get(timeout=120)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert outs[0] == (0, 0, 2037, 20022)
    assert outs[1] == (1000, 10000, 2037, 20022)
