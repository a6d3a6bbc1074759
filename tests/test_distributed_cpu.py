"""Host-side logic of the multi-GPU path on CPU: world_size-2 gloo processes
(sharding, bitmap packing, the single all-gather, the global stop test).  The
per-shard solve is stood in for by the CPU oracle here; on GPUs it is libcno.so."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cppnumericalsolvers_b200 import distributed as cd


def test_shard_range_partitions_batch():
    for B in (0, 1, 7, 64, 1000, 1 << 20):
        for world in (1, 2, 3, 8):
            r = [cd.shard_range(B, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == B
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            sizes = [hi - lo for lo, hi in r]
            assert max(sizes) - min(sizes) <= 1


def test_pack_done_bitmap_matches_bit_definition():
    rng = np.random.default_rng(0)
    for n in (1, 31, 32, 33, 100, 4096):
        st = torch.from_numpy(rng.integers(-1, 5, n).astype(np.int8))
        w = cd.pack_done_bitmap(st).numpy().view(np.uint32)
        for i in range(n):
            assert ((w[i // 32] >> (i % 32)) & 1) == int(st[i] not in (0, -1))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, B, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle_binding as ob
    d = 8
    lo, hi = cd.shard_range(B, rank, world)
    # counter-based starts: shard = a slice of the global stream, no scatter needed
    x0 = ob.fill_uniform((hi - lo, d), lo * d, 12345, -2.0, 2.0)
    r = ob.minimize(ob.LBFGS, ob.FN_ROSENBROCK, x0, threads=1)
    words = cd.pack_done_bitmap(torch.from_numpy(r["status"]))
    gathered = cd.gather_done_bitmaps(words)
    ok = cd.all_done(gathered, B)
    # a not-yet-finished instance on ONE rank must fail the global test on ALL ranks
    st2 = torch.from_numpy(r["status"].copy())
    if rank == world - 1:
        st2[0] = 0
    ok2 = cd.all_done(cd.gather_done_bitmaps(cd.pack_done_bitmap(st2)), B)
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), x=r["x"], it=r["num_iterations"], ok=ok, ok2=ok2,
             lo=lo, hi=hi)
    dist.destroy_process_group()


@pytest.mark.parametrize("B", [64, 101])
def test_two_rank_sharded_solve_equals_single_rank(tmp_path, B):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), B, str(tmp_path)), nprocs=world, join=True)
    from oracle import oracle_binding as ob
    full = ob.minimize(ob.LBFGS, ob.FN_ROSENBROCK, ob.fill_uniform((B, 8), 0, 12345, -2.0, 2.0), threads=1)
    xs, its = [], []
    for r in range(world):
        z = np.load(tmp_path / f"r{r}.npz")
        assert bool(z["ok"]) and not bool(z["ok2"])
        xs.append(z["x"])
        its.append(z["it"])
    # G-way sharded result == 1-rank result, instance for instance (bitwise)
    assert np.array_equal(np.concatenate(xs).view(np.uint64), full["x"].view(np.uint64))
    assert np.array_equal(np.concatenate(its), full["num_iterations"])
