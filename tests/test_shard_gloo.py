"""world_size-2 gloo test of the sharding helpers used by the N>1 path (CPU)."""
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    shard = importlib.import_module("3dfeaturematcher_b200.shard")
    from oracle import oracle_c as orc      # the compute stand-in of this CPU test
    synth = importlib.import_module("3dfeaturematcher_b200.synth")
    q, t, _ = synth.make_float_descriptors(301, 60, 9)
    # replicated inputs: only rank 0 holds the real train set before the broadcast
    tt = torch.from_numpy(t.copy()) if rank == 0 else torch.zeros(t.shape, dtype=torch.float32)
    shard.broadcast_([tt], src=0)
    lo, hi = shard.shard_bounds(q.shape[0], world, rank)
    idx, dd = orc.knn2_f32(q[lo:hi], tt.numpy())
    qi, ti, d = orc.nndr_filter(idx, dd, 0.55)
    n = len(qi)
    cap = max(shard.shard_sizes(q.shape[0], world))
    pad = lambda a, dt: torch.from_numpy(np.concatenate([a, np.zeros(cap - n, a.dtype)]).astype(dt))  # noqa: E731
    gq, gt, gd, counts = shard.gather_matches(pad(qi, np.int32), pad(ti, np.int32), pad(d, np.float32), n, lo, cap)
    # the single-collective form the bench uses: same rows, plus a second result set with its own count
    extra = torch.arange(cap * 3, dtype=torch.float64).reshape(cap, 3) + 1000.0 * rank
    n_extra = max(n - 2, 0)
    (pq, pt_, pd, pe), pc = shard.gather_packed([pad(qi, np.int32) + lo, pad(ti, np.int32), pad(d, np.float32), extra],
                                                [n, n, n, n_extra], cap)
    if rank == 0:
        ret["q"], ret["t"], ret["d"], ret["counts"] = gq.numpy(), gt.numpy(), gd.numpy(), counts
        ret["pq"], ret["pt"], ret["pd"], ret["pe"], ret["pc"] = pq.numpy(), pt_.numpy(), pd.numpy(), pe.numpy(), pc
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_matching_equals_single_process():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    shard = importlib.import_module("3dfeaturematcher_b200.shard")
    assert shard.shard_sizes(10, 3) == [4, 3, 3] and shard.shard_bounds(10, 3, 2) == (7, 10)
    assert sum(shard.shard_sizes(5, 8)) == 5
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(2, 29533 + os.getpid() % 500, ret), nprocs=2, join=True)
        from oracle import oracle_c as orc
        synth = importlib.import_module("3dfeaturematcher_b200.synth")
        q, t, _ = synth.make_float_descriptors(301, 60, 9)
        idx, dd = orc.knn2_f32(q, t)
        qi, ti, d = orc.nndr_filter(idx, dd, 0.55)
        np.testing.assert_array_equal(ret["q"], qi)       # rank-order concatenation == ascending query order
        np.testing.assert_array_equal(ret["t"], ti)
        np.testing.assert_array_equal(ret["d"], d)
        assert sum(ret["counts"]) == len(qi)
        np.testing.assert_array_equal(ret["pq"], qi)
        np.testing.assert_array_equal(ret["pt"], ti)
        np.testing.assert_array_equal(ret["pd"], d)
        pc = ret["pc"]
        assert [c[0] for c in pc] == list(ret["counts"]) and ret["pe"].shape == (sum(c[3] for c in pc), 3)
        assert ret["pe"][0, 0] == 0.0 and ret["pe"][pc[0][3], 0] == 1000.0      # rank 1's block follows rank 0's valid rows


def _worker_one_collective(rank, world, port, ret):
    """The per-step plumbing bench.py uses at N > 1: ONE broadcast of the replicated inputs, ONE all-gather of every
    per-shard result, header written on the device side (no host tensor in the step)."""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    shard = importlib.import_module("3dfeaturematcher_b200.shard")
    from oracle import oracle_c as orc      # the compute stand-in of this CPU test
    synth = importlib.import_module("3dfeaturematcher_b200.synth")
    q, t, _ = synth.make_float_descriptors(301, 60, 9)
    kp2 = np.arange(2 * t.shape[0], dtype=np.float32).reshape(-1, 2)
    img = (np.arange(48 * 37) % 251).astype(np.uint8).reshape(37, 48)
    rep = shard.ReplicatedBuffer([("t", t.shape, torch.float32), ("kp2", kp2.shape, torch.float32), ("img1", img.shape, torch.uint8)], "cpu")
    if rank == 0:
        rep["t"].copy_(torch.from_numpy(t)); rep["kp2"].copy_(torch.from_numpy(kp2)); rep["img1"].copy_(torch.from_numpy(img))
    rep.broadcast_(0)
    assert np.array_equal(rep["t"].numpy(), t) and np.array_equal(rep["kp2"].numpy(), kp2) and np.array_equal(rep["img1"].numpy(), img)
    lo, hi = shard.shard_bounds(q.shape[0], world, rank)
    idx, dd = orc.knn2_f32(q[lo:hi], rep["t"].numpy())
    qi, ti, d = orc.nndr_filter(idx, dd, 0.55)
    n = len(qi)
    cap = max(shard.shard_sizes(q.shape[0], world))
    pad = lambda a, dt: torch.from_numpy(np.concatenate([a, np.zeros(cap - len(a), a.dtype)]).astype(dt))  # noqa: E731
    n_inl = max(n - 3, 0)
    normals = torch.arange(cap * 3, dtype=torch.float64).reshape(cap, 3) + 1000.0 * rank
    g = shard.ShardGather([("qi", (), torch.int32, 0), ("ti", (), torch.int32, 0), ("d", (), torch.float32, 0),
                           ("normals", (3,), torch.float64, 1)], cap, "cpu", rank)
    for _ in range(2):                       # the buffers are reused step after step
        g.gather({"qi": pad(qi, np.int32), "ti": pad(ti, np.int32), "d": pad(d, np.float32), "normals": normals},
                 (torch.tensor([n], dtype=torch.int32), torch.tensor([n_inl], dtype=torch.int32)))
    out, counts = g.unpack()
    if rank == 0:
        los = [shard.shard_bounds(q.shape[0], world, r)[0] for r in range(world)]
        rk = np.repeat(np.arange(world), [c[0] for c in counts])
        ret["q"] = out["qi"].numpy().astype(np.int64) + np.asarray(los)[rk]
        ret["t"], ret["d"], ret["n"], ret["counts"] = out["ti"].numpy(), out["d"].numpy(), out["normals"].numpy(), counts
    dist.barrier()
    dist.destroy_process_group()


def test_one_broadcast_one_gather_equals_single_process():
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker_one_collective, args=(2, 30133 + os.getpid() % 500, ret), nprocs=2, join=True)
        from oracle import oracle_c as orc
        synth = importlib.import_module("3dfeaturematcher_b200.synth")
        q, t, _ = synth.make_float_descriptors(301, 60, 9)
        idx, dd = orc.knn2_f32(q, t)
        qi, ti, d = orc.nndr_filter(idx, dd, 0.55)
        np.testing.assert_array_equal(ret["q"], qi)       # rank-order concatenation == ascending global query order
        np.testing.assert_array_equal(ret["t"], ti)
        np.testing.assert_array_equal(ret["d"], d)
        c = ret["counts"]
        assert sum(x[0] for x in c) == len(qi) and ret["n"].shape == (sum(x[1] for x in c), 3)
        assert ret["n"][0, 0] == 0.0 and ret["n"][c[0][1], 0] == 1000.0     # rank 1's valid rows follow rank 0's
