"""BASELINE.json configs[3] and configs[4] at 1 / 2 / 4 / 8 GPUs (one process per GPU, libfm3d's own NCCL communicator).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/bench_sweeps_multi.py [match] [normals]            (also runs as a plain single process)

match    configs[3]: descriptor matching sweep, n x n for n in 10k..200k, SIFT-128 float (tcgen05 path) and ORB-256 binary
         (popc path).  Queries are sharded over the ranks, the train set is replicated with ONE ncclBroadcast from rank 0 and
         the per-shard (idx, dist) records are collected with ONE ncclAllGather; both collectives are INSIDE the timed region.
         Rank 0 checks a seeded sample of the gathered records bit for bit against the CPU oracle.
normals  configs[4]: NormalOptimizer stress, 50 000 features of one 4K scene sharded over the ranks, pixelsRay 32 / 64 / 128,
         3 / 4 / 5 pyramid images, plus the dense candidate sweep (33 x 33 grid per feature at level 0) on a subset.
Times are CUDA events on the context's stream, max over ranks; one JSON line per case on rank 0."""
import importlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

api = importlib.import_module("3dfeaturematcher_b200.api")
synth = importlib.import_module("3dfeaturematcher_b200.synth")
shard = importlib.import_module("3dfeaturematcher_b200.shard")
PEAKS = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local_rank = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local_rank)
dev = torch.device("cuda", local_rank)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
ctx = api.Context(local_rank)
if world > 1:
    ident = [api.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ident, src=0)
    ctx.comm_init_rank(ident[0], world, rank)
stream = torch.cuda.ExternalStream(ctx.stream, device=dev)


def emit(obj):
    if rank == 0:
        print(json.dumps(obj), flush=True)


def timed_max(fn, reps, warm=1):
    """ms per call of fn (work on the context's stream), max over ranks; all ranks start together."""
    for _ in range(warm):
        fn()
    stream.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
    stream.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def match_sweep(sizes):
    from oracle import oracle_c as orc
    for n in sizes:
        lo, hi = shard.shard_bounds(n, world, rank)
        nq = hi - lo
        cap = max(shard.shard_sizes(n, world))
        g = torch.Generator(device=dev).manual_seed(1003)          # every rank generates the same global problem
        with torch.cuda.stream(stream):
            t_all = torch.randint(0, 256, (n, 128), device=dev, generator=g).float()
            perm = torch.randperm(n, device=dev, generator=g)
            q_all = (t_all[perm] + torch.randint(-6, 7, (n, 128), device=dev, generator=g).float()).clamp_(0, 255)
            tb_all = torch.randint(0, 256, (n, 32), device=dev, generator=g, dtype=torch.uint8)
            qb_all = tb_all[torch.randperm(n, device=dev, generator=g)].clone()
            flip = torch.randint(0, 256, (n, 3), device=dev, generator=g, dtype=torch.uint8)
            qb_all[:, :3] ^= flip
            q, qb = q_all[lo:hi].contiguous(), qb_all[lo:hi].contiguous()
            # replicated operands: only rank 0's copy is meaningful before the broadcast
            t = t_all.clone() if rank == 0 else torch.zeros_like(t_all)
            tb = tb_all.clone() if rank == 0 else torch.zeros_like(tb_all)
            send = torch.zeros((cap, 4), dtype=torch.int32, device=dev)       # idx0, idx1, bits(d0), bits(d1)
            recv = torch.zeros((world, cap, 4), dtype=torch.int32, device=dev) if world > 1 else send.view(1, cap, 4)
            idx, dd = send[:, :2], send[:, 2:].view(torch.float32)
        stream.synchronize()
        h_q, h_t = (q_all.cpu().numpy(), t_all.cpu().numpy()) if rank == 0 else (None, None)
        h_qb, h_tb = (qb_all.cpu().numpy(), tb_all.cpu().numpy()) if rank == 0 else (None, None)
        del t_all, q_all, tb_all, qb_all
        idx_c = torch.empty((nq, 2), dtype=torch.int32, device=dev)
        dist_c = torch.empty((nq, 2), dtype=torch.float32, device=dev)

        def step_f32():
            if world > 1:
                ctx.broadcast_dev(t.data_ptr(), t.numel() * 4, 0)
            ctx.match_knn2_f32_dev(q.data_ptr(), nq, t.data_ptr(), n, 128, idx_c.data_ptr(), dist_c.data_ptr())
            with torch.cuda.stream(stream):
                send[:nq, :2].copy_(idx_c); send[:nq, 2:].copy_(dist_c.view(torch.int32))
            if world > 1:
                ctx.allgather_dev(send.data_ptr(), recv.data_ptr(), send.numel() * 4)

        def step_ham():
            if world > 1:
                ctx.broadcast_dev(tb.data_ptr(), tb.numel(), 0)
            ctx.match_knn2_hamming_dev(qb.data_ptr(), nq, tb.data_ptr(), n, 32, idx_c.data_ptr(), dist_c.data_ptr())
            with torch.cuda.stream(stream):
                send[:nq, :2].copy_(idx_c); send[:nq, 2:].copy_(dist_c.view(torch.int32))
            if world > 1:
                ctx.allgather_dev(send.data_ptr(), recv.data_ptr(), send.numel() * 4)

        reps = 5 if n <= 50000 else 2
        for name, step, hq, ht, knn in (("match_f32_sift128_tcgen05", step_f32, h_q, h_t, orc.knn2_f32),
                                        ("match_hamming_orb256_popc", step_ham, h_qb, h_tb, orc.knn2_hamming)):
            ms = timed_max(step, reps)
            check = None
            if rank == 0:
                got = recv.cpu().numpy()
                sizes_r = shard.shard_sizes(n, world)
                rows = np.concatenate([got[r, :sizes_r[r]] for r in range(world)])           # rank order = ascending query index
                rng = np.random.default_rng(5)
                sel = np.sort(rng.choice(n, 64, replace=False))
                oi, od = knn(hq[sel], ht, threads=os.cpu_count() or 1)
                ok = np.array_equal(rows[sel, :2], oi) and np.array_equal(rows[sel, 2:].view(np.float32), od)
                check = "ok" if ok else "MISMATCH"
            out = {"case": name, "n_gpus": world, "nq": n, "nt": n, "ms": ms, "pairs_per_s": n * n / (ms * 1e-3),
                   "collectives_in_timed_region": ["ncclBroadcast(train set)", "ncclAllGather(idx, dist)"] if world > 1 else [],
                   "oracle_check_64_queries": check}
            if "f32" in name:
                tf = 256.0 * n * n / (ms * 1e-3) / 1e12
                out.update({"tflops": tf, "frac_of_measured_bf16_peak_x_n_gpus": tf / (PEAKS.get("bf16_tflops", 1590.0) * world),
                            "train_set_bytes": n * 128 * 4})
            else:
                out.update({"word_ops_per_s": 8.0 * n * n / (ms * 1e-3), "frac_of_measured_popc_peak_x_n_gpus": 8.0 * n * n / (ms * 1e-3) / (4.54e12 * world),
                            "train_set_bytes": n * 32})
            emit(out)
        del t, tb, q, qb, send, recv
        torch.cuda.empty_cache()


def normals_stress(n_total):
    t0 = time.time()
    # one scene, keypoints safe for the largest disc: the same features serve every pixelsRay
    case = synth.make_stereo_case(3840, 2160, n_total, 1005, pixels_ray=128, n_distractors=0)
    if rank == 0:
        print(f"# C5 scene + {n_total} features generated in {time.time() - t0:.0f} s", file=sys.stderr, flush=True)
    cam = case["scene"].cam
    ctx.set_camera(cam.K, cam.dist, cam.z_min, cam.z_max)
    ctx.set_g12(cam.g12)
    lo, hi = shard.shard_bounds(n_total, world, rank)
    xyz = torch.from_numpy(np.ascontiguousarray(case["X"][lo:hi])).to(dev)
    n = hi - lo
    for r in (32, 64, 128):
        m_disc = sum(2 * int(np.floor(np.sqrt(r * r - j * j))) + 1 for j in range(-r, r + 1))
        for pyr in (2, 3, 4):
            ctx.set_images(case["scene"].img1, case["scene"].img2, pyr)
            normals = torch.empty((n, 3), dtype=torch.float64, device=dev)
            status = torch.empty(n, dtype=torch.int32, device=dev)
            nfev = torch.zeros((n, pyr + 1), dtype=torch.int32, device=dev)
            npen = torch.zeros(n, dtype=torch.int32, device=dev)
            cost = torch.empty(n, dtype=torch.float64, device=dev)
            fn = lambda: ctx.optimize_normals_dev(xyz.data_ptr(), n, r, 1e-10, 1, normals.data_ptr(), status.data_ptr(),  # noqa: E731
                                                  nfev.data_ptr(), npen.data_ptr(), cost.data_ptr())
            ms = timed_max(fn, 1)
            st = ctx.normals_stats()
            nr = normals.cpu().numpy()
            gt = np.degrees(np.arccos(np.clip((nr * case["normal"][lo:hi]).sum(1), -1, 1)))
            ok = status.cpu().numpy() == 0
            agg = torch.tensor([float(ok.sum()), float(st["pixel_evals_value"]), float(st["pixel_evals_jacobian"]), float(nfev.sum().item()),
                                float(np.sum(gt[ok] <= 0.5))], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(agg)
            okc, pv, pj, nf, good = (float(v) for v in agg)
            emit({"case": "normals_stress", "n_gpus": world, "pixels_ray": r, "m": m_disc, "pyramid_images": pyr + 1, "features": n_total,
                  "ms": ms, "features_per_s": n_total / (ms * 1e-3), "ok": int(okc), "within_0p5_deg_of_ground_truth": int(good),
                  "median_angle_to_gt_deg_rank0": float(np.median(gt[ok])) if ok.any() else None,
                  "pixel_evals_per_s": (pv + pj) / (ms * 1e-3), "tflops_fp32_executed": (pv * 64.0 + pj * 152.0) / (ms * 1e-3) / 1e12,
                  "reference_equivalent_tflops": nf * m_disc * 64.0 / (ms * 1e-3) / 1e12,
                  "frac_of_fp32_peak_x_n_gpus": (pv * 64.0 + pj * 152.0) / (ms * 1e-3) / 1e12 / (73.96 * world),
                  "passes_global_taps_rank0": st["passes_slow"]})
            if pyr == 3:
                # dense candidate-normal sampling: a 33 x 33 grid around the initial normal at level 0, 2 368 features in total
                nsw = min(n, max(1, 2368 // world))
                best = torch.empty(n, dtype=torch.int32, device=dev)
                bcost = torch.empty(n, dtype=torch.float64, device=dev)
                fn2 = lambda: ctx.sweep_normals_dev(xyz.data_ptr(), nsw, r, 0, 33, 33, 0.01, 0.01, status.data_ptr(), penalty_mode=1,  # noqa: E731
                                                    best_idx=best.data_ptr(), best_cost=bcost.data_ptr())
                ms2 = timed_max(fn2, 1)
                tot = nsw * world
                emit({"case": "normals_dense_sweep_33x33_level0", "n_gpus": world, "pixels_ray": r, "m": m_disc, "features": tot, "ms": ms2,
                      "pixel_evals_per_s": tot * 1089.0 * m_disc / (ms2 * 1e-3),
                      "tflops_fp32_algorithmic": tot * 1089.0 * m_disc * 64.0 / (ms2 * 1e-3) / 1e12,
                      "candidates_per_s": tot * 1089.0 / (ms2 * 1e-3)})


if __name__ == "__main__":
    which = sys.argv[1:] or ["match", "normals"]
    if "match" in which:
        match_sweep([10000, 50000, 100000, 200000])
    if "normals" in which:
        normals_stress(50000)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()
