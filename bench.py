#!/usr/bin/env python
"""bench.py -- features/sec of the match + triangulate + normal-optimise hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl fm3d|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step is one pass of the hot path over one synthetic stereo pair of BASELINE.json configs[1]:
1280x720, 5 000 SIFT-128 float keypoints per GPU, pixelsRay 64, pyramids 3 (4 LM stages),
NNDR 0.55, penalty wall as the reference's author built it (int abs; the fabs variant of today's
g++ is measured next to it as `penalty_fabs`).  Inputs are resident in HBM for `value`; `e2e` is
the same step through the host-buffer C-ABI calls (pinned host inputs, copies inside the
timed region).  With N ranks every rank owns a contiguous shard of N*5000 query keypoints
(weak scaling), rank 0's train descriptors and images are broadcast with NCCL and the
per-shard matches and normals are all-gathered inside the timed region.

`--impl reference` times the reference's CPU path (the plain-C oracle port: the reference
itself cannot be built in this image) with all host threads on a bounded sample of the same
workload.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WIDTH, HEIGHT, N_KP, N_DISTRACT = 1280, 720, 5000, 1000
PIXELS_RAY, PYRAMIDS, NNDR_EPS, EPS_LMMIN = 64, 3, 0.55, 1e-10
# Semantics of the unqualified abs() in the reference's penalty wall.  The headline
# runs the reference as its author built it (int abs(int): the wall is outside the operating range and
# the normals converge); the same step under today's g++ (fabs: every synthetic feature ends on the
# wall, 12 deg median error in the oracle too) is measured next to it and reported as `penalty_fabs`.
PENALTY = 1          # FM3D_PENALTY_INT_ABS
PENALTY_ALT = 0      # FM3D_PENALTY_FABS
PENALTY_NAMES = {0: "fabs", 1: "int_abs", 2: "off"}
FLOP_PER_PIXEL_EVAL = 64.0    # algorithmic fp32 work of one pixel evaluation of the reference
FLOP_EXEC_VALUE = 58.0        # executed by normals_fast_kernel per pixel of a value-only pass (34 packed ops per pixel pair + the fp64 sum)
FLOP_EXEC_JAC = 113.0         # ... of a value + analytic-Jacobian pass (67 packed ops per pixel pair)
SEED = 1001
# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of normals_fast_kernel on this workload (ncu --set full)
NORMALS_TRAFFIC_BYTES = 0.7585e9
NORMALS_TRAFFIC_SOURCE = "profiles/r02y_normals_pp_kernel_ncu_raw_selected.csv (the kernel as benched, 11.85 ms under ncu): 195 MB read + 564 MB written, the ray scratch leaving L2"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# The contract is ONE JSON line on stdout.  Libraries (NCCL prints its version banner to stdout)
# must not share it: file descriptor 1 is pointed at stderr for the whole run and the JSON line
# goes to a private duplicate of the original stdout.
_JSON_OUT = None


def _claim_stdout():
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(obj):
    _JSON_OUT.write(json.dumps(obj) + "\n")
    _JSON_OUT.flush()


WORKLOAD, SCALING = "c2", "weak"


def select_workload(name, n_ranks):
    """`c2` (default, the driver's contract): BASELINE configs[1], 5 000 keypoints PER GPU (weak scaling).
    `c3`: BASELINE configs[2], one 4K pair with 20 000 keypoints in total, sharded over the ranks (strong scaling)."""
    global WIDTH, HEIGHT, N_KP, N_DISTRACT, SEED, WORKLOAD, SCALING
    if name == "c3":
        WIDTH, HEIGHT, SEED, WORKLOAD, SCALING = 3840, 2160, 1002, "c3", "strong"
        N_KP, N_DISTRACT = 20000 // n_ranks, 4000 // n_ranks


def make_workload(n_ranks, rank):
    """Global problem: one stereo pair with n_ranks*N_KP query keypoints; returns the global
    arrays (every rank generates the same ones; rank 0's copy is the one that gets broadcast)."""
    synth = importlib.import_module("3dfeaturematcher_b200.synth")
    t0 = time.time()
    case = synth.make_stereo_case(WIDTH, HEIGHT, N_KP * n_ranks, SEED, pixels_ray=PIXELS_RAY,
                                  n_distractors=N_DISTRACT * n_ranks)
    log(f"[rank {rank}] workload generated in {time.time() - t0:.1f}s")
    return case


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.samples, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            p = [x.strip() for x in s.split(",")]
            if len(p) < 6:
                continue
            try:
                sm.append(float(p[0]))
                smax = float(p[1])
            except ValueError:
                continue
            for nme, v in zip(names, p[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------- CPU arm
def cpu_reference(case, n_features_sample, threads, penalty=None):
    """Times the CPU port of the reference path (oracle/fm3d_oracle.c) on a bounded sample."""
    from oracle import oracle_c as orc
    cam = case["scene"].cam
    nq = min(N_KP, case["desc1"].shape[0])
    nt = nq + N_DISTRACT
    q, t = case["desc1"][:nq], case["desc2"][:nt]
    t0 = time.perf_counter()
    idx, dist = orc.knn2_f32(q, t, threads=threads)
    qi, ti, d = orc.nndr_filter(idx, dist, NNDR_EPS)
    t_match = time.perf_counter() - t0
    t0 = time.perf_counter()
    xyz_all, mask, xyz = orc.triangulate(cam.K, cam.dist, cam.g12, cam.z_min, cam.z_max, case["kp1"], case["kp2"], qi, ti)
    t_tri = time.perf_counter() - t0
    t0 = time.perf_counter()
    pyr1 = orc.build_pyramid(case["scene"].img1, PYRAMIDS)
    pyr2 = orc.build_pyramid(case["scene"].img2, PYRAMIDS)
    t_pyr = time.perf_counter() - t0
    rng = np.random.default_rng(7)
    sel = np.sort(rng.choice(xyz.shape[0], min(n_features_sample, xyz.shape[0]), replace=False))
    t0 = time.perf_counter()
    res = orc.optimize_normals(cam.K, cam.dist, cam.g12, cam.z_min, cam.z_max, case["scene"].img1, case["scene"].img2,
                               PYRAMIDS, xyz[sel], PIXELS_RAY, EPS_LMMIN, penalty_mode=PENALTY if penalty is None else penalty, threads=threads,
                               pyr1=pyr1, pyr2=pyr2)
    t_norm = time.perf_counter() - t0
    n_match = len(qi)
    per_feature = (t_match + t_tri + t_pyr) / max(n_match, 1) + t_norm / max(len(sel), 1)
    return {
        "features_per_s": 1.0 / per_feature, "n_match": n_match, "n_sample": int(len(sel)),
        "t_match_s": t_match, "t_triangulate_s": t_tri, "t_pyramid_s": t_pyr, "t_normals_sample_s": t_norm,
        "pixel_evals_sample": res["pixel_evals"],
    }


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    case = make_workload(1, 0)
    n_sample = max(16 * threads, 256)       # a seeded random subset of >= 256 features
    vals, t_all = [], []
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        r = cpu_reference(case, n_sample, threads)
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            vals.append(r["features_per_s"])
            t_all.append(dt)
    v = float(np.mean(vals))
    sample = (f"matching {N_KP}x{N_KP + N_DISTRACT} in full, normal optimisation on a seeded sample of {r['n_sample']} "
              f"of {r['n_match']} features, per-feature times summed")
    out = {
        "impl": "reference", "metric": "features/sec (match+triangulate+normal-opt)", "value": v, "unit": "features/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(t_all)),
        "higher_is_better": True, "scaling": SCALING, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(1, impl="reference"),
        "cpu_baseline": {"value": v, "unit": "features/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "features/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(out)


def workload_config(n_ranks, impl="fm3d"):
    cfg = {"workload": f"BASELINE configs[{2 if WORKLOAD == 'c3' else 1}]: {WIDTH}x{HEIGHT} synthetic stereo pair, {N_KP} SIFT-128 float keypoints per GPU "
                       f"(+{N_DISTRACT} distractors), pixelsRay {PIXELS_RAY}, pyramids {PYRAMIDS} (4 LM stages), NNDR {NNDR_EPS}",
           "penalty_mode": PENALTY_NAMES[PENALTY],
           "per_gpu_query_keypoints": N_KP, "global_query_keypoints": N_KP * n_ranks,
           "parallelism": f"keypoint shards x{n_ranks}"}
    if impl == "reference":
        cfg["normal_search"] = ("CPU port of the reference (oracle/fm3d_oracle.c): fp64 geometry, forward-difference Jacobian, "
                                "lmfit restatement, OpenMP over features")
    else:
        cfg["normal_search"] = ("fast kernel: fp32 offset-form geometry, analytic Jacobian, fp64 LM state (the fp64 / forward-difference "
                                "kernel is timed next to it as `faithful_fp64`)")
        cfg["l2_flush_between_steps"] = True
    return cfg


# ------------------------------------------------------------------------------- GPU arm
class PathState:
    """Device buffers of one workload and one pass of the hot path over them (all work on the context's stream)."""

    N_STAGE = 5
    STAGES = ("match+nndr", "triangulate", "pyramids", "normals", "gather")

    def __init__(self, torch, api, shard, ctx, stream, dev, case, world, rank, pixels_ray, pyramids):
        self.torch, self.shard, self.ctx, self.stream, self.dev = torch, shard, ctx, stream, dev
        self.case, self.world, self.rank, self.r, self.pyramids = case, world, rank, pixels_ray, pyramids
        self.cam = case["scene"].cam
        self.H, self.W = case["scene"].img1.shape
        nq_glob, self.nt = case["desc1"].shape[0], case["desc2"].shape[0]
        self.lo, self.hi = shard.shard_bounds(nq_glob, world, rank)
        self.nq = nq = self.hi - self.lo
        self.L1 = pyramids + 1
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()      # noqa: E731
        # pinned host copies (e2e arm)
        self.h_q, self.h_t = pin(case["desc1"][self.lo:self.hi]), pin(case["desc2"])
        self.h_kp1, self.h_kp2 = pin(case["kp1"][self.lo:self.hi]), pin(case["kp2"])
        self.h_img1, self.h_img2 = pin(case["scene"].img1), pin(case["scene"].img2)
        with torch.cuda.stream(stream):
            self.d_q, self.d_kp1 = self.h_q.to(dev), self.h_kp1.to(dev)
            # replicated inputs live in ONE allocation: one broadcast per step; only rank 0's copy is filled
            self.rep = shard.ReplicatedBuffer([("t", (self.nt, 128), torch.float32), ("kp2", (self.nt, 2), torch.float32),
                                               ("img1", (self.H, self.W), torch.uint8), ("img2", (self.H, self.W), torch.uint8)], dev)
            if rank == 0 or world == 1:
                self.rep["t"].copy_(self.h_t); self.rep["kp2"].copy_(self.h_kp2)
                self.rep["img1"].copy_(self.h_img1); self.rep["img2"].copy_(self.h_img2)
            i32, f32, f64 = torch.int32, torch.float32, torch.float64
            self.d_idx = torch.empty((nq, 2), dtype=i32, device=dev)
            self.d_dist = torch.empty((nq, 2), dtype=f32, device=dev)
            self.d_qi, self.d_ti = torch.zeros(nq, dtype=i32, device=dev), torch.zeros(nq, dtype=i32, device=dev)
            self.d_do = torch.zeros(nq, dtype=f32, device=dev)
            self.d_nm = torch.zeros(1, dtype=i32, device=dev)
            self.d_xyz_all = torch.empty((nq, 3), dtype=f64, device=dev)
            self.d_xyz = torch.zeros((nq, 3), dtype=f64, device=dev)
            self.d_mask = torch.empty(nq, dtype=torch.uint8, device=dev)
            self.d_src = torch.zeros(nq, dtype=i32, device=dev)
            self.d_ninl = torch.zeros(1, dtype=i32, device=dev)
            self.d_normals = torch.zeros((nq, 3), dtype=f64, device=dev)
            self.d_status = torch.zeros(nq, dtype=i32, device=dev)
            self.d_nfev = torch.zeros((nq, self.L1), dtype=i32, device=dev)
            self.d_npen = torch.zeros(nq, dtype=i32, device=dev)
            self.d_cost = torch.empty(nq, dtype=f64, device=dev)
            # ONE all-gather per step (libfm3d's own NCCL communicator): a block per rank with (query, train, distance) of its
            # matches and (match row, normal, status) of its inliers, written by fm3d_pack_shard_dev from the device-resident counts
            self.cap = cap = max(shard.shard_sizes(nq_glob, world))
            self.layout = api.shard_block_layout(cap)
            self.send = torch.zeros(self.layout.bytes, dtype=torch.uint8, device=dev)
            self.recv = torch.zeros((world, self.layout.bytes), dtype=torch.uint8, device=dev) if world > 1 else self.send.view(1, -1)
            self.api = api
        stream.synchronize()

    def bind(self):
        self.ctx.set_camera(self.cam.K, self.cam.dist, self.cam.z_min, self.cam.z_max)
        self.ctx.set_g12(self.cam.g12)

    def device_step(self, events=None, penalty=PENALTY):
        """One pass of the hot path on resident inputs."""
        torch, ctx, stream = self.torch, self.ctx, self.stream
        with torch.cuda.stream(stream):
            if events: events[0].record(stream)
            if self.world > 1:      # ONE broadcast: train descriptors + train keypoints + both frames live in one allocation
                ctx.broadcast_dev(self.rep.buf.data_ptr(), self.rep.buf.numel(), 0)
            ctx.match_knn2_f32_dev(self.d_q.data_ptr(), self.nq, self.rep["t"].data_ptr(), self.nt, 128, self.d_idx.data_ptr(), self.d_dist.data_ptr())
            ctx.nndr_filter_dev(self.d_idx.data_ptr(), self.d_dist.data_ptr(), self.nq, NNDR_EPS, self.d_qi.data_ptr(), self.d_ti.data_ptr(),
                                self.d_do.data_ptr(), self.d_nm.data_ptr())
            if events: events[1].record(stream)
            n_match = int(self.d_nm.item())          # sizes the following launches (4-byte D2H, as a host caller needs)
            ctx.triangulate_dev(self.d_kp1.data_ptr(), self.nq, self.rep["kp2"].data_ptr(), self.nt, self.d_qi.data_ptr(), self.d_ti.data_ptr(), n_match,
                                self.d_xyz_all.data_ptr(), self.d_mask.data_ptr(), self.d_xyz.data_ptr(), self.d_src.data_ptr(), self.d_ninl.data_ptr())
            if events: events[2].record(stream)
            ctx.set_images_dev(self.rep["img1"].data_ptr(), self.rep["img2"].data_ptr(), self.W, self.H, self.W, self.pyramids)
            if events: events[3].record(stream)
            n_inl = int(self.d_ninl.item())
            ctx.optimize_normals_dev(self.d_xyz.data_ptr(), n_inl, self.r, EPS_LMMIN, penalty, self.d_normals.data_ptr(),
                                     self.d_status.data_ptr(), self.d_nfev.data_ptr(), self.d_npen.data_ptr(), self.d_cost.data_ptr())
            if events: events[4].record(stream)
            ctx.pack_shard_dev(self.cap, self.rank, self.lo, self.d_nm.data_ptr(), self.d_ninl.data_ptr(), self.d_qi.data_ptr(),
                               self.d_ti.data_ptr(), self.d_do.data_ptr(), self.d_src.data_ptr(), self.d_normals.data_ptr(),
                               self.d_status.data_ptr(), self.send.data_ptr())
            if self.world > 1:
                ctx.allgather_dev(self.send.data_ptr(), self.recv.data_ptr(), self.layout.bytes)
            if events: events[5].record(stream)
        return n_match, n_inl

    def host_step(self):
        """The same pass through the host-buffer C-ABI entry points (copies inside)."""
        ctx = self.ctx
        if self.world > 1:
            ctx.broadcast_dev(self.rep.buf.data_ptr(), self.rep.buf.numel(), 0)
            ctx.sync()
        qi, ti, d = ctx.match_nndr(self.h_q.numpy(), self.h_t.numpy(), NNDR_EPS)
        xyz_all, mask, xyz, src = ctx.triangulate(self.h_kp1.numpy(), self.h_kp2.numpy(), qi, ti)
        ctx.set_images(self.h_img1.numpy(), self.h_img2.numpy(), self.pyramids)
        res = ctx.optimize_normals(xyz, self.r, EPS_LMMIN, PENALTY)
        h2d = (self.h_q.numel() + self.h_t.numel()) * 4 + (self.h_kp1.numel() + self.h_kp2.numel()) * 4 + qi.nbytes + ti.nbytes + \
            self.h_img1.numel() + self.h_img2.numel() + xyz.nbytes
        d2h = qi.nbytes + ti.nbytes + d.nbytes + xyz_all.nbytes + mask.nbytes + xyz.nbytes + src.nbytes + \
            res["normals"].nbytes + res["status"].nbytes + res["nfev"].nbytes + res["npenalty"].nbytes + res["cost"].nbytes
        return len(qi), xyz.shape[0], h2d, d2h

    def timed(self, steps, flush_l2, barrier, penalty=PENALTY, warm=2):
        """`steps` device passes with CUDA events per stage; returns (ms per stage summed over steps, last n_match, n_inl)."""
        torch = self.torch
        for _ in range(warm):
            self.device_step(penalty=penalty)
            flush_l2()
        barrier()
        ev = [[torch.cuda.Event(enable_timing=True) for _ in range(self.N_STAGE + 1)] for _ in range(steps)]
        for s in range(steps):
            n_match, n_inl = self.device_step(ev[s], penalty=penalty)
            flush_l2()                      # not between events 0..5 of a step: excluded from the step time
        barrier()
        stage_ms = np.zeros(self.N_STAGE)
        for s in range(steps):
            for j in range(self.N_STAGE):
                stage_ms[j] += ev[s][j].elapsed_time(ev[s][j + 1])
        return stage_ms, n_match, n_inl

    def check_gathered(self, penalty, n_query_sample=2048, n_normal_sample=64):
        """CONTENT check of what the step gathered from all ranks (rank 0, outside any timed region): the matches of a
        seeded sample of global query indices must be exactly the CPU oracle's (indices and distances bit for bit), and
        the normals / statuses of a seeded sample of gathered inliers, spread over all ranks, must agree with the CPU
        oracle's normal search on the oracle's own triangulation of those matches (status equal, <= 0.5 deg where the
        oracle stays off the penalty wall)."""
        from oracle import oracle_c as orc
        shard, case, cam = self.shard, self.case, self.cam
        out, hdr = self.api.unpack_shard_blocks(self.recv.cpu().numpy(), self.layout)
        counts = hdr[:, :2].tolist()
        los = [shard.shard_bounds(case["desc1"].shape[0], self.world, r)[0] for r in range(self.world)]
        if hdr[:, 2].tolist() != list(range(self.world)) or hdr[:, 4].tolist() != los:
            return f"block headers out of order: ranks {hdr[:, 2].tolist()} offsets {hdr[:, 4].tolist()}"
        q_glob = out["qidx"].astype(np.int64)               # fm3d_pack_shard_dev stores GLOBAL query indices
        ti, dd = out["tidx"], out["dist"]
        if not (np.diff(q_glob) > 0).all():
            return "gathered query indices are not ascending"
        rng = np.random.default_rng(11)
        nq_glob = case["desc1"].shape[0]
        sel = np.sort(rng.choice(nq_glob, min(n_query_sample, nq_glob), replace=False))
        o_idx, o_dist = orc.knn2_f32(case["desc1"][sel], case["desc2"], threads=os.cpu_count() or 1)
        oq, ot, od = orc.nndr_filter(o_idx, o_dist, NNDR_EPS)
        in_sel = np.isin(q_glob, sel)
        if not (np.array_equal(q_glob[in_sel], sel[oq]) and np.array_equal(ti[in_sel], ot) and np.array_equal(dd[in_sel], od)):
            return f"matches of the sampled queries differ from the oracle ({int(in_sel.sum())} gathered, {len(oq)} expected)"
        # normals: inlier k of rank r refers to match row src[k] of rank r
        src = out["src"].astype(np.int64)
        normals, status = out["normals"], out["status"]
        rank_of_inl = np.repeat(np.arange(self.world), [c[1] for c in counts])
        match_base = np.concatenate([[0], np.cumsum([c[0] for c in counts])])[:-1]
        row = src + match_base[rank_of_inl]                       # row in the gathered match list
        pick = np.sort(rng.choice(len(row), min(n_normal_sample, len(row)), replace=False))
        mq, mt = q_glob[row[pick]], ti[row[pick]].astype(np.int64)
        _, o_mask, o_xyz = orc.triangulate(cam.K, cam.dist, cam.g12, cam.z_min, cam.z_max, case["kp1"], case["kp2"],
                                           mq.astype(np.int32), mt.astype(np.int32))
        if not o_mask.astype(bool).all():
            return "a gathered inlier is outside the oracle's depth gate"
        o = orc.optimize_normals(cam.K, cam.dist, cam.g12, cam.z_min, cam.z_max, case["scene"].img1, case["scene"].img2,
                                 self.pyramids, o_xyz, self.r, EPS_LMMIN, penalty_mode=penalty, threads=os.cpu_count() or 1)
        if not np.array_equal(o["status"], status[pick]):
            return "statuses of the sampled normals differ from the oracle"
        interior = (o["status"] == 0) & (o["npenalty"] == 0)
        cosang = np.clip((normals[pick] * o["normals"]).sum(1), -1, 1)
        ang = np.degrees(np.arccos(cosang))
        far = interior & (ang > 0.5)
        better = 0
        if far.any():
            # allowed only where the device found a strictly lower value of the reference's own cost function (at 4K the
            # reference's forward-difference Jacobian stalls lmfit at the coarse levels: tests/test_gpu_parity_scale.py)
            def cost_at(nrm):
                pt = np.stack([np.arctan2(nrm[:, 1], nrm[:, 0]), np.arctan2(nrm[:, 2], np.hypot(nrm[:, 0], nrm[:, 1]))], 1)
                return orc.evaluate_cost(cam.K, cam.dist, cam.g12, cam.z_min, cam.z_max, case["scene"].img1, case["scene"].img2,
                                         self.pyramids, np.ascontiguousarray(o_xyz[far]), pt, self.r, 0, 2)[0]
            gc, oc = cost_at(normals[pick][far]), cost_at(o["normals"][far])
            if not (gc < oc * (1 - 1e-3)).all():
                return f"normals differ from the oracle by {ang[far].max():.3f} deg without a lower cost"
            better = int(far.sum())
        return {"result": "ok", "beyond_0p5_deg_with_lower_cost_than_the_oracle": better, "ranks_with_matches": int(sum(c[0] > 0 for c in counts)), "queries_checked": int(len(sel)),
                "matches_checked": int(in_sel.sum()), "normals_checked": int(len(pick)),
                "normals_interior": int(interior.sum()), "max_angle_deg_interior": float(ang[interior & ~far].max()) if (interior & ~far).any() else None,
                "ranks_in_normal_sample": int(len(np.unique(rank_of_inl[pick])))}


def oracle_agreement_fabs(st, n_sample=96):
    """`penalty_fabs.oracle_agreement`: a seeded sample of THIS rank's inliers of the fabs step against the CPU oracle in the
    same mode: statuses, and how many wall features end within 0.5 deg of the oracle's end state."""
    from oracle import oracle_c as orc
    cam, case = st.cam, st.case
    n_inl = int(st.d_ninl.item())
    xyz = st.d_xyz[:n_inl].cpu().numpy()
    normals, status = st.d_normals[:n_inl].cpu().numpy(), st.d_status[:n_inl].cpu().numpy()
    rng = np.random.default_rng(13)
    pick = np.sort(rng.choice(n_inl, min(n_sample, n_inl), replace=False))
    o = orc.optimize_normals(cam.K, cam.dist, cam.g12, cam.z_min, cam.z_max, case["scene"].img1, case["scene"].img2,
                             st.pyramids, np.ascontiguousarray(xyz[pick]), st.r, EPS_LMMIN, penalty_mode=PENALTY_ALT, threads=os.cpu_count() or 1)
    ok = o["status"] == 0
    wall = ok & (o["npenalty"] > 0)
    ang = np.degrees(np.arccos(np.clip((normals[pick] * o["normals"]).sum(1), -1, 1)))
    return {"features": int(len(pick)), "status_equal": bool(np.array_equal(o["status"], status[pick])),
            "wall_features": int(wall.sum()), "wall_within_0p5_deg": float((ang[wall] <= 0.5).mean()) if wall.any() else None,
            "wall_angle_deg_p50": float(np.median(ang[wall])) if wall.any() else None,
            "wall_angle_deg_p95": float(np.percentile(ang[wall], 95)) if wall.any() else None,
            "interior_features": int((ok & ~wall).sum()),
            "interior_max_angle_deg": float(ang[ok & ~wall].max()) if (ok & ~wall).any() else None}


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    api = importlib.import_module("3dfeaturematcher_b200.api")
    shard = importlib.import_module("3dfeaturematcher_b200.shard")
    synth = importlib.import_module("3dfeaturematcher_b200.synth")
    if not os.path.exists(api.LIB_PATH) and int(os.environ.get("LOCAL_RANK", "0")) == 0:
        # the library normally travels with the tree; nvcc is in the image if it does not
        importlib.import_module("3dfeaturematcher_b200.build").build()

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a B200: the fm3d path has no CPU implementation")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    ctx = api.Context(local_rank)
    info = ctx.device_info()
    if world > 1:
        # libfm3d's own communicator: rank 0 creates the NCCL id, torch.distributed carries the 128 bytes (plumbing only)
        ident = [api.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ident, src=0)
        ctx.comm_init_rank(ident[0], world, rank)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    case = make_workload(world, rank)
    st = PathState(torch, api, shard, ctx, stream, dev, case, world, rank, PIXELS_RAY, PYRAMIDS)
    st.bind()
    H, W, nq, nt, L1 = st.H, st.W, st.nq, st.nt, st.L1
    with torch.cuda.stream(stream):
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2
    stream.synchronize()
    n_stage = st.N_STAGE

    def barrier():
        stream.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def flush_l2():
        with torch.cuda.stream(stream):
            flush.fill_(1)

    # ---- clocks are sampled from the warm-up on (nvidia-smi needs ~0.1 s to deliver its first sample and the
    # timed region of the default run is only ~0.1 s long): every sample is taken under the same load
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.15)

    # ---- warm-up
    for _ in range(max(args.warmup, 3)):
        n_match, n_inl = st.device_step()
        flush_l2()
    barrier()

    # ---- timed region: device-resident arm
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(n_stage + 1)] for _ in range(args.steps)]
    k0, c0 = ctx.launch_counters()
    barrier()
    wall0 = time.perf_counter()
    for s in range(args.steps):
        n_match, n_inl = st.device_step(ev[s])
        flush_l2()                      # not between events 0..5 of a step: excluded from the step time
    barrier()
    wall_dev = time.perf_counter() - wall0
    k1, c1 = ctx.launch_counters()
    stage_ms = np.zeros(n_stage)
    for s in range(args.steps):
        for j in range(n_stage):
            stage_ms[j] += ev[s][j].elapsed_time(ev[s][j + 1])
    t_dev_ms = float(stage_ms.sum())
    clocks = sampler.stop() if rank == 0 else None
    ctx.sync()                          # also surfaces a TMA time-out of the device-resident normal search

    # ---- timed region: end-to-end arm (host buffers)
    for _ in range(2):
        st.host_step()
    barrier()
    t0 = time.perf_counter()
    for s in range(args.steps):
        hm, hi_, h2d, d2h = st.host_step()
    barrier()
    t_e2e = time.perf_counter() - t0

    # ---- what the last headline step gathered from all ranks, checked for CONTENT against the CPU oracle (rank 0)
    st.device_step()
    barrier()
    gather_check = st.check_gathered(PENALTY) if rank == 0 else None
    # ---- executed-work counters of the last headline launch
    stats = ctx.normals_stats()
    nfev_main = st.d_nfev[:n_inl].cpu().numpy().astype(np.int64)
    status_main = st.d_status[:n_inl].cpu().numpy()
    npen_main = st.d_npen[:n_inl].cpu().numpy()
    alt_steps = max(1, min(args.steps, 3))

    # ---- the same step under fabs semantics (the source as today's g++ compiles it)
    stage_alt, n_match_alt, n_inl_alt = st.timed(alt_steps, flush_l2, barrier, penalty=PENALTY_ALT)
    t_alt_ms, t_alt_norm_ms = float(stage_alt.sum()) / alt_steps, float(stage_alt[3]) / alt_steps
    stats_alt = ctx.normals_stats()
    nfev_alt = st.d_nfev[:n_inl_alt].cpu().numpy().astype(np.int64)
    npen_alt = st.d_npen[:n_inl_alt].cpu().numpy()
    fabs_agreement = oracle_agreement_fabs(st) if rank == 0 else None

    # ---- the same step with the faithful kernel (normals_fast = 0: fp64 geometry in the reference's evaluation order,
    # forward-difference Jacobian, one pass per lmfit evaluation group) -- the precision trade of the headline, in the open
    ctx.set_option("normals_fast", 0)
    try:
        stage_f64, n_match_f64, n_inl_f64 = st.timed(alt_steps, flush_l2, barrier, penalty=PENALTY, warm=1)
        nfev_f64 = st.d_nfev[:n_inl_f64].cpu().numpy().astype(np.int64)
        normals_f64 = st.d_normals[:n_inl_f64].cpu().numpy().copy()
    finally:
        ctx.set_option("normals_fast", 1)
    t_f64_ms, t_f64_norm_ms = float(stage_f64.sum()) / alt_steps, float(stage_f64[3]) / alt_steps

    # ---- after the path ("patch extraction separately"): frames -> K8 patches -> K9 SIFT descriptors of
    # this rank's refined features, device-resident, NOT part of `value`
    EPS_M, CM_PP = 0.16, 0.25           # build/settings.yml Neighborhoods: epsilon 0.16 m, cmPerPixel 0.25 -> 128 x 128 patches
    S_patch = api.patch_size(EPS_M, CM_PP)
    st.device_step()                    # headline normals again (the other passes overwrote them)
    stream.synchronize()
    fast_vs_f64 = None
    if n_inl_f64 == n_inl:
        nh = st.d_normals[:n_inl].cpu().numpy()
        ang = np.degrees(np.arccos(np.clip((nh * normals_f64).sum(1), -1, 1)))
        okm = (status_main == 0)
        fast_vs_f64 = {"angle_deg_p50": float(np.median(ang[okm])), "angle_deg_p99": float(np.percentile(ang[okm], 99)),
                       "angle_deg_max_wall_free": float(ang[okm & (npen_main == 0)].max()) if (okm & (npen_main == 0)).any() else None}
    d_xyz, d_normals, d_img1, d_img2 = st.d_xyz, st.d_normals, st.rep["img1"], st.rep["img2"]
    with torch.cuda.stream(stream):
        d_frames = torch.empty((max(n_inl, 1), 16), dtype=torch.float64, device=dev)
        d_patches = torch.empty((max(n_inl, 1), S_patch, S_patch), dtype=torch.uint8, device=dev)
        d_pdesc = torch.empty((max(n_inl, 1), 128), dtype=torch.float32, device=dev)
    gravity = np.array([0.006, 0.99992, -0.011])

    def after_path(events=None):
        with torch.cuda.stream(stream):
            if events: events[0].record(stream)
            ctx.feature_frames_dev(d_xyz.data_ptr(), d_normals.data_ptr(), n_inl, gravity, d_frames.data_ptr())
            if events: events[1].record(stream)
            ctx.extract_patches_dev(d_frames.data_ptr(), n_inl, EPS_M, CM_PP, d_patches.data_ptr(), None)
            if events: events[2].record(stream)
            ctx.describe_patches_sift_dev(d_patches.data_ptr(), n_inl, S_patch, d_pdesc.data_ptr())
            if events: events[3].record(stream)
    for _ in range(2):
        after_path()
        flush_l2()
    barrier()
    ev_after = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(alt_steps)]
    for s_ in range(alt_steps):
        after_path(ev_after[s_])
        flush_l2()
    barrier()
    after_ms = [sum(ev_after[s_][j].elapsed_time(ev_after[s_][j + 1]) for s_ in range(alt_steps)) / alt_steps for j in range(3)]
    pdesc_nonzero = int((d_pdesc[:n_inl] != 0).sum().item())

    # ---- before the path: what compareWithNNDR does first (descriptorsmatcher.cpp:110-115) with
    # DetectorType FAST + ExtractorType SIFT: K10 detection and K11 description of BOTH frames, device-resident,
    # NOT part of `value` (the headline matches the synthetic descriptors BASELINE.json names)
    FAST_T, KP_CAP = 20, 1 << 17
    with torch.cuda.stream(stream):
        d_kxy = torch.zeros((2, KP_CAP, 2), dtype=torch.float32, device=dev)
        d_kresp = torch.zeros((2, KP_CAP), dtype=torch.float32, device=dev)
        d_kn = torch.zeros(2, dtype=torch.int32, device=dev)
        d_k4 = torch.zeros((2, KP_CAP, 4), dtype=torch.float32, device=dev)
        d_kdesc = torch.zeros((2, KP_CAP, 128), dtype=torch.float32, device=dev)
        d_bdesc = torch.zeros((2, KP_CAP, 64), dtype=torch.uint8, device=dev)
        d_bkept = torch.zeros((2, KP_CAP), dtype=torch.uint8, device=dev)
        d_bang = torch.zeros((2, KP_CAP), dtype=torch.float32, device=dev)
        d_odesc = torch.zeros((2, KP_CAP, 32), dtype=torch.uint8, device=dev)
        d_okept = torch.zeros((2, KP_CAP), dtype=torch.uint8, device=dev)
    stream.synchronize()
    n_kp = [0, 0]

    def before_path(events=None, describe=True):
        with torch.cuda.stream(stream):
            if events: events[0].record(stream)
            for a, img in enumerate((d_img1, d_img2)):
                ctx.detect_fast_dev(img.data_ptr(), W, H, W, FAST_T, True, KP_CAP, d_kxy[a].data_ptr(), d_kresp[a].data_ptr(),
                                    d_kn[a:].data_ptr())
            if events: events[1].record(stream)
            if describe:
                for a, img in enumerate((d_img1, d_img2)):
                    if n_kp[a] > 0:
                        ctx.describe_keypoints_sift_dev(img.data_ptr(), W, H, W, d_k4[a].data_ptr(), n_kp[a], d_kdesc[a].data_ptr())
            if events: events[2].record(stream)
            if describe:
                for a, img in enumerate((d_img1, d_img2)):
                    if n_kp[a] > 0:
                        ctx.describe_keypoints_brisk_dev(img.data_ptr(), W, H, W, d_k4[a].data_ptr(), n_kp[a], True, d_bdesc[a].data_ptr(),
                                                         d_bkept[a].data_ptr(), d_bang[a].data_ptr())
            if events: events[3].record(stream)
            if describe:
                for a, img in enumerate((d_img1, d_img2)):
                    if n_kp[a] > 0:
                        ctx.describe_keypoints_orb_dev(img.data_ptr(), W, H, W, d_k4[a].data_ptr(), n_kp[a], d_odesc[a].data_ptr(), d_okept[a].data_ptr())
            if events: events[4].record(stream)
    before_path(describe=False)
    stream.synchronize()
    n_kp = [min(int(v), KP_CAP) for v in d_kn.cpu().tolist()]
    with torch.cuda.stream(stream):
        d_k4[:, :, :2] = d_kxy
        d_k4[:, :, 2] = 7.0             # cv::FastFeatureDetector: KeyPoint(x, y, 7.f, -1, score)
        d_k4[:, :, 3] = -1.0
    for _ in range(2):
        before_path()
        flush_l2()
    barrier()
    ev_before = [[torch.cuda.Event(enable_timing=True) for _ in range(5)] for _ in range(alt_steps)]
    for s_ in range(alt_steps):
        before_path(ev_before[s_])
        flush_l2()
    barrier()
    before_ms = [sum(ev_before[s_][j].elapsed_time(ev_before[s_][j + 1]) for s_ in range(alt_steps)) / alt_steps for j in range(4)]
    brisk_kept = [int(d_bkept[a, :n_kp[a]].sum().item()) for a in range(2)]
    orb_kept = [int(d_okept[a, :n_kp[a]].sum().item()) for a in range(2)]
    kdesc_rows_nonzero = [int((d_kdesc[a, :n_kp[a]] != 0).any(1).sum().item()) for a in range(2)]
    del d_kdesc, d_bdesc, d_odesc, d_patches
    # DetectorType SIFT + ExtractorType SIFT (K14 + K11 on the pyramid layers): host buffers in and out, wall clock (rank 0)
    sift_front = None
    if rank == 0:
        h1 = case["scene"].img1
        ctx.detect_sift(h1)
        t0 = time.perf_counter()
        ksift = ctx.detect_sift(h1)
        t1 = time.perf_counter()
        dsift = ctx.describe_keypoints_sift_oct(h1, ksift[:, :4].astype(np.float32), ksift[:, 5].astype(np.int32))
        t2 = time.perf_counter()
        ctx.detect_and_describe_sift(h1)
        t3 = time.perf_counter()
        kboth, dboth = ctx.detect_and_describe_sift(h1)
        t4 = time.perf_counter()
        sift_front = {"keypoints_frame1": int(len(ksift)), "ms_detect_frame1_host_to_host": 1e3 * (t1 - t0),
                      "ms_describe_frame1_host_to_host": 1e3 * (t2 - t1), "descriptor_rows_nonzero": int((dsift != 0).any(1).sum()),
                      "ms_detect_and_describe_on_one_pyramid_host_to_host": 1e3 * (t4 - t3),
                      "one_pyramid_equals_two_calls": bool(np.array_equal(kboth, ksift) and np.array_equal(dboth, dsift))}

    # ---- the north-star configuration in the same run, at the same N: BASELINE configs[2], ONE 4K pair with 20 000 query
    # keypoints in total sharded over the ranks (strong scaling), pixelsRay 64, pyramids 3
    c3 = None
    if WORKLOAD == "c2" and not args.no_c3:
        t0 = time.time()
        case3 = synth.make_stereo_case(3840, 2160, 20000, 1002, pixels_ray=PIXELS_RAY, n_distractors=4000)
        log(f"[rank {rank}] C3 workload generated in {time.time() - t0:.1f}s")
        st3 = PathState(torch, api, shard, ctx, stream, dev, case3, world, rank, PIXELS_RAY, PYRAMIDS)
        st3.bind()
        c3_steps = max(1, min(args.steps, 3))
        stage3, n_match3, n_inl3 = st3.timed(c3_steps, flush_l2, barrier, penalty=PENALTY, warm=2)
        ctx.sync()
        c3_check = st3.check_gathered(PENALTY, n_query_sample=1024, n_normal_sample=48) if rank == 0 else None
        tm3 = torch.tensor([float(stage3.sum())], dtype=torch.float64, device=dev)
        f3 = torch.tensor([float(n_match3)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tm3, op=dist.ReduceOp.MAX)
            dist.all_reduce(f3, op=dist.ReduceOp.SUM)
        c3 = {"workload": "BASELINE configs[2]: ONE 3840x2160 synthetic stereo pair, 20000 SIFT-128 query keypoints (+4000 distractors) in "
                          f"total, sharded over {world} GPU(s), pixelsRay 64, pyramids 3, NNDR 0.55, penalty int_abs",
              "scaling": "strong", "n_gpus": world, "steps": c3_steps,
              "value": float(f3[0]) * c3_steps / (float(tm3[0]) * 1e-3), "unit": "features/s",
              "ms_per_pair": float(tm3[0]) / c3_steps, "matched_features": float(f3[0]),
              "stage_ms_rank0": {k: float(v) / c3_steps for k, v in zip(PathState.STAGES, stage3)},
              "gather_check": c3_check}
        st.bind()
        del st3, case3

    # ---- per-stage times of every rank (the stage that follows a slow rank's stage absorbs the skew: at N > 1 the `gather`
    # stage of a fast rank is mostly the wait for the slowest rank's normal search, not the collective)
    stage_all = torch.zeros((world, n_stage), dtype=torch.float64, device=dev)
    stage_all[rank] = torch.from_numpy(stage_ms / args.steps).to(dev)
    if world > 1:
        dist.all_reduce(stage_all)
    stage_all = stage_all.cpu().numpy()
    # ---- reduce over ranks: max time, summed features
    tm = torch.tensor([t_dev_ms, t_e2e * 1e3, t_alt_ms, t_f64_ms], dtype=torch.float64, device=dev)
    feats = torch.tensor([float(n_match), float(hm), float(n_match_alt), float(n_match_f64)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        dist.all_reduce(feats, op=dist.ReduceOp.SUM)
    t_dev_ms, t_e2e_ms, t_alt_ms, t_f64_ms = float(tm[0]), float(tm[1]), float(tm[2]), float(tm[3])
    tot_feat, tot_feat_e2e, tot_feat_alt, tot_feat_f64 = float(feats[0]), float(feats[1]), float(feats[2]), float(feats[3])

    if rank == 0:
        nfev, status, npen = nfev_main, status_main, npen_main
        # work of the normal search (rank 0's launch).  Executed: pixel evaluations the kernel made
        # (value-only passes, value + analytic Jacobian passes).  Reference-equivalent: sum over
        # features and levels of nfev * m, nfev = evaluations lmfit would have made with its
        # forward-difference Jacobian; every bench feature has the full disc (m = 12853 at r = 64).
        m_disc = sum(2 * int(np.floor(np.sqrt(PIXELS_RAY ** 2 - j * j))) + 1 for j in range(-PIXELS_RAY, PIXELS_RAY + 1))
        pixel_evals_ref = float(nfev.sum()) * m_disc
        t_norm_s = stage_ms[3] / args.steps * 1e-3
        sm_max = (clocks or {}).get("sm_max_mhz") or 1965.0
        # nominal SMs*128*2*f; measured on this pool with tools/micro/ffma2_rate.cu: 73.96 TFLOP/s at 1965 MHz
        # (profiles/r01_fp32_peak_microbench.txt), i.e. 99.3 % of nominal -> the measured figure, scaled by the clock
        fp32_peak = 0.9934 * info["sm_count"] * 128 * 2 * sm_max * 1e6 / 1e12
        flops_exec = stats["pixel_evals_value"] * FLOP_EXEC_VALUE + stats["pixel_evals_jacobian"] * FLOP_EXEC_JAC
        achieved_exec = flops_exec / t_norm_s / 1e12
        # W = sum over features and levels of nfev * m pixel evaluations of 64 flop each
        achieved = pixel_evals_ref * FLOP_PER_PIXEL_EVAL / t_norm_s / 1e12
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        bf16_peak = peaks.get("bf16_tflops", 1590.0)
        pyr_bytes = sum(2 * 1.25 * (W >> l) * (H >> l) for l in range(PYRAMIDS))
        match_pairs = float(nq) * nt
        n_pass = stats["passes_value"] + stats["passes_jacobian"] + stats["passes_fused"]
        cyc = {k: stats[k] for k in ("cycles_pixels", "cycles_barrier", "cycles_serial", "cycles_lm", "cycles_publish")}
        cyc_tot = max(1, cyc["cycles_pixels"] + cyc["cycles_barrier"] + cyc["cycles_serial"])
        roofline = {
            "kernel": "normals_pp_kernel (K6: LM normal search, persistent CTAs, four features per SM: two per eight warps taking turns)",
            "bound": "fp32", "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s", "frac": achieved / fp32_peak,
            # dram__bytes_read.sum + dram__bytes_write.sum of one launch of this workload under `ncu --set full`
            "traffic": NORMALS_TRAFFIC_BYTES, "traffic_unit": "bytes per launch (ncu --set full, " + NORMALS_TRAFFIC_SOURCE + ")",
            "traffic_algorithmic": n_inl * 44e3,
            "note": "compute-bound kernel; achieved = ALGORITHMIC flops of SURVEY 8(d) (64 flop x sum over features and levels of "
                    "nfev x m pixel evaluations, nfev = what lmfit evaluates with its forward-difference Jacobian: 1 per trial, "
                    "2 per Jacobian -- including the evaluations the kernel answers from memory because their fp32 coefficients equal those of a "
                    "pass already evaluated: `trials_memoized`) / CUDA-event time of the launch; peak = measured FFMA rate (tools/micro/ffma2_rate.cu: "
                    "73.96 TFLOP/s = 99.3 % of SMs*128*2*f_max; MEASURED_PEAKS.json has no fp32 figure).  `executed` is what the "
                    "kernel itself issues: 58 flop per pixel of a value-only pass, 113 per pixel of a value + analytic-Jacobian "
                    "pass, counted by the kernel (rounds 1 and 2a reported that figure as `frac`, with 64 / 152).  HBM traffic is "
                    "compulsory only (~44 KB/feature) and DRAM throughput ~0 (ncu, profiles/)",
            "achieved_executed": achieved_exec, "frac_executed": achieved_exec / fp32_peak,
            "pixel_evals_value": stats["pixel_evals_value"], "pixel_evals_jacobian": stats["pixel_evals_jacobian"],
            "passes": n_pass, "passes_fused": stats["passes_fused"], "fused_accepted": stats["fused_accepted"],
            "passes_global_taps": stats["passes_slow"], "trials_memoized": stats.get("trials_memoized"),
            "reference_equivalent_pixel_evals": pixel_evals_ref,
            "reference_equivalent_tflops": achieved,
            "frac_reference_equivalent": achieved / fp32_peak,
            "thread0_cycle_share": {"pixel_loop": cyc["cycles_pixels"] / cyc_tot, "serial_lm_step": cyc["cycles_serial"] / cyc_tot},
            # the packed FMAs of the pixel loop read {64-bit, 32-bit broadcast constant, 64-bit} register sources: that form
            # sustains 54.6 TFLOP/s on this pool (three distinct 64-bit sources: 47.1), not the 73.96 of re-used operands --
            # vector register-file bandwidth (tools/micro/ffma2_operands.cu, profiles/r01f_ffma2_operand_microbench.txt)
            "peak_operand_limited": 54.63 * sm_max / 1965.0,
            "frac_executed_of_operand_limited_peak": achieved_exec / (54.63 * sm_max / 1965.0),
            "ms_per_launch": t_norm_s * 1e3,
            "hbm_compulsory_gbs": (n_inl * 44e3 / t_norm_s) / 1e9, "hbm_peak_gbs": hbm_peak,
        }
        roofline_other = {
            "matcher_tc": {"bound": "tensor", "achieved": match_pairs * 256 / (stage_ms[0] / args.steps * 1e-3) / 1e12,
                           "peak": bf16_peak, "unit": "TFLOP/s", "ms": stage_ms[0] / args.steps,
                           "note": "includes operand re-tiling, NNDR filter and (N>1) the NCCL broadcast"},
            "pyrdown": {"bound": "hbm", "achieved": pyr_bytes / (stage_ms[2] / args.steps * 1e-3) / 1e9, "peak": hbm_peak,
                        "unit": "GB/s", "ms": stage_ms[2] / args.steps, "note": "0.9 MB images: launch-bound"},
        }
        ncpu = os.cpu_count() or 1
        cpu = cpu_reference(case, max(16 * ncpu, 256), ncpu) if world == 1 else None
        cpu_alt = cpu_reference(case, max(8 * ncpu, 128), ncpu, penalty=PENALTY_ALT) if world == 1 else None
        # the reference is single-threaded -- the same port on ONE host thread, on a smaller bounded sample
        cpu_1t = cpu_reference(case, 32, 1) if world == 1 else None
        value = tot_feat * args.steps / (t_dev_ms * 1e-3)
        out = {
            "metric": "features/sec (match+triangulate+normal-opt)", "value": value, "unit": "features/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": t_dev_ms / args.steps,
            "higher_is_better": True, "scaling": SCALING, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(world),
            "e2e": {"value": tot_feat_e2e * args.steps / (t_e2e_ms * 1e-3), "unit": "features/s",
                    "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "ms_per_step": t_e2e_ms / args.steps},
            "gpu_launches": int(k1 - k0),
            "clocks": clocks,
            "roofline": roofline,
            "roofline_other": roofline_other,
            "cpu_baseline": None if cpu is None else {
                "value": cpu["features_per_s"], "unit": "features/s", "cores": ncpu, "kind": "port",
                "sample": f"full {N_KP}x{N_KP + N_DISTRACT} matching + normal optimisation of {cpu['n_sample']} seeded features "
                          f"of {cpu['n_match']}, per-feature times summed", "detail": cpu,
                "single_thread": {"value": cpu_1t["features_per_s"], "unit": "features/s", "cores": 1,
                                  "sample": f"the same port on ONE host thread (the reference is single-threaded, SURVEY 8d-i): full matching + "
                                            f"normal optimisation of {cpu_1t['n_sample']} seeded features", "detail": cpu_1t}},
            "stage_ms_per_step": {k: float(v) / args.steps for k, v in zip(PathState.STAGES, stage_ms)},
            "stage_ms_per_step_min_max_over_ranks": {k: [float(stage_all[:, j].min()), float(stage_all[:, j].max())]
                                                     for j, k in enumerate(PathState.STAGES)},
            "features_per_step": {"matches": tot_feat, "inliers_rank0": int(n_inl), "ok_rank0": int((status == 0).sum()),
                                  "wall_touching_rank0": int((npen > 0).sum()), "nfev_mean_per_level": nfev.mean(0).tolist()},
            "multi_gpu_check": gather_check["result"] if isinstance(gather_check, dict) else str(gather_check),
            "multi_gpu_check_detail": gather_check,
            "faithful_fp64": {
                "note": "the same step with normals_fast = 0 (fm3d_normals.cu): fp64 geometry in the reference's evaluation order, "
                        "forward-difference Jacobian, one pass per lmfit evaluation group; dtype f64",
                "value": tot_feat_f64 / (t_f64_ms * 1e-3), "unit": "features/s", "ms_per_step": t_f64_ms, "normals_ms": t_f64_norm_ms,
                "nfev_mean_per_level": nfev_f64.mean(0).tolist(), "headline_vs_faithful_normals": fast_vs_f64},
            "c3_strong": c3,
            "penalty_fabs": {
                "note": "the same step with the penalty wall as today's g++ compiles it (abs == fabs): every synthetic "
                        "feature ends on the wall in the reference too (SURVEY fact 11); reported, not the headline",
                "value": tot_feat_alt / (t_alt_ms * 1e-3), "unit": "features/s", "ms_per_step": t_alt_ms,
                "normals_ms": t_alt_norm_ms, "wall_touching_rank0": int((npen_alt > 0).sum()),
                "nfev_mean_per_level": nfev_alt.mean(0).tolist(),
                "passes": stats_alt["passes_value"] + stats_alt["passes_jacobian"] + stats_alt["passes_fused"],
                "oracle_agreement": fabs_agreement,
                "cpu_baseline_features_per_s": None if cpu_alt is None else cpu_alt["features_per_s"]},
            "after_path_rank0": {
                "note": "frames -> rectified patches (K8) -> SIFT descriptors of the patches (K9, extractDescriptorsFromPatches) for "
                        "rank 0's refined features, device-resident, CUDA events; not part of `value`",
                "patch_edge": S_patch, "features": int(n_inl),
                "ms": {"frames": after_ms[0], "patches": after_ms[1], "patch_descriptors": after_ms[2]},
                "patches_per_s": n_inl / (after_ms[1] * 1e-3) if after_ms[1] > 0 else None,
                "patch_descriptors_per_s": n_inl / (after_ms[2] * 1e-3) if after_ms[2] > 0 else None,
                "patches_hbm_write_gbs": n_inl * S_patch * S_patch / (after_ms[1] * 1e-3) / 1e9 if after_ms[1] > 0 else None,
                "descriptor_values_nonzero": pdesc_nonzero},
            "before_path_rank0": {
                "note": "what compareWithNNDR does first with DetectorType FAST + ExtractorType SIFT (descriptorsmatcher.cpp:110-115): "
                        "FAST-9-16 detection (K10), SIFT description at the detected keypoints (K11) and, separately, their BRISK "
                        "and ORB descriptions (K12 / K13, ExtractorType BRISK / ORB) of BOTH frames, "
                        "device-resident, CUDA events; not part of `value`",
                "fast_threshold": FAST_T, "keypoints": n_kp,
                "ms": {"detect_both_frames": before_ms[0], "describe_both_frames": before_ms[1],
                       "describe_brisk_both_frames": before_ms[2], "describe_orb_both_frames": before_ms[3]},
                "brisk_keypoints_kept": brisk_kept, "orb_keypoints_kept": orb_kept,
                "orb_descriptors_per_s": sum(n_kp) / (before_ms[3] * 1e-3) if before_ms[3] > 0 else None,
                "brisk_descriptors_per_s": sum(n_kp) / (before_ms[2] * 1e-3) if before_ms[2] > 0 else None,
                "detect_hbm_gbs": 2 * W * H / (before_ms[0] * 1e-3) / 1e9 if before_ms[0] > 0 else None,
                "descriptors_per_s": sum(n_kp) / (before_ms[1] * 1e-3) if before_ms[1] > 0 else None,
                "descriptor_rows_nonzero": kdesc_rows_nonzero,
                "sift_detector": sift_front},
            "wall_s_device_arm": wall_dev, "gpu": info["name"],
        }
        emit(out)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="fm3d", choices=["fm3d", "reference"])
    ap.add_argument("--no-c3", action="store_true", help="skip the c3_strong block (BASELINE configs[2] in the same run)")
    ap.add_argument("--workload", default="c2", choices=["c2", "c3"],
                    help="c2: BASELINE configs[1], weak scaling (default, the driver's contract); c3: configs[2], 4K / 20k keypoints sharded")
    args = ap.parse_args()
    select_workload(args.workload, 1 if args.impl == "reference" else int(os.environ.get("WORLD_SIZE", "1")))
    _claim_stdout()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
