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
FLOP_PER_PIXEL_EVAL = 64.0    # algorithmic fp32 work of one pixel evaluation
FLOP_PER_PIXEL_JAC = 152.0    # the same plus its analytic derivatives w.r.t. (phi, theta)
SEED = 1001


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
        "config": workload_config(1),
        "cpu_baseline": {"value": v, "unit": "features/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "features/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(out)


def workload_config(n_ranks):
    return {"workload": f"BASELINE configs[{2 if WORKLOAD == 'c3' else 1}]: {WIDTH}x{HEIGHT} synthetic stereo pair, {N_KP} SIFT-128 float keypoints per GPU "
                        f"(+{N_DISTRACT} distractors), pixelsRay {PIXELS_RAY}, pyramids {PYRAMIDS} (4 LM stages), NNDR {NNDR_EPS}",
            "penalty_mode": PENALTY_NAMES[PENALTY], "normal_search": "fast kernel: fp32 offset-form geometry, analytic Jacobian, fp64 LM state",
            "per_gpu_query_keypoints": N_KP, "global_query_keypoints": N_KP * n_ranks,
            "parallelism": f"keypoint shards x{n_ranks}", "l2_flush_between_steps": True}


# ------------------------------------------------------------------------------- GPU arm
def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    api = importlib.import_module("3dfeaturematcher_b200.api")
    shard = importlib.import_module("3dfeaturematcher_b200.shard")
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
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    case = make_workload(world, rank)
    cam = case["scene"].cam
    ctx.set_camera(cam.K, cam.dist, cam.z_min, cam.z_max)
    ctx.set_g12(cam.g12)
    H, W = case["scene"].img1.shape
    nq_glob, nt = case["desc1"].shape[0], case["desc2"].shape[0]
    lo, hi = shard.shard_bounds(nq_glob, world, rank)
    nq = hi - lo
    L1 = PYRAMIDS + 1

    # ---- pinned host copies (e2e arm) and resident device copies (value arm)
    def pin(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t
    h_q, h_t = pin(case["desc1"][lo:hi]), pin(case["desc2"])
    h_kp1, h_kp2 = pin(case["kp1"][lo:hi]), pin(case["kp2"])
    h_img1, h_img2 = pin(case["scene"].img1), pin(case["scene"].img2)
    with torch.cuda.stream(stream):
        d_q, d_kp1 = h_q.to(dev), h_kp1.to(dev)
        # replicated inputs: only rank 0's copy is meaningful before the broadcast
        d_t, d_kp2, d_img1, d_img2 = h_t.to(dev), h_kp2.to(dev), h_img1.to(dev), h_img2.to(dev)
        if rank != 0 and world > 1:
            for x in (d_t, d_kp2, d_img1, d_img2):
                x.zero_()
        d_idx = torch.empty((nq, 2), dtype=torch.int32, device=dev)
        d_dist = torch.empty((nq, 2), dtype=torch.float32, device=dev)
        d_qi = torch.empty(nq, dtype=torch.int32, device=dev)
        d_ti = torch.empty(nq, dtype=torch.int32, device=dev)
        d_do = torch.empty(nq, dtype=torch.float32, device=dev)
        d_nm = torch.zeros(1, dtype=torch.int32, device=dev)
        d_xyz_all = torch.empty((nq, 3), dtype=torch.float64, device=dev)
        d_xyz = torch.empty((nq, 3), dtype=torch.float64, device=dev)
        d_mask = torch.empty(nq, dtype=torch.uint8, device=dev)
        d_src = torch.empty(nq, dtype=torch.int32, device=dev)
        d_ninl = torch.zeros(1, dtype=torch.int32, device=dev)
        d_normals = torch.empty((nq, 3), dtype=torch.float64, device=dev)
        d_status = torch.empty(nq, dtype=torch.int32, device=dev)
        d_nfev = torch.zeros((nq, L1), dtype=torch.int32, device=dev)
        d_npen = torch.zeros(nq, dtype=torch.int32, device=dev)
        d_cost = torch.empty(nq, dtype=torch.float64, device=dev)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2
    stream.synchronize()

    n_stage = 5
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(n_stage + 1)] for _ in range(args.steps)]

    def device_step(events=None, penalty=PENALTY):
        """One pass of the hot path on resident inputs (all work on the context's stream)."""
        with torch.cuda.stream(stream):
            if events: events[0].record(stream)
            if world > 1:
                shard.broadcast_([d_t, d_kp2, d_img1, d_img2], src=0)
            ctx.match_knn2_f32_dev(d_q.data_ptr(), nq, d_t.data_ptr(), nt, 128, d_idx.data_ptr(), d_dist.data_ptr())
            ctx.nndr_filter_dev(d_idx.data_ptr(), d_dist.data_ptr(), nq, NNDR_EPS, d_qi.data_ptr(), d_ti.data_ptr(),
                                d_do.data_ptr(), d_nm.data_ptr())
            if events: events[1].record(stream)
            n_match = int(d_nm.item())          # sizes the following launches (4-byte D2H, as a host caller needs)
            ctx.triangulate_dev(d_kp1.data_ptr(), nq, d_kp2.data_ptr(), nt, d_qi.data_ptr(), d_ti.data_ptr(), n_match,
                                d_xyz_all.data_ptr(), d_mask.data_ptr(), d_xyz.data_ptr(), d_src.data_ptr(), d_ninl.data_ptr())
            if events: events[2].record(stream)
            ctx.set_images_dev(d_img1.data_ptr(), d_img2.data_ptr(), W, H, W, PYRAMIDS)
            if events: events[3].record(stream)
            n_inl = int(d_ninl.item())
            ctx.optimize_normals_dev(d_xyz.data_ptr(), n_inl, PIXELS_RAY, EPS_LMMIN, penalty, d_normals.data_ptr(),
                                     d_status.data_ptr(), d_nfev.data_ptr(), d_npen.data_ptr(), d_cost.data_ptr())
            if events: events[4].record(stream)
            gathered = None
            if world > 1:
                # one collective: (global query index, train index, distance) of the matches, normal and
                # status of the inliers of every shard
                gathered = shard.gather_packed([d_qi + lo, d_ti, d_do, d_normals, d_status],
                                               [n_match, n_match, n_match, n_inl, n_inl], nq, unpack=False)
            if events: events[5].record(stream)
        device_step.gathered = gathered if world > 1 else None
        return n_match, n_inl

    def host_step():
        """The same pass through the host-buffer C-ABI entry points (copies inside)."""
        if world > 1:
            with torch.cuda.stream(stream):
                shard.broadcast_([d_t, d_kp2, d_img1, d_img2], src=0)
            stream.synchronize()
        qi, ti, d = ctx.match_nndr(h_q.numpy(), h_t.numpy(), NNDR_EPS)
        xyz_all, mask, xyz, src = ctx.triangulate(h_kp1.numpy(), h_kp2.numpy(), qi, ti)
        ctx.set_images(h_img1.numpy(), h_img2.numpy(), PYRAMIDS)
        res = ctx.optimize_normals(xyz, PIXELS_RAY, EPS_LMMIN, PENALTY)
        h2d = (h_q.numel() + h_t.numel()) * 4 + (h_kp1.numel() + h_kp2.numel()) * 4 + qi.nbytes + ti.nbytes + \
            h_img1.numel() + h_img2.numel() + xyz.nbytes
        d2h = qi.nbytes + ti.nbytes + d.nbytes + xyz_all.nbytes + mask.nbytes + xyz.nbytes + src.nbytes + \
            res["normals"].nbytes + res["status"].nbytes + res["nfev"].nbytes + res["npenalty"].nbytes + res["cost"].nbytes
        return len(qi), xyz.shape[0], h2d, d2h

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
        n_match, n_inl = device_step()
        flush_l2()
    barrier()

    # ---- timed region: device-resident arm
    k0, c0 = ctx.launch_counters()
    t_dev_ms = 0.0
    barrier()
    wall0 = time.perf_counter()
    for s in range(args.steps):
        n_match, n_inl = device_step(ev[s])
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

    # ---- timed region: end-to-end arm (host buffers)
    for _ in range(2):
        host_step()
    barrier()
    t0 = time.perf_counter()
    for s in range(args.steps):
        hm, hi_, h2d, d2h = host_step()
    barrier()
    t_e2e = time.perf_counter() - t0

    if world > 1:   # the gathered buffer holds every rank's matches and normals: unpack once, outside the timed region
        (g_q, g_t, g_d, g_n, g_s), g_counts = shard.unpack_packed(*device_step.gathered)
        assert g_q.shape[0] == sum(c[0] for c in g_counts) and g_n.shape == (sum(c[3] for c in g_counts), 3)
    # ---- executed-work counters of the last headline launch, then the same step under fabs semantics
    stats = ctx.normals_stats()
    nfev_main = d_nfev[:n_inl].cpu().numpy().astype(np.int64)
    status_main = d_status[:n_inl].cpu().numpy()
    npen_main = d_npen[:n_inl].cpu().numpy()
    alt_steps = max(1, min(args.steps, 3))
    for _ in range(2):
        device_step(penalty=PENALTY_ALT)
        flush_l2()
    barrier()
    ev_alt = [[torch.cuda.Event(enable_timing=True) for _ in range(n_stage + 1)] for _ in range(alt_steps)]
    for s_ in range(alt_steps):
        n_match_alt, n_inl_alt = device_step(ev_alt[s_], penalty=PENALTY_ALT)
        flush_l2()
    barrier()
    t_alt_ms = sum(ev_alt[s_][0].elapsed_time(ev_alt[s_][n_stage]) for s_ in range(alt_steps)) / alt_steps
    t_alt_norm_ms = sum(ev_alt[s_][3].elapsed_time(ev_alt[s_][4]) for s_ in range(alt_steps)) / alt_steps
    stats_alt = ctx.normals_stats()
    nfev_alt = d_nfev[:n_inl_alt].cpu().numpy().astype(np.int64)
    npen_alt = d_npen[:n_inl_alt].cpu().numpy()

    # ---- after the path ("patch extraction separately"): frames -> K8 patches -> K9 SIFT descriptors of
    # this rank's refined features, device-resident, NOT part of `value`
    EPS_M, CM_PP = 0.16, 0.25           # build/settings.yml Neighborhoods: epsilon 0.16 m, cmPerPixel 0.25 -> 128 x 128 patches
    S_patch = api.patch_size(EPS_M, CM_PP)
    device_step()                       # headline normals again (the fabs pass overwrote them)
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

    # ---- reduce over ranks: max time, summed features
    tm = torch.tensor([t_dev_ms, t_e2e * 1e3, t_alt_ms], dtype=torch.float64, device=dev)
    feats = torch.tensor([float(n_match), float(hm), float(n_match_alt)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        dist.all_reduce(feats, op=dist.ReduceOp.SUM)
    t_dev_ms, t_e2e_ms, t_alt_ms = float(tm[0]), float(tm[1]), float(tm[2])
    tot_feat, tot_feat_e2e, tot_feat_alt = float(feats[0]), float(feats[1]), float(feats[2])

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
        flops_exec = stats["pixel_evals_value"] * FLOP_PER_PIXEL_EVAL + stats["pixel_evals_jacobian"] * FLOP_PER_PIXEL_JAC
        achieved = flops_exec / t_norm_s / 1e12
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
            "kernel": "normals_fast_kernel<true> (K6: LM normal search, one persistent CTA per SM, one feature per CTA at a time)",
            "bound": "fp32", "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s", "frac": achieved / fp32_peak,
            # dram__bytes_read.sum + dram__bytes_write.sum of one launch of this workload under `ncu --set full`
            # (profiles/r01f_normals_fast_kernel_ncu_raw_selected.csv: 199 MB + 691 MB, the L2-resident ray scratch
            # being written back); 0.6 % of DRAM throughput -- the kernel is not HBM-bound
            "traffic": 0.8902e9, "traffic_unit": "bytes per launch (ncu, profiles/r01f_*)",
            "note": "compute-bound kernel (SURVEY 8d): algorithmic flops = 64 x value-only pixel evaluations + 152 x "
                    "value+analytic-Jacobian pixel evaluations executed (counted by the kernel) / CUDA-event time; peak = "
                    "measured FFMA rate (tools/micro/ffma2_rate.cu: 73.96 TFLOP/s = 99.3 % of SMs*128*2*f_max; "
                    "MEASURED_PEAKS.json has no fp32 figure); HBM traffic is compulsory only "
                    "(~44 KB/feature) and DRAM throughput ~0 (ncu, profiles/)",
            "pixel_evals_value": stats["pixel_evals_value"], "pixel_evals_jacobian": stats["pixel_evals_jacobian"],
            "passes": n_pass, "passes_fused": stats["passes_fused"], "fused_accepted": stats["fused_accepted"],
            "passes_global_taps": stats["passes_slow"],
            "reference_equivalent_pixel_evals": pixel_evals_ref,
            "reference_equivalent_tflops": pixel_evals_ref * FLOP_PER_PIXEL_EVAL / t_norm_s / 1e12,
            "thread0_cycle_share": {"pixel_loop": cyc["cycles_pixels"] / cyc_tot, "serial_lm_step": cyc["cycles_serial"] / cyc_tot},
            # the packed FMAs of the pixel loop read {64-bit, 32-bit broadcast constant, 64-bit} register sources: that form
            # sustains 54.6 TFLOP/s on this pool (three distinct 64-bit sources: 47.1), not the 73.96 of re-used operands --
            # vector register-file bandwidth (tools/micro/ffma2_operands.cu, profiles/r01f_ffma2_operand_microbench.txt)
            "peak_operand_limited": 54.63 * sm_max / 1965.0,
            "frac_of_operand_limited_peak": achieved / (54.63 * sm_max / 1965.0),
            "ms_per_launch": t_norm_s * 1e3,
            "hbm_compulsory_gbs": (n_inl * 44e3 / t_norm_s) / 1e9, "hbm_peak_gbs": hbm_peak,
        }
        roofline_other = {
            "matcher_tc": {"bound": "tensor", "achieved": match_pairs * 256 / (stage_ms[0] / args.steps * 1e-3) / 1e12,
                           "peak": bf16_peak, "unit": "TFLOP/s", "ms": stage_ms[0] / args.steps,
                           "note": "includes operand re-tiling, NNDR filter and (N>1) the NCCL broadcast"},
            "pyrdown": {"bound": "hbm", "achieved": pyr_bytes / (stage_ms[2] / args.steps * 1e-3) / 1e9, "peak": hbm_peak,
                        "unit": "GB/s", "ms": stage_ms[2] / args.steps, "note": "3 launches + 2 device copies of 0.9 MB images: launch-bound"},
        }
        cpu = cpu_reference(case, max(16 * (os.cpu_count() or 1), 256), os.cpu_count() or 1) if world == 1 else None
        cpu_alt = cpu_reference(case, max(8 * (os.cpu_count() or 1), 128), os.cpu_count() or 1, penalty=PENALTY_ALT) if world == 1 else None
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
                "value": cpu["features_per_s"], "unit": "features/s", "cores": os.cpu_count() or 1, "kind": "port",
                "sample": f"full {N_KP}x{N_KP + N_DISTRACT} matching + normal optimisation of {cpu['n_sample']} seeded features "
                          f"of {cpu['n_match']}, per-feature times summed", "detail": cpu},
            "stage_ms_per_step": {"match+nndr": stage_ms[0] / args.steps, "triangulate": stage_ms[1] / args.steps,
                                  "pyramids": stage_ms[2] / args.steps, "normals": stage_ms[3] / args.steps,
                                  "gather": stage_ms[4] / args.steps},
            "features_per_step": {"matches": tot_feat, "inliers_rank0": int(n_inl), "ok_rank0": int((status == 0).sum()),
                                  "wall_touching_rank0": int((npen > 0).sum()), "nfev_mean_per_level": nfev.mean(0).tolist()},
            "penalty_fabs": {
                "note": "the same step with the penalty wall as today's g++ compiles it (abs == fabs): every synthetic "
                        "feature ends on the wall in the reference too (SURVEY fact 11); reported, not the headline",
                "value": tot_feat_alt / (t_alt_ms * 1e-3), "unit": "features/s", "ms_per_step": t_alt_ms,
                "normals_ms": t_alt_norm_ms, "wall_touching_rank0": int((npen_alt > 0).sum()),
                "nfev_mean_per_level": nfev_alt.mean(0).tolist(),
                "passes": stats_alt["passes_value"] + stats_alt["passes_jacobian"] + stats_alt["passes_fused"],
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
                "descriptor_rows_nonzero": kdesc_rows_nonzero},
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
