"""Debug/inspection: fast vs faithful normal-search kernels against the oracle (run on the GPU box)."""
import sys, importlib, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
from common import *
api = importlib.import_module("3dfeaturematcher_b200.api")
ctx = api.Context(0)

def run(ctx, xyz, r, pen, fast, fuse=1):
    ctx.set_option("normals_fast", fast); ctx.set_option("normals_fuse", fuse)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    res = ctx.optimize_normals(xyz, r, 1e-10, pen)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    st = ctx.normals_stats() if fast else {}
    ctx.set_option("normals_fast", 1); ctx.set_option("normals_fuse", 1)
    return res, dt, st

which = sys.argv[1] if len(sys.argv) > 1 else "clip"
if which == "clip":
    case = stereo_case(640, 480, 40, 1001, 32)
    cam = case["scene"].cam
    setup_ctx(ctx, case, 2)
    pix = np.array([[8.0, 240.0], [300.0, 6.0], [320.0, 472.0], [250.0, 470.0], [400.0, 473.0]])
    x, y, _ = synth.undistort_exact((pix[:, 0] - cam.K[0, 2]) / cam.K[0, 0], (pix[:, 1] - cam.K[1, 2]) / cam.K[1, 1], cam.dist)
    pts, _, hit = synth._raycast(case["scene"].planes, np.zeros(3), np.stack([x, y, np.ones_like(x)], axis=1))
    xyz = np.concatenate([case["X"][:1], pts, np.array([[9.0, 0.3, 1.9]]), case["X"][1:2]])
    o = orc.optimize_normals(*cam_tuple(cam), case["scene"].img1, case["scene"].img2, 2, xyz, 32, 1e-10, penalty_mode=2, threads=4)
    for fast in (0, 1):
        for fuse in ((1,) if not fast else (1, 0)):
            res, dt, st = run(ctx, xyz, 32, 2, fast, fuse)
            print("fast", fast, "fuse", fuse, "status", res["status"], "angle", np.round(angle_deg(res["normals"], o["normals"]), 4))
            print("   nfev", res["nfev"].tolist(), "\n   oracle nfev", o["nfev"].tolist())
            print("   cost", res["cost"], "\n   oracle cost", o["cost"], st)
else:
    W, H, n, r, pyr = (1280, 720, 600, 64, 3)
    case = stereo_case(W, H, n, 1001, r)
    cam = case["scene"].cam
    setup_ctx(ctx, case, pyr)
    xyz = case["X"]
    for pen in (1, 0):
        o = orc.optimize_normals(*cam_tuple(cam), case["scene"].img1, case["scene"].img2, pyr, xyz[:64], r, 1e-10, penalty_mode=pen, threads=16)
        for fast, fuse in ((0, 1), (1, 1), (1, 0)):
            run(ctx, xyz, r, pen, fast, fuse)
            res, dt, st = run(ctx, xyz, r, pen, fast, fuse)
            ang = angle_deg(res["normals"][:64], o["normals"])
            gt = angle_deg(res["normals"], case["normal"])
            ok = res["status"] == 0
            print(f"pen {pen} fast {fast} fuse {fuse}: {dt*1e3:.2f} ms for {n} features; nfev/level {res['nfev'].mean(0).round(1).tolist()} oracle {o['nfev'].mean(0).round(1).tolist()}"
                  f" angle-to-oracle median {np.median(ang):.4f} max {ang.max():.3f}; angle-to-GT median {np.median(gt[ok]):.4f} p90 {np.percentile(gt[ok],90):.3f}"
                  f" oracle-to-GT median {np.median(angle_deg(o['normals'], case['normal'][:64])):.4f}; cost ratio median {np.median(res['cost'][:64]/o['cost']):.6f}; wall {int((res['npenalty']>0).sum())}")
            print("    ", st)
