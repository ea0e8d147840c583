"""Diagnostic (GPU box): where the fast / faithful kernels and the CPU oracle disagree on the 4K scene, who is closer to the
ground truth and who reaches the lower value of the reference's own cost function (evaluated by the oracle at level 0)."""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import angle_deg, cam_tuple, car2sph, orc, setup_ctx, stereo_case  # noqa: E402

api = importlib.import_module("3dfeaturematcher_b200.api")
ctx = api.Context(0)
W, H, n, seed = (int(a) for a in (sys.argv[1:5] if len(sys.argv) > 4 else (3840, 2160, 600, 1002)))
case = stereo_case(W, H, n, seed, 64)
cam = case["scene"].cam
setup_ctx(ctx, case, 3)
rng = np.random.default_rng(77)
X = case["X"]
sel = np.sort(rng.choice(X.shape[0], min(256, X.shape[0]), replace=False))
xyz = np.ascontiguousarray(X[sel])
truth = case["normal"][sel]
for pm in (2, 1):
    o = orc.optimize_normals(*cam_tuple(cam), case["scene"].img1, case["scene"].img2, 3, xyz, 64, 1e-10, penalty_mode=pm, threads=os.cpu_count())
    oc, _, _ = orc.evaluate_cost(*cam_tuple(cam), case["scene"].img1, case["scene"].img2, 3, xyz, car2sph(o["normals"]), 64, 0, 2)
    tc, _, _ = orc.evaluate_cost(*cam_tuple(cam), case["scene"].img1, case["scene"].img2, 3, xyz, car2sph(truth), 64, 0, 2)
    for name, fast, groups in (("fast-4groups", 1, 4), ("fast-1group", 1, 1), ("faithful", 0, 0)):
        ctx.set_option("normals_fast", fast); ctx.set_option("normals_groups", groups)
        res = ctx.optimize_normals(xyz, 64, 1e-10, pm)
        ctx.set_option("normals_fast", 1); ctx.set_option("normals_groups", 0)
        gc, _, _ = orc.evaluate_cost(*cam_tuple(cam), case["scene"].img1, case["scene"].img2, 3, xyz, car2sph(res["normals"]), 64, 0, 2)
        ang = angle_deg(res["normals"], o["normals"])
        g_gt, o_gt = angle_deg(res["normals"], truth), angle_deg(o["normals"], truth)
        ok = (o["status"] == 0) & (res["status"] == 0) & (o["npenalty"] == 0)
        bad = np.nonzero(ok & (ang > 0.1))[0]
        print(f"== penalty {pm} {name}: status equal {np.array_equal(res['status'], o['status'])}, interior {int(ok.sum())}, "
              f"ang p50 {np.median(ang[ok]):.5f} p99 {np.percentile(ang[ok], 99):.4f} max {ang[ok].max():.4f}; "
              f"gpu-vs-truth p50 {np.median(g_gt[ok]):.4f} max {g_gt[ok].max():.4f}; oracle-vs-truth p50 {np.median(o_gt[ok]):.4f} max {o_gt[ok].max():.4f}; "
              f"gpu cost <= oracle cost (1e-9 rel) on {(gc[ok] <= oc[ok] * (1 + 1e-9)).mean():.3f}")
        for i in bad:
            print(f"   f{i}: ang {ang[i]:.4f} gpu_gt {g_gt[i]:.4f} orc_gt {o_gt[i]:.4f} cost gpu {gc[i]:.6f} orc {oc[i]:.6f} truth {tc[i]:.6f} "
                  f"rel {(gc[i] - oc[i]) / oc[i]:+.2e} nfev gpu {res['nfev'][i].tolist()} orc {o['nfev'][i].tolist()} P {xyz[i].round(3).tolist()}")
