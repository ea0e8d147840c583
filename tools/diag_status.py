"""How many features of the C5 stress scene end with status OK under the kernel variants (r = 32, four pyramid images)."""
import sys, os, importlib, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
api = importlib.import_module("3dfeaturematcher_b200.api")
synth = importlib.import_module("3dfeaturematcher_b200.synth")
n_feat = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
r, pyr = 32, 3
case = synth.make_stereo_case(1920, 1080, n_feat, 1004, pixels_ray=r, n_distractors=0)
cam = case["scene"].cam
ctx = api.Context(0)
ctx.set_camera(cam.K, cam.dist, cam.z_min, cam.z_max); ctx.set_g12(cam.g12)
ctx.set_images(case["scene"].img1, case["scene"].img2, pyr)
xyz = np.ascontiguousarray(case["X"])
base = dict(normals_fast=1, normals_pingpong=1, normals_memo=3, normals_fuse=3, normals_groups=0)
ref = None
cfgs = [dict(), dict(normals_pingpong=0), dict(normals_memo=0), dict(normals_fuse=0), dict(normals_fast=0)]
if os.environ.get("FM3D_DIAG_SHORT"):
    cfgs = cfgs[:1]
for cfg in cfgs:
    for k, v in {**base, **cfg}.items():
        try:
            ctx.set_option(k, v)
        except Exception:
            pass
    xs = xyz if cfg.get("normals_fast", 1) else xyz[:4000]
    res = ctx.optimize_normals(xs, r, 1e-10, 1)
    st = res["status"]
    gt = np.degrees(np.arccos(np.clip((res["normals"] * case["normal"][:len(xs)]).sum(1), -1, 1)))
    if ref is None: ref = st.copy()
    print(cfg, "n", len(xs), "status histogram", np.bincount(st, minlength=5).tolist(), "ok within 0.5 deg", int(((st == 0) & (gt < 0.5)).sum()),
          "status equal to default", int((st == ref[:len(xs)]).sum()), "nfev", res["nfev"].mean(0).round(1).tolist(), flush=True)
ctx.close()
