"""A/B timing of the normal-search kernel options on the bench workload (device API, CUDA events)."""
import sys, importlib, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from common import *
api = importlib.import_module("3dfeaturematcher_b200.api")
ctx = api.Context(0)
dev = torch.device("cuda", 0)
stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
n_feat = int(sys.argv[1]) if len(sys.argv) > 1 else 2960
case = stereo_case(1280, 720, n_feat, 1001, 64)
cam = case["scene"].cam
setup_ctx(ctx, case, 3)
xyz_h = np.ascontiguousarray(case["X"]); n = xyz_h.shape[0]
xyz = torch.from_numpy(xyz_h).to(dev)
normals = torch.empty((n, 3), dtype=torch.float64, device=dev); status = torch.empty(n, dtype=torch.int32, device=dev)
nfev = torch.zeros((n, 4), dtype=torch.int32, device=dev); npen = torch.zeros(n, dtype=torch.int32, device=dev); cost = torch.empty(n, dtype=torch.float64, device=dev)
configs = [json.loads(a) for a in sys.argv[2:]] or [dict(normals_groups=1), dict(normals_groups=4)]; sys.argv = sys.argv[:2]
configs += [json.loads(a) for a in sys.argv[2:]]
base = dict(normals_fast=1, normals_memo=3, normals_fuse=3, normals_threads=512, normals_groups=0)
for pen in (1, 0):
    ref = None
    for cfg in configs:
        for k, v in {**base, **cfg}.items(): ctx.set_option(k, v)
        fn = lambda: ctx.optimize_normals_dev(xyz.data_ptr(), n, 64, 1e-10, pen, normals.data_ptr(), status.data_ptr(), nfev.data_ptr(), npen.data_ptr(), cost.data_ptr())
        fn(); stream.synchronize()
        times = []
        for _ in range(int(os.environ.get("AB_REPS", "1"))):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(stream):
                e0.record(stream); fn(); fn(); e1.record(stream)
            stream.synchronize()
            times.append(e0.elapsed_time(e1) / 2)
        ms = float(np.median(times))
        if len(times) > 1: print(f"   [{os.environ.get('FM3D_LIB', 'default lib')}] launches (ms, pairs): min {min(times):.3f} median {ms:.3f} max {max(times):.3f}", flush=True)
        st = ctx.normals_stats() if cfg.get("normals_fast", 1) else {}
        nf = nfev.cpu().numpy(); nr = normals.cpu().numpy()
        gt = angle_deg(nr, case["normal"])
        if ref is None and cfg.get("normals_fast", 1) == 1: ref = (nf.copy(), nr.copy())
        same = None if ref is None else (bool((nf == ref[0]).all()), float(angle_deg(nr, ref[1]).max()))
        passes = st.get("passes_value", 0) + st.get("passes_jacobian", 0) + st.get("passes_fused", 0)
        cyc = {k: round(st[k] / max(passes, 1)) for k in ("cycles_pixels", "cycles_barrier", "cycles_serial", "cycles_lm", "cycles_publish")} if st else {}
        if st: cyc.update({"per_feature_prologue": round(st["cycles_prologue"] / n), "per_feature_level_setup": round(st["cycles_level_setup"] / n), "per_feature_passes": round((st["cycles_pixels"] + st["cycles_serial"] + st["cycles_barrier"]) / n)})
        print(f"pen {pen} {cfg}: {ms:.2f} ms / {n} features = {n/ms:.1f} feat/ms; nfev {nf.mean(0).round(1).tolist()} GT median {np.median(gt):.4f}; "
              f"passes/feature {passes/max(n,1):.1f} memo {st.get('trials_memoized',0)/max(n,1):.1f} fused {st.get('passes_fused',0)/max(n,1):.1f} acc {st.get('fused_accepted',0)/max(n,1):.1f}; cyc/pass {cyc}; same-as-first-fast {same}", flush=True)
ctx.close()
