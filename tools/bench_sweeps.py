"""BASELINE.json configs[3] and configs[4] on one GPU: descriptor-matching sweep (SIFT-128 float on
the tcgen05 path, ORB-256 binary on the popc path) and the NormalOptimizer stress (pixelsRay 32-128,
3-5 pyramid images).  Prints one JSON line per case; CUDA-event timing on the context's stream."""
import importlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

api = importlib.import_module("3dfeaturematcher_b200.api")
synth = importlib.import_module("3dfeaturematcher_b200.synth")
PEAKS = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}


def timed(stream, fn, reps):
    fn(); stream.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
    stream.synchronize()
    return e0.elapsed_time(e1) / reps


def match_sweep(ctx, stream, sizes):
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(1003)
    for n in sizes:
        with torch.cuda.stream(stream):
            t = torch.randint(0, 256, (n, 128), device=dev, generator=g).float()
            q = (t[torch.randperm(n, device=dev, generator=g)] + torch.randint(-6, 7, (n, 128), device=dev, generator=g).float()).clamp_(0, 255)
            idx = torch.empty((n, 2), dtype=torch.int32, device=dev)
            dist = torch.empty((n, 2), dtype=torch.float32, device=dev)
            tb = torch.randint(0, 256, (n, 32), device=dev, generator=g, dtype=torch.uint8)
            qb = tb[torch.randperm(n, device=dev, generator=g)].clone()
        stream.synchronize()
        reps = 5 if n <= 50000 else 2
        ms = timed(stream, lambda: ctx.match_knn2_f32_dev(q.data_ptr(), n, t.data_ptr(), n, 128, idx.data_ptr(), dist.data_ptr()), reps)
        tf = 256.0 * n * n / (ms * 1e-3) / 1e12
        peak = PEAKS.get("bf16_tflops", 1590.0)
        print(json.dumps({"case": "match_f32_sift128_tcgen05", "nq": n, "nt": n, "ms": ms, "pairs_per_s": n * n / (ms * 1e-3),
                          "tflops": tf, "frac_of_measured_bf16_peak": tf / peak, "peak_tflops": peak,
                          "note": "includes operand re-tiling (tc_prep) and finalize"}), flush=True)
        ms = timed(stream, lambda: ctx.match_knn2_hamming_dev(qb.data_ptr(), n, tb.data_ptr(), n, 32, idx.data_ptr(), dist.data_ptr()), reps)
        print(json.dumps({"case": "match_hamming_orb256_popc", "nq": n, "nt": n, "ms": ms, "pairs_per_s": n * n / (ms * 1e-3),
                          "word_ops_per_s": 8.0 * n * n / (ms * 1e-3)}), flush=True)


def normals_stress(ctx, stream, n_feat):
    dev = torch.device("cuda", 0)
    for r in (32, 64, 128):
        case = synth.make_stereo_case(1920, 1080, n_feat, 1004, pixels_ray=r, n_distractors=0)
        cam = case["scene"].cam
        ctx.set_camera(cam.K, cam.dist, cam.z_min, cam.z_max)
        ctx.set_g12(cam.g12)
        xyz = torch.from_numpy(np.ascontiguousarray(case["X"])).to(dev)
        n = xyz.shape[0]
        m_disc = sum(2 * int(np.floor(np.sqrt(r * r - j * j))) + 1 for j in range(-r, r + 1))
        for pyr in (2, 3, 4):
            ctx.set_images(case["scene"].img1, case["scene"].img2, pyr)
            normals = torch.empty((n, 3), dtype=torch.float64, device=dev)
            status = torch.empty(n, dtype=torch.int32, device=dev)
            nfev = torch.zeros((n, pyr + 1), dtype=torch.int32, device=dev)
            npen = torch.zeros(n, dtype=torch.int32, device=dev)
            cost = torch.empty(n, dtype=torch.float64, device=dev)
            fn = lambda: ctx.optimize_normals_dev(xyz.data_ptr(), n, r, 1e-10, 1, normals.data_ptr(), status.data_ptr(),
                                                  nfev.data_ptr(), npen.data_ptr(), cost.data_ptr())
            ms = timed(stream, fn, 2)
            st = ctx.normals_stats()
            gt = np.degrees(np.arccos(np.clip((normals.cpu().numpy() * case["normal"]).sum(1), -1, 1)))
            ok = status.cpu().numpy() == 0
            flops = st["pixel_evals_value"] * 64.0 + st["pixel_evals_jacobian"] * 152.0
            print(json.dumps({"case": "normals_stress", "pixels_ray": r, "m": m_disc, "pyramid_images": pyr + 1, "features": n,
                              "ms": ms, "features_per_s": n / (ms * 1e-3), "ok": int(ok.sum()),
                              "median_angle_to_gt_deg": float(np.median(gt[ok])) if ok.any() else None,
                              "pixel_evals_per_s": (st["pixel_evals_value"] + st["pixel_evals_jacobian"]) / (ms * 1e-3),
                              "tflops_fp32_algorithmic": flops / (ms * 1e-3) / 1e12, "passes_global_taps": st["passes_slow"],
                              "nfev_mean_per_level": nfev.float().mean(0).tolist()}), flush=True)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    ctx = api.Context(0)
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", 0))
    if which in ("all", "match"):
        match_sweep(ctx, stream, [10000, 50000, 100000, 200000])
    if which in ("all", "normals"):
        normals_stress(ctx, stream, int(sys.argv[2]) if len(sys.argv) > 2 else 2000)
    ctx.close()
