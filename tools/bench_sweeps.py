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
                          "word_ops_per_s": 8.0 * n * n / (ms * 1e-3),
                          "frac_of_measured_popc_peak": 8.0 * n * n / (ms * 1e-3) / 4.54e12,
                          "note": "POPC roofline 4.54e12/s measured with tools/micro/popc_rate.cu (16 lanes/clk/SM)"}), flush=True)


def match_splits_sweep(ctx, stream):
    """The integer tensor-core matcher: one CTA per (query tile, train split) with round 1's split rule (`matcher_persistent`
    0; `matcher_splits` > 0 forces the splits) against the persistent kernel over equal tile ranges (`matcher_min_tiles` =
    fewest tiles worth a CTA); results are checked to be identical for every choice."""
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(1003)
    small = ((0, 0, 2), (1, 0, 1), (1, 0, 2), (1, 0, 4), (0, 0, 2), (1, 0, 2))
    big = ((1, 0, 2), (0, 0, 2), (1, 0, 2), (0, 0, 2), (1, 0, 2), (0, 0, 2))       # alternating: the later launches of a size run on a warmer chip
    for nq, nt, cases in ((300, 360, small), (1000, 1200, small), (2000, 2400, small), (5000, 6000, small), (5000, 48000, small), (10000, 10000, small),
                          (20000, 24000, small), (50000, 50000, big), (100000, 100000, big), (150000, 150000, big), (200000, 200000, big)):
        with torch.cuda.stream(stream):
            t = torch.randint(0, 256, (nt, 128), device=dev, generator=g).float()
            q = (t[torch.randperm(nt, device=dev, generator=g)[:nq]] + torch.randint(-6, 7, (nq, 128), device=dev, generator=g).float()).clamp_(0, 255)
            idx = torch.empty((nq, 2), dtype=torch.int32, device=dev)
            dist = torch.empty((nq, 2), dtype=torch.float32, device=dev)
        stream.synchronize()
        ref = None
        for pers, f, mt in cases:
            ctx.set_option("matcher_persistent", pers)
            ctx.set_option("matcher_splits", f)
            ctx.set_option("matcher_min_tiles", mt)
            idx.fill_(-7); stream.synchronize()
            ms = timed(stream, lambda: ctx.match_knn2_f32_dev(q.data_ptr(), nq, t.data_ptr(), nt, 128, idx.data_ptr(), dist.data_ptr()), 5 if nq <= 50000 else 3)
            got = (idx.cpu().numpy().copy(), dist.cpu().numpy().copy())
            if ref is None:
                ref = got
            same = bool((got[0] == ref[0]).all() and (got[1] == ref[1]).all())
            tf = 256.0 * nq * nt / (ms * 1e-3) / 1e12
            print(json.dumps({"case": "match_f32_sift128_tcgen05_splits", "lib": os.environ.get("FM3D_LIB", "default"), "nq": nq, "nt": nt, "matcher_persistent": pers,
                              "matcher_splits": f, "matcher_min_tiles": mt, "ms": ms,
                              "tflops": tf, "frac_of_measured_bf16_peak": tf / PEAKS.get("bf16_tflops", 1590.0), "same_as_first": same}), flush=True)
        ctx.set_option("matcher_splits", 0)
        ctx.set_option("matcher_persistent", 1)
        ctx.set_option("matcher_min_tiles", 1)


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
            flops = st["pixel_evals_value"] * 58.0 + st["pixel_evals_jacobian"] * 113.0      # executed
            print(json.dumps({"case": "normals_stress", "pixels_ray": r, "m": m_disc, "pyramid_images": pyr + 1, "features": n,
                              "ms": ms, "features_per_s": n / (ms * 1e-3), "ok": int(ok.sum()),
                              "median_angle_to_gt_deg": float(np.median(gt[ok])) if ok.any() else None,
                              "pixel_evals_per_s": (st["pixel_evals_value"] + st["pixel_evals_jacobian"]) / (ms * 1e-3),
                              "tflops_fp32_executed": flops / (ms * 1e-3) / 1e12, "passes_global_taps": st["passes_slow"],
                              "nfev_mean_per_level": nfev.float().mean(0).tolist()}), flush=True)
            if pyr == 3:
                # dense candidate-normal sampling: 33 x 33 grid around the initial normal at level 0
                best = torch.empty(n, dtype=torch.int32, device=dev); bcost = torch.empty(n, dtype=torch.float64, device=dev)
                nsw = min(n, 592)
                fn2 = lambda: ctx.sweep_normals_dev(xyz.data_ptr(), nsw, r, 0, 33, 33, 0.01, 0.01, status.data_ptr(), penalty_mode=1,
                                                    best_idx=best.data_ptr(), best_cost=bcost.data_ptr())
                ms2 = timed(stream, fn2, 1)
                print(json.dumps({"case": "normals_dense_sweep_33x33_level0", "pixels_ray": r, "m": m_disc, "features": nsw, "ms": ms2,
                                  "pixel_evals_per_s": nsw * 1089.0 * m_disc / (ms2 * 1e-3),
                                  "tflops_fp32_algorithmic": nsw * 1089.0 * m_disc * 64.0 / (ms2 * 1e-3) / 1e12,
                                  "candidates_per_s": nsw * 1089.0 / (ms2 * 1e-3)}), flush=True)


def c3_pipeline(ctx, stream):
    """BASELINE configs[2] on ONE GPU: 4K stereo pair, 20 000 keypoints (+4 000 distractors), 64 px,
    4 levels; match -> triangulate -> pyramids -> normal search -> frames -> patches (K8)."""
    dev = torch.device("cuda", 0)
    t0 = time.time()
    case = synth.make_stereo_case(3840, 2160, 20000, 1002, pixels_ray=64, n_distractors=4000)
    gen_s = time.time() - t0
    cam = case["scene"].cam
    ctx.set_camera(cam.K, cam.dist, cam.z_min, cam.z_max)
    ctx.set_g12(cam.g12)
    H, W = case["scene"].img1.shape
    to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    q, t, kp1, kp2 = to(case["desc1"]), to(case["desc2"]), to(case["kp1"]), to(case["kp2"])
    img1, img2 = to(case["scene"].img1), to(case["scene"].img2)
    nq, nt = q.shape[0], t.shape[0]
    idx = torch.empty((nq, 2), dtype=torch.int32, device=dev); dist = torch.empty((nq, 2), dtype=torch.float32, device=dev)
    qi = torch.empty(nq, dtype=torch.int32, device=dev); ti = torch.empty(nq, dtype=torch.int32, device=dev)
    do = torch.empty(nq, dtype=torch.float32, device=dev); nm = torch.zeros(1, dtype=torch.int32, device=dev)
    xyz_all = torch.empty((nq, 3), dtype=torch.float64, device=dev); xyz = torch.empty((nq, 3), dtype=torch.float64, device=dev)
    mask = torch.empty(nq, dtype=torch.uint8, device=dev); src = torch.empty(nq, dtype=torch.int32, device=dev)
    ninl = torch.zeros(1, dtype=torch.int32, device=dev)
    normals = torch.empty((nq, 3), dtype=torch.float64, device=dev); status = torch.empty(nq, dtype=torch.int32, device=dev)
    nfev = torch.zeros((nq, 4), dtype=torch.int32, device=dev); npen = torch.zeros(nq, dtype=torch.int32, device=dev)
    cost = torch.empty(nq, dtype=torch.float64, device=dev)
    frames = torch.empty((nq, 16), dtype=torch.float64, device=dev)
    S = api.patch_size(0.16, 0.25)
    patches = torch.empty((nq, S, S), dtype=torch.uint8, device=dev)
    pdesc = torch.empty((nq, 128), dtype=torch.float32, device=dev)
    g = np.array([0.006, 0.99992, -0.011])
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(8)]
    out = {}
    for rep in range(2):
        with torch.cuda.stream(stream):
            ev[0].record(stream)
            ctx.match_knn2_f32_dev(q.data_ptr(), nq, t.data_ptr(), nt, 128, idx.data_ptr(), dist.data_ptr())
            ctx.nndr_filter_dev(idx.data_ptr(), dist.data_ptr(), nq, 0.55, qi.data_ptr(), ti.data_ptr(), do.data_ptr(), nm.data_ptr())
            ev[1].record(stream)
            n_match = int(nm.item())
            ctx.triangulate_dev(kp1.data_ptr(), nq, kp2.data_ptr(), nt, qi.data_ptr(), ti.data_ptr(), n_match, xyz_all.data_ptr(),
                                mask.data_ptr(), xyz.data_ptr(), src.data_ptr(), ninl.data_ptr())
            ev[2].record(stream)
            ctx.set_images_dev(img1.data_ptr(), img2.data_ptr(), W, H, W, 3)
            ev[3].record(stream)
            n_inl = int(ninl.item())
            ctx.optimize_normals_dev(xyz.data_ptr(), n_inl, 64, 1e-10, 1, normals.data_ptr(), status.data_ptr(), nfev.data_ptr(),
                                     npen.data_ptr(), cost.data_ptr())
            ev[4].record(stream)
            ctx.feature_frames_dev(xyz.data_ptr(), normals.data_ptr(), n_inl, g, frames.data_ptr())
            ev[5].record(stream)
            ctx.extract_patches_dev(frames.data_ptr(), n_inl, 0.16, 0.25, patches.data_ptr(), None)
            ev[6].record(stream)
            ctx.describe_patches_sift_dev(patches.data_ptr(), n_inl, S, pdesc.data_ptr())
            ev[7].record(stream)
        stream.synchronize()
    ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(7)]
    st = ctx.normals_stats()
    src_h = src[:n_inl].cpu().numpy(); qi_h = qi[:n_match].cpu().numpy()
    feat = qi_h[src_h]                       # query keypoint of every inlier
    gt_n = case["normal"][feat]
    ok = status[:n_inl].cpu().numpy() == 0
    ang = np.degrees(np.arccos(np.clip((normals[:n_inl].cpu().numpy() * gt_n).sum(1), -1, 1)))
    t_core = ms[0] + ms[1] + ms[2] + ms[3]
    pyr_bytes = sum(2 * 1.25 * (W >> l) * (H >> l) for l in range(3))
    print(json.dumps({"case": "c3_4k_20k_one_gpu", "gen_s": gen_s, "matches": n_match, "inliers": n_inl, "ok": int(ok.sum()),
                      "median_angle_to_gt_deg": float(np.median(ang[ok])), "p99_angle_to_gt_deg": float(np.percentile(ang[ok], 99)),
                      "stage_ms": {"match+nndr": ms[0], "triangulate": ms[1], "pyramids": ms[2], "normals": ms[3], "frames": ms[4], "patches": ms[5], "patch_descriptors": ms[6]},
                      "features_per_s_match_tri_normals": n_match / (t_core * 1e-3),
                      "matcher_tflops": 256.0 * nq * nt / (ms[0] * 1e-3) / 1e12,
                      "pyrdown_gbs": pyr_bytes / (ms[2] * 1e-3) / 1e9, "hbm_peak_gbs": PEAKS.get("hbm_gbs"),
                      "patch_descriptors_per_s": n_inl / (ms[6] * 1e-3), "patch_descriptors_nonzero": int((pdesc[:n_inl] != 0).sum().item()),
                      "patches_per_s": n_inl / (ms[5] * 1e-3), "patch_px_per_s": n_inl * S * S / (ms[5] * 1e-3),
                      "normals_pixel_evals_per_s": (st["pixel_evals_value"] + st["pixel_evals_jacobian"]) / (ms[3] * 1e-3),
                      "passes_global_taps": st["passes_slow"]}), flush=True)


def match_real_sweep(ctx, stream):
    """Real-valued float descriptors (SURF-like unit vectors): tensor-core filter + exact decision vs the exact CUDA-core path."""
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(1007)
    for n, dim in ((10000, 64), (20000, 64), (20000, 128), (50000, 64), (50000, 128), (100000, 128), (200000, 128)):
        q = torch.randn((n, dim), device=dev, generator=g); t = torch.randn((n, dim), device=dev, generator=g)
        q = q / q.norm(dim=1, keepdim=True); t = t / t.norm(dim=1, keepdim=True)
        idx = torch.empty((n, 2), dtype=torch.int32, device=dev); dist = torch.empty((n, 2), dtype=torch.float32, device=dev)
        fn = lambda: ctx.match_knn2_f32_dev(q.data_ptr(), n, t.data_ptr(), n, dim, idx.data_ptr(), dist.data_ptr())
        # one CTA per query tile (`matcher_persistent` 0) and the persistent kernel, alternating (a size's later launches run warmer)
        ab, ref, same = {0: [], 1: []}, None, True
        for pers in (1, 0, 1, 0):
            ctx.set_option("matcher_persistent", pers)
            ab[pers].append(timed(stream, fn, 2))
            got = (idx.cpu().numpy().copy(), dist.cpu().numpy().copy())
            ref = ref or got
            same = same and bool((got[0] == ref[0]).all() and (got[1] == ref[1]).all())
        ctx.set_option("matcher_persistent", 1)
        ms = min(ab[1])
        fb = int(ctx.get_option("matcher_exact_fallback"))
        out = {"case": "match_f32_real_valued_tensor_filter", "nq": n, "nt": n, "dim": dim, "ms": ms, "ms_persistent": ab[1], "ms_one_cta_per_query_tile": ab[0],
               "persistent_equals_one_cta_per_query_tile": same, "pairs_per_s": float(n) * n / (ms * 1e-3),
               "mma_tflops": float(n) * n * (3 * ((dim + 15) // 16) + 1) * 16 * 2 / (ms * 1e-3) / 1e12, "queries_decided_by_exact_path": fb}
        if n <= 50000:
            ctx.set_option("matcher_tensor", 0)
            out["ms_exact_cuda_core_path"] = timed(stream, fn, 1)
            ctx.set_option("matcher_tensor", 1)
        print(json.dumps(out), flush=True)


def describe_sweep(ctx, stream):
    """K9 alone: SIFT descriptors of n rectified patches (extractDescriptorsFromPatches), S = 128 and 64."""
    dev = torch.device("cuda", 0)
    for S, n in ((128, 16384), (64, 65536), (40, 65536)):
        patches = torch.randint(0, 256, (n, S, S), dtype=torch.uint8, device=dev)
        desc = torch.empty((n, 128), dtype=torch.float32, device=dev)
        fn = lambda: ctx.describe_patches_sift_dev(patches.data_ptr(), n, S, desc.data_ptr())
        ms = timed(stream, fn, 3)
        # algorithmic work per pixel: separable 13-tap symmetric blur 2 x 19 flop, gradient + weight + trilinear split ~ 60 flop
        print(json.dumps({"case": "describe_patches_sift", "S": S, "patches": n, "ms": ms, "patches_per_s": n / (ms * 1e-3),
                          "pixels_per_s": n * S * S / (ms * 1e-3), "hbm_gbs_compulsory": (n * S * S + n * 512) / (ms * 1e-3) / 1e9,
                          "hbm_peak_gbs": PEAKS.get("hbm_gbs"), "tflops_fp32_algorithmic": n * S * S * 98.0 / (ms * 1e-3) / 1e12}), flush=True)


def frontend_sweep(ctx, stream):
    """K10 + K11 alone: FAST detection, the SIFT base image and the SIFT descriptors at the detected keypoints of
    one frame (the front end of compareWithNNDR with DetectorType FAST + ExtractorType SIFT), 720p and 4K."""
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(3)
    for W, H in ((1280, 720), (3840, 2160)):
        small = rng.integers(0, 256, (H // 6 + 1, W // 6 + 1)).astype(np.float32)
        img = np.kron(small, np.ones((6, 6), np.float32))[:H, :W]
        img = np.clip(img * 0.7 + rng.integers(0, 77, img.shape), 0, 255).astype(np.uint8)
        d_img = torch.from_numpy(img).to(dev)
        cap = 1 << 20
        xy = torch.zeros((cap, 2), dtype=torch.float32, device=dev)
        resp = torch.zeros(cap, dtype=torch.float32, device=dev)
        cnt = torch.zeros(1, dtype=torch.int32, device=dev)
        base = torch.zeros((H, W), dtype=torch.float32, device=dev)
        det = lambda: ctx.detect_fast_dev(d_img.data_ptr(), W, H, W, 30, True, cap, xy.data_ptr(), resp.data_ptr(), cnt.data_ptr())
        ms_det = timed(stream, det, 5)
        n = min(int(cnt.item()), cap)
        k4 = torch.cat([xy[:n], torch.full((n, 1), 7.0, device=dev), torch.full((n, 1), -1.0, device=dev)], 1).contiguous()
        desc = torch.empty((n, 128), dtype=torch.float32, device=dev)
        ms_base = timed(stream, lambda: ctx.sift_base_image_dev(d_img.data_ptr(), W, H, W, base.data_ptr()), 5)
        ms_all = timed(stream, lambda: ctx.describe_keypoints_sift_dev(d_img.data_ptr(), W, H, W, k4.data_ptr(), n, desc.data_ptr()), 5)
        # the same three steps with OpenCV on this box's host cores, where cv2 is importable (a reported baseline)
        try:
            import cv2
            t0 = time.perf_counter()
            cvk = cv2.FastFeatureDetector_create(threshold=30, nonmaxSuppression=True).detect(img, None)
            t1 = time.perf_counter()
            cv2.SIFT_create().compute(img, cvk)
            t2 = time.perf_counter()
            cv2.BRISK_create(25, 0).compute(img, cvk)
            t3 = time.perf_counter()
            cv2.ORB_create().compute(img, cvk)
            t4 = time.perf_counter()
            print(json.dumps({"case": "frontend_opencv_cpu", "W": W, "H": H, "keypoints": len(cvk), "threads": cv2.getNumThreads(),
                              "ms_detect_fast": (t1 - t0) * 1e3, "ms_sift_compute": (t2 - t1) * 1e3, "ms_brisk_compute": (t3 - t2) * 1e3,
                              "ms_orb_compute": (t4 - t3) * 1e3}), flush=True)
        except ImportError:
            pass
        bdesc = torch.empty((n, 64), dtype=torch.uint8, device=dev)
        bkept = torch.empty(n, dtype=torch.uint8, device=dev)
        bang = torch.empty(n, dtype=torch.float32, device=dev)
        ms_brisk = timed(stream, lambda: ctx.describe_keypoints_brisk_dev(d_img.data_ptr(), W, H, W, k4.data_ptr(), n, True, bdesc.data_ptr(),
                                                                          bkept.data_ptr(), bang.data_ptr()), 5)
        print(json.dumps({"case": "frontend_fast_brisk", "W": W, "H": H, "keypoints": n, "kept": int(bkept.sum().item()),
                          "ms_describe_incl_integral": ms_brisk, "descriptors_per_s": n / (ms_brisk * 1e-3)}), flush=True)
        odesc = torch.empty((n, 32), dtype=torch.uint8, device=dev)
        ms_orb = timed(stream, lambda: ctx.describe_keypoints_orb_dev(d_img.data_ptr(), W, H, W, k4.data_ptr(), n, odesc.data_ptr(), bkept.data_ptr()), 5)
        print(json.dumps({"case": "frontend_fast_orb", "W": W, "H": H, "keypoints": n, "kept": int(bkept.sum().item()),
                          "ms_describe_incl_blur": ms_orb, "descriptors_per_s": n / (ms_orb * 1e-3),
                          "hbm_gbs_compulsory": (2.0 * W * H + 48.0 * n) / (ms_orb * 1e-3) / 1e9, "hbm_peak_gbs": PEAKS.get("hbm_gbs")}), flush=True)
        print(json.dumps({"case": "frontend_fast_sift", "W": W, "H": H, "keypoints": n, "ms_detect": ms_det, "ms_base_image": ms_base,
                          "ms_describe_incl_base": ms_all, "detect_hbm_gbs_compulsory": W * H / (ms_det * 1e-3) / 1e9,
                          "base_hbm_gbs_compulsory": 5.0 * W * H / (ms_base * 1e-3) / 1e9, "hbm_peak_gbs": PEAKS.get("hbm_gbs"),
                          "descriptors_per_s": n / ((ms_all - ms_base) * 1e-3) if ms_all > ms_base else None}), flush=True)


def sift_detect_sweep(ctx, stream):
    """K14 alone: cv::SIFT's detector (fm3d_detect_sift, host buffers in and out: the upload of the frame, the pyramids, the
    extrema / refinement kernels, the download and the host-side sort are all inside the time) and the descriptors of its
    keypoints on the pyramid layers, on rendered frames of 720p and 4K; cv2.SIFT on this box's host cores beside it."""
    import importlib
    synth = importlib.import_module("3dfeaturematcher_b200.synth")
    for W, H in ((1280, 720), (3840, 2160)):
        img = synth.make_stereo_case(W, H, 20, 1001, pixels_ray=32)["scene"].img1
        ctx.detect_sift(img)
        t = []
        for _ in range(3):
            t0 = time.perf_counter(); K = ctx.detect_sift(img); t.append(time.perf_counter() - t0)
        ms_det = 1e3 * min(t)
        kps, oct_ = K[:, :4].astype(np.float32), K[:, 5].astype(np.int32)
        ctx.describe_keypoints_sift_oct(img, kps, oct_)
        t = []
        for _ in range(3):
            t0 = time.perf_counter(); ctx.describe_keypoints_sift_oct(img, kps, oct_); t.append(time.perf_counter() - t0)
        ms_desc = 1e3 * min(t)
        ctx.detect_and_describe_sift(img)
        t = []
        for _ in range(3):
            t0 = time.perf_counter(); ctx.detect_and_describe_sift(img); t.append(time.perf_counter() - t0)
        out = {"case": "sift_detect_and_describe", "W": W, "H": H, "keypoints": int(len(K)), "ms_detect_host_to_host": ms_det,
               "ms_describe_host_to_host": ms_desc, "ms_detect_and_describe_on_one_pyramid_host_to_host": 1e3 * min(t), "pyramid_mb": (2 * W) * (2 * H) * 4 * (6 + 5) * 4 / 3 / 1e6}
        try:
            import cv2
            s = cv2.SIFT_create()
            t0 = time.perf_counter(); kp = s.detect(img, None); t1 = time.perf_counter(); s.compute(img, kp); t2 = time.perf_counter()
            out.update({"cv2_keypoints": len(kp), "cv2_ms_detect": 1e3 * (t1 - t0), "cv2_ms_compute": 1e3 * (t2 - t1), "cv2_threads": cv2.getNumThreads()})
        except ImportError:
            pass
        print(json.dumps(out), flush=True)


def orb_detect_sweep(ctx, stream):
    """K15 alone: cv::ORB's detector + descriptors (fm3d_detect_orb, host buffers in and out) on rendered frames of 720p and 4K;
    cv2.ORB.detectAndCompute on this box's host cores beside it."""
    import importlib
    synth = importlib.import_module("3dfeaturematcher_b200.synth")
    for W, H, nf in ((1280, 720, 5000), (3840, 2160, 20000)):
        img = synth.make_stereo_case(W, H, 20, 1001, pixels_ray=32)["scene"].img1
        ctx.detect_orb(img, nfeatures=nf)
        t = []
        for _ in range(3):
            t0 = time.perf_counter(); K, D = ctx.detect_orb(img, nfeatures=nf); t.append(time.perf_counter() - t0)
        out = {"case": "orb_detect_and_describe", "W": W, "H": H, "nfeatures": nf, "keypoints": int(len(K)), "ms_host_to_host": 1e3 * min(t)}
        try:
            import cv2
            o = cv2.ORB_create(nfeatures=nf)
            t0 = time.perf_counter(); kp, d = o.detectAndCompute(img, None); t1 = time.perf_counter()
            out.update({"cv2_keypoints": len(kp), "cv2_ms": 1e3 * (t1 - t0), "cv2_threads": cv2.getNumThreads()})
        except ImportError:
            pass
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    ctx = api.Context(0)
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", 0))
    if which in ("all", "match"):
        match_sweep(ctx, stream, [int(a) for a in sys.argv[2:]] if which == "match" and len(sys.argv) > 2 else [10000, 50000, 100000, 200000])
    if which in ("all", "normals"):
        normals_stress(ctx, stream, int(sys.argv[2]) if len(sys.argv) > 2 else 2000)
    if which in ("all", "match_real"):
        match_real_sweep(ctx, stream)
    if which in ("all", "describe"):
        describe_sweep(ctx, stream)
    if which in ("all", "frontend"):
        frontend_sweep(ctx, stream)
    if which in ("all", "sift_detect"):
        sift_detect_sweep(ctx, stream)
    if which in ("all", "orb_detect"):
        orb_detect_sweep(ctx, stream)
    if which in ("match_real_one",):        # two calls at one size, for a launch list under ncu
        n, dim = int(sys.argv[2]), int(sys.argv[3])
        dev = torch.device("cuda", 0)
        g = torch.Generator(device=dev).manual_seed(1007)
        q = torch.randn((n, dim), device=dev, generator=g); t = torch.randn((n, dim), device=dev, generator=g)
        q = q / q.norm(dim=1, keepdim=True); t = t / t.norm(dim=1, keepdim=True)
        idx = torch.empty((n, 2), dtype=torch.int32, device=dev); dist = torch.empty((n, 2), dtype=torch.float32, device=dev)
        for _ in range(2):
            ctx.match_knn2_f32_dev(q.data_ptr(), n, t.data_ptr(), n, dim, idx.data_ptr(), dist.data_ptr())
        stream.synchronize()
        print(json.dumps({"n": n, "dim": dim, "queries_decided_by_exact_path": int(ctx.get_option("matcher_exact_fallback"))}))
    if which in ("match_sizes",):           # the integer tensor matcher with its defaults over the sizes of the sweeps, one line each (A/B of library variants: FM3D_LIB)
        dev = torch.device("cuda", 0)
        g = torch.Generator(device=dev).manual_seed(1003)
        for nq, nt in ((300, 360), (1000, 1200), (5000, 6000), (5000, 48000), (10000, 10000), (20000, 24000), (50000, 50000), (100000, 100000), (200000, 200000)):
            t = torch.randint(0, 256, (nt, 128), device=dev, generator=g).float()
            q = (t[torch.randperm(nt, device=dev, generator=g)[:nq]] + torch.randint(-6, 7, (nq, 128), device=dev, generator=g).float()).clamp_(0, 255)
            idx = torch.empty((nq, 2), dtype=torch.int32, device=dev); dist = torch.empty((nq, 2), dtype=torch.float32, device=dev)
            torch.cuda.synchronize()
            ms = min(timed(stream, lambda: ctx.match_knn2_f32_dev(q.data_ptr(), nq, t.data_ptr(), nt, 128, idx.data_ptr(), dist.data_ptr()), 5 if nq <= 50000 else 3) for _ in range(3))
            print(json.dumps({"case": "match_f32_sift128_tcgen05", "lib": os.environ.get("FM3D_LIB", "default"), "nq": nq, "nt": nt, "ms": ms,
                              "tflops": 256.0 * nq * nt / (ms * 1e-3) / 1e12, "checksum": [int(idx.long().sum().item()), float(dist.double().sum().item())]}), flush=True)
    if which in ("splits",):
        match_splits_sweep(ctx, stream)
    if which in ("c3",):
        c3_pipeline(ctx, stream)
    ctx.close()
