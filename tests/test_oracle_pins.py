"""Pins of the oracles themselves (CPU): the LM restatement against MINPACK (scipy), the C port
against the cv2 restatement when cv2 is importable, and the product's LM state machine
(csrc/fm3d_lm2.h, compiled for the host) against the oracle's generic lmmin."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from common import orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _expsin(seed):
    rng = np.random.default_rng(seed)
    t = np.linspace(0, 3, 60)
    p = np.array([1.3, -0.7, 0.5, 2.1])
    y = p[0] * np.exp(p[1] * t) + p[2] * np.sin(p[3] * t) + 0.01 * rng.standard_normal(t.size)
    return t, y, np.array([1.0, -1.0, 1.0, 2.0])


def test_lm_restatement_reproduces_minpack_lmdif():
    leastsq = pytest.importorskip("scipy.optimize").leastsq
    tol = 30 * 2.220446049250313e-16
    for seed in range(4):
        t, y, x0 = _expsin(seed)
        f = lambda p: p[0] * np.exp(p[1] * t) + p[2] * np.sin(p[3] * t) - y  # noqa: E731
        xs, _, info, _, ier = leastsq(f, x0, full_output=True, ftol=tol, xtol=tol, gtol=tol, maxfev=500, epsfcn=1e-10, factor=100)
        x, nfev, inf = orc.lmmin_expsin(t, y, x0, epsilon=1e-10, patience=100, minpack_mode=1)
        # identical iterates; the final ftol test (30 eps) sits at the rounding noise of |f|, so the
        # implementations may stop one or two outer iterations (n+1 evaluations each) apart
        assert abs(nfev - info["nfev"]) <= 20 and inf in (1, 2, 3) and ier in (1, 2, 3)
        np.testing.assert_allclose(x, xs, rtol=0, atol=1e-9)
        from oracle.lmmin_py import LMControl, lmmin
        xp, st = lmmin(4, x0, t.size, lambda p: (f(p), 0), LMControl(epsilon=1e-10), minpack_mode=True, keep_trace=True)
        assert abs(st.nfev - info["nfev"]) <= 20 and st.info in (1, 2, 3)
        np.testing.assert_allclose(xp, xs, rtol=0, atol=1e-9)
        assert [tr[0] for tr in st.trace[:5]] == [6, 11, 16, 21, 26]       # lmdif's evaluation schedule
        xl, nfl, infl = orc.lmmin_expsin(t, y, x0, epsilon=1e-10, patience=100, minpack_mode=0)
        xpl, stl = lmmin(4, x0, t.size, lambda p: (f(p), 0), LMControl(epsilon=1e-10), minpack_mode=False)
        assert abs(nfl - stl.nfev) <= 20 and infl in (1, 2, 3)             # C and Python restatements of lmfit agree
        np.testing.assert_allclose(xl, xpl, rtol=0, atol=1e-9)
    # exact agreement where the stopping test is not at the noise floor: Rosenbrock stops on gtol
    ros = lambda p: np.array([10 * (p[1] - p[0] ** 2), 1 - p[0], 0.0])  # noqa: E731
    xs, _, info, _, ier = leastsq(ros, [-1.2, 1.0], full_output=True, ftol=tol, xtol=tol, gtol=tol, maxfev=300, epsfcn=1e-10, factor=100)
    xp, st = lmmin(2, [-1.2, 1.0], 3, lambda p: (ros(p), 0), LMControl(epsilon=1e-10), minpack_mode=True)
    assert (st.nfev, st.info) == (info["nfev"], ier)
    np.testing.assert_allclose(xp, xs, rtol=0, atol=1e-12)


def test_lm_user_break_is_status_11():
    from oracle.lmmin_py import LMControl, lmmin
    calls = [0]

    def f(p):
        calls[0] += 1
        return (np.array([p[0] - 1, p[1] + 2, 0.1]), -1 if calls[0] == 4 else 0)
    x, st = lmmin(2, [0.0, 0.0], 3, f, LMControl())
    assert st.info == 11 and st.nfev == 4


def _harness():
    so = os.path.join(ROOT, "tests", "_build", "liblm2_harness.so")
    src = os.path.join(ROOT, "tests", "cpp", "lm2_harness.cpp")
    hdr = os.path.join(ROOT, "3dfeaturematcher_b200", "csrc", "fm3d_lm2.h")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        os.makedirs(os.path.dirname(so), exist_ok=True)
        gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.check_call([gxx, "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-I" + os.path.dirname(hdr), src, "-o", so])
    return C.CDLL(so)


def test_product_lm_state_machine_matches_oracle_lmmin():
    h = _harness()
    dp = C.POINTER(C.c_double)
    rng = np.random.default_rng(1)
    exact = 0
    for _ in range(12):
        t = np.linspace(0, 2, 200)
        xt = np.array([rng.uniform(0.5, 3), rng.uniform(-2, 1)])
        y = xt[0] * np.exp(xt[1] * t) + 0.05 * rng.standard_normal(t.size)
        x0 = xt + rng.normal(0, 0.5, 2)
        xo, nf, info = orc.lmmin_exp2(t, y, x0)
        x = x0.copy()
        nfev, inf = C.c_int(), C.c_int()
        h.lm2_run_exp2(t.ctypes.data_as(dp), y.ctypes.data_as(dp), t.size, x.ctypes.data_as(dp), C.c_double(1e-10), 100,
                       C.byref(nfev), C.byref(inf))
        # same iterates; the last ftol decision sits at the rounding noise of |f| (the harness sums J^T J,
        # the oracle factorises J), so the two may stop one outer iteration (3 evaluations) apart
        assert abs(nfev.value - nf) <= 6 and inf.value in (1, 2, 3) and info in (1, 2, 3)
        exact += (nfev.value, inf.value) == (nf, info)
        np.testing.assert_allclose(x, xo, rtol=0, atol=1e-8)
    assert exact >= 8


def test_c_port_against_cv2_restatement():
    pytest.importorskip("cv2")
    import importlib
    from oracle import oracle_cv as oc
    synth = importlib.import_module("3dfeaturematcher_b200.synth")
    case = synth.make_stereo_case(320, 240, 5, 31, pixels_ray=16, n_distractors=4)
    sc = case["scene"]
    cam = oc.Camera(sc.cam.K, sc.cam.dist, sc.cam.z_min, sc.cam.z_max, sc.cam.g12)
    rng = np.random.default_rng(0)
    pts = np.stack([rng.uniform(0, 320, 500), rng.uniform(0, 240, 500)], 1)
    np.testing.assert_allclose(orc.undistort_points(sc.cam.K, sc.cam.dist, pts), oc.undistort_points(cam, pts), rtol=1e-13, atol=1e-14)
    a, b, c = oc.triangulate(cam, case["kp1"], case["kp2_true"])
    a2, b2, c2 = orc.triangulate(sc.cam.K, sc.cam.dist, sc.cam.g12, sc.cam.z_min, sc.cam.z_max, case["kp1"], case["kp2_true"])
    np.testing.assert_allclose(a2, a, rtol=1e-11, atol=1e-13)
    np.testing.assert_array_equal(b2, b)
    odd = sc.img1[:237, :315]
    import cv2
    np.testing.assert_array_equal(orc.pyrdown(odd), cv2.pyrDown(odd))
    pyr1, pyr2 = oc.compute_pyramids(sc.img1, 2), oc.compute_pyramids(sc.img2, 2)
    for mode in (0, 2):
        rp = oc.optimize_normals(cam, a, pyr1, pyr2, 16, 1e-10, mode)
        rc = orc.optimize_normals(sc.cam.K, sc.cam.dist, sc.cam.g12, sc.cam.z_min, sc.cam.z_max, sc.img1, sc.img2, 2, a, 16, 1e-10,
                                  penalty_mode=mode)
        np.testing.assert_array_equal(rc["nfev"], rp["nfev"])
        np.testing.assert_array_equal(rc["status"], rp["status"])
        np.testing.assert_allclose(rc["normals"], rp["normals"], rtol=0, atol=1e-7)


def test_sift_patch_restatement_against_live_cv2():
    """oracle/sift_patch_np.py against cv2.SIFT_create().compute with the reference's keypoint
    (descriptorsmatcher.cpp:150-164), on fresh random patches."""
    pytest.importorskip("cv2")
    from oracle import oracle_cv as oc
    from oracle import sift_patch_np as sp
    rng = np.random.default_rng(99)
    import cv2
    for S in (128, 96, 50, 24, 10):
        patches = []
        for k in range(4):
            img = rng.integers(0, 256, (S, S)).astype(np.uint8)
            if k:
                img = cv2.normalize(cv2.GaussianBlur(img, (0, 0), float(k)), None, 0, 255, cv2.NORM_MINMAX)
            patches.append(img)
        patches = np.array(patches)
        want = oc.describe_patches_sift(patches)
        got = sp.describe_patches_sift(patches)
        diff = np.abs(want - got)
        assert diff.max() <= 1 and (diff == 0).mean() > 0.97
        # blur stage alone
        sigma = (1.6 * 1.6 - 0.25) ** 0.5
        np.testing.assert_allclose(sp.gaussian_blur_f32(patches[1], sigma),
                                   cv2.GaussianBlur(patches[1].astype(np.float32), (0, 0), sigma), rtol=0, atol=2e-4)


def test_fast_restatement_against_live_cv2():
    """oracle/fast_np.py against cv2.FastFeatureDetector (what cv::FastFeatureDetector(threshold, nonmax)
    of descriptorsmatcher.cpp:215-222 runs) on fresh random images, thresholds inside [0, 255]."""
    cv2 = pytest.importorskip("cv2")
    from oracle import fast_np as fo
    rng = np.random.default_rng(2024)
    for k, (h, w) in enumerate(((37, 53), (64, 64), (101, 77), (8, 200))):
        img = rng.integers(0, 256, (h, w)).astype(np.uint8)
        if k % 2:
            img = cv2.normalize(cv2.GaussianBlur(img, (0, 0), 1.2), None, 0, 255, cv2.NORM_MINMAX)
        for t in (0, 5, 17, 60, 254, 255):
            for nm in (False, True):
                kps = cv2.FastFeatureDetector_create(threshold=t, nonmaxSuppression=nm).detect(img, None)
                xy, r = fo.detect_fast(img, t, nm)
                np.testing.assert_array_equal(xy, np.array([p.pt for p in kps], np.float32).reshape(-1, 2))
                np.testing.assert_array_equal(r, np.array([p.response for p in kps], np.float32))


def test_sift_keypoint_restatement_against_live_cv2():
    """oracle/sift_kp_np.py against cv2.SIFT_create().compute on fresh images: FAST keypoints (size 7, angle -1,
    what DetectorType FAST + ExtractorType SIFT of descriptorsmatcher.cpp:215-222, :302-314 hands to compute) and
    keypoints with arbitrary size / angle, including positions on the image border."""
    cv2 = pytest.importorskip("cv2")
    from oracle import sift_kp_np as sk
    rng = np.random.default_rng(77)
    sift = cv2.SIFT_create()
    for h, w in ((90, 131), (64, 64)):
        img = cv2.normalize(cv2.GaussianBlur(rng.integers(0, 256, (h, w)).astype(np.uint8), (0, 0), 1.3), None, 0, 255, cv2.NORM_MINMAX)
        kps = cv2.FastFeatureDetector_create(threshold=15, nonmaxSuppression=True).detect(img, None)[:150]
        arr = np.array([[k.pt[0], k.pt[1], k.size, k.angle] for k in kps], np.float32).reshape(-1, 4)
        extra = np.stack([rng.uniform(0, w - 1, 60), rng.uniform(0, h - 1, 60), rng.uniform(1.5, 30, 60), rng.uniform(0, 360, 60)], 1).astype(np.float32)
        extra[:3, :2] = [[0, 0], [w - 1, h - 1], [w - 1, 0]]
        arr = np.concatenate([arr, extra])
        cvk = [cv2.KeyPoint(float(a[0]), float(a[1]), float(a[2]), float(a[3])) for a in arr]
        cvk2, want = sift.compute(img, cvk)
        assert len(cvk2) == len(cvk)
        got = sk.describe_keypoints_sift(img, arr)
        diff = np.abs(got - want)
        assert diff.max() <= 1 and (diff == 0).mean() >= 0.999


def test_brisk_restatement_against_live_cv2():
    """oracle/brisk_np.py against cv2.BRISK_create().compute on a fresh image: FAST keypoints (what DetectorType FAST +
    ExtractorType BRISK of descriptorsmatcher.cpp:215-222, :343-349 hands to compute) and sub-pixel keypoints of many
    sizes.  Survivors, bits and angles identical."""
    cv2 = pytest.importorskip("cv2")
    from oracle import brisk_np as bn
    rng = np.random.default_rng(78)
    img = cv2.normalize(cv2.GaussianBlur(rng.integers(0, 256, (150, 210)).astype(np.uint8), (0, 0), 1.6), None, 0, 255, cv2.NORM_MINMAX)
    kps = cv2.FastFeatureDetector_create(threshold=12, nonmaxSuppression=True).detect(img, None)[:250]
    arr = np.array([[k.pt[0], k.pt[1], k.size, k.angle] for k in kps], np.float32).reshape(-1, 4)
    extra = np.stack([rng.uniform(0, 209, 120), rng.uniform(0, 149, 120), rng.uniform(2, 30, 120), rng.uniform(0, 360, 120)], 1).astype(np.float32)
    arr = np.concatenate([arr, extra])
    cvk = [cv2.KeyPoint(float(a[0]), float(a[1]), float(a[2]), float(a[3])) for a in arr]
    cvk2, want = cv2.BRISK_create(25, 0).compute(img, cvk)
    kept, ang, got = bn.describe_keypoints_brisk(img, arr)
    assert len(kept) == len(cvk2) > 150
    np.testing.assert_array_equal(arr[kept, :2], np.array([k.pt for k in cvk2], np.float32))
    np.testing.assert_array_equal(got, want)
    np.testing.assert_array_equal(ang, np.array([k.angle for k in cvk2], np.float32))


def test_orb_restatement_against_live_cv2():
    """oracle/orb_np.py against cv2.ORB_create().compute on a fresh image (FAST keypoints + sub-pixel keypoints with
    arbitrary angles), and the recovered pattern against a fresh recovery."""
    cv2 = pytest.importorskip("cv2")
    from oracle import orb_np as on
    rng = np.random.default_rng(79)
    img = cv2.normalize(cv2.GaussianBlur(rng.integers(0, 256, (170, 230)).astype(np.uint8), (0, 0), 1.2), None, 0, 255, cv2.NORM_MINMAX)
    kps = cv2.FastFeatureDetector_create(threshold=12, nonmaxSuppression=True).detect(img, None)
    kps = kps[::max(1, len(kps) // 300)]                 # row-major list: spread over the whole frame
    arr = np.array([[k.pt[0], k.pt[1], k.size, k.angle] for k in kps], np.float32).reshape(-1, 4)
    extra = np.stack([rng.uniform(0, 229, 200), rng.uniform(0, 169, 200), rng.uniform(2, 30, 200), rng.uniform(0, 360, 200)], 1).astype(np.float32)
    arr = np.concatenate([arr, extra])
    cvk = [cv2.KeyPoint(float(a[0]), float(a[1]), float(a[2]), float(a[3]), 0.0, 0, i) for i, a in enumerate(arr)]
    cvk2, want = cv2.ORB_create().compute(img, cvk)
    kept, got = on.describe_keypoints_orb(img, arr)
    np.testing.assert_array_equal(kept, np.array([k.class_id for k in cvk2]))
    ham = np.unpackbits(got ^ want, axis=1).sum(1)
    assert len(kept) > 150 and (ham == 0).mean() >= 0.995 and ham.max() <= 2      # a blurred value at a rounding boundary
    import importlib.util
    spec = importlib.util.spec_from_file_location("recover_orb_pattern", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                                                      "tools", "recover_orb_pattern.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    np.testing.assert_array_equal(mod.recover(), on.PATTERN)
