"""Writes tests/golden/*.npz: inputs and outputs of the reference's arithmetic as computed by the
OpenCV routines the reference calls (cv2 4.13 through oracle/oracle_cv.py).  Run in the build
container (cv2 is needed); the fixtures are committed so that tests never depend on cv2 or on
/root/reference at run time.

    python tools/make_golden.py
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle_cv as oc  # noqa: E402

synth = importlib.import_module("3dfeaturematcher_b200.synth")
OUT = os.path.join(ROOT, "tests", "golden")


def primitives():
    rng = np.random.default_rng(20261018)
    # the reference's own camera (build/settings.yml) and the pose of its two frames
    K = np.array([[572.4765, 0, 549.75189], [0, 572.69354, 411.68039], [0, 0, 1.0]])
    dist = synth.SETTINGS_DIST
    cam = oc.Camera(K, dist, 1.5, 2.4)
    pos1 = [5.301099, 8.031408, 1.977258, 0.153433, 0.149941, -2.658648]
    pos2 = [4.735536, 7.691893, 1.913166, 0.252828, 0.048977, -2.676886]
    g12 = cam.setg12(pos1[:3], pos2[:3], pos1[3:], pos2[3:], synth.SETTINGS_RODRIGUES_IC, synth.SETTINGS_TRANSLATION_IC)
    pts = np.stack([rng.uniform(0, 1024, 400), rng.uniform(0, 768, 400)], 1)
    und = oc.undistort_points(cam, pts)
    X = np.stack([rng.uniform(-1.2, 1.2, 300), rng.uniform(-0.9, 0.9, 300), rng.uniform(1.5, 2.4, 300)], 1)
    r2, t2 = oc.decompose_transformation(g12)
    proj1 = oc.project_points(cam, X, np.zeros(3), np.zeros(3))
    proj2 = oc.project_points(cam, X, r2, t2)
    # matched keypoints = projections of X (+ noise, + gross outliers for the depth gate)
    kp1 = proj1.astype(np.float32)
    kp2 = (proj2 + rng.normal(0, 0.4, proj2.shape)).astype(np.float32)
    kp2[::9, 0] += 35
    xyz_all, mask, xyz = oc.triangulate(cam, kp1, kp2)
    img = rng.integers(0, 256, (53, 71), dtype=np.uint8)
    pyr = oc.compute_pyramids(img, 3)
    q, t, _ = synth.make_float_descriptors(60, 30, 3)
    t[5] = t[11] = t[40]
    q[0] = t[40]
    idx_f, dist_f = oc.knn2(q, t, False)
    qb, tb, _ = synth.make_binary_descriptors(60, 30, 4)
    tb[2] = tb[8] = tb[33]
    qb[0] = tb[33]
    idx_b, dist_b = oc.knn2(qb, tb, True)
    qi, ti, d = oc.nndr_filter(idx_f, dist_f, 0.55)
    np.savez_compressed(os.path.join(OUT, "primitives.npz"), K=K, dist=dist, g12=g12, pos1=pos1, pos2=pos2,
                        pts=pts, undistorted=und, X=X, proj1=proj1, proj2=proj2, kp1=kp1, kp2=kp2, xyz_all=xyz_all,
                        mask=mask, xyz=xyz, img=img, pyr1=pyr[1], pyr2=pyr[2], pyr3=pyr[3], q=q, t=t, idx_f=idx_f,
                        dist_f=dist_f, qb=qb, tb=tb, idx_b=idx_b, dist_b=dist_b, nndr_q=qi, nndr_t=ti, nndr_d=d)


def normals():
    case = synth.make_stereo_case(320, 240, 6, 77, pixels_ray=16, n_distractors=4)
    sc = case["scene"]
    cam = oc.Camera(sc.cam.K, sc.cam.dist, sc.cam.z_min, sc.cam.z_max, sc.cam.g12)
    pyr1, pyr2 = oc.compute_pyramids(sc.img1, 2), oc.compute_pyramids(sc.img2, 2)
    xyz_all, mask, xyz = oc.triangulate(cam, case["kp1"], case["kp2_true"])
    out = {"img1": sc.img1, "img2": sc.img2, "K": sc.cam.K, "dist": sc.cam.dist, "g12": sc.cam.g12,
           "zmin": sc.cam.z_min, "zmax": sc.cam.z_max, "kp1": case["kp1"], "kp2": case["kp2_true"], "xyz": xyz_all,
           "gt_normal": case["normal"]}
    for mode, name in ((oc.PENALTY_FABS, "fabs"), (oc.PENALTY_INT_ABS, "intabs"), (oc.PENALTY_OFF, "off")):
        r = oc.optimize_normals(cam, xyz_all, pyr1, pyr2, 16, 1e-10, mode)
        for k in ("normals", "status", "nfev", "npenalty", "cost", "m"):
            out[f"{name}_{k}"] = r[k]
    pt = np.array([oc.car2sph(x / np.linalg.norm(x)) for x in xyz_all]) + 0.1
    for lvl in range(3):
        out[f"cost_l{lvl}"] = np.array([oc.evaluate_cost(cam, x, p, pyr1, pyr2, 16, lvl, oc.PENALTY_OFF)[0]
                                        for x, p in zip(xyz_all, pt)])
    out["cost_pt"] = pt
    g = oc.gravity_from_settings(synth.SETTINGS_RODRIGUES_IC)
    frames = oc.feature_frames(xyz_all, out["off_normals"], g)
    ref = oc.reference_squared_neighborhood(0.05, 0.25)
    patches, ips = zip(*[oc.project_reference_points(cam, sc.img1, ref, F) for F in frames])
    out.update(gravity=g, frames=frames, patches=np.array(patches), image_points=np.array(ips))
    np.savez_compressed(os.path.join(OUT, "normals.npz"), **out)


def sift_test_patches(S, count, seed):
    """Patches for the SIFT-descriptor fixtures: band-limited textures, white noise, a step edge, a
    constant patch (all-zero histogram) and a linear ramp (one orientation bin)."""
    import cv2
    rng = np.random.default_rng(seed)
    out = []
    for k in range(count):
        img = rng.integers(0, 256, (S, S)).astype(np.uint8)
        if k % 4 != 3:      # every fourth patch stays white noise
            img = cv2.GaussianBlur(img, (0, 0), 1.0 + (k % 3))
            img = cv2.normalize(img, None, 0, 255, cv2.NORM_MINMAX)
        out.append(img)
    edge = np.zeros((S, S), np.uint8)
    edge[:, S // 2 + 1:] = 200
    ramp = np.clip(np.add.outer(np.arange(S) * 1.5, np.arange(S) * 0.5), 0, 255).astype(np.uint8)
    out += [edge, np.full((S, S), 77, np.uint8), ramp]
    return np.array(out)


def sift_patches():
    """K9 fixtures: cv2.SIFT_create().compute with the keypoint of extractDescriptorsFromPatches
    (descriptorsmatcher.cpp:150-157) on patches of several edge lengths, and on the rectified
    patches of normals.npz."""
    out = {}
    for S, count in ((128, 3), (64, 4), (40, 4), (16, 4), (8, 2)):
        p = sift_test_patches(S, count, 5000 + S)
        out[f"p{S}"] = p
        out[f"d{S}"] = oc.describe_patches_sift(p)
    pipe = np.load(os.path.join(OUT, "normals.npz"))["patches"]
    out["d_pipeline"] = oc.describe_patches_sift(pipe)      # patches themselves live in normals.npz
    np.savez_compressed(os.path.join(OUT, "sift_patches.npz"), **out)


def fast_test_images():
    """Images for the FAST fixtures: white noise with odd edge lengths (dense corners, every border case),
    a band-limited texture, a rendered synthetic frame, the smallest image that has an interior pixel and
    one that has none."""
    import cv2
    rng = np.random.default_rng(7100)
    blur = cv2.GaussianBlur(rng.integers(0, 256, (120, 160)).astype(np.uint8), (0, 0), 1.5)
    return {"noise": rng.integers(0, 256, (61, 83)).astype(np.uint8),
            "blur": cv2.normalize(blur, None, 0, 255, cv2.NORM_MINMAX),
            "frame": synth.make_stereo_case(200, 152, 6, 99, pixels_ray=8, n_distractors=10)["scene"].img1,
            "tiny7": rng.integers(0, 2, (7, 7)).astype(np.uint8) * 255,
            "tiny6": rng.integers(0, 256, (6, 9)).astype(np.uint8)}


FAST_THRESHOLDS = (0, 1, 10, 20, 40, 100, 255)


def fast_keypoints():
    """K10 fixtures: cv2.FastFeatureDetector (TYPE_9_16, what cv::FastFeatureDetector(threshold, nonmax) of
    descriptorsmatcher.cpp:215-222 runs) on the images above, thresholds over the whole valid range, with
    and without non-maximum suppression.  xy / response in cv2's output order."""
    import cv2
    out = {}
    for name, img in fast_test_images().items():
        out[f"img_{name}"] = img
        for t in FAST_THRESHOLDS:
            for nm in (0, 1):
                kps = cv2.FastFeatureDetector_create(threshold=t, nonmaxSuppression=bool(nm)).detect(img, None)
                assert all(k.size == 7.0 and k.angle == -1.0 and k.octave == 0 and k.class_id == -1 for k in kps)
                out[f"xy_{name}_{t}_{nm}"] = np.array([k.pt for k in kps], np.float32).reshape(-1, 2).astype(np.uint16)
                out[f"r_{name}_{t}_{nm}"] = np.array([k.response for k in kps], np.float32).astype(np.uint8)
    np.savez_compressed(os.path.join(OUT, "fast_keypoints.npz"), **out)


def sift_general_keypoints(w, h, count, seed):
    """Keypoints with real sizes and angles (what a caller can inject), some on the image border, some with
    angle -1 / 0 / 360-epsilon."""
    rng = np.random.default_rng(seed)
    k = np.stack([rng.uniform(0, w - 1, count), rng.uniform(0, h - 1, count), rng.uniform(2, 24, count),
                  rng.uniform(0, 360, count)], 1).astype(np.float32)
    k[:count // 10, 3] = -1
    k[count // 10:count // 10 + 5, 3] = 0
    k[count // 10 + 5:count // 10 + 8, 3] = 359.99997
    k[-4:, :2] = [[0, 0], [w - 1, h - 1], [0.4, h - 1.4], [w - 1, 3]]
    return k


def sift_keypoints():
    """K11 fixtures: cv2.SIFT_create().compute (descriptor_extractor_->compute of descriptorsmatcher.cpp:114-115
    with ExtractorType SIFT) on the FAST keypoints (threshold 20, non-maximum suppression, at most 300 per
    image) of the images of fast_keypoints.npz, and on keypoints with real sizes / angles.  The images
    themselves live in fast_keypoints.npz."""
    import cv2
    sift = cv2.SIFT_create()
    out = {}
    for name, img in fast_test_images().items():
        kps = cv2.FastFeatureDetector_create(threshold=20, nonmaxSuppression=True).detect(img, None)[:300]
        sets = {"fast": np.array([[k.pt[0], k.pt[1], k.size, k.angle] for k in kps], np.float32).reshape(-1, 4)}
        if name in ("blur", "frame"):
            sets["general"] = sift_general_keypoints(img.shape[1], img.shape[0], 120, 7200 + len(name))
        for tag, arr in sets.items():
            cvk = [cv2.KeyPoint(float(a[0]), float(a[1]), float(a[2]), float(a[3])) for a in arr]
            if cvk:
                cvk2, d = sift.compute(img, cvk)
                assert len(cvk2) == len(cvk)            # nothing filtered: row k belongs to keypoint k
            else:
                d = np.zeros((0, 128), np.float32)
            assert np.array_equal(d, np.rint(d)) and d.min(initial=0) >= 0 and d.max(initial=0) <= 255
            out[f"k_{name}_{tag}"] = arr
            out[f"d_{name}_{tag}"] = d.astype(np.uint8)
    np.savez_compressed(os.path.join(OUT, "sift_keypoints.npz"), **out)


def brisk_keypoints():
    """K12 fixtures: cv2.BRISK_create(25, 0).compute (descriptor_extractor_->compute of descriptorsmatcher.cpp:114-115
    with ExtractorType BRISK, :343-349, knobs of build/settings.yml) on the FAST keypoints (threshold 20, non-maximum
    suppression, at most 400 per image) of the images of fast_keypoints.npz, and on keypoints with sub-pixel
    positions and sizes 3 ... 36 (scale indices 0 ... 30).  cv::BRISK removes keypoints near the border: `kept`
    holds the indices of the survivors, `d` / `a` their rows and angles."""
    import cv2
    brisk = cv2.BRISK_create(25, 0)
    out = {}
    for name, img in fast_test_images().items():
        if name.startswith("tiny"):
            continue
        kps = cv2.FastFeatureDetector_create(threshold=20, nonmaxSuppression=True).detect(img, None)[:400]
        sets = {"fast": np.array([[k.pt[0], k.pt[1], k.size, k.angle] for k in kps], np.float32).reshape(-1, 4)}
        if name in ("blur", "frame"):
            g = sift_general_keypoints(img.shape[1], img.shape[0], 160, 7300 + len(name))
            g[:, 2] = np.random.default_rng(7400).uniform(3, 36, len(g)).astype(np.float32)
            sets["general"] = g
        for tag, arr in sets.items():
            cvk = [cv2.KeyPoint(float(a[0]), float(a[1]), float(a[2]), float(a[3])) for a in arr]
            cvk2, d = brisk.compute(img, cvk)
            d = np.zeros((0, 64), np.uint8) if d is None else d
            # survivors keep their order: recover their indices from the positions
            kept, j = [], 0
            for i, a in enumerate(arr):
                if j < len(cvk2) and cvk2[j].pt == (float(a[0]), float(a[1])) and cvk2[j].size == float(a[2]):
                    kept.append(i)
                    j += 1
            assert j == len(cvk2) == len(d)
            out[f"k_{name}_{tag}"] = arr
            out[f"kept_{name}_{tag}"] = np.array(kept, np.int32)
            out[f"d_{name}_{tag}"] = d
            out[f"a_{name}_{tag}"] = np.array([k.angle for k in cvk2], np.float32)
    np.savez_compressed(os.path.join(OUT, "brisk_keypoints.npz"), **out)


def orb_keypoints():
    """K13 fixtures: cv2.ORB_create().compute (descriptor_extractor_->compute of descriptorsmatcher.cpp:114-115 with
    ExtractorType ORB, :336-342) on the FAST keypoints (threshold 20, non-maximum suppression, at most 500 per image) of
    the images of fast_keypoints.npz and on octave-0 keypoints with sub-pixel positions and arbitrary angles.  cv::ORB
    removes keypoints within 31 pixels of the border: `kept` holds the indices of the survivors (carried through
    KeyPoint::class_id), `d` their rows."""
    import cv2
    orb = cv2.ORB_create()
    out = {}
    for name, img in fast_test_images().items():
        if name.startswith("tiny"):
            continue
        kps = cv2.FastFeatureDetector_create(threshold=20, nonmaxSuppression=True).detect(img, None)[:500]
        sets = {"fast": np.array([[k.pt[0], k.pt[1], k.size, k.angle] for k in kps], np.float32).reshape(-1, 4)}
        if name in ("blur", "frame"):
            sets["general"] = sift_general_keypoints(img.shape[1], img.shape[0], 400, 7500 + len(name))
        for tag, arr in sets.items():
            cvk = [cv2.KeyPoint(float(a[0]), float(a[1]), float(a[2]), float(a[3]), 0.0, 0, i) for i, a in enumerate(arr)]
            cvk2, d = orb.compute(img, cvk)
            d = np.zeros((0, 32), np.uint8) if d is None else d
            kept = np.array([k.class_id for k in cvk2], np.int32)
            assert (np.diff(kept) > 0).all() and len(kept) == len(d)          # survivors keep their order
            out[f"k_{name}_{tag}"] = arr
            out[f"kept_{name}_{tag}"] = kept
            out[f"d_{name}_{tag}"] = d
    # extractDescriptorsFromPatches (descriptorsmatcher.cpp:133-174) with ExtractorType ORB: one keypoint per patch
    for S, count in ((128, 3), (64, 4), (63, 2)):
        p = sift_test_patches(S, count, 6000 + S)
        rows = []
        for patch in p:
            kp = cv2.KeyPoint(float(S // 2), float(S // 2), float(S), -1.0, 1.0, 0, 0)
            k2, d = orb.compute(patch, [kp])
            assert len(k2) == 1
            rows.append(d[0])
        out[f"p{S}"] = p
        out[f"dp{S}"] = np.array(rows, np.uint8)
    np.savez_compressed(os.path.join(OUT, "orb_keypoints.npz"), **out)


def sift_detect_test_images():
    """Frames for the SIFT detector fixtures: band-limited noise, a rendered synthetic frame, an odd-sized resample of it."""
    import cv2
    rng = np.random.default_rng(7400)
    b = cv2.GaussianBlur(rng.integers(0, 256, (120, 160)).astype(np.float32), (0, 0), 2.0)
    frame = synth.make_stereo_case(320, 240, 8, 99, pixels_ray=8, n_distractors=10)["scene"].img1
    return {"blobs": ((b - b.min()) / (b.max() - b.min()) * 255).astype(np.uint8), "frame": frame,
            "odd": cv2.resize(frame, (211, 137), interpolation=cv2.INTER_AREA)}


SIFT_DETECT_PARAMS = {"default": dict(), "best50": dict(nfeatures=50),
                      "knobs": dict(nOctaveLayers=4, contrastThreshold=0.03, edgeThreshold=8.0, sigma=1.4)}


def sift_detect():
    """K14 fixtures: cv2.SIFT_create(...).detect (what cv::SIFT(NumFeatures, NumOctaveLayers, ContrastThreshold, EdgeThreshold,
    Sigma) of descriptorsmatcher.cpp:243-256 runs) on the frames above: rows (x, y, size, angle, response) and the packed octave,
    in cv2's output order (KeyPointsFilter::removeDuplicatedSorted; retainBest leaves its own order: compare as a set)."""
    import cv2
    out = {}
    for name, img in sift_detect_test_images().items():
        out[f"img_{name}"] = img
        for pname, kw in SIFT_DETECT_PARAMS.items():
            kps = cv2.SIFT_create(**kw).detect(img, None)
            out[f"kp_{name}_{pname}"] = np.array([[k.pt[0], k.pt[1], k.size, k.angle, k.response] for k in kps], np.float32).reshape(-1, 5)
            out[f"oct_{name}_{pname}"] = np.array([k.octave for k in kps], np.int32)
            if pname != "best50":       # and what descriptor_extractor_->compute returns for them (K11 on the pyramid layers)
                kps2, desc = cv2.SIFT_create(**kw).compute(img, kps)
                assert len(kps2) == len(kps)
                out[f"desc_{name}_{pname}"] = desc.astype(np.uint8)
    np.savez_compressed(os.path.join(OUT, "sift_detect.npz"), **out)


ORB_DETECT_PARAMS = {"default": dict(), "knobs": dict(nfeatures=1500, scaleFactor=1.3, nlevels=5)}


def orb_detect():
    """K15 fixtures: cv2.ORB_create(nfeatures, scaleFactor, nlevels).detectAndCompute (what cv::ORB(NumFeatures, ScaleFactor,
    NumLevels) of descriptorsmatcher.cpp:273-279 / :336-342 runs as detector and extractor) on two frames: rows (x, y, size,
    angle, response), the octave and the 32-byte descriptors, sorted by (octave, y, x) -- OpenCV's own order is whatever
    std::nth_element leaves."""
    import cv2
    rng = np.random.default_rng(7500)
    b = cv2.GaussianBlur(rng.integers(0, 256, (240, 320)).astype(np.float32), (0, 0), 1.5)
    imgs = {"blobs": ((b - b.min()) / (b.max() - b.min()) * 255).astype(np.uint8),
            "frame": synth.make_stereo_case(320, 240, 8, 99, pixels_ray=8, n_distractors=10)["scene"].img1}
    out = {}
    for name, img in imgs.items():
        out[f"img_{name}"] = img
        for pname, kw in ORB_DETECT_PARAMS.items():
            kps, desc = cv2.ORB_create(**kw).detectAndCompute(img, None)
            K = np.array([[k.pt[0], k.pt[1], k.size, k.angle, k.response, k.octave] for k in kps], np.float64).reshape(-1, 6)
            order = np.lexsort((K[:, 0], K[:, 1], K[:, 5]))
            out[f"kp_{name}_{pname}"] = K[order, :5].astype(np.float32)
            out[f"oct_{name}_{pname}"] = K[order, 5].astype(np.int32)
            out[f"desc_{name}_{pname}"] = desc[order]
    np.savez_compressed(os.path.join(OUT, "orb_detect.npz"), **out)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    if "--orb-only" in sys.argv:
        orb_keypoints()
    elif "--brisk-only" in sys.argv:
        brisk_keypoints()
    elif "--sift-kp-only" in sys.argv:
        sift_keypoints()
    elif "--fast-only" in sys.argv:
        fast_keypoints()
    elif "--sift-detect-only" in sys.argv:
        sift_detect()
    elif "--orb-detect-only" in sys.argv:
        orb_detect()
    else:
        if "--sift-only" not in sys.argv:
            primitives()
            normals()
        sift_patches()
        fast_keypoints()
        sift_keypoints()
        brisk_keypoints()
        orb_keypoints()
        sift_detect()
        orb_detect()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
