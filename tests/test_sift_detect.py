"""K14 -- cv::SIFT's scale-space detector (DetectorType SIFT, DescriptorsMatcher/descriptorsmatcher.cpp:243-256, called at
:110-111): the oracle restatement against the committed outputs of cv2.SIFT_create(...).detect, and the GPU detector
(fm3d_detect_sift) against the same golden vectors and, where cv2 is importable, against cv2 itself on larger frames.

Float work whose summation order is OpenCV's own (SIMD blurs, histogram sums): a keypoint is "the same" when position agrees
to 0.01 px, size to 1e-3 relative, angle to 0.1 degree, response to 1e-4, octave and layer exactly and the third byte of the
packed octave (the rounded sub-layer offset, which nothing downstream reads) to +-1; a candidate within rounding of one of
the detector's thresholds may fall on the other side, so the bar is the matched FRACTION (>= 99 % both ways) and the count."""
import importlib
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PARAMS = {"default": dict(), "best50": dict(nfeatures=50),
          "knobs": dict(n_octave_layers=4, contrast_threshold=0.03, edge_threshold=8.0, sigma=1.4)}


def matched_fraction(A, B, tol_px=0.01, tol_size=1e-3, tol_ang=0.1):
    """fraction of the rows (x, y, size, angle, response, octave) of A that have a partner in B"""
    if len(A) == 0:
        return 1.0
    hit = 0
    for a in A:
        near = np.nonzero(np.abs(B[:, :2] - a[:2]).max(1) < tol_px)[0]
        for j in near:
            da = abs(B[j, 3] - a[3])
            oa, ob = int(a[5]), int(B[j, 5])          # octave | layer << 8 | cvRound((xi + 0.5) * 255) << 16
            if abs(B[j, 2] - a[2]) <= tol_size * max(1.0, a[2]) and min(da, 360 - da) < tol_ang and (oa & 0xffff) == (ob & 0xffff) \
                    and abs((oa >> 16) - (ob >> 16)) <= 1 and abs(B[j, 4] - a[4]) <= 1e-4:
                hit += 1
                break
    return hit / len(A)


def golden_rows(g, name, pname):
    return np.column_stack([g[f"kp_{name}_{pname}"].astype(np.float64), g[f"oct_{name}_{pname}"].astype(np.float64)])


def test_sift_detector_restatement_against_cv2_golden_vectors():
    from oracle import sift_detect_np as sd
    g = np.load(os.path.join(GOLD, "sift_detect.npz"))
    total = 0
    for name in ("blobs", "frame", "odd"):
        for pname, kw in PARAMS.items():
            want = golden_rows(g, name, pname)
            got = sd.detect_sift(g[f"img_{name}"], **kw)
            assert abs(len(got) - len(want)) <= max(1, len(want) // 200), (name, pname, len(got), len(want))
            assert matched_fraction(want, got) >= 0.99 and matched_fraction(got, want) >= 0.99, (name, pname)
            if pname != "best50" and len(got) == len(want):       # removeDuplicatedSorted's order
                assert np.abs(got[:, :2] - want[:, :2]).max() < 0.01
            total += len(want)
    assert total > 4000


@pytest.fixture(scope="module")
def ctx():
    api = importlib.import_module("3dfeaturematcher_b200.api")
    c = api.Context(0)
    yield c
    c.close()


@pytest.mark.gpu
def test_gpu_sift_detector_against_cv2_golden_vectors(ctx):
    g = np.load(os.path.join(GOLD, "sift_detect.npz"))
    for name in ("blobs", "frame", "odd"):
        for pname, kw in PARAMS.items():
            want = golden_rows(g, name, pname)
            got = ctx.detect_sift(g[f"img_{name}"], **kw)
            again = ctx.detect_sift(g[f"img_{name}"], **kw)
            np.testing.assert_array_equal(got, again)                       # reproducible, order included
            assert abs(len(got) - len(want)) <= max(1, len(want) // 200), (name, pname, len(got), len(want))
            f1, f2 = matched_fraction(want, got), matched_fraction(got, want)
            print(f"SIFT detect {name}/{pname}: cv2 {len(want)} gpu {len(got)} matched {f1:.4f} / {f2:.4f}")
            assert f1 >= 0.99 and f2 >= 0.99, (name, pname, f1, f2)


@pytest.mark.gpu
def test_gpu_sift_detector_against_cv2_on_a_720p_frame(ctx):
    cv2 = pytest.importorskip("cv2")
    synth = importlib.import_module("3dfeaturematcher_b200.synth")
    img = synth.make_stereo_case(1280, 720, 50, 1001, pixels_ray=32)["scene"].img1
    want = np.array([[k.pt[0], k.pt[1], k.size, k.angle, k.response, k.octave] for k in cv2.SIFT_create().detect(img, None)])
    got = ctx.detect_sift(img)
    assert len(want) > 5000 and abs(len(got) - len(want)) <= len(want) // 200
    sel = np.random.default_rng(5).choice(len(want), 1500, replace=False)
    assert matched_fraction(want[sel], got) >= 0.99
    sel = np.random.default_rng(6).choice(len(got), 1500, replace=False)
    assert matched_fraction(got[sel], want) >= 0.99


@pytest.mark.gpu
def test_gpu_sift_detector_edge_cases(ctx, ):
    """frames without structure give no keypoints; tiny frames (fewer octaves than layers need) and invalid arguments"""
    api = importlib.import_module("3dfeaturematcher_b200.api")
    assert len(ctx.detect_sift(np.full((64, 80), 93, np.uint8))) == 0
    assert ctx.detect_sift(np.random.default_rng(1).integers(0, 256, (9, 11)).astype(np.uint8)).shape[1] == 6
    with pytest.raises(api.Fm3dError):
        ctx.detect_sift(np.zeros((1, 50), np.uint8))
    with pytest.raises(api.Fm3dError):
        ctx.detect_sift(np.zeros((50, 50), np.uint8), n_octave_layers=9)


@pytest.mark.gpu
def test_gpu_sift_descriptors_of_sift_keypoints_against_cv2_golden_vectors(ctx):
    """descriptor_extractor_->compute (ExtractorType SIFT) for the detector's own keypoints: every descriptor is read from the
    Gaussian image of the keypoint's octave and layer (fm3d_describe_keypoints_sift_oct).  Golden: cv2.SIFT.compute on cv2's
    keypoints.  Quantised values: equal up to +-1 on isolated entries (summation order of the blurs and of the histogram)."""
    g = np.load(os.path.join(GOLD, "sift_detect.npz"))
    for name in ("blobs", "frame", "odd"):
        for pname, kw in PARAMS.items():
            if pname == "best50":
                continue
            kp, oc, want = g[f"kp_{name}_{pname}"], g[f"oct_{name}_{pname}"], g[f"desc_{name}_{pname}"].astype(np.float32)
            got = ctx.describe_keypoints_sift_oct(g[f"img_{name}"], kp[:, :4], oc, n_octave_layers=kw.get("n_octave_layers", 3),
                                                  sigma=kw.get("sigma", 1.6))
            d = np.abs(got - want)
            rows_close = (d.max(1) <= 2).mean()
            print(f"SIFT describe {name}/{pname}: {len(kp)} keypoints, values equal {np.mean(d == 0):.4f}, |diff| <= 1 {np.mean(d <= 1):.4f}, "
                  f"rows within 2: {rows_close:.4f}, max {d.max():.0f}")
            assert np.mean(d == 0) >= 0.95 and np.mean(d <= 1) >= 0.995 and rows_close >= 0.98


@pytest.mark.gpu
def test_gpu_sift_detect_and_describe_on_one_pyramid_equals_the_two_calls(ctx):
    """fm3d_detect_and_describe_sift (cv::SIFT::detectAndCompute: the descriptors read from the pyramid the keypoints were found
    on) returns, bit for bit, the keypoints of fm3d_detect_sift and the rows fm3d_describe_keypoints_sift_oct computes for them on
    a pyramid of its own -- golden frames with three parameter sets (retainBest included) and a rendered 720p frame."""
    synth = importlib.import_module("3dfeaturematcher_b200.synth")
    g = np.load(os.path.join(GOLD, "sift_detect.npz"))
    cases = [(g[f"img_{name}"], kw) for name in ("blobs", "frame", "odd") for kw in PARAMS.values()]
    cases.append((synth.make_stereo_case(1280, 720, 20, 1001, pixels_ray=32)["scene"].img1, {}))
    for img, kw in cases:
        K = ctx.detect_sift(img, **kw)
        K2, D2 = ctx.detect_and_describe_sift(img, **kw)
        np.testing.assert_array_equal(K, K2)
        assert D2.shape == (len(K), 128)
        if len(K):
            D = ctx.describe_keypoints_sift_oct(img, K[:, :4].astype(np.float32), K[:, 5].astype(np.int32),
                                                n_octave_layers=kw.get("n_octave_layers", 3), sigma=kw.get("sigma", 1.6))
            np.testing.assert_array_equal(D, D2)
            assert (D2 != 0).any(1).all()
