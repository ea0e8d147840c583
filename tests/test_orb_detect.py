"""K15 -- cv::ORB's own detector + descriptors (DetectorType ORB + ExtractorType ORB, DescriptorsMatcher/descriptorsmatcher.cpp:273-279,
:336-342, called at :110-115): the oracle restatement against the committed outputs of cv2.ORB_create(...).detectAndCompute, and
the GPU path (fm3d_detect_orb) against the same golden vectors.

Integer work up to the Harris measure and the angle (float formulas on integer sums): the keypoint SET (level, level position)
must be cv2's, angles within 1e-3 degree, responses within 1e-6 relative, sizes and frame coordinates exact; descriptor rows
identical for >= 99.5 % of the keypoints, <= 2 bits otherwise (a blurred value within float rounding of a half-integer, as
for K13).  OpenCV's output ORDER is what std::nth_element leaves: both sides are sorted by (octave, y, x)."""
import importlib
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PARAMS = {"default": dict(), "knobs": dict(nfeatures=1500, scale_factor=1.3, nlevels=5)}


def compare(want_k, want_oct, want_d, K, D, label):
    assert len(K) == len(want_k), (label, len(K), len(want_k))
    np.testing.assert_array_equal(K[:, 5].astype(np.int32), want_oct)
    np.testing.assert_array_equal(K[:, :3].astype(np.float32), want_k[:, :3])              # x, y, size: products of integers and s^l
    da = np.abs(K[:, 3] - want_k[:, 3])
    assert np.minimum(da, 360 - da).max() < 1e-3, label
    np.testing.assert_allclose(K[:, 4], want_k[:, 4], rtol=1e-6, atol=1e-12)
    bits = np.unpackbits(D ^ want_d, axis=1).sum(1)
    print(f"ORB detect {label}: {len(K)} keypoints, rows identical {np.mean(bits == 0):.4f}, max differing bits {bits.max()}")
    assert np.mean(bits == 0) >= 0.995 and bits.max() <= 2, label


def test_orb_detector_restatement_against_cv2_golden_vectors():
    from oracle import orb_detect_np as od
    g = np.load(os.path.join(GOLD, "orb_detect.npz"))
    for name in ("blobs", "frame"):
        for pname, kw in PARAMS.items():
            K, D = od.detect_and_describe_orb(g[f"img_{name}"], **kw)
            compare(g[f"kp_{name}_{pname}"], g[f"oct_{name}_{pname}"], g[f"desc_{name}_{pname}"], K[:, :6], D, f"oracle {name}/{pname}")


def test_resize_linear_exact_restatement():
    cv2 = pytest.importorskip("cv2")
    from oracle import orb_detect_np as od
    rng = np.random.default_rng(2)
    for (h, w), s in (((120, 160), 1.2), ((97, 131), 1.3), ((33, 50), 1.2)):
        src = rng.integers(0, 256, (h, w)).astype(np.uint8)
        dw, dh = int(np.rint(w / s)), int(np.rint(h / s))
        np.testing.assert_array_equal(od.resize_linear_exact(src, dw, dh), cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LINEAR_EXACT))


@pytest.fixture(scope="module")
def ctx():
    api = importlib.import_module("3dfeaturematcher_b200.api")
    c = api.Context(0)
    yield c
    c.close()


@pytest.mark.gpu
def test_gpu_orb_detector_against_cv2_golden_vectors(ctx):
    g = np.load(os.path.join(GOLD, "orb_detect.npz"))
    for name in ("blobs", "frame"):
        for pname, kw in PARAMS.items():
            K, D = ctx.detect_orb(g[f"img_{name}"], **kw)
            K2, D2 = ctx.detect_orb(g[f"img_{name}"], **kw)
            np.testing.assert_array_equal(K, K2)
            np.testing.assert_array_equal(D, D2)
            compare(g[f"kp_{name}_{pname}"], g[f"oct_{name}_{pname}"], g[f"desc_{name}_{pname}"], K, D, f"gpu {name}/{pname}")


@pytest.mark.gpu
def test_gpu_orb_detector_against_cv2_on_a_720p_frame(ctx):
    cv2 = pytest.importorskip("cv2")
    synth = importlib.import_module("3dfeaturematcher_b200.synth")
    img = synth.make_stereo_case(1280, 720, 50, 1001, pixels_ray=32)["scene"].img1
    kps, desc = cv2.ORB_create(nfeatures=5000).detectAndCompute(img, None)
    C = np.array([[k.pt[0], k.pt[1], k.size, k.angle, k.response, k.octave] for k in kps], np.float64)
    order = np.lexsort((C[:, 0], C[:, 1], C[:, 5]))
    K, D = ctx.detect_orb(img, nfeatures=5000)
    compare(C[order, :5].astype(np.float32), C[order, 5].astype(np.int32), desc[order], K, D, "gpu 720p / 5000")


@pytest.mark.gpu
def test_gpu_orb_detector_edge_cases(ctx):
    api = importlib.import_module("3dfeaturematcher_b200.api")
    K, D = ctx.detect_orb(np.full((100, 120), 93, np.uint8))
    assert len(K) == 0 and D.shape == (0, 32)
    K, D = ctx.detect_orb(np.random.default_rng(1).integers(0, 256, (40, 60)).astype(np.uint8))      # smaller than the border
    assert len(K) == 0
    with pytest.raises(api.Fm3dError):
        ctx.detect_orb(np.zeros((50, 50), np.uint8), scale_factor=1.0)
