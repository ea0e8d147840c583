"""The plain-C oracle against the committed golden vectors (tests/golden/*.npz, written by
tools/make_golden.py from cv2.undistortPoints / projectPoints / triangulatePoints / pyrDown /
BFMatcher and the cv2-based restatement of the reference).  CPU only."""
import os

import numpy as np
import pytest

from common import angle_deg, orc

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def prim():
    return np.load(os.path.join(GOLD, "primitives.npz"))


@pytest.fixture(scope="module")
def nrm():
    return np.load(os.path.join(GOLD, "normals.npz"))


def test_undistort_and_project(prim):
    np.testing.assert_allclose(orc.undistort_points(prim["K"], prim["dist"], prim["pts"]), prim["undistorted"], rtol=1e-13, atol=1e-14)
    np.testing.assert_allclose(orc.project_points(prim["K"], prim["dist"], prim["g12"], 1, prim["X"]), prim["proj1"], rtol=0, atol=1e-10)
    np.testing.assert_allclose(orc.project_points(prim["K"], prim["dist"], prim["g12"], 2, prim["X"]), prim["proj2"], rtol=0, atol=1e-10)


def test_triangulate(prim):
    xyz_all, mask, xyz = orc.triangulate(prim["K"], prim["dist"], prim["g12"], 1.5, 2.4, prim["kp1"], prim["kp2"])
    np.testing.assert_array_equal(mask, prim["mask"])
    assert 0 < mask.sum() < mask.size
    np.testing.assert_allclose(xyz_all, prim["xyz_all"], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(xyz, prim["xyz"], rtol=1e-10, atol=1e-12)


def test_pyrdown_bit_exact(prim):
    lv = orc.pyramid_levels(orc.build_pyramid(prim["img"], 3), prim["img"].shape[1], prim["img"].shape[0], 3)
    for k in (1, 2, 3):
        np.testing.assert_array_equal(lv[k], prim[f"pyr{k}"])


def test_matchers(prim):
    idx, dist = orc.knn2_f32(prim["q"], prim["t"])
    np.testing.assert_array_equal(idx, prim["idx_f"])
    np.testing.assert_array_equal(dist, prim["dist_f"])
    assert tuple(idx[0]) == (5, 11)                      # duplicates: lower train index first
    idx, dist = orc.knn2_hamming(prim["qb"], prim["tb"])
    np.testing.assert_array_equal(idx, prim["idx_b"])
    np.testing.assert_array_equal(dist, prim["dist_b"])
    q, t, d = orc.nndr_filter(prim["idx_f"], prim["dist_f"], 0.55)
    np.testing.assert_array_equal(q, prim["nndr_q"])
    np.testing.assert_array_equal(t, prim["nndr_t"])
    np.testing.assert_array_equal(d, prim["nndr_d"])


@pytest.mark.parametrize("mode,name", [(0, "fabs"), (1, "intabs"), (2, "off")])
def test_normal_optimiser(nrm, mode, name):
    r = orc.optimize_normals(nrm["K"], nrm["dist"], nrm["g12"], float(nrm["zmin"]), float(nrm["zmax"]), nrm["img1"], nrm["img2"],
                             2, nrm["xyz"], 16, 1e-10, penalty_mode=mode, threads=2)
    np.testing.assert_array_equal(r["status"], nrm[f"{name}_status"])
    np.testing.assert_array_equal(r["m"], nrm[f"{name}_m"])
    np.testing.assert_array_equal(r["nfev"], nrm[f"{name}_nfev"])          # same LM trajectory as the cv2 restatement
    np.testing.assert_array_equal(r["npenalty"], nrm[f"{name}_npenalty"])
    assert (angle_deg(r["normals"], nrm[f"{name}_normals"]) < 1e-5).all()
    np.testing.assert_allclose(r["cost"], nrm[f"{name}_cost"], rtol=1e-9)
    if name == "off":
        assert (angle_deg(r["normals"], nrm["gt_normal"]) < 0.5).all()    # and it finds the true facet normals


def test_cost_frames_patches(nrm):
    for lvl in range(3):
        c, m, st = orc.evaluate_cost(nrm["K"], nrm["dist"], nrm["g12"], float(nrm["zmin"]), float(nrm["zmax"]), nrm["img1"],
                                     nrm["img2"], 2, nrm["xyz"], nrm["cost_pt"], 16, lvl, 2)
        np.testing.assert_allclose(c, nrm[f"cost_l{lvl}"], rtol=1e-10)
    fr = orc.feature_frames(nrm["xyz"], nrm["off_normals"], nrm["gravity"])
    np.testing.assert_allclose(fr, nrm["frames"], rtol=0, atol=1e-15)
    p, ip = orc.extract_patches(nrm["K"], nrm["dist"], nrm["img1"], nrm["frames"], 0.05, 0.25)
    np.testing.assert_array_equal(p, nrm["patches"])
    np.testing.assert_allclose(ip, nrm["image_points"], rtol=0, atol=1e-10)


# ------------------------------------------------------------------ patch descriptors (K9 oracle)
def _desc_close(got, want, frac_exact=0.97):
    """Quantised SIFT values: +-1 on isolated entries (summation order of float sums), else equal."""
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape
    diff = np.abs(got.astype(np.int32) - want.astype(np.int32))
    assert diff.max(initial=0) <= 1, diff.max()
    assert (diff == 0).mean() >= frac_exact if diff.size else True


def test_sift_patch_restatement_against_cv2_golden_vectors():
    from oracle import sift_patch_np as sp
    g = np.load(os.path.join(GOLD, "sift_patches.npz"))
    for S in (128, 64, 40, 16, 8):
        d = sp.describe_patches_sift(g[f"p{S}"])
        _desc_close(d, g[f"d{S}"])
        # the reference's keypoint (size = S at the patch centre) only ever fills the 2 x 2 central cells,
        # plus orientation bin 1 of column 0 (o0 = -1 under flat addressing, see oracle/sift_patch_np.py)
        cells = g[f"d{S}"].reshape(-1, 4, 4, 8).copy()
        assert np.count_nonzero(cells[:, [0, 3], :, :]) == 0 and np.count_nonzero(cells[:, :, 3, :]) == 0
        assert np.count_nonzero(cells[:, 1:3, 0, 1]) > 0       # the step-edge patch exercises it
        cells[:, 1:3, 0, 1] = 0
        assert np.count_nonzero(cells[:, :, 0, :]) == 0
        assert np.count_nonzero(g[f"d{S}"][-2]) == 0          # constant patch -> zero descriptor
    pipe = np.load(os.path.join(GOLD, "normals.npz"))["patches"]
    _desc_close(sp.describe_patches_sift(pipe), g["d_pipeline"])


FAST_THRESHOLDS = (0, 1, 10, 20, 40, 100, 255)


def test_fast_restatement_against_cv2_golden_vectors():
    """oracle/fast_np.py against the committed outputs of cv2.FastFeatureDetector (integer work: exact,
    order included)."""
    from oracle import fast_np as fo
    g = np.load(os.path.join(GOLD, "fast_keypoints.npz"))
    total = 0
    for name in ("noise", "blur", "frame", "tiny7", "tiny6"):
        img = g[f"img_{name}"]
        for t in FAST_THRESHOLDS:
            for nm in (0, 1):
                xy, r = fo.detect_fast(img, t, bool(nm))
                assert xy.dtype == np.float32 and r.dtype == np.float32
                np.testing.assert_array_equal(xy, g[f"xy_{name}_{t}_{nm}"].astype(np.float32).reshape(-1, 2))
                np.testing.assert_array_equal(r, g[f"r_{name}_{t}_{nm}"].astype(np.float32))
                total += len(xy)
    assert total > 50000


def test_sift_keypoint_restatement_against_cv2_golden_vectors():
    """oracle/sift_kp_np.py (K11: descriptor_extractor_->compute, ExtractorType SIFT, octave-0 keypoints) against
    the committed outputs of cv2.SIFT_create().compute on FAST keypoints and on keypoints with real sizes /
    angles / border positions."""
    from oracle import sift_kp_np as sk
    g = np.load(os.path.join(GOLD, "sift_keypoints.npz"))
    imgs = np.load(os.path.join(GOLD, "fast_keypoints.npz"))
    seen = 0
    for key in g.files:
        if not key.startswith("k_"):
            continue
        name = key.split("_")[1]
        d = sk.describe_keypoints_sift(imgs[f"img_{name}"], g[key])
        _desc_close(d, g["d" + key[1:]], frac_exact=0.999)
        seen += len(d)
    assert seen > 1000
    # FAST keypoints carry angle -1 (ori = 361): the flat-addressing spill fills bins a wrapped ori would not
    k = g["k_frame_fast"].copy()
    k[:, 3] = 359.0                 # 360 - 359 = 1 degree: the same rotation without the out-of-range ori
    alt = sk.describe_keypoints_sift(imgs["img_frame"], k)
    assert (alt != g["d_frame_fast"]).any()


def test_brisk_restatement_against_cv2_golden_vectors():
    """oracle/brisk_np.py (K12: descriptor_extractor_->compute, ExtractorType BRISK) against the committed outputs of
    cv2.BRISK_create(25, 0).compute: surviving keypoints, all 512 bits and the angles identical (scale indices 0 ... 30,
    sub-pixel positions)."""
    from oracle import brisk_np as bn
    g = np.load(os.path.join(GOLD, "brisk_keypoints.npz"))
    imgs = np.load(os.path.join(GOLD, "fast_keypoints.npz"))
    sh, lg = bn.pairs()
    assert len(sh) == 512 and len(lg) == 870 and bn.size_of_scale(0) == 13
    seen = 0
    for key in g.files:
        if not key.startswith("k_"):
            continue
        name, tag = key.split("_")[1:3]
        kept, ang, d = bn.describe_keypoints_brisk(imgs[f"img_{name}"], g[key])
        np.testing.assert_array_equal(kept, g[f"kept_{name}_{tag}"])
        np.testing.assert_array_equal(d, g[f"d_{name}_{tag}"])
        np.testing.assert_array_equal(ang, g[f"a_{name}_{tag}"])
        assert len(kept) < len(g[key])               # the border rule removed some
        seen += len(kept)
    assert seen > 800


def test_orb_restatement_against_cv2_golden_vectors():
    """oracle/orb_np.py (K13: descriptor_extractor_->compute, ExtractorType ORB) against the committed outputs of
    cv2.ORB_create().compute: survivors of the border rule and all 256 bits, for FAST keypoints and for sub-pixel
    keypoints with arbitrary angles.  The pattern itself is recovered from cv2 (tools/recover_orb_pattern.py)."""
    from oracle import orb_np as on
    g = np.load(os.path.join(GOLD, "orb_keypoints.npz"))
    imgs = np.load(os.path.join(GOLD, "fast_keypoints.npz"))
    assert on.PATTERN.shape == (256, 4) and np.abs(on.PATTERN).max() <= 15
    assert on.PATTERN[0].tolist() == [-3, 8, 5, 9]        # (y, x) of both points of bit 0: Rublee et al.'s first pair (8,-3)-(9,5)
    seen = 0
    for key in g.files:
        if not key.startswith("k_"):
            continue
        name, tag = key.split("_")[1:3]
        kept, d = on.describe_keypoints_orb(imgs[f"img_{name}"], g[key])
        np.testing.assert_array_equal(kept, g[f"kept_{name}_{tag}"])
        np.testing.assert_array_equal(d, g[f"d_{name}_{tag}"])
        seen += len(kept)
    assert seen > 600
    # extractDescriptorsFromPatches with ExtractorType ORB: one keypoint per patch
    for S in (128, 64, 63):
        for patch, want in zip(g[f"p{S}"], g[f"dp{S}"]):
            kept, d = on.describe_keypoints_orb(patch, np.array([[S // 2, S // 2, S, -1]], np.float32))
            assert len(kept) == 1 and np.unpackbits(d[0] ^ want).sum() <= 2
