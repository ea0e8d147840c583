"""Sanity of the synthetic generator (CPU): geometry is self-consistent and deterministic."""
import numpy as np

from common import orc, stereo_case, synth


def test_scene_is_deterministic_and_consistent():
    a = synth.make_stereo_case(320, 240, 10, 5, pixels_ray=16)
    b = synth.make_stereo_case(320, 240, 10, 5, pixels_ray=16)
    np.testing.assert_array_equal(a["scene"].img1, b["scene"].img1)
    np.testing.assert_array_equal(a["kp1"], b["kp1"])
    cam = a["scene"].cam
    # ground-truth points project onto the keypoints of both views
    np.testing.assert_allclose(synth.project(cam, a["X"], 1), a["kp1"], atol=1e-3)
    np.testing.assert_allclose(synth.project(cam, a["X"], 2), a["kp2_true"], atol=1e-3)
    # and the oracle triangulates them back (float32 keypoints: ~1e-4 m)
    xyz_all, mask, _ = orc.triangulate(cam.K, cam.dist, cam.g12, cam.z_min, cam.z_max, a["kp1"], a["kp2_true"])
    assert mask.all() and np.abs(xyz_all - a["X"]).max() < 5e-3
    assert a["scene"].img1.std() > 40


def test_descriptor_sets_have_known_matches():
    q, t, gt = synth.make_float_descriptors(200, 50, 1)
    assert q.dtype == np.float32 and (q == np.floor(q)).all() and q.min() >= 0 and q.max() <= 255
    idx, _ = orc.knn2_f32(q, t)
    inl = gt >= 0
    assert (idx[inl, 0] == gt[inl]).mean() > 0.99
    qb, tb, gtb = synth.make_binary_descriptors(200, 50, 1)
    idx, _ = orc.knn2_hamming(qb, tb)
    assert (idx[gtb >= 0, 0] == gtb[gtb >= 0]).mean() > 0.99
