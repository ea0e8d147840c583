"""The reference's own clients on top of the fm3d adapters (north star: "main.cpp and Mosaic run unchanged on top of it").

`__graft_entry__.build_reference_clients()` compiles /root/reference/main.cpp -- byte for byte unchanged, from where it
lies -- and MOSAIC (mosaic.h + mosaic.cpp without lines 81-89) against 3dfeaturematcher_b200/host/compat/ and
the four adapter headers, into tests/_build/ref_main and tests/_build/ref_mosaic.  The reference tree exists in the
build container only, so the CPU test compiles and the GPU test runs the prebuilt binaries:

* `ref_main -s settings.yml` on a synthetic 640x480 pair with DetectorType FAST + ExtractorType SIFT must leave the three
  artefacts main.cpp writes -- matches.pgm (main.cpp:139-141), patch_<k>.pgm (singlecameratriangulator.cpp:799-802) and
  projectedPatches.pgm (main.cpp:190-194) -- equal to what the oracle chain produces from the same frames;
* `ref_mosaic` runs MOSAIC's constructor (steps 1-7) and must leave the same patch files."""
import hashlib
import os
import subprocess

import cv2
import numpy as np
import pytest

from common import angle_deg, orc, stereo_case
from test_host_pipeline import _write_inputs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "tests", "_build")
REF_MAIN, REF_MOSAIC = os.path.join(BUILD, "ref_main"), os.path.join(BUILD, "ref_mosaic")

FAST_SIFT = """FeatureOptions:
   DetectorType: FAST
   DetectorMode: STATIC
   FastDetector:
      Threshold: 25
      NonMaxSuppression: 1
   SiftDetector:
      NumFeatures: 0
      NumOctaveLayers: 3
      ContrastThreshold: 0.04
      EdgeThreshold: 10
      Sigma: 1.6
   ExtractorType: SIFT
"""


def test_the_unchanged_main_cpp_and_mosaic_compile_against_the_adapters():
    """No GPU needed.  Where the reference tree is present (the build container) both clients are compiled from it, the
    recorded digest is that of the file as it lies there, and the binary prints the reference's own usage text."""
    import importlib
    ge = importlib.import_module("__graft_entry__")
    if not os.path.exists(os.path.join(ge.REFERENCE, "main.cpp")):
        pytest.skip("reference tree not present on this box: the prebuilt binaries are used")
    lib = os.path.join(ROOT, "3dfeaturematcher_b200", "libfm3d.so")
    if not os.path.exists(lib):
        importlib.import_module("3dfeaturematcher_b200.build").build()
    assert ge.build_reference_clients()
    assert os.access(REF_MAIN, os.X_OK) and os.access(REF_MOSAIC, os.X_OK)
    with open(os.path.join(ge.REFERENCE, "main.cpp"), "rb") as f:
        digest = hashlib.sha256(f.read()).hexdigest()
    assert open(os.path.join(BUILD, "ref_main.sha256")).read().split()[0] == digest
    assert not os.path.exists(os.path.join(BUILD, "ref")), "no reference source may stay in the tree"
    p = subprocess.run([REF_MAIN], capture_output=True, text=True)
    assert p.returncode == 255 and "Usage: 3dfeaturematcher -s <settings.yml>" in p.stdout      # main.cpp:45-49: help(); exit(-1)
    p = subprocess.run([REF_MAIN, "-s", "/nonexistent.yml"], capture_output=True, text=True)
    assert p.returncode == 255 and "Could not open settings file" in p.stderr                   # main.cpp:67-71


def _read_pnm(path):
    img = cv2.imread(path, cv2.IMREAD_UNCHANGED)
    assert img is not None, path
    return img


def _oracle_chain(ctx, case, r, pyramids, eps_m, cmpp, penalty_mode):
    """What main.cpp computes, by the oracles: FAST keypoints (oracle, bit-identical to K10), their SIFT descriptors (K11
    through the C-ABI: pinned to the oracle within +-1 elsewhere; taken from the device here so that the match list is
    decided by the same values), then matching, triangulation, normals, frames and patches by the CPU oracle."""
    from oracle import fast_np as fo
    cam = case["scene"].cam
    kps, descs = [], []
    for img in (case["scene"].img1, case["scene"].img2):
        xy, resp = fo.detect_fast(img, 25, True)
        k4 = np.concatenate([xy, np.full((len(xy), 1), 7, np.float32), np.full((len(xy), 1), -1, np.float32)], 1)
        kps.append(xy)
        descs.append(ctx.describe_keypoints_sift(img, k4))
    o_idx, o_dist = orc.knn2_f32(descs[0], descs[1])
    oq, ot, od = orc.nndr_filter(o_idx, o_dist, 0.55)
    o_all, o_mask, o_xyz = orc.triangulate(cam.K, cam.dist, cam.g12, cam.z_min, cam.z_max, kps[0], kps[1], oq, ot)
    o = orc.optimize_normals(cam.K, cam.dist, cam.g12, cam.z_min, cam.z_max, case["scene"].img1, case["scene"].img2,
                             pyramids, o_xyz, r, 1e-10, penalty_mode=penalty_mode, threads=os.cpu_count() or 8)
    ok = o["status"] == 0
    import importlib
    synth = importlib.import_module("3dfeaturematcher_b200.synth")
    gravity = np.linalg.inv(cv2.Rodrigues(np.asarray(synth.SETTINGS_RODRIGUES_IC, float).reshape(3, 1))[0]) @ np.array([0.0, 0.0, -1.0])
    # FAST corners lie anywhere in the frame: for a few of them an LM iterate leaves the image or the bounding box and the
    # feature is dropped -- which iterate does is trajectory-dependent, so the device run (analytic Jacobian)
    # and the oracle (forward differences) may drop different marginal features.  The device statuses of the same points are
    # taken through the C-ABI (the kernel is deterministic); they must agree with the oracle on >= 99 % of the features, and
    # the artefacts are compared on the features both keep.
    ctx.set_camera(cam.K, cam.dist, cam.z_min, cam.z_max)
    ctx.set_g12(cam.g12)
    ctx.set_images(case["scene"].img1, case["scene"].img2, pyramids)
    dev = ctx.optimize_normals(o_xyz, r, 1e-10, penalty_mode)
    dev_ok = dev["status"] == 0
    assert (dev["status"] == o["status"]).mean() >= 0.99, (dev["status"] != o["status"]).sum()
    frames = orc.feature_frames(o_xyz[ok], o["normals"][ok], gravity)
    patches, points = orc.extract_patches(cam.K, cam.dist, case["scene"].img1, frames, eps_m, cmpp, want_points=True)
    both = ok & dev_ok
    return dict(kps=kps, q=oq, t=ot, mask=o_mask.astype(bool), xyz=o_xyz, ok=ok, dev_ok=dev_ok, patches=patches, points=points,
                in_oracle=np.cumsum(ok)[both] - 1, in_device=np.cumsum(dev_ok)[both] - 1)


@pytest.mark.gpu
def test_unchanged_main_cpp_runs_and_writes_its_three_artefacts(ctx, tmp_path):
    from oracle import draw_cv
    if not os.access(REF_MAIN, os.X_OK):
        pytest.fail("tests/_build/ref_main missing: run __graft_entry__.build() where /root/reference exists")
    r, pyramids, eps_m, cmpp = 32, 2, 0.05, 0.25
    case = stereo_case(640, 480, 60, 1000, r)
    tmp = str(tmp_path)
    _write_inputs(tmp, case, r, pyramids, eps_m, cmpp, feature_options=FAST_SIFT)
    env = dict(os.environ, FM3D_PENALTY="1")
    p = subprocess.run([REF_MAIN, "-s", os.path.join(tmp, "settings.yml")], capture_output=True, text=True, env=env, cwd=tmp, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    assert p.stdout.startswith("Hello!")                                            # main.cpp:38
    o = _oracle_chain(ctx, case, r, pyramids, eps_m, cmpp, 1)
    assert len(o["q"]) >= 50 and o["ok"].sum() >= 20

    # matches.pgm: binary PPM (three channels), both frames side by side, circles + line per depth-gated match
    window = _read_pnm(os.path.join(tmp, "matches.pgm"))
    o_window, colours = draw_cv.draw_matches(case["scene"].img1, case["scene"].img2, o["kps"][0], o["kps"][1],
                                             np.stack([o["q"], o["t"]], 1), o["mask"])
    assert open(os.path.join(tmp, "matches.pgm"), "rb").read(2) == b"P6"
    np.testing.assert_array_equal(window, o_window)

    # patch_<k>.pgm: one per feature that survived the normal search, in order
    n_dev = int(o["dev_ok"].sum())
    files = sorted(f for f in os.listdir(tmp) if f.startswith("patch_"))
    assert len(files) == n_dev and not os.path.exists(os.path.join(tmp, f"patch_{n_dev}.pgm"))
    S = orc.patch_size(eps_m, cmpp)
    got = np.stack([_read_pnm(os.path.join(tmp, f"patch_{k}.pgm")) for k in range(n_dev)])
    assert got.shape == (n_dev, S, S)
    # the device normals differ from the oracle's (0.002 deg on a facet, more for FAST corners whose disc straddles two facets:
    # both searches are valid, their end states are not pinned to each other there): sampling positions move by a fraction of
    # a pixel and the truncated gray levels by a few units on a small share of the pixels
    diff = np.abs(got[o["in_device"]].astype(int) - o["patches"][o["in_oracle"]].astype(int))
    assert (diff <= 2).mean() > 0.99 and (diff == 0).mean() > 0.9 and np.median(diff.reshape(len(diff), -1).mean(1)) < 0.1, \
        ((diff <= 2).mean(), (diff == 0).mean())

    # projectedPatches.pgm: image points of every patch painted in colours[i] (main.cpp:190-194; colours are per INLIER, D5)
    proj = _read_pnm(os.path.join(tmp, "projectedPatches.pgm"))
    n_ok = int(o["ok"].sum())
    pts = o["points"].reshape(n_ok, S * S, 2)
    if n_dev == n_ok and (o["dev_ok"] == o["ok"]).all():
        o_proj = draw_cv.draw_back_projected_points(case["scene"].img1, pts, np.asarray(colours)[:n_ok, :3])
        assert proj.shape == o_proj.shape and (proj == o_proj).all(axis=2).mean() > 0.995
    else:       # survivor lists differ by a marginal feature: patch i of the device run takes colours[i] of ITS numbering (D5)
        o_proj = draw_cv.draw_back_projected_points(case["scene"].img1, pts[o["in_oracle"]], np.asarray(colours)[o["in_device"], :3])
        painted = (o_proj != cv2.cvtColor(case["scene"].img1, cv2.COLOR_GRAY2BGR)).any(axis=2)
        assert proj.shape == o_proj.shape and (proj[painted] == o_proj[painted]).all(axis=1).mean() > 0.98


@pytest.mark.gpu
def test_mosaic_constructor_runs_the_pipeline_on_the_adapters(ctx, tmp_path):
    if not os.access(REF_MOSAIC, os.X_OK):
        pytest.fail("tests/_build/ref_mosaic missing: run __graft_entry__.build() where /root/reference exists")
    r, pyramids, eps_m, cmpp = 32, 2, 0.05, 0.25
    case = stereo_case(640, 480, 60, 1000, r)
    tmp = str(tmp_path)
    _write_inputs(tmp, case, r, pyramids, eps_m, cmpp, feature_options=FAST_SIFT)
    env = dict(os.environ, FM3D_PENALTY="1")
    p = subprocess.run([REF_MOSAIC, "-s", os.path.join(tmp, "settings.yml")], capture_output=True, text=True, env=env, cwd=tmp, timeout=600)
    assert p.returncode == 0 and "MOSAIC constructed" in p.stdout, p.stdout[-2000:] + p.stderr[-2000:]
    o = _oracle_chain(ctx, case, r, pyramids, eps_m, cmpp, 1)
    n_dev = int(o["dev_ok"].sum())
    files = [f for f in os.listdir(tmp) if f.startswith("patch_")]
    assert len(files) == n_dev >= 20
    got = np.stack([_read_pnm(os.path.join(tmp, f"patch_{k}.pgm")) for k in range(n_dev)])
    diff = np.abs(got[o["in_device"]].astype(int) - o["patches"][o["in_oracle"]].astype(int))
    assert (diff <= 2).mean() > 0.99 and (diff == 0).mean() > 0.9 and np.median(diff.reshape(len(diff), -1).mean(1)) < 0.1
