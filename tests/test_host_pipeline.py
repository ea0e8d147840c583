"""The reference's own call sequence (main.cpp:91-187) on top of the four class adapters in
3dfeaturematcher_b200/host/ (same class names and signatures as the reference, implemented over the
C-ABI), driven through a settings.yml with the reference's layout, compared with the oracle."""
import os
import struct
import subprocess

import numpy as np
import pytest

from common import angle_deg, cam_tuple, orc, stereo_case, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "3dfeaturematcher_b200", "host")
EXE = os.path.join(ROOT, "tests", "_build", "pipeline_main")


def build_pipeline_main():
    srcs = [os.path.join(ROOT, "tests", "cpp", "pipeline_main.cpp"), os.path.join(HOST, "fm3d_host.cpp")]
    deps = srcs + [os.path.join(HOST, "fm3d_cv.h"), os.path.join(ROOT, "include", "fm3d.h")]
    lib = os.path.join(ROOT, "3dfeaturematcher_b200", "libfm3d.so")
    if not os.path.exists(lib):
        __import__("importlib").import_module("3dfeaturematcher_b200.build").build()
    if not os.path.exists(EXE) or os.path.getmtime(EXE) < max(os.path.getmtime(d) for d in deps + [lib]):
        os.makedirs(os.path.dirname(EXE), exist_ok=True)
        gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.check_call([gxx, "-O2", "-std=c++17", "-pthread", "-I" + os.path.join(HOST, "include"), "-I" + HOST,
                               "-I" + os.path.join(ROOT, "include")] + srcs +
                              ["-L" + os.path.dirname(lib), "-lfm3d", "-Wl,-rpath," + os.path.dirname(lib), "-o", EXE])
    return EXE


def test_adapters_compile_and_link_against_the_c_abi():
    """No GPU needed: the reference-shaped classes build against include/fm3d.h and link with libfm3d.so."""
    exe = build_pipeline_main()
    assert os.access(exe, os.X_OK)
    p = subprocess.run([exe], capture_output=True, text=True)
    assert p.returncode == 1 and "usage" in p.stderr


def _rodrigues(r):
    import cv2
    return cv2.Rodrigues(np.asarray(r, np.float64).reshape(3, 1))[0]


def _g(rod, t):
    g = np.eye(4)
    g[:3, :3] = _rodrigues(rod)
    g[:3, 3] = t
    return g


_DEFAULT_FEATURE_OPTIONS = """FeatureOptions:
   DetectorType: SIFT
   DetectorMode: STATIC
   ExtractorType: SIFT
"""


def _write_inputs(tmp, case, r, pyramids, eps_m, cmpp, feature_options=None):
    import cv2
    cam = case["scene"].cam
    # poses such that g12 = g_IC^-1 g2^-1 g1 g_IC (singlecameratriangulator.cpp:140) is the scene's g12
    rIC, tIC = np.asarray(synth.SETTINGS_RODRIGUES_IC, float), np.asarray(synth.SETTINGS_TRANSLATION_IC, float)
    pos1 = np.array([5.301099, 8.031408, 1.977258, 0.153433, 0.149941, -2.658648])   # build/settings.yml:11
    gIC, g1 = _g(rIC, tIC), _g(pos1[3:], pos1[:3])
    g2 = g1 @ gIC @ np.linalg.inv(cam.g12) @ np.linalg.inv(gIC)
    rod2 = cv2.Rodrigues(g2[:3, :3])[0].ravel()
    pos2 = np.concatenate([g2[:3, 3], rod2])
    for name, img in (("img1.pgm", case["scene"].img1), ("img2.pgm", case["scene"].img2)):
        with open(os.path.join(tmp, name), "wb") as f:
            f.write(b"P5\n%d %d\n255\n" % (img.shape[1], img.shape[0]))
            f.write(np.ascontiguousarray(img).tobytes())
    K, d = cam.K, cam.dist
    fmt = lambda v: "[" + ", ".join(repr(float(x)) for x in v) + "]"
    yml = f"""%YAML:1.0
IMAGES:
   img1: {os.path.join(tmp, 'img1.pgm')}
   img2: {os.path.join(tmp, 'img2.pgm')}
   pos1: {fmt(pos1)}
   pos2: {fmt(pos2)}

NNDR:
   epsilon: 0.55

Neighborhoods:
   #Part for normal optimization: take pixels in the image
   epsilonLMMIN: 1e-10
   pixelsRay: {r}
   pyramids: {pyramids}
   method: square
   cmPerPixel: {cmpp}
   epsilon: {eps_m}

{feature_options or _DEFAULT_FEATURE_OPTIONS}

CameraSettings:
   rodriguesIC: {fmt(rIC)}
   translationIC: {fmt(tIC)}
   Fx: {float(K[0, 0])!r}
   Fy: {float(K[1, 1])!r}
   Cx: {float(K[0, 2])!r}
   Cy: {float(K[1, 2])!r}
   p1: {float(d[2])!r}
   p2: {float(d[3])!r}
   k0: {float(d[0])!r}
   k1: {float(d[1])!r}
   k2: {float(d[4])!r}
   zThresholdMin: {float(cam.z_min)!r}
   zThresholdMax: {float(cam.z_max)!r}
"""
    with open(os.path.join(tmp, "settings.yml"), "w") as f:
        f.write(yml)
    n1, n2, dim = case["desc1"].shape[0], case["desc2"].shape[0], case["desc1"].shape[1]
    with open(os.path.join(tmp, "features.bin"), "wb") as f:
        f.write(struct.pack("iii", n1, n2, dim))
        f.write(np.ascontiguousarray(case["kp1"], np.float32).tobytes())
        f.write(np.ascontiguousarray(case["kp2"], np.float32).tobytes())
        f.write(np.ascontiguousarray(case["desc1"], np.float32).tobytes())
        f.write(np.ascontiguousarray(case["desc2"], np.float32).tobytes())


def _read_result(path):
    b = open(path, "rb").read()
    nm, npts, nn, S = struct.unpack_from("iiii", b, 0)
    o = 16
    m = np.frombuffer(b, np.dtype([("q", "i4"), ("t", "i4"), ("d", "f4")]), nm, o); o += 12 * nm
    mask = np.frombuffer(b, np.uint8, nm, o); o += nm
    g12 = np.frombuffer(b, np.float64, 16, o).reshape(4, 4); o += 128
    pts = np.frombuffer(b, np.float64, 3 * npts, o).reshape(-1, 3); o += 24 * npts
    status = np.frombuffer(b, np.int32, npts, o); o += 4 * npts
    kept = np.frombuffer(b, np.float64, 3 * nn, o).reshape(-1, 3); o += 24 * nn
    normals = np.frombuffer(b, np.float64, 3 * nn, o).reshape(-1, 3); o += 24 * nn
    frames = np.frombuffer(b, np.float64, 16 * nn, o).reshape(-1, 4, 4); o += 128 * nn
    patches = np.frombuffer(b, np.uint8, nn * S * S, o).reshape(nn, S, S); o += nn * S * S
    last_nb = np.frombuffer(b, np.float64, 3 * nn, o).reshape(-1, 3); o += 24 * nn
    gravity = np.frombuffer(b, np.float64, 3, o); o += 24
    hm, rc1, rc2, rc3 = struct.unpack_from("iiii", b, o); o += 16
    hcost = struct.unpack_from("d", b, o)[0]; o += 8
    cs = struct.unpack_from("i", b, o)[0]; o += 4
    circ = None
    if cs > 0:
        cn = np.frombuffer(b, np.float64, 3, o); o += 24
        cpts = np.frombuffer(b, np.float64, 3 * cs, o).reshape(-1, 3); o += 24 * cs
        circ = dict(normal=cn, pts=cpts)
    dr, dc = struct.unpack_from("ii", b, o); o += 8
    if dc < 0:
        desc = np.frombuffer(b, np.uint8, dr * -dc, o).reshape(dr, -dc); o += dr * -dc
    else:
        desc = np.frombuffer(b, np.float32, dr * dc, o).reshape(dr, dc); o += 4 * dr * dc
    detected = []
    while o < len(b) and len(detected) < 2:          # "-" mode: the features compareWithNNDR produced itself
        n, cols, es = struct.unpack_from("iii", b, o); o += 12
        k = np.frombuffer(b, np.float32, 5 * n, o).reshape(n, 5); o += 20 * n
        d = np.frombuffer(b, np.float32 if es == 4 else np.uint8, n * cols, o).reshape(n, cols); o += es * n * cols
        detected.append((k, d))
    assert o == len(b)
    return dict(detected=detected, patch_descriptors=desc, circular=circ, helpers=dict(m=hm, rc=(rc1, rc2, rc3), cost=hcost), matches=m, mask=mask, g12=g12, pts=pts, status=status, kept=kept, normals=normals, frames=frames,
                patches=patches, last_nb=last_nb, gravity=gravity, S=S)


@pytest.mark.gpu
def test_main_cpp_call_sequence_on_the_adapters(tmp_path):
    exe = build_pipeline_main()
    r, pyramids, eps_m, cmpp = 32, 2, 0.05, 0.25
    case = stereo_case(640, 480, 60, 1000, r)
    cam = case["scene"].cam
    tmp = str(tmp_path)
    _write_inputs(tmp, case, r, pyramids, eps_m, cmpp)
    env = dict(os.environ, FM3D_PENALTY="2", FM3D_NO_PATCH_FILES="1")
    with open(os.path.join(tmp, "circular.yml"), "w") as f:
        f.write("%YAML:1.0\nNeighborhoods:\n   method: circular\n   epsilon: 0.16\n   thetas: 15\n   rays: 5\n")
    p = subprocess.run([exe, "-s", os.path.join(tmp, "settings.yml"), os.path.join(tmp, "features.bin"),
                        os.path.join(tmp, "result.bin"), os.path.join(tmp, "circular.yml")],
                       capture_output=True, text=True, env=env, cwd=tmp, timeout=300)
    assert p.returncode == 0, p.stdout + p.stderr
    res = _read_result(os.path.join(tmp, "result.bin"))

    # matching + NNDR (descriptorsmatcher.cpp:117-129): exact
    o_idx, o_dist = orc.knn2_f32(case["desc1"], case["desc2"])
    oq, ot, od = orc.nndr_filter(o_idx, o_dist, 0.55)
    np.testing.assert_array_equal(res["matches"]["q"], oq)
    np.testing.assert_array_equal(res["matches"]["t"], ot)
    np.testing.assert_array_equal(res["matches"]["d"], od)
    # setg12 from the poses (singlecameratriangulator.cpp:123-143)
    np.testing.assert_allclose(res["g12"], cam.g12, rtol=0, atol=1e-9)
    # triangulation with the composed g12
    o_all, o_mask, o_xyz = orc.triangulate(cam.K, cam.dist, res["g12"], cam.z_min, cam.z_max, case["kp1"], case["kp2"], oq, ot)
    np.testing.assert_array_equal(res["mask"].astype(bool), o_mask.astype(bool))
    np.testing.assert_allclose(res["pts"], o_xyz, rtol=1e-9, atol=1e-12)
    # normal optimisation: statuses identical, failed features erased from points3D in place
    o = orc.optimize_normals(cam.K, cam.dist, res["g12"], cam.z_min, cam.z_max, case["scene"].img1, case["scene"].img2,
                             pyramids, o_xyz, r, 1e-10, penalty_mode=2, threads=8)
    np.testing.assert_array_equal(res["status"], o["status"])
    ok = o["status"] == 0
    assert ok.sum() >= 20
    np.testing.assert_allclose(res["kept"], o_xyz[ok], rtol=1e-9, atol=1e-12)
    assert (angle_deg(res["normals"], o["normals"][ok]) <= 0.5).all()
    # frames (normaloptimizer.cpp:454-505) with the adapter's gravity = R_IC^-1 (0,0,-1) (:160-176)
    g_expected = np.linalg.inv(_rodrigues(synth.SETTINGS_RODRIGUES_IC)) @ np.array([0.0, 0.0, -1.0])
    np.testing.assert_allclose(res["gravity"], g_expected, atol=1e-12)
    np.testing.assert_allclose(res["frames"], orc.feature_frames(res["kept"], res["normals"], res["gravity"]), rtol=0, atol=1e-12)
    # patches: transposed write, truncation, +-1 gray level at float-cast boundaries
    assert res["S"] == orc.patch_size(eps_m, cmpp)
    o_patches, _ = orc.extract_patches(cam.K, cam.dist, case["scene"].img1, res["frames"], eps_m, cmpp, want_points=False)
    diff = res["patches"].astype(int) - o_patches.astype(int)
    assert (np.abs(diff) <= 1).all() and (diff != 0).mean() < 1e-3
    # the four public per-evaluation helpers chained by hand == one evaluateNormal of the oracle
    h = res["helpers"]
    assert h["rc"] == (0, 0, 0)
    pt = np.array([[np.arctan2(res["normals"][0, 1], res["normals"][0, 0]),
                    np.arctan2(res["normals"][0, 2], np.hypot(res["normals"][0, 0], res["normals"][0, 1]))]])
    oc, om, ost = orc.evaluate_cost(cam.K, cam.dist, res["g12"], cam.z_min, cam.z_max, case["scene"].img1, case["scene"].img2,
                                    pyramids, res["kept"][:1], pt, r, 0, 2)
    assert h["m"] == om[0] and ost[0] == 0
    np.testing.assert_allclose(h["cost"], oc[0], rtol=1e-6)
    # circular neighbourhood of the first feature with the initial-guess normal (neighborhoodsgenerator.cpp:238-277)
    from oracle import oracle_cv as ocv
    o_c, o_n = ocv.circular_neighborhood(res["kept"][0], np.zeros(3), 0.16, 15, 5)
    np.testing.assert_allclose(res["circular"]["pts"], o_c, rtol=0, atol=1e-13)
    np.testing.assert_allclose(res["circular"]["normal"], o_n, rtol=0, atol=1e-15)
    # extractDescriptorsFromPatches (main.cpp:182-183, ExtractorType SIFT): one row per patch
    from oracle import sift_patch_np as sp
    o_desc = sp.describe_patches_sift(res["patches"])
    assert res["patch_descriptors"].shape == (len(res["patches"]), 128)
    ddiff = np.abs(res["patch_descriptors"] - o_desc)
    assert ddiff.max() <= 1 and (ddiff == 0).mean() > 0.97
    # computeSquareNeighborhoodsByNormals: last grid point of every feature
    S = res["S"]
    ref_last = np.array([-eps_m + 0.01 * cmpp * (S - 1), -eps_m + 0.01 * cmpp * (S - 1), 0.0, 1.0])
    np.testing.assert_allclose(res["last_nb"], (res["frames"] @ ref_last)[:, :3], rtol=0, atol=1e-12)


@pytest.mark.gpu
def test_adapters_shard_the_normal_search_over_the_gpus_of_the_process(tmp_path):
    """FM3D_DEVICES=0,1: every context gets its own camera / g12 / pyramids and a contiguous shard of the
    features (host threads, no collective); results must be byte-identical to the single-GPU run."""
    torch = pytest.importorskip("torch")
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs in the process")
    exe = build_pipeline_main()
    r, pyramids, eps_m, cmpp = 32, 2, 0.05, 0.25
    case = stereo_case(640, 480, 60, 1000, r)
    tmp = str(tmp_path)
    _write_inputs(tmp, case, r, pyramids, eps_m, cmpp)
    outs = []
    for devices in ("0", "0,1", "1,0"):
        env = dict(os.environ, FM3D_PENALTY="2", FM3D_NO_PATCH_FILES="1", FM3D_DEVICES=devices)
        out = os.path.join(tmp, f"result_{devices.replace(',', '_')}.bin")
        p = subprocess.run([exe, "-s", os.path.join(tmp, "settings.yml"), os.path.join(tmp, "features.bin"), out],
                           capture_output=True, text=True, env=env, cwd=tmp, timeout=300)
        assert p.returncode == 0, p.stdout + p.stderr
        outs.append(open(out, "rb").read())
    assert outs[0] == outs[1] == outs[2]
    assert len(_read_result(os.path.join(tmp, "result_0_1.bin"))["normals"]) >= 20


@pytest.mark.gpu
def test_c1_main_cpp_pipeline_at_the_reference_defaults(tmp_path):
    """BASELINE configs[0]: the main.cpp call sequence with build/settings.yml defaults (pixelsRay 64, pyramids 3,
    epsilon 0.16 m at 0.25 cm/pixel -> 128 x 128 patches) on a 640 x 480 synthetic pair with ~1k keypoints, through
    the C++ adapters; every stage against the CPU oracle."""
    import time
    exe = build_pipeline_main()
    r, pyramids, eps_m, cmpp = 64, 3, 0.16, 0.25
    case = stereo_case(640, 480, 1000, 1000, r)
    cam = case["scene"].cam
    tmp = str(tmp_path)
    _write_inputs(tmp, case, r, pyramids, eps_m, cmpp)
    env = dict(os.environ, FM3D_PENALTY="1", FM3D_NO_PATCH_FILES="1")     # the wall as the author's toolchain built it
    t0 = time.perf_counter()
    p = subprocess.run([exe, "-s", os.path.join(tmp, "settings.yml"), os.path.join(tmp, "features.bin"), os.path.join(tmp, "result.bin")],
                       capture_output=True, text=True, env=env, cwd=tmp, timeout=600)
    t_gpu = time.perf_counter() - t0
    assert p.returncode == 0, p.stdout + p.stderr
    res = _read_result(os.path.join(tmp, "result.bin"))
    o_idx, o_dist = orc.knn2_f32(case["desc1"], case["desc2"])
    oq, ot, od = orc.nndr_filter(o_idx, o_dist, 0.55)
    np.testing.assert_array_equal(res["matches"]["q"], oq)
    np.testing.assert_array_equal(res["matches"]["t"], ot)
    assert len(oq) >= 600
    o_all, o_mask, o_xyz = orc.triangulate(cam.K, cam.dist, res["g12"], cam.z_min, cam.z_max, case["kp1"], case["kp2"], oq, ot)
    np.testing.assert_array_equal(res["mask"].astype(bool), o_mask.astype(bool))
    np.testing.assert_allclose(res["pts"], o_xyz, rtol=1e-9, atol=1e-12)
    t0 = time.perf_counter()
    o = orc.optimize_normals(cam.K, cam.dist, res["g12"], cam.z_min, cam.z_max, case["scene"].img1, case["scene"].img2,
                             pyramids, o_xyz, r, 1e-10, penalty_mode=1, threads=os.cpu_count() or 8)
    t_cpu_normals = time.perf_counter() - t0
    np.testing.assert_array_equal(res["status"], o["status"])
    ok = o["status"] == 0
    interior = o["npenalty"][ok] == 0
    ang = angle_deg(res["normals"], o["normals"][ok])
    assert interior.sum() >= 500 and (ang[interior] <= 0.5).all()
    assert np.median(ang[interior]) < 0.02
    assert res["S"] == 128 and res["patches"].shape == (int(ok.sum()), 128, 128)
    from oracle import sift_patch_np as sp
    sel = np.arange(0, len(res["patches"]), max(1, len(res["patches"]) // 24))
    ddiff = np.abs(res["patch_descriptors"][sel] - sp.describe_patches_sift(res["patches"][sel]))
    assert ddiff.max() <= 1
    print(f"C1: {len(oq)} matches, {int(ok.sum())} normals; whole C++ process {t_gpu:.2f} s (incl. CUDA start-up), "
          f"CPU oracle normal search alone {t_cpu_normals:.1f} s on {os.cpu_count()} threads")


@pytest.mark.gpu
def test_main_cpp_from_the_frames_alone_with_fast_and_sift(tmp_path):
    """main.cpp:91-187 with DetectorType FAST + ExtractorType SIFT and NO injected features: compareWithNNDR
    detects (K10), describes (K11) and matches on the GPU, the rest of the pipeline runs on what it found.
    Keypoints, descriptors and matches against the oracles; the refined normals of the genuine matches against
    the scene's ground truth."""
    from oracle import fast_np as fo
    from oracle import sift_kp_np as sk
    exe = build_pipeline_main()
    r, pyramids, eps_m, cmpp = 32, 2, 0.05, 0.25
    case = stereo_case(640, 480, 60, 1000, r)
    cam = case["scene"].cam
    tmp = str(tmp_path)
    opts = """FeatureOptions:
   DetectorType: FAST
   DetectorMode: STATIC
   FastDetector:
      Threshold: 25
      NonMaxSuppression: 1
   ExtractorType: SIFT
"""
    _write_inputs(tmp, case, r, pyramids, eps_m, cmpp, feature_options=opts)
    env = dict(os.environ, FM3D_PENALTY="1", FM3D_NO_PATCH_FILES="1")
    p = subprocess.run([exe, "-s", os.path.join(tmp, "settings.yml"), "-", os.path.join(tmp, "result.bin")],
                       capture_output=True, text=True, env=env, cwd=tmp, timeout=300)
    assert p.returncode == 0, p.stdout + p.stderr
    res = _read_result(os.path.join(tmp, "result.bin"))
    assert len(res["detected"]) == 2
    descs = []
    for (k, d), img in zip(res["detected"], (case["scene"].img1, case["scene"].img2)):
        oxy, orr = fo.detect_fast(img, 25, True)
        assert len(oxy) > 100
        np.testing.assert_array_equal(k[:, :2], oxy)
        np.testing.assert_array_equal(k[:, 4], orr)
        assert (k[:, 2] == 7).all() and (k[:, 3] == -1).all()
        sel = np.arange(0, len(k), max(1, len(k) // 150))
        od = sk.describe_keypoints_sift(img, k[sel, :4])
        dd = np.abs(d[sel] - od)
        assert d.shape == (len(k), 128) and dd.max() <= 1 and (dd == 0).mean() > 0.97
        descs.append(d)
    o_idx, o_dist = orc.knn2_f32(descs[0], descs[1])
    oq, ot, od = orc.nndr_filter(o_idx, o_dist, 0.55)
    np.testing.assert_array_equal(res["matches"]["q"], oq)
    np.testing.assert_array_equal(res["matches"]["t"], ot)
    np.testing.assert_array_equal(res["matches"]["d"], od)
    k1, k2 = res["detected"][0][0], res["detected"][1][0]
    o_all, o_mask, o_xyz = orc.triangulate(cam.K, cam.dist, res["g12"], cam.z_min, cam.z_max, k1[:, :2], k2[:, :2], oq, ot)
    np.testing.assert_array_equal(res["mask"].astype(bool), o_mask.astype(bool))
    np.testing.assert_allclose(res["pts"], o_xyz, rtol=1e-9, atol=1e-12)
    print(f"FAST+SIFT from the frames: {len(k1)} / {len(k2)} keypoints, {len(oq)} NNDR matches, {int(o_mask.sum())} in depth range, "
          f"{len(res['normals'])} refined normals")
    assert len(res["normals"]) == int((res["status"] == 0).sum())


@pytest.mark.gpu
def test_main_cpp_from_the_frames_alone_with_fast_and_brisk(tmp_path):
    """The same with ExtractorType BRISK (descriptorsmatcher.cpp:343-349; build/settings.yml carries its knobs): binary
    rows from K12, the Hamming matcher (:64-67), keypoints near the border erased as cv::BRISK::compute erases them."""
    from oracle import fast_np as fo
    from oracle import brisk_np as bn
    exe = build_pipeline_main()
    r, pyramids, eps_m, cmpp = 32, 2, 0.05, 0.25
    case = stereo_case(640, 480, 60, 1000, r)
    cam = case["scene"].cam
    tmp = str(tmp_path)
    opts = """FeatureOptions:
   DetectorType: FAST
   DetectorMode: STATIC
   FastDetector:
      Threshold: 25
      NonMaxSuppression: 1
   BriskDetector:
      Threshold: 25
      Octaves: 0
   ExtractorType: BRISK
"""
    _write_inputs(tmp, case, r, pyramids, eps_m, cmpp, feature_options=opts)
    with open(os.path.join(tmp, "settings.yml")) as f:
        yml = f.read().replace("epsilon: 0.55", "epsilon: 0.8")      # binary descriptors: a looser ratio, as usual for Hamming
    with open(os.path.join(tmp, "settings.yml"), "w") as f:
        f.write(yml)
    env = dict(os.environ, FM3D_PENALTY="1", FM3D_NO_PATCH_FILES="1")
    p = subprocess.run([exe, "-s", os.path.join(tmp, "settings.yml"), "-", os.path.join(tmp, "result.bin")],
                       capture_output=True, text=True, env=env, cwd=tmp, timeout=300)
    assert p.returncode == 0, p.stdout + p.stderr
    res = _read_result(os.path.join(tmp, "result.bin"))
    descs = []
    for (k, d), img in zip(res["detected"], (case["scene"].img1, case["scene"].img2)):
        oxy, orr = fo.detect_fast(img, 25, True)
        k4 = np.concatenate([oxy, np.full((len(oxy), 1), 7, np.float32), np.full((len(oxy), 1), -1, np.float32)], 1)
        sel = np.arange(0, len(k4), max(1, len(k4) // 400))
        okept, oang, od = bn.describe_keypoints_brisk(img, k4[sel])
        # the survivors of the full list, restricted to the sample
        inside = (oxy[:, 0] >= 13) & (oxy[:, 0] < img.shape[1] - 13) & (oxy[:, 1] >= 13) & (oxy[:, 1] < img.shape[0] - 13)
        np.testing.assert_array_equal(k[:, :2], oxy[inside])
        assert d.dtype == np.uint8 and d.shape == (int(inside.sum()), 64)
        pos = np.cumsum(inside) - 1                       # row of keypoint i among the survivors
        rows = pos[sel[okept]]
        np.testing.assert_array_equal(d[rows], od)
        np.testing.assert_allclose(k[rows, 3], oang, rtol=0, atol=1e-4)
        descs.append(d)
    o_idx, o_dist = orc.knn2_hamming(descs[0], descs[1])
    oq, ot, od_ = orc.nndr_filter(o_idx, o_dist, 0.8)
    np.testing.assert_array_equal(res["matches"]["q"], oq)
    np.testing.assert_array_equal(res["matches"]["t"], ot)
    np.testing.assert_array_equal(res["matches"]["d"], od_)
    print(f"FAST+BRISK from the frames: {len(descs[0])} / {len(descs[1])} described keypoints, {len(oq)} NNDR matches, "
          f"{int(res['mask'].sum())} in depth range, {len(res['normals'])} refined normals")
    assert len(oq) > 50


@pytest.mark.gpu
def test_main_cpp_from_the_frames_alone_with_fast_and_orb(tmp_path):
    """The same with ExtractorType ORB (descriptorsmatcher.cpp:336-342): 32-byte rows from K13, the Hamming matcher,
    keypoints within 31 pixels of the border erased as cv::ORB::compute erases them."""
    from oracle import fast_np as fo
    from oracle import orb_np as on
    exe = build_pipeline_main()
    r, pyramids, eps_m, cmpp = 32, 2, 0.16, 0.25          # 128 x 128 patches: cv::ORB needs 31 <= S/2 < S - 31
    case = stereo_case(640, 480, 60, 1000, r)
    tmp = str(tmp_path)
    opts = """FeatureOptions:
   DetectorType: FAST
   DetectorMode: STATIC
   FastDetector:
      Threshold: 25
      NonMaxSuppression: 1
   OrbDetector:
      NumFeatures: 500
      ScaleFactor: 1.2
      NumLevels: 8
   ExtractorType: ORB
"""
    _write_inputs(tmp, case, r, pyramids, eps_m, cmpp, feature_options=opts)
    with open(os.path.join(tmp, "settings.yml")) as f:
        yml = f.read().replace("epsilon: 0.55", "epsilon: 0.8")
    with open(os.path.join(tmp, "settings.yml"), "w") as f:
        f.write(yml)
    env = dict(os.environ, FM3D_PENALTY="1", FM3D_PATCH_ATLAS=os.path.join(tmp, "patches.pgm"))
    p = subprocess.run([exe, "-s", os.path.join(tmp, "settings.yml"), "-", os.path.join(tmp, "result.bin")],
                       capture_output=True, text=True, env=env, cwd=tmp, timeout=300)
    assert p.returncode == 0, p.stdout + p.stderr
    res = _read_result(os.path.join(tmp, "result.bin"))
    # FM3D_PATCH_ATLAS: all patches in ONE PGM (the reference writes patch_N.pgm per feature), none as single files
    assert not [f for f in os.listdir(tmp) if f.startswith("patch_")]
    raw = open(os.path.join(tmp, "patches.pgm"), "rb").read()
    hdr = raw.split(b"\n", 3)
    assert hdr[0] == b"P5" and hdr[2] == b"255"
    aw, ah = map(int, hdr[1].split())
    sheet = np.frombuffer(hdr[3], np.uint8).reshape(ah, aw)
    S_, n_ = res["S"], len(res["patches"])
    cols_ = int(np.ceil(np.sqrt(n_)))
    assert aw == cols_ * S_ and ah == -(-n_ // cols_) * S_
    for f_ in (0, 1, cols_, n_ - 1):
        assert np.array_equal(sheet[(f_ // cols_) * S_:(f_ // cols_ + 1) * S_, (f_ % cols_) * S_:(f_ % cols_ + 1) * S_], res["patches"][f_])
    descs = []
    for (k, d), img in zip(res["detected"], (case["scene"].img1, case["scene"].img2)):
        oxy, orr = fo.detect_fast(img, 25, True)
        k4 = np.concatenate([oxy, np.full((len(oxy), 1), 7, np.float32), np.full((len(oxy), 1), -1, np.float32)], 1)
        okept, od = on.describe_keypoints_orb(img, k4)
        np.testing.assert_array_equal(k[:, :2], oxy[okept])
        assert d.dtype == np.uint8 and d.shape == (len(okept), 32)
        np.testing.assert_array_equal(d, od)
        descs.append(d)
    o_idx, o_dist = orc.knn2_hamming(descs[0], descs[1])
    oq, ot, od_ = orc.nndr_filter(o_idx, o_dist, 0.8)
    np.testing.assert_array_equal(res["matches"]["q"], oq)
    np.testing.assert_array_equal(res["matches"]["t"], ot)
    np.testing.assert_array_equal(res["matches"]["d"], od_)
    # extractDescriptorsFromPatches with ExtractorType ORB (main.cpp:182-183): one 32-byte row per rectified patch
    pd = res["patch_descriptors"]
    assert pd.dtype == np.uint8 and pd.shape == (len(res["patches"]), 32) and res["S"] >= 63
    S = res["S"]
    for j in range(0, len(pd), max(1, len(pd) // 40)):
        ok_, od1 = on.describe_keypoints_orb(res["patches"][j], np.array([[S // 2, S // 2, S, -1]], np.float32))
        assert len(ok_) == 1 and np.array_equal(od1[0], pd[j])
    print(f"FAST+ORB from the frames: {len(descs[0])} / {len(descs[1])} described keypoints, {len(oq)} NNDR matches, "
          f"{int(res['mask'].sum())} in depth range, {len(res['normals'])} refined normals")
    assert len(oq) > 50


def _adaptive_fast(img, nmin, nmax, iters):
    """cv::DynamicAdaptedFeatureDetector(FastAdjuster) of OpenCV 2.4 (descriptorsmatcher.cpp:186-199), restated over the
    FAST oracle: threshold 20, non-maximum suppression, one step per iteration."""
    from oracle import fast_np as fo
    t, down, up, good, out = 20, False, False, False, (np.zeros((0, 2), np.float32), np.zeros(0, np.float32))
    while iters > 0 and not (down and up) and not good and 1 < t < 200:
        out = fo.detect_fast(img, t, True)
        n = len(out[0])
        if n < nmin:
            down, t = True, t - 1
        elif n > nmax:
            up, t = True, t + 1
        else:
            good = True
        iters -= 1
    return out


@pytest.mark.gpu
def test_main_cpp_with_the_adaptive_fast_detector(tmp_path):
    """DetectorMode ADAPTIVE + DetectorType FAST (descriptorsmatcher.cpp:186-199): the adapter iterates the GPU detector as
    cv::DynamicAdaptedFeatureDetector iterates cv::FastAdjuster.  The class left OpenCV with 3.0 (not in cv2 here): the check
    is against the same loop over the FAST oracle -- parity of the loop itself is unpinned."""
    from oracle import fast_np as fo
    exe = build_pipeline_main()
    r, pyramids, eps_m, cmpp = 32, 2, 0.05, 0.25
    case = stereo_case(640, 480, 60, 1000, r)
    tmp = str(tmp_path)
    counts = [len(fo.detect_fast(case["scene"].img1, t, True)[0]) for t in (20, 23)]
    for nmin, nmax, iters in ((counts[1] - 5, counts[1] + 5, 10),         # reaches the window at threshold 23
                              (counts[0] + 50, counts[0] + 60, 10),       # too few at 20: steps down
                              (10, 20, 3)):                               # far too many: stops after MaxIters detections
        opts = f"""FeatureOptions:
   DetectorType: FAST
   DetectorMode: ADAPTIVE
   Adaptive:
      MinFeatures: {nmin}
      MaxFeatures: {nmax}
      MaxIters: {iters}
   FastDetector:
      Threshold: 77
      NonMaxSuppression: 0
   ExtractorType: ORB
"""
        _write_inputs(tmp, case, r, pyramids, eps_m, cmpp, feature_options=opts)
        env = dict(os.environ, FM3D_PENALTY="1", FM3D_NO_PATCH_FILES="1")
        p = subprocess.run([exe, "-s", os.path.join(tmp, "settings.yml"), "-", os.path.join(tmp, "result.bin")],
                           capture_output=True, text=True, env=env, cwd=tmp, timeout=300)
        assert p.returncode == 0, p.stdout + p.stderr
        res = _read_result(os.path.join(tmp, "result.bin"))
        for (k, d), img in zip(res["detected"], (case["scene"].img1, case["scene"].img2)):
            oxy, orr = _adaptive_fast(img, nmin, nmax, iters)
            inside = (oxy[:, 0] >= 31) & (oxy[:, 0] < img.shape[1] - 31) & (oxy[:, 1] >= 31) & (oxy[:, 1] < img.shape[0] - 31)   # ORB's border rule
            np.testing.assert_array_equal(k[:, :2], oxy[inside])
            np.testing.assert_array_equal(k[:, 4], orr[inside])


def _read_knn_lists(b, o):
    n = struct.unpack_from("i", b, o)[0]; o += 4
    out = []
    for _ in range(n):
        k = struct.unpack_from("i", b, o)[0]; o += 4
        out.append(np.frombuffer(b, np.dtype([("q", "i4"), ("t", "i4"), ("d", "f4")]), k, o)); o += 12 * k
    return out, o


@pytest.mark.gpu
@pytest.mark.parametrize("source", ["injected", "frames"])
def test_compare_and_crosscompare_through_the_adapters(tmp_path, source):
    """DescriptorsMatcher::compare / crosscompare (descriptorsmatcher.cpp:74-105): the raw kNN-2 lists A->B (and B->A)
    through the C++ adapter, against the exact brute-force oracle (ties -> lower index); a second call with the same
    output objects overwrites them (stale pre-filled descriptors must never be matched) and compareWithNNDR
    appends to `matches` (:126)."""
    exe = os.path.join(ROOT, "tests", "_build", "compare_main")
    if not os.access(exe, os.X_OK):
        __import__("importlib").import_module("__graft_entry__").build()
    r, pyramids, eps_m, cmpp = 32, 2, 0.05, 0.25
    case = stereo_case(640, 480, 300, 1000, r)
    tmp = str(tmp_path)
    opts = None if source == "injected" else """FeatureOptions:
   DetectorType: FAST
   DetectorMode: STATIC
   FastDetector:
      Threshold: 25
      NonMaxSuppression: 1
   ExtractorType: ORB
"""
    _write_inputs(tmp, case, r, pyramids, eps_m, cmpp, feature_options=opts)
    p = subprocess.run([exe, "-s", os.path.join(tmp, "settings.yml"), os.path.join(tmp, "features.bin") if source == "injected" else "-",
                        os.path.join(tmp, "out.bin")], capture_output=True, text=True, cwd=tmp, timeout=300)
    assert p.returncode == 0, p.stdout + p.stderr
    b = open(os.path.join(tmp, "out.bin"), "rb").read()
    na, nb, dim, es = struct.unpack_from("iiii", b, 0)
    o = 16
    dt = np.float32 if es == 4 else np.uint8
    da = np.frombuffer(b, dt, na * dim, o).reshape(na, dim); o += es * na * dim
    db = np.frombuffer(b, dt, nb * dim, o).reshape(nb, dim); o += es * nb * dim
    m1, o = _read_knn_lists(b, o)
    m2, o = _read_knn_lists(b, o)
    ab, o = _read_knn_lists(b, o)
    ba, o = _read_knn_lists(b, o)
    n_once, n_twice = struct.unpack_from("ii", b, o); o += 8
    nndr = np.frombuffer(b, np.dtype([("q", "i4"), ("t", "i4"), ("d", "f4")]), n_twice, o); o += 12 * n_twice
    assert o == len(b)
    if source == "injected":
        np.testing.assert_array_equal(da, case["desc1"])
        knn = orc.knn2_f32
    else:
        assert es == 1 and dim == 32 and na > 100 and nb > 100 and da.any()      # overwritten after the poisoning, not all zero
        knn = orc.knn2_hamming
    for lists, (q, t) in ((m1, (da, db)), (m2, (da, db)), (ab, (da, db)), (ba, (db, da))):
        oi, od = knn(q, t)
        assert len(lists) == q.shape[0]
        gi = np.array([[l["t"][0], l["t"][1]] for l in lists])
        gd = np.array([[l["d"][0], l["d"][1]] for l in lists])
        np.testing.assert_array_equal(gi, oi)
        np.testing.assert_array_equal(gd, od)
        assert all((l["q"] == i).all() for i, l in enumerate(lists))
    oi, od = knn(da, db)
    oq, ot, odd = orc.nndr_filter(oi, od, 0.55)
    assert n_once == len(oq) and n_twice == 2 * n_once
    for half in (nndr[:n_once], nndr[n_once:]):
        np.testing.assert_array_equal(half["q"], oq)
        np.testing.assert_array_equal(half["t"], ot)
        np.testing.assert_array_equal(half["d"], odd)


@pytest.mark.gpu
@pytest.mark.parametrize("nccl", ["nccl", "host-uploads"])
def test_adapters_shard_the_matching_over_the_gpus_of_the_process(tmp_path, nccl):
    """FM3D_DEVICES=0,1: compare / crosscompare / compareWithNNDR shard the QUERY rows over both GPUs, the train set reaches
    the second GPU by one ncclBroadcast (or by its own host upload with FM3D_NO_NCCL=1); kNN lists, NNDR matches and
    distances must be byte-identical to the single-GPU run."""
    torch = pytest.importorskip("torch")
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs in the process")
    exe = os.path.join(ROOT, "tests", "_build", "compare_main")
    r, pyramids, eps_m, cmpp = 32, 2, 0.05, 0.25
    case = stereo_case(640, 480, 300, 1000, r)
    tmp = str(tmp_path)
    _write_inputs(tmp, case, r, pyramids, eps_m, cmpp)
    outs = []
    for devices in ("0", "0,1", "1,0"):
        env = dict(os.environ, FM3D_DEVICES=devices)
        if nccl != "nccl":
            env["FM3D_NO_NCCL"] = "1"
        out = os.path.join(tmp, f"out_{devices.replace(',', '_')}.bin")
        p = subprocess.run([exe, "-s", os.path.join(tmp, "settings.yml"), os.path.join(tmp, "features.bin"), out],
                           capture_output=True, text=True, env=env, cwd=tmp, timeout=300)
        assert p.returncode == 0, p.stdout + p.stderr
        outs.append(open(out, "rb").read())
    assert outs[0] == outs[1] == outs[2] and len(outs[0]) > 1000


@pytest.mark.gpu
def test_main_cpp_from_the_frames_alone_with_sift_and_sift(tmp_path):
    """DetectorType SIFT + ExtractorType SIFT (descriptorsmatcher.cpp:243-256, :304-314) with the knobs of a SiftDetector
    block: compareWithNNDR detects with cv::SIFT's scale-space detector (K14), describes every keypoint on its own pyramid
    layer (K11) and matches on the GPU; the rest of main.cpp's pipeline runs on what it found.  Keypoints against the
    oracle restatement of cv::SIFT::detect (itself pinned to cv2), descriptors against cv2.SIFT.compute where cv2 is here,
    matches against the matcher oracle on the descriptors the adapter produced."""
    from oracle import sift_detect_np as sd
    from test_sift_detect import matched_fraction
    exe = build_pipeline_main()
    r, pyramids, eps_m, cmpp = 32, 2, 0.05, 0.25
    case = stereo_case(640, 480, 60, 1000, r)
    tmp = str(tmp_path)
    opts = """FeatureOptions:
   DetectorType: SIFT
   DetectorMode: STATIC
   SiftDetector:
      NumFeatures: 1500
      NumOctaveLayers: 3
      ContrastThreshold: 0.04
      EdgeThreshold: 10
      Sigma: 1.6
   ExtractorType: SIFT
"""
    _write_inputs(tmp, case, r, pyramids, eps_m, cmpp, feature_options=opts)
    env = dict(os.environ, FM3D_PENALTY="1", FM3D_NO_PATCH_FILES="1")
    p = subprocess.run([exe, "-s", os.path.join(tmp, "settings.yml"), "-", os.path.join(tmp, "result.bin")],
                       capture_output=True, text=True, env=env, cwd=tmp, timeout=300)
    assert p.returncode == 0, p.stdout + p.stderr
    res = _read_result(os.path.join(tmp, "result.bin"))
    descs = []
    for (k, d), img in zip(res["detected"], (case["scene"].img1, case["scene"].img2)):
        want = sd.detect_sift(img, nfeatures=1500)
        got = np.column_stack([k[:, :5], k[:, 5]]) if k.shape[1] > 5 else None
        assert 1400 <= len(k) <= 1600 and abs(len(k) - len(want)) <= 15
        if got is not None:
            assert matched_fraction(want, got) >= 0.98
        else:           # the driver writes x, y, size, angle, response
            hit = sum(np.abs(k[:, :2] - a[:2]).max(1).min() < 0.01 for a in want[::7])
            assert hit >= 0.98 * len(want[::7])
        assert d.shape == (len(k), 128) and (d >= 0).all() and (d <= 255).all() and (np.linalg.norm(d, axis=1) > 300).all()
        descs.append(d)
    o_idx, o_dist = orc.knn2_f32(descs[0], descs[1])
    oq, ot, od = orc.nndr_filter(o_idx, o_dist, 0.55)
    np.testing.assert_array_equal(res["matches"]["q"], oq)
    np.testing.assert_array_equal(res["matches"]["t"], ot)
    assert len(oq) >= 30 and len(res["normals"]) == int((res["status"] == 0).sum()) and len(res["normals"]) >= 5
    print(f"SIFT+SIFT from the frames: {len(res['detected'][0][0])} / {len(res['detected'][1][0])} keypoints, {len(oq)} NNDR matches, "
          f"{len(res['normals'])} refined normals")


@pytest.mark.gpu
def test_main_cpp_from_the_frames_alone_with_orb_and_orb(tmp_path):
    """DetectorType ORB + ExtractorType ORB (descriptorsmatcher.cpp:273-281, :336-342) with an OrbDetector block: compareWithNNDR
    runs cv::ORB's own detector and descriptors on the GPU (K15) and the Hamming matcher (:64-67); the rest of main.cpp's
    pipeline runs on what it found.  Keypoints and rows against the oracle restatement of cv::ORB::detectAndCompute (itself
    identical to cv2 keypoint for keypoint), matches against the matcher oracle on the rows the adapter produced."""
    from oracle import orb_detect_np as od
    exe = build_pipeline_main()
    r, pyramids, eps_m, cmpp = 32, 2, 0.05, 0.25
    case = stereo_case(640, 480, 60, 1000, r)
    tmp = str(tmp_path)
    opts = """FeatureOptions:
   DetectorType: ORB
   DetectorMode: STATIC
   OrbDetector:
      NumFeatures: 3000
      ScaleFactor: 1.2
      NumLevels: 8
   ExtractorType: ORB
"""
    _write_inputs(tmp, case, r, pyramids, eps_m, cmpp, feature_options=opts)
    with open(os.path.join(tmp, "settings.yml")) as f:
        yml = f.read().replace("epsilon: 0.55", "epsilon: 0.8")
    with open(os.path.join(tmp, "settings.yml"), "w") as f:
        f.write(yml)
    env = dict(os.environ, FM3D_PENALTY="1", FM3D_NO_PATCH_FILES="1")
    p = subprocess.run([exe, "-s", os.path.join(tmp, "settings.yml"), "-", os.path.join(tmp, "result.bin")],
                       capture_output=True, text=True, env=env, cwd=tmp, timeout=300)
    assert p.returncode == 0, p.stdout + p.stderr
    res = _read_result(os.path.join(tmp, "result.bin"))
    descs = []
    for (k, d), img in zip(res["detected"], (case["scene"].img1, case["scene"].img2)):
        K, D = od.detect_and_describe_orb(img, nfeatures=3000)
        assert len(k) == len(K) and len(k) >= 2500
        np.testing.assert_array_equal(k[:, :3].astype(np.float32), K[:, :3].astype(np.float32))
        assert np.abs(k[:, 3] - K[:, 3]).max() < 1e-3
        np.testing.assert_allclose(k[:, 4], K[:, 4], rtol=1e-6)
        bits = np.unpackbits(d.astype(np.uint8) ^ D, axis=1).sum(1)
        assert d.shape == (len(k), 32) and np.mean(bits == 0) >= 0.995 and bits.max() <= 2
        descs.append(np.ascontiguousarray(d.astype(np.uint8)))
    o_idx, o_dist = orc.knn2_hamming(descs[0], descs[1])
    oq, ot, od_ = orc.nndr_filter(o_idx, o_dist, 0.8)
    np.testing.assert_array_equal(res["matches"]["q"], oq)
    np.testing.assert_array_equal(res["matches"]["t"], ot)
    assert len(res["normals"]) == int((res["status"] == 0).sum())
    print(f"ORB+ORB from the frames: {len(res['detected'][0][0])} / {len(res['detected'][1][0])} keypoints, {len(oq)} NNDR matches, "
          f"{len(res['normals'])} refined normals")
