"""Parity of the CUDA path (through the C-ABI) against the CPU oracle on seeded inputs.

Bars: bit-exact for indices / bytes / integer work; 1e-9 relative for fp64 geometry (the
north star allows 1e-4 on 3-D points); 0.5 degrees on refined normals (north star), with the
tighter expectation written next to each check.
"""
import numpy as np
import pytest

from common import angle_deg, cam_tuple, car2sph, orc, setup_ctx, stereo_case, synth

pytestmark = pytest.mark.gpu


# ------------------------------------------------------------------ triangulation (K3)
def test_undistort_points(ctx):
    case = stereo_case(640, 480, 40, 1001, 32)
    cam = case["scene"].cam
    ctx.set_camera(cam.K, cam.dist, cam.z_min, cam.z_max)
    rng = np.random.default_rng(3)
    pts = rng.uniform(0, 640, (5000, 2))
    pts[:, 1] *= 0.75
    np.testing.assert_allclose(ctx.undistort_points(pts), orc.undistort_points(cam.K, cam.dist, pts), rtol=0, atol=1e-13)


@pytest.mark.parametrize("noise", [0.0, 0.7])
def test_triangulate(ctx, noise):
    case = stereo_case(640, 480, 40, 1001, 32)
    cam = case["scene"].cam
    ctx.set_camera(cam.K, cam.dist, cam.z_min, cam.z_max)
    ctx.set_g12(cam.g12)
    rng = np.random.default_rng(5)
    n = 3000
    # matches into shuffled keypoint tables, some far off so that the depth gate drops them
    kp1 = np.repeat(case["kp1"], n // 40, axis=0).astype(np.float32)
    kp2 = np.repeat(case["kp2_true"], n // 40, axis=0).astype(np.float32)
    kp2 += rng.normal(0, noise, kp2.shape).astype(np.float32)
    kp2[::7, 0] += rng.uniform(-40, 40, kp2[::7, 0].shape).astype(np.float32)
    perm1, perm2 = rng.permutation(n), rng.permutation(n)
    t1, t2 = np.empty_like(kp1), np.empty_like(kp2)
    t1[perm1], t2[perm2] = kp1, kp2
    qidx, tidx = perm1.astype(np.int32), perm2.astype(np.int32)
    xyz_all, mask, xyz, src = ctx.triangulate(t1, t2, qidx, tidx)
    o_all, o_mask, o_xyz = orc.triangulate(*cam_tuple(cam)[:2], cam.g12, cam.z_min, cam.z_max, t1, t2, qidx, tidx)
    assert 0 < o_mask.sum() < n
    np.testing.assert_array_equal(mask, o_mask)
    np.testing.assert_allclose(xyz_all, o_all, rtol=1e-9, atol=1e-12)     # bar 1e-4 relative
    np.testing.assert_allclose(xyz, o_xyz, rtol=1e-9, atol=1e-12)
    np.testing.assert_array_equal(src, np.nonzero(o_mask)[0])            # order-preserving compaction


def test_triangulate_empty_and_direct(ctx):
    case = stereo_case(640, 480, 40, 1001, 32)
    cam = case["scene"].cam
    ctx.set_camera(cam.K, cam.dist, cam.z_min, cam.z_max)
    ctx.set_g12(cam.g12)
    xyz_all, mask, xyz, src = ctx.triangulate(np.zeros((0, 2), np.float32), np.zeros((0, 2), np.float32))
    assert xyz_all.shape == (0, 3) and xyz.shape == (0, 3)
    xyz_all, mask, xyz, src = ctx.triangulate(case["kp1"], case["kp2_true"])
    np.testing.assert_allclose(xyz_all, case["X"], rtol=0, atol=2e-3)     # ground truth of the scene
    assert mask.all()


# ------------------------------------------------------------------ pyramids (K4)
@pytest.mark.parametrize("shape", [(480, 640), (477, 635), (67, 131), (5, 9), (1080, 1920)])
def test_pyramid_bit_exact(ctx, shape):
    rng = np.random.default_rng(shape[0])
    img1 = rng.integers(0, 256, shape, dtype=np.uint8)
    img2 = rng.integers(0, 256, shape, dtype=np.uint8)
    levels = 3 if min(shape) > 16 else 1
    ctx.set_images(img1, img2, levels)
    for k, img in ((1, img1), (2, img2)):
        ref = orc.pyramid_levels(orc.build_pyramid(img, levels), shape[1], shape[0], levels)
        for l in range(levels + 1):
            got = ctx.get_pyramid_level(k, l)
            assert got.shape == ref[l].shape
            np.testing.assert_array_equal(got, ref[l])


# ------------------------------------------------------------------ matcher (K1/K2)
def _check_knn(idx, dist, o_idx, o_dist):
    np.testing.assert_array_equal(idx, o_idx)
    np.testing.assert_array_equal(dist, o_dist)


@pytest.mark.parametrize("nq,nt", [(300, 360), (1, 1), (5, 2), (1000, 1300), (129, 257), (2048, 4096), (1000, 60_000), (19_072, 1_700)])
def test_match_f32_integer_descriptors_tensor_path(ctx, nq, nt):
    q, t, _ = synth.make_float_descriptors(nq, max(nt - nq, 0), 7 + nq)
    t = t[:nt]
    ctx.set_option("matcher_tensor", 1)
    idx, dist = ctx.match_knn2_f32(q, t)
    o_idx, o_dist = orc.knn2_f32(q, t, threads=8)
    _check_knn(idx, dist, o_idx, o_dist)


_PERSISTENT_CASE = {}


def _persistent_case():
    if not _PERSISTENT_CASE:
        q, t, _ = synth.make_float_descriptors(40_000, 0, 4711)     # 313 query tiles x 12 train tiles
        t = t[:3_000]
        _PERSISTENT_CASE["v"] = (q, t) + tuple(orc.knn2_f32(q, t, threads=8))
    return _PERSISTENT_CASE["v"]


@pytest.mark.parametrize("persistent,splits,min_tiles", [(1, 0, 1), (1, 0, 7), (1, 0, 40), (1, 0, 5000), (0, 0, 1), (0, 3, 1), (0, 12, 1)])
def test_match_f32_tensor_path_persistent_pieces_and_splits(ctx, persistent, splits, min_tiles):
    """The persistent tensor-core matcher cuts the (query tile, train tile) sequence into equal ranges: a CTA contracts the
    tail of one query tile, whole query tiles and the head of another back to back (query tile double buffer, train ring and
    accumulators running across the boundaries), the lists of a query tile come from consecutive CTAs.  313 query tiles x 12
    train tiles on every SM (25 tiles per CTA), on 93 CTAs (`matcher_min_tiles` 40) and on a single one; bit-exact against
    the oracle, and so is the one-CTA-per-(query tile, split) kernel with 1, 3 and 12 splits."""
    q, t, o_idx, o_dist = _persistent_case()
    ctx.set_option("matcher_tensor", 1)
    ctx.set_option("matcher_persistent", persistent)
    ctx.set_option("matcher_splits", splits)
    ctx.set_option("matcher_min_tiles", min_tiles)
    try:
        idx, dist = ctx.match_knn2_f32(q, t)
    finally:
        ctx.set_option("matcher_persistent", 1)
        ctx.set_option("matcher_splits", 0)
        ctx.set_option("matcher_min_tiles", 1)
    _check_knn(idx, dist, o_idx, o_dist)


def test_match_f32_tensor_path_extremes_and_ties(ctx):
    rng = np.random.default_rng(11)
    t = rng.integers(0, 256, (700, 128)).astype(np.float32)
    t[100] = 255.0
    t[101] = 0.0
    t[3] = t[7] = t[20] = t[650]           # duplicates: ties must resolve to the lower index
    q = t[rng.integers(0, 700, 260)].copy()
    q[0] = 0.0
    q[1] = 255.0                            # |q|^2 = |t|^2 = 128*255^2: largest exact accumulations
    q[2] = t[650]
    idx, dist = ctx.match_knn2_f32(q, t)
    o_idx, o_dist = orc.knn2_f32(q, t, threads=8)
    _check_knn(idx, dist, o_idx, o_dist)
    assert tuple(idx[2]) == (3, 7)


def test_match_f32_generic_path(ctx):
    rng = np.random.default_rng(13)
    for dim in (128, 64, 37):
        q = rng.standard_normal((333, dim)).astype(np.float32)
        t = rng.standard_normal((517, dim)).astype(np.float32)
        idx, dist = ctx.match_knn2_f32(q, t)
        o_idx, o_dist = orc.knn2_f32(q, t, threads=8)
        _check_knn(idx, dist, o_idx, o_dist)
    # integer-valued data through the CUDA-core path must agree with the tensor path
    q, t, _ = synth.make_float_descriptors(500, 100, 3)
    ctx.set_option("matcher_tensor", 0)
    a = ctx.match_knn2_f32(q, t)
    ctx.set_option("matcher_tensor", 1)
    b = ctx.match_knn2_f32(q, t)
    _check_knn(a[0], a[1], b[0], b[1])


def test_match_f32_real_valued_tensor_filter_is_exact(ctx):
    """Real-valued float descriptors (SURF-like; not integer-valued, any dim <= 128): the bf16 hi/lo tensor-core
    filter + exact decision must return what the exact CUDA-core path and the oracle return, bit for bit --
    including exact ties (lower index), duplicates, near-duplicates inside the filter's error (handed to the exact
    path), and train sets smaller than the candidate list."""
    rng = np.random.default_rng(5)

    def both(q, t):
        ctx.set_option("matcher_tensor", 1)
        idx, dist = ctx.match_knn2_f32(q, t)
        fb = int(ctx.get_option("matcher_exact_fallback"))
        ctx.set_option("matcher_tensor", 0)
        idx0, dist0 = ctx.match_knn2_f32(q, t)
        ctx.set_option("matcher_tensor", 1)
        np.testing.assert_array_equal(idx, idx0)
        np.testing.assert_array_equal(dist, dist0)
        return idx, dist, fb

    for dim in (64, 128, 100, 36):
        nq, nt = 2300, 2000                               # >= 2^22 pairs: the filtered path
        t = rng.normal(size=(nt, dim)).astype(np.float32)
        t /= np.linalg.norm(t, axis=1, keepdims=True)
        q = (t[rng.integers(0, nt, nq)] + 0.1 * rng.normal(size=(nq, dim))).astype(np.float32)
        q /= np.linalg.norm(q, axis=1, keepdims=True)
        idx, dist, _ = both(q, t)
        o_idx, o_dist = orc.knn2_f32(q[:200], t)          # the plain-C oracle on a slice
        np.testing.assert_array_equal(idx[:200], o_idx)
        np.testing.assert_array_equal(dist[:200], o_dist)
    dim, nq, nt = 64, 2200, 2100
    t = rng.normal(size=(nt, dim)).astype(np.float32)
    t[100] = t[50]; t[2000] = t[50]; t[7] = t[1999]
    q = t[rng.integers(0, nt, nq)].copy()
    q[5] = t[50]; q[6] = t[7]
    idx, dist, _ = both(q, t)
    assert tuple(idx[5]) == (50, 100) and dist[5, 0] == 0 and dist[5, 1] == 0      # three-way tie: the two lowest indices
    assert tuple(idx[6]) == (7, 1999)
    t2 = t.copy()
    t2[300:340] = t2[299] + rng.normal(size=(40, dim)).astype(np.float32) * 1e-6
    q2 = q.copy()
    q2[:50] = t2[299] + rng.normal(size=(50, dim)).astype(np.float32) * 1e-4
    _, _, fb = both(q2, t2)
    assert fb >= 50                                       # the filter cannot separate them: exact path, same answer
    t3 = rng.uniform(0, 255, size=(2100, 128)).astype(np.float32)                  # SIFT range, not integer-valued
    q3 = (t3[rng.integers(0, 2100, 2000)] + rng.normal(size=(2000, 128)) * 3).astype(np.float32)
    both(q3, t3)
    idx, dist, _ = both(rng.normal(size=(1 << 21, 8)).astype(np.float32), rng.normal(size=(3, 8)).astype(np.float32))
    assert idx.min() >= 0 and idx.max() <= 2
    qn = q3.copy(); qn[3, 5] = np.nan                     # NaN in the data: decided by the exact path
    both(qn, t3)


@pytest.mark.parametrize("dim", [64, 128, 36])
def test_match_f32_real_valued_persistent_pieces(ctx, dim):
    """The real-valued tensor filter as a persistent kernel: more query tiles than SMs (157 x 12 or 24 train tiles), every CTA
    contracts the tail of one query tile and the head of another (two query-tile buffers at dim <= 80, one at 128), the
    candidate lists of a query tile come from consecutive CTAs.  Bit-identical to the one-CTA-per-query-tile kernel, to the
    exact CUDA-core path, and (a slice) to the oracle."""
    rng = np.random.default_rng(17 + dim)
    nq, nt = 20_000, 3_000
    t = rng.normal(size=(nt, dim)).astype(np.float32)
    t /= np.linalg.norm(t, axis=1, keepdims=True)
    q = (t[rng.integers(0, nt, nq)] + 0.1 * rng.normal(size=(nq, dim))).astype(np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    got = {}
    try:
        for key, (tensor, pers, mt) in {"persistent": (1, 1, 1), "persistent_few_ctas": (1, 1, 40), "per_tile": (1, 0, 1), "exact": (0, 1, 1)}.items():
            ctx.set_option("matcher_tensor", tensor)
            ctx.set_option("matcher_persistent", pers)
            ctx.set_option("matcher_min_tiles", mt)
            got[key] = ctx.match_knn2_f32(q, t)
    finally:
        ctx.set_option("matcher_tensor", 1)
        ctx.set_option("matcher_persistent", 1)
        ctx.set_option("matcher_min_tiles", 1)
    for key in ("persistent_few_ctas", "per_tile", "exact"):
        _check_knn(got["persistent"][0], got["persistent"][1], got[key][0], got[key][1])
    o_idx, o_dist = orc.knn2_f32(q[:300], t)
    _check_knn(got["persistent"][0][:300], got["persistent"][1][:300], o_idx, o_dist)


def test_match_f32_integer_and_real_valued_calls_alternate(ctx):
    """The context remembers whether the last 128-d float call held integer descriptors (contraction launched without waiting
    for the operand check) or real-valued ones (check first): every order of the two kinds returns the oracle's answer."""
    rng = np.random.default_rng(23)
    qi, ti, _ = synth.make_float_descriptors(2300, 100, 9)
    tr = rng.uniform(0, 255, size=(2100, 128)).astype(np.float32)
    qr = (tr[rng.integers(0, 2100, 2200)] + rng.normal(size=(2200, 128)) * 3).astype(np.float32)
    oi = orc.knn2_f32(qi, ti, threads=8)
    o_r = orc.knn2_f32(qr, tr, threads=8)
    for kind in "irriiri":
        idx, dist = ctx.match_knn2_f32(qi, ti) if kind == "i" else ctx.match_knn2_f32(qr, tr)
        _check_knn(idx, dist, *(oi if kind == "i" else o_r))


def test_match_f32_real_valued_tile_edges(ctx):
    """The tensor filter at the edges of its tiles: train counts around multiples of 128 / 256, 1-3 train descriptors,
    dimensions that are not multiples of 16, large magnitudes, quantised values with many exact ties, exact copies."""
    rng = np.random.default_rng(11)
    cases = [(4, 1), (8, 3), (20, 127), (36, 128), (48, 129), (64, 255), (84, 256), (100, 257), (112, 1000), (128, 4097), (96, 9000), (12, 5)]
    for it, (dim, nt) in enumerate(cases):
        nq = (1 << 22) // nt + 37
        kind = it % 5
        t = rng.normal(size=(nt, dim)).astype(np.float32)
        if kind == 0:
            t /= np.maximum(np.linalg.norm(t, axis=1, keepdims=True), 1e-9)
        elif kind == 1:
            t *= 1000.0
        elif kind == 2:
            t = np.round(t * 4) / 4
        elif kind == 3:
            t = np.abs(t) * 50 + 0.5
        q = (t[rng.integers(0, nt, nq)] + (0.05 if kind != 1 else 50.0) * rng.normal(size=(nq, dim))).astype(np.float32)
        if kind == 2:
            q = np.round(q * 4) / 4
        if kind == 4:
            q[: nq // 3] = t[rng.integers(0, nt, nq // 3)]
        ctx.set_option("matcher_tensor", 1)
        i1, d1 = ctx.match_knn2_f32(q, t)
        ctx.set_option("matcher_tensor", 0)
        i0, d0 = ctx.match_knn2_f32(q, t)
        ctx.set_option("matcher_tensor", 1)
        np.testing.assert_array_equal(i1, i0, err_msg=f"dim {dim} nt {nt} kind {kind}")
        np.testing.assert_array_equal(d1, d0, err_msg=f"dim {dim} nt {nt} kind {kind}")


@pytest.mark.parametrize("nbytes", [32, 64])
@pytest.mark.parametrize("nq,nt", [(300, 360), (3, 1), (1500, 2100)])
def test_match_hamming(ctx, nbytes, nq, nt):
    q, t, _ = synth.make_binary_descriptors(nq, max(nt - nq, 0), 5 + nq, nbytes=nbytes)
    t = t[:nt]
    if nt > 30:
        t[5] = t[9] = t[21]
        q[0] = t[21]
    idx, dist = ctx.match_knn2_hamming(q, t)
    o_idx, o_dist = orc.knn2_hamming(q, t, threads=8)
    _check_knn(idx, dist, o_idx, o_dist)


def test_match_empty_train_and_query(ctx):
    q = np.zeros((4, 128), np.float32)
    idx, dist = ctx.match_knn2_f32(q, np.zeros((0, 128), np.float32))
    assert (idx == -1).all() and np.isinf(dist).all()
    idx, dist = ctx.match_knn2_f32(np.zeros((0, 128), np.float32), q)
    assert idx.shape == (0, 2)


@pytest.mark.parametrize("hamming", [False, True])
def test_match_nndr_and_mutual(ctx, hamming):
    if hamming:
        q, t, gt = synth.make_binary_descriptors(800, 160, 21)
        eps = 0.8
    else:
        q, t, gt = synth.make_float_descriptors(800, 160, 21)
        eps = 0.55
    qi, ti, d, mu = ctx.match_nndr(q, t, eps, hamming=hamming, with_mutual=True)
    o_idx, o_dist = (orc.knn2_hamming if hamming else orc.knn2_f32)(q, t, threads=8)
    oq, ot, od = orc.nndr_filter(o_idx, o_dist, eps)
    np.testing.assert_array_equal(qi, oq)
    np.testing.assert_array_equal(ti, ot)
    np.testing.assert_array_equal(d, od)
    ba_idx, _ = (orc.knn2_hamming if hamming else orc.knn2_f32)(t, q, threads=8)
    np.testing.assert_array_equal(mu, (ba_idx[ot, 0] == oq).astype(np.uint8))
    # the synthetic inliers are recovered
    inl = gt[qi] >= 0
    assert inl.mean() > 0.95 and (ti[inl] == gt[qi][inl]).all()


def test_match_full_size_properties(ctx):
    """BASELINE configs[3] at its largest size (200 k x 200 k) through size-independent properties:
    a set matched against itself returns every descriptor as its own nearest neighbour at distance
    0, with the second neighbour a different index; 64 random queries agree bit-exactly with the CPU
    oracle over the full train set.  Float (tcgen05 path) and binary (popc path)."""
    n = 200_000
    rng = np.random.default_rng(2003)
    t = rng.integers(0, 256, (n, 128)).astype(np.float32)
    idx, dist = ctx.match_knn2_f32(t, t)
    assert (idx[:, 0] == np.arange(n)).all() and (dist[:, 0] == 0).all()
    assert (idx[:, 1] != np.arange(n)).all() and (dist[:, 1] > 0).all()
    sel = rng.choice(n, 64, replace=False)
    q = np.clip(t[sel] + rng.integers(-8, 9, (64, 128)), 0, 255).astype(np.float32)
    gi, gd = ctx.match_knn2_f32(q, t)
    oi, od = orc.knn2_f32(q, t, threads=8)
    np.testing.assert_array_equal(gi, oi)
    np.testing.assert_array_equal(gd, od)
    tb = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    idx, dist = ctx.match_knn2_hamming(tb, tb)
    assert (dist[:, 0] == 0).all() and (dist[:, 1] >= dist[:, 0]).all()
    dup = idx[:, 0] != np.arange(n)                 # exact duplicates in the random set: the lower index wins
    assert dup.sum() <= 4 and (idx[dup, 0] < np.arange(n)[dup]).all()
    qb = tb[sel].copy(); qb[:, :3] ^= 0x5A
    gi, gd = ctx.match_knn2_hamming(qb, tb)
    oi, od = orc.knn2_hamming(qb, tb, threads=8)
    np.testing.assert_array_equal(gi, oi)
    np.testing.assert_array_equal(gd, od)


def test_normal_search_at_bench_scale_properties(ctx):
    """2 000 features of the 1280x720 bench scene (every group layout the scheduler picks): all survive,
    no pass falls back to global taps, the refined normals are within 0.05 deg (median) of ground truth,
    every level stays below lmfit's evaluation cap, and the 1-group and 4-group layouts agree."""
    case = stereo_case(1280, 720, 2000, 1001, 64)
    setup_ctx(ctx, case, 3)
    xyz = case["X"]
    out = {}
    for groups in (1, 4, 0):
        ctx.set_option("normals_groups", groups)
        try:
            out[groups] = ctx.optimize_normals(xyz, 64, 1e-10, 1)
            st = ctx.normals_stats()
        finally:
            ctx.set_option("normals_groups", 0)
        res = out[groups]
        assert (res["status"] == 0).all() and st["passes_slow"] == 0 and st["features"] == xyz.shape[0]
        gt = angle_deg(res["normals"], case["normal"])
        assert np.median(gt) < 0.05 and np.percentile(gt, 99) < 0.5, (np.median(gt), np.percentile(gt, 99))
        assert (res["nfev"] <= 300 + 2).all() and (res["nfev"] >= 4).all()
        np.testing.assert_allclose(np.linalg.norm(res["normals"], axis=1), 1.0, atol=1e-12)
    ang = angle_deg(out[1]["normals"], out[4]["normals"])
    assert np.percentile(ang, 99) < 0.05 and np.median(ang) < 0.005, (np.median(ang), ang.max())


# ------------------------------------------------------------------ normals (K5-K7)
@pytest.fixture(params=["fast-1group", "fast-2groups", "fast-4groups", "faithful"])
def normals_kernel(request, ctx):
    """The implementations of the plane-normal search: fm3d_normals_fast.cu (default: fp32 offset
    geometry, analytic Jacobian) with one feature per CTA (rays in shared memory) or two feature
    pipelines per CTA (rays streamed from an L2-resident scratch), and fm3d_normals.cu (fp64,
    evaluation by evaluation)."""
    ctx.set_option("normals_fast", 0 if request.param == "faithful" else 1)
    ctx.set_option("normals_groups", {"fast-2groups": 2, "fast-4groups": 4}.get(request.param, 1))
    yield request.param
    ctx.set_option("normals_fast", 1)
    ctx.set_option("normals_groups", 0)


@pytest.mark.parametrize("level", [0, 1, 2])
def test_cost_evaluation_fast_kernel(ctx, level):
    """The fp32 offset-form geometry of the fast kernel evaluates the same cost as the oracle
    (different rounding of the sampling coordinates: a few 1e-4 relative)."""
    case = stereo_case(640, 480, 40, 1001, 32)
    cam = case["scene"].cam
    setup_ctx(ctx, case, 2)
    xyz = case["X"][:24]
    rng = np.random.default_rng(level)
    n0 = xyz / np.linalg.norm(xyz, axis=1, keepdims=True)
    pt = car2sph(n0) + rng.normal(0, 0.15, (xyz.shape[0], 2))
    ctx.set_option("normals_fast", 2)
    try:
        cost, m, status = ctx.evaluate_normals(xyz, pt, 32, level, 2)
    finally:
        ctx.set_option("normals_fast", 1)
    o_cost, o_m, o_status = orc.evaluate_cost(*cam_tuple(cam), case["scene"].img1, case["scene"].img2, 2, xyz, pt, 32, level, 2)
    np.testing.assert_array_equal(m, o_m)
    np.testing.assert_array_equal(status, o_status)
    ok = o_status == 0
    assert ok.sum() >= 20
    rel = np.abs(cost[ok] - o_cost[ok]) / o_cost[ok]
    print("fast-kernel cost: max rel diff", rel.max(), "median", np.median(rel))
    np.testing.assert_allclose(cost[ok], o_cost[ok], rtol=1e-3)


@pytest.mark.parametrize("level", [0, 1, 2])
def test_cost_evaluation_matches_oracle(ctx, level):
    case = stereo_case(640, 480, 40, 1001, 32)
    cam = case["scene"].cam
    setup_ctx(ctx, case, 2)
    xyz = case["X"][:24]
    rng = np.random.default_rng(level)
    n0 = xyz / np.linalg.norm(xyz, axis=1, keepdims=True)
    pt = car2sph(n0) + rng.normal(0, 0.15, (xyz.shape[0], 2))
    cost, m, status = ctx.evaluate_normals(xyz, pt, 32, level, 2)
    o_cost, o_m, o_status = orc.evaluate_cost(*cam_tuple(cam), case["scene"].img1, case["scene"].img2, 2, xyz, pt, 32, level, 2)
    np.testing.assert_array_equal(m, o_m)
    np.testing.assert_array_equal(status, o_status)
    ok = o_status == 0
    assert ok.sum() >= 20
    # identical float samples wherever the fp32 cast of the coordinates agrees; allow a few pixels to differ
    np.testing.assert_allclose(cost[ok], o_cost[ok], rtol=2e-5)


@pytest.mark.parametrize("penalty_mode", [2, 0, 1])
def test_optimize_normals_small_disc(ctx, penalty_mode, normals_kernel):
    case = stereo_case(640, 480, 40, 1001, 32)
    cam = case["scene"].cam
    setup_ctx(ctx, case, 2)
    xyz = case["X"]
    res = ctx.optimize_normals(xyz, 32, 1e-10, penalty_mode)
    o = orc.optimize_normals(*cam_tuple(cam), case["scene"].img1, case["scene"].img2, 2, xyz, 32, 1e-10,
                             penalty_mode=penalty_mode, threads=8)
    np.testing.assert_array_equal(res["status"], o["status"])
    ok = o["status"] == 0
    assert ok.sum() >= 30
    # interior features: the oracle never entered the penalty branch (SURVEY 8c stratification)
    interior = ok & (o["npenalty"] == 0)
    ang = angle_deg(res["normals"], o["normals"])
    assert interior.sum() >= 1 or penalty_mode != 2
    assert (ang[interior] <= 0.5).all(), ang[interior].max()          # north-star bar
    assert np.median(ang[interior]) <= 0.01 if interior.any() else True
    if penalty_mode == 2:
        assert interior.sum() == ok.sum()
        gt = angle_deg(res["normals"], case["normal"])
        o_gt = angle_deg(o["normals"], case["normal"])
        assert (gt[ok] <= o_gt[ok] + 0.1).all()
        np.testing.assert_allclose(res["cost"][ok], o["cost"][ok], rtol=0.01)
    else:
        # wall features are reported, not gated
        wall = ok & ~interior
        print("wall features:", int(wall.sum()), "agree<=0.5deg:", int((ang[wall] <= 0.5).sum()))


def test_optimize_normals_default_settings_r64(ctx, normals_kernel):
    """build/settings.yml defaults: pixelsRay 64, pyramids 3 (4 LM stages)."""
    case = stereo_case(640, 480, 24, 1000, 64)
    cam = case["scene"].cam
    setup_ctx(ctx, case, 3)
    xyz = case["X"]
    res = ctx.optimize_normals(xyz, 64, 1e-10, 2)
    o = orc.optimize_normals(*cam_tuple(cam), case["scene"].img1, case["scene"].img2, 3, xyz, 64, 1e-10,
                             penalty_mode=2, threads=8)
    np.testing.assert_array_equal(res["status"], o["status"])
    ok = o["status"] == 0
    ang = angle_deg(res["normals"], o["normals"])
    assert (ang[ok] <= 0.5).all(), ang[ok].max()
    assert (angle_deg(res["normals"], case["normal"])[ok] <= angle_deg(o["normals"], case["normal"])[ok] + 0.1).all()
    print(normals_kernel, "nfev gpu", res["nfev"].sum(0), "oracle", o["nfev"].sum(0), "max angle", ang[ok].max(),
          "median", np.median(ang[ok]))


@pytest.mark.parametrize("fast", [1, 0])
def test_optimize_normals_edge_inputs(ctx, fast):
    """Empty input, non-finite points, points behind the camera, a single pyramid image, five
    pyramid images, the smallest disc: statuses as the oracle's, no crash, no hang."""
    case = stereo_case(640, 480, 40, 1001, 32)
    cam = case["scene"].cam
    ctx.set_option("normals_fast", fast)
    try:
        setup_ctx(ctx, case, 2)
        res = ctx.optimize_normals(np.zeros((0, 3)), 32, 1e-10, 2)
        assert res["normals"].shape == (0, 3) and res["status"].shape == (0,)
        bad = np.array([[np.nan, 0.1, 1.9], [0.1, np.inf, 1.9], [0.0, 0.0, -1.9], [0.0, 0.0, 0.0], [1e9, 1e9, 1.0]])
        xyz = np.concatenate([case["X"][:2], bad, case["X"][2:4]])
        res = ctx.optimize_normals(xyz, 32, 1e-10, 2)
        o = orc.optimize_normals(*cam_tuple(cam), case["scene"].img1, case["scene"].img2, 2, xyz, 32, 1e-10, penalty_mode=2, threads=4)
        print("edge statuses gpu", res["status"], "oracle", o["status"])
        good = [0, 1, 7, 8]
        np.testing.assert_array_equal(res["status"], o["status"])
        assert (res["status"][2:7] != 0).all()                                        # every degenerate point is dropped
        assert (angle_deg(res["normals"][good], o["normals"][good]) <= 0.5).all()
        for pyramids, r in ((0, 32), (5, 32), (2, 1), (2, 3)):
            setup_ctx(ctx, case, pyramids)
            xyz = case["X"][:6]
            res = ctx.optimize_normals(xyz, r, 1e-10, 2)
            o = orc.optimize_normals(*cam_tuple(cam), case["scene"].img1, case["scene"].img2, pyramids, xyz, r, 1e-10,
                                     penalty_mode=2, threads=4)
            np.testing.assert_array_equal(res["status"], o["status"])
            assert res["nfev"].shape == (6, pyramids + 1)
            if r >= 16:
                assert (angle_deg(res["normals"], o["normals"])[o["status"] == 0] <= 0.5).all()
    finally:
        ctx.set_option("normals_fast", 1)


def test_per_evaluation_helpers_against_oracle(ctx):
    """extractPixelsContour / get3dPointsFromImage1Pixels / updateImage1PixelsIntensity /
    projectPointsToImage2 as stand-alone entry points, against the cv2-based restatement, and
    chained by hand against the fused cost kernel."""
    from oracle import oracle_cv as ocv
    case = stereo_case(640, 480, 40, 1001, 32)
    cam = case["scene"].cam
    setup_ctx(ctx, case, 2)
    ccam = ocv.Camera(cam.K, cam.dist, cam.z_min, cam.z_max, g12=cam.g12)
    pyr1, pyr2 = ocv.compute_pyramids(case["scene"].img1, 2), ocv.compute_pyramids(case["scene"].img2, 2)
    H, W = case["scene"].img1.shape
    for f, level in ((0, 0), (3, 1), (7, 2)):
        P = case["X"][f]
        n = P / np.linalg.norm(P)
        n = n + np.array([0.05, -0.03, 0.0]); n /= np.linalg.norm(n)
        scale = 1.0 / (1 << level)
        pix = ctx.disc_pixels(P, 32)
        o_pix = ocv.extract_pixels_contour(ccam, P, 32, W, H)
        np.testing.assert_allclose(pix, o_pix, rtol=0, atol=1e-9)          # same pixels, same (x-outer) order
        pts, info = ctx.plane_points(P, n, pix)
        u = ocv.undistort_points(ccam, o_pix)
        v = np.concatenate([u, np.ones((u.shape[0], 1))], axis=1)
        o_pts = (float(n @ P) / (v @ n))[:, None] * v
        assert info == 0
        np.testing.assert_allclose(pts, o_pts, rtol=1e-12, atol=1e-12)
        i1, info = ctx.sample_pixels(1, level, scale, pix)
        assert info == 0
        np.testing.assert_array_equal(i1, ocv.bilinear32f(pyr1[level], scale * o_pix[:, 0], scale * o_pix[:, 1]))
        xy2, i2, info = ctx.project_to_image2(pts, level, scale)
        r2, t2 = ocv.decompose_transformation(cam.g12)
        o_xy2 = ocv.project_points(ccam, o_pts, r2, t2)
        assert info == 0
        np.testing.assert_allclose(xy2, o_xy2, rtol=0, atol=1e-9)
        o_i2 = ocv.bilinear32f(pyr2[level], scale * o_xy2[:, 0], scale * o_xy2[:, 1])
        assert (i2 != o_i2).mean() < 2e-3                                  # float-cast boundaries of the coordinates only
        # chained by hand == the fused kernel's cost
        ctx.set_option("normals_fast", 0)
        try:
            pt = car2sph(n[None])
            cost, m, st = ctx.evaluate_normals(P[None], pt, 32, level, 2)
        finally:
            ctx.set_option("normals_fast", 1)
        assert m[0] == pix.shape[0] and st[0] == 0
        np.testing.assert_allclose(((i1.astype(np.float64) - i2.astype(np.float64)) ** 2).sum(), cost[0], rtol=1e-6)
    # gates: a plane almost parallel to the rays leaves the bounding box; a point behind image 2's border is not good
    P = case["X"][0]
    pts, info = ctx.plane_points(P, np.array([1.0, 0.0, 0.02]) / np.hypot(1.0, 0.02), ctx.disc_pixels(P, 32))
    assert info == -1
    xy2, i2, info = ctx.project_to_image2(np.array([[5.0, 0.0, 2.0]]), 0, 1.0)
    assert info == -1


def test_sweep_normals_dense_candidate_grid(ctx):
    """BASELINE configs[4] 'dense candidate-normal sampling': every grid point equals a single cost
    evaluation at that (phi, theta) (bitwise: same kernel path) and the oracle's cost (1e-3)."""
    case = stereo_case(640, 480, 40, 1001, 32)
    cam = case["scene"].cam
    setup_ctx(ctx, case, 2)
    xyz = case["X"][:10]
    n0 = xyz / np.linalg.norm(xyz, axis=1, keepdims=True)
    centre = car2sph(n0)
    n_phi, n_theta, dphi, dtheta = 5, 3, 0.04, 0.03
    for level in (0, 2):
        sw = ctx.sweep_normals(xyz, 32, level, n_phi, n_theta, dphi, dtheta, penalty_mode=2)
        sw2 = ctx.sweep_normals(xyz, 32, level, n_phi, n_theta, dphi, dtheta, center_phi_theta=centre, penalty_mode=2, want_cost=False)
        assert (sw["status"] == 0).all()
        np.testing.assert_array_equal(sw["best_idx"], sw2["best_idx"])     # default centre = the viewing ray
        np.testing.assert_array_equal(sw["best_idx"], np.nanargmin(sw["cost"].reshape(len(xyz), -1), axis=1))
        np.testing.assert_array_equal(sw["best_cost"], np.nanmin(sw["cost"].reshape(len(xyz), -1), axis=1))
        ctx.set_option("normals_fast", 2)
        try:
            for i in (0, 2, 4):
                for j in (0, 1, 2):
                    pt = centre + np.array([(i - 2) * dphi, (j - 1) * dtheta])
                    c, m, st = ctx.evaluate_normals(xyz, pt, 32, level, 2)
                    np.testing.assert_array_equal(c, sw["cost"][:, i, j])
                    oc, om, ost = orc.evaluate_cost(*cam_tuple(cam), case["scene"].img1, case["scene"].img2, 2, xyz, pt, 32, level, 2)
                    ok = ost == 0
                    np.testing.assert_allclose(c[ok], oc[ok], rtol=1e-3)
        finally:
            ctx.set_option("normals_fast", 1)


def test_sweep_normals_batched_equals_one_pass_per_candidate(ctx):
    """normals_sweep_batch = 4 (default) evaluates four candidates per pass over the disc: costs, argmin and
    statuses must be bit-identical to one pass per candidate, for grids that are not multiples of the batch,
    for clipped discs and for every layout of the kernel."""
    case = stereo_case(640, 480, 40, 1001, 32)
    setup_ctx(ctx, case, 2)
    xyz = np.concatenate([case["X"][:12], case["X"][:2] * np.array([1.0, 1.0, 1.6])])   # two points that project elsewhere
    try:
        for groups in (1, 2, 4):
            ctx.set_option("normals_groups", groups)
            for (n_phi, n_theta, level, pen) in ((5, 3, 0, 2), (1, 1, 1, 2), (2, 1, 2, 1), (7, 7, 0, 0), (3, 2, 1, 2)):
                out = []
                for batch in (4, 1):
                    ctx.set_option("normals_sweep_batch", batch)
                    out.append(ctx.sweep_normals(xyz, 32, level, n_phi, n_theta, 0.05, 0.04, penalty_mode=pen))
                np.testing.assert_array_equal(out[0]["status"], out[1]["status"])
                np.testing.assert_array_equal(out[0]["cost"], out[1]["cost"])          # NaNs in the same places too
                np.testing.assert_array_equal(out[0]["best_idx"], out[1]["best_idx"])
                np.testing.assert_array_equal(out[0]["best_cost"], out[1]["best_cost"])
                assert np.isfinite(out[0]["cost"][:12]).all()
        # wide grid (steps of 0.35 rad): some candidates stretch the warp beyond the staged window, the batch
        # then hands the rest of the grid to the one-by-one loop with taps from global memory
        ctx.set_option("normals_groups", 4)
        out, slow = [], []
        for batch in (4, 1):
            ctx.set_option("normals_sweep_batch", batch)
            out.append(ctx.sweep_normals(xyz[:12], 32, 0, 9, 9, 0.35, 0.35, penalty_mode=2))
            slow.append(ctx.normals_stats()["passes_slow"])
        assert slow[0] > 0 and slow[0] == slow[1]
        assert np.array_equal(out[0]["cost"], out[1]["cost"], equal_nan=True)
        np.testing.assert_array_equal(out[0]["best_idx"], out[1]["best_idx"])
        np.testing.assert_array_equal(out[0]["status"], out[1]["status"])
    finally:
        ctx.set_option("normals_groups", 0)
        ctx.set_option("normals_sweep_batch", 4)


def test_optimize_normals_large_disc_r128(ctx):
    """pixelsRay 128 (BASELINE configs[4]): m = 51 433 pixels do not fit in shared memory -- rays and
    image-1 samples stream from the L2-resident scratch and the level-0 window is wider than a TMA tile."""
    case = stereo_case(1280, 720, 6, 1004, 128)
    cam = case["scene"].cam
    setup_ctx(ctx, case, 2)
    xyz = case["X"]
    res = ctx.optimize_normals(xyz, 128, 1e-10, 2)
    o = orc.optimize_normals(*cam_tuple(cam), case["scene"].img1, case["scene"].img2, 2, xyz, 128, 1e-10,
                             penalty_mode=2, threads=8)
    np.testing.assert_array_equal(res["status"], o["status"])
    ok = o["status"] == 0
    assert ok.sum() >= 4
    ang = angle_deg(res["normals"], o["normals"])
    st = ctx.normals_stats()
    print("r=128 angles", np.round(ang[ok], 5), "global-tap passes", st["passes_slow"])
    assert (ang[ok] <= 0.5).all()
    assert st["passes_slow"] == 0        # the full-size window is staged: no pass falls back to global taps
    np.testing.assert_allclose(res["cost"][ok], o["cost"][ok], rtol=0.01)


def test_fast_kernel_schedule_options_do_not_change_results(ctx):
    """Evaluating the Jacobian together with a trial (every trial by default; the first of an iteration only; adaptively; never)
    and answering coefficient-identical trial points and Jacobian requests from memory (iterate + the level's last four passes by
    default; the iterate only; its trials only; nothing)
    are exact: same evaluations counted, bit-identical normals, and every weaker setting costs passes."""
    case = stereo_case(640, 480, 40, 1001, 32)
    setup_ctx(ctx, case, 2)
    xyz = case["X"]
    ctx.set_option("normals_groups", 1)
    try:
        base = ctx.optimize_normals(xyz, 32, 1e-10, 1)
        st0 = ctx.normals_stats()
        passes = lambda s_: s_["passes_value"] + s_["passes_jacobian"] + s_["passes_fused"]
        for key, values, default in (("normals_fuse", (0, 1, 2), 3), ("normals_memo", (0, 1, 2), 3)):
            for v in values:
                ctx.set_option(key, v)
                try:
                    alt = ctx.optimize_normals(xyz, 32, 1e-10, 1)
                    st = ctx.normals_stats()
                finally:
                    ctx.set_option(key, default)
                np.testing.assert_array_equal(alt["nfev"], base["nfev"])
                np.testing.assert_array_equal(alt["normals"], base["normals"])
                np.testing.assert_array_equal(alt["status"], base["status"])
                print(key, v, "passes", passes(st), "default", passes(st0))
                assert passes(st) >= passes(st0) and (passes(st) > passes(st0) or (key, v) == ("normals_memo", 2))   # the default did save passes
    finally:
        ctx.set_option("normals_groups", 0)
    assert st0["trials_memoized"] > 0 and st0["fused_accepted"] > 0


def test_optimize_normals_fp32_geometry(ctx):
    case = stereo_case(640, 480, 40, 1001, 32)
    cam = case["scene"].cam
    setup_ctx(ctx, case, 2)
    xyz = case["X"]
    ctx.set_option("geometry_f32", 1)
    ctx.set_option("normals_fast", 0)
    try:
        res = ctx.optimize_normals(xyz, 32, 1e-10, 2)
    finally:
        ctx.set_option("geometry_f32", 0)
        ctx.set_option("normals_fast", 1)
    o = orc.optimize_normals(*cam_tuple(cam), case["scene"].img1, case["scene"].img2, 2, xyz, 32, 1e-10,
                             penalty_mode=2, threads=8)
    ok = (o["status"] == 0) & (res["status"] == 0)
    assert ok.sum() >= 30
    assert (angle_deg(res["normals"], o["normals"])[ok] <= 0.5).all()


def test_optimize_normals_aborts_and_clipped_discs(ctx, normals_kernel):
    """Features whose disc is clipped by the border (D1), projects outside image 2 (abort, D8)
    or has no pixels at all."""
    case = stereo_case(640, 480, 40, 1001, 32)
    cam = case["scene"].cam
    setup_ctx(ctx, case, 2)
    Z = 1.9
    # 3-D points on the rays of chosen image-1 pixels: left border, outside, top border, corner
    pix = np.array([[8.0, 240.0], [300.0, 6.0], [320.0, 472.0], [250.0, 470.0], [400.0, 473.0]])
    # true surface points under those pixels (ray-cast into the scene), so the photometric problem is well posed
    x, y, _ = synth.undistort_exact((pix[:, 0] - cam.K[0, 2]) / cam.K[0, 0], (pix[:, 1] - cam.K[1, 2]) / cam.K[1, 1], cam.dist)
    pts, _, hit = synth._raycast(case["scene"].planes, np.zeros(3), np.stack([x, y, np.ones_like(x)], axis=1))
    assert hit.all()
    far = np.array([[9.0, 0.3, Z]])                       # projects far outside image 1: no pixels
    xyz = np.concatenate([case["X"][:1], pts, far, case["X"][1:2]])
    res = ctx.optimize_normals(xyz, 32, 1e-10, 2)
    o = orc.optimize_normals(*cam_tuple(cam), case["scene"].img1, case["scene"].img2, 2, xyz, 32, 1e-10,
                             penalty_mode=2, threads=4)
    np.testing.assert_array_equal(res["status"], o["status"])
    assert (o["status"] == 1).any() and ((o["m"] > 0) & (o["m"] < 3209)).sum() >= 3
    print("statuses", o["status"], "m", o["m"])
    ok = o["status"] == 0
    # clipped discs straddle facet edges of the scene, so their cost surface has several minima:
    # the bar is "same minimum as the oracle (0.5 deg) or a strictly better one (lower final cost)"
    ang = angle_deg(res["normals"], o["normals"])
    better = res["cost"] < 0.99 * o["cost"]
    print("angles", np.round(ang[ok], 4), "better-minimum", better[ok])
    assert ((ang <= 0.5) | better)[ok].all()
    assert (ang[ok] <= 0.5).sum() >= ok.sum() - 1
    cost, m, st = ctx.evaluate_normals(xyz, car2sph(xyz / np.linalg.norm(xyz, axis=1, keepdims=True)), 32, 0, 2)
    np.testing.assert_array_equal(m, o["m"])


# ------------------------------------------------------------------ frames + patches (K7/K8)
def test_frames_and_patches(ctx):
    case = stereo_case(640, 480, 40, 1001, 32)
    cam = case["scene"].cam
    setup_ctx(ctx, case, 2)
    xyz, normals = case["X"][:12], case["normal"][:12]
    g = np.array([0.006, 0.99992, -0.011])
    frames = ctx.feature_frames(xyz, normals, g)
    o_frames = orc.feature_frames(xyz, normals, g)
    np.testing.assert_allclose(frames, o_frames, rtol=0, atol=1e-14)
    for eps_m, cmpp in ((0.16, 0.25), (0.05, 0.25), (0.0825, 0.5)):
        patches, ip = ctx.extract_patches(frames, eps_m, cmpp, want_points=True)
        o_patches, o_ip = orc.extract_patches(cam.K, cam.dist, case["scene"].img1, o_frames, eps_m, cmpp)
        assert patches.shape == o_patches.shape
        np.testing.assert_allclose(ip, o_ip, rtol=0, atol=1e-9)
        diff = patches.astype(int) - o_patches.astype(int)
        assert (np.abs(diff) <= 1).all() and (diff != 0).mean() < 1e-3    # float-cast boundaries only
        p2, _ = ctx.extract_patches(frames, eps_m, cmpp, want_points=False)
        np.testing.assert_array_equal(p2, patches)
    nb = ctx.square_neighborhoods(frames, 0.05, 0.25)
    S = orc.patch_size(0.05, 0.25)
    i, j = np.meshgrid(np.arange(S), np.arange(S), indexing="ij")
    ref = np.stack([-0.05 + 0.0025 * i.ravel(), -0.05 + 0.0025 * j.ravel(), np.zeros(S * S), np.ones(S * S)], 1)
    np.testing.assert_allclose(nb, np.einsum("nij,pj->npi", o_frames, ref)[:, :, :3], rtol=0, atol=1e-12)
    # explicit groups into either image (projectPointsToImage)
    for image in (1, 2):
        gp, gip = ctx.project_groups(image, nb)
        o_ip2 = np.stack([orc.project_points(cam.K, cam.dist, cam.g12, image, grp) for grp in nb])
        np.testing.assert_allclose(gip, o_ip2, rtol=0, atol=1e-9)
    gp1, _ = ctx.project_groups(1, nb)
    p1, _ = ctx.extract_patches(frames, 0.05, 0.25, want_points=False)
    assert (np.abs(gp1.astype(int) - p1.astype(int)) <= 1).all()


def test_circular_neighborhoods(ctx):
    """NeighborhoodsGenerator's circular variant (never called by main.cpp, kept for the interface)."""
    from oracle import oracle_cv as ocv
    rng = np.random.default_rng(5)
    P = rng.normal(0, 0.3, (7, 3)) + np.array([0, 0, 2.0])
    N = rng.normal(0, 0.3, (7, 3)) + np.array([0, 0, 1.0])
    N /= np.linalg.norm(N, axis=1, keepdims=True)
    N[2] = 0                                            # "initial guess": P/|P|
    out, used = ctx.circular_neighborhoods(P, N, 0.16, 15, 5)
    assert out.shape == (7, 75, 3)
    for f in range(7):
        o, n_used = ocv.circular_neighborhood(P[f], N[f], 0.16, 15, 5)
        np.testing.assert_allclose(out[f], o, rtol=0, atol=1e-13)
        np.testing.assert_allclose(used[f], n_used, rtol=0, atol=1e-15)
        np.testing.assert_allclose((out[f] - P[f]) @ n_used, 0, atol=1e-12)      # every sample lies in the plane


# ------------------------------------------------------------------ committed golden vectors (cv2)
def test_gpu_against_golden_vectors(ctx):
    """The CUDA path against tests/golden/*.npz: outputs of cv2.undistortPoints / triangulatePoints /
    pyrDown / BFMatcher and of the cv2-based restatement of the reference (tools/make_golden.py)."""
    import os
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    p = np.load(os.path.join(gold, "primitives.npz"))
    ctx.set_camera(p["K"], p["dist"], 1.5, 2.4)
    ctx.set_g12(p["g12"])
    np.testing.assert_allclose(ctx.undistort_points(p["pts"]), p["undistorted"], rtol=1e-12, atol=1e-13)
    xyz_all, mask, xyz, _ = ctx.triangulate(p["kp1"], p["kp2"])
    np.testing.assert_array_equal(mask, p["mask"])
    np.testing.assert_allclose(xyz_all, p["xyz_all"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(xyz, p["xyz"], rtol=1e-9, atol=1e-12)
    ctx.set_images(p["img"], p["img"], 3)
    for k in (1, 2, 3):
        np.testing.assert_array_equal(ctx.get_pyramid_level(1, k), p[f"pyr{k}"])
    idx, dist = ctx.match_knn2_f32(p["q"], p["t"])
    np.testing.assert_array_equal(idx, p["idx_f"])
    np.testing.assert_array_equal(dist, p["dist_f"])
    idx, dist = ctx.match_knn2_hamming(p["qb"], p["tb"])
    np.testing.assert_array_equal(idx, p["idx_b"])
    np.testing.assert_array_equal(dist, p["dist_b"])
    qi, ti, d = ctx.match_nndr(p["q"], p["t"], 0.55)
    np.testing.assert_array_equal(qi, p["nndr_q"])
    np.testing.assert_array_equal(ti, p["nndr_t"])

    g = np.load(os.path.join(gold, "normals.npz"))
    ctx.set_camera(g["K"], g["dist"], float(g["zmin"]), float(g["zmax"]))
    ctx.set_g12(g["g12"])
    ctx.set_images(g["img1"], g["img2"], 2)
    for lvl in range(3):
        cost, m, st = ctx.evaluate_normals(g["xyz"], g["cost_pt"], 16, lvl, 2)
        np.testing.assert_allclose(cost, g[f"cost_l{lvl}"], rtol=1e-5)
    res = ctx.optimize_normals(g["xyz"], 16, 1e-10, 2)
    np.testing.assert_array_equal(res["status"], g["off_status"])
    assert (angle_deg(res["normals"], g["off_normals"]) <= 0.5).all()
    assert (angle_deg(res["normals"], g["gt_normal"]) <= 0.5).all()
    frames = ctx.feature_frames(g["xyz"], g["off_normals"], g["gravity"])
    np.testing.assert_allclose(frames, g["frames"], rtol=0, atol=1e-14)
    patches, ip = ctx.extract_patches(g["frames"], 0.05, 0.25)
    np.testing.assert_allclose(ip, g["image_points"], rtol=0, atol=1e-9)
    diff = patches.astype(int) - g["patches"].astype(int)
    assert (np.abs(diff) <= 1).all() and (diff != 0).mean() < 1e-3


# ------------------------------------------------------------------ patch descriptors (K9)
def _desc_close(got, want, frac_exact=0.97):
    diff = np.abs(np.asarray(got).astype(np.int32) - np.asarray(want).astype(np.int32))
    assert got.shape == want.shape and diff.max(initial=0) <= 1, diff.max(initial=0)
    if diff.size:
        assert (diff == 0).mean() >= frac_exact, (diff == 0).mean()


def test_describe_patches_sift_golden_vectors(ctx):
    """extractDescriptorsFromPatches (descriptorsmatcher.cpp:133-174, SIFT): the committed outputs of
    cv2.SIFT_create().compute.  Values are quantised to integers: +-1 on isolated entries."""
    import os
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    g = np.load(os.path.join(gold, "sift_patches.npz"))
    for S in (128, 64, 40, 16, 8):
        d = ctx.describe_patches_sift(g[f"p{S}"])
        assert d.dtype == np.float32 and d.shape == (g[f"p{S}"].shape[0], 128)
        _desc_close(d, g[f"d{S}"])
        assert np.count_nonzero(d[-2]) == 0            # constant patch
    pipe = np.load(os.path.join(gold, "normals.npz"))["patches"]
    _desc_close(ctx.describe_patches_sift(pipe), g["d_pipeline"])


def test_describe_patches_sift_against_oracle_and_edges(ctx, api):
    from oracle import sift_patch_np as sp
    rng = np.random.default_rng(123)
    for S, n in ((128, 40), (130, 5), (160, 3), (66, 17), (9, 6)):
        patches = rng.integers(0, 256, (n, S, S), dtype=np.uint8)
        # smooth half of them (box filter) so that orientations are not uniform
        sm = patches.astype(np.float32)
        sm = (sm + np.roll(sm, 1, 1) + np.roll(sm, 1, 2) + np.roll(np.roll(sm, 1, 1), 1, 2)) / 4
        patches[::2] = sm[::2].astype(np.uint8)
        patches[-1, :, :] = 0
        patches[-1, :, S // 3:] = 255                  # saturated step edge
        got = ctx.describe_patches_sift(patches)
        _desc_close(got, sp.describe_patches_sift(patches))
        # deterministic: fixed summation order, no atomics
        assert np.array_equal(got, ctx.describe_patches_sift(patches))
        # one patch alone gives the row it gave in the batch
        assert np.array_equal(ctx.describe_patches_sift(patches[1:2])[0], got[1])
    assert ctx.describe_patches_sift(np.zeros((0, 128, 128), np.uint8)).shape == (0, 128)
    for S in (4, 200):                                 # outside the supported patch edges
        with pytest.raises(api.Fm3dError):
            ctx.describe_patches_sift(np.zeros((1, S, S), np.uint8))


def test_describe_patches_pipeline_on_device(ctx, api):
    """frames -> K8 patches -> K9 descriptors without leaving the device equals the host-buffer calls."""
    torch = pytest.importorskip("torch")
    case = stereo_case(640, 480, 40, 1001, 32)
    setup_ctx(ctx, case, 2)
    X = case["X"][:24]
    frames = ctx.feature_frames(X, case["normal"][:24], np.array([0.006, 0.99992, -0.011]))
    patches, _ = ctx.extract_patches(frames, 0.16, 0.25, want_points=False)
    want = ctx.describe_patches_sift(patches)
    S = api.patch_size(0.16, 0.25)
    dev = torch.device("cuda", 0)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    d_frames = torch.from_numpy(frames.reshape(-1, 16)).to(dev)
    d_patches = torch.empty((len(X), S, S), dtype=torch.uint8, device=dev)
    d_desc = torch.empty((len(X), 128), dtype=torch.float32, device=dev)
    torch.cuda.synchronize()
    ctx.extract_patches_dev(d_frames.data_ptr(), len(X), 0.16, 0.25, d_patches.data_ptr(), None)
    ctx.describe_patches_sift_dev(d_patches.data_ptr(), len(X), S, d_desc.data_ptr())
    stream.synchronize()
    assert np.array_equal(d_patches.cpu().numpy(), patches)
    assert np.array_equal(d_desc.cpu().numpy(), want)
    from oracle import sift_patch_np as sp
    _desc_close(want, sp.describe_patches_sift(patches))
    assert np.count_nonzero(want) > 0


# ------------------------------------------------------------------ K10: FAST keypoint detection
def test_detect_fast_golden_vectors(ctx):
    """feature_detector_->detect for DetectorType FAST (descriptorsmatcher.cpp:110-111, :215-222): the committed
    outputs of cv2.FastFeatureDetector.  Integer work: positions, order and responses identical."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fast_keypoints.npz"))
    for name in ("noise", "blur", "frame", "tiny7", "tiny6"):
        img = g[f"img_{name}"]
        for t in (0, 1, 10, 20, 40, 100, 255):
            for nm in (0, 1):
                xy, r, n = ctx.detect_fast(img, t, bool(nm))
                want = g[f"xy_{name}_{t}_{nm}"].astype(np.float32).reshape(-1, 2)
                assert n == len(want), (name, t, nm, n, len(want))
                assert np.array_equal(xy, want) and np.array_equal(r, g[f"r_{name}_{t}_{nm}"].astype(np.float32)), (name, t, nm)


def test_detect_fast_against_oracle_and_edges(ctx, api):
    from oracle import fast_np as fo
    rng = np.random.default_rng(808)
    for h, w in ((480, 640), (33, 1025), (1000, 9), (7, 7), (3, 40), (1, 1)):
        img = rng.integers(0, 256, (h, w), dtype=np.uint8)
        if h >= 32 and w >= 32:
            img[: h // 2] = (img[: h // 2].astype(np.int32) * 3 // 8 + 90).astype(np.uint8)      # a low-contrast half
        for t, nm in ((12, True), (30, False), (0, True)):
            xy, r, n = ctx.detect_fast(img, t, nm)
            oxy, orr = fo.detect_fast(img, t, nm)
            assert n == len(oxy) and np.array_equal(xy, oxy) and np.array_equal(r, orr), (h, w, t, nm)
    # a padded (strided) image gives what its packed copy gives
    big = rng.integers(0, 256, (120, 256), dtype=np.uint8)
    view = big[:, 13:173]
    a = ctx.detect_fast(view, 25, True)
    b = ctx.detect_fast(np.ascontiguousarray(view), 25, True)
    assert a[2] == b[2] > 0 and np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    # truncation keeps the first keypoints in row-major order and still reports the total
    full = ctx.detect_fast(big, 25, True)
    part = ctx.detect_fast(big, 25, True, max_keypoints=10)
    assert part[2] == full[2] and len(part[0]) == 10 and np.array_equal(part[0], full[0][:10]) and np.array_equal(part[1], full[1][:10])
    # thresholds outside [0, 255] are clamped as cv::FAST's C++ path clamps them
    assert ctx.detect_fast(big, 300, True)[2] == 0
    assert ctx.detect_fast(big, -4, False)[2] == ctx.detect_fast(big, 0, False)[2]
    # constant image: no corners
    assert ctx.detect_fast(np.full((64, 64), 128, np.uint8), 0, False)[2] == 0


def test_detect_fast_full_size_properties(ctx):
    """4K frame (BASELINE C3 shape): oracle parity, and size-independent properties -- suppressed keypoints are a
    subset of the unsuppressed ones, raising the threshold only removes corners, detection is deterministic."""
    from oracle import fast_np as fo
    rng = np.random.default_rng(4)
    small = rng.integers(0, 256, (2160 // 8 + 1, 3840 // 8 + 1)).astype(np.float32)
    img = np.kron(small, np.ones((8, 8), np.float32))[:2160, :3840]
    img = np.clip(img * 0.6 + rng.integers(0, 100, img.shape), 0, 255).astype(np.uint8)
    xy, r, n = ctx.detect_fast(img, 20, True)
    oxy, orr = fo.detect_fast(img, 20, True)
    assert n == len(oxy) > 1000 and np.array_equal(xy, oxy) and np.array_equal(r, orr)
    allc = ctx.detect_fast(img, 20, False)
    key = lambda a: set(map(tuple, a.astype(np.int64)))
    assert key(xy) <= key(allc[0])
    higher = ctx.detect_fast(img, 35, False)
    assert key(higher[0]) <= key(allc[0]) and higher[2] < allc[2]
    again = ctx.detect_fast(img, 20, True)
    assert np.array_equal(again[0], xy) and np.array_equal(again[1], r)
    assert (np.diff(xy[:, 1] * 4096 + xy[:, 0]) > 0).all()          # strictly row-major


# ------------------------------------------------------------------ K11: SIFT descriptors at frame keypoints
def test_describe_keypoints_sift_golden_vectors(ctx):
    """descriptor_extractor_->compute (descriptorsmatcher.cpp:114-115, ExtractorType SIFT) on octave-0 keypoints:
    the committed outputs of cv2.SIFT_create().compute.  Values are quantised to integers: +-1 on isolated
    entries (summation order)."""
    import os
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    g = np.load(os.path.join(gold, "sift_keypoints.npz"))
    imgs = np.load(os.path.join(gold, "fast_keypoints.npz"))
    seen = 0
    for key in g.files:
        if not key.startswith("k_"):
            continue
        name = key.split("_")[1]
        d = ctx.describe_keypoints_sift(imgs[f"img_{name}"], g[key])
        assert d.dtype == np.float32 and d.shape == (len(g[key]), 128)
        _desc_close(d, g["d" + key[1:]])
        seen += len(d)
    assert seen > 1000


def test_describe_keypoints_sift_against_oracle_and_edges(ctx):
    from oracle import sift_kp_np as sk
    rng = np.random.default_rng(321)
    for h, w in ((97, 141), (64, 200), (33, 35)):
        img = rng.integers(0, 256, (h, w)).astype(np.float32)
        img = ((img + np.roll(img, 1, 0) + np.roll(img, 1, 1) + np.roll(np.roll(img, 1, 0), 1, 1)) / 4).astype(np.uint8)
        n = 90
        k = np.stack([rng.uniform(0, w - 1, n), rng.uniform(0, h - 1, n), rng.uniform(1.2, 40, n), rng.uniform(0, 360, n)], 1).astype(np.float32)
        k[:15, 3] = -1                                       # FAST's "no orientation"
        k[15:30, 2] = 7                                      # FAST's size
        k[30:34, :2] = [[0, 0], [w - 1, h - 1], [0, h - 1], [w - 1, 0]]
        k[34, 3] = 0
        k[35, 2] = 300                                       # window larger than the image
        got = ctx.describe_keypoints_sift(img, k)
        _desc_close(got, sk.describe_keypoints_sift(img, k))
        # deterministic: fixed summation order, no atomics
        assert np.array_equal(got, ctx.describe_keypoints_sift(img, k))
        # one keypoint alone gives the row it gave in the batch; a strided image gives what its packed copy gives
        assert np.array_equal(ctx.describe_keypoints_sift(img, k[7:8])[0], got[7])
        big = np.zeros((h, w + 37), np.uint8)
        big[:, 5:5 + w] = img
        assert np.array_equal(ctx.describe_keypoints_sift(big[:, 5:5 + w], k), got)
        # keypoints DescriptorExtractor::compute would have removed: zero rows, the others unchanged
        bad = k.copy()
        bad[[2, 40], 0] = [-1.0, w]
        bad[41, 1] = h + 3
        bad[42, 2] = 0
        z = ctx.describe_keypoints_sift(img, bad)
        assert not z[[2, 40, 41, 42]].any()
        keep = np.setdiff1d(np.arange(n), [2, 40, 41, 42])
        assert np.array_equal(z[keep], got[keep])
    assert ctx.describe_keypoints_sift(np.zeros((40, 40), np.uint8), np.zeros((0, 4), np.float32)).shape == (0, 128)
    assert not ctx.describe_keypoints_sift(np.full((40, 40), 9, np.uint8), [[20, 20, 7, -1]]).any()     # constant image


def test_detect_describe_match_chain_full_size(ctx):
    """DetectorType FAST + ExtractorType SIFT on a 1280 x 720 frame (BASELINE C2 shape), all three steps of
    compareWithNNDR (descriptorsmatcher.cpp:110-117) on the GPU: oracle parity on a sample of the keypoints, and
    size-independent properties -- describing a frame against itself matches every keypoint to itself at
    distance 0, a translated copy of the frame gives the same descriptors at the translated keypoints."""
    from oracle import sift_kp_np as sk
    from oracle import fast_np as fo
    rng = np.random.default_rng(11)
    small = rng.integers(0, 256, (720 // 6 + 4, 1280 // 6 + 4)).astype(np.float32)
    big = np.kron(small, np.ones((6, 6), np.float32))
    big = np.clip(big * 0.7 + rng.integers(0, 77, big.shape), 0, 255).astype(np.uint8)
    img = np.ascontiguousarray(big[:720, :1280])
    xy, r, n = ctx.detect_fast(img, 30, True)
    assert n == len(xy) > 2000
    k = np.concatenate([xy, np.full((n, 1), 7, np.float32), np.full((n, 1), -1, np.float32)], 1)
    d = ctx.describe_keypoints_sift(img, k)
    pick = rng.choice(n, 200, replace=False)
    _desc_close(d[pick], sk.describe_keypoints_sift(img, k[pick]))
    idx, dist = ctx.match_knn2_f32(d, d)
    same = idx[:, 0] == np.arange(n)
    assert (dist[:, 0] == 0).all() and (same | (dist[:, 1] == 0)).all()       # itself, or an exact duplicate with a lower index
    # translation by (dx, dy): keypoints whose window stays inside both frames keep their descriptor exactly
    dx, dy = 9, 5
    img2 = np.ascontiguousarray(big[dy:dy + 720, dx:dx + 1280])
    inner = (xy[:, 0] > 60 + dx) & (xy[:, 0] < 1280 - 60) & (xy[:, 1] > 60 + dy) & (xy[:, 1] < 720 - 60)
    k2 = k[inner].copy()
    k2[:, 0] -= dx
    k2[:, 1] -= dy
    assert inner.sum() > 1000 and np.array_equal(ctx.describe_keypoints_sift(img2, k2), d[inner])


# ------------------------------------------------------------------ K12: BRISK descriptors at frame keypoints
def test_describe_keypoints_brisk_golden_vectors(ctx):
    """descriptor_extractor_->compute (descriptorsmatcher.cpp:114-115, ExtractorType BRISK): the committed outputs of
    cv2.BRISK_create(25, 0).compute.  Integer work: removed keypoints and all 512 bits identical."""
    import os
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    g = np.load(os.path.join(gold, "brisk_keypoints.npz"))
    imgs = np.load(os.path.join(gold, "fast_keypoints.npz"))
    seen = 0
    for key in g.files:
        if not key.startswith("k_"):
            continue
        name, tag = key.split("_")[1:3]
        d, kept, ang = ctx.describe_keypoints_brisk(imgs[f"img_{name}"], g[key])
        assert d.dtype == np.uint8 and d.shape == (len(g[key]), 64)
        np.testing.assert_array_equal(np.nonzero(kept)[0], g[f"kept_{name}_{tag}"])
        np.testing.assert_array_equal(d[kept], g[f"d_{name}_{tag}"])
        assert not d[~kept].any()
        np.testing.assert_allclose(ang[kept], g[f"a_{name}_{tag}"], rtol=0, atol=1e-4)
        seen += int(kept.sum())
    assert seen > 800


def test_describe_keypoints_brisk_against_oracle_and_edges(ctx):
    from oracle import brisk_np as bn
    rng = np.random.default_rng(654)
    for h, w in ((120, 171), (90, 300)):
        img = rng.integers(0, 256, (h, w)).astype(np.float32)
        img = ((img + np.roll(img, 1, 0) + np.roll(img, 1, 1) + np.roll(np.roll(img, 1, 0), 1, 1)) / 4).astype(np.uint8)
        n = 150
        k = np.stack([rng.uniform(0, w - 1, n), rng.uniform(0, h - 1, n), rng.uniform(2, 26, n), rng.uniform(0, 360, n)], 1).astype(np.float32)
        k[:40, 2] = 7
        k[:40, 3] = -1                                       # FAST keypoints
        k[:20, :2] = np.floor(k[:20, :2])
        k[40:44, :2] = [[0, 0], [w - 1, h - 1], [13, 13], [w - 14, h - 14]]      # the border rule: 13 <= x < w - 13 at scale 0
        k[40:44, 2] = 7
        k[44, 2] = 0                                         # size 0
        k[45, 2] = 200                                       # scale index saturates at 63: the pattern cannot fit
        for orient in (True, False):
            d, kept, ang = ctx.describe_keypoints_brisk(img, k, compute_orientation=orient)
            okept, oang, od = bn.describe_keypoints_brisk(img, k, compute_orientation=orient)
            np.testing.assert_array_equal(np.nonzero(kept)[0], okept)
            assert kept[42] and kept[43] and not kept[40] and not kept[41] and not kept[44] and not kept[45]
            np.testing.assert_allclose(ang[kept], oang, rtol=0, atol=1e-4)
            np.testing.assert_array_equal(d[kept], od)
            assert not d[~kept].any()
            # deterministic, independent of the batch, independent of the row pitch
            again = ctx.describe_keypoints_brisk(img, k, compute_orientation=orient)
            assert np.array_equal(again[0], d) and np.array_equal(again[2], ang)
            one = ctx.describe_keypoints_brisk(img, k[60:61], compute_orientation=orient)
            assert np.array_equal(one[0][0], d[60]) and one[1][0] == kept[60]
            big = np.zeros((h, w + 19), np.uint8)
            big[:, 7:7 + w] = img
            assert np.array_equal(ctx.describe_keypoints_brisk(big[:, 7:7 + w], k, compute_orientation=orient)[0], d)
        # without the orientation step a keypoint with angle -1 is sampled unrotated: the row of rotation index 0
        d0 = ctx.describe_keypoints_brisk(img, k[:40], compute_orientation=False)
        z = k[:40].copy()
        z[:, 3] = 0.0
        assert np.array_equal(d0[0], ctx.describe_keypoints_brisk(img, z, compute_orientation=False)[0])
    d, kept, ang = ctx.describe_keypoints_brisk(np.zeros((40, 40), np.uint8), np.zeros((0, 4), np.float32))
    assert d.shape == (0, 64) and kept.shape == (0,)


def test_detect_brisk_hamming_chain_full_size(ctx):
    """DetectorType FAST + ExtractorType BRISK on a 1280 x 720 frame (BASELINE C2 shape): detection, binary description
    and the Hamming matcher on the GPU; oracle parity on a sample, and size-independent properties -- a frame matched
    against itself returns every keypoint at distance 0, an integer translation of the frame leaves the rows of the
    translated keypoints unchanged up to the float rounding of pattern point + keypoint position (a changed sub-pixel
    weight can move a smoothed value by one unit: isolated bits)."""
    from oracle import brisk_np as bn
    rng = np.random.default_rng(12)
    small = rng.integers(0, 256, (720 // 6 + 4, 1280 // 6 + 4)).astype(np.float32)
    big = np.kron(small, np.ones((6, 6), np.float32))
    big = np.clip(big * 0.7 + rng.integers(0, 77, big.shape), 0, 255).astype(np.uint8)
    img = np.ascontiguousarray(big[:720, :1280])
    xy, r, n = ctx.detect_fast(img, 30, True)
    k = np.concatenate([xy, np.full((n, 1), 7, np.float32), np.full((n, 1), -1, np.float32)], 1)
    d, kept, ang = ctx.describe_keypoints_brisk(img, k)
    assert kept.sum() > 2000 and (ang[kept] >= 0).all() and (ang[kept] < 360).all()
    pick = rng.choice(np.nonzero(kept)[0], 150, replace=False)
    okept, oang, od = bn.describe_keypoints_brisk(img, k[pick])
    assert len(okept) == len(pick)
    np.testing.assert_array_equal(d[pick], od)
    np.testing.assert_allclose(ang[pick], oang, rtol=0, atol=1e-4)
    dk = d[kept]
    idx, dist = ctx.match_knn2_hamming(dk, dk)
    assert (dist[:, 0] == 0).all() and ((idx[:, 0] == np.arange(len(dk))) | (dist[:, 1] == 0)).all()
    dx, dy = 11, 4
    img2 = np.ascontiguousarray(big[dy:dy + 720, dx:dx + 1280])
    inner = kept & (xy[:, 0] > 20 + dx) & (xy[:, 0] < 1280 - 20) & (xy[:, 1] > 20 + dy) & (xy[:, 1] < 720 - 20)
    k2 = k[inner].copy()
    k2[:, 0] -= dx
    k2[:, 1] -= dy
    d2, kept2, ang2 = ctx.describe_keypoints_brisk(img2, k2)
    assert inner.sum() > 1500 and kept2.all()
    ham = np.unpackbits(d2 ^ d[inner], axis=1).sum(1)
    print("translated frame: rows identical", (ham == 0).mean(), "mean Hamming", ham.mean(), "max", ham.max())
    assert (ham == 0).mean() > 0.9 and ham.mean() < 1.0
    idx2, dist2 = ctx.match_knn2_hamming(d2, d[inner])
    assert (idx2[:, 0] == np.arange(len(d2))).mean() > 0.99          # every translated keypoint finds itself


# ------------------------------------------------------------------ K13: ORB descriptors at frame keypoints
def test_describe_keypoints_orb_golden_vectors_and_oracle(ctx):
    """descriptor_extractor_->compute (descriptorsmatcher.cpp:114-115, ExtractorType ORB): the committed outputs of
    cv2.ORB_create().compute -- survivors identical, rows identical up to a blurred value at a float rounding boundary
    -- and the numpy restatement, which evaluates the blur in the kernel's order: identical."""
    import os
    from oracle import orb_np as on
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    g = np.load(os.path.join(gold, "orb_keypoints.npz"))
    imgs = np.load(os.path.join(gold, "fast_keypoints.npz"))
    seen = 0
    for key in g.files:
        if not key.startswith("k_"):
            continue
        name, tag = key.split("_")[1:3]
        d, kept = ctx.describe_keypoints_orb(imgs[f"img_{name}"], g[key])
        assert d.dtype == np.uint8 and d.shape == (len(g[key]), 32)
        np.testing.assert_array_equal(np.nonzero(kept)[0], g[f"kept_{name}_{tag}"])
        assert not d[~kept].any()
        okept, od = on.describe_keypoints_orb(imgs[f"img_{name}"], g[key])
        np.testing.assert_array_equal(d[kept], od)
        if kept.any():
            ham = np.unpackbits(d[kept] ^ g[f"d_{name}_{tag}"], axis=1).sum(1)
            assert (ham == 0).mean() >= 0.995 and ham.max() <= 2
        seen += int(kept.sum())
    assert seen > 600
    rng = np.random.default_rng(8)
    img = rng.integers(0, 256, (150, 333)).astype(np.uint8)
    k = np.stack([rng.uniform(-5, 340, 300), rng.uniform(-5, 155, 300), rng.uniform(1, 40, 300), rng.uniform(-1, 360, 300)], 1).astype(np.float32)
    k[:5, 0] = [30.4, 30.5, 30.6, 301.49, 301.51]       # the border rule sees the ROUNDED position: 31 <= round(x) < w - 31
    k[:5, 1] = 70
    k[5, 0] = np.nan
    d, kept = ctx.describe_keypoints_orb(img, k)
    okept, od = on.describe_keypoints_orb(img, k)
    np.testing.assert_array_equal(np.nonzero(kept)[0], okept)
    np.testing.assert_array_equal(d[kept], od)
    assert kept[:6].tolist() == [False, False, True, True, False, False]
    big = np.zeros((150, 350), np.uint8)
    big[:, 9:9 + 333] = img
    assert np.array_equal(ctx.describe_keypoints_orb(big[:, 9:9 + 333], k)[0], d)          # row pitch
    assert np.array_equal(ctx.describe_keypoints_orb(img, k[100:101])[0][0], d[100])      # batch independence
    assert ctx.describe_keypoints_orb(img, np.zeros((0, 4), np.float32))[0].shape == (0, 32)


def test_describe_patches_orb(ctx, api):
    """extractDescriptorsFromPatches (descriptorsmatcher.cpp:133-174) with ExtractorType ORB: the committed outputs of
    cv2.ORB_create().compute on one keypoint per patch (centre, size S, angle -1), and the restatement on random patches.
    The patches are described as one stacked image: rows must not depend on their neighbours."""
    import os
    from oracle import orb_np as on
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "orb_keypoints.npz"))
    for S in (128, 64, 63):
        d = ctx.describe_patches_orb(g[f"p{S}"])
        ham = np.unpackbits(d ^ g[f"dp{S}"], axis=1).sum(1)
        assert d.shape == (len(g[f"p{S}"]), 32) and ham.max() <= 2 and (ham == 0).mean() >= 0.8, ham
    rng = np.random.default_rng(77)
    for S, n in ((128, 37), (100, 9), (63, 5)):
        patches = rng.integers(0, 256, (n, S, S), dtype=np.uint8)
        got = ctx.describe_patches_orb(patches)
        want = np.stack([on.describe_keypoints_orb(p, np.array([[S // 2, S // 2, S, -1]], np.float32))[1][0] for p in patches])
        np.testing.assert_array_equal(got, want)
        assert np.array_equal(ctx.describe_patches_orb(patches[3:4])[0], got[3])        # independent of the neighbours in the stack
    assert ctx.describe_patches_orb(np.zeros((0, 128, 128), np.uint8)).shape == (0, 32)
    for S in (62, 40):                                  # cv::ORB would remove the keypoint
        with pytest.raises(api.Fm3dError):
            ctx.describe_patches_orb(np.zeros((2, S, S), np.uint8))


def test_detect_orb_hamming_chain_full_size(ctx):
    """DetectorType FAST + ExtractorType ORB on a 1280 x 720 frame (BASELINE C2 shape; ORB-256 is the binary descriptor the
    matching sweep is quoted on): detection, description and the Hamming matcher on the GPU; oracle parity on a sample;
    a frame matched against itself returns every keypoint at distance 0; an integer translation of the frame leaves the rows
    of the translated keypoints unchanged (integer positions: the same pixels are compared)."""
    from oracle import orb_np as on
    rng = np.random.default_rng(13)
    small = rng.integers(0, 256, (720 // 6 + 4, 1280 // 6 + 4)).astype(np.float32)
    big = np.kron(small, np.ones((6, 6), np.float32))
    big = np.clip(big * 0.7 + rng.integers(0, 77, big.shape), 0, 255).astype(np.uint8)
    img = np.ascontiguousarray(big[:720, :1280])
    xy, r, n = ctx.detect_fast(img, 30, True)
    k = np.concatenate([xy, np.full((n, 1), 7, np.float32), np.full((n, 1), -1, np.float32)], 1)
    d, kept = ctx.describe_keypoints_orb(img, k)
    assert kept.sum() > 2000
    pick = rng.choice(np.nonzero(kept)[0], 300, replace=False)
    okept, od = on.describe_keypoints_orb(img, k[pick])
    assert len(okept) == len(pick)
    np.testing.assert_array_equal(d[pick], od)
    dk = d[kept]
    idx, dist = ctx.match_knn2_hamming(dk, dk)
    assert (dist[:, 0] == 0).all() and ((idx[:, 0] == np.arange(len(dk))) | (dist[:, 1] == 0)).all()
    dx, dy = 11, 4
    img2 = np.ascontiguousarray(big[dy:dy + 720, dx:dx + 1280])
    inner = kept & (xy[:, 0] > 40 + dx) & (xy[:, 0] < 1280 - 40) & (xy[:, 1] > 40 + dy) & (xy[:, 1] < 720 - 40)
    k2 = k[inner].copy()
    k2[:, 0] -= dx
    k2[:, 1] -= dy
    d2, kept2 = ctx.describe_keypoints_orb(img2, k2)
    assert inner.sum() > 1500 and kept2.all() and np.array_equal(d2, d[inner])


@pytest.mark.parametrize("shape,levels", [((480, 640), 3), ((477, 635), 5), ((67, 131), 6), ((5, 9), 2), ((1, 7), 3), ((720, 1280), 3),
                                          ((2160, 3840), 4), ((130, 65), 1), ((64, 64), 0)])
def test_pyramid_fused_kernel_bit_exact(ctx, shape, levels):
    """K4 as ONE launch per three levels (pyramid_fused = 1, the default): every level of both frames bit-identical with the
    cv2.pyrDown chain and with the one-launch-per-level kernel, for frames in device memory with a row stride larger than
    the width (fm3d_set_images_dev, where the kernel also writes level 0) and for host frames."""
    import cv2
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(shape[0] * 7 + shape[1])
    h, w = shape
    imgs = [rng.integers(0, 256, (h, w), dtype=np.uint8) for _ in range(2)]
    refs = []
    for im in imgs:
        chain = [im]
        for _ in range(levels):
            chain.append(cv2.pyrDown(chain[-1]))
        refs.append(chain)
    stride = w + 13
    dev = torch.device("cuda", 0)
    d = [torch.zeros((h, stride), dtype=torch.uint8, device=dev) for _ in range(2)]
    for k in range(2):
        d[k][:, :w] = torch.from_numpy(imgs[k]).to(dev)
    torch.cuda.synchronize()
    for fused in (1, 0):
        ctx.set_option("pyramid_fused", fused)
        try:
            for mode in ("dev", "host"):
                if mode == "dev":
                    ctx.set_images_dev(d[0].data_ptr(), d[1].data_ptr(), w, h, stride, levels)
                    ctx.sync()
                else:
                    ctx.set_images(imgs[0], imgs[1], levels)
                for k in range(2):
                    for l in range(levels + 1):
                        got = ctx.get_pyramid_level(k + 1, l)
                        np.testing.assert_array_equal(got, refs[k][l], err_msg=f"fused={fused} {mode} image {k + 1} level {l}")
        finally:
            ctx.set_option("pyramid_fused", 1)


def test_two_slot_kernel_agrees_with_the_four_group_kernel(ctx, api):
    """normals_pingpong = 1 (default with four windows per CTA): two features per eight warps take turns, the LM step of one
    runs while the seven other warps evaluate the partner's pass.  Same problem, same optimiser: statuses identical to the
    four-group kernel, normals within the parity bar of each other (the pixels are summed over seven warps instead of
    four), results reproducible run to run, for feature counts that leave slots and whole super-groups empty, and for
    features without pixels / dropped features in between."""
    case = stereo_case(640, 480, 90, 1001, 32)
    setup_ctx(ctx, case, 2)
    xyz = case["X"].copy()
    xyz[5] *= np.array([1.0, 1.0, 40.0])        # projects elsewhere, far beyond zmax: dropped by the bounding-box gate
    xyz[17] = np.array([50.0, 0.0, 1.0])        # outside the image: no pixels
    try:
        ctx.set_option("normals_groups", 4)
        for n in (1, 2, 3, 7, 90):
            res = {}
            for pp in (0, 1, 1):
                ctx.set_option("normals_pingpong", pp)
                res.setdefault(pp, []).append(ctx.optimize_normals(xyz[:n], 32, 1e-10, 1))
            a, b, b2 = res[0][0], res[1][0], res[1][1]
            np.testing.assert_array_equal(a["status"], b["status"])
            np.testing.assert_array_equal(b["status"], b2["status"])
            np.testing.assert_array_equal(b["normals"], b2["normals"])          # bit-reproducible
            np.testing.assert_array_equal(b["nfev"], b2["nfev"])
            ok = a["status"] == api.FEAT_OK
            if n == 90:
                assert ok.sum() >= 60 and (a["status"][[5, 17]] != api.FEAT_OK).all()
            ang = angle_deg(a["normals"][ok], b["normals"][ok])
            assert (ang <= 0.05).all(), ang.max()
            np.testing.assert_allclose(b["cost"][ok], a["cost"][ok], rtol=1e-4)     # the flat bottom of the cost: end points differ in the last digits
            np.testing.assert_array_equal(a["normals"][~ok], b["normals"][~ok])  # dropped features keep the viewing ray
    finally:
        ctx.set_option("normals_groups", 0)
        ctx.set_option("normals_pingpong", 1)
