"""cost_mode NCC of the normal search (option "normals_cost" = 1): the zero-mean normalised residual the north star names
next to the reference's SSD (Triangulator/normaloptimizer.cpp:145-148 has only the latter).  The oracle
restates it in oracle/fm3d_oracle.c (eval_normal, COST_NCC) on top of the same lmmin; the fast kernel gets it from twelve
sums per pass (fm3d_normals_fast.cu: ncc_sums_to_normal_equations)."""
import numpy as np
import pytest

from common import angle_deg, cam_tuple, car2sph, orc, setup_ctx, stereo_case

pytestmark = pytest.mark.gpu


@pytest.fixture
def ncc(ctx):
    ctx.set_option("normals_cost", 1)
    yield ctx
    ctx.set_option("normals_cost", 0)
    ctx.set_option("normals_fast", 1)
    ctx.set_option("normals_groups", 0)


@pytest.mark.parametrize("level", [0, 1, 2])
def test_ncc_cost_of_one_evaluation_matches_the_oracle(ncc, level):
    ctx = ncc
    case = stereo_case(640, 480, 40, 1001, 32)
    cam = case["scene"].cam
    setup_ctx(ctx, case, 2)
    xyz = case["X"][:24]
    rng = np.random.default_rng(level)
    n0 = xyz / np.linalg.norm(xyz, axis=1, keepdims=True)
    pt = car2sph(n0) + rng.normal(0, 0.15, (xyz.shape[0], 2))
    ctx.set_option("normals_fast", 2)
    cost, m, status = ctx.evaluate_normals(xyz, pt, 32, level, 2)
    o_cost, o_m, o_status = orc.evaluate_cost(*cam_tuple(cam), case["scene"].img1, case["scene"].img2, 2, xyz, pt, 32, level, 2, cost_mode=1)
    np.testing.assert_array_equal(m, o_m)
    np.testing.assert_array_equal(status, o_status)
    ok = o_status == 0
    assert ok.sum() >= 20 and (o_cost[ok] > 0).all() and (o_cost[ok] < 2).all()
    print("NCC cost: max rel diff", (np.abs(cost[ok] - o_cost[ok]) / o_cost[ok]).max())
    np.testing.assert_allclose(cost[ok], o_cost[ok], rtol=5e-3, atol=1e-7)


@pytest.mark.parametrize("groups", [1, 2, 4])
@pytest.mark.parametrize("penalty_mode", [2, 1])
def test_ncc_normal_search_matches_the_oracle(ncc, groups, penalty_mode):
    ctx = ncc
    case = stereo_case(640, 480, 60, 1001, 32)
    cam = case["scene"].cam
    setup_ctx(ctx, case, 2)
    xyz = case["X"]
    ctx.set_option("normals_groups", groups)
    res = ctx.optimize_normals(xyz, 32, 1e-10, penalty_mode)
    o = orc.optimize_normals(*cam_tuple(cam), case["scene"].img1, case["scene"].img2, 2, xyz, 32, 1e-10,
                             penalty_mode=penalty_mode, threads=8, cost_mode=1)
    np.testing.assert_array_equal(res["status"], o["status"])
    ok = (o["status"] == 0) & (o["npenalty"] == 0)
    assert ok.sum() >= 30
    ang = angle_deg(res["normals"], o["normals"])
    gt, o_gt = angle_deg(res["normals"], case["normal"]), angle_deg(o["normals"], case["normal"])
    print(f"NCC groups={groups} penalty={penalty_mode}: vs oracle p50 {np.median(ang[ok]):.5f} max {ang[ok].max():.4f}; "
          f"vs truth {np.median(gt[ok]):.4f} (oracle {np.median(o_gt[ok]):.4f}); nfev {res['nfev'].sum(0)} / {o['nfev'].sum(0)}")
    assert (ang[ok] <= 0.5).all() and np.median(ang[ok]) <= 0.02
    assert (gt[ok] <= o_gt[ok] + 0.1).all()
    np.testing.assert_allclose(res["cost"][ok], o["cost"][ok], rtol=0.02, atol=1e-8)


def test_ncc_is_invariant_to_gain_and_offset_where_ssd_is_not(ncc):
    """Image 2 with another exposure (0.7 x + 20): the SSD optimum moves away from the true plane, the NCC optimum does not."""
    ctx = ncc
    case = stereo_case(640, 480, 60, 1001, 32)
    cam = case["scene"].cam
    img2 = np.clip(np.rint(case["scene"].img2.astype(np.float64) * 0.7 + 20.0), 0, 255).astype(np.uint8)
    ctx.set_camera(cam.K, cam.dist, cam.z_min, cam.z_max)
    ctx.set_g12(cam.g12)
    ctx.set_images(case["scene"].img1, img2, 2)
    xyz = case["X"]
    res_ncc = ctx.optimize_normals(xyz, 32, 1e-10, 2)
    ctx.set_option("normals_cost", 0)
    res_ssd = ctx.optimize_normals(xyz, 32, 1e-10, 2)
    ok = (res_ncc["status"] == 0) & (res_ssd["status"] == 0)
    g_ncc, g_ssd = angle_deg(res_ncc["normals"], case["normal"])[ok], angle_deg(res_ssd["normals"], case["normal"])[ok]
    print(f"gain/offset: NCC p50 {np.median(g_ncc):.4f} max {g_ncc.max():.4f}; SSD p50 {np.median(g_ssd):.4f} max {g_ssd.max():.4f}")
    assert ok.sum() >= 40 and np.median(g_ncc) < 0.1 and g_ncc.max() < 0.5
    assert np.median(g_ssd) > 5 * np.median(g_ncc)
    o = orc.optimize_normals(*cam_tuple(cam), case["scene"].img1, img2, 2, xyz, 32, 1e-10, penalty_mode=2, threads=8, cost_mode=1)
    np.testing.assert_array_equal(res_ncc["status"], o["status"])
    assert (angle_deg(res_ncc["normals"], o["normals"])[o["status"] == 0] <= 0.5).all()


def test_ncc_flat_patch_and_unsupported_paths(ncc, api):
    ctx = ncc
    case = stereo_case(640, 480, 40, 1001, 32)
    cam = case["scene"].cam
    ctx.set_camera(cam.K, cam.dist, cam.z_min, cam.z_max)
    ctx.set_g12(cam.g12)
    flat = np.full_like(case["scene"].img1, 77)
    ctx.set_images(flat, flat, 2)
    xyz = case["X"][:8]
    res = ctx.optimize_normals(xyz, 32, 1e-10, 2)
    o = orc.optimize_normals(*cam_tuple(cam), flat, flat, 2, xyz, 32, 1e-10, penalty_mode=2, threads=4, cost_mode=1)
    np.testing.assert_array_equal(res["status"], o["status"])
    assert (res["status"] == api.FEAT_ABORT_NAN).all()          # no texture: no normalised residual
    setup_ctx(ctx, case, 2)
    ctx.set_option("normals_fast", 0)
    with pytest.raises(api.Fm3dError):
        ctx.optimize_normals(xyz, 32, 1e-10, 2)
    ctx.set_option("normals_fast", 1)
    with pytest.raises(api.Fm3dError):
        ctx.sweep_normals(xyz, 32, 0, 3, 3, 0.01, 0.01)
