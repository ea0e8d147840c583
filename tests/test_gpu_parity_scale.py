"""Oracle parity of the plane-normal search where round 1 had none ("What's weak" 1 and 3):

* `penalty_mode = fabs` -- the reference as today's g++ compiles it (normaloptimizer.cpp:126-142):
  every feature whose optimum lies across theta = pi/2 ends on the penalty wall, in the oracle too.  The faithful
  kernel (`normals_fast = 0`) must reproduce the oracle's end state ON the wall; the default fast kernel must
  reproduce every status and its agreement rate on the wall is measured and bounded.
* the sizes that are benched: 1280x720 (BASELINE configs[1]) and 3840x2160 (configs[2]) at pixelsRay 64,
  pyramids 3, every group layout of the fast kernel, on a seeded 256-feature sample against the plain-C oracle
  (`normaloptimizer.cpp:321-452` restated in oracle/fm3d_oracle.c, pinned in tests/test_oracle_pins.py).

Feature identity is the index in the input array, never the position in a compacted output.
"""
import os

import numpy as np
import pytest

from common import angle_deg, cam_tuple, car2sph, orc, setup_ctx, stereo_case

pytestmark = pytest.mark.gpu

THREADS = os.cpu_count() or 8


def _oracle(case, pyramids, xyz, r, penalty_mode):
    cam = case["scene"].cam
    return orc.optimize_normals(*cam_tuple(cam), case["scene"].img1, case["scene"].img2, pyramids, xyz, r, 1e-10,
                                penalty_mode=penalty_mode, threads=THREADS)


def _sample(case, n, seed):
    rng = np.random.default_rng(seed)
    X = case["X"]
    sel = np.sort(rng.choice(X.shape[0], min(n, X.shape[0]), replace=False))
    return np.ascontiguousarray(X[sel]), sel


# ------------------------------------------------------------------ fabs mode (the source as compiled today)
def test_fabs_mode_faithful_kernel_reproduces_the_oracle_on_the_wall(ctx):
    """`normals_fast = 0`, penalty wall with fabs semantics: statuses identical, and the end state of the features that
    touch the wall (oracle npenalty > 0) within 0.5 deg of the oracle's on >= 95 % of them; interior features all."""
    case = stereo_case(640, 480, 160, 1003, 32)
    setup_ctx(ctx, case, 2)
    xyz = case["X"]
    ctx.set_option("normals_fast", 0)
    try:
        res = ctx.optimize_normals(xyz, 32, 1e-10, 0)
    finally:
        ctx.set_option("normals_fast", 1)
    o = _oracle(case, 2, xyz, 32, 0)
    np.testing.assert_array_equal(res["status"], o["status"])
    ok = o["status"] == 0
    wall = ok & (o["npenalty"] > 0)
    interior = ok & (o["npenalty"] == 0)
    ang = angle_deg(res["normals"], o["normals"])
    assert wall.sum() >= 40, "the scene must put features on the wall for this test to mean anything"
    rate = float((ang[wall] <= 0.5).mean())
    print(f"fabs/faithful: {int(wall.sum())} wall features, agreement {rate:.3f}, median {np.median(ang[wall]):.4f} deg, "
          f"{int(interior.sum())} interior, max interior {ang[interior].max() if interior.any() else 0:.4f}")
    assert rate >= 0.95, rate
    assert (ang[interior] <= 0.5).all()
    # the wall is where the gpu run spent its penalty evaluations too
    assert ((res["npenalty"] > 0) == (o["npenalty"] > 0))[ok].mean() >= 0.95


@pytest.mark.parametrize("groups", [1, 4])
def test_fabs_mode_fast_kernel_statuses_and_wall_agreement(ctx, groups):
    """The default fast kernel under fabs semantics: statuses identical to the oracle's, interior features within
    0.5 deg, and the wall features end where the oracle's do (<= 0.5 deg) at a rate that is measured here and published
    by bench.py as `penalty_fabs.oracle_agreement`; it must not fall below 0.5, and every wall feature must still be ON
    the wall region the oracle found (within 5 deg), i.e. the kernel never escapes a wall the reference is stuck on."""
    case = stereo_case(640, 480, 160, 1003, 32)
    setup_ctx(ctx, case, 2)
    xyz = case["X"]
    ctx.set_option("normals_groups", groups)
    try:
        res = ctx.optimize_normals(xyz, 32, 1e-10, 0)
    finally:
        ctx.set_option("normals_groups", 0)
    o = _oracle(case, 2, xyz, 32, 0)
    np.testing.assert_array_equal(res["status"], o["status"])
    ok = o["status"] == 0
    wall = ok & (o["npenalty"] > 0)
    interior = ok & (o["npenalty"] == 0)
    ang = angle_deg(res["normals"], o["normals"])
    assert (ang[interior] <= 0.5).all()
    rate = float((ang[wall] <= 0.5).mean()) if wall.any() else 1.0
    print(f"fabs/fast groups={groups}: {int(wall.sum())} wall features, agreement {rate:.3f}, "
          f"p50 {np.median(ang[wall]):.4f} p95 {np.percentile(ang[wall], 95):.4f} max {ang[wall].max():.4f} deg")
    assert rate >= 0.5, rate
    assert np.percentile(ang[wall], 95) <= 5.0


# ------------------------------------------------------------------ the benched sizes
def _oracle_cost_at(case, pyramids, xyz, normals, r):
    """The reference's own cost function (evaluateNormal, level 0, no penalty weight) at given normals, by the oracle."""
    cam = case["scene"].cam
    return orc.evaluate_cost(*cam_tuple(cam), case["scene"].img1, case["scene"].img2, pyramids, xyz, car2sph(normals), r, 0, 2)[0]


def _parity_at_size(ctx, case, pyramids, r, n_sample, penalty_mode, layouts, allow_better_minimum=False, gt_slack=0.1):
    xyz, sel = _sample(case, n_sample, 77)
    o = _oracle(case, pyramids, xyz, r, penalty_mode)
    ok = o["status"] == 0
    interior = ok & (o["npenalty"] == 0)
    assert ok.sum() >= 0.9 * len(sel)
    o_gt = angle_deg(o["normals"], case["normal"][sel])
    out = {}
    for name, fast, groups in layouts:
        ctx.set_option("normals_fast", fast)
        ctx.set_option("normals_groups", groups)
        try:
            res = ctx.optimize_normals(xyz, r, 1e-10, penalty_mode)
        finally:
            ctx.set_option("normals_fast", 1)
            ctx.set_option("normals_groups", 0)
        np.testing.assert_array_equal(res["status"], o["status"], err_msg=name)
        ang = angle_deg(res["normals"], o["normals"])
        gt = angle_deg(res["normals"], case["normal"][sel])
        print(f"{name}: {int(interior.sum())}/{len(sel)} interior, vs oracle p50 {np.median(ang[interior]):.5f} "
              f"max {ang[interior].max():.5f} deg; vs truth p50 {np.median(gt[interior]):.4f} (oracle {np.median(o_gt[interior]):.4f})")
        far = interior & (ang > 0.5)
        if allow_better_minimum and far.any():
            # Where the two end states differ by more than the bar, the kernel must have found a strictly LOWER value of the
            # reference's own cost function (evaluated by the oracle): at 4K the reference's forward-difference Jacobian
            # (step 1e-5 rad) drowns in the float cast of the sampling coordinates at the coarse levels, lmfit runs into its
            # 300-evaluation cap there and stops away from the minimum.  Never the other way round.
            gc = _oracle_cost_at(case, pyramids, xyz[far], res["normals"][far], r)
            oc = _oracle_cost_at(case, pyramids, xyz[far], o["normals"][far], r)
            print(f"{name}: {int(far.sum())} features beyond 0.5 deg, device cost / oracle cost = {np.round(gc / oc, 3).tolist()}, "
                  f"oracle nfev at the coarsest level {o['nfev'][far][:, -1].tolist()}")
            assert (gc < oc * (1 - 1e-3)).all(), (name, gc / oc)
            assert far.sum() <= 0.03 * interior.sum(), (name, int(far.sum()))
        else:
            assert (ang[interior] <= 0.5).all(), (name, ang[interior].max())       # north-star bar
        assert np.median(ang[interior]) <= 0.02, (name, np.median(ang[interior]))
        assert (gt[interior] <= o_gt[interior] + gt_slack).all(), (name, (gt[interior] - o_gt[interior]).max())
        near = interior & (ang <= 0.5)
        rel = np.abs(res["cost"][near] - o["cost"][near]) / np.maximum(o["cost"][near], 1e-30)
        assert np.median(rel) <= 0.01, (name, np.median(rel))
        out[name] = res
    return out, o, interior


LAYOUTS = [("fast-1group", 1, 1), ("fast-2groups", 1, 2), ("fast-4groups", 1, 4), ("fast-auto", 1, 0), ("faithful", 0, 0)]


@pytest.mark.parametrize("penalty_mode", [1, 2])
def test_normals_oracle_parity_at_the_benched_size_720p(ctx, penalty_mode):
    """BASELINE configs[1] (what bench.py times): 1280x720, pixelsRay 64, pyramids 3; a seeded 256-feature sample of the
    bench scene (seed 1001), all group layouts + the faithful kernel, against the oracle in the bench's penalty mode
    (int_abs) and with the wall off (every feature interior)."""
    case = stereo_case(1280, 720, 2000, 1001, 64)
    setup_ctx(ctx, case, 3)
    out, o, interior = _parity_at_size(ctx, case, 3, 64, 256, penalty_mode, LAYOUTS)
    if penalty_mode == 2:
        assert interior.sum() == (o["status"] == 0).sum()
    assert interior.sum() >= 128


def test_normals_oracle_parity_at_4k(ctx):
    """BASELINE configs[2]: 3840x2160 (pixel coordinates up to 3840: the coarsest fp32 quantum the fast kernel's
    geometry meets), pixelsRay 64, pyramids 3, seeded 256-feature sample, wall off and int_abs.

    At this size the ORACLE (= the reference's algorithm) stops short of the minimum on some features: its forward-difference
    step of 1e-5 rad moves a coarse-level sample by less than the fp32 quantum of the float-cast coordinate, the Jacobian is
    noise, lmfit burns its 300 evaluations at level 3 and hands a poor start to the finer levels (measured: 10 % of the
    features hit the cap, 12 % end > 0.1 deg from the ground truth, worst 0.66 deg, with a final cost up to 4x the cost at the
    true normal).  The fast kernel's analytic Jacobian does not have that problem (worst 0.05 deg from the truth).  The gate
    is therefore: within 0.5 deg of the oracle, OR a strictly lower value of the reference's cost function, the latter on at
    most 3 % of the features; the faithful kernel (forward differences like the reference) must stay within 0.5 deg."""
    case = stereo_case(3840, 2160, 600, 1002, 64)
    setup_ctx(ctx, case, 3)
    for penalty_mode in (2, 1):
        out, o, interior = _parity_at_size(ctx, case, 3, 64, 256, penalty_mode, [("fast-4groups", 1, 4), ("fast-1group", 1, 1)],
                                           allow_better_minimum=True)
        # both forward-difference implementations are in the noise regime here: their distance to the truth may differ by 0.2 deg
        _parity_at_size(ctx, case, 3, 64, 256, penalty_mode, [("faithful", 0, 0)], gt_slack=0.2)
        assert interior.sum() >= 128
