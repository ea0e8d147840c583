"""Pins of the host-side plumbing behind main.cpp's output artefacts (matches.pgm, projectedPatches.pgm):

* the integer rasterisers of fm3d_cv.h (cv::circle / cv::line, thickness 1, 8-connected) against cv2.circle / cv2.line,
  including primitives that leave the image;
* the PxM writer / reader against cv2.imwrite / cv2.imread (a CV_8UC3 image written to "*.pgm" becomes a binary PPM with
  RGB channel order; reading it back as grayscale uses cv's fixed-point weights);
* drawMatches / drawBackProjectedPoints (host/fm3d_draw.cpp) against oracle/draw_cv.py, the restatement of
  tools.cpp:116-120,146-239 on top of cv2 calls.

CPU only."""
import ctypes as C
import os

import cv2
import numpy as np
import pytest

from oracle import draw_cv

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "tests", "_build", "libdraw_harness.so")


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(LIB):
        import importlib
        importlib.import_module("__graft_entry__").build()
    return C.CDLL(LIB)


def _p(a, t=C.c_void_p):
    return a.ctypes.data_as(t)


def test_circle_and_line_match_cv2(lib):
    rng = np.random.default_rng(5)
    w, h = 97, 61
    for trial in range(40):
        base = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        n = 60
        prims = np.zeros((n, 6), np.int32)
        prims[:, 0] = rng.integers(0, 2, n)
        prims[:, 1] = rng.integers(-30, w + 30, n)
        prims[:, 2] = rng.integers(-30, h + 30, n)
        prims[:, 3] = np.where(prims[:, 0] == 0, rng.integers(0, 25, n), rng.integers(-30, w + 30, n))
        prims[:, 4] = rng.integers(-30, h + 30, n)
        prims[:, 5] = np.arange(n)
        cols = rng.integers(0, 256, (n, 3)).astype(np.float64)
        ours = base.copy()
        lib.fm3d_test_draw(_p(ours), w, h, _p(prims), n, _p(cols))
        ref = base.copy()
        for p, c in zip(prims, cols):
            col = tuple(float(v) for v in c)
            if p[0] == 0:
                cv2.circle(ref, (int(p[1]), int(p[2])), int(p[3]), col)
            else:
                cv2.line(ref, (int(p[1]), int(p[2])), (int(p[3]), int(p[4])), col)
        np.testing.assert_array_equal(ours, ref, err_msg=f"trial {trial}")


def test_line_directions_and_long_lines_match_cv2(lib):
    w, h = 64, 64
    ends = [(-200, 10), (10, -300), (400, 33), (33, 900), (0, 0), (63, 63), (63, 0), (0, 63), (31, 31), (5, 50)]
    for a in ends:
        for b in ends:
            ours = np.zeros((h, w, 3), np.uint8)
            prims = np.array([[1, a[0], a[1], b[0], b[1], 0]], np.int32)
            cols = np.array([[255.0, 128.0, 7.0]])
            lib.fm3d_test_draw(_p(ours), w, h, _p(prims), 1, _p(cols))
            ref = np.zeros((h, w, 3), np.uint8)
            cv2.line(ref, a, b, (255.0, 128.0, 7.0))
            np.testing.assert_array_equal(ours, ref, err_msg=f"{a} -> {b}")


def test_pxm_writer_and_reader_match_cv2(lib, tmp_path):
    rng = np.random.default_rng(6)
    for ch in (1, 3):
        img = rng.integers(0, 256, (37, 53, ch), dtype=np.uint8)
        # OpenCV 2.4's PxM encoder chose P5 / P6 by the channel count whatever the extension (main.cpp:141 relies on that);
        # cv2 4.13 insists on ".ppm" for three channels -- same bytes
        a, b = str(tmp_path / f"ours{ch}.pgm"), str(tmp_path / (f"cv{ch}.ppm" if ch == 3 else f"cv{ch}.pgm"))
        assert lib.fm3d_test_imwrite(a.encode(), _p(img), 53, 37, ch) == 0
        assert cv2.imwrite(b, img if ch == 3 else img[:, :, 0])
        assert open(a, "rb").read() == open(b, "rb").read()
        out = np.zeros(37 * 53, np.uint8)
        ww, hh = C.c_int(0), C.c_int(0)
        assert lib.fm3d_test_imread_gray(b.encode(), _p(out), out.size, C.byref(ww), C.byref(hh)) == 0
        np.testing.assert_array_equal(out.reshape(37, 53), cv2.imread(b, cv2.IMREAD_GRAYSCALE))


def test_draw_matches_and_back_projected_points_match_the_oracle(lib):
    rng = np.random.default_rng(7)
    w, h = 160, 120
    img1 = rng.integers(0, 256, (h, w), dtype=np.uint8)
    img2 = rng.integers(0, 256, (h, w), dtype=np.uint8)
    n1, n2, nm = 50, 60, 40
    kp1 = (rng.random((n1, 2)) * [w - 1, h - 1]).astype(np.float32)
    kp2 = (rng.random((n2, 2)) * [w - 1, h - 1]).astype(np.float32)
    kp1[0] = (12.5, 7.5)            # ties of the float -> int conversion: round half to even
    kp2[0] = (13.5, 8.5)
    matches = np.stack([rng.integers(0, n1, nm), rng.integers(0, n2, nm)], 1).astype(np.int32)
    matches[0] = (0, 0)
    mask = (rng.random(nm) < 0.7).astype(np.uint8)
    mask[0] = 1
    window = np.zeros((h, 2 * w, 3), np.uint8)
    colours = np.zeros((nm, 3))
    nc = lib.fm3d_test_draw_matches(_p(img1), _p(img2), w, h, _p(kp1), n1, _p(kp2), n2, _p(matches), _p(mask), nm, _p(window), _p(colours))
    o_window, o_colours = draw_cv.draw_matches(img1, img2, kp1, kp2, matches, mask.astype(bool))
    assert nc == int(mask.sum()) == len(o_colours)
    np.testing.assert_array_equal(colours[:nc], np.asarray(o_colours)[:, :3])
    np.testing.assert_array_equal(window, o_window)
    # back-projected patch points, some outside the frame, one on x == cols (skipped: documented deviation)
    npatch, npts = 5, 64
    pts = rng.random((npatch, npts, 2)) * [w + 20, h + 20] - 10
    pts[0, 0] = (w, 5.0)
    pts[0, 1] = (w - 0.5, h - 0.5)      # rounds to (w, h): outside
    out = np.zeros((h, w, 3), np.uint8)
    lib.fm3d_test_draw_points(_p(img1), w, h, _p(pts), npatch, npts, _p(colours), _p(out))
    np.testing.assert_array_equal(out, draw_cv.draw_back_projected_points(img1, pts, colours[:npatch]))
