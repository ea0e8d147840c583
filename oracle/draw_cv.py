"""ORACLE (test infrastructure only): the two drawing helpers behind main.cpp's output artefacts, restated on top of the very
OpenCV calls the reference makes (cv2 4.13 here; the reference linked 2.4 -- parity of the rasterisers across those
versions is unpinned, the integer midpoint circle and Bresenham line are what both document).

  random_color               tools.cpp:116-120   (cv::RNG multiply-with-carry, CV_RGB)
  drawMatches                tools.cpp:146-186
  drawBackProjectedPoints    tools.cpp:188-218   (vector variant, the one main.cpp:190-194 calls)

Only tests/ may import this module."""
import cv2
import numpy as np


class RNG:
    """cv::RNG (operations.hpp): state = (uint32)state * 4164903690 + (state >> 32); next() = (uint32)state."""

    def __init__(self, state):
        self.state = state if state else 0xFFFFFFFF

    def next(self):
        self.state = ((self.state & 0xFFFFFFFF) * 4164903690 + (self.state >> 32)) & 0xFFFFFFFFFFFFFFFF
        return self.state & 0xFFFFFFFF


def random_color(rng):
    c = rng.next()
    if c >= 1 << 31:            # int color = rng.next(): the shifts below are arithmetic
        c -= 1 << 32
    r, g, b = c & 255, (c >> 8) & 255, (c >> 16) & 255
    return (float(b), float(g), float(r), 0.0)     # CV_RGB(r, g, b) = Scalar(b, g, r, 0)


def _round_half_even(v):
    return int(np.rint(np.float64(v)))             # cv::saturate_cast<int>(float) = cvRound


def draw_matches(img1, img2, kp1, kp2, matches, inlier_mask):
    """-> (window h x 2w x 3 BGR, colours).  kp*: (n, 2) float32 pixel coordinates; matches: (m, 2) (queryIdx, trainIdx)."""
    h, w = img1.shape
    window = np.zeros((h, 2 * w, 3), np.uint8)
    window[:, :w] = cv2.cvtColor(img1, cv2.COLOR_GRAY2BGR)
    window[:img2.shape[0], w:w + img2.shape[1]] = cv2.cvtColor(img2, cv2.COLOR_GRAY2BGR)
    rng = RNG(0xFFF0FF0F)
    colours = []
    for (q, t), ok in zip(matches, inlier_mask):
        if not ok:
            continue
        col = random_color(rng)
        colours.append(col)
        x1, y1 = np.float32(kp1[q][0]), np.float32(kp1[q][1])
        x2, y2 = np.float32(np.float32(kp2[t][0]) + np.float32(w)), np.float32(kp2[t][1])     # pt2.x = pt2.x + img1.cols, in float
        p1 = (_round_half_even(x1), _round_half_even(y1))
        p2 = (_round_half_even(x2), _round_half_even(y2))
        cv2.circle(window, p1, 4, col)
        cv2.circle(window, p2, 4, col)
        cv2.line(window, p1, p2, col)
    return window, colours


def draw_back_projected_points(img, points, colours):
    """points: (n_patches, n_points, 2) image points (x, y); colours: per patch (B, G, R).  Points that round onto column
    `cols` or row `rows` are skipped (the reference's `!(x > cols)` test lets them write past the row: deviation, as D2)."""
    out = cv2.cvtColor(img, cv2.COLOR_GRAY2BGR)
    h, w = img.shape
    for pts, col in zip(points, colours):
        x = np.where(pts[:, 0] >= 0, np.floor(pts[:, 0] + 0.5), np.ceil(pts[:, 0] - 0.5)).astype(np.int64)   # C round(): half away from zero
        y = np.where(pts[:, 1] >= 0, np.floor(pts[:, 1] + 0.5), np.ceil(pts[:, 1] - 0.5)).astype(np.int64)
        ok = (x >= 0) & (x < w) & (y >= 0) & (y < h)
        out[y[ok], x[ok]] = np.asarray(col[:3], np.float64).astype(np.uint8)
    return out
