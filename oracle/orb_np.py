"""TEST INFRASTRUCTURE -- CPU oracle for K13 (fm3d_describe_keypoints_orb), numpy only.

Restates what descriptor_extractor_->compute(frame, keypoints, descriptors) computes in
DescriptorsMatcher::compareWithNNDR / compare / crosscompare
(DescriptorsMatcher/descriptorsmatcher.cpp:114-115, :95-96, :80-81) when ExtractorType is ORB (:336-342:
cv::ORB(OrbDetector.NumFeatures, ScaleFactor, NumLevels); the three knobs steer ORB's own detector and pyramid, which a
provided keypoint of octave 0 does not touch).  OpenCV is a third-party dependency of the reference (unpinned, 2.4.x
era); the published algorithm restated here is cv::ORB's descriptor stage (Rublee et al., ICCV 2011;
modules/features2d/src/orb.cpp, OpenCV 4.13 as installed in this image):

  * keypoints whose ROUNDED position is closer than edgeThreshold = 31 pixels to the border are REMOVED
    (KeyPointsFilter::runByImageBorder: Rect(31, 31, w - 62, h - 62).contains(Point(pt)));
  * the level image is blurred in place, GaussianBlur 7 x 7, sigma 2, BORDER_REFLECT_101 -- observed through the bits:
    the float convolution rounded to u8;
  * bit k = [B(c + R a_k) < B(c + R b_k)], c = (cvRound(x), cvRound(y)), R = rotation by KeyPoint::angle (degrees;
    FAST's -1 is a rotation of -1 degree, which moves no pattern point to another pixel), coordinates cvRound-ed;
    (a_k, b_k) = the 256 pairs of ORB's learned pattern, recovered from cv2 by tools/recover_orb_pattern.py.

Pinned by: cv2.ORB_create().compute itself (tests/test_oracle_pins.py, where cv2 is importable) and the committed golden
vectors tests/golden/orb_keypoints.npz written from cv2 by tools/make_golden.py.  Integer comparisons on a rounded float
blur: rows are identical except where a blurred value sits within float rounding of a half-integer (about 1 pixel in
10^5; the summation order of OpenCV's SIMD blur is not published).  Only tests/ may import this module.
"""
from __future__ import annotations

import math
import os

import numpy as np

f32 = np.float32
EDGE = 31
PATTERN = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "orb_pattern.npy")).astype(np.int64)   # dy_a, dx_a, dy_b, dx_b


def gaussian_kernel7() -> np.ndarray:
    """cv::getGaussianKernel(7, 2, CV_32F)."""
    x = np.arange(7) - 3.0
    t = np.exp(-0.5 / 4.0 * x * x)
    return (t / t.sum()).astype(np.float32)


def _fma(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


def orb_blur(img: np.ndarray) -> np.ndarray:
    """7 x 7 sigma-2 float blur (symmetric form with fused multiply-adds, as K13 evaluates it), rounded to u8."""
    k = gaussian_kernel7()
    rows, cols = img.shape

    def reflect(n):
        idx = np.abs(np.arange(-3, n + 3))
        return np.where(idx >= n, 2 * (n - 1) - idx, idx)

    def one(p, axis):
        n = p.shape[axis] - 6
        sl = (lambda q: (slice(None), slice(q, q + n))) if axis == 1 else (lambda q: (slice(q, q + n), slice(None)))
        s = (k[3] * p[sl(3)]).astype(np.float32)
        for i in range(1, 4):
            s = _fma(np.broadcast_to(k[3 + i], s.shape), (p[sl(3 - i)] + p[sl(3 + i)]).astype(np.float32), s)
        return s

    p = np.asarray(img, np.uint8).astype(np.float32)[:, reflect(cols)]
    r = one(p, 1)
    out = one(r[reflect(rows), :], 0)
    return np.clip(np.rint(out), 0, 255).astype(np.uint8)


def describe_keypoints_orb(img: np.ndarray, kps: np.ndarray):
    """img: h x w u8; kps: n x 4 float32 (x, y, size, angle), octave 0.  Returns (kept indices, n_kept x 32 u8)."""
    img = np.asarray(img, np.uint8)
    h, w = img.shape
    kps = np.asarray(kps, np.float32).reshape(-1, 4)
    B = orb_blur(img).astype(np.int32)
    kept, rows = [], []
    pa = PATTERN[:, [1, 0]].astype(np.float32)       # x, y of the first point
    pb = PATTERN[:, [3, 2]].astype(np.float32)
    for k, (x, y, size, angle) in enumerate(kps):
        if not (np.isfinite(x) and np.isfinite(y) and np.isfinite(angle)):
            continue
        cx, cy = int(np.rint(np.float64(x))), int(np.rint(np.float64(y)))     # cvRound; Rect::contains sees the rounded point
        if cx < EDGE or cx >= w - EDGE or cy < EDGE or cy >= h - EDGE:
            continue
        ang = f32(f32(angle) * f32(math.pi / 180.0))
        a, b = f32(math.cos(float(ang))), f32(math.sin(float(ang)))

        def val(p):
            xx = (p[:, 0] * a).astype(np.float32) - (p[:, 1] * b).astype(np.float32)
            yy = (p[:, 0] * b).astype(np.float32) + (p[:, 1] * a).astype(np.float32)
            ix = np.rint(xx.astype(np.float64)).astype(np.int64)
            iy = np.rint(yy.astype(np.float64)).astype(np.int64)
            return B[cy + iy, cx + ix]

        bits = (val(pa) < val(pb)).astype(np.uint8)
        kept.append(k)
        rows.append(np.packbits(bits, bitorder="little"))
    return np.array(kept, np.int64), np.array(rows, np.uint8).reshape(-1, 32)
