"""TEST INFRASTRUCTURE -- CPU oracle for K12 (fm3d_describe_keypoints_brisk), numpy + plain Python.

Restates what descriptor_extractor_->compute(frame, keypoints, descriptors) computes in
DescriptorsMatcher::compareWithNNDR / compare / crosscompare
(DescriptorsMatcher/descriptorsmatcher.cpp:114-115, :95-96, :80-81) when ExtractorType is BRISK (:343-349:
cv::BRISK(FeatureOptions.BriskDetector.Threshold, FeatureOptions.BriskDetector.Octaves); threshold and octaves
only steer BRISK's own detector, which the reference does not use).  OpenCV is a third-party dependency of the
reference (unpinned, 2.4.x era); the published algorithm restated here is cv::BRISK's descriptor stage
(modules/features2d/src/brisk.cpp, OpenCV 4.13 as installed in this image; Leutenegger et al., ICCV 2011):

  * sampling pattern: 60 points on 5 rings (radii 0.85 * {0, 2.9, 4.9, 7.4, 10.8}, {1, 10, 14, 15, 20} points),
    Gaussian-like box smoothing of half width sigma = 1.3 * r * sin(pi / n) (0.65 at the centre), for 64 scales
    2^(k * log2(30) / 64) and 1024 rotations; 512 short pairs (distance < 5.85), 870 long pairs (> 8.2);
  * per keypoint: scale index from KeyPoint::size, keypoints closer than the pattern extent to the border are
    REMOVED (so are their rows); the 60 smoothed intensities (integer box integration with sub-pixel borders over
    the image / the integral image) at rotation 0 give the orientation from the long pairs (integer arithmetic,
    atan2 in float); the intensities at the rotated pattern give bit k = [I(i_k) > I(j_k)] of the 64-byte row.
    compute_orientation=False restates the 2.4-era behaviour for provided keypoints (no orientation step: a
    keypoint with angle -1, as FAST leaves it, is sampled unrotated).

Pinned by: cv2.BRISK_create().compute itself (tests/test_oracle_pins.py, where cv2 is importable) and the committed
golden vectors tests/golden/brisk_keypoints.npz written from cv2 by tools/make_golden.py; integer work, agreement
is exact (surviving keypoints, bits and angles).  Only tests/ may import this module.
"""
from __future__ import annotations

import math

import numpy as np

f32 = np.float32
POINTS, SCALES, N_ROT = 60, 64, 1024
SCALERANGE, BASIC_SIZE = 30.0, 12.0
R_LIST = [f32(0.85 * 0.0), f32(0.85 * 2.9), f32(0.85 * 4.9), f32(0.85 * 7.4), f32(0.85 * 10.8)]
N_LIST = [1, 10, 14, 15, 20]
D_MAX, D_MIN = f32(5.85), f32(8.2)
# scalerange_ is a float member, so std::log(scalerange_) is logf: lb_scale = (float)(logf(30.f) / log(2.0)) =
# 4.906890869140625 (the double logarithm would round to 4.90689039...; pinned through the angles cv2 returns)
LOGF_30 = f32(math.log(SCALERANGE))
LB_SCALE = f32(float(LOGF_30) / math.log(2.0))
LB_SCALE_STEP = f32(LB_SCALE / f32(SCALES))


def scale_factor(scale: int) -> np.float32:
    return f32(math.pow(2.0, float(f32(scale) * LB_SCALE_STEP)))


def pattern(scale: int, rot: int):
    """x, y, sigma (float32 arrays of 60) of the pattern at a scale / rotation index (generateKernel)."""
    s = scale_factor(scale)
    theta = float(rot) * 2 * math.pi / float(N_ROT)
    xs, ys, sg = [], [], []
    for ring in range(5):
        for num in range(N_LIST[ring]):
            alpha = float(num) * 2 * math.pi / float(N_LIST[ring])
            xs.append(f32(float(s * R_LIST[ring]) * math.cos(alpha + theta)))
            ys.append(f32(float(s * R_LIST[ring]) * math.sin(alpha + theta)))
            if ring == 0:
                sg.append(f32(f32(f32(1.3) * s) * f32(0.5)))
            else:
                sg.append(f32(float(f32(1.3) * s) * float(R_LIST[ring]) * math.sin(math.pi / N_LIST[ring])))
    return np.array(xs, np.float32), np.array(ys, np.float32), np.array(sg, np.float32)


def size_of_scale(scale: int) -> int:
    s = scale_factor(scale)
    _, _, sg = pattern(scale, 0)
    out, p = 0, 0
    for ring in range(5):
        for _ in range(N_LIST[ring]):
            out = max(out, int(math.ceil(float(f32(s * R_LIST[ring]) + sg[p]))) + 1)
            p += 1
    return out


_PAIRS = None


def pairs():
    """(short i, short j), (long i, long j, weighted_dx, weighted_dy) from the scale-0 / rotation-0 pattern."""
    global _PAIRS
    if _PAIRS is None:
        x, y, _ = pattern(0, 0)
        dmin_sq, dmax_sq = f32(D_MIN * D_MIN), f32(D_MAX * D_MAX)
        sh, lg = [], []
        for i in range(1, POINTS):
            for j in range(i):
                dx, dy = f32(x[j] - x[i]), f32(y[j] - y[i])
                n2 = f32(f32(dx * dx) + f32(dy * dy))
                if n2 > dmin_sq:
                    lg.append((i, j, int(float(f32(dx / n2)) * 2048.0 + 0.5), int(float(f32(dy / n2)) * 2048.0 + 0.5)))
                elif n2 < dmax_sq:
                    sh.append((i, j))
        _PAIRS = (np.array(sh, np.int32), np.array(lg, np.int32))
    return _PAIRS


def _c_div(a: int, b: int) -> int:
    q = abs(a) // abs(b)
    return q if (a >= 0) == (b >= 0) else -q


def smoothed_intensity(img, integral, key_x, key_y, px, py, sigma_half) -> int:
    rows, cols = img.shape
    xf, yf = f32(px + f32(key_x)), f32(py + f32(key_y))
    x, y = int(xf), int(yf)
    area = f32(f32(f32(4.0) * sigma_half) * sigma_half)
    if sigma_half < 0.5:
        r_x, r_y = int(f32(f32(xf - f32(x)) * f32(1024))), int(f32(f32(yf - f32(y)) * f32(1024)))
        r_x_1, r_y_1 = 1024 - r_x, 1024 - r_y
        v = (r_x_1 * r_y_1 * int(img[y, x]) + r_x * r_y_1 * int(img[y, x + 1]) + r_x * r_y * int(img[y + 1, x + 1]) +
             r_x_1 * r_y * int(img[y + 1, x]))
        return _c_div(v + 512, 1024)
    scaling = int(4194304.0 / float(area))
    scaling2 = int(float(f32(f32(f32(scaling) * area) / f32(1024.0))))
    x_1, x1 = f32(xf - sigma_half), f32(xf + sigma_half)
    y_1, y1 = f32(yf - sigma_half), f32(yf + sigma_half)
    x_left, y_top = int(float(x_1) + 0.5), int(float(y_1) + 0.5)
    x_right, y_bottom = int(float(x1) + 0.5), int(float(y1) + 0.5)
    r_x_1 = f32(f32(f32(x_left) - x_1) + f32(0.5))
    r_y_1 = f32(f32(f32(y_top) - y_1) + f32(0.5))
    r_x1 = f32(f32(x1 - f32(x_right)) + f32(0.5))
    r_y1 = f32(f32(y1 - f32(y_bottom)) + f32(0.5))
    dx, dy = x_right - x_left - 1, y_bottom - y_top - 1
    fs = f32(scaling)
    A, B = int(f32(f32(r_x_1 * r_y_1) * fs)), int(f32(f32(r_x1 * r_y_1) * fs))
    C, D = int(f32(f32(r_x1 * r_y1) * fs)), int(f32(f32(r_x_1 * r_y1) * fs))
    r_x_1_i, r_y_1_i = int(f32(r_x_1 * fs)), int(f32(r_y_1 * fs))
    r_x1_i, r_y1_i = int(f32(r_x1 * fs)), int(f32(r_y1 * fs))
    I = lambda yy, xx: int(img[yy, xx])
    if dx + dy > 2:
        ret = A * I(y_top, x_left) + B * I(y_top, x_right) + C * I(y_bottom, x_right) + D * I(y_bottom, x_left)
        S = lambda ya, yb, xa, xb: int(integral[yb, xb]) - int(integral[ya, xb]) - int(integral[yb, xa]) + int(integral[ya, xa])
        # pixel rows ya .. yb-1, columns xa .. xb-1 of the image
        upper = S(y_top, y_top + 1, x_left + 1, x_right) * r_y_1_i
        middle = S(y_top + 1, y_bottom, x_left + 1, x_right) * scaling
        left = S(y_top + 1, y_bottom, x_left, x_left + 1) * r_x_1_i
        right = S(y_top + 1, y_bottom, x_right, x_right + 1) * r_x1_i
        bottom = S(y_bottom, y_bottom + 1, x_left + 1, x_right) * r_y1_i
        return _c_div(_wrap32(ret + upper + middle + left + right + bottom + scaling2 // 2), scaling2)
    ret = A * I(y_top, x_left)
    for xx in range(x_left + 1, x_right):
        ret += r_y_1_i * I(y_top, xx)
    ret += B * I(y_top, x_right)
    for yy in range(y_top + 1, y_bottom):
        ret += r_x_1_i * I(yy, x_left)
        for xx in range(x_left + 1, x_right):
            ret += I(yy, xx) * scaling
        ret += r_x1_i * I(yy, x_right)
    ret += D * I(y_bottom, x_left)
    for xx in range(x_left + 1, x_right):
        ret += r_y1_i * I(y_bottom, xx)
    ret += C * I(y_bottom, x_right)
    return _c_div(_wrap32(ret + scaling2 // 2), scaling2)


def _wrap32(v: int) -> int:
    v &= 0xFFFFFFFF
    return v - (1 << 32) if v & 0x80000000 else v


def keypoint_scale(size: float) -> int:
    lb_scalerange = f32(LOGF_30 / f32(0.693147180559945))
    basic06 = f32(f32(BASIC_SIZE) * f32(0.6))
    v = float(f32(f32(f32(SCALES) / lb_scalerange) * f32(f32(math.log(float(f32(f32(size) / basic06)))) / f32(0.693147180559945)))) + 0.5
    return min(max(int(v), 0), SCALES - 1)


def describe_keypoints_brisk(img: np.ndarray, kps: np.ndarray, compute_orientation: bool = True):
    """img: h x w u8; kps: n x 4 float32 (x, y, size, angle).  Returns (kept indices, angles, n_kept x 64 u8)."""
    img = np.asarray(img, np.uint8)
    rows, cols = img.shape
    kps = np.asarray(kps, np.float32).reshape(-1, 4)
    integral = np.zeros((rows + 1, cols + 1), np.int64)
    integral[1:, 1:] = img.astype(np.int64).cumsum(0).cumsum(1)
    sh, lg = pairs()
    kept, angles, out = [], [], []
    for k, (x, y, size, angle) in enumerate(kps):
        if not size > 0:                 # DescriptorExtractor::compute (2.4) drops these before the extractor runs
            continue
        scale = keypoint_scale(float(size))
        border = size_of_scale(scale)
        # RoiPredicate: removed if x < minX || x >= maxX || y < minY || y >= maxY
        if x < border or x >= cols - border or y < border or y >= rows - border:
            continue
        if compute_orientation:
            px, py, sg = pattern(scale, 0)
            vals = [smoothed_intensity(img, integral, x, y, px[i], py[i], sg[i]) for i in range(POINTS)]
            d0 = d1 = 0
            for i, j, wdx, wdy in lg:
                dt = vals[i] - vals[j]
                d0 += _c_div(dt * int(wdx), 1024)
                d1 += _c_div(dt * int(wdy), 1024)
            angle = f32(math.atan2(float(f32(d1)), float(f32(d0))) / math.pi * 180.0)     # evaluated in double (pinned: exact at scale 0)
        if angle == -1:
            theta = 0
        else:
            theta = int(N_ROT * (float(angle) / 360.0) + 0.5)
            if theta < 0:
                theta += N_ROT
            if theta >= N_ROT:
                theta -= N_ROT
        if angle < 0:
            angle = f32(angle + f32(360.0))
        px, py, sg = pattern(scale, theta)
        vals = [smoothed_intensity(img, integral, x, y, px[i], py[i], sg[i]) for i in range(POINTS)]
        bits = np.array([vals[i] > vals[j] for i, j in sh], np.uint8)
        row = np.packbits(bits, bitorder="little")
        kept.append(k)
        angles.append(angle)
        out.append(row)
    return (np.array(kept, np.int64), np.array(angles, np.float32),
            np.array(out, np.uint8).reshape(-1, 64))
