"""TEST INFRASTRUCTURE -- CPU oracle for K15 (fm3d_detect_orb), numpy only.

Restates what feature_detector_->detect(frame, keypoints) followed by descriptor_extractor_->compute(frame, keypoints,
descriptors) computes in DescriptorsMatcher::compareWithNNDR / compare / crosscompare (DescriptorsMatcher/
descriptorsmatcher.cpp:110-115, :91-96, :76-81) when DetectorType and ExtractorType are ORB (:273-279, :336-342:
cv::ORB(OrbDetector.NumFeatures, ScaleFactor, NumLevels); the remaining arguments keep cv::ORB's defaults: edgeThreshold 31,
firstLevel 0, WTA_K 2, HARRIS_SCORE, patchSize 31, fastThreshold 20).  OpenCV is a third-party dependency of the reference
(unpinned, 2.4.x era); the published algorithm restated here is cv::ORB_Impl::detectAndCompute of the OpenCV 4.13 in this
image (modules/features2d/src/orb.cpp; Rublee et al., ICCV 2011):

  * pyramid: level l has size cvRound(w / s^l) x cvRound(h / s^l); level 0 is the frame, level l > 0 is
    cv::resize(level l - 1, INTER_LINEAR_EXACT) -- 8.8 fixed-point weights cvRound(256 f), rows then columns, one rounding
    (v + 2^15) >> 16 (restated bit for bit: resize_linear_exact);
  * per level: cv::FAST(fastThreshold, nonmax) keypoints (oracle/fast_np.py), runByImageBorder(edgeThreshold),
    retainBest(2 n_l) by FAST score, then the Harris measure of a 7 x 7 block of Sobel-like integer gradients
    (HarrisResponses, k = 0.04) and retainBest(n_l) by it; n_l = cvRound of a geometric share of nfeatures (the last level
    takes the remainder); retainBest keeps everything that ties with the last kept response;
  * orientation: intensity centroid over the disc of radius 15 (ICAngles: integer moments, cv::fastAtan2);
  * keypoint: pt = level position * s^l, size = 31 s^l, octave = l, response = Harris measure;
  * descriptors: the level image blurred 7 x 7, sigma 2 (oracle/orb_np.py: orb_blur), 256 comparisons of the learned pattern
    rotated by the keypoint's angle around its LEVEL position.

std::nth_element leaves the order of the kept keypoints unspecified: results are compared as sets (keyed by level and level
position).  Pinned by cv2.ORB_create(...).detectAndCompute itself (tests/test_oracle_pins.py, where cv2 is importable) and
the committed golden vectors tests/golden/orb_detect.npz written from cv2 by tools/make_golden.py.  Only tests/ may import
this module.
"""
from __future__ import annotations

import math

import numpy as np

from . import fast_np
from .orb_np import PATTERN, orb_blur
from .sift_patch_np import fast_atan2_deg

f32 = np.float32
HARRIS_K = f32(0.04)
HALF_PATCH = 15


def _cv_round(v) -> int:
    return int(np.rint(np.float64(v)))


def resize_linear_exact(src: np.ndarray, dw: int, dh: int) -> np.ndarray:
    """cv::resize(src, Size(dw, dh), 0, 0, INTER_LINEAR_EXACT) for CV_8UC1."""
    sh, sw = src.shape

    def coeffs(ssize, dsize):
        scale = ssize / dsize
        ofs = np.zeros(dsize, np.int64)
        a1 = np.zeros(dsize, np.int64)
        for d in range(dsize):
            f = (d + 0.5) * scale - 0.5
            s = int(math.floor(f))
            f -= s
            if s < 0:
                s, f = 0, 0.0
            if s >= ssize - 1:
                s, f = ssize - 1, 0.0
            ofs[d] = s
            a1[d] = _cv_round(f * 256)
        return ofs, a1

    ox, ax = coeffs(sw, dw)
    oy, ay = coeffs(sh, dh)
    s = src.astype(np.int64)
    rows = s[:, ox] * (256 - ax) + s[:, np.minimum(ox + 1, sw - 1)] * ax
    v = rows[oy, :] * (256 - ay)[:, None] + rows[np.minimum(oy + 1, sh - 1), :] * ay[:, None]
    return np.clip((v + (1 << 15)) >> 16, 0, 255).astype(np.uint8)


def level_scales(nlevels: int, scale_factor: float):
    return [f32(math.pow(float(f32(scale_factor)), l)) for l in range(nlevels)]       # getScale: (float)pow((double)scaleFactor, level)


def build_pyramid(img: np.ndarray, nlevels: int, scale_factor: float):
    img = np.asarray(img, np.uint8)
    h, w = img.shape
    sc = level_scales(nlevels, scale_factor)
    levels = [img]
    for l in range(1, nlevels):
        dw, dh = _cv_round(w / float(sc[l])), _cv_round(h / float(sc[l]))
        levels.append(resize_linear_exact(levels[-1], dw, dh))
    return levels, sc


def features_per_level(nfeatures: int, nlevels: int, scale_factor: float):
    factor = f32(f32(1.0) / f32(scale_factor))
    nd = f32(nfeatures * (1 - float(factor)) / (1 - float(f32(math.pow(float(factor), nlevels)))))
    out, total = [], 0
    for _ in range(nlevels - 1):
        n = _cv_round(nd)
        out.append(n)
        total += n
        nd = f32(nd * factor)
    out.append(max(nfeatures - total, 0))
    return out


def retain_best(resp: np.ndarray, n: int) -> np.ndarray:
    """KeyPointsFilter::retainBest as a set: indices of everything >= the n-th largest response."""
    if n >= len(resp):
        return np.arange(len(resp))
    if n == 0:
        return np.zeros(0, np.int64)
    thr = np.sort(resp)[::-1][n - 1]
    return np.nonzero(resp >= thr)[0]


def harris_responses(img: np.ndarray, xy: np.ndarray, block: int = 7) -> np.ndarray:
    """HarrisResponses (orb.cpp): img is the level image; the reference reads the reflect-101 border of the pyramid buffer
    beyond it, which an edgeThreshold of 31 never reaches."""
    I = img.astype(np.int64)
    r = block // 2
    scale = f32(1.0) / f32((1 << 2) * block * 255.0)
    scale4 = f32(f32(scale * scale) * f32(scale * scale))
    out = np.zeros(len(xy), np.float32)
    for k, (x0, y0) in enumerate(xy):
        ys, xs = np.mgrid[y0 - r:y0 + r + 1, x0 - r:x0 + r + 1]
        Ix = (I[ys, xs + 1] - I[ys, xs - 1]) * 2 + (I[ys - 1, xs + 1] - I[ys - 1, xs - 1]) + (I[ys + 1, xs + 1] - I[ys + 1, xs - 1])
        Iy = (I[ys + 1, xs] - I[ys - 1, xs]) * 2 + (I[ys + 1, xs - 1] - I[ys - 1, xs - 1]) + (I[ys + 1, xs + 1] - I[ys - 1, xs + 1])
        a, b, c = int((Ix * Ix).sum()), int((Iy * Iy).sum()), int((Ix * Iy).sum())
        af, bf, cf = f32(a), f32(b), f32(c)
        ab = f32(af + bf)
        out[k] = f32(f32(f32(f32(af * bf) - f32(cf * cf)) - f32(f32(HARRIS_K * ab) * ab)) * scale4)
    return out


def umax_table():
    hp = HALF_PATCH
    umax = [0] * (hp + 2)
    vmax = int(math.floor(hp * math.sqrt(2.0) / 2 + 1))
    vmin = int(math.ceil(hp * math.sqrt(2.0) / 2))
    for v in range(vmax + 1):
        umax[v] = _cv_round(math.sqrt(float(hp * hp - v * v)))
    v0 = 0
    for v in range(hp, vmin - 1, -1):
        while umax[v0] == umax[v0 + 1]:
            v0 += 1
        umax[v] = v0
        v0 += 1
    return umax


def ic_angles(img: np.ndarray, xy: np.ndarray) -> np.ndarray:
    I = img.astype(np.int64)
    umax = umax_table()
    out = np.zeros(len(xy), np.float32)
    for k, (x0, y0) in enumerate(xy):
        m01 = m10 = 0
        for u in range(-HALF_PATCH, HALF_PATCH + 1):
            m10 += u * int(I[y0, x0 + u])
        for v in range(1, HALF_PATCH + 1):
            d = umax[v]
            us = np.arange(-d, d + 1)
            plus, minus = I[y0 + v, x0 + us], I[y0 - v, x0 + us]
            m10 += int((us * (plus + minus)).sum())
            m01 += v * int((plus - minus).sum())
        out[k] = fast_atan2_deg(np.array([m01], np.float32), np.array([m10], np.float32))[0]
    return out


def describe_level(B: np.ndarray, xy: np.ndarray, angles: np.ndarray) -> np.ndarray:
    """computeOrbDescriptors on the blurred level image B at integer level positions."""
    pa = PATTERN[:, [1, 0]].astype(np.float32)
    pb = PATTERN[:, [3, 2]].astype(np.float32)
    Bi = B.astype(np.int32)
    rows = np.zeros((len(xy), 32), np.uint8)
    for k, ((cx, cy), angle) in enumerate(zip(xy, angles)):
        ang = f32(f32(angle) * f32(math.pi / 180.0))
        a, b = f32(math.cos(float(ang))), f32(math.sin(float(ang)))

        def val(p):
            xx = (p[:, 0] * a).astype(np.float32) - (p[:, 1] * b).astype(np.float32)
            yy = (p[:, 0] * b).astype(np.float32) + (p[:, 1] * a).astype(np.float32)
            return Bi[cy + np.rint(yy.astype(np.float64)).astype(np.int64), cx + np.rint(xx.astype(np.float64)).astype(np.int64)]

        rows[k] = np.packbits((val(pa) < val(pb)).astype(np.uint8), bitorder="little")
    return rows


def detect_and_describe_orb(img: np.ndarray, nfeatures: int = 500, scale_factor: float = 1.2, nlevels: int = 8,
                            edge_threshold: int = 31, fast_threshold: int = 20):
    """Returns (K, D): K rows (x, y, size, angle, response, octave, level x, level y), D n x 32 u8; sorted by (octave, y, x)."""
    levels, sc = build_pyramid(img, nlevels, scale_factor)
    npl = features_per_level(nfeatures, nlevels, scale_factor)
    K, D = [], []
    for l, lim in enumerate(levels):
        h, w = lim.shape
        if npl[l] == 0 or h <= 2 * edge_threshold or w <= 2 * edge_threshold:
            continue
        xy, resp = fast_np.detect_fast(lim, fast_threshold, True)
        xy = xy.astype(np.int64)
        inb = (xy[:, 0] >= edge_threshold) & (xy[:, 0] < w - edge_threshold) & (xy[:, 1] >= edge_threshold) & (xy[:, 1] < h - edge_threshold)
        xy, resp = xy[inb], resp[inb]
        keep = retain_best(resp, 2 * npl[l])
        xy = xy[keep]
        hr = harris_responses(lim, xy)
        keep = retain_best(hr, npl[l])
        xy, hr = xy[keep], hr[keep]
        if len(xy) == 0:
            continue
        ang = ic_angles(lim, xy)
        rows = describe_level(orb_blur(lim), xy, ang)
        s = sc[l]
        for (x, y), r_, a_, row in zip(xy, hr, ang, rows):
            K.append((float(f32(f32(x) * s)), float(f32(f32(y) * s)), float(f32(f32(31.0) * s)), float(a_), float(r_), l, int(x), int(y)))
            D.append(row)
    K = np.array(K, np.float64).reshape(-1, 8)
    D = np.array(D, np.uint8).reshape(-1, 32)
    order = np.lexsort((K[:, 6], K[:, 7], K[:, 5]))
    return K[order], D[order]
