"""TEST INFRASTRUCTURE -- CPU oracle for K11 (fm3d_describe_keypoints_sift), numpy only.

Restates what descriptor_extractor_->compute(frame, keypoints, descriptors) computes in
DescriptorsMatcher::compareWithNNDR / compare / crosscompare
(DescriptorsMatcher/descriptorsmatcher.cpp:114-115, :95-96, :80-81) when ExtractorType is SIFT (:302-314,
cv::SIFT / cv::SiftDescriptorExtractor with the default 3 octave layers and sigma 1.6) and the keypoints
come from a detector that leaves octave = 0 (FAST, :215-222: size 7, angle -1; also GFTT/Dense-like
keypoints a caller injects).  OpenCV is a third-party dependency of the reference (unpinned, 2.4.x
era); the published algorithm restated here is cv::SIFT's descriptor stage (modules/features2d/src/
sift.*, OpenCV 4.13 as installed in this image):

  * provided keypoints of octave 0 / layer 0 => firstOctave = 0, no up-sampling, and every descriptor
    is read from gpyr[0] = base = GaussianBlur(float(gray), sigma = sqrt(1.6^2 - 0.5^2)), 13 taps,
    BORDER_REFLECT_101 (createInitialImage);
  * per keypoint calcSIFTDescriptor(base, pt, ori = 360 - angle (0 if that is 360), scl = size / 2,
    d = 4, n = 8): window radius round(3 scl sqrt2 (d + 1) / 2) around the ROUNDED keypoint position,
    central differences, cv::fastAtan2, Gaussian weight, trilinear split into a flat
    (d+2)(d+2)(n+2) histogram, circular fold, clip at 0.2 |h|, x 512, saturate to u8.
    A keypoint with angle -1 has ori = 361: orientations below 1 degree keep o0 = -1 after the single
    wrap, i.e. flat slot n+1 of the previous column cell (see oracle/sift_patch_np.py) -- restated.

Pinned by: cv2.SIFT_create().compute itself on FAST keypoints and on keypoints with real angles / sizes
(tests/test_oracle_pins.py, where cv2 is importable) and the committed golden vectors
tests/golden/sift_keypoints.npz written from cv2 by tools/make_golden.py; agreement is exact up to +-1
on isolated quantised values (summation order of float sums).  Only tests/ may import this module.
"""
from __future__ import annotations

import math

import numpy as np

from .sift_patch_np import f32, fast_atan2_deg, gaussian_blur_f32


def sift_base_image(img: np.ndarray) -> np.ndarray:
    """createInitialImage(img, doubleImageSize = false, sigma = 1.6) for a u8 gray image."""
    sigma = math.sqrt(max(1.6 * 1.6 - 0.5 * 0.5, 0.01))
    return gaussian_blur_f32(np.asarray(img, np.uint8), sigma)


def _cv_round(v: float) -> int:
    return int(np.rint(np.float64(v)))          # cvRound: round half to even


def describe_keypoint_sift(base: np.ndarray, x: float, y: float, size: float, angle: float) -> np.ndarray:
    rows, cols = base.shape
    d, n = 4, 8
    ptx, pty = _cv_round(f32(x)), _cv_round(f32(y))
    ori = f32(360.0) - f32(angle)
    if abs(float(ori) - 360.0) < 1.1920929e-07:
        ori = f32(0.0)
    scl = f32(size) * f32(0.5)
    cos_t = f32(np.cos(f32(ori * f32(np.pi / 180))))
    sin_t = f32(np.sin(f32(ori * f32(np.pi / 180))))
    bins_per_rad = f32(n / 360.0)
    exp_scale = f32(-1.0 / (d * d * 0.5))
    hist_width = f32(3.0) * scl
    radius = _cv_round(float(hist_width * f32(1.4142135623730951) * f32(d + 1) * f32(0.5)))
    radius = min(radius, int(math.sqrt(float(cols) * cols + float(rows) * rows)))
    cos_t = f32(cos_t / hist_width)
    sin_t = f32(sin_t / hist_width)
    ii, jj = np.meshgrid(np.arange(-radius, radius + 1), np.arange(-radius, radius + 1), indexing="ij")
    ii, jj = ii.ravel(), jj.ravel()
    r, c = pty + ii, ptx + jj
    inside = (r > 0) & (r < rows - 1) & (c > 0) & (c < cols - 1)
    ii, jj, r, c = ii[inside], jj[inside], r[inside], c[inside]
    i, j = ii.astype(np.float32), jj.astype(np.float32)
    c_rot = j * cos_t - i * sin_t
    r_rot = j * sin_t + i * cos_t
    rbin = r_rot + f32(d // 2) - f32(0.5)
    cbin = c_rot + f32(d // 2) - f32(0.5)
    ok = (rbin > -1) & (rbin < d) & (cbin > -1) & (cbin < d)
    r, c, rbin, cbin, c_rot, r_rot = r[ok], c[ok], rbin[ok], cbin[ok], c_rot[ok], r_rot[ok]
    dx = base[r, c + 1] - base[r, c - 1]
    dy = base[r - 1, c] - base[r + 1, c]
    W = np.exp(((c_rot * c_rot + r_rot * r_rot) * exp_scale).astype(np.float32)).astype(np.float32)
    Ori = fast_atan2_deg(dy, dx)
    Mag = np.sqrt(dx * dx + dy * dy).astype(np.float32)
    obin = ((Ori - ori) * bins_per_rad).astype(np.float32)
    mag = (Mag * W).astype(np.float32)
    r0 = np.floor(rbin).astype(int)
    c0 = np.floor(cbin).astype(int)
    o0 = np.floor(obin).astype(int)
    rb, cb, ob = rbin - r0, cbin - c0, obin - o0
    o0 = np.where(o0 < 0, o0 + n, o0)
    o0 = np.where(o0 >= n, o0 - n, o0)
    hist = np.zeros((d + 2) * (d + 2) * (n + 2), np.float64)
    v_r1 = mag * rb
    v_r0 = mag - v_r1
    v_rc11 = v_r1 * cb
    v_rc10 = v_r1 - v_rc11
    v_rc01 = v_r0 * cb
    v_rc00 = v_r0 - v_rc01
    for vv, dr, dc in ((v_rc00, 0, 0), (v_rc01, 0, 1), (v_rc10, 1, 0), (v_rc11, 1, 1)):
        v1 = vv * ob
        v0 = vv - v1
        idx = ((r0 + 1 + dr) * (d + 2) + c0 + 1 + dc) * (n + 2) + o0      # flat: o0 = -1 lands in the previous cell
        np.add.at(hist, idx, v0)
        np.add.at(hist, idx + 1, v1)
    hist = hist.reshape(d + 2, d + 2, n + 2)
    hist[:, :, 0] += hist[:, :, n]
    hist[:, :, 1] += hist[:, :, n + 1]
    dst = hist[1:d + 1, 1:d + 1, :n].reshape(-1).astype(np.float32)
    nrm2 = float((dst * dst).sum())
    thr = math.sqrt(nrm2) * 0.2
    dst = np.minimum(dst, f32(thr))
    nrm2 = float((dst * dst).sum())
    k = f32(512.0 / max(math.sqrt(nrm2), 1.1920929e-07))
    return np.clip(np.rint(dst * k), 0, 255).astype(np.float32)


def describe_keypoints_sift(img: np.ndarray, kps: np.ndarray) -> np.ndarray:
    """img: h x w u8; kps: n x 4 float32 (x, y, size, angle as cv::KeyPoint) -> n x 128 float32."""
    kps = np.asarray(kps, np.float32).reshape(-1, 4)
    if len(kps) == 0:
        return np.zeros((0, 128), np.float32)
    base = sift_base_image(img)
    return np.stack([describe_keypoint_sift(base, *map(float, k)) for k in kps])
