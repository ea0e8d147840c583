"""TEST INFRASTRUCTURE -- CPU oracle for K9 (fm3d_describe_patches_sift), numpy only.

Restates what DescriptorsMatcher::extractDescriptorsFromPatches
(DescriptorsMatcher/descriptorsmatcher.cpp:133-174) computes when ExtractorType is SIFT (:302-314):
cv::SIFT::compute on each patch with ONE provided keypoint at (floor(S/2), floor(S/2)), size S,
angle -1, octave 0.  OpenCV is a third-party dependency of the reference (unpinned, 2.4.x era); the
published algorithm restated here is cv::SIFT's (modules/features2d/src/sift.*, OpenCV 4.13 as
installed in this image):

  * provided keypoints of octave 0 => firstOctave = 0, one octave, no up-sampling;
    base = GaussianBlur(float(gray), sigma = sqrt(max(1.6^2 - 0.5^2, 0.01))), kernel size
    cvRound(8 sigma + 1) | 1 = 13, BORDER_REFLECT_101;
  * calcSIFTDescriptor(base, pt, ori = 360 - angle = 361, scl = size/2, d = 4, n = 8):
    central differences, cv::fastAtan2's degree polynomial, magnitude, Gaussian weight
    exp(-(r_rot^2 + c_rot^2) / (d^2 / 2)), trilinear interpolation into a (d+2)(d+2)(n+2)
    histogram (flat addressing, see the note in describe_patch_sift), circular fold of the
    orientation bins, clip at 0.2 |h|, scale to 512, saturate to u8 (returned as float32,
    cv::SIFT's default descriptor type).

Pinned by: cv2.SIFT_create().compute itself (tests/test_oracle_pins.py, where cv2 is importable)
and the committed golden vectors tests/golden/sift_patches.npz written from cv2 by
tools/make_golden.py; agreement is exact up to +-1 on isolated quantised values (summation order
of float sums).  Only tests/ may import this module.
"""
from __future__ import annotations

import math

import numpy as np

f32 = np.float32


def gaussian_kernel_f32(ksize: int, sigma: float) -> np.ndarray:
    """cv::getGaussianKernel(ksize, sigma, CV_32F) for sigma > 0."""
    x = np.arange(ksize) - (ksize - 1) * 0.5
    t = np.exp(-0.5 / (sigma * sigma) * x * x)
    return (t / t.sum()).astype(np.float32)


def gaussian_blur_f32(img: np.ndarray, sigma: float) -> np.ndarray:
    """cv::GaussianBlur(CV_32F image, Size(), sigma) with BORDER_REFLECT_101 (separable, float)."""
    ksize = int(round(sigma * 8 + 1)) | 1
    k = gaussian_kernel_f32(ksize, sigma)
    h = ksize // 2
    rows, cols = img.shape

    def reflect(n):
        idx = np.abs(np.arange(-h, n + h))
        return np.where(idx >= n, 2 * (n - 1) - idx, idx)

    p = img.astype(np.float32)[:, reflect(cols)]
    row = np.zeros((rows, cols), np.float32)
    for q in range(ksize):
        row += k[q] * p[:, q:q + cols]
    p2 = row[reflect(rows), :]
    out = np.zeros((rows, cols), np.float32)
    for q in range(ksize):
        out += k[q] * p2[q:q + rows, :]
    return out


def fast_atan2_deg(y: np.ndarray, x: np.ndarray) -> np.ndarray:
    """cv::fastAtan2 on float arrays: degrees in [0, 360)."""
    s = 180.0 / np.pi
    p1, p3, p5, p7 = (f32(0.9997878412794807 * s), f32(-0.3258083974640975 * s),
                      f32(0.1555786518463281 * s), f32(-0.04432655554792128 * s))
    x = x.astype(np.float32)
    y = y.astype(np.float32)
    ax, ay = np.abs(x), np.abs(y)
    eps = f32(2.220446049250313e-16)
    ge = ax >= ay
    c = np.where(ge, ay / (ax + eps), ax / (ay + eps)).astype(np.float32)
    c2 = c * c
    a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c
    a = np.where(ge, a, f32(90.0) - a)
    a = np.where(x < 0, f32(180.0) - a, a)
    a = np.where(y < 0, f32(360.0) - a, a)
    return a.astype(np.float32)


def describe_patch_sift(patch: np.ndarray) -> np.ndarray:
    """One patch (S x S u8) -> 128 float32 values, as cv::SIFT::compute with the reference's keypoint."""
    S = patch.shape[0]
    assert patch.shape == (S, S) and S >= 8
    sigma = math.sqrt(max(1.6 * 1.6 - 0.5 * 0.5, 0.01))
    img = gaussian_blur_f32(patch, sigma)
    d, n = 4, 8
    pt = int(round(float(S // 2)))
    ori = f32(360.0) - f32(-1.0)
    scl = f32(S) * f32(0.5)
    cos_t = f32(np.cos(f32(ori * f32(np.pi / 180))))
    sin_t = f32(np.sin(f32(ori * f32(np.pi / 180))))
    bins_per_rad = f32(n / 360.0)
    exp_scale = f32(-1.0 / (d * d * 0.5))
    hist_width = f32(3.0) * scl
    radius = int(round(float(hist_width * f32(1.4142135623730951) * f32(d + 1) * f32(0.5))))
    radius = min(radius, int(math.sqrt(float(S) * S + float(S) * S)))
    cos_t = f32(cos_t / hist_width)
    sin_t = f32(sin_t / hist_width)
    ii, jj = np.meshgrid(np.arange(-radius, radius + 1), np.arange(-radius, radius + 1), indexing="ij")
    ii, jj = ii.ravel(), jj.ravel()
    r, c = pt + ii, pt + jj
    inside = (r > 0) & (r < S - 1) & (c > 0) & (c < S - 1)      # cheap pre-filter, part of the reference's test
    ii, jj, r, c = ii[inside], jj[inside], r[inside], c[inside]
    i, j = ii.astype(np.float32), jj.astype(np.float32)
    c_rot = j * cos_t - i * sin_t
    r_rot = j * sin_t + i * cos_t
    rbin = r_rot + f32(d // 2) - f32(0.5)
    cbin = c_rot + f32(d // 2) - f32(0.5)
    ok = (rbin > -1) & (rbin < d) & (cbin > -1) & (cbin < d)
    r, c, rbin, cbin, c_rot, r_rot = r[ok], c[ok], rbin[ok], cbin[ok], c_rot[ok], r_rot[ok]
    dx = img[r, c + 1] - img[r, c - 1]
    dy = img[r - 1, c] - img[r + 1, c]
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
    # NOTE the reference's keypoint has angle -1, so ori = 361 lies outside [0, 360): for gradient
    # orientations below 1 degree obin < -n, and the single `if (o0 < 0) o0 += n` of
    # calcSIFTDescriptor leaves o0 = -1.  OpenCV addresses the histogram flat, so that entry is
    # slot n+1 of the PREVIOUS column cell (the padding makes it memory-safe) and the circular fold
    # adds it to orientation bin 1 of that cell.  Restated with the same flat addressing.
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
        idx = ((r0 + 1 + dr) * (d + 2) + c0 + 1 + dc) * (n + 2) + o0
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


def describe_patches_sift(patches: np.ndarray) -> np.ndarray:
    patches = np.asarray(patches, np.uint8)
    return np.stack([describe_patch_sift(p) for p in patches]) if len(patches) else np.zeros((0, 128), np.float32)
