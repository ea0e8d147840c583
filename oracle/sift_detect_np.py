"""TEST INFRASTRUCTURE -- CPU oracle for K14 (fm3d_detect_sift), numpy on top of the OpenCV primitives the reference calls.

Restates what feature_detector_->detect(frame, keypoints) computes in DescriptorsMatcher::compareWithNNDR / compare /
crosscompare (DescriptorsMatcher/descriptorsmatcher.cpp:110-111, :91-92, :76-77) when DetectorType is SIFT (:243-256,
cv::SIFT(NumFeatures, NumOctaveLayers, ContrastThreshold, EdgeThreshold, Sigma) / cv::SiftFeatureDetector).  OpenCV is a
third-party dependency of the reference (unpinned, 2.4.x era); the published algorithm restated here is cv::SIFT's detector
stage as shipped in the OpenCV 4.13 of this image (modules/features2d/src/sift.dispatch.cpp, sift.simd.hpp):

  * createInitialImage: float(gray), doubled with cv::resize(INTER_LINEAR), blurred to sigma: GaussianBlur with
    sqrt(sigma^2 - 4 * 0.5^2);
  * buildGaussianPyramid: nOctaves = cvRound(log2(min(cols, rows) of the doubled image) - 2) + 1 octaves of nOctaveLayers + 3
    images; image i of an octave = GaussianBlur(image i - 1, sqrt((sigma k^i)^2 - (sigma k^(i-1))^2)), k = 2^(1/nOctaveLayers);
    image 0 of the next octave = every second pixel (INTER_NEAREST) of image nOctaveLayers;
  * buildDoGPyramid: differences of neighbouring images;
  * findScaleSpaceExtrema: |v| > floor(0.5 contrastThreshold / nOctaveLayers * 255) and v >= (<=) its 26 neighbours, 5 pixels
    of border; adjustLocalExtrema: at most five Newton steps of the 3-D quadratic fit (3 x 3 system by LU), contrast and
    edge tests, keypoint position / size / packed octave / response; calcOrientationHist: 36 bins over a disc of radius
    round(4.5 scl), smoothed with (1 4 6 4 1)/16; one keypoint per peak >= 0.8 of the maximum, parabolic bin refinement;
  * KeyPointsFilter::removeDuplicatedSorted (sort by x, y, -size, angle, -response, -octave; drop equal x, y, size, angle),
    retainBest(nfeatures) as a SET (std::nth_element leaves the order unspecified), then the keypoints of the doubled image
    are scaled back (pt / 2, size / 2, octave - 1).

The blurs and the resize are cv2's own (the very functions cv::SIFT calls); the exp / fastAtan2 / magnitude of the
orientation histogram are restated (cv::hal::exp32f is a table + polynomial whose last bits are not reproduced: a peak
within float rounding of the 0.8 threshold can differ).  Pinned by cv2.SIFT_create().detect itself
(tests/test_oracle_pins.py, where cv2 is importable) and the committed golden vectors tests/golden/sift_detect.npz written
from cv2 by tools/make_golden.py.  Only tests/ may import this module.
"""
from __future__ import annotations

import math

import numpy as np

from .sift_patch_np import f32, fast_atan2_deg

IMG_BORDER = 5
MAX_INTERP_STEPS = 5
ORI_HIST_BINS = 36
ORI_SIG_FCTR = 1.5
ORI_RADIUS = 4.5          # 3 * ORI_SIG_FCTR
ORI_PEAK_RATIO = 0.8
INIT_SIGMA = 0.5
FLT_EPSILON = 1.1920929e-07


def _cv_round(v) -> int:
    return int(np.rint(np.float64(v)))          # cvRound: round half to even


def gaussian_pyramid(img: np.ndarray, n_octave_layers: int = 3, sigma: float = 1.6):
    """createInitialImage + buildGaussianPyramid: list of octaves, each a list of nOctaveLayers + 3 float32 images."""
    import cv2

    gray = np.asarray(img, np.uint8).astype(np.float32)
    sig_diff = math.sqrt(max(sigma * sigma - INIT_SIGMA * INIT_SIGMA * 4, 0.01))
    dbl = cv2.resize(gray, (gray.shape[1] * 2, gray.shape[0] * 2), interpolation=cv2.INTER_LINEAR)
    base = cv2.GaussianBlur(dbl, (0, 0), sigmaX=float(np.float32(sig_diff)), sigmaY=float(np.float32(sig_diff)))
    n_octaves = _cv_round(math.log(float(min(base.shape))) / math.log(2.0) - 2) + 1
    k = 2.0 ** (1.0 / n_octave_layers)
    sig = [sigma]
    for i in range(1, n_octave_layers + 3):
        sig_prev = (k ** (i - 1)) * sigma
        sig_total = sig_prev * k
        sig.append(math.sqrt(sig_total * sig_total - sig_prev * sig_prev))
    pyr = []
    for o in range(n_octaves):
        octave = []
        for i in range(n_octave_layers + 3):
            if o == 0 and i == 0:
                octave.append(base)
            elif i == 0:
                src = pyr[o - 1][n_octave_layers]
                octave.append(np.ascontiguousarray(src[0:2 * (src.shape[0] // 2):2, 0:2 * (src.shape[1] // 2):2]))   # INTER_NEAREST, half size
            else:
                octave.append(cv2.GaussianBlur(octave[i - 1], (0, 0), sigmaX=sig[i], sigmaY=sig[i]))
        pyr.append(octave)
    return pyr


def _exp_f32(x):
    return np.exp(np.asarray(x, np.float32)).astype(np.float32)


def orientation_hist(img: np.ndarray, px: int, py: int, radius: int, sigma: float):
    n = ORI_HIST_BINS
    rows, cols = img.shape
    expf_scale = f32(-1.0) / (f32(2.0) * f32(sigma) * f32(sigma))
    ii, jj = np.mgrid[-radius:radius + 1, -radius:radius + 1]
    y = py + ii
    x = px + jj
    ok = (y > 0) & (y < rows - 1) & (x > 0) & (x < cols - 1)
    y, x, ii, jj = y[ok], x[ok], ii[ok], jj[ok]
    dx = (img[y, x + 1] - img[y, x - 1]).astype(np.float32)
    dy = (img[y - 1, x] - img[y + 1, x]).astype(np.float32)
    w = _exp_f32((ii * ii + jj * jj).astype(np.float32) * expf_scale)
    ori = fast_atan2_deg(dy, dx)
    mag = np.sqrt(dx * dx + dy * dy).astype(np.float32)
    temphist = np.zeros(n, np.float32)
    bins = np.rint(np.float32(n / 360.0) * ori).astype(np.int64)
    bins[bins >= n] -= n
    bins[bins < 0] += n
    for b, wk, mk in zip(bins, w, mag):                     # sequential float sum, the order of the scalar loop
        temphist[b] = f32(temphist[b] + f32(wk * mk))
    t = np.concatenate([temphist[-2:], temphist, temphist[:2]])
    hist = ((t[0:n] + t[4:n + 4]) * f32(1.0 / 16.0) + (t[1:n + 1] + t[3:n + 3]) * f32(4.0 / 16.0) + t[2:n + 2] * f32(6.0 / 16.0)).astype(np.float32)
    return hist, f32(hist.max())


def _adjust_local_extrema(dog, octv, layer, r, c, n_octave_layers, contrast_threshold, edge_threshold, sigma):
    img_scale = f32(1.0 / 255.0)
    deriv_scale = f32(img_scale * f32(0.5))
    second_deriv_scale = img_scale
    cross_deriv_scale = f32(img_scale * f32(0.25))
    xi = xr = xc = f32(0)
    i = 0
    while i < MAX_INTERP_STEPS:
        img, prev, nxt = dog[octv][layer], dog[octv][layer - 1], dog[octv][layer + 1]
        dD = np.array([(img[r, c + 1] - img[r, c - 1]) * deriv_scale, (img[r + 1, c] - img[r - 1, c]) * deriv_scale,
                       (nxt[r, c] - prev[r, c]) * deriv_scale], np.float32)
        v2 = f32(img[r, c] * f32(2))
        dxx = f32((img[r, c + 1] + img[r, c - 1] - v2) * second_deriv_scale)
        dyy = f32((img[r + 1, c] + img[r - 1, c] - v2) * second_deriv_scale)
        dss = f32((nxt[r, c] + prev[r, c] - v2) * second_deriv_scale)
        dxy = f32((img[r + 1, c + 1] - img[r + 1, c - 1] - img[r - 1, c + 1] + img[r - 1, c - 1]) * cross_deriv_scale)
        dxs = f32((nxt[r, c + 1] - nxt[r, c - 1] - prev[r, c + 1] + prev[r, c - 1]) * cross_deriv_scale)
        dys = f32((nxt[r + 1, c] - nxt[r - 1, c] - prev[r + 1, c] + prev[r - 1, c]) * cross_deriv_scale)
        H = np.array([[dxx, dxy, dxs], [dxy, dyy, dys], [dxs, dys, dss]], np.float32)
        X = _lu_solve3(H, dD)
        xi, xr, xc = f32(-X[2]), f32(-X[1]), f32(-X[0])
        if abs(xi) < 0.5 and abs(xr) < 0.5 and abs(xc) < 0.5:
            break
        big = float(2 ** 31 - 1) / 3
        if abs(xi) > big or abs(xr) > big or abs(xc) > big or not (np.isfinite(xi) and np.isfinite(xr) and np.isfinite(xc)):
            return None
        c += _cv_round(xc)
        r += _cv_round(xr)
        layer += _cv_round(xi)
        rows, cols = dog[octv][0].shape
        if layer < 1 or layer > n_octave_layers or c < IMG_BORDER or c >= cols - IMG_BORDER or r < IMG_BORDER or r >= rows - IMG_BORDER:
            return None
        i += 1
    if i >= MAX_INTERP_STEPS:
        return None
    img, prev, nxt = dog[octv][layer], dog[octv][layer - 1], dog[octv][layer + 1]
    dD = np.array([(img[r, c + 1] - img[r, c - 1]) * deriv_scale, (img[r + 1, c] - img[r - 1, c]) * deriv_scale,
                   (nxt[r, c] - prev[r, c]) * deriv_scale], np.float32)
    t = f32(f32(f32(dD[0] * xc) + f32(dD[1] * xr)) + f32(dD[2] * xi))
    contr = f32(f32(img[r, c] * img_scale) + f32(t * f32(0.5)))
    if abs(contr) * n_octave_layers < contrast_threshold:
        return None
    v2 = f32(img[r, c] * f32(2))
    dxx = f32((img[r, c + 1] + img[r, c - 1] - v2) * second_deriv_scale)
    dyy = f32((img[r + 1, c] + img[r - 1, c] - v2) * second_deriv_scale)
    dxy = f32((img[r + 1, c + 1] - img[r + 1, c - 1] - img[r - 1, c + 1] + img[r - 1, c - 1]) * cross_deriv_scale)
    tr = f32(dxx + dyy)
    det = f32(f32(dxx * dyy) - f32(dxy * dxy))
    if det <= 0 or tr * tr * edge_threshold >= (edge_threshold + 1) * (edge_threshold + 1) * det:
        return None
    kp = {
        "x": f32(f32(c + xc) * f32(1 << octv)), "y": f32(f32(r + xr) * f32(1 << octv)),
        "octave": octv + (layer << 8) + (_cv_round((float(xi) + 0.5) * 255) << 16),
        "size": f32(f32(f32(sigma) * f32(np.power(f32(2.0), f32(f32(layer + xi) / f32(n_octave_layers))))) * f32(1 << octv) * f32(2)),
        "response": f32(abs(contr)), "r": r, "c": c, "layer": layer,
    }
    return kp


def _lu_solve3(A, b):
    """cv::Matx33f::solve(b, DECOMP_LU): for 3 x 3 OpenCV uses the closed form (Matx_FastSolveOp<float, 3, 3, 1>): Cramer's
    rule with the determinant expanded along the first row, every operation in float."""
    a = A.astype(np.float32)
    b = b.astype(np.float32)
    F = f32
    d = F(F(F(a[0, 0] * F(F(a[1, 1] * a[2, 2]) - F(a[2, 1] * a[1, 2]))) - F(a[0, 1] * F(F(a[1, 0] * a[2, 2]) - F(a[2, 0] * a[1, 2])))) +
          F(a[0, 2] * F(F(a[1, 0] * a[2, 1]) - F(a[2, 0] * a[1, 1]))))
    if d == 0:
        return np.zeros(3, np.float32)                      # solve() fails: X stays zero
    d = F(F(1) / d)
    x0 = F(d * F(F(F(b[0] * F(F(a[1, 1] * a[2, 2]) - F(a[1, 2] * a[2, 1]))) - F(a[0, 1] * F(F(b[1] * a[2, 2]) - F(a[1, 2] * b[2])))) +
                 F(a[0, 2] * F(F(b[1] * a[2, 1]) - F(a[1, 1] * b[2])))))
    x1 = F(d * F(F(F(a[0, 0] * F(F(b[1] * a[2, 2]) - F(a[1, 2] * b[2]))) - F(b[0] * F(F(a[1, 0] * a[2, 2]) - F(a[1, 2] * a[2, 0])))) +
                 F(a[0, 2] * F(F(a[1, 0] * b[2]) - F(b[1] * a[2, 0])))))
    x2 = F(d * F(F(F(a[0, 0] * F(F(a[1, 1] * b[2]) - F(b[1] * a[2, 1]))) - F(a[0, 1] * F(F(a[1, 0] * b[2]) - F(b[1] * a[2, 0])))) +
                 F(b[0] * F(F(a[1, 0] * a[2, 1]) - F(a[1, 1] * a[2, 0])))))
    return np.array([x0, x1, x2], np.float32)


def detect_sift(img: np.ndarray, nfeatures: int = 0, n_octave_layers: int = 3, contrast_threshold: float = 0.04,
                edge_threshold: float = 10.0, sigma: float = 1.6) -> np.ndarray:
    """cv::SIFT::detect: rows (x, y, size, angle, response, octave) in the order removeDuplicatedSorted leaves."""
    gpyr = gaussian_pyramid(img, n_octave_layers, sigma)
    dog = [[(o[i + 1] - o[i]).astype(np.float32) for i in range(n_octave_layers + 2)] for o in gpyr]
    threshold = math.floor(0.5 * contrast_threshold / n_octave_layers * 255)
    n = ORI_HIST_BINS
    kpts = []
    for o, octave in enumerate(dog):
        rows, cols = octave[0].shape
        if rows <= 2 * IMG_BORDER or cols <= 2 * IMG_BORDER:
            continue
        for layer in range(1, n_octave_layers + 1):
            cur = octave[layer]
            stack = np.stack([octave[layer - 1], cur, octave[layer + 1]])
            ctr = cur[IMG_BORDER:rows - IMG_BORDER, IMG_BORDER:cols - IMG_BORDER]
            mx = np.full(ctr.shape, -np.inf, np.float32)
            mn = np.full(ctr.shape, np.inf, np.float32)
            for s in range(3):
                for dy in (-1, 0, 1):
                    for dx in (-1, 0, 1):
                        v = stack[s, IMG_BORDER + dy:rows - IMG_BORDER + dy, IMG_BORDER + dx:cols - IMG_BORDER + dx]
                        mx = np.maximum(mx, v)
                        mn = np.minimum(mn, v)
            cand = (np.abs(ctr) > threshold) & (((ctr > 0) & (ctr >= mx)) | ((ctr < 0) & (ctr <= mn)))
            for r0, c0 in zip(*np.nonzero(cand)):           # row-major, as the scalar loop walks
                kp = _adjust_local_extrema(dog, o, layer, int(r0) + IMG_BORDER, int(c0) + IMG_BORDER, n_octave_layers,
                                           contrast_threshold, edge_threshold, sigma)
                if kp is None:
                    continue
                scl_octv = f32(f32(kp["size"] * f32(0.5)) / f32(1 << o))
                hist, omax = orientation_hist(gpyr[o][kp["layer"]], kp["c"], kp["r"], _cv_round(f32(ORI_RADIUS) * scl_octv),
                                              f32(f32(ORI_SIG_FCTR) * scl_octv))
                mag_thr = f32(omax * f32(ORI_PEAK_RATIO))
                for j in range(n):
                    l = j - 1 if j > 0 else n - 1
                    r2 = j + 1 if j < n - 1 else 0
                    if hist[j] > hist[l] and hist[j] > hist[r2] and hist[j] >= mag_thr:
                        b = f32(j + f32(f32(0.5) * f32(hist[l] - hist[r2])) / f32(f32(hist[l] - f32(2) * hist[j]) + hist[r2]))
                        b = f32(n + b) if b < 0 else (f32(b - n) if b >= n else b)
                        ang = f32(f32(360.0) - f32(f32(360.0 / n) * b))
                        if abs(float(ang) - 360.0) < FLT_EPSILON:
                            ang = f32(0)
                        kpts.append((kp["x"], kp["y"], kp["size"], ang, kp["response"], kp["octave"]))
    if not kpts:
        return np.zeros((0, 6), np.float64)
    K = np.array(kpts, np.float64)
    # removeDuplicatedSorted
    order = np.lexsort((-K[:, 5], -K[:, 4], K[:, 3], -K[:, 2], K[:, 1], K[:, 0]))
    K = K[order]
    keep = np.ones(len(K), bool)
    last = 0
    for j in range(1, len(K)):
        if (K[last, :4] == K[j, :4]).all():
            keep[j] = False
        else:
            last = j
    K = K[keep]
    if nfeatures > 0 and len(K) > nfeatures:                # retainBest: the set of the nfeatures best (ties of the last kept)
        thr = np.sort(K[:, 4])[::-1][nfeatures - 1]
        K = K[K[:, 4] >= thr]
    # back from the doubled image (firstOctave = -1)
    oc = K[:, 5].astype(np.int64)
    K[:, 5] = (oc & ~255) | ((oc - 1) & 255)
    K[:, 0] = (K[:, 0].astype(np.float32) * np.float32(0.5)).astype(np.float64)
    K[:, 1] = (K[:, 1].astype(np.float32) * np.float32(0.5)).astype(np.float64)
    K[:, 2] = (K[:, 2].astype(np.float32) * np.float32(0.5)).astype(np.float64)
    return K
