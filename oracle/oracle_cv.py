"""TEST INFRASTRUCTURE -- CPU oracle #1: the reference's hot path restated in Python on top
of the very OpenCV routines the reference calls (cv2 4.13 here; the reference linked
OpenCV 2.4 C++).

The reference (caomw/3DFeatureMatcher) cannot be compiled in this image (needs OpenCV-C++
2.4 `nonfree`, PCL, Boost, lmfit; none present), so this module follows its sources line by
line and delegates to `cv2` wherever the reference delegates to `cv::`:

  matching          DescriptorsMatcher/descriptorsmatcher.cpp:107-131   (cv2.BFMatcher: exact
                    search in place of the reference's randomised FLANN index, as the north
                    star specifies)
  g12               Triangulator/singlecameratriangulator.cpp:123-143, tools.cpp:87-114
  triangulation     Triangulator/singlecameratriangulator.cpp:145-230   (cv2.undistortPoints,
                    cv2.triangulatePoints)
  pyramids          Triangulator/normaloptimizer.cpp:206-221            (cv2.pyrDown)
  disc pixels       Triangulator/singlecameratriangulator.cpp:341-397   (cv2.projectPoints)
  evaluateNormal    Triangulator/normaloptimizer.cpp:65-149 with
                    singlecameratriangulator.cpp:421-470,530-665 and tools.cpp:129-142,767-777
  LM driver         Triangulator/normaloptimizer.cpp:223-292, 321-452  (lmmin: oracle/lmmin_py.py)
  frames            Triangulator/normaloptimizer.cpp:454-505
  patches           Triangulator/neighborhoodsgenerator.cpp:134-158,
                    Triangulator/singlecameratriangulator.cpp:805-849

Pinned by: the cv2 calls themselves (this module IS the reference's arithmetic wherever the
reference calls OpenCV); the hand-written parts (disc lattice, ray/plane intersection,
bilinear sampler, penalty weight, frames, patch layout) are short restatements checked by
tests/test_oracle_pins.py on analytic scenes.  The LM library boundary is unpinned (see
oracle/lmmin_py.py).  Documented deviations from the reference as written:
  D1  the 1024x768 bound hard-coded at singlecameratriangulator.cpp:359 is the image size;
  D2  the sampler reads a continuous cv::Mat with flat addressing like at<uchar>() without
      bounds checks does; bytes beyond the buffer read as 0;
  D12 the PCL visualiser hook (normaloptimizer.cpp:121-123) is not reproduced.

Only tests/, bench.py's cpu_baseline leg and __graft_entry__.smoke() may import this module;
tools/make_golden.py uses it to write tests/golden/*.npz.
"""
from __future__ import annotations

import math

import cv2
import numpy as np

from .lmmin_py import LMControl, lmmin

PENALTY_FABS, PENALTY_INT_ABS, PENALTY_OFF = 0, 1, 2
FEAT_OK, FEAT_NO_PIXELS, FEAT_ABORT_BBOX, FEAT_ABORT_PIXEL, FEAT_ABORT_NAN = 0, 1, 2, 3, 4


# --------------------------------------------------------------------------- settings
def read_settings(path: str) -> dict:
    """settings.yml knobs read with cv::FileStorage like every reference constructor does."""
    fs = cv2.FileStorage(path, cv2.FILE_STORAGE_READ)
    if not fs.isOpened():
        raise FileNotFoundError(path)

    def vec(node):
        return [node.at(i).real() for i in range(node.size())]

    cs = fs.getNode("CameraSettings")
    nb = fs.getNode("Neighborhoods")
    im = fs.getNode("IMAGES")
    s = {
        "img1": im.getNode("img1").string(), "img2": im.getNode("img2").string(),
        "pos1": vec(im.getNode("pos1")), "pos2": vec(im.getNode("pos2")),
        "nndr_epsilon": fs.getNode("NNDR").getNode("epsilon").real(),
        "epsilonLMMIN": nb.getNode("epsilonLMMIN").real(),
        "pixelsRay": int(nb.getNode("pixelsRay").real()),
        "pyramids": int(nb.getNode("pyramids").real()),
        "method": nb.getNode("method").string(),
        "cmPerPixel": nb.getNode("cmPerPixel").real(),
        "epsilon": nb.getNode("epsilon").real(),
        "rodriguesIC": vec(cs.getNode("rodriguesIC")),
        "translationIC": vec(cs.getNode("translationIC")),
        "K": np.array([[cs.getNode("Fx").real(), 0, cs.getNode("Cx").real()],
                       [0, cs.getNode("Fy").real(), cs.getNode("Cy").real()], [0, 0, 1.0]]),
        # singlecameratriangulator.cpp:101-105: (k0,k1,p1,p2,k2) -> OpenCV (k1,k2,p1,p2,k3)
        "dist": np.array([cs.getNode("k0").real(), cs.getNode("k1").real(),
                          cs.getNode("p1").real(), cs.getNode("p2").real(),
                          cs.getNode("k2").real()]),
        "zThresholdMin": cs.getNode("zThresholdMin").real(),
        "zThresholdMax": cs.getNode("zThresholdMax").real(),
        "ExtractorType": fs.getNode("FeatureOptions").getNode("ExtractorType").string(),
    }
    fs.release()
    return s


# --------------------------------------------------------------------------- tools.cpp
def compose_transformation(R, T):
    """tools.cpp:87-99"""
    G = np.eye(4)
    G[:3, :3] = R
    G[:3, 3] = np.asarray(T, dtype=np.float64).ravel()
    return G


def decompose_transformation(G):
    """tools.cpp:101-114 -> (rodrigues r, t)"""
    r, _ = cv2.Rodrigues(np.ascontiguousarray(G[:3, :3]))
    return r.ravel(), G[:3, 3].copy()


def sph2car(phi, theta):
    """tools.cpp:772-777"""
    return np.array([math.cos(theta) * math.cos(phi), math.cos(theta) * math.sin(phi),
                     math.sin(theta)])


def car2sph(v):
    """tools.cpp:767-771 -> (phi, theta)"""
    theta = math.atan2(v[2], math.sqrt(v[0] * v[0] + v[1] * v[1]))
    phi = math.atan2(v[1], v[0])
    return phi, theta


def bilinear32f(img: np.ndarray, x, y):
    """getBilinearInterpPix32f, tools.cpp:129-142, vectorised.  x, y are cast to float32 as
    the reference's parameter types do; arithmetic in float32 with separately rounded
    multiplies and adds, in the reference's expression order.  Flat addressing (D2)."""
    x = np.asarray(x, dtype=np.float64).astype(np.float32)
    y = np.asarray(y, dtype=np.float64).astype(np.float32)
    x0 = np.floor(x.astype(np.float64)).astype(np.int64)
    y0 = np.floor(y.astype(np.float64)).astype(np.int64)
    x1, y1 = x0 + 1, y0 + 1
    h, w = img.shape
    flat = np.ascontiguousarray(img).ravel()

    def at(yy, xx):
        idx = yy * w + xx
        ok = (idx >= 0) & (idx < flat.size)
        return np.where(ok, flat[np.clip(idx, 0, flat.size - 1)], 0).astype(np.float32)

    b00, b01 = at(y0, x0), at(y1, x0)  # bilienar_mat[0], [1]
    b10, b11 = at(y0, x1), at(y1, x1)  # bilienar_mat[2], [3]
    one = np.float32(1.0)
    ax = x - x0.astype(np.float32)
    ay = y - y0.astype(np.float32)
    xm0, xm1 = one - ax, ax
    ym0, ym1 = one - ay, ay
    return xm0 * (b00 * ym0 + b01 * ym1) + xm1 * (b10 * ym0 + b11 * ym1)


# --------------------------------------------------------------------------- matcher
def knn2(desc_a: np.ndarray, desc_b: np.ndarray, hamming: bool):
    """matcher_->knnMatch(a, b, matches, 2), descriptorsmatcher.cpp:117, with the exact
    brute-force matcher.  Returns idx (nq,2) int32 (-1 = missing), dist (nq,2) float32."""
    bf = cv2.BFMatcher(cv2.NORM_HAMMING if hamming else cv2.NORM_L2)
    nq = desc_a.shape[0]
    idx = np.full((nq, 2), -1, np.int32)
    dist = np.full((nq, 2), np.float32(np.inf), np.float32)
    if nq == 0 or desc_b.shape[0] == 0:
        return idx, dist
    res = bf.knnMatch(desc_a, desc_b, 2)
    for i, ms in enumerate(res):
        for k, m in enumerate(ms[:2]):
            idx[i, k] = m.trainIdx
            dist[i, k] = m.distance
    return idx, dist


def nndr_filter(idx, dist, epsilon: float):
    """descriptorsmatcher.cpp:119-129: keep m0 iff two neighbours and d0 <= eps*d1 (double)."""
    qidx, tidx, d = [], [], []
    for i in range(idx.shape[0]):
        if idx[i, 1] >= 0:
            if float(dist[i, 0]) <= epsilon * float(dist[i, 1]):
                qidx.append(i)
                tidx.append(int(idx[i, 0]))
                d.append(dist[i, 0])
    return (np.array(qidx, np.int32), np.array(tidx, np.int32), np.array(d, np.float32))


def compare_with_nndr(desc_a, desc_b, epsilon, hamming=False):
    idx, dist = knn2(desc_a, desc_b, hamming)
    return nndr_filter(idx, dist, epsilon)


def mutual_flags(qidx, tidx, desc_a, desc_b, hamming=False):
    """Cross-check from crosscompare's two raw lists (descriptorsmatcher.cpp:74-87)."""
    idx_ba, _ = knn2(desc_b, desc_a, hamming)
    return (idx_ba[tidx, 0] == qidx).astype(np.uint8)


# --------------------------------------------------------------------------- camera
class Camera:
    """State of SingleCameraTriangulator (singlecameratriangulator.cpp:37-143)."""

    def __init__(self, K, dist, z_min, z_max, g12=None):
        self.K = np.asarray(K, dtype=np.float64).reshape(3, 3)
        self.dist = np.asarray(dist, dtype=np.float64).reshape(5)
        self.z_min, self.z_max = float(z_min), float(z_max)
        self.g12 = None if g12 is None else np.asarray(g12, dtype=np.float64).reshape(4, 4)

    def setg12(self, T1, T2, rod1, rod2, rodIC, tIC):
        """singlecameratriangulator.cpp:123-143"""
        R1, _ = cv2.Rodrigues(np.asarray(rod1, dtype=np.float64))
        R2, _ = cv2.Rodrigues(np.asarray(rod2, dtype=np.float64))
        RIC, _ = cv2.Rodrigues(np.asarray(rodIC, dtype=np.float64))
        g1 = compose_transformation(R1, T1)
        g2 = compose_transformation(R2, T2)
        gIC = compose_transformation(RIC, tIC)
        self.g12 = np.linalg.inv(gIC) @ np.linalg.inv(g2) @ g1 @ gIC
        return self.g12


def undistort_points(cam: Camera, pts):
    """cv::undistortPoints(src, dst, K, dist) without R/P (singlecameratriangulator.cpp:169)."""
    pts = np.asarray(pts, dtype=np.float64).reshape(-1, 1, 2)
    if pts.shape[0] == 0:
        return np.zeros((0, 2))
    return cv2.undistortPoints(pts, cam.K, cam.dist).reshape(-1, 2)


def triangulate(cam: Camera, kp1, kp2, qidx=None, tidx=None):
    """setKeypoints + triangulate (singlecameratriangulator.cpp:145-230).
    Returns xyz_all (n,3), mask (n,) uint8, xyz (ninl,3)."""
    kp1 = np.asarray(kp1, dtype=np.float32)
    kp2 = np.asarray(kp2, dtype=np.float32)
    if qidx is not None:
        a1 = kp1[qidx].astype(np.float64)
        a2 = kp2[tidx].astype(np.float64)
    else:
        a1, a2 = kp1.astype(np.float64), kp2.astype(np.float64)
    n = a1.shape[0]
    if n == 0:
        return np.zeros((0, 3)), np.zeros(0, np.uint8), np.zeros((0, 3))
    u1 = undistort_points(cam, a1)
    u2 = undistort_points(cam, a2)
    P1 = np.eye(3, 4)
    P2 = (np.eye(3, 4) @ cam.g12)
    Xh = cv2.triangulatePoints(P1, P2, u1.T.copy(), u2.T.copy())  # 4 x n
    with np.errstate(divide="ignore", invalid="ignore"):
        xyz_all = (Xh[:3] / Xh[3]).T
    z = xyz_all[:, 2]
    mask = ~((z < cam.z_min) | (z >= cam.z_max))
    mask &= ~np.isnan(z)
    return xyz_all, mask.astype(np.uint8), xyz_all[mask]


# --------------------------------------------------------------------------- pyramids
def compute_pyramids(img, levels: int):
    """normaloptimizer.cpp:206-221: levels+1 images."""
    pyr = [np.ascontiguousarray(img)]
    for _ in range(levels):
        pyr.append(cv2.pyrDown(pyr[-1]))
    return pyr


# --------------------------------------------------------------------------- disc + evaluate
def project_points(cam: Camera, X, rvec, tvec):
    X = np.asarray(X, dtype=np.float64).reshape(-1, 1, 3)
    ip, _ = cv2.projectPoints(X, np.asarray(rvec, dtype=np.float64),
                              np.asarray(tvec, dtype=np.float64), cam.K, cam.dist)
    return ip.reshape(-1, 2)


def extract_pixels_contour(cam: Camera, P, pixels_ray: int, width: int, height: int):
    """extractPixelsContour (singlecameratriangulator.cpp:341-397): centre = distorted
    projection of P; lattice offsets with i (x) outer, j (y) inner; D1 bounds."""
    c = project_points(cam, P, np.zeros(3), np.zeros(3))[0]
    r = pixels_ray
    ii, jj = np.meshgrid(np.arange(-r, r + 1), np.arange(-r, r + 1), indexing="ij")
    keep = (ii * ii + jj * jj) <= r * r
    px = c[0] + ii[keep].astype(np.float64)
    py = c[1] + jj[keep].astype(np.float64)
    ok = ~((px < 0) | (py < 0) | (px >= width) | (py >= height))
    return np.stack([px[ok], py[ok]], axis=1)


def is_pixel_good(x, y, scale, cols, rows):
    """isPixelGood (singlecameratriangulator.cpp:657-665), cols/rows of the CURRENT level."""
    return ~((x < 0) | (x > (1.0 / scale) * cols) | (y < 0) | (y > (1.0 / scale) * rows))


def penalty_weight(phi, theta, mode):
    """normaloptimizer.cpp:126-142.  Returns (w, entered_penalty_branch)."""
    if mode == PENALTY_OFF:
        return 1.0, False
    if mode == PENALTY_INT_ABS:
        a_t, a_p = float(abs(int(theta))), float(abs(int(phi)))
    else:
        a_t, a_p = abs(theta), abs(phi)
    if a_t - math.pi / 2 > 0 or a_p - math.pi > 0:
        w_theta = math.exp(a_t - math.pi / 2) + 1
        w_phi = math.exp(a_p - math.pi + 1) + 1
        return w_phi * w_theta, True
    return 1.0, False


class NormalProblem:
    """Everything evaluateNormal needs for one feature at one pyramid level."""

    def __init__(self, cam: Camera, P, pix, img1_lvl, img2_lvl, scale, penalty_mode):
        self.cam, self.P, self.pix = cam, np.asarray(P, dtype=np.float64), pix
        self.img1, self.img2, self.scale = img1_lvl, img2_lvl, scale
        self.mode = penalty_mode
        self.npenalty = 0
        self.abort = FEAT_OK
        self.r2, self.t2 = decompose_transformation(cam.g12)
        self.cmax = int(2 * cam.z_max)  # singlecameratriangulator.cpp:648 (int truncation, D3)

    def evaluate(self, par):
        """evaluateNormal (normaloptimizer.cpp:65-149) -> (fvec, info)."""
        phi, theta = float(par[0]), float(par[1])
        n = sph2car(phi, theta)
        if np.any(np.isnan(n)):
            self.abort = FEAT_ABORT_NAN
            return None, -1
        cam, P = self.cam, self.P
        # get3dPointsFromImage1Pixels (singlecameratriangulator.cpp:530-565)
        u = undistort_points(cam, self.pix)
        v = np.concatenate([u, np.ones((u.shape[0], 1))], axis=1)
        with np.errstate(divide="ignore", invalid="ignore"):
            k = float(n @ P) / (v @ n)
            X = k[:, None] * v
        if np.any(np.isnan(X)):
            self.abort = FEAT_ABORT_NAN  # exit(-6) in the reference (:465-469)
            return None, -1
        cmax = self.cmax
        inbox = ((X[:, 0] > -cmax) & (X[:, 0] < cmax) & (X[:, 1] > -cmax) & (X[:, 1] < cmax)
                 & (X[:, 2] > 0) & (X[:, 2] < cmax))
        if not np.all(inbox):
            self.abort = FEAT_ABORT_BBOX
            return None, -1
        # updateImage1PixelsIntensity (:576-589)
        rows, cols = self.img1.shape
        s = self.scale
        if not np.all(is_pixel_good(self.pix[:, 0], self.pix[:, 1], s, cols, rows)):
            self.abort = FEAT_ABORT_PIXEL
            return None, -1
        i1 = bilinear32f(self.img1, s * self.pix[:, 0], s * self.pix[:, 1])
        # projectPointsToImage2 (:591-632)
        ip2 = project_points(cam, X, self.r2, self.t2)
        if not np.all(is_pixel_good(ip2[:, 0], ip2[:, 1], s, cols, rows)):
            self.abort = FEAT_ABORT_PIXEL
            return None, -1
        i2 = bilinear32f(self.img2, s * ip2[:, 0], s * ip2[:, 1])
        w, entered = penalty_weight(phi, theta, self.mode)
        if entered:
            self.npenalty += 1
        fvec = w * (i1 - i2).astype(np.float64)  # float32 difference, then double (:147)
        return fvec, 0


def optimize_normal(cam: Camera, P, pyr1, pyr2, pixels_ray, epsilon_lmmin,
                    penalty_mode=PENALTY_FABS, patience=100, minpack_mode=False):
    """One iteration of the loop in computeOptimizedNormals (normaloptimizer.cpp:335-449):
    returns dict(normal, status, nfev[levels+1], npenalty, cost, m)."""
    P = np.asarray(P, dtype=np.float64)
    levels = len(pyr1) - 1
    h, w = pyr1[0].shape
    normal = P / np.linalg.norm(P)
    out = {"normal": normal.copy(), "status": FEAT_OK, "nfev": np.zeros(levels + 1, np.int32),
           "npenalty": 0, "cost": float("nan"), "m": 0}
    pix = extract_pixels_contour(cam, P, pixels_ray, w, h)
    out["m"] = pix.shape[0]
    if pix.shape[0] <= 0:
        out["status"] = FEAT_NO_PIXELS
        return out
    img_scale = float(2.0 ** levels)
    for lvl in range(levels, -1, -1):  # optimize_pyramid (:223-245)
        scale = 1.0 / img_scale
        phi, theta = car2sph(normal)  # optimize (:247-292)
        prob = NormalProblem(cam, P, pix, pyr1[lvl], pyr2[lvl], scale, penalty_mode)
        ctl = LMControl(epsilon=epsilon_lmmin, patience=patience)
        par, st = lmmin(2, [phi, theta], pix.shape[0], prob.evaluate, ctl,
                        minpack_mode=minpack_mode)
        out["nfev"][lvl] = st.nfev
        out["npenalty"] += prob.npenalty
        if st.info == 11:
            out["status"] = prob.abort
            out["normal"] = P / np.linalg.norm(P)
            return out
        normal = sph2car(par[0], par[1])
        out["cost"] = st.fnorm ** 2
        img_scale /= 2.0
    out["normal"] = normal
    return out


def optimize_normals(cam, points, pyr1, pyr2, pixels_ray, epsilon_lmmin,
                     penalty_mode=PENALTY_FABS, patience=100):
    res = [optimize_normal(cam, P, pyr1, pyr2, pixels_ray, epsilon_lmmin, penalty_mode, patience)
           for P in np.asarray(points, dtype=np.float64).reshape(-1, 3)]
    return {
        "normals": np.array([r["normal"] for r in res]).reshape(-1, 3),
        "status": np.array([r["status"] for r in res], np.int32),
        "nfev": np.array([r["nfev"] for r in res], np.int32).reshape(len(res), -1),
        "npenalty": np.array([r["npenalty"] for r in res], np.int32),
        "cost": np.array([r["cost"] for r in res]),
        "m": np.array([r["m"] for r in res], np.int32),
    }


def evaluate_cost(cam, P, phi_theta, pyr1, pyr2, pixels_ray, level, penalty_mode=PENALTY_FABS):
    """Sum of squared residuals of one evaluateNormal call (for parity of the cost kernel)."""
    h, w = pyr1[0].shape
    pix = extract_pixels_contour(cam, P, pixels_ray, w, h)
    if pix.shape[0] == 0:
        return float("nan"), 0, FEAT_NO_PIXELS
    prob = NormalProblem(cam, P, pix, pyr1[level], pyr2[level], 1.0 / (2.0 ** level), penalty_mode)
    fvec, info = prob.evaluate(phi_theta)
    if info < 0:
        return float("nan"), pix.shape[0], prob.abort
    return float(np.dot(fvec, fvec)), pix.shape[0], FEAT_OK


# --------------------------------------------------------------------------- frames + patches
def gravity_from_settings(rodriguesIC):
    """normaloptimizer.cpp:160-176: R_IC^-1 * (0,0,-1)"""
    R, _ = cv2.Rodrigues(np.asarray(rodriguesIC, dtype=np.float64))
    return np.linalg.inv(R) @ np.array([0.0, 0.0, -1.0])


def feature_frames(points, normals, gravity):
    """computeFeaturesFrames (normaloptimizer.cpp:454-505)."""
    frames = []
    g = np.asarray(gravity, dtype=np.float64)
    for P, n in zip(np.asarray(points).reshape(-1, 3), np.asarray(normals).reshape(-1, 3)):
        z = n
        x = np.cross(g, z)
        y = np.cross(z, x)
        x = x / np.linalg.norm(x)
        y = y / np.linalg.norm(y)
        F = np.eye(4)
        F[:3, 0], F[:3, 1], F[:3, 2], F[:3, 3] = x, y, z, P
        frames.append(F)
    return np.array(frames).reshape(-1, 4, 4)


def patch_size(epsilon_m, cm_per_pixel):
    return 2 * int(math.floor(epsilon_m / (0.01 * cm_per_pixel)))


def reference_squared_neighborhood(epsilon_m, cm_per_pixel):
    """getReferenceSquaredNeighborhood (neighborhoodsgenerator.cpp:134-158)."""
    S = patch_size(epsilon_m, cm_per_pixel)
    inc = cm_per_pixel * 0.01
    i, j = np.meshgrid(np.arange(S), np.arange(S), indexing="ij")
    ref = np.zeros((S * S, 3))
    ref[:, 0] = (-epsilon_m + inc * i).ravel()
    ref[:, 1] = (-epsilon_m + inc * j).ravel()
    return ref


def square_neighborhood(frame, epsilon_m, cm_per_pixel):
    """computeSquareNeighborhoodByNormal (neighborhoodsgenerator.cpp:92-132)."""
    ref = reference_squared_neighborhood(epsilon_m, cm_per_pixel)
    h = np.concatenate([ref, np.ones((ref.shape[0], 1))], axis=1) @ np.asarray(frame).T
    return h[:, :3] / h[:, 3:4]


def circular_neighborhood(P, normal, epsilon_m, n_angles, n_rays):
    """computeCircularNeighborhoodByNormal (neighborhoodsgenerator.cpp:238-277) with the look-up
    table of the constructor (:46-66): ray i = 1..rays outer, angle j inner; an all-zero normal
    means P/|P|.  Returns (samples (rays*angles, 3), the normal used)."""
    P = np.asarray(P, dtype=np.float64)
    n = np.asarray(normal, dtype=np.float64).copy()
    if not n.any():
        n = P / np.linalg.norm(P)
    W = np.array([[0, -n[2], n[1]], [n[2], 0, -n[0]], [-n[1], n[0], 0]])   # getSkewMatrix (tools.cpp:122-127)
    s = np.array([0.0, 1.0, -n[1] / n[2]])
    s = s / np.linalg.norm(s) * epsilon_m
    out = []
    for i in range(1, n_rays + 1):
        for j in range(n_angles):
            r, t = i * (epsilon_m / n_rays), j * (2 * math.pi / n_angles)
            st, st2 = math.sin(t), 2 * math.sin(t / 2) * math.sin(t / 2)
            out.append(P + r * (s + W @ s * st + st2 * (W @ W @ s)))
    return np.array(out), n


def project_reference_points(cam: Camera, img1, ref, frame):
    """projectReferencePointsToImageWithFrame (singlecameratriangulator.cpp:805-849).
    Returns patch (S,S) uint8 and image points (S*S,2)."""
    S = int(math.sqrt(ref.shape[0]))
    r, t = decompose_transformation(np.asarray(frame, dtype=np.float64))
    ip = project_points(cam, ref, r, t)
    rows, cols = img1.shape
    good = is_pixel_good(ip[:, 0], ip[:, 1], 1.0, cols, rows)
    val = bilinear32f(img1, np.where(good, ip[:, 0], 0.0), np.where(good, ip[:, 1], 0.0))
    val = np.where(good, val, np.float32(0)).astype(np.float32)
    u8 = val.astype(np.int32).astype(np.uint8)  # static_cast<uchar>: truncation toward zero
    patch = np.zeros((S, S), np.uint8)
    idx = np.arange(S * S)
    col, row = idx % S, idx // S
    patch[col, row] = u8  # patch.at<uchar>(col, row) (:842-846)
    return patch, ip


def project_groups(cam: Camera, img, group, image_id):
    """projectPointsToImage (singlecameratriangulator.cpp:709-767), image_id 1 or 2."""
    S = int(math.sqrt(group.shape[0]))
    if image_id == 1:
        r, t = np.zeros(3), np.zeros(3)
    else:
        r, t = decompose_transformation(cam.g12)
    ip = project_points(cam, group, r, t)
    rows, cols = img.shape
    good = is_pixel_good(ip[:, 0], ip[:, 1], 1.0, cols, rows)
    val = bilinear32f(img, np.where(good, ip[:, 0], 0.0), np.where(good, ip[:, 1], 0.0))
    val = np.where(good, val, np.float32(0)).astype(np.float32)
    u8 = val.astype(np.int32).astype(np.uint8)
    patch = np.zeros((S, S), np.uint8)
    idx = np.arange(S * S)
    patch[idx % S, idx // S] = u8
    return patch, ip


# ---------------------------------------------------------------------------------------------
# patch descriptors (DescriptorsMatcher/descriptorsmatcher.cpp:133-174, ExtractorType SIFT :302-314)
def describe_patches_sift(patches):
    """extractDescriptorsFromPatches: one keypoint per patch at (floor(S/2), floor(S/2)), size = S,
    angle = -1, response = 1, octave = 0, class_id = 0 (:150-157), then
    descriptor_extractor_->compute (:164) and one descriptor row per patch (:166-172)."""
    sift = cv2.SIFT_create()
    out = []
    for patch in np.asarray(patches, np.uint8):
        size = patch.shape[0]
        center = float(int(math.floor(size / 2)))
        kp = cv2.KeyPoint(center, center, float(size), -1.0, 1.0, 0, 0)
        kps, d = sift.compute(np.ascontiguousarray(patch), [kp])
        assert len(kps) == 1 and d.shape == (1, 128)
        out.append(d[0])
    return np.stack(out) if out else np.zeros((0, 128), np.float32)
