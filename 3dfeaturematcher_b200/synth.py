"""Deterministic synthetic stereo pairs for tests and bench.py.

There is no dataset in the reference (its `loop_dataset` images are not in the tree), so
every measurement and parity test runs on scenes generated here: a convex faceted surface
in front of camera 1, textured by a band-limited function of 3-D position, ray-cast exactly
into both views through the reference's lens model (OpenCV k1,k2,p1,p2,k3).  Pure numpy;
no reference code, no oracle.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

# build/settings.yml of the reference: (k0,k1,p1,p2,k2) -> OpenCV order (k1,k2,p1,p2,k3)
SETTINGS_DIST = np.array([-0.299957, 0.124129, -6.6e-05, 0.000567, -0.028357])
SETTINGS_RODRIGUES_IC = np.array([-1.2005, 1.1981, -1.2041])
SETTINGS_TRANSLATION_IC = np.array([0.0, 0.015, -0.051])


def rodrigues(r):
    r = np.asarray(r, dtype=np.float64)
    th = float(np.linalg.norm(r))
    if th < 1e-300:
        return np.eye(3)
    k = r / th
    Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return math.cos(th) * np.eye(3) + (1 - math.cos(th)) * np.outer(k, k) + math.sin(th) * Kx


def distort(x, y, d):
    k1, k2, p1, p2, k3 = d
    r2 = x * x + y * y
    rad = 1 + ((k3 * r2 + k2) * r2 + k1) * r2
    xd = x * rad + 2 * p1 * x * y + p2 * (r2 + 2 * x * x)
    yd = y * rad + p1 * (r2 + 2 * y * y) + 2 * p2 * x * y
    return xd, yd


def undistort_exact(xd, yd, d, iters=60):
    """True inverse of the lens model (fixed point to convergence). Returns x, y, ok."""
    k1, k2, p1, p2, k3 = d
    x, y = xd.copy(), yd.copy()
    for _ in range(iters):
        r2 = x * x + y * y
        icd = 1.0 / (1 + ((k3 * r2 + k2) * r2 + k1) * r2)
        dx = 2 * p1 * x * y + p2 * (r2 + 2 * x * x)
        dy = p1 * (r2 + 2 * y * y) + 2 * p2 * x * y
        x = (xd - dx) * icd
        y = (yd - dy) * icd
    ex, ey = distort(x, y, d)
    ok = (np.abs(ex - xd) < 1e-9) & (np.abs(ey - yd) < 1e-9) & (x * x + y * y < 1.0)
    return x, y, ok


@dataclass
class SynthCamera:
    width: int
    height: int
    K: np.ndarray
    dist: np.ndarray
    g12: np.ndarray
    z_min: float = 1.5
    z_max: float = 2.4


@dataclass
class StereoScene:
    cam: SynthCamera
    img1: np.ndarray
    img2: np.ndarray
    facet1: np.ndarray          # facet id per image-1 pixel (-1 = no hit)
    planes: np.ndarray          # (K,3): Z = a X + b Y + c  in camera-1 coordinates
    tex: dict = field(repr=False, default_factory=dict)


def make_camera(width, height, rng, rot_deg=3.0, t=(-0.25, 0.02, -0.10)):
    fx = 0.559 * width
    K = np.array([[fx, 0, (width - 1) / 2.0], [0, fx, (height - 1) / 2.0], [0, 0, 1.0]])
    axis = rng.standard_normal(3)
    axis /= np.linalg.norm(axis)
    R = rodrigues(axis * math.radians(rot_deg))
    g12 = np.eye(4)
    g12[:3, :3] = R
    g12[:3, 3] = np.asarray(t, dtype=np.float64)
    return SynthCamera(width, height, K, SETTINGS_DIST.copy(), g12)


def make_planes(cam: SynthCamera, rng, pixels_ray=64, z0=1.65, max_slope_deg=33.0):
    """Planes tangent to the bowl Z = z0 + alpha*(X^2+Y^2); Z = max_k plane_k is convex."""
    fx = cam.K[0, 0]
    spacing = 3.5 * (2 * pixels_ray + 1)
    nx = max(2, int(round(cam.width / spacing)))
    ny = max(2, int(round(cam.height / spacing)))
    # extent (metres at depth ~2) of the region where keypoints may fall (norm radius 0.6)
    ext = 0.62 * 2.0
    # slope of the bowl at the rim stays below max_slope
    alpha = math.tan(math.radians(max_slope_deg)) / (2 * ext)
    alpha = min(alpha, (2.28 - z0) / (ext * ext))
    planes = []
    for iy in range(ny):
        for ix in range(nx):
            X0 = ((ix + 0.5) / nx - 0.5) * 2 * ext * min(1.0, cam.width / cam.height * 0.8)
            Y0 = ((iy + 0.5) / ny - 0.5) * 2 * ext * min(1.0, cam.height / cam.width * 1.1)
            X0 += rng.uniform(-0.1, 0.1) * ext / nx
            Y0 += rng.uniform(-0.1, 0.1) * ext / ny
            a, b = 2 * alpha * X0, 2 * alpha * Y0
            c = z0 + alpha * (X0 * X0 + Y0 * Y0) - a * X0 - b * Y0
            planes.append((a, b, c))
    return np.array(planes)


def make_texture(cam: SynthCamera, rng, n_waves=24):
    fx = cam.K[0, 0]
    m_per_px = 2.0 / fx
    lam = np.exp(rng.uniform(math.log(6.0), math.log(96.0), n_waves)) * m_per_px
    dirs = rng.standard_normal((n_waves, 3))
    dirs[:, 2] *= 0.3
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    kvec = dirs * (2 * math.pi / lam)[:, None]
    return {"k": kvec, "phase": rng.uniform(0, 2 * math.pi, n_waves),
            "amp": rng.uniform(0.5, 1.0, n_waves) * (lam / lam.max()) ** 0.35}


def texture_value(tex, X):
    s = np.zeros(X.shape[0])
    for k, ph, a in zip(tex["k"], tex["phase"], tex["amp"]):
        s += a * np.sin(X @ k + ph)
    s /= math.sqrt(0.5 * np.sum(tex["amp"] ** 2))
    return 127.5 + 105.0 * np.tanh(0.8 * s)


def _raycast(planes, o, d):
    """First hit of rays o + s*d with {Z >= a X + b Y + c for all k}. d: (n,3)."""
    a, b, c = planes[:, 0], planes[:, 1], planes[:, 2]
    A = o[2] - a * o[0] - b * o[1] - c                                  # (K,)
    B = d[:, 2:3] - d[:, 0:1] * a[None, :] - d[:, 1:2] * b[None, :]     # (n,K)
    with np.errstate(divide="ignore", invalid="ignore"):
        s = np.where(B > 1e-9, -A[None, :] / B, -np.inf)
    k = np.argmax(s, axis=1)
    sk = s[np.arange(s.shape[0]), k]
    ok = np.isfinite(sk) & (sk > 0) & np.all(B > 1e-9, axis=1)
    X = o[None, :] + sk[:, None] * d
    return X, k, ok


def render(cam: SynthCamera, planes, tex, view: int, noise_sigma=0.0, rng=None, rows_per_chunk=64):
    W, H = cam.width, cam.height
    fx, fy, cx, cy = cam.K[0, 0], cam.K[1, 1], cam.K[0, 2], cam.K[1, 2]
    R, t = cam.g12[:3, :3], cam.g12[:3, 3]
    img = np.zeros((H, W), np.uint8)
    facet = np.full((H, W), -1, np.int32)
    us = np.arange(W, dtype=np.float64)
    for y0 in range(0, H, rows_per_chunk):
        y1 = min(H, y0 + rows_per_chunk)
        uu, vv = np.meshgrid(us, np.arange(y0, y1, dtype=np.float64))
        xd = ((uu - cx) / fx).ravel()
        yd = ((vv - cy) / fy).ravel()
        x, y, ok = undistort_exact(xd, yd, cam.dist)
        d = np.stack([x, y, np.ones_like(x)], axis=1)
        if view == 1:
            o = np.zeros(3)
        else:
            o = -R.T @ t
            d = d @ R          # rows: R^T d
        X, k, hit = _raycast(planes, o, d)
        ok &= hit
        val = texture_value(tex, X)
        if noise_sigma > 0:
            val = val + rng.normal(0, noise_sigma, val.shape)
        val = np.where(ok, np.clip(np.rint(val), 0, 255), 0)
        img[y0:y1] = val.reshape(y1 - y0, W).astype(np.uint8)
        facet[y0:y1] = np.where(ok, k, -1).reshape(y1 - y0, W)
    return img, facet


def _scene_cache_path(width, height, seed, pixels_ray, noise_sigma):
    """Rendering a 4K pair by exact ray casting takes ~30 s of numpy: scenes are cached (compressed pickle) under
    $FM3D_SCENE_CACHE or <repo>/tools/_cache when that directory exists.  The key carries a digest of this file, so a
    cached scene can never come from other generator code."""
    import hashlib
    import os
    root = os.environ.get("FM3D_SCENE_CACHE") or os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "_cache")
    if not os.path.isdir(root):
        return None
    with open(os.path.abspath(__file__), "rb") as f:
        digest = hashlib.sha256(f.read()).hexdigest()[:12]
    return os.path.join(root, f"scene_{width}x{height}_s{seed}_r{pixels_ray}_n{noise_sigma}_{digest}.pkl.gz")


def make_scene(width, height, seed, pixels_ray=64, noise_sigma=0.0) -> StereoScene:
    import gzip
    import os
    import pickle
    path = _scene_cache_path(width, height, seed, pixels_ray, noise_sigma) if width * height >= 1920 * 1080 else None
    if path and os.path.exists(path):
        try:
            with gzip.open(path, "rb") as f:
                return pickle.load(f)
        except Exception:
            pass
    scene = _make_scene(width, height, seed, pixels_ray, noise_sigma)
    if path:
        try:
            tmp = path + f".{os.getpid()}.tmp"
            with gzip.open(tmp, "wb", compresslevel=1) as f:
                pickle.dump(scene, f, protocol=4)
            os.replace(tmp, path)
        except Exception:
            pass
    return scene


def _make_scene(width, height, seed, pixels_ray=64, noise_sigma=0.0) -> StereoScene:
    rng = np.random.default_rng(seed)
    cam = make_camera(width, height, rng)
    planes = make_planes(cam, rng, pixels_ray)
    tex = make_texture(cam, rng)
    img1, facet1 = render(cam, planes, tex, 1, noise_sigma, rng)
    img2, _ = render(cam, planes, tex, 2, noise_sigma, rng)
    return StereoScene(cam, img1, img2, facet1, planes, tex)


def project(cam: SynthCamera, X, view: int):
    X = np.asarray(X, dtype=np.float64).reshape(-1, 3)
    if view == 2:
        X = X @ cam.g12[:3, :3].T + cam.g12[:3, 3]
    x, y = X[:, 0] / X[:, 2], X[:, 1] / X[:, 2]
    xd, yd = distort(x, y, cam.dist)
    return np.stack([cam.K[0, 0] * xd + cam.K[0, 2], cam.K[1, 1] * yd + cam.K[1, 2]], axis=1)


def make_keypoints(scene: StereoScene, n, seed, pixels_ray=64, margin2=24, kp_noise=0.0,
                   max_norm_radius=0.6):
    """n ground-truth features: image-1 pixel, image-2 pixel, 3-D point, facet normal.
    A feature is kept only if its whole disc lies on one facet and its image-2 footprint
    stays `margin2` px inside image 2."""
    rng = np.random.default_rng(seed)
    cam = scene.cam
    W, H = cam.width, cam.height
    fx, cx, cy = cam.K[0, 0], cam.K[0, 2], cam.K[1, 2]
    r = pixels_ray
    ang = np.linspace(0, 2 * math.pi, 16, endpoint=False)
    ring = np.stack([np.cos(ang), np.sin(ang)], axis=1) * (r + 2)
    rad_px = max_norm_radius * 0.9 * fx - r   # distorted radius of norm-radius 0.6 is ~0.545
    rad_px = max(rad_px, 8.0)
    out_p1, out_p2, out_X, out_n, out_f = [], [], [], [], []
    tries = 0
    while len(out_p1) < n and tries < 200:
        tries += 1
        m = max(4 * (n - len(out_p1)), 64)
        u = rng.uniform(max(r + 16, cx - rad_px), min(W - r - 17, cx + rad_px), m)
        v = rng.uniform(max(r + 16, cy - rad_px), min(H - r - 17, cy + rad_px), m)
        keep = (u - cx) ** 2 + (v - cy) ** 2 <= rad_px ** 2 if rad_px > r else np.ones(m, bool)
        u, v = u[keep], v[keep]
        # same facet on the ring
        f0 = scene.facet1[np.rint(v).astype(int), np.rint(u).astype(int)]
        same = f0 >= 0
        for dx, dy in ring:
            uu = np.clip(np.rint(u + dx).astype(int), 0, W - 1)
            vv = np.clip(np.rint(v + dy).astype(int), 0, H - 1)
            same &= scene.facet1[vv, uu] == f0
        u, v, f0 = u[same], v[same], f0[same]
        if u.size == 0:
            continue
        # exact 3-D point on the facet
        x, y, ok = undistort_exact((u - cx) / fx, (v - cy) / cam.K[1, 1], cam.dist)
        a, b, c = scene.planes[f0, 0], scene.planes[f0, 1], scene.planes[f0, 2]
        Z = c / (1 - a * x - b * y)
        X = np.stack([x * Z, y * Z, Z], axis=1)
        p2 = project(cam, X, 2)
        ok &= (Z > cam.z_min + 0.02) & (Z < cam.z_max - 0.02)
        # ring footprint in image 2 (through the facet plane)
        for dx, dy in ring:
            xr, yr, okr = undistort_exact((u + dx - cx) / fx, (v + dy - cy) / cam.K[1, 1], cam.dist)
            Zr = c / (1 - a * xr - b * yr)
            q = project(cam, np.stack([xr * Zr, yr * Zr, Zr], axis=1), 2)
            ok &= okr & (q[:, 0] >= margin2) & (q[:, 0] <= W - 1 - margin2)
            ok &= (q[:, 1] >= margin2) & (q[:, 1] <= H - 1 - margin2)
        nrm = np.stack([-a, -b, np.ones_like(a)], axis=1)
        nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
        for arr, src in ((out_p1, np.stack([u, v], 1)), (out_p2, p2), (out_X, X), (out_n, nrm),
                         (out_f, f0)):
            arr.extend(src[ok])
    if len(out_p1) < n:
        raise RuntimeError(f"could only place {len(out_p1)} of {n} keypoints")
    p1 = np.array(out_p1[:n])
    p2 = np.array(out_p2[:n])
    if kp_noise > 0:
        p2 = p2 + rng.normal(0, kp_noise, p2.shape)
    return {"kp1": p1.astype(np.float32), "kp2": p2.astype(np.float32),
            "X": np.array(out_X[:n]), "normal": np.array(out_n[:n]),
            "facet": np.array(out_f[:n], np.int32)}


def make_float_descriptors(n_query, n_train_extra, seed, dim=128, inlier_frac=0.8, sigma=6.0):
    """SIFT-like integer-valued float descriptors in [0,255].
    Returns query (n_query,dim) f32, train (n_query+extra,dim) f32, gt (n_query,) train index
    of every inlier query (-1 for outlier queries)."""
    rng = np.random.default_rng(seed)
    nt = n_query + n_train_extra
    train = np.floor(rng.gamma(2.0, 18.0, (nt, dim))).clip(0, 255).astype(np.float32)
    perm = rng.permutation(nt)[:n_query]
    inl = rng.random(n_query) < inlier_frac
    query = np.clip(train[perm] + np.rint(rng.normal(0, sigma, (n_query, dim))), 0, 255)
    fresh = np.floor(rng.gamma(2.0, 18.0, (n_query, dim))).clip(0, 255)
    query = np.where(inl[:, None], query, fresh).astype(np.float32)
    gt = np.where(inl, perm, -1).astype(np.int32)
    return query, train, gt


def make_binary_descriptors(n_query, n_train_extra, seed, nbytes=32, inlier_frac=0.8):
    """ORB-like binary descriptors; inlier queries flip 10-25 bits of their train row."""
    rng = np.random.default_rng(seed)
    nt = n_query + n_train_extra
    train = rng.integers(0, 256, (nt, nbytes), dtype=np.uint8)
    perm = rng.permutation(nt)[:n_query]
    inl = rng.random(n_query) < inlier_frac
    query = rng.integers(0, 256, (n_query, nbytes), dtype=np.uint8)
    bits = np.unpackbits(train[perm], axis=1)
    nflip = rng.integers(10, 26, n_query)
    for i in np.nonzero(inl)[0]:
        pos = rng.choice(nbytes * 8, nflip[i], replace=False)
        bits[i, pos] ^= 1
    query = np.where(inl[:, None], np.packbits(bits, axis=1), query).astype(np.uint8)
    gt = np.where(inl, perm, -1).astype(np.int32)
    return query, train, gt


def make_stereo_case(width, height, n_features, seed, pixels_ray=64, n_distractors=None,
                     noise_sigma=0.0, kp_noise=0.0):
    """Full pipeline input: scene + keypoints of both frames + float descriptors such that
    query i (frame A) truly matches train gt[i] (frame B)."""
    scene = make_scene(width, height, seed, pixels_ray, noise_sigma)
    kps = make_keypoints(scene, n_features, seed + 1, pixels_ray, kp_noise=kp_noise)
    nd = n_features // 5 if n_distractors is None else n_distractors
    q, t, gt = make_float_descriptors(n_features, nd, seed + 2)
    rng = np.random.default_rng(seed + 3)
    nt = t.shape[0]
    # frame-B keypoints: train row gt[i] carries the true image-2 position of query i
    kp2 = np.stack([rng.uniform(0, width - 1, nt), rng.uniform(0, height - 1, nt)], 1)
    kp2 = kp2.astype(np.float32)
    has = gt >= 0
    kp2[gt[has]] = kps["kp2"][has]
    return {"scene": scene, "kp1": kps["kp1"], "kp2": kp2, "desc1": q, "desc2": t, "gt": gt,
            "X": kps["X"], "normal": kps["normal"], "facet": kps["facet"],
            "kp2_true": kps["kp2"]}
