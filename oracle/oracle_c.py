"""TEST INFRASTRUCTURE -- ctypes binding of oracle/fm3d_oracle.c (the plain-C oracle).

Only tests/, bench.py's cpu_baseline / --impl reference legs and __graft_entry__.smoke() may
import this.  `build()` compiles the library with gcc if it is missing or stale.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libfm3d_oracle.so")
_lib = None

c_dp = C.POINTER(C.c_double)
c_fp = C.POINTER(C.c_float)
c_ip = C.POINTER(C.c_int32)
c_bp = C.POINTER(C.c_uint8)


def build(force=False):
    src = os.path.join(_HERE, "fm3d_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s"] + (["-s"] if force else []))
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.orc_optimize_normals.restype = C.c_longlong
        _lib.orc_optimize_normals2.restype = C.c_longlong
        _lib.orc_pyramid_bytes.restype = C.c_size_t
    return _lib


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a, t):
    return a.ctypes.data_as(t) if a is not None else None


def knn2_f32(q, t, threads=1):
    q = np.ascontiguousarray(q, np.float32)
    t = np.ascontiguousarray(t, np.float32)
    nq, nt = q.shape[0], t.shape[0]
    dim = q.shape[1] if q.ndim == 2 and nq else (t.shape[1] if nt else 0)
    idx = np.empty((nq, 2), np.int32)
    dist = np.empty((nq, 2), np.float32)
    lib().orc_knn2_f32(_p(q, c_fp), nq, _p(t, c_fp), nt, dim, _p(idx, c_ip), _p(dist, c_fp), threads)
    return idx, dist


def knn2_hamming(q, t, threads=1):
    q = np.ascontiguousarray(q, np.uint8)
    t = np.ascontiguousarray(t, np.uint8)
    nq, nt = q.shape[0], t.shape[0]
    nb = q.shape[1] if nq else (t.shape[1] if nt else 0)
    idx = np.empty((nq, 2), np.int32)
    dist = np.empty((nq, 2), np.float32)
    lib().orc_knn2_hamming(_p(q, c_bp), nq, _p(t, c_bp), nt, nb, _p(idx, c_ip), _p(dist, c_fp), threads)
    return idx, dist


def nndr_filter(idx, dist, eps):
    idx = np.ascontiguousarray(idx, np.int32)
    dist = np.ascontiguousarray(dist, np.float32)
    nq = idx.shape[0]
    qi = np.empty(nq, np.int32)
    ti = np.empty(nq, np.int32)
    d = np.empty(nq, np.float32)
    n = lib().orc_nndr_filter(_p(idx, c_ip), _p(dist, c_fp), nq, C.c_double(eps), _p(qi, c_ip), _p(ti, c_ip), _p(d, c_fp))
    return qi[:n].copy(), ti[:n].copy(), d[:n].copy()


def undistort_points(K, dist, pts):
    pts = _d(pts).reshape(-1, 2)
    out = np.empty_like(pts)
    lib().orc_undistort_points(_p(_d(K), c_dp), _p(_d(dist), c_dp), _p(pts, c_dp), pts.shape[0], _p(out, c_dp))
    return out


def project_points(K, dist, g12, view, X):
    X = _d(X).reshape(-1, 3)
    out = np.empty((X.shape[0], 2))
    lib().orc_project_points(_p(_d(K), c_dp), _p(_d(dist), c_dp), _p(_d(g12), c_dp), view, _p(X, c_dp), X.shape[0], _p(out, c_dp))
    return out


def triangulate(K, dist, g12, zmin, zmax, kp1, kp2, qidx=None, tidx=None):
    kp1 = np.ascontiguousarray(kp1, np.float32)
    kp2 = np.ascontiguousarray(kp2, np.float32)
    if qidx is not None:
        qidx = np.ascontiguousarray(qidx, np.int32)
        tidx = np.ascontiguousarray(tidx, np.int32)
        n = qidx.shape[0]
    else:
        n = kp1.shape[0]
    xyz_all = np.empty((n, 3))
    xyz = np.empty((n, 3))
    mask = np.empty(n, np.uint8)
    ninl = lib().orc_triangulate(_p(_d(K), c_dp), _p(_d(dist), c_dp), _p(_d(g12), c_dp), C.c_double(zmin), C.c_double(zmax),
                                 _p(kp1, c_fp), _p(kp2, c_fp), _p(qidx, c_ip), _p(tidx, c_ip), n,
                                 _p(xyz_all, c_dp), _p(mask, c_bp), _p(xyz, c_dp))
    return xyz_all, mask, xyz[:ninl].copy()


def pyrdown(img):
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    out = np.empty(((h + 1) // 2, (w + 1) // 2), np.uint8)
    lib().orc_pyrdown(_p(img, c_bp), w, h, _p(out, c_bp))
    return out


def build_pyramid(img, levels):
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    nb = lib().orc_pyramid_bytes(w, h, levels)
    out = np.empty(nb, np.uint8)
    lib().orc_build_pyramid(_p(img, c_bp), w, h, levels, _p(out, c_bp))
    return out


def pyramid_levels(buf, w, h, levels):
    out, off = [], 0
    for _ in range(levels + 1):
        out.append(buf[off:off + w * h].reshape(h, w))
        off += w * h
        w, h = (w + 1) // 2, (h + 1) // 2
    return out


def disc_pixels(K, dist, P, r, w, h):
    pix = np.empty(((2 * r + 1) ** 2, 2))
    m = lib().orc_disc_pixels(_p(_d(K), c_dp), _p(_d(dist), c_dp), _p(_d(P), c_dp), r, w, h, _p(pix, c_dp))
    return pix[:m].copy()


def optimize_normals(K, dist, g12, zmin, zmax, img1, img2, levels, xyz, pixels_ray,
                     eps_lmmin, penalty_mode=0, patience=100, as_written=0, threads=1,
                     pyr1=None, pyr2=None, cost_mode=0):
    h, w = img1.shape
    if pyr1 is None:
        pyr1 = build_pyramid(img1, levels)
        pyr2 = build_pyramid(img2, levels)
    xyz = _d(xyz).reshape(-1, 3)
    n = xyz.shape[0]
    normals = np.empty((n, 3))
    status = np.empty(n, np.int32)
    nfev = np.empty((n, levels + 1), np.int32)
    npen = np.empty(n, np.int32)
    cost = np.empty(n)
    m = np.empty(n, np.int32)
    pe = lib().orc_optimize_normals2(_p(_d(K), c_dp), _p(_d(dist), c_dp), _p(_d(g12), c_dp), C.c_double(zmin), C.c_double(zmax),
                                    _p(pyr1, c_bp), _p(pyr2, c_bp), w, h, levels, _p(xyz, c_dp), n, pixels_ray,
                                    C.c_double(eps_lmmin), penalty_mode, cost_mode, patience, as_written, threads,
                                    _p(normals, c_dp), _p(status, c_ip), _p(nfev, c_ip), _p(npen, c_ip), _p(cost, c_dp), _p(m, c_ip))
    return {"normals": normals, "status": status, "nfev": nfev, "npenalty": npen, "cost": cost,
            "m": m, "pixel_evals": int(pe)}


def evaluate_cost(K, dist, g12, zmin, zmax, img1, img2, levels, xyz, phi_theta, pixels_ray,
                  level, penalty_mode=0, cost_mode=0):
    h, w = img1.shape
    pyr1 = build_pyramid(img1, levels)
    pyr2 = build_pyramid(img2, levels)
    xyz = _d(xyz).reshape(-1, 3)
    pt = _d(phi_theta).reshape(-1, 2)
    n = xyz.shape[0]
    cost = np.empty(n)
    m = np.empty(n, np.int32)
    status = np.empty(n, np.int32)
    lib().orc_evaluate_cost2(_p(_d(K), c_dp), _p(_d(dist), c_dp), _p(_d(g12), c_dp), C.c_double(zmin), C.c_double(zmax),
                             _p(pyr1, c_bp), _p(pyr2, c_bp), w, h, levels, _p(xyz, c_dp), _p(pt, c_dp), n, pixels_ray, level,
                             penalty_mode, cost_mode, _p(cost, c_dp), _p(m, c_ip), _p(status, c_ip))
    return cost, m, status


def feature_frames(xyz, normals, gravity):
    xyz = _d(xyz).reshape(-1, 3)
    normals = _d(normals).reshape(-1, 3)
    n = xyz.shape[0]
    frames = np.empty((n, 4, 4))
    lib().orc_feature_frames(_p(xyz, c_dp), _p(normals, c_dp), n, _p(_d(gravity), c_dp), _p(frames, c_dp))
    return frames


def patch_size(eps_m, cm_per_px):
    return lib().orc_patch_size(C.c_double(eps_m), C.c_double(cm_per_px))


def extract_patches(K, dist, img1, frames, eps_m, cm_per_px, want_points=True, threads=1):
    img1 = np.ascontiguousarray(img1, np.uint8)
    h, w = img1.shape
    frames = _d(frames).reshape(-1, 4, 4)
    n = frames.shape[0]
    S = patch_size(eps_m, cm_per_px)
    patches = np.empty((n, S, S), np.uint8)
    ip = np.empty((n, S * S, 2)) if want_points else None
    lib().orc_extract_patches(_p(_d(K), c_dp), _p(_d(dist), c_dp), _p(img1, c_bp), w, h, _p(frames, c_dp), n,
                              C.c_double(eps_m), C.c_double(cm_per_px), _p(patches, c_bp), _p(ip, c_dp), threads)
    return patches, ip


def lmmin_expsin(t, y, x0, epsilon=1e-10, patience=100, minpack_mode=0):
    t, y = _d(t), _d(y)
    x = _d(x0).copy()
    nfev, info = C.c_int(0), C.c_int(0)
    lib().orc_lmmin_expsin(_p(t, c_dp), _p(y, c_dp), t.size, _p(x, c_dp), C.c_double(epsilon), patience, minpack_mode,
                           C.byref(nfev), C.byref(info))
    return x, nfev.value, info.value


def lmmin_exp2(t, y, x0, epsilon=1e-10, patience=100, minpack_mode=0):
    t, y = _d(t), _d(y)
    x = _d(x0).copy()
    nfev, info = C.c_int(0), C.c_int(0)
    lib().orc_lmmin_exp2(_p(t, c_dp), _p(y, c_dp), t.size, _p(x, c_dp), C.c_double(epsilon), patience, minpack_mode,
                         C.byref(nfev), C.byref(info))
    return x, nfev.value, info.value
