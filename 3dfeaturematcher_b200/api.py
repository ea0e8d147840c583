"""ctypes binding of libfm3d.so (include/fm3d.h) -- the only way Python reaches the GPU path.

There is no CPU implementation here: if libfm3d.so is missing or no sm_100 device is
usable, construction of `Context` raises.  numpy arrays are used for the host-pointer entry
points; the `*_dev` methods take raw device addresses (e.g. `torch.Tensor.data_ptr()`), enqueue
on the context's stream and do not synchronise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# FM3D_LIB: another build of the same library (tools/build_variant.py: A/B measurements of kernel variants)
LIB_PATH = os.environ.get("FM3D_LIB") or os.path.join(_HERE, "libfm3d.so")

FEAT_OK, FEAT_NO_PIXELS, FEAT_ABORT_BBOX, FEAT_ABORT_PIXEL, FEAT_ABORT_NAN = 0, 1, 2, 3, 4
PENALTY_FABS, PENALTY_INT_ABS, PENALTY_OFF = 0, 1, 2

_dp, _fp, _ip, _bp = (C.POINTER(C.c_double), C.POINTER(C.c_float), C.POINTER(C.c_int32),
                      C.POINTER(C.c_uint8))

# every symbol include/fm3d.h declares (tests/test_abi.py checks the library exports them all)
SYMBOLS = [
    "fm3d_version", "fm3d_ctx_create", "fm3d_ctx_destroy", "fm3d_last_error", "fm3d_sync",
    "fm3d_stream", "fm3d_device_info", "fm3d_set_option", "fm3d_get_option",
    "fm3d_get_launch_counters", "fm3d_set_camera", "fm3d_set_g12", "fm3d_compose_g12",
    "fm3d_match_knn2_f32", "fm3d_match_knn2_hamming", "fm3d_match_nndr_f32",
    "fm3d_match_nndr_hamming", "fm3d_match_knn2_f32_dev", "fm3d_match_knn2_hamming_dev",
    "fm3d_nndr_filter_dev", "fm3d_triangulate", "fm3d_triangulate_dev", "fm3d_undistort_points",
    "fm3d_set_images", "fm3d_set_images2", "fm3d_set_images_dev", "fm3d_get_pyramid_level", "fm3d_optimize_normals",
    "fm3d_optimize_normals_dev", "fm3d_evaluate_normals", "fm3d_get_normals_stats",
    "fm3d_sweep_normals", "fm3d_sweep_normals_dev",
    "fm3d_disc_pixels", "fm3d_plane_points", "fm3d_sample_pixels", "fm3d_project_to_image2",
    "fm3d_circular_neighborhoods", "fm3d_feature_frames",
    "fm3d_feature_frames_dev", "fm3d_patch_size", "fm3d_extract_patches",
    "fm3d_extract_patches_dev", "fm3d_project_groups", "fm3d_square_neighborhoods",
    "fm3d_describe_patches_sift", "fm3d_describe_patches_sift_dev",
    "fm3d_detect_fast", "fm3d_detect_fast_dev", "fm3d_detect_sift", "fm3d_detect_and_describe_sift", "fm3d_detect_orb",
    "fm3d_describe_keypoints_sift", "fm3d_describe_keypoints_sift_dev", "fm3d_describe_keypoints_sift_oct", "fm3d_sift_base_image_dev",
    "fm3d_describe_keypoints_brisk", "fm3d_describe_keypoints_brisk_dev",
    "fm3d_describe_keypoints_orb", "fm3d_describe_keypoints_orb_dev",
    "fm3d_describe_patches_orb", "fm3d_describe_patches_orb_dev",
    "fm3d_comm_unique_id", "fm3d_comm_init_rank", "fm3d_comm_init_all", "fm3d_comm_info", "fm3d_comm_destroy",
    "fm3d_broadcast_dev", "fm3d_allgather_dev", "fm3d_broadcast_all_dev", "fm3d_allgather_all_dev",
    "fm3d_shard_block_layout", "fm3d_pack_shard_dev",
    "fm3d_dev_malloc", "fm3d_dev_free", "fm3d_copy_h2d", "fm3d_copy_d2h",
]


class Fm3dError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libfm3d error {code}: {msg}")
        self.code = code


_lib = None


def load_library():
    """Loads libfm3d.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -m 3dfeaturematcher_b200.build` "
                               "(__graft_entry__.build()); there is no CPU fallback")
        lib = C.CDLL(LIB_PATH)
        lib.fm3d_last_error.restype = C.c_char_p
        lib.fm3d_stream.restype = C.c_void_p
        lib.fm3d_patch_size.argtypes = [C.c_double, C.c_double]
        _lib = lib
    return _lib


def _arr(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


def _ptr(a, t):
    return a.ctypes.data_as(t) if a is not None else None


def compose_g12(T1, T2, rod1, rod2, rodIC, tIC):
    """SingleCameraTriangulator::setg12 host arithmetic (no GPU needed)."""
    out = np.empty(16)
    args = [_arr(x, np.float64) for x in (T1, T2, rod1, rod2, rodIC, tIC)]
    rc = load_library().fm3d_compose_g12(*[_ptr(a, _dp) for a in args], _ptr(out, _dp))
    if rc:
        raise Fm3dError(rc, "fm3d_compose_g12")
    return out.reshape(4, 4)


def patch_size(epsilon_m, cm_per_pixel):
    return load_library().fm3d_patch_size(float(epsilon_m), float(cm_per_pixel))


class ShardLayout(C.Structure):
    """fm3d_shard_layout (include/fm3d.h): byte offsets of the parts of one gather block."""
    _fields_ = [("cap", C.c_int), ("off_qidx", C.c_size_t), ("off_tidx", C.c_size_t), ("off_dist", C.c_size_t),
                ("off_src", C.c_size_t), ("off_normals", C.c_size_t), ("off_status", C.c_size_t), ("bytes", C.c_size_t)]


COMM_ID_BYTES = 128


def comm_unique_id():
    """fm3d_comm_unique_id: 128 bytes rank 0 creates and hands to the other ranks (any channel)."""
    buf = (C.c_uint8 * COMM_ID_BYTES)()
    rc = load_library().fm3d_comm_unique_id(buf)
    if rc:
        raise Fm3dError(rc, "fm3d_comm_unique_id failed (NCCL not loadable?)")
    return bytes(buf)


def shard_block_layout(cap):
    L = ShardLayout()
    rc = load_library().fm3d_shard_block_layout(int(cap), C.byref(L))
    if rc:
        raise Fm3dError(rc, "fm3d_shard_block_layout failed")
    return L


def unpack_shard_blocks(blocks, layout):
    """Host view of gathered blocks (numpy uint8, world x layout.bytes): the valid rows of all ranks in rank order."""
    world = blocks.shape[0]
    hdr = np.ascontiguousarray(blocks[:, :20]).view(np.int32).reshape(world, 5)
    cap = layout.cap
    out = {k: [] for k in ("qidx", "tidx", "dist", "src", "normals", "status")}
    for r in range(world):
        nm, ni = int(hdr[r, 0]), int(hdr[r, 1])
        b = blocks[r]
        out["qidx"].append(b[layout.off_qidx:layout.off_qidx + 4 * cap].view(np.int32)[:nm])
        out["tidx"].append(b[layout.off_tidx:layout.off_tidx + 4 * cap].view(np.int32)[:nm])
        out["dist"].append(b[layout.off_dist:layout.off_dist + 4 * cap].view(np.float32)[:nm])
        out["src"].append(b[layout.off_src:layout.off_src + 4 * cap].view(np.int32)[:ni])
        out["normals"].append(b[layout.off_normals:layout.off_normals + 24 * cap].view(np.float64).reshape(cap, 3)[:ni])
        out["status"].append(b[layout.off_status:layout.off_status + 4 * cap].view(np.int32)[:ni])
    return {k: np.concatenate(v) for k, v in out.items()}, hdr


class Context:
    """One fm3d_ctx: one GPU, one stream."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        self._h = C.c_void_p()
        rc = self.lib.fm3d_ctx_create(int(device), C.byref(self._h))
        if rc:
            raise Fm3dError(rc, f"fm3d_ctx_create(device={device}) failed: no usable sm_100 GPU "
                                "(libfm3d has no CPU path)")
        self.device = device
        self.levels = None
        self.shape = None

    def close(self):
        if self._h:
            self.lib.fm3d_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc:
            raise Fm3dError(rc, self.lib.fm3d_last_error(self._h).decode())

    # ------------------------------------------------------------------ several GPUs (one process per GPU)
    def comm_init_rank(self, unique_id: bytes, nranks: int, rank: int):
        buf = (C.c_uint8 * COMM_ID_BYTES).from_buffer_copy(unique_id)
        self._ck(self.lib.fm3d_comm_init_rank(self._h, buf, int(nranks), int(rank)))

    def comm_info(self):
        n, r, v = C.c_int(), C.c_int(), C.c_int()
        self._ck(self.lib.fm3d_comm_info(self._h, C.byref(n), C.byref(r), C.byref(v)))
        return {"nranks": n.value, "rank": r.value, "nccl_version": v.value}

    def broadcast_dev(self, buf_ptr, nbytes, root=0):
        self._ck(self.lib.fm3d_broadcast_dev(self._h, C.c_void_p(buf_ptr), C.c_size_t(nbytes), int(root)))

    def allgather_dev(self, send_ptr, recv_ptr, bytes_per_rank):
        self._ck(self.lib.fm3d_allgather_dev(self._h, C.c_void_p(send_ptr), C.c_void_p(recv_ptr), C.c_size_t(bytes_per_rank)))

    def pack_shard_dev(self, cap, rank, query_offset, n_match_dev, n_inl_dev, qidx, tidx, dist, src, normals, status, block):
        self._ck(self.lib.fm3d_pack_shard_dev(self._h, int(cap), int(rank), int(query_offset), C.c_void_p(n_match_dev),
                                              C.c_void_p(n_inl_dev), C.c_void_p(qidx), C.c_void_p(tidx), C.c_void_p(dist),
                                              C.c_void_p(src), C.c_void_p(normals), C.c_void_p(status), C.c_void_p(block)))

    # ------------------------------------------------------------------ context
    def sync(self):
        self._ck(self.lib.fm3d_sync(self._h))

    @property
    def stream(self):
        return self.lib.fm3d_stream(self._h)

    def device_info(self):
        sm, ma, mi = C.c_int(), C.c_int(), C.c_int()
        name = C.create_string_buffer(256)
        self._ck(self.lib.fm3d_device_info(self._h, C.byref(sm), C.byref(ma), C.byref(mi), name, 256))
        return {"sm_count": sm.value, "cc": (ma.value, mi.value), "name": name.value.decode()}

    def set_option(self, key, value):
        self._ck(self.lib.fm3d_set_option(self._h, key.encode(), C.c_double(value)))

    def get_option(self, key):
        v = C.c_double()
        self._ck(self.lib.fm3d_get_option(self._h, key.encode(), C.byref(v)))
        return v.value

    def launch_counters(self):
        k, c = C.c_int64(), C.c_int64()
        self._ck(self.lib.fm3d_get_launch_counters(self._h, C.byref(k), C.byref(c)))
        return k.value, c.value

    # ------------------------------------------------------------------ camera
    def set_camera(self, K, dist, z_min, z_max):
        K, dist = _arr(K, np.float64).reshape(9), _arr(dist, np.float64).reshape(5)
        self._ck(self.lib.fm3d_set_camera(self._h, _ptr(K, _dp), _ptr(dist, _dp), C.c_double(z_min), C.c_double(z_max)))

    def set_g12(self, g12):
        g = _arr(g12, np.float64).reshape(16)
        self._ck(self.lib.fm3d_set_g12(self._h, _ptr(g, _dp)))

    # ------------------------------------------------------------------ matching
    def match_knn2_f32(self, q, t):
        q, t = _arr(q, np.float32), _arr(t, np.float32)
        nq, nt = q.shape[0], t.shape[0]
        dim = q.shape[1] if q.ndim == 2 and q.shape[1] else (t.shape[1] if t.ndim == 2 else 1)
        idx, dist = np.full((nq, 2), -1, np.int32), np.full((nq, 2), np.inf, np.float32)
        self._ck(self.lib.fm3d_match_knn2_f32(self._h, _ptr(q, _fp), nq, _ptr(t, _fp), nt, max(dim, 1), _ptr(idx, _ip), _ptr(dist, _fp)))
        return idx, dist

    def match_knn2_hamming(self, q, t):
        q, t = _arr(q, np.uint8), _arr(t, np.uint8)
        nq, nt = q.shape[0], t.shape[0]
        nb = q.shape[1] if q.ndim == 2 and q.shape[1] else (t.shape[1] if t.ndim == 2 else 32)
        idx, dist = np.full((nq, 2), -1, np.int32), np.full((nq, 2), np.inf, np.float32)
        self._ck(self.lib.fm3d_match_knn2_hamming(self._h, _ptr(q, _bp), nq, _ptr(t, _bp), nt, nb, _ptr(idx, _ip), _ptr(dist, _fp)))
        return idx, dist

    def match_nndr(self, q, t, eps, hamming=False, with_mutual=False):
        dt = np.uint8 if hamming else np.float32
        q, t = _arr(q, dt), _arr(t, dt)
        nq, nt = q.shape[0], t.shape[0]
        dim = q.shape[1] if q.ndim == 2 and q.shape[1] else (t.shape[1] if t.ndim == 2 else 1)
        qi, ti = np.empty(nq, np.int32), np.empty(nq, np.int32)
        d = np.empty(nq, np.float32)
        mu = np.zeros(nq, np.uint8) if with_mutual else None
        n = C.c_int(0)
        if hamming:
            rc = self.lib.fm3d_match_nndr_hamming(self._h, _ptr(q, _bp), nq, _ptr(t, _bp), nt, dim, C.c_double(eps),
                                                  _ptr(qi, _ip), _ptr(ti, _ip), _ptr(d, _fp), _ptr(mu, _bp), C.byref(n))
        else:
            rc = self.lib.fm3d_match_nndr_f32(self._h, _ptr(q, _fp), nq, _ptr(t, _fp), nt, max(dim, 1), C.c_double(eps),
                                              _ptr(qi, _ip), _ptr(ti, _ip), _ptr(d, _fp), _ptr(mu, _bp), C.byref(n))
        self._ck(rc)
        k = n.value
        out = (qi[:k].copy(), ti[:k].copy(), d[:k].copy())
        return out + (mu[:k].copy(),) if with_mutual else out

    # ------------------------------------------------------------------ triangulation
    def triangulate(self, kp1, kp2, qidx=None, tidx=None):
        kp1, kp2 = _arr(kp1, np.float32).reshape(-1, 2), _arr(kp2, np.float32).reshape(-1, 2)
        if qidx is not None:
            qidx, tidx = _arr(qidx, np.int32), _arr(tidx, np.int32)
            n = qidx.shape[0]
        else:
            n = kp1.shape[0]
        xyz_all, xyz = np.empty((n, 3)), np.empty((n, 3))
        mask, src = np.empty(n, np.uint8), np.empty(n, np.int32)
        ninl = C.c_int(0)
        self._ck(self.lib.fm3d_triangulate(self._h, _ptr(kp1, _fp), kp1.shape[0], _ptr(kp2, _fp), kp2.shape[0],
                                           _ptr(qidx, _ip), _ptr(tidx, _ip), n, _ptr(xyz_all, _dp), _ptr(mask, _bp),
                                           _ptr(xyz, _dp), _ptr(src, _ip), C.byref(ninl)))
        k = ninl.value
        return xyz_all, mask, xyz[:k].copy(), src[:k].copy()

    def undistort_points(self, pts):
        pts = _arr(pts, np.float64).reshape(-1, 2)
        out = np.empty_like(pts)
        self._ck(self.lib.fm3d_undistort_points(self._h, _ptr(pts, _dp), pts.shape[0], _ptr(out, _dp)))
        return out

    # ------------------------------------------------------------------ images
    def set_images(self, img1, img2, pyramids):
        img1, img2 = _arr(img1, np.uint8), _arr(img2, np.uint8)
        assert img1.shape == img2.shape and img1.ndim == 2
        h, w = img1.shape
        self._ck(self.lib.fm3d_set_images(self._h, _ptr(img1, _bp), _ptr(img2, _bp), w, h, w, int(pyramids)))
        self.levels, self.shape = int(pyramids), (h, w)

    def set_images_dev(self, img1_ptr, img2_ptr, w, h, stride, pyramids):
        self._ck(self.lib.fm3d_set_images_dev(self._h, C.c_void_p(img1_ptr), C.c_void_p(img2_ptr), w, h, stride, int(pyramids)))
        self.levels, self.shape = int(pyramids), (h, w)

    def get_pyramid_level(self, image, level):
        w, h = C.c_int(), C.c_int()
        self._ck(self.lib.fm3d_get_pyramid_level(self._h, image, level, None, C.byref(w), C.byref(h)))
        out = np.empty((h.value, w.value), np.uint8)
        self._ck(self.lib.fm3d_get_pyramid_level(self._h, image, level, _ptr(out, _bp), C.byref(w), C.byref(h)))
        return out

    # ------------------------------------------------------------------ normals
    def optimize_normals(self, xyz, pixels_ray, epsilon_lmmin=1e-10, penalty_mode=PENALTY_FABS):
        xyz = _arr(xyz, np.float64).reshape(-1, 3)
        n = xyz.shape[0]
        L1 = (self.levels or 0) + 1
        normals, status = np.empty((n, 3)), np.empty(n, np.int32)
        nfev, npen, cost = np.zeros((n, L1), np.int32), np.zeros(n, np.int32), np.empty(n)
        self._ck(self.lib.fm3d_optimize_normals(self._h, _ptr(xyz, _dp), n, int(pixels_ray), C.c_double(epsilon_lmmin),
                                                int(penalty_mode), _ptr(normals, _dp), _ptr(status, _ip), _ptr(nfev, _ip),
                                                _ptr(npen, _ip), _ptr(cost, _dp)))
        return {"normals": normals, "status": status, "nfev": nfev, "npenalty": npen, "cost": cost}

    def normals_stats(self):
        out = (C.c_int64 * 16)()
        self._ck(self.lib.fm3d_get_normals_stats(self._h, out))
        keys = ("passes_value", "passes_jacobian", "passes_fused", "fused_accepted", "passes_slow",
                "pixel_evals_value", "pixel_evals_jacobian", "features", "cycles_pixels", "cycles_barrier",
                "cycles_serial", "cycles_lm", "cycles_publish", "trials_memoized", "cycles_prologue", "cycles_level_setup")
        return dict(zip(keys, [int(v) for v in out]))

    def evaluate_normals(self, xyz, phi_theta, pixels_ray, level, penalty_mode=PENALTY_FABS):
        xyz = _arr(xyz, np.float64).reshape(-1, 3)
        pt = _arr(phi_theta, np.float64).reshape(-1, 2)
        n = xyz.shape[0]
        cost, m, status = np.empty(n), np.empty(n, np.int32), np.empty(n, np.int32)
        self._ck(self.lib.fm3d_evaluate_normals(self._h, _ptr(xyz, _dp), _ptr(pt, _dp), n, int(pixels_ray), int(level),
                                                int(penalty_mode), _ptr(cost, _dp), _ptr(m, _ip), _ptr(status, _ip)))
        return cost, m, status

    def sweep_normals(self, xyz, pixels_ray, level, n_phi, n_theta, dphi, dtheta, center_phi_theta=None,
                      penalty_mode=PENALTY_FABS, want_cost=True):
        xyz = _arr(xyz, np.float64).reshape(-1, 3)
        n = xyz.shape[0]
        pt = None if center_phi_theta is None else _arr(center_phi_theta, np.float64).reshape(-1, 2)
        cost = np.empty((n, n_phi, n_theta)) if want_cost else None
        best, bcost, status = np.empty(n, np.int32), np.empty(n), np.empty(n, np.int32)
        self._ck(self.lib.fm3d_sweep_normals(self._h, _ptr(xyz, _dp), None if pt is None else _ptr(pt, _dp), n, int(pixels_ray),
                                             int(level), int(penalty_mode), int(n_phi), int(n_theta), C.c_double(dphi),
                                             C.c_double(dtheta), None if cost is None else _ptr(cost, _dp), _ptr(best, _ip),
                                             _ptr(bcost, _dp), _ptr(status, _ip)))
        return {"cost": cost, "best_idx": best, "best_cost": bcost, "status": status}

    def sweep_normals_dev(self, xyz, n, pixels_ray, level, n_phi, n_theta, dphi, dtheta, status, center_phi_theta=None,
                          penalty_mode=PENALTY_FABS, cost=None, best_idx=None, best_cost=None):
        self._ck(self.lib.fm3d_sweep_normals_dev(self._h, C.c_void_p(xyz), C.c_void_p(center_phi_theta), n, int(pixels_ray),
                                                 int(level), int(penalty_mode), int(n_phi), int(n_theta), C.c_double(dphi),
                                                 C.c_double(dtheta), C.c_void_p(cost), C.c_void_p(best_idx),
                                                 C.c_void_p(best_cost), C.c_void_p(status)))

    def circular_neighborhoods(self, points, normals, epsilon_m, n_angles, n_rays):
        points = _arr(points, np.float64).reshape(-1, 3)
        normals = _arr(normals, np.float64).reshape(-1, 3).copy()
        n = points.shape[0]
        out = np.empty((n, n_angles * n_rays, 3))
        self._ck(self.lib.fm3d_circular_neighborhoods(self._h, _ptr(points, _dp), _ptr(normals, _dp), n, C.c_double(epsilon_m),
                                                      int(n_angles), int(n_rays), _ptr(out, _dp)))
        return out, normals

    # ------------------------------------------------------------------ per-evaluation helpers
    def disc_pixels(self, P, pixels_ray):
        P = _arr(P, np.float64).reshape(3)
        cap = (2 * int(pixels_ray) + 1) ** 2
        xy = np.empty((cap, 2))
        m = C.c_int()
        self._ck(self.lib.fm3d_disc_pixels(self._h, _ptr(P, _dp), int(pixels_ray), _ptr(xy, _dp), cap, C.byref(m)))
        return xy[:m.value].copy()

    def plane_points(self, P, normal, xy):
        P, normal = _arr(P, np.float64).reshape(3), _arr(normal, np.float64).reshape(3)
        xy = _arr(xy, np.float64).reshape(-1, 2)
        out = np.empty((xy.shape[0], 3))
        info = C.c_int()
        self._ck(self.lib.fm3d_plane_points(self._h, _ptr(P, _dp), _ptr(normal, _dp), _ptr(xy, _dp), xy.shape[0], _ptr(out, _dp), C.byref(info)))
        return out, info.value

    def sample_pixels(self, image, level, scale, xy, gate=True):
        xy = _arr(xy, np.float64).reshape(-1, 2)
        out = np.empty(xy.shape[0], np.float32)
        info = C.c_int()
        self._ck(self.lib.fm3d_sample_pixels(self._h, int(image), int(level), C.c_double(scale), int(bool(gate)), _ptr(xy, _dp),
                                             xy.shape[0], out.ctypes.data_as(C.POINTER(C.c_float)), C.byref(info)))
        return out, info.value

    def project_to_image2(self, xyz, level, scale, want_intensity=True):
        xyz = _arr(xyz, np.float64).reshape(-1, 3)
        xy2 = np.empty((xyz.shape[0], 2))
        inten = np.empty(xyz.shape[0], np.float32) if want_intensity else None
        info = C.c_int()
        self._ck(self.lib.fm3d_project_to_image2(self._h, _ptr(xyz, _dp), xyz.shape[0], int(level), C.c_double(scale), _ptr(xy2, _dp),
                                                 None if inten is None else inten.ctypes.data_as(C.POINTER(C.c_float)), C.byref(info)))
        return xy2, inten, info.value

    def feature_frames(self, xyz, normals, gravity):
        xyz = _arr(xyz, np.float64).reshape(-1, 3)
        normals = _arr(normals, np.float64).reshape(-1, 3)
        g = _arr(gravity, np.float64).reshape(3)
        n = xyz.shape[0]
        frames = np.empty((n, 4, 4))
        self._ck(self.lib.fm3d_feature_frames(self._h, _ptr(xyz, _dp), _ptr(normals, _dp), n, _ptr(g, _dp), _ptr(frames, _dp)))
        return frames

    # ------------------------------------------------------------------ patches
    def extract_patches(self, frames, epsilon_m, cm_per_pixel, want_points=True):
        frames = _arr(frames, np.float64).reshape(-1, 16)
        n = frames.shape[0]
        S = patch_size(epsilon_m, cm_per_pixel)
        patches = np.empty((n, S, S), np.uint8)
        ip = np.empty((n, S * S, 2)) if want_points else None
        self._ck(self.lib.fm3d_extract_patches(self._h, _ptr(frames, _dp), n, C.c_double(epsilon_m), C.c_double(cm_per_pixel),
                                               _ptr(patches, _bp), _ptr(ip, _dp)))
        return patches, ip

    def project_groups(self, image, groups, want_points=True):
        groups = _arr(groups, np.float64)
        n, ss = groups.shape[0], groups.shape[1]
        S = int(round(np.sqrt(ss)))
        patches = np.empty((n, S, S), np.uint8)
        ip = np.empty((n, S * S, 2)) if want_points else None
        self._ck(self.lib.fm3d_project_groups(self._h, int(image), _ptr(groups, _dp), n, S, _ptr(patches, _bp), _ptr(ip, _dp)))
        return patches, ip

    def square_neighborhoods(self, frames, epsilon_m, cm_per_pixel):
        frames = _arr(frames, np.float64).reshape(-1, 16)
        n = frames.shape[0]
        S = patch_size(epsilon_m, cm_per_pixel)
        out = np.empty((n, S * S, 3))
        self._ck(self.lib.fm3d_square_neighborhoods(self._h, _ptr(frames, _dp), n, C.c_double(epsilon_m), C.c_double(cm_per_pixel), _ptr(out, _dp)))
        return out

    def describe_patches_sift(self, patches):
        """extractDescriptorsFromPatches with ExtractorType SIFT: n x S x S u8 -> n x 128 f32."""
        patches = _arr(patches, np.uint8)
        n, S = patches.shape[0], patches.shape[1]
        desc = np.empty((n, 128), np.float32)
        self._ck(self.lib.fm3d_describe_patches_sift(self._h, _ptr(patches, _bp), n, S, _ptr(desc, _fp)))
        return desc

    def detect_fast(self, img, threshold, nonmax=True, max_keypoints=None):
        """feature_detector_->detect for DetectorType FAST: H x W u8 -> (n x 2 f32 xy, n f32 response),
        row-major order like cv::FAST.  max_keypoints=None counts first, then fetches all of them."""
        img = np.asarray(img)
        if img.dtype != np.uint8 or img.ndim != 2 or img.strides[1] != 1:
            img = _arr(img, np.uint8)
        h, w = img.shape
        n = C.c_int(0)
        if max_keypoints is None:
            self._ck(self.lib.fm3d_detect_fast(self._h, _ptr(img, _bp), w, h, img.strides[0], int(threshold), int(bool(nonmax)), 0, None, None, C.byref(n)))
            max_keypoints = n.value
        xy = np.empty((max(max_keypoints, 1), 2), np.float32)
        resp = np.empty(max(max_keypoints, 1), np.float32)
        if max_keypoints > 0:
            self._ck(self.lib.fm3d_detect_fast(self._h, _ptr(img, _bp), w, h, img.strides[0], int(threshold), int(bool(nonmax)), int(max_keypoints),
                                               _ptr(xy, _fp), _ptr(resp, _fp), C.byref(n)))
        got = min(n.value, max_keypoints)
        return xy[:got].copy(), resp[:got].copy(), n.value

    def detect_sift(self, img, nfeatures=0, n_octave_layers=3, contrast_threshold=0.04, edge_threshold=10.0, sigma=1.6):
        """feature_detector_->detect for DetectorType SIFT: H x W u8 -> n x 6 f64 rows (x, y, size, angle, response, octave) in
        the order KeyPointsFilter::removeDuplicatedSorted leaves (cv2.SIFT_create(...).detect)."""
        img = np.asarray(img)
        if img.dtype != np.uint8 or img.ndim != 2 or img.strides[1] != 1:
            img = _arr(img, np.uint8)
        h, w = img.shape
        n = C.c_int(0)
        args = (int(nfeatures), int(n_octave_layers), C.c_double(contrast_threshold), C.c_double(edge_threshold), C.c_double(sigma))
        m = max(4096, (w * h) // 32)            # one call in the usual case; again with the real count if a frame has more
        for _ in range(2):
            xy = np.empty((m, 2), np.float32)
            size, angle, resp = (np.empty(m, np.float32) for _ in range(3))
            octave = np.empty(m, np.int32)
            self._ck(self.lib.fm3d_detect_sift(self._h, _ptr(img, _bp), w, h, img.strides[0], *args, m, _ptr(xy, _fp), _ptr(size, _fp),
                                               _ptr(angle, _fp), _ptr(resp, _fp), _ptr(octave, C.POINTER(C.c_int32)), C.byref(n)))
            if n.value <= m:
                break
            m = n.value
        got = min(n.value, m)
        return np.column_stack([xy[:got].astype(np.float64), size[:got], angle[:got], resp[:got], octave[:got].astype(np.float64)])

    def detect_and_describe_sift(self, img, nfeatures=0, n_octave_layers=3, contrast_threshold=0.04, edge_threshold=10.0, sigma=1.6):
        """feature_detector_->detect + descriptor_extractor_->compute for DetectorType SIFT + ExtractorType SIFT on one pyramid
        (cv2.SIFT_create(...).detectAndCompute): (n x 6 f64 keypoint rows as detect_sift, n x 128 f32 descriptors)."""
        img = np.asarray(img)
        if img.dtype != np.uint8 or img.ndim != 2 or img.strides[1] != 1:
            img = _arr(img, np.uint8)
        h, w = img.shape
        n = C.c_int(0)
        args = (int(nfeatures), int(n_octave_layers), C.c_double(contrast_threshold), C.c_double(edge_threshold), C.c_double(sigma))
        m = max(4096, (w * h) // 32)
        for _ in range(2):
            xy = np.empty((m, 2), np.float32)
            size, angle, resp = (np.empty(m, np.float32) for _ in range(3))
            octave = np.empty(m, np.int32)
            desc = np.empty((m, 128), np.float32)       # rows beyond the keypoints found are never touched (nor their pages)
            self._ck(self.lib.fm3d_detect_and_describe_sift(self._h, _ptr(img, _bp), w, h, img.strides[0], *args, m, _ptr(xy, _fp), _ptr(size, _fp),
                                                            _ptr(angle, _fp), _ptr(resp, _fp), _ptr(octave, C.POINTER(C.c_int32)), C.byref(n),
                                                            _ptr(desc, _fp)))
            if n.value <= m:
                break
            m = n.value
        got = min(n.value, m)
        return (np.column_stack([xy[:got].astype(np.float64), size[:got], angle[:got], resp[:got], octave[:got].astype(np.float64)]),
                desc[:got])

    def detect_orb(self, img, nfeatures=500, scale_factor=1.2, nlevels=8, fast_threshold=20):
        """feature_detector_->detect + descriptor_extractor_->compute for DetectorType ORB + ExtractorType ORB: H x W u8 ->
        (n x 6 f64 rows (x, y, size, angle, response, octave) sorted by (octave, y, x), n x 32 u8 descriptors) --
        cv2.ORB_create(nfeatures, scaleFactor, nlevels).detectAndCompute as a set."""
        img = np.asarray(img)
        if img.dtype != np.uint8 or img.ndim != 2 or img.strides[1] != 1:
            img = _arr(img, np.uint8)
        h, w = img.shape
        n = C.c_int(0)
        m = max(64, int(nfeatures) * 2 + 64)
        for _ in range(2):
            xy = np.empty((m, 2), np.float32)
            size, angle, resp = (np.empty(m, np.float32) for _ in range(3))
            octave = np.empty(m, np.int32)
            desc = np.zeros((m, 32), np.uint8)
            self._ck(self.lib.fm3d_detect_orb(self._h, _ptr(img, _bp), w, h, img.strides[0], int(nfeatures), C.c_double(float(np.float32(scale_factor))),
                                              int(nlevels), int(fast_threshold), m, _ptr(xy, _fp), _ptr(size, _fp), _ptr(angle, _fp), _ptr(resp, _fp),
                                              _ptr(octave, C.POINTER(C.c_int32)), _ptr(desc, _bp), C.byref(n)))
            if n.value <= m:
                break
            m = n.value
        got = min(n.value, m)
        return (np.column_stack([xy[:got].astype(np.float64), size[:got], angle[:got], resp[:got], octave[:got].astype(np.float64)]),
                desc[:got].copy())

    def describe_keypoints_sift_oct(self, img, kps, octaves, n_octave_layers=3, sigma=1.6):
        """descriptor_extractor_->compute for ExtractorType SIFT on keypoints that carry an octave (cv::SIFT's own): kps n x 4
        (x, y, size, angle), octaves n packed cv::KeyPoint::octave -> n x 128 f32."""
        img = np.asarray(img)
        if img.dtype != np.uint8 or img.ndim != 2 or img.strides[1] != 1:
            img = _arr(img, np.uint8)
        h, w = img.shape
        kps = _arr(kps, np.float32).reshape(-1, 4)
        octaves = _arr(octaves, np.int32).reshape(-1)
        n = kps.shape[0]
        assert octaves.shape[0] == n
        desc = np.zeros((n, 128), np.float32)
        self._ck(self.lib.fm3d_describe_keypoints_sift_oct(self._h, _ptr(img, _bp), w, h, img.strides[0], _ptr(kps, _fp),
                                                           _ptr(octaves, C.POINTER(C.c_int32)), n, int(n_octave_layers), C.c_double(sigma),
                                                           _ptr(desc, _fp)))
        return desc

    def describe_keypoints_sift(self, img, kps):
        """descriptor_extractor_->compute for ExtractorType SIFT on octave-0 keypoints: H x W u8 image,
        n x 4 f32 (x, y, size, angle) -> n x 128 f32."""
        img = np.asarray(img)
        if img.dtype != np.uint8 or img.ndim != 2 or img.strides[1] != 1:
            img = _arr(img, np.uint8)
        h, w = img.shape
        kps = _arr(np.asarray(kps, np.float32).reshape(-1, 4), np.float32)
        n = kps.shape[0]
        desc = np.zeros((n, 128), np.float32)
        self._ck(self.lib.fm3d_describe_keypoints_sift(self._h, _ptr(img, _bp), w, h, img.strides[0], _ptr(kps, _fp), n, _ptr(desc, _fp)))
        return desc

    def describe_keypoints_brisk(self, img, kps, compute_orientation=True):
        """descriptor_extractor_->compute for ExtractorType BRISK: H x W u8 image, n x 4 f32 (x, y, size, angle) ->
        (n x 64 u8 rows, n bool kept, n f32 angles); rows of removed keypoints are zero."""
        img = np.asarray(img)
        if img.dtype != np.uint8 or img.ndim != 2 or img.strides[1] != 1:
            img = _arr(img, np.uint8)
        h, w = img.shape
        kps = _arr(np.asarray(kps, np.float32).reshape(-1, 4), np.float32)
        n = kps.shape[0]
        desc = np.zeros((n, 64), np.uint8)
        kept = np.zeros(n, np.uint8)
        ang = np.zeros(n, np.float32)
        self._ck(self.lib.fm3d_describe_keypoints_brisk(self._h, _ptr(img, _bp), w, h, img.strides[0], _ptr(kps, _fp), n,
                                                        int(bool(compute_orientation)), _ptr(desc, _bp), _ptr(kept, _bp), _ptr(ang, _fp)))
        return desc, kept.astype(bool), ang

    def describe_keypoints_orb(self, img, kps):
        """descriptor_extractor_->compute for ExtractorType ORB: H x W u8 image, n x 4 f32 (x, y, size, angle), octave 0 ->
        (n x 32 u8 rows, n bool kept); rows of removed keypoints are zero."""
        img = np.asarray(img)
        if img.dtype != np.uint8 or img.ndim != 2 or img.strides[1] != 1:
            img = _arr(img, np.uint8)
        h, w = img.shape
        kps = _arr(np.asarray(kps, np.float32).reshape(-1, 4), np.float32)
        n = kps.shape[0]
        desc = np.zeros((n, 32), np.uint8)
        kept = np.zeros(n, np.uint8)
        self._ck(self.lib.fm3d_describe_keypoints_orb(self._h, _ptr(img, _bp), w, h, img.strides[0], _ptr(kps, _fp), n, _ptr(desc, _bp), _ptr(kept, _bp)))
        return desc, kept.astype(bool)

    def describe_patches_orb(self, patches):
        """extractDescriptorsFromPatches for ExtractorType ORB: n x S x S u8 -> n x 32 u8."""
        patches = _arr(patches, np.uint8)
        n = patches.shape[0]
        S = patches.shape[1] if patches.ndim == 3 else 0
        desc = np.zeros((n, 32), np.uint8)
        self._ck(self.lib.fm3d_describe_patches_orb(self._h, _ptr(patches, _bp), n, S, _ptr(desc, _bp)))
        return desc

    # ------------------------------------------------------------------ device-pointer entry points
    def describe_patches_orb_dev(self, patches, n, S, descriptors):
        self._ck(self.lib.fm3d_describe_patches_orb_dev(self._h, C.c_void_p(patches), n, S, C.c_void_p(descriptors)))

    def describe_keypoints_orb_dev(self, img, w, h, stride, kps, n, descriptors, kept):
        self._ck(self.lib.fm3d_describe_keypoints_orb_dev(self._h, C.c_void_p(img), w, h, stride, C.c_void_p(kps), n, C.c_void_p(descriptors), C.c_void_p(kept)))

    def describe_keypoints_brisk_dev(self, img, w, h, stride, kps, n, compute_orientation, descriptors, kept, angles):
        self._ck(self.lib.fm3d_describe_keypoints_brisk_dev(self._h, C.c_void_p(img), w, h, stride, C.c_void_p(kps), n, int(bool(compute_orientation)),
                                                            C.c_void_p(descriptors), C.c_void_p(kept), C.c_void_p(angles)))

    def describe_keypoints_sift_dev(self, img, w, h, stride, kps, n, descriptors):
        self._ck(self.lib.fm3d_describe_keypoints_sift_dev(self._h, C.c_void_p(img), w, h, stride, C.c_void_p(kps), n, C.c_void_p(descriptors)))

    def sift_base_image_dev(self, img, w, h, stride, base):
        self._ck(self.lib.fm3d_sift_base_image_dev(self._h, C.c_void_p(img), w, h, stride, C.c_void_p(base)))

    def detect_fast_dev(self, img, w, h, stride, threshold, nonmax, max_keypoints, xy, response, n_dev):
        self._ck(self.lib.fm3d_detect_fast_dev(self._h, C.c_void_p(img), w, h, stride, int(threshold), int(bool(nonmax)), int(max_keypoints),
                                               C.c_void_p(xy), C.c_void_p(response), C.c_void_p(n_dev)))

    def describe_patches_sift_dev(self, patches, n, S, descriptors):
        self._ck(self.lib.fm3d_describe_patches_sift_dev(self._h, C.c_void_p(patches), n, S, C.c_void_p(descriptors)))

    def match_knn2_f32_dev(self, q, nq, t, nt, dim, idx, dist):
        self._ck(self.lib.fm3d_match_knn2_f32_dev(self._h, C.c_void_p(q), nq, C.c_void_p(t), nt, dim, C.c_void_p(idx), C.c_void_p(dist)))

    def match_knn2_hamming_dev(self, q, nq, t, nt, nbytes, idx, dist):
        self._ck(self.lib.fm3d_match_knn2_hamming_dev(self._h, C.c_void_p(q), nq, C.c_void_p(t), nt, nbytes, C.c_void_p(idx), C.c_void_p(dist)))

    def nndr_filter_dev(self, idx, dist, nq, eps, qidx, tidx, dist_out, nmatch):
        self._ck(self.lib.fm3d_nndr_filter_dev(self._h, C.c_void_p(idx), C.c_void_p(dist), nq, C.c_double(eps), C.c_void_p(qidx),
                                               C.c_void_p(tidx), C.c_void_p(dist_out), C.c_void_p(nmatch)))

    def triangulate_dev(self, kp1, n1, kp2, n2, qidx, tidx, n, xyz_all, mask, xyz, src_idx, ninl):
        self._ck(self.lib.fm3d_triangulate_dev(self._h, C.c_void_p(kp1), n1, C.c_void_p(kp2), n2, C.c_void_p(qidx), C.c_void_p(tidx), n,
                                               C.c_void_p(xyz_all), C.c_void_p(mask), C.c_void_p(xyz), C.c_void_p(src_idx), C.c_void_p(ninl)))

    def optimize_normals_dev(self, xyz, n, pixels_ray, epsilon_lmmin, penalty_mode, normals, status, nfev=None, npenalty=None, cost=None):
        self._ck(self.lib.fm3d_optimize_normals_dev(self._h, C.c_void_p(xyz), n, int(pixels_ray), C.c_double(epsilon_lmmin), int(penalty_mode),
                                                    C.c_void_p(normals), C.c_void_p(status), C.c_void_p(nfev), C.c_void_p(npenalty), C.c_void_p(cost)))

    def feature_frames_dev(self, xyz, normals, n, gravity, frames):
        g = _arr(gravity, np.float64).reshape(3)
        self._ck(self.lib.fm3d_feature_frames_dev(self._h, C.c_void_p(xyz), C.c_void_p(normals), n, _ptr(g, _dp), C.c_void_p(frames)))

    def extract_patches_dev(self, frames, n, epsilon_m, cm_per_pixel, patches, image_points=None):
        self._ck(self.lib.fm3d_extract_patches_dev(self._h, C.c_void_p(frames), n, C.c_double(epsilon_m), C.c_double(cm_per_pixel),
                                                   C.c_void_p(patches), C.c_void_p(image_points)))
