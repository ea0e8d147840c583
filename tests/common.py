"""Shared helpers of the test-suite: seeded scenes, oracle adapters, comparisons."""
import functools
import importlib

import numpy as np

synth = importlib.import_module("3dfeaturematcher_b200.synth")
from oracle import oracle_c as orc  # noqa: E402  (the oracle is test infrastructure)


@functools.lru_cache(maxsize=8)
def stereo_case(width, height, n, seed, pixels_ray):
    return synth.make_stereo_case(width, height, n, seed, pixels_ray=pixels_ray)


def cam_tuple(cam):
    return cam.K, cam.dist, cam.g12, cam.z_min, cam.z_max


def setup_ctx(ctx, case, pyramids):
    cam = case["scene"].cam
    ctx.set_camera(cam.K, cam.dist, cam.z_min, cam.z_max)
    ctx.set_g12(cam.g12)
    ctx.set_images(case["scene"].img1, case["scene"].img2, pyramids)


def angle_deg(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    c = np.clip((a * b).sum(-1) / (np.linalg.norm(a, axis=-1) * np.linalg.norm(b, axis=-1)), -1, 1)
    return np.degrees(np.arccos(c))


def car2sph(v):
    v = np.asarray(v, dtype=np.float64).reshape(-1, 3)
    theta = np.arctan2(v[:, 2], np.sqrt(v[:, 0] ** 2 + v[:, 1] ** 2))
    phi = np.arctan2(v[:, 1], v[:, 0])
    return np.stack([phi, theta], axis=1)
