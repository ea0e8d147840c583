"""Kernel-level breakdown of fm3d_detect_sift on a 4K frame (CUDA profiler-free: torch profiler events on the context stream are not visible,
so: ncu launch list)."""
import sys, os, importlib, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
api = importlib.import_module("3dfeaturematcher_b200.api")
synth = importlib.import_module("3dfeaturematcher_b200.synth")
W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (3840, 2160)
img = synth.make_stereo_case(W, H, 20, 1001, pixels_ray=32)["scene"].img1
ctx = api.Context(0)
ctx.detect_sift(img)
t0 = time.perf_counter(); K = ctx.detect_sift(img); t1 = time.perf_counter()
print(f"{W}x{H}: {len(K)} keypoints, {1e3*(t1-t0):.2f} ms host to host")
ctx.close()
