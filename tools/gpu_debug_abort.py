import sys, importlib, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from common import *
api = importlib.import_module("3dfeaturematcher_b200.api")
case = stereo_case(640, 480, 40, 1001, 32)
cam = case["scene"].cam
ctx = api.Context(0)
setup_ctx(ctx, case, 2)
Z = 1.9
pix = np.array([[8.0, 240.0], [300.0, 6.0], [320.0, 472.0], [632.0, 470.0], [600.0, 30.0]])
rays = orc.undistort_points(cam.K, cam.dist, pix)
pts = np.concatenate([rays * Z, np.full((pix.shape[0], 1), Z)], axis=1)
xyz = np.concatenate([case["X"][:1], pts, np.array([[9.0, 0.3, Z]]), case["X"][1:2]])
pt = car2sph(xyz / np.linalg.norm(xyz, axis=1, keepdims=True))
for lvl in (2, 1, 0):
    cost, m, st = ctx.evaluate_normals(xyz, pt, 32, lvl, 2)
    oc, om, ost = orc.evaluate_cost(*cam_tuple(cam), case["scene"].img1, case["scene"].img2, 2, xyz, pt, 32, lvl, 2)
    print(lvl, "gpu", st, m, "oracle", ost, om)
res = ctx.optimize_normals(xyz, 32, 1e-10, 2)
print(res["status"], res["nfev"].tolist())
