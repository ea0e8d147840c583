import sys, importlib, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from common import *
api = importlib.import_module("3dfeaturematcher_b200.api")
tma = int(sys.argv[1]) if len(sys.argv) > 1 else 0
case = stereo_case(640, 480, 40, 1001, 32)
cam = case["scene"].cam
ctx = api.Context(0)
ctx.set_option("normals_tma", tma)
setup_ctx(ctx, case, 2)
xyz = case["X"][:24]
n0 = xyz / np.linalg.norm(xyz, axis=1, keepdims=True)
pt = car2sph(n0)
for level in (2, 1, 0):
    cost, m, status = ctx.evaluate_normals(xyz, pt, 32, level, 2)
    o_cost, o_m, o_status = orc.evaluate_cost(*cam_tuple(cam), case["scene"].img1, case["scene"].img2, 2, xyz, pt, 32, level, 2)
    print("tma", tma, "level", level, "m eq", (m == o_m).all(), "status", status[:8], o_status[:8], "rel", np.nanmax(np.abs(cost / o_cost - 1)))
res = ctx.optimize_normals(case["X"], 32, 1e-10, 2)
o = orc.optimize_normals(*cam_tuple(cam), case["scene"].img1, case["scene"].img2, 2, case["X"], 32, 1e-10, penalty_mode=2, threads=8)
print("status eq", (res["status"] == o["status"]).all(), "max angle", angle_deg(res["normals"], o["normals"]).max())
print("nfev gpu", res["nfev"][:5].tolist(), "oracle", o["nfev"][:5].tolist())
