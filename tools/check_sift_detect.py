"""GPU SIFT detector (fm3d_detect_sift) against cv2.SIFT_create().detect on a few frames: matched fraction and timings."""
import sys, os, importlib, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, cv2
api = importlib.import_module("3dfeaturematcher_b200.api")
synth = importlib.import_module("3dfeaturematcher_b200.synth")


def match(A, B, tol_px=0.01, tol_size=1e-3, tol_ang=0.1):
    """fraction of rows of A with a partner in B"""
    if len(A) == 0:
        return 1.0, 0
    hit = 0
    for a in A:
        d = np.abs(B[:, :2] - a[:2]).max(1)
        c = np.nonzero(d < tol_px)[0]
        ok = False
        for j in c:
            da = abs(B[j, 3] - a[3]); da = min(da, 360 - da)
            if abs(B[j, 2] - a[2]) <= tol_size * max(1.0, a[2]) and da < tol_ang and B[j, 5] == a[5]:
                ok = True; break
        hit += ok
    return hit / len(A), hit


ctx = api.Context(0)
rng = np.random.default_rng(3)
imgs = {}
b = cv2.GaussianBlur(rng.integers(0, 256, (120, 160)).astype(np.float32), (0, 0), 2.0)
imgs["blobs160x120"] = ((b - b.min()) / (b.max() - b.min()) * 255).astype(np.uint8)
case = synth.make_stereo_case(640, 480, 50, 1001, pixels_ray=32)
imgs["synth640x480"] = case["scene"].img1
case = synth.make_stereo_case(1280, 720, 50, 1001, pixels_ray=32)
imgs["synth1280x720"] = case["scene"].img1
imgs["odd333x217"] = cv2.resize(imgs["synth640x480"], (333, 217))
for name, img in imgs.items():
    s = cv2.SIFT_create()
    t0 = time.time(); kp = s.detect(img, None); t_cv = time.time() - t0
    Cc = np.array([[k.pt[0], k.pt[1], k.size, k.angle, k.response, k.octave] for k in kp]).reshape(-1, 6)
    ctx.detect_sift(img)
    t0 = time.time(); G = ctx.detect_sift(img); t_gpu = time.time() - t0
    f1, h1 = match(Cc, G); f2, h2 = match(G, Cc)
    same_order = len(Cc) == len(G) and np.abs(Cc[:, :2] - G[:, :2]).max() < 0.01 if len(Cc) else True
    print(f"{name}: cv2 {len(Cc)} gpu {len(G)}; cv2 found on gpu {f1:.4f}, gpu found in cv2 {f2:.4f}; same order {same_order}; cv2 {t_cv*1e3:.1f} ms gpu (two calls incl. copies) {t_gpu*1e3:.1f} ms", flush=True)
    for nf in (100,):
        kp = cv2.SIFT_create(nfeatures=nf).detect(img, None)
        Cn = np.array([[k.pt[0], k.pt[1], k.size, k.angle, k.response, k.octave] for k in kp]).reshape(-1, 6)
        Gn = ctx.detect_sift(img, nfeatures=nf)
        print(f"   nfeatures={nf}: cv2 {len(Cn)} gpu {len(Gn)} matched {match(Cn, Gn)[0]:.4f}")
ctx.close()
