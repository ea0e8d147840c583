"""CPU-side checks of the drop-in boundary: the library loads, exports every symbol of
include/fm3d.h, refuses to run without a GPU, and the host-only helpers work."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "fm3d.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(fm3d_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol(api):
    lib = api.load_library()
    decl = _declared_symbols()
    assert len(decl) >= 30
    missing = [s for s in decl if not hasattr(lib, s)]
    assert not missing, missing
    assert sorted(api.SYMBOLS) == decl          # the ctypes binding covers the whole header
    assert lib.fm3d_version() == 100


def test_no_cpu_fallback(api):
    """Without a usable sm_100 device the context cannot be created (and nothing else can run)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(api.Fm3dError) as e:
        api.Context(0)
    assert e.value.code == -3


def test_compose_g12_matches_reference_formula(api):
    # g12 = g_IC^-1 g2^-1 g1 g_IC (singlecameratriangulator.cpp:140) on the reference's settings.yml poses
    g = np.load(os.path.join(ROOT, "tests", "golden", "primitives.npz"))
    synth = __import__("importlib").import_module("3dfeaturematcher_b200.synth")
    out = api.compose_g12(g["pos1"][:3], g["pos2"][:3], g["pos1"][3:], g["pos2"][3:], synth.SETTINGS_RODRIGUES_IC,
                          synth.SETTINGS_TRANSLATION_IC)
    np.testing.assert_allclose(out, g["g12"], rtol=0, atol=1e-12)      # golden: cv2.Rodrigues + numpy inverse
    R, t = out[:3, :3], out[:3, 3]
    np.testing.assert_allclose(R @ R.T, np.eye(3), atol=1e-12)
    assert abs(np.linalg.norm(t) - 0.661) < 2e-3                        # |t| = 0.661 m


def test_patch_size(api):
    assert api.patch_size(0.16, 0.25) == 128 and api.patch_size(0.32, 0.5) == 128 and api.patch_size(0.64, 1.0) == 128
    assert api.patch_size(0.05, 0.25) == 40


def _pieces_harness():
    import ctypes as C
    import subprocess
    so = os.path.join(ROOT, "tests", "_build", "libpieces_harness.so")
    src = os.path.join(ROOT, "tests", "cpp", "pieces_harness.cpp")
    hdr = os.path.join(ROOT, "3dfeaturematcher_b200", "csrc", "fm3d_match_pieces.h")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        os.makedirs(os.path.dirname(so), exist_ok=True)
        gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.check_call([gxx, "-O2", "-shared", "-fPIC", "-I" + os.path.dirname(hdr), src, "-o", so])
    return C.CDLL(so)


def test_persistent_matcher_work_distribution_covers_every_tile_pair_once():
    """Host logic of the persistent tensor-core matchers (csrc/fm3d_match_pieces.h, the code the kernels walk): for the sweep's
    sizes, the edge cases (one tile, fewer tiles than CTAs, 149 query tiles on 148 SMs) and 4 000 random shapes every (query
    tile, train tile) pair is contracted exactly once, the pieces of a query tile take consecutive list slots below the planned
    count, exactly one piece closes a query tile, and the CTAs' shares differ by at most one tile."""
    import ctypes as C
    lib = _pieces_harness()
    out = (C.c_int * 4)()
    named = {(40, 24): (148, 6, 7), (391, 196): (148, 517, 518), (782, 391): (148, 2065, 2066), (1563, 782): (148, 8258, 8259),
             (149, 7): (148, 7, 8), (1, 1): (1, 1, 1), (3, 2): (6, 1, 1), (313, 12): (148, 25, 26)}
    for (q, nt), (G, lo, hi) in named.items():
        assert lib.pieces_check(q, nt, 148, 1, out) == 0, (q, nt)
        assert (out[0], out[2], out[3]) == (G, lo, hi), (q, nt, list(out))
    rng = np.random.default_rng(5)
    for _ in range(4000):
        q, nt = int(rng.integers(1, 900)), int(rng.integers(1, 90))
        sms, mt = int(rng.choice([1, 7, 132, 148])), int(rng.choice([1, 2, 6, 40, 5000]))
        assert lib.pieces_check(q, nt, sms, mt, out) == 0, (q, nt, sms, mt)
