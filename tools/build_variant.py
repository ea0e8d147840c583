"""Builds a VARIANT of libfm3d.so for A/B measurements on the GPU box: the listed sources are recompiled with extra nvcc
flags (e.g. -DFM3D_NORMALS_L2HINT=1), everything else is linked from the objects of the regular build.

    python tools/build_variant.py <name> <source.cu>[,<source.cu>...] <nvcc flags...>   ->  tools/_bin/libfm3d_<name>.so

Select it at run time with FM3D_LIB=tools/_bin/libfm3d_<name>.so (3dfeaturematcher_b200/api.py)."""
import importlib
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
b = importlib.import_module("3dfeaturematcher_b200.build")


def main():
    name, sources, flags = sys.argv[1], sys.argv[2].split(","), sys.argv[3:]
    b.build()
    out_dir = os.path.join(ROOT, "tools", "_bin")
    os.makedirs(out_dir, exist_ok=True)
    nvcc = b._nvcc()
    ccbin = ["-ccbin", "/usr/bin/g++"] if os.path.exists("/usr/bin/g++") else []
    objs = []
    for s in b.SOURCES:
        obj = os.path.join(b.OBJ, s.replace(".cu", ".o"))
        if s in sources:
            obj = os.path.join(out_dir, f"{name}_{s.replace('.cu', '.o')}")
            cmd = [nvcc] + ccbin + b.NVCC_FLAGS + flags + ["-c", os.path.join(b.CSRC, s), "-o", obj]
            p = subprocess.run(cmd, capture_output=True, text=True)
            open(obj + ".log", "w").write(" ".join(cmd) + "\n" + p.stdout + p.stderr)
            if p.returncode:
                raise SystemExit(p.stdout + p.stderr)
        objs.append(obj)
    lib = os.path.join(out_dir, f"libfm3d_{name}.so")
    subprocess.check_call([nvcc] + ccbin + ["-shared", "-o", lib] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static", "-ldl"])
    print(lib)


if __name__ == "__main__":
    main()
