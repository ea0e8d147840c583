"""Regenerates profiles/sass/*.txt from the built objects (cuobjdump -sass): the Blackwell-native
instructions of the matcher (tcgen05 / TMEM / bulk copy) and the hot loops of the normal search."""
import os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "3dfeaturematcher_b200", "csrc", "_obj")
OUT = os.path.join(ROOT, "profiles", "sass")
os.makedirs(OUT, exist_ok=True)


def sass(obj, fun=None):
    cmd = ["cuobjdump", "-sass"] + (["-fun", fun] if fun else []) + [os.path.join(OBJ, obj)]
    txt = subprocess.run(cmd, capture_output=True, text=True).stdout
    lines = [re.sub(r"/\* 0x[0-9a-f]+ \*/", "", l).rstrip() for l in txt.splitlines()]
    return [l for l in lines if l.strip() and not re.match(r"^\s*/\* 0x", l)]


def functions(obj):
    txt = subprocess.run(["cuobjdump", "-sass", os.path.join(OBJ, obj)], capture_output=True, text=True).stdout
    return re.findall(r"Function : (\S+)", txt)


# ---- matcher
m = sass("fm3d_match.o")
pat = re.compile(r"UTC[A-Z]*MMA|LDTM|UBLKCP|UTCATOMSWS|UTCBAR|SYNCS\.ARRIVE|POPC")
with open(os.path.join(OUT, "r01_match_kernels.txt"), "w") as f:
    f.write("# cuobjdump -sass fm3d_match.o (sm_100a).  tcgen05.mma -> UTCHMMA, tcgen05.ld -> LDTM, cp.async.bulk -> UBLKCP,\n"
            "# tcgen05.alloc/dealloc -> UTCATOMSWS, tcgen05.commit -> UTCBAR, mbarrier -> SYNCS; Hamming path: POPC.\n")
    counts = {}
    for l in m:
        for k in pat.findall(l):
            counts[k] = counts.get(k, 0) + 1
    f.write("# counts: " + ", ".join(f"{k} x{v}" for k, v in sorted(counts.items())) + "\n")
    for l in m:
        if re.search(r"UTC[A-Z]*MMA|LDTM|UBLKCP|UTCATOMSWS|UTCBAR", l):
            f.write(l[:140] + "\n")

# ---- normal search: the value+Jacobian loop of the layout the bench runs (rays streamed from L2)
for fn in functions("fm3d_normals_fast.o"):
    tag = "rays_smem" if "ILb1ELb1E" in fn else ("rays_l2_i1_smem" if "ILb0ELb1E" in fn else "rays_l2")
    if re.search(r"ILb[01]ELb[01]ELb1EEE", fn):
        tag += "_ncc"                      # cost_mode NCC: its own instantiation
    if "normals_pp_kernel" in fn:          # the two-slot kernel (mode 0, four windows per CTA)
        tag = "two_slot_ncc" if "ILb1EEE" in fn else "two_slot"
    s = sass("fm3d_normals_fast.o", fn)
    # loops: backward branches; pick the ones that contain LDS.U8 (window taps)
    addr = {}
    for i, l in enumerate(s):
        mm = re.match(r"\s*/\*([0-9a-f]+)\*/", l)
        if mm:
            addr[int(mm.group(1), 16)] = i
    loops = []
    for i, l in enumerate(s):
        mm = re.search(r"BRA (0x[0-9a-f]+)", l)
        a = re.match(r"\s*/\*([0-9a-f]+)\*/", l)
        if mm and a and int(mm.group(1), 16) < int(a.group(1), 16) and int(mm.group(1), 16) in addr:
            j = addr[int(mm.group(1), 16)]
            body = s[j:i + 1]
            if sum("LDS.U8" in b for b in body) >= 8 and len(body) < 400:
                loops.append((j, i, body))
    with open(os.path.join(OUT, f"r02_normals_fast_kernel_{tag}.txt"), "w") as f:
        f.write(f"# cuobjdump -sass -fun {fn}\n# TMA window staging: " +
                ", ".join(sorted({k for l in s for k in re.findall(r"UTMALDG\.\w+|SYNCS\.[A-Z0-9.]+", l)})) + "\n")
        for (j, i, body) in loops:
            ops = {}
            for b in body:
                op = re.sub(r"^\s*/\*[0-9a-f]+\*/\s*(@!?U?P\d+\s+)?", "", b).split()[0].split(".")[0]
                ops[op] = ops.get(op, 0) + 1
            npix = max(1, sum("LDS.U8" in b for b in body) // 4)
            fp = ops.get("FFMA", 0) + 2 * (ops.get("FFMA2", 0) + ops.get("FMUL2", 0) + ops.get("FADD2", 0)) + ops.get("FMUL", 0) + ops.get("FADD", 0)
            kind = "value+Jacobian" if fp / npix > 60 else "value-only"
            f.write(f"\n# ---- pixel loop ({kind}, {npix} pixels per iteration): {len(body)} instructions; "
                    + ", ".join(f"{k} {v}" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])) + "\n")
            for b in body:
                f.write(b[:120] + "\n")
print("wrote", os.listdir(OUT))

# ---- the front end (K10-K13): opcode mix per kernel and the inner loops of the two descriptor kernels
import collections
with open(os.path.join(OUT, "r01_frontend_kernels.txt"), "w") as f:
    f.write("# cuobjdump -sass of fm3d_detect.o, fm3d_describe_kp.o, fm3d_describe_brisk.o, fm3d_describe_orb.o (sm_100a):\n"
            "# instruction count and opcode mix per kernel (static), then the full listing of brisk_kp_kernel's and orb_kp_kernel's bodies.\n")
    for obj in ("fm3d_detect.o", "fm3d_describe_kp.o", "fm3d_describe_brisk.o", "fm3d_describe_orb.o"):
        for fn in functions(obj):
            body = [l for l in sass(obj, fn) if re.match(r"^\s+/\*[0-9a-f]{4}\*/", l)]
            ops = collections.Counter()
            for l in body:
                m_ = re.search(r"\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", l)
                if m_:
                    ops[m_.group(1)] += 1
            short = re.sub(r"^_ZN\d+_GLOBAL__N__[0-9a-f]+_\d+_", "", fn)
            f.write(f"\n## {obj}: {short}\n# {len(body)} instructions: " + ", ".join(f"{k} {v}" for k, v in ops.most_common(14)) + "\n")
            if "brisk_kp_kernel" in fn or "orb_kp_kernel" in fn:
                for l in body:
                    f.write(l[:120] + "\n")
