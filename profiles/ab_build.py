"""Build variants of libnerfq.so with extra -D flags into profiles/_ab/<name>/libnerfq.so for A/B timing on one box:

    python profiles/ab_build.py base "" hint4k "-DNERFQ_MBAR_SUSPEND_NS=4000"
    NERFQ_LIB=profiles/_ab/hint4k/libnerfq.so python profiles/time_mlp.py
"""
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "vanilla-nerf-model-compression-using-lsa-enhanced-nncodec_b200", "csrc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]
args = sys.argv[1:]
for name, extra in zip(args[0::2], args[1::2]):
    out = os.path.join(ROOT, "profiles", "_ab", name)
    os.makedirs(out, exist_ok=True)
    procs, objs = [], []
    for s in sorted(glob.glob(os.path.join(CSRC, "*.cu"))):
        o = os.path.join(out, os.path.basename(s)[:-3] + ".o")
        objs.append(o)
        procs.append(subprocess.Popen(["nvcc"] + FLAGS + extra.split() + ["-c", s, "-o", o]))
    for p in procs:
        if p.wait() != 0:
            raise SystemExit(f"nvcc failed for variant {name}")
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", os.path.join(out, "libnerfq.so")] + objs)
    for o in objs:
        os.remove(o)
    print(name, "->", os.path.join(out, "libnerfq.so"))
