"""Developer helper (not a test): build variants/libb2rt_<name>.so with extra nvcc flags for tests/dev_ab.py.
Usage: python tests/dev_variant.py name [-DB2_MIN_BLOCKS=9 ...]"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mini-opencl-raytracer_b200"))
import build
name, extra = sys.argv[1], sys.argv[2:]
out = os.path.join(ROOT, "variants", "libb2rt_%s.so" % name)
objdir = os.path.join(ROOT, "variants", "obj_" + name)
os.makedirs(objdir, exist_ok=True)
objs = []
for f in ("kernels.cu", "lbvh.cu", "api.cu", "multi.cu", "refit.cu", "wide_bvh.cpp"):
    o = os.path.join(objdir, f + ".o")
    subprocess.check_call([build.NVCC] + build.NVCC_FLAGS + extra + ["-x", "cu", "-c", os.path.join(build.CSRC, f), "-o", o])
    objs.append(o)
subprocess.check_call([build.NVCC, "-shared", "-o", out] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-ldl"])
print(out)
