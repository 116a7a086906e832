#!/usr/bin/env python3
"""Recipe for oracle/_ref/libref_oracle.so -- TEST INFRASTRUCTURE ONLY.

Compiles the reference's OWN sources, where they lie under /root/reference,
into a CPU shared library that is used (a) to pin oracle/rt_oracle.cpp (our
restatement) and (b) as the `cpu_baseline.kind == "reference"` arm of bench.py.

Nothing from /root/reference is copied into the repository: a throw-away temp
directory receives *symlinks* to the reference files (needed because the
reference includes "CLBVHnode.h"/"CLCamera.h" while the files on disk are
clBVHnode.h/CLcamera.h -- a Windows-only case mismatch) and one generated file,
kernel_bvh.inc, which is kernel_bvh.cl after the 3-rule mechanical transform of
SURVEY.md Appendix B:
  (i)   the `#include "CLshared_structs.hpp"` line is replaced by that header
        with its `#ifdef __cplusplus` blocks dropped and its `#ifndef
        __cplusplus` blocks kept (i.e. what an OpenCL C compiler would see);
  (ii)  OpenCL vector literals `(float3)(` become `make_float3(`;
  (iii) `return false;` is appended before RayTriangle's closing brace
        (kernel_bvh.cl:152-153 falls off the end of a non-void function, which
        is harmless in OpenCL C because the value is never used, but is UB in
        C++ and makes g++ -O2 emit an endless loop).
The temp directory is deleted afterwards; the only output is the .so under
oracle/_ref/ (git-ignored, NOT gpurun-ignored so it travels to the GPU box).
"""
import os
import re
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("B2RT_REFERENCE_DIR", "/root/reference")
OUT_DIR = os.path.join(HERE, "_ref")
OUT = os.path.join(OUT_DIR, "libref_oracle.so")
RB = os.path.join(HERE, "ref_build")

CXXFLAGS = ["-std=c++17", "-O2", "-fPIC", "-ffp-contract=off", "-fno-fast-math",
            "-Wno-narrowing", "-Wno-unused-result", "-pthread"]


def strip_cplusplus_blocks(text):
    """Rule (i): evaluate #ifdef/#ifndef __cplusplus as an OpenCL C compiler would."""
    out, stack = [], []
    for line in text.splitlines():
        s = line.strip()
        if re.match(r"#\s*ifdef\s+__cplusplus", s):
            stack.append(False)
            continue
        if re.match(r"#\s*ifndef\s+__cplusplus", s):
            stack.append(True)
            continue
        if re.match(r"#\s*if", s):
            stack.append(None)  # unrelated conditional: keep verbatim
            out.append(line)
            continue
        if re.match(r"#\s*else", s) and stack and stack[-1] is not None:
            stack[-1] = not stack[-1]
            continue
        if re.match(r"#\s*endif", s):
            kind = stack.pop() if stack else None
            if kind is None:
                out.append(line)
            continue
        if re.match(r"#\s*pragma\s+once", s):
            continue
        if all(k is not False for k in stack):
            out.append(line)
    return "\n".join(out) + "\n"


def transform_kernel(kernel_src, structs_src):
    src = kernel_src
    inc = '#include "CLshared_structs.hpp"'
    assert inc in src, "kernel no longer includes CLshared_structs.hpp"
    src = src.replace(inc, strip_cplusplus_blocks(structs_src), 1)          # (i)
    n_lit = src.count("(float3)(")
    assert n_lit == 7, "expected 7 OpenCL float3 literals, found %d" % n_lit
    src = src.replace("(float3)(", "make_float3(")                          # (ii)
    # (iii): RayTriangle ends with "        return false;\n    }\n}\n" (for-loop body, for, function)
    m = re.search(r"bool RayTriangle\(.*?\n\{", src, re.S)
    assert m, "RayTriangle not found"
    start = m.end()
    depth, i = 1, start
    while depth:
        c = src[i]
        depth += (c == "{") - (c == "}")
        i += 1
    src = src[: i - 1] + "    return false; /* rule (iii) */\n" + src[i - 1:]
    return src


def build(verbose=True):
    if not os.path.isdir(REF):
        if os.path.exists(OUT):
            if verbose:
                print("[oracle/_ref] %s absent; keeping prebuilt %s" % (REF, OUT))
            return OUT
        raise RuntimeError("reference sources not found at %s and no prebuilt %s" % (REF, OUT))
    os.makedirs(OUT_DIR, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix="b2rt_ref_")
    try:
        # symlinks (no copies) so quote-includes resolve inside the temp dir first
        links = {
            "CLBVHnode.cpp": "CLBVHnode.cpp", "CLBVHnode.h": "clBVHnode.h",
            "CLOBJloader.cpp": "CLOBJloader.cpp", "CLOBJloader.h": "CLOBJloader.h",
            "CLmathlib.hpp": "CLmathlib.hpp", "CLshared_structs.hpp": "CLshared_structs.hpp",
        }
        for name, target in links.items():
            os.symlink(os.path.join(REF, target), os.path.join(tmp, name))
        with open(os.path.join(REF, "kernel_bvh.cl")) as f:
            ksrc = f.read()
        with open(os.path.join(REF, "CLshared_structs.hpp")) as f:
            ssrc = f.read()
        with open(os.path.join(tmp, "kernel_bvh.inc"), "w") as f:
            f.write(transform_kernel(ksrc, ssrc))
        inc = ["-I", tmp, "-I", os.path.join(RB, "shims"), "-I", RB]
        objs = []
        for src in (os.path.join(tmp, "CLBVHnode.cpp"), os.path.join(tmp, "CLOBJloader.cpp"),
                    os.path.join(RB, "ref_host.cpp"), os.path.join(RB, "ref_kernel.cpp")):
            obj = os.path.join(tmp, os.path.basename(src) + ".o")
            cmd = ["g++"] + CXXFLAGS + inc + ["-c", src, "-o", obj]
            if verbose:
                print("[oracle/_ref]", " ".join(cmd))
            subprocess.check_call(cmd)
            objs.append(obj)
        cmd = ["g++", "-shared", "-pthread", "-o", OUT] + objs
        subprocess.check_call(cmd)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return OUT


if __name__ == "__main__":
    print(build(verbose="-q" not in sys.argv))
