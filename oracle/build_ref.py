"""Build recipe: compile the reference's CUDA kernel sources for the CPU -- TEST INFRASTRUCTURE ONLY.

``build_module(code)`` takes the final, %-substituted translation unit the reference
hands to ``pycuda.compiler.SourceModule`` (svirl/parallel/startup.py:62-63), wraps it in
``extern "C" { }`` like pyCUDA does, appends one generated ``simt_launch_<kernel>``
trampoline per ``__global__`` function, and compiles it with g++ behind
``oracle/simt_shim.h``.  Outputs go ONLY to ``oracle/_ref/`` (git-ignored, but shipped to
the GPU box).  No reference source is copied into the repository: the text is read
from /root/reference at build time.

``reference_code(dtype, Nx, Ny, dx, dy, rvl)`` reproduces startup.py's concatenation +
substitution for prebuilding modules without importing the reference package
(used for the bench's CPU baseline configs).
"""
import hashlib
import os
import re
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF_CUDA = "/root/reference/svirl/cuda"
CUDA_FILES = ["common.h", "block_reduction.h", "reduction.h", "utils.h", "observables.h", "td.h", "cg.h"]

_SYNC_RE = re.compile(r"block_reduce_sum|__syncthreads|__shfl")


def _strip_comments(src):
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return re.sub(r"//[^\n]*", "", src)


def parse_kernels(code):
    """-> {name: {'params': [(ctype, name)], 'fiber': bool}} for every __global__ function."""
    clean = _strip_comments(code)
    out = {}
    for m in re.finditer(r"__global__\s+void\s+(\w+)\s*\(([^)]*)\)\s*\{", clean):
        name, plist = m.group(1), m.group(2)
        params = []
        for p in plist.split(","):
            p = " ".join(p.split())
            if not p:
                continue
            mm = re.match(r"(.*?)(\w+)$", p)
            params.append((mm.group(1).strip(), mm.group(2)))
        # body = up to the matching brace
        depth, k = 1, m.end()
        while depth and k < len(clean):
            depth += {"{": 1, "}": -1}.get(clean[k], 0)
            k += 1
        out[name] = {"params": params, "fiber": bool(_SYNC_RE.search(clean[m.end():k]))}
    return out


def _trampolines(kernels):
    lines = []
    for name, info in kernels.items():
        decl = ", ".join("%s %s" % (t, n) for t, n in info["params"])
        args = ", ".join(n for _, n in info["params"])
        lines.append('extern "C" void simt_launch_%s(int gx_, int gy_, int bx_, int mode_%s%s) {' %
                     (name, ", " if decl else "", decl))
        lines.append("    simt_launch(gx_, gy_, bx_, mode_, [&]() { %s(%s); });" % (name, args))
        lines.append("}")
    return "\n".join(lines)


def build_module(code, tag="mod"):
    """Compile (cached by content hash) and return the path of the .so."""
    os.makedirs(OUT, exist_ok=True)
    kernels = parse_kernels(code)
    with open(os.path.join(HERE, "simt_shim.h")) as f:
        shim_txt = f.read()
    h = hashlib.sha1((code + shim_txt).encode()).hexdigest()[:16]
    so = os.path.join(OUT, "ref_%s_%s.so" % (tag, h))
    if os.path.exists(so):
        return so
    tu = 'extern "C" {\n%s\n}\n%s\n' % (code, _trampolines(kernels))
    src = os.path.join(OUT, "ref_%s_%s.cpp" % (tag, h))      # generated TU lives only in the ignored _ref/
    with open(src, "w") as f:
        f.write(tu)
    cmd = ["g++", "-O2", "-std=c++14", "-fPIC", "-shared", "-fopenmp", "-ffp-contract=off", "-w",
           "-include", os.path.join(HERE, "simt_shim.h"), "-I", os.path.join(HERE, "shim_include"),
           src, "-o", so + ".tmp"]
    subprocess.check_call(cmd)
    os.replace(so + ".tmp", so)
    os.remove(src)
    return so


def reference_code(dtype, Nx, Ny, dx, dy, rvl):
    """Same text the reference builds in svirl/parallel/startup.py:41-62."""
    dtype = np.dtype(dtype).type
    tpl = ""
    for fn in CUDA_FILES:
        with open(os.path.join(REF_CUDA, fn)) as f:
            tpl += f.read() + "\n"
    d = {
        "real": {np.float32: "float", np.float64: "double"}[dtype],
        "complex": {np.float32: "pycuda::complex<float>", np.float64: "pycuda::complex<double>"}[dtype],
        "Nx": np.int32(Nx), "Ny": np.int32(Ny), "dx": dtype(dx), "dy": dtype(dy),
        "reduction_vector_length": rvl,
    }
    return tpl % d


def prebuilt_name(dtype, Nx, Ny, dx, dy, rvl):
    return "ref_fixed_%s_%dx%d_%s_%s_%d.so" % (np.dtype(dtype).name, Nx, Ny, str(dx), str(dy), rvl)


def prebuild(dtype, Nx, Ny, dx, dy, rvl):
    """Build a module under a predictable name so it can be found on the GPU box
    (where /root/reference does not exist)."""
    dst = os.path.join(OUT, prebuilt_name(dtype, Nx, Ny, dx, dy, rvl))
    if os.path.exists(dst):
        return dst
    so = build_module(reference_code(dtype, Nx, Ny, dx, dy, rvl), tag="fixed")
    import shutil
    shutil.copyfile(so, dst)
    return dst


# configurations bench.py's CPU baseline uses (cfg2 sample sizes)
PREBUILD = [
    (np.float32, 2048, 2048, 0.5, 0.5, 5),
    (np.float32, 2048, 512, 0.5, 0.5, 5),      # row-band sample of cfg2 for long --impl reference runs
    (np.float32, 512, 512, 0.5, 0.5, 5),
    (np.float64, 129, 129, 0.5, 0.5, 17),
]

if __name__ == "__main__":
    if not os.path.isdir(REF_CUDA):
        print("reference sources not present; nothing to build")
        sys.exit(0)
    for cfg in PREBUILD:
        print(prebuild(*cfg))
