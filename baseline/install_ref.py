"""Install the UNMODIFIED reference package into baseline/_ref (git-ignored, shipped to the GPU box).

The base contract's one offline install: ``pip install --no-index --no-build-isolation --no-deps
--target baseline/_ref <copy of /root/reference>`` (the build writes into the source tree, /root/reference is
read-only, hence the copy under /tmp; --no-deps because pyCUDA -- the reference's only hard dependency that this
image lacks -- cannot be installed offline: baseline/gpu_pycuda stands in for it on the GPU box).  Used by
baseline/bref.py ("B-ref": the reference's own Python + CUDA kernels on the B200); never imported by svirl_b200."""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
DST = os.path.join(HERE, "_ref")


def install(force=False):
    if not os.path.isdir(os.path.join(REF, "svirl")):
        print("reference sources not present; nothing to install")
        return False
    if os.path.isdir(os.path.join(DST, "svirl")) and not force:
        return True
    tmp = tempfile.mkdtemp(prefix="svirl_ref_")
    try:
        src = os.path.join(tmp, "reference")
        shutil.copytree(REF, src)
        shutil.rmtree(DST, ignore_errors=True)
        subprocess.check_call([sys.executable, "-m", "pip", "install", "-q", "--no-index", "--no-build-isolation",
                               "--no-deps", "--find-links", "/opt/wheelhouse", "--target", DST, src])
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return True


# kernel modules compiled ahead for the configurations bench.py / the GPU tests instantiate, so that the GPU box
# does not spend its time in nvcc (anything else is JIT-compiled there, like pyCUDA would)
PREBUILD = [
    ("float32", 2048, 2048, 0.5, 0.5, 5),        # cfg2
    ("float64", 8192, 8192, 0.5, 0.5, 17),       # cfg3, cfg4s
    ("float64", 16384, 16384, 0.5, 0.5, 17),     # cfg4
    ("float64", 129, 129, 0.5, 0.5, 17),         # cfg1
    ("float64", 300, 270, 0.5, 0.5, 17), ("float32", 300, 270, 0.5, 0.5, 5), ("float64", 300, 270, 0.5, 0.5, 5),
    ("float64", 37, 29, 0.5, 0.4, 17),           # fixture td_f64_k5
    ("float64", 4096, 4096, 0.5, 0.5, 17),
]


def prebuild_cubins():
    """Same text as svirl/parallel/startup.py:41-62 builds (oracle/build_ref.reference_code), hashed like the shim does."""
    import numpy as np
    sys.path.insert(0, os.path.join(HERE, "..", "oracle"))
    sys.path.insert(0, os.path.join(HERE, "gpu_pycuda"))
    import build_ref
    from pycuda.compiler import cubin_for, PREBUILT
    for dt, Nx, Ny, dx, dy, rvl in PREBUILD:
        print(cubin_for(build_ref.reference_code(np.dtype(dt).type, Nx, Ny, dx, dy, rvl), out_dir=PREBUILT))


def stage_acceptance_scripts():
    """The reference's acceptance scripts (tests/at_*.py + common.py), copied verbatim next to the installed package
    (baseline/_ref/ref_tests/, git-ignored) so that tests/test_gpu_dropin.py can run them UNMODIFIED against svirl_b200
    on the GPU box, where /root/reference does not exist."""
    import glob
    src = os.path.join(REF, "tests")
    if not os.path.isdir(src):
        return
    dst = os.path.join(DST, "ref_tests")
    os.makedirs(dst, exist_ok=True)
    for f in glob.glob(os.path.join(src, "at_*.py")) + [os.path.join(src, "common.py")]:
        shutil.copyfile(f, os.path.join(dst, os.path.basename(f)))
    print("staged", dst)


if __name__ == "__main__":
    ok = install(force="--force" in sys.argv)
    print("installed" if ok else "skipped", DST)
    if ok:
        prebuild_cubins()
        stage_acceptance_scripts()
