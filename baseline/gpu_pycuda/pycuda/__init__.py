"""GPU stand-in for the pyCUDA surface the reference uses -- BASELINE INFRASTRUCTURE ONLY ("B-ref").

pyCUDA (third-party; svirl's setup.py asks for >= 2018.1) is absent from this image and cannot be installed
offline.  This package provides exactly the calls grepped from the reference on top of
the CUDA driver API (cuda-python, part of the image), so that the UNMODIFIED reference package installed in
baseline/_ref runs on the B200 with its own kernels (JIT-compiled by nvcc for sm_100a like pyCUDA's SourceModule
does), its own launch geometry (block 128, one thread per node), its own per-sweep fill + launch + blocking 4-byte
read-back, and its own host line search.  Never imported by svirl_b200.

LAUNCH_COUNTS counts kernel launches by name (the Jacobi sweep counts of a run)."""
LAUNCH_COUNTS = {}
