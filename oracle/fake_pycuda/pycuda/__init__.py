"""Minimal CPU stand-in for the pyCUDA surface the reference uses -- TEST INFRASTRUCTURE ONLY.

pyCUDA (third-party; svirl's setup.py:18-26 asks for >=2018.1) is absent from this image and
cannot run without a GPU.  This package lets the UNMODIFIED reference package under
/root/reference execute on host cores: "device" arrays are numpy arrays and
``SourceModule`` compiles the reference's kernel text with g++ behind oracle/simt_shim.h
(see oracle/build_ref.py).  Only the calls grepped from the reference are provided.
"""
LAUNCH_COUNTS = {}
