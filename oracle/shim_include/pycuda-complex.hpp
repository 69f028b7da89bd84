// Stand-in for pyCUDA's (third-party, not under /root/reference) pycuda-complex.hpp.
// pycuda::complex<T> is provided by oracle/simt_shim.h (-> std::complex<T>), which is
// force-included before the reference translation unit; nothing to add here.
#pragma once
