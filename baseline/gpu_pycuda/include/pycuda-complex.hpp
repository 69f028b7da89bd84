// Stand-in for pyCUDA's pycuda-complex.hpp (third-party, pycuda >= 2018.1, not under /root/reference) when
// the reference's kernel text is compiled with nvcc for the GPU ("B-ref", oracle/build_ref.py:build_cubin).
// pyCUDA's header is an STLport-derived complex class with the usual operators; thrust::complex provides
// the same operations on the device (SURVEY.md section 8c verified that the reference TU compiles with it).
#pragma once
extern "C++" {
#include <thrust/complex.h>
namespace pycuda {
template <class T> using complex = thrust::complex<T>;
}
}
