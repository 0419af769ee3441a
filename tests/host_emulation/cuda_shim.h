// TEST INFRASTRUCTURE ONLY: lets g++ compile the device headers (diffdart_b200/csrc/*.cuh) for the host so
// that their arithmetic ORDER can be checked on the CPU against the reference's own Dantzig solver
// (oracle/_ref). Nothing here is loaded by the product; the product path is the CUDA library.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstring>
#include <cstdio>
#include <cstdlib>
#define __device__
#define __host__
#define __forceinline__ inline
#define __constant__ static  // one copy per translation unit, as on the device
#define __global__
#define __restrict__
#define __launch_bounds__(...)
#define DDB_HOST_EMULATION 1
struct double2 { double x, y; };
static inline double2 make_double2(double a, double b) { return double2{a, b}; }
template <class T> static inline T __ldg(const T* p) { return *p; }
