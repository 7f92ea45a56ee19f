// INTEGRATION TEST INFRASTRUCTURE: stands where the reference's types_cuda.hpp stands in the `b200` target of
// oracle/ref.mk.  The reference's CLI and driver name their GPU device `Cuda` (gmres_perf_test.cpp:424, gmres.cpp:360);
// here that name is bound to the B200 backend, so the UNMODIFIED reference sources run on libmpgmres_b200.so with --gpu.
#ifndef SHIM_B200_TYPES_CUDA_HPP
#define SHIM_B200_TYPES_CUDA_HPP
#include "b200/types_b200.hpp"
typedef B200 Cuda;
#endif
