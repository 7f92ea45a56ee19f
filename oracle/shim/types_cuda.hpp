// ORACLE — TEST INFRASTRUCTURE ONLY.
// Stand-in for the reference's types_cuda.hpp in the host-only oracle/_ref build.  The real header needs Kokkos::Cuda
// and the legacy cuSPARSE API (csrsv2Info_t, cusparse?csrmv) that CUDA 12 removed, so the reference's CUDA backend
// cannot be compiled in this image at all.  The host build maps the `Cuda` tag onto `MKL`, which
// turns `CREATE_TEST_CONFIGS(Cuda)` (gmres.cpp:360) into a repeat of the MKL instantiations (accepted with
// -fpermissive) and `run_tests<Cuda>` (gmres_perf_test.cpp:424) into the MKL path.
#ifndef ORACLE_SHIM_TYPES_CUDA_HPP
#define ORACLE_SHIM_TYPES_CUDA_HPP

#define Cuda MKL
#endif
