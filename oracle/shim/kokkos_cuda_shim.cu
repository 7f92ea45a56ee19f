// ORACLE / INTEGRATION TEST INFRASTRUCTURE ONLY: device allocation hooks of oracle/shim/Kokkos_Core.hpp for the
// `b200` target of oracle/ref.mk (the reference's unmodified driver and CLI on the B200 backend).  Everything runs on
// the CUDA legacy default stream, like Kokkos::Cuda's default instance and the reference's cuBLAS/cuSPARSE handles.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

static void ck(cudaError_t e, const char* what) {
    if (e != cudaSuccess) {
        std::fprintf(stderr, "kokkos shim: %s: %s\n", what, cudaGetErrorString(e));
        std::abort();
    }
}
extern "C" void* shim_cuda_alloc_zeroed(size_t bytes) {
    void* p = nullptr;
    ck(cudaMalloc(&p, bytes ? bytes : 1), "cudaMalloc");
    ck(cudaMemset(p, 0, bytes ? bytes : 1), "cudaMemset");
    return p;
}
extern "C" void shim_cuda_free(void* p) {
    if (p) cudaFree(p);
}
extern "C" void shim_cuda_memcpy(void* dst, const void* src, size_t bytes, int kind) {
    if (!bytes) return;
    ck(cudaMemcpy(dst, src, bytes, kind == 1 ? cudaMemcpyHostToDevice : (kind == 2 ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice)), "cudaMemcpy");
}
extern "C" void shim_cuda_fence() { ck(cudaDeviceSynchronize(), "cudaDeviceSynchronize"); }
