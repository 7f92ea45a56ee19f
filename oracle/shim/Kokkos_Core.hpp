// ORACLE — TEST INFRASTRUCTURE ONLY.
// Minimal host-only stand-in for the subset of Kokkos 3.x that the reference's MKL path uses, so that the
// reference's own sources (gmres.cpp, Orthogonalization.hpp, IterUtil.hpp, kernels.hpp, types.hpp,
// types_mkl.hpp, kernels_mkl.cpp, gmres_perf_test.cpp, LoadMatrix.hpp) compile UNMODIFIED from where they
// lie under /root/reference (Kokkos itself is not in this image; README.txt:12 pins 3.1.01).
// Semantics kept: View = shallow, reference-counted, zero-initialised, LayoutLeft (column-major) array with
// sub-view constructors; parallel_for/parallel_reduce over a RangePolicy; deep_copy; pair; ALL.
// Written from the public Kokkos API documentation; not derived from Kokkos source.
#ifndef ORACLE_SHIM_KOKKOS_CORE_HPP
#define ORACLE_SHIM_KOKKOS_CORE_HPP

#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <limits>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>

// When the translation unit is compiled by nvcc (oracle/ref.mk target `b200`: the reference driver on the B200
// backend) the same header also provides CudaSpace / Cuda: device-resident Views, lambdas dispatched as kernels.
#if defined(__CUDACC__)
#include <cuda_runtime.h>
#define KOKKOS_LAMBDA [=] __host__ __device__
#define KOKKOS_INLINE_FUNCTION __host__ __device__ inline
#define SHIM_FN __host__ __device__
#else
#define KOKKOS_LAMBDA [=]
#define KOKKOS_INLINE_FUNCTION inline
#define SHIM_FN
#endif

// device allocation hooks: defined in shim/kokkos_cuda_shim.cu for the b200 build, never referenced by the host build
extern "C" void* shim_cuda_alloc_zeroed(size_t bytes);
extern "C" void shim_cuda_free(void* p);
extern "C" void shim_cuda_memcpy(void* dst, const void* src, size_t bytes, int kind);  // 1 H2D, 2 D2H, 3 D2D
extern "C" void shim_cuda_fence();

namespace Kokkos {

struct HostSpace {};
struct CudaSpace {};
struct CudaUVMSpace {};
struct LayoutLeft {};
struct LayoutRight {};
struct OpenMP {
    void fence() const {}
};
struct Cuda {
    void fence() const { shim_cuda_fence(); }
};
struct ALL_t {};
constexpr ALL_t ALL{};

template <class A, class B>
struct pair {
    A first;
    B second;
    pair() : first(), second() {}
    pair(const A& a, const B& b) : first(a), second(b) {}
    template <class U, class V>
    pair(const pair<U, V>& o) : first(static_cast<A>(o.first)), second(static_cast<B>(o.second)) {}
    template <class U, class V>
    pair(const std::pair<U, V>& o) : first(static_cast<A>(o.first)), second(static_cast<B>(o.second)) {}
};

template <class ExecSpace, class Space>
struct SpaceAccessibility {
    enum { accessible = 1 };
};
template <>
struct SpaceAccessibility<OpenMP, CudaSpace> {
    enum { accessible = 0 };
};

inline void abort(const char* msg) {
    std::fprintf(stderr, "%s", msg);
    std::abort();
}
template <class F>
void push_finalize_hook(F) {}

struct ScopeGuard {
    ScopeGuard(int&, char**) {}
    ScopeGuard() {}
};
inline void initialize(int&, char**) {}
inline void finalize() {}

// ---- View -------------------------------------------------------------------------------------------------
namespace detail {
template <class DT> struct data_traits { using value = DT; static constexpr int rank = 0; };
template <class DT> struct data_traits<DT*> { using value = typename data_traits<DT>::value; static constexpr int rank = data_traits<DT>::rank + 1; };

template <class... P> struct has_cuda_space : std::false_type {};
template <class P0, class... P> struct has_cuda_space<P0, P...> : std::integral_constant<bool, std::is_same<P0, CudaSpace>::value || has_cuda_space<P...>::value> {};

struct Range { size_t begin, len; bool scalar; };
inline Range make_range(ALL_t, size_t extent) { return {0, extent, false}; }
template <class A, class B> Range make_range(const pair<A, B>& p, size_t) { return {(size_t)p.first, (size_t)(p.second - p.first), false}; }
template <class A, class B> Range make_range(const std::pair<A, B>& p, size_t) { return {(size_t)p.first, (size_t)(p.second - p.first), false}; }
template <class I, class = typename std::enable_if<std::is_integral<I>::value>::type>
Range make_range(I i, size_t) { return {(size_t)i, 1, true}; }
}  // namespace detail

template <class DataType, class... Props>
class View {
public:
    using value_type = typename detail::data_traits<DataType>::value;
    static constexpr int rank = detail::data_traits<DataType>::rank;
    static constexpr bool on_device = detail::has_cuda_space<Props...>::value;

    std::shared_ptr<value_type> alloc_;
    value_type* ptr_ = nullptr;
    size_t e0_ = 0, e1_ = 0, s1_ = 0;  // extents, column stride (LayoutLeft: stride(0) == 1)

private:
    void allocate(size_t n0, size_t n1) {
        e0_ = n0; e1_ = n1; s1_ = n0;
        const size_t total = (n0 == 0 ? 0 : n0) * (n1 == 0 ? 0 : n1);
        if (total && on_device) {
            value_type* p = static_cast<value_type*>(shim_cuda_alloc_zeroed(total * sizeof(value_type)));
            alloc_ = std::shared_ptr<value_type>(p, [](value_type* q) { shim_cuda_free(q); });
            ptr_ = p;
        } else if (total) {
            value_type* p = static_cast<value_type*>(std::calloc(total, sizeof(value_type)));  // Views zero-fill
            if (!p) throw std::bad_alloc();
            alloc_ = std::shared_ptr<value_type>(p, [](value_type* q) { std::free(q); });
            ptr_ = p;
        }
    }

public:
    View() {}
    explicit View(const std::string&) { static_assert(rank == 0 || rank >= 0, ""); allocate(1, 1); }
    View(const std::string&, size_t n0) { allocate(n0, 1); }
    View(const std::string&, size_t n0, size_t n1) { allocate(n0, n1); }
    // unmanaged view over caller-owned memory (Kokkos: View(pointer, extents...))
    View(value_type* p, size_t n0) : ptr_(p), e0_(n0), e1_(1), s1_(n0) {}
    View(value_type* p, size_t n0, size_t n1) : ptr_(p), e0_(n0), e1_(n1), s1_(n0) {}

    // compatible views convert implicitly (same value type and rank, e.g. default layout -> LayoutLeft)
    template <class DT2, class... P2,
              class = typename std::enable_if<View<DT2, P2...>::rank == rank && std::is_same<typename View<DT2, P2...>::value_type, value_type>::value>::type>
    View(const View<DT2, P2...>& o) : alloc_(o.alloc_), ptr_(o.ptr_), e0_(o.e0_), e1_(o.e1_), s1_(o.s1_) {}

    // sub-views: one argument per source dimension; integral = fix, pair = range, ALL = whole extent
    template <class DT2, class... P2, class A0>
    View(const View<DT2, P2...>& src, const A0& a0) : alloc_(src.alloc_) {
        static_assert(View<DT2, P2...>::rank == 1, "one-argument sub-view needs a rank-1 source");
        const detail::Range r0 = detail::make_range(a0, src.e0_);
        ptr_ = src.ptr_ + r0.begin;
        e0_ = r0.len; e1_ = 1; s1_ = r0.len;
    }
    template <class DT2, class... P2, class A0, class A1>
    View(const View<DT2, P2...>& src, const A0& a0, const A1& a1) : alloc_(src.alloc_) {
        static_assert(View<DT2, P2...>::rank == 2, "two-argument sub-view needs a rank-2 source");
        const detail::Range r0 = detail::make_range(a0, src.e0_);
        const detail::Range r1 = detail::make_range(a1, src.e1_);
        ptr_ = src.ptr_ + r0.begin + r1.begin * src.s1_;
        if (rank == 2) { e0_ = r0.len; e1_ = r1.len; s1_ = src.s1_; }
        else if (rank == 1) {
            if (r1.scalar) { e0_ = r0.len; e1_ = 1; s1_ = r0.len; }
            else throw std::logic_error("shim: strided rank-1 sub-view (row of a LayoutLeft matrix) is not supported");
        } else { e0_ = 1; e1_ = 1; s1_ = 1; }
    }

    SHIM_FN size_t extent(int d) const { return d == 0 ? (rank == 0 ? 1 : e0_) : (d == 1 ? (rank < 2 ? 1 : e1_) : 1); }
    SHIM_FN size_t stride(int d) const { return d == 0 ? 1 : s1_; }
    SHIM_FN size_t size() const { return e0_ * e1_; }
    SHIM_FN value_type* data() const { return ptr_; }

    SHIM_FN value_type& operator()() const { return ptr_[0]; }
    template <class I> SHIM_FN value_type& operator()(const I& i) const { return ptr_[(size_t)i]; }
    template <class I, class J> SHIM_FN value_type& operator()(const I& i, const J& j) const { return ptr_[(size_t)i + (size_t)j * s1_]; }
};

#if defined(__CUDACC__)
namespace detail {
template <class T>
__global__ void shim_fill_kernel(T* p, size_t e0, size_t e1, size_t s1, T v) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < e0 * e1) p[(idx % e0) + (idx / e0) * s1] = v;
}
}  // namespace detail
#endif

template <class DT, class... P, class S>
void deep_copy(const View<DT, P...>& dst, const S& value, typename std::enable_if<std::is_arithmetic<S>::value>::type* = nullptr) {
    using T = typename View<DT, P...>::value_type;
    const size_t e1 = dst.e1_ ? dst.e1_ : 0;
    if (View<DT, P...>::on_device) {
#if defined(__CUDACC__)
        const size_t total = dst.e0_ * e1;
        if (total) detail::shim_fill_kernel<T><<<(unsigned)((total + 255) / 256), 256>>>(dst.ptr_, dst.e0_, e1, dst.s1_, static_cast<T>(value));
#else
        throw std::logic_error("shim: deep_copy(device view, scalar) needs the nvcc build");
#endif
        return;
    }
    for (size_t j = 0; j < e1; ++j)
        for (size_t i = 0; i < dst.e0_; ++i) dst.ptr_[i + j * dst.s1_] = static_cast<T>(value);
}
template <class DT, class... P, class DT2, class... P2>
void deep_copy(const View<DT, P...>& dst, const View<DT2, P2...>& src) {
    using T = typename View<DT, P...>::value_type;
    if (dst.e0_ != src.e0_ || dst.e1_ != src.e1_) throw std::logic_error("shim: deep_copy extent mismatch");
    constexpr bool dd = View<DT, P...>::on_device, sd = View<DT2, P2...>::on_device;
    if (dd || sd) {
        static_assert(std::is_same<T, typename View<DT2, P2...>::value_type>::value || !(dd || sd), "shim: device deep_copy needs equal value types");
        const int kind = (dd && sd) ? 3 : (dd ? 1 : 2);
        for (size_t j = 0; j < dst.e1_; ++j) shim_cuda_memcpy(dst.ptr_ + j * dst.s1_, src.ptr_ + j * src.s1_, dst.e0_ * sizeof(T), kind);
        return;
    }
    for (size_t j = 0; j < dst.e1_; ++j)
        for (size_t i = 0; i < dst.e0_; ++i) dst.ptr_[i + j * dst.s1_] = static_cast<T>(src.ptr_[i + j * src.s1_]);
}
// host mirror of a (possibly device-resident) view: Scalar::access / Vect::access, types.hpp:39-46,104-112
template <class Space, class DT, class... P>
View<DT, LayoutLeft, HostSpace> create_mirror_view_and_copy(const Space&, const View<DT, P...>& v) {
    using T = typename View<DT, P...>::value_type;
    View<DT, LayoutLeft, HostSpace> h;
    if (!View<DT, P...>::on_device) { h.alloc_ = v.alloc_; h.ptr_ = v.ptr_; h.e0_ = v.e0_; h.e1_ = v.e1_; h.s1_ = v.s1_; return h; }
    const size_t e0 = v.e0_ ? v.e0_ : 1, e1 = v.e1_ ? v.e1_ : 1;
    T* p = static_cast<T*>(std::calloc(e0 * e1, sizeof(T)));
    h.alloc_ = std::shared_ptr<T>(p, [](T* q) { std::free(q); });
    h.ptr_ = p; h.e0_ = e0; h.e1_ = e1; h.s1_ = e0;
    for (size_t j = 0; j < e1; ++j) shim_cuda_memcpy(p + j * e0, v.ptr_ + j * v.s1_, e0 * sizeof(T), 2);
    return h;
}

// ---- parallel dispatch -------------------------------------------------------------------------------------
template <class ExecSpace = OpenMP>
struct RangePolicy {
    size_t begin, end;
    RangePolicy(size_t b, size_t e) : begin(b), end(e) {}
};

template <class Exec, class F>
void parallel_for(const RangePolicy<Exec>& p, const F& f) {
    const long b = (long)p.begin, e = (long)p.end;
#pragma omp parallel for schedule(static)
    for (long i = b; i < e; ++i) f((size_t)i);
}
#if defined(__CUDACC__)
namespace detail {
template <class F>
__global__ void shim_for_kernel(size_t b, size_t e, F f) {
    const size_t i = b + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < e) f(i);
}
template <class F, class T>
__global__ void shim_max_kernel(size_t b, size_t e, F f, T* out) {
    __shared__ T sm[256];
    T upd = -3.0e38f;
    for (size_t i = b + threadIdx.x; i < e; i += blockDim.x) f(i, upd);
    sm[threadIdx.x] = upd;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o && sm[threadIdx.x] < sm[threadIdx.x + o]) sm[threadIdx.x] = sm[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = sm[0];
}
}  // namespace detail
template <class F>
void parallel_for(const RangePolicy<Cuda>& p, const F& f) {
    if (p.end <= p.begin) return;
    const size_t n = p.end - p.begin;
    detail::shim_for_kernel<F><<<(unsigned)((n + 255) / 256), 256>>>(p.begin, p.end, f);
}
#endif
template <class Exec, class F>
void parallel_for(const std::string&, const RangePolicy<Exec>& p, const F& f) { parallel_for(p, f); }

template <class T>
struct Max {
    T& ref;
    explicit Max(T& r) : ref(r) {}
};
template <class Exec, class F, class T>
void parallel_reduce(const RangePolicy<Exec>& p, const F& f, Max<T> red) {
    T upd = std::numeric_limits<T>::lowest();
    for (size_t i = p.begin; i < p.end; ++i) f(i, upd);
    red.ref = upd;
}
#if defined(__CUDACC__)
template <class F, class T>
void parallel_reduce(const RangePolicy<Cuda>& p, const F& f, Max<T> red) {
    T* d = static_cast<T*>(shim_cuda_alloc_zeroed(sizeof(T)));
    detail::shim_max_kernel<F, T><<<1, 256>>>(p.begin, p.end, f, d);
    shim_cuda_memcpy(&red.ref, d, sizeof(T), 2);
    shim_cuda_free(d);
}
#endif
template <class Exec, class F, class T>
void parallel_reduce(const RangePolicy<Exec>& p, const F& f, T& sum) {
    T upd = T(0);
    for (size_t i = p.begin; i < p.end; ++i) f(i, upd);
    sum = upd;
}

}  // namespace Kokkos

#endif
