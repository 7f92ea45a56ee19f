// substrate.hpp — Kokkos-free host substrate of the B200 backend: the handle types and the operator surface of the reference
// (types.hpp:15-237, types_cuda.hpp:47-152, kernels.hpp:11-165, Orthogonalization.hpp:17-136, gmres.hpp:15-57) written
// directly on the C ABI of libmpgmres_b200.so.  A C++ caller that does not build Kokkos includes this ONE header and links
// the shared library; nothing here needs nvcc (device memory comes from mpg_malloc, transfers from mpg_memcpy_*).
//
//   b200::Scalar<T> / Vect<T> / MultiVect<T>     types.hpp:15-55 / 57-113 / 115-228
//       shallow, reference-counted handles on device memory (Kokkos::View semantics: copies alias, sub-views alias the
//       parent's storage and keep it alive, allocation zero-fills, access() is a synchronising device-to-host read);
//       sub-range / sub-block constructors, the transpose FLAG of MultiVect (no data movement, types.hpp:209-211) and the
//       "data() of a sub-block = block origin, stride() = parent column stride" rule (types.hpp:193-199) are reproduced
//       exactly, including the quirk that Vect(MultiVect, rows, col) ignores the flag (types.hpp:79-81) while
//       MultiVect::operator()(i, j) honours it (types.hpp:213-219).
//   b200::SparseMatrix<T>                         types_cuda.hpp:47-152 (+ SpMV plan and packed copy, shared with casts)
//   b200::dot / nrm2 / axpy / naxpy / scal / copy / fill / gdmv / rotg / rot / gemv / trsv / spmv     kernels.hpp:11-165
//   b200::Orthogonalization::GS<T, ORTH>          Orthogonalization.hpp:17-74 with the kernels of :76-136 fused on the device
//   b200::gmres_singleUpdate / gmres_baseline     gmres.hpp:15-32 (one call into mpg_gmres_solve)
// Errors: the reference surface returns void; like the reference's library-initialisation failure (types_cuda.hpp:15-24) a
// non-zero C-ABI status aborts with the library's message.  Define B200_SUBSTRATE_THROW to get std::runtime_error instead.
#ifndef B200_SUBSTRATE_HPP
#define B200_SUBSTRATE_HPP

#include <cassert>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../mpgmres_b200.h"

namespace b200 {

// ---- backend singleton: replaces CudaLibSingleton, types_cuda.hpp:9-36 -------------------------------------------------
struct Backend {
    mpg_ctx* ctx = nullptr;
    Backend() {
        if (mpg_ctx_create(0, &ctx) != MPG_OK) {
            std::fprintf(stderr, "mpgmres_b200 initialization failed (no CUDA device?)\n");
            std::abort();
        }
    }
    ~Backend() { mpg_ctx_destroy(ctx); }
    static Backend& singleton() {
        static Backend s;
        return s;
    }
    static mpg_ctx* context() { return singleton().ctx; }
    static void check(int rc, const char* what) {
        if (rc == MPG_OK) return;
        const std::string msg = std::string("mpgmres_b200: ") + what + " failed (" + std::to_string(rc) + "): " + mpg_last_error(singleton().ctx);
#ifdef B200_SUBSTRATE_THROW
        throw std::runtime_error(msg);
#else
        std::fprintf(stderr, "%s\n", msg.c_str());
        std::abort();
#endif
    }
};
#define B200S_CHECK(expr) ::b200::Backend::check((expr), #expr)
inline mpg_ctx* ctx() { return Backend::context(); }
inline void fence() { B200S_CHECK(mpg_sync(ctx())); }   // Device::execution_space().fence(), gmres.cpp:113,225

struct all_t {};
constexpr all_t ALL{};                         // Kokkos::ALL
typedef std::pair<size_t, size_t> range;       // Kokkos::pair<size_t, size_t>: [first, second)

namespace detail {
// zero-filled device allocation with shared ownership (a Kokkos::View allocation, types.hpp:18,60,118)
inline std::shared_ptr<void> device_alloc(size_t bytes) {
    void* p = nullptr;
    B200S_CHECK(mpg_malloc(ctx(), bytes ? bytes : 1, &p));
    return std::shared_ptr<void>(p, [](void* q) { mpg_free(Backend::context(), q); });
}
}  // namespace detail

template <class T> class Vect;
template <class T> class MultiVect;

// ---- Scalar: types.hpp:15-55 ---------------------------------------------------------------------------------------
template <class T>
class Scalar {
    std::shared_ptr<void> own_;
    T* p_ = nullptr;

public:
    Scalar() : own_(detail::device_alloc(sizeof(T))), p_(static_cast<T*>(own_.get())) {}
    Scalar(T val) : own_(detail::device_alloc(sizeof(T))), p_(static_cast<T*>(own_.get())) { B200S_CHECK(mpg_memcpy_h2d(ctx(), p_, &val, sizeof(T))); fence(); }
    Scalar(Vect<T> vec, size_t idx);
    Scalar(MultiVect<T> vec, size_t row, size_t col);   // indices into the BASE view (types.hpp:33-35)
    T access() const {                                  // synchronising device-to-host read (types.hpp:37-44)
        T v;
        B200S_CHECK(mpg_memcpy_d2h(ctx(), &v, p_, sizeof(T)));
        return v;
    }
    T* data() const { return p_; }
    const std::shared_ptr<void>& owner() const { return own_; }
};

// ---- Vect: types.hpp:57-113 ------------------------------------------------------------------------------------------
template <class T>
class Vect {
    std::shared_ptr<void> own_;
    T* p_ = nullptr;
    size_t n_ = 0;

public:
    Vect() {}
    explicit Vect(size_t n) : own_(detail::device_alloc(sizeof(T) * n)), p_(static_cast<T*>(own_.get())), n_(n) {}
    Vect(std::shared_ptr<void> owner, T* p, size_t n) : own_(std::move(owner)), p_(p), n_(n) {}   // wrap existing device memory
    Vect(const std::vector<T>& host) : Vect(host.size()) { upload(host.data()); }
    Vect(Vect<T> vec, range rows) : own_(vec.own_), p_(vec.p_ + rows.first), n_(rows.second - rows.first) { assert(rows.first <= rows.second && rows.second <= vec.n_); }
    Vect(MultiVect<T> vec, all_t, size_t col);
    Vect(MultiVect<T> vec, range rows, size_t col);      // rows / col of the BASE view: the transpose flag is ignored (types.hpp:79-81)
    Scalar<T> operator()(size_t i) const { return Scalar<T>(*this, i); }
    Vect<T> operator()(range rows) const { return Vect<T>(*this, rows); }
    T* data() const { return p_; }
    size_t n() const { return n_; }
    T access(size_t i) const {
        assert(i < n_);
        T v;
        B200S_CHECK(mpg_memcpy_d2h(ctx(), &v, p_ + i, sizeof(T)));
        return v;
    }
    const std::shared_ptr<void>& owner() const { return own_; }
    // Kokkos::deep_copy equivalents (gmres_perf_test.cpp:219-221)
    void upload(const T* host) { B200S_CHECK(mpg_memcpy_h2d(ctx(), p_, host, sizeof(T) * n_)); fence(); }
    std::vector<T> download() const {
        std::vector<T> h(n_);
        if (n_) B200S_CHECK(mpg_memcpy_d2h(ctx(), h.data(), p_, sizeof(T) * n_));
        return h;
    }
};

// ---- MultiVect: types.hpp:115-228 (LayoutLeft = column-major) ----------------------------------------------------------
template <class T>
class MultiVect {
    std::shared_ptr<void> own_;
    T* p_ = nullptr;
    size_t e0_ = 0, e1_ = 0, ld_ = 0;   // extents of THIS view, column stride of the allocation
    bool transposed_ = false;

    MultiVect(std::shared_ptr<void> own, T* p, size_t e0, size_t e1, size_t ld, bool tr) : own_(std::move(own)), p_(p), e0_(e0), e1_(e1), ld_(ld), transposed_(tr) {}
    // sub-view [r0, r1) x [c0, c1) of the base view
    static MultiVect sub(const MultiVect& v, range r, range c, bool tr) {
        assert(r.first <= r.second && r.second <= v.e0_ && c.first <= c.second && c.second <= v.e1_);
        return MultiVect(v.own_, v.p_ + r.first + c.first * v.ld_, r.second - r.first, c.second - c.first, v.ld_, tr);
    }

public:
    MultiVect() {}
    MultiVect(size_t m, size_t n) : own_(detail::device_alloc(sizeof(T) * m * n)), p_(static_cast<T*>(own_.get())), e0_(m), e1_(n), ld_(m) {}
    // the three sub-block constructors address rows / columns of the LOGICAL matrix: with the flag set they select columns /
    // rows of the base view (types.hpp:135-167)
    MultiVect(MultiVect<T> vec, range rows, all_t) { *this = vec.transposed_ ? sub(vec, range(0, vec.e0_), rows, true) : sub(vec, rows, range(0, vec.e1_), false); }
    MultiVect(MultiVect<T> vec, all_t, range cols) { *this = vec.transposed_ ? sub(vec, cols, range(0, vec.e1_), true) : sub(vec, range(0, vec.e0_), cols, false); }
    MultiVect(MultiVect<T> vec, range rows, range cols) { *this = vec.transposed_ ? sub(vec, cols, rows, true) : sub(vec, rows, cols, false); }
    size_t nrows() const { return transposed_ ? e1_ : e0_; }
    size_t ncols() const { return transposed_ ? e0_ : e1_; }
    size_t nrows_base() const { return e0_; }
    size_t ncols_base() const { return e1_; }
    T* data() const { return p_; }                 // block origin (types.hpp:193-195)
    size_t stride() const { return ld_; }          // parent column stride (types.hpp:197-199)
    size_t n() const { return e0_; }
    bool transposed() const { return transposed_; }
    MultiVect<T> transpose_matrix() const { return MultiVect<T>(own_, p_, e0_, e1_, ld_, !transposed_); }   // a flag flip (types.hpp:209-211)
    Scalar<T> operator()(size_t i, size_t j) const { return transposed_ ? Scalar<T>(*this, j, i) : Scalar<T>(*this, i, j); }
    Vect<T> operator()(range rows, size_t col) const { return Vect<T>(*this, rows, col); }
    const std::shared_ptr<void>& owner() const { return own_; }
    void upload(const T* host_colmajor) {          // e0 x e1, leading dimension e0
        for (size_t j = 0; j < e1_; ++j) B200S_CHECK(mpg_memcpy_h2d(ctx(), p_ + j * ld_, host_colmajor + j * e0_, sizeof(T) * e0_));
        fence();
    }
    std::vector<T> download() const {
        std::vector<T> h(e0_ * e1_);
        for (size_t j = 0; j < e1_; ++j) B200S_CHECK(mpg_memcpy_d2h(ctx(), h.data() + j * e0_, p_ + j * ld_, sizeof(T) * e0_));
        return h;
    }
};

template <class T> Scalar<T>::Scalar(Vect<T> vec, size_t idx) : own_(vec.owner()), p_(vec.data() + idx) { assert(idx < vec.n()); }
template <class T> Scalar<T>::Scalar(MultiVect<T> vec, size_t row, size_t col) : own_(vec.owner()), p_(vec.data() + row + col * vec.stride()) {
    assert(row < vec.nrows_base() && col < vec.ncols_base());
}
template <class T> Vect<T>::Vect(MultiVect<T> vec, all_t, size_t col) : own_(vec.owner()), p_(vec.data() + col * vec.stride()), n_(vec.nrows_base()) { assert(col < vec.ncols_base()); }
template <class T> Vect<T>::Vect(MultiVect<T> vec, range rows, size_t col) : own_(vec.owner()), p_(vec.data() + rows.first + col * vec.stride()), n_(rows.second - rows.first) {
    assert(rows.first <= rows.second && rows.second <= vec.nrows_base() && col < vec.ncols_base());
}

// ---- operator surface: kernels.hpp:11-165 ------------------------------------------------------------------------------
#define B200S_SURFACE(T, SFX)                                                                                                               \
    inline T dot(Vect<T> x, Vect<T> y) { assert(x.n() == y.n()); T r; B200S_CHECK(mpg_dot_##SFX(ctx(), (int64_t)x.n(), x.data(), y.data(), &r)); return r; }     \
    inline void dot(Vect<T> x, Vect<T> y, Scalar<T> result) { assert(x.n() == y.n()); B200S_CHECK(mpg_dot_dev_##SFX(ctx(), (int64_t)x.n(), x.data(), y.data(), result.data())); } \
    inline T nrm2(Vect<T> x) { T r; B200S_CHECK(mpg_nrm2_##SFX(ctx(), (int64_t)x.n(), x.data(), &r)); return r; }                                \
    inline void nrm2(Vect<T> x, Scalar<T> result) { B200S_CHECK(mpg_nrm2_dev_##SFX(ctx(), (int64_t)x.n(), x.data(), result.data())); }          \
    inline void axpy(T alpha, Vect<T> x, Vect<T> y) { assert(x.n() == y.n()); B200S_CHECK(mpg_axpy_##SFX(ctx(), (int64_t)x.n(), alpha, x.data(), y.data())); }  \
    inline void axpy(Scalar<T> alpha, Vect<T> x, Vect<T> y) { assert(x.n() == y.n()); B200S_CHECK(mpg_axpy_dev_##SFX(ctx(), (int64_t)x.n(), alpha.data(), x.data(), y.data())); } \
    inline void naxpy(Scalar<T> alpha, Vect<T> x, Vect<T> y) { assert(x.n() == y.n()); B200S_CHECK(mpg_naxpy_dev_##SFX(ctx(), (int64_t)x.n(), alpha.data(), x.data(), y.data())); } \
    inline void scal(T alpha, Vect<T> x) { B200S_CHECK(mpg_scal_##SFX(ctx(), (int64_t)x.n(), alpha, x.data(), x.data())); }                      \
    inline void scal(T alpha, Vect<T> x, Vect<T> y) { assert(x.n() == y.n()); B200S_CHECK(mpg_scal_##SFX(ctx(), (int64_t)x.n(), alpha, x.data(), y.data())); }  \
    inline void scal(Scalar<T> alpha, Vect<T> x, Vect<T> y) { assert(x.n() == y.n()); B200S_CHECK(mpg_scal_dev_##SFX(ctx(), (int64_t)x.n(), alpha.data(), x.data(), y.data())); } \
    inline void scal(T alpha, Scalar<T> x, Scalar<T> y) { B200S_CHECK(mpg_scal_##SFX(ctx(), 1, alpha, x.data(), y.data())); }                    \
    inline void scal(Scalar<T> alpha, Scalar<T> x, Scalar<T> y) { B200S_CHECK(mpg_scal_dev_##SFX(ctx(), 1, alpha.data(), x.data(), y.data())); } \
    inline void fill(T alpha, Vect<T> x) { B200S_CHECK(mpg_fill_##SFX(ctx(), (int64_t)x.n(), alpha, x.data())); }                                \
    inline void fill(T alpha, Scalar<T> x) { B200S_CHECK(mpg_fill_##SFX(ctx(), 1, alpha, x.data())); }                                           \
    inline void gdmv(T alpha, Vect<T> diag, Vect<T> x, T beta, Vect<T> y) {                                                                       \
        assert(diag.n() == x.n() && x.n() == y.n());                                                                                              \
        B200S_CHECK(mpg_gdmv_##SFX(ctx(), (int64_t)diag.n(), alpha, diag.data(), x.data(), beta, y.data()));                                     \
    }                                                                                                                                             \
    inline void rotg(Scalar<T> a, Scalar<T> b, Scalar<T> c, Scalar<T> s) { B200S_CHECK(mpg_rotg_##SFX(ctx(), a.data(), b.data(), c.data(), s.data())); }        \
    inline void rot(Scalar<T> a, Scalar<T> b, Scalar<T> c, Scalar<T> s) { B200S_CHECK(mpg_rot_##SFX(ctx(), a.data(), b.data(), c.data(), s.data())); }          \
    /* c.n() rotations applied to a[0 .. c.n()] (kernels_cuda.cpp:448-494) */                                                                     \
    inline void rot(Vect<T> a, Vect<T> c, Vect<T> s) { assert(c.n() == s.n()); B200S_CHECK(mpg_rot_vec_##SFX(ctx(), (int64_t)c.n(), a.data(), c.data(), s.data())); } \
    /* op(M) from the transpose flag; base dims + stride go to the library like the reference's cublas call (kernels_cuda.cpp:499-535) */         \
    inline void gemv(T alpha, MultiVect<T> matrix, Vect<T> x, T beta, Vect<T> y) {                                                                \
        assert(matrix.ncols() == x.n() && matrix.nrows() == y.n());                                                                               \
        B200S_CHECK(mpg_gemv_##SFX(ctx(), matrix.transposed() ? 1 : 0, (int64_t)matrix.nrows_base(), (int64_t)matrix.ncols_base(), alpha, matrix.data(),       \
                                   (int64_t)matrix.stride(), x.data(), beta, y.data()));                                                          \
    }                                                                                                                                             \
    inline void trsv(const char* upper, MultiVect<T> matrix, Vect<T> x) {                                                                         \
        assert(matrix.ncols() == matrix.nrows() && matrix.ncols() == x.n());                                                                      \
        B200S_CHECK(mpg_trsv_##SFX(ctx(), 'U' == *upper, matrix.transposed() ? 1 : 0, (int64_t)matrix.nrows(), matrix.data(), (int64_t)matrix.stride(), x.data())); \
    }
B200S_SURFACE(float, f32)
B200S_SURFACE(double, f64)
#undef B200S_SURFACE

// copy with implicit type conversion: the fp64 <-> fp32 casts (kernels.hpp:11-30)
inline void copy(Vect<double> x, Vect<float> y) { assert(x.n() == y.n()); B200S_CHECK(mpg_copy_f64_f32(ctx(), (int64_t)x.n(), x.data(), y.data())); }
inline void copy(Vect<float> x, Vect<double> y) { assert(x.n() == y.n()); B200S_CHECK(mpg_copy_f32_f64(ctx(), (int64_t)x.n(), x.data(), y.data())); }
inline void copy(Vect<float> x, Vect<float> y) { assert(x.n() == y.n()); B200S_CHECK(mpg_copy_f32_f32(ctx(), (int64_t)x.n(), x.data(), y.data())); }
inline void copy(Vect<double> x, Vect<double> y) { assert(x.n() == y.n()); B200S_CHECK(mpg_copy_f64_f64(ctx(), (int64_t)x.n(), x.data(), y.data())); }
inline void copy(Scalar<double> x, Scalar<float> y) { B200S_CHECK(mpg_copy_f64_f32(ctx(), 1, x.data(), y.data())); }
inline void copy(Scalar<float> x, Scalar<double> y) { B200S_CHECK(mpg_copy_f32_f64(ctx(), 1, x.data(), y.data())); }
inline void copy(Scalar<float> x, Scalar<float> y) { B200S_CHECK(mpg_copy_f32_f32(ctx(), 1, x.data(), y.data())); }
inline void copy(Scalar<double> x, Scalar<double> y) { B200S_CHECK(mpg_copy_f64_f64(ctx(), 1, x.data(), y.data())); }

// ---- SparseMatrix: types_cuda.hpp:47-152 ------------------------------------------------------------------------------
template <class T>
class SparseMatrix {
    std::shared_ptr<mpg_csr> plan_;        // SpMV plan over (row_map_, inds_); shared with precision-cast copies
    std::shared_ptr<mpg_packed> packed_;   // packed copy (the analogue of create_cuda_handles, types_cuda.hpp:53-60); null if it does not pack
    bool transposed_ = false;

    static int pack_create(const mpg_csr* A, const float* v, mpg_packed** out) { return mpg_pack_create_f32(ctx(), A, v, out); }
    static int pack_create(const mpg_csr* A, const double* v, mpg_packed** out) { return mpg_pack_create_f64(ctx(), A, v, out); }
    void create_plan() {
        mpg_csr* p = nullptr;
        B200S_CHECK(mpg_csr_create(ctx(), m_, n_, (int64_t)nnz_, row_map_.data(), inds_.data(), &p));
        plan_ = std::shared_ptr<mpg_csr>(p, [](mpg_csr* q) { mpg_csr_destroy(q); });
    }
    void create_packed() {
        mpg_packed* q = nullptr;
        B200S_CHECK(pack_create(plan_.get(), vals_.data(), &q));
        if (q) packed_ = std::shared_ptr<mpg_packed>(q, [](mpg_packed* r) { mpg_pack_destroy(r); });
    }
    template <class> friend class SparseMatrix;

public:
    int m_ = 0, n_ = 0, nnz_ = 0;
    Vect<int> row_map_, inds_;
    Vect<T> vals_;

    SparseMatrix() {}
    SparseMatrix(int m, int n, Vect<int> row_map, Vect<int> inds, Vect<T> vals) : m_(m), n_(n), nnz_((int)inds.n()), row_map_(row_map), inds_(inds), vals_(vals) {
        create_plan();
        create_packed();
    }
    // host CSR -> device (types_cuda.hpp:103-114)
    SparseMatrix(int m, int n, const std::vector<int>& row_map, const std::vector<int>& inds, const std::vector<T>& vals)
        : SparseMatrix(m, n, Vect<int>(row_map), Vect<int>(inds), Vect<T>(vals)) {}
    // precision cast: shares row_map / inds / plan, converts the values (types_cuda.hpp:82-101)
    template <class Old>
    SparseMatrix(const SparseMatrix<Old>& old)
        : plan_(old.plan_), transposed_(old.transposed_), m_(old.m_), n_(old.n_), nnz_(old.nnz_), row_map_(old.row_map_), inds_(old.inds_), vals_(old.vals_.n()) {
        copy(old.vals_, vals_);
        create_packed();
    }
    int nrows() const { return m_; }
    int ncols() const { return n_; }
    int nnz() const { return nnz_; }
    int* row_map_data() const { return row_map_.data(); }
    int* inds_data() const { return inds_.data(); }
    T* vals_data() const { return vals_.data(); }
    Vect<T> vals_vect() const { return vals_; }
    const mpg_csr* plan() const { return plan_.get(); }
    const mpg_packed* packed() const { return packed_.get(); }
    void set_transpose(bool t) { transposed_ = t; }   // condest.cpp only; transposed SpMV is not provided
    bool is_transposed() const { return transposed_; }
};

inline void spmv(float alpha, const SparseMatrix<float>& A, Vect<float> x, float beta, Vect<float> y) {   // kernels.hpp:159-160
    assert((size_t)A.ncols() == x.n() && (size_t)A.nrows() == y.n() && !A.is_transposed());
    if (A.packed()) B200S_CHECK(mpg_spmv_packed_f32(ctx(), A.packed(), alpha, x.data(), beta, y.data()));
    else B200S_CHECK(mpg_spmv_f32(ctx(), A.plan(), A.vals_data(), alpha, x.data(), beta, y.data()));
}
inline void spmv(double alpha, const SparseMatrix<double>& A, Vect<double> x, double beta, Vect<double> y) {
    assert((size_t)A.ncols() == x.n() && (size_t)A.nrows() == y.n() && !A.is_transposed());
    if (A.packed()) B200S_CHECK(mpg_spmv_packed_f64(ctx(), A.packed(), alpha, x.data(), beta, y.data()));
    else B200S_CHECK(mpg_spmv_f64(ctx(), A.plan(), A.vals_data(), alpha, x.data(), beta, y.data()));
}

// ---- LinearOperator / Identity / Jacobi: types.hpp:230-237,374-448 -----------------------------------------------------
template <class T>
class LinearOperator {
public:
    virtual ~LinearOperator() {}
    virtual void apply(Vect<T> rhs) = 0;
};
template <class T>
class Identity : public LinearOperator<T> {
public:
    void apply(Vect<T>) override {}
};
template <class T>
class Jacobi : public LinearOperator<T> {
    Vect<T> diag_;

public:
    explicit Jacobi(const SparseMatrix<T>& A) : diag_((size_t)A.nrows()) {   // get_diag_vals, types.hpp:395-430
        if (sizeof(T) == 4) B200S_CHECK(mpg_jacobi_diag_f32(ctx(), A.plan(), (const float*)A.vals_data(), (float*)diag_.data()));
        else B200S_CHECK(mpg_jacobi_diag_f64(ctx(), A.plan(), (const double*)A.vals_data(), (double*)diag_.data()));
    }
    void apply(Vect<T> rhs) override { gdmv(T(1), diag_, rhs, T(0), rhs); }   // types.hpp:444-446
    Vect<T> diag() const { return diag_; }
};

// ---- Orthogonalization::GS: Orthogonalization.hpp:17-74; the kernels of :76-136 are one fused device routine ------------
namespace Orthogonalization {
enum Kind { CGS = MPG_ORTH_CGS, MGS = MPG_ORTH_MGS, CGSR2 = MPG_ORTH_CGSR };

template <class T, Kind ORTH>
class GS {
public:
    MultiVect<T> v;
    GS(size_t n, size_t max_restart_length) : v(n, max_restart_length + 1) {}
    MultiVect<T> basis() { return v; }
    T first_vector(const Vect<T> w) {                    // :36-45
        const T beta = nrm2(w);
        Vect<T> v_col(v, ALL, 0);
        if (beta != 0) scal(1 / beta, w, v_col);
        else fill(T(0), v_col);
        return beta;
    }
    Vect<T> previous_krylov_vector(size_t k) { return Vect<T>(v, ALL, k); }
    // orthogonalise + norm -> h(k+1,k) + V(:,k+1) = w / h(k+1,k) in 3 passes over the basis, no host read-back (:51-60)
    void add_vector(const size_t k, Vect<T> w, MultiVect<T> h) {
        T* hcol = h.data() + k * h.stride();
        if (sizeof(T) == 4) B200S_CHECK(mpg_add_vector_f32(ctx(), ORTH, (int64_t)v.nrows_base(), (int64_t)k, (float*)v.data(), (int64_t)v.stride(), (float*)w.data(), (float*)hcol));
        else B200S_CHECK(mpg_add_vector_f64(ctx(), ORTH, (int64_t)v.nrows_base(), (int64_t)k, (double*)v.data(), (int64_t)v.stride(), (double*)w.data(), (double*)hcol));
    }
    void update_x(const size_t k, const Vect<T> y, Vect<T> x) const {   // :62-65
        MultiVect<T> v_cols(v, ALL, range(0, k));
        gemv(T(1), v_cols, y, T(1), x);
    }
    template <class High>
    void update_x(const size_t k, const Vect<T> y, Vect<High> x, Vect<T> x_inc_temp, Vect<High> x_temp) const {   // :67-73
        MultiVect<T> v_cols(v, ALL, range(0, k));
        gemv(T(1), v_cols, y, T(0), x_inc_temp);
        copy(x_inc_temp, x_temp);
        axpy(High(1), x_temp, x);
    }
};
}  // namespace Orthogonalization

// ---- drivers: gmres.hpp:15-32 behind mpg_gmres_solve ------------------------------------------------------------------
struct SolveResult {
    mpg_gmres_stats stats;
    std::vector<double> hist_inner;   // |s(k+1)| / ||M^-1 b|| per inner iteration
};
struct SolveOptions {                 // gmres_perf_test.cpp:313-339 (--rlen --tol --rtol --max-restarts, the Convergence subclass, --prec)
    int64_t restart_length = 50;
    double tol = 1e-6, restart_tol = 0;
    int64_t max_restarts = 1000000;
    int conv = MPG_CONV_BASE;
    int prec = MPG_PREC_IDENTITY;
    int orth = MPG_ORTH_CGSR;
    int64_t jacobi_steps = 1;         // --jacobi-steps, for prec = MPG_PREC_ILU_JACOBI
};
namespace detail {
inline SolveResult solve(int mode, const SolveOptions& o, const SparseMatrix<double>& A, const float* vals32, Vect<double> b, Vect<double> x) {
    mpg_gmres_params p;
    p.mode = mode; p.orth = o.orth; p.conv = o.conv; p.prec = o.prec;
    p.restart_length = o.restart_length; p.tol = o.tol; p.restart_tol = o.restart_tol; p.max_restarts = o.max_restarts;
    p.jacobi_steps = o.jacobi_steps;
    SolveResult r;
    const int64_t cap = std::min<int64_t>((o.max_restarts + 2) * o.restart_length, 4000000);
    r.hist_inner.assign((size_t)cap, 0.0);
    B200S_CHECK(mpg_gmres_solve(ctx(), &p, A.plan(), A.vals_data(), vals32, b.data(), x.data(), &r.stats, r.hist_inner.data(), cap, nullptr, 0));
    r.hist_inner.resize((size_t)std::min<int64_t>(r.stats.n_hist_inner, cap));
    return r;
}
}  // namespace detail
// GMRES-IR: fp32 inner cycle, fp64 residual and update (gmres.cpp:135-245)
inline SolveResult gmres_singleUpdate(const SolveOptions& o, const SparseMatrix<double>& A, const SparseMatrix<float>& A_single, Vect<double> b, Vect<double> x) {
    return detail::solve(MPG_MODE_MIXED, o, A, A_single.vals_data(), b, x);
}
// uniform precision (gmres.cpp:24-133): Type / PrecType = double/double, double/float, float/float
inline SolveResult gmres_baseline(const SolveOptions& o, const SparseMatrix<double>& A, Vect<double> b, Vect<double> x, bool single_type = false, bool single_prec = false) {
    const int mode = single_type ? MPG_MODE_SINGLE : (single_prec ? MPG_MODE_SINGLE_PREC : MPG_MODE_BASELINE);
    return detail::solve(mode, o, A, nullptr, b, x);
}

}  // namespace b200
#endif  // B200_SUBSTRATE_HPP
