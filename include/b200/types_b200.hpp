// types_b200.hpp — what a maintainer of iamsonderr/icl-mixed-precision-gmres adds next to types_cuda.hpp to get a
// `B200` device backed by libmpgmres_b200.so.  It is written against the REFERENCE's own headers
// (types.hpp, kernels.hpp, Orthogonalization.hpp) and Kokkos, exactly like types_cuda.hpp (types_cuda.hpp:1-255):
//   * B200Backend          replaces CudaLibSingleton            types_cuda.hpp:9-36   (one mpg_ctx per process)
//   * struct B200          replaces struct Cuda                 types_cuda.hpp:39-44
//   * SparseMatrix<T,B200> replaces SparseMatrix<T,Cuda>        types_cuda.hpp:47-152 (+ the SpMV plan, shared between
//                                                               the fp64 matrix and its fp32 copy like row_map/inds)
//   * ILU<T,B200>, ILU_Jacobi_handles<B200>: ILU(0) factors + Jacobi-sweep application; exact triangular solves not provided
// The operator surface itself is specialised in kernels_b200.cpp.
#ifndef TYPES_B200_HPP
#define TYPES_B200_HPP

#include <cstdio>
#include <cstdlib>
#include <memory>

#include "kernels.hpp"
#include "Orthogonalization.hpp"
#include "mpgmres_b200.h"

struct B200Backend {
    mpg_ctx* ctx = nullptr;
    B200Backend() {
        if (mpg_ctx_create(0, &ctx) != MPG_OK) Kokkos::abort("mpgmres_b200 initialization failed (no CUDA device?)\n");
        // Kokkos::Cuda's default instance and the reference's cuBLAS/cuSPARSE handles run on the legacy default stream
        mpg_ctx_set_stream(ctx, nullptr);
    }
    static B200Backend& singleton() {
        static B200Backend s;
        return s;
    }
    static mpg_ctx* context() { return singleton().ctx; }
    static void check(int rc, const char* what) {
        if (rc != MPG_OK) {
            std::fprintf(stderr, "mpgmres_b200: %s failed (%d): %s\n", what, rc, mpg_last_error(singleton().ctx));
            std::abort();
        }
    }
};
#define B200_CHECK(expr) B200Backend::check((expr), #expr)

struct B200 {
public:
    typedef Kokkos::CudaSpace memory_space;
    typedef Kokkos::Cuda execution_space;
};

template <class Type>
class SparseMatrix<Type, B200> {
private:
    std::shared_ptr<mpg_csr> plan_;  // SpMV plan over (row_map_, inds_); shared with precision-cast copies
    // packed (sliced-ELL) copy of this matrix, the analogue of the cusparse handles create_cuda_handles() builds per
    // SparseMatrix (types_cuda.hpp:53-60); null when the structure does not pack well.  Declared after plan_: it refers
    // to the plan and is destroyed first.  Values are taken at construction (the reference never mutates a matrix).
    std::shared_ptr<mpg_packed> packed_;
    bool transposed_ = false;

    static int pack_create(const mpg_csr* A, const float* v, mpg_packed** out) { return mpg_pack_create_f32(B200Backend::context(), A, v, out); }
    static int pack_create(const mpg_csr* A, const double* v, mpg_packed** out) { return mpg_pack_create_f64(B200Backend::context(), A, v, out); }

    void create_plan() {
        mpg_csr* p = nullptr;
        B200_CHECK(mpg_csr_create(B200Backend::context(), m_, n_, nnz_, row_map_.data(), inds_.data(), &p));
        plan_ = std::shared_ptr<mpg_csr>(p, [](mpg_csr* q) { mpg_csr_destroy(q); });
    }
    void create_packed() {
        mpg_packed* q = nullptr;
        B200_CHECK(pack_create(plan_.get(), vals_.data(), &q));
        if (q) packed_ = std::shared_ptr<mpg_packed>(q, [](mpg_packed* r) { mpg_pack_destroy(r); });
    }

    template <class, class>
    friend class SparseMatrix;

public:
    int m_, n_, nnz_;

    Kokkos::View<int*, typename B200::memory_space> row_map_;
    Kokkos::View<int*, typename B200::memory_space> inds_;
    Kokkos::View<Type*, typename B200::memory_space> vals_;

    SparseMatrix(int m, int n, Kokkos::View<int*, typename B200::memory_space> row_map, Kokkos::View<int*, typename B200::memory_space> inds,
                 Kokkos::View<Type*, typename B200::memory_space> vals)
        : m_(m), n_(n), nnz_(inds.extent(0)), row_map_(row_map), inds_(inds), vals_(vals) {
        create_plan();
        create_packed();
    }

    // precision cast: shares row_map / inds / plan, converts the values (types_cuda.hpp:82-101)
    template <class OldType>
    SparseMatrix(SparseMatrix<OldType, B200> old)
        : plan_(old.plan_), transposed_(old.is_transposed()), m_(old.m_), n_(old.n_), nnz_(old.inds_.extent(0)), row_map_(old.row_map_),
          inds_(old.inds_), vals_("vals", old.vals_.extent(0)) {
        copy(Vect<OldType, B200>(old.vals_), Vect<Type, B200>(vals_));
        create_packed();
    }

    // host -> device (types_cuda.hpp:103-114)
    template <class OldDevice>
    SparseMatrix(SparseMatrix<Type, OldDevice> old)
        : transposed_(old.is_transposed()), m_(old.nrows()), n_(old.ncols()), nnz_(old.inds_.extent(0)), row_map_("row_map", old.row_map_.extent(0)),
          inds_("inds", old.inds_.extent(0)), vals_("vals", old.vals_.extent(0)) {
        Kokkos::deep_copy(row_map_, old.row_map_);
        Kokkos::deep_copy(inds_, old.inds_);
        Kokkos::deep_copy(vals_, old.vals_);
        create_plan();
        create_packed();
    }

    int nrows() const { return m_; }
    int ncols() const { return n_; }
    int nnz() const { return nnz_; }
    int* row_map_data() { return row_map_.data(); }
    int* inds_data() { return inds_.data(); }
    Type* vals_data() { return vals_.data(); }
    Vect<Type, B200> vals_vect() { return Vect<Type, B200>(vals_); }
    const mpg_csr* plan() const { return plan_.get(); }
    const mpg_packed* packed() const { return packed_.get(); }
    void set_transpose(bool new_trans) { this->transposed_ = new_trans; }  // only condest.cpp uses it (out of scope)
    bool is_transposed() { return this->transposed_; }
};

// ILU(0): ilu0<Type,B200> factors on the device (level-scheduled IKJ, kernels_b200.cpp); ILU_Jacobi<Type,B200> - the class of
// types.hpp:251-372, unchanged - applies the factors with Jacobi sweeps through ILU_Jacobi_handles<B200>, one fused launch per
// sweep.  The EXACT triangular solves of ILU<>::apply (cusparse csrsv2 in the reference, kernels_cuda.cpp:617-695; the API is gone
// from CUDA 12) are not provided: --prec ilu with this device aborts with a message, --prec ilu_jacobi works.
template <class Type>
class ILU<Type, B200> : public LinearOperator<Type, B200> {
private:
    int n_ = 0, nnz_ = 0;
    Kokkos::View<int*, typename B200::memory_space> row_map_;
    Kokkos::View<int*, typename B200::memory_space> inds_;
    Kokkos::View<Type*, typename B200::memory_space> vals_;

    template <class, class>
    friend class ILU;
    template <class, class>
    friend class ILU_Jacobi;

public:
    ILU() {}
    ILU(int n, int nnz, Kokkos::View<int*, typename B200::memory_space> row_map, Kokkos::View<int*, typename B200::memory_space> inds,
        Kokkos::View<Type*, typename B200::memory_space> vals)
        : n_(n), nnz_(nnz), row_map_(row_map), inds_(inds), vals_(vals) {}
    int n() const { return n_; }
    int nnz() const { return nnz_; }
    int* row_map_data() { return row_map_.data(); }
    int* inds_data() { return inds_.data(); }
    Type* vals_data() { return vals_.data(); }
    void apply(Vect<Type, B200>) { Kokkos::abort("exact ILU triangular solves are not provided by the B200 backend (use --prec ilu_jacobi, jacobi or identity)\n"); }
};
// per-preconditioner library objects: structure plan + the split / packed L and U operators (mpg_ilu_jacobi)
template <>
class ILU_Jacobi_handles<B200> {
public:
    std::shared_ptr<mpg_csr> plan;
    std::shared_ptr<mpg_ilu_jacobi> h;
    template <class Type>
    ILU_Jacobi_handles(const ILU_Jacobi<Type, B200>& ilu_const);
};

template <> void ilusv_jacobi<float, B200>(ILU_Jacobi<float, B200> ilu, Vect<float, B200> x);
template <> void ilusv_jacobi<double, B200>(ILU_Jacobi<double, B200> ilu, Vect<double, B200> x);
template <> void ilu_jacobi_mv<float, B200>(bool lower, float alpha, ILU_Jacobi<float, B200> ilu, Vect<float, B200> x, float beta, Vect<float, B200> y);
template <> void ilu_jacobi_mv<double, B200>(bool lower, double alpha, ILU_Jacobi<double, B200> ilu, Vect<double, B200> x, double beta, Vect<double, B200> y);

template <class Type>
ILU_Jacobi_handles<B200>::ILU_Jacobi_handles(const ILU_Jacobi<Type, B200>& ilu_const) {
    ILU_Jacobi<Type, B200>& ilu = const_cast<ILU_Jacobi<Type, B200>&>(ilu_const);   // the accessors of types.hpp:318-349 are not const
    mpg_ctx* ctx = B200Backend::context();
    mpg_csr* p = nullptr;
    B200_CHECK(mpg_csr_create(ctx, ilu.n(), ilu.n(), ilu.nnz(), ilu.row_map_data(), ilu.inds_data(), &p));
    plan = std::shared_ptr<mpg_csr>(p, [](mpg_csr* q) { mpg_csr_destroy(q); });
    // the library takes the factors in fp64 (what ilu0 produced before type_convert); widening Type -> double is exact
    Kokkos::View<double*, typename B200::memory_space> wide("ilu::vals64", (size_t)ilu.nnz());
    copy(Vect<Type, B200>(ilu.vals_view()), Vect<double, B200>(wide));
    mpg_ilu_jacobi* m = nullptr;
    if (sizeof(Type) == 4) B200_CHECK(mpg_ilu_jacobi_create_f32(ctx, p, wide.data(), ilu.steps(), &m));
    else B200_CHECK(mpg_ilu_jacobi_create_f64(ctx, p, wide.data(), ilu.steps(), &m));
    B200_CHECK(mpg_sync(ctx));
    h = std::shared_ptr<mpg_ilu_jacobi>(m, [](mpg_ilu_jacobi* q) { mpg_ilu_jacobi_destroy(q); });
}

// ---- specialisations of the header-inline generic operators (kernels.hpp:11-20,131-146) -------------------------------
// declared here, before the driver uses them; defined in kernels_b200.cpp
template <> void copy<double, float, B200>(Vect<double, B200> x, Vect<float, B200> y);
template <> void copy<float, double, B200>(Vect<float, B200> x, Vect<double, B200> y);
template <> void copy<float, float, B200>(Vect<float, B200> x, Vect<float, B200> y);
template <> void copy<double, double, B200>(Vect<double, B200> x, Vect<double, B200> y);
template <> void gdmv<float, B200>(float alpha, Vect<float, B200> diag, Vect<float, B200> x, float beta, Vect<float, B200> y);
template <> void gdmv<double, B200>(double alpha, Vect<double, B200> diag, Vect<double, B200> x, double beta, Vect<double, B200> y);

// ---- fused Arnoldi step behind the unchanged class API of Orthogonalization.hpp:51-60 ---------------------------------
// (orthogonalise + norm + normalise in 3 passes over the basis and no host read-back of h(k+1,k))
namespace Orthogonalization {
template <> void GS<float, CGS_Kernel<float, B200>, B200>::add_vector(const size_t k, Vect<float, B200> w, MultiVect<float, B200> h);
template <> void GS<float, MGS_Kernel<float, B200>, B200>::add_vector(const size_t k, Vect<float, B200> w, MultiVect<float, B200> h);
template <> void GS<float, CGSR_Kernel<float, B200, 2>, B200>::add_vector(const size_t k, Vect<float, B200> w, MultiVect<float, B200> h);
template <> void GS<double, CGS_Kernel<double, B200>, B200>::add_vector(const size_t k, Vect<double, B200> w, MultiVect<double, B200> h);
template <> void GS<double, MGS_Kernel<double, B200>, B200>::add_vector(const size_t k, Vect<double, B200> w, MultiVect<double, B200> h);
template <> void GS<double, CGSR_Kernel<double, B200, 2>, B200>::add_vector(const size_t k, Vect<double, B200> w, MultiVect<double, B200> h);
}  // namespace Orthogonalization

#endif  // TYPES_B200_HPP
