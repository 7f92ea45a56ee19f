// ORACLE — TEST INFRASTRUCTURE ONLY.
// Thin C entry point over the REFERENCE'S OWN solver templates (gmres.hpp: gmres_singleUpdate / gmres_baseline, compiled
// unmodified from /root/reference by oracle/ref.mk together with its kernels_mkl.cpp backend).  What is written here
// is only the glue DoMixedPrecisionProblem / DoBaselineProblem do in gmres_perf_test.cpp:53-182 (fp32 copy of A,
// preconditioner choice, chrono window around the solver call, fp64 post-solve norms) plus a Convergence subclass
// that records what the reference's virtual check()/check_initial() hooks are handed — the residual history the
// reference itself never prints.
#include <chrono>
#include <cstdint>
#include <cstring>
#include <cmath>
#include <iostream>
#include <random>
#include <sstream>
#include <vector>

#include <Kokkos_Core.hpp>

#include "types_mkl.hpp"
#include "kernels.hpp"
#include "gmres.hpp"
#include "IterUtil.hpp"
#include "Orthogonalization.hpp"
#include "LoadMatrix.hpp"

namespace {

template <class Base>
struct Logged : public Base {
    using Base::Base;
    std::vector<double> inner, outer;
    iteration_action check_initial(double r, double nrm, double pr, double pb) override {
        outer.push_back(r); outer.push_back(nrm); outer.push_back(pr); outer.push_back(pb);
        return Base::check_initial(r, nrm, pr, pb);
    }
    iteration_action check(size_t k, double r, double bn) override {
        inner.push_back(r / bn);
        return Base::check(k, r, bn);
    }
};

struct RefStats {
    int64_t status, total_iters, total_restarts, outer_i;
    double rel_prec_res, res_norm, err_norm, gmres_seconds, prec_seconds;
    int64_t n_hist_inner, n_hist_outer;
};

template <class T>
Vect<T, MKL> wrap(T* p, size_t n) {
    return Vect<T, MKL>(Kokkos::View<T*, Kokkos::LayoutLeft, Kokkos::HostSpace>(p, n));
}

template <class T, class Conv>
LinearOperator<T, MKL>* make_prec(int prec, SparseMatrix<double, MKL> A) {
    // gmres_perf_test.cpp:69-90,138-159 (identity / jacobi; ilu variants are out of scope)
    if (prec == 1) return new Jacobi<T, MKL>(A);
    return new Identity<T, MKL>();
}

template <class ConvT, class T, class... Args>
ConvT* make_conv(Args... a) { return new ConvT(a...); }

// mode 0: DoMixedPrecisionProblem; modes 1-3: DoBaselineProblem<ORTH, MKL, Type, PrecType>
template <template <class, class> class Kernel, class ConvF, class ConvD>
void run(int mode, int prec, SparseMatrix<double, MKL> A, Vect<double, MKL> b, Vect<double, MKL> x_out, Vect<double, MKL> true_x, ConvF* cf,
         ConvD* cd, RefStats* st) {
    const int n = A.nrows();
    using clk = std::chrono::high_resolution_clock;
    if (mode == 0) {
        using ORTH = Orthogonalization::GS<float, Kernel<float, MKL>, MKL>;
        Vect<double, MKL> x(n);
        copy(x_out, x);  // x0
        auto p0 = clk::now();
        const SparseMatrix<float, MKL> A_single(A);                      // gmres_perf_test.cpp:136
        LinearOperator<float, MKL>* M = make_prec<float, ConvF>(prec, A);
        st->prec_seconds = std::chrono::duration<double>(clk::now() - p0).count();
        auto t0 = clk::now();
        gmres_singleUpdate<ORTH, MKL>(*cf, A, A_single, M, b, x);        // :166
        st->gmres_seconds = std::chrono::duration<double>(clk::now() - t0).count();
        Vect<double, MKL> r(n);
        copy(b, r);
        spmv(-1.0, A, x, 1.0, r);                                        // :169-172
        st->res_norm = nrm2(r);
        copy(x, x_out);
        axpy(-1.0, true_x, x);
        st->err_norm = nrm2(x);
        delete M;
    } else if (mode == 1 || mode == 2) {
        using ORTH = Orthogonalization::GS<double, Kernel<double, MKL>, MKL>;
        auto p0 = clk::now();
        const SparseMatrix<float, MKL> A_type(A);                        // :66 (the fp32-rounded matrix quirk)
        Vect<double, MKL> x_type(n);
        copy(x_out, x_type);
        Vect<double, MKL> b_type(n);
        copy(b, b_type);
        if (mode == 1) {
            LinearOperator<double, MKL>* M = make_prec<double, ConvD>(prec, A);
            st->prec_seconds = std::chrono::duration<double>(clk::now() - p0).count();
            auto t0 = clk::now();
            gmres_baseline<ORTH, MKL, double, double>(*cd, A_type, M, b_type, x_type);   // :101
            st->gmres_seconds = std::chrono::duration<double>(clk::now() - t0).count();
            delete M;
        } else {
            LinearOperator<float, MKL>* M = make_prec<float, ConvD>(prec, A);
            st->prec_seconds = std::chrono::duration<double>(clk::now() - p0).count();
            auto t0 = clk::now();
            gmres_baseline<ORTH, MKL, double, float>(*cd, A_type, M, b_type, x_type);
            st->gmres_seconds = std::chrono::duration<double>(clk::now() - t0).count();
            delete M;
        }
        Vect<double, MKL> x(n), r(n);
        copy(x_type, x);
        copy(b_type, r);
        spmv(-1.0, A, x, 1.0, r);
        st->res_norm = nrm2(r);
        copy(x, x_out);
        axpy(-1.0, true_x, x);
        st->err_norm = nrm2(x);
    } else {
        using ORTH = Orthogonalization::GS<float, Kernel<float, MKL>, MKL>;
        auto p0 = clk::now();
        const SparseMatrix<float, MKL> A_type(A);
        LinearOperator<float, MKL>* M = make_prec<float, ConvF>(prec, A);
        st->prec_seconds = std::chrono::duration<double>(clk::now() - p0).count();
        Vect<float, MKL> x_type(n);
        copy(x_out, x_type);
        Vect<float, MKL> b_type(n);
        copy(b, b_type);
        auto t0 = clk::now();
        gmres_baseline<ORTH, MKL, float, float>(*cf, A_type, M, b_type, x_type);
        st->gmres_seconds = std::chrono::duration<double>(clk::now() - t0).count();
        Vect<double, MKL> x(n), r(n);
        copy(x_type, x);
        copy(b_type, r);
        spmv(-1.0, A, x, 1.0, r);
        st->res_norm = nrm2(r);
        copy(x, x_out);
        axpy(-1.0, true_x, x);
        st->err_norm = nrm2(x);
        delete M;
    }
}

template <class T>
struct CGSR2 {
    template <class A, class B>
    using K = Orthogonalization::CGSR_Kernel<A, B, 2>;
};
template <class A, class B> using CGSR2K = Orthogonalization::CGSR_Kernel<A, B, 2>;

template <class ConvF, class ConvD>
void dispatch_orth(int orth, int mode, int prec, SparseMatrix<double, MKL> A, Vect<double, MKL> b, Vect<double, MKL> x, Vect<double, MKL> xt, ConvF* cf,
                   ConvD* cd, RefStats* st) {
    if (orth == 0) run<Orthogonalization::CGS_Kernel>(mode, prec, A, b, x, xt, cf, cd, st);
    else if (orth == 1) run<Orthogonalization::MGS_Kernel>(mode, prec, A, b, x, xt, cf, cd, st);
    else run<CGSR2K>(mode, prec, A, b, x, xt, cf, cd, st);
}

template <class L>
void export_hist(L* c, RefStats* st, double* hi, int64_t cap_i, double* ho, int64_t cap_o) {
    st->total_iters = (int64_t)c->total_iterations();
    st->total_restarts = (int64_t)c->total_restarts;
    st->n_hist_inner = (int64_t)c->inner.size();
    st->n_hist_outer = (int64_t)c->outer.size() / 4;
    for (int64_t i = 0; i < st->n_hist_inner && i < cap_i; ++i) hi[i] = c->inner[i];
    for (int64_t i = 0; i < (int64_t)c->outer.size() && i < 4 * cap_o; ++i) ho[i] = c->outer[i];
}

}  // namespace

extern "C" {

// the reference's own LoadMatrix<double>() (LoadMatrix.hpp:17-154).  Call with null arrays to get the sizes.
// returns 0, or 1 with the exception text in err
int ref_load_matrix(const char* path, int* n, long* nnz, int* row_map, int* inds, double* vals, char* err, int errlen) {
    try {
        SparseMatrix<double, MKL> A = LoadMatrix<double>(const_cast<char*>(path));
        *n = A.nrows();
        *nnz = (long)A.inds_.extent(0);
        if (row_map) for (int i = 0; i <= *n; ++i) row_map[i] = A.row_map_(i);
        if (inds) for (long k = 0; k < *nnz; ++k) { inds[k] = A.inds_(k); vals[k] = A.vals_(k); }
        return 0;
    } catch (const std::exception& e) {
        if (err && errlen > 0) std::snprintf(err, (size_t)errlen, "%s", e.what());
        return 1;
    }
}

// the reference's own LoadVector<double>(file, col) (LoadMatrix.hpp:156-233); call with out == null for the length
int ref_load_vector(const char* path, int col, long* n, double* out, char* err, int errlen) {
    try {
        Vect<double, MKL> v = LoadVector<double>(const_cast<char*>(path), col);
        *n = (long)v.n();
        if (out) for (long i = 0; i < *n; ++i) out[i] = v.data()[i];
        return 0;
    } catch (const std::exception& e) {
        if (err && errlen > 0) std::snprintf(err, (size_t)errlen, "%s", e.what());
        return 1;
    }
}

int ref_num_threads() { return MKL_Get_Max_Threads(); }

// ---- per-op timings of the reference's OWN MKL backend (BASELINE.json configs[4], the kernel_perf_test sweep) ----------------
// spmv<T,MKL> (kernels_mkl.cpp:326-352) and Orthogonalization::GS<T, Kernel<T,MKL>, MKL>::add_vector (Orthogonalization.hpp:51-60 with
// the kernels of :76-136) - one warm pass, then `trials` timed passes (kernel_perf_test.cpp:170-179 shape), basis and vectors
// filled from mt19937 floats as kernel_perf_test.cpp:11-39 does.
}  // extern "C" (templates below)
namespace {
template <class T>
void fill_random(T* p, size_t n, uint32_t seed) {
    std::mt19937 engine(seed);
    std::uniform_real_distribution<float> dist;
    for (size_t i = 0; i < n; ++i) p[i] = dist(engine);
}
template <class T>
void sweep_spmv(int n, int* rm, int* ind, double* vals64, int trials, double* seconds) {
    const size_t nnz = rm[n];
    using IV = Kokkos::View<int*, Kokkos::HostSpace>;
    using DV = Kokkos::View<double*, Kokkos::HostSpace>;
    SparseMatrix<double, MKL> A64(n, n, IV(rm, n + 1), IV(ind, nnz), DV(vals64, nnz));
    SparseMatrix<T, MKL> A(A64);
    Vect<T, MKL> x(n), y(n);
    fill_random(x.data(), n, 42);
    spmv(T(1), A, x, T(0), y);   // warm
    for (int t = 0; t < trials; ++t) {
        auto t0 = std::chrono::high_resolution_clock::now();
        spmv(T(1), A, x, T(0), y);
        seconds[t] = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count();
    }
}
template <class T, template <class, class> class Kernel>
void sweep_add_vector(size_t n, size_t m, const int* ks, int nk, int trials, double* seconds) {
    Orthogonalization::GS<T, Kernel<T, MKL>, MKL> gs(n, m);
    T* v = gs.v.data();
    // an orthonormal-ish basis is not needed for timing; scale the random columns so that norms stay O(1)
    fill_random(v, n * (m + 1), 45);
    const T sc = T(1) / std::sqrt((T)n);
    for (size_t i = 0; i < n * (m + 1); ++i) v[i] *= sc;
    MultiVect<T, MKL> h(m + 1, m);
    Vect<T, MKL> w(n), w0(n);
    fill_random(w0.data(), n, 43);
    for (int q = 0; q < nk; ++q) {
        const size_t k = (size_t)ks[q];
        copy(w0, w);
        gs.add_vector(k, w, h);   // warm
        for (int t = 0; t < trials; ++t) {
            copy(w0, w);
            auto t0 = std::chrono::high_resolution_clock::now();
            gs.add_vector(k, w, h);
            seconds[(size_t)q * trials + t] = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count();
        }
    }
}
template <class A, class B> using CGSR2K_sweep = Orthogonalization::CGSR_Kernel<A, B, 2>;
}  // namespace
extern "C" {

int ref_sweep_spmv(int n, int* rm, int* ind, double* vals64, int is_float, int trials, double* seconds) {
    if (is_float) sweep_spmv<float>(n, rm, ind, vals64, trials, seconds); else sweep_spmv<double>(n, rm, ind, vals64, trials, seconds);
    return 0;
}
// orth: 0 CGS, 1 MGS, 2 CGSR<2>; ks[nk]: the values of k (basis width k + 1) to time; seconds[nk * trials]
int ref_sweep_add_vector(int64_t n, int m, int orth, int is_float, const int* ks, int nk, int trials, double* seconds) {
    if (is_float) {
        if (orth == 0) sweep_add_vector<float, Orthogonalization::CGS_Kernel>(n, m, ks, nk, trials, seconds);
        else if (orth == 1) sweep_add_vector<float, Orthogonalization::MGS_Kernel>(n, m, ks, nk, trials, seconds);
        else sweep_add_vector<float, CGSR2K_sweep>(n, m, ks, nk, trials, seconds);
    } else {
        if (orth == 0) sweep_add_vector<double, Orthogonalization::CGS_Kernel>(n, m, ks, nk, trials, seconds);
        else if (orth == 1) sweep_add_vector<double, Orthogonalization::MGS_Kernel>(n, m, ks, nk, trials, seconds);
        else sweep_add_vector<double, CGSR2K_sweep>(n, m, ks, nk, trials, seconds);
    }
    return 0;
}

// same argument meaning as orc_gmres (oracle/oracle.cpp); true_x may be null (then err_norm = ||x||)
int ref_gmres(int mode, int orth, int conv_kind, int prec, int64_t rlen, double tol, double rtol, int64_t max_restarts, int n, int* row_map, int* inds,
              double* vals64, double* b_in, double* x_io, double* true_x_in, RefStats* st, double* hist_inner, int64_t cap_inner, double* hist_outer,
              int64_t cap_outer) {
    std::memset(st, 0, sizeof(RefStats));
    const size_t nnz = row_map[n];
    using IV = Kokkos::View<int*, Kokkos::HostSpace>;
    using DV = Kokkos::View<double*, Kokkos::HostSpace>;
    SparseMatrix<double, MKL> A(n, n, IV(row_map, n + 1), IV(inds, nnz), DV(vals64, nnz));
    Vect<double, MKL> b = wrap(b_in, n), x = wrap(x_io, n);
    Vect<double, MKL> xt(n);
    if (true_x_in) copy(wrap(true_x_in, n), xt);

    // the reference reports through std::cout (gmres.cpp:186-190,230-239); capture it to recover status / k / i
    std::ostringstream cap;
    std::streambuf* old = std::cout.rdbuf(cap.rdbuf());
    // alloc_convergence, gmres_perf_test.cpp:185-196
    if (conv_kind == 0 || rtol == 0) {
        auto* cf = new Logged<Convergence<float, MKL>>(tol, (size_t)rlen, (size_t)max_restarts);
        auto* cd = new Logged<Convergence<double, MKL>>(tol, (size_t)rlen, (size_t)max_restarts);
        dispatch_orth(orth, mode, prec, A, b, x, xt, cf, cd, st);
        if (mode == 0 || mode == 3) export_hist(cf, st, hist_inner, cap_inner, hist_outer, cap_outer); else export_hist(cd, st, hist_inner, cap_inner, hist_outer, cap_outer);
        delete cf; delete cd;
    } else if (conv_kind == 1) {
        auto* cf = new Logged<RelPrecRes_Convergence<float, MKL>>(tol, rtol, (size_t)rlen, (size_t)max_restarts);
        auto* cd = new Logged<RelPrecRes_Convergence<double, MKL>>(tol, rtol, (size_t)rlen, (size_t)max_restarts);
        dispatch_orth(orth, mode, prec, A, b, x, xt, cf, cd, st);
        if (mode == 0 || mode == 3) export_hist(cf, st, hist_inner, cap_inner, hist_outer, cap_outer); else export_hist(cd, st, hist_inner, cap_inner, hist_outer, cap_outer);
        delete cf; delete cd;
    } else if (conv_kind == 2) {
        auto* cf = new Logged<RepeatIteration_Convergence<float, MKL>>(tol, rtol, (size_t)rlen, (size_t)max_restarts);
        auto* cd = new Logged<RepeatIteration_Convergence<double, MKL>>(tol, rtol, (size_t)rlen, (size_t)max_restarts);
        dispatch_orth(orth, mode, prec, A, b, x, xt, cf, cd, st);
        if (mode == 0 || mode == 3) export_hist(cf, st, hist_inner, cap_inner, hist_outer, cap_outer); else export_hist(cd, st, hist_inner, cap_inner, hist_outer, cap_outer);
        delete cf; delete cd;
    } else {
        auto* cf = new Logged<LostOrthogonality_Convergence<float, MKL>>(tol, rtol, (size_t)rlen, (size_t)max_restarts);
        auto* cd = new Logged<LostOrthogonality_Convergence<double, MKL>>(tol, rtol, (size_t)rlen, (size_t)max_restarts);
        dispatch_orth(orth, mode, prec, A, b, x, xt, cf, cd, st);
        if (mode == 0 || mode == 3) export_hist(cf, st, hist_inner, cap_inner, hist_outer, cap_outer); else export_hist(cd, st, hist_inner, cap_inner, hist_outer, cap_outer);
        delete cf; delete cd;
    }
    std::cout.rdbuf(old);
    const std::string out = cap.str();
    st->status = out.find("Found solution") != std::string::npos ? 1 : (out.find("Aborting") != std::string::npos ? 3 : 0);
    const size_t p = out.find("rel prec res norm = ");
    if (p != std::string::npos) {
        std::istringstream is(out.substr(p + 20));
        std::string w1, w2, w3, w4;
        long kk = 0, ii = 0;
        is >> st->rel_prec_res >> w1 >> w2 >> w3 >> kk >> w4 >> w1 >> w2 >> ii;   // "<v> when k = <k> and i = <i>"
        st->outer_i = ii;
    }
    return 0;
}

}  // extern "C"
