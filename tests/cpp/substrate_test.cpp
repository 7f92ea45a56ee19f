// substrate_test.cpp — drives include/b200/substrate.hpp (the Kokkos-free Scalar / Vect / MultiVect / SparseMatrix handles and the
// operator surface) the way the reference's own code uses its types.hpp handles:
//   Orthogonalization.hpp:38,48,58      Vect(v, ALL, col) columns of the basis
//   Orthogonalization.hpp:63,69         MultiVect(v, ALL, pair(0, k)) column blocks for update_x
//   Orthogonalization.hpp:83-87,121-133 v_prevCols.transpose_matrix() gemv pairs, Vect(h, pair(0, k+1), k), weights sub-range
//   gmres.cpp:219-222                   rot(h(range, k), cos(range), sin(range)); rotg(h(k,k), h(k+1,k), cos(k), sin(k)); rot(s(k), s(k+1), ...)
//   gmres.cpp:276-303                   y(s, pair(0, k)); h_temp(h, pair(0, k), pair(0, k)); trsv("Upper", h_temp, y); update_x
// and checks every step against plain host loops in double.  Built with g++ only (no nvcc, no Kokkos) and linked to
// libmpgmres_b200.so; run by tests/test_substrate_gpu.py on the GPU box.  Exit code 0 = all checks passed.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define B200_SUBSTRATE_THROW
#include "b200/substrate.hpp"

using namespace b200;

static int g_fail = 0;
#define CHECK(cond)                                                                  \
    do {                                                                             \
        if (!(cond)) { std::printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond); ++g_fail; } \
    } while (0)

// ---- 2-D 5-point Laplacian, LoadMatrix-canonical CSR ---------------------------------------------------------------------
static void lap2d(int N, std::vector<int>& rm, std::vector<int>& ind, std::vector<double>& val) {
    rm.assign(1, 0);
    for (int y = 0; y < N; ++y)
        for (int x = 0; x < N; ++x) {
            const int i = x + N * y;
            if (y > 0) { ind.push_back(i - N); val.push_back(-1); }
            if (x > 0) { ind.push_back(i - 1); val.push_back(-1); }
            ind.push_back(i); val.push_back(4);
            if (x < N - 1) { ind.push_back(i + 1); val.push_back(-1); }
            if (y < N - 1) { ind.push_back(i + N); val.push_back(-1); }
            rm.push_back((int)ind.size());
        }
}

// ---- handle semantics (types.hpp:15-228) ----------------------------------------------------------------------------------
static void test_handles() {
    MultiVect<double> M(5, 4);   // zero-filled like a Kokkos::View
    for (double v : M.download()) CHECK(v == 0.0);
    std::vector<double> h(20);
    for (int j = 0; j < 4; ++j)
        for (int i = 0; i < 5; ++i) h[i + 5 * j] = 10 * i + j;
    M.upload(h.data());
    CHECK(M.nrows() == 5 && M.ncols() == 4 && M.stride() == 5 && M.n() == 5 && !M.transposed());
    // element access honours the flag; the flag flip moves no data
    CHECK(M(3, 2).access() == 32.0);
    MultiVect<double> Mt = M.transpose_matrix();
    CHECK(Mt.data() == M.data() && Mt.transposed() && Mt.nrows() == 4 && Mt.ncols() == 5 && Mt.nrows_base() == 5 && Mt.ncols_base() == 4);
    CHECK(Mt(2, 3).access() == 32.0);
    // sub-block: data() = block origin, stride() = parent column stride, base extents = the block's
    MultiVect<double> B(M, range(1, 4), range(2, 4));
    CHECK(B.data() == M.data() + 1 + 2 * 5 && B.stride() == 5 && B.nrows_base() == 3 && B.ncols_base() == 2 && B.nrows() == 3 && B.ncols() == 2);
    CHECK(B(0, 0).access() == 12.0 && B(2, 1).access() == 33.0);
    // sub-block of a transposed handle addresses LOGICAL rows/cols (types.hpp:135-167)
    MultiVect<double> Bt(Mt, range(2, 4), range(1, 4));   // logical rows 2..3 (base cols), logical cols 1..3 (base rows)
    CHECK(Bt.transposed() && Bt.nrows() == 2 && Bt.ncols() == 3 && Bt.nrows_base() == 3 && Bt.ncols_base() == 2 && Bt.data() == B.data());
    CHECK(Bt(1, 2).access() == 33.0);
    MultiVect<double> Cc(M, ALL, range(1, 3));
    CHECK(Cc.nrows() == 5 && Cc.ncols() == 2 && Cc.data() == M.data() + 5);
    MultiVect<double> Rr(M, range(2, 5), ALL);
    CHECK(Rr.nrows() == 3 && Rr.ncols() == 4 && Rr.data() == M.data() + 2 && Rr.stride() == 5);
    MultiVect<double> Rt(Mt, range(1, 3), ALL);   // logical rows of the transposed matrix = base columns
    CHECK(Rt.nrows() == 2 && Rt.ncols() == 5 && Rt.data() == M.data() + 5 && Rt.nrows_base() == 5 && Rt.ncols_base() == 2);
    // columns / column pieces as vectors; Vect(MultiVect, rows, col) ignores the flag (types.hpp:79-81)
    Vect<double> c2(M, ALL, 2);
    CHECK(c2.n() == 5 && c2.data() == M.data() + 10 && c2.access(4) == 42.0);
    Vect<double> piece = M(range(1, 3), 3);
    CHECK(piece.n() == 2 && piece.access(0) == 13.0 && piece.access(1) == 23.0);
    Vect<double> piece_t = Mt(range(1, 3), 3);
    CHECK(piece_t.data() == piece.data());
    // sub-range of a vector, scalar views, aliasing
    Vect<double> sub = c2(range(1, 4));
    CHECK(sub.n() == 3 && sub.data() == c2.data() + 1);
    Scalar<double> e = sub(2);
    CHECK(e.data() == M.data() + 13 && e.access() == 32.0);
    fill(7.0, sub);
    CHECK(M(1, 2).access() == 7.0 && M(3, 2).access() == 7.0 && M(0, 2).access() == 2.0 && M(4, 2).access() == 42.0 && e.access() == 7.0);
    // a sub-view keeps the allocation alive after the parent handle is gone; copies are shallow
    Vect<double> keep;
    {
        Vect<double> parent(100);
        fill(3.0, parent);
        Vect<double> copy_of_parent = parent;
        CHECK(copy_of_parent.data() == parent.data());
        keep = parent(range(40, 60));
    }
    CHECK(keep.n() == 20 && keep.access(19) == 3.0);
    // Scalar(val) and device-scalar forms
    Scalar<float> a(3.0f), b(4.0f), c, s;
    rotg(a, b, c, s);   // r = 5 (sign of b since |b| > |a|), then b := 0 (kernels_cuda.cpp:394-420)
    CHECK(std::fabs(a.access() - 5.0f) < 1e-6f && b.access() == 0.0f && std::fabs(c.access() - 0.6f) < 1e-6f && std::fabs(s.access() - 0.8f) < 1e-6f);
    Scalar<float> two(2.0f), out;
    scal(1.5f, two, out);
    CHECK(out.access() == 3.0f);
    scal(two, out, out);
    CHECK(out.access() == 6.0f);
}

// ---- one GMRES(m) cycle written with the reference's own sub-view idioms, unfused kernels -----------------------------------
// ORTH: 0 CGS (Orthogonalization.hpp:82-88), 1 MGS (:98-106), 2 CGSR<2> (:120-135)
template <class T>
static void orthogonalize_like_reference(int orth, MultiVect<T> v, size_t k, Vect<T> w, MultiVect<T> h, Vect<T> weights) {
    if (orth == 1) {
        for (size_t j = 0; j < k + 1; ++j) {
            Vect<T> v_col(v, ALL, j);
            dot(w, v_col, h(j, k));
            naxpy(h(j, k), v_col, w);
        }
        return;
    }
    MultiVect<T> v_prevCols(v, ALL, range(0, k + 1));
    Vect<T> h_col(h, range(0, k + 1), k);
    gemv(T(1), v_prevCols.transpose_matrix(), w, T(0), h_col);
    gemv(T(-1), v_prevCols, h_col, T(1), w);
    if (orth == 2) {
        Vect<T> weights_view(weights, range(0, k + 1));
        gemv(T(1), v_prevCols.transpose_matrix(), w, T(0), weights_view);
        gemv(T(-1), v_prevCols, weights_view, T(1), w);
        axpy(T(1), weights_view, h_col);
    }
}

template <class T>
static void test_cycle(int orth) {
    const int N = 24, n = N * N;
    const size_t m = 30;
    std::vector<int> rm, ind;
    std::vector<double> val;
    lap2d(N, rm, ind, val);
    SparseMatrix<double> A64(n, n, rm, ind, val);
    SparseMatrix<T> A(A64);   // precision cast shares the structure (types_cuda.hpp:82-101)
    CHECK(A.row_map_data() == A64.row_map_data() && A.inds_data() == A64.inds_data() && A.plan() == A64.plan());
    std::vector<T> bh(n);
    for (int i = 0; i < n; ++i) bh[i] = (T)(1.0 + 0.37 * std::sin(0.11 * i));
    Vect<T> b(bh);

    // (a) the reference's loop, unfused kernels through sub-views
    MultiVect<T> v(n, m + 1), h(m + 1, m);
    Vect<T> w(n), cosv(m + 1), sinv(m + 1), s(m + 1), weights(m);
    copy(b, w);
    const T beta = nrm2(w);
    {
        Vect<T> v_col(v, ALL, 0);
        scal(1 / beta, w, v_col);
    }
    fill(T(0), s);
    fill(beta, s(0));
    std::vector<double> resid_a;
    for (size_t k = 0; k < m; ++k) {
        {
            Vect<T> v_col(v, ALL, k);
            spmv(T(1), A, v_col, T(0), w);
        }
        orthogonalize_like_reference<T>(orth, v, k, w, h, weights);
        nrm2(w, h(k + 1, k));
        const T h_final = h(k + 1, k).access();
        Vect<T> v_next(v, ALL, k + 1);
        scal(1 / h_final, w, v_next);
        const range r(0, k);
        rot(h(r, k), cosv(r), sinv(r));
        rotg(h(k, k), h(k + 1, k), cosv(k), sinv(k));
        rot(s(k), s(k + 1), cosv(k), sinv(k));
        fence();
        resid_a.push_back(std::fabs((double)s.access(k + 1)));
    }
    // basis orthonormality through gemv on column blocks: G = V_k^T V_k (host check)
    {
        std::vector<T> vh = v.download();
        double worst = 0;
        for (size_t i = 0; i < 6; ++i)
            for (size_t j = 0; j < 6; ++j) {
                double d = 0;
                for (int r = 0; r < n; ++r) d += (double)vh[r + i * n] * (double)vh[r + j * n];
                worst = std::fmax(worst, std::fabs(d - (i == j ? 1.0 : 0.0)));
            }
        CHECK(worst < (sizeof(T) == 4 ? 2e-5 : 1e-12));
    }
    CHECK(resid_a.back() < resid_a.front());
    for (size_t k = 1; k < m; ++k) CHECK(resid_a[k] <= resid_a[k - 1] * (1 + 1e-6));

    // solution_update (gmres.cpp:291-303) through sub-views: y = triu(H[0:k,0:k])^-1 s[0:k] in place, x += V[:,0:k] y
    std::vector<T> Hh = h.download(), sh = s.download();
    Vect<T> y(s, range(0, m));
    MultiVect<T> h_temp(h, range(0, m), range(0, m));
    CHECK(h_temp.stride() == m + 1 && h_temp.nrows_base() == m);
    trsv("Upper", h_temp, y);
    {   // host back-substitution in double on the same (rotated) H and s
        std::vector<double> yy(m);
        for (size_t i = m; i-- > 0;) {
            double t = sh[i];
            for (size_t j = i + 1; j < m; ++j) t -= (double)Hh[i + j * (m + 1)] * yy[j];
            yy[i] = t / (double)Hh[i + i * (m + 1)];
        }
        std::vector<T> yd = y.download();
        double num = 0, den = 0;
        for (size_t i = 0; i < m; ++i) { num += (yd[i] - yy[i]) * (yd[i] - yy[i]); den += yy[i] * yy[i]; }
        CHECK(std::sqrt(num / den) < (sizeof(T) == 4 ? 1e-3 : 1e-10));
    }
    Vect<T> x(n);
    {
        MultiVect<T> v_cols(v, ALL, range(0, m));
        gemv(T(1), v_cols, y, T(1), x);
    }
    // true residual ||b - A x|| equals the Givens estimate |s(m)| (one cycle from x0 = 0)
    Vect<T> r(n);
    copy(b, r);
    spmv(T(-1), A, x, T(1), r);
    const double true_res = (double)nrm2(r);
    CHECK(std::fabs(true_res - resid_a.back()) <= (sizeof(T) == 4 ? 5e-3 : 1e-8) * resid_a.front() + 1e-3 * resid_a.back());

    // (b) the same cycle with the fused GS::add_vector of the substrate (what the drop-in uses): same history
    typedef Orthogonalization::GS<T, Orthogonalization::CGS> GS0;
    typedef Orthogonalization::GS<T, Orthogonalization::MGS> GS1;
    typedef Orthogonalization::GS<T, Orthogonalization::CGSR2> GS2;
    MultiVect<T> h2(m + 1, m);
    Vect<T> w2(n), cos2(m + 1), sin2(m + 1), s2(m + 1);
    GS0 g0(n, m); GS1 g1(n, m); GS2 g2(n, m);
    copy(b, w2);
    const T beta2 = orth == 0 ? g0.first_vector(w2) : (orth == 1 ? g1.first_vector(w2) : g2.first_vector(w2));
    CHECK(beta2 == beta);
    fill(T(0), s2);
    fill(beta2, s2(0));
    for (size_t k = 0; k < m; ++k) {
        Vect<T> v_col = orth == 0 ? g0.previous_krylov_vector(k) : (orth == 1 ? g1.previous_krylov_vector(k) : g2.previous_krylov_vector(k));
        spmv(T(1), A, v_col, T(0), w2);
        if (orth == 0) g0.add_vector(k, w2, h2); else if (orth == 1) g1.add_vector(k, w2, h2); else g2.add_vector(k, w2, h2);
        const range r(0, k);
        rot(h2(r, k), cos2(r), sin2(r));
        rotg(h2(k, k), h2(k + 1, k), cos2(k), sin2(k));
        rot(s2(k), s2(k + 1), cos2(k), sin2(k));
        const double rb = std::fabs((double)s2.access(k + 1));
        CHECK(std::fabs(rb - resid_a[k]) <= (sizeof(T) == 4 ? 2e-3 : 1e-9) * resid_a[k] + 1e-7 * resid_a[0]);
    }
    // update through the mixed-precision overload (Orthogonalization.hpp:67-73): x64 += (double)(V y)
    Vect<T> y2(s2, range(0, m));
    MultiVect<T> h2_temp(h2, range(0, m), range(0, m));
    trsv("Upper", h2_temp, y2);
    Vect<double> x64(n), x_temp(n);
    Vect<T> x_inc(n);
    if (orth == 0) g0.update_x(m, y2, x64, x_inc, x_temp); else if (orth == 1) g1.update_x(m, y2, x64, x_inc, x_temp); else g2.update_x(m, y2, x64, x_inc, x_temp);
    std::vector<double> xa = x64.download();
    std::vector<T> xb = x.download();
    double num = 0, den = 0;
    for (int i = 0; i < n; ++i) { num += (xa[i] - xb[i]) * (xa[i] - xb[i]); den += (double)xb[i] * xb[i]; }
    CHECK(std::sqrt(num / den) < (sizeof(T) == 4 ? 2e-3 : 1e-9));
}

// ---- the drivers (gmres.hpp:15-32) and Jacobi -----------------------------------------------------------------------------
static void test_drivers() {
    const int N = 40, n = N * N;
    std::vector<int> rm, ind;
    std::vector<double> val;
    lap2d(N, rm, ind, val);
    SparseMatrix<double> A(n, n, rm, ind, val);
    SparseMatrix<float> A32(A);
    std::vector<double> xt(n);
    for (int i = 0; i < n; ++i) xt[i] = 0.5 + 0.25 * std::cos(0.05 * i);
    Vect<double> x_true(xt), b(n), x(n);
    spmv(1.0, A, x_true, 0.0, b);
    SolveOptions o;
    o.restart_length = 40; o.tol = 1e-9;
    SolveResult r = gmres_singleUpdate(o, A, A32, b, x);
    CHECK(r.stats.status == 1 && r.stats.total_iters > 0 && r.stats.total_iters % 40 == 0 && (int64_t)r.hist_inner.size() == r.stats.total_iters);
    Vect<double> res(n);
    copy(b, res);
    spmv(-1.0, A, x, 1.0, res);
    CHECK(nrm2(res) / (r.stats.b_norm + r.stats.A_norm * nrm2(x)) <= 1e-9 * 1.01);   // the reference's stopping rule, IterUtil.hpp:42-51
    axpy(-1.0, x_true, x);
    CHECK(nrm2(x) < 1e-4 * nrm2(x_true));
    Vect<double> xb(n);
    o.prec = MPG_PREC_JACOBI;
    SolveResult rb = gmres_baseline(o, A, b, xb);
    CHECK(rb.stats.status == 1);
    Jacobi<double> J(A);
    CHECK(J.diag().access(5) == 0.25);
    // error path: a bad restart length surfaces as an exception (B200_SUBSTRATE_THROW), never as a crash
    o.restart_length = 0;
    bool threw = false;
    try { gmres_baseline(o, A, b, xb); } catch (const std::runtime_error&) { threw = true; }
    CHECK(threw);
}

int main() {
    std::printf("%s\n", mpg_version());
    test_handles();
    for (int orth = 0; orth < 3; ++orth) {
        test_cycle<float>(orth);
        test_cycle<double>(orth);
    }
    test_drivers();
    if (g_fail) { std::printf("substrate: %d check(s) FAILED\n", g_fail); return 1; }
    std::printf("substrate ok\n");
    return 0;
}
