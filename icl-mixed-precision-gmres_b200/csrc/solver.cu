// solver.cu — restarted GMRES drivers behind the reference's gmres.hpp signatures:
//   gmres_singleUpdate<Orth,Device>            gmres.hpp:26-32, gmres.cpp:135-245   (GMRES-IR: fp32 inner, fp64 outer)
//   gmres_baseline<Orth,Device,Type,PrecType>  gmres.hpp:15-20, gmres.cpp:24-133    (uniform precision)
//   solution_update (both overloads)           gmres.hpp:44-57, gmres.cpp:276-303
// with GS<...> / CGS / MGS / CGSR<2> (Orthogonalization.hpp) and the Convergence family (IterUtil.hpp)
// as run-time options instead of template parameters.
//
// What is different from the reference's driver (behaviour-preserving):
//   * nothing is read back per inner iteration: h(k+1,k), 1/h(k+1,k), the Givens rotations and |s(k+1)| stay
//     on the device (the reference blocks twice per iteration, Orthogonalization.hpp:56 and gmres.cpp:226).
//     With the base Convergence class — whose check() looks at the iteration count only, IterUtil.hpp:57-65 —
//     a whole restart cycle is enqueued without a host synchronisation; the residual-driven restart
//     policies read one 8-byte value per iteration, as the reference does;
//   * outer residual, cast and the per-restart norms are 1 SpMV-shaped kernel + 3 reductions + ONE readback;
//   * CGS2 makes 3 passes over the basis instead of 4 (ortho.cu);  V(:,k+1) = w/h is one pass, not copy+scal;
//   * the workspace is cached in the context (the reference allocates and zero-fills 2 bases per call,
//     gmres.cpp:147,156 — the second one is never used).
#include <algorithm>
#include <thread>
#include <vector>

#include "common.cuh"

using namespace mpg;

namespace mpg {
int dot_dev(mpg_ctx*, int64_t, const float*, const float*, float*);
int dot_dev(mpg_ctx*, int64_t, const double*, const double*, double*);
int nrm2_dev(mpg_ctx*, int64_t, const float*, float*);
int nrm2_dev(mpg_ctx*, int64_t, const double*, double*);
int scal_host(mpg_ctx*, int64_t, float, const float*, float*);
int scal_host(mpg_ctx*, int64_t, double, const double*, double*);
int cast_copy(mpg_ctx*, int64_t, const double*, float*);
int cast_copy(mpg_ctx*, int64_t, const float*, double*);
int cast_copy(mpg_ctx*, int64_t, const float*, float*);
int cast_copy(mpg_ctx*, int64_t, const double*, double*);
int fill_host(mpg_ctx*, int64_t, float, float*);
int fill_host(mpg_ctx*, int64_t, double, double*);
int gdmv_host(mpg_ctx*, int64_t, float, const float*, const float*, float, float*);
int gdmv_host(mpg_ctx*, int64_t, double, const double*, const double*, double, double*);
template <class T> int givens_step(mpg_ctx*, int64_t, T*, int64_t, T*, T*, T*, double*, double*);
template <class T> int trsv(mpg_ctx*, int, int, int64_t, const T*, int64_t, T*);
template <class T> int spmv(mpg_ctx*, const mpg_csr*, const T*, T, const T*, T, const T*, T*, float*, const T* rowscale = nullptr, int part = SPMV_ALL);
template <class T> int gemvn(mpg_ctx*, int64_t, int, const T*, int64_t, T, const T*, T, T*, bool, T*, T*, double*);
template <class T> int gemvt(mpg_ctx*, int64_t, int, const T*, int64_t, T, const T*, T, T*);
template <class T> int add_vector(mpg_ctx*, int, int64_t, int64_t, T*, int64_t, T*, T*, T*, bool);
template <class T> int arnoldi_tail(mpg_ctx*, int64_t, const T*, const T*, T*, int64_t, T*, int64_t, T*, T*, T*, double*, double*, const PushArgs* = nullptr);
template <class T> int pack_create(mpg_ctx*, const mpg_csr*, const T*, mpg_packed**);
template <class T> int pack_update(mpg_ctx*, mpg_packed*, const T*);
void pack_free(mpg_packed*);
bool pack_matches(const mpg_packed*, const mpg_csr*, int tsize);
int pack_boundary_slices(const mpg_packed*);
template <class T> int spmv_packed(mpg_ctx*, const mpg_packed*, T, const T*, T, const T*, T*, float*, const T*, int, const HaloWait* = nullptr, const T* xadd = nullptr, const PushArgs* push = nullptr);
template <class T> int ilu_jacobi_apply_t(mpg_ctx*, mpg_ilu_jacobi*, T*);
template <class T> int halo_exchange(mpg_ctx*, T*);
template <class T> int halo_begin(mpg_ctx*, T*);
template <class T> int halo_finish(mpg_ctx*, T*);
int64_t dist_halo(mpg_ctx*);
int dist_exchange_basis(mpg_ctx*, void* V, int64_t ldv, int tsize);
bool dist_basis_ready(mpg_ctx*);
template <class T> int halo_direct_args(mpg_ctx*, int64_t col, PushArgs*);
template <class T> int halo_push_direct(mpg_ctx*, const T* x, int64_t col);
int halo_wait_args(mpg_ctx*, HaloWait*);
int halo_wait_only(mpg_ctx*);
int64_t dist_nglobal(mpg_ctx*);
int dist_world(mpg_ctx*);
}  // namespace mpg

namespace {

// ---- restart / stop policy: IterUtil.hpp restated for the host side of the driver --------------------------
enum Action { NEXT = 0, CONVERGED = 1, RESTART = 2, ABORTED = 3 };  // iteration_action, IterUtil.hpp:10-15

struct Policy {
    int kind;
    double tol, rtol;
    int64_t rlen, max_restarts;
    int64_t total_iters = 0, total_restarts = 0;
    double restart_tol;
    int64_t second_len = 0;
    bool first_iteration = true;
    double loss_sq = 0;
    double rows_per_rank = 0, nnz_per_rank = 0;   // rank-invariant size of one rank's share (look-ahead depth), set by size_estimate()

    Policy(const mpg_gmres_params& p)
        : kind(p.conv), tol(p.tol), rtol(p.restart_tol), rlen(p.restart_length), max_restarts(p.max_restarts), restart_tol(p.restart_tol) {}

    bool needs_residual() const { return kind == MPG_CONV_RELPRECRES || kind == MPG_CONV_REPEAT; }

    Action check_initial(double res, double normalization, double pres, double pb) {
        if (kind == MPG_CONV_RELPRECRES) restart_tol = pres / pb * rtol;                    // IterUtil.hpp:150-153
        if (kind == MPG_CONV_REPEAT && first_iteration) restart_tol = pres / pb * rtol;     // :99-104
        if (kind == MPG_CONV_ORTHLOSS) loss_sq = 0;                                         // :195-198
        total_restarts++;                                                                   // :42-51
        if (total_restarts > max_restarts) return ABORTED;
        if (res / normalization > tol) return NEXT;
        return CONVERGED;
    }
    Action base_check(int64_t k) {  // :57-65
        total_iters++;
        return (rlen <= k) ? RESTART : NEXT;
    }
    // `loss_inc` is only consulted for ORTHLOSS and must be evaluated lazily by the caller (it costs a gemv)
    template <class LossFn>
    Action check(int64_t k, double res, double bnorm, LossFn loss_inc) {
        const Action a = base_check(k);
        switch (kind) {
            case MPG_CONV_RELPRECRES:  // :155-165
                if (a != NEXT) return a;
                return (res / bnorm <= restart_tol) ? RESTART : NEXT;
            case MPG_CONV_REPEAT:      // :106-133
                if (first_iteration) {
                    if (a != NEXT) { first_iteration = false; second_len = k; return a; }
                    if (res / bnorm <= restart_tol) { first_iteration = false; second_len = k; return RESTART; }
                    return NEXT;
                }
                if (a != NEXT) return a;
                return (second_len <= k) ? RESTART : NEXT;
            case MPG_CONV_ORTHLOSS:    // :200-223
                if (a != NEXT) return a;
                loss_sq += loss_inc(k);
                return (loss_sq >= rtol * rtol) ? RESTART : NEXT;
            default:
                return a;
        }
    }
};

// ---- cached workspace ---------------------------------------------------------------------------------------
struct Workspace {
    int64_t n = 0, m = 0, ldv = 0;
    int tsize = 0;
    void* V = nullptr;        // ldv x (m+1) of T        Orthogonalization.hpp:27-30
    void* w = nullptr;        // n (padded) of T          gmres.cpp:155
    void* h = nullptr;        // (m+1) x m of T           gmres.cpp:157
    void* cs = nullptr;       // m+1                      gmres.cpp:150
    void* sn = nullptr;       // m+1                      gmres.cpp:151
    void* s = nullptr;        // m+1                      gmres.cpp:152
    void* scratch = nullptr;  // m+8: CGSR weights + 1/norm
    double* hist = nullptr;   // m+1 device |s(k+1)|
    double* hist_host = nullptr;  // pinned
    void* S = nullptr;        // (m+1)^2 of T, ORTHLOSS   IterUtil.hpp:178,185
    void* u = nullptr;        // m+1 of T
    float* tmp32 = nullptr;   // n floats: single-prec preconditioner bridge (typesafe_apply, gmres.cpp:12-17)
    int64_t halo = 0;         // multi-GPU: halo slots appended to every SpMV input (tail of each basis column)
    void* xext = nullptr;     // multi-GPU: [x_local ; halo] copy of the iterate for the outer residual
    // packed copies of the matrix (sell.cu): slot 0 = the operator of the inner iterations, slot 1 = the fp64 operator of the
    // mixed-precision outer residual.  Values are refreshed at every solve unless the caller set values_static.
    mpg_packed* packed[2] = {nullptr, nullptr};
    const void* packed_src[2] = {nullptr, nullptr};   // value array the slot was last packed from
};

// replicated (not row-distributed) data: reductions over it must not be all-reduced
struct LocalScope {
    mpg_ctx* ctx;
    mpg_dist* saved;
    explicit LocalScope(mpg_ctx* c) : ctx(c), saved(c->dist) { c->dist = nullptr; }
    ~LocalScope() { ctx->dist = saved; }
};

void ws_release(void* p) {
    Workspace* ws = static_cast<Workspace*>(p);
    if (!ws) return;
    cudaFree(ws->V); cudaFree(ws->w); cudaFree(ws->h); cudaFree(ws->cs); cudaFree(ws->sn); cudaFree(ws->s);
    cudaFree(ws->scratch); cudaFree(ws->hist); cudaFreeHost(ws->hist_host); cudaFree(ws->S); cudaFree(ws->u); cudaFree(ws->tmp32);
    cudaFree(ws->xext);
    pack_free(ws->packed[0]);
    pack_free(ws->packed[1]);
    delete ws;
}

int get_workspace(mpg_ctx* ctx, int64_t n, int64_t m, int tsize, bool need_S, bool need_tmp32, Workspace** out) {
    Workspace* ws = static_cast<Workspace*>(ctx->ws);
    const int64_t halo = dist_halo(ctx);
    if (ws && (ws->n != n || ws->m != m || ws->tsize != tsize || ws->halo != halo)) {
        MPG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        ws_release(ws);
        ctx->ws = ws = nullptr;
    }
    if (!ws) {
        ws = new Workspace();
        ws->n = n; ws->m = m; ws->tsize = tsize; ws->halo = halo;
        ws->ldv = (n + halo + 31) & ~int64_t(31);  // 128-byte aligned columns; rows [n, n+halo) of a column hold its SpMV halo
        ctx->ws = ws;
        ctx->ws_free = ws_release;
        const size_t ts = (size_t)tsize;
        MPG_CUDA(ctx, cudaMalloc(&ws->V, ts * (size_t)ws->ldv * (size_t)(m + 1)));
        MPG_CUDA(ctx, cudaMalloc(&ws->w, ts * (size_t)(ws->ldv + 32)));
        MPG_CUDA(ctx, cudaMalloc(&ws->h, ts * (size_t)(m + 1) * (size_t)std::max<int64_t>(m, 1)));
        MPG_CUDA(ctx, cudaMalloc(&ws->cs, ts * (size_t)(m + 1)));
        MPG_CUDA(ctx, cudaMalloc(&ws->sn, ts * (size_t)(m + 1)));
        MPG_CUDA(ctx, cudaMalloc(&ws->s, ts * (size_t)(m + 1)));
        MPG_CUDA(ctx, cudaMalloc(&ws->scratch, ts * (size_t)(m + 8)));
        MPG_CUDA(ctx, cudaMalloc(&ws->hist, sizeof(double) * (size_t)(m + 1)));
        MPG_CUDA(ctx, cudaMallocHost(&ws->hist_host, sizeof(double) * (size_t)(m + 1)));
        // Kokkos::View zero-fills on construction; ORTHLOSS even reads a column before it is
        // written (IterUtil.hpp:207), so the basis must start as zeros
        MPG_CUDA(ctx, cudaMemsetAsync(ws->V, 0, ts * (size_t)ws->ldv * (size_t)(m + 1), ctx->stream));
        MPG_CUDA(ctx, cudaMemsetAsync(ws->w, 0, ts * (size_t)(ws->ldv + 32), ctx->stream));
        MPG_CUDA(ctx, cudaMemsetAsync(ws->h, 0, ts * (size_t)(m + 1) * (size_t)std::max<int64_t>(m, 1), ctx->stream));
        if (halo > 0) MPG_CUDA(ctx, cudaMalloc(&ws->xext, sizeof(double) * (size_t)(n + halo)));
    }
    if (need_S && !ws->S) {
        MPG_CUDA(ctx, cudaMalloc(&ws->S, (size_t)tsize * (size_t)(m + 1) * (size_t)(m + 1)));
        MPG_CUDA(ctx, cudaMalloc(&ws->u, (size_t)tsize * (size_t)(m + 1)));
    }
    if (need_tmp32 && !ws->tmp32) MPG_CUDA(ctx, cudaMalloc(&ws->tmp32, sizeof(float) * (size_t)n));
    *out = ws;
    return MPG_OK;
}

struct History {
    double* inner; int64_t cap_inner;
    double* outer; int64_t cap_outer;
    int64_t ni = 0, no = 0;
    void push_inner(double v) { if (inner && ni < cap_inner) inner[ni] = v; ni++; }
    void push_outer(double a, double b, double c, double d) {
        if (outer && no < cap_outer) { outer[4 * no] = a; outer[4 * no + 1] = b; outer[4 * no + 2] = c; outer[4 * no + 3] = d; }
        no++;
    }
};

// read `count` scalars of type T that kernels left in ctx->dscal[slot0 ...] (one 8-byte slot each)
template <class T>
int read_scalars(mpg_ctx* ctx, int count) {
    MPG_CUDA(ctx, cudaMemcpyAsync(ctx->hscal, ctx->dscal, sizeof(double) * (size_t)count, cudaMemcpyDeviceToHost, ctx->stream));
    MPG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return check_dev_err(ctx);
}
template <class T> T hs(mpg_ctx* ctx, int slot) { T v; memcpy(&v, ctx->hscal + slot, sizeof(T)); return v; }
template <class T> T* ds(mpg_ctx* ctx, int slot) { return reinterpret_cast<T*>(ctx->dscal + slot); }

// Rank-invariant estimate of one rank's share of the problem: local sizes on one GPU; with a communicator attached the global
// row count / world and the all-reduced nonzero count / world (the slabs themselves differ between ranks).
int size_estimate(mpg_ctx* ctx, const mpg_csr* A, Policy& pol) {
    pol.rows_per_rank = (double)A->nrows;
    pol.nnz_per_rank = (double)A->nnz;
    if (!ctx->dist || !pol.needs_residual()) return MPG_OK;
    const double h[2] = {(double)A->nnz, 1.0};
    MPG_CUDA(ctx, cudaMemcpyAsync(ds<double>(ctx, 20), h, sizeof(h), cudaMemcpyHostToDevice, ctx->stream));
    MPG_TRY(dot_dev(ctx, 1, ds<double>(ctx, 20), ds<double>(ctx, 21), ds<double>(ctx, 22)));   // all-reduced over the ranks
    MPG_CUDA(ctx, cudaMemcpyAsync(ctx->hscal + 22, ctx->dscal + 22, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    MPG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const double world = (double)dist_world(ctx);
    pol.nnz_per_rank = ctx->hscal[22] / world;
    pol.rows_per_rank = (double)dist_nglobal(ctx) / world;
    return MPG_OK;
}

// Packed copy of the matrix the inner iterations multiply with (T = the inner precision).  The structure is cached in
// the mpg_csr plan, the values are re-packed at every solve (one pass over the matrix; the caller may have changed them).
// *out stays null when packing is switched off or the structure does not pack well: the CSR kernel is used then.
template <class T>
int get_packed(mpg_ctx* ctx, Workspace* ws, int slot, const mpg_csr* A, const T* vals, const mpg_packed** out) {
    *out = nullptr;
    if (!ctx->tune.spmv_packed) return MPG_OK;
    mpg_packed*& P = ws->packed[slot];
    if (P && !pack_matches(P, A, (int)sizeof(T))) {
        MPG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        pack_free(P);
        P = nullptr;
    }
    if (!P) MPG_TRY(pack_create<T>(ctx, A, vals, &P));
    else if (!(ctx->tune.values_static && ws->packed_src[slot] == (const void*)vals)) MPG_TRY(pack_update<T>(ctx, P, vals));
    ws->packed_src[slot] = vals;
    *out = P;
    return MPG_OK;
}

// Preconditioner application M(w) in the working type T.
//   jac != null: Jacobi, gdmv(1, diag, w, 0, w)  (types.hpp:444-446)
//   bridge:      typesafe_apply with PrecType = float and Type = double (gmres.cpp:12-17): cast, apply, cast back
template <class T>
int apply_prec(mpg_ctx* ctx, int64_t n, T* w, const T* jac, const float* jac32, bool bridge, float* tmp32, mpg_ilu_jacobi* ilu = nullptr) {
    if (bridge) {
        MPG_TRY(cast_copy(ctx, n, reinterpret_cast<const double*>(w), tmp32));
        if (ilu) MPG_TRY(ilu_jacobi_apply_t<float>(ctx, ilu, tmp32));
        else if (jac32) MPG_TRY(gdmv_host(ctx, n, 1.f, jac32, tmp32, 0.f, tmp32));
        MPG_TRY(cast_copy(ctx, n, tmp32, reinterpret_cast<double*>(w)));
        return MPG_OK;
    }
    if (ilu) return ilu_jacobi_apply_t<T>(ctx, ilu, w);   // ILU_Jacobi::apply = ilusv_jacobi (types.hpp:365-367)
    if (jac) MPG_TRY(gdmv_host(ctx, n, T(1), jac, w, T(0), w));
    return MPG_OK;
}

// One restart cycle shared by both drivers: first_vector, s init, Arnoldi loop.  On return *k_out is the
// number of inner iterations performed (the `k` handed to solution_update).
template <class T>
int run_cycle(mpg_ctx* ctx, const mpg_gmres_params& p, Policy& pol, Workspace* ws, const mpg_csr* A, const T* vals, const mpg_packed* P, const T* jac,
              const float* jac32, bool bridge, mpg_ilu_jacobi* ilu, T beta, double Minvb_norm, History& hist, int64_t* k_out, Action* act_out) {
    const int64_t n = ws->n, m = ws->m, ldv = ws->ldv, ldh = m + 1;
    T* V = static_cast<T*>(ws->V);
    T* w = static_cast<T*>(ws->w);
    T* h = static_cast<T*>(ws->h);
    T* cs = static_cast<T*>(ws->cs);
    T* sn = static_cast<T*>(ws->sn);
    T* s = static_cast<T*>(ws->s);
    T* scratch = static_cast<T*>(ws->scratch);

    // GS::first_vector, Orthogonalization.hpp:36-45 (beta was computed by the caller from the same w)
    if (beta != T(0)) MPG_TRY(scal_host(ctx, n, T(1) / beta, w, V));
    else MPG_TRY(fill_host(ctx, n, T(0), V));
    // s = [beta, 0, ...]   gmres.cpp:86-93,198-206
    MPG_TRY(fill_host(ctx, m + 1, T(0), s));
    MPG_TRY(fill_host(ctx, 1, beta, s));

    int64_t pushed_col = -1;   // fused halo: basis column whose boundary rows are already on their way to the neighbours
    // where the fused push rides: in the head of the SpMV kernel that consumes the column (stencil halos: its round trips hide under
    // the interior product) or in the Arnoldi tail that produces it (all-to-all halos, where nearly every slice needs the halo at once)
    const bool push_in_spmv = ctx->tune.dist_push_in_spmv && P && pack_boundary_slices(P) <= 16384 && ctx->tune.dist_spmv_one_launch;
    // one Arnoldi step, enqueued asynchronously
    auto enqueue_iteration = [&](int64_t kk, double* resid_host) -> int {
        // w = A v_k ; M(w)            gmres.cpp:98-102,210-215
        // Jacobi in the working precision rides in the SpMV store (same rounding as the separate gdmv pass)
        const T* rowscale = (jac && !bridge) ? jac : nullptr;
        T* vk = V + (size_t)kk * ldv;
        auto mult = [&](int part) -> int {
            if (P) return spmv_packed<T>(ctx, P, T(1), vk, T(0), w, w, nullptr, rowscale, part);
            return spmv<T>(ctx, A, vals, T(1), vk, T(0), w, w, nullptr, rowscale, part);
        };
        const bool slab = A->ncols > A->nrows;
        const bool direct = slab && dist_basis_ready(ctx);   // fused halo: the neighbours write the halo tail of v_k themselves
        if (direct) {
            // v_k's boundary rows were pushed by the Arnoldi tail that produced v_k; the first vector of a cycle (and an unfused tail)
            // is pushed here.  No wait-and-move launch: the SpMV on the slices that read halo columns waits for the flags itself.
            const bool one_launch = P && pack_boundary_slices(P) <= 16384 && ctx->tune.dist_spmv_one_launch;
            PushArgs spa;
            const bool push_here = one_launch && push_in_spmv && pushed_col != kk;
            if (push_here) MPG_TRY(halo_direct_args<T>(ctx, kk, &spa));   // the SpMV kernel itself sends v_k's boundary rows first
            else if (pushed_col != kk) MPG_TRY(halo_push_direct<T>(ctx, vk, kk));
            HaloWait hw;
            MPG_TRY(halo_wait_args(ctx, &hw));
            if (one_launch) {
                // ONE launch: (push CTAs,) the slices without halo columns, then the CTAs of the others (last in the grid), which wait for the flags
                MPG_TRY(spmv_packed<T>(ctx, P, T(1), vk, T(0), w, w, nullptr, rowscale, SPMV_ORDERED, &hw, nullptr, push_here ? &spa : nullptr));
            } else if (P && pack_boundary_slices(P) <= 16384) {
                MPG_TRY(spmv_packed<T>(ctx, P, T(1), vk, T(0), w, w, nullptr, rowscale, SPMV_INTERIOR));
                MPG_TRY(spmv_packed<T>(ctx, P, T(1), vk, T(0), w, w, nullptr, rowscale, SPMV_BOUNDARY, &hw));
            } else if (P) {
                // nearly every slice reads halo columns (all-to-all halos of power-law matrices): tens of thousands of CTAs would each
                // spin on system-scope flags - one tiny wait kernel in front of the product is cheaper than that
                MPG_TRY(spmv_packed<T>(ctx, P, T(1), vk, T(0), w, w, nullptr, rowscale, SPMV_INTERIOR));
                MPG_TRY(halo_wait_only(ctx));
                MPG_TRY(spmv_packed<T>(ctx, P, T(1), vk, T(0), w, w, nullptr, rowscale, SPMV_BOUNDARY));
            } else {
                MPG_TRY(mult(SPMV_INTERIOR));
                MPG_TRY(halo_wait_only(ctx));
                MPG_TRY(mult(SPMV_BOUNDARY));
            }
        } else if (slab && ctx->tune.dist_overlap) {
            // multi-GPU: send v_k's boundary rows, multiply the rows that need no halo while they travel, then the rest
            MPG_TRY(halo_begin<T>(ctx, vk));
            MPG_TRY(mult(SPMV_INTERIOR));
            MPG_TRY(halo_finish<T>(ctx, vk));
            MPG_TRY(mult(SPMV_BOUNDARY));
        } else {
            MPG_TRY(halo_exchange<T>(ctx, vk));   // fill the halo tail of v_k (no-op on one GPU)
            MPG_TRY(mult(SPMV_ALL));
        }
        if (!rowscale) MPG_TRY(apply_prec<T>(ctx, n, w, jac, jac32, bridge, ws->tmp32, ilu));
        // orth.add_vector(k, w, h)     gmres.cpp:104,217
        if (ctx->tune.fuse_tail) {
            // orthogonalise, then ONE launch for V(:,k+1) = w / h(k+1,k) and the rotations (independent of each other)
            MPG_TRY(add_vector<T>(ctx, p.orth, n, kk, V, ldv, w, h + (size_t)kk * ldh, scratch, true));
            PushArgs pa;
            const bool tail_push = direct && !push_in_spmv;
            if (tail_push) { MPG_TRY(halo_direct_args<T>(ctx, kk + 1, &pa)); pushed_col = kk + 1; }
            return arnoldi_tail<T>(ctx, n, scratch + (kk + 1), w, V + (size_t)(kk + 1) * ldv, kk, h, ldh, cs, sn, s, ws->hist + kk, resid_host,
                                   tail_push ? &pa : nullptr);
        }
        MPG_TRY(add_vector<T>(ctx, p.orth, n, kk, V, ldv, w, h + (size_t)kk * ldh, scratch, false));
        // rot / rotg / rot             gmres.cpp:106-110,219-222 ; |s(k+1)| stays on the device
        return givens_step<T>(ctx, kk, h, ldh, cs, sn, s, ws->hist + kk, resid_host);
    };
    auto no_loss = [](int64_t) -> double { return 0.0; };

    int64_t k = 0;
    Action act = NEXT;
    if (pol.kind == MPG_CONV_BASE) {
        // Convergence::check only counts (IterUtil.hpp:57-65): the whole cycle is enqueued without a read-back
        for (k = 0;; ++k) {
            MPG_TRY(enqueue_iteration(k, nullptr));
            act = pol.check(k + 1, 0.0, Minvb_norm, no_loss);   // gmres.cpp:115,227
            if (act != NEXT) { ++k; break; }
        }
        MPG_CUDA(ctx, cudaMemcpyAsync(ws->hist_host, ws->hist, sizeof(double) * (size_t)k, cudaMemcpyDeviceToHost, ctx->stream));
        MPG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        for (int64_t j = 0; j < k; ++j) hist.push_inner(ws->hist_host[j] / Minvb_norm);
    } else if (pol.needs_residual()) {
        // RelPrecRes_ / RepeatIteration_ (IterUtil.hpp:84-169) restart on the Arnoldi residual.  Instead of the
        // reference's blocking read per iteration (gmres.cpp:225-226) the Givens kernel also writes |s(k+1)| to mapped
        // pinned memory and the host follows it while up to kLookahead further iterations are already enqueued.  A
        // restart decided at step k* leaves at most kLookahead-1 speculative steps behind; they only touch basis columns
        // > k*, Hessenberg columns >= k* and s[k*..], none of which solution_update(k*) reads.  The issue order is a
        // function of the decisions and of kLookahead alone.
        // Speculation wastes up to kLookahead-1 steps per restart; it pays when a step is short compared with the
        // ~30-50 us the GPU idles across a blocking read-back, so the depth follows a bandwidth estimate of the step time.
        // Multi-GPU: every step contains all-reduces, so all ranks MUST enqueue the same number of speculative steps - the depth
        // is derived from rank-invariant quantities only (global rows and nonzeros / world size), never from the local slab.
        const double step_us = (pol.nnz_per_rank * (sizeof(T) + 4) + 1.5 * (double)m * pol.rows_per_rank * sizeof(T)) / 6.0e6;
        const int64_t kLookahead = ctx->tune.lookahead > 0 ? ctx->tune.lookahead : (step_us < 200.0 ? 3 : (step_us < 1000.0 ? 2 : 1));
        volatile double* hh = ws->hist_host;
        for (int64_t j = 0; j <= m; ++j) hh[j] = -1.0;   // |s| >= 0: negative = not written yet (stream is idle here)
        int64_t issued = 0, checked = 0;
        while (issued < std::min<int64_t>(m, kLookahead)) { MPG_TRY(enqueue_iteration(issued, ws->hist_host + issued)); ++issued; }
        for (;;) {
            int64_t spins = 0;
            while (hh[checked] < 0.0) {
                if ((++spins & 0xFFFF) == 0) {   // surface a device fault instead of spinning forever
                    const cudaError_t e = cudaStreamQuery(ctx->stream);
                    if (e != cudaSuccess && e != cudaErrorNotReady) return fail(ctx, MPG_ERR_CUDA, std::string("stream error: ") + cudaGetErrorString(e));
                }
            }
            const double ares = hh[checked];
            hist.push_inner(ares / Minvb_norm);
            act = pol.check(checked + 1, ares, Minvb_norm, no_loss);
            ++checked;
            if (act != NEXT) break;
            if (issued < m) { MPG_TRY(enqueue_iteration(issued, ws->hist_host + issued)); ++issued; }
        }
        k = checked;
    } else {
        // LostOrthogonality_ (IterUtil.hpp:172-227) needs a gemv-T on the basis per check: one read-back per iteration
        for (k = 0;; ++k) {
            MPG_TRY(enqueue_iteration(k, nullptr));
            MPG_CUDA(ctx, cudaMemcpyAsync(ws->hist_host + k, ws->hist + k, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
            MPG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            const double ares = ws->hist_host[k];
            hist.push_inner(ares / Minvb_norm);
            int rc_loss = MPG_OK;
            auto loss_inc = [&](int64_t kk) -> double {
                // LostOrthogonality_Convergence::check, IterUtil.hpp:205-216 (kk = k+1).  Note the reference reads basis
                // column kk+1, which add_vector has not written yet in this cycle; reproduced as is.
                T* S = static_cast<T*>(ws->S);
                T* u = static_cast<T*>(ws->u);
                const int64_t ldS = m + 1;
                const int k1 = (int)kk + 1;
                T* scol = S + (size_t)(kk + 1) * ldS;
                rc_loss = gemvt<T>(ctx, n, k1, V, ldv, T(1), V + (size_t)(kk + 1) * ldv, T(0), u);
                LocalScope replicated(ctx);   // S, u, s_col are replicated on every rank
                if (rc_loss == MPG_OK) rc_loss = cast_copy(ctx, k1, u, scol);
                if (rc_loss == MPG_OK) rc_loss = gemvn<T>(ctx, k1, k1, S, ldS, T(-1), u, T(1), scol, false, nullptr, nullptr, nullptr);
                if (rc_loss == MPG_OK) rc_loss = dot_dev(ctx, k1, scol, scol, ds<T>(ctx, 16));
                if (rc_loss != MPG_OK) return 0.0;
                cudaMemcpyAsync(ctx->hscal + 16, ctx->dscal + 16, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
                cudaStreamSynchronize(ctx->stream);
                return (double)hs<T>(ctx, 16);
            };
            act = pol.check(k + 1, ares, Minvb_norm, loss_inc);   // gmres.cpp:115,227
            MPG_TRY(rc_loss);
            if (act != NEXT) { ++k; break; }
        }
    }
    *k_out = k;
    *act_out = act;
    return MPG_OK;
}

// ---- GMRES-IR: gmres_singleUpdate, gmres.cpp:135-245 ----------------------------------------------------------
int solve_mixed(mpg_ctx* ctx, const mpg_gmres_params& p, const mpg_csr* A, const double* vals64, const float* vals32, const float* jac32,
                mpg_ilu_jacobi* ilu, const double* b, double* x, mpg_gmres_stats* st, History& hist) {
    const int64_t n = A->nrows, m = p.restart_length;
    Workspace* ws = nullptr;
    Trace tr(ctx);
    MPG_TRY(get_workspace(ctx, n, m, 4, p.conv == MPG_CONV_ORTHLOSS, false, &ws));
    tr.mark("workspace");
    MPG_TRY(dist_exchange_basis(ctx, ws->V, ws->ldv, 4));
    tr.mark("exchange_basis");
    float* w = static_cast<float*>(ws->w);
    float* h = static_cast<float*>(ws->h);
    float* s = static_cast<float*>(ws->s);
    float* V = static_cast<float*>(ws->V);
    const mpg_packed* packed = nullptr;
    MPG_TRY(get_packed<float>(ctx, ws, 0, A, vals32, &packed));
    // fp64 operator of the outer residual on the same packed structure (one more pass over the values per solve; the CSR kernel
    // gathers x once per nonzero and runs at ~0.55 of the roofline, the packed one at ~0.9)
    const mpg_packed* packed64 = nullptr;
    tr.mark("packed fp32 (plan + values)");
    // host-buffer entry point with host_overlap: vals64 may still be in flight on the copy stream.  Nothing before the first fp64
    // residual that needs the operator touches it; need_v64() makes the compute stream wait for the copy and packs it then.
    DeferredV64* df = ctx->defer;
    bool v64_ready = (df == nullptr);
    auto need_v64 = [&]() -> int {
        if (!v64_ready) {
            int state;
            while ((state = df->recorded.load(std::memory_order_acquire)) == 0) std::this_thread::yield();
            if (state < 0) return fail(ctx, MPG_ERR_CUDA, "gmres_solve_host: the host-to-device copy of the fp64 values failed");
            MPG_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, df->ev, 0));
            v64_ready = true;
        }
        if (!packed64 && ctx->tune.residual_packed && packed) MPG_TRY(get_packed<double>(ctx, ws, 1, A, vals64, &packed64));
        return MPG_OK;
    };
    if (v64_ready) MPG_TRY(need_v64());
    tr.mark("packed fp64");
    Policy pol(p);
    MPG_TRY(size_estimate(ctx, A, pol));
    if (p.conv == MPG_CONV_ORTHLOSS) {
        MPG_TRY(fill_host(ctx, (m + 1) * (m + 1), 0.f, static_cast<float*>(ws->S)));  // IterUtil.hpp:191
        MPG_CUDA(ctx, cudaMemsetAsync(ws->V, 0, sizeof(float) * (size_t)ws->ldv * (size_t)(m + 1), ctx->stream));  // fresh View (see get_workspace)
    }

    // setup norms, gmres.cpp:162-168 — three reductions, one readback
    MPG_TRY(nrm2_dev(ctx, n, b, ds<double>(ctx, 0)));
    MPG_TRY(cast_copy(ctx, n, b, w));
    MPG_TRY(apply_prec<float>(ctx, n, w, jac32, nullptr, false, nullptr, ilu));
    MPG_TRY(nrm2_dev(ctx, n, w, ds<float>(ctx, 1)));
    MPG_TRY(nrm2_dev(ctx, A->nnz, vals32, ds<float>(ctx, 2)));
    MPG_TRY(read_scalars<double>(ctx, 3));
    const double b_norm = hs<double>(ctx, 0);
    const double Minvb_norm = hs<float>(ctx, 1);
    const double A_norm = hs<float>(ctx, 2);
    st->b_norm = b_norm; st->Minvb_norm = Minvb_norm; st->A_norm = A_norm;
    tr.mark("set-up norms");

    for (int64_t i = 0;; ++i) {
        // r = b - A x (fp64), w = (float) r : one fused kernel  (gmres.cpp:173-175)
        const double* xin = x;
        if (ws->halo > 0) {   // multi-GPU: [x_local ; halo]
            double* xe = static_cast<double*>(ws->xext);
            MPG_TRY(cast_copy(ctx, n, x, xe));
            MPG_TRY(halo_exchange<double>(ctx, xe));
            xin = xe;
        }
        if (i == 0 && df && df->x0_zero && ws->halo == 0) {
            MPG_TRY(cast_copy(ctx, n, b, w));   // x0 == 0 (checked on the host): r = b - A*0 = b, so w = (float) b - same bits, no operator needed yet
        } else {
            MPG_TRY(need_v64());
            if (packed64) MPG_TRY(spmv_packed<double>(ctx, packed64, -1.0, xin, 1.0, b, nullptr, w, nullptr, SPMV_ALL));
            else MPG_TRY(spmv<double>(ctx, A, vals64, -1.0, xin, 1.0, b, nullptr, w));
        }
        MPG_TRY(nrm2_dev(ctx, n, w, ds<float>(ctx, 0)));                       // r_norm   :176
        MPG_TRY(apply_prec<float>(ctx, n, w, jac32, nullptr, false, nullptr, ilu)); // M(w)     :177
        if (jac32 || ilu) MPG_TRY(nrm2_dev(ctx, n, w, ds<float>(ctx, 1)));     // beta     :179 (same vector when M = I)
        MPG_TRY(nrm2_dev(ctx, n, x, ds<double>(ctx, 2)));                      // x_norm   :181
        MPG_TRY(read_scalars<double>(ctx, 3));
        const double r_norm = hs<float>(ctx, 0);
        const float beta = (jac32 || ilu) ? hs<float>(ctx, 1) : hs<float>(ctx, 0);
        const double x_norm = hs<double>(ctx, 2);
        hist.push_outer(r_norm, b_norm + A_norm * x_norm, beta, x_norm);
        const Action a0 = pol.check_initial(r_norm, b_norm + A_norm * x_norm, beta, Minvb_norm);   // :184
        if (a0 == CONVERGED) { st->status = 1; st->rel_prec_res = double(beta / Minvb_norm); st->outer_i = i; break; }
        if (a0 == ABORTED) { st->status = 3; st->outer_i = i; break; }

        int64_t k = 0;
        Action act = NEXT;
        MPG_TRY(run_cycle<float>(ctx, p, pol, ws, A, vals32, packed, jac32, nullptr, false, ilu, beta, Minvb_norm, hist, &k, &act));
        if (act == ABORTED) { st->status = 3; st->outer_i = i; break; }

        // solution_update, gmres.cpp:276-290: y = triu(H)^-1 s ; x += (double)(V_k y)  (Orthogonalization.hpp:67-73)
        MPG_TRY(trsv<float>(ctx, 1, 0, k, h, m + 1, s));
        MPG_TRY(gemvn<float>(ctx, n, (int)k, V, ws->ldv, 1.f, s, 0.f, w, false, nullptr, nullptr, x));
        tr.mark("cycle + update");
    }
    st->total_iters = pol.total_iters;
    st->total_restarts = pol.total_restarts;
    return MPG_OK;
}

// ---- uniform precision: gmres_baseline<Orth,Device,Type,PrecType>, gmres.cpp:24-133 ---------------------------
template <class T>
int solve_uniform(mpg_ctx* ctx, const mpg_gmres_params& p, const mpg_csr* A, const T* vals, const T* jac, const float* jac32, bool bridge,
                  mpg_ilu_jacobi* ilu, const T* b, T* x, mpg_gmres_stats* st, History& hist) {
    const int64_t n = A->nrows, m = p.restart_length;
    Workspace* ws = nullptr;
    MPG_TRY(get_workspace(ctx, n, m, (int)sizeof(T), p.conv == MPG_CONV_ORTHLOSS, bridge, &ws));
    MPG_TRY(dist_exchange_basis(ctx, ws->V, ws->ldv, (int)sizeof(T)));
    T* w = static_cast<T*>(ws->w);
    T* h = static_cast<T*>(ws->h);
    T* s = static_cast<T*>(ws->s);
    T* V = static_cast<T*>(ws->V);
    const mpg_packed* packed = nullptr;
    MPG_TRY(get_packed<T>(ctx, ws, 0, A, vals, &packed));
    Policy pol(p);
    MPG_TRY(size_estimate(ctx, A, pol));
    if (p.conv == MPG_CONV_ORTHLOSS) {
        MPG_TRY(fill_host(ctx, (m + 1) * (m + 1), T(0), static_cast<T*>(ws->S)));
        MPG_CUDA(ctx, cudaMemsetAsync(ws->V, 0, sizeof(T) * (size_t)ws->ldv * (size_t)(m + 1), ctx->stream));
    }

    MPG_TRY(nrm2_dev(ctx, n, b, ds<T>(ctx, 0)));                               // b_norm      :54
    MPG_TRY(cast_copy(ctx, n, b, w));                                          // copy(b, w)  :56
    MPG_TRY(apply_prec<T>(ctx, n, w, jac, jac32, bridge, ws->tmp32, ilu));     //             :57
    MPG_TRY(nrm2_dev(ctx, n, w, ds<T>(ctx, 1)));                               // Minvb_norm  :58
    MPG_TRY(nrm2_dev(ctx, A->nnz, vals, ds<T>(ctx, 2)));                       // A_norm      :60
    MPG_TRY(read_scalars<T>(ctx, 3));
    const T b_norm = hs<T>(ctx, 0), Minvb_norm = hs<T>(ctx, 1), A_norm = hs<T>(ctx, 2);
    st->b_norm = b_norm; st->Minvb_norm = Minvb_norm; st->A_norm = A_norm;
    const bool have_prec = bridge || jac != nullptr || ilu != nullptr;

    for (int64_t i = 0;; ++i) {
        // w = b - A x in Type: one kernel (gmres.cpp:62-63)
        const T* xin = x;
        if (ws->halo > 0) {
            T* xe = static_cast<T*>(ws->xext);
            MPG_TRY(cast_copy(ctx, n, x, xe));
            MPG_TRY(halo_exchange<T>(ctx, xe));
            xin = xe;
        }
        if (packed && ctx->tune.residual_packed) MPG_TRY(spmv_packed<T>(ctx, packed, T(-1), xin, T(1), b, w, nullptr, nullptr, SPMV_ALL));
        else MPG_TRY(spmv<T>(ctx, A, vals, T(-1), xin, T(1), b, w, nullptr));
        MPG_TRY(nrm2_dev(ctx, n, w, ds<T>(ctx, 0)));                           // r_norm :67
        MPG_TRY(apply_prec<T>(ctx, n, w, jac, jac32, bridge, ws->tmp32, ilu)); //        :68
        if (have_prec) MPG_TRY(nrm2_dev(ctx, n, w, ds<T>(ctx, 1)));            // beta   :70
        MPG_TRY(nrm2_dev(ctx, n, x, ds<T>(ctx, 2)));                           // x_norm :72
        MPG_TRY(read_scalars<T>(ctx, 3));
        const T r_norm = hs<T>(ctx, 0);
        const T beta = have_prec ? hs<T>(ctx, 1) : r_norm;
        const T x_norm = hs<T>(ctx, 2);
        const double normalization = b_norm + A_norm * x_norm;                 // Type arithmetic, :74
        hist.push_outer(r_norm, normalization, beta, x_norm);
        const Action a0 = pol.check_initial(r_norm, normalization, beta, Minvb_norm);
        if (a0 == CONVERGED) { st->status = 1; st->rel_prec_res = double(T(beta / Minvb_norm)); st->outer_i = i; break; }
        if (a0 == ABORTED) { st->status = 3; st->outer_i = i; break; }

        int64_t k = 0;
        Action act = NEXT;
        MPG_TRY(run_cycle<T>(ctx, p, pol, ws, A, vals, packed, jac, jac32, bridge, ilu, beta, (double)Minvb_norm, hist, &k, &act));
        if (act == ABORTED) { st->status = 3; st->outer_i = i; break; }

        // solution_update, gmres.cpp:291-303: y = triu(H)^-1 s ; x = 1*V_k y + 1*x (Orthogonalization.hpp:62-65)
        MPG_TRY(trsv<T>(ctx, 1, 0, k, h, m + 1, s));
        MPG_TRY(gemvn<T>(ctx, n, (int)k, V, ws->ldv, T(1), s, T(1), x, false, nullptr, nullptr, nullptr));
    }
    st->total_iters = pol.total_iters;
    st->total_restarts = pol.total_restarts;
    return MPG_OK;
}

}  // namespace

extern "C" int mpg_gmres_solve(mpg_ctx* ctx, const mpg_gmres_params* pp, const mpg_csr* A, const double* vals64, const float* vals32_in,
                               const double* b, double* x, mpg_gmres_stats* st, double* hist_inner, int64_t cap_inner, double* hist_outer,
                               int64_t cap_outer) {
    MPG_REQUIRE(ctx, pp && A && vals64 && b && x && st, "gmres_solve: null argument");
    const mpg_gmres_params p = *pp;
    MPG_REQUIRE(ctx, p.restart_length >= 1 && p.restart_length + 1 <= kMaxCols, "gmres_solve: restart length must be in [1, 255]");
    MPG_REQUIRE(ctx, p.mode >= 0 && p.mode <= 3 && p.orth >= 0 && p.orth <= 2 && p.conv >= 0 && p.conv <= 3 && p.prec >= 0 && p.prec <= 2,
                "gmres_solve: bad enum");
    MPG_REQUIRE(ctx, p.prec != MPG_PREC_ILU_JACOBI || (ctx->dist == nullptr && p.jacobi_steps >= 0 && p.jacobi_steps <= 1000),
                "gmres_solve: ilu_jacobi needs the whole matrix on one GPU and 0 <= jacobi_steps <= 1000");
    MPG_REQUIRE(ctx, (int64_t)A->ncols == (int64_t)A->nrows + dist_halo(ctx), "gmres_solve: matrix must be square (local slab: nrows + halo columns)");
    memset(st, 0, sizeof(*st));
    History hist{hist_inner, cap_inner, hist_outer, cap_outer};
    const int64_t n = A->nrows, nnz = A->nnz;
    const int64_t launches0 = ctx->launches;

    cudaEvent_t e0, e1;
    MPG_CUDA(ctx, cudaEventCreate(&e0));
    MPG_CUDA(ctx, cudaEventCreate(&e1));

    // temporaries that depend on the mode (freed below)
    float* vals32_own = nullptr;
    double* vals_rt = nullptr;
    float *jac32 = nullptr, *b32 = nullptr, *x32 = nullptr;
    double* jac64 = nullptr;
    double* ilu_vals = nullptr;
    mpg_ilu_jacobi* ilu = nullptr;
    int rc = MPG_OK;
    auto cleanup = [&]() {
        cudaStreamSynchronize(ctx->stream);
        mpg_ilu_jacobi_destroy(ilu);
        pool_free(ilu_vals);
        pool_free(vals32_own); pool_free(vals_rt); pool_free(jac32); pool_free(jac64); pool_free(b32); pool_free(x32);
        cudaEventDestroy(e0); cudaEventDestroy(e1);
    };
#define MPG_TRY_C(expr) do { rc = (expr); if (rc != MPG_OK) { cleanup(); return rc; } } while (0)
#define MPG_CUDA_C(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { cleanup(); return fail(ctx, MPG_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); } } while (0)

    const float* vals32 = vals32_in;
    if (!vals32) {  // SparseMatrix<float,Device>(A): cast values, share structure (types_cuda.hpp:82-101)
        MPG_CUDA_C(pool_alloc(ctx, &vals32_own, sizeof(float) * (size_t)std::max<int64_t>(nnz, 1)));
        MPG_TRY_C(cast_copy(ctx, nnz, vals64, vals32_own));
        vals32 = vals32_own;
    }
    if (p.prec == MPG_PREC_ILU_JACOBI) {
        // ILU_Jacobi<PrecType>(ilu0<PrecType>(A), jacobi_steps) on the fp64 matrix (gmres_perf_test.cpp:75-78,145-148): the "ilu took" window
        const bool prec_f64 = (p.mode == MPG_MODE_BASELINE);   // PrecType = double only for <double,double>
        MPG_CUDA_C(pool_alloc(ctx, &ilu_vals, sizeof(double) * (size_t)std::max<int64_t>(nnz, 1)));
        MPG_TRY_C(mpg_ilu0_f64(ctx, A, vals64, prec_f64 ? 0 : 1, ilu_vals));
        if (prec_f64) MPG_TRY_C(mpg_ilu_jacobi_create_f64(ctx, A, ilu_vals, (int)p.jacobi_steps, &ilu));
        else MPG_TRY_C(mpg_ilu_jacobi_create_f32(ctx, A, ilu_vals, (int)p.jacobi_steps, &ilu));
    }
    if (p.mode == MPG_MODE_MIXED) {
        if (p.prec == MPG_PREC_JACOBI) {  // Jacobi<float>(A) on the fp32-cast matrix, gmres_perf_test.cpp:149
            MPG_CUDA_C(pool_alloc(ctx, &jac32, sizeof(float) * (size_t)n));
            MPG_TRY_C(mpg_jacobi_diag_f32(ctx, A, vals32, jac32));
        }
        MPG_CUDA_C(cudaEventRecord(e0, ctx->stream));
        MPG_TRY_C(solve_mixed(ctx, p, A, vals64, vals32, jac32, ilu, b, x, st, hist));
    } else if (p.mode == MPG_MODE_BASELINE || p.mode == MPG_MODE_SINGLE_PREC) {
        // DoBaselineProblem hands the solver the fp32-rounded matrix converted back to double
        // (gmres_perf_test.cpp:66,101 + implicit conversion)
        MPG_CUDA_C(pool_alloc(ctx, &vals_rt, sizeof(double) * (size_t)std::max<int64_t>(nnz, 1)));
        MPG_TRY_C(cast_copy(ctx, nnz, vals32, vals_rt));
        const bool bridge = (p.mode == MPG_MODE_SINGLE_PREC);
        if (p.prec == MPG_PREC_JACOBI) {
            if (bridge) { MPG_CUDA_C(pool_alloc(ctx, &jac32, sizeof(float) * (size_t)n)); MPG_TRY_C(mpg_jacobi_diag_f32(ctx, A, vals32, jac32)); }
            else { MPG_CUDA_C(pool_alloc(ctx, &jac64, sizeof(double) * (size_t)n)); MPG_TRY_C(mpg_jacobi_diag_f64(ctx, A, vals64, jac64)); }
        }
        MPG_CUDA_C(cudaEventRecord(e0, ctx->stream));
        MPG_TRY_C(solve_uniform<double>(ctx, p, A, vals_rt, jac64, jac32, bridge, ilu, b, x, st, hist));
    } else {
        if (p.prec == MPG_PREC_JACOBI) { MPG_CUDA_C(pool_alloc(ctx, &jac32, sizeof(float) * (size_t)n)); MPG_TRY_C(mpg_jacobi_diag_f32(ctx, A, vals32, jac32)); }
        MPG_CUDA_C(pool_alloc(ctx, &b32, sizeof(float) * (size_t)n));
        MPG_CUDA_C(pool_alloc(ctx, &x32, sizeof(float) * (size_t)n));
        MPG_TRY_C(cast_copy(ctx, n, b, b32));   // copy(b, b_type)  gmres_perf_test.cpp:97-98
        MPG_TRY_C(cast_copy(ctx, n, x, x32));
        MPG_CUDA_C(cudaEventRecord(e0, ctx->stream));
        MPG_TRY_C(solve_uniform<float>(ctx, p, A, vals32, jac32, nullptr, false, ilu, b32, x32, st, hist));
        MPG_CUDA_C(cudaEventRecord(e1, ctx->stream));
        MPG_TRY_C(cast_copy(ctx, n, x32, x));   // copy(x_type, x)  gmres_perf_test.cpp:104-105
    }
    if (p.mode != MPG_MODE_SINGLE) MPG_CUDA_C(cudaEventRecord(e1, ctx->stream));
    MPG_CUDA_C(cudaStreamSynchronize(ctx->stream));
    MPG_CUDA_C(cudaEventSynchronize(e1));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    st->solve_ms = ms;
    st->n_hist_inner = hist.ni;
    st->n_hist_outer = hist.no;
    st->launches = ctx->launches - launches0;
    cleanup();
    return MPG_OK;
}

