// sell.cu — packed SpMV operator: the CSR matrix re-laid as 32-lane slices stored position-major (sliced ELLPACK).
// Reference surface: the same y = alpha*A*x + beta*y of kernels.hpp:159-165; the packed form plays the role of the
// reference's per-matrix library handles (create_cuda_handles, types_cuda.hpp:53-60: cusparse descriptors built once
// per SparseMatrix and re-used by every spmv call).
//
// Why: the CSR kernel (spmv.cu) streams the matrix perfectly but gathers x with one thread per NONZERO, so one warp
// gather touches ~10 different 128-byte lines of x (ncu: the kernel is L1TEX-tag bound at 73-77 % of the HBM
// roofline, profiles/r01_ncu_summary.md).  In the packed layout one warp owns 32 rows and lane = row:
//   * the p-th nonzeros of the 32 rows are adjacent in memory, 4 at a time per lane, so indices and values are still
//     read with fully coalesced 16-byte streaming loads;
//   * a warp gather reads x[col_p(row)] for 32 rows - for stencil-like matrices 32 nearly consecutive entries of x:
//     1-2 lines instead of ~10;
//   * every row is summed by its own lane in nonzero order: no shared memory, no cross-thread reduction, no fix-up.
//
// Two structures, chosen per matrix (mpg_sell_plan::mode):
//   PLAIN  slice s = rows [32 s, 32 s + 32).  Stencil-like matrices (padding <= 25 %).
//   SIGMA  SELL-C-sigma for uneven row lengths (power-law): rows longer than CHUNK = 256 nonzeros are cut into
//          "virtual rows" of at most CHUNK nonzeros; the virtual rows are sorted by length (descending, ties by index)
//          inside windows of SIGMA = 4096 virtual rows; slice s = sorted positions [32 s, 32 s + 32).  A lane writes its
//          sum to the row it stands for (vout[pos] >= 0) or, for a piece of a cut row, to partial[-1 - vout[pos]]; a
//          small fix-up kernel adds the pieces of each cut row in nonzero order and applies the epilogue.  Per-row
//          nonzero order is preserved inside every piece, results are bit-reproducible, y stays in caller order.
// Layout of a slice (both modes): 32 * L_s elements, L_s = longest (virtual) row of the slice, rounded up to a whole
// group of G = 4 when that pads <= 10 %.  The first floor(L_s / 4) * 4 positions are stored in groups of 4 per lane:
// index of (lane l, position p) at slice_off[s] + (p / 4) * 128 + l * 4 + p % 4; the L_s % 4 trailing positions one per
// lane: slice_off[s] + floor(L_s / 4) * 128 + (p % 4) * 32 + l.  fp32 values use the same addresses.  fp64 values share the
// SAME index array (the reference's SparseMatrix<float> aliases row_map / inds of the fp64 matrix, types_cuda.hpp:82-91)
// and store each group as two half-groups so that both 16-byte loads of a warp are fully coalesced:
// (p / 4) * 128 + ((p % 4) / 2) * 64 + l * 2 + p % 2.  Shorter rows are padded by repeating their first column with
// value 0.  The packed INDICES depend only on the structure and are cached in the mpg_csr plan; the packed VALUES belong
// to an mpg_packed object and are refreshed with mpg_pack_update.
#include <vector>

#include "common.cuh"

using namespace mpg;

namespace {

constexpr int SLICE = 32;
constexpr int G = 4;
constexpr int CHUNK = 256;     // longest virtual row (multiple of G)
constexpr int SIGMA = 4096;    // sorting window in virtual rows (multiple of SLICE, power of two for the bitonic network)
enum { MODE_PLAIN = 0, MODE_SIGMA = 1 };

__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }

// where lane `pos` (a row in PLAIN mode, a sorted virtual row in SIGMA mode) finds its nonzeros in the CSR arrays
struct LaneSrc {
    int nrows;                 // PLAIN: rows; SIGMA: virtual rows
    const int* row_map;        // PLAIN
    const int* vstart;         // SIGMA: CSR offset of the first nonzero of the sorted virtual row
    const int* vlen;           // SIGMA
    __device__ __forceinline__ void get(int pos, int& start, int& len) const {
        if (pos >= nrows) { start = 0; len = 0; return; }
        if (vstart) { start = __ldg(vstart + pos); len = __ldg(vlen + pos); }
        else { start = __ldg(row_map + pos); len = __ldg(row_map + pos + 1) - start; }
    }
};

__device__ __forceinline__ int round_len(int len) {
    const int rem = len % G;
    if (rem > 0 && (G - rem) * 10 <= len) len += G - rem;   // one more 16-byte group is cheaper than up to 3 single-element loads: cd27 27 -> 28; lap2d stays 4 + 1
    return len;
}

// slice_len[s] = 32 * L_s (elements)
__global__ void sell_len_kernel(LaneSrc src, int nslices, int64_t* __restrict__ slice_len, int* has_rem) {
    const int s = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (s >= nslices) return;
    int start, len;
    src.get(s * SLICE + lane, start, len);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) len = max(len, __shfl_xor_sync(0xffffffffu, len, o));
    len = round_len(len);
    if (lane == 0) {
        slice_len[s] = (int64_t)len * SLICE;
        if (len % G) *has_rem = 1;   // benign race: every writer stores 1
    }
}

// exclusive prefix sum of n values, out[n] = total.  One 1024-thread block: per-thread chunk sums, Hillis-Steele over the
// 1024 chunk sums, per-thread chunk write-out.  Plan-time only.
template <class TI, class TO>
__global__ void __launch_bounds__(1024) scan_kernel(int64_t n, const TI* __restrict__ in, TO* __restrict__ out) {
    __shared__ int64_t part[1024];
    const int t = threadIdx.x;
    const int64_t chunk = (n + 1023) / 1024;
    const int64_t lo = min((int64_t)t * chunk, n), hi = min(lo + chunk, n);
    int64_t s = 0;
    for (int64_t i = lo; i < hi; ++i) s += (int64_t)in[i];
    part[t] = s;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {
        const int64_t v = t >= off ? part[t - off] : 0;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    int64_t run = t ? part[t - 1] : 0;
    for (int64_t i = lo; i < hi; ++i) {
        const int64_t v = (int64_t)in[i];
        out[i] = (TO)run;
        run += v;
    }
    if (t == 1023) out[n] = (TO)part[1023];
}

// ---- SIGMA structure -------------------------------------------------------------------------------------------------
// per row: number of virtual rows, and that number again if the row is cut (0 otherwise)
__global__ void vrow_count_kernel(int nrows, const int* __restrict__ row_map, int* __restrict__ nch, int* __restrict__ is_split, int* __restrict__ nch_split) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrows) return;
    const int len = row_map[r + 1] - row_map[r];
    const int c = max(1, (len + CHUNK - 1) / CHUNK);
    nch[r] = c;
    is_split[r] = c > 1;
    nch_split[r] = c > 1 ? c : 0;
}
// virtual rows in (row, piece) order: CSR offset, length, destination
__global__ void vrow_fill_kernel(int nrows, const int* __restrict__ row_map, const int* __restrict__ vrow_off, const int* __restrict__ split_off,
                                 const int* __restrict__ chunk_off, int* __restrict__ vstart, int* __restrict__ vlen, int* __restrict__ vout,
                                 int* __restrict__ split_rows, int* __restrict__ chunk_base) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrows) return;
    const int rs = row_map[r], len = row_map[r + 1] - rs;
    const int v0 = vrow_off[r], c = vrow_off[r + 1] - v0;
    if (c == 1) { vstart[v0] = rs; vlen[v0] = len; vout[v0] = r; return; }
    const int j = split_off[r], p0 = chunk_off[r];
    split_rows[j] = r;
    chunk_base[j] = p0;
    for (int q = 0; q < c; ++q) {
        vstart[v0 + q] = rs + q * CHUNK;
        vlen[v0 + q] = min(CHUNK, len - q * CHUNK);
        vout[v0 + q] = -1 - (p0 + q);
    }
}
// one CTA per window of SIGMA virtual rows: bitonic sort of the unique keys ((CHUNK - len) << 12 | index in window), i.e.
// by length descending, ties by index ascending; writes the three per-lane arrays in sorted order
__global__ void __launch_bounds__(512) vrow_sort_kernel(int nv, const int* __restrict__ vstart, const int* __restrict__ vlen, const int* __restrict__ vout,
                                                         int* __restrict__ sstart, int* __restrict__ slen, int* __restrict__ sout) {
    __shared__ unsigned int key[SIGMA];
    const int base = blockIdx.x * SIGMA;
    for (int i = threadIdx.x; i < SIGMA; i += blockDim.x) {
        const int v = base + i;
        key[i] = v < nv ? (((unsigned)(CHUNK - vlen[v])) << 12) | (unsigned)i : 0xffffffffu;   // past the end: sorts last
    }
    __syncthreads();
    for (int k = 2; k <= SIGMA; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < SIGMA; i += blockDim.x) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const unsigned a = key[i], b = key[ixj];
                    const bool up = (i & k) == 0;
                    if ((a > b) == up) { key[i] = b; key[ixj] = a; }
                }
            }
            __syncthreads();
        }
    }
    for (int i = threadIdx.x; i < SIGMA; i += blockDim.x) {
        const int pos = base + i;
        if (pos >= nv) continue;
        const int v = base + (int)(key[i] & 0xfffu);
        sstart[pos] = vstart[v];
        slen[pos] = vlen[v];
        sout[pos] = vout[v];
    }
}

// packed indices (structure); warp = slice, lane = (virtual) row.  halo_flag[s] = 1 if the slice references a column >= halo_from
__global__ void sell_fill_inds_kernel(LaneSrc src, int nslices, int halo_from, const int* __restrict__ inds, const int64_t* __restrict__ slice_off,
                                      int* __restrict__ sinds, int* __restrict__ halo_flag) {
    const int s = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (s >= nslices) return;
    const int64_t off = slice_off[s];
    const int L = (int)((slice_off[s + 1] - off) / SLICE);
    int rs, len;
    src.get(s * SLICE + lane, rs, len);
    const int pad = len > 0 ? __ldg(inds + rs) : 0;
    int any_halo = 0;
    const int ng = L / G;
    for (int g = 0; g < ng; ++g) {
        int c[G];
#pragma unroll
        for (int q = 0; q < G; ++q) {
            const int p = g * G + q;
            c[q] = p < len ? __ldg(inds + rs + p) : pad;
            any_halo |= c[q] >= halo_from;
        }
        *reinterpret_cast<int4*>(sinds + off + (int64_t)g * SLICE * G + lane * G) = make_int4(c[0], c[1], c[2], c[3]);
    }
    for (int p = ng * G; p < L; ++p) {
        const int c = p < len ? __ldg(inds + rs + p) : pad;
        any_halo |= c >= halo_from;
        sinds[off + (int64_t)ng * SLICE * G + (int64_t)(p - ng * G) * SLICE + lane] = c;
    }
    any_halo = __any_sync(0xffffffffu, any_halo);
    if (lane == 0 && halo_flag) halo_flag[s] = any_halo;
}

template <class T>
__global__ void sell_fill_vals_kernel(LaneSrc src, int nslices, const T* __restrict__ vals, const int64_t* __restrict__ slice_off, T* __restrict__ svals) {
    const int s = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (s >= nslices) return;
    const int64_t off = slice_off[s];
    const int L = (int)((slice_off[s + 1] - off) / SLICE);
    int rs, len;
    src.get(s * SLICE + lane, rs, len);
    const int ng = L / G;
    for (int g = 0; g < ng; ++g) {
        T v[G];
#pragma unroll
        for (int q = 0; q < G; ++q) {
            const int p = g * G + q;
            v[q] = p < len ? __ldg(vals + rs + p) : T(0);
        }
        T* dst = svals + off + (int64_t)g * SLICE * G;
        if (sizeof(T) == 4) {
            *reinterpret_cast<float4*>(dst + lane * G) = make_float4((float)v[0], (float)v[1], (float)v[2], (float)v[3]);
        } else {
            *reinterpret_cast<double2*>(dst + lane * 2) = make_double2((double)v[0], (double)v[1]);
            *reinterpret_cast<double2*>(dst + 64 + lane * 2) = make_double2((double)v[2], (double)v[3]);
        }
    }
    for (int p = ng * G; p < L; ++p) svals[off + (int64_t)ng * SLICE * G + (int64_t)(p - ng * G) * SLICE + lane] = p < len ? __ldg(vals + rs + p) : T(0);
}

// one group of a lane: 4 values
__device__ __forceinline__ void load_group(const float* gbase, int lane, float v[4]) {
    const float4 t = ldg_stream(reinterpret_cast<const float4*>(gbase) + lane);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
__device__ __forceinline__ void load_group(const double* gbase, int lane, double v[4]) {
    const double2 a = ldg_stream(reinterpret_cast<const double2*>(gbase) + lane);
    const double2 b = ldg_stream(reinterpret_cast<const double2*>(gbase + 64) + lane);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}

// xadd != null: Jacobi sweep of the ILU factors (ilu.cu) - the SpMV result v = b - Op x is the correction, the stored value is
//   x + v (axpy(1, temp, x), kernels.hpp:239) or, with rowscale = D^-1, 1*x + (1*d)*v (gdmv(1, diag, temp, 1, x), kernels.hpp:247)
template <class T>
__device__ __forceinline__ T sell_epilogue(int r, T sum, T alpha, T beta, const T* y_in, const T* __restrict__ rowscale, const T* xadd) {
    T v = (beta == T(0)) ? alpha * sum : fma(alpha, sum, beta * y_in[r]);
    if (xadd) {
        const T xo = xadd[r];
        if (rowscale) v = add_rn(mul_rn(T(1), xo), mul_rn(mul_rn(T(1), __ldg(rowscale + r)), v));
        else v = fma(T(1), v, xo);
    } else if (rowscale) {   // Jacobi: gdmv(1, d, v, 0, v), rounding sequence of the stand-alone kernel (kernels.hpp:143-145)
        const T d = __ldg(rowscale + r);
        v = add_rn(mul_rn(T(0), v), mul_rn(mul_rn(T(1), d), v));
    }
    return v;
}

// gather of x: through L1 (stencil matrices: neighbouring lanes and rows share lines) or around it (NA: random columns never hit)
template <bool NA> __device__ __forceinline__ float ldx(const float* p) { return NA ? ldg_stream(p) : __ldg(p); }
template <bool NA> __device__ __forceinline__ double ldx(const double* p) { return NA ? ldg_stream(p) : __ldg(p); }

template <class T, bool HAS_REM, bool NA, int UN>
__global__ void __launch_bounds__(256, UN == 4 ? 5 : (sizeof(T) == 8 ? 6 : 8)) spmv_sell_kernel(int nlanes, int nslices, const int64_t* __restrict__ slice_off, const int* __restrict__ sinds,
                                                         const T* __restrict__ svals, const T* __restrict__ x, T alpha, T beta, const T* y_in,
                                                         T* y_out, float* out32, const T* __restrict__ rowscale, const int* __restrict__ slice_list,
                                                         const int* __restrict__ vout, T* __restrict__ partial, const T* xadd, const __grid_constant__ HaloWait hw,
                                                         const __grid_constant__ PushArgs push, int npush) {
    pdl_trigger_early(nlanes);
    pdl_wait();
    // multi-GPU: the first npush CTAs send this rank's boundary rows of x to the neighbours (straight into the halo tail of THEIR copy of
    // the same basis column) while the slices without halo columns are multiplied; the CTAs of the boundary slices, last in the grid,
    // wait for the neighbours' flags.  The NVLink round trips of the push hide under the interior product instead of sitting at the
    // end of the Arnoldi tail kernel.
    if ((int)blockIdx.x < npush) { halo_push_block<T>(push, (int)blockIdx.x, x, nullptr); return; }
    const int bx = (int)blockIdx.x - npush;
    // CTA-uniform: does this CTA hold a slice position that reads halo columns?
    if (hw.npeers > 0 && (int)((((int64_t)bx + 1) * blockDim.x - 1) >> 5) >= hw.wait_from) halo_wait_block(hw);
    const int ws = (int)(((int64_t)bx * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (ws >= nslices) return;
    const int s = slice_list ? __ldg(slice_list + ws) : ws;
    const int64_t off = __ldg(slice_off + s);
    const int L = (int)((__ldg(slice_off + s + 1) - off) / SLICE);
    const int ng = L / G;
    const int4* ip = reinterpret_cast<const int4*>(sinds + off) + lane;
    const T* vbase = svals + off;
    T sum = T(0);
    int g = 0;
    // UN groups per step: the streaming loads of all of them in flight, then 4 UN gathers, then the row sum in nonzero order
    for (; g + UN <= ng; g += UN) {
        int4 c[UN];
        T v[UN][4];
#pragma unroll
        for (int u = 0; u < UN; ++u) c[u] = ldg_stream(ip + (size_t)(g + u) * SLICE);
#pragma unroll
        for (int u = 0; u < UN; ++u) load_group(vbase + (size_t)(g + u) * SLICE * G, lane, v[u]);
        T xv[UN][4];
#pragma unroll
        for (int u = 0; u < UN; ++u) { xv[u][0] = ldx<NA>(x + c[u].x); xv[u][1] = ldx<NA>(x + c[u].y); xv[u][2] = ldx<NA>(x + c[u].z); xv[u][3] = ldx<NA>(x + c[u].w); }
#pragma unroll
        for (int u = 0; u < UN; ++u)
#pragma unroll
            for (int q = 0; q < 4; ++q) sum = add_rn(sum, mul_rn(v[u][q], xv[u][q]));
    }
    for (; g < ng; ++g) {
        const int4 c0 = ldg_stream(ip + (size_t)g * SLICE);
        T v0[4];
        load_group(vbase + (size_t)g * SLICE * G, lane, v0);
        const T x0 = ldx<NA>(x + c0.x), x1 = ldx<NA>(x + c0.y), x2 = ldx<NA>(x + c0.z), x3 = ldx<NA>(x + c0.w);
        sum = add_rn(sum, mul_rn(v0[0], x0));
        sum = add_rn(sum, mul_rn(v0[1], x1));
        sum = add_rn(sum, mul_rn(v0[2], x2));
        sum = add_rn(sum, mul_rn(v0[3], x3));
    }
    // the L % 4 trailing positions, one element per lane (compiled out for plans whose slices are all whole groups)
    if (HAS_REM) {
        const int* it = sinds + off + (int64_t)ng * SLICE * G + lane;
        const T* vt = svals + off + (int64_t)ng * SLICE * G + lane;
        const int nt = L - ng * G;
        int ct[3];
        T vv[3];
#pragma unroll
        for (int t = 0; t < 3; ++t) if (t < nt) { ct[t] = ldg_stream(it + t * SLICE); vv[t] = ldg_stream(vt + t * SLICE); }
        T xt[3];
#pragma unroll
        for (int t = 0; t < 3; ++t) if (t < nt) xt[t] = ldx<NA>(x + ct[t]);
#pragma unroll
        for (int t = 0; t < 3; ++t) if (t < nt) sum = add_rn(sum, mul_rn(vv[t], xt[t]));
    }
    pdl_trigger();
    const int pos = s * SLICE + lane;
    if (pos >= nlanes) return;
    int r = pos;
    if (vout) {
        r = __ldg(vout + pos);
        if (r < 0) { partial[-1 - r] = sum; return; }   // a piece of a cut row: sell_fixup_kernel finishes it
    }
    const T v = sell_epilogue<T>(r, sum, alpha, beta, y_in, rowscale, xadd);
    if (y_out) y_out[r] = v;
    if (out32) out32[r] = (float)v;
}

// SIGMA mode: one thread per cut row adds its pieces in nonzero order and applies the epilogue
template <class T>
__global__ void sell_fixup_kernel(int nsplit, const int* __restrict__ split_rows, const int* __restrict__ chunk_base, const T* __restrict__ partial,
                                  T alpha, T beta, const T* y_in, T* y_out, float* out32, const T* __restrict__ rowscale, const T* xadd) {
    pdl_trigger();
    pdl_wait();
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nsplit) return;
    const int r = split_rows[j];
    T sum = T(0);
    for (int p = chunk_base[j], pe = chunk_base[j + 1]; p < pe; ++p) sum = add_rn(sum, partial[p]);
    const T v = sell_epilogue<T>(r, sum, alpha, beta, y_in, rowscale, xadd);
    if (y_out) y_out[r] = v;
    if (out32) out32[r] = (float)v;
}

}  // namespace

struct mpg_sell_plan {
    int mode = MODE_PLAIN;
    int nlanes = 0;                // PLAIN: rows; SIGMA: virtual rows
    int nslices = 0;
    int64_t total = 0;             // padded element count
    int64_t* slice_off = nullptr;  // [nslices + 1]
    int* sinds = nullptr;          // [total]
    int* slice_list = nullptr;     // partitioned matrices: slices without halo columns first (SIGMA: longest first inside each part)
    int n_interior = 0;
    int* order = nullptr;          // SIGMA: all slices, longest first (launch order of an undivided product); PLAIN: null = natural order
    int has_rem = 0;               // some slice length is not a multiple of G
    // SIGMA
    int* vstart = nullptr;         // [nlanes] sorted order
    int* vlen = nullptr;
    int* vout = nullptr;
    int nsplit = 0;                // rows cut into pieces
    int nchunks = 0;               // pieces of cut rows
    int* split_rows = nullptr;     // [nsplit]
    int* chunk_base = nullptr;     // [nsplit + 1]
    void* partial = nullptr;       // [nchunks] doubles (either precision)
    unsigned long long uid = 0;    // distinguishes plans that happen to be allocated at the same address
};

struct mpg_packed {
    const mpg_csr* A = nullptr;
    const mpg_sell_plan* plan = nullptr;
    unsigned long long plan_uid = 0;
    int tsize = 0;
    void* svals = nullptr;
    int device = 0;
};

namespace mpg {

void sell_plan_free(mpg_sell_plan* p) {
    if (!p) return;
    pool_free(p->slice_off); pool_free(p->sinds); pool_free(p->slice_list); pool_free(p->order);
    pool_free(p->vstart); pool_free(p->vlen); pool_free(p->vout); pool_free(p->split_rows); pool_free(p->chunk_base); pool_free(p->partial);
    delete p;
}

static LaneSrc lane_src(const mpg_csr* A, const mpg_sell_plan* p) {
    return p->mode == MODE_SIGMA ? LaneSrc{p->nlanes, nullptr, p->vstart, p->vlen} : LaneSrc{A->nrows, A->row_map, nullptr, nullptr};
}

// slice lengths -> slice_off, total, has_rem for the plan's current lane source
static int plan_slices(mpg_ctx* ctx, const mpg_csr* A, mpg_sell_plan* p) {
    p->nslices = (int)cdiv(p->nlanes, SLICE);
    int64_t* len = nullptr;
    MPG_CUDA(ctx, pool_alloc(ctx, &len, sizeof(int64_t) * (size_t)(p->nslices + 2)));
    pool_free(p->slice_off);
    p->slice_off = nullptr;
    MPG_CUDA(ctx, pool_alloc(ctx, &p->slice_off, sizeof(int64_t) * (size_t)(p->nslices + 1)));
    MPG_CUDA(ctx, cudaMemsetAsync(len, 0, sizeof(int64_t) * (size_t)(p->nslices + 2), ctx->stream));
    int* rem_flag = reinterpret_cast<int*>(len + p->nslices + 1);
    sell_len_kernel<<<(int)cdiv((int64_t)p->nslices * 32, 256), 256, 0, ctx->stream>>>(lane_src(A, p), p->nslices, len, rem_flag);
    MPG_CHECK_LAUNCH(ctx);
    scan_kernel<int64_t, int64_t><<<1, 1024, 0, ctx->stream>>>((int64_t)p->nslices, len, p->slice_off);
    MPG_CHECK_LAUNCH(ctx);
    MPG_CUDA(ctx, cudaMemcpyAsync(&p->total, p->slice_off + p->nslices, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
    MPG_CUDA(ctx, cudaMemcpyAsync(&p->has_rem, rem_flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    MPG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    pool_free(len);
    return MPG_OK;
}

// virtual rows, sorted inside windows (see the header comment)
static int plan_sigma(mpg_ctx* ctx, const mpg_csr* A, mpg_sell_plan* p) {
    const int n = A->nrows;
    int *nch = nullptr, *is_split = nullptr, *nch_split = nullptr, *vrow_off = nullptr, *split_off = nullptr, *chunk_off = nullptr;
    int *vstart = nullptr, *vlen = nullptr, *vout = nullptr;
    auto drop = [&]() { pool_free(nch); pool_free(is_split); pool_free(nch_split); pool_free(vrow_off); pool_free(split_off); pool_free(chunk_off);
                        pool_free(vstart); pool_free(vlen); pool_free(vout); };
#define MPG_CUDA_D(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { (void)cudaGetLastError(); drop(); return fail(ctx, MPG_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); } } while (0)
    const size_t nb = sizeof(int) * (size_t)(n + 1);
    MPG_CUDA_D(pool_alloc(ctx, &nch, nb)); MPG_CUDA_D(pool_alloc(ctx, &is_split, nb)); MPG_CUDA_D(pool_alloc(ctx, &nch_split, nb));
    MPG_CUDA_D(pool_alloc(ctx, &vrow_off, nb)); MPG_CUDA_D(pool_alloc(ctx, &split_off, nb)); MPG_CUDA_D(pool_alloc(ctx, &chunk_off, nb));
    vrow_count_kernel<<<(int)cdiv(n, 256), 256, 0, ctx->stream>>>(n, A->row_map, nch, is_split, nch_split);
    scan_kernel<int, int><<<1, 1024, 0, ctx->stream>>>((int64_t)n, nch, vrow_off);
    scan_kernel<int, int><<<1, 1024, 0, ctx->stream>>>((int64_t)n, is_split, split_off);
    scan_kernel<int, int><<<1, 1024, 0, ctx->stream>>>((int64_t)n, nch_split, chunk_off);
    ctx->launches += 4;
    int tot[3] = {0, 0, 0};
    MPG_CUDA_D(cudaMemcpyAsync(&tot[0], vrow_off + n, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    MPG_CUDA_D(cudaMemcpyAsync(&tot[1], split_off + n, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    MPG_CUDA_D(cudaMemcpyAsync(&tot[2], chunk_off + n, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    MPG_CUDA_D(cudaStreamSynchronize(ctx->stream));
    p->nlanes = tot[0]; p->nsplit = tot[1]; p->nchunks = tot[2];
    const size_t vb = sizeof(int) * (size_t)std::max(p->nlanes, 1);
    MPG_CUDA_D(pool_alloc(ctx, &vstart, vb)); MPG_CUDA_D(pool_alloc(ctx, &vlen, vb)); MPG_CUDA_D(pool_alloc(ctx, &vout, vb));
    MPG_CUDA_D(pool_alloc(ctx, &p->vstart, vb)); MPG_CUDA_D(pool_alloc(ctx, &p->vlen, vb)); MPG_CUDA_D(pool_alloc(ctx, &p->vout, vb));
    MPG_CUDA_D(pool_alloc(ctx, &p->split_rows, sizeof(int) * (size_t)std::max(p->nsplit, 1)));
    MPG_CUDA_D(pool_alloc(ctx, &p->chunk_base, sizeof(int) * (size_t)(p->nsplit + 1)));
    MPG_CUDA_D(pool_alloc(ctx, &p->partial, sizeof(double) * (size_t)std::max(p->nchunks, 1)));
    vrow_fill_kernel<<<(int)cdiv(n, 256), 256, 0, ctx->stream>>>(n, A->row_map, vrow_off, split_off, chunk_off, vstart, vlen, vout, p->split_rows, p->chunk_base);
    MPG_CUDA_D(cudaMemcpyAsync(p->chunk_base + p->nsplit, &p->nchunks, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    vrow_sort_kernel<<<(int)cdiv(p->nlanes, SIGMA), 512, 0, ctx->stream>>>(p->nlanes, vstart, vlen, vout, p->vstart, p->vlen, p->vout);
    ctx->launches += 2;
    MPG_CUDA_D(cudaStreamSynchronize(ctx->stream));
    MPG_CUDA_D(cudaGetLastError());
#undef MPG_CUDA_D
    drop();
    p->mode = MODE_SIGMA;
    return MPG_OK;
}

// build (or fetch) the packed structure; *out = nullptr if the matrix does not pack well
int sell_plan_get(mpg_ctx* ctx, const mpg_csr* A, const mpg_sell_plan** out) {
    *out = nullptr;
    mpg_csr* Am = const_cast<mpg_csr*>(A);
    if (Am->sell_tried) { *out = Am->sell; return MPG_OK; }
    Am->sell_tried = 1;
    if (A->nrows == 0 || A->nnz == 0) return MPG_OK;
    static std::atomic<unsigned long long> next_uid{0};   // plans are created from any thread, any context
    mpg_sell_plan* p = new mpg_sell_plan();
    struct Guard { mpg_sell_plan* p; ~Guard() { if (p) sell_plan_free(p); } } guard{p};   // error paths release what was built so far
    p->uid = ++next_uid;
    p->nlanes = A->nrows;
    MPG_TRY(plan_slices(ctx, A, p));
    const auto too_padded = [&]() { return (double)p->total > 1.25 * (double)A->nnz + 4096.0 || p->total >= (int64_t)1 << 40; };
    if (too_padded()) {
        if (!ctx->tune.spmv_sigma) return MPG_OK;
        MPG_TRY(plan_sigma(ctx, A, p));
        MPG_TRY(plan_slices(ctx, A, p));
        if (too_padded()) return MPG_OK;   // keep CSR
    }
    MPG_CUDA(ctx, pool_alloc(ctx, &p->sinds, sizeof(int) * (size_t)std::max<int64_t>(p->total, 1)));
    const bool slab = A->ncols > A->nrows;
    if (slab) MPG_CUDA(ctx, pool_alloc(ctx, &p->slice_list, sizeof(int) * (size_t)p->nslices));
    const int wgrid = (int)cdiv((int64_t)p->nslices * 32, 256);
    sell_fill_inds_kernel<<<wgrid, 256, 0, ctx->stream>>>(lane_src(A, p), p->nslices, A->nrows, A->inds, p->slice_off, p->sinds, p->slice_list);
    MPG_CHECK_LAUNCH(ctx);
    // SIGMA plans choose the LAUNCH ORDER of their slices.  Slice lengths range from 1 to CHUNK positions and a warp walks its slice alone:
    // under load one step of 16 gathers per lane takes ~10 us (the kernel is L1TEX / L2-sector bound, profiles/r02p_ncu_spmv_powerlaw.md),
    // so a CHUNK-long slice lives for 16 such steps whatever the size of the matrix.  In window order the long slices of the last windows
    // start when the grid is nearly drained and the product ends on a tail of a few lonely warps (sm__warps_active 52 % of a theoretical
    // 62 %); on the slab of an 8-GPU run that tail is as long as the whole product.  With M resident warps and W slice-positions of work
    // a slice of length L occupies its warp for a fraction L M / W of the kernel, so it has to START before 1 - L M / W of the grid has
    // been handed out.  sell_lpt = 1 (default): slices whose place in window order misses that deadline (with a 1.5x margin) move to the
    // front, longest first; everything else keeps window order - that keeps long (streaming) and short (latency-bound) slices mixed, which
    // a full sort does not (measured: longest-first over ALL slices gains 3 % in fp32 and LOSES 6 % in fp64 on 8 M rows).  2: full sort.
    // 0: window order.  Results do not depend on the order (every lane owns its row or piece).
    const int lpt = p->mode == MODE_SIGMA ? ctx->tune.sell_lpt : 0;
    std::vector<int64_t> off_h;
    if (lpt) {
        off_h.resize((size_t)p->nslices + 1);
        MPG_CUDA(ctx, cudaMemcpyAsync(off_h.data(), p->slice_off, sizeof(int64_t) * off_h.size(), cudaMemcpyDeviceToHost, ctx->stream));
        MPG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    const auto len_of = [&](int sl) { return (int)((off_h[(size_t)sl + 1] - off_h[(size_t)sl]) / SLICE); };
    // stable counting sort of a list of slices by length, descending
    const auto by_length = [&](int* first, int* last) {
        if (last - first < 2) return;
        int maxl = 0;
        for (int* q = first; q < last; ++q) maxl = std::max(maxl, len_of(*q));
        std::vector<int> start((size_t)maxl + 2, 0), tmp(first, last);
        for (int sl : tmp) ++start[(size_t)(maxl - len_of(sl)) + 1];
        for (size_t b = 1; b < start.size(); ++b) start[b] += start[b - 1];
        for (int sl : tmp) first[start[(size_t)(maxl - len_of(sl))]++] = sl;
    };
    const auto longest_first = [&](int* first, int* last) {
        const int64_t cnt = last - first;
        if (!lpt || cnt < 2) return;
        if (lpt >= 2) { by_length(first, last); return; }
        double work = 0.0;
        for (int* q = first; q < last; ++q) work += len_of(*q);
        if (work <= 0.0) return;
        const double warps = 40.0 * ctx->num_sms;   // resident warps of the UN = 4 kernel (launch bounds: 1280 threads per SM)
        std::vector<int> urgent, rest;
        for (int64_t i = 0; i < cnt; ++i) {
            const int sl = first[i];
            const double deadline = 1.0 - 1.5 * len_of(sl) * warps / work;
            ((double)i / (double)cnt > deadline ? urgent : rest).push_back(sl);
        }
        by_length(urgent.data(), urgent.data() + urgent.size());
        std::copy(urgent.begin(), urgent.end(), first);
        std::copy(rest.begin(), rest.end(), first + urgent.size());
    };
    if (slab) {
        // local slab of a partitioned matrix: order the slices [no halo column | some halo column]
        std::vector<int> flag((size_t)p->nslices), list((size_t)p->nslices);
        MPG_CUDA(ctx, cudaMemcpyAsync(flag.data(), p->slice_list, sizeof(int) * flag.size(), cudaMemcpyDeviceToHost, ctx->stream));
        MPG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        int ni = 0;
        for (int s = 0; s < p->nslices; ++s) if (!flag[(size_t)s]) list[(size_t)ni++] = s;
        p->n_interior = ni;
        for (int s = 0; s < p->nslices; ++s) if (flag[(size_t)s]) list[(size_t)ni++] = s;
        longest_first(list.data(), list.data() + p->n_interior);
        longest_first(list.data() + p->n_interior, list.data() + p->nslices);
        MPG_CUDA(ctx, cudaMemcpyAsync(p->slice_list, list.data(), sizeof(int) * list.size(), cudaMemcpyHostToDevice, ctx->stream));
        MPG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    if (lpt) {
        std::vector<int> all((size_t)p->nslices);
        for (int s = 0; s < p->nslices; ++s) all[(size_t)s] = s;
        longest_first(all.data(), all.data() + p->nslices);
        MPG_CUDA(ctx, pool_alloc(ctx, &p->order, sizeof(int) * all.size()));
        MPG_CUDA(ctx, cudaMemcpyAsync(p->order, all.data(), sizeof(int) * all.size(), cudaMemcpyHostToDevice, ctx->stream));
        MPG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    guard.p = nullptr;
    Am->sell = p;
    *out = p;
    return MPG_OK;
}

template <class T>
int pack_update(mpg_ctx* ctx, mpg_packed* P, const T* vals) {
    const mpg_sell_plan* p = P->plan;
    const mpg_csr* A = P->A;
    ProfScope prof(ctx, MPG_PROF_ELEMENTWISE, (double)A->nnz * sizeof(T) + (double)p->total * sizeof(T));
    sell_fill_vals_kernel<T><<<(int)cdiv((int64_t)p->nslices * 32, 256), 256, 0, ctx->stream>>>(lane_src(A, p), p->nslices, vals, p->slice_off,
                                                                                             static_cast<T*>(P->svals));
    MPG_CHECK_LAUNCH(ctx);
    return MPG_OK;
}
template int pack_update<float>(mpg_ctx*, mpg_packed*, const float*);
template int pack_update<double>(mpg_ctx*, mpg_packed*, const double*);

// *out = nullptr (and MPG_OK) when the structure does not pack well
template <class T>
int pack_create(mpg_ctx* ctx, const mpg_csr* A, const T* vals, mpg_packed** out) {
    *out = nullptr;
    const mpg_sell_plan* plan = nullptr;
    MPG_TRY(sell_plan_get(ctx, A, &plan));
    if (!plan) return MPG_OK;
    mpg_packed* P = new mpg_packed();
    P->A = A; P->plan = plan; P->plan_uid = plan->uid; P->tsize = (int)sizeof(T); P->device = ctx->device;
    cudaError_t e = pool_alloc(ctx, &P->svals, sizeof(T) * (size_t)std::max<int64_t>(plan->total, 1));
    if (e != cudaSuccess) { (void)cudaGetLastError(); delete P; return fail(ctx, MPG_ERR_CUDA, std::string("pack_create: ") + cudaGetErrorString(e)); }
    *out = P;
    return pack_update<T>(ctx, P, vals);
}
template int pack_create<float>(mpg_ctx*, const mpg_csr*, const float*, mpg_packed**);
template int pack_create<double>(mpg_ctx*, const mpg_csr*, const double*, mpg_packed**);

void pack_free(mpg_packed* P) {
    if (!P) return;
    pool_free(P->svals);
    delete P;
}

// slices of a partitioned slab that reference halo columns (0 without a slice list)
int pack_boundary_slices(const mpg_packed* P) { return (P && P->plan && P->plan->slice_list) ? P->plan->nslices - P->plan->n_interior : 0; }

bool pack_matches(const mpg_packed* P, const mpg_csr* A, int tsize) {
    const mpg_sell_plan* cur = A->sell;
    return P && P->A == A && P->tsize == tsize && cur && cur == P->plan && cur->uid == P->plan_uid;
}

template <class T>
int spmv_packed(mpg_ctx* ctx, const mpg_packed* P, T alpha, const T* x, T beta, const T* y_in, T* y_out, float* out32, const T* rowscale, int part,
                const HaloWait* hw_in, const T* xadd, const PushArgs* push_in) {
    const mpg_sell_plan* p = P->plan;
    const mpg_csr* A = P->A;
    if (part != SPMV_ALL && !p->slice_list) {
        if (part == SPMV_INTERIOR) return MPG_OK;
        part = SPMV_ALL;
    }
    const int s_first = part == SPMV_BOUNDARY ? p->n_interior : 0;
    const int s_count = part == SPMV_INTERIOR ? p->n_interior : p->nslices - s_first;
    const int* list = part == SPMV_ALL ? p->order : p->slice_list + s_first;   // ORDERED: the whole list, interior slices first; ALL: longest first (SIGMA) or natural order
    // algorithmic bytes: the CSR figure of SURVEY.md §8d (padding and per-lane destinations the packed layout reads on top are not counted)
    const double n_ = A->nrows, s_ = sizeof(T);
    const double bytes = (double)A->nnz * (s_ + 4) + 4 * (n_ + 1) + n_ * s_ + (y_out ? n_ * s_ : 0) + (beta != T(0) ? n_ * s_ : 0) + (out32 ? 4 * n_ : 0) +
                         (rowscale ? n_ * s_ : 0);
    ProfScope prof(ctx, sizeof(T) == 4 ? MPG_PROF_SPMV_F32 : MPG_PROF_SPMV_F64, bytes * ((double)s_count / p->nslices));
    HaloWait hw;
    hw.npeers = 0;
    if (hw_in) hw = *hw_in;
    hw.wait_from = (part == SPMV_ORDERED) ? p->n_interior : 0;
    if (s_count > 0) {
        // measured (profiles/r02_tune_sell_sigma.txt): SIGMA plans (random columns) want 4 groups = 16 gathers in flight per lane, stencils 2;
        // x always through L1 - the no-allocate hint on the gathers also demotes x in L2 and doubles the DRAM traffic
        const int variant = ctx->tune.sell_variant >= 0 ? ctx->tune.sell_variant : (p->mode == MODE_SIGMA ? 2 : 0);
        const int block = ctx->tune.sell_block > 0 ? std::min(ctx->tune.sell_block, 256) : 128;
        PushArgs pa;
        if (push_in) pa = *push_in;
        const int npush = push_in ? pa.npeers * pa.bpp : 0;
        const int grid = (int)cdiv((int64_t)s_count * 32, block) + npush;
        void (*kern)(int, int, const int64_t*, const int*, const T*, const T*, T, T, const T*, T*, float*, const T*, const int*, const int*, T*, const T*, const HaloWait,
                     const PushArgs, int);
        switch ((variant & 3) * 2 + (p->has_rem ? 1 : 0)) {
            case 0: kern = spmv_sell_kernel<T, false, false, 2>; break;
            case 1: kern = spmv_sell_kernel<T, true, false, 2>; break;
            case 2: kern = spmv_sell_kernel<T, false, true, 2>; break;
            case 3: kern = spmv_sell_kernel<T, true, true, 2>; break;
            case 4: kern = spmv_sell_kernel<T, false, false, 4>; break;
            case 5: kern = spmv_sell_kernel<T, true, false, 4>; break;
            case 6: kern = spmv_sell_kernel<T, false, true, 4>; break;
            default: kern = spmv_sell_kernel<T, true, true, 4>; break;
        }
        MPG_CUDA(ctx, launch_pdl(ctx, (int64_t)A->nrows, kern, grid, block, 0, p->nlanes, s_count, (const int64_t*)p->slice_off, (const int*)p->sinds,
                                 static_cast<const T*>(P->svals), x, alpha, beta, y_in, y_out, out32, rowscale, list, (const int*)p->vout,
                                 static_cast<T*>(p->partial), xadd, hw, pa, npush));
        MPG_CHECK_LAUNCH(ctx);
    }
    if (p->nsplit > 0 && part != SPMV_INTERIOR) {
        MPG_CUDA(ctx, launch_pdl(ctx, (int64_t)A->nrows, sell_fixup_kernel<T>, (int)cdiv(p->nsplit, 128), 128, 0, p->nsplit, (const int*)p->split_rows,
                                 (const int*)p->chunk_base, static_cast<const T*>(p->partial), alpha, beta, y_in, y_out, out32, rowscale, xadd));
        MPG_CHECK_LAUNCH(ctx);
    }
    return MPG_OK;
}
template int spmv_packed<float>(mpg_ctx*, const mpg_packed*, float, const float*, float, const float*, float*, float*, const float*, int, const HaloWait*, const float*, const PushArgs*);
template int spmv_packed<double>(mpg_ctx*, const mpg_packed*, double, const double*, double, const double*, double*, float*, const double*, int, const HaloWait*, const double*, const PushArgs*);

int scan_i32(mpg_ctx* ctx, int64_t n, const int* in, int* out) {
    scan_kernel<int, int><<<1, 1024, 0, ctx->stream>>>(n, in, out);
    MPG_CHECK_LAUNCH(ctx);
    return MPG_OK;
}

}  // namespace mpg

// ---- C ABI -------------------------------------------------------------------------------------------------
#define MPG_DEF_PACK(SFX, T)                                                                                                       \
    extern "C" int mpg_pack_create_##SFX(mpg_ctx* ctx, const mpg_csr* A, const T* vals, mpg_packed** out) {                       \
        MPG_REQUIRE(ctx, A && out && (vals || A->nnz == 0), "pack_create: null argument");                                        \
        return mpg::pack_create<T>(ctx, A, vals, out);                                                                            \
    }                                                                                                                              \
    extern "C" int mpg_pack_update_##SFX(mpg_ctx* ctx, mpg_packed* P, const T* vals) {                                            \
        MPG_REQUIRE(ctx, P && vals && P->tsize == (int)sizeof(T), "pack_update: null argument or wrong precision");               \
        return mpg::pack_update<T>(ctx, P, vals);                                                                                 \
    }                                                                                                                              \
    extern "C" int mpg_spmv_packed_##SFX(mpg_ctx* ctx, const mpg_packed* P, T alpha, const T* x, T beta, T* y) {                  \
        MPG_REQUIRE(ctx, P && x && y && P->tsize == (int)sizeof(T), "spmv_packed: null argument or wrong precision");             \
        return mpg::spmv_packed<T>(ctx, P, alpha, x, beta, y, y, nullptr, nullptr, mpg::SPMV_ALL, nullptr, nullptr, nullptr);              \
    }
MPG_DEF_PACK(f32, float)
MPG_DEF_PACK(f64, double)

extern "C" int mpg_pack_describe(const mpg_packed* P, int* group, int* nslices, int64_t* total, const int64_t** slice_off, const int** inds,
                                 const void** vals) {
    if (!P || !P->plan) return MPG_ERR_ARG;
    if (group) *group = G;
    if (nslices) *nslices = P->plan->nslices;
    if (total) *total = P->plan->total;
    if (slice_off) *slice_off = P->plan->slice_off;
    if (inds) *inds = P->plan->sinds;
    if (vals) *vals = P->svals;
    return MPG_OK;
}

extern "C" int mpg_pack_describe_rows(const mpg_packed* P, int* mode, int* chunk, int* sigma, int* nlanes, const int** lane_start, const int** lane_len,
                                      const int** lane_out, int* nsplit, int* nchunks, const int** split_rows, const int** chunk_base) {
    if (!P || !P->plan) return MPG_ERR_ARG;
    const mpg_sell_plan* p = P->plan;
    if (mode) *mode = p->mode;
    if (chunk) *chunk = CHUNK;
    if (sigma) *sigma = SIGMA;
    if (nlanes) *nlanes = p->nlanes;
    if (lane_start) *lane_start = p->vstart;
    if (lane_len) *lane_len = p->vlen;
    if (lane_out) *lane_out = p->vout;
    if (nsplit) *nsplit = p->nsplit;
    if (nchunks) *nchunks = p->nchunks;
    if (split_rows) *split_rows = p->split_rows;
    if (chunk_base) *chunk_base = p->chunk_base;
    return MPG_OK;
}

extern "C" int mpg_pack_destroy(mpg_packed* P) {
    if (!P) return MPG_OK;
    cudaSetDevice(P->device);
    mpg::pack_free(P);
    return MPG_OK;
}
