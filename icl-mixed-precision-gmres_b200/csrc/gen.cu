// gen.cu — device generators for the synthetic inputs of SURVEY.md §8d (the reference ships none).
// The definitions are frozen in oracle/oracle.cpp; tests compare these bit-exactly against it.
// All CSR in LoadMatrix-canonical form (LoadMatrix.hpp:62-145): ascending columns, diagonal present.
#include <vector>

#include "common.cuh"

using namespace mpg;

namespace {

// CSR offset of the first entry of grid row i (i == n: the total), closed forms - every rank can generate ANY row range
__host__ __device__ __forceinline__ int64_t lap2d_offset(int64_t i, int64_t N) {
    if (i >= N * N) return 5 * N * N - 4 * N;
    const int64_t x = i % N, y = i / N;
    const int64_t vy_pre = (y - 1 > 0 ? y - 1 : 0) + (y < N - 1 ? y : N - 1);    // sum_{y'<y} ([y'>0] + [y'<N-1])
    const int64_t rows_before = y * (N + 2 * (N - 1)) + N * vy_pre;
    const int64_t vy = (y > 0) + (y < N - 1);
    return rows_before + x * (1 + vy) + (x - 1 > 0 ? x - 1 : 0) + (x < N - 1 ? x : N - 1);
}
__host__ __device__ __forceinline__ int64_t cd27_pre1(int64_t t, int64_t N) { return t + (t - 1 > 0 ? t - 1 : 0) + (t < N - 1 ? t : N - 1); }
__host__ __device__ __forceinline__ int64_t cd27_offset(int64_t i, int64_t N) {
    const int64_t S = 3 * N - 2;
    if (i >= N * N * N) return S * S * S;
    const int64_t x = i % N, y = (i / N) % N, z = i / (N * N);
    const int64_t cz = 1 + (z > 0) + (z < N - 1), cy = 1 + (y > 0) + (y < N - 1);
    return cd27_pre1(z, N) * S * S + cz * (cd27_pre1(y, N) * S + cy * cd27_pre1(x, N));
}

// rows [lo, hi): row_map[i - lo] = offset(i) - offset(lo); entries with GLOBAL column indices
__global__ void lap2d_kernel(int64_t N, int64_t lo, int64_t hi, int* row_map, int* inds, double* vals) {
    const int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > hi) return;
    const int64_t base = lap2d_offset(lo, N);
    int64_t p = lap2d_offset(i, N) - base;
    row_map[i - lo] = (int)p;
    if (i == hi || !inds) return;
    const int64_t x = i % N, y = i / N;
    if (y > 0) { inds[p] = (int)(i - N); vals[p++] = -1.0; }
    if (x > 0) { inds[p] = (int)(i - 1); vals[p++] = -1.0; }
    inds[p] = (int)i; vals[p++] = 4.0;
    if (x < N - 1) { inds[p] = (int)(i + 1); vals[p++] = -1.0; }
    if (y < N - 1) { inds[p] = (int)(i + N); vals[p++] = -1.0; }
}

// 27-point convection-diffusion: diag 26, off-diagonals -1, plus convection +-c on the six face neighbours
// (oracle/oracle.cpp CD27_*).
__global__ void cd27_kernel(int64_t N, int64_t lo, int64_t hi, int* row_map, int* inds, double* vals) {
    const int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > hi) return;
    const int64_t base = cd27_offset(lo, N);
    int64_t p = cd27_offset(i, N) - base;
    row_map[i - lo] = (int)p;
    if (i == hi || !inds) return;
    const int64_t x = i % N, y = (i / N) % N, z = i / (N * N);
    for (int dz = -1; dz <= 1; ++dz)
        for (int dy = -1; dy <= 1; ++dy)
            for (int dx = -1; dx <= 1; ++dx) {
                const int64_t xx = x + dx, yy = y + dy, zz = z + dz;
                if (xx < 0 || xx >= N || yy < 0 || yy >= N || zz < 0 || zz >= N) continue;
                double v = -1.0;
                const int na = (dx != 0) + (dy != 0) + (dz != 0);
                if (na == 0) v = 26.0;
                else if (na == 1) v += dx * 0.5 + dy * 0.25 + dz * 0.125;
                inds[p] = (int)(xx + N * (yy + N * zz));
                vals[p++] = v;
            }
}

__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ uint64_t pl_hash(uint64_t seed, uint64_t row, uint64_t slot) {
    return splitmix64(splitmix64(seed ^ (row * 0xD1342543DE82EF95ull)) + slot);
}
__host__ __device__ __forceinline__ int64_t pl_rowlen(uint64_t seed, int64_t i, int64_t n, int lmin, int gmax) {
    const uint64_t hsh = pl_hash(seed, (uint64_t)i, 0);
    int g = 0;
    while (g < gmax && !((hsh >> (63 - g)) & 1ull)) ++g;
    const int64_t base = (int64_t)lmin << g;
    const int64_t frac = (int64_t)(hsh & 0xFFFFull);
    int64_t len = base + ((base * frac) >> 16);
    if (len > n - 1) len = n - 1;
    return len;
}

__global__ void powerlaw_len_kernel(int64_t n, int64_t lo, int64_t hi, uint64_t seed, int lmin, int gmax, int* lens) {
    const int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < hi) lens[i - lo] = (int)(pl_rowlen(seed, i, n, lmin, gmax) + 1);
}

// one warp per row: lanes stride over the off-diagonal slots; the diagonal position is the number of
// off-diagonal columns below i (columns are ascending in the slot index, so it is a ballot/count).
__global__ void powerlaw_fill_kernel(int64_t n, int64_t lo, int64_t hi, uint64_t seed, int lmin, int gmax, const int* __restrict__ row_map, int* inds,
                                     double* vals) {
    const int64_t i = lo + (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (i >= hi) return;
    const int64_t len = pl_rowlen(seed, i, n, lmin, gmax);
    const int64_t p0 = row_map[i - lo];
    int64_t below = 0;   // off-diagonal columns < i handled by this lane
    int64_t abs64 = 0;   // sum |v| * 64, exact integer
    for (int64_t s = lane; s < len; s += 32) {
        const int64_t lo = (s * (n - 1)) / len, hi = ((s + 1) * (n - 1)) / len;   // < 2^62: s < 2^31, n < 2^31
        int64_t c = lo + (int64_t)(pl_hash(seed, (uint64_t)i, (uint64_t)(2 * s + 1)) % (uint64_t)(hi - lo));
        c += (c >= i) ? 1 : 0;
        int kq = (int)(pl_hash(seed, (uint64_t)i, (uint64_t)(2 * s + 2)) % 127u) - 63;
        if (kq == 0) kq = 1;
        const int64_t pos = p0 + s + ((c > i) ? 1 : 0);
        inds[pos] = (int)c;
        vals[pos] = (double)kq / 64.0;
        below += (c < i) ? 1 : 0;
        abs64 += (kq < 0) ? -kq : kq;
    }
    for (int o = 16; o > 0; o >>= 1) {
        below += __shfl_xor_sync(0xffffffffu, below, o);
        abs64 += __shfl_xor_sync(0xffffffffu, abs64, o);
    }
    if (lane == 0) {
        inds[p0 + below] = (int)i;
        vals[p0 + below] = 1.0 + (double)abs64 / 64.0;
    }
}

}  // namespace

extern "C" int64_t mpg_lap2d_nnz(int64_t N) { return 5 * N * N - 4 * N; }
extern "C" int64_t mpg_cd27_nnz(int64_t N) { const int64_t t = 3 * N - 2; return t * t * t; }

extern "C" int mpg_gen_lap2d(mpg_ctx* ctx, int64_t N, int* row_map, int* inds, double* vals) {
    MPG_REQUIRE(ctx, N >= 1 && mpg_lap2d_nnz(N) < 2147483647LL, "gen_lap2d: N out of range");
    lap2d_kernel<<<(int)cdiv(N * N + 1, 256), 256, 0, ctx->stream>>>(N, 0, N * N, row_map, inds, vals);
    MPG_CHECK_LAUNCH(ctx);
    return MPG_OK;
}
extern "C" int mpg_gen_cd27(mpg_ctx* ctx, int64_t N, int* row_map, int* inds, double* vals) {
    MPG_REQUIRE(ctx, N >= 1 && mpg_cd27_nnz(N) < 2147483647LL, "gen_cd27: N out of range");
    cd27_kernel<<<(int)cdiv(N * N * N + 1, 256), 256, 0, ctx->stream>>>(N, 0, N * N * N, row_map, inds, vals);
    MPG_CHECK_LAUNCH(ctx);
    return MPG_OK;
}
namespace {
// lens[0 .. hi-lo) on the device -> exclusive scan on the host (setup code, 4 bytes per row each way) -> row_map
int powerlaw_rowmap_range(mpg_ctx* ctx, int64_t n, int64_t lo, int64_t hi, uint64_t seed, int lmin, int gmax, int* row_map, int64_t* nnz_host) {
    const int64_t m = hi - lo;
    if (m > 0) {
        powerlaw_len_kernel<<<(int)cdiv(m, 256), 256, 0, ctx->stream>>>(n, lo, hi, seed, lmin, gmax, row_map + 1);
        MPG_CHECK_LAUNCH(ctx);
    }
    std::vector<int> h((size_t)m + 1);
    if (m > 0) MPG_CUDA(ctx, cudaMemcpyAsync(h.data() + 1, row_map + 1, sizeof(int) * (size_t)m, cudaMemcpyDeviceToHost, ctx->stream));
    MPG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    int64_t p = 0;
    h[0] = 0;
    for (int64_t i = 1; i <= m; ++i) {
        p += h[(size_t)i];
        if (p >= 2147483647LL) return fail(ctx, MPG_ERR_ARG, "gen_powerlaw: nnz overflows int32");
        h[(size_t)i] = (int)p;
    }
    MPG_CUDA(ctx, cudaMemcpyAsync(row_map, h.data(), sizeof(int) * (size_t)(m + 1), cudaMemcpyHostToDevice, ctx->stream));
    MPG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (nnz_host) *nnz_host = p;
    return MPG_OK;
}
}  // namespace
extern "C" int mpg_gen_powerlaw_rowmap(mpg_ctx* ctx, int64_t n, uint64_t seed, int lmin, int gmax, int* row_map, int64_t* nnz_host) {
    MPG_REQUIRE(ctx, n >= 2 && lmin >= 1 && gmax >= 0 && gmax < 40, "gen_powerlaw: bad parameters");
    return powerlaw_rowmap_range(ctx, n, 0, n, seed, lmin, gmax, row_map, nnz_host);
}
extern "C" int mpg_gen_powerlaw_fill(mpg_ctx* ctx, int64_t n, uint64_t seed, int lmin, int gmax, const int* row_map, int* inds, double* vals) {
    MPG_REQUIRE(ctx, n >= 2, "gen_powerlaw: bad n");
    powerlaw_fill_kernel<<<(int)cdiv(n * 32, 256), 256, 0, ctx->stream>>>(n, 0, n, seed, lmin, gmax, row_map, inds, vals);
    MPG_CHECK_LAUNCH(ctx);
    return MPG_OK;
}

// ---- row-range generators: rows [lo, hi) of the same matrices, GLOBAL column indices, local row map (first entry 0).
// Every rank of a multi-GPU run builds only its slab (mpg_dist_setup renumbers the columns); nobody holds the global matrix.
// kind: 0 lap2d (size = N), 1 cd27 (size = N), 2 powerlaw (size = n rows; seed, lmin, gmax as mpg_gen_powerlaw_*).
extern "C" int mpg_gen_slab_rowmap(mpg_ctx* ctx, int kind, int64_t size, uint64_t seed, int lmin, int gmax, int64_t lo, int64_t hi, int* row_map_local,
                                   int64_t* nnz_local_host) {
    MPG_REQUIRE(ctx, kind >= 0 && kind <= 2 && size >= 1 && row_map_local, "gen_slab: bad argument");
    const int64_t n = kind == 0 ? size * size : (kind == 1 ? size * size * size : size);
    MPG_REQUIRE(ctx, 0 <= lo && lo <= hi && hi <= n, "gen_slab: row range out of bounds");
    if (kind == 2) {
        MPG_REQUIRE(ctx, n >= 2 && lmin >= 1 && gmax >= 0 && gmax < 40, "gen_powerlaw: bad parameters");
        return powerlaw_rowmap_range(ctx, n, lo, hi, seed, lmin, gmax, row_map_local, nnz_local_host);
    }
    const int64_t nnz = kind == 0 ? lap2d_offset(hi, size) - lap2d_offset(lo, size) : cd27_offset(hi, size) - cd27_offset(lo, size);
    MPG_REQUIRE(ctx, nnz < 2147483647LL, "gen_slab: local nnz overflows int32");
    if (kind == 0) lap2d_kernel<<<(int)cdiv(hi - lo + 1, 256), 256, 0, ctx->stream>>>(size, lo, hi, row_map_local, nullptr, nullptr);
    else cd27_kernel<<<(int)cdiv(hi - lo + 1, 256), 256, 0, ctx->stream>>>(size, lo, hi, row_map_local, nullptr, nullptr);
    MPG_CHECK_LAUNCH(ctx);
    if (nnz_local_host) *nnz_local_host = nnz;
    return MPG_OK;
}
extern "C" int mpg_gen_slab_fill(mpg_ctx* ctx, int kind, int64_t size, uint64_t seed, int lmin, int gmax, int64_t lo, int64_t hi, int* row_map_local,
                                 int* inds_global, double* vals) {
    MPG_REQUIRE(ctx, kind >= 0 && kind <= 2 && size >= 1 && row_map_local, "gen_slab: bad argument");
    const int64_t n = kind == 0 ? size * size : (kind == 1 ? size * size * size : size);
    MPG_REQUIRE(ctx, 0 <= lo && lo <= hi && hi <= n, "gen_slab: row range out of bounds");
    if (hi == lo) return MPG_OK;   // an empty slab has no entries (null arrays are fine)
    MPG_REQUIRE(ctx, inds_global && vals, "gen_slab: null arrays");
    if (kind == 0) lap2d_kernel<<<(int)cdiv(hi - lo + 1, 256), 256, 0, ctx->stream>>>(size, lo, hi, row_map_local, inds_global, vals);
    else if (kind == 1) cd27_kernel<<<(int)cdiv(hi - lo + 1, 256), 256, 0, ctx->stream>>>(size, lo, hi, row_map_local, inds_global, vals);
    else powerlaw_fill_kernel<<<(int)cdiv((hi - lo) * 32, 256), 256, 0, ctx->stream>>>(n, lo, hi, seed, lmin, gmax, row_map_local, inds_global, vals);
    MPG_CHECK_LAUNCH(ctx);
    return MPG_OK;
}
// global row map only (4 (n + 1) bytes): what nnz-balanced split points need
extern "C" int mpg_gen_rowmap(mpg_ctx* ctx, int kind, int64_t size, uint64_t seed, int lmin, int gmax, int* row_map) {
    const int64_t n = kind == 0 ? size * size : (kind == 1 ? size * size * size : size);
    return mpg_gen_slab_rowmap(ctx, kind, size, seed, lmin, gmax, 0, n, row_map, nullptr);
}

// ---- 1-D row partition (host) -----------------------------------------------------------------------------
#include <algorithm>
extern "C" int mpg_partition_bounds(int64_t n, int P, int64_t* bounds) {
    if (!bounds || P < 1 || n < 0) return MPG_ERR_ARG;
    for (int r = 0; r <= P; ++r) bounds[r] = (int64_t)(((__int128)r * n) / P);
    return MPG_OK;
}
extern "C" int mpg_partition_bounds_nnz(int64_t n, int P, const int* row_map, int64_t* bounds) {
    if (!bounds || !row_map || P < 1 || n < 0) return MPG_ERR_ARG;
    const int64_t nnz = row_map[n];
    for (int k = 0; k <= P; ++k) {
        const int64_t target = (int64_t)(((__int128)k * nnz) / P);
        bounds[k] = std::lower_bound(row_map, row_map + n + 1, target, [](int a, int64_t t) { return (int64_t)a < t; }) - row_map;
    }
    bounds[0] = 0;
    bounds[P] = n;
    for (int k = 1; k <= P; ++k) bounds[k] = std::max(bounds[k], bounds[k - 1]);
    return MPG_OK;
}
extern "C" int mpg_partition_local(int64_t n, int P, int r, const int* row_map, const int* inds, int64_t* n_halo, int64_t* halo_cols,
                                   int* local_inds) {
    if (!row_map || !inds || !n_halo || P < 1 || r < 0 || r >= P) return MPG_ERR_ARG;
    const int64_t lo = (int64_t)(((__int128)r * n) / P), hi = (int64_t)(((__int128)(r + 1) * n) / P);
    const int64_t p0 = row_map[lo], p1 = row_map[hi];
    // remote columns: mark-and-compact over the global column range keeps this O(nnz_local + n) with no sort
    std::vector<unsigned char> mark((size_t)n, 0);
    for (int64_t p = p0; p < p1; ++p) {
        const int64_t c = inds[p];
        if (c < lo || c >= hi) mark[(size_t)c] = 1;
    }
    std::vector<int> rank_of;  // halo rank of each marked global column
    if (local_inds) rank_of.assign((size_t)n, -1);
    int64_t nh = 0;
    for (int64_t c = 0; c < n; ++c)
        if (mark[(size_t)c]) {
            if (halo_cols) halo_cols[nh] = c;
            if (local_inds) rank_of[(size_t)c] = (int)nh;
            ++nh;
        }
    *n_halo = nh;
    if (local_inds) {
        const int64_t nl = hi - lo;
        for (int64_t p = p0; p < p1; ++p) {
            const int64_t c = inds[p];
            local_inds[p - p0] = (c >= lo && c < hi) ? (int)(c - lo) : (int)(nl + rank_of[(size_t)c]);
        }
    }
    return MPG_OK;
}
