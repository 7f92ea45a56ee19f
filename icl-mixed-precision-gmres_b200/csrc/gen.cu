// gen.cu — device generators for the synthetic inputs of SURVEY.md §8d (the reference ships none).
// The definitions are frozen in oracle/oracle.cpp; tests compare these bit-exactly against it.
// All CSR in LoadMatrix-canonical form (LoadMatrix.hpp:62-145): ascending columns, diagonal present.
#include <vector>

#include "common.cuh"

using namespace mpg;

namespace {

// number of valid offsets {-1,0,1} at coordinate t of an N-grid, and the prefix sum over t' < t
__device__ __forceinline__ int64_t cnt1(int64_t t, int64_t N) { return 1 + (t > 0) + (t < N - 1); }
__device__ __forceinline__ int64_t pre1(int64_t t, int64_t N) { return t + max(t - 1, (int64_t)0) + min(t, N - 1); }

__global__ void lap2d_kernel(int64_t N, int* row_map, int* inds, double* vals) {
    const int64_t n = N * N;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    if (i == n) { row_map[n] = (int)(5 * N * N - 4 * N); return; }
    const int64_t x = i % N, y = i / N;
    // entries before grid row y, then before column x inside it
    const int64_t vy_pre = max(y - 1, (int64_t)0) + min(y, N - 1);              // sum_{y'<y} ([y'>0] + [y'<N-1])
    const int64_t rows_before = y * (N + 2 * (N - 1)) + N * vy_pre;
    const int64_t vy = (y > 0) + (y < N - 1);
    const int64_t in_row = x * (1 + vy) + max(x - 1, (int64_t)0) + min(x, N - 1);
    int64_t p = rows_before + in_row;
    row_map[i] = (int)p;
    if (y > 0) { inds[p] = (int)(i - N); vals[p++] = -1.0; }
    if (x > 0) { inds[p] = (int)(i - 1); vals[p++] = -1.0; }
    inds[p] = (int)i; vals[p++] = 4.0;
    if (x < N - 1) { inds[p] = (int)(i + 1); vals[p++] = -1.0; }
    if (y < N - 1) { inds[p] = (int)(i + N); vals[p++] = -1.0; }
}

// 27-point convection-diffusion: diag 26, off-diagonals -1, plus convection +-c on the six face neighbours
// (oracle/oracle.cpp CD27_*).
__global__ void cd27_kernel(int64_t N, int* row_map, int* inds, double* vals) {
    const int64_t n = N * N * N;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    const int64_t S = 3 * N - 2;
    if (i == n) { row_map[n] = (int)(S * S * S); return; }
    const int64_t x = i % N, y = (i / N) % N, z = i / (N * N);
    int64_t p = pre1(z, N) * S * S + cnt1(z, N) * (pre1(y, N) * S + cnt1(y, N) * pre1(x, N));
    row_map[i] = (int)p;
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

__global__ void powerlaw_len_kernel(int64_t n, uint64_t seed, int lmin, int gmax, int* lens) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) lens[i] = (int)(pl_rowlen(seed, i, n, lmin, gmax) + 1);
}

// one warp per row: lanes stride over the off-diagonal slots; the diagonal position is the number of
// off-diagonal columns below i (columns are ascending in the slot index, so it is a ballot/count).
__global__ void powerlaw_fill_kernel(int64_t n, uint64_t seed, int lmin, int gmax, const int* __restrict__ row_map, int* inds, double* vals) {
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (i >= n) return;
    const int64_t len = pl_rowlen(seed, i, n, lmin, gmax);
    const int64_t p0 = row_map[i];
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
    lap2d_kernel<<<(int)cdiv(N * N + 1, 256), 256, 0, ctx->stream>>>(N, row_map, inds, vals);
    MPG_CHECK_LAUNCH(ctx);
    return MPG_OK;
}
extern "C" int mpg_gen_cd27(mpg_ctx* ctx, int64_t N, int* row_map, int* inds, double* vals) {
    MPG_REQUIRE(ctx, N >= 1 && mpg_cd27_nnz(N) < 2147483647LL, "gen_cd27: N out of range");
    cd27_kernel<<<(int)cdiv(N * N * N + 1, 256), 256, 0, ctx->stream>>>(N, row_map, inds, vals);
    MPG_CHECK_LAUNCH(ctx);
    return MPG_OK;
}
extern "C" int mpg_gen_powerlaw_rowmap(mpg_ctx* ctx, int64_t n, uint64_t seed, int lmin, int gmax, int* row_map, int64_t* nnz_host) {
    MPG_REQUIRE(ctx, n >= 2 && lmin >= 1 && gmax >= 0 && gmax < 40, "gen_powerlaw: bad parameters");
    // row lengths on the device, exclusive scan on the host (setup code, 4 bytes per row each way)
    powerlaw_len_kernel<<<(int)cdiv(n, 256), 256, 0, ctx->stream>>>(n, seed, lmin, gmax, row_map + 1);
    MPG_CHECK_LAUNCH(ctx);
    std::vector<int> h(n + 1);
    MPG_CUDA(ctx, cudaMemcpyAsync(h.data() + 1, row_map + 1, sizeof(int) * n, cudaMemcpyDeviceToHost, ctx->stream));
    MPG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    int64_t p = 0;
    h[0] = 0;
    for (int64_t i = 1; i <= n; ++i) {
        p += h[i];
        if (p >= 2147483647LL) return fail(ctx, MPG_ERR_ARG, "gen_powerlaw: nnz overflows int32");
        h[i] = (int)p;
    }
    MPG_CUDA(ctx, cudaMemcpyAsync(row_map, h.data(), sizeof(int) * (n + 1), cudaMemcpyHostToDevice, ctx->stream));
    MPG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (nnz_host) *nnz_host = p;
    return MPG_OK;
}
extern "C" int mpg_gen_powerlaw_fill(mpg_ctx* ctx, int64_t n, uint64_t seed, int lmin, int gmax, const int* row_map, int* inds, double* vals) {
    MPG_REQUIRE(ctx, n >= 2, "gen_powerlaw: bad n");
    powerlaw_fill_kernel<<<(int)cdiv(n * 32, 256), 256, 0, ctx->stream>>>(n, seed, lmin, gmax, row_map, inds, vals);
    MPG_CHECK_LAUNCH(ctx);
    return MPG_OK;
}

// ---- 1-D row partition (host) -----------------------------------------------------------------------------
#include <algorithm>
extern "C" int mpg_partition_bounds(int64_t n, int P, int64_t* bounds) {
    if (!bounds || P < 1 || n < 0) return MPG_ERR_ARG;
    for (int r = 0; r <= P; ++r) bounds[r] = (int64_t)(((__int128)r * n) / P);
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
