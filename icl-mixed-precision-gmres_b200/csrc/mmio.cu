// mmio.cu — MatrixMarket ingest into LoadMatrix-canonical CSR (host code).  SURVEY.md §8f-2 ("next" row).
// Restates LoadMatrix<S>() of the reference (LoadMatrix.hpp:17-154) without its quadratic parts:
//   * coordinate, real|integer, general|symmetric only (LoadMatrix.hpp:48-54); anything else is an error;
//   * every row gets a diagonal entry, explicit 0 if the file has none (:62-66,94-101); a file diagonal overwrites it,
//     the last one wins (:110-111);
//   * symmetric files are mirrored (:79-82,118-124); duplicates are NOT merged;
//   * columns ascending within a row; the reference's bubble sort (:128-145) is stable, so is the sort used here:
//     equal columns keep insertion order (diagonal slot first, then file order with mirrored entries interleaved as read).
// The reference parses with fscanf and sorts each row in O(len^2); here the file is read in one block, parsed with
// strtol/strtod, and rows are sorted with std::stable_sort — same result, usable at 10^8 nonzeros.
#include <algorithm>
#include <cctype>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <string>
#include <vector>

#include "common.cuh"

namespace {
struct Entry { int col; int seq; double val; };
}

extern "C" int mpg_mm_read_host(const char* path, int* nrows_out, int* ncols_out, int64_t* nnz_out, int** row_map_out, int** inds_out,
                                double** vals_out, char* errbuf, int errlen) {
    auto fail_msg = [&](const char* m) {
        if (errbuf && errlen > 0) { std::snprintf(errbuf, (size_t)errlen, "%s", m); }
        return MPG_ERR_ARG;
    };
    if (!path || !nrows_out || !ncols_out || !nnz_out || !row_map_out || !inds_out || !vals_out) return fail_msg("null argument");
    FILE* f = std::fopen(path, "rb");
    if (!f) return fail_msg("Could not access file");                                  // LoadMatrix.hpp:22-25
    std::fseek(f, 0, SEEK_END);
    const long sz = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    std::vector<char> buf((size_t)sz + 1);
    const size_t got = std::fread(buf.data(), 1, (size_t)sz, f);
    std::fclose(f);
    buf[got] = 0;
    char* p = buf.data();
    // ---- banner (mmio.c mm_read_banner semantics: "%%MatrixMarket matrix <format> <field> <symmetry>") ----
    char* eol = std::strchr(p, '\n');
    if (!eol) return fail_msg("Missing values in banner");
    std::string banner(p, eol);
    for (auto& c : banner) c = (char)std::tolower((unsigned char)c);
    char b0[64], b1[64], b2[64], b3[64], b4[64];
    if (std::sscanf(banner.c_str(), "%63s %63s %63s %63s %63s", b0, b1, b2, b3, b4) != 5) return fail_msg("Missing values in banner");
    if (std::string(b0) != "%%matrixmarket") return fail_msg("Banner is missing");
    if (std::string(b1) != "matrix") return fail_msg("Unrecognized description");
    const bool coordinate = std::string(b2) == "coordinate";
    const bool real_or_int = std::string(b3) == "real" || std::string(b3) == "integer";
    const bool general = std::string(b4) == "general", symmetric = std::string(b4) == "symmetric";
    if (!(coordinate && real_or_int && (general || symmetric))) return fail_msg("Unsupported matrix type");   // :48-54
    p = eol + 1;
    while (*p == '%') { eol = std::strchr(p, '\n'); if (!eol) return fail_msg("Malformed matrix size information"); p = eol + 1; }
    char* q;
    const long M = std::strtol(p, &q, 10); p = q;
    const long N = std::strtol(p, &q, 10); p = q;
    const long nz = std::strtol(p, &q, 10);
    if (q == p || M <= 0 || N <= 0 || nz < 0) return fail_msg("Malformed matrix size information");
    p = q;
    // ---- entries: count, then place ----
    std::vector<int> I((size_t)nz), J((size_t)nz);
    std::vector<double> V((size_t)nz);
    std::vector<int64_t> cnt((size_t)N + 1, 1);   // one diagonal per row (:62-64)
    cnt[0] = 0;
    for (long e = 0; e < nz; ++e) {
        const long i = std::strtol(p, &q, 10); if (q == p) return fail_msg("premature end of entries"); p = q;
        const long j = std::strtol(p, &q, 10); if (q == p) return fail_msg("premature end of entries"); p = q;
        const double v = std::strtod(p, &q); if (q == p) return fail_msg("premature end of entries"); p = q;
        if (i < 1 || i > N || j < 1 || j > N) return fail_msg("entry index out of range");
        I[(size_t)e] = (int)(i - 1); J[(size_t)e] = (int)(j - 1); V[(size_t)e] = v;
        if (i != j) { cnt[(size_t)i] += 1; if (symmetric) cnt[(size_t)j] += 1; }
    }
    for (long r = 0; r < N; ++r) cnt[(size_t)r + 1] += cnt[(size_t)r];
    const int64_t nnz = cnt[(size_t)N];
    if (nnz >= 2147483647LL) return fail_msg("nnz overflows int32 (types_cuda.hpp:66-70)");
    int* row_map = (int*)std::malloc(sizeof(int) * ((size_t)N + 1));
    int* inds = (int*)std::malloc(sizeof(int) * (size_t)std::max<int64_t>(nnz, 1));
    double* vals = (double*)std::malloc(sizeof(double) * (size_t)std::max<int64_t>(nnz, 1));
    if (!row_map || !inds || !vals) { std::free(row_map); std::free(inds); std::free(vals); return fail_msg("out of memory"); }
    for (long r = 0; r <= N; ++r) row_map[r] = (int)cnt[(size_t)r];
    std::vector<int> fill((size_t)N, 1);
    for (long r = 0; r < N; ++r) { inds[row_map[r]] = (int)r; vals[row_map[r]] = 0.0; }   // base diagonal (:94-101)
    for (long e = 0; e < nz; ++e) {
        const int row = I[(size_t)e], col = J[(size_t)e];
        const double v = V[(size_t)e];
        if (row == col) { vals[row_map[row]] = v; continue; }                              // :110-111
        int k = fill[(size_t)row]++;
        inds[row_map[row] + k] = col; vals[row_map[row] + k] = v;
        if (symmetric) { k = fill[(size_t)col]++; inds[row_map[col] + k] = row; vals[row_map[col] + k] = v; }
    }
    // ---- stable sort of every row by column (:128-145) ----
    std::vector<Entry> tmp;
    for (long r = 0; r < N; ++r) {
        const int s = row_map[r], len = row_map[r + 1] - s;
        bool sorted = true;
        for (int k = 1; k < len && sorted; ++k) sorted = inds[s + k - 1] <= inds[s + k];
        if (sorted) continue;
        tmp.resize((size_t)len);
        for (int k = 0; k < len; ++k) tmp[(size_t)k] = {inds[s + k], k, vals[s + k]};
        std::stable_sort(tmp.begin(), tmp.end(), [](const Entry& a, const Entry& b) { return a.col < b.col; });
        for (int k = 0; k < len; ++k) { inds[s + k] = tmp[(size_t)k].col; vals[s + k] = tmp[(size_t)k].val; }
    }
    *nrows_out = (int)M; *ncols_out = (int)N; *nnz_out = nnz;
    *row_map_out = row_map; *inds_out = inds; *vals_out = vals;
    return MPG_OK;
}

// LoadVector<S>(file, col), LoadMatrix.hpp:156-233: column `col` of a MatrixMarket ARRAY file (dense, column-major) or of a
// COORDINATE file (entries of other columns are skipped, missing entries are 0, the last duplicate wins).  Same exception
// texts as the reference; the size line is read like mm_read_mtx_array_size / mm_read_mtx_crd_size (comment lines skipped).
extern "C" int mpg_mm_read_vector_host(const char* path, int col, int64_t* n_out, double** vals_out, char* errbuf, int errlen) {
    auto fail_msg = [&](const std::string& m) {
        if (errbuf && errlen > 0) { std::snprintf(errbuf, (size_t)errlen, "%s", m.c_str()); }
        return MPG_ERR_ARG;
    };
    if (!path || !n_out || !vals_out || col < 0) return fail_msg("null argument");
    FILE* f = std::fopen(path, "rb");
    if (!f) return fail_msg("Could not access file");                                  // :158-161
    std::fseek(f, 0, SEEK_END);
    const long sz = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    std::vector<char> buf((size_t)sz + 1);
    const size_t got = std::fread(buf.data(), 1, (size_t)sz, f);
    std::fclose(f);
    buf[got] = 0;
    char* p = buf.data();
    char* eol = std::strchr(p, '\n');
    if (!eol) return fail_msg("Missing values in banner");                             // :165-177
    std::string banner(p, eol);
    for (auto& c : banner) c = (char)std::tolower((unsigned char)c);
    char b0[64], b1[64], b2[64], b3[64], b4[64];
    if (std::sscanf(banner.c_str(), "%63s %63s %63s %63s %63s", b0, b1, b2, b3, b4) != 5) return fail_msg("Missing values in banner");
    if (std::string(b0) != "%%matrixmarket") return fail_msg("Banner is missing");
    if (std::string(b1) != "matrix") return fail_msg("Unrecognized description");
    const bool array = std::string(b2) == "array", coordinate = std::string(b2) == "coordinate";
    if (!array && !coordinate) return fail_msg("Unrecognized description");
    p = eol + 1;
    while (*p == '%') { eol = std::strchr(p, '\n'); if (!eol) return fail_msg("Malformed matrix size information"); p = eol + 1; }
    char* q;
    const long M = std::strtol(p, &q, 10); if (q == p) return fail_msg("Malformed matrix size information"); p = q;   // :181-190
    const long N = std::strtol(p, &q, 10); if (q == p) return fail_msg("Malformed matrix size information"); p = q;
    long nz = 0;
    if (coordinate) { nz = std::strtol(p, &q, 10); if (q == p) return fail_msg("Malformed matrix size information"); p = q; }
    if (M < 0 || N < 0 || nz < 0) return fail_msg("Malformed matrix size information");
    if (col >= N) return fail_msg("Column " + std::to_string(col) + " is too large for the " + std::to_string(N) + " vectors");   // :192-197
    double* vals = (double*)std::calloc((size_t)std::max<long>(M, 1), sizeof(double));
    if (!vals) return fail_msg("out of memory");
    if (array) {                                                                       // :202-214
        for (long i = 0; i < (long)col * M; ++i) { std::strtod(p, &q); if (q == p) { std::free(vals); return fail_msg("premature end of entries"); } p = q; }
        for (long j = 0; j < M; ++j) { vals[j] = std::strtod(p, &q); if (q == p) { std::free(vals); return fail_msg("premature end of entries"); } p = q; }
    } else {                                                                           // :215-227
        for (long e = 0; e < nz; ++e) {
            const long i = std::strtol(p, &q, 10); if (q == p) { std::free(vals); return fail_msg("premature end of entries"); } p = q;
            const long j = std::strtol(p, &q, 10); if (q == p) { std::free(vals); return fail_msg("premature end of entries"); } p = q;
            const double v = std::strtod(p, &q); if (q == p) { std::free(vals); return fail_msg("premature end of entries"); } p = q;
            if (i < 1 || i > M) { std::free(vals); return fail_msg("entry index out of range"); }
            if (j - 1 == col) vals[i - 1] = v;
        }
    }
    *n_out = M;
    *vals_out = vals;
    return MPG_OK;
}

// ---- partition-aware ingest ("a partition-aware scatter to P GPUs") -------------------------------------------------
// Rows [lo, hi) of the canonical CSR of a MatrixMarket file WITHOUT ever holding the other rows: what one rank of a multi-GPU run reads.
// The file is streamed in blocks (a coordinate file lists entries in any order, so every rank scans all of it, but keeps only what lands
// in its rows: memory = O(entries of the slab) + 4 (n + 1) bytes for the optional global row map).  The result equals rows [lo, hi) of
// mpg_mm_read_host bit for bit (tests/test_loader_cpu.py): a row holds its diagonal slot first, then its entries in file order with the
// mirrored entries of a symmetric file interleaved as read, then the stable sort by column - all of which is decided per row.
namespace {
// whitespace-separated tokens of a FILE read in blocks; a block is cut at its last white-space character so that no token straddles two
// blocks (the rest is carried over), and a NUL at the cut stops strtol / strtod
struct TokenStream {
    FILE* f;
    std::vector<char> buf;
    size_t len = 0, limit = 0, pos = 0;
    char saved = 0;
    bool eof = false;
    explicit TokenStream(FILE* file, size_t cap) : f(file), buf(cap + 1) {}
    bool refill() {
        if (eof && limit == len) return false;
        buf[limit] = saved;
        const size_t tail = len - pos;
        std::memmove(buf.data(), buf.data() + pos, tail);
        const size_t want = buf.size() - 1 - tail;
        const size_t got = want ? std::fread(buf.data() + tail, 1, want, f) : 0;
        if (got < want) eof = true;
        len = tail + got;
        pos = 0;
        limit = len;
        if (!eof) {
            while (limit > 0 && !std::isspace((unsigned char)buf[limit - 1])) --limit;
            if (limit == 0) return false;   // one token longer than a block: not a MatrixMarket file
        }
        saved = buf[limit];
        buf[limit] = 0;
        return true;
    }
    // next token through conv (strtol / strtod); false at the end of the file or on a token that does not convert
    template <class Conv> bool next(Conv conv) {
        for (;;) {
            while (pos < limit && std::isspace((unsigned char)buf[pos])) ++pos;
            if (pos < limit) {
                char* p = buf.data() + pos;
                char* q = p;
                conv(p, &q);
                if (q == p) return false;
                pos = (size_t)(q - buf.data());
                return true;
            }
            if (!refill()) return false;
        }
    }
    bool next_long(long& v) { return next([&](char* p, char** q) { v = std::strtol(p, q, 10); }); }
    bool next_double(double& v) { return next([&](char* p, char** q) { v = std::strtod(p, q); }); }
};
}  // namespace

extern "C" int mpg_mm_read_slab_host(const char* path, int64_t lo, int64_t hi, int* nrows_out, int* ncols_out, int64_t* nnz_global_out,
                                     int64_t* nnz_local_out, int** row_map_local_out, int** inds_out, double** vals_out, int** row_map_global_out,
                                     char* errbuf, int errlen) {
    auto fail_msg = [&](const char* m) {
        if (errbuf && errlen > 0) { std::snprintf(errbuf, (size_t)errlen, "%s", m); }
        return MPG_ERR_ARG;
    };
    if (!path || !nrows_out || !ncols_out || !nnz_global_out || !nnz_local_out || !row_map_local_out || !inds_out || !vals_out) return fail_msg("null argument");
    FILE* f = std::fopen(path, "rb");
    if (!f) return fail_msg("Could not access file");                                  // LoadMatrix.hpp:22-25
    struct Closer { FILE* f; ~Closer() { std::fclose(f); } } closer{f};
    // ---- banner and comment lines, line by line (the same checks and texts as mpg_mm_read_host) ----
    std::string line;
    auto read_line = [&]() {
        line.clear();
        int c;
        while ((c = std::fgetc(f)) != EOF) { if (c == '\n') return true; line.push_back((char)c); }
        return false;   // no newline before the end of the file
    };
    if (!read_line()) return fail_msg("Missing values in banner");
    for (auto& c : line) c = (char)std::tolower((unsigned char)c);
    char b0[64], b1[64], b2[64], b3[64], b4[64];
    if (std::sscanf(line.c_str(), "%63s %63s %63s %63s %63s", b0, b1, b2, b3, b4) != 5) return fail_msg("Missing values in banner");
    if (std::string(b0) != "%%matrixmarket") return fail_msg("Banner is missing");
    if (std::string(b1) != "matrix") return fail_msg("Unrecognized description");
    const bool coordinate = std::string(b2) == "coordinate";
    const bool real_or_int = std::string(b3) == "real" || std::string(b3) == "integer";
    const bool general = std::string(b4) == "general", symmetric = std::string(b4) == "symmetric";
    if (!(coordinate && real_or_int && (general || symmetric))) return fail_msg("Unsupported matrix type");   // :48-54
    for (;;) {
        const int c = std::fgetc(f);
        if (c == EOF) return fail_msg("Malformed matrix size information");
        if (c != '%') { std::ungetc(c, f); break; }
        if (!read_line()) return fail_msg("Malformed matrix size information");
    }
    TokenStream ts(f, (size_t)1 << 24);
    long M = 0, N = 0, nz = -1;
    if (!ts.next_long(M) || !ts.next_long(N) || !ts.next_long(nz) || M <= 0 || N <= 0 || nz < 0) return fail_msg("Malformed matrix size information");
    if (hi < 0) hi = N;
    if (lo < 0 || lo > hi || hi > N) return fail_msg("row range outside the matrix");
    const long nl = (long)(hi - lo);
    // ---- one pass: keep what lands in rows [lo, hi); count every row when the global row map is wanted ----
    struct Kept { int row_local; int col; double val; };
    std::vector<Kept> kept;
    std::vector<double> diag((size_t)std::max<long>(nl, 1), 0.0);   // base diagonal 0 (:94-101); a file diagonal overwrites it, the last one wins (:110-111)
    std::vector<int64_t> cnt_local((size_t)nl + 1, 1);
    cnt_local[0] = 0;
    std::vector<int64_t> cnt_global;
    if (row_map_global_out) { cnt_global.assign((size_t)N + 1, 1); cnt_global[0] = 0; }
    int64_t nnz_global = N;
    for (long e = 0; e < nz; ++e) {
        long i, j;
        double v;
        if (!ts.next_long(i) || !ts.next_long(j) || !ts.next_double(v)) return fail_msg("premature end of entries");
        if (i < 1 || i > N || j < 1 || j > N) return fail_msg("entry index out of range");
        const long r = i - 1, c = j - 1;
        if (r == c) { if (r >= lo && r < hi) diag[(size_t)(r - lo)] = v; continue; }
        nnz_global += symmetric ? 2 : 1;
        if (row_map_global_out) { cnt_global[(size_t)r + 1] += 1; if (symmetric) cnt_global[(size_t)c + 1] += 1; }
        if (r >= lo && r < hi) { kept.push_back({(int)(r - lo), (int)c, v}); cnt_local[(size_t)(r - lo) + 1] += 1; }
        if (symmetric && c >= lo && c < hi) { kept.push_back({(int)(c - lo), (int)r, v}); cnt_local[(size_t)(c - lo) + 1] += 1; }
    }
    if (nnz_global >= 2147483647LL) return fail_msg("nnz overflows int32 (types_cuda.hpp:66-70)");
    for (long r = 0; r < nl; ++r) cnt_local[(size_t)r + 1] += cnt_local[(size_t)r];
    const int64_t nnz_local = cnt_local[(size_t)nl];
    int* row_map = (int*)std::malloc(sizeof(int) * ((size_t)nl + 1));
    int* inds = (int*)std::malloc(sizeof(int) * (size_t)std::max<int64_t>(nnz_local, 1));
    double* vals = (double*)std::malloc(sizeof(double) * (size_t)std::max<int64_t>(nnz_local, 1));
    int* row_map_global = row_map_global_out ? (int*)std::malloc(sizeof(int) * ((size_t)N + 1)) : nullptr;
    if (!row_map || !inds || !vals || (row_map_global_out && !row_map_global)) {
        std::free(row_map); std::free(inds); std::free(vals); std::free(row_map_global);
        return fail_msg("out of memory");
    }
    for (long r = 0; r <= nl; ++r) row_map[r] = (int)cnt_local[(size_t)r];
    std::vector<int> fill((size_t)std::max<long>(nl, 1), 1);
    for (long r = 0; r < nl; ++r) { inds[row_map[r]] = (int)(lo + r); vals[row_map[r]] = diag[(size_t)r]; }
    for (const Kept& k : kept) {
        const int at = row_map[k.row_local] + fill[(size_t)k.row_local]++;
        inds[at] = k.col; vals[at] = k.val;
    }
    std::vector<Entry> tmp;
    for (long r = 0; r < nl; ++r) {   // stable sort of every row by column (:128-145)
        const int s = row_map[r], len = row_map[r + 1] - s;
        bool sorted = true;
        for (int k = 1; k < len && sorted; ++k) sorted = inds[s + k - 1] <= inds[s + k];
        if (sorted) continue;
        tmp.resize((size_t)len);
        for (int k = 0; k < len; ++k) tmp[(size_t)k] = {inds[s + k], k, vals[s + k]};
        std::stable_sort(tmp.begin(), tmp.end(), [](const Entry& a, const Entry& b) { return a.col < b.col; });
        for (int k = 0; k < len; ++k) { inds[s + k] = tmp[(size_t)k].col; vals[s + k] = tmp[(size_t)k].val; }
    }
    if (row_map_global_out) {
        for (long r = 0; r < N; ++r) cnt_global[(size_t)r + 1] += cnt_global[(size_t)r];
        for (long r = 0; r <= N; ++r) row_map_global[r] = (int)cnt_global[(size_t)r];
        *row_map_global_out = row_map_global;
    }
    *nrows_out = (int)M; *ncols_out = (int)N; *nnz_global_out = nnz_global; *nnz_local_out = nnz_local;
    *row_map_local_out = row_map; *inds_out = inds; *vals_out = vals;
    return MPG_OK;
}

extern "C" void mpg_host_free(void* p) { std::free(p); }
