// mmio.cu — MatrixMarket ingest into LoadMatrix-canonical CSR (host code).  SURVEY.md §8f-2 ("next" row).
// Restates LoadMatrix<S>() of the reference (LoadMatrix.hpp:17-154) without its quadratic parts:
//   * coordinate, real|integer, general|symmetric only (LoadMatrix.hpp:48-54); anything else is an error;
//   * every row gets a diagonal entry, explicit 0 if the file has none (:62-66,94-101); a file diagonal overwrites it,
//     the last one wins (:110-111);
//   * symmetric files are mirrored (:79-82,118-124); duplicates are NOT merged;
//   * columns ascending within a row; the reference's bubble sort (:128-145) is stable, so is the sort used here:
//     equal columns keep insertion order (diagonal slot first, then file order with mirrored entries interleaved as read).
// The reference parses with fscanf and sorts each row in O(len^2); here the entry stream is converted by all host threads over byte
// ranges (strtol / strtod per token, parse_block), rows are counted, filled and stable-sorted by the thread that owns them — same
// result, ~9 M entries / s on 8 cores (one thread: ~3 M / s), usable at 10^8 nonzeros.  Environment: MPG_LOADER_THREADS,
// MPG_LOADER_BLOCK (bytes of text per block of the slab reader) - for tests.  A token that strtol / strtod does not consume entirely is
// an error ("premature end of entries"), where fscanf would re-read its rest as the next field.
#include <algorithm>
#include <cctype>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <climits>
#include <string>
#include <thread>
#include <vector>

#include <sched.h>

#include "common.cuh"

namespace {
struct Entry { int col; int seq; double val; };

// ---- parallel parsing ("needs a fast parallel loader"; the reference's fscanf loop does ~1 M entries / s) ------------
int loader_threads(int64_t work_items) {
    if (const char* e = std::getenv("MPG_LOADER_THREADS")) { const int v = std::atoi(e); if (v > 0) return std::min(v, 64); }
    int c = 0;
    cpu_set_t set;
    CPU_ZERO(&set);
    if (sched_getaffinity(0, sizeof(set), &set) == 0) c = CPU_COUNT(&set);
    if (c <= 0) { const unsigned h = std::thread::hardware_concurrency(); c = h ? (int)h : 1; }
    c = std::min(c, 32);
    return (int)std::max<int64_t>(1, std::min<int64_t>(c, work_items / 65536));   // one thread per 64 K items at least
}
template <class F> void run_threads(int nt, F f) {
    if (nt <= 1) { f(0); return; }
    std::vector<std::thread> th;
    for (int t = 1; t < nt; ++t) th.emplace_back([&f, t]() { f(t); });
    f(0);
    for (auto& x : th) x.join();
}

enum { PARSE_OK = 0, PARSE_PREMATURE = 1, PARSE_RANGE = 2 };
// state that travels from one block of the file to the next: how many tokens of the entry stream were consumed (entry = token / 3, field
// = token % 3) and the fields of an entry whose tokens straddle two blocks
struct ParseCarry { int64_t tokens = 0; long i = 0, j = 0; };
struct ParsedEntries { std::vector<int> I, J; std::vector<double> V; size_t count = 0; };   // complete entries, file order, 0-based indices

// Entries "i j v" of buf[0, limit): buf[limit] == 0 and the block ends on white space or at the end of the file, so no token is torn.
// A token is a maximal run of non-white-space characters; the k-th token of the entry stream is field k % 3 of entry k / 3 (the
// reference's fscanf("%d %d %lg") reads the same stream, LoadMatrix.hpp:68-84).  Threads take byte ranges; a token belongs to the range
// its first character lies in; pass 1 counts tokens per range, pass 2 converts them in place of their entry.  Tokens beyond 3 * nz are
// ignored like the reference ignores what follows the last entry.  A token that strtol / strtod does not consume entirely, or an index
// outside [1, N], is an error; the FIRST one in file order is reported (what a sequential reader would hit first).
int parse_block(char* buf, size_t limit, int64_t max_tokens, long N, ParseCarry& carry, ParsedEntries& out) {
    const int nt = loader_threads((int64_t)(limit / 24));
    const auto is_sp = [](char c) { return c == ' ' || (c >= '\t' && c <= '\r'); };   // isspace() of the C locale, inlined
    std::vector<size_t> lo((size_t)nt + 1);
    for (int t = 0; t <= nt; ++t) lo[(size_t)t] = limit / (size_t)nt * (size_t)t;
    lo[(size_t)nt] = limit;
    std::vector<int64_t> ntok((size_t)nt + 1, 0);
    run_threads(nt, [&](int t) {
        int64_t c = 0;
        for (size_t k = lo[(size_t)t]; k < lo[(size_t)t + 1]; ++k) c += !is_sp(buf[k]) && (k == 0 || is_sp(buf[k - 1]));
        ntok[(size_t)t + 1] = c;
    });
    for (int t = 0; t < nt; ++t) ntok[(size_t)t + 1] += ntok[(size_t)t];
    const int64_t first = carry.tokens;                                  // stream index of this block's first token
    const int64_t use = std::max<int64_t>(0, std::min<int64_t>(ntok[(size_t)nt], max_tokens - first));
    const int64_t e_first = first / 3, e_last = (first + use) / 3;       // complete entries end before e_last
    const size_t nent = (size_t)(e_last - e_first) + 1;                  // + a possibly incomplete last one
    std::vector<long> Il(nent, 0), Jl(nent, 0);
    std::vector<double> Vl(nent, 0.0);
    if (first % 3 >= 1) Il[0] = carry.i;
    if (first % 3 == 2) Jl[0] = carry.j;
    std::vector<int64_t> bad_conv((size_t)nt, INT64_MAX), bad_range((size_t)nt, INT64_MAX);   // stream index of the first offending token per range
    run_threads(nt, [&](int t) {
        int64_t k = ntok[(size_t)t];                                     // block-local index of the next token of this range
        for (size_t b = lo[(size_t)t]; b < lo[(size_t)t + 1] && k < use; ++b) {
            if (is_sp(buf[b]) || !(b == 0 || is_sp(buf[b - 1]))) continue;
            const int64_t g = first + k;
            const size_t e = (size_t)(g / 3 - e_first);
            char* q = buf + b;
            bool in_range = true;
            if (g % 3 == 2) { Vl[e] = std::strtod(buf + b, &q); }
            else {
                const long v = std::strtol(buf + b, &q, 10);
                (g % 3 == 0 ? Il[e] : Jl[e]) = v;
                in_range = v >= 1 && v <= N;
            }
            if (q == buf + b || !(*q == 0 || is_sp(*q))) { if (bad_conv[(size_t)t] == INT64_MAX) bad_conv[(size_t)t] = g; }
            else if (!in_range && bad_range[(size_t)t] == INT64_MAX) bad_range[(size_t)t] = g;
            ++k;
        }
    });
    {
        // the sequential reader converts the three fields of an entry, then checks its indices: entries in file order, and inside one
        // entry a conversion error comes before a range error
        int64_t conv = INT64_MAX, range = INT64_MAX;
        for (int t = 0; t < nt; ++t) { conv = std::min(conv, bad_conv[(size_t)t]); range = std::min(range, bad_range[(size_t)t]); }
        if (range != INT64_MAX && (conv == INT64_MAX || range / 3 < conv / 3)) return PARSE_RANGE;
        if (conv != INT64_MAX) return PARSE_PREMATURE;
    }
    out.count = (size_t)(e_last - e_first);
    out.I.resize(out.count); out.J.resize(out.count); out.V.resize(out.count);
    run_threads(nt, [&](int t) {
        const size_t a = out.count / (size_t)nt * (size_t)t, b = t == nt - 1 ? out.count : out.count / (size_t)nt * (size_t)(t + 1);
        for (size_t e = a; e < b; ++e) { out.I[e] = (int)(Il[e] - 1); out.J[e] = (int)(Jl[e] - 1); out.V[e] = Vl[e]; }
    });
    carry.tokens = first + use;
    if (carry.tokens % 3 >= 1) carry.i = Il[nent - 1];
    if (carry.tokens % 3 == 2) carry.j = Jl[nent - 1];
    return PARSE_OK;
}

// stable sort of every row in [r0, r1) by column (LoadMatrix.hpp:128-145: the reference's bubble sort is stable)
void sort_rows(const int* row_map, int* inds, double* vals, long r0, long r1) {
    std::vector<Entry> tmp;
    for (long r = r0; r < r1; ++r) {
        const int s = row_map[r], len = row_map[r + 1] - s;
        bool sorted = true;
        for (int k = 1; k < len && sorted; ++k) sorted = inds[s + k - 1] <= inds[s + k];
        if (sorted) continue;
        tmp.resize((size_t)len);
        for (int k = 0; k < len; ++k) tmp[(size_t)k] = {inds[s + k], k, vals[s + k]};
        std::stable_sort(tmp.begin(), tmp.end(), [](const Entry& a, const Entry& b) { return a.col < b.col; });
        for (int k = 0; k < len; ++k) { inds[s + k] = tmp[(size_t)k].col; vals[s + k] = tmp[(size_t)k].val; }
    }
}
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
    // ---- entries: parsed by all host threads, then counted, placed and sorted by the thread that OWNS the row (each thread scans the
    // entry list in file order and acts on its rows only: per-row insertion order = file order, no races, same bits as one thread) ----
    ParsedEntries ent;
    {
        ParseCarry carry;
        const size_t limit = got;   // buf[got] == 0
        const int pk = parse_block(p, limit - (size_t)(p - buf.data()), 3 * (int64_t)nz, N, carry, ent);
        if (pk == PARSE_RANGE) return fail_msg("entry index out of range");
        if (pk != PARSE_OK || ent.count < (size_t)nz) return fail_msg("premature end of entries");
    }
    const int* I = ent.I.data();
    const int* J = ent.J.data();
    const double* V = ent.V.data();
    const int nt = loader_threads((int64_t)nz / 4);
    const auto own_lo = [&](int t) { return (long)((int64_t)N * t / nt); };
    std::vector<int64_t> cnt((size_t)N + 1, 1);   // one diagonal per row (:62-64)
    cnt[0] = 0;
    run_threads(nt, [&](int t) {
        const int r0 = (int)own_lo(t), r1 = (int)own_lo(t + 1);
        for (long e = 0; e < nz; ++e) {
            const int i = I[e], j = J[e];
            if (i == j) continue;
            if (i >= r0 && i < r1) cnt[(size_t)i + 1] += 1;
            if (symmetric && j >= r0 && j < r1) cnt[(size_t)j + 1] += 1;
        }
    });
    for (long r = 0; r < N; ++r) cnt[(size_t)r + 1] += cnt[(size_t)r];
    const int64_t nnz = cnt[(size_t)N];
    if (nnz >= 2147483647LL) return fail_msg("nnz overflows int32 (types_cuda.hpp:66-70)");
    int* row_map = (int*)std::malloc(sizeof(int) * ((size_t)N + 1));
    int* inds = (int*)std::malloc(sizeof(int) * (size_t)std::max<int64_t>(nnz, 1));
    double* vals = (double*)std::malloc(sizeof(double) * (size_t)std::max<int64_t>(nnz, 1));
    if (!row_map || !inds || !vals) { std::free(row_map); std::free(inds); std::free(vals); return fail_msg("out of memory"); }
    for (long r = 0; r <= N; ++r) row_map[r] = (int)cnt[(size_t)r];
    std::vector<int> fill((size_t)N, 1);
    run_threads(nt, [&](int t) {
        const int r0 = (int)own_lo(t), r1 = (int)own_lo(t + 1);
        for (int r = r0; r < r1; ++r) { inds[row_map[r]] = r; vals[row_map[r]] = 0.0; }   // base diagonal (:94-101)
        for (long e = 0; e < nz; ++e) {
            const int row = I[e], col = J[e];
            if (row == col) { if (row >= r0 && row < r1) vals[row_map[row]] = V[e]; continue; }   // :110-111, the last one wins
            if (row >= r0 && row < r1) { const int k = fill[(size_t)row]++; inds[row_map[row] + k] = col; vals[row_map[row] + k] = V[e]; }
            if (symmetric && col >= r0 && col < r1) { const int k = fill[(size_t)col]++; inds[row_map[col] + k] = row; vals[row_map[col] + k] = V[e]; }
        }
        sort_rows(row_map, inds, vals, r0, r1);   // :128-145
    });
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
    size_t block = (size_t)1 << 26;   // 64 MiB of text per block: every host thread gets a few MB to convert
    if (const char* e = std::getenv("MPG_LOADER_BLOCK")) { const long v = std::atol(e); if (v >= 256) block = (size_t)v; }
    TokenStream ts(f, block);
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
    ParseCarry carry;
    ParsedEntries ent;
    for (;;) {
        if (ts.limit > ts.pos && carry.tokens < 3 * (int64_t)nz) {
            // this block: converted by all host threads, then walked once in file order
            const int pk = parse_block(ts.buf.data() + ts.pos, ts.limit - ts.pos, 3 * (int64_t)nz, N, carry, ent);
            if (pk == PARSE_RANGE) return fail_msg("entry index out of range");
            if (pk != PARSE_OK) return fail_msg("premature end of entries");
            for (size_t e = 0; e < ent.count; ++e) {
                const long r = ent.I[e], c = ent.J[e];
                const double v = ent.V[e];
                if (r == c) { if (r >= lo && r < hi) diag[(size_t)(r - lo)] = v; continue; }
                nnz_global += symmetric ? 2 : 1;
                if (row_map_global_out) { cnt_global[(size_t)r + 1] += 1; if (symmetric) cnt_global[(size_t)c + 1] += 1; }
                if (r >= lo && r < hi) { kept.push_back({(int)(r - lo), (int)c, v}); cnt_local[(size_t)(r - lo) + 1] += 1; }
                if (symmetric && c >= lo && c < hi) { kept.push_back({(int)(c - lo), (int)r, v}); cnt_local[(size_t)(c - lo) + 1] += 1; }
            }
        }
        ts.pos = ts.limit;
        if (carry.tokens >= 3 * (int64_t)nz || !ts.refill()) break;
    }
    if (carry.tokens < 3 * (int64_t)nz) return fail_msg("premature end of entries");
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
    {
        const int nts = loader_threads(nnz_local / 4);
        run_threads(nts, [&](int t) { sort_rows(row_map, inds, vals, (long)((int64_t)nl * t / nts), (long)((int64_t)nl * (t + 1) / nts)); });   // :128-145
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
