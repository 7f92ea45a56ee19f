// host-side probe (GPU box): how fast can the host cores cast an fp64 value array to fp32 (candidate for overlapping the fp32
// operator's transfer with the fp64 one in mpg_gmres_solve_host)?  gcc -O3 -march=native -fopenmp hostcast.c -o hostcast
#include <omp.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
static double now() { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }
int main(int argc, char** argv) {
    size_t n = argc > 1 ? strtoull(argv[1], 0, 10) : 449455096ull;
    double* a = malloc(n * 8);
    float* b = malloc(n * 4);
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; ++i) { a[i] = 1.0 / (double)(i + 1); b[i] = 0.f; }
    for (int nt = 1; nt <= omp_get_max_threads(); nt *= 2) {
        omp_set_num_threads(nt);
        double best = 1e9;
        for (int rep = 0; rep < 3; ++rep) {
            double t0 = now();
#pragma omp parallel for schedule(static)
            for (size_t i = 0; i < n; ++i) b[i] = (float)a[i];
            double t = now() - t0;
            if (t < best) best = t;
        }
        printf("threads %d: cast %zu doubles in %.1f ms = %.1f GB/s read + %.1f GB/s write\n", nt, n, best * 1e3, n * 8 / best / 1e9, n * 4 / best / 1e9);
    }
    double t0 = now();
    memcpy(b, a, n * 4);
    printf("memcpy 1 thread %.1f GB/s\n", n * 4 / (now() - t0) / 1e9);
    printf("max threads %d, procs %d\n", omp_get_max_threads(), omp_get_num_procs());
    return 0;
}
