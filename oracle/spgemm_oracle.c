/*
 * spgemm_oracle.c -- CPU restatement of spada-sim's functional SpGEMM path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (spada-sim_b200/) may
 * import, link or call this file; it is used by tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs as the checker.
 *
 * PARITY STATUS: "parity unpinned" by the reference's own tests -- the
 * reference (tsinghua-ideal/spada-sim, a single-threaded Rust cycle simulator)
 * ships no tests, golden vectors or recorded outputs,
 * and it cannot be compiled here (no cargo/rustc, un-vendored crates).  The
 * restatement is therefore pinned against (1) scipy 1.18.1 `A @ B` +
 * `sort_indices()` -- bit-identical structure and f64 bits, tests/test_oracle.py
 * -- and (2) the survey-derived known answers for matrices/cari.mtx
 * (tests/golden/cari_known_answers.json).
 *
 * What is restated (reference file:line, relative to /root/reference):
 *   - operands are canonical CSR: indptr / indices / data (storage.rs:150-160,
 *     214-239); "columns" of a window are positions in the stored row
 *     (storage.rs:279-323).
 *   - a product is ONE rounded f64 multiply, no FMA: simulator.rs:100-101
 *     (`a.value * b.value`), adder_tree.rs:37-57.
 *   - products of one A-row group are sorted by column (stable) and equal
 *     columns are summed left to right: simulator.rs:143-171 (sort_by col),
 *     simulator.rs:199-230 (`m.last_mut().value += e.value`).
 *   - partial rows of the same C row are k-way merged by ascending column and
 *     equal [row,col] are summed: adder_tree.rs:73-83, 145-188;
 *     scheduler.rs:381-480, 820-920.
 *   - partial rows are concatenated, never pruned (structural zeros stay):
 *     storage.rs:81-90, 685-735.
 *   - result assembly: one CsrRow per A row in raw row order, empty when the A
 *     row is empty or no product was ever written: simulator.rs:1034-1062.
 *   - the simplest equivalent statement of the per-row math in the reference
 *     is storage_traffic_model.rs:1668-1697 (sorted insert / `+= sf * value`).
 *
 * Canonical association fixed by this oracle: for each row i,
 * for p over A.row(i) in stored (ascending-k) order, for q over B.row(k) in
 * stored order:  acc[j] = first ? fl(a*b) : fl(acc[j] + fl(a*b)).
 * The reference's own association is schedule dependent (HashMap iteration
 * order), hence north_star's 1e-12 relative tolerance.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_API __attribute__((visibility("default")))

static int cmp_i32(const void *a, const void *b) {
    int32_t x = *(const int32_t *)a, y = *(const int32_t *)b;
    return (x > y) - (x < y);
}

ORACLE_API int oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/*
 * Intermediate-product count per A row: sum over stored nonzeros (i,k) of
 * len(B.row(k)).  This is what the scheduler's row-length tables
 * (scheduler.rs:197-202, b_row_lens) add up to per window; it is the quantity
 * the engine's K1 pass bins on.  Returns the total.
 */
ORACLE_API int64_t oracle_flops(int64_t m, const int64_t *Ap, const int32_t *Aj,
                                const int64_t *Bp, int64_t *flops) {
    int64_t total = 0;
    for (int64_t i = 0; i < m; ++i) {
        int64_t f = 0;
        for (int64_t p = Ap[i]; p < Ap[i + 1]; ++p) {
            int32_t k = Aj[p];
            f += Bp[k + 1] - Bp[k];
        }
        flops[i] = f;
        total += f;
    }
    return total;
}

/*
 * Structure of C: per-row number of distinct columns (the dedupe of
 * simulator.rs:209-221 / adder_tree.rs:73-83 -- no numerical pruning).
 * Writes Cp[0..m]; returns nnz(C).  `threads` <= 1 runs sequentially.
 */
ORACLE_API int64_t oracle_spgemm_symbolic(int64_t m, int64_t n, const int64_t *Ap,
                                          const int32_t *Aj, const int64_t *Bp,
                                          const int32_t *Bj, int64_t *Cp, int threads) {
    if (threads < 1) threads = 1;
    Cp[0] = 0;
#pragma omp parallel num_threads(threads) if (threads > 1)
    {
        int64_t *mark = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
        for (int64_t j = 0; j < n; ++j) mark[j] = -1;
#pragma omp for schedule(dynamic, 256)
        for (int64_t i = 0; i < m; ++i) {
            int64_t cnt = 0;
            for (int64_t p = Ap[i]; p < Ap[i + 1]; ++p) {
                int32_t k = Aj[p];
                for (int64_t q = Bp[k]; q < Bp[k + 1]; ++q) {
                    int32_t j = Bj[q];
                    if (mark[j] != i) {
                        mark[j] = i;
                        ++cnt;
                    }
                }
            }
            Cp[i + 1] = cnt;
        }
        free(mark);
    }
    for (int64_t i = 0; i < m; ++i) Cp[i + 1] += Cp[i];
    return Cp[m];
}

/*
 * Values and column ids of C, rows canonical (ascending unique columns),
 * association as stated in the header.  Cp must come from
 * oracle_spgemm_symbolic.
 */
ORACLE_API void oracle_spgemm_numeric(int64_t m, int64_t n, const int64_t *Ap,
                                      const int32_t *Aj, const double *Ax, const int64_t *Bp,
                                      const int32_t *Bj, const double *Bx, const int64_t *Cp,
                                      int32_t *Cj, double *Cx, int threads) {
    if (threads < 1) threads = 1;
#pragma omp parallel num_threads(threads) if (threads > 1)
    {
        int64_t *mark = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
        double *acc = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
        for (int64_t j = 0; j < n; ++j) mark[j] = -1;
#pragma omp for schedule(dynamic, 256)
        for (int64_t i = 0; i < m; ++i) {
            int32_t *cj = Cj + Cp[i];
            double *cx = Cx + Cp[i];
            int64_t cnt = 0;
            for (int64_t p = Ap[i]; p < Ap[i + 1]; ++p) {
                int32_t k = Aj[p];
                double a = Ax[p];
                for (int64_t q = Bp[k]; q < Bp[k + 1]; ++q) {
                    int32_t j = Bj[q];
                    double prod = a * Bx[q]; /* one rounded multiply (simulator.rs:101) */
                    if (mark[j] != i) {
                        mark[j] = i;
                        acc[j] = prod;
                        cj[cnt++] = j;
                    } else {
                        acc[j] = acc[j] + prod; /* separate rounded add (simulator.rs:217) */
                    }
                }
            }
            qsort(cj, (size_t)cnt, sizeof(int32_t), cmp_i32);
            for (int64_t t = 0; t < cnt; ++t) cx[t] = acc[cj[t]];
        }
        free(mark);
        free(acc);
    }
}

/*
 * B = A^T as CSR for the non-square SS workloads (gemm.rs:41-53,
 * `transpose_into().to_csr()`): counting transpose, rows of B come out with
 * ascending column ids because A's rows are visited in ascending order.
 */
ORACLE_API void oracle_transpose(int64_t m, int64_t n, const int64_t *Ap, const int32_t *Aj,
                                 const double *Ax, int64_t *Bp, int32_t *Bj, double *Bx) {
    for (int64_t j = 0; j <= n; ++j) Bp[j] = 0;
    for (int64_t p = 0; p < Ap[m]; ++p) Bp[Aj[p] + 1]++;
    for (int64_t j = 0; j < n; ++j) Bp[j + 1] += Bp[j];
    int64_t *cur = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
    memcpy(cur, Bp, sizeof(int64_t) * (size_t)n);
    for (int64_t i = 0; i < m; ++i)
        for (int64_t p = Ap[i]; p < Ap[i + 1]; ++p) {
            int64_t d = cur[Aj[p]]++;
            Bj[d] = (int32_t)i;
            Bx[d] = Ax[p];
        }
    free(cur);
}

/*
 * Canonical-CSR check used by the loaders' contract (SURVEY.md 8a "unsorted /
 * duplicate columns"): returns 0 if every row has strictly ascending column
 * ids inside [0, ncols) and indptr is monotone from 0; otherwise 1 + the first
 * offending row.
 */
ORACLE_API int64_t oracle_validate_csr(int64_t m, int64_t ncols, const int64_t *Ap,
                                       const int32_t *Aj) {
    if (Ap[0] != 0) return 1;
    for (int64_t i = 0; i < m; ++i) {
        if (Ap[i + 1] < Ap[i]) return 1 + i;
        for (int64_t p = Ap[i]; p < Ap[i + 1]; ++p) {
            if (Aj[p] < 0 || Aj[p] >= ncols) return 1 + i;
            if (p > Ap[i] && Aj[p] <= Aj[p - 1]) return 1 + i;
        }
    }
    return 0;
}

/*
 * Row groups of similar length (rowwise_perf_adjust.rs:36-77, parse_group with
 * var_factor 1.5, simulator.rs:449): consecutive rows stay in one group while
 * each non-empty row's length is within x var_factor of the previous non-empty
 * row's.  Writes the start row of every group into group_start (capacity m+1)
 * and returns the number of groups.  f32 arithmetic like the reference.
 */
ORACLE_API int64_t oracle_parse_group(int64_t m, const int64_t *Ap, float var_factor,
                                      int64_t *group_start) {
    int64_t ng = 0;
    int64_t prev = -1, row_s = 0;
    for (int64_t idx = 0; idx < m; ++idx) {
        int64_t len = Ap[idx + 1] - Ap[idx];
        if (len == 0) continue;
        if (prev < 0) {
            prev = len;
        } else if ((float)prev * var_factor < (float)len || (float)prev > var_factor * (float)len) {
            group_start[ng++] = row_s;
            prev = len;
            row_s = idx;
        } else {
            prev = len;
        }
    }
    if (m > 0) group_start[ng++] = row_s;
    return ng;
}
