/*
 * oracle.c -- CPU restatement of the three VexCL hot paths.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this; the product (vexcl_b200/, include/) never does.
 *
 * Why a restatement: the reference itself cannot be built in this image
 * (Boost and the OpenCL headers are absent: vexcl/vector.hpp:44 fails on
 * <boost/proto/proto.hpp>, vexcl/types.hpp:47 on <CL/cl_platform.h>), and its
 * kernels are source text generated at run time.  Each function below cites the
 * reference lines whose arithmetic and ordering it follows (paths relative to
 * /root/reference).  The reference holds no stored golden vectors; parity is
 * pinned by the closed-form known answers and inline restatements of its own
 * tests (tests/vector_arithmetics.cpp, tests/spmv.cpp,
 * tests/sparse_matrices.cpp, examples/benchmark.cpp) -- see tests/test_oracle_kat.py.
 *
 * Execution model restated: the OpenCL-CPU backend.  A kernel runs on
 * G = 8 * compute_units work-groups of ONE work-item
 * (vexcl/backend/opencl/kernel.hpp:166-171, :193-194); each work-item owns a
 * contiguous chunk [g*ceil(n/G), ...) (vexcl/backend/opencl/source.hpp:255-268).
 * Groups are spread over host threads with OpenMP.
 *
 * Compiled with -ffp-contract=off: `a*b + c` is a multiply then an add, each
 * rounded (the *_fma variants use fma() explicitly).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

void orc_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ---- partition: vexcl/vector.hpp:131-167, vexcl/util.hpp:91-93 ---------------------- */
static size_t alignup16(size_t n) { return (n + 15) / 16 * 16; }

void orc_partition(size_t n, int nparts, const double *weights, size_t *part) {
    part[0] = 0;
    if (nparts > 1) {
        double *cumsum = (double *)malloc(sizeof(double) * (size_t)(nparts + 1));
        cumsum[0] = 0;
        for (int d = 0; d < nparts; ++d) cumsum[d + 1] = cumsum[d] + (weights ? weights[d] : 1.0);
        for (int d = 1; d < nparts; ++d) {
            size_t b = alignup16((size_t)(n * cumsum[d] / cumsum[nparts]));
            part[d] = b < n ? b : n;
        }
        free(cumsum);
    }
    part[nparts] = n;
}

/* ---- random inputs: std::default_random_engine (= minstd_rand0) feeding
 * std::uniform_real_distribution<double>(0,1), as tests/random_vector.hpp:12-19 and
 * examples/benchmark.cpp:72-80 do.  libstdc++'s generate_canonical<double,53> draws
 * k = 2 values: S = (x1-1) + (x2-1)*R, R = 2147483646, result S / R^2
 * (and nextafter(1,0) if that rounds to 1). */
static uint32_t minstd_next(uint32_t *s) {
    *s = (uint32_t)(((uint64_t)*s * 16807u) % 2147483647u);
    return *s;
}

void orc_uniform_real(uint32_t seed, size_t n, double *out) {
    uint32_t s = seed % 2147483647u;
    if (s == 0) s = 1;
    const double R = 2147483646.0;
    for (size_t i = 0; i < n; ++i) {
        double sum = 0.0, tmp = 1.0;
        for (int k = 0; k < 2; ++k) {
            sum += (double)(minstd_next(&s) - 1u) * tmp;
            tmp *= R;
        }
        double r = sum / tmp;
        if (r >= 1.0) r = nextafter(1.0, 0.0);
        out[i] = r;
    }
}

/* ---- matrices: examples/benchmark.cpp:357-415 (3-D 7-point Poisson on an n^3 grid,
 * boundary rows are identity, h2i = (n-1)^2) and its 2-D 5-point analogue. */
void orc_poisson_sizes(int dim, size_t n, size_t *nrows, size_t *nnz) {
    if (dim == 2) {
        *nrows = n * n;
        size_t in = n > 2 ? (n - 2) * (n - 2) : 0;
        *nnz = in * 5 + (n * n - in);
    } else {
        *nrows = n * n * n;
        size_t in = n > 2 ? (n - 2) * (n - 2) * (n - 2) : 0;
        *nnz = in * 7 + (n * n * n - in);
    }
}

void orc_poisson(int dim, size_t n, int64_t *row, int64_t *col, double *val) {
    const double h2i = (double)((n - 1) * (n - 1));
    size_t idx = 0, p = 0;
    row[0] = 0;
    const size_t nk = dim == 3 ? n : 1;
    for (size_t k = 0; k < nk; ++k)
        for (size_t j = 0; j < n; ++j)
            for (size_t i = 0; i < n; ++i, ++idx) {
                const int bnd = i == 0 || i == n - 1 || j == 0 || j == n - 1 || (dim == 3 && (k == 0 || k == n - 1));
                if (bnd) {
                    col[p] = (int64_t)idx; val[p] = 1; ++p;
                } else {
                    if (dim == 3) { col[p] = (int64_t)(idx - n * n); val[p] = -h2i; ++p; }
                    col[p] = (int64_t)(idx - n); val[p] = -h2i; ++p;
                    col[p] = (int64_t)(idx - 1); val[p] = -h2i; ++p;
                    col[p] = (int64_t)idx; val[p] = (dim == 3 ? 6 : 4) * h2i; ++p;
                    col[p] = (int64_t)(idx + 1); val[p] = -h2i; ++p;
                    col[p] = (int64_t)(idx + n); val[p] = -h2i; ++p;
                    if (dim == 3) { col[p] = (int64_t)(idx + n * n); val[p] = -h2i; ++p; }
                }
                row[idx + 1] = (int64_t)p;
            }
}

/* ---- elementwise: the emitted statement is the parenthesised transcription of the
 * expression (vexcl/operations.hpp:1209-1238), e.g. prm_1[idx] = ( prm_2[idx] + ( prm_3[idx] * prm_4[idx] ) );
 * run by every work-item over its chunk.  mode 0: a = b + c*d   1: a += b + c*d. */
static void chunk_bounds(size_t n, int G, int g, size_t *lo, size_t *hi) {
    size_t chunk = (n + (size_t)G - 1) / (size_t)G;
    *lo = (size_t)g * chunk;
    *hi = *lo + chunk;
    if (*lo > n) *lo = n;
    if (*hi > n) *hi = n;
}

void orc_vec_muladd(double *a, const double *b, const double *c, const double *d, size_t n, int mode, int use_fma, int G) {
#pragma omp parallel for schedule(static)
    for (int g = 0; g < G; ++g) {
        size_t lo, hi; chunk_bounds(n, G, g, &lo, &hi);
        if (mode == 0) {
            if (use_fma) for (size_t i = lo; i < hi; ++i) a[i] = fma(c[i], d[i], b[i]);
            else         for (size_t i = lo; i < hi; ++i) a[i] = b[i] + c[i] * d[i];
        } else {
            if (use_fma) for (size_t i = lo; i < hi; ++i) a[i] += fma(c[i], d[i], b[i]);
            else         for (size_t i = lo; i < hi; ++i) a[i] += b[i] + c[i] * d[i];
        }
    }
}

/* a = alpha * a + b   (examples/benchmark.cpp:102-107) */
void orc_vec_saxpy(double *a, double alpha, const double *b, size_t n, int use_fma, int G) {
#pragma omp parallel for schedule(static)
    for (int g = 0; g < G; ++g) {
        size_t lo, hi; chunk_bounds(n, G, g, &lo, &hi);
        if (use_fma) for (size_t i = lo; i < hi; ++i) a[i] = fma(alpha, a[i], b[i]);
        else         for (size_t i = lo; i < hi; ++i) a[i] = alpha * a[i] + b[i];
    }
}

/* ---- Reductor: vexcl/reductor.hpp.  Each work-item folds its chunk sequentially
 * (:511-533 plain, :537-564 Kahan); on a CPU device g_odata[group] = mySum (:358-363);
 * the host folds the partials in order starting from initial() (:420-436).
 * op: 0 SUM, 1 SUM_Kahan, 2 MAX, 3 MIN.  x holds the already-evaluated expression. */
double orc_reduce(const double *x, size_t n, int op, int G) {
    double *part = (double *)malloc(sizeof(double) * (size_t)G);
#pragma omp parallel for schedule(static)
    for (int g = 0; g < G; ++g) {
        size_t lo, hi; chunk_bounds(n, G, g, &lo, &hi);
        double s;
        if (op == 0) { s = 0; for (size_t i = lo; i < hi; ++i) s = s + x[i]; }
        else if (op == 1) {
            s = 0; double c = 0;
            for (size_t i = lo; i < hi; ++i) { double y = x[i] - c; double t = s + y; c = (t - s) - y; s = t; }
        }
        else if (op == 2) { s = -1.7976931348623157e308; for (size_t i = lo; i < hi; ++i) s = s > x[i] ? s : x[i]; }
        else              { s =  1.7976931348623157e308; for (size_t i = lo; i < hi; ++i) s = s < x[i] ? s : x[i]; }
        part[g] = s;
    }
    double r = (op == 2) ? -1.7976931348623157e308 : (op == 3) ? 1.7976931348623157e308 : 0.0;
    for (int g = 0; g < G; ++g) {
        if (op <= 1) r = r + part[g];
        else if (op == 2) r = r > part[g] ? r : part[g];
        else r = r < part[g] ? r : part[g];
    }
    free(part);
    return r;
}

/* sum(a * b) without materialising the product (examples/benchmark.cpp:236-241). */
double orc_reduce_dot(const double *a, const double *b, size_t n, int kahan, int G) {
    double *part = (double *)malloc(sizeof(double) * (size_t)G);
#pragma omp parallel for schedule(static)
    for (int g = 0; g < G; ++g) {
        size_t lo, hi; chunk_bounds(n, G, g, &lo, &hi);
        double s = 0, c = 0;
        if (!kahan) for (size_t i = lo; i < hi; ++i) s = s + a[i] * b[i];
        else for (size_t i = lo; i < hi; ++i) { double y = a[i] * b[i] - c; double t = s + y; c = (t - s) - y; s = t; }
        part[g] = s;
    }
    double r = 0;
    for (int g = 0; g < G; ++g) r = r + part[g];
    free(part);
    return r;
}

/* Reference accumulator used by tests/vector_arithmetics.cpp:82-86: a Kahan sum over all elements. */
double orc_kahan_sum(const double *x, size_t n) {
    double s = 0, c = 0;
    for (size_t i = 0; i < n; ++i) { double y = x[i] - c; double t = s + y; c = (t - s) - y; s = t; }
    return s;
}

/* ---- SpMV: vexcl/spmat/csr.inl:163-170
 *     sum = 0; for (j = row[i]; j < row[i+1]; ++j) sum += val[j] * in[col[j]]; out[i] OP scale * sum;
 * The same loop is the inline check of tests/spmv.cpp:28-34.  row may start at a non-zero offset. */
void orc_csr_spmv(size_t n, const int64_t *row, const int64_t *col, const double *val, const double *x, double *y,
                  double alpha, int append, int use_fma, int G) {
#pragma omp parallel for schedule(static)
    for (int g = 0; g < G; ++g) {
        size_t lo, hi; chunk_bounds(n, G, g, &lo, &hi);
        for (size_t i = lo; i < hi; ++i) {
            double sum = 0;
            if (use_fma) for (int64_t j = row[i]; j < row[i + 1]; ++j) sum = fma(val[j], x[col[j]], sum);
            else         for (int64_t j = row[i]; j < row[i + 1]; ++j) sum += val[j] * x[col[j]];
            if (append) y[i] += alpha * sum; else y[i] = alpha * sum;
        }
    }
}

/* Row magnitude sum_j |val_j * x_col_j| -- the scale against which a 1e-10 relative
 * tolerance is meaningful for rows that cancel (Poisson interior rows with constant x). */
void orc_csr_absrow(size_t n, const int64_t *row, const int64_t *col, const double *val, const double *x, double *out) {
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; ++i) {
        double s = 0;
        for (int64_t j = row[i]; j < row[i + 1]; ++j) s += fabs(val[j] * x[col[j]]);
        out[i] = s;
    }
}

/* Hybrid ELL product: vexcl/spmat/hybrid_ell.inl:252-268 (ELL columns first, sentinel -1, then the CSR tail). */
void orc_hell_spmv(size_t n, size_t ell_w, size_t ell_pitch, const int64_t *ell_col, const double *ell_val,
                   const int64_t *csr_row, const int64_t *csr_col, const double *csr_val,
                   const double *x, double *y, double alpha, int append) {
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; ++i) {
        double sum = 0;
        for (size_t j = 0; j < ell_w; ++j) {
            int64_t c = ell_col[i + j * ell_pitch];
            if (c != -1) sum += ell_val[i + j * ell_pitch] * x[c];
        }
        if (csr_row) for (int64_t j = csr_row[i]; j < csr_row[i + 1]; ++j) sum += csr_val[j] * x[csr_col[j]];
        if (append) y[i] += alpha * sum; else y[i] = alpha * sum;
    }
}

/* ELL width: vexcl/spmat/hybrid_ell.inl:66-113.  widths[i] = number of entries of row i in this part. */
size_t orc_hell_width(const int64_t *widths, size_t n) {
    size_t maxw = 0;
    for (size_t i = 0; i < n; ++i) if ((size_t)widths[i] > maxw) maxw = (size_t)widths[i];
    size_t *hist = (size_t *)calloc(maxw + 1, sizeof(size_t));
    for (size_t i = 0; i < n; ++i) ++hist[widths[i]];
    size_t rows = n, w = maxw;
    for (size_t i = 0; i < maxw; ++i) {
        rows -= hist[i];
        if (3.0 * rows < n) { w = i; break; }
    }
    free(hist);
    return w;
}

/* The reference benchmark's own single-thread "C++" loops, verbatim in meaning
 * (examples/benchmark.cpp:190-195, :255-260, :445-454). */
void orc_cpp_vec(double *A, const double *B, const double *C, const double *D, size_t N) {
    for (size_t j = 0; j < N; ++j) A[j] += B[j] + C[j] * D[j];
}
double orc_cpp_dot(const double *A, const double *B, size_t N) {
    double s = 0;
    for (size_t j = 0; j < N; ++j) s = s + A[j] * B[j];
    return s;
}
void orc_cpp_spmv(size_t N, const int64_t *row, const int64_t *col, const double *val, const double *X, double *Y) {
    for (size_t i = 0; i < N; ++i) {
        double s = 0;
        for (int64_t j = row[i]; j < row[i + 1]; ++j) s += val[j] * X[col[j]];
        Y[i] += s;
    }
}

/* ---- timing helpers (bench.py's CPU baseline only) -------------------------------------------
 * First-touch placement: on a multi-socket host a buffer filled by one thread lives on one NUMA node
 * and every other core reads it remotely.  These clones are written by the same static work-group ->
 * thread mapping the kernels above use, so each thread's chunk is local to it, which is what an
 * OpenCL CPU runtime's own buffers would look like after a first kernel has written them.
 * row_chunks != NULL: elements are nonzeros, chunked by the ROW ranges of the groups. */
void *orc_numa_clone(const void *src, size_t n, size_t elem, const int64_t *row, size_t nrows, int G) {
    char *dst = (char *)malloc(n * elem + 64);
    if (!dst) return NULL;
#pragma omp parallel for schedule(static)
    for (int g = 0; g < G; ++g) {
        size_t lo, hi;
        if (row) { size_t rl, rh; chunk_bounds(nrows, G, g, &rl, &rh); lo = (size_t)(row[rl] - row[0]); hi = (size_t)(row[rh] - row[0]); }
        else chunk_bounds(n, G, g, &lo, &hi);
        if (hi > lo) memcpy(dst + lo * elem, (const char *)src + lo * elem, (hi - lo) * elem);
    }
    return dst;
}

void orc_free(void *p) { free(p); }

/* One timed pass on raw (cloned) pointers. */
void orc_csr_spmv_raw(size_t n, const int64_t *row, const int64_t *col, const double *val, const double *x, double *y, int G) {
    orc_csr_spmv(n, row, col, val, x, y, 1.0, 0, 0, G);
}
