// The hot-path functions of the reference's examples/benchmark.cpp (benchmark_saxpy :83-148,
// benchmark_vector :152-216, benchmark_reductor :219-278, benchmark_spmv :352-477), written against
// include/vexcl.  Same problem sizes, same reporting formulae, same self-check ("res").
//   g++ -std=c++17 -O2 -I include examples/hotpath_benchmark.cpp -L vexcl_b200 -lvexb200 -o hotpath_benchmark
#include <cmath>
#include <iostream>
#include <numeric>
#include <random>
#include <vexcl/vexcl.hpp>

typedef double real;

static std::vector<real> random_vector(size_t n) {
    std::default_random_engine rng(std::rand());
    std::uniform_real_distribution<real> rnd(0.0, 1.0);
    std::vector<real> x(n);
    for (auto &v : x) v = rnd(rng);
    return x;
}

static void report(const char *title, double gflops, double bwidth, double res) {
    std::cout << title << " (double)\n  B200\n    GFLOPS:    " << gflops << "\n    Bandwidth: " << bwidth
              << "\n  res = " << res << "\n" << std::endl;
}

static void benchmark_saxpy(const vex::Context &ctx, vex::profiler<> &prof, size_t N, size_t M) {
    std::vector<real> A(N, 0), B = random_vector(N);
    real alpha = random_vector(1)[0];
    vex::vector<real> a(ctx, A), b(ctx, B);
    auto ta = vex::tag<1>(a);
    ta = alpha * ta + b;
    ta = static_cast<real>(0);
    prof.tic_cpu("saxpy");
    for (size_t i = 0; i < M; i++) ta = alpha * ta + b;
    ctx.finish();
    double t = prof.toc("saxpy");
    for (size_t i = 0; i < M; i++) for (size_t j = 0; j < N; j++) A[j] = alpha * A[j] + B[j];
    vex::copy(A, b);
    vex::Reductor<real, vex::SUM> sum(ctx);
    a -= b;
    report("Vector SAXPY", 2.0 * N * M / t / 1e9, 3.0 * N * M * sizeof(real) / t / 1e9, sum(a * a));
}

static void benchmark_vector(const vex::Context &ctx, vex::profiler<> &prof, size_t N, size_t M) {
    std::vector<real> A(N, 0), B = random_vector(N), C = random_vector(N), D = random_vector(N);
    vex::vector<real> a(ctx, A), b(ctx, B), c(ctx, C), d(ctx, D);
    a += b + c * d;
    a = 0;
    prof.tic_cpu("vector");
    for (size_t i = 0; i < M; i++) a += b + c * d;
    ctx.finish();
    double t = prof.toc("vector");
    for (size_t i = 0; i < M; i++) for (size_t j = 0; j < N; j++) A[j] += B[j] + C[j] * D[j];
    vex::copy(A, b);
    vex::Reductor<real, vex::SUM> sum(ctx);
    a -= b;
    report("Vector arithmetic", 3.0 * N * M / t / 1e9, 5.0 * N * M * sizeof(real) / t / 1e9, sum(a * a));
}

static void benchmark_reductor(const vex::Context &ctx, vex::profiler<> &prof, size_t N, size_t M) {
    std::vector<real> A = random_vector(N), B = random_vector(N);
    vex::vector<real> a(ctx, A), b(ctx, B);
    vex::Reductor<real, vex::SUM> sum(ctx);
    real sum_cl = sum(a * b);
    sum_cl = 0;
    prof.tic_cpu("reductor");
    for (size_t i = 0; i < M; i++) sum_cl += sum(a * b);
    ctx.finish();
    double t = prof.toc("reductor");
    real sum_cpp = 0;
    for (size_t i = 0; i < M; i++) sum_cpp += std::inner_product(A.begin(), A.end(), B.begin(), static_cast<real>(0));
    report("Reduction", 2.0 * N * M / t / 1e9, 2.0 * N * M * sizeof(real) / t / 1e9, std::fabs((sum_cl - sum_cpp) / sum_cpp));
}

static void benchmark_spmv(const vex::Context &ctx, vex::profiler<> &prof, size_t n, size_t M) {
    const size_t N = n * n * n;
    const real h2i = (n - 1) * (n - 1);
    std::vector<size_t> row; std::vector<uint> col; std::vector<real> val;
    std::vector<real> X(N, static_cast<real>(1e-2)), Y(N, 0);
    row.push_back(0);
    for (size_t k = 0, idx = 0; k < n; k++) for (size_t j = 0; j < n; j++) for (size_t i = 0; i < n; i++, idx++) {
        if (i == 0 || i == n - 1 || j == 0 || j == n - 1 || k == 0 || k == n - 1) {
            col.push_back(idx); val.push_back(1);
        } else {
            const long off[] = {-(long)(n * n), -(long)n, -1, 0, 1, (long)n, (long)(n * n)};
            for (int t = 0; t < 7; ++t) { col.push_back(idx + off[t]); val.push_back(t == 3 ? 6 * h2i : -h2i); }
        }
        row.push_back(col.size());
    }
    const size_t nnz = row.back();
    vex::SpMat<real, uint> A(ctx, N, N, row.data(), col.data(), val.data());
    vex::vector<real> x(ctx, X), y(ctx, Y);
    y += A * x;
    y = 0;
    prof.tic_cpu("spmv");
    for (size_t i = 0; i < M; i++) y += A * x;
    ctx.finish();
    double t = prof.toc("spmv");
    const size_t Mc = std::min<size_t>(M, 16);                 // the CPU check loop is slow; scale the device side to match
    y = 0;
    for (size_t i = 0; i < Mc; i++) y += A * x;
    for (size_t k = 0; k < Mc; k++) for (size_t i = 0; i < N; i++) {
        real s = 0;
        for (size_t j = row[i]; j < row[i + 1]; j++) s += val[j] * X[col[j]];
        Y[i] += s;
    }
    vex::copy(Y, x);
    y -= x;
    vex::Reductor<real, vex::SUM> sum(ctx);
    report("SpMV", M / t / 1e9 * (2.0 * nnz + N), M / t / 1e9 * (nnz * (2 * sizeof(real) + sizeof(size_t)) + 4 * N * sizeof(real)), sum(y * y));
}

int main(int argc, char **argv) {
    try {
        vex::Context ctx(vex::Filter::Env && vex::Filter::DoublePrecision);
        if (!ctx) { std::cerr << "No compute devices" << std::endl; return 1; }
        std::cout << ctx << std::endl;
        const bool quick = argc > 1 && std::string(argv[1]) == "--quick";
        vex::profiler<> prof(ctx);
        benchmark_saxpy(ctx, prof, 1024 * 1024, quick ? 64 : 1024);
        benchmark_vector(ctx, prof, 1024 * 1024, quick ? 64 : 1024);
        benchmark_reductor(ctx, prof, 16 * 1024 * 1024, quick ? 8 : 64);
        benchmark_spmv(ctx, prof, quick ? 64 : 128, quick ? 32 : 1024);
        std::cout << prof << std::endl;
    } catch (const vex::error &e) {
        std::cerr << e << std::endl;
        return 2;
    }
    return 0;
}
