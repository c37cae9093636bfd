// Minimal stand-in for the Boost.Test macros the reference tests use, and their fixtures
// (tests/context_setup.hpp, tests/random_vector.hpp, tests/random_matrix.hpp).
#pragma once
#include <cmath>
#include <csignal>
#include <execinfo.h>
#include <unistd.h>
#include <cstdlib>
#include <ctime>
#include <functional>
#include <iostream>
#include <random>
#include <set>
#include <string>
#include <vector>
#include <vexcl/vexcl.hpp>

namespace testing {
struct registry {
    std::vector<std::pair<std::string, std::function<void()>>> cases;
    int failures = 0, checks = 0;
    static registry& get() { static registry r; return r; }
};
struct registrar { registrar(const char *n, std::function<void()> f) { registry::get().cases.emplace_back(n, f); } };
inline void fail(const char *file, int line, const std::string &what) {
    ++registry::get().failures;
    std::cerr << file << ":" << line << ": check failed: " << what << std::endl;
}
}
#define BOOST_AUTO_TEST_CASE(name) static void name(); static testing::registrar reg_##name(#name, name); static void name()
#define BOOST_CHECK(c) do { ++testing::registry::get().checks; if (!(c)) testing::fail(__FILE__, __LINE__, #c); } while (0)
#define BOOST_REQUIRE(c) do { ++testing::registry::get().checks; if (!(c)) { testing::fail(__FILE__, __LINE__, #c); throw std::runtime_error("requirement failed"); } } while (0)
#define BOOST_CHECK_EQUAL(a, b) BOOST_CHECK((a) == (b))
// third argument is a PERCENT tolerance, as in Boost.Test
#define BOOST_CHECK_CLOSE(a, b, pct) do { ++testing::registry::get().checks; double a_ = (a), b_ = (b); \
    double d_ = std::fabs(a_ - b_), t_ = (pct) / 100.0; \
    if (!(d_ <= t_ * std::fabs(a_) && d_ <= t_ * std::fabs(b_)) && d_ != 0) testing::fail(__FILE__, __LINE__, std::string(#a " ~ " #b " : ") + std::to_string(a_) + " vs " + std::to_string(b_)); } while (0)
#define BOOST_CHECK_SMALL(a, tol) do { ++testing::registry::get().checks; if (!(std::fabs(a) <= (tol))) testing::fail(__FILE__, __LINE__, #a " is not small"); } while (0)
#define BOOST_CHECK_THROW(expr, exc) do { ++testing::registry::get().checks; bool t_ = false; try { expr; } catch (const exc&) { t_ = true; } catch (...) {} \
    if (!t_) testing::fail(__FILE__, __LINE__, #expr " did not throw " #exc); } while (0)

// The context: all devices, and -- like the reference fixture -- when only one device is present a
// second queue on the same device so that every multi-device path runs (context_setup.hpp:24-39).
inline vex::Context& test_context() {
    static std::unique_ptr<vex::Context> ctx;
    if (!ctx) {
        ctx.reset(new vex::Context(vex::Filter::DoublePrecision && vex::Filter::Env));
        if (ctx->size() == 1) {
            std::vector<vex::backend::command_queue> q = ctx->queue();
            q.push_back(vex::backend::duplicate_queue(q[0]));
            if (std::getenv("VEXCL_TEST_PARTS") && std::atoi(std::getenv("VEXCL_TEST_PARTS")) == 1) q.pop_back();
            ctx.reset(new vex::Context(q));
        }
    }
    return *ctx;
}
#define ctx test_context()

template <class T> struct generator {
    static T get() {
        static std::default_random_engine rng(std::rand());
        if constexpr (std::is_floating_point<T>::value) { static std::uniform_real_distribution<T> rnd((T)0, (T)1); return rnd(rng); }
        else { static std::uniform_int_distribution<T> rnd(0, 100); return rnd(rng); }
    }
};
template <class T> std::vector<T> random_vector(size_t n) { std::vector<T> x(n); for (auto &v : x) v = generator<T>::get(); return x; }

template <typename RT, typename CT, typename VT>
void random_matrix(size_t n, size_t m, size_t nnz_per_row, std::vector<RT> &row, std::vector<CT> &col, std::vector<VT> &val) {
    row.clear(); col.clear();
    std::default_random_engine rng(std::rand());
    std::uniform_int_distribution<size_t> random_width(0, nnz_per_row - 1), random_column(0, m - 1);
    row.push_back(0);
    for (size_t k = 0; k < n; k++) {
        size_t width = random_width(rng);
        std::set<CT> cs;
        while (cs.size() < width) cs.insert(static_cast<CT>(random_column(rng)));
        for (auto c : cs) col.push_back(c);
        row.push_back(static_cast<RT>(col.size()));
    }
    random_vector<VT>(col.size()).swap(val);
}

#define SAMPLE_SIZE 32
template <class V, class F> void check_sample(const V &v, F f) {
    for (size_t i = 0; i < SAMPLE_SIZE; ++i) { size_t idx = rand() % v.size(); f(idx, v[idx]); }
}
template <class V1, class V2, class F> void check_sample(const V1 &v1, const V2 &v2, F f) {
    BOOST_REQUIRE(v1.size() == v2.size());
    for (size_t i = 0; i < SAMPLE_SIZE; ++i) { size_t idx = rand() % v1.size(); f(idx, v1[idx], v2[idx]); }
}
template <class V1, class V2, class V3, class F> void check_sample(const V1 &v1, const V2 &v2, const V3 &v3, F f) {
    for (size_t i = 0; i < SAMPLE_SIZE; ++i) { size_t idx = rand() % v1.size(); f(idx, v1[idx], v2[idx], v3[idx]); }
}

// A crash (also one during static destruction after main) leaves a backtrace on stderr instead of a bare status 139.
inline void crash_backtrace(int sig) {
    void *bt[64];
    const int n = backtrace(bt, 64);
    const char msg[] = "fatal signal, backtrace:\n";
    if (write(2, msg, sizeof(msg) - 1) < 0) {}
    backtrace_symbols_fd(bt, n, 2);
    _exit(128 + sig);
}

int main(int argc, char **argv) {
    std::signal(SIGSEGV, crash_backtrace);
    std::signal(SIGABRT, crash_backtrace);
    unsigned seed = argc > 1 ? std::atoi(argv[1]) : static_cast<unsigned>(time(0));
    std::cout << "seed: " << seed << std::endl;
    srand(seed);
    try {
        std::cout << test_context() << std::endl;
        for (auto &c : testing::registry::get().cases) {
            std::cout << "  " << c.first << std::endl;
            try { c.second(); }
            catch (const std::exception &e) { testing::fail("", 0, c.first + " threw: " + e.what()); }
        }
    } catch (const vex::error &e) { std::cerr << "Error: " << e << std::endl; return 2; }
    auto &r = testing::registry::get();
    std::cout << r.cases.size() << " cases, " << r.checks << " checks, " << r.failures << " failures" << std::endl;
    return r.failures ? 1 : 0;
}
