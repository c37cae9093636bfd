#ifndef VEXCL_PROFILER_HPP
#define VEXCL_PROFILER_HPP
// Nested named timers with the reference's interface (vexcl/profiler.hpp:92-358):
// tic_cpu / tic_cl / toc, printed as an indented tree.  tic_cl and its toc bracket the
// region with queue.finish(), like the reference does.
#include <chrono>
#include <iomanip>
#include <iostream>
#include <map>
#include <memory>
#include <stack>
#include <string>
#include <vector>
#include "backend.hpp"

namespace vex {

template <class Clock = std::chrono::high_resolution_clock>
class stopwatch {
    public:
        stopwatch() : n(0), total(0) { tic(); }
        void tic() { start = Clock::now(); }
        double toc() {
            const double d = std::chrono::duration<double>(Clock::now() - start).count();
            total += d; ++n;
            return d;
        }
        double average() const { return n ? total / n : 0.0; }
        double total_time() const { return total; }
        size_t tics() const { return n; }
    private:
        size_t n; double total;
        typename Clock::time_point start;
};

template <class Clock = std::chrono::high_resolution_clock>
class profiler {
    struct unit {
        std::string name; bool cl = false; double total = 0; size_t hits = 0;
        typename Clock::time_point start;
        std::vector<std::unique_ptr<unit>> children;
        unit* child(const std::string &nm) {
            for (auto &c : children) if (c->name == nm) return c.get();
            children.emplace_back(new unit()); children.back()->name = nm; return children.back().get();
        }
        void print(std::ostream &os, int depth, double parent) const {
            os << std::string(2 * depth, ' ') << name << ": " << std::setw(12 - 2 * std::min(depth, 5)) << std::fixed << std::setprecision(3) << total << " sec.";
            if (parent > 0) os << " (" << std::setprecision(2) << 100 * total / parent << "%)";
            if (hits > 1) os << " [" << hits << "x]";
            os << std::endl;
            for (auto &c : children) c->print(os, depth + 1, total);
        }
    };
    public:
        profiler(const std::vector<backend::command_queue> &queue = std::vector<backend::command_queue>(),
                 const std::string &name = "Profile") : queue(queue) {
            root.name = name; root.start = Clock::now(); stack.push(&root);
        }
        void tic_cpu(const std::string &name) { push(name, false); }
        void tic_cl(const std::string &name) { for (auto &q : queue) q.finish(); push(name, true); }
        double toc(const std::string& = "") {
            unit *u = stack.top();
            if (u == &root) return 0;
            if (u->cl) for (auto &q : queue) q.finish();
            const double d = std::chrono::duration<double>(Clock::now() - u->start).count();
            u->total += d; ++u->hits; stack.pop();
            return d;
        }
        void reset() { root.children.clear(); while (stack.size() > 1) stack.pop(); root.start = Clock::now(); }
        void print(std::ostream &os) const {
            unit copy; copy.name = root.name;
            const double t = std::chrono::duration<double>(Clock::now() - root.start).count();
            os << std::endl << root.name << ": " << std::fixed << std::setprecision(3) << t << " sec." << std::endl;
            for (auto &c : root.children) c->print(os, 1, t);
        }
    private:
        std::vector<backend::command_queue> queue;
        unit root;
        std::stack<unit*> stack;
        void push(const std::string &name, bool cl) {
            unit *u = stack.top()->child(name); u->cl = cl; u->start = Clock::now(); stack.push(u);
        }
};

template <class Clock>
inline std::ostream& operator<<(std::ostream &os, const profiler<Clock> &p) { p.print(os); return os; }

} // namespace vex
#endif
