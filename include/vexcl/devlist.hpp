#ifndef VEXCL_DEVLIST_HPP
#define VEXCL_DEVLIST_HPP
// vex::Context: the device list a program computes on (vexcl/devlist.hpp:229-412).
#include <iostream>
#include <memory>
#include <vector>
#include "backend.hpp"
#include "util.hpp"

namespace vex {

class Context;

/// Holder of the "current" context used by the queue-less constructors (devlist.hpp:229-252).
template <bool dummy = true>
class StaticContext {
    public:
        static void set(const Context &ctx) { instance() = &ctx; }
        static const Context& get() {
            precondition(instance() != 0, "Uninitialized static context");
            return *instance();
        }
    private:
        static const Context*& instance() { static const Context *ctx = 0; return ctx; }
};
inline const Context& current_context() { return StaticContext<>::get(); }

class Context {
    public:
        /// Devices that satisfy the filter, one queue each.
        template <class DevFilter, class = typename std::enable_if<
            !std::is_same<typename std::decay<DevFilter>::type, std::vector<backend::command_queue>>::value &&
            !std::is_same<typename std::decay<DevFilter>::type, Context>::value>::type>
        explicit Context(DevFilter &&filter, unsigned queue_flags = 0) {
            std::tie(c, q) = backend::queue_list(std::forward<DevFilter>(filter), queue_flags);
#ifdef VEXCL_THROW_ON_EMPTY_CONTEXT
            precondition(!q.empty(), "No compute devices found");
#endif
            StaticContext<>::set(*this);
        }
        /// A user-assembled list (the reference's tests build a two-queue context on one device this way).
        Context(const std::vector<backend::context> &c, const std::vector<backend::command_queue> &q) : c(c), q(q) {
            StaticContext<>::set(*this);
        }
        explicit Context(const std::vector<backend::command_queue> &queues) : q(queues) {
            for (auto &x : q) c.push_back(x.context());
            StaticContext<>::set(*this);
        }

        const std::vector<backend::context>& context() const { return c; }
        const backend::context& context(unsigned d) const { return c[d]; }
        const std::vector<backend::command_queue>& queue() const { return q; }
        operator const std::vector<backend::command_queue>&() const { return q; }
        const backend::command_queue& queue(unsigned d) const { return q[d]; }
        backend::device device(unsigned d) const { return q[d].device(); }
        size_t size() const { return q.size(); }
        bool empty() const { return q.empty(); }
        operator bool() const { return !q.empty(); }
        void finish() const { for (auto &x : q) x.finish(); }
    private:
        std::vector<backend::context> c;
        std::vector<backend::command_queue> q;
};

inline std::ostream& operator<<(std::ostream &os, const std::vector<backend::command_queue> &queue) {
    unsigned p = 0;
    for (auto &x : queue) os << ++p << ". " << x.device().name() << std::endl;
    return os;
}
inline std::ostream& operator<<(std::ostream &os, const Context &ctx) { return os << ctx.queue(); }

} // namespace vex
#endif
