#ifndef VEXCL_STENCIL_HPP
#define VEXCL_STENCIL_HPP
/*
 * vex::stencil<T> (vexcl/stencil.hpp:168-330): convolution of a vector with a small stencil,
 *     y = x * s;   y += x * s;   y = 42 * (x * s);   y = x * s + x * s;        (also for multivectors)
 *     y[i] = sum_k s[k] * x[clamp(i + k - center, 0, n-1)]
 * and vex::StencilOperator / VEX_STENCIL_OPERATOR (stencil.hpp:510-680): y = f(X) with a user-supplied body,
 *     VEX_STENCIL_OPERATOR(oscillate, double, 3, 1, "return sin(X[1] - X[0]) + sin(X[0] - X[-1]);", ctx);
 *     y = oscillate(x);
 *
 * The convolution is libvexb200's stencil_kernel (csrc/stencil.cu); a user-defined operator is generated and compiled
 * by NVRTC at first use, as the reference generates its kernel per operator.  With several device slices the elements
 * a slice needs from its neighbours are copied device to device into a per-slice halo buffer before the launch (the
 * reference stages them through the host and finishes every queue twice, stencil.hpp:86-150); here the exchange is ordered
 * by events between the queues and never waits on the host.
 */
#include <algorithm>
#include <initializer_list>
#include <string>
#include "vector.hpp"
#include "multivector.hpp"

namespace vex {
namespace detail {

/// Halo buffers and their exchange, shared by stencil and StencilOperator (stencil_base, stencil.hpp:43-150).
template <typename T>
class stencil_halos {
    protected:
        std::vector<backend::command_queue> queue;
        std::vector<backend::device_vector<T>> dbuf;
        int width = 0, lhalo = 0, rhalo = 0;

        stencil_halos(const std::vector<backend::command_queue> &q, size_t n, unsigned center) : queue(q) {
            precondition(!queue.empty() && n >= 1 && center < n, "stencil needs width >= 1 and center < width");   // stencil.hpp:70-74
            width = static_cast<int>(n); lhalo = static_cast<int>(center); rhalo = width - lhalo - 1;
            for (unsigned d = 0; d < queue.size(); ++d)
                dbuf.emplace_back(queue[d], static_cast<size_t>(width));          // one more than needed, never empty (stencil.hpp:81-82)
        }

        /// Fills left[d] / right[d] with the halo pointers of every slice that is not at an end of the vector.
        /// Ordering is by events, nothing waits on the host (the reference finishes every queue twice, stencil.hpp:113,:147):
        /// a slice's copies wait for the event "x complete" of the queues they read from, and every queue that was read
        /// from waits -- on the device -- for the readers' "copies done" events before it runs anything later.
        void exchange_halos(const vex::vector<T> &x, std::vector<const T*> &left, std::vector<const T*> &right) const {
            precondition(x.nparts() == queue.size(), "stencil: the vector lives on other queues");
            if (queue.size() <= 1 || width <= 1) return;                          // stencil.hpp:89
            const size_t n = x.size();
            const unsigned nd = static_cast<unsigned>(queue.size());
            if (ready.empty()) {
                ready.assign(nd, nullptr); done.assign(nd, nullptr);
                for (unsigned d = 0; d < nd; ++d) {
                    VEXB_CHECKED(vexb_event_create(queue[d].ordinal(), &ready[d]));
                    VEXB_CHECKED(vexb_event_create(queue[d].ordinal(), &done[d]));
                }
            }
            for (unsigned d = 0; d < nd; ++d) VEXB_CHECKED(vexb_event_record(queue[d].ordinal(), ready[d], queue[d].raw()));
            std::vector<std::vector<char>> reads(nd, std::vector<char>(nd, 0));   // reads[d][p]: slice d copied from slice p
            for (unsigned d = 0; d < nd; ++d) {
                const size_t start = x.part_start(d), size = x.part_size(d);
                if (!size) continue;
                if (start > 0 && lhalo > 0) {
                    const size_t have = std::min<size_t>(start, lhalo);           // elements that exist before this slice
                    if (have < static_cast<size_t>(lhalo)) fill(d, 0, lhalo - have, x[0]);
                    gather(x, d, lhalo - have, start - have, start, reads[d]);
                    left[d] = dbuf[d].raw_ptr();
                }
                if (start + size < n && rhalo > 0) {
                    const size_t g0 = start + size, g1 = std::min(g0 + rhalo, n);
                    gather(x, d, lhalo, g0, g1, reads[d]);
                    if (g1 - g0 < static_cast<size_t>(rhalo)) fill(d, lhalo + (g1 - g0), rhalo - (g1 - g0), x[n - 1]);
                    right[d] = dbuf[d].raw_ptr() + lhalo;
                }
            }
            // nobody overwrites x while a neighbour still copies from it
            for (unsigned d = 0; d < nd; ++d) {
                bool any = false;
                for (unsigned p = 0; p < nd; ++p) any = any || reads[d][p];
                if (any) VEXB_CHECKED(vexb_event_record(queue[d].ordinal(), done[d], queue[d].raw()));
            }
            for (unsigned d = 0; d < nd; ++d)
                for (unsigned p = 0; p < nd; ++p)
                    if (p != d && reads[d][p]) VEXB_CHECKED(vexb_stream_wait_event(queue[p].ordinal(), queue[p].raw(), done[d]));
        }
        ~stencil_halos() {
            for (unsigned d = 0; d < ready.size(); ++d) { vexb_event_destroy(queue[d].ordinal(), ready[d]); vexb_event_destroy(queue[d].ordinal(), done[d]); }
        }
        stencil_halos(const stencil_halos&) = delete;
        stencil_halos& operator=(const stencil_halos&) = delete;
    private:
        /// Copy global elements [g0, g1) of x into dbuf[d] at element offset `at`.
        mutable std::vector<void*> ready, done;      // per queue: "x complete" / "my copies of the neighbours' x are done"
        void gather(const vex::vector<T> &x, unsigned d, size_t at, size_t g0, size_t g1, std::vector<char> &reads) const {
            for (unsigned p = 0; p < queue.size(); ++p) {
                const size_t a = std::max(g0, x.part_start(p)), b = std::min(g1, x.part_start(p) + x.part_size(p));
                if (a < b) {
                    if (p != d && !reads[p]) VEXB_CHECKED(vexb_stream_wait_event(queue[d].ordinal(), queue[d].raw(), ready[p]));
                    reads[p] = 1;
                    VEXB_CHECKED(vexb_copy_peer(queue[d].ordinal(), dbuf[d].raw_ptr() + at + (a - g0), queue[p].ordinal(),
                                                x(p).raw_ptr() + (a - x.part_start(p)), (b - a) * sizeof(T), queue[d].raw()));
                }
            }
        }
        void fill(unsigned d, size_t at, size_t count, T value) const {
            std::vector<T> h(count, value);
            dbuf[d].write(queue[d], at, count, h.data(), true);
        }
};

} // namespace detail

template <typename T>
class stencil : private detail::stencil_halos<T> {
        typedef detail::stencil_halos<T> Base;
    public:
        typedef T value_type;

        /// queue list, stencil values, index of the center element (stencil.hpp:179-228).
        stencil(const std::vector<backend::command_queue> &queue, const std::vector<T> &st, unsigned center)
            : Base(queue, st.size(), center) { init(st.data()); }
        template <class Iterator>
        stencil(const std::vector<backend::command_queue> &queue, Iterator begin, Iterator end, unsigned center)
            : Base(queue, static_cast<size_t>(std::distance(begin, end)), center) { std::vector<T> st(begin, end); init(st.data()); }
        stencil(const std::vector<backend::command_queue> &queue, std::initializer_list<T> list, unsigned center)
            : Base(queue, list.size(), center) { std::vector<T> st(list); init(st.data()); }

        /// y = alpha * (x * s)  or  y += alpha * (x * s)   (stencil<T>::apply, stencil.hpp:258-330).
        void apply(const vex::vector<T> &x, vex::vector<T> &y, T alpha = 1, bool append = false) const {
            precondition(x.size() == y.size() && y.nparts() == this->queue.size(), "stencil: vectors differ in size or queues");
            std::vector<const T*> left(this->queue.size(), nullptr), right(this->queue.size(), nullptr);
            this->exchange_halos(x, left, right);
            for (unsigned d = 0; d < this->queue.size(); ++d)
                VEXB_CHECKED(vexb_stencil_apply(this->queue[d].ordinal(), this->queue[d].raw(), dtype_of<T>::value, s[d].raw(),
                                                this->width, this->lhalo, x(d).raw(), x.part_size(d), left[d], right[d], y(d).raw(),
                                                static_cast<double>(alpha), append));
        }
        unsigned size() const { return static_cast<unsigned>(this->width); }
    private:
        std::vector<backend::device_vector<T>> s;
        void init(const T *st) {
            for (unsigned d = 0; d < this->queue.size(); ++d) s.emplace_back(this->queue[d], static_cast<size_t>(this->width), st);
            for (auto &q : this->queue) q.finish();
        }
};

template <typename T>
additive_operator<stencil<T>, vector<T>> operator*(const stencil<T> &s, const vector<T> &x) { return additive_operator<stencil<T>, vector<T>>(s, x); }
template <typename T>
additive_operator<stencil<T>, vector<T>> operator*(const vector<T> &x, const stencil<T> &s) { return additive_operator<stencil<T>, vector<T>>(s, x); }
template <typename T, size_t N>
additive_operator<stencil<T>, multivector<T, N>> operator*(const stencil<T> &s, const multivector<T, N> &x) { return additive_operator<stencil<T>, multivector<T, N>>(s, x); }
template <typename T, size_t N>
additive_operator<stencil<T>, multivector<T, N>> operator*(const multivector<T, N> &x, const stencil<T> &s) { return additive_operator<stencil<T>, multivector<T, N>>(s, x); }

/// User-defined stencil operator; Impl::body() is the C source of `T f(const T *X)` (stencil.hpp:510-545).
template <typename T, unsigned width, unsigned center, class Impl>
class StencilOperator : private detail::stencil_halos<T> {
        typedef detail::stencil_halos<T> Base;
    public:
        typedef T value_type;
        StencilOperator(const std::vector<backend::command_queue> &queue) : Base(queue, width, center), id(-1) {
            VEXB_CHECKED(vexb_stencil_operator_register(dtype_of<T>::value, width, center, std::string(Impl::body()).c_str(), &id));
        }
        additive_operator<StencilOperator, vector<T>> operator()(const vector<T> &x) const {
            return additive_operator<StencilOperator, vector<T>>(*this, x);
        }
        template <size_t N>
        additive_operator<StencilOperator, multivector<T, N>> operator()(const multivector<T, N> &x) const {
            return additive_operator<StencilOperator, multivector<T, N>>(*this, x);
        }
        /// y = alpha * f(x)  or  y += alpha * f(x)
        void apply(const vex::vector<T> &x, vex::vector<T> &y, T alpha = 1, bool append = false) const {
            precondition(x.size() == y.size() && y.nparts() == this->queue.size(), "stencil operator: vectors differ in size or queues");
            std::vector<const T*> left(this->queue.size(), nullptr), right(this->queue.size(), nullptr);
            this->exchange_halos(x, left, right);
            for (unsigned d = 0; d < this->queue.size(); ++d)
                VEXB_CHECKED(vexb_stencil_operator_apply(this->queue[d].ordinal(), this->queue[d].raw(), id, x(d).raw(), x.part_size(d),
                                                         left[d], right[d], y(d).raw(), static_cast<double>(alpha), append));
        }
    private:
        int id;
};

} // namespace vex

/// Declare a user-defined stencil operator type (stencil.hpp:652-657).
#define VEX_STENCIL_OPERATOR_TYPE(name, type, width, center, body_str) \
    struct name : vex::StencilOperator<type, width, center, name> { \
        name(const std::vector<vex::backend::command_queue> &q) : vex::StencilOperator<type, width, center, name>(q) {} \
        static std::string body() { return body_str; } \
    }
/// Declare a user-defined stencil operator (stencil.hpp:667-671).
#define VEX_STENCIL_OPERATOR(name, type, width, center, body, queue) \
    VEX_STENCIL_OPERATOR_TYPE(stencil_operator_##name##_t, type, width, center, body) const name(queue)

#endif
