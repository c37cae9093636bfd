#ifndef VEXCL_REDUCTOR_HPP
#define VEXCL_REDUCTOR_HPP
/*
 * vex::Reductor<T, RDC> (vexcl/reductor.hpp:289-439).  The reference reduces each slice to 8*SM
 * partials, copies them to the host and folds them there (:412-436).  Here each device leaves ONE
 * value in device memory (vexb_reduce_all: warp-shuffle fold, last block combines); with several
 * distinct devices the last blocks also exchange their values through peer memory over NVLink
 * inside the same kernel (ncclAllReduce if peer access is unavailable), and one 8-byte copy
 * brings the result back.  When the slices share a device (the reference's own single-GPU test
 * trick) or NCCL is unavailable, the nparts values are folded on the host in device order.
 */
#include <array>
#include <limits>
#include <map>
#include <memory>
#include <mutex>
#include "vector.hpp"

namespace vex {

struct SUM        { static const int op = VEXB_SUM; };
struct SUM_Kahan  { static const int op = VEXB_SUM_KAHAN; };
struct MAX        { static const int op = VEXB_MAX; };
struct MIN        { static const int op = VEXB_MIN; };
struct MIN_MAX    { static const int op = VEXB_MINMAX; };

namespace detail {

/// NCCL communicators for a queue list, created on first use; empty when not applicable.
struct comm_set {
    std::vector<vexb_comm*> comms;
    ~comm_set() { for (auto c : comms) vexb_comm_destroy(c); }
};
inline std::shared_ptr<comm_set> communicators(const std::vector<backend::command_queue> &queue) {
    static std::mutex mx;
    static std::map<std::vector<int>, std::shared_ptr<comm_set>> cache;
    std::vector<int> devs;
    for (auto &q : queue) devs.push_back(q.ordinal());
    std::lock_guard<std::mutex> lock(mx);
    auto it = cache.find(devs);
    if (it != cache.end()) return it->second;
    auto cs = std::make_shared<comm_set>();
    std::vector<int> sorted(devs);
    std::sort(sorted.begin(), sorted.end());
    const bool distinct = std::adjacent_find(sorted.begin(), sorted.end()) == sorted.end();
    if (devs.size() > 1 && distinct && !std::getenv("VEXCL_NO_NCCL")) {
        cs->comms.resize(devs.size(), nullptr);
        if (vexb_comm_create_all(static_cast<int>(devs.size()), devs.data(), cs->comms.data()) != VEXB_OK) cs->comms.clear();
    }
    cache[devs] = cs;
    return cs;
}

/// Peer-memory groups (mailboxes mapped between the devices) for a queue list; empty when not applicable.
struct peer_set {
    std::vector<vexb_peer*> peers;
    ~peer_set() { for (auto p : peers) vexb_peer_destroy(p); }
};
inline std::shared_ptr<peer_set> peer_group(const std::vector<backend::command_queue> &queue) {
    static std::mutex mx;
    static std::map<std::vector<int>, std::shared_ptr<peer_set>> cache;
    std::vector<int> devs;
    for (auto &q : queue) devs.push_back(q.ordinal());
    std::lock_guard<std::mutex> lock(mx);
    auto it = cache.find(devs);
    if (it != cache.end()) return it->second;
    auto ps = std::make_shared<peer_set>();
    std::vector<int> sorted(devs);
    std::sort(sorted.begin(), sorted.end());
    const bool distinct = std::adjacent_find(sorted.begin(), sorted.end()) == sorted.end();
    if (devs.size() > 1 && devs.size() <= 16 && distinct && !std::getenv("VEXCL_NO_PEER")) {
        ps->peers.resize(devs.size(), nullptr);
        if (vexb_peer_create_all(static_cast<int>(devs.size()), devs.data(), ps->peers.data()) != VEXB_OK) ps->peers.clear();
    }
    cache[devs] = ps;
    return ps;
}

template <class T, class RDC> struct reduce_result { typedef T type; static T make(const T *v) { return v[0]; } };
template <class T> struct reduce_result<T, MIN_MAX> { typedef vec2<T> type; static type make(const T *v) { type r; r.s[0] = v[0]; r.s[1] = v[1]; return r; } };

template <class T> inline T host_fold(int op, T a, T b) {
    switch (op) { case VEXB_MAX: return a > b ? a : b; case VEXB_MIN: return a < b ? a : b; default: return a + b; }
}

} // namespace detail

template <typename ScalarType, class RDC = SUM>
class Reductor {
    public:
        typedef typename detail::reduce_result<ScalarType, RDC>::type result_type;

        Reductor(const std::vector<backend::command_queue> &queue
#ifndef VEXCL_NO_STATIC_CONTEXT_CONSTRUCTORS
                = current_context().queue()
#endif
                ) : queue(queue), ws(queue.size()), res(queue.size())
        {
            for (unsigned d = 0; d < queue.size(); ++d) {
                size_t bytes = 0;
                VEXB_CHECKED(vexb_reduce_workspace_bytes(queue[d].ordinal(), &bytes));
                ws[d] = backend::device_vector<char>(queue[d], bytes);
                VEXB_CHECKED(vexb_memset(queue[d].ordinal(), ws[d].raw(), 0, bytes, queue[d].raw()));
                res[d] = backend::device_vector<ScalarType>(queue[d], 8);
            }
        }

        template <class Expr>
        typename std::enable_if<is_vector_expr<Expr>::value && detail::ncomp<Expr>::value == 0, result_type>::type
        operator()(const Expr &expr) const { return reduce(expr, -1); }

        /// Multi-expressions reduce component by component (reductor.hpp:341-349).
        template <class Expr>
        typename std::enable_if<is_vector_expr<Expr>::value && (detail::ncomp<Expr>::value > 0),
                                std::array<result_type, detail::ncomp<Expr>::value> >::type
        operator()(const Expr &expr) const {
            std::array<result_type, detail::ncomp<Expr>::value> r;
            for (size_t i = 0; i < r.size(); ++i) r[i] = reduce(expr, static_cast<int>(i));
            return r;
        }
    private:
        template <class Expr>
        result_type reduce(const Expr &expr, int comp) const {
            detail::expr_props p;
            p.comp = comp;
            expr.props(p);
            const int op = RDC::op, dt = dtype_of<ScalarType>::value;
            const int cnt = op == VEXB_MINMAX ? 2 : 1;
            ScalarType out[2] = {identity(op == VEXB_MINMAX ? VEXB_MIN : op), identity(VEXB_MAX)};
            if (!p.sized || p.size == 0) return detail::reduce_result<ScalarType, RDC>::make(out);   // reductor.hpp:318-321
            if (p.part.empty()) p.part = vex::partition(p.size, queue);                                // :323-325

            // distinct devices: the combine across GPUs happens inside the reduction kernel (peer memory)
            auto ps = queue.size() > 1 ? detail::peer_group(queue) : std::shared_ptr<detail::peer_set>();
            const bool fused = ps && !ps->peers.empty();
            for (unsigned d = 0; d < queue.size(); ++d) {
                detail::ir_builder b(d, comp);
                expr.lower(b);
                const int st = vexb_reduce_all(queue[d].ordinal(), queue[d].raw(), &b.e, dt, p.part_size(d), p.part_start(d),
                                               op, res[d].raw(), ws[d].raw(), fused ? ps->peers[d] : nullptr);
                if (st == VEXB_ERR_UNSUPPORTED && d == 0) {
                    // the expression calls a user function: evaluate it into a temporary (NVRTC side path), reduce that
                    vex::vector<typename Expr::value_type> tmp(queue, p.size);
                    detail::assign_expression<assign::SET>(tmp, expr, comp);
                    return reduce(tmp, -1);
                }
                if (st != VEXB_OK && fused && d > 0) {
                    // devices 0..d-1 have already launched and will wait for everybody in the kernel: keep the group in step
                    // (identity + the standalone combine on the devices that did not launch), then report the failure
                    const std::string why = vexb_last_error();
                    for (unsigned e = d; e < queue.size(); ++e) {
                        vexb_reduce_identity(queue[e].ordinal(), queue[e].raw(), dt, op, res[e].raw());
                        vexb_peer_allreduce(ps->peers[e], queue[e].raw(), res[e].raw(), dt, op);
                    }
                    throw backend::error(st, why);
                }
                VEXB_CHECKED(st);
            }
            auto cs = (queue.size() > 1 && !fused) ? detail::communicators(queue) : std::shared_ptr<detail::comm_set>();
            if (fused) {
                VEXB_CHECKED(vexb_reduce_fetch(queue[0].ordinal(), queue[0].raw(), res[0].raw(), dt, cnt, out));
            } else if (cs && !cs->comms.empty()) {
                std::vector<void*> bufs, streams;
                for (unsigned d = 0; d < queue.size(); ++d) { bufs.push_back(res[d].raw()); streams.push_back(queue[d].raw()); }
                VEXB_CHECKED(vexb_comm_allreduce(static_cast<int>(queue.size()), cs->comms.data(), bufs.data(), streams.data(), 1, dt, op));
                VEXB_CHECKED(vexb_reduce_fetch(queue[0].ordinal(), queue[0].raw(), res[0].raw(), dt, cnt, out));
            } else {
                for (unsigned d = 0; d < queue.size(); ++d) {
                    ScalarType v[2];
                    VEXB_CHECKED(vexb_reduce_fetch(queue[d].ordinal(), queue[d].raw(), res[d].raw(), dt, cnt, v));
                    if (op == VEXB_MINMAX) { out[0] = detail::host_fold(VEXB_MIN, out[0], v[0]); out[1] = detail::host_fold(VEXB_MAX, out[1], v[1]); }
                    else out[0] = detail::host_fold(op, out[0], v[0]);
                }
            }
            return detail::reduce_result<ScalarType, RDC>::make(out);
        }
        std::vector<backend::command_queue> queue;
        mutable std::vector<backend::device_vector<char>> ws;
        mutable std::vector<backend::device_vector<ScalarType>> res;

        static ScalarType identity(int op) {                     // reductor.hpp:55, :87, :111
            switch (op) {
                case VEXB_MAX: return std::numeric_limits<ScalarType>::lowest();
                case VEXB_MIN: return std::numeric_limits<ScalarType>::max();
                default: return ScalarType();
            }
        }
};

template <typename T, class R>
Reductor<T, R> get_reductor(const std::vector<backend::command_queue> &queue) { return Reductor<T, R>(queue); }

} // namespace vex
#endif
