/*
 * vexb200.h -- C ABI of libvexb200.so, the Blackwell (sm_100a) compute back end
 * that sits behind the vex:: C++ front end in include/vexcl/.
 *
 * This is the drop-in boundary.  Every entry point
 * replaces one thing the reference's hot path does through its
 * `vex::backend` layer (reference paths are relative to /root/reference):
 *
 *   lifecycle / props   <- backend::queue_list, backend::device
 *                          (vexcl/backend/cuda/context.hpp:96-155, :385-413)
 *   streams / events    <- backend::command_queue, backend::event
 *                          (vexcl/backend/cuda/context.hpp:205-260, event.hpp:54-123)
 *   memory              <- backend::device_vector<T> ctor/read/write
 *                          (vexcl/backend/cuda/device_vector.hpp:66-210)
 *   vexb_eval           <- detail::assign_expression launch loop body
 *                          (vexcl/operations.hpp:1818-1897)
 *   vexb_reduce*        <- Reductor::operator() device stage + host fold
 *                          (vexcl/reductor.hpp:302-439)
 *   vexb_partition      <- partitioning_scheme<>::get (vexcl/vector.hpp:131-167)
 *   vexb_halo_plan_*    <- SpMat::setup_exchange (vexcl/spmat.hpp:291-378)
 *   vexb_csr_create /
 *   vexb_spmv           <- SpMatCSR / SpMatHELL ctor + mul_local/mul_remote
 *                          (vexcl/spmat/csr.inl:45-209, hybrid_ell.inl:53-330)
 *   vexb_dspmat_*       <- SpMat ctor + SpMat::apply (vexcl/spmat.hpp:71-185)
 *   vexb_ccsr_*         <- SpMatCCSR ctor + its generated product function
 *                          (vexcl/spmat/ccsr.hpp:70-78, :176-201)
 *   vexb_comm_*         <- the host-staged D2H/H2D halo and the host fold of
 *                          reduction partials (spmat.hpp:149-176,
 *                          reductor.hpp:412-436), moved to NCCL over NVLink.
 *
 * Conventions: plain pointers and sizes only; every function returns 0
 * (VEXB_OK) on success and a vexb_status code otherwise and never throws;
 * vexb_last_error() returns a thread-local "file:line: message" string for
 * the last failure on the calling thread.  All `stream` arguments are
 * cudaStream_t passed as void* (NULL = the legacy default stream).  Device
 * pointers are owned by the caller unless stated otherwise.  The library is
 * thread-compatible: concurrent calls that touch different streams/handles
 * are legal, there is no shared argument stack (contrast
 * vexcl/backend/cuda/kernel.hpp:44-241).
 *
 * There is NO CPU fallback: compute entry points fail with VEXB_ERR_CUDA when
 * no CUDA device is usable.
 */
#ifndef VEXB200_H
#define VEXB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VEXB_ABI_VERSION 1

typedef enum {
    VEXB_OK = 0,
    VEXB_ERR_CUDA = 1,        /* a CUDA runtime call failed               */
    VEXB_ERR_INVALID = 2,     /* bad argument / precondition violated     */
    VEXB_ERR_NCCL = 3,        /* NCCL missing or an NCCL call failed      */
    VEXB_ERR_UNSUPPORTED = 4, /* valid request the back end cannot serve  */
    VEXB_ERR_NOMEM = 5,
    VEXB_ERR_PEER = 6         /* a peer GPU did not arrive in a fused combine / peer-memory halo */
} vexb_status;

/* Scalar element types (subset of vexcl/types.hpp:202-260). */
typedef enum {
    VEXB_F64 = 0, VEXB_F32 = 1, VEXB_I32 = 2, VEXB_U32 = 3, VEXB_I64 = 4, VEXB_U64 = 5
} vexb_dtype;

/* Assignment operators, same set and order as vexcl/operations.hpp:70-80. */
typedef enum {
    VEXB_SET = 0, VEXB_ADD = 1, VEXB_SUB = 2, VEXB_MUL = 3, VEXB_DIV = 4, VEXB_MOD = 5,
    VEXB_AND = 6, VEXB_OR = 7, VEXB_XOR = 8, VEXB_LSH = 9, VEXB_RSH = 10
} vexb_assign;

/* Reduction kinds (vexcl/reductor.hpp:47-128, :132-280). */
typedef enum {
    VEXB_SUM = 0, VEXB_SUM_KAHAN = 1, VEXB_MAX = 2, VEXB_MIN = 3,
    VEXB_MINMAX = 4 /* result[0]=min, result[1]=max, as MIN_MAX (reductor.hpp:262-280) */
} vexb_reduce_op;

/* ------------------------------------------------------------------------
 * Expression IR.  The front end lowers a vex:: expression tree to a postfix
 * program over a table of terminals; this replaces the source text the
 * reference emits in vexcl/operations.hpp:1209-1353.
 * ---------------------------------------------------------------------- */
#define VEXB_MAX_TERMS 16
#define VEXB_MAX_CODE  64
#define VEXB_MAX_STACK 12

typedef enum {
    VEXB_TERM_VEC = 0,    /* v.ptr: device array of `dtype`, element i of this device slice */
    VEXB_TERM_SCALAR = 1, /* by-value scalar of `dtype` (operations.hpp:168-175)            */
    VEXB_TERM_INDEX = 2,  /* element_index: index_offset + i + v.i64 (element_index.hpp:40-111), type u64 */
    VEXB_TERM_DSCALAR = 3,/* v.ptr: ONE device-resident value of `dtype`, broadcast to every element.  Lets the
                             result of vexb_reduce feed the next expression without a host round trip. */
    VEXB_TERM_SPMV = 4    /* v.ptr: a vexb_spmat (host handle; CSR or hybrid ELL, plain strip); pad[0]: slot of the
                             VEXB_TERM_VEC holding x.  Element i evaluates to row i of A*x (products added in storage
                             order, as vexb_spmv does): the sparse product as a *terminal* of the consumer's kernel --
                             `y = x + A*x` is one launch, y is written once and A*x never goes to memory
                             (vexcl/sparse/product.hpp:45-130, sparse/csr.hpp:102-132, sparse/ell.hpp:207-265,
                             spmat/inline_spmv.hpp:68-76).  Expressions with such terminals run on the NVRTC side path
                             (the row loop is generated into the kernel, specialised to the strip's format and width);
                             vexb_reduce does not take them. */
} vexb_term_kind;

typedef struct {
    uint8_t kind;   /* vexb_term_kind */
    uint8_t dtype;  /* vexb_dtype     */
    uint8_t pad[6];
    union { const void *ptr; double f64; float f32; int32_t i32; uint32_t u32; int64_t i64; uint64_t u64; } v;
} vexb_term;

/* Opcodes.  `type` of an instruction is the vexb_dtype the node evaluates in
 * (operands of binary ops must already have that type: the front end inserts
 * VEXB_OP_CVT following the usual arithmetic conversions).  Comparisons and
 * logical ops take operands of `type` and produce I32 0/1. */
typedef enum {
    VEXB_OP_TERM = 0,  /* push term[arg] converted to nothing: its own dtype            */
    VEXB_OP_CVT,       /* convert top from dtype `arg` to `type`                         */
    /* unary */
    VEXB_OP_NEG, VEXB_OP_LNOT,
    /* binary arithmetic / bitwise */
    VEXB_OP_ADD, VEXB_OP_SUB, VEXB_OP_MUL, VEXB_OP_DIV, VEXB_OP_MOD,
    VEXB_OP_BAND, VEXB_OP_BOR, VEXB_OP_BXOR, VEXB_OP_SHL, VEXB_OP_SHR,
    /* comparisons / logical (result I32) */
    VEXB_OP_LT, VEXB_OP_GT, VEXB_OP_LE, VEXB_OP_GE, VEXB_OP_EQ, VEXB_OP_NE,
    VEXB_OP_LAND, VEXB_OP_LOR,
    /* ternary: stack [cond(I32) a b] -> cond ? a : b  (if_else, operations.hpp:1277-1301) */
    VEXB_OP_SELECT,
    /* builtin functions (vexcl/function.hpp:287-...), floating types only unless noted */
    VEXB_OP_SIN, VEXB_OP_COS, VEXB_OP_TAN, VEXB_OP_ASIN, VEXB_OP_ACOS, VEXB_OP_ATAN,
    VEXB_OP_SINH, VEXB_OP_COSH, VEXB_OP_TANH, VEXB_OP_EXP, VEXB_OP_EXP2, VEXB_OP_LOG,
    VEXB_OP_LOG2, VEXB_OP_LOG10, VEXB_OP_SQRT, VEXB_OP_RSQRT, VEXB_OP_CBRT, VEXB_OP_FABS /* also ints: abs */,
    VEXB_OP_FLOOR, VEXB_OP_CEIL, VEXB_OP_ROUND, VEXB_OP_TRUNC,
    VEXB_OP_POW, VEXB_OP_ATAN2, VEXB_OP_FMOD, VEXB_OP_HYPOT,
    VEXB_OP_FMIN /* also ints: min */, VEXB_OP_FMAX /* also ints: max */,
    VEXB_OP_FMA,      /* ternary: a*b+c with one rounding (builtin fma)                 */
    VEXB_OP_CALL,     /* user function (VEX_FUNCTION): arg = id from vexb_function_register; pops its
                         declared number of arguments (already converted to the declared types), pushes
                         `type` = its return type.  Expressions with calls run on the NVRTC side path. */
    VEXB_OP_COUNT_
} vexb_opcode;

typedef struct {
    uint8_t  op;    /* vexb_opcode */
    uint8_t  type;  /* vexb_dtype  */
    uint16_t arg;   /* TERM: term slot; CVT: source dtype */
} vexb_instr;

typedef struct {
    int32_t    n_terms;
    int32_t    n_code;
    vexb_term  term[VEXB_MAX_TERMS];
    vexb_instr code[VEXB_MAX_CODE];
} vexb_expr;

/* ------------------------------------------------------------------------
 * Lifecycle, device properties
 * ---------------------------------------------------------------------- */
typedef struct {
    char     name[256];
    int32_t  cc_major, cc_minor;
    int32_t  sm_count;
    int32_t  max_threads_per_block;
    int32_t  warp_size;
    int32_t  pad;
    size_t   smem_per_block_optin;
    size_t   total_mem;
    size_t   l2_bytes;
} vexb_devprops;

int         vexb_abi_version(void);
const char *vexb_last_error(void);
int vexb_init(void);                       /* idempotent; fails loudly when no CUDA device */
int vexb_shutdown(void);
int vexb_device_count(int *n);
int vexb_device_props(int dev, vexb_devprops *p);
/* Tunables for experiments ("sweep.blocks_per_sm", "spmv.tile_nnz", ...). */
int vexb_set_param(const char *name, long value);
int vexb_get_param(const char *name, long *value);
/* Count of kernels this library has launched from the calling process. */
int vexb_launch_count(uint64_t *n);

/* ------------------------------------------------------------------------
 * Streams and events
 * ---------------------------------------------------------------------- */
int vexb_stream_create(int dev, void **stream);
int vexb_stream_destroy(int dev, void *stream);
int vexb_stream_sync(int dev, void *stream);
int vexb_device_sync(int dev);
int vexb_event_create(int dev, void **event);
int vexb_event_destroy(int dev, void *event);
int vexb_event_record(int dev, void *event, void *stream);
int vexb_event_sync(int dev, void *event);
int vexb_stream_wait_event(int dev, void *stream, void *event);
int vexb_event_elapsed_ms(void *start, void *stop, float *ms);

/* ------------------------------------------------------------------------
 * Memory
 * ---------------------------------------------------------------------- */
int vexb_malloc(int dev, size_t bytes, void **p);
int vexb_free(int dev, void *p);
int vexb_host_alloc(size_t bytes, void **p);   /* pinned host memory */
int vexb_host_free(void *p);
int vexb_h2d(int dev, void *dst, const void *src, size_t bytes, void *stream, int blocking);
int vexb_d2h(int dev, void *dst, const void *src, size_t bytes, void *stream, int blocking);
int vexb_d2d(int dev, void *dst, const void *src, size_t bytes, void *stream);
int vexb_memset(int dev, void *dst, int byte, size_t bytes, void *stream);

/* ------------------------------------------------------------------------
 * Partitioning (host only; no GPU needed)
 * vexcl/vector.hpp:131-167 with util.hpp:91-93 (alignup 16).
 * weights == NULL means equal weights (vector.hpp:79-81).
 * part must hold nparts+1 entries.
 * ---------------------------------------------------------------------- */
int vexb_partition(size_t n, int nparts, const double *weights, size_t *part);

/* ------------------------------------------------------------------------
 * Elementwise: lhs[i] OP= expr(i), i in [0,n) of one device slice.
 * Replaces the per-device kernel launch of assign_expression
 * (operations.hpp:1886-1895).  Asynchronous on `stream`.
 * ---------------------------------------------------------------------- */
int vexb_eval(int dev, void *stream, void *lhs, int lhs_dtype, int assign_op,
              const vexb_expr *expr, size_t n, size_t index_offset);
/* User-defined device functions (VEX_FUNCTION, vexcl/function.hpp:225): `body` is C source that refers to its
 * arguments as prm1, prm2, ... (the reference's convention) and returns a value of `ret_dtype`.
 * Such functions cannot be pre-compiled: an expression that calls one is turned into CUDA source
 * (the sweep skeleton with the expression inlined), compiled once with NVRTC for sm_100a, cached, and launched
 * through the driver API.  Setting the tunable "eval.jit" = 1 sends every non-sweep expression down the same
 * path instead of the interpreter.  Registering the same definition twice returns the same id. */
int vexb_function_register(const char *name, int ret_dtype, int nargs, const int *arg_dtypes,
                           const char *body, int *id);
/* Generated source and NVRTC build log of the kernel vexb_eval would JIT for this request (for inspection
 * and for tests on machines without a GPU: NVRTC needs no device).  Two-call pattern on *len. */
int vexb_jit_source(int lhs_dtype, int assign_op, const vexb_expr *expr, char *buf, size_t *len, int compile);
/* The same for the one kernel vexb_eval_multi generates for ncomp (2..8) components. */
int vexb_jit_source_multi(int lhs_dtype, int assign_op, int ncomp, const vexb_expr *const *exprs, char *buf, size_t *len, int compile);
/* Expressions without a hand-written sweep are served by the pre-compiled interpreter while NVRTC builds a kernel
 * specialised to the expression on a background thread (started at the first use of a new expression shape; tunable
 * "eval.jit": 0 = interpreter only, 1 = compile synchronously, 2 = this, the default).  *pending = 1 while any such
 * compilation is still running: benchmarks and tests wait on it to time / check the specialised kernel. */
int vexb_jit_pending(int *pending);
/* Compile the kernel of a request shape before its first use (background != 0: on a background thread, as the first
 * vexb_eval of a new shape does in the default mode).  Needs NVRTC, no device.  A process may exit while background
 * compilations run: the thread that started them waits for the one in flight and cancels the rest on its way out. */
int vexb_jit_precompile(int lhs_dtype, int assign_op, const vexb_expr *expr, int background);
/* All components of a multi-expression assignment in one launch (assign_multiexpression, operations.hpp:2081-2185):
 * lhs[k][i] OP= exprs[k](i), k < ncomp <= 8, all of one type; every right-hand side of element i is evaluated before
 * any left-hand side of element i is written.  *handled = 0 when the request is not served here (NVRTC missing or
 * still compiling in the background, sparse-product terminals): evaluate component by component then. */
int vexb_eval_multi(int dev, void *stream, int ncomp, void *const *lhs, int lhs_dtype, int assign_op,
                    const vexb_expr *const *exprs, size_t n, size_t index_offset, int *handled);
/* Which kernel vexb_eval would take for this request: writes a short
 * name ("sweep:muladd", "interp", "jit", ...) to buf. */
int vexb_eval_path(int lhs_dtype, int assign_op, const vexb_expr *expr, char *buf, size_t buflen);

/* ------------------------------------------------------------------------
 * Reduction of an expression over one device slice to ONE device-resident
 * value (two for VEXB_MINMAX) of type `dtype`, written to d_result.
 * d_workspace: at least vexb_reduce_workspace_bytes() bytes of device memory
 * private to the (device, stream) in use; it must be zero-initialised once
 * (vexb_memset) before first use and is left reusable after each call.
 * Replaces reductor.hpp:327-410; the host fold :412-436 is replaced by
 * vexb_reduce_fetch (single device) or vexb_comm_allreduce (+ fetch).
 * ---------------------------------------------------------------------- */
typedef struct vexb_peer vexb_peer;   /* a group of GPUs that write each other's memory; see "Peer memory" below */
int vexb_reduce_workspace_bytes(int dev, size_t *bytes);
int vexb_reduce(int dev, void *stream, const vexb_expr *expr, int dtype, size_t n,
                size_t index_offset, int op, void *d_result, void *d_workspace);
/* Several reductions of the SAME expression in one pass over memory: vex::CombineReductors<R...> (reductor.hpp:132-280).
 * ops: nops <= VEXB_MAX_COMBINED of VEXB_SUM / VEXB_SUM_KAHAN / VEXB_MAX / VEXB_MIN; d_result receives nops values of
 * `dtype` in that order; d_workspace must hold nops * vexb_reduce_workspace_bytes() zero-initialised bytes.  With a
 * peer group every value is combined across the GPUs inside the kernel, as in vexb_reduce_all. */
#define VEXB_MAX_COMBINED 16
int vexb_reduce_multi(int dev, void *stream, const vexb_expr *expr, int dtype, size_t n, size_t index_offset,
                      int nops, const int *ops, void *d_result, void *d_workspace, vexb_peer *peer);
/* Fill d_result with the identity of `op` (reductor.hpp:55,87,111): used for empty slices. */
int vexb_reduce_identity(int dev, void *stream, int dtype, int op, void *d_result);
/* D2H of `count` values + stream sync. */
int vexb_reduce_fetch(int dev, void *stream, const void *d_result, int dtype, int count, void *host_out);

/* ------------------------------------------------------------------------
 * Communication (NCCL over NVLink).  NCCL is dlopen()ed on first use.
 *   single process, n devices : vexb_comm_create_all
 *   one process per device    : rank 0 calls vexb_comm_unique_id, the 128-byte
 *                               id is broadcast by the host program, every
 *                               rank calls vexb_comm_create_rank
 * ---------------------------------------------------------------------- */
typedef struct vexb_comm vexb_comm;
#define VEXB_UNIQUE_ID_BYTES 128
int vexb_comm_unique_id(void *id128);
int vexb_comm_create_rank(int dev, int nranks, int rank, const void *id128, vexb_comm **comm);
int vexb_comm_create_all(int ndev, const int *devs, vexb_comm **comms /* ndev out */);
int vexb_comm_destroy(vexb_comm *comm);
int vexb_comm_rank(const vexb_comm *comm, int *rank, int *nranks, int *dev);
/* In-place all-reduce of `count` values on each local device.  op: VEXB_SUM / VEXB_MAX / VEXB_MIN. */
int vexb_comm_allreduce(int nlocal, vexb_comm *const *comms, void *const *bufs, void *const *streams,
                        int count, int dtype, int op);
int vexb_comm_barrier(int nlocal, vexb_comm *const *comms, void *const *streams);

/* ------------------------------------------------------------------------
 * CUDA graphs: record everything enqueued on `stream` (and on streams that
 * fork from / join it through events, e.g. the halo side stream and its NCCL
 * calls) between begin and end, then replay it with one launch.  Replaces the
 * per-kernel host launch loop of the reference for launch-bound inner loops
 * (a CG iteration at 8 GPUs).  Only asynchronous entry points may be called
 * while capturing (no blocking copies, no vexb_reduce_fetch).
 * ---------------------------------------------------------------------- */
typedef struct vexb_graph vexb_graph;
int vexb_graph_begin(int dev, void *stream);
int vexb_graph_end(int dev, void *stream, vexb_graph **graph);
int vexb_graph_launch(vexb_graph *graph, void *stream);
int vexb_graph_destroy(vexb_graph *graph);

/* ------------------------------------------------------------------------
 * Peer memory: a group of ranks (one per GPU) whose kernels write into each
 * other's device memory over NVLink.  Used to fuse the combine of
 * Reductor across GPUs into the reduction kernel itself (vexb_reduce_all):
 * the last block of every rank pushes its value into all peers' mailboxes
 * and folds what it received, in rank order -- one kernel, no NCCL call, no
 * host (replaces reductor.hpp:412-436).
 *   one process per GPU : vexb_peer_create (returns a 64-byte CUDA IPC handle),
 *                         the launcher all-gathers the handles (rank order),
 *                         vexb_peer_connect maps the other ranks' mailboxes;
 *   one process, n GPUs : vexb_peer_create_all (peer access, distinct devices).
 * A rank that never shows up makes the waiting kernel give up after ~20 s:
 * it stores NaN / all-ones instead of a partial fold and raises a sticky fault
 * (vexb_peer_error, vexb_peer_fault) instead of hanging.
 * ---------------------------------------------------------------------- */
#define VEXB_IPC_HANDLE_BYTES 64
int vexb_peer_create(int dev, int rank, int nranks, vexb_peer **peer, void *handle64);
int vexb_peer_connect(vexb_peer *peer, const void *handles /* nranks * 64 bytes, rank order */);
int vexb_peer_create_all(int ndev, const int *devs, vexb_peer **peers /* ndev out */);
int vexb_peer_destroy(vexb_peer *peer);
int vexb_peer_error(vexb_peer *peer, unsigned long long *epoch_of_timeout /* 0 = none */);
/* Process-wide sticky fault: non-zero once any kernel of this process (fused reduction combine, peer-memory halo)
 * gave up waiting for a peer.  Such a kernel never folds or multiplies stale data: the values that needed the peer
 * are written as NaN (floating types) or all-ones (integers).  vexb_reduce_fetch returns VEXB_ERR_PEER while the
 * fault is set; the front ends turn that into an exception.  clear != 0 resets it (after the group was rebuilt). */
int vexb_peer_fault(unsigned long long *epoch, int clear);
/* In-place all-reduce of one value (two for VEXB_MINMAX) per rank. */
int vexb_peer_allreduce(vexb_peer *peer, void *stream, void *d_buf, int dtype, int op);
/* vexb_reduce + the combine across the peer group in the same kernel; every rank ends with the
 * same bits in d_result.  peer == NULL (or a group of one) is plain vexb_reduce. */
int vexb_reduce_all(int dev, void *stream, const vexb_expr *expr, int dtype, size_t n,
                    size_t index_offset, int op, void *d_result, void *d_workspace, vexb_peer *peer);

/* ------------------------------------------------------------------------
 * Halo plan (host only; no GPU needed): who sends which x entries to whom.
 * Input: column partition and, for every part d, the sorted unique list of
 * global column ids outside [col_part[d], col_part[d+1]) referenced by the
 * rows of part d (the `ghost_cols[d]` sets of spmat.hpp:300-316),
 * concatenated, with ghost_off[d]..ghost_off[d+1] delimiting part d.
 * ---------------------------------------------------------------------- */
typedef struct vexb_halo_plan vexb_halo_plan;
/* Ghost columns of one row strip: two-call pattern (out == NULL returns the count). */
int vexb_strip_ghost_cols(size_t nrows, const void *ptr, int ptr_bytes, const void *col, int col_bytes,
                          size_t col_begin, size_t col_end, int64_t *out, size_t *count);
int vexb_halo_plan_create(int nparts, const size_t *col_part, const int64_t *ghost_cols,
                          const size_t *ghost_off, vexb_halo_plan **plan);
int vexb_halo_plan_destroy(vexb_halo_plan *plan);
/* Reference-equivalent tables (spmat.hpp:319-371), for parity checks:
 * cols_to_send: global sorted union (n_send_total entries) with the owner's
 * col_part start subtracted; cidx: nparts+1 offsets into it. */
int vexb_halo_plan_ref_sizes(const vexb_halo_plan *plan, size_t *n_send_total);
int vexb_halo_plan_ref_tables(const vexb_halo_plan *plan, int64_t *cols_to_send, size_t *cidx);
int vexb_halo_plan_ref_recv(const vexb_halo_plan *plan, int part, int64_t *cols_to_recv /* n_ghost(part) */);
/* Pairwise form used by the exchange: for `part`, send_counts[p] values go to
 * peer p (send_cols lists the local x indices, grouped by ascending p);
 * recv_counts[p] values arrive from peer p and land contiguously, in
 * ascending p order, in the ghost buffer (which is ordered like the sorted
 * ghost list, exactly the renumbering of csr.inl:92-96). */
int vexb_halo_plan_counts(const vexb_halo_plan *plan, int part, size_t *send_counts, size_t *recv_counts);
int vexb_halo_plan_send_cols(const vexb_halo_plan *plan, int part, int64_t *send_cols);

/* ------------------------------------------------------------------------
 * Sparse strips (one device).  Input: host CSR with column ids already local
 * to the strip's x (ptr[0] may be non-zero; it is subtracted).  Index width
 * on the device is 32-bit whenever that is lossless.
 * ---------------------------------------------------------------------- */
typedef struct vexb_spmat vexb_spmat;
typedef enum {
    VEXB_FMT_AUTO = 0,  /* CSR row-block stream kernel unless ELL is clearly better */
    VEXB_FMT_CSR = 1,   /* row-block CSR, tiles staged through shared memory by TMA bulk copies */
    VEXB_FMT_HELL = 2,  /* hybrid ELL + CSR tail, width by hybrid_ell.inl:66-114 */
    VEXB_FMT_PATTERNS = 3,/* the strip's unique rows (column offsets from the diagonal + values) and one pattern id per
                             row: the reference's CCSR (spmat/ccsr.hpp) found automatically.  Used when the strip has at
                             most "spmv.max_patterns" (256) distinct rows and no row map; otherwise as VEXB_FMT_AUTO.
                             Opt-in in round 1 (not yet run on a GPU). */
    VEXB_FMT_SELL = 4     /* sliced ELL (SELL-32-sigma): slices of 32 rows stored column-major at the width of their longest
                             row, rows sorted by length inside windows of "spmv.sell_sigma" (1024) rows so that a slice
                             holds rows of similar length.  One warp per slice, one lane per row, coalesced loads, no
                             shared memory, products added in storage order (same bits as the reference loop).  What
                             VEXB_FMT_AUTO picks for rows too irregular for hybrid ELL. */
} vexb_spfmt;
/* Host-only: the row patterns VEXB_FMT_PATTERNS would use.  *n_patterns = number of distinct rows; idx (optional,
 * nrows entries) = pattern of each row.  Returns VEXB_ERR_UNSUPPORTED when there are more than max_patterns. */
int vexb_csr_row_patterns(size_t nrows, const void *ptr, int ptr_bytes, const void *col, int col_bytes,
                          const void *val, int val_dtype, size_t max_patterns, size_t *n_patterns, int32_t *idx);

/* Host-only: the sliced-ELL layout VEXB_FMT_SELL would use.  perm (optional, 32 * n_slices entries): the row each slice
 * lane multiplies, -1 = none; slice_ptr (optional, n_slices + 1): first slot of each slice.  sigma = sorting window. */
int vexb_csr_sell_layout(size_t nrows, const void *ptr, int ptr_bytes, long sigma, size_t *n_slices, size_t *n_slots,
                         int32_t *perm, int32_t *slice_ptr);

int vexb_csr_create(int dev, void *stream, size_t nrows, size_t ncols,
                    const void *ptr, int ptr_bytes, const void *col, int col_bytes,
                    const void *val, int val_dtype, int fmt, vexb_spmat **out);
int vexb_spmat_destroy(vexb_spmat *A);
typedef struct {
    size_t  nrows, ncols, nnz;
    int32_t fmt;          /* VEXB_FMT_CSR, VEXB_FMT_HELL, VEXB_FMT_PATTERNS or VEXB_FMT_SELL */
    int32_t val_dtype;
    size_t  ell_width, ell_pitch, csr_tail_nnz;   /* HELL only */
    size_t  n_tiles, tile_nnz;                    /* CSR: tiles, nnz per tile; PATTERNS: unique rows, entries in their table */
    size_t  device_bytes;                         /* bytes of matrix data resident in HBM */
} vexb_spmat_info;
int vexb_spmat_get_info(const vexb_spmat *A, vexb_spmat_info *info);
/* Copy the HELL arrays back (parity with hybrid_ell.inl:132-193); any pointer may be NULL. */
int vexb_spmat_hell_download(const vexb_spmat *A, int32_t *ell_col, void *ell_val,
                             int64_t *csr_ptr, int32_t *csr_col, void *csr_val);
/* y (=|+=) alpha * A x     (csr.inl:188-209: append ? "+=" : "=") */
int vexb_spmv(int dev, void *stream, const vexb_spmat *A, const void *x, void *y, double alpha, int append);

/* ------------------------------------------------------------------------
 * Compressed CSR for stencil-like matrices: vex::SpMatCCSR
 * (vexcl/spmat/ccsr.hpp:54-86 ctor, :176-201 kernel).  Single device.
 *   y[i] (=|+=) alpha * sum_{j = row[idx[i]] .. row[idx[i]+1]} val[j] * x[i + col[j]]
 * idx: n entries naming one of the m unique rows; row: m+1 offsets; col: SIGNED
 * offsets from the diagonal.  Unlike the reference, create() rejects a matrix
 * whose rows reach outside [0, n).
 * ---------------------------------------------------------------------- */
typedef struct vexb_ccsr vexb_ccsr;
typedef struct {
    size_t  nrows, unique_rows, nnz;
    int32_t idx_bytes;       /* width idx was re-encoded to on the device (1, 2 or 4) */
    int32_t table_in_smem;   /* unique-row table staged in shared memory by the kernel */
    size_t  device_bytes;
} vexb_ccsr_info;
int vexb_ccsr_create(int dev, void *stream, size_t n, size_t m, const void *idx, int idx_bytes,
                     const void *row, int row_bytes, const void *col, int col_bytes,
                     const void *val, int val_dtype, vexb_ccsr **out);
int vexb_ccsr_destroy(vexb_ccsr *A);
int vexb_ccsr_get_info(const vexb_ccsr *A, vexb_ccsr_info *info);
int vexb_ccsr_spmv(int dev, void *stream, const vexb_ccsr *A, const void *x, void *y, double alpha, int append);
/* Source of the matrix-specialised product kernel (tunable "ccsr.jit"; the unique rows become code, compiled by
 * NVRTC at first use -- the counterpart of the reference's generated "<prm>_spmv" function, ccsr.hpp:176-201).
 * Host-only: m unique rows, row[m+1] offsets, col/val entries; idx_bytes = device width of idx (1, 2 or 4).
 * compile != 0 also runs NVRTC for sm_100a (no device needed).  *len in: capacity, out: bytes needed. */
int vexb_ccsr_jit_source(size_t m, const int32_t *row, const int32_t *col, const void *val, int val_dtype,
                         int idx_bytes, char *buf, size_t *len, int compile);

/* ------------------------------------------------------------------------
 * Stencil convolution: vex::stencil<T> (vexcl/stencil.hpp:168-330), one device
 * slice per call:
 *   y[i] (=|+=) alpha * sum_{k<width} s[k] * X(i + k - center)
 * X = the slice x[0..n) extended by `left` (the `center` elements before it)
 * and `right` (the width-1-center elements after it); a NULL side clamps to the
 * slice's first / last element (the ends of the whole vector).  s, x, left,
 * right, y are device pointers of `dtype` (VEXB_F64 or VEXB_F32).
 * vexb_copy_peer moves halo pieces between devices (replaces the D2H/H2D
 * staging of stencil_base::exchange_halos, stencil.hpp:86-150).
 * ---------------------------------------------------------------------- */
int vexb_stencil_apply(int dev, void *stream, int dtype, const void *s, int width, int center,
                       const void *x, size_t n, const void *left, const void *right,
                       void *y, double alpha, int append);
int vexb_copy_peer(int dst_dev, void *dst, int src_dev, const void *src, size_t bytes, void *stream);
/* User-defined stencil operators: vex::StencilOperator / VEX_STENCIL_OPERATOR (stencil.hpp:510-680).
 *   y[i] (=|+=) alpha * f(X),  X[k] = the element k places from i (k in [-center, width-1-center]); `body` is the C
 * source of f with X a pointer into the staged window, e.g. "return sin(X[1] - X[0]) + sin(X[0] - X[-1]);".
 * register() is host-only and returns the same id for the same definition; the kernel is compiled by NVRTC at the first
 * apply() on a device (needs libnvrtc; there is no pre-compiled alternative).  source() returns the generated kernel
 * (compile != 0: also runs NVRTC for sm_100a, no device needed).  Opt-in in round 1: not yet run on a GPU. */
int vexb_stencil_operator_register(int dtype, int width, int center, const char *body, int *id);
int vexb_stencil_operator_source(int id, char *buf, size_t *len, int compile);
int vexb_stencil_operator_apply(int dev, void *stream, int id, const void *x, size_t n, const void *left,
                                const void *right, void *y, double alpha, int append);

/* ------------------------------------------------------------------------
 * Distributed SpMat part: the slice of a vex::SpMat owned by one device
 * (spmat.hpp:71-106 ctor body for one d, :120-185 apply).
 * `col` holds GLOBAL column ids for the strip's rows, indexed by ptr values
 * relative to ptr[0].
 * ---------------------------------------------------------------------- */
typedef struct vexb_dspmat vexb_dspmat;
int vexb_dspmat_create(int dev, void *stream, int part, const vexb_halo_plan *plan,
                       size_t nrows, const void *ptr, int ptr_bytes, const void *col, int col_bytes,
                       const void *val, int val_dtype, int fmt, vexb_dspmat **out);
int vexb_dspmat_destroy(vexb_dspmat *A);
typedef struct {
    size_t nrows, ncols_local, n_ghost, n_send, loc_nnz, rem_nnz;
    vexb_spmat_info loc, rem;
} vexb_dspmat_info;
int vexb_dspmat_get_info(const vexb_dspmat *A, vexb_dspmat_info *info);
/* Split tables back on the host for parity with csr.inl:70-112 (any pointer may be NULL). */
int vexb_dspmat_download_split(const vexb_dspmat *A, int64_t *loc_ptr, int64_t *loc_col, void *loc_val,
                               int64_t *rem_ptr, int64_t *rem_col, void *rem_val);
/* The part's strip for use as a VEXB_TERM_SPMV terminal: set when the part has no ghost columns and its rows are
 * stored plainly in CSR or hybrid ELL (then row i of the strip is element i of the part's slice); NULL otherwise. */
int vexb_dspmat_inline_strip(const vexb_dspmat *A, const vexb_spmat **strip);
void *vexb_dspmat_send_buffer(const vexb_dspmat *A);  /* device, n_send values  */
void *vexb_dspmat_ghost_buffer(const vexb_dspmat *A); /* device, n_ghost values */
/* Steps of SpMat::apply, all asynchronous on `stream`: */
int vexb_dspmat_pack(const vexb_dspmat *A, void *stream, const void *x);                       /* spmat.hpp:127-135 */
int vexb_dspmat_mul_local(const vexb_dspmat *A, void *stream, const void *x, void *y, double alpha, int append); /* :142-146 */
int vexb_dspmat_mul_remote(const vexb_dspmat *A, void *stream, void *y, double alpha);         /* :178-183 */
/* Halo exchange for the local parts (grouped ncclSend/ncclRecv): send buffers -> peers' ghost buffers.
 * Replaces spmat.hpp:149-176. */
int vexb_halo_exchange(int nlocal, vexb_comm *const *comms, vexb_dspmat *const *parts, void *const *streams);
/* Peer-memory halo (csrc/distapply.cu): when every part is connected, vexb_dspmat_apply is ONE kernel per GPU -- it
 * stores the x values its neighbours need straight into their ghost buffers over NVLink, multiplies the interior rows,
 * waits (in the kernel) for its own ghosts and finishes the boundary rows; no NCCL call, no copy, no extra launch, CUDA-
 * graph replayable.  Replaces spmat.hpp:127-183 in one launch.  Up to 16 parts on distinct devices.
 *   one process per GPU : vexb_dspmat_halo_handle (64-byte CUDA IPC handle of this part's ghost box), all-gather the
 *                         handles in part order, vexb_dspmat_halo_connect;
 *   one process, n GPUs : vexb_dspmat_halo_connect_local (peer access).
 * Every part of the matrix must then call apply the same number of times (as with NCCL).  A neighbour that never
 * arrives makes the kernel give up after ~20 s, store NaN in the rows it could not finish and raise vexb_peer_fault. */
int vexb_dspmat_halo_handle(vexb_dspmat *A, void *handle64);
int vexb_dspmat_halo_connect(vexb_dspmat *A, const void *handles /* nparts * 64 bytes, part order */);
int vexb_dspmat_halo_connect_local(int nlocal, vexb_dspmat *const *parts);
int vexb_dspmat_halo_connected(const vexb_dspmat *A, int *connected);
int vexb_dspmat_halo_disconnect(vexb_dspmat *A);   /* back to NCCL / copies (e.g. when another rank failed to connect) */
/* Whole apply for the local parts: pack -> (side stream: exchange) || mul_local -> mul_remote.
 * x[k], y[k] are the device slices of part k.  comms may be NULL when there are no ghosts. */
int vexb_dspmat_apply(int nlocal, vexb_comm *const *comms, vexb_dspmat *const *parts, void *const *streams,
                      const void *const *x, void *const *y, double alpha, int append);

/* Several right-hand sides in one pass over the matrix: vex::SpMat * vex::multivector<T,N> (multivector.hpp;
 * the reference multiplies component by component, operations.hpp:876-880).  Hybrid-ELL strips take groups of up to 4
 * vectors per launch (columns and values loaded once, same per-component bits as vexb_spmv); anything else falls back to
 * one product per vector.  vexb_dspmat_apply_multi: x[k * nrhs + r] / y[k * nrhs + r] = slice of component r on part k;
 * parts with a halo multiply component by component. */
int vexb_spmv_multi(int dev, void *stream, const vexb_spmat *A, int nrhs, const void *const *x, void *const *y,
                    double alpha, int append);
int vexb_dspmat_apply_multi(int nlocal, vexb_comm *const *comms, vexb_dspmat *const *parts, void *const *streams,
                            int nrhs, const void *const *x, void *const *y, double alpha, int append);

/* The product and a dot product with its result without re-reading the vectors: y (=|+=) alpha*A*x, and
 * d_result[k][0] = sum over all parts of dot_with . y (same bits on every GPU).  The product kernel leaves one partial
 * per block; a one-block second launch folds them and combines across the peer group.  With dot_with = x this is q = A*p, (p, q) of a CG iteration.  Needs the peer-memory halo on every part
 * (or a single part) and a hybrid-ELL interior strip; returns VEXB_ERR_UNSUPPORTED otherwise (compose apply + reduce). */
int vexb_dspmat_apply_dot(int nlocal, vexb_dspmat *const *parts, void *const *streams, const void *const *x,
                          void *const *y, double alpha, int append, const void *const *dot_with,
                          void *const *d_result, vexb_peer *const *peers);

/* The vector half of a conjugate-gradient iteration (BASELINE configs[4]) in two sweeps, scalars device-resident:
 *   vexb_cg_update_r  : alpha = *d_rho / *d_pq;  r -= alpha q;  *d_rho_new = (r, r)   (folded like vexb_reduce, combined
 *                       over `peer` in the same kernel; peer == NULL: this slice only)             24 bytes per row
 *   vexb_cg_update_xp : alpha as above, beta = *d_rho_new / *d_rho;  x += alpha p;  p = r + beta p  40 bytes per row
 * Same unfused per-element arithmetic as the vexb_eval / vexb_reduce composition (viennacl.hpp:36-64 composes CG from
 * those); d_workspace as for vexb_reduce.  Vectors must be 32-byte aligned (vexb_malloc's are). */
int vexb_cg_update_r(int dev, void *stream, int dtype, size_t n, void *r, const void *q,
                     const void *d_rho, const void *d_pq, void *d_rho_new, void *d_workspace, vexb_peer *peer);
int vexb_cg_update_xp(int dev, void *stream, int dtype, size_t n, void *x, void *p, const void *r,
                      const void *d_rho, const void *d_pq, const void *d_rho_new);

#ifdef __cplusplus
}
#endif
#endif /* VEXB200_H */
