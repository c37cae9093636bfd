// Device-resident sparse strip (internal layout shared by spmv.cu and dspmat.cu).
#pragma once
#include "hostlogic.hpp"
#include <vector>

namespace vexb {
// 16-bit ELL columns are stored as distances from (row + shift of the slot): one shift per ELL slot (a 7-point stencil
// has its neighbours n*n, n, 1 away -- no single shift brings them all within 16 bits, one shift per slot does).  Slots
// past the last entry share it.
constexpr int kEllShiftSlots = 16;
struct EllShifts { int s[kEllShiftSlots]; };
// Widths for which EVERY kernel that walks an ELL strip has an unrolled instantiation (hell_kernel, hell_multi_kernel,
// dist_apply_kernel: keep their switches in step with this).  Only those strips get one shift per slot: the kernels'
// run-time loop over the slots reads shift.s[0] for all of them (indexing the by-value table at run time spills it).
constexpr bool ell_width_is_unrolled_everywhere(size_t w) { return w == 3 || w == 5 || w == 7 || w == 9; }
}

struct vexb_spmat {
    int dev = 0;
    int fmt = VEXB_FMT_CSR;
    int val_dtype = VEXB_F64;
    size_t nrows = 0, ncols = 0, nnz = 0;
    // CSR stream
    void *val = nullptr; int *col = nullptr; int *rowptr = nullptr; int2 *tile = nullptr;
    size_t n_tiles = 0, tile_nnz = 0, tile_rows = 0;
    int2 *tile_x = nullptr; size_t xwin = 0, n_windowed_tiles = 0;   // per CTA tile: {first column, length} of its x window (csr_window_kernel)
    int2 *wtile = nullptr; size_t n_wtiles = 0;   // warp tiles (<= 256 nnz, <= 256 rows) for csr_warp_kernel
    int csr_variant = 0;                          // kernel picked for this strip when spmv.kernel is not set (see build())
    size_t max_row_nnz = 0;
    // HELL
    size_t ell_width = 0, ell_pitch = 0, tail_nnz = 0;
    int *ell_col = nullptr; void *ell_val = nullptr;
    short *ell_col16 = nullptr; vexb::EllShifts ell_shifts = {};   // optional: columns as 16-bit offsets from (row + shift of the slot); see spmv.col16
    int *tail_ptr = nullptr; int *tail_col = nullptr; void *tail_val = nullptr;
    // SELL-32-sigma: slice s holds stored rows perm[32 s .. 32 s + 31] (-1: none) in sell_col / sell_val at
    // [slice_ptr[s], slice_ptr[s+1]), slot k of lane l at slice_ptr[s] + 32 k + l
    int *sell_ptr = nullptr; int *sell_perm = nullptr; int *sell_col = nullptr; void *sell_val = nullptr;
    short *sell_col16 = nullptr; int sell_shift = 0;   // optional 16-bit columns: distance from (row + sell_shift), -32768 = padding
    size_t n_slices = 0, sell_slots = 0;
    vexb_ccsr *patterns = nullptr; // VEXB_FMT_PATTERNS: the strip as unique row patterns + one pattern id per row (csrc/ccsr.cu)
    size_t n_patterns = 0;
    int *row_ids = nullptr;        // optional: compressed rows, y index of stored row r (remote strips)
    size_t y_offset = 0;           // y index of stored row 0 when the strip covers a contiguous row range
    size_t nrows_stored = 0;       // rows held in the arrays (== nrows unless row_ids)
    size_t device_bytes = 0;
    void *d_desc = nullptr;        // device copy of vexb::SpmvDesc (what a generated kernel needs to walk the rows); see jit.cu
};

namespace vexb {
// Everything a kernel generated for a VEXB_TERM_SPMV terminal reads about the strip (uniform loads through one pointer).
struct SpmvDesc {
    const void *ell_col; const void *ell_val; const int *tail_ptr; const int *tail_col; const void *tail_val;
    const int *rowptr; const int *col; const void *val;
    unsigned long long pitch; int width; int shifts[kEllShiftSlots];
};
}

namespace vexb {
// Build a strip from 32-bit host CSR.  row_ids (optional) maps stored row r to its y index;
// nrows is then the length of y and rowptr.size()-1 the number of stored rows.
int spmat_from_csr(int dev, size_t nrows, size_t ncols, std::vector<int> &rowptr, std::vector<int> &col,
                   const void *val, int val_dtype, int fmt, const std::vector<int> *row_ids, vexb_spmat **out);
}
