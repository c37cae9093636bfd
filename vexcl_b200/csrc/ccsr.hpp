// Internal entry point of the CCSR code shared with spmv.cu (row-pattern strips of vex::SpMat).
#pragma once
#include <cstddef>
#include "../../include/vexb200.h"

namespace vexb {
/// vexb_ccsr_create for a matrix whose rows index a vector of `xlen` elements (the public call has xlen = n).
int ccsr_create_ex(int dev, size_t n, size_t xlen, size_t m, const void *idx, int idx_bytes,
                   const void *row, int row_bytes, const void *col, int col_bytes,
                   const void *val, int val_dtype, vexb_ccsr **out);
}
