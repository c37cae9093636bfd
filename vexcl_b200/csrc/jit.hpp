// Run-time compilation helpers shared by the NVRTC side paths (csrc/jit.cu): expression kernels and the
// matrix-specialised CCSR kernel (csrc/ccsr.cu).
#pragma once
#include <string>
#include <cuda_runtime.h>

namespace vexb {
/// Compile `src` for sm_100a (--fmad=false) without touching a device; *cubin_bytes and *log are optional.
int jit_compile_only(const std::string &src, size_t *cubin_bytes, std::string *log);
/// Compile `src`, load it on the CURRENT device and return the entry point `name`.  Cached by (source, device).
int jit_build(int dev, const std::string &src, const char *name, void **fn);
/// cuLaunchKernel on a function returned by jit_build.
int jit_launch(void *fn, unsigned grid, unsigned block, unsigned smem, cudaStream_t st, void **args);
}
