// Host-side launchers of the kernels in kernels.cu.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

struct MrParams;

namespace maray {

// Dynamic shared memory the interpreter needs for a launch configuration.
size_t interp_smem_bytes(unsigned int block, unsigned int pixels_per_thread, unsigned int n_slots, unsigned int n_consts,
                         unsigned int n_uniform = 0);

// d_code must be padded to an even number of instructions (16-byte cp.async granules).
cudaError_t launch_interp(const MrParams& p, const uint64_t* d_code, unsigned int n_instr, const double* d_consts,
                          unsigned int n_consts, unsigned int n_slots, unsigned int block, unsigned int pixels_per_thread,
                          cudaStream_t stream, unsigned int n_uniform = 0, bool row_uniform = false);
// row_uniform: the bytecode is the row-uniform form (bytecode.hpp); fails with cudaErrorInvalidValue unless
// every block of the launch lies inside one image row and is full (W, p0 and n multiples of block * pixels_per_thread).

cudaError_t launch_fp64_issue_rate(bool fma, double* d_sink, int iters, int blocks, cudaStream_t stream);

}  // namespace maray
