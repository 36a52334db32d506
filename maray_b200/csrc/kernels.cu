// Hand-written sm_100a kernels of the render path:
//   maray_interp<P>   -- the bytecode interpreter (MARAY_BACKEND_INTERP)
//   fp64_issue_rate   -- FP64-pipe issue-rate microbenchmark (the roofline denominator)
// The NVRTC back end's kernel is generated at run time (codegen.cpp) from the same device_sem.cuh.
//
// Compiled with --fmad=false: the reference never fuses a*b+c.
#include <cuda_runtime.h>

#include <cstdint>

#include "bytecode.hpp"
#include "device_sem.cuh"
#include "kernels.hpp"

namespace maray {

// ------------------------------------------------------------------------------------------------
// Bytecode interpreter (bytecode.hpp, version 2).
//
// Mapping: one thread evaluates P pixels (P independent dependency chains); a block of B threads
// covers B*P consecutive pixels of the linear image index.  Per-pixel value slots live in shared
// memory as slots[slot][k][tid] (consecutive threads hit consecutive 8-byte words: conflict-free),
// the constant pool is copied to shared memory once per block (read as a broadcast).  The
// instruction stream is warp-uniform: it is staged from global memory into shared memory in
// double-buffered chunks with cp.async and every warp reads the same word (a broadcast), so there is
// no divergence anywhere in the loop.
//
// The loop is software-pipelined: while instruction i executes, instruction i+1 is fetched and BOTH
// of its operands are loaded (speculatively -- flags decide later which of them are used).  The one
// read-after-write case this cannot see, an operand stored by instruction i itself, is marked by the
// host compiler (ACC_A / FWD_B: take the accumulator instead).

constexpr int kChunk = 512;   // instructions per staged chunk (4 KiB)

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned int s = (unsigned int)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// (Measured and rejected: moving sqrt/1/x/sin/exp/ln/texture bodies out of line to shrink the loop --
// chess_1k 37.3 vs 40.6 ms, but the transcendental-heavy deep scene 28.9 vs 23.5 ms.)
//
// U = true is the row-uniform form of the bytecode (bytecode.hpp, BC_F_*_UNI): values that depend on y
// only live in n_uniform per-BLOCK words placed after the constants; every block of such a launch lies
// inside one image row (the host checks it).  Every warp computes and stores every uniform value itself
// with identical bits and uniform slots are never recycled, so no barrier is needed.  With U = false
// none of that code exists in the kernel.
template <int P, bool U>
__global__ void maray_interp(const MrParams p, const uint64_t* __restrict__ code, unsigned int n_instr,
                             const double* __restrict__ consts, unsigned int n_consts, unsigned int n_slots,
                             unsigned int n_uniform) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // layout: [2][kChunk] instruction words | staging tile (3*B*P bytes, 16-aligned) | constants | uniform slots | slots
    uint64_t* code_s = reinterpret_cast<uint64_t*>(smem_raw);
    unsigned char* stage = smem_raw + 2 * kChunk * sizeof(uint64_t);
    const unsigned int B = blockDim.x;
    const unsigned int tid = threadIdx.x;
    double* consts_s = reinterpret_cast<double*>(stage + ((3u * B * P + 15u) & ~15u));
    double* uslots = consts_s + ((n_consts + 1u) & ~1u);
    double* slots = uslots + (U ? ((n_uniform + 1u) & ~1u) : 0u);

    const unsigned int first = blockIdx.x * B * P;
    double acc[P], va[P], vb[P];
    // channel values go to three extra slots (n_slots .. n_slots+2): keeping them in registers makes the
    // compiler copy them at every handler join
    double* outs = slots + (size_t)n_slots * P * B + tid;
#pragma unroll
    for (int k = 0; k < P; k++) {
        unsigned int j = first + k * B + tid;
        unsigned int pix = p.p0 + (j < p.n ? j : 0u);   // idle lanes redo pixel p0
        unsigned int yi = pix / p.W;
        unsigned int xi = pix - yi * p.W;
        slots[(0 * P + k) * B + tid] = (double)xi;      // `x as f64`, reference src/render.rs:25
        slots[(1 * P + k) * B + tid] = (double)yi;
        acc[k] = 0.0; va[k] = 0.0; vb[k] = 0.0;
    }
    for (unsigned int i = tid; i < n_consts; i += B) consts_s[i] = consts[i];

    const unsigned int n_chunks = (n_instr + kChunk - 1) / kChunk;
    for (unsigned int i = tid; i < kChunk / 2; i += B) {   // prefetch chunk 0
        unsigned int idx = i * 2;
        if (idx < n_instr) cp_async16(code_s + idx, code + idx);
    }
    cp_async_commit();

    // operand fetch for one instruction word
    auto fetch = [&](uint64_t w, double* fa, double* fb) {
        const unsigned int a = (unsigned int)(w >> 32) & 0xffffu, b = (unsigned int)(w >> 48);
        const bool bk = (w >> 8) & BC_F_B_CONST;
        const double* pa = slots + (size_t)a * P * B + tid;
        const double* pb = bk ? consts_s + b : slots + (size_t)b * P * B + tid;
        unsigned int sa = B, sb = bk ? 0u : B;
        if constexpr (U) {
            if ((w >> 8) & BC_F_A_UNI) { pa = uslots + a; sa = 0u; }
            if ((w >> 8) & BC_F_B_UNI) { pb = uslots + b; sb = 0u; }
        }
#pragma unroll
        for (int k = 0; k < P; k++) { fa[k] = pa[k * sa]; fb[k] = pb[k * sb]; }
    };

    for (unsigned int c = 0; c < n_chunks; c++) {
        cp_async_wait_all();
        __syncthreads();   // chunk c (and, first time, the constants and X/Y slots) landed; chunk c-1 is done
        if (c + 1 < n_chunks) {
            uint64_t* dst = code_s + ((c + 1) & 1) * kChunk;
            const uint64_t* src = code + (size_t)(c + 1) * kChunk;
            unsigned int left = n_instr - (c + 1) * kChunk;
            for (unsigned int i = tid; i < kChunk / 2; i += B) {
                unsigned int idx = i * 2;
                if (idx < left) cp_async16(dst + idx, src + idx);
            }
            cp_async_commit();
        }
        const uint64_t* cs = code_s + (c & 1) * kChunk;
        const unsigned int cnt = (n_instr - c * kChunk < (unsigned)kChunk) ? (n_instr - c * kChunk) : (unsigned)kChunk;
        uint64_t w = cs[0];
        fetch(w, va, vb);                                  // the chunk's first instruction: not overlapped
#pragma unroll 2
        for (unsigned int i = 0; i < cnt; i++) {
            // ---- stage 1: next instruction's word and operands (overlaps stage 2) -----------------
            const uint64_t wn = cs[i + 1 < cnt ? i + 1 : i];
            double na[P], nb[P];
            fetch(wn, na, nb);
            // ---- stage 2: execute the current instruction -------------------------------------------
            const unsigned int lo = (unsigned int)w;
            const unsigned int op = lo & 0xffu, fl = (lo >> 8) & 0xffu;
            double x[P], y[P];
#pragma unroll
            for (int k = 0; k < P; k++) {
                const double f = (fl & BC_F_ACC_A) ? acc[k] : va[k];
                const double s = (fl & BC_F_FWD_B) ? acc[k] : vb[k];
                x[k] = (fl & BC_F_SWAP) ? s : f;
                y[k] = (fl & BC_F_SWAP) ? f : s;
            }
            if (op == BC_MUL) {
#pragma unroll
                for (int k = 0; k < P; k++) acc[k] = x[k] * y[k];
            } else if (op == BC_ADD) {
#pragma unroll
                for (int k = 0; k < P; k++) acc[k] = x[k] + y[k];
            } else if (op == BC_MOV) {
#pragma unroll
                for (int k = 0; k < P; k++) acc[k] = x[k];
            } else if (op == BC_STEP) {
#pragma unroll
                for (int k = 0; k < P; k++) acc[k] = mr_step(x[k]);
            } else if (op == BC_NEG) {
#pragma unroll
                for (int k = 0; k < P; k++) acc[k] = -x[k];
            } else if (op == BC_MIN) {
#pragma unroll
                for (int k = 0; k < P; k++) acc[k] = mr_min(x[k], y[k]);
            } else if (op == BC_MAX) {
#pragma unroll
                for (int k = 0; k < P; k++) acc[k] = mr_max(x[k], y[k]);
            } else {
                switch (op) {
                case BC_ABS:
#pragma unroll
                    for (int k = 0; k < P; k++) acc[k] = fabs(x[k]);
                    break;
                case BC_RECIP:
#pragma unroll
                    for (int k = 0; k < P; k++) acc[k] = mr_recip(x[k]);
                    break;
                case BC_SQRT:
#pragma unroll
                    for (int k = 0; k < P; k++) acc[k] = mr_sqrt(x[k]);
                    break;
                case BC_SIN:
#pragma unroll
                    for (int k = 0; k < P; k++) acc[k] = mr_sin(x[k]);
                    break;
                case BC_EXP:
#pragma unroll
                    for (int k = 0; k < P; k++) acc[k] = mr_exp(x[k]);
                    break;
                case BC_LN:
#pragma unroll
                    for (int k = 0; k < P; k++) acc[k] = mr_log(x[k]);
                    break;
                case BC_TEX: {
                    const unsigned int imm = lo >> 16;
                    const MrTexture t = p.tex[imm >> 2];
#pragma unroll
                    for (int k = 0; k < P; k++) acc[k] = mr_tex(t.data, t.w, t.h, imm & 3u, x[k], y[k]);
                } break;
                case BC_OUT_R: case BC_OUT_G: case BC_OUT_B:
#pragma unroll
                    for (int k = 0; k < P; k++) outs[((op - BC_OUT_R) * P + k) * B] = x[k];
                    break;
                default: break;   // BC_END
                }
            }
            if (fl & BC_F_STORE) {
                bool uni = false;
                if constexpr (U) uni = (fl & BC_F_ST_UNI) != 0;
                if (uni) {
                    uslots[lo >> 16] = acc[0];       // the same bits in every thread and for every k: one row
                } else {
                    double* dp = slots + (size_t)(lo >> 16) * P * B + tid;
#pragma unroll
                    for (int k = 0; k < P; k++) dp[k * B] = acc[k];
                }
            }
            w = wn;
#pragma unroll
            for (int k = 0; k < P; k++) { va[k] = na[k]; vb[k] = nb[k]; }
        }
    }

    // `as u8` + RGB pack through the staging tile, then coalesced 16-byte stores.
#pragma unroll
    for (int k = 0; k < P; k++) {
        unsigned int l = k * B + tid;   // pixel within the block
        const double o_r = outs[(0 * P + k) * B], o_g = outs[(1 * P + k) * B], o_b = outs[(2 * P + k) * B];
        stage[3u * l + 0u] = (unsigned char)mr_as_u8(o_r);
        stage[3u * l + 1u] = (unsigned char)mr_as_u8(o_g);
        stage[3u * l + 2u] = (unsigned char)mr_as_u8(o_b);
        unsigned int j = first + l;
        if (p.f64_out != nullptr && j < p.n) {
            p.f64_out[j] = o_r;
            p.f64_out[(size_t)p.f64_plane + j] = o_g;
            p.f64_out[2u * (size_t)p.f64_plane + j] = o_b;
        }
    }
    __syncthreads();
    const unsigned int span = B * P;
    const unsigned int valid = (p.n - first < span) ? (p.n - first) : span;
    unsigned char* dst = p.out + 3u * (size_t)first;
    if (p.out_aligned && valid == span && (span & 15u) == 0u) {
        for (unsigned int i = tid; i < 3u * span / 16u; i += B)
            reinterpret_cast<uint4*>(dst)[i] = reinterpret_cast<const uint4*>(stage)[i];
    } else {
        for (unsigned int i = tid; i < 3u * valid; i += B) dst[i] = stage[i];
    }
}

size_t interp_smem_bytes(unsigned int block, unsigned int pixels_per_thread, unsigned int n_slots, unsigned int n_consts,
                         unsigned int n_uniform) {
    size_t stage = (3u * (size_t)block * pixels_per_thread + 15u) & ~size_t(15);
    return 2 * kChunk * sizeof(uint64_t) + stage + (size_t)((n_consts + 1u) & ~1u) * sizeof(double) +
           (size_t)((n_uniform + 1u) & ~1u) * sizeof(double) +
           (size_t)(n_slots + 3u) * pixels_per_thread * block * sizeof(double);   // + 3 channel-output slots
}

template <int P, bool U>
static cudaError_t launch_interp_as(const MrParams& p, const uint64_t* d_code, unsigned int n_instr, const double* d_consts,
                                    unsigned int n_consts, unsigned int n_slots, unsigned int n_uniform, unsigned int block,
                                    unsigned int grid, size_t smem, cudaStream_t stream) {
    cudaError_t e = cudaFuncSetAttribute(maray_interp<P, U>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    maray_interp<P, U><<<grid, block, smem, stream>>>(p, d_code, n_instr, d_consts, n_consts, n_slots, n_uniform);
    return cudaGetLastError();
}

cudaError_t launch_interp(const MrParams& p, const uint64_t* d_code, unsigned int n_instr, const double* d_consts,
                          unsigned int n_consts, unsigned int n_slots, unsigned int block, unsigned int pixels_per_thread,
                          cudaStream_t stream, unsigned int n_uniform, bool row_uniform) {
    if (p.n == 0) return cudaSuccess;
    size_t smem = interp_smem_bytes(block, pixels_per_thread, n_slots, n_consts, row_uniform ? n_uniform : 0u);
    unsigned int span = block * pixels_per_thread;
    unsigned int grid = (p.n + span - 1) / span;
    if (row_uniform) {
        // every block must lie inside one image row, and there must be no idle lanes (they redo pixel p0)
        if (p.W % span != 0 || p.p0 % span != 0 || p.n % span != 0) return cudaErrorInvalidValue;
        switch (pixels_per_thread) {
        case 1: return launch_interp_as<1, true>(p, d_code, n_instr, d_consts, n_consts, n_slots, n_uniform, block, grid, smem, stream);
        case 2: return launch_interp_as<2, true>(p, d_code, n_instr, d_consts, n_consts, n_slots, n_uniform, block, grid, smem, stream);
        case 4: return launch_interp_as<4, true>(p, d_code, n_instr, d_consts, n_consts, n_slots, n_uniform, block, grid, smem, stream);
        default: return cudaErrorInvalidValue;
        }
    }
    switch (pixels_per_thread) {
    case 1: return launch_interp_as<1, false>(p, d_code, n_instr, d_consts, n_consts, n_slots, 0u, block, grid, smem, stream);
    case 2: return launch_interp_as<2, false>(p, d_code, n_instr, d_consts, n_consts, n_slots, 0u, block, grid, smem, stream);
    case 4: return launch_interp_as<4, false>(p, d_code, n_instr, d_consts, n_consts, n_slots, 0u, block, grid, smem, stream);
    default: return cudaErrorInvalidValue;
    }
}

// ------------------------------------------------------------------------------------------------
// FP64 issue-rate microbenchmark: 8 independent DADD/DMUL chains per thread (no FMA: the render
// path has none), or 8 DFMA chains for the nominal figure.  Counts warp-instructions * 32.
template <bool FMA>
__global__ void __launch_bounds__(256) fp64_issue_rate(double* sink, int iters, double m, double a) {
    double v0 = threadIdx.x * 1e-9, v1 = v0 + 1, v2 = v0 + 2, v3 = v0 + 3, v4 = v0 + 4, v5 = v0 + 5, v6 = v0 + 6, v7 = v0 + 7;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            if (FMA) {
                v0 = __fma_rn(v0, m, a); v1 = __fma_rn(v1, m, a); v2 = __fma_rn(v2, m, a); v3 = __fma_rn(v3, m, a);
                v4 = __fma_rn(v4, m, a); v5 = __fma_rn(v5, m, a); v6 = __fma_rn(v6, m, a); v7 = __fma_rn(v7, m, a);
            } else {
                v0 = __dmul_rn(v0, m); v1 = __dadd_rn(v1, a); v2 = __dmul_rn(v2, m); v3 = __dadd_rn(v3, a);
                v4 = __dmul_rn(v4, m); v5 = __dadd_rn(v5, a); v6 = __dmul_rn(v6, m); v7 = __dadd_rn(v7, a);
            }
        }
    }
    double s = v0 + v1 + v2 + v3 + v4 + v5 + v6 + v7;
    if (s == 123.456) sink[0] = s;   // never true; keeps the chains alive
}

cudaError_t launch_fp64_issue_rate(bool fma, double* d_sink, int iters, int blocks, cudaStream_t stream) {
    if (fma) fp64_issue_rate<true><<<blocks, 256, 0, stream>>>(d_sink, iters, 1.0000001, 1e-9);
    else fp64_issue_rate<false><<<blocks, 256, 0, stream>>>(d_sink, iters, 1.0000001, 1e-9);
    return cudaGetLastError();
}

}  // namespace maray
