// Instruction-supply probe: runs a pre-built cubin whose kernel `probe` is N straight-line FP64
// instructions (8 independent DMUL/DADD chains, no loop, no memory traffic) on every SM and reports
// warp-instructions per cycle per SM.  The only thing that changes with N is the code footprint, so
// the IPC-vs-bytes curve is the chip's instruction-fetch capability for code that is never re-used
// by the same warp -- which is what a straight-line NVRTC kernel is.
//   build: nvcc -arch=sm_100a -o ifetch_probe ifetch_probe.cu -lcuda
//   run:   ifetch_probe file.cubin n_fp64_statements [blocks_per_sm=2] [waves=8]
#include <cuda.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { CUresult r_ = (x); if (r_ != CUDA_SUCCESS) { const char* s; cuGetErrorString(r_, &s); std::fprintf(stderr, "%s: %s\n", #x, s); return 1; } } while (0)

int main(int argc, char** argv) {
    if (argc < 3) { std::fprintf(stderr, "usage: %s file.cubin n_statements [blocks_per_sm] [waves]\n", argv[0]); return 2; }
    const double n_stmt = std::atof(argv[2]);
    const int per_sm = argc > 3 ? std::atoi(argv[3]) : 2, waves = argc > 4 ? std::atoi(argv[4]) : 8;
    FILE* f = std::fopen(argv[1], "rb");
    if (!f) { std::perror(argv[1]); return 1; }
    std::fseek(f, 0, SEEK_END); long sz = std::ftell(f); std::fseek(f, 0, SEEK_SET);
    std::vector<char> img(sz); if (std::fread(img.data(), 1, sz, f) != (size_t)sz) return 1; std::fclose(f);
    CK(cuInit(0));
    CUdevice dev; CK(cuDeviceGet(&dev, 0));
    CUcontext ctx; CK(cuDevicePrimaryCtxRetain(&ctx, dev)); CK(cuCtxSetCurrent(ctx));
    int sms = 0, khz = 0;
    CK(cuDeviceGetAttribute(&sms, CU_DEVICE_ATTRIBUTE_MULTIPROCESSOR_COUNT, dev));
    CK(cuDeviceGetAttribute(&khz, CU_DEVICE_ATTRIBUTE_CLOCK_RATE, dev));
    CUmodule mod; CK(cuModuleLoadData(&mod, img.data()));
    CUfunction fn; CK(cuModuleGetFunction(&fn, mod, "probe"));
    CUdeviceptr sink; CK(cuMemAlloc(&sink, 64));
    double m = 1.0000001, a = 1e-9;
    void* args[] = {&sink, &m, &a};
    const unsigned grid = unsigned(sms * per_sm * waves);
    CUevent e0, e1; CK(cuEventCreate(&e0, 0)); CK(cuEventCreate(&e1, 0));
    float best = 1e30f;
    for (int it = 0; it < 4; it++) {
        CK(cuEventRecord(e0, 0));
        CK(cuLaunchKernel(fn, grid, 1, 1, 256, 1, 1, 0, 0, args, nullptr));
        CK(cuEventRecord(e1, 0));
        CK(cuEventSynchronize(e1));
        float ms; CK(cuEventElapsedTime(&ms, e0, e1));
        if (it > 0 && ms < best) best = ms;
    }
    const double warp_instr = double(grid) * 8.0 * n_stmt;            // 8 warps per block
    const double cycles = best * 1e-3 * khz * 1e3;
    std::printf("{\"statements\": %.0f, \"code_kb\": %.0f, \"blocks_per_sm\": %d, \"ms\": %.3f, \"fp64_ipc_per_sm\": %.3f, \"frac_of_fp64_peak\": %.3f}\n",
                n_stmt, n_stmt * 16 / 1024, per_sm, best, warp_instr / cycles / sms, warp_instr / cycles / sms / 2.0);
    return 0;
}
