// fp64_ubench.cu -- what the fp64 pipe of an sm_100a SM really sustains, measured.
//
// The multidim kernels are bound by fp64 issue.  Their ceiling depends on three facts that
// are not in any guide for B200 and are measured here:
//   1. warp-wide DADD / DMUL / DFMA per clock per SM sub-partition with enough independent chains,
//   2. whether instructions of OTHER pipes (integer, fp32, shared-memory loads, register moves) issue in the shadow of
//      an fp64 instruction or take issue slots away from the pipe,
//   3. how many independent fp64 instructions one scheduler needs in flight (warps x chains) to saturate the pipe.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/fp64_ubench tools/fp64_ubench.cu
// Every case runs one CTA per SM; rates are warp instructions per clock per SM sub-partition from clock64().
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x)                                                                                   \
    do                                                                                          \
    {                                                                                           \
        cudaError_t e_ = (x);                                                                   \
        if (cudaSuccess != e_)                                                                  \
        {                                                                                       \
            std::fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_));     \
            std::exit(1);                                                                       \
        }                                                                                       \
    } while (0)

constexpr int ITER = 2048;

// OP: 0 DADD, 1 DMUL, 2 DFMA.  CH independent chains per thread.
// MIX: 0 none, 1 one IADD3-class integer op per fp64 op, 2 one FFMA per fp64 op, 3 one LDS per fp64 op,
//      4 one integer op per TWO fp64 ops, 5 one MOV-class (64-bit select) per fp64 op
template <int OP, int CH, int MIX>
__global__ void __launch_bounds__(1024) k(double * out, long long * cyc, double seed, int iseed)
{
    __shared__ double sm[1024];
    sm[threadIdx.x] = seed + threadIdx.x;
    __syncthreads();
    double x[CH];
    int n[CH];
    float f[CH];
    double l[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c)
    {
        x[c] = seed + c + threadIdx.x;
        n[c] = iseed + c + threadIdx.x;
        f[c] = static_cast<float>(seed) + c + threadIdx.x;
        l[c] = 0.0;
    }
    int const tid3 = iseed + threadIdx.x;
    double const a = seed * 1.0000001;
    double const b = seed * 0.5;
    long long const t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it)
    {
#pragma unroll
        for (int c = 0; c < CH; ++c)
        {
            if (0 == OP) { asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(x[c]) : "d"(a)); }
            if (1 == OP) { asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(x[c]) : "d"(a)); }
            if (2 == OP) { asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x[c]) : "d"(a), "d"(b)); }
            if (1 == MIX || (4 == MIX && (c & 1))) { asm volatile("add.s32 %0, %0, %1;" : "+r"(n[c]) : "r"(tid3)); }
            if (2 == MIX) { asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(f[c]) : "f"(static_cast<float>(seed))); }
            if (3 == MIX) { asm volatile("ld.shared.f64 %0, [%1];" : "=d"(l[c]) : "r"(static_cast<unsigned>(__cvta_generic_to_shared(&sm[(threadIdx.x + c) & 1023])))); }
            if (5 == MIX) { asm volatile("xor.b32 %0, %0, %1;" : "+r"(n[c]) : "r"(tid3)); }
        }
    }
    long long const t1 = clock64();
    double s = 0.0;
#pragma unroll
    for (int c = 0; c < CH; ++c) { s += x[c] + n[c] + f[c] + l[c]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (0 == threadIdx.x) { cyc[blockIdx.x] = t1 - t0; }
}

template <int OP, int CH, int MIX>
void run(char const * name, int warps_per_smsp, double * out, long long * cyc, int nsm)
{
    int const threads = 128 * warps_per_smsp;
    k<OP, CH, MIX><<<nsm, threads>>>(out, cyc, 1.0, 3);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    k<OP, CH, MIX><<<nsm, threads>>>(out, cyc, 1.0, 3);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    long long h[256];
    CK(cudaMemcpy(h, cyc, sizeof(long long) * nsm, cudaMemcpyDeviceToHost));
    long long mx = 0;
    for (int i = 0; i < nsm; ++i) { mx = h[i] > mx ? h[i] : mx; }
    double const fp64_inst = static_cast<double>(ITER) * CH * warps_per_smsp; // per SMSP
    double other = 0;
    if (1 == MIX || 2 == MIX || 3 == MIX || 5 == MIX) { other = fp64_inst; }
    if (4 == MIX) { other = fp64_inst / 2; }
    std::printf("%-34s warps/smsp %d chains %2d : fp64 %.3f inst/clk/smsp, other %.3f, total %.3f  (%.1f us)\n", name,
                warps_per_smsp, CH, fp64_inst / mx, other / mx, (fp64_inst + other) / mx, ms * 1e3);
    CK(cudaEventDestroy(e0));
    CK(cudaEventDestroy(e1));
}

int main()
{
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    int const nsm = p.multiProcessorCount;
    std::printf("%s, %d SMs, clock %d kHz\n", p.name, nsm, p.clockRate);
    double * out;
    long long * cyc;
    CK(cudaMalloc(&out, sizeof(double) * 1024 * nsm));
    CK(cudaMalloc(&cyc, sizeof(long long) * nsm));
    for (int w : {1, 2, 4, 8})
    {
        run<0, 8, 0>("DADD", w, out, cyc, nsm);
        run<1, 8, 0>("DMUL", w, out, cyc, nsm);
        run<2, 8, 0>("DFMA", w, out, cyc, nsm);
    }
    for (int w : {1, 2, 3})
    {
        run<0, 1, 0>("DADD latency-bound", w, out, cyc, nsm);
        run<0, 2, 0>("DADD", w, out, cyc, nsm);
        run<0, 4, 0>("DADD", w, out, cyc, nsm);
        run<0, 16, 0>("DADD", w, out, cyc, nsm);
    }
    for (int w : {2, 4, 8})
    {
        run<0, 8, 1>("DADD + IADD 1:1", w, out, cyc, nsm);
        run<0, 8, 4>("DADD + IADD 2:1", w, out, cyc, nsm);
        run<0, 8, 2>("DADD + FFMA 1:1", w, out, cyc, nsm);
        run<0, 8, 3>("DADD + LDS.64 1:1", w, out, cyc, nsm);
        run<0, 8, 5>("DADD + LOP 1:1", w, out, cyc, nsm);
        run<2, 8, 1>("DFMA + IADD 1:1", w, out, cyc, nsm);
    }
    for (int w : {2, 3})
    {
        run<0, 2, 1>("DADD + IADD 1:1", w, out, cyc, nsm);
        run<0, 4, 1>("DADD + IADD 1:1", w, out, cyc, nsm);
        run<0, 4, 4>("DADD + IADD 2:1", w, out, cyc, nsm);
    }
    return 0;
}
