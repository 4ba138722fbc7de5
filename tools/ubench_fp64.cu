// ubench_fp64.cu — fp64 pipe micro-benchmarks on B200 that size the trace kernel's latency model:
//   1. dependent-issue latency of DFMA / DMUL / DADD and of the MUFU.RCP64H / RSQ64H seeds (1 warp, clock64)
//   2. DFMA throughput per SM sub-partition as a function of resident warps x independent chains per thread
// Build + run (GPU box):  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ubench tools/ubench_fp64.cu && /tmp/ubench
#include <cstdio>
#include <cuda_runtime.h>

#define CHAIN 4096

template <int OP>
__global__ void lat_kernel(double* out, long long* cyc, double a, double b)
{
    double x = a + threadIdx.x*1e-9;
    long long t0 = clock64();
#pragma unroll 64
    for (int i = 0; i < CHAIN; ++i) {
        if (OP == 0) x = __fma_rn(x, b, a);
        else if (OP == 1) x = __dmul_rn(x, b);
        else if (OP == 2) x = __dadd_rn(x, b);
        else if (OP == 3) { double y; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x)); x = y; }
        else if (OP == 4) { double y; asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x)); x = y; }
        else if (OP == 5) x = sqrt(x) + b;
        else if (OP == 6) x = a/x;
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

// ILP independent DFMA chains per thread
template <int ILP>
__global__ void thr_kernel(double* out, double a, double b, int iters)
{
    double x[ILP];
#pragma unroll
    for (int k = 0; k < ILP; ++k) x[k] = a + (threadIdx.x + k)*1e-9;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u)
#pragma unroll
            for (int k = 0; k < ILP; ++k) x[k] = __fma_rn(x[k], b, a);
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < ILP; ++k) s += x[k];
    out[blockIdx.x*blockDim.x + threadIdx.x] = s;
}

template <int ILP>
void run_thr(double* out, int warps_per_smsp, int sms)
{
    const int iters = 4096;
    const int threads = warps_per_smsp*4*32;     // one block per SM
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    thr_kernel<ILP><<<sms, threads>>>(out, 1.0, 0.999, 16);
    cudaEventRecord(e0);
    thr_kernel<ILP><<<sms, threads>>>(out, 1.0, 0.999, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double warp_instr_per_smsp = (double)iters*16*ILP*warps_per_smsp;
    const double cycles = ms*1e-3*1.965e9;
    printf("thr warps/SMSP=%d ILP=%d : %.3f ms, %.2f cycles per DFMA warp-instr per SMSP (at 1965 MHz)\n",
           warps_per_smsp, ILP, ms, cycles/warp_instr_per_smsp);
}

int main()
{
    double* out;
    long long* cyc;
    cudaMalloc(&out, 148*1024*sizeof(double));
    cudaMalloc(&cyc, sizeof(long long));
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    printf("%s, %d SMs, clock %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
    const char* names[] = {"DFMA", "DMUL", "DADD", "MUFU.RCP64H", "MUFU.RSQ64H", "sqrt()+DADD", "a/x"};
    long long h;
#define LAT(OP) lat_kernel<OP><<<1, 32>>>(out, cyc, 1.0000001, 0.9999999); lat_kernel<OP><<<1, 32>>>(out, cyc, 1.0000001, 0.9999999); \
    cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost); printf("lat %-12s : %.2f cycles per dependent op\n", names[OP], (double)h/CHAIN);
    LAT(0) LAT(1) LAT(2) LAT(3) LAT(4) LAT(5) LAT(6)
    for (int w = 1; w <= 8; w *= 2) {
        run_thr<1>(out, w, p.multiProcessorCount);
        run_thr<2>(out, w, p.multiProcessorCount);
        run_thr<4>(out, w, p.multiProcessorCount);
        run_thr<8>(out, w, p.multiProcessorCount);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
