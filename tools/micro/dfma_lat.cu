// FP64 FMA dependent-issue latency and per-SMSP throughput on sm_100a: W warps per SM sub-partition, I independent chains per thread.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dfma_lat dfma_lat.cu ; run: ./dfma_lat
#include <cstdio>
#include <cuda_runtime.h>
template <int I>
__global__ void k(double* out, double a, double b, int iters, long long* cyc) {
    double x[I];
#pragma unroll
    for (int j = 0; j < I; j++) x[j] = threadIdx.x * 1e-3 + j;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
            for (int j = 0; j < I; j++) x[j] = fma(x[j], a, b);
    }
    long long t1 = clock64();
    double s = 0; 
#pragma unroll
    for (int j = 0; j < I; j++) s += x[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int I> void run(int warps_per_smsp) {
    double* out; long long* cyc; cudaMalloc(&out, 8 * 148 * 1024); cudaMalloc(&cyc, 8);
    int iters = 2000, threads = 32 * 4 * warps_per_smsp;
    k<I><<<148, threads>>>(out, 0.999, 1e-3, iters, cyc); cudaDeviceSynchronize();
    k<I><<<148, threads>>>(out, 0.999, 1e-3, iters, cyc); cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    double per_warp_instr = (double)c / (iters * 8.0 * I);
    printf("warps/SMSP %d  chains %d : %.2f cycles per warp-DFMA per warp, SMSP issue rate %.3f DFMA/clk (pipe %.0f%%)\n", warps_per_smsp, I, per_warp_instr,
           warps_per_smsp / per_warp_instr, 100.0 * 2.0 * warps_per_smsp / per_warp_instr);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    for (int w : {1, 2, 3}) { run<1>(w); run<2>(w); run<3>(w); run<4>(w); run<6>(w); run<8>(w); }
    return 0;
}
