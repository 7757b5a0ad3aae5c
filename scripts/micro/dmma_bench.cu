// Microbenchmark: FP64 DMMA (mma.sync m8n8k4 / m16n8k8 / m16n8k16) vs DFMA vs SHFL throughput on sm_100a.
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o dmma_bench dmma_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double* c, const double* a, const double* b) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
template <int NACC> __global__ void k_dmma884(double* out, double a, double b) {
  double c[NACC][2];
  for (int i = 0; i < NACC; i++) c[i][0] = c[i][1] = threadIdx.x;
  for (int it = 0; it < ITERS; it++)
#pragma unroll
    for (int i = 0; i < NACC; i++) dmma884(c[i][0], c[i][1], a, b);
  double s = 0; for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC> __global__ void k_dmma1688(double* out, double a, double b) {
  double c[NACC][4]; double av[4] = {a, a, a, a}, bv[2] = {b, b};
  for (int i = 0; i < NACC; i++) for (int j = 0; j < 4; j++) c[i][j] = threadIdx.x;
  for (int it = 0; it < ITERS; it++)
#pragma unroll
    for (int i = 0; i < NACC; i++) dmma1688(c[i], av, bv);
  double s = 0; for (int i = 0; i < NACC; i++) for (int j = 0; j < 4; j++) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC> __global__ void k_dfma(double* out, double a, double b) {
  double c[NACC];
  for (int i = 0; i < NACC; i++) c[i] = threadIdx.x + i;
  for (int it = 0; it < ITERS; it++)
#pragma unroll
    for (int i = 0; i < NACC; i++) c[i] = fma(c[i], a, b);
  double s = 0; for (int i = 0; i < NACC; i++) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC> __global__ void k_shfl(double* out, int src) {
  int c[NACC];
  for (int i = 0; i < NACC; i++) c[i] = threadIdx.x + i;
  for (int it = 0; it < ITERS; it++)
#pragma unroll
    for (int i = 0; i < NACC; i++) c[i] = __shfl_sync(0xffffffffu, c[i], (src + c[i]) & 31);
  int s = 0; for (int i = 0; i < NACC; i++) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// dependent latency of one dmma chain
__global__ void k_dmma_lat(double* out, double a, double b) {
  double c0 = threadIdx.x, c1 = 1;
  for (int it = 0; it < ITERS; it++) dmma884(c0, c1, a, b);
  out[blockIdx.x * blockDim.x + threadIdx.x] = c0 + c1;
}
// LDS.128 broadcast vs LDS.64 per-lane
__global__ void k_lds(double* out, int stride, int mode) {
  __shared__ double sm[4096];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i;
  __syncthreads();
  double s = 0; int idx = (threadIdx.x & 31) * stride;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (mode == 0) { s += sm[(idx + i * 64 + it) & 4095]; }
      else { double2 v = *(const double2*)&sm[((idx + i * 64 + it * 2) & 4094)]; s += v.x + v.y; }
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <class F> float timeit(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount; double* out; cudaMalloc(&out, sizeof(double) * sms * 1024 * 8);
  printf("%s SMs %d clock %d kHz\n", p.name, sms, p.clockRate);
  for (int wpb : {4, 8, 16, 32}) {
    int grid = sms * 2, thr = wpb * 32;
    double nw = (double)grid * wpb;
    float ms = timeit([&] { k_dfma<8><<<grid, thr>>>(out, 1.0000001, 1e-9); });
    printf("warps/blk %2d dfma  x8acc : %.2f TFLOP/s\n", wpb, nw * 32 * 8 * ITERS * 2 / ms / 1e9);
    ms = timeit([&] { k_dmma884<4><<<grid, thr>>>(out, 1.0000001, 1e-9); });
    printf("warps/blk %2d dmma884 x4acc: %.2f TFLOP/s\n", wpb, nw * 256 * 4 * ITERS * 2 / ms / 1e9);
    ms = timeit([&] { k_dmma884<8><<<grid, thr>>>(out, 1.0000001, 1e-9); });
    printf("warps/blk %2d dmma884 x8acc: %.2f TFLOP/s\n", wpb, nw * 256 * 8 * ITERS * 2 / ms / 1e9);
    ms = timeit([&] { k_dmma1688<4><<<grid, thr>>>(out, 1.0000001, 1e-9); });
    printf("warps/blk %2d dmma1688 x4acc: %.2f TFLOP/s\n", wpb, nw * 1024 * 4 * ITERS * 2 / ms / 1e9);
    ms = timeit([&] { k_shfl<8><<<grid, thr>>>(out, 1); });
    printf("warps/blk %2d shfl.32: %.2f warp-instr/clk/SM\n", wpb, nw * 8 * ITERS / (ms * 1e-3 * 1.965e9) / sms);
    ms = timeit([&] { k_lds<<<grid, thr>>>(out, 1, 0); });
    printf("warps/blk %2d lds.64 stride1: %.2f warp-instr/clk/SM\n", wpb, nw * 8 * ITERS / (ms * 1e-3 * 1.965e9) / sms);
    ms = timeit([&] { k_lds<<<grid, thr>>>(out, 0, 1); });
    printf("warps/blk %2d lds.128 broadcast: %.2f warp-instr/clk/SM\n", wpb, nw * 8 * ITERS / (ms * 1e-3 * 1.965e9) / sms);
    ms = timeit([&] { k_lds<<<grid, thr>>>(out, 2, 1); });
    printf("warps/blk %2d lds.128 stride2: %.2f warp-instr/clk/SM\n", wpb, nw * 8 * ITERS / (ms * 1e-3 * 1.965e9) / sms);
  }
  float ms = timeit([&] { k_dmma_lat<<<sms, 32>>>(out, 1.0000001, 1e-9); });
  printf("dmma884 dependent latency: %.1f clk\n", ms * 1e-3 * 1.965e9 / ITERS);
  ms = timeit([&] { k_dfma<1><<<sms, 32>>>(out, 1.0000001, 1e-9); });
  printf("dfma dependent latency: %.1f clk\n", ms * 1e-3 * 1.965e9 / ITERS);
  return 0;
}
