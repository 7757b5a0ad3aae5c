// Unit test + timing of gpmp2_b200/csrc/mma_solve.cuh against a host dense Cholesky.
// build: nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -I gpmp2_b200/csrc -o scripts/micro/mma_solve_test scripts/micro/mma_solve_test.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <random>
#include <cuda_runtime.h>
#include "mma_solve.cuh"

template <int D, bool TWO>
__global__ void __launch_bounds__(TWO ? 64 : 32, TWO ? 8 : 1) k_solve(const double* __restrict__ Hd_in, const double* __restrict__ Ho_in, const double* __restrict__ g_in,
                                              double* __restrict__ out, int N, int nsys, double lambda, int reps, unsigned long long* clk) {
  extern __shared__ double smem[];
  constexpr int b = 2 * D, BD = b * (b + 1) / 2, BB = b * b;
  double* g = smem; double* dl = g + N * b; double* scr = dl + N * b; double* Ho = scr + 64; double* Hd = Ho + (N - 1) * BB;
  mma::Solver<D> S;
  const int lane = threadIdx.x, NTH = TWO ? 64 : 32;
  long long tsolve = 0, nsolve = 0;
  for (int sys = blockIdx.x; sys < nsys; sys += gridDim.x) {
    for (int rep = 0; rep < reps; rep++) {
      for (int i = lane; i < N * BD; i += NTH) Hd[i] = Hd_in[(size_t)sys * N * BD + i];
      for (int i = lane; i < (N - 1) * BB; i += NTH) Ho[i] = Ho_in[(size_t)sys * (N - 1) * BB + i];
      for (int i = lane; i < N * b; i += NTH) g[i] = g_in[(size_t)sys * N * b + i];
      if (TWO) __syncthreads(); else __syncwarp();
      const long long t0 = clock64();
      if (TWO) S.solve2(Hd, Ho, g, dl, lambda, N, scr, [] {}); else S.solve(Hd, Ho, g, dl, lambda, N, scr);
      if (TWO) __syncthreads(); else __syncwarp();
      tsolve += clock64() - t0; nsolve++;
    }
    for (int i = lane; i < N * b; i += NTH) out[(size_t)sys * N * b + i] = dl[i];
    if (TWO) __syncthreads(); else __syncwarp();
  }
  if (lane == 0 && clk) { atomicAdd(clk, (unsigned long long)tsolve); atomicAdd(clk + 1, (unsigned long long)nsolve); }
#ifdef MMA_SOLVE_PROFILE
  if (threadIdx.x == 0 && clk) for (int k = 0; k < 6; k++) atomicAdd(clk + 2 + k, (unsigned long long)S.pt[k]);
#endif
}

template <int D, bool TWO = false>
int run(int N, int nsys, int reps, int wps = 0) {
  constexpr int b = 2 * D, BD = b * (b + 1) / 2, BB = b * b;
  const int n = N * b;
  std::mt19937_64 rng(1234 + D * 100 + N);
  std::normal_distribution<double> nd(0.0, 1.0);
  std::vector<double> Hd((size_t)nsys * N * BD), Ho((size_t)nsys * std::max(1, N - 1) * BB), g((size_t)nsys * n), ref((size_t)nsys * n);
  const double lambda = 0.37;
  const int nref = std::min(nsys, 64);
  const int ngen = std::min(nsys, 256);   // distinct systems generated on the host; the rest are copies
  for (int s = 0; s < nsys; s++) {
    if (s >= ngen) {
      const int q = s % ngen;
      std::copy(Hd.begin() + (size_t)q * N * BD, Hd.begin() + (size_t)(q + 1) * N * BD, Hd.begin() + (size_t)s * N * BD);
      if (N > 1) std::copy(Ho.begin() + (size_t)q * (N - 1) * BB, Ho.begin() + (size_t)(q + 1) * (N - 1) * BB, Ho.begin() + (size_t)s * (N - 1) * BB);
      std::copy(g.begin() + (size_t)q * n, g.begin() + (size_t)(q + 1) * n, g.begin() + (size_t)s * n);
      continue;
    }
    // A = J^T J restricted to the block-tridiagonal pattern + diagonal shift (SPD by diagonal dominance of the shift)
    std::vector<double> A((size_t)n * n, 0.0), Jm((size_t)2 * n * n);
    for (auto& v : Jm) v = nd(rng);
    for (int i = 0; i < n; i++)
      for (int j2 = 0; j2 <= i; j2++) {
        if (i / b - j2 / b > 1) continue;
        double acc = 0; for (int k = 0; k < 2 * n; k++) acc += Jm[(size_t)k * n + i] * Jm[(size_t)k * n + j2];
        if (i / b != j2 / b) acc *= 0.4;
        if (i == j2) acc += 0.6 * n;
        A[(size_t)i * n + j2] = A[(size_t)j2 * n + i] = acc;
      }
    for (int i = 0; i < N; i++)
      for (int r = 0; r < b; r++)
        for (int c = 0; c <= r; c++) Hd[((size_t)s * N + i) * BD + r * (r + 1) / 2 + c] = A[(size_t)(i * b + r) * n + i * b + c];
    for (int i = 0; i + 1 < N; i++)
      for (int r = 0; r < b; r++)
        for (int c = 0; c < b; c++) Ho[((size_t)s * (N - 1) + i) * BB + r * b + c] = A[(size_t)(i * b + r) * n + (i + 1) * b + c];
    for (int i = 0; i < n; i++) g[(size_t)s * n + i] = nd(rng);
    if (s < nref) {
      // dense Cholesky solve of (A + lambda I) x = -g
      std::vector<long double> L((size_t)n * n, 0.0L), x(n);
      for (int i = 0; i < n; i++)
        for (int j2 = 0; j2 <= i; j2++) {
          long double acc = A[(size_t)i * n + j2] + (i == j2 ? lambda : 0.0);
          for (int k = 0; k < j2; k++) acc -= L[(size_t)i * n + k] * L[(size_t)j2 * n + k];
          L[(size_t)i * n + j2] = (i == j2) ? sqrtl(acc) : acc / L[(size_t)j2 * n + j2];
        }
      for (int i = 0; i < n; i++) { long double acc = -g[(size_t)s * n + i]; for (int k = 0; k < i; k++) acc -= L[(size_t)i * n + k] * x[k]; x[i] = acc / L[(size_t)i * n + i]; }
      for (int i = n - 1; i >= 0; i--) { long double acc = x[i]; for (int k = i + 1; k < n; k++) acc -= L[(size_t)k * n + i] * x[k]; x[i] = acc / L[(size_t)i * n + i]; }
      for (int i = 0; i < n; i++) ref[(size_t)s * n + i] = (double)x[i];
    }
  }
  double *dHd, *dHo, *dg, *dout; unsigned long long* dclk; cudaMalloc(&dclk, 128); cudaMemset(dclk, 0, 128);
  cudaMalloc(&dHd, Hd.size() * 8); cudaMalloc(&dHo, Ho.size() * 8); cudaMalloc(&dg, g.size() * 8); cudaMalloc(&dout, g.size() * 8);
  cudaMemcpy(dHd, Hd.data(), Hd.size() * 8, cudaMemcpyHostToDevice); cudaMemcpy(dHo, Ho.data(), Ho.size() * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(dg, g.data(), g.size() * 8, cudaMemcpyHostToDevice);
  const size_t smem = sizeof(double) * (2 * n + 64 + (size_t)(N - 1) * BB + (size_t)N * BD + 2);
  cudaFuncSetAttribute(k_solve<D, TWO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_solve<D, TWO>, TWO ? 64 : 32, smem);
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
  const int grid = std::min(nsys, (wps ? std::min(wps, occ) : occ) * prop.multiProcessorCount);
  k_solve<D, TWO><<<grid, TWO ? 64 : 32, smem>>>(dHd, dHo, dg, dout, N, nsys, lambda, 1, nullptr);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("D=%d N=%d CUDA error %s\n", D, N, cudaGetErrorString(e)); return 1; }
  std::vector<double> out(g.size());
  cudaMemcpy(out.data(), dout, out.size() * 8, cudaMemcpyDeviceToHost);
  double maxerr = 0, maxref = 0;
  for (size_t i = 0; i < (size_t)nref * n; i++) { maxerr = std::max(maxerr, std::fabs(out[i] - ref[i])); maxref = std::max(maxref, std::fabs(ref[i])); }
  bool nan = false; for (double v : out) if (!std::isfinite(v)) nan = true;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms = 0;
  if (reps > 0) {
    cudaEventRecord(e0);
    k_solve<D, TWO><<<grid, TWO ? 64 : 32, smem>>>(dHd, dHo, dg, dout, N, nsys, lambda, reps, dclk);
    cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
  }
  unsigned long long hclk[8] = {0, 1}; if (reps > 0) cudaMemcpy(hclk, dclk, 64, cudaMemcpyDeviceToHost);
#ifdef MMA_SOLVE_PROFILE
  if (reps > 0) printf("   warp 0 per solve: forward %.0f, wait %.0f, middle %.0f, middle back %.0f, wait %.0f, chain back %.0f clk\n", (double)hclk[2] / hclk[1], (double)hclk[3] / hclk[1], (double)hclk[4] / hclk[1], (double)hclk[5] / hclk[1], (double)hclk[6] / hclk[1], (double)hclk[7] / hclk[1]);
#endif
  if (reps > 0) printf("   in-kernel: %.0f clk per solve per warp (%llu solves)\n", (double)hclk[0] / (double)hclk[1], hclk[1]);
  printf("%s D=%d N=%2d nsys=%6d occ=%d smem=%zu  max rel err %.2e %s   %s\n", TWO ? "2-warp" : "1-warp", D, N, nsys, occ, smem, maxerr / maxref, nan ? "NaN!" : "",
         reps > 0 ? (std::to_string(ms * 1e3 / ((double)nsys * reps) * 1e3) + " ns/solve (" + std::to_string(ms) + " ms, incl. H reload)").c_str() : "");
  cudaFree(dHd); cudaFree(dHo); cudaFree(dg); cudaFree(dout);
  return (maxerr / maxref > 1e-9 || nan) ? 1 : 0;
}

int main(int argc, char** argv) {
  int bad = 0;
  if (argc > 1) { if (argc > 3) run<7>(11, 16384, atoi(argv[1]), atoi(argv[2])); else run<7, true>(11, 16384, atoi(argv[1]), argc > 2 ? atoi(argv[2]) : 0); return 0; }
  bad += run<7>(11, 256, 0);
  bad += run<7>(1, 64, 0); bad += run<7>(2, 64, 0); bad += run<7>(3, 64, 0); bad += run<7>(4, 64, 0); bad += run<7>(10, 64, 0); bad += run<7>(21, 64, 0);
  bad += run<1>(11, 64, 0); bad += run<2>(11, 64, 0); bad += run<3>(11, 64, 0); bad += run<4>(11, 64, 0); bad += run<5>(11, 64, 0); bad += run<6>(11, 64, 0);
  bad += run<3>(2, 64, 0); bad += run<4>(6, 64, 0); bad += run<5>(7, 64, 0);
  bad += run<7, true>(11, 256, 0);
  bad += run<7, true>(1, 64, 0); bad += run<7, true>(2, 64, 0); bad += run<7, true>(3, 64, 0); bad += run<7, true>(4, 64, 0); bad += run<7, true>(10, 64, 0); bad += run<7, true>(21, 64, 0);
  bad += run<1, true>(11, 64, 0); bad += run<2, true>(11, 64, 0); bad += run<3, true>(11, 64, 0); bad += run<4, true>(11, 64, 0); bad += run<5, true>(11, 64, 0); bad += run<6, true>(11, 64, 0);
  bad += run<3, true>(2, 64, 0); bad += run<4, true>(6, 64, 0); bad += run<5, true>(7, 64, 0);
  printf(bad ? "FAILED %d\n" : "all ok\n", bad);
  run<7>(11, 65536, 4);
  run<7, true>(11, 65536, 4);
  run<3>(11, 65536, 4);
  run<3, true>(11, 65536, 4);
  return bad;
}
