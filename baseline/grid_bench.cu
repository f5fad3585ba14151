// grid_bench.cu -- timing harness for the reference's own GPU prior art: the forward-dynamics-gradient kernel that
// the reference's GRiD code generator EMITS (GRiD/GRiDCodeGenerator/GRiDCodeGenerator.py:261 gen_all_code -> grid.cuh; the kernel is
// launched like the emitted host wrapper does, cf. GRiD/printGRiD.cu:62), compiled unchanged for sm_100a.  grid.cuh is generated
// by baseline/make_grid.py from the fixed arm6 URDF with the reference's generator and lives in baseline/_ref/ (git-ignored).
// This file is ours: it only allocates device-resident inputs, launches grid::forward_dynamics_gradient_kernel<T> over K knot
// points and times it with CUDA events.  Output: one JSON line.
//   grid_bench [knots=524288] [blocks=0 (one block per knot point)]
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <random>
#include "grid.cuh"

template <typename T>
double run(int K, int blocks, const char* name) {
  std::vector<T> h((size_t)K * 18);
  std::mt19937 rng(1337);
  std::uniform_real_distribution<double> dist(-0.8, 0.8);
  for (auto& v : h) v = (T)dist(rng);
  T *d_in = nullptr, *d_out = nullptr;
  gpuErrchk(cudaMalloc(&d_in, h.size() * sizeof(T)));
  gpuErrchk(cudaMalloc(&d_out, (size_t)K * 72 * sizeof(T)));
  gpuErrchk(cudaMemcpy(d_in, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  cudaStream_t* streams = grid::init_grid<T>();          // raises the dynamic shared-memory limits of the gradient kernels
  grid::robotModel<T>* d_model = grid::init_robotModel<T>();
  const dim3 threads(grid::SUGGESTED_THREADS, 1, 1), grd(blocks > 0 ? blocks : K, 1, 1);
  const size_t smem = grid::FD_DU_DYNAMIC_SHARED_MEM_COUNT * sizeof(T);
  auto kern = static_cast<void (*)(T*, const T*, const int, const grid::robotModel<T>*, const T, const int)>(&grid::forward_dynamics_gradient_kernel<T>);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  std::vector<float> ms;
  for (int rep = 0; rep < 8; ++rep) {
    cudaEventRecord(e0);
    kern<<<grd, threads, smem>>>(d_out, d_in, 18, d_model, (T)9.81, K);
    cudaEventRecord(e1);
    gpuErrchk(cudaEventSynchronize(e1));
    gpuErrchk(cudaGetLastError());
    float t = 0; cudaEventElapsedTime(&t, e0, e1);
    if (rep >= 3) ms.push_back(t);
  }
  std::sort(ms.begin(), ms.end());
  std::vector<T> out(72);
  gpuErrchk(cudaMemcpy(out.data(), d_out, 72 * sizeof(T), cudaMemcpyDeviceToHost));
  double chk = 0; for (T v : out) chk += (double)v;
  printf("{\"kernel\": \"grid::forward_dynamics_gradient_kernel<%s>\", \"knots\": %d, \"blocks\": %d, \"threads\": %d, \"smem_bytes\": %zu, "
         "\"ms_best\": %.4f, \"ms_median\": %.4f, \"ns_per_knot\": %.3f, \"checksum_knot0\": %.9g}\n",
         name, K, (int)grd.x, (int)threads.x, smem, ms.front(), ms[ms.size() / 2], 1e6 * ms.front() / K, chk);
  cudaFree(d_in); cudaFree(d_out);
  return ms.front();
}

// one scalar type per binary: the emitted kernels declare `extern __shared__ T s_XITemp[]`, which cannot be instantiated for two
// types in one translation unit
#ifndef GRID_T
#define GRID_T double
#endif
#define GRID_STR2(x) #x
#define GRID_STR(x) GRID_STR2(x)
int main(int argc, char** argv) {
  const int K = argc > 1 ? atoi(argv[1]) : 524288;
  const int blocks = argc > 2 ? atoi(argv[2]) : 0;
  run<GRID_T>(K, blocks, GRID_STR(GRID_T));
  return 0;
}
