// micro-benchmark: cost of cooperative launches and of the hand-rolled grid barrier (gf_ingest.cuh) on B200
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
__device__ __forceinline__ unsigned ld_acq(const unsigned *p) { unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_rel(unsigned *p, unsigned v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
struct Bar { unsigned *arrive, *release; unsigned abase, rbase, grid; };
__device__ __forceinline__ void gbar(const Bar &b, unsigned k) {
  __shared__ unsigned s_last;
  __syncthreads();
  if (threadIdx.x == 0) { __threadfence(); unsigned old = atomicAdd(b.arrive, 1u); s_last = (old - b.abase) == (k + 1u) * b.grid - 1u; }
  __syncthreads();
  unsigned epoch = b.rbase + k + 1u;
  if (s_last) { if (threadIdx.x == 0) { __threadfence(); st_rel(b.release, epoch); } }
  else if (threadIdx.x == 0) { while ((int)(ld_acq(b.release) - epoch) < 0) {} __threadfence(); }
  __syncthreads();
}
__global__ void k_empty(int) {}
__global__ void k_bars(Bar b, int nb) { for (int k = 0; k < nb; k++) gbar(b, k); }
int main() {
  unsigned *w; cudaMalloc(&w, 8); cudaMemset(w, 0, 8);
  cudaStream_t st; cudaStreamCreate(&st);
  auto now = [] { return std::chrono::steady_clock::now(); };
  auto us = [](auto a, auto b) { return std::chrono::duration<double, std::micro>(b - a).count(); };
  const int N = 2000;
  for (int coop = 0; coop < 2; coop++) for (int grid : {148, 296}) for (int nb : {0, 4, 40}) {
    unsigned ab = 0, rb = 0; cudaMemset(w, 0, 8); cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto t0 = now(); cudaEventRecord(e0, st);
    for (int i = 0; i < N; i++) {
      Bar b = {w, w + 1, ab, rb, (unsigned)grid}; int nbb = nb; void *args[] = {&b, &nbb};
      if (coop) cudaLaunchCooperativeKernel((const void *)k_bars, dim3(grid), dim3(256), args, 0, st);
      else k_bars<<<grid, 256, 0, st>>>(b, nb);
      ab += nb * grid; rb += nb;
    }
    auto t1 = now(); cudaEventRecord(e1, st); cudaStreamSynchronize(st); auto t2 = now();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("coop=%d grid=%d barriers=%d: host enqueue %.2f us/launch, device %.2f us/launch, wall %.2f us/launch, err=%s\n", coop, grid, nb,
           us(t0, t1) / N, ms * 1000 / N, us(t0, t2) / N, cudaGetErrorString(cudaGetLastError()));
  }
  // launch + sync round trip
  for (int coop = 0; coop < 2; coop++) {
    Bar b = {w, w + 1, 0, 0, 148}; int nb = 0; void *args[] = {&b, &nb};
    auto t0 = now();
    for (int i = 0; i < N; i++) {
      if (coop) cudaLaunchCooperativeKernel((const void *)k_bars, dim3(148), dim3(256), args, 0, st); else k_bars<<<148, 256, 0, st>>>(b, nb);
      cudaStreamSynchronize(st);
    }
    printf("coop=%d launch+sync: %.2f us\n", coop, us(t0, now()) / N);
  }
  return 0;
}
