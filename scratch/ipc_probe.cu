// probe: can two processes share device memory through cudaIpc handles, and peer-write over NVLink?
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <unistd.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("rank %d: %s -> %s\n", rank, #x, cudaGetErrorString(e)); return 1; } } while (0)
__global__ void fill(int *p, int v, int n) { int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) p[i] = v + i; }
int main(int argc, char **argv) {
  int rank = atoi(argv[1]);
  const char *path = argv[2];
  CK(cudaSetDevice(rank));
  int can = 0; CK(cudaDeviceCanAccessPeer(&can, rank, 1 - rank));
  printf("rank %d canAccessPeer=%d\n", rank, can);
  int *buf; CK(cudaMalloc(&buf, 1 << 20)); CK(cudaMemset(buf, 0, 1 << 20));
  cudaIpcMemHandle_t h; CK(cudaIpcGetMemHandle(&h, buf));
  char fn[256]; snprintf(fn, 256, "%s.%d", path, rank);
  FILE *f = fopen(fn, "wb"); fwrite(&h, sizeof(h), 1, f); fclose(f);
  snprintf(fn, 256, "%s.%d", path, 1 - rank);
  cudaIpcMemHandle_t ph;
  for (int t = 0; t < 200; t++) { f = fopen(fn, "rb"); if (f && fread(&ph, sizeof(ph), 1, f) == 1) { fclose(f); break; } if (f) fclose(f); usleep(50000); }
  int *peer; CK(cudaIpcOpenMemHandle((void **)&peer, ph, cudaIpcMemLazyEnablePeerAccess));
  fill<<<4, 256>>>(peer, 1000 * (rank + 1), 1024); CK(cudaDeviceSynchronize());
  sleep(2);
  int host[4]; CK(cudaMemcpy(host, buf, 16, cudaMemcpyDeviceToHost));
  printf("rank %d sees in own buffer: %d %d %d (written by peer)\n", rank, host[0], host[1], host[2]);
  CK(cudaIpcCloseMemHandle(peer));
  return 0;
}
