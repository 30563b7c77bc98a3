// mma.sync.m16n8k8 TF32 on sm_100a: latency of a dependent chain and throughput with independent accumulators,
// per SM (8 warps, 256 threads), measured with clock64. Calibrates the row-parallel kernel's cost model.
#include <cuda_runtime.h>
#include <cstdio>
__device__ __forceinline__ void mma(float (&c)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
template <int CH>
__global__ void k(float* out, long long* cyc, int iters) {
  unsigned a[4] = {threadIdx.x, 2, 3, 4}, b[2] = {5, 6};
  float c[CH][4];
  for (int i = 0; i < CH; ++i) for (int q = 0; q < 4; ++q) c[i][q] = 0.f;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) mma(c[i], a, b);
  }
  long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < CH; ++i) for (int q = 0; q < 4; ++q) s += c[i][q];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 1000;
  for (int warps : {1, 4, 8, 16}) {
    long long h;
#define RUN(CH) k<CH><<<1, warps * 32>>>(out, cyc, iters); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); \
    printf("warps %2d chains %d: %.1f cycles per mma per warp, SM rate %.2f mma/cycle (%.0f MAC/cycle/SM)\n", warps, CH, (double)h / (iters * CH), \
           (double)iters * CH * warps / h, 1024.0 * iters * CH * warps / h);
    RUN(1) RUN(2) RUN(3) RUN(6)
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
