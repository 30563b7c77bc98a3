// How many thread-block clusters of a given size are co-resident on this GPU (per dynamic shared memory size)?
#include <cuda_runtime.h>
#include <cstdio>
__global__ void k(int* p) { extern __shared__ float s[]; if (p && threadIdx.x == 999) p[0] = (int)s[0]; }
int main() {
  cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  for (int smem : {0, 100 * 1024, 200 * 1024, 226 * 1024}) {
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int cs : {1, 2, 4, 6, 8, 10, 12, 14, 16}) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(cs * 64); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      int n = -1;
      cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &cfg);
      printf("smem %6d cluster %2d -> max active clusters %d (%d CTAs) %s\n", smem, cs, n, n * cs, e == cudaSuccess ? "" : cudaGetErrorString(e));
      cudaGetLastError();
    }
  }
  return 0;
}
