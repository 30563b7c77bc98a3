// Standalone micro-benchmark of the GEMM tile (nvcc -DSACX_MATH_VARIANT=n ...): every CTA runs one tile of a
// 256x256x256 layer per mode on L2-resident operands; prints median clock64 cycles per tile section.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include "../soft-actor-critic_b200/csrc/sacx_gemm.cuh"
using namespace sacx;

template <int MODE>
__global__ void __launch_bounds__(256, 1) bench_kernel(Op op, float* base, unsigned long long* ts, int reps) {
  extern __shared__ __align__(16) float smem[];
  AgentScalars* scal = reinterpret_cast<AgentScalars*>(base);
  Hyper hp; hp.tau = 0.005f; hp.one_minus_tau = 0.995f;
  EpiCtx ctx{base, scal, &hp, ts + blockIdx.x * 8};
  for (int r = 0; r < reps; ++r) gemm_tile_impl<CfgSmall, MODE>(op, ctx, blockIdx.x % op.ntiles, smem);
}

int main() {
  const int B = 256, H = 256;
  size_t floats = 64 + 8 * (size_t)B * H;
  float* base; cudaMalloc(&base, floats * 4);
  std::vector<float> h(floats);
  for (auto& x : h) x = (rand() / (float)RAND_MAX - 0.5f) * 0.1f;
  cudaMemcpy(base, h.data(), floats * 4, cudaMemcpyHostToDevice);
  unsigned long long* ts; cudaMalloc(&ts, 148 * 8 * 8);
  const size_t smem = CfgSmall::SMEM_FLOATS * 4;
  auto region = [&](int i) { return (i64)64 + (i64)i * B * H; };
  for (int mode = 0; mode < 3; ++mode) {
    Op op; memset(&op, 0, sizeof op);
    op.type = OP_GEMM; op.M = B; op.N = H; op.K = H; op.tiles_n = H / 32; op.ntiles = (B / 32) * (H / 32);
    op.a = region(0); op.b = region(1); op.c = region(2); op.bias = region(3); op.aux = region(3); op.zout = -1;
    op.ldc = H; op.ld_aux = H; op.a_vec = op.b_vec = 1; op.act = 1;
    op.p = region(4); op.pm = region(5); op.pv = region(6); op.pt = region(7); op.pg = -1;
    op.pb = region(3); op.pbm = region(3) + 256; op.pbv = region(3) + 512; op.pbt = region(3) + 768; op.pbg = -1;
    if (mode == 0) { op.epi = EPI_FWD; op.a_sm = H; op.a_sk = 1; op.b_sk = 1; op.b_sn = H; }
    if (mode == 1) { op.epi = EPI_DACT; op.a_sm = H; op.a_sk = 1; op.b_sk = H; op.b_sn = 1; }
    if (mode == 2) { op.epi = EPI_DW; op.flags = DW_ADAM | DW_POLYAK; op.a_sm = 1; op.a_sk = H; op.b_sk = H; op.b_sn = 1; }
    for (int rep = 0; rep < 3; ++rep) {
      if (mode == 0) { cudaFuncSetAttribute(bench_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); bench_kernel<0><<<148, 256, smem>>>(op, base, ts, 4); }
      if (mode == 1) { cudaFuncSetAttribute(bench_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); bench_kernel<1><<<148, 256, smem>>>(op, base, ts, 4); }
      if (mode == 2) { cudaFuncSetAttribute(bench_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); bench_kernel<2><<<148, 256, smem>>>(op, base, ts, 4); }
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    }
    std::vector<unsigned long long> t(148 * 8);
    cudaMemcpy(t.data(), ts, 148 * 8 * 8, cudaMemcpyDeviceToHost);
    const char* names[4] = {"load0", "kloop", "reduce", "epilogue"};
    printf("variant %d mode %d:", SACX_MATH_VARIANT, mode);
    for (int k = 0; k < 4; ++k) {
      std::vector<long long> d;
      for (int c = 0; c < 148; ++c) d.push_back((long long)(t[c * 8 + k + 1] - t[c * 8 + k]));
      std::sort(d.begin(), d.end());
      printf(" %s med %lld max %lld |", names[k], d[74], d[147]);
    }
    std::vector<long long> d;
    for (int c = 0; c < 148; ++c) d.push_back((long long)(t[c * 8 + 4] - t[c * 8 + 0]));
    std::sort(d.begin(), d.end());
    printf(" total med %lld max %lld\n", d[74], d[147]);
  }
  return 0;
}
