// Groundwork for the population / large-batch tensor-core path:
// a hand-written sm_100a GEMM  D[M,N] = A[M,K] . B[N,K]^T  (both operands K-contiguous fp32, consumed as TF32)
//   * operands staged by TMA (cp.async.bulk.tensor, 128B swizzle) through a 4-stage mbarrier ring,
//   * tcgen05.mma.cta_group::1.kind::tf32 issued by one thread, accumulators in TMEM (128 lanes x BN columns),
//   * epilogue tcgen05.ld (32x32b) -> registers -> global.
// This is the forward-layer shape of the MLP (activations [batch, in] x nn.Linear weight [out, in]).
// Standalone on purpose: it validates descriptors / barriers / TMEM handling against a CPU reference and measures
// the tensor pipe before the path is wired into the engine. Not part of libsacx.so.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/bin/tc_gemm_bench tools/tc_gemm_bench.cu
#include <cuda.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

constexpr int BM = 128, BN = 128, BK = 32, NS = 4, UMMA_K = 8;
constexpr int STAGE_A = BM * BK * 4, STAGE_B = BN * BK * 4, STAGE = STAGE_A + STAGE_B;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(smem_u32(b)), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
// K-major operand, 128-byte swizzle: 8-row groups are 1024 B apart (SBO), LBO unused, version 1 (sm_100), layout 2
__device__ __forceinline__ uint64_t umma_desc(const void* smem_ptr) {
  const uint64_t addr = (smem_u32(smem_ptr) & 0x3FFFF) >> 4;
  return addr | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// MN-major operand staged as 32-wide (128 B) slabs of [BK rows (k) x 128 B]; lbo / sbo in bytes (probed from the host)
__device__ __forceinline__ uint64_t umma_desc_mn(const void* smem_ptr, uint32_t lbo, uint32_t sbo, uint32_t lt) {
  const uint64_t addr = (smem_u32(smem_ptr) & 0x3FFFF) >> 4;
  return addr | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46) | ((uint64_t)lt << 61);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(128, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, float* __restrict__ D, int M, int N, int K,
               int mode, uint32_t lbo, uint32_t sbo, uint32_t kstep, uint32_t lt) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t full[NS], empty[NS], acc_full;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int nkb = K / BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NS; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(&acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {   // one warp allocates the accumulator columns in tensor memory
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"((uint32_t)BN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0 && lane == 0) {
    // TMA producer
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % NS;
      mbar_wait(&empty[s], ((kb / NS) & 1) ^ 1);
      mbar_expect_tx(&full[s], STAGE);
      if (mode < 2) tma_load_2d(smem + s * STAGE, &mapA, &full[s], kb * BK, m0);
      else for (int j = 0; j < BM / 32; ++j) tma_load_2d(smem + s * STAGE + j * (BK * 128), &mapA, &full[s], m0 + j * 32, kb * BK);
      if (mode < 1) tma_load_2d(smem + s * STAGE + STAGE_A, &mapB, &full[s], kb * BK, n0);
      else for (int j = 0; j < BN / 32; ++j) tma_load_2d(smem + s * STAGE + STAGE_A + j * (BK * 128), &mapB, &full[s], n0 + j * 32, kb * BK);
    }
  } else if (warp == 1 && lane == 0) {
    // MMA issuer: instruction descriptor = F32 accumulate, TF32 x TF32, both K-major, N = BN, M = BM
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24) |
                           (mode >= 2 ? (1u << 15) : 0u) | (mode >= 1 ? (1u << 16) : 0u);     // bit 15 / 16: A / B is MN-major
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % NS;
      mbar_wait(&full[s], (kb / NS) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint8_t* a = smem + s * STAGE;
      const uint8_t* b = a + STAGE_A;
#pragma unroll
      for (int k = 0; k < BK / UMMA_K; ++k)
      {
        const uint64_t ad = mode >= 2 ? umma_desc_mn(a + k * kstep, lbo, sbo, lt) : umma_desc(a + k * UMMA_K * 4);
        const uint64_t bd = mode >= 1 ? umma_desc_mn(b + k * kstep, lbo, sbo, lt) : umma_desc(b + k * UMMA_K * 4);
        umma_tf32(tmem_base, ad, bd, idesc, (kb | k) != 0);
      }
      umma_commit(&empty[s]);          // smem slot is free once these MMAs have read it
    }
    umma_commit(&acc_full);            // accumulator complete
  }
  __syncwarp();
  // epilogue: warp w owns TMEM lanes 32w .. 32w+31  =  output rows m0 + 32w + lane
  mbar_wait(&acc_full, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int row = m0 + warp * 32 + lane;
#pragma unroll 1
  for (int c = 0; c < BN; c += 8) {
    uint32_t v[8];
    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (row < M) {
      float4* out = reinterpret_cast<float4*>(D + (size_t)row * N + n0 + c);
      out[0] = make_float4(__uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2]), __uint_as_float(v[3]));
      out[1] = make_float4(__uint_as_float(v[4]), __uint_as_float(v[5]), __uint_as_float(v[6]), __uint_as_float(v[7]));
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)BN) : "memory");
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static bool make_map(EncodeFn enc, CUtensorMap* map, float* base, int rows, int cols, int box_rows, int box_cols = BK, CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_128B) {
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); return false; }
  return true;
}

int main(int argc, char** argv) {
  const int M = argc > 1 ? atoi(argv[1]) : 8192, N = argc > 2 ? atoi(argv[2]) : 256, K = argc > 3 ? atoi(argv[3]) : 256;
  const int mode = argc > 4 ? atoi(argv[4]) : 0;
  const uint32_t lbo = argc > 5 ? atoi(argv[5]) : 4096, sbo = argc > 6 ? atoi(argv[6]) : 512, kstep = argc > 7 ? atoi(argv[7]) : 1024;
  const uint32_t lt = argc > 8 ? atoi(argv[8]) : 1;
  const CUtensorMapSwizzle swmn = lt == 1 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B;
  if (M % BM || N % BN || K % BK) { printf("M, N, K must be multiples of %d, %d, %d\n", BM, BN, BK); return 1; }
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) {
    printf("no cuTensorMapEncodeTiled entry point\n");
    return 1;
  }
  EncodeFn enc = (EncodeFn)fn;
  std::vector<float> hA((size_t)M * K), hB((size_t)N * K);
  srand(1);
  for (auto& x : hA) x = (rand() / (float)RAND_MAX - 0.5f);
  for (auto& x : hB) x = (rand() / (float)RAND_MAX - 0.5f);
  float *dA, *dB, *dD;
  cudaMalloc(&dA, hA.size() * 4); cudaMalloc(&dB, hB.size() * 4); cudaMalloc(&dD, (size_t)M * N * 4);
  // device layouts: mode 0: A[M,K], B[N,K]; mode 1: A[M,K], Bt[K,N]; mode 2: At[K,M], Bt[K,N]
  std::vector<float> tA(hA.size()), tB(hB.size());
  for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) tA[(size_t)k * M + m] = hA[(size_t)m * K + k];
  for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) tB[(size_t)k * N + n] = hB[(size_t)n * K + k];
  cudaMemcpy(dA, (mode >= 2 ? tA : hA).data(), hA.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, (mode >= 1 ? tB : hB).data(), hB.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0, (size_t)M * N * 4);
  CUtensorMap mapA, mapB;
  if (!(mode >= 2 ? make_map(enc, &mapA, dA, K, M, BK, 32, swmn) : make_map(enc, &mapA, dA, M, K, BM))) return 1;
  if (!(mode >= 1 ? make_map(enc, &mapB, dB, K, N, BK, 32, swmn) : make_map(enc, &mapB, dB, N, K, BN))) return 1;
  const int smem = NS * STAGE + 1024;
  cudaFuncSetAttribute(tc_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  dim3 grid(M / BM, N / BN);
  tc_gemm_kernel<<<grid, 128, smem>>>(mapA, mapB, dD, M, N, K, mode, lbo, sbo, kstep, lt);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
  std::vector<float> hD((size_t)M * N);
  cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
  // check a sample of rows against fp64
  double max_rel = 0, ref_norm = 0, err_norm = 0;
  for (int r = 0; r < M; r += 37) {
    for (int c = 0; c < N; ++c) {
      double s = 0;
      for (int k = 0; k < K; ++k) s += (double)hA[(size_t)r * K + k] * hB[(size_t)c * K + k];
      const double d = hD[(size_t)r * N + c] - s;
      ref_norm += s * s; err_norm += d * d;
      max_rel = fmax(max_rel, fabs(d));
    }
  }
  printf("mode %d lbo %u sbo %u kstep %u lt %u | ", mode, lbo, sbo, kstep, lt);
  printf("tcgen05 tf32 GEMM %dx%dx%d: rel-L2 error vs fp64 = %.3e (tf32 inputs: ~5e-4 expected), max abs err %.3e\n", M, N, K,
         sqrt(err_norm / ref_norm), max_rel);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 5; ++i) tc_gemm_kernel<<<grid, 128, smem>>>(mapA, mapB, dD, M, N, K, mode, lbo, sbo, kstep, lt);
  cudaEventRecord(e0);
  const int reps = 50;
  for (int i = 0; i < reps; ++i) tc_gemm_kernel<<<grid, 128, smem>>>(mapA, mapB, dD, M, N, K, mode, lbo, sbo, kstep, lt);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double tf = 2.0 * M * N * K * reps / (ms * 1e-3) / 1e12;
  printf("time %.3f us per GEMM, %.1f TFLOP/s (tf32), grid %d x %d CTAs\n", ms * 1e3 / reps, tf, grid.x, grid.y);
  return sqrt(err_norm / ref_norm) < 5e-3 ? 0 : 2;
}
