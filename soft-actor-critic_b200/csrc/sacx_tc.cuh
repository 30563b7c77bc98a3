// Tensor-core path for the MLP GEMMs at large batch (SURVEY section 8d regime 2, north_star item 3): tcgen05.mma with
// TMEM accumulators, operands staged by TMA, fp32-level accuracy through a 3xTF32 split.
//
// One persistent CTA per SM walks the tiles of up to four GEMM ops of one plan phase (sacx_types.cuh: Op). A tile is
// 128 rows x N (N <= 256: the whole layer width, so the A operand -- the 64 MB activation matrix at batch 65536 -- is
// read from HBM exactly once) and the reduction runs in blocks of 16 (four 48 KB stages hide the TMA -> split -> MMA
// chain better than two 96 KB ones: measured 1.7x on the K = 256 layers). Three shapes, all C = A . B^T with K the
// reduction index (reference: sac/models.py:30-33,73-77 forward; the autograd backward of sac/agent.py:230,235,256):
//   EPI_FWD   h = act(x W^T + b)          A = x  [batch][in]   K-major     B = W [out][in]    K-major
//   EPI_DACT  dx = (dy W) * act'(h)       A = dy [batch][out]  K-major     B = W [out][in]    MN-major (n = in)
//   EPI_DW    dW = dy^T x (+ db = 1^T dy) A = dy [batch][out]  MN-major    B = x [batch][in]  MN-major, batch range split
// K-major tiles (16 floats = 64 B rows) use the 64-byte TMA/UMMA swizzle; MN-major fp32 tiles need the 128B-swizzle-with-32B-atoms layout
// (CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B / UMMA layout type 1, LBO = slab stride, SBO = 512 B: tools/tc_gemm_bench.cu
// validated all three against fp64 on a B200).
//
// Accuracy: TF32 keeps 10 mantissa bits; the parity contract is rel 1e-4 after K updates against the fp32 reference.
// Every operand tile is therefore split in shared memory into hi = tf32(x) and lo = x - hi by eight splitter warps
// (in place + a second buffer) and each k-step issues three MMAs (lo.hi, hi.lo, hi.hi) into the same fp32 TMEM
// accumulator; the dropped lo.lo term is ~2^-22 relative.
//
// Warp roles (448 threads): 0-3 epilogue (TMEM -> registers -> swizzled smem -> TMA store; DACT pulls act'(h) tiles
// in by TMA as well), 4 TMA producer, 5 MMA issuer + TMEM owner, 6-13 splitters. Four 48 KB operand stages, two
// 256-column TMEM accumulators (epilogue of tile i overlaps the main loop of tile i+1), two 16 KB staging buffers.
// dW partial tiles ([split][out][in]) and bias partials go to a scratch buffer; tc_dw_reduce_kernel sums the
// splits in a fixed order and applies the same fused optimiser epilogue as the FFMA tiles (Adam, Polyak).
#pragma once
#include <cuda.h>

#include <cstdio>

#include "sacx_gemm.cuh"

namespace sacx {

constexpr int TC_BM = 128, TC_BK = 16, TC_NMAX = 256, TC_STAGES = 4, TC_MAX_OPS = 4;
constexpr int TC_A_BYTES = TC_BM * TC_BK * 4;                 // 8 KB
constexpr int TC_B_BYTES = TC_NMAX * TC_BK * 4;               // 16 KB
constexpr int TC_HALF = TC_A_BYTES + TC_B_BYTES;              // hi (or lo) part of one stage
constexpr int TC_STAGE_BYTES = 2 * TC_HALF;                   // 48 KB
constexpr int TC_STG_BYTES = TC_BM * 32 * 4;                  // epilogue staging chunk: 128 rows x 32 columns
constexpr int TC_SMEM_BYTES = TC_STAGES * TC_STAGE_BYTES + 2 * TC_STG_BYTES;      // 224 KB; + ~3 KB static = the 227 KB limit
constexpr int TC_THREADS = 448;
constexpr int TC_EPI_WARPS = 4, TC_SPLIT_WARPS = 8, TC_SPLIT_THREADS = 256;
constexpr int TC_SLAB = 32 * TC_BK * 4;                       // MN-major slab: 16 k-rows x 128 B

struct TcOp {
  int kind, act;
  int M, N, K;               // output rows, output columns, reduction length
  int n_mma;                 // UMMA N (multiple of 16)
  int tile0, ntiles;         // this op's tiles inside the launch
  int m_tiles, splits, k_per_split;
  int a_mn, b_mn;            // operand is MN-major
  int b_rows;                // K-major B: box rows;  MN-major B: number of 32-wide slabs
  int a_bytes, b_bytes;      // bytes per stage
  unsigned idesc;
  int has_aux, pad;
  i64 bias;                  // FWD: arena offset of the bias row
  i64 bias_part;             // DW: scratch offset of the bias partials [split * m_tiles][8 splitter warps][128] (-1: none)
  // FWD: projection of every output row onto one weight vector (critic head riding on the last hidden layer's GEMM: the epilogue
  // thread owns the whole 256-wide row): out[m] = sum_n act(...)[m][n] * w[n] + b.  Arena offsets, -1: none
  i64 proj_w, proj_b, proj_out;
};

struct TcParams {
  int n_ops, total_tiles;          // total_tiles = n_agents * tiles_per_agent (agent-major)
  int n_agents, tiles_per_agent;
  int dbg, pad;                    // SACX_TC_DBG (timing experiments only): 1 no split, 2 one MMA per k-step, 4 no output store
  i64 agent_stride, scratch_stride;   // floats between two agents' arenas / scratch blocks
  float* arena;
  float* scratch;
  TcOp ops[TC_MAX_OPS];
};

struct TcMaps {
  CUtensorMap a[TC_MAX_OPS], b[TC_MAX_OPS], c[TC_MAX_OPS], aux[TC_MAX_OPS];
};

// ---- PTX wrappers --------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t tc_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tc_mbar_init(uint64_t* b, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tc_smem(b)), "r"(count));
}
__device__ __forceinline__ void tc_mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tc_smem(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tc_mbar_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc_smem(b)) : "memory");
}
__device__ __forceinline__ void tc_mbar_wait(uint64_t* b, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(tc_smem(b)), "r"(parity) : "memory");
  } while (!done);
}
// every tensor map is 3-D: (inner, rows, agent); a single agent is the degenerate case with one slice
__device__ __forceinline__ void tc_tma_load(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(tc_smem(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(tc_smem(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tc_tma_store(const CUtensorMap* map, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(tc_smem(src)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tc_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tc_bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N>
__device__ __forceinline__ void tc_bulk_wait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void tc_fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_bar(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// shared-memory matrix descriptors (sm_100 version bit 46)
__device__ __forceinline__ uint64_t tc_desc_k(uint32_t saddr) {          // K-major rows of 16 floats: 64B swizzle, 8-row groups 512 B apart
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (4ull << 61);
}
__device__ __forceinline__ uint64_t tc_desc_mn(uint32_t saddr) {         // MN-major fp32: 128B swizzle with 32B atoms
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(TC_SLAB >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (1ull << 61);
}
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tc_smem(bar)) : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, "
      "%19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// tile -> (agent, op, m tile, split)
struct TcTile { int agent, op, mt, split, m0, k0, nkb; };
__device__ __forceinline__ TcTile tc_decode(const TcParams& P, int tile) {
  TcTile t;
  t.agent = tile / P.tiles_per_agent;
  tile -= t.agent * P.tiles_per_agent;
  t.op = 0;
#pragma unroll
  for (int i = 1; i < TC_MAX_OPS; ++i)
    if (i < P.n_ops && tile >= P.ops[i].tile0) t.op = i;
  const TcOp& o = P.ops[t.op];
  const int lt = tile - o.tile0;
  t.mt = lt % o.m_tiles;
  t.split = lt / o.m_tiles;
  t.m0 = t.mt * TC_BM;
  t.k0 = t.split * o.k_per_split;
  const int klen = min(o.k_per_split, o.K - t.k0);
  t.nkb = (klen + TC_BK - 1) / TC_BK;
  return t;
}

// timing experiments / role trace (SACX_TC_DBG) exist only in builds with -DSACX_DEBUG_HOOKS: the production kernel carries none
#ifdef SACX_DEBUG_HOOKS
#define TC_DBG P.dbg
#define TC_TS(stmt) stmt
#else
#define TC_DBG 0
#define TC_TS(stmt)
#endif

__global__ void __launch_bounds__(TC_THREADS, 1)
sacx_tc_kernel(const __grid_constant__ TcParams P, const __grid_constant__ TcMaps maps) {
  extern __shared__ __align__(1024) uint8_t tc_raw[];
  __shared__ __align__(8) uint64_t full[TC_STAGES], ready[TC_STAGES], empty[TC_STAGES], acc_full[2], acc_empty[2], aux_bar;
  __shared__ uint32_t tmem_base_s;
#ifdef SACX_DEBUG_HOOKS
  __shared__ long long ts[4][4][3];            // SACX_TC_DBG & 128: per-role timestamps of CTA 0's first 4 tiles
  const long long t_start = clock64();
#endif
  uint8_t* smem = tc_raw;
  if ((tc_smem(tc_raw) & 1023u) != 0u) __trap();        // swizzled tiles need 1024-byte alignment (declared on tc_raw)
  __shared__ __align__(16) float bias_sm[TC_NMAX];     // FWD: bias row of the epilogue's current tile
  __shared__ __align__(16) float proj_sm[TC_NMAX];     // FWD: projection vector of the current tile (critic head)
  uint8_t* stg = smem + TC_STAGES * TC_STAGE_BYTES;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < TC_STAGES; ++s) { tc_mbar_init(&full[s], 1); tc_mbar_init(&ready[s], TC_SPLIT_WARPS); tc_mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) { tc_mbar_init(&acc_full[b], 1); tc_mbar_init(&acc_empty[b], TC_EPI_WARPS); }
    tc_mbar_init(&aux_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 5) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem(&tmem_base_s)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 4) {
    // ---------------------------------------------------------------- TMA producer
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
        const TcTile t = tc_decode(P, tile);
        const TcOp& o = P.ops[t.op];
        const CUtensorMap* ma = &maps.a[t.op];
        const CUtensorMap* mb = &maps.b[t.op];
        TC_TS(const int tn_ = (tile - blockIdx.x) / gridDim.x;)
        TC_TS(if (tn_ < 4) ts[0][tn_][0] = clock64() - t_start;)
        for (int kb = 0; kb < t.nkb; ++kb, ++it) {
          const int s = it % TC_STAGES;
          tc_mbar_wait(&empty[s], ((it / TC_STAGES) & 1) ^ 1);
          TC_TS(if (tn_ < 4 && kb == 0) ts[0][tn_][1] = clock64() - t_start;)
          tc_mbar_expect_tx(&full[s], (uint32_t)(o.a_bytes + o.b_bytes));
          uint8_t* st = smem + s * TC_STAGE_BYTES;
          const int k = t.k0 + kb * TC_BK;
          if (!o.a_mn) tc_tma_load(st, ma, &full[s], k, t.m0, t.agent);
          else
            for (int j = 0; j < TC_BM / 32; ++j) tc_tma_load(st + j * TC_SLAB, ma, &full[s], t.m0 + 32 * j, k, t.agent);
          if (!o.b_mn) tc_tma_load(st + TC_A_BYTES, mb, &full[s], k, 0, t.agent);
          else
            for (int j = 0; j < o.b_rows; ++j) tc_tma_load(st + TC_A_BYTES + j * TC_SLAB, mb, &full[s], 32 * j, k, t.agent);
          TC_TS(if (tn_ < 4) ts[0][tn_][2] = clock64() - t_start;)
        }
      }
    }
  } else if (warp == 5) {
    // ---------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      uint32_t it = 0, tc = 0;
      for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x, ++tc) {
        const TcTile t = tc_decode(P, tile);
        const TcOp& o = P.ops[t.op];
        const uint32_t buf = tc & 1;
        tc_mbar_wait(&acc_empty[buf], ((tc >> 1) & 1) ^ 1);
        tc_fence_after();
        TC_TS(if (tc < 4) ts[2][tc][0] = clock64() - t_start;)
        const uint32_t tacc = tmem_base + buf * TC_NMAX;
        for (int kb = 0; kb < t.nkb; ++kb, ++it) {
          const int s = it % TC_STAGES;
          tc_mbar_wait(&ready[s], (it / TC_STAGES) & 1);
          tc_fence_after();
          TC_TS(if (tc < 4 && kb == 0) ts[2][tc][1] = clock64() - t_start;)
          const uint32_t hi = tc_smem(smem + s * TC_STAGE_BYTES), lo = hi + TC_HALF;
#pragma unroll
          for (int k8 = 0; k8 < TC_BK / 8; ++k8) {
            const uint32_t aoff = o.a_mn ? k8 * 1024 : k8 * 32, boff = TC_A_BYTES + (o.b_mn ? k8 * 1024 : k8 * 32);
            const uint64_t ah = o.a_mn ? tc_desc_mn(hi + aoff) : tc_desc_k(hi + aoff);
            const uint64_t al = o.a_mn ? tc_desc_mn(lo + aoff) : tc_desc_k(lo + aoff);
            const uint64_t bh = o.b_mn ? tc_desc_mn(hi + boff) : tc_desc_k(hi + boff);
            const uint64_t bl = o.b_mn ? tc_desc_mn(lo + boff) : tc_desc_k(lo + boff);
            tc_mma(tacc, al, bh, o.idesc, (kb | k8) != 0);
            if (!(TC_DBG & 2)) {
              tc_mma(tacc, ah, bl, o.idesc, 1u);
              tc_mma(tacc, ah, bh, o.idesc, 1u);
            }
          }
          tc_commit(&empty[s]);
        }
        tc_commit(&acc_full[buf]);
        TC_TS(if (tc < 4) ts[2][tc][2] = clock64() - t_start;)
      }
    }
  } else if (warp >= 6) {
    // ---------------------------------------------------------------- splitters: x -> (tf32(x), x - tf32(x))
    const int st = tid - 6 * 32, sw = warp - 6;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
      const TcTile t = tc_decode(P, tile);
      const TcOp& o = P.ops[t.op];
      const bool want_bias = (o.kind == EPI_DW) && (o.bias_part >= 0);
      float bs[2][4];       // dy tile = 4 slabs of 128 float4; this thread meets slabs (st >> 7) and (st >> 7) + 2
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) bs[j][e] = 0.f;
      const int n4 = (TC_DBG & 512) ? (TC_A_BYTES >> 4) : ((TC_A_BYTES + o.b_bytes) >> 4);      // 512: timing experiment, B operand not split
      for (int kb = 0; kb < t.nkb; ++kb, ++it) {
        const int s = it % TC_STAGES;
        tc_mbar_wait(&full[s], (it / TC_STAGES) & 1);
        TC_TS({ const int tn_ = (tile - blockIdx.x) / gridDim.x; if (st == 0 && tn_ < 4 && kb == 0) ts[1][tn_][0] = clock64() - t_start; })
        float4* hi = reinterpret_cast<float4*>(smem + s * TC_STAGE_BYTES);
        float4* lo = reinterpret_cast<float4*>(smem + s * TC_STAGE_BYTES + TC_HALF);
        // six float4 per thread cover a full-width stage (1536 float4): all loads first, then the arithmetic and stores
        constexpr int U = 6;
        for (int base = st; base < ((TC_DBG & 1) ? 0 : n4); base += TC_SPLIT_THREADS * U) {
          float4 x[U];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int i = base + u * TC_SPLIT_THREADS;
            x[u] = (i < n4) ? hi[i] : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int i = base + u * TC_SPLIT_THREADS;
            if (i < n4) {
              float4 h, l;
              if (TC_DBG & 256) {     // experiment: leave the raw value as "hi" (the tensor core drops the low 13 bits itself)
                l.x = x[u].x - __uint_as_float(__float_as_uint(x[u].x) & 0xffffe000u);
                l.y = x[u].y - __uint_as_float(__float_as_uint(x[u].y) & 0xffffe000u);
                l.z = x[u].z - __uint_as_float(__float_as_uint(x[u].z) & 0xffffe000u);
                l.w = x[u].w - __uint_as_float(__float_as_uint(x[u].w) & 0xffffe000u);
                lo[i] = l;
                continue;
              }
              h.x = __uint_as_float((__float_as_uint(x[u].x) + 0x1000u) & 0xffffe000u); l.x = x[u].x - h.x;
              h.y = __uint_as_float((__float_as_uint(x[u].y) + 0x1000u) & 0xffffe000u); l.y = x[u].y - h.y;
              h.z = __uint_as_float((__float_as_uint(x[u].z) + 0x1000u) & 0xffffe000u); l.z = x[u].z - h.z;
              h.w = __uint_as_float((__float_as_uint(x[u].w) + 0x1000u) & 0xffffe000u); l.w = x[u].w - h.w;
              hi[i] = h;
              lo[i] = l;
            }
          }
          if (want_bias && base == st) {         // column sums of the dy tile (float4 0..511 = u 0, 1) = bias gradient
            bs[0][0] += x[0].x; bs[0][1] += x[0].y; bs[0][2] += x[0].z; bs[0][3] += x[0].w;
            bs[1][0] += x[1].x; bs[1][1] += x[1].y; bs[1][2] += x[1].z; bs[1][3] += x[1].w;
          }
        }
        tc_fence_async();
        __syncwarp();
        if (lane == 0) tc_mbar_arrive(&ready[s]);
        TC_TS({ const int tn_ = (tile - blockIdx.x) / gridDim.x; if (st == 0 && tn_ < 4) ts[1][tn_][kb == 0 ? 1 : 2] = clock64() - t_start; })
      }
      if (want_bias) {
        // thread (k-row r = (st % 128) / 8, 16-byte chunk q = st % 8) holds out-features slab*32 + ((q>>1) ^ (r&3))*8 + (q&1)*4 + e
        // of its two slabs; a warp covers 4 k-rows of one slab pair: sum them (lanes holding the same logical chunk)
        const int rr = lane >> 3, q = lane & 7, c = (q >> 1) ^ rr, hbit = q & 1;
#pragma unroll
        for (int d = 1; d <= 2; d <<= 1) {
          const int r2 = rr ^ d, src = r2 * 8 + (((c ^ r2) << 1) | hbit);
#pragma unroll
          for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) bs[j][e] += __shfl_sync(0xffffffffu, bs[j][e], src);
        }
        // per-warp shares go straight to the scratch ([row block][warp][128], zeros for the slabs a warp does not meet);
        // tc_dw_reduce_kernel sums them in a fixed order
        if (rr == 0) {
          float* dst = P.scratch + t.agent * P.scratch_stride + o.bias_part + ((i64)(t.split * o.m_tiles + t.mt) * TC_SPLIT_WARPS + sw) * TC_BM;
          const int sl0 = sw >> 2;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (j == sl0) v = make_float4(bs[0][0], bs[0][1], bs[0][2], bs[0][3]);
            if (j == sl0 + 2) v = make_float4(bs[1][0], bs[1][1], bs[1][2], bs[1][3]);
            *reinterpret_cast<float4*>(dst + j * 32 + c * 8 + hbit * 4) = v;
          }
        }
      }
    }
  } else {
    // ---------------------------------------------------------------- epilogue (warps 0-3: TMEM lanes 32w .. 32w+31)
    uint32_t tc = 0, chunk = 0, auxn = 0;
    for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x, ++tc) {
      const TcTile t = tc_decode(P, tile);
      const TcOp& o = P.ops[t.op];
      const uint32_t buf = tc & 1;
      tc_mbar_wait(&acc_full[buf], (tc >> 1) & 1);
      tc_fence_after();
      TC_TS(if (tid == 0 && tc < 4) ts[3][tc][0] = clock64() - t_start;)
      const int row = tid;                                  // row of the tile == TMEM lane
      const int crow = (o.kind == EPI_DW) ? (t.split * o.m_tiles + t.mt) * TC_BM : t.m0;
      const int ncols = (o.N + 31) & ~31;
      const int act = o.act;
      if (o.kind == EPI_FWD) {                 // the tile's bias row -> shared memory (the previous tile's readers are past their last chunk: barrier below)
        tc_bar(1, TC_EPI_WARPS * 32);
        for (int n = tid; n < ncols; n += TC_EPI_WARPS * 32) {
          bias_sm[n] = (n < o.N) ? __ldg(P.arena + t.agent * P.agent_stride + o.bias + n) : 0.f;
          if (o.proj_w >= 0) proj_sm[n] = (n < o.N) ? __ldg(P.arena + t.agent * P.agent_stride + o.proj_w + n) : 0.f;
        }
      }
      const bool proj = (o.kind == EPI_FWD) && (o.proj_w >= 0);
      float pacc = 0.f;
      for (int c0 = 0; c0 < ncols; c0 += 32, ++chunk) {
        uint8_t* sb = stg + (chunk & 1) * TC_STG_BYTES;
        if (tid == 0) tc_bulk_wait_read<1>();               // the store that last read this buffer has drained it
        tc_bar(1, TC_EPI_WARPS * 32);
        if (o.has_aux) {
          if (tid == 0) {
            tc_mbar_expect_tx(&aux_bar, TC_STG_BYTES);
            tc_tma_load(sb, &maps.aux[t.op], &aux_bar, c0, t.m0, t.agent);
          }
          tc_mbar_wait(&aux_bar, auxn & 1);
          ++auxn;
        }
        uint32_t v[32];
        if (!(TC_DBG & 32)) tc_ld32(tmem_base + buf * TC_NMAX + c0 + ((uint32_t)(warp * 32) << 16), v);
        else { for (int z = 0; z < 32; ++z) v[z] = 0u; }
        float4* srow = reinterpret_cast<float4*>(sb + row * 128);
        const int sx = row & 7;
        if (o.kind == EPI_FWD) {
          const float4* b4 = reinterpret_cast<const float4*>(bias_sm + c0);       // broadcast reads
          const float4* p4 = reinterpret_cast<const float4*>(proj_sm + c0);
          if (act == SACX_ACT_RELU) {
#pragma unroll
            for (int qd = 0; qd < 8; ++qd) {
              const float4 b = b4[qd];
              const float4 hv = make_float4(fmaxf(__uint_as_float(v[4 * qd]) + b.x, 0.f), fmaxf(__uint_as_float(v[4 * qd + 1]) + b.y, 0.f),
                                            fmaxf(__uint_as_float(v[4 * qd + 2]) + b.z, 0.f), fmaxf(__uint_as_float(v[4 * qd + 3]) + b.w, 0.f));
              srow[qd ^ sx] = hv;
              if (proj) { const float4 w = p4[qd]; pacc = fmaf(hv.x, w.x, fmaf(hv.y, w.y, fmaf(hv.z, w.z, fmaf(hv.w, w.w, pacc)))); }
            }
          } else {
#pragma unroll
            for (int qd = 0; qd < 8; ++qd) {
              const float4 b = b4[qd];
              const float4 hv = make_float4(act_fwd(act, __uint_as_float(v[4 * qd]) + b.x), act_fwd(act, __uint_as_float(v[4 * qd + 1]) + b.y),
                                            act_fwd(act, __uint_as_float(v[4 * qd + 2]) + b.z), act_fwd(act, __uint_as_float(v[4 * qd + 3]) + b.w));
              srow[qd ^ sx] = hv;
              if (proj) { const float4 w = p4[qd]; pacc = fmaf(hv.x, w.x, fmaf(hv.y, w.y, fmaf(hv.z, w.z, fmaf(hv.w, w.w, pacc)))); }
            }
          }
        } else if (o.kind == EPI_DACT) {
          if (act == SACX_ACT_RELU) {
#pragma unroll
            for (int qd = 0; qd < 8; ++qd) {
              const float4 h = srow[qd ^ sx];
              srow[qd ^ sx] = make_float4(h.x > 0.f ? __uint_as_float(v[4 * qd]) : 0.f, h.y > 0.f ? __uint_as_float(v[4 * qd + 1]) : 0.f,
                                          h.z > 0.f ? __uint_as_float(v[4 * qd + 2]) : 0.f, h.w > 0.f ? __uint_as_float(v[4 * qd + 3]) : 0.f);
            }
          } else {
#pragma unroll
            for (int qd = 0; qd < 8; ++qd) {
              const float4 h = srow[qd ^ sx];
              srow[qd ^ sx] = make_float4(__uint_as_float(v[4 * qd]) * act_dz(act, h.x), __uint_as_float(v[4 * qd + 1]) * act_dz(act, h.y),
                                          __uint_as_float(v[4 * qd + 2]) * act_dz(act, h.z), __uint_as_float(v[4 * qd + 3]) * act_dz(act, h.w));
            }
          }
        } else {
#pragma unroll
          for (int qd = 0; qd < 8; ++qd)
            srow[qd ^ sx] = make_float4(__uint_as_float(v[4 * qd]), __uint_as_float(v[4 * qd + 1]), __uint_as_float(v[4 * qd + 2]),
                                        __uint_as_float(v[4 * qd + 3]));
        }
        if (!(TC_DBG & 8)) tc_fence_async();
        if (!(TC_DBG & 64)) tc_bar(1, TC_EPI_WARPS * 32);
        if (tid == 0 && !(TC_DBG & 4)) {
          tc_tma_store(&maps.c[t.op], sb, c0, crow, t.agent);
          tc_bulk_commit();
        }
      }
      if (proj && t.m0 + row < o.M)
        P.arena[t.agent * P.agent_stride + o.proj_out + t.m0 + row] = pacc + __ldg(P.arena + t.agent * P.agent_stride + o.proj_b);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) tc_mbar_arrive(&acc_empty[buf]);
      TC_TS(if (tid == 0 && tc < 4) ts[3][tc][1] = clock64() - t_start;)
    }
    if (tid == 0) tc_bulk_wait<0>();
  }
  tc_fence_before();
  __syncthreads();
#ifdef SACX_DEBUG_HOOKS
  if ((TC_DBG & 128) && blockIdx.x == 0 && tid == 0) {
    const int nt = min(4, (P.total_tiles + (int)gridDim.x - 1) / (int)gridDim.x);
    printf("tc kernel: %d ops, %d tiles, K0 %d; end %lld cycles\n", P.n_ops, P.total_tiles, P.ops[0].K, clock64() - t_start);
    for (int i = 0; i < nt; ++i)
      printf("  tile %d: tma start %lld first-empty %lld issued %lld | split first-full %lld first-done %lld last-done %lld | mma acc-empty %lld "
             "first-ready %lld committed %lld | epi acc-full %lld done %lld\n", i, ts[0][i][0], ts[0][i][1], ts[0][i][2], ts[1][i][0], ts[1][i][1],
             ts[1][i][2], ts[2][i][0], ts[2][i][1], ts[2][i][2], ts[3][i][0], ts[3][i][1]);
  }
#endif
  if (warp == 5) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ---- dW: sum the split partials, then the optimiser epilogue of the FFMA dW tile (sacx_gemm.cuh: epilogue_row4<2>) -----
struct TcRedOp {
  Op op;
  i64 part, bias_part;      // scratch offsets (floats)
  int splits, m_pad, n_ld, pad;
};
struct TcRedParams {
  int n_ops, pad;
  float* arena;
  const float* scratch;
  i64 scal_off, agent_stride, scratch_stride;
  Hyper hp;
  TcRedOp ops[TC_MAX_OPS];
};

__global__ void __launch_bounds__(256) tc_dw_reduce_kernel(const __grid_constant__ TcRedParams R) {
  const TcRedOp& r = R.ops[blockIdx.y];
  const Op& op = r.op;
  float* base = R.arena + (i64)blockIdx.z * R.agent_stride;
  const float* scratch = R.scratch + (i64)blockIdx.z * R.scratch_stride;
  const AgentScalars* scal = reinterpret_cast<const AgentScalars*>(base + R.scal_off);
  const bool vec = ((op.N & 3) == 0) && (((op.p | op.pm | op.pv | op.pg) & 3) == 0) && (op.pt < 0 || (op.pt & 3) == 0);
  const int n4 = (op.N + 3) >> 2;
  const i64 total = (i64)op.M * n4 + op.M;           // weight float4 groups, then one bias element per output row
  const float ss = (op.flags & DW_ADAM) ? __ldcg(&scal->adam_step_size[op.opt]) : 0.f;
  const float bc = (op.flags & DW_ADAM) ? __ldcg(&scal->adam_bc2_sqrt[op.opt]) : 1.f;
  const float tau = __ldcg(&scal->tau), omt = __ldcg(&scal->one_minus_tau);
  for (i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (i64)gridDim.x * blockDim.x) {
    if (e < (i64)op.M * n4) {
      const int m = (int)(e / n4), n = (int)(e % n4) * 4;
      float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
      const float* p = scratch + r.part + (i64)m * r.n_ld + n;           // n_ld is a multiple of 4: the group is in range
      const i64 sstep = (i64)r.m_pad * r.n_ld;
      int s = 0;
      for (; s + 7 < r.splits; s += 8) {            // eight partial tiles in flight (the loop is latency-bound); summed in split order
        float4 x[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) x[u] = __ldcs(reinterpret_cast<const float4*>(p + (s + u) * sstep));
#pragma unroll
        for (int u = 0; u < 8; ++u) { g.x += x[u].x; g.y += x[u].y; g.z += x[u].z; g.w += x[u].w; }
      }
      for (; s < r.splits; ++s) {
        const float4 x = __ldcs(reinterpret_cast<const float4*>(p + s * sstep));
        g.x += x.x; g.y += x.y; g.z += x.z; g.w += x.w;
      }
      const i64 w = (i64)m * op.N + n;
      if (!vec) {                                    // narrow or unaligned weight rows (first layer of a 5-wide input, ...)
        const float gg[4] = {g.x, g.y, g.z, g.w};
        for (int j = 0; j < 4 && n + j < op.N; ++j) {
          if (op.flags & DW_STORE_GRAD) base[op.pg + w + j] = gg[j];
          if (op.flags & DW_ADAM) {
            float pp = base[op.p + w + j], mm = base[op.pm + w + j], vv = base[op.pv + w + j];
            adam_update(gg[j], pp, mm, vv, ss, bc);
            base[op.p + w + j] = pp; base[op.pm + w + j] = mm; base[op.pv + w + j] = vv;
            if (op.flags & DW_POLYAK) base[op.pt + w + j] = polyak_mix(tau, omt, pp, base[op.pt + w + j]);
          }
        }
        continue;
      }
      if (op.flags & DW_STORE_GRAD) *reinterpret_cast<float4*>(base + op.pg + w) = g;
      if (op.flags & DW_ADAM) {
        const float4 p4 = *reinterpret_cast<const float4*>(base + op.p + w), m4 = *reinterpret_cast<const float4*>(base + op.pm + w),
                     v4 = *reinterpret_cast<const float4*>(base + op.pv + w);
        float gg[4] = {g.x, g.y, g.z, g.w}, pp[4] = {p4.x, p4.y, p4.z, p4.w}, mm[4] = {m4.x, m4.y, m4.z, m4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w};
        float tt[4] = {0.f, 0.f, 0.f, 0.f};
        if (op.flags & DW_POLYAK) {
          const float4 t4 = *reinterpret_cast<const float4*>(base + op.pt + w);
          tt[0] = t4.x; tt[1] = t4.y; tt[2] = t4.z; tt[3] = t4.w;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          adam_update(gg[j], pp[j], mm[j], vv[j], ss, bc);
          if (op.flags & DW_POLYAK) tt[j] = polyak_mix(tau, omt, pp[j], tt[j]);
        }
        *reinterpret_cast<float4*>(base + op.p + w) = make_float4(pp[0], pp[1], pp[2], pp[3]);
        *reinterpret_cast<float4*>(base + op.pm + w) = make_float4(mm[0], mm[1], mm[2], mm[3]);
        *reinterpret_cast<float4*>(base + op.pv + w) = make_float4(vv[0], vv[1], vv[2], vv[3]);
        if (op.flags & DW_POLYAK) *reinterpret_cast<float4*>(base + op.pt + w) = make_float4(tt[0], tt[1], tt[2], tt[3]);
      }
    } else if (op.pb >= 0) {
      const int m = (int)(e - (i64)op.M * n4);
      float g = 0.f;
      const float* bp = scratch + r.bias_part + (i64)(m / TC_BM) * TC_SPLIT_WARPS * TC_BM + (m % TC_BM);
      for (int s = 0; s < r.splits; ++s)
        for (int w = 0; w < TC_SPLIT_WARPS; ++w) g += __ldcs(bp + ((i64)s * (r.m_pad / TC_BM) * TC_SPLIT_WARPS + w) * TC_BM);
      if (op.flags & DW_STORE_GRAD) base[op.pbg + m] = g;
      if (op.flags & DW_ADAM) {
        float p = base[op.pb + m], mm = base[op.pbm + m], vv = base[op.pbv + m];
        adam_update(g, p, mm, vv, ss, bc);
        base[op.pb + m] = p; base[op.pbm + m] = mm; base[op.pbv + m] = vv;
        if (op.flags & DW_POLYAK) base[op.pbt + m] = polyak_mix(tau, omt, p, base[op.pbt + m]);
      }
    }
  }
}

}  // namespace sacx
