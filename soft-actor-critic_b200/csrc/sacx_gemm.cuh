// FP32 tiled GEMM for the MLP layers of the SAC update, with the layer's elementwise work fused into
// the epilogue (bias+activation forward, activation-derivative backward, Adam(+Polyak) on dW).
//
// One CTA (256 threads) computes a BM x BN output tile. The K range is split over KG thread groups
// inside the CTA (intra-CTA split-K): at batch 256 / hidden 256 a layer is only 256x256x256, so tiles
// must be small (32x32) to spread one layer over the 148 SMs, and split-K keeps all 8 warps busy on
// such a tile. Partial tiles are reduced through shared memory; the epilogue then runs on coalesced
// float4 rows. Operands are staged through double-buffered shared memory as S[k][m] / S[k][n]
// (K-major) with register prefetch of the next K block.
//
// Precision: plain FFMA, fp32 accumulate -- the parity contract is rel 1e-4 against the fp32
// reference, which rules out TF32/BF16 tensor-core inputs for this path (SURVEY 2b note).
#pragma once
#include "sacx_math.cuh"
#include "sacx_types.cuh"

namespace sacx {

template <int BM_, int BN_, int TM_, int TN_, int KG_, int BKG_>
struct TileCfg {
  static constexpr int BM = BM_, BN = BN_, TM = TM_, TN = TN_, KG = KG_, BKG = BKG_;
  static constexpr int TPG = (BM / TM) * (BN / TN);   // threads per K group
  static constexpr int THREADS = TPG * KG;
  static constexpr int BKT = KG * BKG;                // K extent of one smem stage
  static constexpr int SA = BM + 4, SB = BN + 4;      // padded leading dims (keep 16B alignment)
  static constexpr int RS = BN + 4;                   // reduction row stride (+1 col used for bias sums)
  static constexpr int STAGE_FLOATS = BKT * (SA + SB);
  static constexpr int RED_FLOATS = KG * BM * RS;
  static constexpr int SMEM_FLOATS = 2 * STAGE_FLOATS + RED_FLOATS;
  static_assert(THREADS == 256, "tile configs are written for 256-thread CTAs");
  static_assert(BKT == 64, "loaders assume 64-wide K stages");
};

using CfgSmall = TileCfg<32, 32, 4, 4, 4, 16>;     // latency config: many small tiles (single agent)
using CfgLarge = TileCfg<64, 64, 8, 8, 4, 16>;     // throughput config: population / large batch

struct EpiCtx {
  float* base;                  // agent arena base
  const AgentScalars* scal;
  const Hyper* hp;
};

// ---- global -> register tile loads -------------------------------------------------------------
// A tile is R (rows along m or n) x 64 (k). NV = float4 per thread = R/16.
// contig_k: element (row, k) at base[row*s_row + k]   (K contiguous)  -> transposed smem store
// else     : element (row, k) at base[k*s_k + row]    (row contiguous) -> direct float4 smem store
template <int R>
__device__ __forceinline__ void tile_load(const float* __restrict__ base, bool contig_k, int s_other, bool vec,
                                          int row0, int rows_total, int k0, int k_total, float4 (&v)[R / 16]) {
  const int tid = threadIdx.x;
#pragma unroll
  for (int i = 0; i < R / 16; ++i) {
    const int id = tid + i * 256;
    int row, k;
    if (contig_k) {
      const int r_lo = id & 15, q = (id >> 4) & 1, rest = id >> 5;
      row = r_lo + 16 * (rest >> 3);
      k = (((rest & 7) << 1) + q) << 2;
    } else {
      row = (id % (R / 4)) << 2;
      k = id / (R / 4);
    }
    const int gr = row0 + row, gk = k0 + k;
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    if (contig_k) {
      if (gr < rows_total) {
        const float* p = base + (i64)gr * s_other + gk;
        if (vec && gk + 3 < k_total) {
          t = __ldcg(reinterpret_cast<const float4*>(p));
        } else {
          if (gk + 0 < k_total) t.x = __ldcg(p + 0);
          if (gk + 1 < k_total) t.y = __ldcg(p + 1);
          if (gk + 2 < k_total) t.z = __ldcg(p + 2);
          if (gk + 3 < k_total) t.w = __ldcg(p + 3);
        }
      }
    } else {
      if (gk < k_total) {
        const float* p = base + (i64)gk * s_other + gr;
        if (vec && gr + 3 < rows_total) {
          t = __ldcg(reinterpret_cast<const float4*>(p));
        } else {
          if (gr + 0 < rows_total) t.x = __ldcg(p + 0);
          if (gr + 1 < rows_total) t.y = __ldcg(p + 1);
          if (gr + 2 < rows_total) t.z = __ldcg(p + 2);
          if (gr + 3 < rows_total) t.w = __ldcg(p + 3);
        }
      }
    }
    v[i] = t;
  }
}

template <int R, int LD>
__device__ __forceinline__ void tile_store(float* __restrict__ S, bool contig_k, const float4 (&v)[R / 16]) {
  const int tid = threadIdx.x;
#pragma unroll
  for (int i = 0; i < R / 16; ++i) {
    const int id = tid + i * 256;
    if (contig_k) {
      const int r_lo = id & 15, q = (id >> 4) & 1, rest = id >> 5;
      const int row = r_lo + 16 * (rest >> 3);
      const int k = (((rest & 7) << 1) + q) << 2;
      S[(k + 0) * LD + row] = v[i].x;
      S[(k + 1) * LD + row] = v[i].y;
      S[(k + 2) * LD + row] = v[i].z;
      S[(k + 3) * LD + row] = v[i].w;
    } else {
      const int row = (id % (R / 4)) << 2;
      const int k = id / (R / 4);
      *reinterpret_cast<float4*>(&S[k * LD + row]) = v[i];
    }
  }
}

// ---- epilogues ---------------------------------------------------------------------------------
__device__ __forceinline__ void epilogue_row4(const Op& op, const EpiCtx& ctx, int m, int n, float4 acc) {
  float* base = ctx.base;
  float v[4] = {acc.x, acc.y, acc.z, acc.w};
  if (m >= op.M) return;
  if (op.epi == EPI_FWD) {
    float* c = base + op.c + (i64)m * op.ldc;
    float* z = op.zout >= 0 ? base + op.zout + (i64)m * op.ldc : nullptr;
    const float* bias = base + op.bias;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (n + j < op.N) {
        const float zz = v[j] + __ldcg(bias + n + j);
        if (z) z[n + j] = zz;
        c[n + j] = act_fwd(op.act, zz);
      }
    }
  } else if (op.epi == EPI_DACT) {
    float* c = base + op.c + (i64)m * op.ldc;
    const float* aux = base + op.aux + (i64)m * op.ld_aux;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (n + j < op.N) c[n + j] = v[j] * act_dz(op.act, __ldcg(aux + n + j));
  } else {  // EPI_DW: (m, n) = (out neuron, in feature); parameter leading dim = N (= K_in)
    const i64 e = (i64)m * op.N + n;
    const float ss = __ldcg(&ctx.scal->adam_step_size[op.opt]), bc = __ldcg(&ctx.scal->adam_bc2_sqrt[op.opt]);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (n + j < op.N) {
        const float g = v[j];
        if (op.flags & DW_STORE_GRAD) base[op.pg + e + j] = g;
        if (op.flags & DW_ADAM) {
          float p = __ldcg(base + op.p + e + j), mm = __ldcg(base + op.pm + e + j), vv = __ldcg(base + op.pv + e + j);
          adam_update(g, p, mm, vv, ss, bc);
          base[op.p + e + j] = p;
          base[op.pm + e + j] = mm;
          base[op.pv + e + j] = vv;
          if (op.flags & DW_POLYAK) {
            float* t = base + op.pt + e + j;
            *t = polyak_mix(ctx.hp->tau, ctx.hp->one_minus_tau, p, __ldcg(t));
          }
        }
      }
    }
  }
}

__device__ __forceinline__ void epilogue_bias(const Op& op, const EpiCtx& ctx, int m, float g) {
  if (m >= op.M) return;
  float* base = ctx.base;
  if (op.flags & DW_STORE_GRAD) base[op.pbg + m] = g;
  if (op.flags & DW_ADAM) {
    const float ss = __ldcg(&ctx.scal->adam_step_size[op.opt]), bc = __ldcg(&ctx.scal->adam_bc2_sqrt[op.opt]);
    float p = __ldcg(base + op.pb + m), mm = __ldcg(base + op.pbm + m), vv = __ldcg(base + op.pbv + m);
    adam_update(g, p, mm, vv, ss, bc);
    base[op.pb + m] = p;
    base[op.pbm + m] = mm;
    base[op.pbv + m] = vv;
    if (op.flags & DW_POLYAK) {
      float* t = base + op.pbt + m;
      *t = polyak_mix(ctx.hp->tau, ctx.hp->one_minus_tau, p, __ldcg(t));
    }
  }
}

// ---- the tile ------------------------------------------------------------------------------------
template <class C>
__device__ void gemm_tile(const Op& op, const EpiCtx& ctx, int tile, float* __restrict__ smem) {
  constexpr int BM = C::BM, BN = C::BN, TM = C::TM, TN = C::TN, KG = C::KG, BKG = C::BKG, BKT = C::BKT;
  constexpr int SA = C::SA, SB = C::SB, RS = C::RS;
  const int tid = threadIdx.x;
  const int tm = tile / op.tiles_n, tn = tile % op.tiles_n;
  const int m0 = tm * BM, n0 = tn * BN;
  const float* __restrict__ A = ctx.base + op.a;
  const float* __restrict__ Bp = ctx.base + op.b;
  const bool a_ck = (op.a_sk == 1), b_ck = (op.b_sk == 1);
  const int a_other = a_ck ? op.a_sm : op.a_sk;
  const int b_other = b_ck ? op.b_sn : op.b_sk;
  const bool bias_tile = (op.epi == EPI_DW) && (tn == 0) && (op.pb >= 0);

  float* As = smem;
  float* Bs = smem + 2 * BKT * SA;
  float* red = smem + 2 * C::STAGE_FLOATS;

  const int kg = tid / C::TPG, t = tid % C::TPG;
  const int ty = t / (BN / TN), tx = t % (BN / TN);

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
  float bsum[TM];
#pragma unroll
  for (int i = 0; i < TM; ++i) bsum[i] = 0.f;

  float4 ra[BM / 16], rb[BN / 16];
  const int nk = (op.K + BKT - 1) / BKT;
  tile_load<BM>(A, a_ck, a_other, op.a_vec != 0, m0, op.M, 0, op.K, ra);
  tile_load<BN>(Bp, b_ck, b_other, op.b_vec != 0, n0, op.N, 0, op.K, rb);
  tile_store<BM, SA>(As, a_ck, ra);
  tile_store<BN, SB>(Bs, b_ck, rb);
  __syncthreads();

  for (int it = 0; it < nk; ++it) {
    const int cur = it & 1;
    const bool more = (it + 1 < nk);
    if (more) {
      tile_load<BM>(A, a_ck, a_other, op.a_vec != 0, m0, op.M, (it + 1) * BKT, op.K, ra);
      tile_load<BN>(Bp, b_ck, b_other, op.b_vec != 0, n0, op.N, (it + 1) * BKT, op.K, rb);
    }
    const float* as = As + cur * BKT * SA + (kg * BKG) * SA + ty * TM;
    const float* bs = Bs + cur * BKT * SB + (kg * BKG) * SB + tx * TN;
#pragma unroll
    for (int kk = 0; kk < BKG; ++kk) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; i += 4) {
        const float4 q = *reinterpret_cast<const float4*>(as + kk * SA + i);
        a[i] = q.x; a[i + 1] = q.y; a[i + 2] = q.z; a[i + 3] = q.w;
      }
#pragma unroll
      for (int j = 0; j < TN; j += 4) {
        const float4 q = *reinterpret_cast<const float4*>(bs + kk * SB + j);
        b[j] = q.x; b[j + 1] = q.y; b[j + 2] = q.z; b[j + 3] = q.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      if (bias_tile) {
#pragma unroll
        for (int i = 0; i < TM; ++i) bsum[i] += a[i];
      }
    }
    if (more) {
      tile_store<BM, SA>(As + (cur ^ 1) * BKT * SA, a_ck, ra);
      tile_store<BN, SB>(Bs + (cur ^ 1) * BKT * SB, b_ck, rb);
    }
    __syncthreads();
  }

  // intra-CTA split-K reduction through shared memory
  float* myred = red + kg * (BM * RS);
#pragma unroll
  for (int i = 0; i < TM; ++i) {
#pragma unroll
    for (int j = 0; j < TN; j += 4)
      *reinterpret_cast<float4*>(&myred[(ty * TM + i) * RS + tx * TN + j]) =
          make_float4(acc[i][j], acc[i][j + 1], acc[i][j + 2], acc[i][j + 3]);
    if (bias_tile && tx == 0) myred[(ty * TM + i) * RS + BN] = bsum[i];
  }
  __syncthreads();
  constexpr int NV = BM * BN / 4 / 256;   // float4 outputs per thread
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int g = tid + i * 256;
    const int row = g / (BN / 4), c4 = (g % (BN / 4)) * 4;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int q = 0; q < KG; ++q) {
      const float4 p = *reinterpret_cast<const float4*>(&red[q * (BM * RS) + row * RS + c4]);
      s.x += p.x; s.y += p.y; s.z += p.z; s.w += p.w;
    }
    epilogue_row4(op, ctx, m0 + row, n0 + c4, s);
  }
  if (bias_tile && tid < BM) {
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < KG; ++q) s += red[q * (BM * RS) + tid * RS + BN];
    epilogue_bias(op, ctx, m0 + tid, s);
  }
  __syncthreads();   // smem is reused by the next tile
}

}  // namespace sacx
