// Row-wise operations of the SAC update: one warp owns one batch row. These are the narrow output
// layers (policy head N=2A, critic head N=1) and everything the reference computes right after them:
// the tanh-squashed Gaussian rsample/log_prob (sac/models.py:79-87), the soft Bellman target
// (sac/agent.py:207-210), the critic/actor loss gradients and the closed-form head backward
// (SURVEY section 8 a5/a8), the temperature step (agent.py:263-280). Reductions use warp shuffles.
#pragma once
#include "sacx_math.cuh"
#include "sacx_types.cuh"

namespace sacx {

constexpr int ROWS_PER_TILE = 8;   // 256 threads = 8 warps = 8 rows

struct RowCtx {
  float* base;
  AgentScalars* scal;
  const RunArgs* args;
  int agent, step;
  float* wsm;          // per-warp shared scratch, 2*SACX_MAX_ACT*2 floats
};

// dot(h[0:K], w[0:K]) over the lanes of a warp (all lanes receive the result)
__device__ __forceinline__ float warp_dot(const float* __restrict__ h, const float* __restrict__ w, int K, int lane) {
  float s = 0.f;
  if (((K & 3) == 0) && ((((uintptr_t)h | (uintptr_t)w) & 15) == 0)) {
    for (int k = lane * 4; k < K; k += 128) {
      const float4 a = __ldcg(reinterpret_cast<const float4*>(h + k));
      const float4 b = __ldcg(reinterpret_cast<const float4*>(w + k));
      s = fmaf(a.x, b.x, s); s = fmaf(a.y, b.y, s); s = fmaf(a.z, b.z, s); s = fmaf(a.w, b.w, s);
    }
  } else {
    for (int k = lane; k < K; k += 32) s = fmaf(__ldcg(h + k), __ldcg(w + k), s);
  }
  return warp_sum(s);
}

// ---------------------------------------------------------------- OP_GATHER
// o[0]=X_sa o[1]=X_s2 o[2]=X_pi o[3]=r o[4]=d o[5]=idx(i64)  i[0]=ldx
__device__ __forceinline__ void op_gather(const Op& op, const RowCtx& c, int row, int lane) {
  const RunArgs& a = *c.args;
  const Hyper& hp = a.hp;
  if (row >= hp.B) return;
  const float* ring = a.ring + (i64)c.agent * a.ring_stride;
  const RingMeta* meta = reinterpret_cast<const RingMeta*>(ring);
  const i64 pushes = meta->pushes;
  const i64 cap = a.ring_capacity;
  const i64 n = pushes < cap ? pushes : cap;
  const i64 upd = __ldcg(&c.scal->updates);
  i64 j;
  if (a.idx_ext) {
    j = a.idx_ext[((i64)c.step * a.n_agents + c.agent) * hp.B + row];
  } else {
    // throughput mode: position (global row) of a keyed bijection on [0, n) -> distinct indices
    j = (i64)feistel_index((unsigned long long)(hp.row0_global + row), (unsigned long long)n, hp.seed,
                           (unsigned long long)upd, (uint32_t)c.agent);
  }
  const i64 oldest = pushes > cap ? pushes - cap : 0;
  const i64 slot = (oldest + j) % cap;
  float* base = c.base;
  const int O = hp.obs, A = hp.act, ldx = op.i[0];
  if (lane == 0) reinterpret_cast<i64*>(base + op.o[5])[row] = j;
  const float* rs = ring + a.ring_s + slot * O;
  const float* rs2 = ring + a.ring_s2 + slot * O;
  float* xsa = base + op.o[0] + (i64)row * ldx;
  float* xs2 = base + op.o[1] + (i64)row * ldx;
  float* xpi = base + op.o[2] + (i64)row * ldx;
  for (int k = lane; k < O; k += 32) {
    const float v = __ldcs(rs + k);
    xsa[k] = v;
    xpi[k] = v;
    xs2[k] = __ldcs(rs2 + k);
  }
  const float* ra = ring + a.ring_a + slot * A;
  for (int k = lane; k < A; k += 32) xsa[O + k] = __ldcs(ra + k);
  if (lane == 0) {
    base[op.o[3] + row] = __ldcs(ring + a.ring_r + slot);
    base[op.o[4] + row] = __ldcs(ring + a.ring_d + slot);
  }
}

// ---------------------------------------------------------------- OP_PI_HEAD
// o[0]=h(last hidden) o[1]=W_L o[2]=b_L o[3]=X(dest, action at col obs+j) o[4]=lp o[5]=eps buf
// o[6]=tz o[7]=se o[8]=mask o[9]=headz (each -1 when not saved)   i[0]=ldh i[1]=K i[2]=ldx   mode: 1 target / 2 actor
__device__ __forceinline__ void op_pi_head(const Op& op, const RowCtx& c, int row, int lane) {
  const RunArgs& a = *c.args;
  const Hyper& hp = a.hp;
  if (row >= hp.B) return;
  float* base = c.base;
  const int A = hp.act, K = op.i[1];
  const float* h = base + op.o[0] + (i64)row * op.i[0];
  const float* W = base + op.o[1];
  const float* bias = base + op.o[2];
  float* hs = c.wsm;                      // head post-activation [2A], then pre-activation [2A]
  for (int j = 0; j < 2 * A; ++j) {
    const float z = warp_dot(h, W + (i64)j * K, K, lane) + __ldcg(bias + j);
    if (lane == 0) {
      hs[j] = act_fwd(op.act_out, z);
      hs[2 * SACX_MAX_ACT + j] = z;
    }
  }
  __syncwarp();
  const float* eps_ext = (op.mode == 1) ? a.eps1_ext : a.eps2_ext;
  const i64 upd = __ldcg(&c.scal->updates);
  float lp_part = 0.f;
  bool bad = false;
  for (int j = lane; j < A; j += 32) {
    const float mu = hs[j], ls_raw = hs[A + j];
    const float ls = fminf(fmaxf(ls_raw, hp.log_std_min), hp.log_std_max);
    const float sd = expf(ls);
    float e;
    if (eps_ext) e = eps_ext[(((i64)c.step * a.n_agents + c.agent) * hp.B + row) * A + j];
    else e = philox_normal(hp.seed, (unsigned long long)upd, op.mode, (uint32_t)(hp.row0_global + row), (uint32_t)j,
                           (uint32_t)c.agent);
    base[op.o[5] + (i64)row * A + j] = e;
    const float z = mu + e * sd;                       // Normal.rsample: loc + eps * scale
    const float tz = tanhf(z);
    base[op.o[3] + (i64)row * op.i[2] + hp.obs + j] = tz * hp.action_scale;
    const float dzm = z - mu;
    // Normal.log_prob: -(z-mu)^2 / (2 var) - log(std) - log(sqrt(2 pi))
    float lp = -(dzm * dzm) / (2.f * (sd * sd)) - logf(sd) - 0.91893853320467274178f;
    // tanh Jacobian, softplus form: 2 (log 2 - z - softplus(-2 z))   (F7: no -log(action_scale))
    lp -= 2.f * (0.69314718055994530942f - z - softplus20(-2.f * z));
    lp_part += lp;
    bad |= !(isfinite(mu) && isfinite(sd));
    if (op.o[6] >= 0) {
      base[op.o[6] + (i64)row * A + j] = tz;
      base[op.o[7] + (i64)row * A + j] = sd * e;
      base[op.o[8] + (i64)row * A + j] = (ls_raw >= hp.log_std_min && ls_raw <= hp.log_std_max) ? 1.f : 0.f;
    }
  }
  if (op.o[9] >= 0)
    for (int j = lane; j < 2 * A; j += 32) base[op.o[9] + (i64)row * 2 * A + j] = hs[2 * SACX_MAX_ACT + j];
  const float lp = warp_sum(lp_part);
  if (lane == 0) base[op.o[4] + row] = lp;
  if (bad) atomicOr(&c.scal->nonfinite, 1);
  __syncwarp();
}

// ---------------------------------------------------------------- OP_Q_ROW (target)
// o[0..1]=hqt(last hidden) o[2..3]=Wt_L o[4..5]=bt_L o[6]=r o[7]=d o[8]=lp2 o[9]=y o[10..11]=tq   i[0]=ldh i[1]=K
__device__ __forceinline__ void op_q_target(const Op& op, const RowCtx& c, int row, int lane) {
  const Hyper& hp = c.args->hp;
  if (row >= hp.B) return;
  float* base = c.base;
  float tq[2];
#pragma unroll
  for (int n = 0; n < 2; ++n) {
    const float z = warp_dot(base + op.o[n] + (i64)row * op.i[0], base + op.o[2 + n], op.i[1], lane) + __ldcg(base + op.o[4 + n]);
    tq[n] = act_fwd(op.act_out, z);
  }
  if (lane == 0) {
    const float alpha = __ldcg(&c.scal->alpha_f32);
    const float r = __ldcg(base + op.o[6] + row), d = __ldcg(base + op.o[7] + row), lp2 = __ldcg(base + op.o[8] + row);
    const float minq = fminf(tq[0], tq[1]);
    // y = r + gamma * (1 - d) * (minq - alpha * logpi')      (agent.py:208-210)
    const float y = r + (hp.gamma * (1.f - d)) * (minq - alpha * lp2);
    base[op.o[9] + row] = y;
    base[op.o[10] + row] = tq[0];
    base[op.o[11] + row] = tq[1];
  }
}

// ---------------------------------------------------------------- OP_CRITIC_ROW
// o[0..1]=hq(last hidden) o[2..3]=aux(z or h) o[4..5]=W_L o[6..7]=b_L o[8]=y o[9..10]=q out
// o[11..12]=dout o[13..14]=delta(last hidden) o[15..16]=lossrow    i[0]=ldh i[1]=K
__device__ __forceinline__ void op_critic_row(const Op& op, const RowCtx& c, int row, int lane) {
  const Hyper& hp = c.args->hp;
  if (row >= hp.B) return;
  float* base = c.base;
  const float y = c.args->y_ext ? c.args->y_ext[row] : __ldcg(base + op.o[8] + row);
  const int K = op.i[1], ld = op.i[0];
#pragma unroll
  for (int n = 0; n < 2; ++n) {
    const float* W = base + op.o[4 + n];
    const float z = warp_dot(base + op.o[n] + (i64)row * ld, W, K, lane) + __ldcg(base + op.o[6 + n]);
    const float q = act_fwd(op.act_out, z);
    const float diff = q - y;
    // d mse / d q = 2 (q - y) / B_global; through the output activation
    const float dout = (2.f * diff / (float)hp.B_global) * act_dz2(op.act_out, z, q);
    if (lane == 0) {
      base[op.o[9 + n] + row] = q;
      base[op.o[11 + n] + row] = dout;
      base[op.o[15 + n] + row] = diff * diff;
    }
    const float* aux = base + op.o[2 + n] + (i64)row * ld;
    float* dl = base + op.o[13 + n] + (i64)row * ld;
    for (int k = lane; k < K; k += 32) dl[k] = dout * __ldcg(W + k) * act_dz(op.act, __ldcg(aux + k));
  }
}

// ---------------------------------------------------------------- OP_ACTOR_Q
// o[0..1]=hq(last hidden) o[2..3]=aux o[4..5]=W_L o[6..7]=b_L o[8]=lp o[9..10]=q out o[13..14]=delta o[15]=plossrow
__device__ __forceinline__ void op_actor_q(const Op& op, const RowCtx& c, int row, int lane) {
  const Hyper& hp = c.args->hp;
  if (row >= hp.B) return;
  float* base = c.base;
  const int K = op.i[1], ld = op.i[0];
  float q[2], z[2];
#pragma unroll
  for (int n = 0; n < 2; ++n) {
    z[n] = warp_dot(base + op.o[n] + (i64)row * ld, base + op.o[4 + n], K, lane) + __ldcg(base + op.o[6 + n]);
    q[n] = act_fwd(op.act_out, z[n]);
  }
  const float alpha = __ldcg(&c.scal->alpha_f32);
  // torch.min backward: gradient to the smaller input, ties split 1/2 - 1/2
  const float w1 = q[0] < q[1] ? 1.f : (q[0] == q[1] ? 0.5f : 0.f);
  const float g[2] = {-w1 / (float)hp.B_global, -(1.f - w1) / (float)hp.B_global};
  if (lane == 0) {
    base[op.o[9] + row] = q[0];
    base[op.o[10] + row] = q[1];
    base[op.o[15] + row] = alpha * __ldcg(base + op.o[8] + row) - fminf(q[0], q[1]);
  }
#pragma unroll
  for (int n = 0; n < 2; ++n) {
    const float dout = g[n] * act_dz2(op.act_out, z[n], q[n]);
    const float* W = base + op.o[4 + n];
    const float* aux = base + op.o[2 + n] + (i64)row * ld;
    float* dl = base + op.o[13 + n] + (i64)row * ld;
    for (int k = lane; k < K; k += 32) dl[k] = dout * __ldcg(W + k) * act_dz(op.act, __ldcg(aux + k));
  }
}

// ---------------------------------------------------------------- OP_ACTOR_BWD
// o[0..1]=delta0 of critics [B,H0q] o[2..3]=W_0 of critics [H0q, obs+act]
// o[4]=tz o[5]=se o[6]=mask o[7]=headz o[8]=dhead out o[9]=Wpi_L o[10]=unused o[11]=aux pi(last hidden) o[12]=delta pi(last hidden)
// i[0]=ld delta0  i[1]=H0q  i[2]=ldW0 (=obs+act)  i[3]=ld pi hidden  i[4]=Kpi
__device__ __forceinline__ void op_actor_bwd(const Op& op, const RowCtx& c, int row, int lane) {
  const Hyper& hp = c.args->hp;
  if (row >= hp.B) return;
  float* base = c.base;
  const int A = hp.act, O = hp.obs, H0 = op.i[1];
  float* da = c.wsm;                          // [A] dQ/da, then dhead [2A]
  float* dh = c.wsm + SACX_MAX_ACT;
  // d(-minQ)/d a_j = sum_c sum_h delta0_c[h] * W0_c[h, obs + j]
  for (int j = 0; j < A; ++j) {
    float s = 0.f;
#pragma unroll
    for (int n = 0; n < 2; ++n) {
      const float* d0 = base + op.o[n] + (i64)row * op.i[0];
      const float* W0 = base + op.o[2 + n] + O + j;
      for (int k = lane; k < H0; k += 32) s = fmaf(__ldcg(d0 + k), __ldcg(W0 + (i64)k * op.i[2]), s);
    }
    s = warp_sum(s);
    if (lane == 0) da[j] = s;
  }
  __syncwarp();
  const float alpha = __ldcg(&c.scal->alpha_f32);
  const float ab = alpha / (float)hp.B_global;
  for (int j = lane; j < A; j += 32) {
    const float tz = __ldcg(base + op.o[4] + (i64)row * A + j);
    const float se = __ldcg(base + op.o[5] + (i64)row * A + j);
    const float mk = __ldcg(base + op.o[6] + (i64)row * A + j);
    // dL/dz = (alpha/B) 2 tanh z + dL/da * c (1 - tanh^2 z);  dL/dmu = dL/dz;
    // dL/dlogstd_raw = (sigma eps dL/dz - alpha/B) * 1[lo <= raw <= hi]
    const float dz = ab * (2.f * tz) + da[j] * (hp.action_scale * (1.f - tz * tz));
    float dmu = dz, dls = (se * dz - ab) * mk;
    if (op.act_out != SACX_ACT_IDENTITY) {
      const float zm = __ldcg(base + op.o[7] + (i64)row * 2 * A + j), zl = __ldcg(base + op.o[7] + (i64)row * 2 * A + A + j);
      dmu *= act_dz2(op.act_out, zm, act_fwd(op.act_out, zm));
      dls *= act_dz2(op.act_out, zl, act_fwd(op.act_out, zl));
    }
    dh[j] = dmu;
    dh[A + j] = dls;
    base[op.o[8] + (i64)row * 2 * A + j] = dmu;
    base[op.o[8] + (i64)row * 2 * A + A + j] = dls;
  }
  __syncwarp();
  // delta of the policy's last hidden layer: (dhead . Wpi_L) * act'(.)
  const int Kp = op.i[4];
  const float* W = base + op.o[9];
  const float* aux = base + op.o[11] + (i64)row * op.i[3];
  float* dl = base + op.o[12] + (i64)row * op.i[3];
  for (int k = lane; k < Kp; k += 32) {
    float s = 0.f;
    for (int j = 0; j < 2 * A; ++j) s = fmaf(dh[j], __ldcg(W + (i64)j * Kp + k), s);
    dl[k] = s * act_dz(op.act, __ldcg(aux + k));
  }
  __syncwarp();
}

// ---------------------------------------------------------------- OP_PROLOGUE (one thread)
// mode = bitmask of optimisers (1 << OptId) whose step advances in this update
__device__ __forceinline__ void op_prologue(const Op& op, const RowCtx& c) {
  AgentScalars* s = c.scal;
  const Hyper& hp = c.args->hp;
  for (int o = 0; o < 3; ++o) {
    if (!(op.mode & (1 << o))) continue;
    const i64 t = s->step[o] + 1;
    s->step[o] = t;
    // python-double bias corrections of torch's _single_tensor_adam
    const double bc1 = 1.0 - pow(0.9, (double)t);
    const double bc2 = 1.0 - pow(0.999, (double)t);
    s->adam_step_size[o] = (float)(hp.lr[o] / bc1);
    s->adam_bc2_sqrt[o] = (float)sqrt(bc2);
  }
}

// ---------------------------------------------------------------- OP_FINAL (one warp)
// mode bits: 1 critic loss means, 2 policy loss mean, 4 temperature step, 8 count the update
// o[0..1]=lossrow o[2]=plossrow o[3]=lp o[4..5]=q o[6]=y
__device__ __forceinline__ void op_final(const Op& op, const RowCtx& c, int lane) {
  const Hyper& hp = c.args->hp;
  float* base = c.base;
  AgentScalars* s = c.scal;
  const int B = hp.B;
  const float invB = 1.f / (float)hp.B_global;
  auto mean_of = [&](const float* p) {
    float acc = 0.f;
    for (int b = lane; b < B; b += 32) acc += __ldcg(p + b);
    return warp_sum(acc) * invB;
  };
  if (op.mode & 1) {
    const float l1 = mean_of(base + op.o[0]), l2 = mean_of(base + op.o[1]);
    const float q1 = mean_of(base + op.o[4]), q2 = mean_of(base + op.o[5]);
    const float ym = mean_of(base + op.o[6]);
    if (lane == 0) { s->metrics[0] = l1; s->metrics[1] = l2; s->metrics[6] = q1; s->metrics[7] = q2; s->metrics[9] = ym; }
  }
  if (op.mode & 2) {
    const float pl = mean_of(base + op.o[2]);
    if (lane == 0) s->metrics[2] = pl;
  }
  if (op.mode & 4) {
    const float* lp = c.args->lp_ext ? c.args->lp_ext : base + op.o[3];
    float acc = 0.f, accl = 0.f;
    const float la32 = (float)s->log_alpha;
    for (int b = lane; b < B; b += 32) {
      const float t = __ldcg(lp + b) + hp.target_entropy;
      acc += t;
      accl += la32 * t;
    }
    const float mean_t = warp_sum(acc) * invB;           // f32 mean, as in the reference
    const float mean_lt = warp_sum(accl) * invB;
    const float lpm = mean_t - hp.target_entropy;
    if (lane == 0) {
      s->metrics[8] = lpm;
      if (hp.auto_alpha) {
        // alpha_loss = -(log_alpha * (logpi + H).detach()).mean();  d/dlog_alpha = -mean(logpi + H)
        const double g = -(double)mean_t;
        const i64 t = s->step[OPT_ALPHA] + 1;
        s->step[OPT_ALPHA] = t;
        s->alpha_m = s->alpha_m + (1.0 - 0.9) * (g - s->alpha_m);
        s->alpha_v = s->alpha_v * 0.999 + (1.0 - 0.999) * g * g;
        const double bc1 = 1.0 - pow(0.9, (double)t), bc2 = 1.0 - pow(0.999, (double)t);
        const double denom = sqrt(s->alpha_v) / sqrt(bc2) + 1e-8;
        s->log_alpha = s->log_alpha - (hp.alpha_lr / bc1) * (s->alpha_m / denom);
        s->alpha = exp(s->log_alpha);
        s->alpha_f32 = (float)s->alpha;
        s->metrics[3] = -mean_lt;
      }
      s->metrics[4] = (float)s->alpha;
      s->metrics[5] = (float)s->log_alpha;
    }
  }
  if ((op.mode & 8) && lane == 0) s->updates += 1;
}

// ---------------------------------------------------------------- OP_POLYAK / OP_ADAM_FLAT (elementwise tiles)
constexpr int FLAT_TILE = 256 * 8;
// OP_POLYAK: o[0]=online o[1]=target o[2]=count
__device__ __forceinline__ void op_polyak(const Op& op, const RowCtx& c, int tile) {
  const Hyper& hp = c.args->hp;
  float* base = c.base;
  const i64 n = op.o[2];
  for (i64 e = (i64)tile * FLAT_TILE + threadIdx.x; e < n && e < (i64)(tile + 1) * FLAT_TILE; e += 256) {
    const float p = __ldcg(base + op.o[0] + e);
    float* t = base + op.o[1] + e;
    *t = polyak_mix(hp.tau, hp.one_minus_tau, p, __ldcg(t));
  }
}
// OP_ADAM_FLAT: o[0]=p o[1]=m o[2]=v o[3]=g o[4]=count o[5]=target or -1 (Polyak after the step)
__device__ __forceinline__ void op_adam_flat(const Op& op, const RowCtx& c, int tile) {
  const Hyper& hp = c.args->hp;
  float* base = c.base;
  const i64 n = op.o[4];
  const float ss = __ldcg(&c.scal->adam_step_size[op.opt]), bc = __ldcg(&c.scal->adam_bc2_sqrt[op.opt]);
  for (i64 e = (i64)tile * FLAT_TILE + threadIdx.x; e < n && e < (i64)(tile + 1) * FLAT_TILE; e += 256) {
    float p = __ldcg(base + op.o[0] + e), m = __ldcg(base + op.o[1] + e), v = __ldcg(base + op.o[2] + e);
    adam_update(__ldcg(base + op.o[3] + e), p, m, v, ss, bc);
    base[op.o[0] + e] = p;
    base[op.o[1] + e] = m;
    base[op.o[2] + e] = v;
    if (op.o[5] >= 0) {
      float* t = base + op.o[5] + e;
      *t = polyak_mix(hp.tau, hp.one_minus_tau, p, __ldcg(t));
    }
  }
}

}  // namespace sacx
