// ce.cu — K6 (v1): label-smoothed softmax cross-entropy over fp32 logits, forward + d_logits in one kernel,
// followed by the per-sentence / batch-mean reduction.
// Replaces util.label_smooth (utils/util.py:88-103), softmax_cross_entropy_with_logits_v2 and the masked
// per-sample mean (models/transformer.py:198-211).  The reference materialises a dense one-hot soft-label tensor
// [rows, V]; here the smoothed target is implicit:
//   nll = -(p * lsm[gold] + q * (sum_j lsm[j] - lsm[gold])) - normaliser,  p = 1 - eps, q = eps / (V - 1)
//   d nll / d logit_j = softmax_j - (j == gold ? p : q)
// One CTA per token row; pass 1 = online max / sum-exp / sum-logit, pass 2 = gradient (row re-read from L2).
// Algorithmic bytes / row: V * 4 (logits read) + V * 2 (bf16 gradient write).
#include <math.h>
#include <stdlib.h>

#include "zb_common.h"
#include "zb_ptx.cuh"

namespace zb {

constexpr int kCeThreads = 256;

__device__ __forceinline__ void online_merge(float& m, float& s, float m2, float s2) {
  if (m2 == -INFINITY) return;
  if (m == -INFINITY) {
    m = m2;
    s = s2;
    return;
  }
  const float nm = fmaxf(m, m2);
  s = s * __expf(m - nm) + s2 * __expf(m2 - nm);
  m = nm;
}

__global__ void __launch_bounds__(kCeThreads)
softmax_ce_kernel(const float* __restrict__ logits, long long ld, const int32_t* __restrict__ labels, int batch,
                  int seq_len, float* __restrict__ nll, __nv_bfloat16* __restrict__ d_logits, long long ldd, int vocab,
                  float smooth, float loss_scale) {
  grid_dep_wait();
  __shared__ float sm_m[kCeThreads / 32], sm_s[kCeThreads / 32], sm_t[kCeThreads / 32];
  __shared__ float sh_lse, sh_sumlogit, sh_w;
  const long long row = blockIdx.x;
  const float* lr = logits + row * ld;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float m = -INFINITY, s = 0.f, tot = 0.f;
  const int v4 = (vocab % 4 == 0 && (ld % 4) == 0) ? vocab / 4 : 0;
  for (int i = tid; i < v4; i += kCeThreads) {
    const float4 x = __ldg(reinterpret_cast<const float4*>(lr) + i);
    const float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      tot += xs[e];
      if (xs[e] > m) {
        s = s * __expf(m - xs[e]) + 1.f;
        m = xs[e];
      } else {
        s += __expf(xs[e] - m);
      }
    }
  }
  for (int i = v4 * 4 + tid; i < vocab; i += kCeThreads) {
    const float x = lr[i];
    tot += x;
    if (x > m) {
      s = s * __expf(m - x) + 1.f;
      m = x;
    } else {
      s += __expf(x - m);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
    online_merge(m, s, m2, s2);
    tot += __shfl_xor_sync(0xffffffffu, tot, o);
  }
  if (lane == 0) {
    sm_m[warp] = m;
    sm_s[warp] = s;
    sm_t[warp] = tot;
  }
  __syncthreads();
  if (tid == 0) {
    float M = sm_m[0], S = sm_s[0], T = sm_t[0];
    for (int w = 1; w < kCeThreads / 32; ++w) {
      online_merge(M, S, sm_m[w], sm_s[w]);
      T += sm_t[w];
    }
    const float lse = M + logf(S);
    sh_lse = lse;
    sh_sumlogit = T;
    const int gold = labels[row];
    const float lg = lr[gold < 0 ? 0 : (gold >= vocab ? vocab - 1 : gold)] - lse;
    float val;
    if (smooth > 0.f && smooth < 1.f) {
      const float n = (float)(vocab - 1);
      const float p = 1.f - smooth, q = smooth / n;
      const float norm = -(p * logf(p) + n * q * logf(q + 1e-20f));
      const float sum_lsm = T - (float)vocab * lse;
      val = -(p * lg + q * (sum_lsm - lg)) - norm;
    } else {
      val = -lg;
    }
    nll[row] = val;
    // d loss / d nll for this token: mask / (len_b * batch)   (models/transformer.py:208-210)
    float w = 0.f;
    if (d_logits) {
      const int b = (int)(row / seq_len);
      int len = 0;
      for (int t = 0; t < seq_len; ++t) len += labels[(long long)b * seq_len + t] != 0;
      if (labels[row] != 0 && len > 0) w = loss_scale / ((float)len * (float)batch);
    }
    sh_w = w;
  }
  if (!d_logits) return;
  __syncthreads();
  const float lse = sh_lse, w = sh_w;
  const int gold = labels[row];
  float p = 1.f, q = 0.f;
  if (smooth > 0.f && smooth < 1.f) {
    p = 1.f - smooth;
    q = smooth / (float)(vocab - 1);
  }
  __nv_bfloat16* dr = d_logits + row * ldd;
  const int v8 = (vocab % 8 == 0 && ld % 4 == 0 && ldd % 8 == 0) ? vocab / 8 : 0;
  for (int i = tid; i < v8; i += kCeThreads) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(lr) + 2 * i);
    const float4 b = __ldg(reinterpret_cast<const float4*>(lr) + 2 * i + 1);
    const float xs[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    float g[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) g[e] = w * (__expf(xs[e] - lse) - ((i * 8 + e) == gold ? p : q));
    uint4 o;
    o.x = pack_bf16x2(g[0], g[1]);
    o.y = pack_bf16x2(g[2], g[3]);
    o.z = pack_bf16x2(g[4], g[5]);
    o.w = pack_bf16x2(g[6], g[7]);
    reinterpret_cast<uint4*>(dr)[i] = o;
  }
  for (int i = v8 * 8 + tid; i < vocab; i += kCeThreads)
    dr[i] = __float2bfloat16(w * (__expf(lr[i] - lse) - (i == gold ? p : q)));
}

// Register-resident variant for vocab <= 32768 (vocab % 4 == 0): one 1024-thread CTA per token row, the whole row
// (<= 8 float4 per thread) is loaded ONCE with every load in flight before the first use, exponentiated once, and
// the gradient is written from registers — one HBM pass (V * 4 B read + V * 2 B write per row), no L2 re-read.
constexpr int kCeRegThreads = 1024;
constexpr int kCeRegF4 = 8;

__device__ __forceinline__ float block_reduce_1024(float v, float* sm, bool is_max) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float t = __shfl_xor_sync(0xffffffffu, v, o);
    v = is_max ? fmaxf(v, t) : v + t;
  }
  if (lane == 0) sm[warp] = v;
  __syncthreads();
  v = sm[lane];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float t = __shfl_xor_sync(0xffffffffu, v, o);
    v = is_max ? fmaxf(v, t) : v + t;
  }
  __syncthreads();  // sm is reused by the next reduction
  return v;
}

__global__ void __launch_bounds__(kCeRegThreads, 1)
softmax_ce_reg_kernel(const float* __restrict__ logits, long long ld, const int32_t* __restrict__ labels, int batch,
                      int seq_len, float* __restrict__ nll, __nv_bfloat16* __restrict__ d_logits, long long ldd,
                      int vocab, float smooth, float loss_scale) {
  grid_dep_wait();
  __shared__ float sm[32];
  __shared__ int sh_len;
  const long long row = blockIdx.x;
  const float* lr = logits + row * ld;
  const int tid = threadIdx.x;
  const int v4 = vocab >> 2;
  float4 x[kCeRegF4];
#pragma unroll
  for (int j = 0; j < kCeRegF4; ++j) {
    const int i = tid + j * kCeRegThreads;
    x[j] = (i < v4) ? __ldg(reinterpret_cast<const float4*>(lr) + i)
                    : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
  }
  const int gold = labels[row];
  if (tid < 32 && d_logits) {  // number of non-pad target tokens of this row's sentence (models/transformer.py:208)
    const int b = (int)(row / seq_len);
    int len = 0;
    for (int t = tid; t < seq_len; t += 32) len += labels[(long long)b * seq_len + t] != 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) len += __shfl_xor_sync(0xffffffffu, len, o);
    if (tid == 0) sh_len = len;
  }
  float m = -INFINITY, tot = 0.f;
#pragma unroll
  for (int j = 0; j < kCeRegF4; ++j) {
    if (tid + j * kCeRegThreads < v4) {
      m = fmaxf(m, fmaxf(fmaxf(x[j].x, x[j].y), fmaxf(x[j].z, x[j].w)));
      tot += (x[j].x + x[j].y) + (x[j].z + x[j].w);
    }
  }
  const float M = block_reduce_1024(m, sm, true);
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < kCeRegF4; ++j) {
    x[j].x = __expf(x[j].x - M);  // exp(-inf) = 0 for the slots beyond the row
    x[j].y = __expf(x[j].y - M);
    x[j].z = __expf(x[j].z - M);
    x[j].w = __expf(x[j].w - M);
    s += (x[j].x + x[j].y) + (x[j].z + x[j].w);
  }
  const float S = block_reduce_1024(s, sm, false);
  const float T = block_reduce_1024(tot, sm, false);
  const float lse = M + logf(S);
  float p = 1.f, q = 0.f;
  if (smooth > 0.f && smooth < 1.f) {
    p = 1.f - smooth;
    q = smooth / (float)(vocab - 1);
  }
  if (tid == 0) {
    const float lg = lr[gold < 0 ? 0 : (gold >= vocab ? vocab - 1 : gold)] - lse;
    float val;
    if (smooth > 0.f && smooth < 1.f) {
      const float n = (float)(vocab - 1);
      const float norm = -(p * logf(p) + n * q * logf(q + 1e-20f));
      const float sum_lsm = T - (float)vocab * lse;
      val = -(p * lg + q * (sum_lsm - lg)) - norm;
    } else {
      val = -lg;
    }
    nll[row] = val;
  }
  if (!d_logits) return;
  float w = 0.f;
  {
    const int len = sh_len;  // written before the first __syncthreads of block_reduce_1024
    if (gold != 0 && len > 0) w = loss_scale / ((float)len * (float)batch);
  }
  const float wS = w / S;
  const float wp = w * p, wq = w * q;
  __nv_bfloat16* dr = d_logits + row * ldd;
#pragma unroll
  for (int j = 0; j < kCeRegF4; ++j) {
    const int i = tid + j * kCeRegThreads;
    if (i < v4) {
      const int c = i * 4;
      const float g0 = x[j].x * wS - (c == gold ? wp : wq), g1 = x[j].y * wS - (c + 1 == gold ? wp : wq),
                  g2 = x[j].z * wS - (c + 2 == gold ? wp : wq), g3 = x[j].w * wS - (c + 3 == gold ? wp : wq);
      uint2 o;
      o.x = pack_bf16x2(g0, g1);
      o.y = pack_bf16x2(g2, g3);
      reinterpret_cast<uint2*>(dr)[i] = o;
    }
  }
}

// per_sample[b] = sum_t nll * mask / sum_t mask ; loss = mean_b per_sample  (models/transformer.py:208-216)
__global__ void ce_reduce_kernel(const float* __restrict__ nll, const int32_t* __restrict__ labels, int batch,
                                 int seq_len, float* __restrict__ per_sample, float* __restrict__ loss) {
  grid_dep_wait();
  __shared__ float acc[32];
  float local = 0.f;
  for (int b = threadIdx.x; b < batch; b += blockDim.x) {
    float s = 0.f, n = 0.f;
    for (int t = 0; t < seq_len; ++t) {
      const long long i = (long long)b * seq_len + t;
      if (labels[i] != 0) {
        s += nll[i];
        n += 1.f;
      }
    }
    const float ps = s / n;  // 0/0 = NaN exactly like the reference for an all-pad row
    if (per_sample) per_sample[b] = ps;
    local += ps;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  if ((threadIdx.x & 31) == 0) acc[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0 && loss) {
    float t = 0.f;
    for (int w = 0; w < (int)(blockDim.x + 31) / 32; ++w) t += acc[w];
    loss[0] = batch > 0 ? t / (float)batch : 0.f;
  }
}

// per-sentence / batch means of the per-token NLL (also used by vocab_ce.cu)
int ce_reduce_launch(const float* nll, const int32_t* labels, int batch, int seq_len, float* per_sample, float* loss,
                     cudaStream_t st) {
  ZB_LAUNCH(ce_reduce_kernel, 1, 256, 0, st, nll, labels, batch, seq_len, per_sample, loss);
  return check_launch("zb_ce_reduce");
}

}  // namespace zb

extern "C" int zb_softmax_ce(const zb_ce_args* a, zb_stream_t stream) {
  using namespace zb;
  ZB_REQUIRE(a && a->logits && a->labels && a->nll, "zb_softmax_ce: null pointer");
  ZB_REQUIRE(a->vocab > 1 && a->batch >= 0 && a->seq_len > 0, "zb_softmax_ce: bad shape");
  const long long rows = (long long)a->batch * a->seq_len;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  static const bool no_reg = getenv("ZB_CE_TWO_PASS") != nullptr;
  const bool reg_ok = !no_reg && a->vocab % 4 == 0 && a->ld % 4 == 0 && a->vocab <= 4 * kCeRegF4 * kCeRegThreads &&
                      (reinterpret_cast<uintptr_t>(a->logits) & 15) == 0 &&
                      (!a->d_logits || (a->ldd % 4 == 0 && (reinterpret_cast<uintptr_t>(a->d_logits) & 7) == 0));
  if (rows > 0 && reg_ok) {
    ZB_LAUNCH(softmax_ce_reg_kernel, (unsigned)rows, kCeRegThreads, 0, st, a->logits, a->ld, a->labels, a->batch,
              a->seq_len, a->nll, (__nv_bfloat16*)a->d_logits, a->ldd, a->vocab, a->smooth, a->loss_scale);
    int rc = check_launch("zb_softmax_ce(reg)");
    if (rc) return rc;
  } else if (rows > 0) {
    ZB_LAUNCH(softmax_ce_kernel, (unsigned)rows, kCeThreads, 0, st, a->logits, a->ld, a->labels, a->batch, a->seq_len, a->nll,
                                                            (__nv_bfloat16*)a->d_logits, a->ldd, a->vocab, a->smooth,
                                                            a->loss_scale);
    int rc = check_launch("zb_softmax_ce");
    if (rc) return rc;
  }
  if (a->per_sample || a->loss) {
    ZB_LAUNCH(ce_reduce_kernel, 1, 256, 0, st, a->nll, a->labels, a->batch, a->seq_len, a->per_sample, a->loss);
    return check_launch("zb_ce_reduce");
  }
  return ZB_OK;
}
