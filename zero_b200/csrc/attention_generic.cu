// attention_generic.cu — K2/K3 (feature-complete path): fused scaled-dot attention forward / backward with every
// variant Zero's Transformer family uses, logits never leaving the SM:
//   * softmax attention with additive -inf masks from key lengths / causal index  (func.py:218-256, 372-388)
//   * Shaw relative-position terms on keys and values, bucket = clip(i - j, -k, k) + k  (modules/rpr.py:10-75)
//   * ReLA: relu(logits * keep) instead of softmax                                 (modules/rela.py:52-75)
//   * cached decode (lq = 1, q_offset = time)                                       (func.py:199-205)
// CUDA-core (fp32 FMA) implementation: one warp per query row (fwd, dq) or key row (dk/dv), 32-wide key/query
// tiles staged in shared memory, online softmax.  attention_mma.cu holds the tensor-core fast path for the
// plain softmax case; this file is the path used when rpr / ReLA / odd head sizes are requested.
#include <math.h>
#include <stdlib.h>

#include "zb_common.h"
#include "zb_ptx.cuh"

namespace zb {

constexpr int kAttnWarps = 8;

struct AttnP {
  const __nv_bfloat16 *q, *k, *v, *d_o, *rpr_k, *rpr_v;
  __nv_bfloat16 *o, *dq, *dk, *dv;
  const __nv_bfloat16* o_in;
  long long ldq, ldk, ldv, ldo, bsq, bsk, bsv, bso;
  long long lddo, lddq, lddk, lddv, bsdo, bsdq, bsdk, bsdv;
  int batch, heads, lq, lk;
  const int32_t* key_len;
  int causal, q_offset, max_rel, relu_attn, kv_group;
  float scale, inf_value;
  float *lse, *delta, *d_rpr_k, *d_rpr_v;
  float drop_rate;                        // attention dropout (func.py:245): 0 = off
  uint32_t drop_site;
  const unsigned long long* drop_seed;
};

// multiplier of attention weight (b, h, i, j) under dropout: 0 or 1 / keep
struct DropCtx {
  bool on;
  uint64_t seed;
  uint32_t site, thr;
  float inv_keep;
  __device__ __forceinline__ explicit DropCtx(const AttnP& p) {
    on = p.drop_seed != nullptr && p.drop_rate > 0.f;
    seed = on ? *p.drop_seed : 0ull;
    site = p.drop_site;
    thr = dropout_threshold(p.drop_rate);
    inv_keep = on ? 1.f / (1.f - p.drop_rate) : 1.f;
  }
  __device__ __forceinline__ float mul(const AttnP& p, int b, int h, int i, int j) const {
    if (!on) return 1.f;
    const uint64_t idx = (((uint64_t)b * p.heads + h) * p.lq + i) * (uint64_t)p.lk + j;
    return dropout_mul(seed, site, idx, thr, inv_keep);
  }
};

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float wmax(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ int rel_bucket(int i_abs, int j, int R) {
  int d = i_abs - j;
  d = d < -R ? -R : (d > R ? R : d);
  return d + R;
}

// cooperative load of `rows` x DH bf16 (row pitch ld) into fp32 smem [32][DH+1]; rows beyond `valid` are zeroed
template <int DH>
__device__ __forceinline__ void load_tile(float (*dst)[DH + 1], const __nv_bfloat16* src, long long ld, int valid,
                                          float mul) {
  for (int idx = threadIdx.x; idx < 32 * DH; idx += kAttnWarps * 32) {
    const int r = idx / DH, d = idx % DH;
    dst[r][d] = r < valid ? __bfloat162float(src[(long long)r * ld + d]) * mul : 0.f;
  }
}

template <int DH>
struct Smem {
  float a[32][DH + 1];
  float b[32][DH + 1];
  float aux0[32];
  float aux1[32];
};

// ------------------------------------------------------------------------------------------------ forward
template <int DH>
__global__ void __launch_bounds__(kAttnWarps * 32, 1) attn_fwd_generic(const AttnP p) {
  grid_dep_wait();
  constexpr int ND = (DH + 31) / 32;
  extern __shared__ float smem_f[];
  Smem<DH>& sm = *reinterpret_cast<Smem<DH>*>(smem_f);
  const int NB = 2 * p.max_rel + 1;
  float(*ek)[DH + 1] = reinterpret_cast<float(*)[DH + 1]>(smem_f + sizeof(Smem<DH>) / 4);
  float(*ev)[DH + 1] = ek + NB;
  const bool rpr = p.rpr_k != nullptr;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.z, h = blockIdx.y;
  const int i = blockIdx.x * kAttnWarps + warp;
  const bool active = i < p.lq;
  const int kl = p.key_len ? p.key_len[b / p.kv_group] : p.lk;
  if (rpr) {
    for (int idx = threadIdx.x; idx < NB * DH; idx += blockDim.x) {
      ek[idx / DH][idx % DH] = __bfloat162float(p.rpr_k[idx]);
      ev[idx / DH][idx % DH] = __bfloat162float(p.rpr_v[idx]);
    }
  }
  float qr[DH];
  if (active) {
    const __nv_bfloat16* qp = p.q + (long long)b * p.bsq + (long long)i * p.ldq + h * DH;
#pragma unroll
    for (int d = 0; d < DH; ++d) qr[d] = __bfloat162float(qp[d]) * p.scale;
  }
  float m = -INFINITY, l = 0.f, o[ND];
#pragma unroll
  for (int dd = 0; dd < ND; ++dd) o[dd] = 0.f;
  const int i_abs = i + p.q_offset;
  const DropCtx drop(p);

  for (int kt = 0; kt < p.lk; kt += 32) {
    __syncthreads();
    const int valid_rows = min(32, p.lk - kt);
    load_tile<DH>(sm.a, p.k + (long long)(b / p.kv_group) * p.bsk + (long long)kt * p.ldk + h * DH, p.ldk, valid_rows, 1.f);
    load_tile<DH>(sm.b, p.v + (long long)(b / p.kv_group) * p.bsv + (long long)kt * p.ldv + h * DH, p.ldv, valid_rows, 1.f);
    __syncthreads();
    if (!active) continue;
    const int j = kt + lane;
    const bool inb = j < p.lk;
    float s = 0.f;
#pragma unroll
    for (int d = 0; d < DH; ++d) s += qr[d] * sm.a[lane][d];
    int bucket = 0;
    if (rpr) {
      bucket = rel_bucket(i_abs, j, p.max_rel);
#pragma unroll
      for (int d = 0; d < DH; ++d) s += qr[d] * ek[bucket][d];
    }
    const bool valid = inb && j < kl && (!p.causal || j <= i_abs);
    float pj;
    if (p.relu_attn) {
      pj = valid ? fmaxf(s, 0.f) : 0.f;
    } else {
      s = inb ? (valid ? s : s - p.inf_value) : -INFINITY;
      const float m_new = fmaxf(m, wmax(s));
      pj = inb ? __expf(s - m_new) : 0.f;
      const float corr = __expf(m - m_new);
      l = l * corr + wsum(pj);
      m = m_new;
#pragma unroll
      for (int dd = 0; dd < ND; ++dd) o[dd] *= corr;
    }
    // dropout acts on the normalised weights (the 1 / l is applied at the end), not on the normaliser
    if (inb) pj *= drop.mul(p, b, h, i, j);
#pragma unroll 4
    for (int jj = 0; jj < 32; ++jj) {
      const float pv = __shfl_sync(0xffffffffu, pj, jj);
      const int bk = __shfl_sync(0xffffffffu, bucket, jj);
#pragma unroll
      for (int dd = 0; dd < ND; ++dd) {
        const int d = lane + 32 * dd;
        if (d < DH) {
          float vv = sm.b[jj][d];
          if (rpr) vv += ev[bk][d];
          o[dd] += pv * vv;
        }
      }
    }
  }
  if (!active) return;
  const float inv = p.relu_attn ? 1.f : 1.f / l;
  __nv_bfloat16* op = p.o + (long long)b * p.bso + (long long)i * p.ldo + h * DH;
#pragma unroll
  for (int dd = 0; dd < ND; ++dd) {
    const int d = lane + 32 * dd;
    if (d < DH) op[d] = __float2bfloat16(o[dd] * inv);
  }
  if (lane == 0 && p.lse) p.lse[((long long)b * p.heads + h) * p.lq + i] = p.relu_attn ? 0.f : m + __logf(l);
}

// ------------------------------------------------------------------------------------------------ backward: dq
template <int DH>
__global__ void __launch_bounds__(kAttnWarps * 32, 1) attn_bwd_dq_generic(const AttnP p) {
  grid_dep_wait();
  constexpr int ND = (DH + 31) / 32;
  extern __shared__ float smem_f[];
  Smem<DH>& sm = *reinterpret_cast<Smem<DH>*>(smem_f);
  const int NB = 2 * p.max_rel + 1;
  float(*ek)[DH + 1] = reinterpret_cast<float(*)[DH + 1]>(smem_f + sizeof(Smem<DH>) / 4);
  float(*ev)[DH + 1] = ek + NB;
  float(*dek)[DH + 1] = ev + NB;
  float(*dev)[DH + 1] = dek + NB;
  const bool rpr = p.rpr_k != nullptr;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.z, h = blockIdx.y;
  const int i = blockIdx.x * kAttnWarps + warp;
  const bool active = i < p.lq;
  const int kl = p.key_len ? p.key_len[b / p.kv_group] : p.lk;
  if (rpr) {
    for (int idx = threadIdx.x; idx < NB * DH; idx += blockDim.x) {
      ek[idx / DH][idx % DH] = __bfloat162float(p.rpr_k[idx]);
      ev[idx / DH][idx % DH] = __bfloat162float(p.rpr_v[idx]);
      dek[idx / DH][idx % DH] = 0.f;
      dev[idx / DH][idx % DH] = 0.f;
    }
  }
  float qr[DH], dor[DH], qo[ND], doo[ND], dq[ND];
  float lse = 0.f, delta = 0.f;
  if (active) {
    const __nv_bfloat16* qp = p.q + (long long)b * p.bsq + (long long)i * p.ldq + h * DH;
    const __nv_bfloat16* dop = p.d_o + (long long)b * p.bsdo + (long long)i * p.lddo + h * DH;
    const __nv_bfloat16* oin = p.o_in + (long long)b * p.bso + (long long)i * p.ldo + h * DH;
    float dl = 0.f;
#pragma unroll
    for (int d = 0; d < DH; ++d) {
      qr[d] = __bfloat162float(qp[d]) * p.scale;
      dor[d] = __bfloat162float(dop[d]);
    }
#pragma unroll
    for (int dd = 0; dd < ND; ++dd) {
      const int d = lane + 32 * dd;
      qo[dd] = d < DH ? __bfloat162float(qp[d]) * p.scale : 0.f;
      doo[dd] = d < DH ? __bfloat162float(dop[d]) : 0.f;
      dl += d < DH ? doo[dd] * __bfloat162float(oin[d]) : 0.f;
      dq[dd] = 0.f;
    }
    delta = wsum(dl);
    const long long li = ((long long)b * p.heads + h) * p.lq + i;
    lse = p.lse[li];
    if (lane == 0) p.delta[li] = delta;
  }
  const int i_abs = i + p.q_offset;
  const DropCtx drop(p);
  for (int kt = 0; kt < p.lk; kt += 32) {
    __syncthreads();
    const int valid_rows = min(32, p.lk - kt);
    load_tile<DH>(sm.a, p.k + (long long)(b / p.kv_group) * p.bsk + (long long)kt * p.ldk + h * DH, p.ldk, valid_rows, 1.f);
    load_tile<DH>(sm.b, p.v + (long long)(b / p.kv_group) * p.bsv + (long long)kt * p.ldv + h * DH, p.ldv, valid_rows, 1.f);
    __syncthreads();
    if (!active) continue;
    const int j = kt + lane;
    const bool inb = j < p.lk;
    float s = 0.f, dp = 0.f;
#pragma unroll
    for (int d = 0; d < DH; ++d) {
      s += qr[d] * sm.a[lane][d];
      dp += dor[d] * sm.b[lane][d];
    }
    int bucket = 0;
    if (rpr) {
      bucket = rel_bucket(i_abs, j, p.max_rel);
#pragma unroll
      for (int d = 0; d < DH; ++d) {
        s += qr[d] * ek[bucket][d];
        dp += dor[d] * ev[bucket][d];
      }
    }
    const bool valid = inb && j < kl && (!p.causal || j <= i_abs);
    const float dm = inb ? drop.mul(p, b, h, i, j) : 0.f;
    dp *= dm;  // gradient wrt the undropped weight
    float pj, ds;
    if (p.relu_attn) {
      pj = valid ? fmaxf(s, 0.f) : 0.f;
      ds = (valid && s > 0.f) ? dp : 0.f;
    } else {
      const float sm_ = valid ? s : s - p.inf_value;
      pj = inb ? __expf(sm_ - lse) : 0.f;
      ds = pj * (dp - delta);
    }
    pj *= dm;  // the dropped weight is what multiplied the values
#pragma unroll 4
    for (int jj = 0; jj < 32; ++jj) {
      const float dsv = __shfl_sync(0xffffffffu, ds, jj);
      const float pv = __shfl_sync(0xffffffffu, pj, jj);
      const int bk = __shfl_sync(0xffffffffu, bucket, jj);
#pragma unroll
      for (int dd = 0; dd < ND; ++dd) {
        const int d = lane + 32 * dd;
        if (d < DH) {
          float kk = sm.a[jj][d];
          if (rpr) {
            kk += ek[bk][d];
            if (dsv != 0.f) atomicAdd(&dek[bk][d], dsv * qo[dd]);
            if (pv != 0.f) atomicAdd(&dev[bk][d], pv * doo[dd]);
          }
          dq[dd] += dsv * kk;
        }
      }
    }
  }
  if (active) {
    __nv_bfloat16* dqp = p.dq + (long long)b * p.bsdq + (long long)i * p.lddq + h * DH;
#pragma unroll
    for (int dd = 0; dd < ND; ++dd) {
      const int d = lane + 32 * dd;
      if (d < DH) dqp[d] = __float2bfloat16(dq[dd] * p.scale);
    }
  }
  if (rpr) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < NB * DH; idx += blockDim.x) {
      const float a = dek[idx / DH][idx % DH], c = dev[idx / DH][idx % DH];
      if (a != 0.f) atomicAdd(p.d_rpr_k + idx, a);
      if (c != 0.f) atomicAdd(p.d_rpr_v + idx, c);
    }
  }
}

// ------------------------------------------------------------------------------------------------ backward: dk, dv
template <int DH>
__global__ void __launch_bounds__(kAttnWarps * 32, 1) attn_bwd_dkv_generic(const AttnP p) {
  grid_dep_wait();
  constexpr int ND = (DH + 31) / 32;
  extern __shared__ float smem_f[];
  Smem<DH>& sm = *reinterpret_cast<Smem<DH>*>(smem_f);
  const int NB = 2 * p.max_rel + 1;
  float(*ek)[DH + 1] = reinterpret_cast<float(*)[DH + 1]>(smem_f + sizeof(Smem<DH>) / 4);
  float(*ev)[DH + 1] = ek + NB;
  const bool rpr = p.rpr_k != nullptr;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.z, h = blockIdx.y;
  const int j = blockIdx.x * kAttnWarps + warp;
  const bool active = j < p.lk;
  const int kl = p.key_len ? p.key_len[b / p.kv_group] : p.lk;
  if (rpr) {
    for (int idx = threadIdx.x; idx < NB * DH; idx += blockDim.x) {
      ek[idx / DH][idx % DH] = __bfloat162float(p.rpr_k[idx]);
      ev[idx / DH][idx % DH] = __bfloat162float(p.rpr_v[idx]);
    }
  }
  float kr[DH], vr[DH], dk[ND], dv[ND];
  if (active) {
    const __nv_bfloat16* kp = p.k + (long long)b * p.bsk + (long long)j * p.ldk + h * DH;
    const __nv_bfloat16* vp = p.v + (long long)b * p.bsv + (long long)j * p.ldv + h * DH;
#pragma unroll
    for (int d = 0; d < DH; ++d) {
      kr[d] = __bfloat162float(kp[d]);
      vr[d] = __bfloat162float(vp[d]);
    }
  }
#pragma unroll
  for (int dd = 0; dd < ND; ++dd) dk[dd] = dv[dd] = 0.f;
  const bool key_ok = j < kl;
  const DropCtx drop(p);
  for (int qt = 0; qt < p.lq; qt += 32) {
    __syncthreads();
    const int valid_rows = min(32, p.lq - qt);
    load_tile<DH>(sm.a, p.q + (long long)b * p.bsq + (long long)qt * p.ldq + h * DH, p.ldq, valid_rows, p.scale);
    load_tile<DH>(sm.b, p.d_o + (long long)b * p.bsdo + (long long)qt * p.lddo + h * DH, p.lddo, valid_rows, 1.f);
    if (threadIdx.x < 32) {
      const int i = qt + threadIdx.x;
      const long long li = ((long long)b * p.heads + h) * p.lq + i;
      sm.aux0[threadIdx.x] = i < p.lq ? p.lse[li] : 0.f;
      sm.aux1[threadIdx.x] = i < p.lq ? p.delta[li] : 0.f;
    }
    __syncthreads();
    if (!active) continue;
    const int i = qt + lane;
    const int i_abs = i + p.q_offset;
    const bool inb = i < p.lq;
    float s = 0.f, dp = 0.f;
    int bucket = 0;
    if (rpr) bucket = rel_bucket(i_abs, j, p.max_rel);
#pragma unroll
    for (int d = 0; d < DH; ++d) {
      float kk = kr[d], vv = vr[d];
      if (rpr) {
        kk += ek[bucket][d];
        vv += ev[bucket][d];
      }
      s += sm.a[lane][d] * kk;
      dp += sm.b[lane][d] * vv;
    }
    const bool valid = inb && key_ok && (!p.causal || j <= i_abs);
    const float dm = inb ? drop.mul(p, b, h, i, j) : 0.f;
    dp *= dm;
    float pj, ds;
    if (p.relu_attn) {
      pj = valid ? fmaxf(s, 0.f) : 0.f;
      ds = (valid && s > 0.f) ? dp : 0.f;
    } else {
      const float sm_ = valid ? s : s - p.inf_value;
      pj = inb ? __expf(sm_ - sm.aux0[lane]) : 0.f;
      ds = pj * (dp - sm.aux1[lane]);
    }
    pj *= dm;
#pragma unroll 4
    for (int ii = 0; ii < 32; ++ii) {
      const float pv = __shfl_sync(0xffffffffu, pj, ii);
      const float dsv = __shfl_sync(0xffffffffu, ds, ii);
#pragma unroll
      for (int dd = 0; dd < ND; ++dd) {
        const int d = lane + 32 * dd;
        if (d < DH) {
          dv[dd] += pv * sm.b[ii][d];
          dk[dd] += dsv * sm.a[ii][d];
        }
      }
    }
  }
  if (!active) return;
  __nv_bfloat16* dkp = p.dk + (long long)b * p.bsdk + (long long)j * p.lddk + h * DH;
  __nv_bfloat16* dvp = p.dv + (long long)b * p.bsdv + (long long)j * p.lddv + h * DH;
#pragma unroll
  for (int dd = 0; dd < ND; ++dd) {
    const int d = lane + 32 * dd;
    if (d < DH) {
      dkp[d] = __float2bfloat16(dk[dd]);
      dvp[d] = __float2bfloat16(dv[dd]);
    }
  }
}

static AttnP to_params(const zb_attention_args* a) {
  AttnP p;
  p.q = (const __nv_bfloat16*)a->q; p.k = (const __nv_bfloat16*)a->k; p.v = (const __nv_bfloat16*)a->v;
  p.o = (__nv_bfloat16*)a->o; p.o_in = (const __nv_bfloat16*)a->o;
  p.d_o = (const __nv_bfloat16*)a->d_o; p.dq = (__nv_bfloat16*)a->dq; p.dk = (__nv_bfloat16*)a->dk;
  p.dv = (__nv_bfloat16*)a->dv;
  p.rpr_k = (const __nv_bfloat16*)a->rpr_k; p.rpr_v = (const __nv_bfloat16*)a->rpr_v;
  p.ldq = a->ldq; p.ldk = a->ldk; p.ldv = a->ldv; p.ldo = a->ldo;
  p.bsq = a->bsq; p.bsk = a->bsk; p.bsv = a->bsv; p.bso = a->bso;
  p.lddo = a->lddo; p.lddq = a->lddq; p.lddk = a->lddk; p.lddv = a->lddv;
  p.bsdo = a->bsdo; p.bsdq = a->bsdq; p.bsdk = a->bsdk; p.bsdv = a->bsdv;
  p.batch = a->batch; p.heads = a->heads; p.lq = a->lq; p.lk = a->lk;
  p.key_len = a->key_len; p.causal = a->causal; p.q_offset = a->q_offset;
  p.max_rel = a->rpr_k ? a->max_rel : 0; p.relu_attn = a->relu_attn;
  p.kv_group = a->kv_group > 0 ? a->kv_group : 1;
  p.scale = a->scale; p.inf_value = a->inf_value;
  p.lse = a->lse; p.delta = a->delta; p.d_rpr_k = a->d_rpr_k; p.d_rpr_v = a->d_rpr_v;
  p.drop_rate = a->dropout_seed ? a->dropout_rate : 0.f; p.drop_site = a->dropout_site;
  p.drop_seed = reinterpret_cast<const unsigned long long*>(a->dropout_seed);
  return p;
}

template <int DH>
static size_t smem_bytes(const AttnP& p, int tables) {
  return sizeof(Smem<DH>) + (size_t)tables * (2 * p.max_rel + 1) * (DH + 1) * sizeof(float);
}

template <typename K>
static int set_smem(K kern, size_t bytes) {
  if (bytes > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) {
      set_error("attention: cudaFuncSetAttribute(%zu): %s", bytes, cudaGetErrorString(e));
      return ZB_ECUDA;
    }
  }
  return ZB_OK;
}

// ------------------------------------------------------------------------------------------------ cached decode
// lq = 1 (one new target position per row, func.py:199-216 with the cache): the tiled kernel above would leave seven
// of its eight warps idle and run one CTA per (row, head).  Here a CTA owns one (memory, head): the rows that share
// the memory (the beams of a sentence, kv_group) are its warps, and the 64-key tiles of K and V are staged ONCE in
// shared memory by all threads with 16-byte cp.async (every load of the tile in flight at once).  Per warp: lanes
// split the keys for q.k (padded rows, conflict-free 16-byte reads), the softmax runs on shuffles, and for p.v each
// lane owns two head channels so a value row is one 128-byte shared-memory read.
__device__ __forceinline__ void dec_cp16(void* smem, const void* gmem, bool pred) {
  const uint32_t s = smem_u32(smem);
  const int sz = pred ? 16 : 0;  // src-size 0 -> zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(sz) : "memory");
}

template <int DH>
__global__ void __launch_bounds__(256) attn_decode_kernel(const AttnP p) {
  grid_dep_wait();
  constexpr int KT = 64;        // keys per tile
  constexpr int PK = DH + 8;    // K row pitch: 16-byte reads of 8 consecutive rows fall in distinct banks
  constexpr int CH = DH / 8;    // 16-byte chunks per row
  __shared__ __align__(16) __nv_bfloat16 sK[KT * PK];
  __shared__ __align__(16) __nv_bfloat16 sV[KT * DH];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const int g = blockIdx.x, h = blockIdx.y;
  const int kl = p.key_len ? p.key_len[g] : p.lk;
  const __nv_bfloat16* kbase = p.k + (long long)g * p.bsk + h * DH;
  const __nv_bfloat16* vbase = p.v + (long long)g * p.bsv + h * DH;
  const int i_abs = p.q_offset;
  for (int r0 = 0; r0 < p.kv_group; r0 += nwarp) {
    const int r = r0 + warp;
    const bool row_on = r < p.kv_group;          // warp-uniform
    const int b = g * p.kv_group + (row_on ? r : 0);
    float qr[DH];
    {
      const uint4* qp = reinterpret_cast<const uint4*>(p.q + (long long)b * p.bsq + h * DH);
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        const uint4 u = qp[c];
        const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = __bfloat1622float2(h2[e]);
          qr[8 * c + 2 * e] = f.x * p.scale;
          qr[8 * c + 2 * e + 1] = f.y * p.scale;
        }
      }
    }
    float m = -INFINITY, l = 0.f, o0 = 0.f, o1 = 0.f;
    for (int kt = 0; kt < p.lk; kt += KT) {
      __syncthreads();                           // the previous tile has been consumed by every warp
      const int nvalid = min(KT, p.lk - kt);
      for (int c = threadIdx.x; c < 2 * KT * CH; c += blockDim.x) {
        const int which = c / (KT * CH), cc = c % (KT * CH), row = cc / CH, ch = cc % CH;
        const bool ok = row < nvalid;
        const long long grow = kt + (ok ? row : 0);
        if (which == 0) dec_cp16(sK + row * PK + ch * 8, kbase + grow * p.ldk + ch * 8, ok);
        else dec_cp16(sV + row * DH + ch * 8, vbase + grow * p.ldv + ch * 8, ok);
      }
      asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
      __syncthreads();
      if (!row_on) continue;
      float sc[2], pj[2];
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int jl = half * 32 + lane, j = kt + jl;
        const bool inb = j < p.lk;
        float s = 0.f;
        const uint4* kp = reinterpret_cast<const uint4*>(sK + jl * PK);
#pragma unroll
        for (int c = 0; c < CH; ++c) {
          const uint4 u = kp[c];
          const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 f = __bfloat1622float2(h2[e]);
            s += qr[8 * c + 2 * e] * f.x;
            s += qr[8 * c + 2 * e + 1] * f.y;
          }
        }
        const bool valid = inb && j < kl && (!p.causal || j <= i_abs);
        if (p.relu_attn) {
          pj[half] = valid ? fmaxf(s, 0.f) : 0.f;
          sc[half] = 0.f;
        } else {
          sc[half] = inb ? (valid ? s : s - p.inf_value) : -INFINITY;
        }
      }
      if (!p.relu_attn) {
        const float m_new = fmaxf(m, wmax(fmaxf(sc[0], sc[1])));
        pj[0] = sc[0] == -INFINITY ? 0.f : __expf(sc[0] - m_new);
        pj[1] = sc[1] == -INFINITY ? 0.f : __expf(sc[1] - m_new);
        const float corr = __expf(m - m_new);
        l = l * corr + wsum(pj[0] + pj[1]);
        m = m_new;
        o0 *= corr;
        o1 *= corr;
      }
      // p.v: lane owns channels 2 * lane, 2 * lane + 1 (dh = 64) or channel lane (dh <= 32)
      const int n0 = min(32, nvalid), n1 = nvalid - n0;
#pragma unroll 8
      for (int jj = 0; jj < n0; ++jj) {
        const float pv = __shfl_sync(0xffffffffu, pj[0], jj);
        if (DH == 64) {
          const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(sV + jj * DH + 2 * lane));
          o0 += pv * f.x;
          o1 += pv * f.y;
        } else if (lane < DH) {
          o0 += pv * __bfloat162float(sV[jj * DH + lane]);
        }
      }
#pragma unroll 8
      for (int jj = 0; jj < n1; ++jj) {
        const float pv = __shfl_sync(0xffffffffu, pj[1], jj);
        if (DH == 64) {
          const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(sV + (32 + jj) * DH + 2 * lane));
          o0 += pv * f.x;
          o1 += pv * f.y;
        } else if (lane < DH) {
          o0 += pv * __bfloat162float(sV[(32 + jj) * DH + lane]);
        }
      }
    }
    if (!row_on) continue;
    const float inv = p.relu_attn ? 1.f : 1.f / l;
    __nv_bfloat16* op = p.o + (long long)b * p.bso + h * DH;
    if (DH == 64) {
      *reinterpret_cast<__nv_bfloat162*>(op + 2 * lane) = __floats2bfloat162_rn(o0 * inv, o1 * inv);
    } else if (lane < DH) {
      op[lane] = __float2bfloat16(o0 * inv);
    }
    if (lane == 0 && p.lse) p.lse[(long long)b * p.heads + h] = p.relu_attn ? 0.f : m + __logf(l);
  }
}

static bool decode_path_enabled() {
  static const bool on = [] {
    const char* e = getenv("ZB_DECODE_ATTN");
    return !(e && e[0] == '0');
  }();
  return on;
}

// lq = 1, plain or ReLU attention without relative positions or dropout, 16-byte aligned head slices
bool attention_decode_supported(const zb_attention_args* a) {
  if (!decode_path_enabled()) return false;
  if (a->lq != 1 || a->rpr_k || (a->dropout_seed && a->dropout_rate > 0.f)) return false;
  if (a->dh != 16 && a->dh != 32 && a->dh != 64) return false;
  const int grp = a->kv_group > 0 ? a->kv_group : 1;
  if (a->batch % grp) return false;
  auto al16 = [](const void* ptr) { return (reinterpret_cast<uintptr_t>(ptr) & 15u) == 0; };
  if (!al16(a->q) || !al16(a->k) || !al16(a->v)) return false;
  if ((a->bsq % 8) || (a->bsk % 8) || (a->ldk % 8) || (a->bsv % 8) || (a->ldv % 8)) return false;
  if ((a->bso % 2) || (reinterpret_cast<uintptr_t>(a->o) & 3u)) return false;
  return true;
}

int attention_decode_fwd(const zb_attention_args* a, cudaStream_t st) {
  AttnP p = to_params(a);
  // one warp per row of the group; at least four so the cooperative tile loads have enough threads in flight
  const int warps = p.kv_group < 4 ? 4 : (p.kv_group < 8 ? p.kv_group : 8);
  const dim3 grid(a->batch / p.kv_group, a->heads);
  switch (a->dh) {
    case 16: ZB_LAUNCH(attn_decode_kernel<16>, grid, warps * 32, 0, st, p); break;
    case 32: ZB_LAUNCH(attn_decode_kernel<32>, grid, warps * 32, 0, st, p); break;
    default: ZB_LAUNCH(attn_decode_kernel<64>, grid, warps * 32, 0, st, p); break;
  }
  note_path(ZB_PATH_ATTN_DECODE);
  return check_launch("zb_attention_fwd(decode)");
}

int attention_generic_fwd(const zb_attention_args* a, cudaStream_t st) {
  AttnP p = to_params(a);
  const dim3 grid((a->lq + kAttnWarps - 1) / kAttnWarps, a->heads, a->batch);
  const int tables = a->rpr_k ? 2 : 0;
#define LAUNCH(DH)                                                               \
  do {                                                                           \
    const size_t sb = smem_bytes<DH>(p, tables);                                 \
    int rc = set_smem(attn_fwd_generic<DH>, sb);                                 \
    if (rc) return rc;                                                           \
    ZB_LAUNCH(attn_fwd_generic<DH>, grid, kAttnWarps * 32, sb, st, p);                  \
  } while (0)
  switch (a->dh) {
    case 16: LAUNCH(16); break;
    case 32: LAUNCH(32); break;
    case 64: LAUNCH(64); break;
    default: set_error("zb_attention: head size %d unsupported (16/32/64)", a->dh); return ZB_EUNSUPPORTED;
  }
#undef LAUNCH
  note_path(ZB_PATH_ATTN_GENERIC);
  return check_launch("zb_attention_fwd");
}

int attention_generic_bwd(const zb_attention_args* a, cudaStream_t st) {
  AttnP p = to_params(a);
  const dim3 gq((a->lq + kAttnWarps - 1) / kAttnWarps, a->heads, a->batch);
  const dim3 gk((a->lk + kAttnWarps - 1) / kAttnWarps, a->heads, a->batch);
  const bool rpr = a->rpr_k != nullptr;
#define LAUNCH(DH)                                                               \
  do {                                                                           \
    const size_t s1 = smem_bytes<DH>(p, rpr ? 4 : 0);                            \
    int rc = set_smem(attn_bwd_dq_generic<DH>, s1);                              \
    if (rc) return rc;                                                           \
    ZB_LAUNCH(attn_bwd_dq_generic<DH>, gq, kAttnWarps * 32, s1, st, p);                 \
    rc = check_launch("zb_attention_bwd(dq)");                                   \
    if (rc) return rc;                                                           \
    const size_t s2 = smem_bytes<DH>(p, rpr ? 2 : 0);                            \
    rc = set_smem(attn_bwd_dkv_generic<DH>, s2);                                 \
    if (rc) return rc;                                                           \
    ZB_LAUNCH(attn_bwd_dkv_generic<DH>, gk, kAttnWarps * 32, s2, st, p);                \
  } while (0)
  switch (a->dh) {
    case 16: LAUNCH(16); break;
    case 32: LAUNCH(32); break;
    case 64: LAUNCH(64); break;
    default: set_error("zb_attention: head size %d unsupported (16/32/64)", a->dh); return ZB_EUNSUPPORTED;
  }
#undef LAUNCH
  return check_launch("zb_attention_bwd(dkv)");
}

}  // namespace zb
