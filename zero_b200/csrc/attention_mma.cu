// attention_mma.cu — K2/K3 on the warp-level tensor cores (mma.sync m16n8k16 bf16 -> fp32) for the plain softmax
// case with dh = 64 (func.py:218-256).  Since round 2 the DEFAULT for these problems is attention_tc.cu (tcgen05 /
// TMEM / TMA); this file is what ZB_ATTN_TC=0 selects (the A/B baseline) and what serves the shapes the tcgen05
// kernels decline (unaligned views, causal with a query offset, no workspace for a multi-block backward).  Logits / probabilities live only in registers;
// the backward recomputes them from q, k and the saved log-sum-exp.
//   fwd : CTA = 64 query rows of one (batch, head), 4 warps x 16 rows, 64-key tiles, online softmax
//   bwd : dq kernel (CTA per 64 queries, loops over key tiles) + dk/dv kernel (CTA per 64 keys, loops over
//         query tiles); no atomics, deterministic
// Shared-memory tiles are [64][64] bf16 with a 16-byte-chunk XOR swizzle (chunk ^= row & 7) so ldmatrix is
// bank-conflict free.  Masks come from key lengths / causal index (func.attention_bias, func.py:372-388) with the
// reference's additive -inf_value.
#include <math.h>
#include <stdlib.h>

#include "zb_common.h"
#include "zb_ptx.cuh"

namespace zb {

namespace fa {

constexpr int BQ = 64, BK = 64, DH = 64, NT = 128;

struct Params {
  const __nv_bfloat16 *q, *k, *v, *o, *d_o;
  __nv_bfloat16 *out, *dq, *dk, *dv;
  long long ldq, ldk, ldv, ldo, bsq, bsk, bsv, bso;
  long long lddo, lddq, lddk, lddv, bsdo, bsdq, bsdk, bsdv;
  int heads, lq, lk, causal, q_offset, kv_group;
  const int32_t* key_len;
  float scale, inf_value;
  float *lse, *delta;
  float drop_rate;  // attention dropout (func.py:245); the kernels are instantiated with DROP = true when > 0
  uint32_t drop_site;
  const unsigned long long* drop_seed;
};

// Dropout multiplier (0 or 1 / keep) of attention weight (query row i, key j) of one (batch, head): the same pure
// function of (*seed, site, flat index into [batch, heads, lq, lk]) in every forward / backward kernel.
template <bool DROP>
struct Drop {
  uint64_t seed, base;
  uint32_t site, thr;
  float inv_keep;
  int lk;
  __device__ __forceinline__ Drop(const Params& p, int b, int h) {
    if (DROP) {
      seed = *p.drop_seed;
      base = ((uint64_t)b * p.heads + h) * (uint64_t)p.lq;
      site = p.drop_site;
      thr = dropout_threshold(p.drop_rate);
      inv_keep = 1.f / (1.f - p.drop_rate);
      lk = p.lk;
    }
  }
  __device__ __forceinline__ float mul(int i, int j) const {
    if (!DROP) return 1.f;
    return dropout_mul(seed, site, (base + i) * (uint64_t)lk + j, thr, inv_keep);
  }
};

__device__ __forceinline__ int swz(int r, int c) { return r * 64 + ((((c >> 3) ^ (r & 7)) << 3) | (c & 7)); }

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool pred) {
  const uint32_t s = smem_u32(smem);
  const int sz = pred ? 16 : 0;  // src-size 0 -> zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

// 64 x 64 bf16 tile, rows [r0, r0 + 64) of a [rows, ld] view starting at column 0 of `src`
__device__ __forceinline__ void load_tile(__nv_bfloat16* dst, const __nv_bfloat16* src, long long ld, int rows_valid) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int chunk = threadIdx.x + i * NT;  // 512 chunks of 16 B
    const int r = chunk >> 3, ch = chunk & 7;
    const bool ok = r < rows_valid;
    cp_async16(dst + r * 64 + ((ch ^ (r & 7)) << 3), src + (long long)(ok ? r : 0) * ld + ch * 8, ok);
  }
}

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const __nv_bfloat16* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const __nv_bfloat16* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(p)));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// A fragments (16 rows x 64 k) of the warp's row block from a row-major tile
__device__ __forceinline__ void load_a_frags(uint32_t (&a)[4][4], const __nv_bfloat16* tile, int row0) {
  const int lane = threadIdx.x & 31;
  const int r = row0 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) ldsm_x4(a[ks], tile + swz(r, ks * 16 + (lane >> 4) * 8));
}

// acc[nt] += A(16 x 64) * B^T where B tile is stored [n][k] (k contiguous): acc is 16 x 64 (8 n-tiles)
__device__ __forceinline__ void gemm_nk(float (&acc)[8][4], const uint32_t (&a)[4][4], const __nv_bfloat16* tile) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t b[4];
      const int n = np * 16 + (lane & 7) + (lane >> 4) * 8;
      const int kc = ks * 16 + ((lane >> 3) & 1) * 8;
      ldsm_x4(b, tile + swz(n, kc));
      mma16816(acc[2 * np], a[ks], b[0], b[1]);
      mma16816(acc[2 * np + 1], a[ks], b[2], b[3]);
    }
  }
}
// acc[nt] += P(16 x 64, A fragments) * B where B tile is stored [k][n] (n contiguous): acc is 16 x 64
__device__ __forceinline__ void gemm_kn(float (&acc)[8][4], const uint32_t (&a)[4][4], const __nv_bfloat16* tile) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t b[4];
      const int kr = ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
      const int nc = np * 16 + (lane >> 4) * 8;
      ldsm_x4_t(b, tile + swz(kr, nc));
      mma16816(acc[2 * np], a[ks], b[0], b[1]);
      mma16816(acc[2 * np + 1], a[ks], b[2], b[3]);
    }
  }
}
// C-layout fp32 (16 x 64) -> A-layout bf16 fragments for the next GEMM's k dimension
__device__ __forceinline__ void c_to_a(uint32_t (&a)[4][4], const float (&c)[8][4]) {
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    a[ks][0] = pack_bf16x2(c[2 * ks][0], c[2 * ks][1]);
    a[ks][1] = pack_bf16x2(c[2 * ks][2], c[2 * ks][3]);
    a[ks][2] = pack_bf16x2(c[2 * ks + 1][0], c[2 * ks + 1][1]);
    a[ks][3] = pack_bf16x2(c[2 * ks + 1][2], c[2 * ks + 1][3]);
  }
}
__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}
__device__ __forceinline__ void store_c_bf16(__nv_bfloat16* base, long long ld, int row0, int rows_valid,
                                             const float (&c)[8][4], float mul0, float mul1) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    const int col = nt * 8 + 2 * t;
    if (row0 + g < rows_valid)
      *reinterpret_cast<uint32_t*>(base + (long long)(row0 + g) * ld + col) = pack_bf16x2(c[nt][0] * mul0, c[nt][1] * mul0);
    if (row0 + g + 8 < rows_valid)
      *reinterpret_cast<uint32_t*>(base + (long long)(row0 + g + 8) * ld + col) =
          pack_bf16x2(c[nt][2] * mul1, c[nt][3] * mul1);
  }
}

// ------------------------------------------------------------------------------------------------ forward
template <bool DROP>
__global__ void __launch_bounds__(NT) fwd_kernel(const Params p) {
  grid_dep_wait();
  __shared__ __align__(128) __nv_bfloat16 sQ[64 * 64], sK[64 * 64], sV[64 * 64];
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * BQ;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const Drop<DROP> drop(p, b, h);
  const int kb = b / p.kv_group;
  const int kl = p.key_len ? p.key_len[kb] : p.lk;
  const __nv_bfloat16* qb = p.q + (long long)b * p.bsq + (long long)q0 * p.ldq + h * DH;
  load_tile(sQ, qb, p.ldq, p.lq - q0);
  cp_async_wait_all();
  __syncthreads();
  uint32_t qa[4][4];
  load_a_frags(qa, sQ, warp * 16);
  float m[2] = {-INFINITY, -INFINITY}, l[2] = {0.f, 0.f};
  float o[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
  const int row_abs0 = q0 + warp * 16 + g + p.q_offset;
  int k_end = p.lk;
  if (p.causal) k_end = min(p.lk, q0 + BQ + p.q_offset);  // keys beyond the last query row contribute exp(-inf) = 0
  for (int kt = 0; kt < k_end; kt += BK) {
    __syncthreads();
    load_tile(sK, p.k + (long long)kb * p.bsk + (long long)kt * p.ldk + h * DH, p.ldk, p.lk - kt);
    load_tile(sV, p.v + (long long)kb * p.bsv + (long long)kt * p.ldv + h * DH, p.ldv, p.lk - kt);
    cp_async_wait_all();
    __syncthreads();
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
    gemm_nk(s, qa, sK);
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int col = kt + nt * 8 + 2 * t + (j & 1);
        const int ra = row_abs0 + (j >> 1) * 8;
        const bool inb = col < p.lk;
        const bool valid = inb && col < kl && (!p.causal || col <= ra);
        float v = s[nt][j] * p.scale;
        v = inb ? (valid ? v : v - p.inf_value) : -INFINITY;
        s[nt][j] = v;
        mx[j >> 1] = fmaxf(mx[j >> 1], v);
      }
    float corr[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const float mn = fmaxf(m[r], quad_max(mx[r]));
      corr[r] = __expf(m[r] - mn);
      m[r] = mn;
      l[r] *= corr[r];
    }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float pv = __expf(s[nt][j] - m[j >> 1]);
        l[j >> 1] += pv;  // the normaliser sums the undropped weights
        s[nt][j] = pv * drop.mul(q0 + warp * 16 + g + (j >> 1) * 8, kt + nt * 8 + 2 * t + (j & 1));
        o[nt][j] *= corr[j >> 1];
      }
    uint32_t pa[4][4];
    c_to_a(pa, s);
    gemm_kn(o, pa, sV);
  }
  const float l0 = quad_sum(l[0]), l1 = quad_sum(l[1]);
  __nv_bfloat16* ob = p.out + (long long)b * p.bso + (long long)q0 * p.ldo + h * DH;
  store_c_bf16(ob, p.ldo, warp * 16, p.lq - q0, o, 1.f / l0, 1.f / l1);
  if (t == 0 && p.lse) {
    const int r0 = q0 + warp * 16 + g;
    float* lp = p.lse + ((long long)b * p.heads + h) * p.lq;
    if (r0 < p.lq) lp[r0] = m[0] + __logf(l0);
    if (r0 + 8 < p.lq) lp[r0 + 8] = m[1] + __logf(l1);
  }
}

// ------------------------------------------------------------------------------------------------ backward: dq
template <bool DROP>
__global__ void __launch_bounds__(NT) bwd_dq_kernel(const Params p) {
  grid_dep_wait();
  __shared__ __align__(128) __nv_bfloat16 sQ[64 * 64], sK[64 * 64], sV[64 * 64];
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * BQ;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const Drop<DROP> drop(p, b, h);
  const int kl = p.key_len ? p.key_len[b] : p.lk;
  const int rows = p.lq - q0;
  uint32_t qa[4][4], doa[4][4];
  // dO and O tiles -> delta = rowsum(dO * O); dO fragments stay in registers
  load_tile(sK, p.d_o + (long long)b * p.bsdo + (long long)q0 * p.lddo + h * DH, p.lddo, rows);
  load_tile(sV, p.o + (long long)b * p.bso + (long long)q0 * p.ldo + h * DH, p.ldo, rows);
  load_tile(sQ, p.q + (long long)b * p.bsq + (long long)q0 * p.ldq + h * DH, p.ldq, rows);
  cp_async_wait_all();
  __syncthreads();
  load_a_frags(qa, sQ, warp * 16);
  load_a_frags(doa, sK, warp * 16);
  float delta[2] = {0.f, 0.f};
  {
    uint32_t oa[4][4];
    load_a_frags(oa, sV, warp * 16);
#pragma unroll
    for (int ks = 0; ks < 4; ++ks)
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float2 x = unpack_bf16x2(doa[ks][r]), y = unpack_bf16x2(oa[ks][r]);
        delta[r & 1] += x.x * y.x + x.y * y.y;  // regs 0,2 -> row g ; regs 1,3 -> row g + 8
      }
    delta[0] = quad_sum(delta[0]);
    delta[1] = quad_sum(delta[1]);
  }
  const int r0 = q0 + warp * 16 + g;
  const long long lbase = ((long long)b * p.heads + h) * p.lq;
  float lse[2] = {r0 < p.lq ? p.lse[lbase + r0] : 0.f, r0 + 8 < p.lq ? p.lse[lbase + r0 + 8] : 0.f};
  if (t == 0) {
    if (r0 < p.lq) p.delta[lbase + r0] = delta[0];
    if (r0 + 8 < p.lq) p.delta[lbase + r0 + 8] = delta[1];
  }
  float dq[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) dq[i][j] = 0.f;
  const int row_abs0 = r0 + p.q_offset;
  int k_end = p.lk;
  if (p.causal) k_end = min(p.lk, q0 + BQ + p.q_offset);
  for (int kt = 0; kt < k_end; kt += BK) {
    __syncthreads();
    load_tile(sK, p.k + (long long)b * p.bsk + (long long)kt * p.ldk + h * DH, p.ldk, p.lk - kt);
    load_tile(sV, p.v + (long long)b * p.bsv + (long long)kt * p.ldv + h * DH, p.ldv, p.lk - kt);
    cp_async_wait_all();
    __syncthreads();
    float s[8][4], dp[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[i][j] = dp[i][j] = 0.f;
    gemm_nk(s, qa, sK);
    gemm_nk(dp, doa, sV);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int col = kt + nt * 8 + 2 * t + (j & 1);
        const int ra = row_abs0 + (j >> 1) * 8;
        const bool inb = col < p.lk;
        const bool valid = inb && col < kl && (!p.causal || col <= ra);
        float v = s[nt][j] * p.scale;
        v = valid ? v : v - p.inf_value;
        const float pv = inb ? __expf(v - lse[j >> 1]) : 0.f;
        s[nt][j] = pv * (dp[nt][j] * drop.mul(r0 + (j >> 1) * 8, col) - delta[j >> 1]);
      }
    uint32_t dsa[4][4];
    c_to_a(dsa, s);
    gemm_kn(dq, dsa, sK);
  }
  store_c_bf16(p.dq + (long long)b * p.bsdq + (long long)q0 * p.lddq + h * DH, p.lddq, warp * 16, rows, dq, p.scale,
               p.scale);
}

// ------------------------------------------------------------------------------------------------ backward: dk, dv
template <bool DROP>
__global__ void __launch_bounds__(NT) bwd_dkv_kernel(const Params p) {
  grid_dep_wait();
  __shared__ __align__(128) __nv_bfloat16 sQ[64 * 64], sDO[64 * 64];
  __shared__ float sLse[64], sDelta[64];
  const int b = blockIdx.z, h = blockIdx.y, k0 = blockIdx.x * BK;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const Drop<DROP> drop(p, b, h);
  const int kl = p.key_len ? p.key_len[b] : p.lk;
  const int krows = p.lk - k0;
  uint32_t ka[4][4], va[4][4];
  load_tile(sQ, p.k + (long long)b * p.bsk + (long long)k0 * p.ldk + h * DH, p.ldk, krows);
  load_tile(sDO, p.v + (long long)b * p.bsv + (long long)k0 * p.ldv + h * DH, p.ldv, krows);
  cp_async_wait_all();
  __syncthreads();
  load_a_frags(ka, sQ, warp * 16);
  load_a_frags(va, sDO, warp * 16);
  float dk[8][4], dv[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) dk[i][j] = dv[i][j] = 0.f;
  const int key0 = k0 + warp * 16 + g;
  const long long lbase = ((long long)b * p.heads + h) * p.lq;
  int q_begin = 0;
  if (p.causal) q_begin = max(0, (k0 - p.q_offset) / BQ * BQ);  // queries before the first key of the block see none of it
  for (int qt = q_begin; qt < p.lq; qt += BQ) {
    __syncthreads();
    load_tile(sQ, p.q + (long long)b * p.bsq + (long long)qt * p.ldq + h * DH, p.ldq, p.lq - qt);
    load_tile(sDO, p.d_o + (long long)b * p.bsdo + (long long)qt * p.lddo + h * DH, p.lddo, p.lq - qt);
    if (threadIdx.x < 64) {
      const int i = qt + threadIdx.x;
      sLse[threadIdx.x] = i < p.lq ? p.lse[lbase + i] : 0.f;
      sDelta[threadIdx.x] = i < p.lq ? p.delta[lbase + i] : 0.f;
    }
    cp_async_wait_all();
    __syncthreads();
    float st[8][4], dpt[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) st[i][j] = dpt[i][j] = 0.f;
    gemm_nk(st, ka, sQ);     // S^T[key, query] = K Q^T
    gemm_nk(dpt, va, sDO);   // dP^T[key, query] = V dO^T
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int qi = nt * 8 + 2 * t + (j & 1);
        const int key = key0 + (j >> 1) * 8;
        const bool inb = qt + qi < p.lq && key < p.lk;
        const bool valid = key < kl && (!p.causal || key <= qt + qi + p.q_offset);
        float v = st[nt][j] * p.scale;
        v = valid ? v : v - p.inf_value;
        const float pv = inb ? __expf(v - sLse[qi]) : 0.f;
        const float dm = drop.mul(qt + qi, key);
        st[nt][j] = pv * dm;
        dpt[nt][j] = pv * (dpt[nt][j] * dm - sDelta[qi]);
      }
    uint32_t pa[4][4], dsa[4][4];
    c_to_a(pa, st);
    c_to_a(dsa, dpt);
    gemm_kn(dv, pa, sDO);   // dV += P^T dO
    gemm_kn(dk, dsa, sQ);   // dK += dS^T Q
  }
  store_c_bf16(p.dk + (long long)b * p.bsdk + (long long)k0 * p.lddk + h * DH, p.lddk, warp * 16, krows, dk, p.scale,
               p.scale);
  store_c_bf16(p.dv + (long long)b * p.bsdv + (long long)k0 * p.lddv + h * DH, p.lddv, warp * 16, krows, dv, 1.f, 1.f);
}

// ================================================================================================ single-tile path
// lq <= 64 and lk <= 64 (every attention of the reference's 64-token training batches): one (batch, head) problem is
// one 64 x 64 tile, so a CTA is handed a stream of (batch, head) items and double-buffers them through shared memory
// with cp.async — the loads of item i+1 overlap the math of item i and the kernel runs at HBM speed instead of
// load -> wait -> compute per CTA.  The backward is fused: dQ, dK and dV of an item come from ONE load of
// Q, K, V, dO, O (the two-kernel general path loads them twice).
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

struct Tile64Fwd {
  __nv_bfloat16 q[64 * 64], k[64 * 64], v[64 * 64];
};
struct Tile64Bwd {
  __nv_bfloat16 q[64 * 64], k[64 * 64], v[64 * 64], d_o[64 * 64], o[64 * 64];
  float lse[64];
};

// These kernels are latency-bound per warp (ncu: 2 warps / scheduler, 26 % issue slots used), so they are sized for
// residency rather than prefetch: one tile set per CTA (24 / 41 KB), <= 128 registers, 4 CTAs per SM; at batch 64 x 8
// heads every (batch, head) item of the step is resident at once.
template <bool DROP>
__global__ void __launch_bounds__(NT, 4) fwd64_kernel(const Params p, const int items) {
  grid_dep_wait();
  extern __shared__ __align__(128) uint8_t smem64[];
  Tile64Fwd& T = *reinterpret_cast<Tile64Fwd*>(smem64);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  for (int item = blockIdx.x; item < items; item += gridDim.x) {
    const int b = item / p.heads, h = item % p.heads;
    load_tile(T.q, p.q + (long long)b * p.bsq + h * DH, p.ldq, p.lq);
    load_tile(T.k, p.k + (long long)b * p.bsk + h * DH, p.ldk, p.lk);
    load_tile(T.v, p.v + (long long)b * p.bsv + h * DH, p.ldv, p.lk);
    cp_async_commit();
    const Drop<DROP> drop(p, b, h);
    cp_async_wait<0>();
    __syncthreads();
    const int kl = p.key_len ? p.key_len[b] : p.lk;
    uint32_t qa[4][4];
    load_a_frags(qa, T.q, warp * 16);
    float s[8][4], o[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[i][j] = o[i][j] = 0.f;
    gemm_nk(s, qa, T.k);
    const int row_abs0 = warp * 16 + g + p.q_offset;
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int col = nt * 8 + 2 * t + (j & 1);
        const int ra = row_abs0 + (j >> 1) * 8;
        const bool inb = col < p.lk;
        const bool valid = inb && col < kl && (!p.causal || col <= ra);
        float v = s[nt][j] * p.scale;
        v = inb ? (valid ? v : v - p.inf_value) : -INFINITY;
        s[nt][j] = v;
        mx[j >> 1] = fmaxf(mx[j >> 1], v);
      }
    const float m0 = quad_max(mx[0]), m1 = quad_max(mx[1]);
    float l0 = 0.f, l1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float pv = __expf(s[nt][j] - ((j >> 1) ? m1 : m0));
        if (j >> 1) l1 += pv; else l0 += pv;
        s[nt][j] = pv * drop.mul(warp * 16 + g + (j >> 1) * 8, nt * 8 + 2 * t + (j & 1));
      }
    uint32_t pa[4][4];
    c_to_a(pa, s);
    gemm_kn(o, pa, T.v);
    l0 = quad_sum(l0);
    l1 = quad_sum(l1);
    store_c_bf16(p.out + (long long)b * p.bso + h * DH, p.ldo, warp * 16, p.lq, o, 1.f / l0, 1.f / l1);
    if (t == 0 && p.lse) {
      const int r0 = warp * 16 + g;
      float* lp = p.lse + ((long long)b * p.heads + h) * p.lq;
      if (r0 < p.lq) lp[r0] = m0 + __logf(l0);
      if (r0 + 8 < p.lq) lp[r0 + 8] = m1 + __logf(l1);
    }
    __syncthreads();  // everyone is done with the tiles before the next item's loads overwrite them
  }
}

template <bool DROP>
__global__ void __launch_bounds__(NT, 4) bwd64_kernel(const Params p, const int items) {
  grid_dep_wait();
  extern __shared__ __align__(128) uint8_t smem64[];
  Tile64Bwd& T = *reinterpret_cast<Tile64Bwd*>(smem64);
  __shared__ float sDelta[64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  for (int item = blockIdx.x; item < items; item += gridDim.x) {
    const int b = item / p.heads, h = item % p.heads;
    load_tile(T.q, p.q + (long long)b * p.bsq + h * DH, p.ldq, p.lq);
    load_tile(T.k, p.k + (long long)b * p.bsk + h * DH, p.ldk, p.lk);
    load_tile(T.v, p.v + (long long)b * p.bsv + h * DH, p.ldv, p.lk);
    load_tile(T.d_o, p.d_o + (long long)b * p.bsdo + h * DH, p.lddo, p.lq);
    load_tile(T.o, p.o + (long long)b * p.bso + h * DH, p.ldo, p.lq);
    if (threadIdx.x < 16) {  // 64 fp32 log-sum-exps (lq may be < 64: rows beyond lq read as 0 through the predicate)
      const int r = threadIdx.x * 4;
      const float* src = p.lse + ((long long)b * p.heads + h) * p.lq + r;
      if (r + 3 < p.lq && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
        cp_async16(&T.lse[r], src, true);
      } else {
        for (int e = 0; e < 4; ++e) T.lse[r + e] = (r + e < p.lq) ? src[e] : 0.f;
      }
    }
    cp_async_commit();
    const Drop<DROP> drop(p, b, h);
    cp_async_wait<0>();
    __syncthreads();
    const int kl = p.key_len ? p.key_len[b] : p.lk;
    // ---- delta = rowsum(dO * O) for this warp's 16 query rows -> shared (every warp needs all 64 below)
    uint32_t doa[4][4];
    load_a_frags(doa, T.d_o, warp * 16);
    float delta[2] = {0.f, 0.f};
    {
      uint32_t oa[4][4];
      load_a_frags(oa, T.o, warp * 16);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const float2 x = unpack_bf16x2(doa[ks][r]), y = unpack_bf16x2(oa[ks][r]);
          delta[r & 1] += x.x * y.x + x.y * y.y;
        }
      delta[0] = quad_sum(delta[0]);
      delta[1] = quad_sum(delta[1]);
      if (t == 0) {
        sDelta[warp * 16 + g] = delta[0];
        sDelta[warp * 16 + g + 8] = delta[1];
      }
    }
    __syncthreads();
    // ---- dQ for query rows [16w, 16w + 16)
    {
      uint32_t qa[4][4];
      load_a_frags(qa, T.q, warp * 16);
      float s[8][4], dp[8][4], dq[8][4];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s[i][j] = dp[i][j] = dq[i][j] = 0.f;
      gemm_nk(s, qa, T.k);
      gemm_nk(dp, doa, T.v);
      const int r0 = warp * 16 + g;
      const float lse0 = T.lse[r0], lse1 = T.lse[r0 + 8];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int col = nt * 8 + 2 * t + (j & 1);
          const int ra = r0 + p.q_offset + (j >> 1) * 8;
          const bool inb = col < p.lk;
          const bool valid = inb && col < kl && (!p.causal || col <= ra);
          float v = s[nt][j] * p.scale;
          v = valid ? v : v - p.inf_value;
          const float pv = inb ? __expf(v - ((j >> 1) ? lse1 : lse0)) : 0.f;
          s[nt][j] = pv * (dp[nt][j] * drop.mul(r0 + (j >> 1) * 8, col) - delta[j >> 1]);
        }
      uint32_t dsa[4][4];
      c_to_a(dsa, s);
      gemm_kn(dq, dsa, T.k);
      store_c_bf16(p.dq + (long long)b * p.bsdq + h * DH, p.lddq, warp * 16, p.lq, dq, p.scale, p.scale);
    }
    // ---- dK, dV for key rows [16w, 16w + 16)
    {
      uint32_t ka[4][4], va[4][4];
      load_a_frags(ka, T.k, warp * 16);
      load_a_frags(va, T.v, warp * 16);
      float st[8][4], dpt[8][4];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) st[i][j] = dpt[i][j] = 0.f;
      gemm_nk(st, ka, T.q);
      gemm_nk(dpt, va, T.d_o);
      const int key0 = warp * 16 + g;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int qi = nt * 8 + 2 * t + (j & 1);
          const int key = key0 + (j >> 1) * 8;
          const bool inb = qi < p.lq && key < p.lk;
          const bool valid = key < kl && (!p.causal || key <= qi + p.q_offset);
          float v = st[nt][j] * p.scale;
          v = valid ? v : v - p.inf_value;
          const float pv = inb ? __expf(v - T.lse[qi]) : 0.f;
          const float dm = drop.mul(qi, key);
          st[nt][j] = pv * dm;
          dpt[nt][j] = pv * (dpt[nt][j] * dm - sDelta[qi]);
        }
      uint32_t pa[4][4], dsa[4][4];
      c_to_a(pa, st);
      c_to_a(dsa, dpt);
      float dk[8][4], dv[8][4];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dk[i][j] = dv[i][j] = 0.f;
      gemm_kn(dv, pa, T.d_o);
      gemm_kn(dk, dsa, T.q);
      store_c_bf16(p.dk + (long long)b * p.bsdk + h * DH, p.lddk, warp * 16, p.lk, dk, p.scale, p.scale);
      store_c_bf16(p.dv + (long long)b * p.bsdv + h * DH, p.lddv, warp * 16, p.lk, dv, 1.f, 1.f);
    }
    __syncthreads();  // the tiles and sDelta are free again
  }
}

static Params to_params(const zb_attention_args* a) {
  Params p;
  p.q = (const __nv_bfloat16*)a->q; p.k = (const __nv_bfloat16*)a->k; p.v = (const __nv_bfloat16*)a->v;
  p.o = (const __nv_bfloat16*)a->o; p.out = (__nv_bfloat16*)a->o; p.d_o = (const __nv_bfloat16*)a->d_o;
  p.dq = (__nv_bfloat16*)a->dq; p.dk = (__nv_bfloat16*)a->dk; p.dv = (__nv_bfloat16*)a->dv;
  p.ldq = a->ldq; p.ldk = a->ldk; p.ldv = a->ldv; p.ldo = a->ldo;
  p.bsq = a->bsq; p.bsk = a->bsk; p.bsv = a->bsv; p.bso = a->bso;
  p.lddo = a->lddo; p.lddq = a->lddq; p.lddk = a->lddk; p.lddv = a->lddv;
  p.bsdo = a->bsdo; p.bsdq = a->bsdq; p.bsdk = a->bsdk; p.bsdv = a->bsdv;
  p.heads = a->heads; p.lq = a->lq; p.lk = a->lk; p.causal = a->causal; p.q_offset = a->q_offset;
  p.kv_group = a->kv_group > 0 ? a->kv_group : 1;
  p.key_len = a->key_len; p.scale = a->scale; p.inf_value = a->inf_value; p.lse = a->lse; p.delta = a->delta;
  p.drop_rate = (a->dropout_seed && a->dropout_rate > 0.f) ? a->dropout_rate : 0.f;
  p.drop_site = a->dropout_site;
  p.drop_seed = reinterpret_cast<const unsigned long long*>(a->dropout_seed);
  return p;
}

}  // namespace fa

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

bool attention_mma_supported(const zb_attention_args* a, bool bwd) {
  if (a->dh != 64 || a->rpr_k || a->relu_attn) return false;
  if (a->lq < 16 && !bwd) return false;  // decode steps (lq = 1) are better served by the warp-per-row kernel
  if (a->ldq % 8 || a->ldk % 8 || a->ldv % 8 || a->ldo % 8 || a->bsq % 8 || a->bsk % 8 || a->bsv % 8 || a->bso % 8)
    return false;
  if (!aligned16(a->q) || !aligned16(a->k) || !aligned16(a->v) || !aligned16(a->o)) return false;
  if (bwd) {
    if (a->lddo % 8 || a->lddq % 8 || a->lddk % 8 || a->lddv % 8 || a->bsdo % 8 || a->bsdq % 8 || a->bsdk % 8 ||
        a->bsdv % 8)
      return false;
    if (!aligned16(a->d_o) || !aligned16(a->dq) || !aligned16(a->dk) || !aligned16(a->dv)) return false;
    if (a->lq < 16) return false;
  }
  return true;
}

static bool tile64_ok(const zb_attention_args* a) {
  static const bool off = getenv("ZB_NO_TILE64") != nullptr;
  return !off && a->lq <= 64 && a->lk <= 64 && a->kv_group <= 1;
}

int attention_mma_fwd(const zb_attention_args* a, cudaStream_t st) {
  const fa::Params p = fa::to_params(a);
  const bool drop = p.drop_rate > 0.f;
  if (tile64_ok(a)) {
    const int items = a->batch * a->heads;
    const int smem = (int)sizeof(fa::Tile64Fwd);
    static bool attr = false;
    if (!attr) {
      cudaFuncSetAttribute(fa::fwd64_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      cudaFuncSetAttribute(fa::fwd64_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      attr = true;
    }
    const int grid = items < 8 * num_sms() ? items : 8 * num_sms();
    if (drop) ZB_LAUNCH(fa::fwd64_kernel<true>, grid, fa::NT, smem, st, p, items);
    else ZB_LAUNCH(fa::fwd64_kernel<false>, grid, fa::NT, smem, st, p, items);
    note_path(ZB_PATH_ATTN_MMA);
    return check_launch("zb_attention_fwd(tile64)");
  }
  const dim3 grid((a->lq + fa::BQ - 1) / fa::BQ, a->heads, a->batch);
  if (drop) ZB_LAUNCH(fa::fwd_kernel<true>, grid, fa::NT, 0, st, p);
  else ZB_LAUNCH(fa::fwd_kernel<false>, grid, fa::NT, 0, st, p);
  note_path(ZB_PATH_ATTN_MMA);
  return check_launch("zb_attention_fwd(mma)");
}

int attention_mma_bwd(const zb_attention_args* a, cudaStream_t st) {
  const fa::Params p = fa::to_params(a);
  const bool drop = p.drop_rate > 0.f;
  if (tile64_ok(a)) {
    const int items = a->batch * a->heads;
    const int smem = (int)sizeof(fa::Tile64Bwd);
    static bool attr = false;
    if (!attr) {
      cudaFuncSetAttribute(fa::bwd64_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      cudaFuncSetAttribute(fa::bwd64_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      attr = true;
    }
    const int grid = items < 4 * num_sms() ? items : 4 * num_sms();
    if (drop) ZB_LAUNCH(fa::bwd64_kernel<true>, grid, fa::NT, smem, st, p, items);
    else ZB_LAUNCH(fa::bwd64_kernel<false>, grid, fa::NT, smem, st, p, items);
    return check_launch("zb_attention_bwd(tile64)");
  }
  const dim3 gq((a->lq + fa::BQ - 1) / fa::BQ, a->heads, a->batch);
  const dim3 gk((a->lk + fa::BK - 1) / fa::BK, a->heads, a->batch);
  if (drop) ZB_LAUNCH(fa::bwd_dq_kernel<true>, gq, fa::NT, 0, st, p);
  else ZB_LAUNCH(fa::bwd_dq_kernel<false>, gq, fa::NT, 0, st, p);
  int rc = check_launch("zb_attention_bwd(mma dq)");
  if (rc) return rc;
  if (drop) ZB_LAUNCH(fa::bwd_dkv_kernel<true>, gk, fa::NT, 0, st, p);
  else ZB_LAUNCH(fa::bwd_dkv_kernel<false>, gk, fa::NT, 0, st, p);
  return check_launch("zb_attention_bwd(mma dkv)");
}

}  // namespace zb
