// addln.cu — K4: fused residual-add + LayerNorm, forward and backward (HBM-bound).
// Replaces func.residual_fn + func.layer_norm (func.py:321-324, 289-303): s = x + y,
// out = scale * (s - mean) * rsqrt(var + eps) + offset, biased variance, eps inside the rsqrt.
// One warp per row, 16-byte vector loads, the row lives in registers (cols <= 2048, cols % 8 == 0).
// Algorithmic bytes / row (bf16): fwd 3 * cols * 2 (read x, y; write out); bwd 5 * cols * 2.
#include "zb_common.h"
#include "zb_ptx.cuh"
#include <stdlib.h>

namespace zb {

constexpr int kLnWarps = 8;
constexpr int kLnMaxVec = 8;  // 8 vectors of 8 bf16 per lane -> cols <= 2048

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ void load8(const __nv_bfloat16* p, float (&f)[8]) {
  const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float2 t = unpack_bf16x2(w[e]);
    f[2 * e] = t.x;
    f[2 * e + 1] = t.y;
  }
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float (&f)[8]) {
  uint4 u;
  u.x = pack_bf16x2(f[0], f[1]);
  u.y = pack_bf16x2(f[2], f[3]);
  u.z = pack_bf16x2(f[4], f[5]);
  u.w = pack_bf16x2(f[6], f[7]);
  *reinterpret_cast<uint4*>(p) = u;
}

// Decode-step variant: the branch output arrives as the fp32 accumulator of a split-K projection (+ the projection's
// bias, which a split-K GEMM cannot add): s = x + y32 + ybias.  The accumulator is cleared on the way (zeros written
// back), so the next split-K projection of the step finds a clean buffer without a memset in between.
template <int NV>
__global__ void __launch_bounds__(kLnWarps * 32, NV <= 2 ? 6 : 2)
add_ln_fwd_acc_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ y32, const float* __restrict__ ybias,
                      __nv_bfloat16* __restrict__ out, float* __restrict__ mean, float* __restrict__ rstd,
                      const float* __restrict__ scale, const float* __restrict__ offset, long long rows, int cols,
                      float eps) {
  grid_dep_wait();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nvec = cols >> 3;
  for (long long row = (long long)blockIdx.x * kLnWarps + warp; row < rows; row += (long long)gridDim.x * kLnWarps) {
    float s[NV][8];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = lane + 32 * i;
      if (v < nvec) {
        load8(x + row * cols + v * 8, s[i]);
        float4* yp = reinterpret_cast<float4*>(y32 + row * cols + v * 8);
        const float4 y0 = yp[0], y1 = yp[1];
        yp[0] = make_float4(0.f, 0.f, 0.f, 0.f);
        yp[1] = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 b0 = ybias ? __ldg(reinterpret_cast<const float4*>(ybias + v * 8)) : make_float4(0.f, 0.f, 0.f, 0.f),
                     b1 = ybias ? __ldg(reinterpret_cast<const float4*>(ybias + v * 8) + 1)
                                : make_float4(0.f, 0.f, 0.f, 0.f);
        // the projection's output is bf16 in the unsplit path (GEMM epilogue): round the same way, so both paths agree
        const float t[8] = {y0.x + b0.x, y0.y + b0.y, y0.z + b0.z, y0.w + b0.w,
                            y1.x + b1.x, y1.y + b1.y, y1.z + b1.z, y1.w + b1.w};
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          s[i][e] += __bfloat162float(__float2bfloat16(t[e]));
          sum += s[i][e];
        }
      }
    }
    const float mu = warp_sum(sum) / cols;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (lane + 32 * i < nvec) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float d = s[i][e] - mu;
          sq += d * d;
        }
      }
    }
    const float rs = rsqrtf(warp_sum(sq) / cols + eps);
    if (lane == 0) {
      if (mean) mean[row] = mu;
      if (rstd) rstd[row] = rs;
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = lane + 32 * i;
      if (v < nvec) {
        float o[8];
        const float4 sc0 = __ldg(reinterpret_cast<const float4*>(scale + v * 8)),
                     sc1 = __ldg(reinterpret_cast<const float4*>(scale + v * 8) + 1),
                     of0 = __ldg(reinterpret_cast<const float4*>(offset + v * 8)),
                     of1 = __ldg(reinterpret_cast<const float4*>(offset + v * 8) + 1);
        const float sc[8] = {sc0.x, sc0.y, sc0.z, sc0.w, sc1.x, sc1.y, sc1.z, sc1.w};
        const float of[8] = {of0.x, of0.y, of0.z, of0.w, of1.x, of1.y, of1.z, of1.w};
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = sc[e] * (s[i][e] - mu) * rs + of[e];
        store8(out + row * cols + v * 8, o);
      }
    }
  }
}

template <int NV>
__global__ void __launch_bounds__(kLnWarps * 32, NV <= 2 ? 6 : 2)
add_ln_fwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ y,
                  __nv_bfloat16* __restrict__ out, float* __restrict__ mean, float* __restrict__ rstd,
                  const float* __restrict__ scale, const float* __restrict__ offset, long long rows, int cols,
                  float eps) {
  grid_dep_wait();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nvec = cols >> 3;
  for (long long row = (long long)blockIdx.x * kLnWarps + warp; row < rows; row += (long long)gridDim.x * kLnWarps) {
    float s[NV][8];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = lane + 32 * i;
      if (v < nvec) {
        load8(x + row * cols + v * 8, s[i]);
        if (y) {
          float t[8];
          load8(y + row * cols + v * 8, t);
#pragma unroll
          for (int e = 0; e < 8; ++e) s[i][e] += t[e];
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) sum += s[i][e];
      }
    }
    const float mu = warp_sum(sum) / cols;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (lane + 32 * i < nvec) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float d = s[i][e] - mu;
          sq += d * d;
        }
      }
    }
    const float rs = rsqrtf(warp_sum(sq) / cols + eps);
    if (lane == 0) {
      if (mean) mean[row] = mu;
      if (rstd) rstd[row] = rs;
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = lane + 32 * i;
      if (v < nvec) {
        float o[8];
        const float4 sc0 = __ldg(reinterpret_cast<const float4*>(scale + v * 8)),
                     sc1 = __ldg(reinterpret_cast<const float4*>(scale + v * 8) + 1),
                     of0 = __ldg(reinterpret_cast<const float4*>(offset + v * 8)),
                     of1 = __ldg(reinterpret_cast<const float4*>(offset + v * 8) + 1);
        const float sc[8] = {sc0.x, sc0.y, sc0.z, sc0.w, sc1.x, sc1.y, sc1.z, sc1.w};
        const float of[8] = {of0.x, of0.y, of0.z, of0.w, of1.x, of1.y, of1.z, of1.w};
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = sc[e] * (s[i][e] - mu) * rs + of[e];
        store8(out + row * cols + v * 8, o);
      }
    }
  }
}

// ds = rstd * (g - mean(g) - s_hat * mean(g * s_hat)), g = d_out * scale;  dscale += sum_rows d_out * s_hat;
// doffset += sum_rows d_out.  d_out may arrive as two addends (gradient through the next sublayer + the
// gradient through the residual skip), which fuses the "x + y" gradient fan-in of func.residual_fn.
template <int NV>
__global__ void __launch_bounds__(kLnWarps * 32)
add_ln_bwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ y,
                  const __nv_bfloat16* __restrict__ d_out, const __nv_bfloat16* __restrict__ d_out2,
                  const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ scale,
                  __nv_bfloat16* __restrict__ ds, float* __restrict__ dscale, float* __restrict__ doffset,
                  float* __restrict__ dbias, long long rows, int cols) {
  grid_dep_wait();
  extern __shared__ float red[];  // [kLnWarps][cols] x 2
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nvec = cols >> 3;
  float acc_s[NV][8], acc_o[NV][8], acc_b[NV][8];
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int e = 0; e < 8; ++e) acc_s[i][e] = acc_o[i][e] = acc_b[i][e] = 0.f;

  for (long long row = (long long)blockIdx.x * kLnWarps + warp; row < rows; row += (long long)gridDim.x * kLnWarps) {
    const float mu = mean[row], rs = rstd[row];
    float sh[NV][8], g[NV][8];
    float sg = 0.f, sgs = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = lane + 32 * i;
      if (v < nvec) {
        load8(x + row * cols + v * 8, sh[i]);
        if (y) {
          float t[8];
          load8(y + row * cols + v * 8, t);
#pragma unroll
          for (int e = 0; e < 8; ++e) sh[i][e] += t[e];
        }
        float d[8];
        load8(d_out + row * cols + v * 8, d);
        if (d_out2) {
          float t[8];
          load8(d_out2 + row * cols + v * 8, t);
#pragma unroll
          for (int e = 0; e < 8; ++e) d[e] += t[e];
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          sh[i][e] = (sh[i][e] - mu) * rs;
          acc_s[i][e] += d[e] * sh[i][e];
          acc_o[i][e] += d[e];
          g[i][e] = d[e] * __ldg(scale + v * 8 + e);
          sg += g[i][e];
          sgs += g[i][e] * sh[i][e];
        }
      }
    }
    sg = warp_sum(sg) / cols;
    sgs = warp_sum(sgs) / cols;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = lane + 32 * i;
      if (v < nvec) {
        float o[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          o[e] = rs * (g[i][e] - sg - sh[i][e] * sgs);
          acc_b[i][e] += o[e];  // column sum of ds = bias gradient of the linear layer that produced y
        }
        store8(ds + row * cols + v * 8, o);
      }
    }
  }
  // block-level column reduction, then one atomic per column per block
  float* rs_ = red;
  float* ro_ = red + kLnWarps * cols;
  float* rb_ = red + 2 * kLnWarps * cols;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int v = lane + 32 * i;
    if (v < nvec) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        rs_[warp * cols + v * 8 + e] = acc_s[i][e];
        ro_[warp * cols + v * 8 + e] = acc_o[i][e];
        if (dbias) rb_[warp * cols + v * 8 + e] = acc_b[i][e];
      }
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < cols; c += blockDim.x) {
    float a = 0.f, b = 0.f, d = 0.f;
#pragma unroll
    for (int w = 0; w < kLnWarps; ++w) {
      a += rs_[w * cols + c];
      b += ro_[w * cols + c];
      if (dbias) d += rb_[w * cols + c];
    }
    atomicAdd(dscale + c, a);
    atomicAdd(doffset + c, b);
    if (dbias) atomicAdd(dbias + c, d);
  }
}

// Single-pass variant for cols <= 512 and rows <= 32 * gridDim.x (every add+LN backward of a 4096-token training
// batch): one CTA of 32 warps per 32 rows, ONE row per warp, so every row of the batch is in flight at once (the
// looping kernel above needs two latency-bound passes at 4096 rows).  The column sums (dscale, doffset, dbias) of
// the CTA's 32 rows are combined through shared memory and leave as 16-byte vector reductions: the per-address
// serialisation of the L2 atomic unit made the 296-CTA x 1536 scalar atomics of the looping kernel cost ~6 us.
constexpr int kLn1pWarps = 32;

__device__ __forceinline__ void red_add_f32x4(float* addr, const float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

// W = warps (= rows) per CTA: 32 is the measured default; 16 (ZB_LN1P_WARPS=16, opt-in, untimed) halves the shared
// memory per CTA (two CTAs per SM, twice the CTAs) at the price of twice the vector reductions per column.
template <int NV, int W = kLn1pWarps>
__global__ void __launch_bounds__(W * 32)
add_ln_bwd_1pass_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ y,
                        const __nv_bfloat16* __restrict__ d_out, const __nv_bfloat16* __restrict__ d_out2,
                        const float* __restrict__ mean, const float* __restrict__ rstd,
                        const float* __restrict__ scale, __nv_bfloat16* __restrict__ ds, float* __restrict__ dscale,
                        float* __restrict__ doffset, float* __restrict__ dbias, long long rows, int cols) {
  grid_dep_wait();
  extern __shared__ __align__(16) float red[];  // [3][W][cols]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nvec = cols >> 3;
  const long long row = (long long)blockIdx.x * W + warp;
  const bool live = row < rows;
  float* r_s = red + (size_t)warp * cols;
  float* r_o = red + (size_t)(W + warp) * cols;
  float* r_b = red + (size_t)(2 * W + warp) * cols;
  float sh[NV][8], d[NV][8];
  float sg = 0.f, sgs = 0.f;
  float mu = 0.f, rs = 0.f;
  if (live) {
    mu = mean[row];
    rs = rstd[row];
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int v = lane + 32 * i;
#pragma unroll
    for (int e = 0; e < 8; ++e) sh[i][e] = d[i][e] = 0.f;
    if (live && v < nvec) {
      load8(x + row * cols + v * 8, sh[i]);
      if (y) {
        float t[8];
        load8(y + row * cols + v * 8, t);
#pragma unroll
        for (int e = 0; e < 8; ++e) sh[i][e] += t[e];
      }
      load8(d_out + row * cols + v * 8, d[i]);
      if (d_out2) {
        float t[8];
        load8(d_out2 + row * cols + v * 8, t);
#pragma unroll
        for (int e = 0; e < 8; ++e) d[i][e] += t[e];
      }
      const float4 sc0 = __ldg(reinterpret_cast<const float4*>(scale + v * 8)),
                   sc1 = __ldg(reinterpret_cast<const float4*>(scale + v * 8) + 1);
      const float sc[8] = {sc0.x, sc0.y, sc0.z, sc0.w, sc1.x, sc1.y, sc1.z, sc1.w};
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        sh[i][e] = (sh[i][e] - mu) * rs;
        const float g = d[i][e] * sc[e];
        sg += g;
        sgs += g * sh[i][e];
      }
    }
  }
  sg = warp_sum(sg) / cols;
  sgs = warp_sum(sgs) / cols;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int v = lane + 32 * i;
    if (v < nvec) {
      float o[8], ps[8];
      if (live) {
        const float4 sc0 = __ldg(reinterpret_cast<const float4*>(scale + v * 8)),
                     sc1 = __ldg(reinterpret_cast<const float4*>(scale + v * 8) + 1);
        const float sc[8] = {sc0.x, sc0.y, sc0.z, sc0.w, sc1.x, sc1.y, sc1.z, sc1.w};
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = rs * (d[i][e] * sc[e] - sg - sh[i][e] * sgs);
        store8(ds + row * cols + v * 8, o);
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = 0.f;
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) ps[e] = d[i][e] * sh[i][e];
      float4* ws = reinterpret_cast<float4*>(r_s + v * 8);
      float4* wo = reinterpret_cast<float4*>(r_o + v * 8);
      ws[0] = make_float4(ps[0], ps[1], ps[2], ps[3]);
      ws[1] = make_float4(ps[4], ps[5], ps[6], ps[7]);
      wo[0] = make_float4(d[i][0], d[i][1], d[i][2], d[i][3]);
      wo[1] = make_float4(d[i][4], d[i][5], d[i][6], d[i][7]);
      if (dbias) {
        float4* wb = reinterpret_cast<float4*>(r_b + v * 8);
        wb[0] = make_float4(o[0], o[1], o[2], o[3]);   // bf16-rounded ds is what the GEMMs see; the sum uses fp32
        wb[1] = make_float4(o[4], o[5], o[6], o[7]);
      }
    }
  }
  __syncthreads();
  // thread (q, c4): quantity q in {dscale, doffset, dbias}, 4 columns; sums the CTA's 32 row partials
  const int c4n = cols >> 2;
  for (int idx = threadIdx.x; idx < (dbias ? 3 : 2) * c4n; idx += W * 32) {
    const int q = idx / c4n, c4 = idx - q * c4n;
    const float* src = red + (size_t)q * W * cols + c4 * 4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
    for (int w = 0; w < W; ++w) {
      const float4 t = *reinterpret_cast<const float4*>(src + (size_t)w * cols);
      acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
    }
    float* dst = (q == 0 ? dscale : (q == 1 ? doffset : dbias)) + c4 * 4;
    red_add_f32x4(dst, acc);
  }
}

// Rows-per-warp variant (opt-in, ZB_LN1P_WARPS=8, untimed): 8 warps per CTA, R consecutive rows per warp.  A lane owns
// the same columns in every row, so the three column sums of a warp's R rows accumulate in registers for free and only
// one partial per WARP (not per row) goes through shared memory: R x less shared-memory traffic, a 256-thread CTA with
// 48 KB (d = 512) that can share an SM with its neighbours under programmatic launch, the same bytes in flight per SM
// (all R rows' loads are issued before the first reduction).  Same arithmetic per element as the kernel above; the
// column sums are taken in a different order (rows of a warp first).
constexpr int kLnRowsWarps = 8;
template <int NV, int R>
__global__ void __launch_bounds__(kLnRowsWarps * 32)
add_ln_bwd_rows_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ y,
                       const __nv_bfloat16* __restrict__ d_out, const __nv_bfloat16* __restrict__ d_out2,
                       const float* __restrict__ mean, const float* __restrict__ rstd,
                       const float* __restrict__ scale, __nv_bfloat16* __restrict__ ds, float* __restrict__ dscale,
                       float* __restrict__ doffset, float* __restrict__ dbias, long long rows, int cols) {
  grid_dep_wait();
  constexpr int W = kLnRowsWarps;
  extern __shared__ __align__(16) float red[];  // [3][W][cols]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nvec = cols >> 3;
  const long long row0 = ((long long)blockIdx.x * W + warp) * R;
  float mu[R], rs[R], sg[R], sgs[R];
  float sc[NV][8];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int v = lane + 32 * i;
#pragma unroll
    for (int e = 0; e < 8; ++e) sc[i][e] = 0.f;
    if (v < nvec) {
      const float4 sc0 = __ldg(reinterpret_cast<const float4*>(scale + v * 8)),
                   sc1 = __ldg(reinterpret_cast<const float4*>(scale + v * 8) + 1);
      sc[i][0] = sc0.x; sc[i][1] = sc0.y; sc[i][2] = sc0.z; sc[i][3] = sc0.w;
      sc[i][4] = sc1.x; sc[i][5] = sc1.y; sc[i][6] = sc1.z; sc[i][7] = sc1.w;
    }
  }
  // every load of the warp's R rows is in flight before anything is consumed; the rows stay packed (bf16) in
  // registers and are unpacked once for the row statistics and once for the outputs
  uint4 px[R][NV], py[R][NV], pd[R][NV], pd2[R][NV];
  const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const long long row = row0 + r;
    const bool live = row < rows;
    mu[r] = live ? mean[row] : 0.f;
    rs[r] = live ? rstd[row] : 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = lane + 32 * i;
      const bool on = live && v < nvec;
      const long long off = row * cols + v * 8;
      px[r][i] = on ? __ldg(reinterpret_cast<const uint4*>(x + off)) : zero4;
      pd[r][i] = on ? __ldg(reinterpret_cast<const uint4*>(d_out + off)) : zero4;
      py[r][i] = (on && y) ? __ldg(reinterpret_cast<const uint4*>(y + off)) : zero4;
      pd2[r][i] = (on && d_out2) ? __ldg(reinterpret_cast<const uint4*>(d_out2 + off)) : zero4;
    }
  }
  // xhat = (x + y - mu) * rstd and d = d_out + d_out2 of vector (r, i), same operation order as the kernel above
  auto unpack = [&](int r, int i, float (&xh)[8], float (&dd)[8]) {
    const uint32_t wx[4] = {px[r][i].x, px[r][i].y, px[r][i].z, px[r][i].w};
    const uint32_t wy[4] = {py[r][i].x, py[r][i].y, py[r][i].z, py[r][i].w};
    const uint32_t wd[4] = {pd[r][i].x, pd[r][i].y, pd[r][i].z, pd[r][i].w};
    const uint32_t w2[4] = {pd2[r][i].x, pd2[r][i].y, pd2[r][i].z, pd2[r][i].w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 a = unpack_bf16x2(wx[e]), b = unpack_bf16x2(wy[e]), c = unpack_bf16x2(wd[e]), g = unpack_bf16x2(w2[e]);
      xh[2 * e] = (a.x + b.x - mu[r]) * rs[r];
      xh[2 * e + 1] = (a.y + b.y - mu[r]) * rs[r];
      dd[2 * e] = c.x + g.x;
      dd[2 * e + 1] = c.y + g.y;
    }
  };
#pragma unroll
  for (int r = 0; r < R; ++r) {
    sg[r] = sgs[r] = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float xh[8], dd[8];
      unpack(r, i, xh, dd);
      if (row0 + r < rows && lane + 32 * i < nvec) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float g = dd[e] * sc[i][e];
          sg[r] += g;
          sgs[r] += g * xh[e];
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < R; ++r) {   // 2 R independent butterfly reductions in flight
    sg[r] = warp_sum(sg[r]) / cols;
    sgs[r] = warp_sum(sgs[r]) / cols;
  }
  // per-lane column partials over the warp's rows: dscale += d * xhat, doffset += d, dbias += ds
  float cs[NV][8], co[NV][8], cb[NV][8];
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int e = 0; e < 8; ++e) cs[i][e] = co[i][e] = cb[i][e] = 0.f;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const long long row = row0 + r;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = lane + 32 * i;
      if (row < rows && v < nvec) {
        float xh[8], dd[8], o[8];
        unpack(r, i, xh, dd);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          o[e] = rs[r] * (dd[e] * sc[i][e] - sg[r] - xh[e] * sgs[r]);
          cs[i][e] += dd[e] * xh[e];
          co[i][e] += dd[e];
          cb[i][e] += o[e];
        }
        store8(ds + row * cols + v * 8, o);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int v = lane + 32 * i;
    if (v < nvec) {
      float4* ws = reinterpret_cast<float4*>(red + (size_t)warp * cols + v * 8);
      float4* wo = reinterpret_cast<float4*>(red + (size_t)(W + warp) * cols + v * 8);
      ws[0] = make_float4(cs[i][0], cs[i][1], cs[i][2], cs[i][3]);
      ws[1] = make_float4(cs[i][4], cs[i][5], cs[i][6], cs[i][7]);
      wo[0] = make_float4(co[i][0], co[i][1], co[i][2], co[i][3]);
      wo[1] = make_float4(co[i][4], co[i][5], co[i][6], co[i][7]);
      if (dbias) {
        float4* wb = reinterpret_cast<float4*>(red + (size_t)(2 * W + warp) * cols + v * 8);
        wb[0] = make_float4(cb[i][0], cb[i][1], cb[i][2], cb[i][3]);
        wb[1] = make_float4(cb[i][4], cb[i][5], cb[i][6], cb[i][7]);
      }
    }
  }
  __syncthreads();
  // (quantity q, 4 columns) pairs are dealt round-robin to the CTA's 256 threads
  const int c4n = cols >> 2;
  const int items = (dbias ? 3 : 2) * c4n;
  for (int it = threadIdx.x; it < items; it += W * 32) {
    const int q = it / c4n, c4 = it % c4n;
    const float* src = red + (size_t)q * W * cols + c4 * 4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int w = 0; w < W; ++w) {
      const float4 t = *reinterpret_cast<const float4*>(src + (size_t)w * cols);
      acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
    }
    float* dst = (q == 0 ? dscale : (q == 1 ? doffset : dbias)) + c4 * 4;
    red_add_f32x4(dst, acc);
  }
}

static int check(const zb_add_ln_args* a, const char* who) {
  ZB_REQUIRE(a && a->x && a->scale, "%s: null pointer", who);
  ZB_REQUIRE(a->rows >= 0 && a->cols > 0 && a->cols % 8 == 0 && a->cols <= 8 * 32 * kLnMaxVec,
             "%s: cols must be a multiple of 8 and <= %d (got %lld)", who, 8 * 32 * kLnMaxVec, (long long)a->cols);
  return ZB_OK;
}

}  // namespace zb

#define ZB_LN_DISPATCH(NVEC, CALL)          \
  do {                                      \
    if (NVEC <= 1) { CALL(1); }             \
    else if (NVEC <= 2) { CALL(2); }        \
    else if (NVEC <= 4) { CALL(4); }        \
    else { CALL(8); }                       \
  } while (0)

extern "C" int zb_add_ln_fwd(const zb_add_ln_args* a, zb_stream_t stream) {
  using namespace zb;
  int rc = check(a, "zb_add_ln_fwd");
  if (rc) return rc;
  ZB_REQUIRE(a->out && a->offset, "zb_add_ln_fwd: null pointer");
  if (a->rows == 0) return ZB_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int nv = (int)((a->cols / 8 + 31) / 32);
  long long blocks = (a->rows + kLnWarps - 1) / kLnWarps;
  const long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (a->y32) {
    ZB_REQUIRE(!a->y && (reinterpret_cast<uintptr_t>(a->y32) & 15) == 0 &&
                   (!a->ybias || (reinterpret_cast<uintptr_t>(a->ybias) & 15) == 0),
               "zb_add_ln_fwd: y32 excludes y and needs 16-byte aligned y32 / ybias");
#define CALLA(N)                                                                                             \
  ZB_LAUNCH(add_ln_fwd_acc_kernel<N>, (int)blocks, kLnWarps * 32, 0, st, (const __nv_bfloat16*)a->x, a->y32, \
            a->ybias, (__nv_bfloat16*)a->out, a->mean, a->rstd, a->scale, a->offset, a->rows, (int)a->cols, a->eps)
    ZB_LN_DISPATCH(nv, CALLA);
#undef CALLA
    return check_launch("zb_add_ln_fwd(fp32 accumulator)");
  }
#define CALL(N)                                                                                              \
  ZB_LAUNCH(add_ln_fwd_kernel<N>, (int)blocks, kLnWarps * 32, 0, st,                                                \
      (const __nv_bfloat16*)a->x, (const __nv_bfloat16*)a->y, (__nv_bfloat16*)a->out, a->mean, a->rstd, a->scale, \
      a->offset, a->rows, (int)a->cols, a->eps)
  ZB_LN_DISPATCH(nv, CALL);
#undef CALL
  return check_launch("zb_add_ln_fwd");
}

extern "C" int zb_add_ln_bwd(const zb_add_ln_args* a, zb_stream_t stream) {
  using namespace zb;
  int rc = check(a, "zb_add_ln_bwd");
  if (rc) return rc;
  ZB_REQUIRE(a->d_out && a->ds && a->dscale && a->doffset && a->mean && a->rstd, "zb_add_ln_bwd: null pointer");
  if (a->rows == 0) return ZB_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int nv = (int)((a->cols / 8 + 31) / 32);
  static const bool no_1pass = getenv("ZB_LN_BWD_LOOP") != nullptr;
  const bool vec_ok = ((reinterpret_cast<uintptr_t>(a->dscale) | reinterpret_cast<uintptr_t>(a->doffset) |
                        reinterpret_cast<uintptr_t>(a->dbias)) & 15) == 0;
  static const bool rows8 = getenv("ZB_LN1P_WARPS") != nullptr && atoi(getenv("ZB_LN1P_WARPS")) == 8;
  if (!no_1pass && rows8 && nv <= 2 && vec_ok && a->rows <= (long long)num_sms() * 64) {
    constexpr int R = 4;
    const int grid = (int)((a->rows + kLnRowsWarps * R - 1) / (kLnRowsWarps * R));
    const size_t smem1 = (size_t)3 * kLnRowsWarps * a->cols * sizeof(float);
#define CALLR(N)                                                                                             \
  do {                                                                                                       \
    static bool attr = false;                                                                                \
    if (!attr) {                                                                                             \
      cudaFuncSetAttribute(add_ln_bwd_rows_kernel<N, R>, cudaFuncAttributeMaxDynamicSharedMemorySize,        \
                           3 * kLnRowsWarps * 512 * (int)sizeof(float));                                     \
      attr = true;                                                                                           \
    }                                                                                                        \
    ZB_LAUNCH((add_ln_bwd_rows_kernel<N, R>), grid, kLnRowsWarps * 32, smem1, st,                            \
        (const __nv_bfloat16*)a->x, (const __nv_bfloat16*)a->y, (const __nv_bfloat16*)a->d_out,              \
        (const __nv_bfloat16*)a->d_out2, a->mean, a->rstd, a->scale, (__nv_bfloat16*)a->ds, a->dscale,       \
        a->doffset, a->dbias, a->rows, (int)a->cols);                                                        \
  } while (0)
    if (nv <= 1) CALLR(1); else CALLR(2);
#undef CALLR
    return check_launch("zb_add_ln_bwd(rows per warp)");
  }
  // ZB_LN1P_WARPS = 16 / 4 (calibration): fewer rows per CTA -> less shared memory, several CTAs per SM in different
  // phases, more vector reductions per column (rows8 above is the rows-per-warp kernel)
  static const int wsel = getenv("ZB_LN1P_WARPS") ? atoi(getenv("ZB_LN1P_WARPS")) : 0;
  if (!no_1pass && (wsel == 16 || wsel == 4 || wsel == 2) && nv <= 2 && vec_ok && a->rows <= (long long)num_sms() * 64) {
#define CALLW(N, WW)                                                                                         \
  do {                                                                                                       \
    static bool attr = false;                                                                                \
    if (!attr) {                                                                                             \
      cudaFuncSetAttribute(add_ln_bwd_1pass_kernel<N, WW>, cudaFuncAttributeMaxDynamicSharedMemorySize,      \
                           3 * WW * 512 * (int)sizeof(float));                                               \
      attr = true;                                                                                           \
    }                                                                                                        \
    ZB_LAUNCH((add_ln_bwd_1pass_kernel<N, WW>), (int)((a->rows + WW - 1) / WW), WW * 32,                     \
        (size_t)3 * WW * a->cols * sizeof(float), st,                                                        \
        (const __nv_bfloat16*)a->x, (const __nv_bfloat16*)a->y, (const __nv_bfloat16*)a->d_out,              \
        (const __nv_bfloat16*)a->d_out2, a->mean, a->rstd, a->scale, (__nv_bfloat16*)a->ds, a->dscale,       \
        a->doffset, a->dbias, a->rows, (int)a->cols);                                                        \
  } while (0)
    if (wsel == 16) { if (nv <= 1) CALLW(1, 16); else CALLW(2, 16); }
    else if (wsel == 4) { if (nv <= 1) CALLW(1, 4); else CALLW(2, 4); }
    else { if (nv <= 1) CALLW(1, 2); else CALLW(2, 2); }
#undef CALLW
    return check_launch("zb_add_ln_bwd(1pass, W rows)");
  }
  if (!no_1pass && nv <= 2 && vec_ok && 3 * (a->cols / 4) <= kLn1pWarps * 32 &&
      a->rows <= (long long)num_sms() * kLn1pWarps) {
    const int grid = (int)((a->rows + kLn1pWarps - 1) / kLn1pWarps);
    const size_t smem1 = (size_t)3 * kLn1pWarps * a->cols * sizeof(float);
#define CALL1(N)                                                                                             \
  do {                                                                                                       \
    static bool attr = false;                                                                                \
    if (!attr) {                                                                                             \
      cudaFuncSetAttribute(add_ln_bwd_1pass_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize,          \
                           3 * kLn1pWarps * 512 * (int)sizeof(float));                                       \
      attr = true;                                                                                           \
    }                                                                                                        \
    ZB_LAUNCH(add_ln_bwd_1pass_kernel<N>, grid, kLn1pWarps * 32, smem1, st,                                  \
        (const __nv_bfloat16*)a->x, (const __nv_bfloat16*)a->y, (const __nv_bfloat16*)a->d_out,              \
        (const __nv_bfloat16*)a->d_out2, a->mean, a->rstd, a->scale, (__nv_bfloat16*)a->ds, a->dscale,       \
        a->doffset, a->dbias, a->rows, (int)a->cols);                                                        \
  } while (0)
    if (nv <= 1) CALL1(1); else CALL1(2);
#undef CALL1
    return check_launch("zb_add_ln_bwd(1pass)");
  }
  long long blocks = (a->rows + kLnWarps - 1) / kLnWarps;
  const long long cap = (long long)num_sms() * 2;
  if (blocks > cap) blocks = cap;
  const size_t smem = (size_t)3 * kLnWarps * a->cols * sizeof(float);
#define CALL(N)                                                                                              \
  do {                                                                                                       \
    if (smem > 48 * 1024)                                                                                    \
      cudaFuncSetAttribute(add_ln_bwd_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);    \
    ZB_LAUNCH(add_ln_bwd_kernel<N>, (int)blocks, kLnWarps * 32, smem, st,                                           \
        (const __nv_bfloat16*)a->x, (const __nv_bfloat16*)a->y, (const __nv_bfloat16*)a->d_out,              \
        (const __nv_bfloat16*)a->d_out2, a->mean, a->rstd, a->scale, (__nv_bfloat16*)a->ds, a->dscale,       \
        a->doffset, a->dbias, a->rows, (int)a->cols);                                                                  \
  } while (0)
  ZB_LN_DISPATCH(nv, CALL);
#undef CALL
  return check_launch("zb_add_ln_bwd");
}
