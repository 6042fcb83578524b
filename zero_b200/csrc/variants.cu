// variants.cu — HBM-bound pieces of the Transformer variants on the hot path:
//   * ReLA's gated RMS norm  scale * x * rsqrt(mean(x^2) + eps) * sigmoid(gate * x)   (modules/rela.py:95-109)
//   * Average Attention: masked prefix mean (models/transformer_aan.py:99-108, func.py:389-398), the cached
//     running sum of decode (transformer_aan.py:110-112) and the input/forget gate (transformer_aan.py:185-189)
//   * plain bf16 add (merged attention: o_cross + aan_o, func.py:274-275)
// One warp per row with 16-byte vectors, like addln.cu.
#include "zb_common.h"
#include "zb_ptx.cuh"

namespace zb {

constexpr int kVWarps = 8;

__device__ __forceinline__ float vsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ void ld8(const __nv_bfloat16* p, float (&f)[8]) {
  const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float2 t = unpack_bf16x2(w[e]);
    f[2 * e] = t.x;
    f[2 * e + 1] = t.y;
  }
}
__device__ __forceinline__ void st8(__nv_bfloat16* p, const float (&f)[8]) {
  uint4 u;
  u.x = pack_bf16x2(f[0], f[1]);
  u.y = pack_bf16x2(f[2], f[3]);
  u.z = pack_bf16x2(f[4], f[5]);
  u.w = pack_bf16x2(f[6], f[7]);
  *reinterpret_cast<uint4*>(p) = u;
}
__device__ __forceinline__ float sigm(float x) { return 1.f / (1.f + __expf(-x)); }

// ---------------------------------------------------------------- gated RMS norm (cols <= 2048)
__global__ void __launch_bounds__(kVWarps * 32)
gated_rms_fwd_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out, float* __restrict__ rstd,
                     const float* __restrict__ scale, const float* __restrict__ gate, long long rows, int cols,
                     float eps) {
  grid_dep_wait();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nvec = cols >> 3;
  for (long long row = (long long)blockIdx.x * kVWarps + warp; row < rows; row += (long long)gridDim.x * kVWarps) {
    float xs[8][8];
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int v = lane + 32 * i;
      if (v < nvec) {
        ld8(x + row * cols + v * 8, xs[i]);
#pragma unroll
        for (int e = 0; e < 8; ++e) sq += xs[i][e] * xs[i][e];
      }
    }
    const float rs = rsqrtf(vsum(sq) / cols + eps);
    if (lane == 0 && rstd) rstd[row] = rs;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int v = lane + 32 * i;
      if (v < nvec) {
        float o[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int c = v * 8 + e;
          o[e] = __ldg(scale + c) * xs[i][e] * rs * sigm(__ldg(gate + c) * xs[i][e]);
        }
        st8(out + row * cols + v * 8, o);
      }
    }
  }
}

// y = s * x * r * sig(g x), r = rsqrt(mean(x^2) + eps)
// dx = dy*s*r*sig*(1 + g x (1 - sig))  -  x * r^3 / cols * sum_c(dy*s*x*sig)
// dscale = sum_rows dy * x * r * sig ;  dgate = sum_rows dy * s * x * r * sig * (1 - sig) * x
__global__ void __launch_bounds__(kVWarps * 32)
gated_rms_bwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy,
                     const float* __restrict__ rstd, const float* __restrict__ scale,
                     const float* __restrict__ gate, __nv_bfloat16* __restrict__ dx, float* __restrict__ dscale,
                     float* __restrict__ dgate, long long rows, int cols) {
  grid_dep_wait();
  extern __shared__ float red[];  // 2 x [kVWarps][cols]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nvec = cols >> 3;
  for (int c = threadIdx.x; c < 2 * kVWarps * cols; c += blockDim.x) red[c] = 0.f;
  __syncthreads();
  float* rs_ = red + warp * cols;
  float* rg_ = red + kVWarps * cols + warp * cols;
  for (long long row = (long long)blockIdx.x * kVWarps + warp; row < rows; row += (long long)gridDim.x * kVWarps) {
    const float r = rstd[row];
    float xs[8][8], ds[8][8];
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int v = lane + 32 * i;
      if (v < nvec) {
        ld8(x + row * cols + v * 8, xs[i]);
        ld8(dy + row * cols + v * 8, ds[i]);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int c = v * 8 + e;
          const float s = __ldg(scale + c), g = __ldg(gate + c);
          const float sg = sigm(g * xs[i][e]);
          dot += ds[i][e] * s * xs[i][e] * sg;
          rs_[c] += ds[i][e] * xs[i][e] * r * sg;
          rg_[c] += ds[i][e] * s * xs[i][e] * r * sg * (1.f - sg) * xs[i][e];
        }
      }
    }
    dot = vsum(dot);
    const float k = r * r * r / cols * dot;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int v = lane + 32 * i;
      if (v < nvec) {
        float o[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int c = v * 8 + e;
          const float s = __ldg(scale + c), g = __ldg(gate + c);
          const float sg = sigm(g * xs[i][e]);
          o[e] = ds[i][e] * s * r * sg * (1.f + g * xs[i][e] * (1.f - sg)) - xs[i][e] * k;
        }
        st8(dx + row * cols + v * 8, o);
      }
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < cols; c += blockDim.x) {
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int w = 0; w < kVWarps; ++w) {
      a += red[w * cols + c];
      b += red[kVWarps * cols + w * cols + c];
    }
    atomicAdd(dscale + c, a);
    atomicAdd(dgate + c, b);
  }
}

// ---------------------------------------------------------------- AAN gate: out = sig(i) * x + sig(f) * y, z = [i | f]
__global__ void aan_gate_fwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ y,
                                    const __nv_bfloat16* __restrict__ z, __nv_bfloat16* __restrict__ out,
                                    long long rows, int d) {
  grid_dep_wait();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int nvec = d >> 3;
  if (idx >= rows * nvec) return;
  const long long row = idx / nvec;
  const int v = (int)(idx % nvec);
  float xv[8], yv[8], iv[8], fv[8], o[8];
  ld8(x + row * d + v * 8, xv);
  ld8(y + row * d + v * 8, yv);
  ld8(z + row * 2 * d + v * 8, iv);
  ld8(z + row * 2 * d + d + v * 8, fv);
#pragma unroll
  for (int e = 0; e < 8; ++e) o[e] = sigm(iv[e]) * xv[e] + sigm(fv[e]) * yv[e];
  st8(out + row * d + v * 8, o);
}
__global__ void aan_gate_bwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ y,
                                    const __nv_bfloat16* __restrict__ z, const __nv_bfloat16* __restrict__ dout,
                                    __nv_bfloat16* __restrict__ dx, __nv_bfloat16* __restrict__ dy,
                                    __nv_bfloat16* __restrict__ dz, long long rows, int d) {
  grid_dep_wait();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int nvec = d >> 3;
  if (idx >= rows * nvec) return;
  const long long row = idx / nvec;
  const int v = (int)(idx % nvec);
  float xv[8], yv[8], iv[8], fv[8], g[8], ox[8], oy[8], oi[8], of[8];
  ld8(x + row * d + v * 8, xv);
  ld8(y + row * d + v * 8, yv);
  ld8(z + row * 2 * d + v * 8, iv);
  ld8(z + row * 2 * d + d + v * 8, fv);
  ld8(dout + row * d + v * 8, g);
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const float si = sigm(iv[e]), sf = sigm(fv[e]);
    ox[e] = g[e] * si;
    oy[e] = g[e] * sf;
    oi[e] = g[e] * xv[e] * si * (1.f - si);
    of[e] = g[e] * yv[e] * sf * (1.f - sf);
  }
  st8(dx + row * d + v * 8, ox);
  st8(dy + row * d + v * 8, oy);
  st8(dz + row * 2 * d + v * 8, oi);
  st8(dz + row * 2 * d + d + v * 8, of);
}

// ---------------------------------------------------------------- masked prefix mean over time
// mode 0 ("aan" bias, func.py:389-398): y[t] = mean_{s<=t} x[s] for t < len, 0 for pad rows
// mode 1 (cumsum / count, transformer_aan.py:103-108): y[t] = cumsum(x)[t] / max(count_valid(<=t), 1)
__global__ void prefix_mean_fwd_kernel2(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                        const int32_t* __restrict__ lens, int batch, int len, int dim, int mode) {
  grid_dep_wait();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int half = dim >> 1;
  if (i >= (long long)batch * half) return;
  const int b = (int)(i / half), c = (int)(i % half) * 2;
  const int L = lens ? lens[b] : len;
  float2 acc = make_float2(0.f, 0.f);
  for (int t = 0; t < len; ++t) {
    const long long o = ((long long)b * len + t) * dim + c;
    const float2 v = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(x + o));
    float2 r;
    if (mode == 0) {
      if (t < L) {
        acc.x += v.x;
        acc.y += v.y;
        const float inv = 1.f / (float)(t + 1);
        r = make_float2(acc.x * inv, acc.y * inv);
      } else {
        r = make_float2(0.f, 0.f);
      }
    } else {
      acc.x += v.x;
      acc.y += v.y;
      const int cnt = t < L ? t + 1 : L;
      const float inv = 1.f / (float)(cnt < 1 ? 1 : cnt);
      r = make_float2(acc.x * inv, acc.y * inv);
    }
    *reinterpret_cast<uint32_t*>(y + o) = pack_bf16x2(r.x, r.y);
  }
}
__global__ void prefix_mean_bwd_kernel2(const __nv_bfloat16* __restrict__ dy, __nv_bfloat16* __restrict__ dx,
                                        const int32_t* __restrict__ lens, int batch, int len, int dim, int mode) {
  grid_dep_wait();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int half = dim >> 1;
  if (i >= (long long)batch * half) return;
  const int b = (int)(i / half), c = (int)(i % half) * 2;
  const int L = lens ? lens[b] : len;
  float2 acc = make_float2(0.f, 0.f);
  for (int t = len - 1; t >= 0; --t) {
    const long long o = ((long long)b * len + t) * dim + c;
    const float2 v = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(dy + o));
    float2 r;
    if (mode == 0) {
      if (t < L) {
        const float inv = 1.f / (float)(t + 1);
        acc.x += v.x * inv;
        acc.y += v.y * inv;
        r = acc;
      } else {
        r = make_float2(0.f, 0.f);
      }
    } else {
      const int cnt = t < L ? t + 1 : L;
      const float inv = 1.f / (float)(cnt < 1 ? 1 : cnt);
      acc.x += v.x * inv;
      acc.y += v.y * inv;
      r = acc;
    }
    *reinterpret_cast<uint32_t*>(dx + o) = pack_bf16x2(r.x, r.y);
  }
}

// cached decode: y = (x + sum) / (t + 1); sum += x   (transformer_aan.py:110-112), fp32 running sum
__global__ void aan_step_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ sum,
                                __nv_bfloat16* __restrict__ y, long long n, float inv) {
  grid_dep_wait();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float s = sum[i] + __bfloat162float(x[i]);
  sum[i] = s;
  y[i] = __float2bfloat16(s * inv);
}

// The three launches at the head of the cached average-attention sublayer in one: sum += x; y = sum / (t + 1);
// cat = [x | y] (tf.concat of transformer_aan.py:185) and a contiguous copy of y for the gate.  Same arithmetic and
// rounding as aan_step_kernel + two add2d copies.
__global__ void aan_cat_step_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ sum,
                                    __nv_bfloat16* __restrict__ cat, long long ldcat, __nv_bfloat16* __restrict__ y,
                                    long long rows, int d, float inv) {
  grid_dep_wait();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int nvec = d >> 3;
  if (i >= rows * nvec) return;
  const long long r = i / nvec;
  const int v = (int)(i % nvec);
  const uint4 xr = __ldg(reinterpret_cast<const uint4*>(x + r * d + v * 8));
  const uint32_t w[4] = {xr.x, xr.y, xr.z, xr.w};
  float4* sp = reinterpret_cast<float4*>(sum + r * d + v * 8);
  float4 s0 = sp[0], s1 = sp[1];
  float s[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
  float o[8];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float2 t = unpack_bf16x2(w[e]);
    s[2 * e] += t.x;
    s[2 * e + 1] += t.y;
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) o[e] = s[e] * inv;
  sp[0] = make_float4(s[0], s[1], s[2], s[3]);
  sp[1] = make_float4(s[4], s[5], s[6], s[7]);
  *reinterpret_cast<uint4*>(cat + r * ldcat + v * 8) = xr;
  st8(cat + r * ldcat + d + v * 8, o);
  st8(y + r * d + v * 8, o);
}

// Gate + residual + LayerNorm of the average-attention sublayer in one pass over the row (one warp per row):
// g = sigmoid(i) x + sigmoid(f) y with z = [i | f] (transformer_aan.py:185-189), out = LN(x + g) (func.py:289-324).
// g is rounded to bf16 before the residual add, exactly what aan_gate_fwd_kernel hands to add_ln_fwd_kernel, and the
// reductions run in the same order, so the fused and the two-kernel paths agree bit for bit.
template <int NV>
__global__ void __launch_bounds__(kVWarps * 32)
aan_gate_ln_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ y,
                   const __nv_bfloat16* __restrict__ z, __nv_bfloat16* __restrict__ out,
                   const float* __restrict__ scale, const float* __restrict__ offset, long long rows, int d, float eps) {
  grid_dep_wait();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nvec = d >> 3;
  const long long row = (long long)blockIdx.x * kVWarps + warp;
  if (row >= rows) return;
  float s[NV][8];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int v = lane + 32 * i;
    if (v < nvec) {
      float xv[8], yv[8], iv[8], fv[8];
      ld8(x + row * d + v * 8, xv);
      ld8(y + row * d + v * 8, yv);
      ld8(z + row * 2 * d + v * 8, iv);
      ld8(z + row * 2 * d + d + v * 8, fv);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float g = sigm(iv[e]) * xv[e] + sigm(fv[e]) * yv[e];
        s[i][e] = xv[e] + __bfloat162float(__float2bfloat16(g));
        sum += s[i][e];
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mu = sum / d;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    if (lane + 32 * i < nvec) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float dd = s[i][e] - mu;
        sq += dd * dd;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float rs = rsqrtf(sq / d + eps);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int v = lane + 32 * i;
    if (v < nvec) {
      float o[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = __ldg(scale + v * 8 + e) * (s[i][e] - mu) * rs + __ldg(offset + v * 8 + e);
      st8(out + row * d + v * 8, o);
    }
  }
}

// out[r, :cols] = a[r, :cols] (+ b[r, :cols]); every operand a strided 2-D bf16 view (pitches in elements)
__global__ void add2d_kernel(const __nv_bfloat16* __restrict__ a, long long lda, const __nv_bfloat16* __restrict__ b,
                             long long ldb, __nv_bfloat16* __restrict__ out, long long ldo, long long rows, int cols) {
  grid_dep_wait();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int nvec = cols >> 3;
  if (i >= rows * nvec) return;
  const long long r = i / nvec;
  const int v = (int)(i % nvec);
  float x[8];
  ld8(a + r * lda + v * 8, x);
  if (b) {
    float y[8];
    ld8(b + r * ldb + v * 8, y);
#pragma unroll
    for (int e = 0; e < 8; ++e) x[e] += y[e];
  }
  st8(out + r * ldo + v * 8, x);
}

}  // namespace zb

using namespace zb;
#define ST(s) reinterpret_cast<cudaStream_t>(s)

extern "C" int zb_gated_rms_fwd(const void* x, void* out, float* rstd, const float* scale, const float* gate,
                                int64_t rows, int64_t cols, float eps, zb_stream_t stream) {
  ZB_REQUIRE(x && out && scale && gate && cols > 0 && cols % 8 == 0 && cols <= 2048, "zb_gated_rms_fwd: bad args");
  if (rows == 0) return ZB_OK;
  long long blocks = (rows + kVWarps - 1) / kVWarps;
  if (blocks > 16ll * num_sms()) blocks = 16ll * num_sms();
  ZB_LAUNCH(gated_rms_fwd_kernel, (int)blocks, kVWarps * 32, 0, ST(stream), (const __nv_bfloat16*)x, (__nv_bfloat16*)out, rstd,
                                                                    scale, gate, rows, (int)cols, eps);
  return check_launch("zb_gated_rms_fwd");
}
extern "C" int zb_gated_rms_bwd(const void* x, const void* dy, const float* rstd, const float* scale, const float* gate,
                                void* dx, float* dscale, float* dgate, int64_t rows, int64_t cols,
                                zb_stream_t stream) {
  ZB_REQUIRE(x && dy && rstd && scale && gate && dx && dscale && dgate && cols > 0 && cols % 8 == 0 && cols <= 2048,
             "zb_gated_rms_bwd: bad args");
  if (rows == 0) return ZB_OK;
  long long blocks = (rows + kVWarps - 1) / kVWarps;
  if (blocks > 2ll * num_sms()) blocks = 2ll * num_sms();
  const size_t smem = (size_t)2 * kVWarps * cols * sizeof(float);
  if (smem > 48 * 1024) cudaFuncSetAttribute(gated_rms_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  ZB_LAUNCH(gated_rms_bwd_kernel, (int)blocks, kVWarps * 32, smem, ST(stream), 
      (const __nv_bfloat16*)x, (const __nv_bfloat16*)dy, rstd, scale, gate, (__nv_bfloat16*)dx, dscale, dgate, rows,
      (int)cols);
  return check_launch("zb_gated_rms_bwd");
}
extern "C" int zb_aan_gate_fwd(const void* x, const void* y, const void* z, void* out, int64_t rows, int32_t dim,
                               zb_stream_t stream) {
  ZB_REQUIRE(x && y && z && out && dim > 0 && dim % 8 == 0, "zb_aan_gate_fwd: bad args");
  const long long n = rows * (dim / 8);
  if (n == 0) return ZB_OK;
  ZB_LAUNCH(aan_gate_fwd_kernel, (unsigned)((n + 255) / 256), 256, 0, ST(stream), 
      (const __nv_bfloat16*)x, (const __nv_bfloat16*)y, (const __nv_bfloat16*)z, (__nv_bfloat16*)out, rows, dim);
  return check_launch("zb_aan_gate_fwd");
}
extern "C" int zb_aan_gate_bwd(const void* x, const void* y, const void* z, const void* dout, void* dx, void* dy,
                               void* dz, int64_t rows, int32_t dim, zb_stream_t stream) {
  ZB_REQUIRE(x && y && z && dout && dx && dy && dz && dim > 0 && dim % 8 == 0, "zb_aan_gate_bwd: bad args");
  const long long n = rows * (dim / 8);
  if (n == 0) return ZB_OK;
  ZB_LAUNCH(aan_gate_bwd_kernel, (unsigned)((n + 255) / 256), 256, 0, ST(stream), 
      (const __nv_bfloat16*)x, (const __nv_bfloat16*)y, (const __nv_bfloat16*)z, (const __nv_bfloat16*)dout,
      (__nv_bfloat16*)dx, (__nv_bfloat16*)dy, (__nv_bfloat16*)dz, rows, dim);
  return check_launch("zb_aan_gate_bwd");
}
extern "C" int zb_prefix_mean_fwd(const void* x, void* y, const int32_t* lens, int32_t batch, int32_t len, int32_t dim,
                                  int32_t mode, zb_stream_t stream) {
  ZB_REQUIRE(x && y && batch >= 0 && len > 0 && dim % 2 == 0 && (mode == 0 || mode == 1), "zb_prefix_mean_fwd: bad args");
  const long long n = (long long)batch * (dim / 2);
  if (n == 0) return ZB_OK;
  ZB_LAUNCH(prefix_mean_fwd_kernel2, (unsigned)((n + 127) / 128), 128, 0, ST(stream), (const __nv_bfloat16*)x, (__nv_bfloat16*)y,
                                                                              lens, batch, len, dim, mode);
  return check_launch("zb_prefix_mean_fwd");
}
extern "C" int zb_prefix_mean_bwd(const void* dy, void* dx, const int32_t* lens, int32_t batch, int32_t len, int32_t dim,
                                  int32_t mode, zb_stream_t stream) {
  ZB_REQUIRE(dy && dx && batch >= 0 && len > 0 && dim % 2 == 0 && (mode == 0 || mode == 1), "zb_prefix_mean_bwd: bad args");
  const long long n = (long long)batch * (dim / 2);
  if (n == 0) return ZB_OK;
  ZB_LAUNCH(prefix_mean_bwd_kernel2, (unsigned)((n + 127) / 128), 128, 0, ST(stream), (const __nv_bfloat16*)dy, (__nv_bfloat16*)dx,
                                                                              lens, batch, len, dim, mode);
  return check_launch("zb_prefix_mean_bwd");
}
extern "C" int zb_aan_step(const void* x, float* sum, void* y, int64_t n, int32_t time, zb_stream_t stream) {
  ZB_REQUIRE(x && sum && y && n >= 0 && time >= 0, "zb_aan_step: bad args");
  if (n == 0) return ZB_OK;
  ZB_LAUNCH(aan_step_kernel, (unsigned)((n + 255) / 256), 256, 0, ST(stream), (const __nv_bfloat16*)x, sum, (__nv_bfloat16*)y, n,
                                                                      1.f / (float)(time + 1));
  return check_launch("zb_aan_step");
}
extern "C" int zb_aan_cat_step(const void* x, float* sum, void* cat, int64_t ldcat, void* y, int64_t rows, int32_t dim,
                               int32_t time, zb_stream_t stream) {
  ZB_REQUIRE(x && sum && cat && y && rows >= 0 && dim > 0 && dim % 8 == 0 && ldcat % 8 == 0 && ldcat >= 2 * dim &&
                 time >= 0,
             "zb_aan_cat_step: bad args");
  ZB_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(sum) | reinterpret_cast<uintptr_t>(cat) |
               reinterpret_cast<uintptr_t>(y)) & 15) == 0,
             "zb_aan_cat_step: operands must be 16-byte aligned");
  const long long n = rows * (dim / 8);
  if (n == 0) return ZB_OK;
  ZB_LAUNCH(aan_cat_step_kernel, (unsigned)((n + 255) / 256), 256, 0, ST(stream), (const __nv_bfloat16*)x, sum,
            (__nv_bfloat16*)cat, (long long)ldcat, (__nv_bfloat16*)y, (long long)rows, (int)dim,
            1.f / (float)(time + 1));
  return check_launch("zb_aan_cat_step");
}
extern "C" int zb_aan_gate_ln(const void* x, const void* y, const void* z, void* out, const float* scale,
                              const float* offset, int64_t rows, int32_t dim, float eps, zb_stream_t stream) {
  ZB_REQUIRE(x && y && z && out && scale && offset && rows >= 0 && dim > 0 && dim % 8 == 0 && dim <= 8 * 32 * 8,
             "zb_aan_gate_ln: bad args (dim must be a multiple of 8, <= 2048)");
  if (rows == 0) return ZB_OK;
  const unsigned grid = (unsigned)((rows + kVWarps - 1) / kVWarps);
  const int nv = (dim / 8 + 31) / 32;
#define ZB_GLN(N)                                                                                                   \
  ZB_LAUNCH(aan_gate_ln_kernel<N>, grid, kVWarps * 32, 0, ST(stream), (const __nv_bfloat16*)x, (const __nv_bfloat16*)y, \
            (const __nv_bfloat16*)z, (__nv_bfloat16*)out, scale, offset, (long long)rows, (int)dim, eps)
  if (nv <= 1) ZB_GLN(1);
  else if (nv <= 2) ZB_GLN(2);
  else if (nv <= 4) ZB_GLN(4);
  else ZB_GLN(8);
#undef ZB_GLN
  return check_launch("zb_aan_gate_ln");
}
extern "C" int zb_add2d(const void* a, int64_t lda, const void* b, int64_t ldb, void* out, int64_t ldo, int64_t rows,
                        int64_t cols, zb_stream_t stream) {
  ZB_REQUIRE(a && out && rows >= 0 && cols > 0 && cols % 8 == 0 && lda % 8 == 0 && ldo % 8 == 0 && (!b || ldb % 8 == 0),
             "zb_add2d: cols and pitches must be multiples of 8");
  const long long n = rows * (cols / 8);
  if (n == 0) return ZB_OK;
  ZB_LAUNCH(add2d_kernel, (unsigned)((n + 255) / 256), 256, 0, ST(stream), (const __nv_bfloat16*)a, lda, (const __nv_bfloat16*)b, ldb,
                                                                   (__nv_bfloat16*)out, ldo, rows, (int)cols);
  return check_launch("zb_add2d");
}
