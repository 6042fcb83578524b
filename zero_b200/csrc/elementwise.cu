// elementwise.cu — small HBM-bound helpers around the hot path: dtype casts of the fp32 master weights
// (utils/dtype.py:55-69), bias-gradient column sums (tf.nn.bias_add grad), global-norm partial sums
// (utils/cycle.py:94), TF-semantics Adam (main.py:178-181), row gathers for beam reordering
// (search.py:205-209) and the Average-Attention prefix mean (models/transformer_aan.py:99-108).
#include "zb_common.h"
#include "zb_ptx.cuh"

namespace zb {

__global__ void cast_f32_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long n) {
  grid_dep_wait();
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    const float4 v = *reinterpret_cast<const float4*>(src + i);
    uint2 o;
    o.x = pack_bf16x2(v.x, v.y);
    o.y = pack_bf16x2(v.z, v.w);
    *reinterpret_cast<uint2*>(dst + i) = o;
  } else {
    for (long long j = i; j < n; ++j) dst[j] = __float2bfloat16(src[j]);
  }
}
__global__ void cast_bf16_f32_kernel(const __nv_bfloat16* __restrict__ src, float* __restrict__ dst, long long n) {
  grid_dep_wait();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = __bfloat162float(src[i]);
}

// out[n] += sum_m x[m, n]; grid (col-chunks, row-chunks); each thread owns 2 columns, block reduces over rows.
__global__ void __launch_bounds__(256)
colsum_kernel(const __nv_bfloat16* __restrict__ x, long long m, long long n, long long ld, float* __restrict__ out,
              long long rows_per_block) {
  grid_dep_wait();
  __shared__ float2 red[8][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long col = ((long long)blockIdx.x * 32 + lane) * 2;
  const long long r0 = (long long)blockIdx.y * rows_per_block;
  long long r1 = r0 + rows_per_block;
  if (r1 > m) r1 = m;
  float2 acc = make_float2(0.f, 0.f);
  if (col + 1 < n) {
    for (long long r = r0 + warp; r < r1; r += 8) {
      const float2 t = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(x + r * ld + col));
      acc.x += t.x;
      acc.y += t.y;
    }
  } else if (col < n) {
    for (long long r = r0 + warp; r < r1; r += 8) acc.x += __bfloat162float(x[r * ld + col]);
  }
  red[warp][lane] = acc;
  __syncthreads();
  if (warp == 0) {
    float2 t = red[0][lane];
#pragma unroll
    for (int w = 1; w < 8; ++w) {
      t.x += red[w][lane].x;
      t.y += red[w][lane].y;
    }
    if (col < n) atomicAdd(out + col, t.x);
    if (col + 1 < n) atomicAdd(out + col + 1, t.y);
  }
}

// Vectorised variant for n % 8 == 0: a warp reads 512 contiguous bytes of one row (32 x 16 B), 8 row lanes per
// block, 4 rows in flight per thread; block-level reduction, one atomic per column per block.
__device__ __forceinline__ void colsum_vec_body(const __nv_bfloat16* __restrict__ x, long long m, long long nvec,
                                                long long ld, float* __restrict__ out, long long rows_per_block,
                                                unsigned bx, unsigned by) {
  __shared__ float red[8][32][9];
  const int lane = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const long long vc = (long long)bx * 32 + lane;
  const long long r0 = (long long)by * rows_per_block;
  long long r1 = r0 + rows_per_block;
  if (r1 > m) r1 = m;
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  if (vc < nvec) {
    const __nv_bfloat16* base = x + vc * 8;
    long long r = r0 + rl;
    for (; r + 24 < r1; r += 32) {
      uint4 u[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) u[i] = __ldg(reinterpret_cast<const uint4*>(base + (r + 8 * i) * ld));
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t w[4] = {u[i].x, u[i].y, u[i].z, u[i].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 t = unpack_bf16x2(w[e]);
          acc[2 * e] += t.x;
          acc[2 * e + 1] += t.y;
        }
      }
    }
    for (; r < r1; r += 8) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(base + r * ld));
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 t = unpack_bf16x2(w[e]);
        acc[2 * e] += t.x;
        acc[2 * e + 1] += t.y;
      }
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) red[rl][lane][e] = acc[e];
  __syncthreads();
  const int c = threadIdx.x;  // column within the block's 256-column group
  float t = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) t += red[w][c >> 3][c & 7];
  const long long col = (long long)blockIdx.x * 256 + c;
  if (col < nvec * 8) atomicAdd(out + col, t);
}

__global__ void __launch_bounds__(256)
colsum_vec_kernel(const __nv_bfloat16* __restrict__ x, long long m, long long nvec, long long ld,
                  float* __restrict__ out, long long rows_per_block) {
  grid_dep_wait();
  colsum_vec_body(x, m, nvec, ld, out, rows_per_block, blockIdx.x, blockIdx.y);
}

// Several column sums in one launch (blockIdx.z = problem; blocks outside a problem's own grid exit at once).
struct ColsumGroup {
  const __nv_bfloat16* x[8];
  float* out[8];
  long long m[8], nvec[8], ld[8], rows_pb[8];
  unsigned bx[8], by[8];
};
__global__ void __launch_bounds__(256) colsum_vec_group_kernel(const ColsumGroup g) {
  grid_dep_wait();
  const int z = blockIdx.z;
  if (blockIdx.x >= g.bx[z] || blockIdx.y >= g.by[z]) return;
  colsum_vec_body(g.x[z], g.m[z], g.nvec[z], g.ld[z], g.out[z], g.rows_pb[z], blockIdx.x, blockIdx.y);
}

// out = dropout(x (+ x2)): 8 elements per thread (two mask hashes), 16-byte accesses when aligned.
__global__ void __launch_bounds__(256)
dropout_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ x2, __nv_bfloat16* out, long long n,
               float rate, const unsigned long long* __restrict__ seed_p, uint32_t site) {
  grid_dep_wait();
  const long long i0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (i0 >= n) return;
  const uint64_t seed = *seed_p;
  const uint32_t thr = dropout_threshold(rate);
  const float inv_keep = 1.f / (1.f - rate);
  const bool vec = i0 + 8 <= n && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out) |
                                    reinterpret_cast<uintptr_t>(x2)) & 15) == 0;
  float f[8];
  if (vec) {
    const uint4 u = *reinterpret_cast<const uint4*>(x + i0);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 t = unpack_bf16x2(w[e]);
      f[2 * e] = t.x;
      f[2 * e + 1] = t.y;
    }
    if (x2) {
      const uint4 u2 = *reinterpret_cast<const uint4*>(x2 + i0);
      const uint32_t w2[4] = {u2.x, u2.y, u2.z, u2.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 t = unpack_bf16x2(w2[e]);
        f[2 * e] += t.x;
        f[2 * e + 1] += t.y;
      }
    }
  } else {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      f[e] = i0 + e < n ? __bfloat162float(x[i0 + e]) : 0.f;
      if (x2 && i0 + e < n) f[e] += __bfloat162float(x2[i0 + e]);
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) f[e] *= dropout_mul(seed, site, (uint64_t)(i0 + e), thr, inv_keep);
  if (vec) {
    uint4 o;
    o.x = pack_bf16x2(f[0], f[1]);
    o.y = pack_bf16x2(f[2], f[3]);
    o.z = pack_bf16x2(f[4], f[5]);
    o.w = pack_bf16x2(f[6], f[7]);
    *reinterpret_cast<uint4*>(out + i0) = o;
  } else {
#pragma unroll
    for (int e = 0; e < 8; ++e)
      if (i0 + e < n) out[i0 + e] = __float2bfloat16(f[e]);
  }
}

__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ x, long long n, float* __restrict__ out) {
  grid_dep_wait();
  __shared__ float red[8];
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = x[i];
    acc += v * v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    atomicAdd(out, t);
  }
}

// util.gumbel_noise (utils/util.py:189-195) added to the step logits (search.py:143-145): x += -log(-log(u + eps) + eps),
// u uniform in [0, 1) with 24 random bits per element, a pure function of (*seed, site, element index) like the dropout
// masks (one splitmix64 hash per two elements), so a replayed decode graph draws fresh noise when the host bumps *seed.
__global__ void __launch_bounds__(256)
gumbel_add_kernel(float* __restrict__ x, long long n, float eps, const unsigned long long* __restrict__ seed_p,
                  uint32_t site) {
  grid_dep_wait();
  const long long i0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 2;
  if (i0 >= n) return;
  const uint64_t h = dropout_hash4(*seed_p, site ^ 0x6A09E667u, (uint64_t)(i0 >> 1));
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    if (i0 + e < n) {
      const float u = (float)((uint32_t)(h >> (32 * e)) >> 8) * (1.f / 16777216.f);
      x[i0 + e] += -logf(-logf(u + eps) + eps);
    }
  }
}

// tf.train.AdamOptimizer: lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t) (host, passed in scalars[0]);
// m <- b1 m + (1-b1) g; v <- b2 v + (1-b2) g^2; p <- p - lr_t * m / (sqrt(v) + eps).
__global__ void __launch_bounds__(256)
adam_tf_kernel(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v, const float* __restrict__ g,
               __nv_bfloat16* __restrict__ pb, long long n, float b1, float b2, float eps, float lr_t, float gscale,
               const float* __restrict__ clip_scale, float* __restrict__ norms) {
  grid_dep_wait();
  const float gs = clip_scale ? gscale * clip_scale[0] : gscale;
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (norms) {
    // tf.global_norm of the (averaged, unclipped) gradients and of the pre-update parameters
    // (utils/cycle.py:94-95), fused into the pass that reads both anyway
    float sg = 0.f, sp = 0.f;
    for (long long j = i; j < n && j < i + 4; ++j) {
      const float gr = g[j] * gscale;
      sg += gr * gr;
      sp += p[j] * p[j];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sg += __shfl_xor_sync(0xffffffffu, sg, o);
      sp += __shfl_xor_sync(0xffffffffu, sp, o);
    }
    __shared__ float rg[8], rp[8];
    if ((threadIdx.x & 31) == 0) {
      rg[threadIdx.x >> 5] = sg;
      rp[threadIdx.x >> 5] = sp;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      float a = 0.f, b = 0.f;
      for (int w = 0; w < 8; ++w) {
        a += rg[w];
        b += rp[w];
      }
      atomicAdd(norms, a);
      atomicAdd(norms + 1, b);
    }
  }
  if (i + 3 < n) {
    float4 pp = *reinterpret_cast<float4*>(p + i), mm = *reinterpret_cast<float4*>(m + i),
           vv = *reinterpret_cast<float4*>(v + i);
    const float4 gg = *reinterpret_cast<const float4*>(g + i);
    float* P = &pp.x; float* M = &mm.x; float* V = &vv.x; const float* G = &gg.x;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float gr = G[e] * gs;
      M[e] = b1 * M[e] + (1.f - b1) * gr;
      V[e] = b2 * V[e] + (1.f - b2) * gr * gr;
      P[e] -= lr_t * M[e] / (sqrtf(V[e]) + eps);
    }
    *reinterpret_cast<float4*>(p + i) = pp;
    *reinterpret_cast<float4*>(m + i) = mm;
    *reinterpret_cast<float4*>(v + i) = vv;
    if (pb) {
      uint2 o;
      o.x = pack_bf16x2(pp.x, pp.y);
      o.y = pack_bf16x2(pp.z, pp.w);
      *reinterpret_cast<uint2*>(pb + i) = o;
    }
  } else {
    for (long long j = i; j < n; ++j) {
      const float gr = g[j] * gs;
      m[j] = b1 * m[j] + (1.f - b1) * gr;
      v[j] = b2 * v[j] + (1.f - b2) * gr * gr;
      p[j] -= lr_t * m[j] / (sqrtf(v[j]) + eps);
      if (pb) pb[j] = __float2bfloat16(p[j]);
    }
  }
}

// Same update, kIt float4 groups per thread: 8x fewer blocks, so the two norm atomics per block stop mattering.
constexpr int kAdamIt = 8;
__global__ void __launch_bounds__(256)
adam_tf_kernel_v2(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v, const float* __restrict__ g,
                  __nv_bfloat16* __restrict__ pb, long long n, float b1, float b2, float eps, float lr_t, float gscale,
                  const float* __restrict__ clip_scale, float* __restrict__ norms) {
  grid_dep_wait();
  __shared__ float rg[8], rp[8];
  const float gs = clip_scale ? gscale * clip_scale[0] : gscale;
  float sg = 0.f, sp = 0.f;
#pragma unroll 2
  for (int it = 0; it < kAdamIt; ++it) {
    const long long i = (((long long)blockIdx.x * kAdamIt + it) * 256 + threadIdx.x) * 4;
    if (i + 3 < n) {
      float4 pp = *reinterpret_cast<float4*>(p + i), mm = *reinterpret_cast<float4*>(m + i),
             vv = *reinterpret_cast<float4*>(v + i);
      const float4 gg = *reinterpret_cast<const float4*>(g + i);
      float* P = &pp.x; float* M = &mm.x; float* V = &vv.x; const float* G = &gg.x;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float gu = G[e] * gscale;
        sg += gu * gu;
        sp += P[e] * P[e];
        const float gr = G[e] * gs;
        M[e] = b1 * M[e] + (1.f - b1) * gr;
        V[e] = b2 * V[e] + (1.f - b2) * gr * gr;
        P[e] -= lr_t * M[e] / (sqrtf(V[e]) + eps);
      }
      *reinterpret_cast<float4*>(p + i) = pp;
      *reinterpret_cast<float4*>(m + i) = mm;
      *reinterpret_cast<float4*>(v + i) = vv;
      if (pb) {
        uint2 o;
        o.x = pack_bf16x2(pp.x, pp.y);
        o.y = pack_bf16x2(pp.z, pp.w);
        *reinterpret_cast<uint2*>(pb + i) = o;
      }
    } else {
      for (long long j = i; j < n; ++j) {
        const float gu = g[j] * gscale;
        sg += gu * gu;
        sp += p[j] * p[j];
        const float gr = g[j] * gs;
        m[j] = b1 * m[j] + (1.f - b1) * gr;
        v[j] = b2 * v[j] + (1.f - b2) * gr * gr;
        p[j] -= lr_t * m[j] / (sqrtf(v[j]) + eps);
        if (pb) pb[j] = __float2bfloat16(p[j]);
      }
    }
  }
  if (norms) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sg += __shfl_xor_sync(0xffffffffu, sg, o);
      sp += __shfl_xor_sync(0xffffffffu, sp, o);
    }
    if ((threadIdx.x & 31) == 0) {
      rg[threadIdx.x >> 5] = sg;
      rp[threadIdx.x >> 5] = sp;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      float a = 0.f, b = 0.f;
      for (int w = 0; w < 8; ++w) {
        a += rg[w];
        b += rp[w];
      }
      atomicAdd(norms, a);
      atomicAdd(norms + 1, b);
    }
  }
}

__global__ void gather_rows_kernel(const uint4* __restrict__ src, const int32_t* __restrict__ index,
                                   uint4* __restrict__ dst, long long rows, long long vec_per_row,
                                   long long pitch_vec) {
  grid_dep_wait();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * vec_per_row) return;
  const long long r = i / vec_per_row, c = i % vec_per_row;
  dst[r * pitch_vec + c] = src[(long long)index[r] * pitch_vec + c];
}

}  // namespace zb

using namespace zb;
#define ST(s) reinterpret_cast<cudaStream_t>(s)

extern "C" int zb_cast_f32_bf16(const float* src, void* dst, int64_t n, zb_stream_t stream) {
  ZB_REQUIRE(src && dst && n >= 0, "zb_cast_f32_bf16: bad args");
  if (n == 0) return ZB_OK;
  const long long blocks = ((n + 3) / 4 + 255) / 256;
  ZB_LAUNCH(cast_f32_bf16_kernel, (unsigned)blocks, 256, 0, ST(stream), src, (__nv_bfloat16*)dst, n);
  return check_launch("zb_cast_f32_bf16");
}
extern "C" int zb_cast_bf16_f32(const void* src, float* dst, int64_t n, zb_stream_t stream) {
  ZB_REQUIRE(src && dst && n >= 0, "zb_cast_bf16_f32: bad args");
  if (n == 0) return ZB_OK;
  ZB_LAUNCH(cast_bf16_f32_kernel, (unsigned)((n + 255) / 256), 256, 0, ST(stream), (const __nv_bfloat16*)src, dst, n);
  return check_launch("zb_cast_bf16_f32");
}
extern "C" int zb_dropout(const void* x, const void* x2, void* out, int64_t n, float rate, const uint64_t* seed,
                          uint32_t site, zb_stream_t stream) {
  ZB_REQUIRE(x && out && seed && n >= 0, "zb_dropout: null pointer");
  ZB_REQUIRE(rate >= 0.f && rate < 1.f, "zb_dropout: rate must be in [0, 1) (got %g)", (double)rate);
  if (n == 0) return ZB_OK;
  const long long threads = (n + 7) / 8;
  ZB_LAUNCH(dropout_kernel, (unsigned)((threads + 255) / 256), 256, 0, ST(stream), (const __nv_bfloat16*)x,
            (const __nv_bfloat16*)x2, (__nv_bfloat16*)out, (long long)n, rate,
            reinterpret_cast<const unsigned long long*>(seed), site);
  return check_launch("zb_dropout");
}
extern "C" int zb_colsum(const void* x, int64_t m, int64_t n, int64_t ld, float* out, zb_stream_t stream) {
  ZB_REQUIRE(x && out && m >= 0 && n > 0 && ld % 2 == 0, "zb_colsum: bad args");
  if (m == 0) return ZB_OK;
  if (n % 8 == 0 && ld % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
    const unsigned bx = (unsigned)((n / 8 + 31) / 32);
    long long by = (4ll * num_sms() + bx - 1) / bx;
    if (by > (m + 31) / 32) by = (m + 31) / 32;
    if (by < 1) by = 1;
    const long long rows_pb = ((m + by - 1) / by + 7) / 8 * 8;
    by = (m + rows_pb - 1) / rows_pb;
    ZB_LAUNCH(colsum_vec_kernel, dim3(bx, (unsigned)by), 256, 0, ST(stream), (const __nv_bfloat16*)x, m, n / 8, ld, out, rows_pb);
    return check_launch("zb_colsum");
  }
  const unsigned gx = (unsigned)((n + 63) / 64);
  long long gy = (2ll * num_sms() + gx - 1) / gx;
  if (gy > (m + 63) / 64) gy = (m + 63) / 64;
  if (gy < 1) gy = 1;
  const long long rpb = (m + gy - 1) / gy;
  ZB_LAUNCH(colsum_kernel, dim3(gx, (unsigned)gy), 256, 0, ST(stream), (const __nv_bfloat16*)x, m, n, ld, out, rpb);
  return check_launch("zb_colsum");
}
extern "C" int zb_colsum_grouped(const zb_colsum_args* problems, int32_t count, zb_stream_t stream) {
  ZB_REQUIRE(count >= 0 && (count == 0 || problems), "zb_colsum_grouped: null pointer");
  int done = 0;
  while (done < count) {
    int n = count - done < 8 ? count - done : 8;
    bool vec = n > 1;
    for (int i = 0; i < n && vec; ++i) {
      const zb_colsum_args& a = problems[done + i];
      vec = a.x && a.out && a.m > 0 && a.n > 0 && a.n % 8 == 0 && a.ld % 8 == 0 &&
            (reinterpret_cast<uintptr_t>(a.x) & 15) == 0;
    }
    if (!vec) {
      for (int i = 0; i < n; ++i) {
        const zb_colsum_args& a = problems[done + i];
        int rc = zb_colsum(a.x, a.m, a.n, a.ld, a.out, stream);
        if (rc) return rc;
      }
    } else {
      ColsumGroup g;
      unsigned gx = 1, gy = 1;
      for (int i = 0; i < n; ++i) {
        const zb_colsum_args& a = problems[done + i];
        const unsigned bx = (unsigned)((a.n / 8 + 31) / 32);
        long long by = (4ll * num_sms() / n + bx - 1) / bx;  // the group shares the machine
        if (by > (a.m + 31) / 32) by = (a.m + 31) / 32;
        if (by < 1) by = 1;
        const long long rows_pb = ((a.m + by - 1) / by + 7) / 8 * 8;
        by = (a.m + rows_pb - 1) / rows_pb;
        g.x[i] = (const __nv_bfloat16*)a.x; g.out[i] = a.out; g.m[i] = a.m; g.nvec[i] = a.n / 8; g.ld[i] = a.ld;
        g.rows_pb[i] = rows_pb; g.bx[i] = bx; g.by[i] = (unsigned)by;
        gx = bx > gx ? bx : gx;
        gy = (unsigned)by > gy ? (unsigned)by : gy;
      }
      for (int i = n; i < 8; ++i) {
        g.x[i] = g.x[0]; g.out[i] = g.out[0]; g.m[i] = 0; g.nvec[i] = 0; g.ld[i] = 0; g.rows_pb[i] = 8;
        g.bx[i] = 0; g.by[i] = 0;
      }
      ZB_LAUNCH(colsum_vec_group_kernel, dim3(gx, gy, (unsigned)n), 256, 0, ST(stream), g);
      int rc = check_launch("zb_colsum_grouped");
      if (rc) return rc;
    }
    done += n;
  }
  return ZB_OK;
}
extern "C" int zb_sumsq(const float* x, int64_t n, float* out, zb_stream_t stream) {
  ZB_REQUIRE(x && out && n >= 0, "zb_sumsq: bad args");
  if (n == 0) return ZB_OK;
  long long blocks = (n + 256 * 16 - 1) / (256 * 16);
  if (blocks > 4ll * num_sms()) blocks = 4ll * num_sms();
  ZB_LAUNCH(sumsq_kernel, (unsigned)blocks, 256, 0, ST(stream), x, n, out);
  return check_launch("zb_sumsq");
}
extern "C" int zb_gumbel_add(float* x, int64_t n, float eps, const uint64_t* seed, uint32_t site, zb_stream_t stream) {
  ZB_REQUIRE(x && seed && n >= 0 && eps > 0.f, "zb_gumbel_add: bad args");
  if (n == 0) return ZB_OK;
  ZB_LAUNCH(gumbel_add_kernel, (unsigned)((n + 511) / 512), 256, 0, ST(stream), x, (long long)n, eps,
            reinterpret_cast<const unsigned long long*>(seed), site);
  return check_launch("zb_gumbel_add");
}
extern "C" int zb_adam_tf(const zb_adam_args* a, zb_stream_t stream) {
  ZB_REQUIRE(a && a->param && a->m && a->v && a->grad && a->n >= 0, "zb_adam_tf: bad args");
  if (a->n == 0) return ZB_OK;
  const long long per_block = 256ll * 4 * kAdamIt;
  const long long blocks = (a->n + per_block - 1) / per_block;
  ZB_LAUNCH(adam_tf_kernel_v2, (unsigned)blocks, 256, 0, ST(stream), a->param, a->m, a->v, a->grad, (__nv_bfloat16*)a->param_bf16,
                                                          a->n, a->beta1, a->beta2, a->eps, a->lr_t, a->grad_scale,
                                                          a->clip_scale, a->norms);
  return check_launch("zb_adam_tf");
}
extern "C" int zb_gather_rows(const void* src, const int32_t* index, void* dst, int64_t rows, int64_t row_bytes,
                              int64_t pitch_bytes, zb_stream_t stream) {
  ZB_REQUIRE(src && index && dst && rows >= 0 && row_bytes > 0 && row_bytes % 16 == 0 && pitch_bytes % 16 == 0 &&
                 pitch_bytes >= row_bytes,
             "zb_gather_rows: row_bytes / pitch_bytes must be multiples of 16");
  if (rows == 0) return ZB_OK;
  const long long vec = row_bytes / 16, total = rows * vec;
  ZB_LAUNCH(gather_rows_kernel, (unsigned)((total + 255) / 256), 256, 0, ST(stream), (const uint4*)src, index, (uint4*)dst, rows,
                                                                             vec, pitch_bytes / 16);
  return check_launch("zb_gather_rows");
}
