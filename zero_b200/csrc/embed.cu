// embed.cu — K5: fused embedding gather * sqrt(d) + shared bias + sinusoidal timing signal (+ decoder shift),
// and its backward (scatter-add into the fp32 table gradient, bias gradient).
// Replaces tf.gather / bias_add / pad-shift / func.add_timing_signal
// (models/transformer.py:29-31, 104-117; func.py:341-369).  One warp per token row, 16 B vectors.
#include "zb_common.h"
#include "zb_ptx.cuh"

namespace zb {

constexpr int kEmbWarps = 8;

// func.add_timing_signal: channel c < dim/2 -> sin(pos * w_c), else cos(pos * w_{c - dim/2}),
// w_i = exp(-i * ln(1e4) / (dim/2 - 1)); odd trailing channel is zero.
__device__ __forceinline__ float timing_value(float pos, int c, int dim) {
  const int nts = dim >> 1;
  if (c >= 2 * nts) return 0.f;
  const int i = c < nts ? c : c - nts;
  const float inc = 9.210340371976184f / (float)(nts - 1);  // ln(1e4)
  const float ang = pos * expf(-(float)i * inc);
  return c < nts ? sinf(ang) : cosf(ang);
}

__global__ void __launch_bounds__(kEmbWarps * 32)
embed_fwd_kernel(const int32_t* __restrict__ ids, const __nv_bfloat16* __restrict__ table,
                 const float* __restrict__ bias, __nv_bfloat16* __restrict__ out, int batch, int len, int dim,
                 int vocab, int shift, int zero_if_all_pad, int time, float mult) {
  grid_dep_wait();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  bool all_pad = false;
  if (zero_if_all_pad) {
    int any = 0;
    for (int i = threadIdx.x; i < batch * len; i += blockDim.x) any |= (ids[i] != 0);
    all_pad = !__syncthreads_or(any);
  }
  const long long rows = (long long)batch * len;
  const int nvec = dim >> 3;
  for (long long row = (long long)blockIdx.x * kEmbWarps + warp; row < rows; row += (long long)gridDim.x * kEmbWarps) {
    const int l = (int)(row % len);
    const float pos = time >= 0 ? (float)time : (float)l;
    const bool zero_row = (shift && l < shift) || all_pad;
    int id = 0;
    if (!zero_row) {
      id = ids[row - shift];
      id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
    }
    for (int v = lane; v < nvec; v += 32) {
      float f[8];
      if (!zero_row) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(table + (long long)id * dim + v * 8));
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 t = unpack_bf16x2(w[e]);
          f[2 * e] = t.x * mult + __ldg(bias + v * 8 + 2 * e);
          f[2 * e + 1] = t.y * mult + __ldg(bias + v * 8 + 2 * e + 1);
        }
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = 0.f;
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] += timing_value(pos, v * 8 + e, dim);
      uint4 o;
      o.x = pack_bf16x2(f[0], f[1]);
      o.y = pack_bf16x2(f[2], f[3]);
      o.z = pack_bf16x2(f[4], f[5]);
      o.w = pack_bf16x2(f[6], f[7]);
      *reinterpret_cast<uint4*>(out + row * dim + v * 8) = o;
    }
  }
}

__global__ void __launch_bounds__(kEmbWarps * 32)
embed_bwd_kernel(const int32_t* __restrict__ ids, const __nv_bfloat16* __restrict__ d_out,
                 const __nv_bfloat16* __restrict__ d_out2,
                 float* __restrict__ d_table, float* __restrict__ d_bias, int batch, int len, int dim, int vocab,
                 int shift, float mult) {
  grid_dep_wait();
  extern __shared__ float red[];  // [kEmbWarps][dim]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long rows = (long long)batch * len;
  const int nvec = dim >> 3;
  for (int c = threadIdx.x; c < kEmbWarps * dim; c += blockDim.x) red[c] = 0.f;
  __syncthreads();
  for (long long row = (long long)blockIdx.x * kEmbWarps + warp; row < rows; row += (long long)gridDim.x * kEmbWarps) {
    const int l = (int)(row % len);
    if (shift && l < shift) continue;
    int id = ids[row - shift];
    id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
    for (int v = lane; v < nvec; v += 32) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(d_out + row * dim + v * 8));
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
      float f[8];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 t = unpack_bf16x2(w[e]);
        f[2 * e] = t.x;
        f[2 * e + 1] = t.y;
      }
      if (d_out2) {
        const uint4 u2 = __ldg(reinterpret_cast<const uint4*>(d_out2 + row * dim + v * 8));
        const uint32_t w2[4] = {u2.x, u2.y, u2.z, u2.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 t = unpack_bf16x2(w2[e]);
          f[2 * e] += t.x;
          f[2 * e + 1] += t.y;
        }
      }
      float* dst = d_table + (long long)id * dim + v * 8;
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(f[0] * mult), "f"(f[1] * mult),
                   "f"(f[2] * mult), "f"(f[3] * mult)
                   : "memory");
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4), "f"(f[4] * mult),
                   "f"(f[5] * mult), "f"(f[6] * mult), "f"(f[7] * mult)
                   : "memory");
#pragma unroll
      for (int e = 0; e < 8; ++e) red[warp * dim + v * 8 + e] += f[e];
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < dim; c += blockDim.x) {
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < kEmbWarps; ++w) a += red[w * dim + c];
    atomicAdd(d_bias + c, a);
  }
}

}  // namespace zb

extern "C" int zb_embed_fwd(const zb_embed_args* a, zb_stream_t stream) {
  using namespace zb;
  ZB_REQUIRE(a && a->ids && a->table && a->bias && a->out, "zb_embed_fwd: null pointer");
  ZB_REQUIRE(a->dim > 0 && a->dim % 8 == 0 && a->vocab > 0, "zb_embed_fwd: dim must be a multiple of 8");
  ZB_REQUIRE(a->shift == 0 || a->shift == 1, "zb_embed_fwd: shift must be 0 or 1");
  const long long rows = (long long)a->batch * a->len;
  if (rows == 0) return ZB_OK;
  long long blocks = (rows + kEmbWarps - 1) / kEmbWarps;
  const long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  ZB_LAUNCH(embed_fwd_kernel, (int)blocks, kEmbWarps * 32, 0, reinterpret_cast<cudaStream_t>(stream), 
      a->ids, (const __nv_bfloat16*)a->table, a->bias, (__nv_bfloat16*)a->out, a->batch, a->len, a->dim, a->vocab,
      a->shift, a->zero_if_all_pad, a->time, a->mult);
  return check_launch("zb_embed_fwd");
}

extern "C" int zb_embed_bwd(const zb_embed_args* a, zb_stream_t stream) {
  using namespace zb;
  ZB_REQUIRE(a && a->ids && a->d_out && a->d_table && a->d_bias, "zb_embed_bwd: null pointer");
  ZB_REQUIRE(a->dim > 0 && a->dim % 8 == 0 && a->dim <= 1536, "zb_embed_bwd: dim must be a multiple of 8, <= 1536");
  const long long rows = (long long)a->batch * a->len;
  if (rows == 0) return ZB_OK;
  long long blocks = (rows + kEmbWarps - 1) / kEmbWarps;
  const long long cap = (long long)num_sms() * 2;
  if (blocks > cap) blocks = cap;
  const size_t smem = (size_t)kEmbWarps * a->dim * sizeof(float);
  ZB_LAUNCH(embed_bwd_kernel, (int)blocks, kEmbWarps * 32, smem, reinterpret_cast<cudaStream_t>(stream), 
      a->ids, (const __nv_bfloat16*)a->d_out, (const __nv_bfloat16*)a->d_out2, a->d_table, a->d_bias, a->batch, a->len, a->dim, a->vocab, a->shift,
      a->mult);
  return check_launch("zb_embed_bwd");
}
