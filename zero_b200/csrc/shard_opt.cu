// shard_opt.cu — K10: gradient aggregation fused with the optimizer step over NVLink / NVSwitch peer memory.
//
// The reference averages the towers' gradients variable by variable on one device and then runs Adam on every
// variable (utils/parallel.py:134-208, main.py:42-43, :178-181).  With one process per GPU and the flat arenas of
// ParamStore, the same step is ONE kernel per rank over that rank's 1/N shard of the arena:
//   reduce-scatter : g = sum_r grad_r[i], taken inside the NVSwitch (multimem.ld_reduce on the multicast mapping of the
//                    symmetric gradient arena) or from the ranks' unicast mappings (N loads in flight per element);
//   Adam           : TF semantics, on the shard's master / m / v only — 1/N of the 2.3 GB the replicated update moves;
//   all-gather     : the refreshed bf16 compute copy of the shard is stored into EVERY rank's mirror arena
//                    (multimem.st, or one store per rank) — bf16, half the bytes of the fp32 gradients.
// Cross-rank ordering is the caller's: a barrier before (every rank's backward has finished) and after (every copy has
// landed, every read of this rank's gradients is done).  The kernel ends with a system-scope fence.
#include "zb_common.h"
#include "zb_ptx.cuh"

namespace zb {

struct ShardK {
  long long lo, n;
  int world, rank, sources, flags;
  const float* grad_mc;
  const float* grad_peer[ZB_SHARD_MAX_WORLD];
  float* param; float* m; float* v;
  const uint8_t* wide;
  float* param_mc;
  float* param_peer[ZB_SHARD_MAX_WORLD];
  __nv_bfloat16* mirror_mc;
  __nv_bfloat16* mirror_peer[ZB_SHARD_MAX_WORLD];
  float* grad_out;
  float b1, b2, eps, lr_t, gscale;
  const float* clip_scale;
  float* norms;
  float* norm_parts_peer[ZB_SHARD_MAX_WORLD];
  unsigned int* done_counter;
};

// peer memory is read / written with system-scope accesses: never served from this SM's (non-coherent) L1
__device__ __forceinline__ float4 ld_sys_f4(const float* p) {
  float4 r;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ void st_sys_b128(void* p, const uint4 v) {
  asm volatile("st.relaxed.sys.global.v4.b32 [%0], {%1, %2, %3, %4};"
               :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void st_sys_f32(float* p, float v) {
  asm volatile("st.relaxed.sys.global.f32 [%0], %1;" :: "l"(p), "f"(v) : "memory");
}
// multicast mapping: one load returns the sum of the element over every GPU bound to the multicast object (reduced in
// the switch), one store lands in every GPU's copy
__device__ __forceinline__ float4 mc_ld_reduce_f4(const float* p) {
  float4 r;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ void mc_st_b128(void* p, const uint4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
               :: "l"(p), "f"(__uint_as_float(v.x)), "f"(__uint_as_float(v.y)), "f"(__uint_as_float(v.z)),
                  "f"(__uint_as_float(v.w)) : "memory");
}

__device__ __forceinline__ void adam4(float4& pp, float4& mm, float4& vv, const float4 gg, const ShardK& k,
                                      const float gs, float& sg, float& sp) {
  float* P = &pp.x; float* M = &mm.x; float* V = &vv.x; const float* G = &gg.x;
#pragma unroll
  for (int e = 0; e < 4; ++e) {   // same expression order as adam_tf_kernel_v2 (elementwise.cu)
    const float gu = G[e] * k.gscale;
    sg += gu * gu;
    sp += P[e] * P[e];
    const float gr = G[e] * gs;
    M[e] = k.b1 * M[e] + (1.f - k.b1) * gr;
    V[e] = k.b2 * V[e] + (1.f - k.b2) * gr * gr;
    P[e] -= k.lr_t * M[e] / (sqrtf(V[e]) + k.eps);
  }
}

// One unit = 8 consecutive elements = two 16-byte gradient vectors in, one 16-byte bf16 vector out per rank.
__global__ void __launch_bounds__(256)
shard_adam_kernel(const __grid_constant__ ShardK k) {
  grid_dep_wait();
  __shared__ float rg[8], rp[8];
  __shared__ bool last_cta;
  const float gs = k.clip_scale ? k.gscale * k.clip_scale[0] : k.gscale;
  const bool update = (k.flags & ZB_SHARD_UPDATE) != 0;
  float sg = 0.f, sp = 0.f;
  const long long units = k.n >> 3;
  for (long long u = (long long)blockIdx.x * 256 + threadIdx.x; u < units; u += (long long)gridDim.x * 256) {
    const long long i = k.lo + (u << 3);
    float4 g0, g1;
    if (k.grad_mc) {
      g0 = mc_ld_reduce_f4(k.grad_mc + i);
      g1 = mc_ld_reduce_f4(k.grad_mc + i + 4);
    } else {
      g0 = make_float4(0.f, 0.f, 0.f, 0.f);
      g1 = g0;
#pragma unroll 4
      for (int r = 0; r < k.sources; ++r) {   // rank order: the one owner of an element fixes the summation order
        const float4 a = ld_sys_f4(k.grad_peer[r] + i), b = ld_sys_f4(k.grad_peer[r] + i + 4);
        g0.x += a.x; g0.y += a.y; g0.z += a.z; g0.w += a.w;
        g1.x += b.x; g1.y += b.y; g1.z += b.z; g1.w += b.w;
      }
    }
    if (k.flags & ZB_SHARD_STORE_GRAD) {
      *reinterpret_cast<float4*>(k.grad_out + i) = g0;
      *reinterpret_cast<float4*>(k.grad_out + i + 4) = g1;
    }
    if (update) {
      float4 p0 = *reinterpret_cast<float4*>(k.param + i), p1 = *reinterpret_cast<float4*>(k.param + i + 4);
      float4 m0 = *reinterpret_cast<float4*>(k.m + i), m1 = *reinterpret_cast<float4*>(k.m + i + 4);
      float4 v0 = *reinterpret_cast<float4*>(k.v + i), v1 = *reinterpret_cast<float4*>(k.v + i + 4);
      adam4(p0, m0, v0, g0, k, gs, sg, sp);
      adam4(p1, m1, v1, g1, k, gs, sg, sp);
      *reinterpret_cast<float4*>(k.param + i) = p0; *reinterpret_cast<float4*>(k.param + i + 4) = p1;
      *reinterpret_cast<float4*>(k.m + i) = m0;     *reinterpret_cast<float4*>(k.m + i + 4) = m1;
      *reinterpret_cast<float4*>(k.v + i) = v0;     *reinterpret_cast<float4*>(k.v + i + 4) = v1;
      uint4 o;
      o.x = pack_bf16x2(p0.x, p0.y); o.y = pack_bf16x2(p0.z, p0.w);
      o.z = pack_bf16x2(p1.x, p1.y); o.w = pack_bf16x2(p1.z, p1.w);
      if (k.mirror_mc) {
        mc_st_b128(k.mirror_mc + i, o);
      } else {
#pragma unroll 4
        for (int r = 0; r < k.world; ++r) st_sys_b128(k.mirror_peer[r] + i, o);
      }
      if (k.wide && k.wide[i >> 6]) {   // a 1-D variable: the forward pass reads its fp32 master on every rank
        const uint4 w0 = *reinterpret_cast<uint4*>(&p0), w1 = *reinterpret_cast<uint4*>(&p1);
        if (k.param_mc) {
          mc_st_b128(k.param_mc + i, w0);
          mc_st_b128(k.param_mc + i + 4, w1);
        } else {
          for (int r = 0; r < k.world; ++r) {
            if (r == k.rank) continue;
            st_sys_b128(k.param_peer[r] + i, w0);
            st_sys_b128(k.param_peer[r] + i + 4, w1);
          }
        }
      }
    } else {
      const float* G0 = &g0.x; const float* G1 = &g1.x;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float a = G0[e] * k.gscale, b = G1[e] * k.gscale;
        sg += a * a + b * b;
      }
    }
  }
  if (k.norms) {
    if (!(k.flags & ZB_SHARD_NORM_G)) sg = 0.f;
    if (!(k.flags & ZB_SHARD_NORM_P)) sp = 0.f;
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
      if (k.flags & ZB_SHARD_NORM_G) atomicAdd(k.norms, a);
      if (k.flags & ZB_SHARD_NORM_P) atomicAdd(k.norms + 1, b);
      bool last = false;
      if (k.done_counter) {
        __threadfence();
        last = atomicAdd(k.done_counter, 1u) == gridDim.x - 1;
      }
      last_cta = last;
    }
    __syncthreads();
    if (last_cta && threadIdx.x < 2 * k.world) {
      // this rank's totals into row `rank` of every rank's table (thread t: rank t / 2, column t % 2)
      __threadfence();
      const int r = threadIdx.x >> 1, c = threadIdx.x & 1;
      const float val = atomicAdd(k.norms + c, 0.f);   // L2-coherent read of the finished sum
      if (k.norm_parts_peer[r]) st_sys_f32(k.norm_parts_peer[r] + 2 * k.rank + c, val);
    }
    if (last_cta && threadIdx.x == 0) *k.done_counter = 0u;
  }
  __threadfence_system();   // peer stores ordered before the kernel's completion is observed through the barrier
}

}  // namespace zb

using namespace zb;

extern "C" int zb_shard_adam(const zb_shard_adam_args* a, zb_stream_t stream) {
  ZB_REQUIRE(a && a->n >= 0 && a->lo >= 0 && (a->lo & 7) == 0 && (a->n & 7) == 0,
             "zb_shard_adam: lo and n must be non-negative multiples of 8 elements");
  ZB_REQUIRE(a->world >= 1 && a->world <= ZB_SHARD_MAX_WORLD && a->rank >= 0 && a->rank < a->world,
             "zb_shard_adam: world must be 1..%d and rank inside it", ZB_SHARD_MAX_WORLD);
  ZB_REQUIRE(a->grad_mc || (a->grad_sources >= 1 && a->grad_sources <= a->world),
             "zb_shard_adam: grad_sources must be 1..world");
  const bool update = (a->flags & ZB_SHARD_UPDATE) != 0;
  ZB_REQUIRE(!update || (a->param && a->m && a->v), "zb_shard_adam: ZB_SHARD_UPDATE needs param / m / v");
  ZB_REQUIRE(!(a->flags & ZB_SHARD_STORE_GRAD) || a->grad_out, "zb_shard_adam: ZB_SHARD_STORE_GRAD needs grad_out");
  ZB_REQUIRE(!(a->flags & (ZB_SHARD_NORM_G | ZB_SHARD_NORM_P)) || a->norms, "zb_shard_adam: norm flags need norms");
  ShardK k = {};
  k.lo = a->lo; k.n = a->n; k.world = a->world; k.rank = a->rank; k.flags = a->flags;
  k.sources = a->grad_sources;
  k.grad_mc = a->grad_mc;
  k.mirror_mc = (__nv_bfloat16*)a->mirror_mc;
  bool parts = false;
  for (int r = 0; r < a->world; ++r) {
    ZB_REQUIRE(a->grad_mc || r >= a->grad_sources || a->grad_peer[r],
               "zb_shard_adam: grad_peer[%d] is null and there is no grad_mc", r);
    ZB_REQUIRE(!update || a->mirror_mc || a->mirror_peer[r],
               "zb_shard_adam: mirror_peer[%d] is null and there is no mirror_mc", r);
    k.grad_peer[r] = a->grad_peer[r];
    k.mirror_peer[r] = (__nv_bfloat16*)a->mirror_peer[r];
    k.norm_parts_peer[r] = a->norm_parts_peer[r];
    parts = parts || a->norm_parts_peer[r];
  }
  ZB_REQUIRE(!parts || (a->norms && a->done_counter), "zb_shard_adam: norm_parts_peer needs norms and done_counter");
  k.param = a->param; k.m = a->m; k.v = a->v;
  k.wide = a->wide_mask;
  k.param_mc = a->param_mc;
  for (int r = 0; r < a->world; ++r) {
    ZB_REQUIRE(!update || !a->wide_mask || a->param_mc || r == a->rank || a->param_peer[r],
               "zb_shard_adam: wide_mask needs param_mc or param_peer[%d]", r);
    k.param_peer[r] = a->param_peer[r];
  }
  k.grad_out = a->grad_out;
  k.b1 = a->beta1; k.b2 = a->beta2; k.eps = a->eps; k.lr_t = a->lr_t; k.gscale = a->grad_scale;
  k.clip_scale = a->clip_scale;
  k.norms = a->norms;
  k.done_counter = parts ? a->done_counter : nullptr;
  // an empty shard still takes part in the norm exchange (one CTA, no loop iterations)
  const long long units = a->n >> 3;
  long long blocks = (units + 2 * 256 - 1) / (2 * 256);   // ~2 units per thread: 6 + 2N vector loads in flight each
  const long long cap = 8ll * num_sms();
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  if (a->n == 0 && !parts) return ZB_OK;
  ZB_LAUNCH(shard_adam_kernel, (unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream), k);
  return check_launch("zb_shard_adam");
}
