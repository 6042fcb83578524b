// gemm_skinny.cu — K1 for the cached decode step: y = x W (+ b) (relu) with only a few hundred rows of x
// (batch * beam = 256 at BASELINE configs[2]; func.linear, func.py:14-65, called once per sublayer per step).
//
// At 256 rows a projection is 0.13 - 0.5 GFLOP against 0.5 - 2 MB of weights: neither the tensor pipe nor HBM is
// the limit, the launch-to-first-MMA latency is.  The persistent tcgen05 kernel pays tensor-map fetches, TMEM
// allocation, barrier set-up and a 2-CTA cluster launch for at most eight 256 x 128 tiles (measured 10 us per
// projection, 30 projections per decode step).  This kernel is the small-problem complement: 64 x 64 output tiles so
// that even a 256 x 512 projection spreads over 32 SMs, operands streamed through a 4-stage cp.async ring of
// XOR-swizzled 64 x 64 bf16 tiles, warp-level mma.sync m16n8k16 (fp32 accumulation) — no set-up beyond the first
// cp.async.  A 256 x 512 x 512 projection is still only 32 tiles with a serial chain of 8 k-chunks each (measured:
// no faster than the tcgen05 kernel), so the k dimension is additionally split over a thread-block cluster of up to
// 8 CTAs (grid z): every CTA streams 1 - 4 chunks, all in flight at once, and the partial accumulators are summed by
// the cluster's rank 0 through distributed shared memory in a fixed order (deterministic), which then applies the
// epilogue.  zb_gemm routes a problem here only when m <= kSkinnyMaxM; everything token-sized stays on tcgen05.
#include <cooperative_groups.h>
#include <stdlib.h>

#include "zb_common.h"
#include "zb_ptx.cuh"

namespace zb {
namespace sk {

constexpr int NT = 128, BM = 64, BN = 64, BK = 64, STAGES = 4;
constexpr int kSkinnyMaxM = 384;

struct Params {
  const __nv_bfloat16 *a, *b;
  void* d;
  const float* bias;
  int M, N, K;
  long long lda, ldb, ldd;
  float alpha;
  int relu, d_f32;
  int split;   // cluster size along grid z = number of k slices
};

// [64][64] bf16 tile, 16-byte chunks XOR-swizzled by (row & 7): conflict-free for ldmatrix
__device__ __forceinline__ int swz(int r, int c) { return r * 64 + ((((c >> 3) ^ (r & 7)) << 3) | (c & 7)); }

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool pred) {
  const uint32_t s = smem_u32(smem);
  const int sz = pred ? 16 : 0;  // src-size 0 -> zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// rows [0, rows_valid) x 64 columns of a row-major view -> swizzled tile (rows beyond rows_valid are zero-filled)
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

// B_MN: B stored [k][n] (weights [in, out]); else [n][k]
template <bool B_MN>
__global__ void __launch_bounds__(NT) gemm_skinny_kernel(const Params p) {
  grid_dep_wait();
  extern __shared__ __align__(128) __nv_bfloat16 sk_smem[];
  __nv_bfloat16* sA = sk_smem;                          // STAGES x [64 m][64 k]
  __nv_bfloat16* sB = sk_smem + STAGES * BM * BK;       // STAGES x ([64 k][64 n] or [64 n][64 k])
  const int n0 = blockIdx.x * BN, m0 = blockIdx.y * BM;
  const int rows_valid = min(BM, p.M - m0);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  namespace cg = cooperative_groups;
  const int S = p.split;
  const int z = S > 1 ? (int)cg::this_cluster().block_rank() : 0;
  const int nk = p.K / BK / S;                           // chunks of this CTA's k slice
  const __nv_bfloat16* ag = p.a + (long long)m0 * p.lda + (long long)z * nk * BK;
  const __nv_bfloat16* bg = B_MN ? p.b + n0 + (long long)z * nk * BK * p.ldb
                                 : p.b + (long long)n0 * p.ldb + (long long)z * nk * BK;

  auto issue = [&](int kc) {
    const int s = kc % STAGES;
    load_tile(sA + s * BM * BK, ag + kc * BK, p.lda, rows_valid);
    if (B_MN) load_tile(sB + s * BK * BN, bg + (long long)kc * BK * p.ldb, p.ldb, BK);
    else load_tile(sB + s * BK * BN, bg + kc * BK, p.ldb, BN);
  };
#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < nk) issue(s);
    cp_async_commit();
  }
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int kc = 0; kc < nk; ++kc) {
    cp_async_wait<STAGES - 2>();   // chunk kc has landed (for this thread's copies)
    __syncthreads();               // ... for everyone's; and everyone is done with the stage refilled below
    if (kc + STAGES - 1 < nk) issue(kc + STAGES - 1);
    cp_async_commit();
    const __nv_bfloat16* ta = sA + (kc % STAGES) * BM * BK;
    const __nv_bfloat16* tb = sB + (kc % STAGES) * BK * BN;
    uint32_t a[4][4];
    {
      const int r = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) ldsm_x4(a[ks], ta + swz(r, ks * 16 + (lane >> 4) * 8));
    }
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t b[4];
        if (B_MN) {
          const int kr = ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
          const int nc = np * 16 + (lane >> 4) * 8;
          ldsm_x4_t(b, tb + swz(kr, nc));
        } else {
          const int n = np * 16 + (lane & 7) + (lane >> 4) * 8;
          const int kcol = ks * 16 + ((lane >> 3) & 1) * 8;
          ldsm_x4(b, tb + swz(n, kcol));
        }
        mma16816(acc[2 * np], a[ks], b[0], b[1]);
        mma16816(acc[2 * np + 1], a[ks], b[2], b[3]);
      }
    }
  }
  if (S > 1) {
    // k slices -> rank 0: [32 values][128 threads] fp32 in the (now idle) operand ring, read over DSMEM
    cg::cluster_group cluster = cg::this_cluster();
    float* red = reinterpret_cast<float*>(sk_smem);
    __syncthreads();
    if (z != 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) red[(i * 4 + j) * NT + threadIdx.x] = acc[i][j];
    }
    cluster.sync();
    if (z == 0) {
      for (int r = 1; r < S; ++r) {
        const float* rr = cluster.map_shared_rank(red, r);
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] += rr[(i * 4 + j) * NT + threadIdx.x];
      }
    }
    cluster.sync();   // the other ranks' shared memory must outlive rank 0's reads
    if (z != 0) return;
  }
  // epilogue: alpha, bias, relu; the thread holds rows g and g + 8 of its warp's 16, columns nt * 8 + 2t, +1
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    const int col = n0 + nt * 8 + 2 * t;
    float b0 = 0.f, b1 = 0.f;
    if (p.bias) {
      b0 = p.bias[col];
      b1 = p.bias[col + 1];
    }
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int row = warp * 16 + g + half * 8;
      if (row >= rows_valid) continue;
      float v0 = acc[nt][2 * half] * p.alpha + b0, v1 = acc[nt][2 * half + 1] * p.alpha + b1;
      if (p.relu) {
        v0 = fmaxf(v0, 0.f);
        v1 = fmaxf(v1, 0.f);
      }
      const long long off = (long long)(m0 + row) * p.ldd + col;
      if (p.d_f32) *reinterpret_cast<float2*>(reinterpret_cast<float*>(p.d) + off) = make_float2(v0, v1);
      else *reinterpret_cast<uint32_t*>(reinterpret_cast<__nv_bfloat16*>(p.d) + off) = pack_bf16x2(v0, v1);
    }
  }
}

}  // namespace sk

// Opt-in (ZB_SKINNY_GEMM=1): in the decode step of BASELINE configs[2] this kernel measured 10.2 us per projection
// (10.9 us with the cluster k split) against 9.1 us for the tcgen05 kernel it was meant to undercut — all three sit on
// the same floor, so the k-loop is not what bounds a 256-row projection (profiles/r01_decode_ab_v2.jsonl).  Kept,
// parity-tested (tests/test_kernels_gpu.py::test_gemm_skinny_rows sets the switch), for the follow-up that finds the floor.
static bool skinny_enabled() {
  const char* e = getenv("ZB_SKINNY_GEMM");   // read per call: the parity test flips it inside one process
  return e && e[0] == '1';
}

// K-major A, whole 64-wide tiles in n and k, plain / bias / relu epilogue, a few hundred rows at most
bool gemm_skinny_wanted(const zb_gemm_args* a) {
  if (!skinny_enabled()) return false;
  if (a->m > sk::kSkinnyMaxM || a->n > 4096 || (a->n % sk::BN) || (a->k % sk::BK)) return false;
  if (a->a_layout != ZB_K_MAJOR) return false;
  if (a->flags & ~(ZB_EPI_BIAS | ZB_EPI_RELU)) return false;
  if (a->split_k > 1) return false;
  return true;
}

int gemm_skinny_launch(const zb_gemm_args* a, cudaStream_t st) {
  sk::Params p;
  p.a = reinterpret_cast<const __nv_bfloat16*>(a->a);
  p.b = reinterpret_cast<const __nv_bfloat16*>(a->b);
  p.d = a->d;
  p.bias = (a->flags & ZB_EPI_BIAS) ? a->bias : nullptr;
  p.M = (int)a->m; p.N = (int)a->n; p.K = (int)a->k;
  p.lda = a->lda; p.ldb = a->ldb; p.ldd = a->ldd;
  p.alpha = a->alpha;
  p.relu = (a->flags & ZB_EPI_RELU) ? 1 : 0;
  p.d_f32 = a->d_dtype == ZB_F32;
  constexpr size_t smem = (size_t)sk::STAGES * (sk::BM * sk::BK + sk::BK * sk::BN) * sizeof(__nv_bfloat16);
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(sk::gemm_skinny_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(sk::gemm_skinny_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr_set = true;
  }
  // k slices: double while the grid does not fill the machine or a CTA would still chain more than 4 chunks
  const int tiles = (p.N / sk::BN) * ((p.M + sk::BM - 1) / sk::BM), nk = p.K / sk::BK;
  const char* split_env = getenv("ZB_SKINNY_SPLIT");
  const bool split_on = !(split_env && split_env[0] == '0');
  int S = 1;
  while (split_on && S < 8 && nk % (2 * S) == 0 && (tiles * S < num_sms() || nk / S > 4)) S *= 2;
  p.split = S;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(p.N / sk::BN, (p.M + sk::BM - 1) / sk::BM, S);
  cfg.blockDim = dim3(sk::NT);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (S > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 1;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = S;
    ++na;
  }
  if (pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  cudaError_t le = a->b_layout == ZB_MN_MAJOR ? cudaLaunchKernelEx(&cfg, sk::gemm_skinny_kernel<true>, p)
                                              : cudaLaunchKernelEx(&cfg, sk::gemm_skinny_kernel<false>, p);
  if (le != cudaSuccess) {
    set_error("zb_gemm (skinny, %d k slices) launch: %s", S, cudaGetErrorString(le));
    return ZB_ECUDA;
  }
  note_path(ZB_PATH_GEMM_SKINNY);
  return check_launch("zb_gemm(skinny)");
}

}  // namespace zb
