// gemm_tcgen05.cu — K1: persistent, warp-specialised bf16 GEMM on the 5th-gen tensor cores.
//
//   D[m,n] (+)= alpha * sum_k A(m,k) B(n,k)   (+bias[n]) (relu) (* (mask[m,n] > 0))
//
// Replaces every tf.matmul on Zero's Transformer path: func.linear (func.py:49,59), the tied-softmax
// projection (models/transformer.py:194) and, as dgrad/wgrad, their tf.gradients (main.py:28).
//
// Structure (one CTA per SM, 320 threads):
//   warp 0      TMA producer  : cp.async.bulk.tensor -> 128B-swizzled smem ring (A 128x64, B BNx64 bf16)
//   warp 1      MMA issuer    : one elected lane issues tcgen05.mma (128 x BN x 16), accumulators in TMEM,
//                               tcgen05.commit releases smem slots / publishes the accumulator
//   warps 2..9  epilogue      : tcgen05.ld (32 lanes x 32 columns) -> alpha/bias/relu/mask -> 16 B stores
//                               or fp32 red.add (split-K / gradient accumulation)
// Two TMEM accumulator stages (2 x BN columns) let the MMAs of tile i+1 overlap the epilogue of tile i.
// Operands may be K-major ([rows][k]) or MN-major ([k][rows]); both map onto SWIZZLE_128B canonical layouts,
// so forward (x W), dgrad (dy W^T) and wgrad (x^T dy) all read the tensors where they lie — no transposes.
#include <stdlib.h>

#include "zb_common.h"
#include "zb_ptx.cuh"

namespace zb {

bool gemm_skinny_wanted(const zb_gemm_args* a);
int gemm_skinny_launch(const zb_gemm_args* a, cudaStream_t st);
bool gemm2_wanted(const zb_gemm_args* a);
int gemm2_launch(const zb_gemm_args* a, cudaStream_t st);
int gemm2_launch_group(const zb_gemm_args* args, int count, cudaStream_t st);

struct GemmKParams {
  int M, N, K;
  int mt, nt, splits, kb_per_split, kb_total;
  int gn;  // rasterisation: tiles are walked in column panels of `gn` n-blocks, n fastest inside a panel, so the
           // CTAs running at the same time share A row-panels and B column-panels in L2 instead of re-reading them
  void* d;
  long long ldd;
  const float* bias;
  const __nv_bfloat16* mask;
  long long ldmask;
  float alpha;
  int flags;
  int d_f32;
  unsigned long long* trace;  // debug (ZB_GEMM_TRACE=1): CTA 0 records globaltimer at 7 points of its life
};

constexpr int kBM = 128;
constexpr int kBK = 64;
constexpr int kGemmThreads = 320;  // TMA warp + MMA warp + 8 epilogue warps
constexpr int kSmemBudget = 200 * 1024;

template <int BN, int BM = kBM>
struct GemmCfg {
  static constexpr int A_BYTES = BM * kBK * 2;
  static constexpr int B_BYTES = BN * kBK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES_RAW = kSmemBudget / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int TMEM_COLS = 2 * BN;  // two accumulator stages; 128 / 256 / 512: powers of two
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

// linear tile index -> (m block, n block, k split)
__device__ __forceinline__ void tile_coords(const GemmKParams& p, int tile, int& m_blk, int& n_blk, int& ks) {
  const int per_split = p.mt * p.nt;
  ks = tile / per_split;
  const int t = tile - ks * per_split;
  const int panel_tiles = p.mt * p.gn;
  const int pn = t / panel_tiles;
  const int r = t - pn * panel_tiles;
  const int w = min(p.gn, p.nt - pn * p.gn);
  m_blk = r / w;
  n_blk = pn * p.gn + (r - m_blk * w);
}

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}

// BM = 128 is the production tile.  BM = 64 (opt-in, ZB_GEMM_BM64=1, for problems of a few hundred rows) halves the
// A bytes a CTA pulls per k-block and doubles the number of CTAs a skinny problem spreads over; tcgen05.mma with M = 64
// leaves the accumulator in the lower 16 lanes of each 32-lane TMEM quadrant (row r -> lane 32 * (r / 16) + r % 16),
// so an epilogue warp of quadrant q owns rows 16 q .. 16 q + 15 in its lanes 0..15.
template <int BN, bool A_MN, bool B_MN, int BM = kBM>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_tcgen05(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                  const GemmKParams p) {
  static_assert(BM == 64 || BM == 128, "tcgen05.mma cta_group::1: M is 64 or 128");
  using Cfg = GemmCfg<BN, BM>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.mt * p.nt * p.splits;
  const bool tracing = p.trace != nullptr && blockIdx.x == 0;
  if (tracing && threadIdx.x == 0) p.trace[0] = globaltimer_ns();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 8);  // one arrival per epilogue warp
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Everything above (barrier init, TMEM allocation, descriptor prefetch) overlapped the tail of the previous
  // kernel when launched with programmatic stream serialisation; from here on we touch its results.
  grid_dep_wait();
  if (tracing && threadIdx.x == 0) p.trace[1] = globaltimer_ns();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int m_blk, n_blk, ks;
        tile_coords(p, tile, m_blk, n_blk, ks);
        const int kb0 = ks * p.kb_per_split;
        const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
        const int m0 = m_blk * BM, n0 = n_blk * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + Cfg::A_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          const int k0 = kb * kBK;
          if constexpr (!A_MN) {
            tma_load_2d(sa, &tma_a, &full_bar[stage], k0, m0);  // box {64 k, BM m}
          } else {
#pragma unroll
            for (int c = 0; c < BM / 64; ++c)  // box {64 m, 64 k}
              tma_load_2d(sa + c * 8192, &tma_a, &full_bar[stage], m0 + 64 * c, k0);
          }
          if constexpr (!B_MN) {
            tma_load_2d(sb, &tma_b, &full_bar[stage], k0, n0);  // box {64 k, BN n}
          } else {
#pragma unroll
            for (int c = 0; c < BN / 64; ++c)  // box {64 n, 64 k}
              tma_load_2d(sb + c * 8192, &tma_b, &full_bar[stage], n0 + 64 * c, k0);
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, A_MN ? 1u : 0u, B_MN ? 1u : 0u);
      // canonical SWIZZLE_128B layouts: K-major  -> 8-row groups every 1024 B (SBO), k-step = 32 B
      //                                 MN-major -> 64-wide MN atoms every BK*128 B (LBO), 8-k groups every
      //                                             1024 B (SBO), k-step (16 rows of 128 B) = 2048 B
      constexpr uint32_t A_LBO = A_MN ? kBK * 128 : 0, A_KSTEP = A_MN ? 2048 : 32;
      constexpr uint32_t B_LBO = B_MN ? kBK * 128 : 0, B_KSTEP = B_MN ? 2048 : 32;
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int m_blk, n_blk, ks;
        tile_coords(p, tile, m_blk, n_blk, ks);
        const int kb0 = ks * p.kb_per_split;
        const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          if (tracing && tile == blockIdx.x && kb == kb0) p.trace[2] = globaltimer_ns();
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint32_t sb = sa + Cfg::A_BYTES;
#pragma unroll
          for (int kk = 0; kk < kBK / 16; ++kk) {
            const uint64_t da = umma_smem_desc(sa + kk * A_KSTEP, A_LBO, 1024);
            const uint64_t db = umma_smem_desc(sb + kk * B_KSTEP, B_LBO, 1024);
            umma_bf16_ss(tmem_d, da, db, idesc, (kb > kb0 || kk > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // smem slot free once these MMAs retire
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tfull_bar[acc]);  // accumulator complete
        if (tracing && tile == blockIdx.x) p.trace[3] = globaltimer_ns();
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..9)
    // Per tile each warp owns 32 accumulator rows (its TMEM lane quadrant) and walks the BN columns in chunks of
    // 32.  Two register sets ping-pong so the tcgen05.ld of chunk c+1 is in flight while chunk c is converted
    // and stored; the per-column bias is fetched once per chunk by one coalesced load (lane j <- bias[col0+j]),
    // issued before the TMEM wait, and broadcast with shuffles.
    const int quad = warp & 3;          // TMEM lane quadrant this warp may access
    const int half = (warp - 2) >> 2;   // which half of the tile's columns (two warps per quadrant)
    const int flags = p.flags;
    const float alpha = p.alpha;
    const bool d_f32 = p.d_f32 != 0;
    int acc = 0;
    uint32_t acc_phase = 0;

    auto finish_chunk = [&](uint32_t (&r)[32], int row, bool row_ok, int col0, float bias_lane, const uint4 (&mk)[4]) {
      if (col0 >= p.N) return;
      const bool full_cols = (col0 + 32 <= p.N);
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) * alpha;
      if (flags & ZB_EPI_BIAS) {
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += __shfl_sync(0xffffffffu, bias_lane, j);
      }
      if (flags & ZB_EPI_RELU) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
      }
      if (!row_ok) return;
      if (flags & ZB_EPI_RELU_MASK) {
        if (full_cols) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint32_t w[4] = {mk[q].x, mk[q].y, mk[q].z, mk[q].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 f = unpack_bf16x2(w[e]);
              if (!(f.x > 0.f)) v[q * 8 + e * 2] = 0.f;
              if (!(f.y > 0.f)) v[q * 8 + e * 2 + 1] = 0.f;
            }
          }
        } else {
          const __nv_bfloat16* mrow = p.mask + static_cast<long long>(row) * p.ldmask + col0;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (col0 + j < p.N && !(__bfloat162float(mrow[j]) > 0.f)) v[j] = 0.f;
        }
      }
      if (d_f32) {
        float* drow = reinterpret_cast<float*>(p.d) + static_cast<long long>(row) * p.ldd + col0;
        if (flags & ZB_EPI_ACCUM) {
          if (full_cols) {
#pragma unroll
            for (int q = 0; q < 8; ++q) red_add_v4(drow + q * 4, v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j < p.N) atomicAdd(drow + j, v[j]);
          }
        } else {
          if (full_cols) {
#pragma unroll
            for (int q = 0; q < 8; ++q)
              reinterpret_cast<float4*>(drow)[q] = make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j < p.N) drow[j] = v[j];
          }
        }
      } else {
        __nv_bfloat16* drow = reinterpret_cast<__nv_bfloat16*>(p.d) + static_cast<long long>(row) * p.ldd + col0;
        if (full_cols) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 o;
            o.x = pack_bf16x2(v[q * 8 + 0], v[q * 8 + 1]);
            o.y = pack_bf16x2(v[q * 8 + 2], v[q * 8 + 3]);
            o.z = pack_bf16x2(v[q * 8 + 4], v[q * 8 + 5]);
            o.w = pack_bf16x2(v[q * 8 + 6], v[q * 8 + 7]);
            reinterpret_cast<uint4*>(drow)[q] = o;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (col0 + j < p.N) drow[j] = __float2bfloat16(v[j]);
        }
      }
    };
    // operands of a chunk that come from global memory: issued early so their latency hides behind tcgen05.ld
    auto prefetch_chunk = [&](int row, bool row_ok, int col0, float& bias_lane, uint4 (&mk)[4]) {
      bias_lane = 0.f;
      if ((flags & ZB_EPI_BIAS) && col0 + lane < p.N) bias_lane = __ldg(p.bias + col0 + lane);
      if ((flags & ZB_EPI_RELU_MASK) && row_ok && col0 + 32 <= p.N) {
        const uint4* mrow = reinterpret_cast<const uint4*>(p.mask + static_cast<long long>(row) * p.ldmask + col0);
#pragma unroll
        for (int q = 0; q < 4; ++q) mk[q] = __ldg(mrow + q);
      }
    };

    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      int m_blk, n_blk, ks;
      tile_coords(p, tile, m_blk, n_blk, ks);
      const int m0 = m_blk * BM, n0 = n_blk * BN;
      const int row = BM == 128 ? m0 + quad * 32 + lane : m0 + quad * 16 + lane;
      const bool row_ok = BM == 128 ? row < p.M : (lane < 16 && row < p.M);
      float bias_a, bias_b;
      uint4 mk_a[4], mk_b[4];
      // this warp's column range: the two warps of a lane quadrant split the BN columns in halves
      constexpr int NCH = BN / 64;  // 32-column chunks per warp
      const int c0 = half * NCH;
      prefetch_chunk(row, row_ok, n0 + c0 * 32, bias_a, mk_a);
      mbar_wait(&tfull_bar[acc], acc_phase);
      if (tracing && tile == blockIdx.x && warp == 2 && lane == 0) p.trace[4] = globaltimer_ns();
      tc_fence_after();
      const uint32_t tbase = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BN + c0 * 32;
      uint32_t ra[32], rb[32];
      tmem_ld_32x32b_x32(tbase, ra);
#pragma unroll 1
      for (int c = 0; c < NCH; c += 2) {
        __syncwarp();
        tmem_ld_wait();  // ra (chunk c) has landed
        if (c + 1 < NCH) {
          tmem_ld_32x32b_x32(tbase + (c + 1) * 32, rb);
          prefetch_chunk(row, row_ok, n0 + (c0 + c + 1) * 32, bias_b, mk_b);
        }
        finish_chunk(ra, row, row_ok, n0 + (c0 + c) * 32, bias_a, mk_a);
        if (c + 1 < NCH) {
          __syncwarp();
          tmem_ld_wait();  // rb (chunk c + 1) has landed
          if (c + 2 < NCH) {
            tmem_ld_32x32b_x32(tbase + (c + 2) * 32, ra);
            prefetch_chunk(row, row_ok, n0 + (c0 + c + 2) * 32, bias_a, mk_a);
          }
          finish_chunk(rb, row, row_ok, n0 + (c0 + c + 1) * 32, bias_b, mk_b);
        }
      }
      // all TMEM reads of this accumulator stage are complete (every tcgen05.ld was waited on)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      if (tracing && tile == blockIdx.x && warp == 2 && lane == 0) p.trace[5] = globaltimer_ns();
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (tracing && threadIdx.x == 0) p.trace[6] = globaltimer_ns();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------ host
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

// 2-D bf16 tensor map: dim0 (contiguous) x dim1 with row pitch `ld` elements; box {64, box1}; SWIZZLE_128B.
int make_map(CUtensorMap* m, const void* ptr, uint64_t dim0, uint64_t dim1, uint64_t ld, uint32_t box1) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled unavailable");
    return ZB_ECUDA;
  }
  cuuint64_t dims[2] = {dim0, dim1};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {64, box1};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): dims %llu x %llu ld %llu", (int)r, (unsigned long long)dim0,
              (unsigned long long)dim1, (unsigned long long)ld);
    return ZB_ECUDA;
  }
  return ZB_OK;
}

// 3-D bf16 tensor map over a [batch, len, width] view (width contiguous, row pitch `ld`, batch pitch `bs` elements);
// box {64 channels, box1 positions, 1 batch element}; SWIZZLE_128B.  Positions past `len` read as zeros (the tile
// never runs into the next batch element), which is what the attention kernels rely on for ragged lengths.
int make_map_3d(CUtensorMap* m, const void* ptr, uint64_t width, uint64_t len, uint64_t batch, uint64_t ld, uint64_t bs,
                uint32_t box1) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled unavailable");
    return ZB_ECUDA;
  }
  cuuint64_t dims[3] = {width, len, batch};
  cuuint64_t strides[2] = {ld * 2, (batch > 1 ? bs : ld * len) * 2};
  cuuint32_t box[3] = {64, box1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (3-D) failed (%d): dims %llu x %llu x %llu ld %llu bs %llu", (int)r,
              (unsigned long long)width, (unsigned long long)len, (unsigned long long)batch, (unsigned long long)ld,
              (unsigned long long)bs);
    return ZB_ECUDA;
  }
  return ZB_OK;
}

// 4-D bf16 tensor map over a [batch, len, heads * 64] view seen as {64 channels, len, heads, batch} (row pitch `ld`,
// head pitch 64, batch pitch `bs` elements); box {64, box_rows, box_heads, 1}; SWIZZLE_128B.  A box lands in shared
// memory as [box_heads][box_rows][64 channels]: the stacked-heads operand layout of the attention kernels in ONE copy.
// Positions past `len` read as zeros and are not written by stores.
int make_map_heads(CUtensorMap* m, const void* ptr, uint64_t len, uint64_t heads, uint64_t batch, uint64_t ld,
                   uint64_t bs, uint32_t box_rows, uint32_t box_heads) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled unavailable");
    return ZB_ECUDA;
  }
  cuuint64_t dims[4] = {64, len, heads, batch};
  cuuint64_t strides[3] = {ld * 2, 128, (batch > 1 ? bs : ld * len) * 2};
  cuuint32_t box[4] = {64, box_rows, box_heads, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (heads) failed (%d): len %llu heads %llu batch %llu ld %llu bs %llu", (int)r,
              (unsigned long long)len, (unsigned long long)heads, (unsigned long long)batch, (unsigned long long)ld,
              (unsigned long long)bs);
    return ZB_ECUDA;
  }
  return ZB_OK;
}

// fp32 output tiles for the epilogue's TMA stores: box {32 floats = 128 B, box1 rows}, 128B swizzle
int make_map_f32(CUtensorMap* m, const void* ptr, uint64_t dim0, uint64_t dim1, uint64_t ld, uint32_t box1) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled unavailable");
    return ZB_ECUDA;
  }
  cuuint64_t dims[2] = {dim0, dim1};
  cuuint64_t strides[1] = {ld * 4};
  cuuint32_t box[2] = {32, box1};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (f32) failed (%d): dims %llu x %llu ld %llu", (int)r, (unsigned long long)dim0,
              (unsigned long long)dim1, (unsigned long long)ld);
    return ZB_ECUDA;
  }
  return ZB_OK;
}

template <int BN, bool A_MN, bool B_MN, int BM = kBM>
static int launch(const CUtensorMap& ta, const CUtensorMap& tb, const GemmKParams& p, int grid, cudaStream_t st) {
  auto kern = gemm_bf16_tcgen05<BN, A_MN, B_MN, BM>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<BN, BM>::SMEM_BYTES);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(smem=%d): %s", GemmCfg<BN, BM>::SMEM_BYTES, cudaGetErrorString(e));
      return ZB_ECUDA;
    }
    attr_done = true;
  }
  static const bool use_pdl = getenv("ZB_NO_PDL") == nullptr;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = GemmCfg<BN, BM>::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = use_pdl ? 1 : 0;
  cudaError_t le = cudaLaunchKernelEx(&cfg, kern, ta, tb, p);
  if (le != cudaSuccess) {
    set_error("zb_gemm launch: %s", cudaGetErrorString(le));
    return ZB_ECUDA;
  }
  note_path(ZB_PATH_GEMM_TCGEN05);
  return check_launch("zb_gemm");
}

template <int BN, int BM = kBM>
static int dispatch_layout(int a_mn, int b_mn, const CUtensorMap& ta, const CUtensorMap& tb, const GemmKParams& p,
                           int grid, cudaStream_t st) {
  if (!a_mn && !b_mn) return launch<BN, false, false, BM>(ta, tb, p, grid, st);
  if (!a_mn && b_mn) return launch<BN, false, true, BM>(ta, tb, p, grid, st);
  if (a_mn && !b_mn) return launch<BN, true, false, BM>(ta, tb, p, grid, st);
  return launch<BN, true, true, BM>(ta, tb, p, grid, st);
}

}  // namespace zb

static int validate_gemm(const zb_gemm_args* a) {
  using namespace zb;
  ZB_REQUIRE(a && a->a && a->b && a->d, "zb_gemm: null pointer");
  ZB_REQUIRE(a->m > 0 && a->n > 0 && a->k > 0, "zb_gemm: empty problem m=%lld n=%lld k=%lld", (long long)a->m,
             (long long)a->n, (long long)a->k);
  ZB_REQUIRE(a->lda % 8 == 0 && a->ldb % 8 == 0, "zb_gemm: lda/ldb must be multiples of 8 elements (TMA 16 B pitch)");
  ZB_REQUIRE((reinterpret_cast<uintptr_t>(a->a) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->b) & 15) == 0,
             "zb_gemm: operands must be 16-byte aligned");
  ZB_REQUIRE(a->d_dtype == ZB_BF16 || a->d_dtype == ZB_F32, "zb_gemm: bad d_dtype");
  const bool accum = a->flags & ZB_EPI_ACCUM;
  ZB_REQUIRE(!accum || a->d_dtype == ZB_F32, "zb_gemm: ZB_EPI_ACCUM needs an fp32 destination");
  ZB_REQUIRE(!(a->flags & ZB_EPI_BIAS) || a->bias, "zb_gemm: ZB_EPI_BIAS without bias");
  ZB_REQUIRE(!(a->flags & ZB_EPI_RELU_MASK) || (a->mask && a->ldmask % 8 == 0 &&
                                                (reinterpret_cast<uintptr_t>(a->mask) & 15) == 0),
             "zb_gemm: ZB_EPI_RELU_MASK needs an aligned mask");
  const int esz = a->d_dtype == ZB_F32 ? 4 : 2;
  ZB_REQUIRE((a->ldd * esz) % 16 == 0 && (reinterpret_cast<uintptr_t>(a->d) & 15) == 0,
             "zb_gemm: destination must be 16-byte aligned with a 16-byte pitch");
  ZB_REQUIRE(a->m < (1ll << 31) && a->n < (1ll << 31) && a->k < (1ll << 31), "zb_gemm: dimension too large");
  return ZB_OK;
}

// Same result as zb_gemm on every problem (outputs must not overlap); accumulate-into-fp32 problems with MN-major
// operands (the weight gradients of one layer) share ONE persistent launch.
extern "C" int zb_gemm_grouped(const zb_gemm_args* problems, int32_t count, zb_stream_t stream) {
  using namespace zb;
  ZB_REQUIRE(count >= 0 && (count == 0 || problems), "zb_gemm_grouped: null pointer");
  for (int i = 0; i < count; ++i) {
    int rc = validate_gemm(&problems[i]);
    if (rc) return rc;
  }
  int done = 0;
  while (done < count) {
    int n = count - done < 8 ? count - done : 8;
    int rc = gemm2_launch_group(problems + done, n, reinterpret_cast<cudaStream_t>(stream));
    if (rc < 0) return rc;
    if (rc > 0) {  // not groupable: one by one
      for (int i = 0; i < n; ++i) {
        rc = zb_gemm(&problems[done + i], stream);
        if (rc) return rc;
      }
    }
    done += n;
  }
  return ZB_OK;
}

extern "C" int zb_gemm(const zb_gemm_args* a, zb_stream_t stream) {
  using namespace zb;
  int vrc = validate_gemm(a);
  if (vrc) return vrc;
  const bool accum = a->flags & ZB_EPI_ACCUM;
  // a few hundred rows (the cached decode step): latency-bound, small-tile kernel of gemm_skinny.cu
  if (gemm_skinny_wanted(a)) return gemm_skinny_launch(a, reinterpret_cast<cudaStream_t>(stream));
  // CTA-pair kernel (tcgen05 cta_group::2, gemm2_tcgen05.cu) for everything with at least one 256 x 128 pair tile
  if (gemm2_wanted(a)) return gemm2_launch(a, reinterpret_cast<cudaStream_t>(stream));

  GemmKParams p;
  p.M = (int)a->m; p.N = (int)a->n; p.K = (int)a->k;
  p.d = a->d; p.ldd = a->ldd; p.bias = a->bias;
  p.mask = reinterpret_cast<const __nv_bfloat16*>(a->mask); p.ldmask = a->ldmask;
  p.alpha = a->alpha; p.flags = a->flags; p.d_f32 = a->d_dtype == ZB_F32;
  p.trace = nullptr;
  static const bool trace_on = getenv("ZB_GEMM_TRACE") != nullptr;
  static unsigned long long* trace_buf = nullptr;
  if (trace_on) {
    if (!trace_buf) cudaMalloc(&trace_buf, 8 * sizeof(unsigned long long));
    p.trace = trace_buf;
  }
  p.kb_total = (p.K + kBK - 1) / kBK;
  // 64-row tiles (opt-in): a problem of a few hundred rows spreads over twice the CTAs, each pulling half the A bytes
  const char* bm64_env = getenv("ZB_GEMM_BM64");   // per call: the parity test flips it inside one process
  const bool bm64 = bm64_env && bm64_env[0] == '1' && !accum && p.M <= 512;
  const int bm = bm64 ? 64 : kBM;
  p.mt = (p.M + bm - 1) / bm;

  const int sms = num_sms_compute();
  // Tile width: the widest BN whose tile count still fills the machine (wave quantisation dominates at the
  // reference's 4096-token batches); 256 when there is plenty of work.
  int bn = 256;
  auto tiles_for = [&](int b) { return (long long)p.mt * ((p.N + b - 1) / b); };
  if (p.N <= 64 || bm64) bn = 64;
  else if (p.N <= 128) bn = 128;
  if (!accum) {  // with an accumulating epilogue split-K fills the machine, so keep the widest (most L2-frugal) tile
    if (bn == 256 && tiles_for(256) < 2ll * sms) bn = 128;
    if (bn == 128 && tiles_for(128) < sms && p.N > 64) bn = 64;
  }
  p.nt = (p.N + bn - 1) / bn;
  // panel width minimising the bytes the ~#SM concurrently running tiles pull through L2:
  // (sms / gn) A row-blocks of 128 rows + gn B column-blocks of bn rows  ->  gn ~ sqrt(sms * 128 / bn)
  p.gn = bn == 256 ? 8 : (bn == 128 ? 12 : 16);
  if (p.gn > p.nt) p.gn = p.nt;

  int splits = a->split_k;
  const long long tiles = (long long)p.mt * p.nt;
  if (splits <= 0) {
    splits = 1;
    if (accum && tiles < sms) {
      // pick the split count whose tiles * splits fills whole waves of SMs best (ties -> fewer splits), keeping
      // at least 4 k-blocks (256 k) per split so the pipeline has something to stream
      int max_splits = p.kb_total / 4;
      if (max_splits < 1) max_splits = 1;
      if (max_splits > 64) max_splits = 64;
      double best = 0.0;
      for (int s = 1; s <= max_splits; ++s) {
        const long long work = tiles * s;
        const long long waves = (work + sms - 1) / sms;
        const double eff = (double)work / (double)(waves * sms);
        if (eff > best + 0.02) {
          best = eff;
          splits = s;
        }
      }
    }
  }
  ZB_REQUIRE(splits == 1 || accum, "zb_gemm: split_k > 1 requires ZB_EPI_ACCUM");
  if (splits > p.kb_total) splits = p.kb_total;
  p.kb_per_split = (p.kb_total + splits - 1) / splits;
  p.splits = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;

  CUtensorMap ta, tb;
  int rc;
  const bool a_mn = a->a_layout == ZB_MN_MAJOR, b_mn = a->b_layout == ZB_MN_MAJOR;
  if (!a_mn) rc = make_map(&ta, a->a, p.K, p.M, a->lda, bm);
  else rc = make_map(&ta, a->a, p.M, p.K, a->lda, kBK);
  if (rc) return rc;
  if (!b_mn) rc = make_map(&tb, a->b, p.K, p.N, a->ldb, bn);
  else rc = make_map(&tb, a->b, p.N, p.K, a->ldb, kBK);
  if (rc) return rc;

  const long long total = tiles * p.splits;
  const int grid = (int)(total < sms ? total : sms);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int lrc;
  if (bm64) {
    note_path(ZB_PATH_GEMM_BM64);
    return dispatch_layout<64, 64>(a_mn, b_mn, ta, tb, p, grid, st);
  }
  switch (bn) {
    case 64: lrc = dispatch_layout<64>(a_mn, b_mn, ta, tb, p, grid, st); break;
    case 128: lrc = dispatch_layout<128>(a_mn, b_mn, ta, tb, p, grid, st); break;
    default: lrc = dispatch_layout<256>(a_mn, b_mn, ta, tb, p, grid, st); break;
  }
  if (trace_on && lrc == ZB_OK) {
    unsigned long long h[8];
    cudaStreamSynchronize(st);
    cudaMemcpy(h, trace_buf, sizeof(h), cudaMemcpyDeviceToHost);
    fprintf(stderr,
            "[zb_gemm trace] m=%d n=%d k=%d bn=%d splits=%d grid=%d | ns since entry: dep_wait %llu, first data %llu, "
            "tile0 mma issued %llu, tile0 acc ready %llu, tile0 epilogue done %llu, exit %llu\n",
            p.M, p.N, p.K, bn, p.splits, grid, h[1] - h[0], h[2] - h[0], h[3] - h[0], h[4] - h[0], h[5] - h[0], h[6] - h[0]);
  }
  return lrc;
}
