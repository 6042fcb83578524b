// attention_tc.cu — K2/K3 on the 5th-generation tensor cores: fused scaled-dot attention forward and backward for
// dh = 64 at ANY sequence length (func.dot_attention, func.py:218-256, and its gradient), with the reference's masks
// (func.attention_bias, func.py:372-388: additive -inf_value from key lengths / the causal index) and attention dropout
// (func.py:245).  tcgen05.mma with accumulators in TMEM, q / k / v / dO tiles by TMA (cp.async.bulk.tensor, 3-D maps
// over the [batch, len, heads * dh] views, read in place from the fused [tokens, 3d] projection buffer; positions past
// the sequence end arrive as zeros), logits and probabilities never touch HBM.
//
// Both kernels are PERSISTENT and warp-specialised (352 threads, one CTA per SM):
//   warp 0     TMA producer: runs up to two 128 x 128 blocks ahead of the math through a two-stage shared-memory ring
//              (4-D tensor maps {64 channels, position, head, batch}: one copy per operand, also for stacked heads)
//   warp 1     TMEM allocation + MMA issue (one elected thread)
//   warps 2-9  two threads per row of the block (forward: a query row; backward: a query row of S / dP, a key row of
//              dK / dV), each owning half of the row's columns: tcgen05.ld, masks, exp2, dropout, P / dS written to
//              shared memory in the canonical SWIZZLE_128B K-major layout a TMA load would produce (so they feed the
//              next MMA as the A operand); results staged as bf16 rows in shared memory
//   warp 10    TMA tile stores of the staged results (rows past the sequence end are clipped by the copy engine)
// In-kernel timelines (ZB_ATTN_TRACE=1, profiles/r02*_attn_trace.log) shaped this: one thread per row with 4 row warps
// spent 5 us per block in the row arithmetic and 3 us in 16-byte-per-lane global stores.
//
// A 128-row block is either 128 consecutive positions of ONE head, or — when lq, lk <= 64, the reference's 64-token
// training batches — the stacked 64 + 64 positions of TWO heads (h, h + 1): S = [Q_h; Q_h+1][K_h; K_h+1]^T fills one
// 128 x 128 tcgen05.mma, the off-diagonal 64 x 64 blocks are ignored and P is written with zeros there, so
// O = P [V_h; V_h+1] needs no special case.
//
//   forward   unit = (batch, head selector, 128-query block); loop over 128-key blocks with the online softmax.
//             S is double-buffered in TMEM (2 x 128 columns): the MMA of block n + 1 runs while the rows of block n
//             are exponentiated; O = P V of a block lands in 64 more columns and is folded into registers
//             (acc = acc * exp(m_acc - m) + O) after the NEXT block's row maxima are known.
//   backward  unit = (batch, head selector, 128-key block); loop over the 128-query blocks that see it.  Per block:
//             S = Q K^T and dP = dO V^T (2 x 128 columns), P = exp(S * scale + mask - lse), dS = P (dP - delta) per
//             row, then dQ = dS K (fresh), dK += dS^T Q, dV += P^T dO (accumulated in TMEM over the unit's query
//             blocks: no atomics, written once).  P and dS exist as ONE shared-memory image each, read through K-major
//             (rows = m) and MN-major (rows = k) descriptors.  dQ goes straight to its bf16 destination when a
//             (batch, head) has one key block, else it is reduced in an fp32 workspace (red.global.add.v4.f32) and
//             cast by a second tiny kernel.
#include <math.h>
#include <stdlib.h>

#include "zb_common.h"
#include "zb_ptx.cuh"

namespace zb {

int make_map_heads(CUtensorMap* m, const void* ptr, uint64_t len, uint64_t heads, uint64_t batch, uint64_t ld,
                   uint64_t bs, uint32_t box_rows, uint32_t box_heads);  // gemm_tcgen05.cu
int make_map(CUtensorMap* m, const void* ptr, uint64_t dim0, uint64_t dim1, uint64_t ld, uint32_t box1);

namespace fat {

constexpr int kRowWarps = 8;   // two threads per row of a 128-row block
constexpr int kThreads = 32 * (2 + kRowWarps + 1);  // warp 0 TMA loads, warp 1 MMA, warps 2-9 rows, warp 10 TMA stores
constexpr int kTile = 128 * 64;  // elements of a [128 rows][64 channels] bf16 tile (16 KB)

struct Params {
  const __nv_bfloat16 *o, *d_o;
  __nv_bfloat16 *out, *dq, *dk, *dv;
  float* dq32;  // fp32 [batch, lq, heads * 64] reduction workspace (backward, more than one key block)
  long long ldo, bso, lddo, bsdo, lddq, bsdq, lddk, bsdk, lddv, bsdv;
  int batch, heads, lq, lk, causal, q_offset;
  int relu;   // 1: ReLA (modules/rela.py:52-75): weights = relu(logits * keep), no normaliser
  int pair;   // 1: two heads stacked per 128-row block (lq, lk <= 64)
  int nq, nk; // 128-row query / key blocks per (batch, head selector); 1 in pair mode
  int hsel;   // head selectors per batch element: heads, or heads / 2 in pair mode
  int units;
  const int32_t* key_len;
  float scale, inf_value;
  float* lse;
  float drop_rate;
  uint32_t drop_site;
  const unsigned long long* drop_seed;
  unsigned long long* trace;  // debug (ZB_ATTN_TRACE=1): CTA 0 stamps globaltimer at fixed points of its first items
};
// trace slots: 0 entry, 1 setup done, 2 producer past the dependency wait, 3 exit; per item n < 4 at 8 + 8 n:
// +0 loads issued, +1 tiles landed (MMA warp), +2 first MMA batch committed, +3 row warp sees S, +4 row warp arrived on
// bar_p, +5 second MMA batch committed, +6 row warp sees O / dQ, +7 row warp done with the item
#define FAT_TRACE(slot)                                                       \
  do {                                                                        \
    if (p.trace != nullptr && blockIdx.x == 0) p.trace[slot] = globaltimer_ns(); \
  } while (0)
#define FAT_TRACE_ITEM(n, ev)                                   \
  do {                                                          \
    if ((n) < 4) FAT_TRACE(8 + 8 * (int)(n) + (ev));            \
  } while (0)

// The flat sequence of (unit, inner block) items of this CTA; every role walks the same sequence.
//   forward : outer = query block, inner = key block     backward: outer = key block, inner = query block
struct Walk {
  const Params& p;
  const bool fwd;
  int unit, b, hs, outer, inner, inner_begin, inner_end;
  __device__ __forceinline__ Walk(const Params& p_, bool fwd_) : p(p_), fwd(fwd_) {}
  __device__ __forceinline__ bool load() {
    while (unit < p.units) {
      const int nouter = fwd ? p.nq : p.nk;
      b = unit / (p.hsel * nouter);
      const int rem = unit - b * (p.hsel * nouter);
      hs = rem / nouter;
      outer = rem - hs * nouter;
      if (fwd) {
        inner_begin = 0;
        inner_end = p.nk;
        if (p.causal && !p.pair) inner_end = min(p.nk, (outer * 128 + 127 + p.q_offset) / 128 + 1);
      } else {
        inner_begin = 0;
        inner_end = p.nq;
        if (p.causal && !p.pair) inner_begin = max(0, (outer * 128 - p.q_offset) / 128);
      }
      inner = inner_begin;
      if (inner < inner_end) return true;
      unit += gridDim.x;
    }
    return false;
  }
  __device__ __forceinline__ bool start() {
    unit = blockIdx.x;
    return load();
  }
  __device__ __forceinline__ bool next() {
    if (++inner < inner_end) return true;
    unit += gridDim.x;
    return load();
  }
  __device__ __forceinline__ bool first() const { return inner == inner_begin; }
  __device__ __forceinline__ bool last() const { return inner + 1 == inner_end; }
  __device__ __forceinline__ int qblock() const { return fwd ? outer : inner; }
  __device__ __forceinline__ int kblock() const { return fwd ? inner : outer; }
};

// attention dropout multiplier (0 or 1 / keep) of weight (query i, key j) of (batch b, head h): the same pure function
// of (*seed, site, flat index into [batch, heads, lq, lk]) as in attention_mma.cu / attention_generic.cu
template <bool DROP>
struct Drop {
  uint64_t seed;
  uint32_t site, thr;
  float inv_keep;
  __device__ __forceinline__ Drop(const Params& p) {
    if (DROP) {
      seed = *p.drop_seed;
      site = p.drop_site;
      thr = dropout_threshold(p.drop_rate);
      inv_keep = 1.f / (1.f - p.drop_rate);
    }
  }
  __device__ __forceinline__ float mul(const Params& p, int b, int h, int i, int j) const {
    if (!DROP) return 1.f;
    const uint64_t idx = (((uint64_t)b * p.heads + h) * (uint64_t)p.lq + i) * (uint64_t)p.lk + j;
    return dropout_mul(seed, site, idx, thr, inv_keep);
  }
};

// 16-byte unit `u` (0..7) of row `row` inside a [rows][128 B] SWIZZLE_128B atom
__device__ __forceinline__ uint32_t swz_off(int row, int u) { return (uint32_t)row * 128u + (uint32_t)((u ^ (row & 7)) << 4); }

__device__ __forceinline__ void sts128(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
constexpr float kLog2e = 1.4426950408889634f, kLn2 = 0.6931471805599453f;

// 32 consecutive fp32 values (one tcgen05.ld chunk) -> bf16 -> the four 16-byte units [u0, u0 + 4) of `row`
__device__ __forceinline__ void store_chunk_bf16(uint32_t atom_saddr, int row, int u0, const float (&v)[32]) {
#pragma unroll
  for (int q = 0; q < 4; ++q)
    sts128(atom_saddr + swz_off(row, u0 + q), pack_bf16x2(v[8 * q + 0], v[8 * q + 1]),
           pack_bf16x2(v[8 * q + 2], v[8 * q + 3]), pack_bf16x2(v[8 * q + 4], v[8 * q + 5]),
           pack_bf16x2(v[8 * q + 6], v[8 * q + 7]));
}
__device__ __forceinline__ void store_chunk_zero(uint32_t atom_saddr, int row, int u0) {
#pragma unroll
  for (int q = 0; q < 4; ++q) sts128(atom_saddr + swz_off(row, u0 + q), 0u, 0u, 0u, 0u);
}
// 64 fp32 accumulator values of one row (two tcgen05.ld chunks) * mul -> bf16 -> row `row` of a [128][128 B] swizzled
// staging tile (the layout a TMA tile store reads)
__device__ __forceinline__ void stage_row_bf16(uint32_t tile_saddr, int row, const uint32_t (&ra)[32],
                                               const uint32_t (&rb)[32], float mul) {
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const uint32_t* s = c < 4 ? ra + 8 * c : rb + 8 * (c - 4);
    sts128(tile_saddr + swz_off(row, c), pack_bf16x2(__uint_as_float(s[0]) * mul, __uint_as_float(s[1]) * mul),
           pack_bf16x2(__uint_as_float(s[2]) * mul, __uint_as_float(s[3]) * mul),
           pack_bf16x2(__uint_as_float(s[4]) * mul, __uint_as_float(s[5]) * mul),
           pack_bf16x2(__uint_as_float(s[6]) * mul, __uint_as_float(s[7]) * mul));
  }
}

// Logits of one row, one 32-column chunk, in base-2 units: t = raw * sl2 (sl2 = scale * log2 e), minus inf2 where the
// key is masked (columns [nv, nb) of the chunk), -inf where the column is no key of this row at all (>= nb).
// The mode is warp-uniform:
//   kFast  every row of the warp has the whole chunk visible: no per-element mask work
//   kMid   the chunk is in bounds for every row and every row has a visible key elsewhere (in the block for the
//          forward's running maximum, in the sequence for the backward's log-sum-exp), so a masked logit's weight
//          2^(t - inf2 - m) is exactly 0 in fp32 and it cannot be the maximum: masked columns are simply dropped
//   kSlow  the reference's arithmetic literally (rows whose keys are all masked, ragged last block)
enum { kFast = 0, kMid = 1, kSlow = 2 };
__device__ __forceinline__ int chunk_mode(int nv, int nb, bool row_has_visible) {
  if (__all_sync(0xffffffffu, nv >= 32)) return kFast;
  if (__all_sync(0xffffffffu, nb >= 32 && row_has_visible)) return kMid;
  return kSlow;
}
__device__ __forceinline__ float chunk_max(const uint32_t (&r)[32], int mode, int nv, int nb, float sl2, float inf2) {
  float m[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};  // four independent chains: few warps per scheduler need ILP
  if (mode == kFast) {
#pragma unroll
    for (int jj = 0; jj < 32; ++jj) m[jj & 3] = fmaxf(m[jj & 3], __uint_as_float(r[jj]));
    return fmaxf(fmaxf(m[0], m[1]), fmaxf(m[2], m[3])) * sl2;  // sl2 > 0
  }
  if (mode == kMid) {
#pragma unroll
    for (int jj = 0; jj < 32; ++jj) m[jj & 3] = fmaxf(m[jj & 3], jj < nv ? __uint_as_float(r[jj]) : -INFINITY);
    return fmaxf(fmaxf(m[0], m[1]), fmaxf(m[2], m[3])) * sl2;
  }
#pragma unroll
  for (int jj = 0; jj < 32; ++jj) {
    float t = __uint_as_float(r[jj]) * sl2;
    t = jj < nv ? t : t - inf2;
    t = jj < nb ? t : -INFINITY;
    m[jj & 3] = fmaxf(m[jj & 3], t);
  }
  return fmaxf(fmaxf(m[0], m[1]), fmaxf(m[2], m[3]));
}
// e[jj] = 2^(t[jj] - off) of the same chunk; returns the sum of e
__device__ __forceinline__ float chunk_exp(const uint32_t (&r)[32], float (&e)[32], int mode, int nv, int nb, float sl2,
                                           float inf2, float off) {
  float l[4] = {0.f, 0.f, 0.f, 0.f};
  if (mode == kFast) {
#pragma unroll
    for (int jj = 0; jj < 32; ++jj) {
      e[jj] = ex2(fmaf(__uint_as_float(r[jj]), sl2, -off));
      l[jj & 3] += e[jj];
    }
  } else if (mode == kMid) {
#pragma unroll
    for (int jj = 0; jj < 32; ++jj) {
      const float x = ex2(fmaf(__uint_as_float(r[jj]), sl2, -off));
      e[jj] = jj < nv ? x : 0.f;
      l[jj & 3] += e[jj];
    }
  } else {
#pragma unroll
    for (int jj = 0; jj < 32; ++jj) {
      float t = __uint_as_float(r[jj]) * sl2;
      t = jj < nv ? t : t - inf2;
      e[jj] = jj < nb ? ex2(t - off) : 0.f;
      l[jj & 3] += e[jj];
    }
  }
  return (l[0] + l[1]) + (l[2] + l[3]);
}
// ReLA (modules/rela.py:65-70): e[jj] = relu(raw * scale) where the key is visible, 0 elsewhere (multiplicative 0 / 1
// mask); `g` (optional) = 1 where the weight is positive — the gate of its gradient
__device__ __forceinline__ void chunk_relu(const uint32_t (&r)[32], float (&e)[32], int nv, float scale) {
#pragma unroll
  for (int jj = 0; jj < 32; ++jj) e[jj] = jj < nv ? fmaxf(__uint_as_float(r[jj]) * scale, 0.f) : 0.f;
}
// barrier among the 8 row warps (named barrier 1; barrier 0 is __syncthreads)
__device__ __forceinline__ void row_warps_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// ================================================================================================ forward
struct FwdStage {
  __nv_bfloat16 q[kTile], k[kTile], v[kTile];
};
struct FwdSmem {
  FwdStage st[2];
  __nv_bfloat16 p[2][kTile];    // A operand of O = P V: two 64-key atoms of [128 rows][64 keys]
  __nv_bfloat16 ostage[kTile];  // O rows of a finished unit on their way out (TMA tile store)
  float xmax[2][2][128];        // [item parity][column half][row]: the two threads of a row exchange their maxima
  float xl[2][2][128];          // [store parity][column half][row]: ... and their partial normalisers
  uint64_t full[2], empty[2], bar_s[2], bar_p, bar_o, ost_full, ost_free;
  uint32_t tmem_slot;
};
// TMEM columns: S[0] 0..127, S[1] 128..255, O 256..319
constexpr uint32_t kFwdTmemCols = 512;

template <bool DROP>
__global__ void __launch_bounds__(kThreads, 1)
fwd_tc_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_k,
              const __grid_constant__ CUtensorMap tma_v, const __grid_constant__ CUtensorMap tma_o, const Params p) {
  extern __shared__ uint8_t fat_raw[];
  FwdSmem& T = *reinterpret_cast<FwdSmem*>((reinterpret_cast<uintptr_t>(fat_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) FAT_TRACE(0);
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_q);
    tma_prefetch_desc(&tma_k);
    tma_prefetch_desc(&tma_v);
    tma_prefetch_desc(&tma_o);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&T.full[s], 1);
      mbar_init(&T.empty[s], 1);
      mbar_init(&T.bar_s[s], 1);
    }
    mbar_init(&T.bar_p, kRowWarps);  // one arrival per row warp
    mbar_init(&T.bar_o, 1);
    mbar_init(&T.ost_full, kRowWarps);
    mbar_init(&T.ost_free, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(&T.tmem_slot, kFwdTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = T.tmem_slot;
  if (threadIdx.x == 0) FAT_TRACE(1);

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      Walk w(p, true);
      bool ok = w.start();
      grid_dep_wait();
      FAT_TRACE(2);
      int stage = 0;
      uint32_t phase = 0;
      long n = 0;
      for (; ok; ok = w.next(), ++n) {
        mbar_wait(&T.empty[stage], phase ^ 1);
        FwdStage& S = T.st[stage];
        mbar_arrive_expect_tx(&T.full[stage], (uint32_t)sizeof(FwdStage));
        const int h0 = p.pair ? w.hs * 2 : w.hs;
        const int q0 = p.pair ? 0 : w.qblock() * 128, k0 = p.pair ? 0 : w.kblock() * 128;
        tma_load_4d(S.q, &tma_q, &T.full[stage], 0, q0, h0, w.b);
        tma_load_4d(S.k, &tma_k, &T.full[stage], 0, k0, h0, w.b);
        tma_load_4d(S.v, &tma_v, &T.full[stage], 0, k0, h0, w.b);
        FAT_TRACE_ITEM(n, 0);
        stage ^= 1;
        if (stage == 0) phase ^= 1;
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t id_s = umma_idesc_bf16(128, 128, 0u, 0u);  // S = Q K^T: both operands K-major
      constexpr uint32_t id_o = umma_idesc_bf16(128, 64, 0u, 1u);   // O = P V  : A K-major, B MN-major
      auto issue_pv = [&](int st) {
        const uint32_t sv = smem_u32(T.st[st].v);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint32_t sa = smem_u32(T.p[kk >> 2]) + (kk & 3) * 32;
          umma_bf16_ss(tmem_base + 256, umma_smem_desc(sa, 0, 1024), umma_smem_desc(sv + kk * 2048, 64 * 128, 1024), id_o,
                       kk > 0 ? 1u : 0u);
        }
        umma_commit(&T.bar_o);
        umma_commit(&T.empty[st]);
      };
      Walk w(p, true);
      int stage = 0, prev_stage = 0;
      uint32_t phase = 0;
      long n = 0;
      for (bool ok = w.start(); ok; ok = w.next(), ++n) {
        mbar_wait(&T.full[stage], phase);
        tc_fence_after();
        FAT_TRACE_ITEM(n, 1);
        {
          const uint32_t sq = smem_u32(T.st[stage].q), sk = smem_u32(T.st[stage].k);
          const uint32_t d = tmem_base + (uint32_t)(n & 1) * 128;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_bf16_ss(d, umma_smem_desc(sq + kk * 32, 0, 1024), umma_smem_desc(sk + kk * 32, 0, 1024), id_s,
                         kk > 0 ? 1u : 0u);
          umma_commit(&T.bar_s[n & 1]);
          FAT_TRACE_ITEM(n, 2);
        }
        if (n > 0) {
          mbar_wait(&T.bar_p, (uint32_t)((n - 1) & 1));
          tc_fence_after();
          issue_pv(prev_stage);
          FAT_TRACE_ITEM(n - 1, 5);
        }
        prev_stage = stage;
        stage ^= 1;
        if (stage == 0) phase ^= 1;
      }
      if (n > 0) {
        mbar_wait(&T.bar_p, (uint32_t)((n - 1) & 1));
        tc_fence_after();
        issue_pv(prev_stage);
        FAT_TRACE_ITEM(n - 1, 5);
      }
    }
  } else if (warp == 2 + kRowWarps) {
    // ------------------------------------------------------------------ store warp: finished units leave by TMA
    if (lane == 0) {
      Walk w(p, true);
      long cnt = 0;
      for (bool ok = w.start(); ok; ok = w.next()) {
        if (!w.last()) continue;
        mbar_wait(&T.ost_full, (uint32_t)(cnt & 1));
        tma_store_4d(&tma_o, T.ostage, 0, p.pair ? 0 : w.qblock() * 128, p.pair ? w.hs * 2 : w.hs, w.b);
        bulk_commit_group();
        bulk_wait_read_all();
        mbar_arrive(&T.ost_free);
        ++cnt;
      }
      bulk_wait_all();
    }
  } else {
    // ------------------------------------------------------------------ row warps: softmax + accumulation.
    // Two threads per row (warps w and w + 4 share a TMEM lane quadrant): each owns half of the row's key columns
    // of S and half of the 64 output channels of O.
    const int quad = warp & 3;          // TMEM lane quadrant this warp may read
    const int half = (warp - 2) >> 2;   // which half of the row's columns
    const int row = quad * 32 + lane;   // row of the 128-row block
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const int item_r = p.pair ? (row >> 6) : 0;
    // 32-column chunks of S owned by this thread: one of its head's two (pair mode), or two of the block's four
    const int c_begin = p.pair ? 2 * item_r + half : 2 * half, c_cnt = p.pair ? 1 : 2;
    const uint32_t p_atom0 = smem_u32(T.p[0]), p_atom1 = smem_u32(T.p[1]), ost = smem_u32(T.ostage);
    const float sl2 = p.scale * kLog2e, inf2 = p.inf_value * kLog2e;
    grid_dep_wait();
    const Drop<DROP> drop(p);
    float m_run = -INFINITY, l_run = 0.f;  // base-2 running maximum (whole row), normaliser (this thread's columns)
    float acc[32];                          // channels [32 half, 32 half + 32) of the row's output
    float m_acc = 0.f;
    bool acc_empty = true;
    long stores = 0;
    // the block whose O = P V has been issued but not folded into acc yet
    bool pend = false, pend_last = false;
    long pend_n = 0;
    float pend_m = 0.f, pend_l = 0.f;
    int pend_b = 0, pend_h = 0, pend_i = 0;

    auto fold_pending = [&]() {
      mbar_wait(&T.bar_o, (uint32_t)(pend_n & 1));
      tc_fence_after();
      if (threadIdx.x == 64) FAT_TRACE_ITEM(pend_n, 6);
      uint32_t ra[32];
      tmem_ld_32x32b_x32(t_lane + 256 + 32 * half, ra);
      tmem_ld_wait();
      if (acc_empty) {
#pragma unroll
        for (int c = 0; c < 32; ++c) acc[c] = __uint_as_float(ra[c]);
        acc_empty = false;
      } else {
        const float corr = ex2(m_acc - pend_m);
#pragma unroll
        for (int c = 0; c < 32; ++c) acc[c] = fmaf(acc[c], corr, __uint_as_float(ra[c]));
      }
      m_acc = pend_m;
      if (pend_last) {
        // the unit's rows are complete: normaliser = both halves' sums; bf16 into the staging tile, out by one TMA
        // tile store (rows past lq are clipped)
        float* xl = &T.xl[stores & 1][0][0];
        xl[half * 128 + row] = pend_l;
        mbar_wait(&T.ost_free, (uint32_t)((stores & 1) ^ 1));
        row_warps_sync();
        const float l_tot = p.relu ? 1.f : xl[row] + xl[128 + row];
        const float inv = 1.f / l_tot;
#pragma unroll
        for (int c = 0; c < 4; ++c)
          sts128(ost + swz_off(row, 4 * half + c), pack_bf16x2(acc[8 * c + 0] * inv, acc[8 * c + 1] * inv),
                 pack_bf16x2(acc[8 * c + 2] * inv, acc[8 * c + 3] * inv),
                 pack_bf16x2(acc[8 * c + 4] * inv, acc[8 * c + 5] * inv),
                 pack_bf16x2(acc[8 * c + 6] * inv, acc[8 * c + 7] * inv));
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&T.ost_full);
        ++stores;
        if (half == 0 && pend_i < p.lq && p.lse)
          p.lse[((long long)pend_b * p.heads + pend_h) * p.lq + pend_i] = p.relu ? 0.f : pend_m * kLn2 + __logf(l_tot);
        acc_empty = true;
      }
      if (threadIdx.x == 64) FAT_TRACE_ITEM(pend_n, 7);
      pend = false;
    };

    Walk w(p, true);
    long n = 0;
    for (bool ok = w.start(); ok; ok = w.next(), ++n) {
      const int b = w.b;
      const int h = p.pair ? w.hs * 2 + item_r : w.hs;
      const int i = p.pair ? (row & 63) : w.qblock() * 128 + row;  // query position of this thread's row
      const int kbase = p.pair ? 0 : w.kblock() * 128;
      const int kl = p.key_len ? __ldg(p.key_len + b) : p.lk;
      // keys [0, jv) are visible to this row, [jv, lk) are masked (func.attention_bias), >= lk do not exist
      const int jv = min(kl, p.causal ? i + p.q_offset + 1 : p.lk);
      const bool vis = jv - kbase >= 1;  // the row sees at least one key of this block
      const uint32_t t_s = t_lane + (uint32_t)(n & 1) * 128;
      mbar_wait(&T.bar_s[n & 1], (uint32_t)((n >> 1) & 1));
      tc_fence_after();
      if (threadIdx.x == 64) FAT_TRACE_ITEM(n, 3);
      // ---- pass 1: the maximum of this thread's columns, then of the row (ReLA: no normalisation, m = 0 throughout)
      float mt = p.relu ? 0.f : -INFINITY;
      if (!p.relu) {
#pragma unroll 1
        for (int cc = 0; cc < c_cnt; ++cc) {
          const int c32 = c_begin + cc;
          uint32_t ra[32];
          tmem_ld_32x32b_x32(t_s + c32 * 32, ra);
          tmem_ld_wait();
          const int base = p.pair ? (c32 & 1) * 32 : kbase + c32 * 32;  // key position of the chunk's first column
          const int nv = jv - base, nb = p.lk - base;
          mt = fmaxf(mt, chunk_max(ra, chunk_mode(nv, nb, vis), nv, nb, sl2, inf2));
        }
        float* xm = &T.xmax[n & 1][0][0];
        xm[half * 128 + row] = mt;
        row_warps_sync();
        mt = fmaxf(xm[row], xm[128 + row]);
      }
      // ---- the previous block's O is complete by now (its P V ran under pass 1): fold it, finish its row if last
      if (pend) fold_pending();
      const bool first = w.first();
      const float m_new = first ? mt : fmaxf(m_run, mt);
      l_run = first ? 0.f : l_run * ex2(m_run - m_new);
      m_run = m_new;
      // ---- pass 2: P = 2^(t - m) (bf16, dropout applied) -> shared memory; l += sum of the undropped weights
      if (p.pair) store_chunk_zero(item_r ? p_atom0 : p_atom1, row, 4 * half);  // the other head's keys: zeros
#pragma unroll 1
      for (int cc = 0; cc < c_cnt; ++cc) {
        const int c32 = c_begin + cc;
        uint32_t ra[32];
        tmem_ld_32x32b_x32(t_s + c32 * 32, ra);
        tmem_ld_wait();
        const int base = p.pair ? (c32 & 1) * 32 : kbase + c32 * 32;
        const int nv = jv - base, nb = p.lk - base;
        float e[32];
        if (p.relu) chunk_relu(ra, e, nv, p.scale);
        else l_run += chunk_exp(ra, e, chunk_mode(nv, nb, vis), nv, nb, sl2, inf2, m_run);
        if (DROP) {
#pragma unroll
          for (int jj = 0; jj < 32; ++jj) e[jj] *= drop.mul(p, b, h, i, base + jj);
        }
        store_chunk_bf16((c32 >> 1) ? p_atom1 : p_atom0, row, (c32 & 1) * 4, e);
      }
      fence_proxy_async_smem();  // generic-proxy writes of P -> visible to the tensor core's async-proxy reads
      tc_fence_before();         // this thread's tcgen05.ld of S / O are complete before the MMA warp proceeds
      __syncwarp();
      if (lane == 0) mbar_arrive(&T.bar_p);
      if (threadIdx.x == 64) FAT_TRACE_ITEM(n, 4);
      pend = true;
      pend_last = w.last();
      pend_n = n;
      pend_m = m_run;
      pend_l = l_run;
      pend_b = b;
      pend_h = h;
      pend_i = i;
    }
    if (pend) fold_pending();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kFwdTmemCols);
  }
  if (threadIdx.x == 0) FAT_TRACE(3);
}

// ================================================================================================ backward
struct BwdStage {
  __nv_bfloat16 q[kTile], k[kTile], v[kTile], d_o[kTile];  // after the MMAs: staging of dQ, dK, dV on their way out
};
struct BwdSmem {
  BwdStage st[2];
  __nv_bfloat16 p[2][kTile], ds[2][kTile];  // two 64-key atoms of [128 query rows][64 keys] each
  float xdelta[2][2][128];                   // [item parity][channel half][row]: partial rowsum(dO * O)
  uint64_t full[2], empty[2], bar_s, bar_p, bar_o, stg_full;
  uint32_t tmem_slot;
};
// TMEM columns: S 0..127, dP 128..255, dQ 256..319, dK 320..383, dV 384..447
constexpr uint32_t kBwdTmemCols = 512;

template <bool DROP>
__global__ void __launch_bounds__(kThreads, 1)
bwd_tc_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_k,
              const __grid_constant__ CUtensorMap tma_v, const __grid_constant__ CUtensorMap tma_do,
              const __grid_constant__ CUtensorMap tma_dq, const __grid_constant__ CUtensorMap tma_dk,
              const __grid_constant__ CUtensorMap tma_dv, const Params p) {
  extern __shared__ uint8_t fat_raw[];
  BwdSmem& T = *reinterpret_cast<BwdSmem*>((reinterpret_cast<uintptr_t>(fat_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) FAT_TRACE(0);
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_q);
    tma_prefetch_desc(&tma_k);
    tma_prefetch_desc(&tma_v);
    tma_prefetch_desc(&tma_do);
    tma_prefetch_desc(&tma_dq);
    tma_prefetch_desc(&tma_dk);
    tma_prefetch_desc(&tma_dv);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&T.full[s], 1);
      mbar_init(&T.empty[s], 1);  // released by the store warp once the outputs staged in the slot have left
    }
    mbar_init(&T.bar_s, 1);
    mbar_init(&T.bar_p, kRowWarps);
    mbar_init(&T.bar_o, 1);
    mbar_init(&T.stg_full, kRowWarps);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(&T.tmem_slot, kBwdTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = T.tmem_slot;
  if (threadIdx.x == 0) FAT_TRACE(1);

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      Walk w(p, false);
      bool ok = w.start();
      grid_dep_wait();
      FAT_TRACE(2);
      int stage = 0;
      uint32_t phase = 0;
      long n = 0;
      for (; ok; ok = w.next(), ++n) {
        mbar_wait(&T.empty[stage], phase ^ 1);
        BwdStage& S = T.st[stage];
        mbar_arrive_expect_tx(&T.full[stage], (uint32_t)sizeof(BwdStage));
        const int h0 = p.pair ? w.hs * 2 : w.hs;
        const int q0 = p.pair ? 0 : w.qblock() * 128, k0 = p.pair ? 0 : w.kblock() * 128;
        tma_load_4d(S.q, &tma_q, &T.full[stage], 0, q0, h0, w.b);
        tma_load_4d(S.k, &tma_k, &T.full[stage], 0, k0, h0, w.b);
        tma_load_4d(S.d_o, &tma_do, &T.full[stage], 0, q0, h0, w.b);
        tma_load_4d(S.v, &tma_v, &T.full[stage], 0, k0, h0, w.b);
        FAT_TRACE_ITEM(n, 0);
        stage ^= 1;
        if (stage == 0) phase ^= 1;
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t id_s = umma_idesc_bf16(128, 128, 0u, 0u);
      constexpr uint32_t id_kmaj = umma_idesc_bf16(128, 64, 0u, 1u);  // A K-major,  B MN-major
      constexpr uint32_t id_mn = umma_idesc_bf16(128, 64, 1u, 1u);    // A MN-major, B MN-major
      const uint32_t sp = smem_u32(T.p[0]), sds = smem_u32(T.ds[0]);
      Walk w(p, false);
      int stage = 0;
      uint32_t phase = 0;
      long n = 0;
      for (bool ok = w.start(); ok; ok = w.next(), ++n) {
        const uint32_t sq = smem_u32(T.st[stage].q), sk = smem_u32(T.st[stage].k), sv = smem_u32(T.st[stage].v),
                       sdo = smem_u32(T.st[stage].d_o);
        mbar_wait(&T.full[stage], phase);
        tc_fence_after();
        FAT_TRACE_ITEM(n, 1);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)  // S = Q K^T
          umma_bf16_ss(tmem_base, umma_smem_desc(sq + kk * 32, 0, 1024), umma_smem_desc(sk + kk * 32, 0, 1024), id_s,
                       kk > 0 ? 1u : 0u);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)  // dP = dO V^T
          umma_bf16_ss(tmem_base + 128, umma_smem_desc(sdo + kk * 32, 0, 1024), umma_smem_desc(sv + kk * 32, 0, 1024),
                       id_s, kk > 0 ? 1u : 0u);
        umma_commit(&T.bar_s);
        FAT_TRACE_ITEM(n, 2);
        mbar_wait(&T.bar_p, (uint32_t)(n & 1));
        tc_fence_after();
        const uint32_t keep = w.first() ? 0u : 1u;  // dK / dV accumulate over the unit's query blocks
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)  // dQ = dS K : k = keys; A atom (kk / 4), 32 B per k-step inside an atom
          umma_bf16_ss(tmem_base + 256, umma_smem_desc(sds + (kk >> 2) * 16384 + (kk & 3) * 32, 0, 1024),
                       umma_smem_desc(sk + kk * 2048, 64 * 128, 1024), id_kmaj, kk > 0 ? 1u : 0u);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)  // dK += dS^T Q : k = queries (16 rows = 2048 B per k-step), m atoms 16 KB apart
          umma_bf16_ss(tmem_base + 320, umma_smem_desc(sds + kk * 2048, 16384, 1024),
                       umma_smem_desc(sq + kk * 2048, 64 * 128, 1024), id_mn, (kk > 0 ? 1u : 0u) | keep);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)  // dV += P^T dO
          umma_bf16_ss(tmem_base + 384, umma_smem_desc(sp + kk * 2048, 16384, 1024),
                       umma_smem_desc(sdo + kk * 2048, 64 * 128, 1024), id_mn, (kk > 0 ? 1u : 0u) | keep);
        umma_commit(&T.bar_o);
        FAT_TRACE_ITEM(n, 5);
        stage ^= 1;
        if (stage == 0) phase ^= 1;
      }
    }
  } else if (warp == 2 + kRowWarps) {
    // ------------------------------------------------------------------ store warp: dQ / dK / dV leave by TMA from
    // the slot's own q / k / v buffers, then the slot goes back to the producer
    if (lane == 0) {
      Walk w(p, false);
      int stage = 0;
      long n = 0;
      for (bool ok = w.start(); ok; ok = w.next(), ++n) {
        mbar_wait(&T.stg_full, (uint32_t)(n & 1));
        const int h0 = p.pair ? w.hs * 2 : w.hs;
        if (p.nk == 1) tma_store_4d(&tma_dq, T.st[stage].q, 0, p.pair ? 0 : w.qblock() * 128, h0, w.b);
        if (w.last()) {
          tma_store_4d(&tma_dk, T.st[stage].k, 0, p.pair ? 0 : w.kblock() * 128, h0, w.b);
          tma_store_4d(&tma_dv, T.st[stage].v, 0, p.pair ? 0 : w.kblock() * 128, h0, w.b);
        }
        bulk_commit_group();
        bulk_wait_read_all();
        mbar_arrive(&T.empty[stage]);
        stage ^= 1;
      }
      bulk_wait_all();
    }
  } else {
    // ------------------------------------------------------------------ row warps: two threads per row, each owns
    // half of the row's key columns of S / dP and half of the 64 channels of dQ / dK / dV
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = quad * 32 + lane;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const int item_r = p.pair ? (row >> 6) : 0;
    const int c_begin = p.pair ? 2 * item_r + half : 2 * half, c_cnt = p.pair ? 1 : 2;
    const uint32_t p_atom0 = smem_u32(T.p[0]), p_atom1 = smem_u32(T.p[1]);
    const uint32_t ds_atom0 = smem_u32(T.ds[0]), ds_atom1 = smem_u32(T.ds[1]);
    const float sl2 = p.scale * kLog2e, inf2 = p.inf_value * kLog2e;
    grid_dep_wait();
    const Drop<DROP> drop(p);
    Walk w(p, false);
    long n = 0;
    int stage = 0;
    for (bool ok = w.start(); ok; ok = w.next(), ++n, stage ^= 1) {
      const int b = w.b;
      const int h = p.pair ? w.hs * 2 + item_r : w.hs;
      const int i = p.pair ? (row & 63) : w.qblock() * 128 + row;   // query position of row `row` of S / dP / dQ
      const int kbase = p.pair ? 0 : w.kblock() * 128;
      const int kl = p.key_len ? __ldg(p.key_len + b) : p.lk;
      const bool row_ok = i < p.lq;
      const int jv = row_ok ? min(kl, p.causal ? i + p.q_offset + 1 : p.lk) : 0;
      const int jb = row_ok ? p.lk : 0;  // rows past lq hold zeros (TMA fill): P = dS = 0
      // delta = rowsum(dO * O) (this thread: 32 of the 64 channels) and the row's log-sum-exp, from global memory
      // while the tiles are in flight
      float delta = 0.f, lse2 = 0.f;
      if (row_ok) {
        const long long ch = h * 64 + 32 * half;
        const uint4* orow = reinterpret_cast<const uint4*>(p.o + (long long)b * p.bso + (long long)i * p.ldo + ch);
        const uint4* drow = reinterpret_cast<const uint4*>(p.d_o + (long long)b * p.bsdo + (long long)i * p.lddo + ch);
        float d4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const uint4 a = __ldg(orow + c), d = __ldg(drow + c);
          const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, dw[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 x = unpack_bf16x2(aw[e]), y = unpack_bf16x2(dw[e]);
            d4[e] = fmaf(x.x, y.x, fmaf(x.y, y.y, d4[e]));
          }
        }
        delta = (d4[0] + d4[1]) + (d4[2] + d4[3]);
        lse2 = p.lse[((long long)b * p.heads + h) * p.lq + i] * kLog2e;
      }
      {
        float* xd = &T.xdelta[n & 1][0][0];
        xd[half * 128 + row] = delta;
        row_warps_sync();
        delta = xd[row] + xd[128 + row];
      }
      mbar_wait(&T.bar_s, (uint32_t)(n & 1));
      tc_fence_after();
      if (threadIdx.x == 64) FAT_TRACE_ITEM(n, 3);
      if (p.pair) {  // the other head's keys: zeros
        store_chunk_zero(item_r ? p_atom0 : p_atom1, row, 4 * half);
        store_chunk_zero(item_r ? ds_atom0 : ds_atom1, row, 4 * half);
      }
#pragma unroll 1
      for (int cc = 0; cc < c_cnt; ++cc) {
        const int c32 = c_begin + cc;
        uint32_t rs[32], rp[32];
        tmem_ld_32x32b_x32(t_lane + c32 * 32, rs);
        tmem_ld_32x32b_x32(t_lane + 128 + c32 * 32, rp);
        tmem_ld_wait();
        const int base = p.pair ? (c32 & 1) * 32 : kbase + c32 * 32;
        const int nv = jv - base, nb = jb - base;
        float e[32];
        float dsv[32];
        if (p.relu) {
          // weights relu(s) where visible; their gradient passes where the weight is positive (modules/rela.py:65-70)
          chunk_relu(rs, e, nv, p.scale);
#pragma unroll
          for (int jj = 0; jj < 32; ++jj) {
            const float dm = DROP ? drop.mul(p, b, h, i, base + jj) : 1.f;
            dsv[jj] = e[jj] > 0.f ? __uint_as_float(rp[jj]) * dm : 0.f;
            e[jj] *= dm;
          }
        } else {
          chunk_exp(rs, e, chunk_mode(nv, nb, jv >= 1), nv, nb, sl2, inf2, lse2);  // P = 2^(t - lse2)
          if (DROP) {
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) {
              const float dm = drop.mul(p, b, h, i, base + jj);
              dsv[jj] = e[jj] * fmaf(__uint_as_float(rp[jj]), dm, -delta);
              e[jj] *= dm;
            }
          } else {
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) dsv[jj] = e[jj] * (__uint_as_float(rp[jj]) - delta);
          }
        }
        const int u0 = (c32 & 1) * 4;
        store_chunk_bf16((c32 >> 1) ? p_atom1 : p_atom0, row, u0, e);
        store_chunk_bf16((c32 >> 1) ? ds_atom1 : ds_atom0, row, u0, dsv);
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&T.bar_p);
      if (threadIdx.x == 64) FAT_TRACE_ITEM(n, 4);
      mbar_wait(&T.bar_o, (uint32_t)(n & 1));
      tc_fence_after();
      if (threadIdx.x == 64) FAT_TRACE_ITEM(n, 6);
      // the MMAs that read this slot's q / k / v are complete: the buffers now stage dQ / dK / dV (bf16, swizzled rows)
      auto stage_half = [&](uint32_t tile, const uint32_t (&r)[32], float mul) {
#pragma unroll
        for (int c = 0; c < 4; ++c)
          sts128(tile + swz_off(row, 4 * half + c),
                 pack_bf16x2(__uint_as_float(r[8 * c + 0]) * mul, __uint_as_float(r[8 * c + 1]) * mul),
                 pack_bf16x2(__uint_as_float(r[8 * c + 2]) * mul, __uint_as_float(r[8 * c + 3]) * mul),
                 pack_bf16x2(__uint_as_float(r[8 * c + 4]) * mul, __uint_as_float(r[8 * c + 5]) * mul),
                 pack_bf16x2(__uint_as_float(r[8 * c + 6]) * mul, __uint_as_float(r[8 * c + 7]) * mul));
      };
      {
        uint32_t ra[32];
        tmem_ld_32x32b_x32(t_lane + 256 + 32 * half, ra);
        tmem_ld_wait();
        if (p.nk == 1) {
          stage_half(smem_u32(T.st[stage].q), ra, p.scale);
        } else if (row_ok) {
          float* dst = p.dq32 + (((long long)b * p.lq + i) * p.heads + h) * 64 + 32 * half;
#pragma unroll
          for (int c = 0; c < 8; ++c)
            red_add_v4(dst + 4 * c, __uint_as_float(ra[4 * c]) * p.scale, __uint_as_float(ra[4 * c + 1]) * p.scale,
                       __uint_as_float(ra[4 * c + 2]) * p.scale, __uint_as_float(ra[4 * c + 3]) * p.scale);
        }
      }
      if (w.last()) {  // the unit's dK / dV are complete: rows are keys
        uint32_t ra[32], rb[32];
        tmem_ld_32x32b_x32(t_lane + 320 + 32 * half, ra);
        tmem_ld_32x32b_x32(t_lane + 384 + 32 * half, rb);
        tmem_ld_wait();
        stage_half(smem_u32(T.st[stage].k), ra, p.scale);
        stage_half(smem_u32(T.st[stage].v), rb, 1.f);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&T.stg_full);
      if (threadIdx.x == 64) FAT_TRACE_ITEM(n, 7);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kBwdTmemCols);
  }
  if (threadIdx.x == 0) FAT_TRACE(3);
}

// ================================================================================================ relative positions
// modules/rpr.py:10-75 (Shaw et al.) on the tensor cores, for blocks of one head with lq, lk <= 128 (BASELINE
// configs[3]) and max_relative_position <= 16.  The reference gathers z[i, j, :] = E[clip(i - j, -R, R) + R] into an
// [L, L, dh] tensor and runs two extra batched matmuls over it; here the 2R + 1 rows of E are a GEMM operand:
//   logits   S[i, j] += QE[i, u(i, j)]      with QE = Q E_k^T          [128 x 64 slots]  one more MMA per block
//   context  O[i, :] += sum_u W[i, u] E_v[u]  with W[i, u] = sum of the (dropped) weights of bucket u: P is
//            bucket-summed per row into a [128 x 64 slots] bf16 A operand, O = P V + W E_v is ONE accumulation
//   backward dP[i, j] += (dO E_v^T)[i, u],  dQ += DSb E_k,  dE_k = DSb^T Q,  dE_v = W^T dO  with DSb = bucket-summed dS
// A row's buckets 1 .. 2R - 1 hold exactly one key each (j = i - u + R): the bucket sum is a skewed copy, done by the
// row's two threads with 2-byte shared-memory stores; the clipped buckets 0 / 2R collect everything further away and are
// summed in registers.  Each of the row's two threads has its own pair of clipped slots (0, 2R for the first half of
// the columns; 2R + 1, 2R + 2 for the second, rows 2R + 1 / 2R + 2 of the E tiles being copies of rows 0 / 2R), so the
// halves never have to be added.  QE is staged once per block as fp32 [128][36] in shared memory (row pitch 36 words:
// the skewed read row * 36 + (row + c) is bank-conflict free).
constexpr int kQeStride = 36;   // >= 2 * 16 + 3 slots; 4 * odd: conflict-free 16-byte row writes and skewed reads
constexpr int kMaxRel = 16;

struct RprGeom {
  int R, i_abs, ibase;  // max_rel; absolute position of this thread's row; of its warp's first row
};
// out[jj] = tab[clip(i_abs - (base + jj), -R, R) + R] for the chunk's 32 keys; lo / hi = tab[0] / tab[2R]
__device__ __forceinline__ void rpr_gather(float (&out)[32], const float* tab, float lo, float hi, const RprGeom& g,
                                           int base) {
  const int dmin = g.ibase - base - 31, dmax = g.ibase + 31 - base;  // over the warp's rows x the chunk's keys
  if (dmin >= g.R) {
#pragma unroll
    for (int jj = 0; jj < 32; ++jj) out[jj] = hi;
  } else if (dmax <= -g.R) {
#pragma unroll
    for (int jj = 0; jj < 32; ++jj) out[jj] = lo;
  } else {
    const int u0 = g.i_abs - base + g.R;
#pragma unroll
    for (int jj = 0; jj < 32; ++jj) {
      const int u = u0 - jj;
      out[jj] = u <= 0 ? lo : (u >= 2 * g.R ? hi : tab[u]);
    }
  }
}
// bucket sums of one chunk's values: in-band buckets are 2-byte stores into row `row` of the [128][64 slots] bf16 atom,
// the clipped ones accumulate in lo / hi
__device__ __forceinline__ void rpr_scatter(uint32_t atom, int row, const float (&val)[32], float& lo, float& hi,
                                            const RprGeom& g, int base) {
  const int dmin = g.ibase - base - 31, dmax = g.ibase + 31 - base;
  if (dmin >= g.R) {
#pragma unroll
    for (int jj = 0; jj < 32; ++jj) hi += val[jj];
  } else if (dmax <= -g.R) {
#pragma unroll
    for (int jj = 0; jj < 32; ++jj) lo += val[jj];
  } else {
    const int u0 = g.i_abs - base + g.R;
#pragma unroll
    for (int jj = 0; jj < 32; ++jj) {
      const int u = u0 - jj;
      if (u <= 0) lo += val[jj];
      else if (u >= 2 * g.R) hi += val[jj];
      else {
        const __nv_bfloat16 hv = __float2bfloat16(val[jj]);
        asm volatile("st.shared.b16 [%0], %1;" ::"r"(atom + swz_off(row, u >> 3) + (uint32_t)((u & 7) * 2)),
                     "h"(*reinterpret_cast<const unsigned short*>(&hv))
                     : "memory");
      }
    }
  }
}
__device__ __forceinline__ void rpr_store_slot(uint32_t atom, int row, int u, float v) {
  const __nv_bfloat16 hv = __float2bfloat16(v);
  asm volatile("st.shared.b16 [%0], %1;" ::"r"(atom + swz_off(row, u >> 3) + (uint32_t)((u & 7) * 2)),
               "h"(*reinterpret_cast<const unsigned short*>(&hv))
               : "memory");
}
// rows 2R + 1 / 2R + 2 of an E tile ([64 slots][64 channels] bf16, swizzled) = copies of rows 0 / 2R (one warp)
__device__ __forceinline__ void rpr_dup_rows(uint8_t* tile, int R, int lane) {
  if (lane < 16) {
    const int u = lane & 7, src = lane < 8 ? 0 : 2 * R, dst = lane < 8 ? 2 * R + 1 : 2 * R + 2;
    *reinterpret_cast<uint4*>(tile + swz_off(dst, u)) = *reinterpret_cast<const uint4*>(tile + swz_off(src, u));
  }
}
// 32 + 4 fp32 values of one TMEM row (columns [col, col + 36)) -> tab[0 .. 36)
__device__ __forceinline__ void rpr_stage_row(float* tab, uint32_t taddr) {
  uint32_t ra[32], rb[32];
  tmem_ld_32x32b_x32(taddr, ra);
  tmem_ld_32x32b_x32(taddr + 32, rb);
  tmem_ld_wait();
#pragma unroll
  for (int c = 0; c < 8; ++c)
    *reinterpret_cast<float4*>(tab + 4 * c) = make_float4(__uint_as_float(ra[4 * c]), __uint_as_float(ra[4 * c + 1]),
                                                           __uint_as_float(ra[4 * c + 2]), __uint_as_float(ra[4 * c + 3]));
  *reinterpret_cast<float4*>(tab + 32) =
      make_float4(__uint_as_float(rb[0]), __uint_as_float(rb[1]), __uint_as_float(rb[2]), __uint_as_float(rb[3]));
}

struct RprFwdSmem {
  __nv_bfloat16 q[kTile], k[kTile], v[kTile];
  __nv_bfloat16 p[2][kTile];
  __nv_bfloat16 w[kTile];                      // bucket-summed weights [128 rows][64 slots]
  __nv_bfloat16 ek[kTile / 2], ev[kTile / 2];  // [64 slots][64 channels]; slots past 2R + 2 are zero
  __nv_bfloat16 ostage[kTile];
  float qe[128 * kQeStride];
  float xmax[2][128], xl[2][128];
  uint64_t full, empty, bar_tab, bar_dup, bar_s, bar_p, bar_o, ost_full, ost_free;
  uint32_t tmem_slot;
};
// TMEM columns: S 0..127, O 128..191, QE 192..255

template <bool DROP>
__global__ void __launch_bounds__(kThreads, 1)
fwd_tc_rpr_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_k,
                  const __grid_constant__ CUtensorMap tma_v, const __grid_constant__ CUtensorMap tma_o,
                  const __grid_constant__ CUtensorMap tma_ek, const __grid_constant__ CUtensorMap tma_ev, const Params p,
                  const int R) {
  extern __shared__ uint8_t fat_raw[];
  RprFwdSmem& T = *reinterpret_cast<RprFwdSmem*>((reinterpret_cast<uintptr_t>(fat_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_q);
    tma_prefetch_desc(&tma_k);
    tma_prefetch_desc(&tma_v);
    tma_prefetch_desc(&tma_o);
    tma_prefetch_desc(&tma_ek);
    tma_prefetch_desc(&tma_ev);
    mbar_init(&T.full, 1);
    mbar_init(&T.empty, 1);
    mbar_init(&T.bar_tab, 1);
    mbar_init(&T.bar_dup, kRowWarps);
    mbar_init(&T.bar_s, 1);
    mbar_init(&T.bar_p, kRowWarps);
    mbar_init(&T.bar_o, 1);
    mbar_init(&T.ost_full, kRowWarps);
    mbar_init(&T.ost_free, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(&T.tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = T.tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      Walk w(p, true);
      bool ok = w.start();
      grid_dep_wait();
      mbar_arrive_expect_tx(&T.bar_tab, 2 * (kTile / 2) * 2);
      tma_load_2d(T.ek, &tma_ek, &T.bar_tab, 0, 0);
      tma_load_2d(T.ev, &tma_ev, &T.bar_tab, 0, 0);
      long n = 0;
      for (; ok; ok = w.next(), ++n) {
        mbar_wait(&T.empty, (uint32_t)((n & 1) ^ 1));
        mbar_arrive_expect_tx(&T.full, 3 * kTile * 2);
        tma_load_4d(T.q, &tma_q, &T.full, 0, 0, w.hs, w.b);
        tma_load_4d(T.k, &tma_k, &T.full, 0, 0, w.hs, w.b);
        tma_load_4d(T.v, &tma_v, &T.full, 0, 0, w.hs, w.b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t id_s = umma_idesc_bf16(128, 128, 0u, 0u);
      constexpr uint32_t id_qe = umma_idesc_bf16(128, 64, 0u, 0u);  // QE = Q E_k^T: both K-major
      constexpr uint32_t id_o = umma_idesc_bf16(128, 64, 0u, 1u);
      const uint32_t sq = smem_u32(T.q), sk = smem_u32(T.k), sv = smem_u32(T.v), sw = smem_u32(T.w),
                     sek = smem_u32(T.ek), sev = smem_u32(T.ev);
      mbar_wait(&T.bar_dup, 0);  // implies the tables have landed and their duplicate rows are written
      Walk w(p, true);
      long n = 0;
      for (bool ok = w.start(); ok; ok = w.next(), ++n) {
        mbar_wait(&T.full, (uint32_t)(n & 1));
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          umma_bf16_ss(tmem_base, umma_smem_desc(sq + kk * 32, 0, 1024), umma_smem_desc(sk + kk * 32, 0, 1024), id_s,
                       kk > 0 ? 1u : 0u);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          umma_bf16_ss(tmem_base + 192, umma_smem_desc(sq + kk * 32, 0, 1024), umma_smem_desc(sek + kk * 32, 0, 1024),
                       id_qe, kk > 0 ? 1u : 0u);
        umma_commit(&T.bar_s);
        mbar_wait(&T.bar_p, (uint32_t)(n & 1));
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)  // O = P V
          umma_bf16_ss(tmem_base + 128, umma_smem_desc(smem_u32(T.p[kk >> 2]) + (kk & 3) * 32, 0, 1024),
                       umma_smem_desc(sv + kk * 2048, 64 * 128, 1024), id_o, kk > 0 ? 1u : 0u);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)  // O += W E_v (k = slots)
          umma_bf16_ss(tmem_base + 128, umma_smem_desc(sw + kk * 32, 0, 1024),
                       umma_smem_desc(sev + kk * 2048, 64 * 128, 1024), id_o, 1u);
        umma_commit(&T.bar_o);
        umma_commit(&T.empty);
      }
    }
  } else if (warp == 2 + kRowWarps) {
    if (lane == 0) {
      Walk w(p, true);
      long n = 0;
      for (bool ok = w.start(); ok; ok = w.next(), ++n) {
        mbar_wait(&T.ost_full, (uint32_t)(n & 1));
        tma_store_4d(&tma_o, T.ostage, 0, 0, w.hs, w.b);
        bulk_commit_group();
        bulk_wait_read_all();
        mbar_arrive(&T.ost_free);
      }
      bulk_wait_all();
    }
  } else {
    const int quad = warp & 3, half = (warp - 2) >> 2, row = quad * 32 + lane;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const uint32_t p_atom0 = smem_u32(T.p[0]), p_atom1 = smem_u32(T.p[1]), w_atom = smem_u32(T.w), ost = smem_u32(T.ostage);
    const float sl2 = p.scale * kLog2e, inf2 = p.inf_value * kLog2e;
    float* qrow = T.qe + row * kQeStride;
    grid_dep_wait();
    const Drop<DROP> drop(p);
    mbar_wait(&T.bar_tab, 0);
    if (warp == 2) {
      rpr_dup_rows(reinterpret_cast<uint8_t*>(T.ek), R, lane);
      rpr_dup_rows(reinterpret_cast<uint8_t*>(T.ev), R, lane);
      fence_proxy_async_smem();
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&T.bar_dup);
    Walk w(p, true);
    long n = 0;
    for (bool ok = w.start(); ok; ok = w.next(), ++n) {
      const int b = w.b, h = w.hs, i = row;
      const int kl = p.key_len ? __ldg(p.key_len + b) : p.lk;
      const int jv = min(kl, p.causal ? i + p.q_offset + 1 : p.lk);
      const bool vis = jv >= 1;
      RprGeom g;
      g.R = R;
      g.i_abs = i + p.q_offset;
      g.ibase = quad * 32 + p.q_offset;
      mbar_wait(&T.bar_s, (uint32_t)(n & 1));
      tc_fence_after();
      // this row's QE: the first thread stages slots 0..31, the second 32..35; both clear their half of W
      if (half == 0) {
        uint32_t ra[32];
        tmem_ld_32x32b_x32(t_lane + 192, ra);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 8; ++c)
          *reinterpret_cast<float4*>(qrow + 4 * c) = make_float4(__uint_as_float(ra[4 * c]), __uint_as_float(ra[4 * c + 1]),
                                                                  __uint_as_float(ra[4 * c + 2]), __uint_as_float(ra[4 * c + 3]));
      } else {
        uint32_t ra[32];
        tmem_ld_32x32b_x32(t_lane + 192 + 32, ra);
        tmem_ld_wait();
        *reinterpret_cast<float4*>(qrow + 32) =
            make_float4(__uint_as_float(ra[0]), __uint_as_float(ra[1]), __uint_as_float(ra[2]), __uint_as_float(ra[3]));
      }
      store_chunk_zero(w_atom, row, 4 * half);
      row_warps_sync();
      const float qlo = qrow[0], qhi = qrow[2 * R];
      // ---- pass 1: the maximum of this thread's two chunks, then of the row
      float mt = -INFINITY;
#pragma unroll 1
      for (int cc = 0; cc < 2; ++cc) {
        const int c32 = 2 * half + cc, base = c32 * 32;
        uint32_t ra[32];
        tmem_ld_32x32b_x32(t_lane + c32 * 32, ra);
        tmem_ld_wait();
        float bias[32];
        rpr_gather(bias, qrow, qlo, qhi, g, base);
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) ra[jj] = __float_as_uint(__uint_as_float(ra[jj]) + bias[jj]);
        const int nv = jv - base, nb = p.lk - base;
        mt = fmaxf(mt, chunk_max(ra, chunk_mode(nv, nb, vis), nv, nb, sl2, inf2));
      }
      {
        float* xm = &T.xmax[0][0];
        xm[half * 128 + row] = mt;
        row_warps_sync();
        mt = fmaxf(xm[row], xm[128 + row]);
      }
      // ---- pass 2: weights, bucket sums
      float l_run = 0.f, wlo = 0.f, whi = 0.f;
#pragma unroll 1
      for (int cc = 0; cc < 2; ++cc) {
        const int c32 = 2 * half + cc, base = c32 * 32;
        uint32_t ra[32];
        tmem_ld_32x32b_x32(t_lane + c32 * 32, ra);
        tmem_ld_wait();
        float e[32];
        rpr_gather(e, qrow, qlo, qhi, g, base);
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) ra[jj] = __float_as_uint(__uint_as_float(ra[jj]) + e[jj]);
        const int nv = jv - base, nb = p.lk - base;
        l_run += chunk_exp(ra, e, chunk_mode(nv, nb, vis), nv, nb, sl2, inf2, mt);
        if (DROP) {
#pragma unroll
          for (int jj = 0; jj < 32; ++jj) e[jj] *= drop.mul(p, b, h, i, base + jj);
        }
        store_chunk_bf16((c32 >> 1) ? p_atom1 : p_atom0, row, (c32 & 1) * 4, e);
        rpr_scatter(w_atom, row, e, wlo, whi, g, base);
      }
      rpr_store_slot(w_atom, row, half ? 2 * R + 1 : 0, wlo);
      rpr_store_slot(w_atom, row, half ? 2 * R + 2 : 2 * R, whi);
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&T.bar_p);
      // ---- O = P V + W E_v
      mbar_wait(&T.bar_o, (uint32_t)(n & 1));
      tc_fence_after();
      uint32_t ro[32];
      tmem_ld_32x32b_x32(t_lane + 128 + 32 * half, ro);
      tmem_ld_wait();
      {
        float* xl = &T.xl[0][0];
        xl[half * 128 + row] = l_run;
        mbar_wait(&T.ost_free, (uint32_t)((n & 1) ^ 1));
        row_warps_sync();
        const float l_tot = xl[row] + xl[128 + row];
        const float inv = 1.f / l_tot;
#pragma unroll
        for (int c = 0; c < 4; ++c)
          sts128(ost + swz_off(row, 4 * half + c),
                 pack_bf16x2(__uint_as_float(ro[8 * c + 0]) * inv, __uint_as_float(ro[8 * c + 1]) * inv),
                 pack_bf16x2(__uint_as_float(ro[8 * c + 2]) * inv, __uint_as_float(ro[8 * c + 3]) * inv),
                 pack_bf16x2(__uint_as_float(ro[8 * c + 4]) * inv, __uint_as_float(ro[8 * c + 5]) * inv),
                 pack_bf16x2(__uint_as_float(ro[8 * c + 6]) * inv, __uint_as_float(ro[8 * c + 7]) * inv));
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&T.ost_full);
        if (half == 0 && i < p.lq && p.lse) p.lse[((long long)b * p.heads + h) * p.lq + i] = mt * kLn2 + __logf(l_tot);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

struct RprBwdSmem {
  __nv_bfloat16 q[kTile], k[kTile], v[kTile], d_o[kTile];
  __nv_bfloat16 p[2][kTile], ds[2][kTile];
  __nv_bfloat16 w[kTile], dsb[kTile];          // bucket-summed dropped weights / logit gradients [128 rows][64 slots]
  __nv_bfloat16 ek[kTile / 2], ev[kTile / 2];
  float qe[128 * kQeStride], dov[128 * kQeStride];
  float xdelta[2][128];
  uint64_t full, empty, bar_tab, bar_dup, bar_s, bar_p, bar_o, stg_full, bar_tfree;
  uint32_t tmem_slot;
};
// TMEM columns: S 0..127, dP 128..255, QE 256..319 (then dQ), dO E_v^T 320..383 (then dK), dV 384..447;
// after the row warps have written P / dS: dE_k 0..63 and dE_v 64..127 (M = 64: rows 16 q + t in lanes t < 16 of quadrant q)

template <bool DROP>
__global__ void __launch_bounds__(kThreads, 1)
bwd_tc_rpr_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_k,
                  const __grid_constant__ CUtensorMap tma_v, const __grid_constant__ CUtensorMap tma_do,
                  const __grid_constant__ CUtensorMap tma_dq, const __grid_constant__ CUtensorMap tma_dk,
                  const __grid_constant__ CUtensorMap tma_dv, const __grid_constant__ CUtensorMap tma_ek,
                  const __grid_constant__ CUtensorMap tma_ev, const Params p, const int R, float* __restrict__ d_ek,
                  float* __restrict__ d_ev) {
  extern __shared__ uint8_t fat_raw[];
  RprBwdSmem& T = *reinterpret_cast<RprBwdSmem*>((reinterpret_cast<uintptr_t>(fat_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_q);
    tma_prefetch_desc(&tma_k);
    tma_prefetch_desc(&tma_v);
    tma_prefetch_desc(&tma_do);
    tma_prefetch_desc(&tma_dq);
    tma_prefetch_desc(&tma_dk);
    tma_prefetch_desc(&tma_dv);
    tma_prefetch_desc(&tma_ek);
    tma_prefetch_desc(&tma_ev);
    mbar_init(&T.full, 1);
    mbar_init(&T.empty, 1);
    mbar_init(&T.bar_tab, 1);
    mbar_init(&T.bar_dup, kRowWarps);
    mbar_init(&T.bar_s, 1);
    mbar_init(&T.bar_p, kRowWarps);
    mbar_init(&T.bar_o, 1);
    mbar_init(&T.stg_full, kRowWarps);
    mbar_init(&T.bar_tfree, kRowWarps);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(&T.tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = T.tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      Walk w(p, false);
      bool ok = w.start();
      grid_dep_wait();
      mbar_arrive_expect_tx(&T.bar_tab, 2 * (kTile / 2) * 2);
      tma_load_2d(T.ek, &tma_ek, &T.bar_tab, 0, 0);
      tma_load_2d(T.ev, &tma_ev, &T.bar_tab, 0, 0);
      long n = 0;
      for (; ok; ok = w.next(), ++n) {
        mbar_wait(&T.empty, (uint32_t)((n & 1) ^ 1));
        mbar_arrive_expect_tx(&T.full, 4 * kTile * 2);
        tma_load_4d(T.q, &tma_q, &T.full, 0, 0, w.hs, w.b);
        tma_load_4d(T.k, &tma_k, &T.full, 0, 0, w.hs, w.b);
        tma_load_4d(T.d_o, &tma_do, &T.full, 0, 0, w.hs, w.b);
        tma_load_4d(T.v, &tma_v, &T.full, 0, 0, w.hs, w.b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t id_s = umma_idesc_bf16(128, 128, 0u, 0u);
      constexpr uint32_t id_qe = umma_idesc_bf16(128, 64, 0u, 0u);
      constexpr uint32_t id_kmaj = umma_idesc_bf16(128, 64, 0u, 1u);
      constexpr uint32_t id_mn = umma_idesc_bf16(128, 64, 1u, 1u);
      constexpr uint32_t id_mn64 = umma_idesc_bf16(64, 64, 1u, 1u);  // dE = (bucket sums)^T x rows: 64 slots tall
      const uint32_t sq = smem_u32(T.q), sk = smem_u32(T.k), sv = smem_u32(T.v), sdo = smem_u32(T.d_o);
      const uint32_t sp = smem_u32(T.p[0]), sds = smem_u32(T.ds[0]), sw = smem_u32(T.w), sdsb = smem_u32(T.dsb);
      const uint32_t sek = smem_u32(T.ek), sev = smem_u32(T.ev);
      mbar_wait(&T.bar_dup, 0);
      Walk w(p, false);
      long n = 0;
      for (bool ok = w.start(); ok; ok = w.next(), ++n) {
        mbar_wait(&T.full, (uint32_t)(n & 1));
        if (n > 0) mbar_wait(&T.bar_tfree, (uint32_t)((n - 1) & 1));  // the row warps have drained dE_k / dE_v
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)  // S = Q K^T
          umma_bf16_ss(tmem_base, umma_smem_desc(sq + kk * 32, 0, 1024), umma_smem_desc(sk + kk * 32, 0, 1024), id_s,
                       kk > 0 ? 1u : 0u);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)  // dP = dO V^T
          umma_bf16_ss(tmem_base + 128, umma_smem_desc(sdo + kk * 32, 0, 1024), umma_smem_desc(sv + kk * 32, 0, 1024),
                       id_s, kk > 0 ? 1u : 0u);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)  // QE = Q E_k^T
          umma_bf16_ss(tmem_base + 256, umma_smem_desc(sq + kk * 32, 0, 1024), umma_smem_desc(sek + kk * 32, 0, 1024),
                       id_qe, kk > 0 ? 1u : 0u);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)  // dO E_v^T
          umma_bf16_ss(tmem_base + 320, umma_smem_desc(sdo + kk * 32, 0, 1024), umma_smem_desc(sev + kk * 32, 0, 1024),
                       id_qe, kk > 0 ? 1u : 0u);
        umma_commit(&T.bar_s);
        mbar_wait(&T.bar_p, (uint32_t)(n & 1));
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)  // dQ = dS K
          umma_bf16_ss(tmem_base + 256, umma_smem_desc(sds + (kk >> 2) * 16384 + (kk & 3) * 32, 0, 1024),
                       umma_smem_desc(sk + kk * 2048, 64 * 128, 1024), id_kmaj, kk > 0 ? 1u : 0u);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)  // dQ += DSb E_k (k = slots)
          umma_bf16_ss(tmem_base + 256, umma_smem_desc(sdsb + kk * 32, 0, 1024),
                       umma_smem_desc(sek + kk * 2048, 64 * 128, 1024), id_kmaj, 1u);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)  // dK = dS^T Q
          umma_bf16_ss(tmem_base + 320, umma_smem_desc(sds + kk * 2048, 16384, 1024),
                       umma_smem_desc(sq + kk * 2048, 64 * 128, 1024), id_mn, kk > 0 ? 1u : 0u);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)  // dV = P^T dO
          umma_bf16_ss(tmem_base + 384, umma_smem_desc(sp + kk * 2048, 16384, 1024),
                       umma_smem_desc(sdo + kk * 2048, 64 * 128, 1024), id_mn, kk > 0 ? 1u : 0u);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)  // dE_k = DSb^T Q (k = the block's 128 query rows)
          umma_bf16_ss(tmem_base, umma_smem_desc(sdsb + kk * 2048, 16384, 1024),
                       umma_smem_desc(sq + kk * 2048, 64 * 128, 1024), id_mn64, kk > 0 ? 1u : 0u);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)  // dE_v = W^T dO
          umma_bf16_ss(tmem_base + 64, umma_smem_desc(sw + kk * 2048, 16384, 1024),
                       umma_smem_desc(sdo + kk * 2048, 64 * 128, 1024), id_mn64, kk > 0 ? 1u : 0u);
        umma_commit(&T.bar_o);
      }
    }
  } else if (warp == 2 + kRowWarps) {
    if (lane == 0) {
      Walk w(p, false);
      long n = 0;
      for (bool ok = w.start(); ok; ok = w.next(), ++n) {
        mbar_wait(&T.stg_full, (uint32_t)(n & 1));
        tma_store_4d(&tma_dq, T.q, 0, 0, w.hs, w.b);
        tma_store_4d(&tma_dk, T.k, 0, 0, w.hs, w.b);
        tma_store_4d(&tma_dv, T.v, 0, 0, w.hs, w.b);
        bulk_commit_group();
        bulk_wait_read_all();
        mbar_arrive(&T.empty);
      }
      bulk_wait_all();
    }
  } else {
    const int quad = warp & 3, half = (warp - 2) >> 2, row = quad * 32 + lane;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const uint32_t p_atom0 = smem_u32(T.p[0]), p_atom1 = smem_u32(T.p[1]);
    const uint32_t ds_atom0 = smem_u32(T.ds[0]), ds_atom1 = smem_u32(T.ds[1]);
    const uint32_t w_atom = smem_u32(T.w), dsb_atom = smem_u32(T.dsb);
    const float sl2 = p.scale * kLog2e, inf2 = p.inf_value * kLog2e;
    float* qrow = T.qe + row * kQeStride;
    float* vrow = T.dov + row * kQeStride;
    grid_dep_wait();
    const Drop<DROP> drop(p);
    mbar_wait(&T.bar_tab, 0);
    if (warp == 2) {
      rpr_dup_rows(reinterpret_cast<uint8_t*>(T.ek), R, lane);
      rpr_dup_rows(reinterpret_cast<uint8_t*>(T.ev), R, lane);
      fence_proxy_async_smem();
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&T.bar_dup);
    Walk w(p, false);
    long n = 0;
    for (bool ok = w.start(); ok; ok = w.next(), ++n) {
      const int b = w.b, h = w.hs, i = row;
      const int kl = p.key_len ? __ldg(p.key_len + b) : p.lk;
      const bool row_ok = i < p.lq;
      const int jv = row_ok ? min(kl, p.causal ? i + p.q_offset + 1 : p.lk) : 0;
      const int jb = row_ok ? p.lk : 0;
      RprGeom g;
      g.R = R;
      g.i_abs = i + p.q_offset;
      g.ibase = quad * 32 + p.q_offset;
      float delta = 0.f, lse2 = 0.f;
      if (row_ok) {
        const long long ch = h * 64 + 32 * half;
        const uint4* orow = reinterpret_cast<const uint4*>(p.o + (long long)b * p.bso + (long long)i * p.ldo + ch);
        const uint4* drow = reinterpret_cast<const uint4*>(p.d_o + (long long)b * p.bsdo + (long long)i * p.lddo + ch);
        float d4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const uint4 a = __ldg(orow + c), d = __ldg(drow + c);
          const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, dw[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 x = unpack_bf16x2(aw[e]), y = unpack_bf16x2(dw[e]);
            d4[e] = fmaf(x.x, y.x, fmaf(x.y, y.y, d4[e]));
          }
        }
        delta = (d4[0] + d4[1]) + (d4[2] + d4[3]);
        lse2 = p.lse[((long long)b * p.heads + h) * p.lq + i] * kLog2e;
      }
      T.xdelta[half][row] = delta;
      mbar_wait(&T.bar_s, (uint32_t)(n & 1));
      tc_fence_after();
      // stage this row's QE (first thread) / dO E_v^T (second thread); clear the bucket-sum rows
      if (half == 0) rpr_stage_row(qrow, t_lane + 256);
      else rpr_stage_row(vrow, t_lane + 320);
      store_chunk_zero(w_atom, row, 4 * half);
      store_chunk_zero(dsb_atom, row, 4 * half);
      row_warps_sync();
      delta = T.xdelta[0][row] + T.xdelta[1][row];
      const float qlo = qrow[0], qhi = qrow[2 * R], vlo = vrow[0], vhi = vrow[2 * R];
      float wlo = 0.f, whi = 0.f, dlo = 0.f, dhi = 0.f;
#pragma unroll 1
      for (int cc = 0; cc < 2; ++cc) {
        const int c32 = 2 * half + cc, base = c32 * 32;
        uint32_t rs[32], rp[32];
        tmem_ld_32x32b_x32(t_lane + c32 * 32, rs);
        tmem_ld_32x32b_x32(t_lane + 128 + c32 * 32, rp);
        tmem_ld_wait();
        float e[32], dsv[32];
        rpr_gather(e, qrow, qlo, qhi, g, base);
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) rs[jj] = __float_as_uint(__uint_as_float(rs[jj]) + e[jj]);
        rpr_gather(dsv, vrow, vlo, vhi, g, base);
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) rp[jj] = __float_as_uint(__uint_as_float(rp[jj]) + dsv[jj]);
        const int nv = jv - base, nb = jb - base;
        chunk_exp(rs, e, chunk_mode(nv, nb, jv >= 1), nv, nb, sl2, inf2, lse2);
        if (DROP) {
#pragma unroll
          for (int jj = 0; jj < 32; ++jj) {
            const float dm = drop.mul(p, b, h, i, base + jj);
            dsv[jj] = e[jj] * fmaf(__uint_as_float(rp[jj]), dm, -delta);
            e[jj] *= dm;
          }
        } else {
#pragma unroll
          for (int jj = 0; jj < 32; ++jj) dsv[jj] = e[jj] * (__uint_as_float(rp[jj]) - delta);
        }
        const int u0 = (c32 & 1) * 4;
        store_chunk_bf16((c32 >> 1) ? p_atom1 : p_atom0, row, u0, e);
        store_chunk_bf16((c32 >> 1) ? ds_atom1 : ds_atom0, row, u0, dsv);
        rpr_scatter(w_atom, row, e, wlo, whi, g, base);
        rpr_scatter(dsb_atom, row, dsv, dlo, dhi, g, base);
      }
      rpr_store_slot(w_atom, row, half ? 2 * R + 1 : 0, wlo);
      rpr_store_slot(w_atom, row, half ? 2 * R + 2 : 2 * R, whi);
      rpr_store_slot(dsb_atom, row, half ? 2 * R + 1 : 0, dlo);
      rpr_store_slot(dsb_atom, row, half ? 2 * R + 2 : 2 * R, dhi);
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&T.bar_p);
      mbar_wait(&T.bar_o, (uint32_t)(n & 1));
      tc_fence_after();
      auto stage_half = [&](uint32_t tile, const uint32_t (&r)[32], float mul) {
#pragma unroll
        for (int c = 0; c < 4; ++c)
          sts128(tile + swz_off(row, 4 * half + c),
                 pack_bf16x2(__uint_as_float(r[8 * c + 0]) * mul, __uint_as_float(r[8 * c + 1]) * mul),
                 pack_bf16x2(__uint_as_float(r[8 * c + 2]) * mul, __uint_as_float(r[8 * c + 3]) * mul),
                 pack_bf16x2(__uint_as_float(r[8 * c + 4]) * mul, __uint_as_float(r[8 * c + 5]) * mul),
                 pack_bf16x2(__uint_as_float(r[8 * c + 6]) * mul, __uint_as_float(r[8 * c + 7]) * mul));
      };
      {
        uint32_t ra[32], rb[32];
        tmem_ld_32x32b_x32(t_lane + 256 + 32 * half, ra);
        tmem_ld_32x32b_x32(t_lane + 320 + 32 * half, rb);
        tmem_ld_wait();
        stage_half(smem_u32(T.q), ra, p.scale);
        stage_half(smem_u32(T.k), rb, p.scale);
        tmem_ld_32x32b_x32(t_lane + 384 + 32 * half, ra);
        tmem_ld_wait();
        stage_half(smem_u32(T.v), ra, 1.f);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&T.stg_full);
      {
        // dE_k / dE_v of this block: slot u = 16 quad + lane lives in the lanes < 16 of the quadrant
        uint32_t ra[32], rb[32];
        tmem_ld_32x32b_x32(t_lane + 32 * half, ra);
        tmem_ld_32x32b_x32(t_lane + 64 + 32 * half, rb);
        tmem_ld_wait();
        const int u = 16 * quad + lane;
        if (lane < 16 && u < 2 * R + 3) {
          const int bucket = u <= 2 * R ? u : (u == 2 * R + 1 ? 0 : 2 * R);
          float* dk_ = d_ek + bucket * 64 + 32 * half;
          float* dv_ = d_ev + bucket * 64 + 32 * half;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            red_add_v4(dk_ + 4 * c, __uint_as_float(ra[4 * c]) * p.scale, __uint_as_float(ra[4 * c + 1]) * p.scale,
                       __uint_as_float(ra[4 * c + 2]) * p.scale, __uint_as_float(ra[4 * c + 3]) * p.scale);
            red_add_v4(dv_ + 4 * c, __uint_as_float(rb[4 * c]), __uint_as_float(rb[4 * c + 1]),
                       __uint_as_float(rb[4 * c + 2]), __uint_as_float(rb[4 * c + 3]));
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&T.bar_tfree);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// dq (bf16, strided) = dq32 (fp32 [batch, lq, heads * 64], already scaled)
__global__ void dq_cast_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long lddq,
                               long long bsdq, int lq, int width, long long total4) {
  grid_dep_wait();
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total4; t += (long long)gridDim.x * blockDim.x) {
    const long long e = t * 4;
    const int c = (int)(e % width);
    const long long r = e / width;
    const int i = (int)(r % lq);
    const long long b = r / lq;
    const float4 v = *reinterpret_cast<const float4*>(src + e);
    uint2 o;
    o.x = pack_bf16x2(v.x, v.y);
    o.y = pack_bf16x2(v.z, v.w);
    *reinterpret_cast<uint2*>(dst + b * bsdq + (long long)i * lddq + c) = o;
  }
}

}  // namespace fat

// ------------------------------------------------------------------------------------------------ host side
static bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

static bool tc_enabled() {
  const char* e = getenv("ZB_ATTN_TC");  // per call: the parity tests flip it inside one process
  return !(e && e[0] == '0');
}

long long attention_tc_bwd_workspace_bytes(const zb_attention_args* a) {
  if (!a || a->dh != 64 || a->lk <= 128) return 0;
  return (long long)a->batch * a->lq * a->heads * 64 * (long long)sizeof(float);
}

bool attention_tc_supported(const zb_attention_args* a, bool bwd) {
  if (!tc_enabled()) return false;
  if (a->dh != 64 || a->kv_group > 1) return false;
  if (a->rpr_k) {
    // relative positions: one 128 x 128 block per (batch, head), at most 2 * 16 + 1 buckets (attention_tc.cu)
    if (a->max_rel > fat::kMaxRel || a->lq > 128 || a->lk > 128) return false;
    if (!al16(a->rpr_k) || !al16(a->rpr_v)) return false;
  }
  if (a->lq < 16) return false;  // decode steps (lq = 1) are served by the warp-per-row kernel
  if (a->causal && (a->q_offset != 0 || a->lq != a->lk)) return false;
  if (a->ldq % 8 || a->ldk % 8 || a->ldv % 8 || a->ldo % 8 || a->bsq % 8 || a->bsk % 8 || a->bsv % 8 || a->bso % 8)
    return false;
  if (!al16(a->q) || !al16(a->k) || !al16(a->v) || !al16(a->o)) return false;
  if ((long long)a->heads * 64 > a->ldq || (long long)a->heads * 64 > a->ldk || (long long)a->heads * 64 > a->ldv)
    return false;
  if (bwd) {
    if (a->lddo % 8 || a->lddq % 8 || a->lddk % 8 || a->lddv % 8 || a->bsdo % 8 || a->bsdq % 8 || a->bsdk % 8 ||
        a->bsdv % 8)
      return false;
    if (!al16(a->d_o) || !al16(a->dq) || !al16(a->dk) || !al16(a->dv)) return false;
    if (a->lk > 128 && (!a->workspace || a->workspace_bytes < attention_tc_bwd_workspace_bytes(a))) return false;
  }
  return true;
}

static unsigned long long* trace_buffer(cudaStream_t st) {
  static const bool on = getenv("ZB_ATTN_TRACE") != nullptr;
  static unsigned long long* buf = nullptr;
  if (!on) return nullptr;
  if (!buf) cudaMalloc(&buf, 64 * sizeof(unsigned long long));
  cudaMemsetAsync(buf, 0, 64 * sizeof(unsigned long long), st);
  return buf;
}
static void trace_report(const char* what, const fat::Params& p, int grid, cudaStream_t st) {
  if (!p.trace) return;
  unsigned long long h[64];
  cudaStreamSynchronize(st);
  cudaMemcpy(h, p.trace, sizeof(h), cudaMemcpyDeviceToHost);
  auto d = [&](int i) { return h[i] ? (long long)(h[i] - h[0]) : -1ll; };
  fprintf(stderr, "[zb_attention trace] %s lq=%d lk=%d pair=%d units=%d grid=%d | ns since entry: setup %lld, dep wait %lld, exit %lld\n",
          what, p.lq, p.lk, p.pair, p.units, grid, d(1), d(2), d(3));
  for (int n = 0; n < 4; ++n) {
    if (!h[8 + 8 * n]) break;
    fprintf(stderr, "    item %d: loads issued %lld, tiles landed %lld, MMA-1 committed %lld, rows see S %lld, rows arrived %lld, "
                    "MMA-2 committed %lld, rows see result %lld, rows done %lld\n",
            n, d(8 + 8 * n), d(9 + 8 * n), d(10 + 8 * n), d(11 + 8 * n), d(12 + 8 * n), d(13 + 8 * n), d(14 + 8 * n),
            d(15 + 8 * n));
  }
}

static fat::Params tc_params(const zb_attention_args* a, bool bwd) {
  fat::Params p = {};
  p.o = (const __nv_bfloat16*)a->o; p.out = (__nv_bfloat16*)a->o; p.d_o = (const __nv_bfloat16*)a->d_o;
  p.dq = (__nv_bfloat16*)a->dq; p.dk = (__nv_bfloat16*)a->dk; p.dv = (__nv_bfloat16*)a->dv;
  p.dq32 = (float*)a->workspace;
  p.ldo = a->ldo; p.bso = a->bso; p.lddo = a->lddo; p.bsdo = a->bsdo;
  p.lddq = a->lddq; p.bsdq = a->bsdq; p.lddk = a->lddk; p.bsdk = a->bsdk; p.lddv = a->lddv; p.bsdv = a->bsdv;
  p.batch = a->batch; p.heads = a->heads; p.lq = a->lq; p.lk = a->lk; p.causal = a->causal; p.q_offset = a->q_offset;
  p.relu = a->relu_attn ? 1 : 0;
  p.pair = (a->lq <= 64 && a->lk <= 64 && (a->heads & 1) == 0 && !a->rpr_k) ? 1 : 0;
  p.nq = p.pair ? 1 : (a->lq + 127) / 128;
  p.nk = p.pair ? 1 : (a->lk + 127) / 128;
  p.hsel = p.pair ? a->heads / 2 : a->heads;
  p.units = a->batch * p.hsel * (bwd ? p.nk : p.nq);
  p.key_len = a->key_len; p.scale = a->scale; p.inf_value = a->inf_value; p.lse = a->lse;
  p.drop_rate = (a->dropout_seed && a->dropout_rate > 0.f) ? a->dropout_rate : 0.f;
  p.drop_site = a->dropout_site;
  p.drop_seed = reinterpret_cast<const unsigned long long*>(a->dropout_seed);
  return p;
}

int attention_tc_fwd(const zb_attention_args* a, cudaStream_t st) {
  fat::Params p = tc_params(a, false);
  p.trace = trace_buffer(st);
  CUtensorMap mq, mk, mv, mo;
  const uint32_t br = p.pair ? 64 : 128, bh = p.pair ? 2 : 1;
  int rc = make_map_heads(&mq, a->q, a->lq, a->heads, a->batch, a->ldq, a->bsq, br, bh);
  if (!rc) rc = make_map_heads(&mk, a->k, a->lk, a->heads, a->batch, a->ldk, a->bsk, br, bh);
  if (!rc) rc = make_map_heads(&mv, a->v, a->lk, a->heads, a->batch, a->ldv, a->bsv, br, bh);
  if (!rc) rc = make_map_heads(&mo, a->o, a->lq, a->heads, a->batch, a->ldo, a->bso, br, bh);
  if (rc) return rc;
  if (a->rpr_k) {
    CUtensorMap mek, mev;
    const uint64_t nbk = 2 * (uint64_t)a->max_rel + 1;
    rc = make_map(&mek, a->rpr_k, 64, nbk, 64, 64);
    if (!rc) rc = make_map(&mev, a->rpr_v, 64, nbk, 64, 64);
    if (rc) return rc;
    const int smem_r = (int)sizeof(fat::RprFwdSmem) + 1024;
    static bool attr_r = false;
    if (!attr_r) {
      cudaFuncSetAttribute(fat::fwd_tc_rpr_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_r);
      cudaFuncSetAttribute(fat::fwd_tc_rpr_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_r);
      attr_r = true;
    }
    const int grid_r = p.units < num_sms_compute() ? p.units : num_sms_compute();
    if (p.drop_rate > 0.f)
      ZB_LAUNCH(fat::fwd_tc_rpr_kernel<true>, grid_r, fat::kThreads, smem_r, st, mq, mk, mv, mo, mek, mev, p, (int)a->max_rel);
    else
      ZB_LAUNCH(fat::fwd_tc_rpr_kernel<false>, grid_r, fat::kThreads, smem_r, st, mq, mk, mv, mo, mek, mev, p, (int)a->max_rel);
    note_path(ZB_PATH_ATTN_TC);
    return check_launch("zb_attention_fwd(tcgen05, relative positions)");
  }
  const int smem = (int)sizeof(fat::FwdSmem) + 1024;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(fat::fwd_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(fat::fwd_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    attr = true;
  }
  const int grid = p.units < num_sms_compute() ? p.units : num_sms_compute();
  if (p.drop_rate > 0.f) ZB_LAUNCH(fat::fwd_tc_kernel<true>, grid, fat::kThreads, smem, st, mq, mk, mv, mo, p);
  else ZB_LAUNCH(fat::fwd_tc_kernel<false>, grid, fat::kThreads, smem, st, mq, mk, mv, mo, p);
  note_path(ZB_PATH_ATTN_TC);
  trace_report("fwd", p, grid, st);
  return check_launch("zb_attention_fwd(tcgen05)");
}

int attention_tc_bwd(const zb_attention_args* a, cudaStream_t st) {
  fat::Params p = tc_params(a, true);
  p.trace = trace_buffer(st);
  CUtensorMap mq, mk, mv, mdo, mdq, mdk, mdv;
  const uint32_t br = p.pair ? 64 : 128, bh = p.pair ? 2 : 1;
  int rc = make_map_heads(&mq, a->q, a->lq, a->heads, a->batch, a->ldq, a->bsq, br, bh);
  if (!rc) rc = make_map_heads(&mk, a->k, a->lk, a->heads, a->batch, a->ldk, a->bsk, br, bh);
  if (!rc) rc = make_map_heads(&mv, a->v, a->lk, a->heads, a->batch, a->ldv, a->bsv, br, bh);
  if (!rc) rc = make_map_heads(&mdo, a->d_o, a->lq, a->heads, a->batch, a->lddo, a->bsdo, br, bh);
  if (!rc) rc = make_map_heads(&mdq, a->dq, a->lq, a->heads, a->batch, a->lddq, a->bsdq, br, bh);
  if (!rc) rc = make_map_heads(&mdk, a->dk, a->lk, a->heads, a->batch, a->lddk, a->bsdk, br, bh);
  if (!rc) rc = make_map_heads(&mdv, a->dv, a->lk, a->heads, a->batch, a->lddv, a->bsdv, br, bh);
  if (rc) return rc;
  if (a->rpr_k) {
    CUtensorMap mek, mev;
    const uint64_t nbk = 2 * (uint64_t)a->max_rel + 1;
    rc = make_map(&mek, a->rpr_k, 64, nbk, 64, 64);
    if (!rc) rc = make_map(&mev, a->rpr_v, 64, nbk, 64, 64);
    if (rc) return rc;
    const int smem_r = (int)sizeof(fat::RprBwdSmem) + 1024;
    static bool attr_r = false;
    if (!attr_r) {
      cudaFuncSetAttribute(fat::bwd_tc_rpr_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_r);
      cudaFuncSetAttribute(fat::bwd_tc_rpr_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_r);
      attr_r = true;
    }
    const int grid_r = p.units < num_sms_compute() ? p.units : num_sms_compute();
    if (p.drop_rate > 0.f)
      ZB_LAUNCH(fat::bwd_tc_rpr_kernel<true>, grid_r, fat::kThreads, smem_r, st, mq, mk, mv, mdo, mdq, mdk, mdv, mek, mev, p,
                (int)a->max_rel, a->d_rpr_k, a->d_rpr_v);
    else
      ZB_LAUNCH(fat::bwd_tc_rpr_kernel<false>, grid_r, fat::kThreads, smem_r, st, mq, mk, mv, mdo, mdq, mdk, mdv, mek, mev,
                p, (int)a->max_rel, a->d_rpr_k, a->d_rpr_v);
    note_path(ZB_PATH_ATTN_TC);
    return check_launch("zb_attention_bwd(tcgen05, relative positions)");
  }
  const int smem = (int)sizeof(fat::BwdSmem) + 1024;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(fat::bwd_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(fat::bwd_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    attr = true;
  }
  const long long ws_elems = (long long)a->batch * a->lq * a->heads * 64;
  if (p.nk > 1) {
    cudaError_t e = cudaMemsetAsync(a->workspace, 0, (size_t)ws_elems * sizeof(float), st);
    if (e != cudaSuccess) {
      set_error("zb_attention_bwd: clearing the dq workspace: %s", cudaGetErrorString(e));
      return ZB_ECUDA;
    }
  }
  const int grid = p.units < num_sms_compute() ? p.units : num_sms_compute();
  if (p.drop_rate > 0.f)
    ZB_LAUNCH(fat::bwd_tc_kernel<true>, grid, fat::kThreads, smem, st, mq, mk, mv, mdo, mdq, mdk, mdv, p);
  else
    ZB_LAUNCH(fat::bwd_tc_kernel<false>, grid, fat::kThreads, smem, st, mq, mk, mv, mdo, mdq, mdk, mdv, p);
  note_path(ZB_PATH_ATTN_TC);
  trace_report("bwd", p, grid, st);
  rc = check_launch("zb_attention_bwd(tcgen05)");
  if (rc || p.nk == 1) return rc;
  const long long total4 = ws_elems / 4;
  const int blocks = (int)((total4 + 255) / 256 < 4 * num_sms() ? (total4 + 255) / 256 : 4 * num_sms());
  ZB_LAUNCH(fat::dq_cast_kernel, blocks, 256, 0, st, (const float*)a->workspace, (__nv_bfloat16*)a->dq, (long long)a->lddq,
            (long long)a->bsdq, (int)a->lq, (int)(a->heads * 64), total4);
  return check_launch("zb_attention_bwd(dq cast)");
}

}  // namespace zb
