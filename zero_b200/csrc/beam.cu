// beam.cu — K8: one fused beam-search expansion step per call, one CTA per source sentence.
// Replaces search.py:141-228 (log-softmax, EOS ban at t = 0, length penalty, top-2k over beam*V, //V and %V,
// candidate sequences, alive top-k, finished 3k -> k merge) and search.py:85-113 (_not_finished).
// The reference materialises [B, beam*V] score tensors and runs three tf.nn.top_k + six gather_nd per step;
// here a CTA streams its beam*V logits twice (log-sum-exp, then candidate scan with per-thread top-2k lists)
// and finishes the bookkeeping in shared memory.  All index arithmetic is int32; all scores fp32 with the
// reference's constants (float32.min masking, additive -inf on EOS), ties -> lower flat index like tf.nn.top_k.
#include <cooperative_groups.h>
#include <float.h>
#include <math.h>
#include <stdlib.h>

#include "zb_common.h"
#include "zb_ptx.cuh"

namespace zb {

constexpr int kBeamThreads = 512;
constexpr int kMaxBeam = 8;  // top-2k list per thread lives in registers: 2 * beam <= 16
#define F32_MIN (-FLT_MAX)

struct Cand {
  float s;
  int i;
};
__device__ __forceinline__ bool better(float s1, int i1, float s2, int i2) {
  return s1 > s2 || (s1 == s2 && i1 < i2);
}

template <int N>
__device__ __forceinline__ void list_insert(float (&ls)[N], int (&li)[N], float s, int i) {
  if (!better(s, i, ls[N - 1], li[N - 1])) return;
  ls[N - 1] = s;
  li[N - 1] = i;
#pragma unroll
  for (int k = N - 1; k > 0; --k) {
    if (better(ls[k], li[k], ls[k - 1], li[k - 1])) {
      const float ts = ls[k]; ls[k] = ls[k - 1]; ls[k - 1] = ts;
      const int ti = li[k]; li[k] = li[k - 1]; li[k - 1] = ti;
    }
  }
}

// stable top-k of a short array in shared memory by one thread: descending, ties -> lower index
__device__ void small_topk(const float* v, int n, int k, float* out_v, int* out_i) {
  unsigned long long used = 0ull;
  for (int r = 0; r < k; ++r) {
    int best = -1;
    for (int c = 0; c < n; ++c) {
      if (used >> c & 1ull) continue;
      if (best < 0 || v[c] > v[best]) best = c;
    }
    used |= 1ull << best;
    out_v[r] = v[best];
    out_i[r] = best;
  }
}

template <int N2>
__global__ void __launch_bounds__(kBeamThreads) beam_step_kernel(const zb_beam_args a) {
  grid_dep_wait();
  if (a.active && a.active[0] == 0) return;
  const int b = blockIdx.x, K = a.beam, V = a.vocab, t = a.time, cap = a.seq_cap;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = kBeamThreads / 32;
  __shared__ float red_m[NW], red_s[NW];
  __shared__ float lse[kMaxBeam];
  __shared__ float cand_s[NW * 32];  // reused: per-thread list heads
  __shared__ int cand_i[NW * 32];
  __shared__ float top_s[2 * kMaxBeam];
  __shared__ int top_i[2 * kMaxBeam];
  __shared__ int bi[2 * kMaxBeam], wi[2 * kMaxBeam], done[2 * kMaxBeam];
  __shared__ float tmpv[3 * kMaxBeam], a_s[kMaxBeam], f_s[kMaxBeam];
  __shared__ int a_i[kMaxBeam], f_i[kMaxBeam], new_flag[kMaxBeam];
  __shared__ int win_thread;

  const float invT_is_one = a.temperature == 1.f;
  // ---- phase 1: log-sum-exp of (logits / T) per beam
  for (int k = 0; k < K; ++k) {
    const float* row = a.logits + ((long long)b * K + k) * V;
    float m = -INFINITY, s = 0.f;
    for (int w = tid; w < V; w += kBeamThreads) {
      const float x = invT_is_one ? row[w] : row[w] / a.temperature;
      if (x > m) {
        s = s * __expf(m - x) + 1.f;
        m = x;
      } else {
        s += __expf(x - m);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
      if (m2 != -INFINITY) {
        if (m == -INFINITY) {
          m = m2;
          s = s2;
        } else {
          const float nm = fmaxf(m, m2);
          s = s * __expf(m - nm) + s2 * __expf(m2 - nm);
          m = nm;
        }
      }
    }
    if (lane == 0) {
      red_m[warp] = m;
      red_s[warp] = s;
    }
    __syncthreads();
    if (tid == 0) {
      float M = -INFINITY;
      for (int w = 0; w < NW; ++w) M = fmaxf(M, red_m[w]);
      float S = 0.f;
      for (int w = 0; w < NW; ++w)
        if (red_m[w] != -INFINITY) S += red_s[w] * expf(red_m[w] - M);
      lse[k] = M + logf(S);
    }
    __syncthreads();
  }
  // ---- phase 2: scan the beam*V candidates, per-thread sorted top-2k lists
  float ls[N2];
  int li[N2];
#pragma unroll
  for (int k = 0; k < N2; ++k) {
    ls[k] = -INFINITY;
    li[k] = 0x7fffffff;
  }
  const float pen = a.length_penalty;
  for (int k = 0; k < K; ++k) {
    const float* row = a.logits + ((long long)b * K + k) * V;
    const float lp_prev = a.alive_logp[b * K + k];
    const float l = lse[k];
    for (int w = tid; w < V; w += kBeamThreads) {
      const float x = invT_is_one ? row[w] : row[w] / a.temperature;
      float lp = x - l;
      if (t < 1 && w == a.eos_id) lp = lp + (-a.inf_value);
      const float sc = (lp_prev + lp) / pen;
      list_insert<N2>(ls, li, sc, k * V + w);
    }
  }
  // ---- phase 3: 2k rounds of block arg-max over the list heads
  const int n2 = 2 * K;
  int head = 0;
  for (int r = 0; r < n2; ++r) {
    float hs = -INFINITY;
    int hi = 0x7fffffff;
#pragma unroll
    for (int k = 0; k < N2; ++k)
      if (k == head) {
        hs = ls[k];
        hi = li[k];
      }
    if (head >= N2) {
      hs = -INFINITY;
      hi = 0x7fffffff;
    }
    float bs = hs;
    int bidx = hi, bthr = tid;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float s2 = __shfl_xor_sync(0xffffffffu, bs, o);
      const int i2 = __shfl_xor_sync(0xffffffffu, bidx, o);
      const int t2 = __shfl_xor_sync(0xffffffffu, bthr, o);
      if (better(s2, i2, bs, bidx)) {
        bs = s2;
        bidx = i2;
        bthr = t2;
      }
    }
    if (lane == 0) {
      cand_s[warp] = bs;
      cand_i[warp] = bidx;
      cand_i[NW + warp] = bthr;
    }
    __syncthreads();
    if (tid == 0) {
      float s0 = cand_s[0];
      int i0 = cand_i[0], t0 = cand_i[NW];
      for (int w = 1; w < NW; ++w)
        if (better(cand_s[w], cand_i[w], s0, i0)) {
          s0 = cand_s[w];
          i0 = cand_i[w];
          t0 = cand_i[NW + w];
        }
      top_s[r] = s0;
      top_i[r] = i0;
      win_thread = t0;
    }
    __syncthreads();
    if (tid == win_thread) ++head;
    __syncthreads();
  }
  // ---- phase 4: bookkeeping (search.py:179-228)
  const int max_len = a.max_len[b];
  if (tid < n2) {
    bi[tid] = top_i[tid] / V;
    wi[tid] = top_i[tid] % V;
    done[tid] = (wi[tid] == a.eos_id) || (t >= max_len);
  }
  __syncthreads();
  if (tid == 0) {
    // alive: k best not-finished candidates
    for (int c = 0; c < n2; ++c) tmpv[c] = top_s[c] + (float)done[c] * F32_MIN;
    small_topk(tmpv, n2, K, a_s, a_i);
    // finished: k best of (previous finished, newly finished)
    for (int k = 0; k < K; ++k) tmpv[k] = a.fin_score[b * K + k];
    for (int c = 0; c < n2; ++c) tmpv[K + c] = top_s[c] + (1.0f - (float)done[c]) * F32_MIN;
    small_topk(tmpv, 3 * K, K, f_s, f_i);
    for (int k = 0; k < K; ++k) new_flag[k] = f_i[k] < K ? a.fin_flag[b * K + f_i[k]] : done[f_i[k] - K];
  }
  __syncthreads();
  // tmp rows: [0,K) previous finished sequences padded with pad at position t+1; [K,3K) candidate sequences
  int* tmp = a.tmp_seq + (long long)b * 3 * K * cap;
  const int newlen = t + 2;
  for (int idx = tid; idx < 3 * K * newlen; idx += kBeamThreads) {
    const int r = idx / newlen, pos = idx % newlen;
    int val;
    if (r < K) {
      val = pos <= t ? a.fin_seq[((long long)b * K + r) * cap + pos] : a.pad_id;
    } else {
      const int c = r - K;
      val = pos <= t ? a.alive_seq[((long long)b * K + bi[c]) * cap + pos] : wi[c];
    }
    tmp[r * cap + pos] = val;
  }
  __syncthreads();
  for (int idx = tid; idx < K * newlen; idx += kBeamThreads) {
    const int k = idx / newlen, pos = idx % newlen;
    a.alive_seq[((long long)b * K + k) * cap + pos] = tmp[(K + a_i[k]) * cap + pos];
    a.fin_seq[((long long)b * K + k) * cap + pos] = tmp[f_i[k] * cap + pos];
  }
  if (tid < K) {
    a.alive_logp[b * K + tid] = a_s[tid] * pen;
    a.alive_score[b * K + tid] = a_s[tid];
    a.fin_score[b * K + tid] = f_s[tid];
    a.fin_flag[b * K + tid] = new_flag[tid];
    a.parent[b * K + tid] = b * K + bi[a_i[tid]];
  }
}

// ------------------------------------------------------------------------------------------------ row-parallel step
// One CTA per (sentence, beam) row instead of one per sentence: batch * beam CTAs fill the machine where `batch` CTAs
// left most SMs idle, and each CTA reads its row of logits from HBM / L2 exactly once (staged in shared memory, then
// max, sum-exp and candidate scan run out of shared memory).  A row's top-2k candidates are exact (same score
// arithmetic and (score, flat index) order as the reference's top_k over beam*V), so the top-2k of the sentence is
// the top-2k of the union of its rows' lists.  The last CTA of a sentence to publish its list (atomic ticket in
// row_ws, the classic fence + counter hand-off) merges the lists and does the bookkeeping of search.py:179-228.

// Pop the n best entries of the lanes' sorted lists (head-of-list arg-max over the warp, n rounds); every lane
// returns with the r-th winner in (ws[r], wi[r]) if `keep`, else lane 0 stores it to out_s / out_i.
template <int N>
__device__ __forceinline__ void warp_pop(const float (&ls)[N], const int (&li)[N], int n, float* out_s, int* out_i) {
  const int lane = threadIdx.x & 31;
  int head = 0;
  for (int r = 0; r < n; ++r) {
    float hs = -INFINITY;
    int hi = 0x7fffffff;
#pragma unroll
    for (int k = 0; k < N; ++k)
      if (k == head) {
        hs = ls[k];
        hi = li[k];
      }
    float bs = hs;
    int bidx = hi, bl = lane;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float s2 = __shfl_xor_sync(0xffffffffu, bs, o);
      const int i2 = __shfl_xor_sync(0xffffffffu, bidx, o);
      const int l2 = __shfl_xor_sync(0xffffffffu, bl, o);
      if (better(s2, i2, bs, bidx)) {
        bs = s2;
        bidx = i2;
        bl = l2;
      }
    }
    bs = __shfl_sync(0xffffffffu, bs, 0);
    bidx = __shfl_sync(0xffffffffu, bidx, 0);
    bl = __shfl_sync(0xffffffffu, bl, 0);
    if (lane == bl) ++head;
    if (lane == 0) {
      out_s[r] = bs;
      out_i[r] = bidx;
    }
  }
}

constexpr int kRowThreads = 512;

template <int N2>
__global__ void __launch_bounds__(kRowThreads) beam_row_kernel(const zb_beam_args a, const int stage) {
  grid_dep_wait();
  if (a.active && a.active[0] == 0) return;
  const int K = a.beam, V = a.vocab, t = a.time, cap = a.seq_cap;
  const int b = blockIdx.x / K, k = blockIdx.x % K;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = kRowThreads / 32;
  const int n2 = 2 * K;
  extern __shared__ __align__(16) float row_smem[];
  __shared__ float red[NW];
  __shared__ float bcast;
  __shared__ float wl_s[NW * N2];   // per-warp winners, later the row's / the sentence's winners
  __shared__ int wl_i[NW * N2];
  __shared__ int is_last;
  __shared__ float top_s[2 * kMaxBeam];
  __shared__ int top_i[2 * kMaxBeam];
  __shared__ int bi[2 * kMaxBeam], wi[2 * kMaxBeam], done[2 * kMaxBeam];
  __shared__ float tmpv[3 * kMaxBeam], a_s[kMaxBeam], f_s[kMaxBeam];
  __shared__ int a_i[kMaxBeam], f_i[kMaxBeam], new_flag[kMaxBeam];

  const float* grow = a.logits + ((long long)b * K + k) * V;
  const bool t_one = a.temperature == 1.f;
  // ---- pass 1: stage x = logits / T in shared memory (when it fits), row max
  float m = -INFINITY;
  const float* src = grow;
  if (stage) {
    if ((V & 3) == 0 && (reinterpret_cast<uintptr_t>(grow) & 15u) == 0) {
      const float4* g4 = reinterpret_cast<const float4*>(grow);
      float4* s4 = reinterpret_cast<float4*>(row_smem);
      const int V4 = V / 4;
      constexpr int U = 4;   // loads in flight per thread: issue U independent 16-byte loads, then consume them
      for (int w0 = tid; w0 < V4; w0 += U * kRowThreads) {
        float4 x[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int w = w0 + u * kRowThreads;
          x[u] = w < V4 ? __ldg(g4 + w) : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int w = w0 + u * kRowThreads;
          if (w < V4) {
            if (!t_one) {
              x[u].x = x[u].x / a.temperature; x[u].y = x[u].y / a.temperature;
              x[u].z = x[u].z / a.temperature; x[u].w = x[u].w / a.temperature;
            }
            s4[w] = x[u];
            m = fmaxf(m, fmaxf(fmaxf(x[u].x, x[u].y), fmaxf(x[u].z, x[u].w)));
          }
        }
      }
    } else {
      for (int w = tid; w < V; w += kRowThreads) {
        const float x = t_one ? grow[w] : grow[w] / a.temperature;
        row_smem[w] = x;
        m = fmaxf(m, x);
      }
    }
    src = row_smem;
  } else {
    for (int w = tid; w < V; w += kRowThreads) m = fmaxf(m, t_one ? grow[w] : grow[w] / a.temperature);
  }
  const bool scaled = stage || t_one;   // src already holds logits / T
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) red[warp] = m;
  __syncthreads();
  if (tid == 0) {
    float M = red[0];
    for (int w = 1; w < NW; ++w) M = fmaxf(M, red[w]);
    bcast = M;
  }
  __syncthreads();
  const float M = bcast;
  // ---- pass 2: sum exp(x - max)
  float sum = 0.f;
  for (int w = tid; w < V; w += kRowThreads) {
    const float x = scaled ? src[w] : src[w] / a.temperature;
    sum += __expf(x - M);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  __syncthreads();   // red / bcast reuse
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  if (tid == 0) {
    float S = 0.f;
    for (int w = 0; w < NW; ++w) S += red[w];
    bcast = M + logf(S);
  }
  __syncthreads();
  const float l = bcast;
  // ---- pass 3: candidate scores of the row (search.py:148-170), per-thread sorted top-2k lists
  float ls[N2];
  int li[N2];
#pragma unroll
  for (int c = 0; c < N2; ++c) {
    ls[c] = -INFINITY;
    li[c] = 0x7fffffff;
  }
  const float pen = a.length_penalty;
  const float lp_prev = a.alive_logp[b * K + k];
  for (int w = tid; w < V; w += kRowThreads) {
    const float x = scaled ? src[w] : src[w] / a.temperature;
    float lp = x - l;
    if (t < 1 && w == a.eos_id) lp = lp + (-a.inf_value);
    const float sc = (lp_prev + lp) / pen;
    list_insert<N2>(ls, li, sc, k * V + w);
  }
  // warp winners -> shared, then warp 0 reduces the NW * n2 warp winners to the row's n2
  warp_pop<N2>(ls, li, n2, wl_s + warp * N2, wl_i + warp * N2);
  __syncthreads();
  float* ws_s = a.row_ws + (long long)b * (4 * K * K + 1);
  int* ws_i = reinterpret_cast<int*>(ws_s) + 2 * K * K;
  unsigned* ticket = reinterpret_cast<unsigned*>(ws_s) + 4 * K * K;
  if (warp == 0) {
#pragma unroll
    for (int c = 0; c < N2; ++c) {
      ls[c] = -INFINITY;
      li[c] = 0x7fffffff;
    }
    for (int c = lane; c < NW * n2; c += 32) {
      const int w = c / n2, r = c % n2;
      list_insert<N2>(ls, li, wl_s[w * N2 + r], wl_i[w * N2 + r]);
    }
    __syncwarp();
    warp_pop<N2>(ls, li, n2, top_s, top_i);
    __syncwarp();
    if (lane < n2) {
      ws_s[k * n2 + lane] = top_s[lane];
      ws_i[k * n2 + lane] = top_i[lane];
    }
    __threadfence();
    __syncwarp();
    if (lane == 0) {
      const unsigned prev = atomicAdd(ticket, 1u);
      is_last = prev == (unsigned)(K - 1);
    }
  }
  __syncthreads();
  if (!is_last) return;
  // ---- the sentence's last row: merge the K lists (K * 2K <= 128 candidates), then search.py:179-228
  __threadfence();
  if (warp == 0) {
#pragma unroll
    for (int c = 0; c < N2; ++c) {
      ls[c] = -INFINITY;
      li[c] = 0x7fffffff;
    }
    for (int c = lane; c < K * n2; c += 32) list_insert<N2>(ls, li, __ldcg(ws_s + c), __ldcg(ws_i + c));
    __syncwarp();
    warp_pop<N2>(ls, li, n2, top_s, top_i);
    if (lane == 0) *ticket = 0u;   // ready for the next step
  }
  __syncthreads();
  const int max_len = a.max_len[b];
  if (tid < n2) {
    bi[tid] = top_i[tid] / V;
    wi[tid] = top_i[tid] % V;
    done[tid] = (wi[tid] == a.eos_id) || (t >= max_len);
  }
  __syncthreads();
  if (tid == 0) {
    for (int c = 0; c < n2; ++c) tmpv[c] = top_s[c] + (float)done[c] * F32_MIN;
    small_topk(tmpv, n2, K, a_s, a_i);
    for (int c = 0; c < K; ++c) tmpv[c] = a.fin_score[b * K + c];
    for (int c = 0; c < n2; ++c) tmpv[K + c] = top_s[c] + (1.0f - (float)done[c]) * F32_MIN;
    small_topk(tmpv, 3 * K, K, f_s, f_i);
    for (int c = 0; c < K; ++c) new_flag[c] = f_i[c] < K ? a.fin_flag[b * K + f_i[c]] : done[f_i[c] - K];
  }
  __syncthreads();
  int* tmp = a.tmp_seq + (long long)b * 3 * K * cap;
  const int newlen = t + 2;
  for (int idx = tid; idx < 3 * K * newlen; idx += kRowThreads) {
    const int r = idx / newlen, pos = idx % newlen;
    int val;
    if (r < K) {
      val = pos <= t ? a.fin_seq[((long long)b * K + r) * cap + pos] : a.pad_id;
    } else {
      const int c = r - K;
      val = pos <= t ? a.alive_seq[((long long)b * K + bi[c]) * cap + pos] : wi[c];
    }
    tmp[r * cap + pos] = val;
  }
  __syncthreads();
  for (int idx = tid; idx < K * newlen; idx += kRowThreads) {
    const int c = idx / newlen, pos = idx % newlen;
    a.alive_seq[((long long)b * K + c) * cap + pos] = tmp[(K + a_i[c]) * cap + pos];
    a.fin_seq[((long long)b * K + c) * cap + pos] = tmp[f_i[c] * cap + pos];
  }
  if (tid < K) {
    a.alive_logp[b * K + tid] = a_s[tid] * pen;
    a.alive_score[b * K + tid] = a_s[tid];
    a.fin_score[b * K + tid] = f_s[tid];
    a.fin_flag[b * K + tid] = new_flag[tid];
    a.parent[b * K + tid] = b * K + bi[a_i[tid]];
  }
}

// ------------------------------------------------------------------------------------------------ part-parallel step
// (opt-in, ZB_BEAM_PARTS=1 — written after the round's last GPU visit, parity-tested but not yet timed.)
// What the timeline says about beam_row_kernel (96 us at batch 64, beam 4, V = 32000): a 128 KB row in shared memory
// means one CTA per SM and two waves of 148 + 108 CTAs, and inside a CTA the per-thread sorted lists dominate — with
// 62 elements per thread almost every warp iteration has some lane inserting, so the whole warp pays the ~130-instruction
// insertion path 62 times (the same divergence model reproduces the 152 us of the one-CTA-per-sentence kernel).
// This kernel attacks both:
//  * a row is split over a cluster of kParts CTAs (a quarter row = 32 KB of shared memory, 256 threads: seven CTAs per
//    SM, all batch * beam * kParts CTAs resident in one wave); the parts exchange (max, sum-exp) through distributed
//    shared memory, so every part scores with the same log-sum-exp, and rank 0 merges the parts' candidate lists;
//  * a threshold pass first: T = the 2k-th largest of the threads' maximum scores.  At least 2k elements score >= T,
//    so the part's exact top-2k all pass `score >= T` (ties included) while on average only a handful of the
//    8000 elements do — the sorted-list insertion runs for those only.
// The hand-off to the sentence's last arriver and the bookkeeping of search.py:179-228 are those of beam_row_kernel.
constexpr int kParts = 4;
constexpr int kPartThreads = 256;

// The last arriver of sentence b: merge the K rows' lists from row_ws, then search.py:179-228.  NT = blockDim.x.
template <int N2, int NT>
__device__ void beam_sentence_tail(const zb_beam_args& a, int b, float* ws_s, int* ws_i, unsigned* ticket) {
  const int K = a.beam, V = a.vocab, t = a.time, cap = a.seq_cap;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n2 = 2 * K;
  const float pen = a.length_penalty;
  __shared__ float top_s[2 * kMaxBeam];
  __shared__ int top_i[2 * kMaxBeam];
  __shared__ int bi[2 * kMaxBeam], wi[2 * kMaxBeam], done[2 * kMaxBeam];
  __shared__ float tmpv[3 * kMaxBeam], a_s[kMaxBeam], f_s[kMaxBeam];
  __shared__ int a_i[kMaxBeam], f_i[kMaxBeam], new_flag[kMaxBeam];
  __threadfence();
  if (warp == 0) {
    float ls[N2];
    int li[N2];
#pragma unroll
    for (int c = 0; c < N2; ++c) {
      ls[c] = -INFINITY;
      li[c] = 0x7fffffff;
    }
    for (int c = lane; c < K * n2; c += 32) list_insert<N2>(ls, li, __ldcg(ws_s + c), __ldcg(ws_i + c));
    __syncwarp();
    warp_pop<N2>(ls, li, n2, top_s, top_i);
    if (lane == 0) *ticket = 0u;   // ready for the next step
  }
  __syncthreads();
  const int max_len = a.max_len[b];
  if (tid < n2) {
    bi[tid] = top_i[tid] / V;
    wi[tid] = top_i[tid] % V;
    done[tid] = (wi[tid] == a.eos_id) || (t >= max_len);
  }
  __syncthreads();
  if (tid == 0) {
    for (int c = 0; c < n2; ++c) tmpv[c] = top_s[c] + (float)done[c] * F32_MIN;
    small_topk(tmpv, n2, K, a_s, a_i);
    for (int c = 0; c < K; ++c) tmpv[c] = a.fin_score[b * K + c];
    for (int c = 0; c < n2; ++c) tmpv[K + c] = top_s[c] + (1.0f - (float)done[c]) * F32_MIN;
    small_topk(tmpv, 3 * K, K, f_s, f_i);
    for (int c = 0; c < K; ++c) new_flag[c] = f_i[c] < K ? a.fin_flag[b * K + f_i[c]] : done[f_i[c] - K];
  }
  __syncthreads();
  int* tmp = a.tmp_seq + (long long)b * 3 * K * cap;
  const int newlen = t + 2;
  for (int idx = tid; idx < 3 * K * newlen; idx += NT) {
    const int r = idx / newlen, pos = idx % newlen;
    int val;
    if (r < K) {
      val = pos <= t ? a.fin_seq[((long long)b * K + r) * cap + pos] : a.pad_id;
    } else {
      const int c = r - K;
      val = pos <= t ? a.alive_seq[((long long)b * K + bi[c]) * cap + pos] : wi[c];
    }
    tmp[r * cap + pos] = val;
  }
  __syncthreads();
  for (int idx = tid; idx < K * newlen; idx += NT) {
    const int c = idx / newlen, pos = idx % newlen;
    a.alive_seq[((long long)b * K + c) * cap + pos] = tmp[(K + a_i[c]) * cap + pos];
    a.fin_seq[((long long)b * K + c) * cap + pos] = tmp[f_i[c] * cap + pos];
  }
  if (tid < K) {
    a.alive_logp[b * K + tid] = a_s[tid] * pen;
    a.alive_score[b * K + tid] = a_s[tid];
    a.fin_score[b * K + tid] = f_s[tid];
    a.fin_flag[b * K + tid] = new_flag[tid];
    a.parent[b * K + tid] = b * K + bi[a_i[tid]];
  }
}

template <int N2>
__global__ void __launch_bounds__(kPartThreads) beam_part_kernel(const zb_beam_args a, const int vp) {
  namespace cg = cooperative_groups;
  grid_dep_wait();
  if (a.active && a.active[0] == 0) return;    // uniform over the grid: no cluster barrier is left waiting
  cg::cluster_group cluster = cg::this_cluster();
  const int part = (int)cluster.block_rank();
  const int K = a.beam, V = a.vocab, t = a.time;
  const int row = blockIdx.x / kParts;          // b * K + k
  const int b = row / K, k = row % K;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = kPartThreads / 32;
  const int n2 = 2 * K;
  extern __shared__ __align__(16) float part_smem[];   // this part's logits / T
  __shared__ float red[NW];
  __shared__ float bcast;
  __shared__ float stat[2];                            // (max, sum exp(x - max)) of this part, read by the cluster
  __shared__ float wl_s[NW * N2];
  __shared__ int wl_i[NW * N2];
  __shared__ float part_s[2 * kMaxBeam];               // this part's top-2k, read by rank 0
  __shared__ int part_i[2 * kMaxBeam];
  __shared__ float row_s[2 * kMaxBeam];
  __shared__ int row_i[2 * kMaxBeam];
  __shared__ int is_last;

  const int w_lo = part * vp, w_hi = min(V, w_lo + vp), n = max(0, w_hi - w_lo);
  const float* grow = a.logits + (long long)row * V + w_lo;
  const bool t_one = a.temperature == 1.f;
  // ---- pass 1: stage the part, part max
  float m = -INFINITY;
  if ((n & 3) == 0 && (reinterpret_cast<uintptr_t>(grow) & 15u) == 0) {
    const float4* g4 = reinterpret_cast<const float4*>(grow);
    float4* s4 = reinterpret_cast<float4*>(part_smem);
    const int n4 = n / 4;
    constexpr int U = 4;
    for (int w0 = tid; w0 < n4; w0 += U * kPartThreads) {
      float4 x[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int w = w0 + u * kPartThreads;
        x[u] = w < n4 ? __ldg(g4 + w) : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int w = w0 + u * kPartThreads;
        if (w < n4) {
          if (!t_one) {
            x[u].x = x[u].x / a.temperature; x[u].y = x[u].y / a.temperature;
            x[u].z = x[u].z / a.temperature; x[u].w = x[u].w / a.temperature;
          }
          s4[w] = x[u];
          m = fmaxf(m, fmaxf(fmaxf(x[u].x, x[u].y), fmaxf(x[u].z, x[u].w)));
        }
      }
    }
  } else {
    for (int w = tid; w < n; w += kPartThreads) {
      const float x = t_one ? grow[w] : grow[w] / a.temperature;
      part_smem[w] = x;
      m = fmaxf(m, x);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) red[warp] = m;
  __syncthreads();
  if (tid == 0) {
    float M = red[0];
    for (int w = 1; w < NW; ++w) M = fmaxf(M, red[w]);
    bcast = M;
  }
  __syncthreads();
  const float mp = bcast;
  // ---- pass 2: part sum of exp(x - part max)
  float sum = 0.f;
  for (int w = tid; w < n; w += kPartThreads) sum += __expf(part_smem[w] - mp);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  __syncthreads();
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  if (tid == 0) {
    float S = 0.f;
    for (int w = 0; w < NW; ++w) S += red[w];
    stat[0] = mp;
    stat[1] = S;
  }
  cluster.sync();
  // every part combines the kParts statistics in the same order -> bit-identical log-sum-exp in the whole row
  float l;
  {
    float pm[kParts], psum[kParts];
    float M = -INFINITY;
#pragma unroll
    for (int p = 0; p < kParts; ++p) {
      const float* st = cluster.map_shared_rank(stat, p);
      pm[p] = st[0];
      psum[p] = st[1];
      M = fmaxf(M, pm[p]);
    }
    float S = 0.f;
#pragma unroll
    for (int p = 0; p < kParts; ++p)
      if (pm[p] != -INFINITY) S += psum[p] * expf(pm[p] - M);
    l = M + logf(S);
  }
  // ---- pass 3a: thread maxima of the candidate scores (search.py:148-170), T = their 2k-th largest
  const float pen = a.length_penalty;
  const float lp_prev = a.alive_logp[row];
  float tm = -INFINITY;
  for (int w = tid; w < n; w += kPartThreads) {
    float lp = part_smem[w] - l;
    if (t < 1 && w_lo + w == a.eos_id) lp = lp + (-a.inf_value);
    tm = fmaxf(tm, (lp_prev + lp) / pen);
  }
  {
    float l1[1] = {tm};
    int i1[1] = {tid};
    warp_pop<1>(l1, i1, n2, wl_s + warp * N2, wl_i + warp * N2);
  }
  __syncthreads();
  float ls[N2];
  int li[N2];
  if (warp == 0) {
#pragma unroll
    for (int c = 0; c < N2; ++c) {
      ls[c] = -INFINITY;
      li[c] = 0x7fffffff;
    }
    for (int c = lane; c < NW * n2; c += 32) {
      const int w = c / n2, r = c % n2;
      list_insert<N2>(ls, li, wl_s[w * N2 + r], wl_i[w * N2 + r]);
    }
    __syncwarp();
    warp_pop<N2>(ls, li, n2, part_s, part_i);   // part_s is scratch here; its last entry is the threshold
    __syncwarp();
    if (lane == 0) bcast = part_s[n2 - 1];
  }
  __syncthreads();
  const float T = bcast;
  // ---- pass 3b: exact top-2k of the part among the few elements scoring >= T
#pragma unroll
  for (int c = 0; c < N2; ++c) {
    ls[c] = -INFINITY;
    li[c] = 0x7fffffff;
  }
  for (int w = tid; w < n; w += kPartThreads) {
    float lp = part_smem[w] - l;
    if (t < 1 && w_lo + w == a.eos_id) lp = lp + (-a.inf_value);
    const float sc = (lp_prev + lp) / pen;
    if (sc >= T) list_insert<N2>(ls, li, sc, k * V + w_lo + w);
  }
  __syncthreads();   // wl_* (threshold scratch) is rewritten below
  warp_pop<N2>(ls, li, n2, wl_s + warp * N2, wl_i + warp * N2);
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int c = 0; c < N2; ++c) {
      ls[c] = -INFINITY;
      li[c] = 0x7fffffff;
    }
    for (int c = lane; c < NW * n2; c += 32) {
      const int w = c / n2, r = c % n2;
      list_insert<N2>(ls, li, wl_s[w * N2 + r], wl_i[w * N2 + r]);
    }
    __syncwarp();
    warp_pop<N2>(ls, li, n2, part_s, part_i);
  }
  cluster.sync();     // every part's top-2k is in its shared memory
  float* ws_s = a.row_ws + (long long)b * (4 * K * K + 1);
  int* ws_i = reinterpret_cast<int*>(ws_s) + 2 * K * K;
  unsigned* ticket = reinterpret_cast<unsigned*>(ws_s) + 4 * K * K;
  if (part == 0 && warp == 0) {
#pragma unroll
    for (int c = 0; c < N2; ++c) {
      ls[c] = -INFINITY;
      li[c] = 0x7fffffff;
    }
    for (int c = lane; c < kParts * n2; c += 32) {
      const int p = c / n2, r = c % n2;
      const float* ps = cluster.map_shared_rank(part_s, p);
      const int* pi = cluster.map_shared_rank(part_i, p);
      list_insert<N2>(ls, li, ps[r], pi[r]);
    }
    __syncwarp();
    warp_pop<N2>(ls, li, n2, row_s, row_i);
    __syncwarp();
    if (lane < n2) {
      ws_s[k * n2 + lane] = row_s[lane];
      ws_i[k * n2 + lane] = row_i[lane];
    }
    __threadfence();
    __syncwarp();
    if (lane == 0) {
      const unsigned prev = atomicAdd(ticket, 1u);
      is_last = prev == (unsigned)(K - 1);
    }
  }
  cluster.sync();     // the parts' shared memory outlives rank 0's reads; also orders is_last for rank 0's threads
  if (part != 0 || !is_last) return;
  beam_sentence_tail<N2, kPartThreads>(a, b, ws_s, ws_i, ticket);
}

// ------------------------------------------------------------------------------------------------------------------
// Candidate kernel: the step after zb_vocab_topk (vocab_topk.cu).  One CTA per (sentence, beam) row; the row's V
// logits were reduced by the GEMM epilogue to `parts` x {max, sum exp, top-8 (x, column)} with x = logit / T.  The row's
// log-sum-exp is the fold of the parts' statistics; every candidate is scored like the logits kernels score every word
// (search.py:148-170) and the row's top-2k goes to the sentence's last arriver exactly as in beam_part_kernel.
constexpr int kCandThreads = 256;

template <int N2>
__global__ void __launch_bounds__(kCandThreads) beam_cand_kernel(const zb_beam_args a, const int parts) {
  grid_dep_wait();
  if (a.active && a.active[0] == 0) return;
  const int K = a.beam, V = a.vocab;
  const int row = blockIdx.x;                   // b * K + k
  const int b = row / K, k = row % K;
  const long long rows = (long long)a.batch * K;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = kCandThreads / 32;
  const int n2 = 2 * K;
  __shared__ float red_m[NW], red_s[NW];
  __shared__ float bcast;
  __shared__ float wl_s[NW * N2];
  __shared__ int wl_i[NW * N2];
  __shared__ float row_s[2 * kMaxBeam];
  __shared__ int row_i[2 * kMaxBeam];
  __shared__ int is_last;
  const float4* stats = reinterpret_cast<const float4*>(a.cand);
  const float* cval = reinterpret_cast<const float*>(stats + (long long)parts * rows);
  const int* cidx = reinterpret_cast<const int*>(cval + (long long)parts * rows * 8);
  // ---- log-sum-exp of the row: fold the parts in a fixed order (thread-strided, butterfly, warps in order)
  float m = -INFINITY, s = 0.f;
  for (int p = tid; p < parts; p += kCandThreads) {
    const float4 t = __ldcg(stats + (long long)p * rows + row);
    if (t.x > -INFINITY) {
      const float nm = fmaxf(m, t.x);
      s = s * expf(m - nm) + t.y * expf(t.x - nm);
      m = nm;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
    const float nm = fmaxf(m, m2);
    const float sa = m > -INFINITY ? s * expf(m - nm) : 0.f, sb = m2 > -INFINITY ? s2 * expf(m2 - nm) : 0.f;
    // both lanes of a pair must add in the same order to agree bit for bit: the lower lane's term first
    s = (lane & o) ? sb + sa : sa + sb;
    m = nm;
  }
  if (lane == 0) {
    red_m[warp] = m;
    red_s[warp] = s;
  }
  __syncthreads();
  if (tid == 0) {
    float M = -INFINITY;
    for (int w = 0; w < NW; ++w) M = fmaxf(M, red_m[w]);
    float S = 0.f;
    for (int w = 0; w < NW; ++w)
      if (red_m[w] > -INFINITY) S += red_s[w] * expf(red_m[w] - M);
    bcast = M + logf(S);
  }
  __syncthreads();
  const float l = bcast;
  // ---- the row's top-2k among the 8 * parts candidates
  const float pen = a.length_penalty;
  const float lp_prev = a.alive_logp[row];
  float ls[N2];
  int li[N2];
#pragma unroll
  for (int c = 0; c < N2; ++c) {
    ls[c] = -INFINITY;
    li[c] = 0x7fffffff;
  }
  // A beam that is not really alive (score float32.min: beams 1.. at t = 0; -inf: a sentence past its max_len) absorbs
  // every log-probability: all V continuations tie at prev / penalty and top_k keeps the lowest columns, whatever the
  // logits are.  The logits kernels get that from the arithmetic; here it has to be said.
  const bool dead = !(lp_prev > F32_MIN);
  if (dead && tid < n2) list_insert<N2>(ls, li, lp_prev / pen, k * V + tid);
  for (int p = tid; p < (dead ? 0 : parts); p += kCandThreads) {
    const long long slot = ((long long)p * rows + row) * 8;
    const float4 v0 = __ldcg(reinterpret_cast<const float4*>(cval + slot));
    const float4 v1 = __ldcg(reinterpret_cast<const float4*>(cval + slot) + 1);
    const int4 i0 = __ldcg(reinterpret_cast<const int4*>(cidx + slot));
    const int4 i1 = __ldcg(reinterpret_cast<const int4*>(cidx + slot) + 1);
    const float xv[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
    const int xi[8] = {i0.x, i0.y, i0.z, i0.w, i1.x, i1.y, i1.z, i1.w};
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      if (xv[c] > -INFINITY) {
        const float lp = xv[c] - l;
        list_insert<N2>(ls, li, (lp_prev + lp) / pen, k * V + xi[c]);
      }
    }
  }
  warp_pop<N2>(ls, li, n2, wl_s + warp * N2, wl_i + warp * N2);
  __syncthreads();
  float* ws_s = a.row_ws + (long long)b * (4 * K * K + 1);
  int* ws_i = reinterpret_cast<int*>(ws_s) + 2 * K * K;
  unsigned* ticket = reinterpret_cast<unsigned*>(ws_s) + 4 * K * K;
  if (warp == 0) {
#pragma unroll
    for (int c = 0; c < N2; ++c) {
      ls[c] = -INFINITY;
      li[c] = 0x7fffffff;
    }
    for (int c = lane; c < NW * n2; c += 32) {
      const int w = c / n2, r = c % n2;
      list_insert<N2>(ls, li, wl_s[w * N2 + r], wl_i[w * N2 + r]);
    }
    __syncwarp();
    warp_pop<N2>(ls, li, n2, row_s, row_i);
    __syncwarp();
    if (lane < n2) {
      ws_s[k * n2 + lane] = row_s[lane];
      ws_i[k * n2 + lane] = row_i[lane];
    }
    __threadfence();
    __syncwarp();
    if (lane == 0) {
      const unsigned prev = atomicAdd(ticket, 1u);
      is_last = prev == (unsigned)(K - 1);
    }
  }
  __syncthreads();
  if (!is_last) return;
  beam_sentence_tail<N2, kCandThreads>(a, b, ws_s, ws_i, ticket);
}

// search.py:85-113 _not_finished(time): not(all_b(worst finished > best alive bound)) and any_b(time < max_len)
__global__ void beam_cond_kernel(const zb_beam_args a) {
  grid_dep_wait();
  __shared__ int s_all, s_any;
  if (threadIdx.x == 0) {
    s_all = 1;
    s_any = 0;
  }
  __syncthreads();
  for (int b = threadIdx.x; b < a.batch; b += blockDim.x) {
    const float best_alive = a.alive_logp[b * a.beam] / a.max_penalty[b];
    float worst = INFINITY;
    int any_fin = 0;
    for (int k = 0; k < a.beam; ++k) {
      const int fl = a.fin_flag[b * a.beam + k];
      worst = fminf(worst, a.fin_score[b * a.beam + k] * (float)fl);
      any_fin |= fl;
    }
    worst += (1.0f - (float)(any_fin != 0)) * F32_MIN;
    if (!(worst > best_alive)) atomicAnd(&s_all, 0);
    if (a.time < a.max_len[b]) atomicOr(&s_any, 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) a.active[0] = (!s_all && s_any) ? 1 : 0;
}

}  // namespace zb

extern "C" int zb_beam_step(const zb_beam_args* a, zb_stream_t stream) {
  using namespace zb;
  ZB_REQUIRE(a && (a->logits || a->cand) && a->max_len && a->alive_seq && a->alive_logp && a->alive_score && a->fin_seq &&
                 a->fin_score && a->fin_flag && a->parent && a->tmp_seq,
             "zb_beam_step: null pointer");
  ZB_REQUIRE(a->beam >= 1 && a->beam <= kMaxBeam, "zb_beam_step: beam must be in [1, %d]", kMaxBeam);
  ZB_REQUIRE(a->vocab >= 2 * a->beam && a->time >= 0 && a->time + 2 <= a->seq_cap, "zb_beam_step: bad vocab/time/seq_cap");
  ZB_REQUIRE((long long)a->beam * a->vocab < (1ll << 31), "zb_beam_step: beam * vocab overflows int32");
  if (a->batch == 0) return ZB_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (a->cand) {
    ZB_REQUIRE(!a->logits, "zb_beam_step: give logits or cand, not both");
    ZB_REQUIRE(a->row_ws && 2 * a->beam <= 8 && a->vocab >= 128 && a->vocab - 1 >= 2 * a->beam,
               "zb_beam_step(cand): needs row_ws, beam <= 4 and vocab >= 128");
    ZB_REQUIRE((reinterpret_cast<uintptr_t>(a->cand) & 15) == 0, "zb_beam_step(cand): workspace must be 16-byte aligned");
    ZB_LAUNCH(beam_cand_kernel<8>, a->batch * a->beam, kCandThreads, 0, st, *a, 2 * ((a->vocab + 255) / 256));
    note_path(ZB_PATH_BEAM_CAND);
    return check_launch("zb_beam_step(cand)");
  }
  const char* parts_env = getenv("ZB_BEAM_PARTS");   // per call: the parity test flips it inside one process
  const int vp = (((a->vocab + kParts - 1) / kParts) + 3) & ~3;   // elements per part, 16-byte granular
  // default since the r02a A/B (decode step 0.589 -> 0.563 ms); ZB_BEAM_PARTS=0 keeps the one-CTA-per-row kernel
  if (a->row_ws && !(parts_env && parts_env[0] == '0') && (size_t)vp * sizeof(float) <= 64 * 1024) {
    static bool part_attr = false;
    if (!part_attr) {
      cudaFuncSetAttribute(beam_part_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
      cudaFuncSetAttribute(beam_part_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
      part_attr = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(a->batch * a->beam * kParts);
    cfg.blockDim = dim3(kPartThreads);
    cfg.dynamicSmemBytes = (size_t)vp * sizeof(float);
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kParts;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    cudaError_t le = 2 * a->beam <= 8 ? cudaLaunchKernelEx(&cfg, beam_part_kernel<8>, *a, vp)
                                      : cudaLaunchKernelEx(&cfg, beam_part_kernel<16>, *a, vp);
    if (le != cudaSuccess) {
      set_error("zb_beam_step (parts) launch: %s", cudaGetErrorString(le));
      return ZB_ECUDA;
    }
    note_path(ZB_PATH_BEAM_PARTS);
    return check_launch("zb_beam_step(parts)");
  }
  if (a->row_ws) {
    // row-parallel kernel: stage the row in shared memory when it fits beside the static buffers
    const size_t row_bytes = (size_t)a->vocab * sizeof(float);
    const int stage = row_bytes <= 200 * 1024 ? 1 : 0;
    const size_t smem = stage ? row_bytes : 0;
    static bool attr_set = false;
    if (!attr_set) {
      cudaFuncSetAttribute(beam_row_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      cudaFuncSetAttribute(beam_row_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      attr_set = true;
    }
    if (2 * a->beam <= 8) ZB_LAUNCH(beam_row_kernel<8>, a->batch * a->beam, kRowThreads, smem, st, *a, stage);
    else ZB_LAUNCH(beam_row_kernel<16>, a->batch * a->beam, kRowThreads, smem, st, *a, stage);
    note_path(ZB_PATH_BEAM_ROWS);
    return check_launch("zb_beam_step(rows)");
  }
  if (2 * a->beam <= 8) ZB_LAUNCH(beam_step_kernel<8>, a->batch, kBeamThreads, 0, st, *a);
  else ZB_LAUNCH(beam_step_kernel<16>, a->batch, kBeamThreads, 0, st, *a);
  note_path(ZB_PATH_BEAM_SENTENCE);
  return check_launch("zb_beam_step");
}

extern "C" int zb_beam_cond(const zb_beam_args* a, zb_stream_t stream) {
  using namespace zb;
  ZB_REQUIRE(a && a->active && a->max_len && a->max_penalty && a->alive_logp && a->fin_score && a->fin_flag,
             "zb_beam_cond: null pointer");
  ZB_LAUNCH(beam_cond_kernel, 1, 256, 0, reinterpret_cast<cudaStream_t>(stream), *a);
  return check_launch("zb_beam_cond");
}
