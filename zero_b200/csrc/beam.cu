// beam.cu — K8: one fused beam-search expansion step per call, one CTA per source sentence.
// Replaces search.py:141-228 (log-softmax, EOS ban at t = 0, length penalty, top-2k over beam*V, //V and %V,
// candidate sequences, alive top-k, finished 3k -> k merge) and search.py:85-113 (_not_finished).
// The reference materialises [B, beam*V] score tensors and runs three tf.nn.top_k + six gather_nd per step;
// here a CTA streams its beam*V logits twice (log-sum-exp, then candidate scan with per-thread top-2k lists)
// and finishes the bookkeeping in shared memory.  All index arithmetic is int32; all scores fp32 with the
// reference's constants (float32.min masking, additive -inf on EOS), ties -> lower flat index like tf.nn.top_k.
#include <float.h>
#include <math.h>

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
  ZB_REQUIRE(a && a->logits && a->max_len && a->alive_seq && a->alive_logp && a->alive_score && a->fin_seq &&
                 a->fin_score && a->fin_flag && a->parent && a->tmp_seq,
             "zb_beam_step: null pointer");
  ZB_REQUIRE(a->beam >= 1 && a->beam <= kMaxBeam, "zb_beam_step: beam must be in [1, %d]", kMaxBeam);
  ZB_REQUIRE(a->vocab >= 2 * a->beam && a->time >= 0 && a->time + 2 <= a->seq_cap, "zb_beam_step: bad vocab/time/seq_cap");
  ZB_REQUIRE((long long)a->beam * a->vocab < (1ll << 31), "zb_beam_step: beam * vocab overflows int32");
  if (a->batch == 0) return ZB_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (2 * a->beam <= 8) ZB_LAUNCH(beam_step_kernel<8>, a->batch, kBeamThreads, 0, st, *a);
  else ZB_LAUNCH(beam_step_kernel<16>, a->batch, kBeamThreads, 0, st, *a);
  return check_launch("zb_beam_step");
}

extern "C" int zb_beam_cond(const zb_beam_args* a, zb_stream_t stream) {
  using namespace zb;
  ZB_REQUIRE(a && a->active && a->max_len && a->max_penalty && a->alive_logp && a->fin_score && a->fin_flag,
             "zb_beam_cond: null pointer");
  ZB_LAUNCH(beam_cond_kernel, 1, 256, 0, reinterpret_cast<cudaStream_t>(stream), *a);
  return check_launch("zb_beam_cond");
}
