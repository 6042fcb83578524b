// gemm2_tcgen05.cu — K1, CTA-pair variant: tcgen05.mma.cta_group::2 (UMMA 256 x BN x 16) over a 2-CTA cluster.
//
// Why: timeline traces of the 1-CTA kernel (profiles/r01_gemm_trace.log) show every GEMM of the training step
// waiting on operand delivery — a 128 x BN tile needs (128 + BN) * 128 B per k-block per SM, more than the
// ~50-60 B/cycle an SM ingests from L2.  With a CTA pair the B tile is split across the two SMs (each loads
// BN/2 rows) and the tensor cores of both read it in place, so a 256 x BN pair-tile costs each SM only
// (128 + BN/2) * 128 B per k-block: 1.5x (BN = 256) to 1.33x (BN = 128) fewer bytes per FLOP.
//
// Roles per CTA (320 threads): warp 0 TMA producer (both CTAs; the .cta_group::2 form of cp.async.bulk.tensor
// signals the LEADER's full barrier), warp 1 MMA issuer (leader CTA only) / TMEM allocator (both),
// warps 2..9 epilogue (each CTA drains its own 128 accumulator rows from its own TMEM).
// tcgen05.commit multicasts "slot free" / "accumulator ready" to both CTAs; the peer's epilogue warps hand the
// accumulator stage back with a remote mbarrier arrive on the leader's barrier.
#include <stdlib.h>

#include "zb_common.h"
#include "zb_ptx.cuh"

namespace zb {

int make_map(CUtensorMap* m, const void* ptr, uint64_t dim0, uint64_t dim1, uint64_t ld, uint32_t box1);
int make_map_f32(CUtensorMap* m, const void* ptr, uint64_t dim0, uint64_t dim1, uint64_t ld, uint32_t box1);

struct Gemm2Params {
  int M, N, K;
  int mt, nt, splits, kb_per_split, kb_total;  // mt counts 256-row pair tiles
  int gn;
  void* d;
  long long ldd;
  const float* bias;
  const __nv_bfloat16* mask;
  long long ldmask;
  float alpha;
  int flags;
  int d_f32;
  int tile_begin;  // first work unit of this problem in the launch (grouped launches)
  int d_tma;  // output written through shared memory + TMA tile stores (full 128 B lines): 1 = bf16, 2 = fp32,
              // 3 = fp32 bulk reduce-add
  // K6 (vocab_ce.cu): the vocabulary projection fused with the label-smoothed cross-entropy.
  //   ce_mode 1: nothing is stored; every epilogue thread reduces its columns of its row to {max, sum exp(x - max),
  //              sum x, x[gold]} and writes them to ce_stats[(n_blk * 2 + column half) * M + row]
  //   ce_mode 2: the tile is turned into d_logits = (exp(x - lse[row]) - soft_label) * weight[row] before the store
  // K8 (vocab_topk.cu): the decode step's vocabulary projection reduced to what the beam step needs.
  //   ce_mode 3: nothing is stored; x = acc / ce_p (the temperature); per (row, column half) {max, sum exp(x - max), 0, 0}
  //              go to ce_stats as in mode 1 and the half's 8 largest x (ties -> lower column; column ce_skip left out)
  //              with their columns to ce_cval / ce_cidx[((n_blk * 2 + half) * M + row) * 8 ...]
  int ce_mode;
  const float* ce_aux;        // mode 2: [M][2] = {log-sum-exp, d loss / d nll} per row
  const int32_t* ce_labels;   // [M] gold class per row
  float4* ce_stats;           // mode 1
  float ce_p, ce_q;           // smoothed target: p on the gold class, q elsewhere (mode 3: ce_p = temperature)
  float* ce_cval;             // mode 3
  int32_t* ce_cidx;           // mode 3
  int ce_skip;                // mode 3: -1 = none
};

// One launch may carry several independent problems of the same operand layouts (the weight-gradient GEMMs of a
// layer): work units are numbered across the problems and dealt round-robin to the CTA pairs.
template <int NP>
struct Gemm2Group {
  CUtensorMap maps[3 * NP];  // a, b, d of each problem
  Gemm2Params prob[NP];
  int count, total_tiles;
  unsigned long long* trace;  // debug (ZB_GEMM_TRACE=1): CTA 0 stamps globaltimer at 10 points of its life
};
constexpr int k2MaxGroup = 8;

constexpr int k2BK = 64;
constexpr int k2Threads = 320;

template <int BN>
struct Gemm2Cfg {
  static constexpr int A_BYTES = 128 * k2BK * 2;        // this CTA's 128 rows of A
  static constexpr int B_BYTES = (BN / 2) * k2BK * 2;   // this CTA's half of the B tile
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int EPI_BYTES = 8 * 4096;  // one 32-row x 128 B (64 bf16) staging tile per epilogue warp
  static constexpr int STAGES_RAW = (192 * 1024) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int TMEM_COLS = 2 * BN;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + 1024 + 256;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// Cluster barrier with a RELAXED arrive: ptxas lowers the release form to MEMBAR.ALL.GPU + ERRBAR before the
// UCGABAR_ARV (SASS), a GPU-scope fence the kernel does not need — what crosses the pair is ordered by
// fence.mbarrier_init / the tcgen05 fences, and the CTA-local part by the __syncthreads in front of it.
__device__ __forceinline__ void cluster_sync_all() {
  __syncthreads();
  asm volatile("barrier.cluster.arrive.relaxed.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `p` (a shared::cta pointer of this CTA) as seen in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(const void* p, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
  return r;
}
// Default semantics (.release.cta), like CUTLASS' ClusterBarrier::arrive(cta_id): the cluster-scope release form
// costs a MEMBAR.ALL.GPU per accumulator hand-back (~1.2 us between "store issued" and "arrived" in the trace);
// the hand-back only has to order the TMEM reads, which tcgen05.fence::before_thread_sync does.
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  // acquire at cluster scope: the arrivals come from the peer CTA
  uint32_t spins = 0;
  uint64_t t0 = 0;
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (ok) return;
    if ((++spins & 0x3FF) == 0) {
      const uint64_t now = globaltimer_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000ull) asm volatile("trap;");
    }
  }
}
// 2-CTA TMA load: data lands in THIS CTA's shared memory, completion bytes are credited to the barrier at the
// same offset in the LEADER CTA (peer bit of the shared::cluster address cleared)
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int32_t c0,
                                                int32_t c1) {
  const uint32_t mbar = smem_u32(bar) & 0xFEFFFFFFu;
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], "
      "[%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(mbar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_ss_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once all previously issued MMAs retire) on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
  const uint16_t mask = 3;
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(mask)
      : "memory");
}

// explicit shared-space 16 B store (the staging pointer arithmetic otherwise compiles to generic ST.E.128)
__device__ __forceinline__ void sts128(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

__device__ __forceinline__ void tile2_coords(const Gemm2Params& p, int tile, int& m_blk, int& n_blk, int& ks) {
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

__device__ __forceinline__ float ex2_(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void red_add_v4_(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}

template <int NP>
__device__ __forceinline__ int find_problem(const Gemm2Group<NP>& G, int tile) {
  int pi = 0;
#pragma unroll
  for (int i = 1; i < NP; ++i)
    if (i < G.count && tile >= G.prob[i].tile_begin) pi = i;
  return pi;
}

// CE: 0 = the plain GEMM; 1 / 2 = the K6 cross-entropy epilogues (Gemm2Params::ce_mode), separate instantiations so that
// the plain kernels keep their register allocation
template <int BN, bool A_MN, bool B_MN, int NP, int CE = 0>
__global__ void __launch_bounds__(k2Threads, 1)
gemm2_bf16_tcgen05(const __grid_constant__ Gemm2Group<NP> G) {
  using Cfg = Gemm2Cfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* epi_smem = smem + STAGES * Cfg::STAGE_BYTES;  // 1024-aligned (stage sizes are multiples of 1 KB)
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(epi_smem + Cfg::EPI_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();  // 0 = leader
  const bool leader = rank == 0;
  const int num_tiles = G.total_tiles;
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const bool tracing = G.trace != nullptr && blockIdx.x == 0;
  if (tracing && threadIdx.x == 0) G.trace[0] = globaltimer_ns();

  if (warp == 0 && lane < 3 * NP && lane < 3 * G.count) tma_prefetch_desc(&G.maps[lane]);
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);   // leader's producer arms it; both CTAs' TMA bytes are credited to the leader's
      mbar_init(&empty_bar[s], 1);  // one multicast commit per phase
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 16);  // 8 epilogue warps of each CTA (used in the leader only)
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc_2sm(tmem_slot, Cfg::TMEM_COLS);
  {
    // The problem descriptors live in the kernel-parameter constant bank; their first read after the dependency wait
    // cost the producer ~0.5 us (trace: dep_wait -> first TMA).  Reading them here, in the shadow of the TMEM
    // allocation and of the predecessor's tail, pulls them into the constant cache.
    int warm = G.prob[0].M + G.prob[0].tile_begin + G.prob[0].d_tma;
    if (NP > 1) warm += G.prob[NP - 1].M + G.prob[NP - 1].d_tma + G.prob[NP / 2].M;
    asm volatile("" ::"r"(warm));
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (tracing && threadIdx.x == 0) G.trace[1] = globaltimer_ns();
  // griddepcontrol.wait is executed only by the threads that touch global memory (the producer and the epilogue
  // warps), each AFTER the parameter-only part of its prologue: the producer's tile coordinates (three integer
  // divisions + constant-bank reads, ~0.3 us) are computed in the shadow of the predecessor's tail.

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int pi = 0, m0 = 0, n0 = 0, kb0 = 0, kb1 = 0;
      auto locate = [&](int tile) {
        pi = find_problem<NP>(G, tile);
        const Gemm2Params& q = G.prob[pi];
        int m_blk, n_blk, ks;
        tile2_coords(q, tile - q.tile_begin, m_blk, n_blk, ks);
        kb0 = ks * q.kb_per_split;
        kb1 = min(q.kb_total, kb0 + q.kb_per_split);
        m0 = m_blk * 256 + (int)rank * 128;       // this CTA's rows of A
        n0 = n_blk * BN + (int)rank * (BN / 2);   // this CTA's rows of B
      };
      if (pair < num_tiles) locate(pair);
      grid_dep_wait();
      if (tracing) G.trace[2] = globaltimer_ns();
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        if (tile != pair) locate(tile);
        const CUtensorMap* tma_a = &G.maps[3 * pi];
        const CUtensorMap* tma_b = &G.maps[3 * pi + 1];
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (tracing && tile == pair && kb == kb0) G.trace[3] = globaltimer_ns();
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + Cfg::A_BYTES;
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
          const int k0 = kb * k2BK;
          if constexpr (!A_MN) {
            tma_load_2d_2sm(sa, tma_a, &full_bar[stage], k0, m0);  // box {64 k, 128 m}
          } else {
#pragma unroll
            for (int c = 0; c < 2; ++c) tma_load_2d_2sm(sa + c * 8192, tma_a, &full_bar[stage], m0 + 64 * c, k0);
          }
          if constexpr (!B_MN) {
            tma_load_2d_2sm(sb, tma_b, &full_bar[stage], k0, n0);  // box {64 k, BN/2 n}
          } else {
#pragma unroll
            for (int c = 0; c < BN / 128; ++c)
              tma_load_2d_2sm(sb + c * 8192, tma_b, &full_bar[stage], n0 + 64 * c, k0);
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    if (leader && lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(256, BN, A_MN ? 1u : 0u, B_MN ? 1u : 0u);
      constexpr uint32_t A_LBO = A_MN ? k2BK * 128 : 0, A_KSTEP = A_MN ? 2048 : 32;
      constexpr uint32_t B_LBO = B_MN ? k2BK * 128 : 0, B_KSTEP = B_MN ? 2048 : 32;
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        const Gemm2Params& p = G.prob[find_problem<NP>(G, tile)];
        int m_blk, n_blk, ks;
        tile2_coords(p, tile - p.tile_begin, m_blk, n_blk, ks);
        const int kb0 = ks * p.kb_per_split;
        const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
        mbar_wait_cluster(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (tracing && tile == pair && kb == kb0) G.trace[4] = globaltimer_ns();
          const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint32_t sb = sa + Cfg::A_BYTES;
#pragma unroll
          for (int kk = 0; kk < k2BK / 16; ++kk) {
            const uint64_t da = umma_smem_desc(sa + kk * A_KSTEP, A_LBO, 1024);
            const uint64_t db = umma_smem_desc(sb + kk * B_KSTEP, B_LBO, 1024);
            umma_bf16_ss_2sm(tmem_d, da, db, idesc, (kb > kb0 || kk > 0) ? 1u : 0u);
          }
          umma_commit_2sm(&empty_bar[stage]);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit_2sm(&tfull_bar[acc]);
        if (tracing && tile == pair) G.trace[5] = globaltimer_ns();
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..9, both CTAs)
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    uint8_t* stg = epi_smem + (warp - 2) * 4096;
    const uint32_t stg_s = smem_u32(stg);
    bool store_pending = false;  // warp-uniform: a bulk tile store of this warp may still be reading `stg`
    int acc = 0;
    uint32_t acc_phase = 0;

    const uint32_t leader_tempty0 = mapa_u32(&tempty_bar[0], 0), leader_tempty1 = mapa_u32(&tempty_bar[1], 0);
    grid_dep_wait();  // bias / mask reads and the output writes below; hidden behind the first tile's main loop

    for (int tile = pair; tile < num_tiles; tile += num_pairs) {
      const int pi = find_problem<NP>(G, tile);
      const Gemm2Params& p = G.prob[pi];
      const CUtensorMap* tma_d = &G.maps[3 * pi + 2];
    const int flags = p.flags;
    const float alpha = p.alpha;
    const bool d_f32 = p.d_f32 != 0;
    const bool d_tma = p.d_tma != 0;
    const bool tma_f32 = p.d_tma >= 2;  // fp32: one 32-column chunk (128 B per row) per tile store
    const bool tma_red = p.d_tma == 3;  // ... as a bulk reduce-add (split-K / gradient accumulation)
    // TMA mode: chunk pairs (64 columns) are staged as a 32-row x 128 B tile in the 128B-swizzle layout (16 B unit u
    // of row r lives at unit u ^ (r & 7): conflict-free for row-per-lane writes) and written with one bulk tile store
    auto issue_store = [&](int row0, int col_even) {
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        if (tma_red) tma_reduce_add_2d(tma_d, stg, col_even, row0);
        else tma_store_2d(tma_d, stg, col_even, row0);
        bulk_commit_group();
      }
      store_pending = true;
    };
    // K6 state of this thread's row in this tile
    float ce_m = -INFINITY, ce_s = 0.f, ce_t = 0.f, ce_g = 0.f, ce_lse = 0.f, ce_w = 0.f;
    int ce_gold = -1;
    float tk_v[CE == 3 ? 8 : 1];   // K8: this thread's row, this column half: the 8 largest so far, descending
    int tk_i[CE == 3 ? 8 : 1];
    auto finish_chunk = [&](uint32_t (&r)[32], int row, bool row_ok, int col0, float bias_lane, const uint4 (&mk)[4],
                            int cpar) {
      if (col0 >= p.N) {
        if (d_tma && !tma_f32 && cpar == 1 && col0 - 32 < p.N) issue_store(row - lane, col0 - 32);
        return;
      }
      const bool full_cols = (col0 + 32 <= p.N);
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) * alpha;
      if constexpr (CE == 1) {
        // running {max, sum exp, sum, gold logit} over this thread's columns (four chains each)
        float cm[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY}, ct[4] = {0.f, 0.f, 0.f, 0.f};
        if (full_cols) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            cm[j & 3] = fmaxf(cm[j & 3], v[j]);
            ct[j & 3] += v[j];
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const bool on = col0 + j < p.N;
            v[j] = on ? v[j] : -INFINITY;  // exp(-inf - m) = 0 below
            cm[j & 3] = fmaxf(cm[j & 3], v[j]);
            ct[j & 3] += on ? v[j] : 0.f;
          }
        }
        ce_t += (ct[0] + ct[1]) + (ct[2] + ct[3]);
        const int gj = ce_gold - col0;
        if (gj >= 0 && gj < 32) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j == gj) ce_g = v[j];
        }
        const float nm = fmaxf(ce_m, fmaxf(fmaxf(cm[0], cm[1]), fmaxf(cm[2], cm[3])));
        const float nm2 = nm * 1.4426950408889634f;
        float acc4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < 32; ++j) acc4[j & 3] += ex2_(fmaf(v[j], 1.4426950408889634f, -nm2));
        ce_s = ce_s * __expf(ce_m - nm) + ((acc4[0] + acc4[1]) + (acc4[2] + acc4[3]));
        ce_m = nm;
        return;
      }
      if constexpr (CE == 3) {
        if (p.ce_p != 1.f) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = v[j] / p.ce_p;   // the beam kernels' logits / temperature
        }
        float cm[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          if (!full_cols && col0 + j >= p.N) v[j] = -INFINITY;   // exp(-inf - m) = 0; never a candidate
          cm[j & 3] = fmaxf(cm[j & 3], v[j]);
        }
        const float nm = fmaxf(ce_m, fmaxf(fmaxf(cm[0], cm[1]), fmaxf(cm[2], cm[3])));
        const float nm2 = nm * 1.4426950408889634f;
        float acc4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < 32; ++j) acc4[j & 3] += ex2_(fmaf(v[j], 1.4426950408889634f, -nm2));
        ce_s = ce_s * __expf(ce_m - nm) + ((acc4[0] + acc4[1]) + (acc4[2] + acc4[3]));
        ce_m = nm;
        const int gj = p.ce_skip - col0;   // the soft-max statistics above keep the column, the candidates do not
        if (gj >= 0 && gj < 32) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j == gj) v[j] = -INFINITY;
        }
        // sorted insertion without branches: columns arrive in ascending order and `>` is strict, so equal values
        // keep the lower column first (tf.nn.top_k's order); -inf never displaces the empty slots
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float x = v[j];
          int xi = col0 + j;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const bool gt = x > tk_v[c];
            const float ov = tk_v[c];
            const int oi = tk_i[c];
            tk_v[c] = gt ? x : ov;
            tk_i[c] = gt ? xi : oi;
            x = gt ? ov : x;
            xi = gt ? oi : xi;
          }
        }
        return;
      }
      if constexpr (CE == 2) {
        const float l2 = ce_lse * 1.4426950408889634f, qw = p.ce_q * ce_w;
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaf(ex2_(fmaf(v[j], 1.4426950408889634f, -l2)), ce_w, -qw);
        const int gj = ce_gold - col0;
        if (gj >= 0 && gj < 32) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j == gj) v[j] -= (p.ce_p - p.ce_q) * ce_w;
        }
      }
      if (CE == 0 && (flags & ZB_EPI_BIAS)) {
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += __shfl_sync(0xffffffffu, bias_lane, j);
      }
      if (CE == 0 && (flags & ZB_EPI_RELU)) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
      }
      if (!row_ok && !d_tma) return;
      if (CE == 0 && (flags & ZB_EPI_RELU_MASK) && row_ok) {
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
      if (tma_f32) {
        if (store_pending) {
          if (lane == 0) bulk_wait_read_all();
          __syncwarp();
        }
        const uint32_t rowp = stg_s + lane * 128;
#pragma unroll
        for (int q = 0; q < 8; ++q)
          sts128(rowp + ((q ^ (lane & 7)) << 4), __float_as_uint(v[q * 4]), __float_as_uint(v[q * 4 + 1]),
                 __float_as_uint(v[q * 4 + 2]), __float_as_uint(v[q * 4 + 3]));
        issue_store(row - lane, col0);
      } else if (d_tma) {
        if (cpar == 0 && store_pending) {  // the previous tile store must have drained the staging buffer
          if (lane == 0) bulk_wait_read_all();
          __syncwarp();
        }
        const uint32_t rowp = stg_s + lane * 128;
#pragma unroll
        for (int q = 0; q < 4; ++q)
          sts128(rowp + (((cpar * 4 + q) ^ (lane & 7)) << 4), pack_bf16x2(v[q * 8 + 0], v[q * 8 + 1]),
                 pack_bf16x2(v[q * 8 + 2], v[q * 8 + 3]), pack_bf16x2(v[q * 8 + 4], v[q * 8 + 5]),
                 pack_bf16x2(v[q * 8 + 6], v[q * 8 + 7]));
        if (cpar == 1) issue_store(row - lane, col0 - 32);
      } else if (d_f32) {
        float* drow = reinterpret_cast<float*>(p.d) + static_cast<long long>(row) * p.ldd + col0;
        if (flags & ZB_EPI_ACCUM) {
          if (full_cols) {
#pragma unroll
            for (int q = 0; q < 8; ++q) red_add_v4_(drow + q * 4, v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
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
    auto prefetch_chunk = [&](int row, bool row_ok, int col0, float& bias_lane, uint4 (&mk)[4]) {
      bias_lane = 0.f;
      if constexpr (CE != 0) return;  // the cross-entropy epilogues take no bias / mask
      if ((flags & ZB_EPI_BIAS) && col0 + lane < p.N) bias_lane = __ldg(p.bias + col0 + lane);
      if ((flags & ZB_EPI_RELU_MASK) && row_ok && col0 + 32 <= p.N) {
        const uint4* mrow = reinterpret_cast<const uint4*>(p.mask + static_cast<long long>(row) * p.ldmask + col0);
#pragma unroll
        for (int q = 0; q < 4; ++q) mk[q] = __ldg(mrow + q);
      }
    };
      int m_blk, n_blk, ks;
      tile2_coords(p, tile - p.tile_begin, m_blk, n_blk, ks);
      const int m0 = m_blk * 256 + (int)rank * 128, n0 = n_blk * BN;
      const int row = m0 + quad * 32 + lane;
      const bool row_ok = row < p.M;
      float bias_a, bias_b;
      uint4 mk_a[4], mk_b[4];
      constexpr int NCH = BN / 64;
      const int c0 = half * NCH;
      if constexpr (CE == 3) {
        ce_m = -INFINITY;
        ce_s = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          tk_v[c] = -INFINITY;
          tk_i[c] = 0x7fffffff;
        }
      } else if constexpr (CE != 0) {
        ce_m = -INFINITY;
        ce_s = ce_t = ce_g = 0.f;
        ce_gold = row_ok ? __ldg(p.ce_labels + row) : -1;
        if (CE == 2 && row_ok) {
          const float2 aux = __ldg(reinterpret_cast<const float2*>(p.ce_aux) + row);
          ce_lse = aux.x;
          ce_w = aux.y;
        }
      }
      prefetch_chunk(row, row_ok, n0 + c0 * 32, bias_a, mk_a);
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t tbase = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BN + c0 * 32;
      uint32_t ra[32], rb[32];
      if (tracing && tile == pair && warp == 2 && lane == 0) G.trace[6] = globaltimer_ns();
      tmem_ld_32x32b_x32(tbase, ra);
#pragma unroll 1
      const bool tr = tracing && tile == pair && warp == 2 && lane == 0;
      for (int c = 0; c < NCH; c += 2) {
        __syncwarp();
        tmem_ld_wait();
        if (tr && c == 0) G.trace[10] = globaltimer_ns();
        if (c + 1 < NCH) {
          tmem_ld_32x32b_x32(tbase + (c + 1) * 32, rb);
          prefetch_chunk(row, row_ok, n0 + (c0 + c + 1) * 32, bias_b, mk_b);
        }
        finish_chunk(ra, row, row_ok, n0 + (c0 + c) * 32, bias_a, mk_a, 0);
        if (tr && c == 0) G.trace[11] = globaltimer_ns();
        if (c + 1 < NCH) {
          __syncwarp();
          tmem_ld_wait();
          if (tr && c == 0) G.trace[12] = globaltimer_ns();
          if (c + 2 < NCH) {
            tmem_ld_32x32b_x32(tbase + (c + 2) * 32, ra);
            prefetch_chunk(row, row_ok, n0 + (c0 + c + 2) * 32, bias_a, mk_a);
          }
          finish_chunk(rb, row, row_ok, n0 + (c0 + c + 1) * 32, bias_b, mk_b, 1);
          if (tr && c == 0) G.trace[13] = globaltimer_ns();
        }
      }
      if (CE == 1 && row_ok)
        p.ce_stats[(static_cast<long long>(n_blk) * 2 + half) * p.M + row] = make_float4(ce_m, ce_s, ce_t, ce_g);
      if constexpr (CE == 3) {
        if (row_ok) {
          const long long slot = (static_cast<long long>(n_blk) * 2 + half) * p.M + row;
          p.ce_stats[slot] = make_float4(ce_m, ce_s, 0.f, 0.f);
          float4* cv = reinterpret_cast<float4*>(p.ce_cval + slot * 8);
          int4* ci = reinterpret_cast<int4*>(p.ce_cidx + slot * 8);
          cv[0] = make_float4(tk_v[0], tk_v[1], tk_v[2], tk_v[3]);
          cv[1] = make_float4(tk_v[4], tk_v[5], tk_v[6], tk_v[7]);
          ci[0] = make_int4(tk_i[0], tk_i[1], tk_i[2], tk_i[3]);
          ci[1] = make_int4(tk_i[4], tk_i[5], tk_i[6], tk_i[7]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(acc == 0 ? leader_tempty0 : leader_tempty1);
      if (tracing && tile == pair && warp == 2 && lane == 0) G.trace[7] = globaltimer_ns();
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
    // the tile stores of this warp have finished reading the staging buffer before the CTA exits (their global
    // writes complete with the grid, like CUTLASS' tma_store_wait<0>)
    if (lane == 0) bulk_wait_read_all();
    if (tracing && warp == 2 && lane == 0) G.trace[8] = globaltimer_ns();
  }

  tc_fence_before();
  cluster_sync_all();  // the peer may still signal this CTA's barriers / read its shared memory until here
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, Cfg::TMEM_COLS);
  }
  if (tracing && threadIdx.x == 0) G.trace[9] = globaltimer_ns();
}

template <int BN, bool A_MN, bool B_MN, int NP, int CE = 0>
static int launch2(const Gemm2Group<NP>& G, int grid, cudaStream_t st) {
  auto kern = gemm2_bf16_tcgen05<BN, A_MN, B_MN, NP, CE>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Gemm2Cfg<BN>::SMEM_BYTES);
    if (e != cudaSuccess) {
      set_error("gemm2 cudaFuncSetAttribute(smem=%d): %s", Gemm2Cfg<BN>::SMEM_BYTES, cudaGetErrorString(e));
      return ZB_ECUDA;
    }
    attr_done = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(k2Threads);
  cfg.dynamicSmemBytes = Gemm2Cfg<BN>::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  cudaError_t le = cudaLaunchKernelEx(&cfg, kern, G);
  if (le != cudaSuccess) {
    set_error("zb_gemm (cta pair) launch: %s", cudaGetErrorString(le));
    return ZB_ECUDA;
  }
  note_path(ZB_PATH_GEMM_PAIR);
  return check_launch("zb_gemm(2cta)");
}

template <int BN>
static int dispatch2(int a_mn, int b_mn, const Gemm2Group<1>& G, int grid, cudaStream_t st) {
  if (!a_mn && !b_mn) return launch2<BN, false, false, 1>(G, grid, st);
  if (!a_mn && b_mn) return launch2<BN, false, true, 1>(G, grid, st);
  if (a_mn && !b_mn) return launch2<BN, true, false, 1>(G, grid, st);
  return launch2<BN, true, true, 1>(G, grid, st);
}

// Whether the CTA-pair kernel should take this problem.  ZB_GEMM2=0 disables it, ZB_GEMM2=1 forces it whenever legal.
bool gemm2_wanted(const zb_gemm_args* a) {
  static const char* env = getenv("ZB_GEMM2");
  if (env && env[0] == '0') return false;
  // ZB_GEMM2_MIN_M: smallest m the pair kernel takes (default 512).  A few hundred rows make only n / 128 pair
  // tiles (8 CTAs for the 256 x 512 projections of a decode step), each bound by its own SM's L2 ingest (~0.27 us per
  // k-block); the single-CTA kernel's 128 x 64 tiles spread the same problem over 2 - 4x as many SMs: decode step
  // 0.646 -> 0.604 ms at BASELINE configs[2] (profiles/r01_decode_ab_v3.jsonl).
  static const char* min_m_env = getenv("ZB_GEMM2_MIN_M");
  const long long min_m = min_m_env ? atoll(min_m_env) : 512;
  if (a->n < 128 || a->m < min_m) return false;
  return true;
}

// Split-K choice for `tiles` output tiles of `kb_total` k-blocks each.  Cost model fitted to measurements on the
// weight-gradient shapes (tools/gemm_selftest --one ... with ZB_FORCE_SPLITS): a wave of CTA pairs costs
// ~0.27 us per k-block plus ~1.5 us of fill / drain, so e.g. 16 tiles take 4 splits (one wave of 64) rather than
// 9 (two waves of 72).
static int choose_splits(long long tiles, int kb_total, int pairs_hw) {
  static const char* force_splits = getenv("ZB_FORCE_SPLITS");  // calibration
  if (force_splits) return atoi(force_splits);
  int max_splits = kb_total / 4;
  if (max_splits < 1) max_splits = 1;
  if (max_splits > 64) max_splits = 64;
  int best_s = 1;
  double best = 1e30;
  for (int s = 1; s <= max_splits; ++s) {
    const long long waves = (tiles * s + pairs_hw - 1) / pairs_hw;
    const double cost = (double)waves * (((kb_total + s - 1) / s) * 0.27 + 1.5);
    if (cost < best - 1e-9) {
      best = cost;
      best_s = s;
    }
  }
  return best_s;
}

// Fills the tile geometry (everything but the split fields) and the three tensor maps of one problem.
static int setup_problem(const zb_gemm_args* a, int bn, Gemm2Params& p, CUtensorMap* maps) {
  p.M = (int)a->m; p.N = (int)a->n; p.K = (int)a->k;
  p.d = a->d; p.ldd = a->ldd; p.bias = a->bias;
  p.mask = reinterpret_cast<const __nv_bfloat16*>(a->mask); p.ldmask = a->ldmask;
  p.alpha = a->alpha; p.flags = a->flags; p.d_f32 = a->d_dtype == ZB_F32;
  p.ce_mode = 0; p.ce_aux = nullptr; p.ce_labels = nullptr; p.ce_stats = nullptr; p.ce_p = 1.f; p.ce_q = 0.f;
  p.ce_cval = nullptr; p.ce_cidx = nullptr; p.ce_skip = -1;
  p.kb_total = (p.K + k2BK - 1) / k2BK;
  p.mt = (p.M + 255) / 256;
  p.nt = (p.N + bn - 1) / bn;
  p.gn = bn == 256 ? 6 : 8;
  if (p.gn > p.nt) p.gn = p.nt;
  p.tile_begin = 0;
  const bool accum = a->flags & ZB_EPI_ACCUM;
  int rc;
  const bool a_mn = a->a_layout == ZB_MN_MAJOR, b_mn = a->b_layout == ZB_MN_MAJOR;
  if (!a_mn) rc = make_map(&maps[0], a->a, p.K, p.M, a->lda, 128);
  else rc = make_map(&maps[0], a->a, p.M, p.K, a->lda, k2BK);
  if (rc) return rc;
  if (!b_mn) rc = make_map(&maps[1], a->b, p.K, p.N, a->ldb, bn / 2);
  else rc = make_map(&maps[1], a->b, p.N, p.K, a->ldb, k2BK);
  if (rc) return rc;
  // outputs with 16 B-aligned rows go through shared memory + TMA tile stores / bulk reduce-adds
  static const char* no_tma_d = getenv("ZB_GEMM_NO_TMA_STORE");
  static const char* no_tma_red = getenv("ZB_GEMM_NO_TMA_REDUCE");
  maps[2] = maps[0];
  p.d_tma = 0;
  const bool d_al = a->d != nullptr && (reinterpret_cast<uintptr_t>(a->d) & 15) == 0;   // no output tensor: K6 mode 1
  if (accum && p.d_f32 && !no_tma_d && !no_tma_red && d_al && (a->ldd * 4) % 16 == 0) {
    rc = make_map_f32(&maps[2], a->d, p.N, p.M, a->ldd, 32);
    if (rc) return rc;
    p.d_tma = 3;
  }
  if (!accum && !no_tma_d && d_al) {
    if (!p.d_f32 && (a->ldd * 2) % 16 == 0) {
      rc = make_map(&maps[2], a->d, p.N, p.M, a->ldd, 32);
      if (rc) return rc;
      p.d_tma = 1;
    } else if (p.d_f32 && (a->ldd * 4) % 16 == 0) {
      rc = make_map_f32(&maps[2], a->d, p.N, p.M, a->ldd, 32);
      if (rc) return rc;
      p.d_tma = 2;
    }
  }
  return ZB_OK;
}

static void set_splits(Gemm2Params& p, int splits) {
  if (splits > p.kb_total) splits = p.kb_total;
  if (splits < 1) splits = 1;
  p.kb_per_split = (p.kb_total + splits - 1) / splits;
  p.splits = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;
}

int gemm2_launch(const zb_gemm_args* a, cudaStream_t st) {
  const bool accum = a->flags & ZB_EPI_ACCUM;
  const int pairs_hw = num_sms_compute() / 2;
  const int mt = (int)((a->m + 255) / 256);
  int bn = 256;
  auto tiles_for = [&](int b) { return (long long)mt * ((a->n + b - 1) / b); };
  if (a->n <= 128) bn = 128;
  if (!accum && bn == 256) {
    // wave quantisation: cost ~ (#waves of pair tiles) x (tile width); e.g. N = 1536 at M = 4096 is 96 tiles of 256
    // (2 waves of 74 pairs) but 192 tiles of 128 (3 waves of half the work each)
    // The k-loop of these kernels runs at the L2 -> SM ingest cap, where a tile costs its operand bytes
    // (256 + bn) * K * 2 rather than its flops (tools/gemm_l2_model.py), so a wave of 256-wide tiles costs 512 / 384
    // of a wave of 128-wide ones, not 2x: +1.2 % tokens/s at configs[1] against the flop-weighted rule
    // (profiles/r02a_summary.txt).  ZB_GEMM2_TILE_MODEL=flops restores the latter.
    static const bool l2_model = !(getenv("ZB_GEMM2_TILE_MODEL") != nullptr && getenv("ZB_GEMM2_TILE_MODEL")[0] == 'f');
    const long long c256 = l2_model ? 256 + 256 : 256, c128 = l2_model ? 256 + 128 : 128;
    const long long w256 = (tiles_for(256) + pairs_hw - 1) / pairs_hw * c256;
    const long long w128 = (tiles_for(128) + pairs_hw - 1) / pairs_hw * c128;
    if (w128 < w256) bn = 128;
  }
  static const char* force_bn = getenv("ZB_GEMM2_BN");   // calibration: 128 / 256 for every problem wide enough
  if (force_bn && a->n > 128) bn = atoi(force_bn) == 128 ? 128 : 256;
  Gemm2Group<1> G;
  Gemm2Params& p = G.prob[0];
  int rc = setup_problem(a, bn, p, G.maps);
  if (rc) return rc;
  const long long tiles = (long long)p.mt * p.nt;
  int splits = a->split_k;
  if (splits <= 0) splits = (accum && tiles < pairs_hw) ? choose_splits(tiles, p.kb_total, pairs_hw) : 1;
  if (splits > 1 && !accum) {
    set_error("zb_gemm: split_k > 1 requires ZB_EPI_ACCUM");
    return ZB_EINVAL;
  }
  set_splits(p, splits);
  const long long total = tiles * p.splits;
  G.count = 1;
  G.total_tiles = (int)total;
  G.trace = nullptr;
  static const bool trace_on = getenv("ZB_GEMM_TRACE") != nullptr;
  static unsigned long long* trace_buf = nullptr;
  if (trace_on) {
    if (!trace_buf) cudaMalloc(&trace_buf, 16 * sizeof(unsigned long long));
    cudaMemsetAsync(trace_buf, 0, 16 * sizeof(unsigned long long), st);
    G.trace = trace_buf;
  }
  const int grid = 2 * (int)(total < pairs_hw ? total : pairs_hw);
  const bool a_mn = a->a_layout == ZB_MN_MAJOR, b_mn = a->b_layout == ZB_MN_MAJOR;
  const int lrc = bn == 128 ? dispatch2<128>(a_mn, b_mn, G, grid, st) : dispatch2<256>(a_mn, b_mn, G, grid, st);
  if (trace_on && lrc == ZB_OK) {
    unsigned long long h[16];
    cudaStreamSynchronize(st);
    cudaMemcpy(h, trace_buf, sizeof(h), cudaMemcpyDeviceToHost);
    auto d = [&](int i) { return h[i] ? (long long)(h[i] - h[0]) : -1ll; };
    fprintf(stderr,
            "[zb_gemm2 trace] m=%d n=%d k=%d bn=%d splits=%d grid=%d tiles=%lld | ns since entry: setup %lld, dep_wait %lld, "
            "first tma %lld, first data %lld, tile0 mma issued %lld, tile0 acc ready %lld, tile0 epilogue done %lld, "
            "stores drained %lld, exit %lld | epilogue warp 2: ld0 done %lld, chunk0 staged %lld, ld1 done %lld, "
            "chunk1 staged + store issued %lld\n",
            p.M, p.N, p.K, bn, p.splits, grid, total, d(1), d(2), d(3), d(4), d(5), d(6), d(7), d(8), d(9), d(10), d(11),
            d(12), d(13));
  }
  return lrc;
}

// K6: logits = a @ b^T (both K-major, 256-wide tiles) with one of the cross-entropy epilogues (vocab_ce.cu).
// mode 1 needs no output tensor (a->d is ignored); mode 2 writes bf16 d_logits to a->d.
int gemm2_launch_ce(const zb_gemm_args* a, int mode, const float* aux, const int32_t* labels, float4* stats, float ce_p,
                    float ce_q, cudaStream_t st) {
  const int pairs_hw = num_sms_compute() / 2;
  Gemm2Group<1> G;
  Gemm2Params& p = G.prob[0];
  zb_gemm_args b = *a;
  if (mode == 1) b.d = nullptr;  // nothing is stored: no output tensor map
  int rc = setup_problem(&b, 256, p, G.maps);
  if (rc) return rc;
  if (mode == 1) p.d_tma = 0;
  else if (p.d_tma != 1) {
    set_error("zb_vocab_ce: d_logits must be 16-byte aligned with a pitch that is a multiple of 8 elements");
    return ZB_EINVAL;
  }
  p.ce_mode = mode; p.ce_aux = aux; p.ce_labels = labels; p.ce_stats = stats; p.ce_p = ce_p; p.ce_q = ce_q;
  set_splits(p, 1);
  const long long total = (long long)p.mt * p.nt;
  G.count = 1;
  G.total_tiles = (int)total;
  G.trace = nullptr;
  const int grid = 2 * (int)(total < pairs_hw ? total : pairs_hw);
  return mode == 1 ? launch2<256, false, false, 1, 1>(G, grid, st) : launch2<256, false, false, 1, 2>(G, grid, st);
}

// K8: logits = a @ b^T / temperature reduced to per-(row, 128-column half tile) soft-max statistics and top-8 candidates
// (vocab_topk.cu); a->d is ignored.
int gemm2_launch_topk(const zb_gemm_args* a, float4* stats, float* cval, int32_t* cidx, int skip_col, float temperature,
                      cudaStream_t st) {
  const int pairs_hw = num_sms_compute() / 2;
  Gemm2Group<1> G;
  Gemm2Params& p = G.prob[0];
  zb_gemm_args b = *a;
  b.d = nullptr;
  int rc = setup_problem(&b, 256, p, G.maps);
  if (rc) return rc;
  p.d_tma = 0;
  p.ce_mode = 3; p.ce_stats = stats; p.ce_cval = cval; p.ce_cidx = cidx; p.ce_skip = skip_col; p.ce_p = temperature;
  set_splits(p, 1);
  const long long total = (long long)p.mt * p.nt;
  G.count = 1;
  G.total_tiles = (int)total;
  G.trace = nullptr;
  const int grid = 2 * (int)(total < pairs_hw ? total : pairs_hw);
  return launch2<256, false, false, 1, 3>(G, grid, st);
}

// One launch for several accumulate-into-fp32 problems with MN-major operands (the weight gradients of a layer).
// Returns ZB_OK after launching, a negative status on error, or +1 when the set does not qualify (the caller then
// launches the problems one by one).
int gemm2_launch_group(const zb_gemm_args* args, int count, cudaStream_t st) {
  static const char* off = getenv("ZB_NO_GEMM_GROUP");
  if (off || count < 2 || count > k2MaxGroup) return 1;
  for (int i = 0; i < count; ++i) {
    const zb_gemm_args* a = &args[i];
    if (a->a_layout != ZB_MN_MAJOR || a->b_layout != ZB_MN_MAJOR || !(a->flags & ZB_EPI_ACCUM) ||
        a->d_dtype != ZB_F32 || a->split_k > 0 || a->n < 256 || a->m < 256 || !gemm2_wanted(a))
      return 1;
  }
  const int pairs_hw = num_sms_compute() / 2;
  Gemm2Group<k2MaxGroup> G;
  long long tiles = 0;
  int kb_min = 1 << 30, kb_max = 0;
  for (int i = 0; i < count; ++i) {
    int rc = setup_problem(&args[i], 256, G.prob[i], &G.maps[3 * i]);
    if (rc) return rc;
    tiles += (long long)G.prob[i].mt * G.prob[i].nt;
    kb_min = G.prob[i].kb_total < kb_min ? G.prob[i].kb_total : kb_min;
    kb_max = G.prob[i].kb_total > kb_max ? G.prob[i].kb_total : kb_max;
  }
  int splits = choose_splits(tiles, kb_max, pairs_hw);
  if (splits > kb_min / 4) splits = kb_min / 4 > 0 ? kb_min / 4 : 1;
  long long total = 0;
  for (int i = 0; i < count; ++i) {
    set_splits(G.prob[i], splits);
    G.prob[i].tile_begin = (int)total;
    total += (long long)G.prob[i].mt * G.prob[i].nt * G.prob[i].splits;
  }
  for (int i = count; i < k2MaxGroup; ++i) {
    G.prob[i] = G.prob[0];
    G.prob[i].tile_begin = 1 << 30;
  }
  G.count = count;
  G.total_tiles = (int)total;
  G.trace = nullptr;
  const int grid = 2 * (int)(total < pairs_hw ? total : pairs_hw);
  return launch2<256, true, true, k2MaxGroup>(G, grid, st);
}

}  // namespace zb
