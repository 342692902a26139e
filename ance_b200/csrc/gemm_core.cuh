// gemm_core.cuh — the one tcgen05 mainloop of this repo.
//
//   D[M,N] = A[M,K] * B[N,K]^T        A, B: 16-bit (bf16 or fp16), K-major, fp32 accumulate in TMEM
//
// Used by (i) the encoder's linear layers (x * W^T, W stored [out,in] as in the checkpoint) and
// (ii) the coarse pass of the flat inner-product search (Q * P^T).  The two differ only in the
// epilogue functor `Ep` (bias/GELU/residual store vs. per-query running top-k).
//
// Structure (one CTA per SM, persistent, warp-specialised):
//   warp 0   : TMA producer   — streams 128x64 A and (BN/CG)x64 B tiles through a STAGES-deep smem ring
//   warp 1   : MMA issuer     — one thread issues tcgen05.mma (M = 128*CG, N = BN, K = 16) x 4 per k-block
//   warp 2   : TMEM allocator — 2 accumulator stages of BN fp32 columns each (double buffered)
//   warp 3   : idle
//   warps 4+ : epilogue       — tcgen05.ld the accumulator, run Ep, release the TMEM stage
// CG = 2 pairs two CTAs (cluster 2x1x1) on one 256-row tile: each CTA loads its own 128 A rows and
// half of the B rows, the leader CTA issues the MMAs, commits are multicast to both CTAs.
//
// A "work item" is (m_blk, split): one M tile swept over a contiguous range of N blocks.  Plain
// GEMMs use one N block per work item; the search sweeps thousands, carrying top-k state.
#pragma once
#include "tc05.cuh"

namespace gemm {

using namespace tc05;

constexpr int BM = 128;  // rows per CTA
constexpr int BK = 64;   // 64 x 16-bit = one 128-byte swizzle span
constexpr int UMMA_K = 16;

struct WorkShape {
  int M, N, K;
  int num_m_blks;        // ceil(M / (BM*CG))
  int num_n_blks;        // ceil(N / BN)
  int n_splits;          // work items per m block
  int n_blks_per_split;  // ceil(num_n_blks / n_splits)
  unsigned long long hint_a, hint_b;  // L2 cache-policy words of the A / B tile loads (0 = the epilogue's default)
  // Soft barrier between CTA pairs that sweep the SAME B rows (search, n_splits == 1): pace[c] = progress of cluster c
  // in tiles; a producer does not run more than pace_window tiles ahead of the slowest cluster, so that a B tile fetched
  // from HBM by the first pair is still in L2 when the last pair asks for it.  null = off.
  int* pace;
  int pace_window;
  int pace_stride;   // clusters c, c + stride, c + 2 stride, ... sweep the same rows (1: all of them; n_splits: one wave of
                     // (query tile, row range) items, item w on cluster w, range = w % n_splits)
};

// Publish this cluster's progress and wait (bounded: ~100 us, then go on regardless — the barrier is a bandwidth
// optimisation, never a correctness requirement, and must not hang if the grid is not fully co-resident).
static __device__ __noinline__ void pace_wait(int* pace, int window, int me, int n_clusters, int stride, int seq) {
  volatile int* vp = pace;
  vp[me] = seq;
  const long long t0 = clock64();
  for (;;) {
    int mn = 0x7fffffff;
    for (int c = me % stride; c < n_clusters; c += stride) mn = min(mn, vp[c]);
    if (seq - mn <= window || clock64() - t0 > 200000ll) break;
  }
}

struct EpiCtx {
  int m_blk, split;
  int nb0, nb1;     // n-block range of this work item
  int row0;         // first global row of this CTA's 128-row tile
  int quad;         // TMEM lane quadrant of this warp (warp_idx % 4)
  int epi_warp;     // 0 .. EPI_WARPS-1
  int lane;
  int work_seq;     // how many work items this CTA has processed before this one
  uint8_t* ep_smem; // Ep::kSmemBytes of shared memory owned by the epilogue (1024-B aligned)
};

template <int BN, int STAGES, int CG, int EP_SMEM = 0>
struct SmemPlan {
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBRows = BN / CG;
  static constexpr int kBBytes = kBRows * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kRingBytes = STAGES * kStageBytes;
  static constexpr int kEpOffset = kRingBytes;
  static constexpr int kBarOffset = kRingBytes + EP_SMEM;
  // full[STAGES] empty[STAGES] tmem_full[2] tmem_empty[2] + tmem ptr
  static constexpr int kBarBytes = (2 * STAGES + 4) * 8 + 16;
  static constexpr int kTotal = kBarOffset + kBarBytes;
  static constexpr int kDynamicBytes = kTotal + 1024;  // slack for manual 1024-B alignment
};

template <class Ep, int BN, int STAGES, int CG, int EPI_WARPS, uint32_t FMT>
__global__ void __launch_bounds__(128 + 32 * EPI_WARPS, 1)
tc05_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const WorkShape ws, const __grid_constant__ typename Ep::Params ep) {
  static_assert(BN % 32 == 0 && BN >= 32 && BN <= 256, "BN");
  static_assert(CG == 1 || CG == 2, "CG");
  using Plan = SmemPlan<BN, STAGES, CG, Ep::kSmemBytes>;
  constexpr uint32_t kTmemCols = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Plan::kBarOffset);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = (CG == 2) ? cluster_ctarank() : 0u;
  const bool leader = (cta_rank == 0);
  const int cluster_id = (CG == 2) ? (blockIdx.x >> 1) : blockIdx.x;
  const int num_clusters = (CG == 2) ? (gridDim.x >> 1) : gridDim.x;
  const int total_work = ws.num_m_blks * ws.n_splits;
  const int num_kb = (ws.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);   // the leader CTA's producer arrives once and expects both CTAs' bytes
      mbar_init(&empty_bar[s], 1);  // one tcgen05.commit
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], CG * EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<CG>(tmem_ptr_smem, kTmemCols);
  tc_fence_before_sync();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      const uint64_t hint_a = ws.hint_a ? ws.hint_a : Ep::kHintA, hint_b = ws.hint_b ? ws.hint_b : Ep::kHintB;
      Ring<STAGES> ring;
      for (int w = cluster_id; w < total_work; w += num_clusters) {
        const int m_blk = w / ws.n_splits, split = w - m_blk * ws.n_splits;
        const int nb0 = split * ws.n_blks_per_split;
        const int nb1 = min(nb0 + ws.n_blks_per_split, ws.num_n_blks);
        const int a_row = (m_blk * CG + (int)cta_rank) * BM;
        for (int nb = nb0; nb < nb1; ++nb) {
          if (ws.pace != nullptr && leader && ((nb - nb0) & 7) == 0)
            pace_wait(ws.pace, ws.pace_window, cluster_id, num_clusters, ws.pace_stride,
                      ((w - cluster_id) / num_clusters) * ws.num_n_blks + (nb - nb0));
          const int b_row = nb * BN + (int)cta_rank * Plan::kBRows;
          for (int kb = 0; kb < num_kb; ++kb) {
            mbar_wait(&empty_bar[ring.stage], ring.phase ^ 1, 1);
            uint8_t* sa = smem + ring.stage * Plan::kStageBytes;
            uint8_t* sb = sa + Plan::kABytes;
            if constexpr (CG == 1) {
              mbar_arrive_expect_tx(&full_bar[ring.stage], Plan::kStageBytes);
              tma_load_2d(sa, &tmA, &full_bar[ring.stage], kb * BK, a_row, hint_a);
              tma_load_2d(sb, &tmB, &full_bar[ring.stage], kb * BK, b_row, hint_b);
            } else {
              // Both CTAs' TMA bytes complete on the LEADER's barrier; only the leader arrives on it.  The
              // peer cannot run a phase ahead: it refills a stage only after the MMA that consumed it
              // (multicast commit on its own empty barrier).
              if (leader) mbar_arrive_expect_tx(&full_bar[ring.stage], 2 * Plan::kStageBytes);
              tma_load_2d_2sm(sa, &tmA, &full_bar[ring.stage], kb * BK, a_row, hint_a);
              tma_load_2d_2sm(sb, &tmB, &full_bar[ring.stage], kb * BK, b_row, hint_b);
            }
            ring.advance();
          }
        }
      }
      if (ws.pace != nullptr && leader) *reinterpret_cast<volatile int*>(ws.pace + cluster_id) = 0x7fffffff;   // done: never the slowest
    }
  } else if (warp == 1) {
    // ====================================== MMA issuer ======================================
    if (lane == 0 && leader) {
      constexpr uint32_t idesc = make_idesc_f16(BM * CG, BN, FMT, 0, 0);
      Ring<STAGES> ring;
      Ring<2> acc;
      for (int w = cluster_id; w < total_work; w += num_clusters) {
        const int m_blk = w / ws.n_splits, split = w - m_blk * ws.n_splits;
        const int nb0 = split * ws.n_blks_per_split;
        const int nb1 = min(nb0 + ws.n_blks_per_split, ws.num_n_blks);
        for (int nb = nb0; nb < nb1; ++nb) {
          mbar_wait(&tempty_bar[acc.stage], acc.phase ^ 1, 2);
          tc_fence_after_sync();
          const uint32_t tmem_d = tmem_base + acc.stage * BN;
          for (int kb = 0; kb < num_kb; ++kb) {
            mbar_wait(&full_bar[ring.stage], ring.phase, 3);
            tc_fence_after_sync();
            const uint32_t sa = smem_u32(smem + ring.stage * Plan::kStageBytes);
            const uint32_t sb = sa + Plan::kABytes;
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              const uint64_t adesc = make_desc_k_sw128(sa + k * UMMA_K * 2);
              const uint64_t bdesc = make_desc_k_sw128(sb + k * UMMA_K * 2);
              umma_ss<CG>(tmem_d, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
            }
            if constexpr (CG == 1) umma_commit(&empty_bar[ring.stage]);
            else umma_commit_2sm(&empty_bar[ring.stage], 0x3);
            ring.advance();
          }
          if constexpr (CG == 1) umma_commit(&tfull_bar[acc.stage]);
          else umma_commit_2sm(&tfull_bar[acc.stage], 0x3);
          acc.advance();
        }
      }
    }
  } else if (warp >= 4) {
    // ======================================= epilogue =======================================
    Ep epi;
    EpiCtx cx;
    cx.quad = warp & 3;
    cx.epi_warp = warp - 4;
    cx.lane = lane;
    cx.work_seq = 0;
    cx.ep_smem = smem + Plan::kEpOffset;
    Ring<2> acc;
    for (int w = cluster_id; w < total_work; w += num_clusters, ++cx.work_seq) {
      cx.m_blk = w / ws.n_splits;
      cx.split = w - cx.m_blk * ws.n_splits;
      cx.nb0 = cx.split * ws.n_blks_per_split;
      cx.nb1 = min(cx.nb0 + ws.n_blks_per_split, ws.num_n_blks);
      cx.row0 = (cx.m_blk * CG + (int)cta_rank) * BM;
      epi.begin_work(ep, ws, cx);
      for (int nb = cx.nb0; nb < cx.nb1; ++nb) {
        mbar_wait(&tfull_bar[acc.stage], acc.phase, 4);
        tc_fence_after_sync();
        const uint32_t tacc = tmem_base + acc.stage * BN + (static_cast<uint32_t>(cx.quad * 32) << 16);
        epi.tile(ep, ws, cx, tacc, nb);
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) {
          if (CG == 1 || leader) mbar_arrive(&tempty_bar[acc.stage]);
          else mbar_arrive_cluster(&tempty_bar[acc.stage], 0);
        }
        acc.advance();
      }
      epi.end_work(ep, ws, cx);
    }
    epi.end_kernel(ep, cx);
  }

  tc_fence_before_sync();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 2) tmem_dealloc<CG>(tmem_base, kTmemCols);
}

// ------------------------------------------------------------------------------------------------
// host-side launch helper
// ------------------------------------------------------------------------------------------------
constexpr int kMaxDevices = 64;

inline int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return (dev >= 0 && dev < kMaxDevices) ? dev : 0;
}

// SM count of the CURRENT device (cached per device ordinal: one process may drive several GPUs through the C ABI)
inline int sm_count() {
  static int n[kMaxDevices] = {};
  const int dev = current_device();
  if (!n[dev]) cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, dev);
  return n[dev];
}

template <class Ep, int BN, int STAGES, int CG, int EPI_WARPS, uint32_t FMT>
cudaError_t launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const WorkShape& ws,
                   const typename Ep::Params& ep, int max_ctas, cudaStream_t stream) {
  using Plan = SmemPlan<BN, STAGES, CG, Ep::kSmemBytes>;
  auto kern = tc05_gemm_kernel<Ep, BN, STAGES, CG, EPI_WARPS, FMT>;
  static bool configured[kMaxDevices] = {};   // the opt-in is per device, not per process
  const int dev = current_device();
  if (!configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Plan::kDynamicBytes);
    if (e != cudaSuccess) return e;
    configured[dev] = true;
  }
  const int total = ws.num_m_blks * ws.n_splits;
  int clusters = min(total, (max_ctas > 0 ? max_ctas : sm_count()) / CG);
  if (clusters < 1) clusters = 1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(clusters * CG);
  cfg.blockDim = dim3(128 + 32 * EPI_WARPS);
  cfg.dynamicSmemBytes = Plan::kDynamicBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, tmA, tmB, ws, ep);
}

inline WorkShape make_shape(int M, int N, int K, int BN, int CG, int n_splits /* <=0: one N block per item */) {
  WorkShape ws;
  ws.M = M;
  ws.N = N;
  ws.K = K;
  ws.num_m_blks = (M + BM * CG - 1) / (BM * CG);
  ws.num_n_blks = (N + BN - 1) / BN;
  if (n_splits <= 0 || n_splits > ws.num_n_blks) n_splits = ws.num_n_blks;
  ws.n_blks_per_split = (ws.num_n_blks + n_splits - 1) / n_splits;
  ws.n_splits = (ws.num_n_blks + ws.n_blks_per_split - 1) / ws.n_blks_per_split;
  ws.hint_a = ws.hint_b = 0;
  ws.pace = nullptr;
  ws.pace_window = 0;
  ws.pace_stride = 1;
  return ws;
}

}  // namespace gemm
