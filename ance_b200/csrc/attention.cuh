// attention.cuh — fused multi-head self-attention for the encoder (K3 of SURVEY.md §2.3).
//
//   ctx = softmax(Q K^T / sqrt(64) + (1 - mask) * -10000) V         per (sequence, head)
//
// replaces HF RobertaSelfAttention / BertSelfAttention (transformers==2.3.0, eager) as reached from
// model/models.py:150-151,188-189,237-238.  The additive -10000 mask is reproduced literally
// (bias added in fp32 before the softmax), so an all-padding sequence gives the same finite
// "softmax of the raw scores" the reference gives, not NaN.
//
// One work item = (128-token query tile, head); a persistent CTA per SM keeps two items in flight, one per
// softmax warpgroup (see attention_kernel).  Per 128-key block of the same sequence:
//   S = Q K^T          tcgen05.mma  M128 N128 K16 x4   (Q, K: TMA boxes of the [tokens, 3H] QKV buffer)
//   softmax            4 warps per group, thread = query row: L <= 128 one trip to TMEM (row in registers: max, exp2),
//                      L > 128 two passes per block with online rescaling;
//                      P written as 16-bit (FMT) into shared memory in the K-major SWIZZLE_128B layout
//   O_blk = P V        tcgen05.mma  M128 N64 K16 x8    (V is the MN-major B operand)
//   o = o*alpha + O_blk in registers; after the last block ctx = o / l  (16-bit)
// Sequence lengths: a multiple of 128, or a divisor of 128 (then a tile holds 128/L sequences and
// cross-sequence scores are excluded).  head_dim is fixed at 64 (BERT/RoBERTa-base).
#pragma once
#include "act16.cuh"
#include "tc05.cuh"

namespace attn {

using namespace tc05;

constexpr int kTile = 128;   // query rows / keys per block
constexpr int kDh = 64;      // head dim

__device__ __forceinline__ float ex2_ftz(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

struct Params {
  int n_tokens;        // B * L
  int L;               // sequence length
  int heads;
  int hidden;          // heads * 64
  const float* kbias;  // [n_tokens] additive key bias * log2(e): 0 or -10000*log2e
  float scale_log2;    // log2(e) / sqrt(64)
  // Variable-length packing (kPacked kernels only; null = uniform L): token row r of a 128-row tile belongs to a sequence
  // that occupies the tile-local rows [row_lo[r], row_hi[r]) — whole sequences of ANY length <= 128 share a tile, nothing is
  // padded inside a sequence (kbias is all zero), rows after the last sequence of a tile attend to themselves only.
  const uint8_t* row_lo;
  const uint8_t* row_hi;
};

struct Smem {
  static constexpr int kGroups = 2;                            // softmax warpgroups, each with its own Q / S / P / O
  static constexpr int kStages = 4;                            // (K, V) stages shared by both groups
  static constexpr int kTileBytes = kTile * kDh * 2;           // 16 KB: one 128 x 64 16-bit operand tile
  static constexpr int kQ = 0;
  static constexpr int kKV = kQ + kGroups * kTileBytes;        // stage s: K at +0, V at +16 KB
  static constexpr int kP = kKV + kStages * 2 * kTileBytes;    // per group 128 x 128 16-bit (two 64-key halves)
  static constexpr int kBias = kP + kGroups * kTile * kTile * 2;   // per group 128 floats + 4 ballots
  static constexpr int kBiasStride = kTile * 4 + 16;
  static constexpr int kBar = kBias + kGroups * kBiasStride;
  // q_full[2] q_empty[2] kv_full[S] kv_empty[S] s[2] p[2] o[2]  + tmem ptr
  static constexpr int kNumBars = 4 + 2 * kStages + 6;
  static constexpr int kTotal = kBar + kNumBars * 8 + 16;
  static constexpr int kDynamic = kTotal + 1024;
  static_assert(kDynamic <= 232448, "attention smem exceeds 227 KB");
};

// a running maximum above this is the score of an unmasked key: masked keys sit near -10000 log2(e) = -14427
constexpr float kRealMax = -7000.0f;

constexpr int kThreads = 384;   // warp 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle, 4-7 softmax group 0, 8-11 softmax group 1

// Persistent, one CTA per SM.  The CTA walks its work items w = blockIdx.x + n * gridDim.x; item n belongs to
// softmax group n & 1, so two items are always in flight: while one group runs its softmax on the CUDA cores the
// tensor core serves the other, and the TMA producer runs up to kStages key/value blocks ahead (the HBM latency
// of a 48 KB Q/K/V fetch is longer than one item's arithmetic).  Blocks are issued in the interleaved order
// (a,0) (b,0) (a,1) (b,1) ... for the item pair (a, b); producer, MMA issuer and both groups derive that order
// from the same loop nest.
// kPacked: L < 128, a tile holds 128/L sequences.  kSingle: L <= 128, one key block per item.
// FMT: 16-bit format of Q / K / V, of the probabilities P and of the output (act16.cuh).
template <bool kPacked, bool kSingle, uint32_t FMT>
__global__ void __launch_bounds__(kThreads, 1)
attention_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmCTX, const Params p) {
  static_assert(kSingle || !kPacked, "a packed tile has a single key block");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Smem::kBar);
  uint64_t* q_full = bars + 0;                       // [2]
  uint64_t* q_empty = bars + 2;                      // [2]
  uint64_t* kv_full = bars + 4;                      // [kStages]
  uint64_t* kv_empty = bars + 4 + Smem::kStages;     // [kStages]
  uint64_t* bar_s = bars + 4 + 2 * Smem::kStages;    // [2]
  uint64_t* bar_p = bar_s + 2;                       // [2]
  uint64_t* bar_o = bar_s + 4;                       // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + Smem::kNumBars);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = (p.n_tokens + kTile - 1) / kTile;
  const int total_work = n_tiles * p.heads;
  const int nkv = (p.L >= kTile) ? p.L / kTile : 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmCTX);
    for (int g = 0; g < Smem::kGroups; ++g) {
      mbar_init(&q_full[g], 1);
      mbar_init(&q_empty[g], 1);
      mbar_init(&bar_s[g], 1);
      mbar_init(&bar_p[g], 4);
      mbar_init(&bar_o[g], 1);
    }
    for (int s = 0; s < Smem::kStages; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<1>(tmem_ptr_smem, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr_smem;
  // group g: S at columns [128 g, 128 g + 128), O at [256 + 64 g, 256 + 64 g + 64)

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      Ring<Smem::kStages> kv;
      uint32_t qc[2] = {0, 0};
      for (int wa = blockIdx.x; wa < total_work; wa += 2 * gridDim.x) {
        const int n_in_pair = (wa + static_cast<int>(gridDim.x) < total_work) ? 2 : 1;
        for (int j = 0; j < nkv; ++j) {
          for (int g = 0; g < n_in_pair; ++g) {
            const int w = wa + g * gridDim.x;
            const int tile = w / p.heads, h = w - tile * p.heads;
            const int tok0 = tile * kTile;
            const int kv_tok0 = (p.L >= kTile) ? (tok0 / p.L) * p.L : tok0;
            if (j == 0) {
              mbar_wait(&q_empty[g], (qc[g] & 1) ^ 1, 10);
              ++qc[g];
              mbar_arrive_expect_tx(&q_full[g], Smem::kTileBytes);
              tma_load_2d(smem + Smem::kQ + g * Smem::kTileBytes, &tmQKV, &q_full[g], h * kDh, tok0, kEvictFirst);
            }
            mbar_wait(&kv_empty[kv.stage], kv.phase ^ 1, 11);
            mbar_arrive_expect_tx(&kv_full[kv.stage], 2 * Smem::kTileBytes);
            uint8_t* st = smem + Smem::kKV + kv.stage * 2 * Smem::kTileBytes;
            tma_load_2d(st, &tmQKV, &kv_full[kv.stage], p.hidden + h * kDh, kv_tok0 + j * kTile, kEvictNormal);
            tma_load_2d(st + Smem::kTileBytes, &tmQKV, &kv_full[kv.stage], 2 * p.hidden + h * kDh, kv_tok0 + j * kTile,
                        kEvictNormal);
            kv.advance();
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================= MMA issuer =================================
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_f16(kTile, kTile, FMT, 0, 0);  // Q K^T : both K-major
      constexpr uint32_t idesc_o = make_idesc_f16(kTile, kDh, FMT, 0, 1);    // P V   : V is MN-major
      Ring<Smem::kStages> kv;
      uint32_t qc[2] = {0, 0}, bc[2] = {0, 0};
      int prev_g = -1, prev_stage = 0;   // block whose S is issued and whose P V is still owed
      auto issue_pv = [&](int g, int stage) {
        mbar_wait(&bar_p[g], bc[g] & 1, 14);
        ++bc[g];
        tc_fence_after_sync();
        const uint32_t sp = smem_u32(smem + Smem::kP + g * (kTile * kTile * 2));
        const uint32_t sv = smem_u32(smem + Smem::kKV + stage * 2 * Smem::kTileBytes + Smem::kTileBytes);
        const uint32_t tmem_O = tmem_base + 256 + g * kDh;
#pragma unroll
        for (int k = 0; k < kTile / 16; ++k) {
          // A = P: keys [16k, 16k+16) live in 64-key half (k/4), 32 bytes per K step inside the span
          const uint64_t adesc = make_desc_k_sw128(sp + (k >> 2) * (kTile * 128) + (k & 3) * 32);
          // B = V (MN-major): 16 keys = two 8-row groups of 1024 bytes
          const uint64_t bdesc = make_desc_mn_sw128(sv + k * 2048, kTile * 128, 1024);
          umma_ss<1>(tmem_O, adesc, bdesc, idesc_o, k != 0);
        }
        umma_commit(&kv_empty[stage]);
        umma_commit(&bar_o[g]);
      };
      for (int wa = blockIdx.x; wa < total_work; wa += 2 * gridDim.x) {
        const int n_in_pair = (wa + static_cast<int>(gridDim.x) < total_work) ? 2 : 1;
        for (int j = 0; j < nkv; ++j) {
          for (int g = 0; g < n_in_pair; ++g) {
            // S of group g may be overwritten only after the group's previous block published its P
            if (prev_g == g) { issue_pv(prev_g, prev_stage); prev_g = -1; }
            if (j == 0) { mbar_wait(&q_full[g], qc[g] & 1, 12); ++qc[g]; }
            mbar_wait(&kv_full[kv.stage], kv.phase, 13);
            tc_fence_after_sync();
            const uint32_t sq = smem_u32(smem + Smem::kQ + g * Smem::kTileBytes);
            const uint32_t sk = smem_u32(smem + Smem::kKV + kv.stage * 2 * Smem::kTileBytes);
#pragma unroll
            for (int k = 0; k < kDh / 16; ++k)
              umma_ss<1>(tmem_base + g * kTile, make_desc_k_sw128(sq + k * 32), make_desc_k_sw128(sk + k * 32), idesc_s, k != 0);
            umma_commit(&bar_s[g]);
            if (j == nkv - 1) umma_commit(&q_empty[g]);
            if (prev_g >= 0) issue_pv(prev_g, prev_stage);
            prev_g = g;
            prev_stage = kv.stage;
            kv.advance();
          }
        }
      }
      if (prev_g >= 0) issue_pv(prev_g, prev_stage);
    }
  } else if (warp >= 4) {
    // =============================== softmax / output ==============================
    const int g = (warp - 4) >> 2;             // softmax group
    const int quad = warp & 3;                 // TMEM lane quarter this warp may read
    const int row = quad * 32 + lane;          // query row inside the tile
    const uint32_t lane_sel = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t tmem_S = tmem_base + g * kTile;
    const uint32_t tmem_O = tmem_base + 256 + g * kDh;
    uint8_t* sP = smem + Smem::kP + g * (kTile * kTile * 2);
    float* sbias = reinterpret_cast<float*>(smem + Smem::kBias + g * Smem::kBiasStride);
    unsigned* smask = reinterpret_cast<unsigned*>(sbias + kTile);  // per-warp ballots of masked keys
    const int bar_id = 1 + g;
    // additive key bias of block (w2, j2) for key `row` of that block (fetched one block ahead of its use)
    auto load_bias = [&](int w2, int j2) -> float {
      if (w2 >= total_work) return 0.f;
      const int t0 = (w2 / p.heads) * kTile;
      const int kt = ((p.L >= kTile) ? (t0 / p.L) * p.L : t0) + j2 * kTile + row;
      return (kt < p.n_tokens) ? __ldg(p.kbias + kt) : -INFINITY;
    };
    // P (and the output tile) rows in shared memory: 128-byte spans, 16-byte chunk index ^= (row & 7)  (SWIZZLE_128B)
    auto store_chunks = [&](uint8_t* rowp, int chunk0, const uint32_t* pk) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int ch = (chunk0 + q) ^ (row & 7);
        *reinterpret_cast<uint4*>(rowp + ch * 16) = make_uint4(pk[q * 4], pk[q * 4 + 1], pk[q * 4 + 2], pk[q * 4 + 3]);
      }
    };
    uint32_t it = 0;
    const int w_first = blockIdx.x + g * gridDim.x;
    float bv = load_bias(w_first, 0);
    for (int w = w_first; w < total_work; w += 2 * gridDim.x) {
      const int tile = w / p.heads, h = w - tile * p.heads;
      const int tok0 = tile * kTile;
      const bool varlen = kPacked && p.row_lo != nullptr;
      int seq_lo = (p.L >= kTile) ? 0 : (row / p.L) * p.L;   // keys of this row's own sequence
      int seq_hi = (p.L >= kTile) ? kTile : seq_lo + p.L;
      if (varlen) {
        seq_lo = __ldg(p.row_lo + tok0 + row);
        seq_hi = __ldg(p.row_hi + tok0 + row);
      }
      float m_run = -INFINITY, l_run = 0.f;
      float o[kDh];
      if constexpr (!kSingle) {
#pragma unroll
        for (int i = 0; i < kDh; ++i) o[i] = 0.f;
      }
      for (int j = 0; j < nkv; ++j, ++it) {
        // key bias of this block -> smem (the previous block's readers are past their last use: they all arrived
        // on bar_p before the PV MMA whose completion this thread has waited for)
        sbias[row] = bv;
        {
          const unsigned mk = __ballot_sync(0xffffffffu, bv < 0.f);
          if (lane == 0) smask[quad] = mk;
        }
        if (row == 0) bulk_wait_read_all();   // the previous item's output tile has left sP
        named_bar_sync(bar_id, kTile);
        bv = (j + 1 < nkv) ? load_bias(w, j + 1) : load_bias(w + 2 * gridDim.x, 0);
        // Per 32-key chunk, warp-uniform:  0 = every p is exactly 0 (keys of another packed sequence, or all keys
        // masked while the row has an unmasked key somewhere: exp2(-10000 log2e + s - m) flushes to zero, as
        // exp(-10000 + s - m) does in the reference's fp32 softmax),  1 = no key masked (no bias term),  2 = general.
        const unsigned mk[4] = {smask[0], smask[1], smask[2], smask[3]};
        // (variable-length tiles: a row whose sequence fills the whole tile takes the same arithmetic as a full-length
        // sequence of the dense L = 128 kernel, so that the two paths agree bit for bit)
        const bool plain = kPacked ? (varlen && seq_lo == 0 && seq_hi == kTile) : ((mk[0] | mk[1] | mk[2] | mk[3]) == 0u);
        int st[4];
        if (varlen) {   // per row: chunk outside / inside / straddling the boundary of the row's own sequence
#pragma unroll
          for (int c = 0; c < 4; ++c)
            st[c] = (seq_hi <= c * 32 || seq_lo >= c * 32 + 32) ? 0 : (seq_lo <= c * 32 && seq_hi >= c * 32 + 32) ? 1 : 2;
        } else {
          bool own[4], any_unmasked = false;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            own[c] = !kPacked || (p.L >= 32 ? (c * 32 >= seq_lo && c * 32 < seq_hi) : c == quad);
            any_unmasked |= own[c] && mk[c] != 0xffffffffu;
          }
          const bool skip_ok = (kPacked && p.L < 32) ? false : (any_unmasked || (!kSingle && m_run > kRealMax));
#pragma unroll
          for (int c = 0; c < 4; ++c)
            st[c] = !own[c] ? 0 : (mk[c] == 0xffffffffu && skip_ok) ? 0 : (mk[c] == 0u && !(kPacked && p.L < 32)) ? 1 : 2;
        }
        mbar_wait(&bar_s[g], it & 1, 15);
        tc_fence_after_sync();
        float rsum, alpha = 1.f, m_new;
        bool all_skip = false;
        if constexpr (kSingle) {
          // one key block per item: the whole score row lives in registers, one trip to TMEM
          uint32_t v[kTile];
#pragma unroll
          for (int c = 0; c < kTile; c += 32) tmem_ld_32x32b_x32p(tmem_S + lane_sel + c, v + c);
          tmem_ld_wait();
          float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
          if (plain) {   // max of the raw scores: scale > 0 commutes with max
#pragma unroll
            for (int i = 0; i < kTile; ++i) mx[i & 3] = fmaxf(mx[i & 3], __uint_as_float(v[i]));
          } else {
#pragma unroll
            for (int c = 0; c < kTile; c += 32) {
              if (st[c >> 5] == 1) {
#pragma unroll
                for (int i = c; i < c + 32; ++i) {
                  const float t = __uint_as_float(v[i]) * p.scale_log2;
                  v[i] = __float_as_uint(t);
                  mx[i & 3] = fmaxf(mx[i & 3], t);
                }
              } else if (st[c >> 5] == 2) {
#pragma unroll
                for (int i = c; i < c + 32; ++i) {
                  float t = fmaf(__uint_as_float(v[i]), p.scale_log2, sbias[i]);
                  if (kPacked && (i < seq_lo || i >= seq_hi)) t = -INFINITY;
                  v[i] = __float_as_uint(t);
                  mx[i & 3] = fmaxf(mx[i & 3], t);
                }
              }
            }
          }
          m_new = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
          const float sc = plain ? p.scale_log2 : 1.0f;
          if (plain) m_new *= p.scale_log2;
          const float nm = -m_new;
          float rs[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int c = 0; c < kTile; c += 32) {
            uint32_t pk[16];
            if (st[c >> 5] == 0) {
#pragma unroll
              for (int i = 0; i < 16; ++i) pk[i] = 0u;
            } else {
#pragma unroll
              for (int i = 0; i < 32; i += 2) {
                const float p0 = ex2_ftz(fmaf(__uint_as_float(v[c + i]), sc, nm));
                const float p1 = ex2_ftz(fmaf(__uint_as_float(v[c + i + 1]), sc, nm));
                rs[(i >> 1) & 3] += p0 + p1;
                pk[i >> 1] = act16::Act<FMT>::pack2(p0, p1);
              }
            }
            store_chunks(sP + (c >> 6) * (kTile * 128) + row * 128, (c & 63) >> 3, pk);
          }
          rsum = (rs[0] + rs[1]) + (rs[2] + rs[3]);
        } else {
          const int stb = st[0] | (st[1] << 2) | (st[2] << 4) | (st[3] << 6);
          all_skip = stb == 0;   // a fully masked block of a row that has real keys: contributes nothing
          // p = exp2(t - m_ref) for the whole block: row sum, 16-bit P into swizzled smem; returns max(t - m_ref)
          auto exp_pass = [&](float m_ref, float& rs_out) -> float {
            float rs_ = 0.f, mu = -INFINITY;
            const float nm = -m_ref;
#pragma unroll 1
            for (int c = 0; c < kTile; c += 32) {
              const int sc_ = (stb >> (c >> 4)) & 3;
              uint32_t pk[16];
              if (sc_ == 0) {
#pragma unroll
                for (int i = 0; i < 16; ++i) pk[i] = 0u;
              } else {
                uint32_t v[32];
                tmem_ld_32x32b_x32(tmem_S + lane_sel + c, v);
                tmem_ld_wait();
                if (sc_ == 1) {
#pragma unroll
                  for (int i = 0; i < 32; i += 2) {
                    const float u0 = fmaf(__uint_as_float(v[i]), p.scale_log2, nm);
                    const float u1 = fmaf(__uint_as_float(v[i + 1]), p.scale_log2, nm);
                    mu = fmaxf(mu, fmaxf(u0, u1));
                    const float p0 = ex2_ftz(u0), p1 = ex2_ftz(u1);
                    rs_ += p0 + p1;
                    pk[i >> 1] = act16::Act<FMT>::pack2(p0, p1);
                  }
                } else {
#pragma unroll
                  for (int i = 0; i < 32; i += 2) {
                    const float u0 = fmaf(__uint_as_float(v[i]), p.scale_log2, sbias[c + i]) + nm;
                    const float u1 = fmaf(__uint_as_float(v[i + 1]), p.scale_log2, sbias[c + i + 1]) + nm;
                    mu = fmaxf(mu, fmaxf(u0, u1));
                    const float p0 = ex2_ftz(u0), p1 = ex2_ftz(u1);
                    rs_ += p0 + p1;
                    pk[i >> 1] = act16::Act<FMT>::pack2(p0, p1);
                  }
                }
              }
              store_chunks(sP + (c >> 6) * (kTile * 128) + row * 128, (c & 63) >> 3, pk);
            }
            rs_out = rs_;
            return mu;
          };
          // Blocks after the first: ONE trip over the scores, relative to the running maximum (the block's own maximum is
          // tracked on the way).  exp2(t - m_run) stays <= 2^8 unless a score exceeds every earlier one by more than 8
          // (log2 units) — then, and only then, the warp redoes the block relative to the true maximum.  The softmax is
          // the same function either way (numerator and denominator carry the same factor 2^(m_true - m_ref)).
          const bool optimistic = __all_sync(0xffffffffu, j > 0 && m_run > kRealMax);
          if (optimistic) {
            m_new = m_run;
            alpha = 1.f;
            rsum = 0.f;
            if (!all_skip) {
              const float over = exp_pass(m_run, rsum);
              if (__any_sync(0xffffffffu, over > 8.0f)) {
                m_new = fmaxf(m_run, m_run + over);
                alpha = exp2f(m_run - m_new);
                exp_pass(m_new, rsum);
              }
            } else {
              float dummy;
              exp_pass(m_run, dummy);   // (writes the zero P tile; no TMEM reads: every chunk state is 0)
            }
          } else {
            // first block of an item (or no real key seen yet): pass 1 = row max (of the raw scores when `plain`:
            // scale > 0 commutes with max), pass 2 = exp relative to it
            float m_blk = -INFINITY;
            if (!all_skip) {
#pragma unroll 1
              for (int c = 0; c < kTile; c += 32) {
                const int sc_ = (stb >> (c >> 4)) & 3;
                if (sc_ == 0) continue;
                uint32_t v[32];
                tmem_ld_32x32b_x32(tmem_S + lane_sel + c, v);
                tmem_ld_wait();
                if (plain) {
#pragma unroll
                  for (int i = 0; i < 32; ++i) m_blk = fmaxf(m_blk, __uint_as_float(v[i]));
                } else if (sc_ == 1) {
#pragma unroll
                  for (int i = 0; i < 32; ++i) m_blk = fmaxf(m_blk, __uint_as_float(v[i]) * p.scale_log2);
                } else {
#pragma unroll
                  for (int i = 0; i < 32; ++i) m_blk = fmaxf(m_blk, fmaf(__uint_as_float(v[i]), p.scale_log2, sbias[c + i]));
                }
              }
            }
            if (plain) m_blk *= p.scale_log2;
            m_new = fmaxf(m_run, m_blk);
            alpha = (m_run == -INFINITY) ? 0.f : exp2f(m_run - m_new);
            exp_pass(m_new, rsum);
          }
        }
        fence_proxy_async_smem();
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_p[g]);
        // O_blk
        mbar_wait(&bar_o[g], it & 1, 16);
        tc_fence_after_sync();
        if constexpr (kSingle) {
          l_run = rsum;
        } else {
          if (!all_skip) {   // (a skipped block has P = 0: its O_blk is exactly 0 and alpha is exactly 1)
#pragma unroll
            for (int c = 0; c < kDh; c += 32) {
              uint32_t v[32];
              tmem_ld_32x32b_x32(tmem_O + lane_sel + c, v);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) o[c + i] = fmaf(o[c + i], alpha, __uint_as_float(v[i]));
            }
            tc_fence_before_sync();
            l_run = fmaf(l_run, alpha, rsum);
            m_run = m_new;
          }
        }
      }
      // ctx tile = o / l in 16 bits: staged in this group's P buffer (the PV MMA has finished reading it), one TMA store
      {
        const float inv = 1.0f / l_run;
        if constexpr (kSingle) {
          uint32_t v[kDh];
          tmem_ld_32x32b_x32p(tmem_O + lane_sel, v);
          tmem_ld_32x32b_x32p(tmem_O + lane_sel + 32, v + 32);
          tmem_ld_wait();
          tc_fence_before_sync();
#pragma unroll
          for (int i = 0; i < kDh; ++i) o[i] = __uint_as_float(v[i]);
        }
#pragma unroll
        for (int c = 0; c < kDh; c += 32) {
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            pk[i >> 1] = act16::Act<FMT>::pack2(o[c + i] * inv, o[c + i + 1] * inv);
          }
          store_chunks(sP + row * 128, c >> 3, pk);
        }
        fence_proxy_async_smem();
        named_bar_sync(bar_id, kTile);
        if (row == 0) {
          tma_store_2d(&tmCTX, sP, h * kDh, tok0);   // rows past n_tokens are clipped by the tensor map
          bulk_commit_group();
        }
      }
    }
    if (row == 0) bulk_wait_read_all();
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) tmem_dealloc<1>(tmem_base, 512);
}

}  // namespace attn
