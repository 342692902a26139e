// attention.cuh — fused multi-head self-attention for the encoder (K3 of SURVEY.md §2.3).
//
//   ctx = softmax(Q K^T / sqrt(64) + (1 - mask) * -10000) V         per (sequence, head)
//
// replaces HF RobertaSelfAttention / BertSelfAttention (transformers==2.3.0, eager) as reached from
// model/models.py:150-151,188-189,237-238.  The additive -10000 mask is reproduced literally
// (bias added in fp32 before the softmax), so an all-padding sequence gives the same finite
// "softmax of the raw scores" the reference gives, not NaN.
//
// One work item = (128-token query tile, head).  Per 128-key block of the same sequence:
//   S = Q K^T          tcgen05.mma  M128 N128 K16 x4   (Q, K: TMA boxes of the [tokens, 3H] QKV buffer)
//   online softmax     4 warps, thread = query row: two passes over S in TMEM (max, then exp2),
//                      P written as bf16 into shared memory in the K-major SWIZZLE_128B layout
//   O_blk = P V        tcgen05.mma  M128 N64 K16 x8    (V is the MN-major B operand)
//   o = o*alpha + O_blk in registers; after the last block ctx = o / l  (bf16)
// Sequence lengths: a multiple of 128, or a divisor of 128 (then a tile holds 128/L sequences and
// cross-sequence scores are excluded).  head_dim is fixed at 64 (BERT/RoBERTa-base).
#pragma once
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
  __nv_bfloat16* ctx;  // [n_tokens, hidden]
  float scale_log2;    // log2(e) / sqrt(64)
};

struct Smem {
  static constexpr int kQ = 0;
  static constexpr int kStages = 1;                            // K/V stages (1 keeps two CTAs resident per SM)
  static constexpr int kK = kQ + kTile * kDh * 2;
  static constexpr int kV = kK + kStages * kTile * kDh * 2;
  static constexpr int kP = kV + kStages * kTile * kDh * 2;    // 128 x 128 bf16 (two 64-key halves)
  static constexpr int kBias = kP + kTile * kTile * 2;       // 128 floats
  static constexpr int kBar = kBias + kTile * 4 + 16;
  // q_full q_empty kv_full[2] kv_empty[2] s p o  + tmem ptr
  static constexpr int kTotal = kBar + 9 * 8 + 16;
  static constexpr int kDynamic = kTotal + 1024;
};

template <bool kPacked>  // kPacked: L < 128, a tile holds 128/L sequences
__global__ void __launch_bounds__(256, 2)
attention_kernel(const __grid_constant__ CUtensorMap tmQKV, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Smem::kBar);
  uint64_t* q_full = bars + 0;
  uint64_t* q_empty = bars + 1;
  uint64_t* kv_full = bars + 2;
  uint64_t* kv_empty = bars + 4;
  uint64_t* bar_s = bars + 6;
  uint64_t* bar_p = bars + 7;
  uint64_t* bar_o = bars + 8;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 9);
  float* sbias = reinterpret_cast<float*>(smem + Smem::kBias);
  unsigned* smask = reinterpret_cast<unsigned*>(smem + Smem::kBias + kTile * 4);  // per-warp ballots of masked keys

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = (p.n_tokens + kTile - 1) / kTile;
  const int total_work = n_tiles * p.heads;
  const int nkv = (p.L >= kTile) ? p.L / kTile : 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    for (int s = 0; s < Smem::kStages; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    mbar_init(bar_s, 1);
    mbar_init(bar_p, 4);
    mbar_init(bar_o, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<1>(tmem_ptr_smem, 256);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const uint32_t tmem_S = tmem_base;        // 128 columns
  const uint32_t tmem_O = tmem_base + 128;  // 64 columns

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      Ring<Smem::kStages> kv;
      uint32_t wk = 0;
      for (int w = blockIdx.x; w < total_work; w += gridDim.x, ++wk) {
        const int tile = w / p.heads, h = w - tile * p.heads;
        const int tok0 = tile * kTile;
        const int kv_tok0 = (p.L >= kTile) ? (tok0 / p.L) * p.L : tok0;
        mbar_wait(q_empty, (wk & 1) ^ 1, 10);
        mbar_arrive_expect_tx(q_full, kTile * kDh * 2);
        tma_load_2d(smem + Smem::kQ, &tmQKV, q_full, h * kDh, tok0, kEvictFirst);
        for (int j = 0; j < nkv; ++j) {
          mbar_wait(&kv_empty[kv.stage], kv.phase ^ 1, 11);
          mbar_arrive_expect_tx(&kv_full[kv.stage], 2 * kTile * kDh * 2);
          tma_load_2d(smem + Smem::kK + kv.stage * kTile * kDh * 2, &tmQKV, &kv_full[kv.stage],
                      p.hidden + h * kDh, kv_tok0 + j * kTile, kEvictNormal);
          tma_load_2d(smem + Smem::kV + kv.stage * kTile * kDh * 2, &tmQKV, &kv_full[kv.stage],
                      2 * p.hidden + h * kDh, kv_tok0 + j * kTile, kEvictNormal);
          kv.advance();
        }
      }
    }
  } else if (warp == 1) {
    // ================================= MMA issuer =================================
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_f16(kTile, kTile, kFmtBF16, 0, 0);  // Q K^T : both K-major
      constexpr uint32_t idesc_o = make_idesc_f16(kTile, kDh, kFmtBF16, 0, 1);    // P V   : V is MN-major
      Ring<Smem::kStages> kv;
      uint32_t wk = 0, it = 0;
      for (int w = blockIdx.x; w < total_work; w += gridDim.x, ++wk) {
        mbar_wait(q_full, wk & 1, 12);
        for (int j = 0; j < nkv; ++j, ++it) {
          mbar_wait(&kv_full[kv.stage], kv.phase, 13);
          tc_fence_after_sync();
          const uint32_t sq = smem_u32(smem + Smem::kQ);
          const uint32_t sk = smem_u32(smem + Smem::kK + kv.stage * kTile * kDh * 2);
          const uint32_t sv = smem_u32(smem + Smem::kV + kv.stage * kTile * kDh * 2);
#pragma unroll
          for (int k = 0; k < kDh / 16; ++k)
            umma_ss<1>(tmem_S, make_desc_k_sw128(sq + k * 32), make_desc_k_sw128(sk + k * 32), idesc_s, k != 0);
          umma_commit(bar_s);
          if (j == nkv - 1) umma_commit(q_empty);
          mbar_wait(bar_p, it & 1, 14);
          tc_fence_after_sync();
          const uint32_t sp = smem_u32(smem + Smem::kP);
#pragma unroll
          for (int k = 0; k < kTile / 16; ++k) {
            // A = P: keys [16k, 16k+16) live in 64-key half (k/4), 32 bytes per K step inside the span
            const uint64_t adesc = make_desc_k_sw128(sp + (k >> 2) * (kTile * 128) + (k & 3) * 32);
            // B = V (MN-major): 16 keys = two 8-row groups of 1024 bytes
            const uint64_t bdesc = make_desc_mn_sw128(sv + k * 2048, kTile * 128, 1024);
            umma_ss<1>(tmem_O, adesc, bdesc, idesc_o, k != 0);
          }
          umma_commit(&kv_empty[kv.stage]);
          umma_commit(bar_o);
          kv.advance();
        }
      }
    }
  } else if (warp >= 4) {
    // =============================== softmax / output ==============================
    const int quad = warp & 3;
    const int row = quad * 32 + lane;          // query row inside the tile
    const int tid128 = (warp - 4) * 32 + lane;
    const uint32_t lane_sel = static_cast<uint32_t>(quad * 32) << 16;
    uint8_t* sP = smem + Smem::kP;
    uint32_t it = 0;
    for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
      const int tile = w / p.heads, h = w - tile * p.heads;
      const int tok0 = tile * kTile;
      const int kv_tok0 = (p.L >= kTile) ? (tok0 / p.L) * p.L : tok0;
      const int seq_lo = (p.L >= kTile) ? 0 : (row / p.L) * p.L;   // keys of this row's own sequence
      const int seq_hi = (p.L >= kTile) ? kTile : seq_lo + p.L;
      float m_run = -INFINITY, l_run = 0.f;
      float o[kDh];
#pragma unroll
      for (int i = 0; i < kDh; ++i) o[i] = 0.f;
      for (int j = 0; j < nkv; ++j, ++it) {
        // key bias of this block -> smem (previous block's readers are past their last use: they
        // all arrived on bar_p before the PV MMA whose completion we waited for below)
        {
          const int kt = kv_tok0 + j * kTile + tid128;
          const float bv = (kt < p.n_tokens) ? __ldg(p.kbias + kt) : -INFINITY;
          sbias[tid128] = bv;
          const unsigned mk = __ballot_sync(0xffffffffu, bv < 0.f);
          if (lane == 0) smask[warp - 4] = mk;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        // no key of this block carries a mask bias (and the tile is one sequence): the bias terms vanish
        const bool plain = !kPacked && ((smask[0] | smask[1] | smask[2] | smask[3]) == 0u);
        mbar_wait(bar_s, it & 1, 15);
        tc_fence_after_sync();
        // pass 1: row max (of the raw scores when `plain`: scale > 0 commutes with max)
        float m_blk = -INFINITY;
#pragma unroll 1
        for (int c = 0; c < kTile; c += 32) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(tmem_S + lane_sel + c, v);
          tmem_ld_wait();
          if (plain) {
#pragma unroll
            for (int i = 0; i < 32; ++i) m_blk = fmaxf(m_blk, __uint_as_float(v[i]));
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const int key = c + i;
              float t = fmaf(__uint_as_float(v[i]), p.scale_log2, sbias[key]);
              if (kPacked && (key < seq_lo || key >= seq_hi)) t = -INFINITY;
              m_blk = fmaxf(m_blk, t);
            }
          }
        }
        if (plain) m_blk *= p.scale_log2;
        const float m_new = fmaxf(m_run, m_blk);
        const float alpha = (m_run == -INFINITY) ? 0.f : exp2f(m_run - m_new);
        // pass 2: p = exp2(t - m_new), row sum, bf16 P into swizzled smem
        float rsum = 0.f;
#pragma unroll 1
        for (int c = 0; c < kTile; c += 32) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(tmem_S + lane_sel + c, v);
          tmem_ld_wait();
          uint32_t pk[16];
          if (plain) {
            const float nm = -m_new;
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              const float p0 = ex2_ftz(fmaf(__uint_as_float(v[i]), p.scale_log2, nm));
              const float p1 = ex2_ftz(fmaf(__uint_as_float(v[i + 1]), p.scale_log2, nm));
              rsum += p0 + p1;
              const __nv_bfloat162 h2 = __floats2bfloat162_rn(p0, p1);
              pk[i >> 1] = *reinterpret_cast<const uint32_t*>(&h2);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              float t0 = fmaf(__uint_as_float(v[i]), p.scale_log2, sbias[c + i]);
              float t1 = fmaf(__uint_as_float(v[i + 1]), p.scale_log2, sbias[c + i + 1]);
              if (kPacked && (c + i < seq_lo || c + i >= seq_hi)) t0 = -INFINITY;
              if (kPacked && (c + i + 1 < seq_lo || c + i + 1 >= seq_hi)) t1 = -INFINITY;
              const float p0 = ex2_ftz(t0 - m_new), p1 = ex2_ftz(t1 - m_new);
              rsum += p0 + p1;
              const __nv_bfloat162 h2 = __floats2bfloat162_rn(p0, p1);
              pk[i >> 1] = *reinterpret_cast<const uint32_t*>(&h2);
            }
          }
          // four 16-byte chunks (8 keys each); swizzle: chunk index ^= (row & 7) inside the 128-byte span
          const int half = c >> 6;                 // which 64-key half
          const int chunk0 = (c & 63) >> 3;        // first 16-byte chunk of these 32 keys within the span
          uint8_t* rowp = sP + half * (kTile * 128) + row * 128;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int ch = (chunk0 + q) ^ (row & 7);
            *reinterpret_cast<uint4*>(rowp + ch * 16) = make_uint4(pk[q * 4], pk[q * 4 + 1], pk[q * 4 + 2], pk[q * 4 + 3]);
          }
        }
        fence_proxy_async_smem();
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_p);
        // O_blk
        mbar_wait(bar_o, it & 1, 16);
        tc_fence_after_sync();
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
      const int tok = tok0 + row;
      if (tok < p.n_tokens) {
        const float inv = 1.0f / l_run;
        __nv_bfloat16* out = p.ctx + static_cast<size_t>(tok) * p.hidden + h * kDh;
#pragma unroll
        for (int q = 0; q < kDh / 8; ++q) {
          uint4 u;
          __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
          for (int t = 0; t < 4; ++t) h2[t] = __floats2bfloat162_rn(o[q * 8 + t * 2] * inv, o[q * 8 + t * 2 + 1] * inv);
          reinterpret_cast<uint4*>(out)[q] = u;
        }
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) tmem_dealloc<1>(tmem_base, 256);
}

}  // namespace attn
