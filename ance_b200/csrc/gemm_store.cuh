// gemm_store.cuh — "store" epilogue of the tcgen05 GEMM: bias, erf-GELU (two closed forms, see gelu_erf2 / gelu_logistic2), residual add, bf16
// output through shared memory + TMA store (or fp32 output by direct stores).  Covers the encoder's
// linear layers K2/K4/K5/K6/K7 of SURVEY.md §2.3 (HF RobertaSelfAttention / RobertaSelfOutput /
// RobertaIntermediate / RobertaOutput reached from model/models.py:150-151, and embeddingHead,
// models.py:152-153) and the debug GEMM used by the bring-up tests.
//
// bf16 path: each group of 4 epilogue warps (= the 128 rows of the tile) turns 64 accumulator columns
// at a time into a 128 x 64 bf16 slab in SWIZZLE_128B shared memory and one thread hands it to the TMA
// (cp.async.bulk.tensor.2d.global.shared::cta); the stores to HBM are then full 128-byte lines and run
// asynchronously under the next slab's math.  Rows / columns past the tensor edge are clipped by TMA.
#pragma once
#include "act16.cuh"
#include "gemm_core.cuh"

namespace gemm {

__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// Packed fp32 pairs (Blackwell FFMA2 / fma.rn.f32x2): the FP32 pipe, not the issue rate, bounds the GELU
// epilogue, and FFMA2 does two elements per slot at full fp32 precision.
__device__ __forceinline__ uint64_t pack2(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// HF "gelu" = x * 0.5 * (1 + erf(x / sqrt(2))) = relu(x) - 0.5 |x| erfc(|x| / sqrt(2)), with
//   erfc(z) ~= (1 + a1 z + ... + a6 z^6)^-16        (Abramowitz & Stegun 7.1.28, |error| <= 3e-7)
// and the 1/sqrt(2) folded into the coefficients.  Measured |error| of the GELU <= 7.1e-7 absolute (three
// orders of magnitude below the bf16 rounding of the output).  The epilogue of the FFN-up GEMM is bound by
// the SFU (MUFU) rate, so this form uses ONE MUFU (rcp) per element and no exponential; everything else is
// packed FFMA2/FMUL2: per PAIR of elements 14 packed fp32 ops + 2 LOP + 2 MUFU.RCP.
__device__ __forceinline__ void gelu_erf2(float& x0, float& x1) {
  const uint64_t x = pack2(x0, x1);
  const uint64_t ax = x & 0x7FFFFFFF7FFFFFFFull;
  uint64_t p = fma2(pack2(5.38297490e-06f, 5.38297490e-06f), ax, pack2(4.88906371e-05f, 4.88906371e-05f));
  p = fma2(p, ax, pack2(3.80035744e-05f, 3.80035744e-05f));
  p = fma2(p, ax, pack2(3.27762635e-03f, 3.27762635e-03f));
  p = fma2(p, ax, pack2(2.11410057e-02f, 2.11410057e-02f));
  p = fma2(p, ax, pack2(4.98673469e-02f, 4.98673469e-02f));
  p = fma2(p, ax, pack2(1.0f, 1.0f));
  p = mul2(p, p);
  p = mul2(p, p);
  p = mul2(p, p);
  p = mul2(p, p);  // (1 + ...)^16 ; overflows to +inf for |x| > ~25, whose reciprocal is the correct 0
  float p0, p1;
  unpack2(p, p0, p1);
  const uint64_t r = pack2(rcp_approx(p0), rcp_approx(p1));
  const uint64_t relu = fma2(ax, pack2(0.5f, 0.5f), mul2(x, pack2(0.5f, 0.5f)));
  unpack2(fma2(mul2(ax, pack2(-0.5f, -0.5f)), r, relu), x0, x1);
}

// The same function in logistic form:  gelu(x) = x * Phi(x) = x / (1 + exp(-g(x))),  g = logit(Phi) fitted by an odd
// degree-9 polynomial (minimax on the GELU itself over |x| <= 8; the leading coefficient is positive, so g -> +-inf
// and the form saturates correctly for any |x|).  |error| <= 3.7e-6 absolute in fp32 (bf16 rounds the result at 2^-9
// relative).  Per PAIR: 8 packed fp32 ops + 2 MUFU.EX2 + 2 MUFU.RCP  (gelu_erf2: 14 packed + 2 LOP + 2 MUFU.RCP).
__device__ __forceinline__ void gelu_logistic2(float& x0, float& x1) {
  const uint64_t x = pack2(x0, x1);
  const uint64_t t = mul2(x, x);
  // -log2(e) * (c0 + c1 t + c2 t^2 + c3 t^3 + c4 t^4)
  uint64_t p = fma2(pack2(-3.290565928e-06f, -3.290565928e-06f), t, pack2(8.930850163e-05f, 8.930850163e-05f));
  p = fma2(p, t, pack2(3.548454260e-04f, 3.548454260e-04f));
  p = fma2(p, t, pack2(-1.052178442e-01f, -1.052178442e-01f));
  p = fma2(p, t, pack2(-2.302048445e+00f, -2.302048445e+00f));
  float g0, g1;
  unpack2(mul2(p, x), g0, g1);
  const uint64_t d = fma2(pack2(ex2_approx(g0), ex2_approx(g1)), pack2(1.0f, 1.0f), pack2(1.0f, 1.0f));   // 1 + exp(-g)
  float d0, d1;
  unpack2(d, d0, d1);
  unpack2(mul2(x, pack2(rcp_approx(d0), rcp_approx(d1))), x0, x1);
}

// FMT: 16-bit format of the OUTPUT and of the residual (act16.cuh); the operand format of the GEMM itself is the
// FMT parameter of gemm::launch.
template <int BN, int EPI_WARPS, uint32_t FMT = tc05::kFmtBF16>
struct EpStore {
  using A16 = act16::Act<FMT>;
  static constexpr uint64_t kHintA = tc05::kEvictNormal;
  static constexpr uint64_t kHintB = tc05::kEvictLast;  // weights: keep in L2
  static constexpr int kColGroups = EPI_WARPS / 4;
  static constexpr int kColsPerGroup = BN / kColGroups;
  static constexpr int kSlabBytes = BM * 128;            // 128 rows x 64 bf16
  static constexpr int kSmemBytes = kColGroups * kSlabBytes + 64;  // slabs + one residual mbarrier per column group
  static_assert(EPI_WARPS % 4 == 0 && kColsPerGroup % 64 == 0, "epilogue warp layout");

  struct alignas(64) Params {
    CUtensorMap tmC;         // 16-bit output [M, N], box {64, 128}, SWIZZLE_128B (valid when C != null)
    CUtensorMap tmR;         // 16-bit residual, same geometry (valid when R != null and C != null)
    uint16_t* C;             // [M, ldc] 16-bit (FMT) or null
    float* C32;              // [M, ldc32] fp32 or null (direct stores)
    const float* bias;       // [N] or null
    const uint16_t* R;       // residual [M, ldr] (FMT) or null
    int ldc, ldc32, ldr;
    int act;                 // 0 none, 1 gelu (erfc form, |err| <= 7e-7), 2 gelu (logistic form, |err| <= 3.7e-6)
  };

  uint32_t rphase;

  __device__ __forceinline__ void begin_work(const Params& p, const WorkShape&, const EpiCtx& cx) {
    if (cx.work_seq == 0 && p.R && p.C && !p.C32) {  // residual tiles arrive by TMA: one mbarrier per column group
      const int cgi = cx.epi_warp >> 2;
      uint64_t* rbar = reinterpret_cast<uint64_t*>(cx.ep_smem + kColGroups * kSlabBytes) + cgi;
      if ((cx.epi_warp & 3) == 0 && cx.lane == 0) {
        tc05::mbar_init(rbar, 1);
        tc05::fence_barrier_init();
      }
      tc05::named_bar_sync(2 + cgi, 128);
      rphase = 0;
    }
  }
  __device__ __forceinline__ void end_work(const Params&, const WorkShape&, const EpiCtx&) {}
  __device__ __forceinline__ void end_kernel(const Params& p, const EpiCtx& cx) {
    if (p.C && (cx.epi_warp & 3) == 0 && cx.lane == 0) tc05::bulk_wait_all();
  }

  // bias + activation + residual on one 32-column chunk held as fp32
  __device__ __forceinline__ void finish_chunk(const Params& p, const WorkShape& ws, float (&f)[32], int row,
                                               bool row_ok, int col0, bool direct_residual) {
    const bool full = (col0 + 32 <= ws.N);
    if (p.bias) {
      if (full) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col0) + j);
          f[j * 4] += b.x; f[j * 4 + 1] += b.y; f[j * 4 + 2] += b.z; f[j * 4 + 3] += b.w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (col0 + i < ws.N) f[i] += __ldg(p.bias + col0 + i);
      }
    }
    if (p.act == 1) {
#pragma unroll
      for (int i = 0; i < 32; i += 2) gelu_erf2(f[i], f[i + 1]);
    } else if (p.act == 2) {
#pragma unroll
      for (int i = 0; i < 32; i += 2) gelu_logistic2(f[i], f[i + 1]);
    }
    if (direct_residual && p.R && row_ok) {
      const uint16_t* r = p.R + (size_t)row * p.ldr + col0;
      if (full) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint4 q = __ldg(reinterpret_cast<const uint4*>(r) + j);
          const uint32_t h[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const float2 x = A16::unpack2(h[t]);
            f[j * 8 + t * 2] += x.x;
            f[j * 8 + t * 2 + 1] += x.y;
          }
        }
      } else {
        for (int i = 0; i < 32; ++i)
          if (col0 + i < ws.N) f[i] += A16::to_float(r[i]);
      }
    }
  }

  __device__ __forceinline__ void tile(const Params& p, const WorkShape& ws, const EpiCtx& cx, uint32_t tacc,
                                       int nb) {
    const int cgi = cx.epi_warp >> 2;
    const int r_in_tile = cx.quad * 32 + cx.lane;
    const int row = cx.row0 + r_in_tile;
    const bool row_ok = row < ws.M;
    uint8_t* slab = cx.ep_smem + cgi * kSlabBytes;
    const bool issuer = (cx.epi_warp & 3) == 0 && cx.lane == 0;
#pragma unroll 1
    for (int s = 0; s < kColsPerGroup; s += 64) {
      const int col_in_tile = cgi * kColsPerGroup + s;
      const int col0 = nb * BN + col_in_tile;
      if (col0 >= ws.N) break;  // uniform across the column group
      const bool tma_res = (p.R != nullptr) && (p.C != nullptr) && (p.C32 == nullptr);
      uint64_t* rbar = reinterpret_cast<uint64_t*>(cx.ep_smem + kColGroups * kSlabBytes) + cgi;
      uint8_t* rowp = slab + r_in_tile * 128;
      if (tma_res && issuer) {
        // residual slab -> the staging buffer (free once the previous TMA store has read it)
        tc05::bulk_wait_read_all();
        tc05::mbar_arrive_expect_tx(rbar, kSlabBytes);
        tc05::tma_load_2d(slab, &p.tmR, rbar, col0, cx.row0, tc05::kEvictFirst);
      }
      // 8 epilogue warps (168 registers each): both 32-column halves of the slab are requested from TMEM before the
      // first is consumed.  16 warps (102 registers each): one half at a time.
      constexpr bool kLean = EPI_WARPS > 8;
      uint32_t va[32], vb[kLean ? 1 : 32];
      tc05::tmem_ld_32x32b_x32(tacc + col_in_tile, va);
      if constexpr (!kLean) tc05::tmem_ld_32x32b_x32(tacc + col_in_tile + 32, vb);
      tc05::tmem_ld_wait();
      if (!tma_res && p.C) {
        // the slab is free once the previous TMA store has finished reading it
        if (issuer) tc05::bulk_wait_read_all();
        tc05::named_bar_sync(2 + cgi, 128);
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float f[32];
        if constexpr (kLean) {
          if (h == 1) {
            tc05::tmem_ld_32x32b_x32(tacc + col_in_tile + 32, va);
            tc05::tmem_ld_wait();
          }
#pragma unroll
          for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(va[i]);
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(h == 0 ? va[i] : vb[i]);
        }
        finish_chunk(p, ws, f, row, row_ok, col0 + h * 32, !tma_res);
        if (p.C32 && row_ok) {
          float* o = p.C32 + (size_t)row * p.ldc32 + col0 + h * 32;
          if (col0 + h * 32 + 32 <= ws.N) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              reinterpret_cast<float4*>(o)[j] = make_float4(f[j * 4], f[j * 4 + 1], f[j * 4 + 2], f[j * 4 + 3]);
          } else {
            for (int i = 0; i < 32; ++i)
              if (col0 + h * 32 + i < ws.N) o[i] = f[i];
          }
        }
        if (p.C) {
          if (tma_res && h == 0) tc05::mbar_wait(rbar, rphase, 20);
#pragma unroll
          for (int q = 0; q < 4; ++q) {  // 16-byte chunk (h*4 + q) of this row, 128-byte swizzle
            uint4* cp = reinterpret_cast<uint4*>(rowp + (((h * 4 + q) ^ (r_in_tile & 7)) * 16));
            if (tma_res) {
              const uint4 rv = *cp;
              const uint32_t rh[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                const float2 x = A16::unpack2(rh[t]);
                f[q * 8 + t * 2] += x.x;
                f[q * 8 + t * 2 + 1] += x.y;
              }
            }
            uint4 o;
            o.x = A16::pack2(f[q * 8 + 0], f[q * 8 + 1]);
            o.y = A16::pack2(f[q * 8 + 2], f[q * 8 + 3]);
            o.z = A16::pack2(f[q * 8 + 4], f[q * 8 + 5]);
            o.w = A16::pack2(f[q * 8 + 6], f[q * 8 + 7]);
            *cp = o;
          }
        }
      }
      if (p.C) {
        if (tma_res) rphase ^= 1;
        tc05::fence_proxy_async_smem();
        tc05::named_bar_sync(2 + cgi, 128);
        if (issuer) {
          tc05::tma_store_2d(&p.tmC, slab, col0, cx.row0);
          tc05::bulk_commit_group();
        }
      }
    }
  }
};

// host helper: output tensor map for EpStore (bf16 [M, N] with row pitch ldc, box 64 x 128, SW128)
inline bool make_store_tmap(CUtensorMap* tm, void* C, int M, int N, int ldc) {
  return tc05_host::make_tmap_2d_16b(tm, C, M, N, ldc, BM, 64);
}

}  // namespace gemm
