// gemm_store.cuh — "store" epilogue of the tcgen05 GEMM: bias, exact-erf GELU, residual add, bf16
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
#include "gemm_core.cuh"

namespace gemm {

// erf with |error| <= 1.5e-7 (Abramowitz & Stegun 7.1.26): one MUFU.RCP + one MUFU.EX2 + 7 FMA —
// three orders of magnitude below the bf16 rounding applied to the result.
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

__device__ __forceinline__ float erf_fast(float x) {
  const float ax = fabsf(x);
  const float t = rcp_approx(fmaf(0.3275911f, ax, 1.0f));   // MUFU.RCP, rel. error 2^-23
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  p *= t;
  const float e = ex2_approx(-1.4426950408889634f * ax * ax);  // MUFU.EX2, rel. error 2^-22
  return copysignf(fmaf(-p, e, 1.0f), x);
}

__device__ __forceinline__ float gelu_erf(float x) {
  // HF "gelu": x * 0.5 * (1 + erf(x / sqrt(2)))
  return 0.5f * x * (1.0f + erf_fast(x * 0.70710678118654752440f));
}

template <int BN, int EPI_WARPS>
struct EpStore {
  static constexpr uint64_t kHintA = tc05::kEvictNormal;
  static constexpr uint64_t kHintB = tc05::kEvictLast;  // weights: keep in L2
  static constexpr int kColGroups = EPI_WARPS / 4;
  static constexpr int kColsPerGroup = BN / kColGroups;
  static constexpr int kSlabBytes = BM * 128;            // 128 rows x 64 bf16
  static constexpr int kSmemBytes = kColGroups * kSlabBytes;
  static_assert(EPI_WARPS % 4 == 0 && kColsPerGroup % 64 == 0, "epilogue warp layout");

  struct alignas(64) Params {
    CUtensorMap tmC;         // bf16 output [M, N], box {64, 128}, SWIZZLE_128B (valid when C != null)
    __nv_bfloat16* C;        // [M, ldc] bf16 or null
    float* C32;              // [M, ldc32] fp32 or null (direct stores)
    const float* bias;       // [N] or null
    const __nv_bfloat16* R;  // residual [M, ldr] or null
    int ldc, ldc32, ldr;
    int act;                 // 0 none, 1 gelu(erf)
  };

  __device__ __forceinline__ void begin_work(const Params&, const WorkShape&, const EpiCtx&) {}
  __device__ __forceinline__ void end_work(const Params&, const WorkShape&, const EpiCtx&) {}
  __device__ __forceinline__ void end_kernel(const Params& p, const EpiCtx& cx) {
    if (p.C && (cx.epi_warp & 3) == 0 && cx.lane == 0) tc05::bulk_wait_all();
  }

  // bias + activation + residual on one 32-column chunk held as fp32
  __device__ __forceinline__ void finish_chunk(const Params& p, const WorkShape& ws, float (&f)[32], int row,
                                               bool row_ok, int col0) {
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
      for (int i = 0; i < 32; ++i) f[i] = gelu_erf(f[i]);
    }
    if (p.R && row_ok) {
      const __nv_bfloat16* r = p.R + (size_t)row * p.ldr + col0;
      if (full) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint4 q = __ldg(reinterpret_cast<const uint4*>(r) + j);
          const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const float2 x = __bfloat1622float2(h[t]);
            f[j * 8 + t * 2] += x.x;
            f[j * 8 + t * 2 + 1] += x.y;
          }
        }
      } else {
        for (int i = 0; i < 32; ++i)
          if (col0 + i < ws.N) f[i] += __bfloat162float(r[i]);
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
      uint32_t pk[32];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t v[32];
        tc05::tmem_ld_32x32b_x32(tacc + col_in_tile + h * 32, v);
        tc05::tmem_ld_wait();
        float f[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
        finish_chunk(p, ws, f, row, row_ok, col0 + h * 32);
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
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const __nv_bfloat162 h2 = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
          pk[h * 16 + i] = *reinterpret_cast<const uint32_t*>(&h2);
        }
      }
      if (p.C) {
        // the slab is free once the previous TMA store has finished reading it
        if (issuer) tc05::bulk_wait_read_all();
        tc05::named_bar_sync(2 + cgi, 128);
        uint8_t* rowp = slab + r_in_tile * 128;
#pragma unroll
        for (int q = 0; q < 8; ++q)  // 16-byte chunk q of this row, 128-byte swizzle
          *reinterpret_cast<uint4*>(rowp + ((q ^ (r_in_tile & 7)) * 16)) =
              make_uint4(pk[q * 4], pk[q * 4 + 1], pk[q * 4 + 2], pk[q * 4 + 3]);
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
