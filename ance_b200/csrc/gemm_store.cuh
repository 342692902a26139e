// gemm_store.cuh — "store" epilogue of the tcgen05 GEMM: bias, exact-erf GELU, residual add, bf16
// and/or fp32 output.  Covers the encoder's linear layers K2/K4/K5/K6 of SURVEY.md §2.3 (HF
// RobertaSelfAttention / RobertaSelfOutput / RobertaIntermediate / RobertaOutput, called from
// model/models.py:150-151 in the reference) and the debug GEMM used by the bring-up tests.
#pragma once
#include "gemm_core.cuh"

namespace gemm {

__device__ __forceinline__ float gelu_erf(float x) {
  // HF "gelu": x * 0.5 * (1 + erf(x / sqrt(2)))
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}

template <int BN, int EPI_WARPS>
struct EpStore {
  static constexpr uint64_t kHintA = tc05::kEvictNormal;
  static constexpr uint64_t kHintB = tc05::kEvictLast;  // weights: keep in L2
  struct Params {
    __nv_bfloat16* C;        // [M, ldc] bf16 or null
    float* C32;              // [M, ldc32] fp32 or null
    const float* bias;       // [N] or null
    const __nv_bfloat16* R;  // residual [M, ldr] or null
    int ldc, ldc32, ldr;
    int act;                 // 0 none, 1 gelu(erf)
  };

  __device__ __forceinline__ void begin_work(const Params&, const WorkShape&, const EpiCtx&) {}
  __device__ __forceinline__ void end_work(const Params&, const WorkShape&, const EpiCtx&) {}

  __device__ __forceinline__ void tile(const Params& p, const WorkShape& ws, const EpiCtx& cx, uint32_t tacc,
                                       int nb) {
    constexpr int kColGroups = EPI_WARPS / 4;
    constexpr int kColsPerGroup = BN / kColGroups;
    static_assert(EPI_WARPS % 4 == 0 && kColsPerGroup % 32 == 0, "epilogue warp layout");
    const int cgi = cx.epi_warp >> 2;
    const int row = cx.row0 + cx.quad * 32 + cx.lane;
    const bool row_ok = row < ws.M;
#pragma unroll 1
    for (int c = 0; c < kColsPerGroup; c += 32) {
      const int col_in_tile = cgi * kColsPerGroup + c;
      const int col0 = nb * BN + col_in_tile;
      if (col0 >= ws.N) break;  // warp-uniform
      uint32_t v[32];
      tc05::tmem_ld_32x32b_x32(tacc + col_in_tile, v);
      tc05::tmem_ld_wait();
      float f[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
      const bool full = (col0 + 32 <= ws.N);
      if (p.bias) {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (full || col0 + i < ws.N) f[i] += __ldg(p.bias + col0 + i);
      }
      if (p.act == 1) {
#pragma unroll
        for (int i = 0; i < 32; ++i) f[i] = gelu_erf(f[i]);
      }
      if (row_ok) {
      if (p.R) {
        const __nv_bfloat16* r = p.R + (size_t)row * p.ldr + col0;
        if (full) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 q = __ldg(reinterpret_cast<const uint4*>(r) + j);
            const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              float2 x = __bfloat1622float2(h[t]);
              f[j * 8 + t * 2] += x.x;
              f[j * 8 + t * 2 + 1] += x.y;
            }
          }
        } else {
          for (int i = 0; i < 32; ++i)
            if (col0 + i < ws.N) f[i] += __bfloat162float(r[i]);
        }
      }
      if (p.C) {
        __nv_bfloat16* o = p.C + (size_t)row * p.ldc + col0;
        if (full) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 q;
            __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&q);
#pragma unroll
            for (int t = 0; t < 4; ++t) h[t] = __floats2bfloat162_rn(f[j * 8 + t * 2], f[j * 8 + t * 2 + 1]);
            reinterpret_cast<uint4*>(o)[j] = q;
          }
        } else {
          for (int i = 0; i < 32; ++i)
            if (col0 + i < ws.N) o[i] = __float2bfloat16_rn(f[i]);
        }
      }
      if (p.C32) {
        float* o = p.C32 + (size_t)row * p.ldc32 + col0;
        if (full) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            reinterpret_cast<float4*>(o)[j] = make_float4(f[j * 4], f[j * 4 + 1], f[j * 4 + 2], f[j * 4 + 3]);
        } else {
          for (int i = 0; i < 32; ++i)
            if (col0 + i < ws.N) o[i] = f[i];
        }
      }
      }  // row_ok
      __syncwarp();
    }
  }
};

}  // namespace gemm
