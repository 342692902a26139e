// encoder.cu — BERT/RoBERTa-base dual-encoder forward on sm_100a.
//
// Replaces the library calls behind the reference's
//   model/models.py:149-157  RobertaDot_NLL_LN.query_emb/body_emb  (HF RobertaModel -> CLS -> embeddingHead -> norm)
//   model/models.py:165-199  MultiChunk body_emb (caller reshapes [B,2048] -> [4B,512]; token 0 of each chunk)
//   model/models.py:223-259  BiEncoder / HFBertEncoder (CLS of the last layer)
// Per layer:
//   QKV  = X Wqkv^T + b                       tcgen05 GEMM (gemm_core.cuh), bias epilogue
//   CTX  = softmax(QK^T/8 + mask) V           attention.cuh
//   T    = CTX Wo^T + b + X ; X1 = LN(T)      GEMM with bias+residual epilogue, then ln_rows_kernel
//   F    = gelu_erf(X1 W1^T + b1)             GEMM with bias+GELU epilogue
//   T    = F W2^T + b2 + X1 ; X = LN(T)       GEMM with bias+residual epilogue, then ln_rows_kernel
// Activations are bf16 in HBM; all accumulation, LayerNorm statistics and softmax are fp32.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <vector>

#include "attention.cuh"
#include "common.h"
#include "gemm_store.cuh"

namespace {

constexpr float kLog2e = 1.4426950408889634f;

// ------------------------------------------------------------------------------------------------
// K1: embeddings gather + LayerNorm, position ids, key-bias
// ------------------------------------------------------------------------------------------------
struct EmbedParams {
  const int32_t* ids;    // [B, L]
  const int32_t* lens;   // [B] or null
  const uint8_t* mask;   // [B, L] or null
  int B, L, H;
  int roberta;           // 1: pos = cumsum(ids != pad) * (ids != pad) + pad ; 0: pos = 0..L-1
  int pad_id, vocab, max_pos;
  const __nv_bfloat16* word;  // [vocab, H]
  const __nv_bfloat16* pos;   // [max_pos, H]
  const __nv_bfloat16* type;  // [type_vocab, H] (row 0)
  const float* gamma;
  const float* beta;
  float eps;
  __nv_bfloat16* X;      // [B*L, H]
  float* kbias;          // [B*L]  (1 - mask) * -10000 * log2e
  int* err_flag;
};

template <int NV>  // H = NV * 256
__global__ void __launch_bounds__(256) embed_ln_kernel(const EmbedParams p) {
  __shared__ int s_pos[512];
  __shared__ int s_warp_cnt[8];
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int32_t* ids = p.ids + static_cast<size_t>(b) * p.L;
  // position ids (L <= 512): inclusive scan of (id != pad)
  for (int base = 0, carry = 0; base < p.L; base += 256) {
    const int t = base + threadIdx.x;
    const int flag = (t < p.L && ids[t] != p.pad_id) ? 1 : 0;
    const unsigned bal = __ballot_sync(0xffffffffu, flag);
    if (lane == 0) s_warp_cnt[warp] = __popc(bal);
    __syncthreads();
    int pre = carry;
    for (int w2 = 0; w2 < warp; ++w2) pre += s_warp_cnt[w2];
    const int incl = pre + __popc(bal & ((2u << lane) - 1u));
    if (t < p.L) s_pos[t] = p.roberta ? (flag ? incl + p.pad_id : p.pad_id) : t;
    int tot = 0;
    for (int w2 = 0; w2 < 8; ++w2) tot += s_warp_cnt[w2];
    carry += tot;
    __syncthreads();
  }
  const int len = p.lens ? p.lens[b] : 0;
  for (int t = warp; t < p.L; t += 8) {
    const size_t tok = static_cast<size_t>(b) * p.L + t;
    int id = ids[t];
    int ps = s_pos[t];
    if (id < 0 || id >= p.vocab || ps >= p.max_pos) {
      if (lane == 0) atomicExch(p.err_flag, 1);
      id = min(max(id, 0), p.vocab - 1);
      ps = min(ps, p.max_pos - 1);
    }
    const uint4* wr = reinterpret_cast<const uint4*>(p.word + static_cast<size_t>(id) * p.H);
    const uint4* pr = reinterpret_cast<const uint4*>(p.pos + static_cast<size_t>(ps) * p.H);
    const uint4* tr = reinterpret_cast<const uint4*>(p.type);
    float x[NV * 8];
    float sum = 0.f;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const uint4 a = __ldg(wr + v * 32 + lane), c = __ldg(pr + v * 32 + lane), d = __ldg(tr + v * 32 + lane);
      const __nv_bfloat162* ah = reinterpret_cast<const __nv_bfloat162*>(&a);
      const __nv_bfloat162* ch = reinterpret_cast<const __nv_bfloat162*>(&c);
      const __nv_bfloat162* dh = reinterpret_cast<const __nv_bfloat162*>(&d);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float2 fa = __bfloat1622float2(ah[q]), fc = __bfloat1622float2(ch[q]), fd = __bfloat1622float2(dh[q]);
        x[v * 8 + q * 2] = fa.x + fc.x + fd.x;
        x[v * 8 + q * 2 + 1] = fa.y + fc.y + fd.y;
        sum += x[v * 8 + q * 2] + x[v * 8 + q * 2 + 1];
      }
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, s);
    const float mean = sum / p.H;
    float var = 0.f;
#pragma unroll
    for (int i = 0; i < NV * 8; ++i) {
      const float dlt = x[i] - mean;
      var = fmaf(dlt, dlt, var);
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) var += __shfl_xor_sync(0xffffffffu, var, s);
    const float rstd = rsqrtf(var / p.H + p.eps);
    uint4* out = reinterpret_cast<uint4*>(p.X + tok * p.H);
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int col = (v * 32 + lane) * 8;
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.gamma + col)), g1 = __ldg(reinterpret_cast<const float4*>(p.gamma + col + 4));
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.beta + col)), b1 = __ldg(reinterpret_cast<const float4*>(p.beta + col + 4));
      const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      uint4 u;
      __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
      for (int q = 0; q < 4; ++q)
        h2[q] = __floats2bfloat162_rn((x[v * 8 + q * 2] - mean) * rstd * g[q * 2] + bb[q * 2],
                                      (x[v * 8 + q * 2 + 1] - mean) * rstd * g[q * 2 + 1] + bb[q * 2 + 1]);
      out[v * 32 + lane] = u;
    }
    if (lane == 0) {
      const bool keep = p.mask ? (p.mask[tok] != 0) : (t < len);
      p.kbias[tok] = keep ? 0.f : -10000.0f * kLog2e;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm over rows: bf16 or fp32 in, bf16 and/or fp32 out; row r read at in + r * in_ld
// ------------------------------------------------------------------------------------------------
template <int NV, bool kInF32>
__global__ void __launch_bounds__(256) ln_rows_kernel(const void* __restrict__ in, size_t in_ld, int n_rows, int H,
                                                      const float* __restrict__ gamma, const float* __restrict__ beta,
                                                      float eps, __nv_bfloat16* __restrict__ out16,
                                                      float* __restrict__ out32) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= n_rows) return;
  float x[NV * 8];
  float sum = 0.f;
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int col = (v * 32 + lane) * 8;
    if (kInF32) {
      const float* r = reinterpret_cast<const float*>(in) + static_cast<size_t>(row) * in_ld + col;
      const float4 a = __ldg(reinterpret_cast<const float4*>(r)), b = __ldg(reinterpret_cast<const float4*>(r + 4));
      x[v * 8 + 0] = a.x; x[v * 8 + 1] = a.y; x[v * 8 + 2] = a.z; x[v * 8 + 3] = a.w;
      x[v * 8 + 4] = b.x; x[v * 8 + 5] = b.y; x[v * 8 + 6] = b.z; x[v * 8 + 7] = b.w;
    } else {
      const __nv_bfloat16* r = reinterpret_cast<const __nv_bfloat16*>(in) + static_cast<size_t>(row) * in_ld + col;
      const uint4 a = __ldg(reinterpret_cast<const uint4*>(r));
      const __nv_bfloat162* ah = reinterpret_cast<const __nv_bfloat162*>(&a);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float2 f = __bfloat1622float2(ah[q]);
        x[v * 8 + q * 2] = f.x;
        x[v * 8 + q * 2 + 1] = f.y;
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) sum += x[v * 8 + i];
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, s);
  const float mean = sum / H;
  float var = 0.f;
#pragma unroll
  for (int i = 0; i < NV * 8; ++i) {
    const float d = x[i] - mean;
    var = fmaf(d, d, var);
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) var += __shfl_xor_sync(0xffffffffu, var, s);
  const float rstd = rsqrtf(var / H + eps);
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int col = (v * 32 + lane) * 8;
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + col)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + col + 4));
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + col)), b1 = __ldg(reinterpret_cast<const float4*>(beta + col + 4));
    const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    float y[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) y[i] = (x[v * 8 + i] - mean) * rstd * g[i] + bb[i];
    if (out16) {
      uint4 u;
      __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
      for (int q = 0; q < 4; ++q) h2[q] = __floats2bfloat162_rn(y[q * 2], y[q * 2 + 1]);
      *reinterpret_cast<uint4*>(out16 + static_cast<size_t>(row) * H + col) = u;
    }
    if (out32) {
      float* o = out32 + static_cast<size_t>(row) * H + col;
      *reinterpret_cast<float4*>(o) = make_float4(y[0], y[1], y[2], y[3]);
      *reinterpret_cast<float4*>(o + 4) = make_float4(y[4], y[5], y[6], y[7]);
    }
  }
}

// bf16 -> bf16 LayerNorm, kB rows per warp written as independent instruction streams (their shuffle / FMA chains
// overlap and gamma / beta are fetched once); the arithmetic of a row is that of ln_rows_kernel, bit for bit.
template <int NV, int kB>
__global__ void __launch_bounds__(256) ln_rows_multi_kernel(const __nv_bfloat16* __restrict__ in, size_t in_ld, int n_rows, int H,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            float eps, __nv_bfloat16* __restrict__ out16) {
  const int lane = threadIdx.x & 31;
  const int row0 = (blockIdx.x * 8 + (threadIdx.x >> 5)) * kB;
  if (row0 >= n_rows) return;
  float x[kB][NV * 8], sum[kB], var[kB], mean[kB], rstd[kB];
#pragma unroll
  for (int b = 0; b < kB; ++b) {
    const int row = min(row0 + b, n_rows - 1);
    const uint4* src = reinterpret_cast<const uint4*>(in + static_cast<size_t>(row) * in_ld);
    uint4 raw[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) raw[v] = __ldg(src + v * 32 + lane);
    sum[b] = 0.f;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const __nv_bfloat162* ah = reinterpret_cast<const __nv_bfloat162*>(&raw[v]);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float2 f = __bfloat1622float2(ah[q]);
        x[b][v * 8 + q * 2] = f.x;
        x[b][v * 8 + q * 2 + 1] = f.y;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) sum[b] += x[b][v * 8 + i];
    }
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
#pragma unroll
    for (int b = 0; b < kB; ++b) sum[b] += __shfl_xor_sync(0xffffffffu, sum[b], s);
  }
#pragma unroll
  for (int b = 0; b < kB; ++b) {
    mean[b] = sum[b] / H;
    var[b] = 0.f;
#pragma unroll
    for (int i = 0; i < NV * 8; ++i) {
      const float d = x[b][i] - mean[b];
      var[b] = fmaf(d, d, var[b]);
    }
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
#pragma unroll
    for (int b = 0; b < kB; ++b) var[b] += __shfl_xor_sync(0xffffffffu, var[b], s);
  }
#pragma unroll
  for (int b = 0; b < kB; ++b) rstd[b] = rsqrtf(var[b] / H + eps);
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int col = (v * 32 + lane) * 8;
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + col)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + col + 4));
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + col)), b1 = __ldg(reinterpret_cast<const float4*>(beta + col + 4));
    const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int b = 0; b < kB; ++b) {
      float y[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) y[i] = (x[b][v * 8 + i] - mean[b]) * rstd[b] * g[i] + bb[i];
      uint4 u;
      __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
      for (int q = 0; q < 4; ++q) h2[q] = __floats2bfloat162_rn(y[q * 2], y[q * 2 + 1]);
      if (row0 + b < n_rows) *reinterpret_cast<uint4*>(out16 + static_cast<size_t>(row0 + b) * H + col) = u;
    }
  }
}

// rows r*stride of a bf16 matrix -> fp32 [n, H]   (DPR: CLS of the last layer, models.py:239)
__global__ void gather_rows_f32_kernel(const __nv_bfloat16* __restrict__ X, size_t row_stride, int n, int H,
                                       float* __restrict__ out) {
  const int r = blockIdx.x;
  for (int c = threadIdx.x; c < H; c += blockDim.x)
    out[static_cast<size_t>(r) * H + c] = __bfloat162float(X[static_cast<size_t>(r) * row_stride + c]);
}

__global__ void bf16_to_f32_kernel(const __nv_bfloat16* __restrict__ in, float* __restrict__ out, size_t n) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __bfloat162float(in[i]);
}

}  // namespace

// ================================================================================================
// handle
// ================================================================================================
struct LayerDev {
  __nv_bfloat16 *wqkv, *wo, *w1, *w2;    // [3H,H] [H,H] [F,H] [H,F]
  float *bqkv, *bo, *b1, *b2, *ln1g, *ln1b, *ln2g, *ln2b;
};

struct ance_encoder {
  ance_encoder_config cfg{};
  int max_tokens = 0;
  __nv_bfloat16 *word = nullptr, *pos = nullptr, *type = nullptr;
  float *eg = nullptr, *eb = nullptr;
  std::vector<LayerDev> layers;
  __nv_bfloat16* head_w = nullptr;
  float *head_b = nullptr, *head_g = nullptr, *head_bt = nullptr;
  // activations
  __nv_bfloat16 *X = nullptr, *QKV = nullptr, *CTX = nullptr, *T = nullptr, *X1 = nullptr, *FF = nullptr;
  float* kbias = nullptr;
  float* head_tmp = nullptr;  // [max_seqs, H] fp32
  int* err_flag = nullptr;
  __nv_bfloat16* dbg = nullptr;  // [(n_layer+1), max_tokens, H] when debugging
  int dbg_tokens = 0;
  int prune_last_layer = 1;  // last layer: only the CLS rows go through out-proj / FFN (identical result)
  std::vector<void*> allocs;
};

namespace {

template <class T>
T* dev_alloc(ance_encoder* e, size_t n) {
  void* p = nullptr;
  if (cudaMalloc(&p, n * sizeof(T)) != cudaSuccess) return nullptr;
  e->allocs.push_back(p);
  return reinterpret_cast<T*>(p);
}

float* upload_f32(ance_encoder* e, const float* h, size_t n) {
  float* d = dev_alloc<float>(e, n);
  if (d) cudaMemcpy(d, h, n * 4, cudaMemcpyHostToDevice);
  return d;
}

__nv_bfloat16* upload_bf16(ance_encoder* e, const float* h, size_t n) {
  std::vector<__nv_bfloat16> tmp(n);
  for (size_t i = 0; i < n; ++i) tmp[i] = __float2bfloat16_rn(h[i]);
  __nv_bfloat16* d = dev_alloc<__nv_bfloat16>(e, n);
  if (d) cudaMemcpy(d, tmp.data(), n * 2, cudaMemcpyHostToDevice);
  return d;
}

// one GEMM of the forward: C[M,N] = act(A[M,K] W[N,K]^T + bias) (+ R)
// cta_group::2: 256 x 256 tile per CTA pair.  EW = 8 epilogue warps (two 128-column groups, 6 smem stages) or EW = 16
// (four 64-column groups, each with its own staging slab, 5 stages): with 16 a tile's epilogue is ONE
// LDTM -> math -> TMA-store round per warp instead of two in sequence.  Measured at 592 x 128 (ms per forward, 8 -> 16):
// FFN-up (GELU) 4.15 -> 3.62, out-proj 1.15 -> 1.06, but QKV 2.55 -> 2.90 and FFN-down 2.81 -> 2.90: the heavy / short-K
// epilogues want the shorter chain, the light ones the deeper operand pipeline.
template <int EW, int STAGES>
int linear_cfg(const __nv_bfloat16* A, size_t lda, int M, const __nv_bfloat16* W, int N, int K, const float* bias,
               const __nv_bfloat16* R, int act, __nv_bfloat16* C, float* C32, cudaStream_t st, int cls, size_t ldr) {
  constexpr int BN = 256, CG = 2;
  using Ep = gemm::EpStore<BN, EW>;
  CUtensorMap tmA, tmB;
  if (!tc05_host::make_tmap_2d_16b(&tmA, A, M, K, lda, gemm::BM) || !tc05_host::make_tmap_2d_16b(&tmB, W, N, K, K, BN / CG)) {
    ance::set_error("encoder: cuTensorMapEncodeTiled failed (M=%d N=%d K=%d)", M, N, K);
    return ANCE_ERR_CUDA;
  }
  gemm::WorkShape ws = gemm::make_shape(M, N, K, BN, CG, 0);
  typename Ep::Params p;
  memset(&p, 0, sizeof(p));
  if (C && !gemm::make_store_tmap(&p.tmC, C, M, N, N)) {
    ance::set_error("encoder: cuTensorMapEncodeTiled failed for the output (M=%d N=%d)", M, N);
    return ANCE_ERR_CUDA;
  }
  if (ldr == 0) ldr = N;
  if (C && R && !gemm::make_store_tmap(&p.tmR, const_cast<__nv_bfloat16*>(R), M, N, static_cast<int>(ldr))) {
    ance::set_error("encoder: cuTensorMapEncodeTiled failed for the residual (M=%d N=%d)", M, N);
    return ANCE_ERR_CUDA;
  }
  p.C = C;
  p.C32 = C32;
  p.bias = bias;
  p.R = R;
  p.ldc = N;
  p.ldc32 = N;
  p.ldr = static_cast<int>(ldr);
  // 2 = logistic form (|err| <= 3.7e-6, FFN-up 3.66 -> 3.49 ms per forward, A/B on one box), 1 = erfc form (|err| <= 7e-7)
  static const int gelu_form = getenv("ANCE_B200_GELU") ? atoi(getenv("ANCE_B200_GELU")) : 2;
  p.act = act ? gelu_form : 0;
  {
    ance::ProfScope ps(cls, st);
    ANCE_CUDA((gemm::launch<Ep, BN, STAGES, CG, EW, tc05::kFmtBF16>(tmA, tmB, ws, p, 0, st)));
  }
  ance::count_launch(1);
  return ANCE_OK;
}

int linear(const __nv_bfloat16* A, size_t lda, int M, const __nv_bfloat16* W, int N, int K, const float* bias,
           const __nv_bfloat16* R, int act, __nv_bfloat16* C, float* C32, cudaStream_t st, int cls = ance::kClsGemm,
           size_t ldr = 0) {
  bool short_chain = (act != 0) || (R != nullptr && K <= 1024);   // FFN-up, out-proj
  static const char* force = getenv("ANCE_B200_EPI_MASK");   // tuning aid: bit per GEMM class (qkv, out, ffn1, ffn2)
  if (force && cls >= ance::kClsGemmQkv && cls <= ance::kClsGemmFfn2) short_chain = (atoi(force) >> (cls - ance::kClsGemmQkv)) & 1;
  if (short_chain) return linear_cfg<16, 5>(A, lda, M, W, N, K, bias, R, act, C, C32, st, cls, ldr);
  return linear_cfg<8, 6>(A, lda, M, W, N, K, bias, R, act, C, C32, st, cls, ldr);
}

int g_ln_rows_per_warp = 2;   // ance_encoder_set_param("ln_rows_per_warp"): 1.58 -> 1.25 ms per forward at 592 x 128 (4: 1.65)

int layer_norm(const void* in, bool in_f32, size_t in_ld, int rows, int H, const float* g, const float* b, float eps,
               __nv_bfloat16* out16, float* out32, cudaStream_t st) {
  const int blocks = (rows + 7) / 8;
  const int nv = H / 256;
  ance::ProfScope ps(ance::kClsNorm, st);
  if (!in_f32 && out16 && !out32 && nv == 3 && g_ln_rows_per_warp > 1 && rows >= 4096) {
    const __nv_bfloat16* src = reinterpret_cast<const __nv_bfloat16*>(in);
    if (g_ln_rows_per_warp == 2) ln_rows_multi_kernel<3, 2><<<(rows + 15) / 16, 256, 0, st>>>(src, in_ld, rows, H, g, b, eps, out16);
    else ln_rows_multi_kernel<3, 4><<<(rows + 31) / 32, 256, 0, st>>>(src, in_ld, rows, H, g, b, eps, out16);
    ANCE_CUDA(cudaGetLastError());
    ance::count_launch(1);
    return ANCE_OK;
  }
#define LN_CASE(NV_)                                                                                       \
  if (in_f32) ln_rows_kernel<NV_, true><<<blocks, 256, 0, st>>>(in, in_ld, rows, H, g, b, eps, out16, out32); \
  else ln_rows_kernel<NV_, false><<<blocks, 256, 0, st>>>(in, in_ld, rows, H, g, b, eps, out16, out32)
  if (nv == 3) { LN_CASE(3); }
  else if (nv == 4) { LN_CASE(4); }
  else if (nv == 1) { LN_CASE(1); }
  else if (nv == 2) { LN_CASE(2); }
  else { ance::set_error("encoder: hidden size %d unsupported", H); return ANCE_ERR_UNSUPPORTED; }
#undef LN_CASE
  ANCE_CUDA(cudaGetLastError());
  ance::count_launch(1);
  return ANCE_OK;
}

}  // namespace

extern "C" int ance_encoder_create(const ance_encoder_config* cfg, const ance_encoder_weights* w, int max_tokens,
                                   ance_encoder_t* out) {
  ANCE_REQUIRE(cfg && w && out, "ance_encoder_create: null argument");
  ANCE_REQUIRE(cfg->hidden % 256 == 0 && cfg->hidden <= 1024, "ance_encoder_create: hidden must be a multiple of 256 (<= 1024), got %d", cfg->hidden);
  ANCE_REQUIRE(cfg->heads * 64 == cfg->hidden, "ance_encoder_create: head_dim must be 64 (hidden %d, heads %d)", cfg->hidden, cfg->heads);
  ANCE_REQUIRE(cfg->ffn % 8 == 0 && cfg->n_layer > 0 && cfg->vocab > 0 && cfg->max_pos > 0, "ance_encoder_create: bad config");
  ANCE_REQUIRE(max_tokens >= 128, "ance_encoder_create: max_tokens must be >= 128");
  ANCE_REQUIRE(!cfg->has_head || (w->head_w && w->head_b && w->head_ln_g && w->head_ln_b), "ance_encoder_create: has_head without head weights");
  {
    int dev = 0, major = 0;
    ANCE_CUDA(cudaGetDevice(&dev));
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (major != 10) {
      ance::set_error("device %d has compute capability %d.x; libance_b200 is built for sm_100a only (no CPU fallback)", dev, major);
      return ANCE_ERR_CUDA;
    }
  }
  ance_encoder* e = new ance_encoder();
  e->cfg = *cfg;
  e->max_tokens = (max_tokens + 127) / 128 * 128;
  const size_t H = cfg->hidden, F = cfg->ffn, T = e->max_tokens;
  bool ok = true;
  auto chk = [&](const void* p) { ok = ok && (p != nullptr); };
  chk(e->word = upload_bf16(e, w->word_emb, static_cast<size_t>(cfg->vocab) * H));
  chk(e->pos = upload_bf16(e, w->pos_emb, static_cast<size_t>(cfg->max_pos) * H));
  chk(e->type = upload_bf16(e, w->type_emb, static_cast<size_t>(cfg->type_vocab) * H));
  chk(e->eg = upload_f32(e, w->emb_ln_g, H));
  chk(e->eb = upload_f32(e, w->emb_ln_b, H));
  e->layers.resize(cfg->n_layer);
  for (int l = 0; l < cfg->n_layer && ok; ++l) {
    const ance_layer_weights& lw = w->layers[l];
    LayerDev& d = e->layers[l];
    std::vector<float> wqkv(3 * H * H), bqkv(3 * H);
    memcpy(wqkv.data(), lw.q_w, H * H * 4);
    memcpy(wqkv.data() + H * H, lw.k_w, H * H * 4);
    memcpy(wqkv.data() + 2 * H * H, lw.v_w, H * H * 4);
    memcpy(bqkv.data(), lw.q_b, H * 4);
    memcpy(bqkv.data() + H, lw.k_b, H * 4);
    memcpy(bqkv.data() + 2 * H, lw.v_b, H * 4);
    chk(d.wqkv = upload_bf16(e, wqkv.data(), wqkv.size()));
    chk(d.bqkv = upload_f32(e, bqkv.data(), bqkv.size()));
    chk(d.wo = upload_bf16(e, lw.ao_w, H * H));
    chk(d.bo = upload_f32(e, lw.ao_b, H));
    chk(d.ln1g = upload_f32(e, lw.ln1_g, H));
    chk(d.ln1b = upload_f32(e, lw.ln1_b, H));
    chk(d.w1 = upload_bf16(e, lw.ff1_w, F * H));
    chk(d.b1 = upload_f32(e, lw.ff1_b, F));
    chk(d.w2 = upload_bf16(e, lw.ff2_w, H * F));
    chk(d.b2 = upload_f32(e, lw.ff2_b, H));
    chk(d.ln2g = upload_f32(e, lw.ln2_g, H));
    chk(d.ln2b = upload_f32(e, lw.ln2_b, H));
  }
  if (cfg->has_head && ok) {
    chk(e->head_w = upload_bf16(e, w->head_w, H * H));
    chk(e->head_b = upload_f32(e, w->head_b, H));
    chk(e->head_g = upload_f32(e, w->head_ln_g, H));
    chk(e->head_bt = upload_f32(e, w->head_ln_b, H));
  }
  chk(e->X = dev_alloc<__nv_bfloat16>(e, T * H));
  chk(e->QKV = dev_alloc<__nv_bfloat16>(e, T * 3 * H));
  chk(e->CTX = dev_alloc<__nv_bfloat16>(e, T * H));
  chk(e->T = dev_alloc<__nv_bfloat16>(e, T * H));
  chk(e->X1 = dev_alloc<__nv_bfloat16>(e, T * H));
  chk(e->FF = dev_alloc<__nv_bfloat16>(e, T * F));
  chk(e->kbias = dev_alloc<float>(e, T));
  chk(e->head_tmp = dev_alloc<float>(e, T / 16 * H));
  chk(e->err_flag = dev_alloc<int>(e, 1));
  if (!ok || cudaGetLastError() != cudaSuccess) {
    ance::set_error("ance_encoder_create: device allocation / upload failed (max_tokens %d)", max_tokens);
    ance_encoder_destroy(e);
    return ANCE_ERR_NOMEM;
  }
  cudaMemset(e->err_flag, 0, sizeof(int));
  static bool attr = false;
  if (!attr) {
    ANCE_CUDA(cudaFuncSetAttribute(attn::attention_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, attn::Smem::kDynamic));
    ANCE_CUDA(cudaFuncSetAttribute(attn::attention_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, attn::Smem::kDynamic));
    ANCE_CUDA(cudaFuncSetAttribute(attn::attention_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, attn::Smem::kDynamic));
    attr = true;
  }
  *out = e;
  return ANCE_OK;
}

extern "C" int ance_encoder_destroy(ance_encoder_t e) {
  if (!e) return ANCE_OK;
  for (void* p : e->allocs) cudaFree(p);
  delete e;
  return ANCE_OK;
}

extern "C" int ance_encoder_forward(ance_encoder_t e, const int32_t* ids_dev, const int32_t* lens_dev,
                                    const uint8_t* mask_dev, int B, int L, float* out_dev, void* stream) {
  ANCE_REQUIRE(e != nullptr, "ance_encoder_forward: null handle");
  ANCE_REQUIRE(ids_dev && out_dev, "ance_encoder_forward: null buffer");
  ANCE_REQUIRE((lens_dev != nullptr) != (mask_dev != nullptr), "ance_encoder_forward: pass exactly one of lens_dev / mask_dev");
  ANCE_REQUIRE(B > 0 && L > 0, "ance_encoder_forward: empty batch");
  ANCE_REQUIRE(L <= 512 && ((L % 128 == 0) || (128 % L == 0 && L >= 8)), "ance_encoder_forward: L = %d unsupported (need a multiple of 128 up to 512, or a divisor of 128)", L);
  const ance_encoder_config& c = e->cfg;
  ANCE_REQUIRE(L + (c.arch == ANCE_ARCH_ROBERTA ? c.pad_id + 1 : 0) <= c.max_pos, "ance_encoder_forward: L = %d exceeds max_position_embeddings %d", L, c.max_pos);
  const long long tokens = static_cast<long long>(B) * L;
  ANCE_REQUIRE(tokens <= e->max_tokens, "ance_encoder_forward: %lld tokens exceed max_tokens %d", tokens, e->max_tokens);
  ANCE_REQUIRE(B <= e->max_tokens / 16, "ance_encoder_forward: batch %d too large for the head buffer", B);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int M = static_cast<int>(tokens), H = c.hidden, F = c.ffn;
  int rc;
  // K1
  EmbedParams ep;
  ep.ids = ids_dev; ep.lens = lens_dev; ep.mask = mask_dev;
  ep.B = B; ep.L = L; ep.H = H;
  ep.roberta = (c.arch == ANCE_ARCH_ROBERTA);
  ep.pad_id = c.pad_id; ep.vocab = c.vocab; ep.max_pos = c.max_pos;
  ep.word = e->word; ep.pos = e->pos; ep.type = e->type;
  ep.gamma = e->eg; ep.beta = e->eb; ep.eps = c.ln_eps;
  ep.X = e->X; ep.kbias = e->kbias; ep.err_flag = e->err_flag;
  ance::prof_begin(ance::kClsNorm, st);
  switch (H / 256) {
    case 1: embed_ln_kernel<1><<<B, 256, 0, st>>>(ep); break;
    case 2: embed_ln_kernel<2><<<B, 256, 0, st>>>(ep); break;
    case 3: embed_ln_kernel<3><<<B, 256, 0, st>>>(ep); break;
    default: embed_ln_kernel<4><<<B, 256, 0, st>>>(ep); break;
  }
  ance::prof_end(ance::kClsNorm, st);
  ANCE_CUDA(cudaGetLastError());
  ance::count_launch(1);
  if (e->dbg && M <= e->dbg_tokens) ANCE_CUDA(cudaMemcpyAsync(e->dbg, e->X, static_cast<size_t>(M) * H * 2, cudaMemcpyDeviceToDevice, st));
  // attention tensor map over the QKV buffer
  CUtensorMap tmQKV;
  if (!tc05_host::make_tmap_2d_16b(&tmQKV, e->QKV, M, 3 * H, 3 * H, attn::kTile)) {
    ance::set_error("encoder: cuTensorMapEncodeTiled failed for QKV");
    return ANCE_ERR_CUDA;
  }
  CUtensorMap tmCTX;
  if (!tc05_host::make_tmap_2d_16b(&tmCTX, e->CTX, M, H, H, attn::kTile)) {
    ance::set_error("encoder: cuTensorMapEncodeTiled failed for the attention output");
    return ANCE_ERR_CUDA;
  }
  attn::Params ap;
  ap.n_tokens = M; ap.L = L; ap.heads = c.heads; ap.hidden = H;
  ap.kbias = e->kbias;
  ap.scale_log2 = kLog2e / 8.0f;
  const int attn_work = ((M + 127) / 128) * c.heads;
  const int attn_grid = std::min(attn_work, gemm::sm_count());
  for (int l = 0; l < c.n_layer; ++l) {
    const LayerDev& d = e->layers[l];
    if ((rc = linear(e->X, H, M, d.wqkv, 3 * H, H, d.bqkv, nullptr, 0, e->QKV, nullptr, st, ance::kClsGemmQkv))) return rc;
    ance::prof_begin(ance::kClsAttn, st);
    if (L < attn::kTile) attn::attention_kernel<true, true><<<attn_grid, attn::kThreads, attn::Smem::kDynamic, st>>>(tmQKV, tmCTX, ap);
    else if (L == attn::kTile) attn::attention_kernel<false, true><<<attn_grid, attn::kThreads, attn::Smem::kDynamic, st>>>(tmQKV, tmCTX, ap);
    else attn::attention_kernel<false, false><<<attn_grid, attn::kThreads, attn::Smem::kDynamic, st>>>(tmQKV, tmCTX, ap);
    ance::prof_end(ance::kClsAttn, st);
    ANCE_CUDA(cudaGetLastError());
    ance::count_launch(1);
    // In the last layer only token 0 of every sequence is read downstream (models.py:49,193): run the
    // out-projection, FFN and both LayerNorms on those B rows only (strided TMA views, compact outputs).
    const bool cls_only = e->prune_last_layer && (l == c.n_layer - 1);
    const int Mr = cls_only ? B : M;                                    // rows processed from here on
    const size_t pitch = cls_only ? static_cast<size_t>(L) * H : H;     // row pitch of CTX / X views
    if ((rc = linear(e->CTX, pitch, Mr, d.wo, H, H, d.bo, e->X, 0, e->T, nullptr, st, ance::kClsGemmOut, pitch))) return rc;
    if ((rc = layer_norm(e->T, false, H, Mr, H, d.ln1g, d.ln1b, c.ln_eps, e->X1, nullptr, st))) return rc;
    if ((rc = linear(e->X1, H, Mr, d.w1, F, H, d.b1, nullptr, 1, e->FF, nullptr, st, ance::kClsGemmFfn1))) return rc;
    if ((rc = linear(e->FF, F, Mr, d.w2, H, F, d.b2, e->X1, 0, e->T, nullptr, st, ance::kClsGemmFfn2))) return rc;
    if ((rc = layer_norm(e->T, false, H, Mr, H, d.ln2g, d.ln2b, c.ln_eps, e->X, nullptr, st))) return rc;
    if (e->dbg && M <= e->dbg_tokens)  // with cls_only the first B rows hold the CLS rows of the last layer
      ANCE_CUDA(cudaMemcpyAsync(e->dbg + static_cast<size_t>(l + 1) * e->dbg_tokens * H, e->X, static_cast<size_t>(Mr) * H * 2, cudaMemcpyDeviceToDevice, st));
  }
  const size_t cls_pitch = e->prune_last_layer ? static_cast<size_t>(H) : static_cast<size_t>(L) * H;
  // K7: CLS rows (token 0 of every sequence) -> head
  if (c.has_head) {
    // A = the CLS rows of X ([B, H] compact after the pruned last layer, else row pitch L*H)
    if ((rc = linear(e->X, cls_pitch, B, e->head_w, H, H, e->head_b, nullptr, 0, nullptr, e->head_tmp, st))) return rc;
    if ((rc = layer_norm(e->head_tmp, true, H, B, H, e->head_g, e->head_bt, 1e-5f, nullptr, out_dev, st))) return rc;
  } else {
    ance::ProfScope ps(ance::kClsNorm, st);
    gather_rows_f32_kernel<<<B, 256, 0, st>>>(e->X, cls_pitch, B, H, out_dev);
    ANCE_CUDA(cudaGetLastError());
    ance::count_launch(1);
  }
  return ANCE_OK;
}

extern "C" int ance_encoder_set_param(ance_encoder_t e, const char* name, double value) {
  ANCE_REQUIRE(e != nullptr && name != nullptr, "ance_encoder_set_param: null argument");
  if (!strcmp(name, "prune_last_layer")) e->prune_last_layer = value != 0;
  else if (!strcmp(name, "ln_rows_per_warp")) g_ln_rows_per_warp = static_cast<int>(value);
  else { ance::set_error("ance_encoder_set_param: unknown parameter '%s'", name); return ANCE_ERR_INVALID; }
  return ANCE_OK;
}

extern "C" int ance_encoder_check(ance_encoder_t e, void* stream) {
  ANCE_REQUIRE(e != nullptr, "ance_encoder_check: null handle");
  int err = 0;
  ANCE_CUDA(cudaStreamSynchronize(reinterpret_cast<cudaStream_t>(stream)));
  ANCE_CUDA(cudaMemcpy(&err, e->err_flag, sizeof(int), cudaMemcpyDeviceToHost));
  if (err) {
    ANCE_CUDA(cudaMemset(e->err_flag, 0, sizeof(int)));
    ance::set_error("ance_encoder_forward: a token id outside [0, vocab_size) or a position past max_position_embeddings "
                    "was seen since the last check (the reference's embedding lookup raises an index error there)");
    return ANCE_ERR_INVALID;
  }
  return ANCE_OK;
}

extern "C" int ance_encoder_debug_hidden(ance_encoder_t e, int layer, float* out_dev, void* stream) {
  ANCE_REQUIRE(e != nullptr, "ance_encoder_debug_hidden: null handle");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t H = e->cfg.hidden;
  if (layer < 0) {  // enable capture for batches up to 4096 tokens
    if (!e->dbg) {
      e->dbg_tokens = std::min(e->max_tokens, 4096);
      e->dbg = dev_alloc<__nv_bfloat16>(e, static_cast<size_t>(e->cfg.n_layer + 1) * e->dbg_tokens * H);
      ANCE_REQUIRE(e->dbg != nullptr, "ance_encoder_debug_hidden: allocation failed");
    }
    return ANCE_OK;
  }
  ANCE_REQUIRE(e->dbg != nullptr, "ance_encoder_debug_hidden: capture not enabled (call with layer = -1 first)");
  ANCE_REQUIRE(layer <= e->cfg.n_layer && out_dev, "ance_encoder_debug_hidden: bad layer or null buffer");
  const size_t n = static_cast<size_t>(e->dbg_tokens) * H;
  bf16_to_f32_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(e->dbg + static_cast<size_t>(layer) * n, out_dev, n);
  ANCE_CUDA(cudaGetLastError());
  return ANCE_OK;
}
