// Fused GEMM epilogues.  One call handles 32 consecutive accumulator columns of one output row, all held by one
// thread -- exactly what a tcgen05.ld.32x32b.x32 hands an epilogue thread, and what the validation path's
// stand-alone epilogue kernel reads back from the raw accumulator buffer.  All arithmetic is fp32; the value is
// rounded once when stored.
#pragma once
#include "common.cuh"

namespace cfb {

template <typename T>
struct OutTraits;
template <>
struct OutTraits<float> {
  static constexpr bool kFast = false;  // validation path: accurate expf
};
template <>
struct OutTraits<bf16> {
  static constexpr bool kFast = true;
};

__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// sigmoid(x) = 0.5 + 0.5 tanh(x/2): one MUFU op on the fast (bf16-output) path instead of ex2 + rcp
template <bool kFast>
__device__ __forceinline__ float sigmoidf_(float x) {
  if constexpr (kFast) {
    return fmaf(0.5f, tanh_approx(0.5f * x), 0.5f);
  } else {
    return 1.f / (1.f + expf(-x));
  }
}
// swish(x) = x sigmoid(x) = h + h tanh(h), h = x/2
template <bool kFast>
__device__ __forceinline__ float swishf_(float x) {
  if constexpr (kFast) {
    const float h = 0.5f * x;
    return fmaf(h, tanh_approx(h), h);
  } else {
    return x / (1.f + expf(-x));
  }
}

__device__ __forceinline__ void store_run(float* dst, const float* v, int n) {  // n % 4 == 0, dst 16 B aligned
#pragma unroll
  for (int j = 0; j < 32; j += 4)
    if (j < n) *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
}
__device__ __forceinline__ void store_run(bf16* dst, const float* v, int n) {  // n % 8 == 0, dst 16 B aligned
#pragma unroll
  for (int j = 0; j < 32; j += 8)
    if (j < n) {
      uint4 u;
      __nv_bfloat162 a = __floats2bfloat162_rn(v[j], v[j + 1]);
      __nv_bfloat162 b = __floats2bfloat162_rn(v[j + 2], v[j + 3]);
      __nv_bfloat162 c = __floats2bfloat162_rn(v[j + 4], v[j + 5]);
      __nv_bfloat162 d = __floats2bfloat162_rn(v[j + 6], v[j + 7]);
      u.x = *reinterpret_cast<uint32_t*>(&a);
      u.y = *reinterpret_cast<uint32_t*>(&b);
      u.z = *reinterpret_cast<uint32_t*>(&c);
      u.w = *reinterpret_cast<uint32_t*>(&d);
      *reinterpret_cast<uint4*>(dst + j) = u;
    }
}
template <typename T>
__device__ __forceinline__ void store_tail(T* dst, const float* v, int n) {
  for (int j = 0; j < n; ++j) {
    if constexpr (sizeof(T) == 4)
      dst[j] = v[j];
    else
      dst[j] = __float2bfloat16_rn(v[j]);
  }
}
template <typename T>
__device__ __forceinline__ void store_cols(T* dst, const float* v, int n) {  // n <= 32 valid values
  if ((n & 7) == 0)
    store_run(dst, v, n);
  else
    store_tail(dst, v, n);
}

__device__ __forceinline__ void load_bias32(const float* bias, int col0, int n, float (&b)[32]) {
  if (bias == nullptr) {
#pragma unroll
    for (int j = 0; j < 32; ++j) b[j] = 0.f;
    return;
  }
  if (n == 32) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      float4 t = __ldg(reinterpret_cast<const float4*>(bias + col0 + j));
      b[j] = t.x, b[j + 1] = t.y, b[j + 2] = t.z, b[j + 3] = t.w;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) b[j] = (j < n) ? __ldg(bias + col0 + j) : 0.f;
  }
}

// Arithmetic part of the epilogue.  acc: 32 accumulator columns [col0, col0+32) of the output row `out_row`
// (masking uses out_row / frames_per_seq).  On return acc holds the values to store:
//   LINEAR / SWISH / RELU : 32 values for output columns col0..          (zeroed at padded frames when p.lens is set)
//   RESID                 : 32 increments alpha * (acc + bias) to be ADDED to the fp32 residual stream
//   QKV                   : pass 0 -> acc + bias (q+u | k | v), pass 1 -> acc + bias2 (q+v; only for col0 < qkv_dp)
//   GLU                   : 16 values a * sigmoid(g) for output columns col0/2.. (zeroed at padded frames)
// epi_math: same, with the 32 bias values already in registers (the tensor-core kernels stage a tile's bias in shared
// memory while they wait for the accumulator: a per-chunk global load is an L1 miss per tile and was measured to be
// ~300 of the ~650 cycles an epilogue chunk took).
template <int EPI, bool kFast>
__device__ __forceinline__ void epi_math(const EpiParams& p, long long out_row, float (&acc)[32], const float (&b)[32]);

template <int EPI, bool kFast>
__device__ __forceinline__ void epi_compute(const EpiParams& p, long long out_row, int col0, float (&acc)[32],
                                            int pass) {
  const int n = min(32, p.N - col0);
  float b[32];
  load_bias32((EPI == EPI_QKV && pass == 1) ? p.bias2 : p.bias, col0, n, b);
  epi_math<EPI, kFast>(p, out_row, acc, b);
}

// Is output row `out_row` a valid frame (false: the masked epilogues store zeros)?  A dependent global load: the
// tensor-core kernels ask once per tile BEFORE they wait for the accumulator (epi_math_keep), not in front of the
// first arithmetic of the tile.
template <int EPI>
__device__ __forceinline__ bool epi_keep(const EpiParams& p, long long out_row) {
  bool keep = true;
  if constexpr (EPI == EPI_LINEAR || EPI == EPI_SWISH || EPI == EPI_RELU || EPI == EPI_GLU) {
    if (p.row_t != nullptr) {
      keep = p.row_t[out_row] >= 0;
    } else if (p.lens != nullptr) {
      const int seq = static_cast<int>(out_row / p.frames_per_seq);
      const int t = static_cast<int>(out_row - static_cast<long long>(seq) * p.frames_per_seq);
      keep = t < p.lens[seq];
    }
  }
  return keep;
}

template <int EPI, bool kFast>
__device__ __forceinline__ void epi_math_keep(const EpiParams& p, bool keep, float (&acc)[32], const float (&b)[32]);

template <int EPI, bool kFast>
__device__ __forceinline__ void epi_math(const EpiParams& p, long long out_row, float (&acc)[32], const float (&b)[32]) {
  epi_math_keep<EPI, kFast>(p, epi_keep<EPI>(p, out_row), acc, b);
}

template <int EPI, bool kFast>
__device__ __forceinline__ void epi_math_keep(const EpiParams& p, bool keep, float (&acc)[32], const float (&b)[32]) {
  if constexpr (EPI == EPI_LINEAR || EPI == EPI_SWISH || EPI == EPI_RELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      float x = acc[j] + b[j];
      if constexpr (EPI == EPI_SWISH) x = swishf_<kFast>(x);
      if constexpr (EPI == EPI_RELU) x = fmaxf(x, 0.f);
      acc[j] = keep ? x : 0.f;
    }
  } else if constexpr (EPI == EPI_RESID) {
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[j] = fmaf(p.alpha, acc[j], p.alpha * b[j]);
  } else if constexpr (EPI == EPI_QKV) {
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[j] = acc[j] + b[j];
  } else if constexpr (EPI == EPI_GLU) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float a = acc[j] + b[j];
      const float g = acc[16 + j] + b[16 + j];
      acc[j] = keep ? a * sigmoidf_<kFast>(g) : 0.f;
    }
  }
}

// Direct-to-global form used by the validation path's stand-alone epilogue kernel.
template <int EPI, typename TOut>
__device__ __forceinline__ void epi_apply(const EpiParams& p, long long out_row, int col0, float (&acc)[32]) {
  constexpr bool kFast = OutTraits<TOut>::kFast;
  const int n = min(32, p.N - col0);
  if (n <= 0) return;
  if constexpr (EPI == EPI_RESID) {
    epi_compute<EPI, kFast>(p, out_row, col0, acc, 0);
    float* r = reinterpret_cast<float*>(p.out) + out_row * p.ldo + col0;
    for (int j = 0; j < n; ++j) r[j] += acc[j];
  } else if constexpr (EPI == EPI_QKV) {
    TOut* o = reinterpret_cast<TOut*>(p.out) + out_row * p.ldo;
    if (col0 < p.qkv_dp) {
      float v2[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v2[j] = acc[j];
      epi_compute<EPI, kFast>(p, out_row, col0, acc, 0);
      epi_compute<EPI, kFast>(p, out_row, col0, v2, 1);
      store_cols(o + col0, acc, n);
      store_cols(o + p.qkv_dp + col0, v2, n);
    } else {
      epi_compute<EPI, kFast>(p, out_row, col0, acc, 0);
      store_cols(o + p.qkv_dp + col0, acc, n);
    }
  } else if constexpr (EPI == EPI_GLU) {
    epi_compute<EPI, kFast>(p, out_row, col0, acc, 0);
    store_cols(reinterpret_cast<TOut*>(p.out) + out_row * p.ldo + (col0 >> 1), acc, 16);
  } else {
    epi_compute<EPI, kFast>(p, out_row, col0, acc, 0);
    store_cols(reinterpret_cast<TOut*>(p.out) + out_row * p.ldo + col0, acc, n);
  }
}

}  // namespace cfb
