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

template <bool kFast>
__device__ __forceinline__ float sigmoidf_(float x) {
  if constexpr (kFast) {
    return __fdividef(1.f, 1.f + __expf(-x));
  } else {
    return 1.f / (1.f + expf(-x));
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

// acc: 32 accumulator columns [col0, col0+32) of logical row `row` (0 <= row < p.M, checked by the caller);
// out_row: the row of the output matrix it maps to (differs from `row` only for the implicit-GEMM convolution).
template <int EPI, typename TOut>
__device__ __forceinline__ void epi_apply(const EpiParams& p, long long out_row, int col0, float (&acc)[32]) {
  constexpr bool kFast = OutTraits<TOut>::kFast;
  const int n = min(32, p.N - col0);
  if (n <= 0) return;
  float b[32];
  load_bias32(p.bias, col0, n, b);

  if constexpr (EPI == EPI_LINEAR || EPI == EPI_SWISH || EPI == EPI_RELU) {
    bool keep = true;
    if (p.lens != nullptr) {
      const int seq = static_cast<int>(out_row / p.frames_per_seq);
      const int t = static_cast<int>(out_row - static_cast<long long>(seq) * p.frames_per_seq);
      keep = t < p.lens[seq];
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      float x = acc[j] + b[j];
      if constexpr (EPI == EPI_SWISH) x = x * sigmoidf_<kFast>(x);
      if constexpr (EPI == EPI_RELU) x = fmaxf(x, 0.f);
      acc[j] = keep ? x : 0.f;
    }
    store_cols(reinterpret_cast<TOut*>(p.out) + out_row * p.ldo + col0, acc, n);
  } else if constexpr (EPI == EPI_RESID) {
    float* r = reinterpret_cast<float*>(p.out) + out_row * p.ldo + col0;
    if (n == 32) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        float4 t = *reinterpret_cast<const float4*>(r + j);
        t.x += p.alpha * (acc[j] + b[j]);
        t.y += p.alpha * (acc[j + 1] + b[j + 1]);
        t.z += p.alpha * (acc[j + 2] + b[j + 2]);
        t.w += p.alpha * (acc[j + 3] + b[j + 3]);
        *reinterpret_cast<float4*>(r + j) = t;
      }
    } else {
      for (int j = 0; j < n; ++j) r[j] += p.alpha * (acc[j] + b[j]);
    }
  } else if constexpr (EPI == EPI_QKV) {
    // accumulator columns: [0,Dp) q, [Dp,2Dp) k, [2Dp,3Dp) v ; output columns: [q+u | q+v | k | v]
    TOut* o = reinterpret_cast<TOut*>(p.out) + out_row * p.ldo;
    if (col0 < p.qkv_dp) {
      float b2[32], v2[32];
      load_bias32(p.bias2, col0, n, b2);
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        v2[j] = acc[j] + b2[j];
        acc[j] = acc[j] + b[j];
      }
      store_cols(o + col0, acc, n);
      store_cols(o + p.qkv_dp + col0, v2, n);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) acc[j] = acc[j] + b[j];
      store_cols(o + p.qkv_dp + col0, acc, n);
    }
  } else if constexpr (EPI == EPI_GLU) {
    // accumulator columns come in groups of 32 = [16 'a' channels | the matching 16 gate channels]
    bool keep = true;
    if (p.lens != nullptr) {
      const int seq = static_cast<int>(out_row / p.frames_per_seq);
      const int t = static_cast<int>(out_row - static_cast<long long>(seq) * p.frames_per_seq);
      keep = t < p.lens[seq];
    }
    float o[32];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float a = acc[j] + b[j];
      const float g = acc[16 + j] + b[16 + j];
      o[j] = keep ? a * sigmoidf_<kFast>(g) : 0.f;
    }
    store_cols(reinterpret_cast<TOut*>(p.out) + out_row * p.ldo + (col0 >> 1), o, 16);
  }
}

}  // namespace cfb
