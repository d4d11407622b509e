// CUDA-core fp32 relative-position attention for the validation path.  One warp per query row, the row of scores is
// kept in shared memory (never in HBM), exact two-pass softmax.  Semantics follow multi_head_attention.py:195-210 and
// :104-113: keys j >= len are excluded, query rows i >= len produce a zero context vector.
#include <math.h>

#include "common.cuh"

namespace cfb {
namespace {

constexpr int kWarps = 4;

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// qkv (B*T, 4*Dp) fp32: [q+u | q+v | k | v]; pos (2T-1, ld_pos) fp32; ctx (B*T, Dp) fp32
__global__ void __launch_bounds__(kWarps * 32) rel_attn_simt_kernel(const float* __restrict__ qkv,
                                                                    const float* __restrict__ pos, long long ld_pos,
                                                                    float* __restrict__ ctx,
                                                                    const int32_t* __restrict__ lens, int T, int H,
                                                                    int dk, int dkp, float sqrt_dk) {
  extern __shared__ float smem[];  // kWarps * (T + dkp)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z, h = blockIdx.y;
  const int i = blockIdx.x * kWarps + warp;
  if (i >= T) return;
  const int Dp = H * dkp;
  const int len = min(lens[b], T);
  float* sc = smem + warp * (T + 2 * dkp);
  float* qu = sc + T;
  float* qv = qu + dkp;
  const long long row = static_cast<long long>(b) * T + i;
  float* out = ctx + row * Dp + h * dkp;
  if (i >= len) {
    for (int c = lane; c < dkp; c += 32) out[c] = 0.f;
    return;
  }
  for (int c = lane; c < dkp; c += 32) {
    qu[c] = qkv[row * 4 * Dp + h * dkp + c];
    qv[c] = qkv[row * 4 * Dp + Dp + h * dkp + c];
  }
  __syncwarp();
  float mx = -INFINITY;
  for (int j = lane; j < len; j += 32) {
    const float* kr = qkv + (static_cast<long long>(b) * T + j) * 4 * Dp + 2 * Dp + h * dkp;
    const float* pr = pos + static_cast<long long>(T - 1 + j - i) * ld_pos + h * dkp;  // rel_shift as an index remap
    float ac = 0.f, bd = 0.f;
    for (int c = 0; c < dk; ++c) {
      ac = fmaf(qu[c], kr[c], ac);
      bd = fmaf(qv[c], pr[c], bd);
    }
    const float s = (ac + bd) / sqrt_dk;
    sc[j] = s;
    mx = fmaxf(mx, s);
  }
  mx = warp_max(mx);
  float sum = 0.f;
  for (int j = lane; j < len; j += 32) {
    const float e = expf(sc[j] - mx);
    sc[j] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  __syncwarp();
  const float inv = 1.f / sum;
  for (int c = lane; c < dkp; c += 32) {
    float acc = 0.f;
    if (c < dk) {
      const float* vcol = qkv + static_cast<long long>(b) * T * 4 * Dp + 3 * Dp + h * dkp + c;
      for (int j = 0; j < len; ++j) acc = fmaf(sc[j], vcol[static_cast<long long>(j) * 4 * Dp], acc);
    }
    out[c] = acc * inv;
  }
}

}  // namespace

int launch_attn_simt(const AttnDesc& a, cudaStream_t st, std::string* err) {
  if (a.B <= 0 || a.T <= 0) return 0;
  const size_t smem = static_cast<size_t>(kWarps) * (a.T + 2 * a.dkp) * sizeof(float);
  if (smem > 200 * 1024) {
    if (err) *err = "attn_simt: sequence too long for the validation kernel";
    return -1;
  }
  static bool configured[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!configured[dev & 63]) {
    cudaFuncSetAttribute(rel_attn_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    configured[dev & 63] = true;
  }
  dim3 grid((a.T + kWarps - 1) / kWarps, a.H, a.B);
  rel_attn_simt_kernel<<<grid, kWarps * 32, smem, st>>>(
      reinterpret_cast<const float*>(a.qkv), reinterpret_cast<const float*>(a.pos), a.ld_pos,
      reinterpret_cast<float*>(a.ctx), a.lens, a.T, a.H, a.dk, a.dkp, sqrtf(static_cast<float>(a.dk)));
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    if (err) *err = std::string("attn_simt launch: ") + cudaGetErrorString(e);
    return static_cast<int>(e);
  }
  return 0;
}

}  // namespace cfb
