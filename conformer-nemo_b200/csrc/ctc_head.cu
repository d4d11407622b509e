// CTC head on the encoder output: ConvASRDecoder = 1x1 Conv1d (d -> V+1) + log_softmax over
// the classes (modules/conv_asr.py:437-444), plus the greedy argmax the CTC models take right after it
// (models/ctc_models.py:593-594).  The projection runs on the tcgen05 GEMM (bf16 operands, fp32 logits); this file
// holds the two memory-bound helpers around it and the C-ABI entry point.
#include <float.h>

#include "../../include/cfb.h"
#include "common.cuh"

namespace cfb {
namespace {

__global__ void __launch_bounds__(256) cast_f32_bf16_kernel(const float* __restrict__ x, bf16* __restrict__ y, long long n4) {
  pdl_launch_dependents();
  pdl_wait();
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
    __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&lo);
    u.y = *reinterpret_cast<uint32_t*>(&hi);
    reinterpret_cast<uint2*>(y)[i] = u;
  }
}

// One warp per frame: log_probs[r, :] = logits[r, :] - max - log(sum exp(logits - max)); argmax = first index of the max
// (torch.argmax's tie rule).  Two passes over the row (it stays in L1/L2 between them).
__global__ void __launch_bounds__(256) logsoftmax_argmax_kernel(const float* __restrict__ logits, long long ldl,
                                                                float* __restrict__ logprobs, int32_t* __restrict__ best,
                                                                int rows, int v1) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  pdl_launch_dependents();
  pdl_wait();
  if (row >= rows) return;
  const float* in = logits + static_cast<long long>(row) * ldl;
  float m = -FLT_MAX;
  int mi = 0x7fffffff;
  for (int c = lane; c < v1; c += 32) {
    const float v = in[c];
    if (v > m) {  // strict: keeps the first index within the lane (columns ascend)
      m = v;
      mi = c;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, m, o);
    const int oi = __shfl_xor_sync(0xffffffffu, mi, o);
    if (om > m || (om == m && oi < mi)) {
      m = om;
      mi = oi;
    }
  }
  float s = 0.f;
  for (int c = lane; c < v1; c += 32) s += expf(in[c] - m);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float lse = m + logf(s);
  float* out = logprobs + static_cast<long long>(row) * v1;
  for (int c = lane; c < v1; c += 32) out[c] = in[c] - lse;
  if (lane == 0 && best != nullptr) best[row] = mi;
}

// Greedy CTC collapse on the device (metrics/wer.py:152-164): per utterance keep frame t < len iff its arg-max p is not the
// blank and differs from the previous frame's (the reference's `previous` starts as the blank).  One CTA per utterance:
// keep flags -> block-wide exclusive scan (warp shuffles + one shared array) -> stable compaction in frame order.
__global__ void __launch_bounds__(256) ctc_collapse_kernel(const int32_t* __restrict__ best, const int32_t* __restrict__ lens,
                                                           int T, int blank, int32_t* __restrict__ tokens,
                                                           int32_t* __restrict__ n_tokens) {
  __shared__ int s_warp[8];
  __shared__ int s_base;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int32_t* in = best + static_cast<long long>(b) * T;
  int32_t* out = tokens + static_cast<long long>(b) * T;
  int len = lens != nullptr ? lens[b] : T;
  len = len < 0 ? 0 : (len > T ? T : len);
  if (tid == 0) s_base = 0;
  __syncthreads();
  for (int t0 = 0; t0 < len; t0 += 256) {
    const int t = t0 + tid;
    int p = blank, prev = blank;
    if (t < len) {
      p = in[t];
      prev = t > 0 ? in[t - 1] : blank;
    }
    const int keep = (t < len && p != blank && p != prev) ? 1 : 0;
    int incl = keep;  // inclusive scan inside the warp
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    int before = s_base;
    for (int w = 0; w < warp; ++w) before += s_warp[w];
    if (keep) out[before + incl - 1] = p;
    __syncthreads();
    if (tid == 0) {
      int total = 0;
      for (int w = 0; w < 8; ++w) total += s_warp[w];
      s_base += total;
    }
    __syncthreads();
  }
  if (tid == 0) n_tokens[b] = s_base;
}

std::string g_ctc_error;

}  // namespace
}  // namespace cfb

using namespace cfb;

extern "C" {

size_t cfb_ctc_head_scratch_bytes(int M, int d, int v1) {
  const size_t ldl = (static_cast<size_t>(v1) + 3) / 4 * 4;
  return (static_cast<size_t>(M) * d * 2 + 255) / 256 * 256 + static_cast<size_t>(M) * ldl * 4 + 256;
}

int cfb_op_ctc_head(const void* x, int x_dtype, const void* W, const float* bias, int M, int d, int v1, float* logprobs,
                    int32_t* best, void* scratch, size_t scratch_bytes, cfb_stream stream) {
  if (!x || !W || !logprobs || !scratch || M < 1 || d < 8 || v1 < 1 || (d % 8) != 0) return CFB_ERR_INVALID_ARG;
  if (x_dtype != CFB_F32 && x_dtype != CFB_BF16) return CFB_ERR_INVALID_ARG;
  if (scratch_bytes < cfb_ctc_head_scratch_bytes(M, d, v1) || (reinterpret_cast<uintptr_t>(scratch) & 255))
    return CFB_ERR_WORKSPACE;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  uint8_t* ws = reinterpret_cast<uint8_t*>(scratch);
  const void* a = x;
  size_t off = 0;
  if (x_dtype == CFB_F32) {
    const long long n4 = static_cast<long long>(M) * d / 4;
    int blocks = static_cast<int>((n4 + 255) / 256 < 148 * 8 ? (n4 + 255) / 256 : 148 * 8);
    launch_pdl(cast_f32_bf16_kernel, dim3(blocks), dim3(256), 0, st, reinterpret_cast<const float*>(x),
               reinterpret_cast<bf16*>(ws), n4);
    a = ws;
    off = (static_cast<size_t>(M) * d * 2 + 255) / 256 * 256;
  }
  const long long ldl = (static_cast<long long>(v1) + 3) / 4 * 4;
  float* logits = reinterpret_cast<float*>(ws + off);
  GemmDesc g;
  g.A = a;
  g.lda = d;
  g.W = W;
  g.ldw = d;
  g.M = M;
  g.N = v1;
  g.K = d;
  g.epi = EPI_LINEAR;
  g.out_bf16 = false;
  g.ep.bias = bias;
  g.ep.out = logits;
  g.ep.ldo = ldl;
  std::string err;
  int rc = launch_gemm_tc(g, st, &err);
  if (rc != 0) return CFB_ERR_CUDA;
  launch_pdl(logsoftmax_argmax_kernel, dim3((M + 7) / 8), dim3(256), 0, st, static_cast<const float*>(logits), ldl, logprobs,
             best, M, v1);
  return cudaGetLastError() == cudaSuccess ? CFB_OK : CFB_ERR_CUDA;
}

int cfb_op_ctc_collapse(const int32_t* best, const int32_t* lens, int B, int T, int blank, int32_t* tokens, int32_t* n_tokens,
                        cfb_stream stream) {
  if (!best || !tokens || !n_tokens || B < 1 || T < 1) return CFB_ERR_INVALID_ARG;
  ctc_collapse_kernel<<<B, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(best, lens, T, blank, tokens, n_tokens);
  return cudaGetLastError() == cudaSuccess ? CFB_OK : CFB_ERR_CUDA;
}

}  // extern "C"
