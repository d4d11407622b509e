// CUDA-core fp32 GEMM for the validation path (cfb_precision = CFB_PREC_FP32_VALIDATE): fp32 operands, fp32
// accumulation, the same epilogues as the tensor-core kernel (applied by a second kernel over the raw
// accumulators).  Slow by design -- it exists so that every tcgen05 kernel has an independent witness on the GPU.
#include "common.cuh"
#include "epilogue.cuh"

namespace cfb {
namespace {

constexpr int TM = 64, TN = 64, TK = 16;

// C[M][N] (ldc = N) = A[M][K] * W[N][K]^T
__global__ void __launch_bounds__(256) gemm_raw_kernel(const float* __restrict__ A, long long lda,
                                                       const float* __restrict__ W, long long ldw,
                                                       float* __restrict__ C, int M, int N, int K) {
  __shared__ float As[TK][TM + 4];
  __shared__ float Ws[TK][TN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  const int tx = tid % 16, ty = tid / 16;  // 16 x 16 threads, 4 x 4 outputs each
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += TK) {
    for (int i = tid; i < TM * TK; i += 256) {
      const int r = i / TK, c = i % TK;
      const int gm = m0 + r, gk = k0 + c;
      As[c][r] = (gm < M && gk < K) ? A[static_cast<long long>(gm) * lda + gk] : 0.f;
      const int gn = n0 + r;
      Ws[c][r] = (gn < N && gk < K) ? W[static_cast<long long>(gn) * ldw + gk] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TK; ++k) {
      float a[4], w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) w[j] = Ws[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn < N) C[static_cast<long long>(gm) * N + gn] = acc[i][j];
    }
  }
}

template <int EPI, typename TOut>
__global__ void __launch_bounds__(128) epilogue_kernel(const float* __restrict__ C, EpiParams p) {
  const int chunks = (p.N + 31) / 32;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<long long>(p.M) * chunks) return;
  const long long row = idx / chunks;
  const int col0 = static_cast<int>(idx % chunks) * 32;
  float acc[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) acc[j] = (col0 + j < p.N) ? C[row * p.N + col0 + j] : 0.f;
  epi_apply<EPI, TOut>(p, row, col0, acc);
}

template <int EPI, typename TOut>
void run_epilogue(const float* C, const EpiParams& p, cudaStream_t st) {
  const long long total = static_cast<long long>(p.M) * ((p.N + 31) / 32);
  const int blocks = static_cast<int>((total + 127) / 128);
  epilogue_kernel<EPI, TOut><<<blocks, 128, 0, st>>>(C, p);
}

}  // namespace

int launch_gemm_simt(const GemmDesc& g, float* scratch, cudaStream_t st, std::string* err) {
  if (g.M <= 0 || g.N <= 0) return 0;
  dim3 grid((g.N + TN - 1) / TN, (g.M + TM - 1) / TM);
  gemm_raw_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const float*>(g.A), g.lda,
                                        reinterpret_cast<const float*>(g.W), g.ldw, scratch, g.M, g.N, g.K);
  EpiParams p = g.ep;
  p.M = g.M;
  p.N = g.N;
  const bool b = g.out_bf16;
  switch (g.epi) {
    case EPI_LINEAR:
      b ? run_epilogue<EPI_LINEAR, bf16>(scratch, p, st) : run_epilogue<EPI_LINEAR, float>(scratch, p, st);
      break;
    case EPI_SWISH:
      b ? run_epilogue<EPI_SWISH, bf16>(scratch, p, st) : run_epilogue<EPI_SWISH, float>(scratch, p, st);
      break;
    case EPI_RELU:
      b ? run_epilogue<EPI_RELU, bf16>(scratch, p, st) : run_epilogue<EPI_RELU, float>(scratch, p, st);
      break;
    case EPI_RESID:
      run_epilogue<EPI_RESID, float>(scratch, p, st);
      break;
    case EPI_QKV:
      b ? run_epilogue<EPI_QKV, bf16>(scratch, p, st) : run_epilogue<EPI_QKV, float>(scratch, p, st);
      break;
    case EPI_GLU:
      b ? run_epilogue<EPI_GLU, bf16>(scratch, p, st) : run_epilogue<EPI_GLU, float>(scratch, p, st);
      break;
    default:
      if (err) *err = "gemm_simt: unknown epilogue";
      return -1;
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    if (err) *err = std::string("gemm_simt launch: ") + cudaGetErrorString(e);
    return static_cast<int>(e);
  }
  return 0;
}

}  // namespace cfb
