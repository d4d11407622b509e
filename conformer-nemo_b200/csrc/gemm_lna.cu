// tcgen05 GEMM whose A operand is produced in shared memory by a LayerNorm prologue and stays resident:
//
//   [y = LN1(x); x_out = y]                 optional: norm_out of the previous layer (conformer_modules.py:120)
//   a = bf16(LN2(y or x))                   the block's input LayerNorm (:98, :103, :112, :116)
//   D = a W^T  with the fused epilogues of epilogue.cuh (Swish / QKV split / GLU + mask)
//
// One CTA owns a 128-row block: the eight epilogue warps first normalise the block's rows (one warp per row, row in
// registers, two-pass statistics exactly like layernorm_kernel) and write them as bf16 into the UMMA K-major
// 128-byte-swizzled layout, 16 KB per 64-column k-block.  The block then runs ALL its N tiles against that resident
// operand: only W streams through the TMA ring, so the L2 -> SM operand stream per k-block drops from A + W (48 KB) to
// W (32 KB), the bf16 copy of LN(x) never exists in HBM, and the stand-alone LayerNorm launch disappears.
//   warp 0      TMA producer (W boxes only)      warp 1      MMA issuer (A descriptors point into the resident tile)
//   warps 2..9  LayerNorm prologue, then the epilogue of every N tile (accumulators double-buffered in TMEM)
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "epilogue.cuh"
#include "ptx.cuh"

namespace cfb {
namespace {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kBN = 256;
constexpr int kEpiWarps = 8;
constexpr int kThreads = 64 + 32 * kEpiWarps;
constexpr int kMaxStages = 6;
constexpr int kKbBytes = kBlockM * kBlockK * 2;  // one resident k-block of A: 16 KB
constexpr int kBBytes = kBN * kBlockK * 2;       // one W stage: 32 KB
constexpr int kStagingBytes = kEpiWarps * 4096;  // one output box per epilogue warp
constexpr int kMaxVec = 4;                       // float4 per lane: d <= 512

struct LnaParams {
  int num_m_blocks, num_n_tiles, num_k_blocks, stages;
  int d;                 // K = LayerNorm width (d % 4 == 0, d <= 512)
  int off_b, off_staging, off_bar, smem_needed;
  const float* x;        // (M, d) fp32
  long long ldx;
  const float* gamma1;   // optional first LayerNorm (norm_out): y = LN1(x) is written back to x_out
  const float* beta1;
  float* x_out;
  const float* gamma2;   // the LayerNorm whose output is the A operand
  const float* beta2;
  EpiParams ep;
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ void sts64(uint32_t addr, uint32_t a, uint32_t b) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void tma_store_2d_a(const void* tmap, uint32_t smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(tmap)),
               "r"(smem_src), "r"(c0), "r"(c1)
               : "memory");
}

template <int EPI>
__global__ void __launch_bounds__(kThreads, 1)
gemm_lna_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmO, const LnaParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  {
    uint32_t dyn;
    asm volatile("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn));
    if (sbase - ptx::smem_u32(smem_raw) + p.smem_needed > dyn) __trap();  // alignment slack did not fit
  }
  const uint32_t bar0 = sbase + p.off_bar;
  const uint32_t full_bar = bar0;                     // [kMaxStages]
  const uint32_t empty_bar = bar0 + 8 * kMaxStages;   // [kMaxStages]
  const uint32_t acc_full = bar0 + 16 * kMaxStages;   // [2]
  const uint32_t acc_empty = acc_full + 16;           // [2]
  const uint32_t a_ready = acc_empty + 16;            // the resident A tile of the current row block is written
  const uint32_t tmem_slot = a_ready + 8;
  uint32_t tid;
  asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid));
  const int warp = tid >> 5;
  const int lane = tid & 31;

  if (warp == 0) {
    if (lane == 0) {
      ptx::prefetch_tmap(&tmW);
      ptx::prefetch_tmap(&tmO);
      for (int s = 0; s < kMaxStages; ++s) {
        ptx::mbar_init_a(full_bar + 8 * s, 1);
        ptx::mbar_init_a(empty_bar + 8 * s, 1);
      }
      for (int b = 0; b < 2; ++b) {
        ptx::mbar_init_a(acc_full + 8 * b, 1);
        ptx::mbar_init_a(acc_empty + 8 * b, 32 * kEpiWarps);
      }
      ptx::mbar_init_a(a_ready, 32 * kEpiWarps);
      ptx::fence_mbar_init();
    }
    __syncwarp();
  } else if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  } else {
    // zero the resident tile once: row / column tails (M % 128, d % 64) are never written by the prologue
    const uint32_t a_bytes = static_cast<uint32_t>(p.num_k_blocks) * kKbBytes;
    for (uint32_t o = (tid - 64) * 16; o < a_bytes; o += 32 * kEpiWarps * 16) ptx::sts128(sbase + o, 0u, 0u, 0u, 0u);
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer: W only, runs ahead freely
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int mb = blockIdx.x; mb < p.num_m_blocks; mb += gridDim.x) {
        for (int nt = 0; nt < p.num_n_tiles; ++nt) {
          for (int kb = 0; kb < p.num_k_blocks; ++kb) {
            ptx::mbar_wait_a(empty_bar + 8 * stage, phase ^ 1);
            ptx::mbar_arrive_expect_tx_a(full_bar + 8 * stage, kBBytes);
            ptx::tma_load_2d_a(sbase + p.off_b + stage * kBBytes, &tmW, full_bar + 8 * stage, kb * kBlockK, nt * kBN);
            if (++stage == p.stages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = ptx::make_idesc_bf16(kBlockM, kBN, 0, 0);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;   // N tiles issued so far (accumulator buffer = it & 1)
    int blk = 0;  // row blocks issued so far
    for (int mb = blockIdx.x; mb < p.num_m_blocks; mb += gridDim.x, ++blk) {
      ptx::mbar_wait_a(a_ready, blk & 1);
      ptx::tc_fence_after();
      for (int nt = 0; nt < p.num_n_tiles; ++nt, ++it) {
        const int buf = it & 1;
        ptx::mbar_wait_a(acc_empty + 8 * buf, ((it >> 1) & 1) ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * kBN;
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          ptx::mbar_wait_a(full_bar + 8 * stage, phase);
          ptx::tc_fence_after();
          const uint64_t da = ptx::make_sdesc_sw128(sbase + kb * kKbBytes, 16, 1024);
          const uint64_t db = ptx::make_sdesc_sw128(sbase + p.off_b + stage * kBBytes, 16, 1024);
          if (ptx::elect_one()) {
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k)
              ptx::umma_bf16(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            ptx::tc_commit_a(empty_bar + 8 * stage);
            if (kb == p.num_k_blocks - 1) ptx::tc_commit_a(acc_full + 8 * buf);
          }
          __syncwarp();
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ prologue + epilogue warps
    constexpr bool kFast = true;
    constexpr int kBoxCols = 64;                                            // bf16 output columns per 128-byte box row
    constexpr int kAccPerBox = (EPI == EPI_GLU) ? 2 * kBoxCols : kBoxCols;  // accumulator columns feeding one box
    constexpr int kChunks = kAccPerBox / 32;
    constexpr int kBoxes = kBN / kAccPerBox;
    const int ew = warp - 2;
    const int quarter = warp & 3;
    const int half = ew >> 2;
    const int row_in_tile = quarter * 32 + lane;
    const uint32_t sbuf = sbase + p.off_staging + ew * 4096;
    const uint32_t srow = sbuf + lane * 128;
    const uint32_t swz = static_cast<uint32_t>(lane & 7);
    const int d = p.d;
    const int nvec = d >> 2;
    const float inv_d = 1.0f / static_cast<float>(d);
    int it = 0;
    for (int mb = blockIdx.x; mb < p.num_m_blocks; mb += gridDim.x) {
      // ---- LayerNorm prologue: warp ew normalises rows ew, ew + 8, ... of the block into the resident A tile.
      // (The previous block's MMAs have completed: this warp has waited for the acc_full of its last N tile.)
      float4 nx[kMaxVec];
      auto fetch = [&](int r) {
        const long long row = static_cast<long long>(mb) * kBlockM + r;
        if (row < p.ep.M) {
          const float4* xr = reinterpret_cast<const float4*>(p.x + row * p.ldx);
#pragma unroll
          for (int k = 0; k < kMaxVec; ++k)
            if (lane + 32 * k < nvec) nx[k] = __ldg(xr + lane + 32 * k);
        }
      };
      fetch(ew);
      for (int r = ew; r < kBlockM; r += kEpiWarps) {
        const long long row = static_cast<long long>(mb) * kBlockM + r;
        float4 v[kMaxVec];
#pragma unroll
        for (int k = 0; k < kMaxVec; ++k) v[k] = nx[k];
        if (r + kEpiWarps < kBlockM) fetch(r + kEpiWarps);
        if (row >= p.ep.M) continue;  // tail rows keep the zeros written at kernel start
        auto stats = [&](float& mean, float& rstd) {
          float s = 0.f;
#pragma unroll
          for (int k = 0; k < kMaxVec; ++k)
            if (lane + 32 * k < nvec) s += (v[k].x + v[k].y) + (v[k].z + v[k].w);
          mean = warp_sum(s) * inv_d;
          float q = 0.f;
#pragma unroll
          for (int k = 0; k < kMaxVec; ++k)
            if (lane + 32 * k < nvec) {
              v[k].x -= mean, v[k].y -= mean, v[k].z -= mean, v[k].w -= mean;
              q += (v[k].x * v[k].x + v[k].y * v[k].y) + (v[k].z * v[k].z + v[k].w * v[k].w);
            }
          rstd = 1.0f / sqrtf(warp_sum(q) * inv_d + 1e-5f);
        };
        float mean, rstd;
        stats(mean, rstd);
        if (p.gamma1 != nullptr) {
          // y = LN1(x) goes back to the fp32 residual stream and is normalised again for the GEMM operand
          float4* yo = reinterpret_cast<float4*>(p.x_out + row * p.ldx);
#pragma unroll
          for (int k = 0; k < kMaxVec; ++k) {
            const int i = lane + 32 * k;
            if (i < nvec) {
              const float4 g = __ldg(reinterpret_cast<const float4*>(p.gamma1) + i);
              const float4 b = __ldg(reinterpret_cast<const float4*>(p.beta1) + i);
              v[k].x = fmaf(v[k].x * rstd, g.x, b.x), v[k].y = fmaf(v[k].y * rstd, g.y, b.y);
              v[k].z = fmaf(v[k].z * rstd, g.z, b.z), v[k].w = fmaf(v[k].w * rstd, g.w, b.w);
              yo[i] = v[k];
            }
          }
          stats(mean, rstd);
        }
#pragma unroll
        for (int k = 0; k < kMaxVec; ++k) {
          const int i = lane + 32 * k;  // float4 index: columns 4 i .. 4 i + 3
          if (i < nvec) {
            const float4 g = __ldg(reinterpret_cast<const float4*>(p.gamma2) + i);
            const float4 b = __ldg(reinterpret_cast<const float4*>(p.beta2) + i);
            const uint32_t lo = ptx::pack_bf16x2(fmaf(v[k].x * rstd, g.x, b.x), fmaf(v[k].y * rstd, g.y, b.y));
            const uint32_t hi = ptx::pack_bf16x2(fmaf(v[k].z * rstd, g.z, b.z), fmaf(v[k].w * rstd, g.w, b.w));
            // K-major SWIZZLE_128B tile: k-block = col / 64, 128-byte row r, 16-byte chunk ((col % 64) / 8) ^ (r & 7)
            const int kb = i >> 4;
            const uint32_t chunk = static_cast<uint32_t>((i & 15) >> 1);
            sts64(sbase + kb * kKbBytes + r * 128 + ((chunk ^ static_cast<uint32_t>(r & 7)) << 4) + (i & 1) * 8, lo, hi);
          }
        }
      }
      ptx::fence_proxy_async_smem();  // generic-proxy writes of A -> visible to the tensor core's operand reads
      ptx::mbar_arrive_a(a_ready);

      // ---- epilogue of every N tile of this row block
      const long long out_row = static_cast<long long>(mb) * kBlockM + row_in_tile;
      const bool row_ok = out_row < p.ep.M;
      const int row0 = mb * kBlockM + quarter * 32;
      for (int nt = 0; nt < p.num_n_tiles; ++nt, ++it) {
        const int buf = it & 1;
        ptx::mbar_wait_a(acc_full + 8 * buf, (it >> 1) & 1);
        ptx::tc_fence_after();
        const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + buf * kBN;
#pragma unroll 1
        for (int box = half; box < kBoxes; box += kEpiWarps / 4) {
          const int acc_col0 = nt * kBN + box * kAccPerBox;
          if (acc_col0 >= p.ep.N) break;  // warp-uniform: nothing of this box is inside the matrix
          const int n_pass = (EPI == EPI_QKV && acc_col0 < p.ep.qkv_dp) ? 2 : 1;
#pragma unroll 1
          for (int pass = 0; pass < n_pass; ++pass) {
            if (lane == 0) ptx::bulk_wait_read<0>();  // the store that last read this buffer has drained it
            __syncwarp();
#pragma unroll
            for (int ch = 0; ch < kChunks; ++ch) {
              uint32_t v[32];
              ptx::tmem_ld_x32(t_addr + box * kAccPerBox + ch * 32, v);
              ptx::tc_wait_ld();
              float acc[32];
#pragma unroll
              for (int j = 0; j < 32; ++j) acc[j] = __uint_as_float(v[j]);
              if (row_ok) epi_compute<EPI, kFast>(p.ep, out_row, acc_col0 + ch * 32, acc, pass);
              constexpr int kPieces = (EPI == EPI_GLU) ? 2 : 4;  // 16-byte pieces produced by this chunk
#pragma unroll
              for (int j = 0; j < kPieces; ++j)
                ptx::sts128(srow + ((static_cast<uint32_t>(ch * kPieces + j) ^ swz) << 4),
                            ptx::pack_bf16x2(acc[8 * j + 0], acc[8 * j + 1]), ptx::pack_bf16x2(acc[8 * j + 2], acc[8 * j + 3]),
                            ptx::pack_bf16x2(acc[8 * j + 4], acc[8 * j + 5]), ptx::pack_bf16x2(acc[8 * j + 6], acc[8 * j + 7]));
            }
            ptx::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              if constexpr (EPI == EPI_QKV) {
                const int oc = (pass == 1 || acc_col0 >= p.ep.qkv_dp) ? acc_col0 + p.ep.qkv_dp : acc_col0;
                tma_store_2d_a(&tmO, sbuf, oc, row0);
              } else if constexpr (EPI == EPI_GLU) {
                tma_store_2d_a(&tmO, sbuf, acc_col0 >> 1, row0);
              } else {
                tma_store_2d_a(&tmO, sbuf, acc_col0, row0);
              }
              ptx::bulk_commit();
            }
          }
        }
        ptx::tc_fence_before();
        ptx::mbar_arrive_a(acc_empty + 8 * buf);
      }
    }
    if (lane == 0) ptx::bulk_wait<0>();  // all output writes complete before the CTA retires
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

template <int EPI>
int launch_instance(const CUtensorMap& tmW, const CUtensorMap& tmO, const LnaParams& p, int grid, int smem,
                    cudaStream_t st, std::string* err) {
  auto kern = gemm_lna_kernel<EPI>;
  static bool configured[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!configured[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) {
      if (err) *err = std::string("cudaFuncSetAttribute(gemm_lna): ") + cudaGetErrorString(e);
      return static_cast<int>(e);
    }
    configured[dev & 63] = true;
  }
  cudaError_t e = launch_pdl(kern, dim3(grid), dim3(kThreads), smem, st, tmW, tmO, p);
  if (e != cudaSuccess) {
    if (err) *err = std::string("gemm_lna launch: ") + cudaGetErrorString(e);
    return static_cast<int>(e);
  }
  return 0;
}

}  // namespace

int launch_gemm_lna(const GemmLnaDesc& g, cudaStream_t st, std::string* err) {
  if (g.M <= 0 || g.N <= 0) return 0;
  if (g.d % 8 != 0 || g.d > 512 || g.d < 16 || (g.ldw % 8) || (g.ldx % 4)) {
    if (err) *err = "gemm_lna: d must be a multiple of 8 and <= 512";
    return -1;
  }
  if (g.epi != EPI_SWISH && g.epi != EPI_QKV && g.epi != EPI_GLU && g.epi != EPI_LINEAR) {
    if (err) *err = "gemm_lna: unsupported epilogue";
    return -1;
  }
  LnaParams p{};
  p.d = g.d;
  p.num_m_blocks = (g.M + kBlockM - 1) / kBlockM;
  p.num_n_tiles = (g.N + kBN - 1) / kBN;
  p.num_k_blocks = (g.d + kBlockK - 1) / kBlockK;
  p.off_b = p.num_k_blocks * kKbBytes;
  constexpr int kBarBytes = 256;
  int stages = (227 * 1024 - p.off_b - kStagingBytes - kBarBytes) / kBBytes;
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) {
    if (err) *err = "gemm_lna: shared memory budget";
    return -1;
  }
  p.stages = stages;
  p.off_staging = p.off_b + stages * kBBytes;
  p.off_bar = p.off_staging + kStagingBytes;
  p.smem_needed = p.off_bar + kBarBytes;
  const int smem_total = p.smem_needed + 1024 <= 227 * 1024 ? p.smem_needed + 1024 : 227 * 1024;
  p.x = g.x;
  p.ldx = g.ldx;
  p.gamma1 = g.gamma1;
  p.beta1 = g.beta1;
  p.x_out = g.x_out;
  p.gamma2 = g.gamma2;
  p.beta2 = g.beta2;
  p.ep = g.ep;
  p.ep.M = g.M;
  p.ep.N = g.N;

  CUtensorMap tmW, tmO;
  {
    uint64_t dims[2] = {static_cast<uint64_t>(g.d), static_cast<uint64_t>(g.N)};
    uint64_t strides[1] = {static_cast<uint64_t>(g.ldw) * 2};
    uint32_t box[2] = {kBlockK, kBN};
    if (!encode_tmap_bf16(&tmW, g.W, 2, dims, strides, box, err)) return -1;
  }
  {
    uint64_t cols = static_cast<uint64_t>(g.N);
    if (g.epi == EPI_QKV) cols = static_cast<uint64_t>(g.N) + g.ep.qkv_dp;  // [q+u | q+v | k | v]
    if (g.epi == EPI_GLU) cols = static_cast<uint64_t>(g.N) / 2;
    if ((g.ep.ldo * 2) % 16) {
      if (err) *err = "gemm_lna: output leading dimension must be a multiple of 16 bytes";
      return -1;
    }
    uint64_t dims[2] = {cols, static_cast<uint64_t>(g.M)};
    uint64_t strides[1] = {static_cast<uint64_t>(g.ep.ldo) * 2};
    uint32_t box[2] = {64u, 32u};
    if (!encode_tmap_bf16(&tmO, g.ep.out, 2, dims, strides, box, err)) return -1;
  }
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (sms <= 0) sms = 148;
  const int grid = p.num_m_blocks < sms ? p.num_m_blocks : sms;
  switch (g.epi) {
    case EPI_SWISH:
      return launch_instance<EPI_SWISH>(tmW, tmO, p, grid, smem_total, st, err);
    case EPI_QKV:
      return launch_instance<EPI_QKV>(tmW, tmO, p, grid, smem_total, st, err);
    case EPI_GLU:
      return launch_instance<EPI_GLU>(tmW, tmO, p, grid, smem_total, st, err);
    default:
      return launch_instance<EPI_LINEAR>(tmW, tmO, p, grid, smem_total, st, err);
  }
}

}  // namespace cfb
