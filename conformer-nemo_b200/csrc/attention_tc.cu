// Fused relative-position attention for sm_100a (flash-style; no T x T matrix ever reaches HBM).
//
//   ctx[i,:] = softmax_j( ((q_i+u).k_j + (q_i+v).p_{T-1+j-i}) / sqrt(dk) ) v_j        (multi_head_attention.py:195-210)
//
// One CTA per (128-query tile, head, sequence); 192 threads:
//   warp 0      TMA producer: Q+u, Q+v once; per 64-key tile K, V and the 192-row band of linear_pos(pos_emb) that
//               the tile can touch (rows T-1+j0-i0-127 .. +191), 2-stage mbarrier ring
//   warp 1      MMA issuer (tcgen05, cta_group::1): S = (Q+u) K^T (128x64) and G = (Q+v) Pband^T (128x192) into one
//               of two TMEM buffers, later O_part = P V (128x64) into the S columns of the same buffer
//   warps 2..5  softmax, one query row per thread: rel_shift is an index remap -- row ii needs G[ii][127-ii+jj] --
//               done as a warp-uniform TMEM column offset plus a 5-stage register barrel shift by (31 - lane);
//               online softmax in fp32 (exp2), probabilities written as bf16 into a 128-byte-swizzled smem tile for
//               the PV MMA, running output kept in registers.
// Keys j >= len[b] are masked to -inf (the reference's -10000 underflows to exactly 0 for valid rows); query rows
// i >= len[b] are written as zeros (multi_head_attention.py:104-113).
#include <math.h>
#include <stdio.h>

#include "common.cuh"
#include "ptx.cuh"

namespace cfb {
namespace {

constexpr int kBM = 128;   // queries per CTA
constexpr int kBN = 64;    // keys per tile
constexpr int kDK = 64;    // padded head dim
constexpr int kBand = 192; // >= kBM + kBN - 1, multiple of 64
constexpr int kThreads = 192;
constexpr int kQBytes = kBM * kDK * 2;     // 16 KB each for Q+u, Q+v
constexpr int kKBytes = kBN * kDK * 2;     // 8 KB
constexpr int kBandBytes = kBand * kDK * 2;  // 24 KB
constexpr int kStageBytes = 2 * kKBytes + kBandBytes;  // K, V, band = 40 KB
constexpr int kPBytes = kBM * kBN * 2;     // 16 KB probabilities
constexpr int kOffStage = 2 * kQBytes;
constexpr int kOffP = kOffStage + 2 * kStageBytes;
constexpr int kOffBar = kOffP + kPBytes;
constexpr int kSmemTotal = kOffBar + 256 + 1024;
constexpr uint32_t kTmemCols = 512;  // two buffers of [S 64 | G 192]

struct AttnParams {
  const int32_t* lens;
  bf16* ctx;
  int T, Dp;
  int pos_col0;  // first column of this layer inside the positional projection buffer
  float scale_log2;  // log2(e) / sqrt(dk)
};

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(kThreads, 1)
rel_attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                   const __grid_constant__ CUtensorMap tmP, const AttnParams p) {
  const int i0 = blockIdx.x * kBM;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int T = p.T;
  const int len = min(p.lens[b], T);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (i0 >= len) {
    // the whole query tile is padding: context is zero (block-uniform exit, nothing allocated yet)
    if (warp >= 2) {
      const int i = i0 + (warp & 3) * 32 + lane;
      if (i < T) {
        uint4* o = reinterpret_cast<uint4*>(p.ctx + (static_cast<long long>(b) * T + i) * p.Dp + h * kDK);
#pragma unroll
        for (int c = 0; c < 8; ++c) o[c] = make_uint4(0, 0, 0, 0);
      }
    }
    return;
  }

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQu = smem;
  uint8_t* sQv = smem + kQBytes;
  uint8_t* sP = smem + kOffP;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint64_t* q_full = bars + 0;
  uint64_t* kv_full = bars + 1;   // [2]
  uint64_t* kv_empty = bars + 3;  // [2]
  uint64_t* sg_full = bars + 5;   // [2]
  uint64_t* sg_free = bars + 7;   // [2]
  uint64_t* p_ready = bars + 9;
  uint64_t* o_full = bars + 10;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int n_kt = (len + kBN - 1) / kBN;

  if (warp == 0) {
    if (lane == 0) {
      ptx::prefetch_tmap(&tmQ);
      ptx::prefetch_tmap(&tmKV);
      ptx::prefetch_tmap(&tmP);
      ptx::mbar_init(q_full, 1);
      for (int s = 0; s < 2; ++s) {
        ptx::mbar_init(&kv_full[s], 1);
        ptx::mbar_init(&kv_empty[s], 1);
        ptx::mbar_init(&sg_full[s], 1);
        ptx::mbar_init(&sg_free[s], 128);
      }
      ptx::mbar_init(p_ready, 128);
      ptx::mbar_init(o_full, 1);
      ptx::fence_mbar_init();
    }
    __syncwarp();
    ptx::tmem_alloc(tmem_slot, kTmemCols);
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ---------------------------------------------------------------------------------- TMA producer
    if (lane == 0) {
      const int row_q = b * T + i0;
      ptx::mbar_arrive_expect_tx(q_full, 2 * kQBytes);
      ptx::tma_load_2d(sQu, &tmQ, q_full, h * kDK, row_q);
      ptx::tma_load_2d(sQv, &tmQ, q_full, p.Dp + h * kDK, row_q);
      for (int kt = 0; kt < n_kt; ++kt) {
        const int s = kt & 1;
        ptx::mbar_wait(&kv_empty[s], ((kt >> 1) & 1) ^ 1);
        ptx::mbar_arrive_expect_tx(&kv_full[s], kStageBytes);
        uint8_t* st = smem + kOffStage + s * kStageBytes;
        const int j0 = kt * kBN;
        ptx::tma_load_2d(st, &tmKV, &kv_full[s], 2 * p.Dp + h * kDK, b * T + j0);
        ptx::tma_load_2d(st + kKBytes, &tmKV, &kv_full[s], 3 * p.Dp + h * kDK, b * T + j0);
        const int r_lo = T - 1 + j0 - i0 - (kBM - 1);  // band row of G column 0 (may be < 0: TMA zero-fills)
#pragma unroll
        for (int q = 0; q < kBand / 64; ++q)
          ptx::tma_load_2d(st + 2 * kKBytes + q * (64 * kDK * 2), &tmP, &kv_full[s], p.pos_col0 + h * kDK,
                           r_lo + q * 64);
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_s = ptx::make_idesc_bf16(kBM, kBN, 0, 0);
      constexpr uint32_t idesc_g = ptx::make_idesc_bf16(kBM, kBand, 0, 0);
      constexpr uint32_t idesc_o = ptx::make_idesc_bf16(kBM, kDK, 0, 1);  // B = V is MN-major (keys x dk rows)
      const uint64_t dQu = ptx::make_sdesc_sw128(ptx::smem_u32(sQu), 16, 1024);
      const uint64_t dQv = ptx::make_sdesc_sw128(ptx::smem_u32(sQv), 16, 1024);
      const uint64_t dP = ptx::make_sdesc_sw128(ptx::smem_u32(sP), 16, 1024);
      ptx::mbar_wait(q_full, 0);

      auto issue_sg = [&](int kt) {
        const int s = kt & 1;
        ptx::mbar_wait(&kv_full[s], (kt >> 1) & 1);
        ptx::mbar_wait(&sg_free[s], ((kt >> 1) & 1) ^ 1);
        ptx::tc_fence_after();
        const uint32_t st = ptx::smem_u32(smem + kOffStage + s * kStageBytes);
        const uint64_t dK = ptx::make_sdesc_sw128(st, 16, 1024);
        const uint64_t dB = ptx::make_sdesc_sw128(st + 2 * kKBytes, 16, 1024);
        const uint32_t tS = tmem_base + s * 256;
#pragma unroll
        for (int k = 0; k < kDK / 16; ++k) ptx::umma_bf16(tS, dQu + 2 * k, dK + 2 * k, idesc_s, k != 0);
#pragma unroll
        for (int k = 0; k < kDK / 16; ++k) ptx::umma_bf16(tS + kBN, dQv + 2 * k, dB + 2 * k, idesc_g, k != 0);
        ptx::tc_commit(&sg_full[s]);
      };

      issue_sg(0);
      for (int kt = 0; kt < n_kt; ++kt) {
        const int s = kt & 1;
        if (kt + 1 < n_kt) issue_sg(kt + 1);
        ptx::mbar_wait(p_ready, kt & 1);
        ptx::tc_fence_after();
        const uint32_t st = ptx::smem_u32(smem + kOffStage + s * kStageBytes);
        // V tile: 64 keys (K of this MMA) x 64 dk (N), 128-byte rows along N -> MN-major, 8-key groups 1024 B apart
        const uint64_t dV = ptx::make_sdesc_sw128(st + kKBytes, 1024, 1024);
#pragma unroll
        for (int k = 0; k < kBN / 16; ++k)
          ptx::umma_bf16(tmem_base + s * 256, dP + 2 * k, dV + static_cast<uint64_t>(k) * (2048 >> 4), idesc_o, k != 0);
        ptx::tc_commit(o_full);
        ptx::tc_commit(&kv_empty[s]);
      }
    }
  } else {
    // ---------------------------------------------------------------------------------- softmax warps
    const int quarter = warp & 3;
    const int ii = quarter * 32 + lane;  // query row inside the tile == TMEM lane
    const int i = i0 + ii;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const int g_base = kBN + (96 - 32 * quarter);  // warp-uniform part of the rel_shift column offset
    const int sh = 31 - lane;                      // per-lane part, applied by the barrel shifter
    float o_acc[kDK];
#pragma unroll
    for (int c = 0; c < kDK; ++c) o_acc[c] = 0.f;
    float m_run = -INFINITY, l_run = 0.f, alpha_pending = 1.f;

    for (int kt = 0; kt < n_kt; ++kt) {
      const int s = kt & 1;
      const int j0 = kt * kBN;
      ptx::mbar_wait(&sg_full[s], (kt >> 1) & 1);
      ptx::tc_fence_after();
      const uint32_t tS = lane_addr + s * 256;
      float sv[kBN];
#pragma unroll
      for (int qd = 0; qd < kBN / 16; ++qd) {
        uint32_t a[16], w[48];
        ptx::tmem_ld_x16(tS + qd * 16, a);
        ptx::tmem_ld_x16(tS + g_base + qd * 16, w);
        ptx::tmem_ld_x16(tS + g_base + qd * 16 + 16, w + 16);
        ptx::tmem_ld_x16(tS + g_base + qd * 16 + 32, w + 32);
        ptx::tc_wait_ld();
        // out[jj] = w[jj + sh], sh in [0,31]
        if (sh & 16) {
#pragma unroll
          for (int c = 0; c < 31; ++c) w[c] = w[c + 16];
        }
        if (sh & 8) {
#pragma unroll
          for (int c = 0; c < 23; ++c) w[c] = w[c + 8];
        }
        if (sh & 4) {
#pragma unroll
          for (int c = 0; c < 19; ++c) w[c] = w[c + 4];
        }
        if (sh & 2) {
#pragma unroll
          for (int c = 0; c < 17; ++c) w[c] = w[c + 2];
        }
        if (sh & 1) {
#pragma unroll
          for (int c = 0; c < 16; ++c) w[c] = w[c + 1];
        }
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          const int j = j0 + qd * 16 + c;
          const float x = (__uint_as_float(a[c]) + __uint_as_float(w[c])) * p.scale_log2;
          sv[qd * 16 + c] = (j < len) ? x : -INFINITY;
        }
      }
      float mx = sv[0];
#pragma unroll
      for (int c = 1; c < kBN; ++c) mx = fmaxf(mx, sv[c]);
      const float m_new = fmaxf(m_run, mx);  // finite: every tile holds at least one key j < len
      const float alpha = fast_exp2(m_run - m_new);
      float rsum = 0.f;
#pragma unroll
      for (int c = 0; c < kBN; ++c) {
        sv[c] = fast_exp2(sv[c] - m_new);
        rsum += sv[c];
      }
      l_run = l_run * alpha + rsum;
      m_run = m_new;

      // the probability tile is single-buffered: PV of the previous key tile must have drained it
      if (kt > 0) {
        ptx::mbar_wait(o_full, (kt - 1) & 1);
        ptx::tc_fence_after();
      }
      {
        uint8_t* prow = sP + ii * 128;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          uint4 u;
          u.x = ptx::pack_bf16x2(sv[8 * c + 0], sv[8 * c + 1]);
          u.y = ptx::pack_bf16x2(sv[8 * c + 2], sv[8 * c + 3]);
          u.z = ptx::pack_bf16x2(sv[8 * c + 4], sv[8 * c + 5]);
          u.w = ptx::pack_bf16x2(sv[8 * c + 6], sv[8 * c + 7]);
          *reinterpret_cast<uint4*>(prow + ((c ^ (ii & 7)) << 4)) = u;  // 128-byte swizzle, K-major
        }
      }
      ptx::fence_proxy_async_smem();
      ptx::tc_fence_before();  // orders this thread's TMEM reads of S/G before the MMA that overwrites S
      ptx::mbar_arrive(p_ready);

      if (kt > 0) {
        // fold in O_part of the previous tile (it sits in the S columns of the other TMEM buffer)
        const uint32_t tO = lane_addr + ((kt - 1) & 1) * 256;
#pragma unroll
        for (int qd = 0; qd < kDK / 16; ++qd) {
          uint32_t a[16];
          ptx::tmem_ld_x16(tO + qd * 16, a);
          ptx::tc_wait_ld();
#pragma unroll
          for (int c = 0; c < 16; ++c) o_acc[qd * 16 + c] = o_acc[qd * 16 + c] * alpha_pending + __uint_as_float(a[c]);
        }
        ptx::tc_fence_before();
        ptx::mbar_arrive(&sg_free[(kt - 1) & 1]);
      }
      alpha_pending = alpha;
    }
    // drain the last tile
    {
      const int kt = n_kt - 1;
      ptx::mbar_wait(o_full, kt & 1);
      ptx::tc_fence_after();
      const uint32_t tO = lane_addr + (kt & 1) * 256;
      const float inv = (i < len && l_run > 0.f) ? 1.f / l_run : 0.f;  // padded query rows -> zeros
#pragma unroll
      for (int qd = 0; qd < kDK / 16; ++qd) {
        uint32_t a[16];
        ptx::tmem_ld_x16(tO + qd * 16, a);
        ptx::tc_wait_ld();
#pragma unroll
        for (int c = 0; c < 16; ++c)
          o_acc[qd * 16 + c] = (o_acc[qd * 16 + c] * alpha_pending + __uint_as_float(a[c])) * inv;
      }
      if (i < T) {
        uint4* o = reinterpret_cast<uint4*>(p.ctx + (static_cast<long long>(b) * T + i) * p.Dp + h * kDK);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          uint4 u;
          u.x = ptx::pack_bf16x2(o_acc[8 * c + 0], o_acc[8 * c + 1]);
          u.y = ptx::pack_bf16x2(o_acc[8 * c + 2], o_acc[8 * c + 3]);
          u.z = ptx::pack_bf16x2(o_acc[8 * c + 4], o_acc[8 * c + 5]);
          u.w = ptx::pack_bf16x2(o_acc[8 * c + 6], o_acc[8 * c + 7]);
          o[c] = u;
        }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace

int launch_attn_tc(const AttnDesc& a, cudaStream_t st, std::string* err) {
  if (a.B <= 0 || a.T <= 0) return 0;
  if (a.dkp != kDK) {
    if (err) *err = "attn_tc: padded head dim must be 64";
    return -1;
  }
  const int Dp = a.H * a.dkp;
  const long long rows = static_cast<long long>(a.B) * a.T;
  CUtensorMap tmQ, tmKV, tmP;
  {
    uint64_t dims[2] = {static_cast<uint64_t>(4 * Dp), static_cast<uint64_t>(rows)};
    uint64_t strides[1] = {static_cast<uint64_t>(4 * Dp) * 2};
    uint32_t boxq[2] = {kDK, kBM};
    uint32_t boxk[2] = {kDK, kBN};
    if (!encode_tmap_bf16(&tmQ, a.qkv, 2, dims, strides, boxq, err)) return -1;
    if (!encode_tmap_bf16(&tmKV, a.qkv, 2, dims, strides, boxk, err)) return -1;
  }
  {
    // a.pos points at this layer's first column inside the (2T-1, ld_pos) positional projection buffer
    uint64_t dims[2] = {static_cast<uint64_t>(Dp), static_cast<uint64_t>(2 * a.T - 1)};
    uint64_t strides[1] = {static_cast<uint64_t>(a.ld_pos) * 2};
    uint32_t box[2] = {kDK, 64};
    if (!encode_tmap_bf16(&tmP, a.pos, 2, dims, strides, box, err)) return -1;
  }
  static bool configured[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!configured[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(rel_attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal);
    if (e != cudaSuccess) {
      if (err) *err = std::string("cudaFuncSetAttribute(attn_tc): ") + cudaGetErrorString(e);
      return static_cast<int>(e);
    }
    configured[dev & 63] = true;
  }
  AttnParams p;
  p.lens = a.lens;
  p.ctx = reinterpret_cast<bf16*>(a.ctx);
  p.T = a.T;
  p.Dp = Dp;
  p.pos_col0 = 0;
  p.scale_log2 = 1.4426950408889634f / sqrtf(static_cast<float>(a.dk));
  dim3 grid((a.T + kBM - 1) / kBM, a.H, a.B);
  rel_attn_tc_kernel<<<grid, kThreads, kSmemTotal, st>>>(tmQ, tmKV, tmP, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    if (err) *err = std::string("attn_tc launch: ") + cudaGetErrorString(e);
    return static_cast<int>(e);
  }
  return 0;
}

}  // namespace cfb
