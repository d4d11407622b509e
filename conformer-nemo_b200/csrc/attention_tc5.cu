// Fused relative-position attention, TWO THREADS PER QUERY ROW (round 2).  Same math, tensor-memory map, TMA producer and
// MMA issuer roles as attention_tc.cu (read that header first); what changes is the softmax side:
//
//   attention_tc.cu   8 softmax warps: one thread owns a query row and all 64 keys of its set's key tile.  ncu on it:
//                     issue slots 36 % busy, 3 warps per scheduler, 0.47 eligible, no hot instruction -- the softmax
//                     warps are latency-bound chains (~700 instructions per thread and tile, ~2 100 cycles per tile
//                     against ~700 at full issue rate).
//   here              16 softmax warps: warps w and w + 4 of a set share a TMEM lane quarter (a warp may only touch
//                     lanes 32 (w % 4) ..) and split every row's tile in halves of 32 keys: each thread reads 32 S
//                     columns and a 64-column G window, exchanges its row maximum with its partner through shared
//                     memory (one 64-thread named barrier per tile), writes 16 of the 32 P columns and folds 32 of the
//                     64 O columns.  Half the serial work per thread and four softmax warps per scheduler instead of two.
//
// Registers: 640 threads are launched with 96 each and setmaxnreg only redistributes what the CTA was given at launch
// (USETMAXREG.TRY_ALLOC.CTAPOOL: asking for more than 640 * 96 in total never completes): the four producer / issuer
// warps drop to 64, the softmax warps rise to 104 (4 * 32 * 64 + 16 * 32 * 104 = 61 440 = 640 * 96).  Shared memory: K/V ring 3 x 16 KB, band slots 3 x 8 KB, 16
// private shift areas of 32 rows x 68 words (64-column window; 16-byte stores and 4-byte loads at word offset 31 - lane
// are conflict-free at this pitch), the per-tile maximum exchange and the end-of-kernel set exchange (in the K/V ring).
#include <cuda_fp16.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"

namespace cfb {
namespace {

constexpr int kBM = 128;   // queries per CTA
constexpr int kBN = 64;    // keys per tile
constexpr int kHN = 32;    // keys per thread and tile (half a tile)
constexpr int kDK = 64;    // padded head dim
constexpr int kIssuerWarps = 4;
constexpr int kSoftmaxWarps = 16;
constexpr int kThreads = 32 * (kIssuerWarps + kSoftmaxWarps);
constexpr int kKBytes = kBN * kDK * 2;     // 8 KB
constexpr int kKVStages = 3;               // K + V tiles, 16 KB per stage
constexpr int kKVBytes = 2 * kKBytes;
constexpr int kBandSlots = 3;
constexpr int kBlockBytes = 64 * kDK * 2;
constexpr int kGSlots = 4;
constexpr int kShiftPitch = 68;            // words per private shift row (64-column fp32 window + pad)
constexpr int kShiftBytes = 32 * kShiftPitch * 4;
constexpr int kXPitch = 68;                // words per row of the end-of-kernel exchange (m1 + pad, O1[64])
constexpr int kOffKV = 0;
constexpr int kOffBand = kOffKV + kKVStages * kKVBytes;
constexpr int kOffShift = kOffBand + kBandSlots * kBlockBytes;
constexpr int kOffMax = kOffShift + kSoftmaxWarps * kShiftBytes;  // float [2 sets][2 buffers][2 halves][128 rows]
constexpr int kMaxBytes = 2 * 2 * 2 * kBM * 4;
constexpr int kOffSum = kOffMax + kMaxBytes;                      // float [2 sets][2 halves][128 rows]: partial row sums
constexpr int kSumBytes = 2 * 2 * kBM * 4;
constexpr int kOffBar = kOffSum + kSumBytes;
constexpr int kBarBytes = 512;
constexpr int kSmemTotal = kOffBar + kBarBytes + 1024;
static_assert(kBM * kXPitch * 4 <= kKVStages * kKVBytes, "set exchange must fit in the K/V ring");
static_assert(kSmemTotal <= 227 * 1024, "shared memory budget");
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kColQ = 0, kColS = 64, kColP = 192, kColG = 256;

struct Attn5Params {
  const bf16* qkv;
  const int32_t* lens;
  bf16* ctx;
  int T, Dp;
  const int4* tiles;  // packed batches: blockIdx.x -> (sequence, i0, first token row of its slot, rows in the slot)
  float scale_log2;   // log2(e) / sqrt(dk)
  int flags;          // experiments (CFB_ATTN5_FLAGS): 1 = no MUFU turn-taking between the sets, 2 = one wait per window
};

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
  return d;
}
__device__ __forceinline__ float4 lds_f32x4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
// registers -> TMEM, 16 consecutive 32-bit columns of this thread's lane
__device__ __forceinline__ void tmem_st_x16(uint32_t taddr, const uint32_t* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void pair_barrier(int id) {  // the two warps that share a row group (64 threads)
  asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory");
}

__global__ void __launch_bounds__(kThreads, 1)
rel_attn_tc5_kernel(const __grid_constant__ CUtensorMap tmKV, const __grid_constant__ CUtensorMap tmP,
                    const Attn5Params p) {
  const int h = blockIdx.y;
  const int T = p.T;  // extent of the positional table (2T - 1 band rows)
  uint32_t tid;
  asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid));
  const int warp = tid >> 5;
  const int lane = tid & 31;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar0 = sbase + kOffBar;
  const uint32_t qu_ready = bar0 + 0;    // Q+u is in TMEM (8 warp arrivals, softmax set 0)
  const uint32_t qv_ready = bar0 + 8;    // Q+v is in TMEM (8 warp arrivals, softmax set 1)
  const uint32_t sg_full = bar0 + 16;    // [2] per set: S of the set's current key tile is in TMEM
  const uint32_t s_free = bar0 + 32;     // [2] O_part folded: the S columns may be overwritten
  const uint32_t g_free = bar0 + 48;     // [2] per set: the G window of the set's tile has been read out of the ring
  const uint32_t p_ready = bar0 + 64;    // [2]
  const uint32_t o_full = bar0 + 80;     // [2]
  const uint32_t exp_done = bar0 + 96;   // [2] per set: the exponentials of the set's tile have issued
  const uint32_t kv_full = bar0 + 112;                   // [kKVStages]
  const uint32_t kv_empty = kv_full + 8 * kKVStages;     // [kKVStages]
  const uint32_t band_full = kv_empty + 8 * kKVStages;   // [kBandSlots]
  const uint32_t band_empty = band_full + 8 * kBandSlots;  // [kBandSlots]
  const uint32_t g_full = band_empty + 8 * kBandSlots;   // [kGSlots] G block in TMEM
  const uint32_t tmem_slot = g_full + 8 * kGSlots;
  static_assert(112 + 8 * (2 * kKVStages + 2 * kBandSlots + kGSlots) + 8 <= kBarBytes, "barrier area");
  constexpr int kSetWarps = kSoftmaxWarps / 2;  // warp arrivals per set

  if (warp == 0) {
    if (lane == 0) {
      ptx::prefetch_tmap(&tmKV);
      ptx::prefetch_tmap(&tmP);
      ptx::mbar_init_a(qu_ready, kSetWarps);
      ptx::mbar_init_a(qv_ready, kSetWarps);
      for (int s = 0; s < kKVStages; ++s) {
        ptx::mbar_init_a(kv_full + 8 * s, 1);
        ptx::mbar_init_a(kv_empty + 8 * s, 1);
      }
      for (int s = 0; s < kBandSlots; ++s) {
        ptx::mbar_init_a(band_full + 8 * s, 1);
        ptx::mbar_init_a(band_empty + 8 * s, 1);
      }
      for (int s = 0; s < kGSlots; ++s) ptx::mbar_init_a(g_full + 8 * s, 1);
      for (int s = 0; s < 2; ++s) {
        ptx::mbar_init_a(sg_full + 8 * s, 1);
        ptx::mbar_init_a(s_free + 8 * s, kSetWarps);   // one elected arrival per softmax warp
        ptx::mbar_init_a(g_free + 8 * s, kSetWarps);
        ptx::mbar_init_a(p_ready + 8 * s, kSetWarps);
        ptx::mbar_init_a(o_full + 8 * s, 1);
        ptx::mbar_init_a(exp_done + 8 * s, kSetWarps);
      }
      ptx::fence_mbar_init();
    }
    __syncwarp();
  } else if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  pdl_launch_dependents();
  pdl_wait();
  // dense layout: sequence b owns rows [b T, b T + T); packed: the slot [rb, rb + S) of the tile table
  int i0 = blockIdx.x * kBM, b = blockIdx.z, S = T;
  long long rb = static_cast<long long>(b) * T;
  if (p.tiles != nullptr) {
    const int4 t = __ldg(p.tiles + blockIdx.x);
    b = t.x, i0 = t.y, rb = t.z, S = t.w;
  }
  const int len = min(p.lens[b], S);
  const bool active = i0 < len;  // otherwise the whole query tile is padding: the context rows are zero
  const int n_kt = active ? (len + kBN - 1) / kBN : 0;
  const int n_gb = n_kt + 2;  // G blocks 0 .. n_kt+1

  // softmax thread -> (set, key half, lane quarter)
  const int sw = warp - kIssuerWarps;
  const int quarter = warp & 3;
  const int half = (sw >> 2) & 1;
  const int set = sw >> 3;

  uint32_t qw[16];  // softmax warps: this thread's half row of Q+u (set 0) / Q+v (set 1), 32 bf16, fetched before the setup
  if (warp >= kIssuerWarps && active) {
    const int i = i0 + quarter * 32 + lane;
    if (i < S) {
      const uint4* src = reinterpret_cast<const uint4*>(p.qkv + (rb + i) * (4 * p.Dp) + set * p.Dp + h * kDK + half * 32);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const uint4 u = __ldg(src + c);
        qw[4 * c] = u.x, qw[4 * c + 1] = u.y, qw[4 * c + 2] = u.z, qw[4 * c + 3] = u.w;
      }
    } else {
#pragma unroll
      for (int c = 0; c < 16; ++c) qw[c] = 0u;
    }
  }

  if (warp == 0 && lane == 0 && active) {
    // the first loads only need their own barriers: issue them before the CTA-wide setup barrier
    const int r0 = T - 1 - i0 - (kBM - 1);  // band row of G column 0 of block 0 (may be < 0: TMA zero-fills)
#pragma unroll
    for (int g = 0; g < 3; ++g) {
      ptx::mbar_arrive_expect_tx_a(band_full + 8 * g, kBlockBytes);
      ptx::tma_load_2d_a(sbase + kOffBand + g * kBlockBytes, &tmP, band_full + 8 * g, h * kDK, r0 + 64 * g);
    }
    ptx::mbar_arrive_expect_tx_a(kv_full, kKVBytes);
    ptx::tma_load_2d_a(sbase + kOffKV, &tmKV, kv_full, 2 * p.Dp + h * kDK, static_cast<int>(rb));
    ptx::tma_load_2d_a(sbase + kOffKV + kKBytes, &tmKV, kv_full, 3 * p.Dp + h * kDK, static_cast<int>(rb));
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");

  if (!active) {
    if (warp >= kIssuerWarps && set == 0) {
      const int i = i0 + quarter * 32 + lane;
      if (i < S) {
        uint4* o = reinterpret_cast<uint4*>(p.ctx + (rb + i) * p.Dp + h * kDK + half * 32);
#pragma unroll
        for (int c = 0; c < 4; ++c) o[c] = make_uint4(0, 0, 0, 0);
      }
    }
  } else if (warp < kIssuerWarps) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 64;" ::: "memory");
    if (warp == 0) {
      // ---------------------------------------------------------------------------------- TMA producer
      if (lane == 0) {
        const int r0 = T - 1 - i0 - (kBM - 1);
        auto load_band_block = [&](int g) {
          if (g >= n_gb) return;
          const int slot = g % kBandSlots, use = g / kBandSlots;
          ptx::mbar_wait_a(band_empty + 8 * slot, (use & 1) ^ 1);
          ptx::mbar_arrive_expect_tx_a(band_full + 8 * slot, kBlockBytes);
          ptx::tma_load_2d_a(sbase + kOffBand + slot * kBlockBytes, &tmP, band_full + 8 * slot, h * kDK, r0 + 64 * g);
        };
        load_band_block(3);
        for (int kt = 1; kt < n_kt; ++kt) {
          const int st = kt % kKVStages, use = kt / kKVStages;
          ptx::mbar_wait_a(kv_empty + 8 * st, (use & 1) ^ 1);
          ptx::mbar_arrive_expect_tx_a(kv_full + 8 * st, kKVBytes);
          const uint32_t dst = sbase + kOffKV + st * kKVBytes;
          ptx::tma_load_2d_a(dst, &tmKV, kv_full + 8 * st, 2 * p.Dp + h * kDK, static_cast<int>(rb) + kt * kBN);
          ptx::tma_load_2d_a(dst + kKBytes, &tmKV, kv_full + 8 * st, 3 * p.Dp + h * kDK, static_cast<int>(rb) + kt * kBN);
          load_band_block(kt + 3);
        }
      }
    } else if (warp == 1 || warp == 2) {
      // ---------------------------------------------------------------------------------- S / PV issuer of a set
      const int s = warp - 1;
      constexpr uint32_t idesc_s = ptx::make_idesc_bf16(kBM, kBN, 0, 0);
      constexpr uint32_t idesc_o = ptx::make_idesc_bf16(kBM, kDK, 0, 1);  // B = V is MN-major (keys x dk rows)
      const uint32_t tQu = tmem_base + kColQ;
      const uint32_t tS = tmem_base + kColS + s * 64;
      const uint32_t tP = tmem_base + kColP + s * 32;
      ptx::mbar_wait_a(qu_ready, 0);
      ptx::tc_fence_after();
      int it = 0;
      for (int kt = s; kt < n_kt; kt += 2, ++it) {
        const int kvs = kt % kKVStages;
        const uint32_t st = sbase + kOffKV + kvs * kKVBytes;
        const uint64_t dK = ptx::make_sdesc_sw128(st, 16, 1024);
        const uint64_t dV = ptx::make_sdesc_sw128(st + kKBytes, 1024, 1024);
        ptx::mbar_wait_a(kv_full + 8 * kvs, (kt / kKVStages) & 1);
        ptx::mbar_wait_a(s_free + 8 * s, (it & 1) ^ 1);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
#pragma unroll
          for (int k = 0; k < kDK / 16; ++k) ptx::umma_bf16_ts(tS, tQu + 8 * k, dK + 2 * k, idesc_s, k != 0);
          ptx::tc_commit_a(sg_full + 8 * s);
        }
        __syncwarp();
        // ---- O_part = P V once both halves of the set have stored their probabilities
        ptx::mbar_wait_a(p_ready + 8 * s, it & 1);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
#pragma unroll
          for (int k = 0; k < kBN / 16; ++k)
            ptx::umma_bf16_ts(tS, tP + 8 * k, dV + static_cast<uint64_t>(k) * (2048 >> 4), idesc_o, k != 0);
          ptx::tc_commit_a(o_full + 8 * s);
          ptx::tc_commit_a(kv_empty + 8 * kvs);
        }
        __syncwarp();
      }
    } else {
      // ---------------------------------------------------------------------------------- G ring issuer
      constexpr uint32_t idesc_g = ptx::make_idesc_bf16(kBM, 64, 0, 0);
      constexpr uint32_t idesc_g3 = ptx::make_idesc_bf16(kBM, 192, 0, 0);
      const uint32_t tQv = tmem_base + kColQ + 32;
      const uint32_t band_base = sbase + kOffBand;
      ptx::mbar_wait_a(qv_ready, 0);
      ptx::mbar_wait_a(band_full + 0, 0);
      ptx::mbar_wait_a(band_full + 8, 0);
      ptx::mbar_wait_a(band_full + 16, 0);
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
        const uint64_t dB = ptx::make_sdesc_sw128(band_base, 16, 1024);
#pragma unroll
        for (int k = 0; k < kDK / 16; ++k)
          ptx::umma_bf16_ts(tmem_base + kColG, tQv + 8 * k, dB + 2 * k, idesc_g3, k != 0);
        ptx::tc_commit_a(g_full + 0);
        ptx::tc_commit_a(g_full + 8);
        ptx::tc_commit_a(g_full + 16);
        ptx::tc_commit_a(band_empty + 0);
        ptx::tc_commit_a(band_empty + 8);
        ptx::tc_commit_a(band_empty + 16);
      }
      __syncwarp();
      for (int g = 3; g < n_gb; ++g) {
        const int bs = g % kBandSlots;
        const uint64_t dB = ptx::make_sdesc_sw128(band_base + bs * kBlockBytes, 16, 1024);
        const uint32_t tG = tmem_base + kColG + (g % kGSlots) * 64;
        ptx::mbar_wait_a(band_full + 8 * bs, (g / kBandSlots) & 1);
        // see attention_tc.cu: one g_free wait per step covers the three readers of the ring slot
        if (g >= 4 && g - 4 < n_kt) ptx::mbar_wait_a(g_free + 8 * ((g - 4) & 1), ((g - 4) >> 1) & 1);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
#pragma unroll
          for (int k = 0; k < kDK / 16; ++k) ptx::umma_bf16_ts(tG, tQv + 8 * k, dB + 2 * k, idesc_g, k != 0);
          ptx::tc_commit_a(g_full + 8 * (g % kGSlots));
          ptx::tc_commit_a(band_empty + 8 * bs);
        }
        __syncwarp();
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;" ::: "memory");
    // ---------------------------------------------------------------------------------- softmax warps
    const int ii = quarter * 32 + lane;  // query row inside the tile == TMEM lane
    const int i = i0 + ii;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const uint32_t tS = t_lane + kColS + set * 64 + half * kHN;   // this thread's 32 score / O columns
    const uint32_t tP = t_lane + kColP + set * 32 + half * 16;    // and its 16 columns of the bf16 P operand
    const int pair_id = 2 + set * 4 + quarter;                    // named barrier of the two warps that share the rows

    // ---- this thread's half row of Q+u (set 0) or Q+v (set 1) -> TMEM A operand
    tmem_st_x16(t_lane + kColQ + set * 32 + half * 16, qw);
    ptx::tc_wait_st();
    ptx::tc_fence_before();
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive_a(set == 0 ? qu_ready : qv_ready);

    // rel_shift: row ii needs, for its keys jj = 32 half .. 32 half + 31 of the tile, ring column 127 - ii + jj counted
    // from block kt: window of 64 columns starting at 96 - 32 quarter + 32 half, read back at word offset 31 - lane.
    const int sh = 31 - lane;
    const uint32_t shift_row = sbase + kOffShift + sw * kShiftBytes + lane * kShiftPitch * 4;
    const int wcol = 96 - 32 * quarter + 32 * half;
    const uint32_t xmax = sbase + kOffMax + (set * 4) * (kBM * 4) + ii * 4;  // + (buffer * 2 + half) * kBM * 4

    auto fetch_window = [&](int kt) {
      ptx::mbar_wait_a(g_full + 8 * ((kt + 2) % kGSlots), ((kt + 2) / kGSlots) & 1);  // blocks complete in order
      ptx::tc_fence_after();
      if (p.flags & 2) {
        uint32_t w0[32], w1[32];
        const int wc0 = wcol, wc1 = wcol + 32;
        ptx::tmem_ld_x32(t_lane + kColG + ((kt + (wc0 >> 6)) % kGSlots) * 64 + (wc0 & 63), w0);
        ptx::tmem_ld_x32(t_lane + kColG + ((kt + (wc1 >> 6)) % kGSlots) * 64 + (wc1 & 63), w1);
        ptx::tc_wait_ld();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive_a(g_free + 8 * set);
#pragma unroll
        for (int q = 0; q < 8; ++q) ptx::sts128(shift_row + q * 16, w0[4 * q], w0[4 * q + 1], w0[4 * q + 2], w0[4 * q + 3]);
#pragma unroll
        for (int q = 0; q < 8; ++q) ptx::sts128(shift_row + 128 + q * 16, w1[4 * q], w1[4 * q + 1], w1[4 * q + 2], w1[4 * q + 3]);
        return;
      }
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t w[32];
        const int wc = wcol + 32 * c;
        const int blk = kt + (wc >> 6);
        ptx::tmem_ld_x32(t_lane + kColG + (blk % kGSlots) * 64 + (wc & 63), w);
        ptx::tc_wait_ld();
        if (c == 1) {
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive_a(g_free + 8 * set);
        }
#pragma unroll
        for (int q = 0; q < 8; ++q)
          ptx::sts128(shift_row + c * 128 + q * 16, w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
      }
    };

    float o_acc[kHN];
#pragma unroll
    for (int c = 0; c < kHN; ++c) o_acc[c] = 0.f;
    float m_run = -INFINITY, l_run = 0.f;
    const float scale = p.scale_log2;

    if (set < n_kt) fetch_window(set);
    int it = 0;
    for (int kt = set; kt < n_kt; kt += 2, ++it) {
      const int j0 = kt * kBN + half * kHN;
      ptx::mbar_wait_a(sg_full + 8 * set, it & 1);
      ptx::tc_fence_after();
      float sv[kHN];
      {
        uint32_t sr[32];
        ptx::tmem_ld_x32(tS, sr);
        float g[32];
        ptx::lds_f32x32(shift_row + sh * 4, g);
        ptx::tc_wait_ld();
#pragma unroll
        for (int c = 0; c < kHN; ++c) sv[c] = __uint_as_float(sr[c]) + g[c];
      }
      if (j0 + kHN > len) {  // only the last key tile can contain masked keys
#pragma unroll
        for (int c = 0; c < kHN; ++c)
          if (j0 + c >= len) sv[c] = -INFINITY;
      }
      float mx4[4] = {sv[0], sv[1], sv[2], sv[3]};
#pragma unroll
      for (int c = 4; c < kHN; ++c) mx4[c & 3] = fmaxf(mx4[c & 3], sv[c]);
      float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
      // ---- row maximum over both halves: exchange through shared memory (double-buffered by tile parity)
      {
        const uint32_t mine = xmax + ((it & 1) * 2 + half) * (kBM * 4);
        const uint32_t theirs = xmax + ((it & 1) * 2 + (half ^ 1)) * (kBM * 4);
        ptx::sts_f32(mine, mx);
        pair_barrier(pair_id);
        mx = fmaxf(mx, ptx::lds_f32(theirs));
      }
      // the exponentials of consecutive key tiles take turns on the MUFU pipe (see attention_tc.cu)
      if (kt > 0 && !(p.flags & 1)) ptx::mbar_wait_a(exp_done + 8 * (set ^ 1), (set == 0 ? it - 1 : it) & 1);
      const float m_new = fmaxf(m_run, mx);  // finite: every tile holds at least one key j < len (in half 0)
      const float ms = m_new * scale;
      const float alpha = fast_exp2(fmaf(m_run, scale, -ms));
      float rs4[4] = {0.f, 0.f, 0.f, 0.f};
      uint32_t pw[16];
#pragma unroll
      for (int m = 0; m < 16; ++m) {
        const float e0 = fast_exp2(fmaf(sv[2 * m], scale, -ms)), e1 = fast_exp2(fmaf(sv[2 * m + 1], scale, -ms));
        rs4[m & 3] += e0 + e1;
        // bf16 pair without the conversion unit (it shares the MUFU pipe): round half up with integer adds, one PRMT
        pw[m] = prmt(__float_as_uint(e0) + 0x8000u, __float_as_uint(e1) + 0x8000u, 0x7632u);
      }
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_a(exp_done + 8 * set);
      tmem_st_x16(tP, pw);
      const float rsum = (rs4[0] + rs4[1]) + (rs4[2] + rs4[3]);
      l_run = fmaf(l_run, alpha, rsum);  // partial sum over this thread's keys; the halves are added at the end
      m_run = m_new;
      ptx::tc_wait_st();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_a(p_ready + 8 * set);
      // ---- the G window of this set's next tile, while the P V MMA runs
      if (kt + 2 < n_kt) fetch_window(kt + 2);
      // ---- o_acc = o_acc * alpha + O_part (this thread's 32 of the 64 columns)
      ptx::mbar_wait_a(o_full + 8 * set, it & 1);
      ptx::tc_fence_after();
      {
        uint32_t a0[32];
        ptx::tmem_ld_x32(tS, a0);
        ptx::tc_wait_ld();
#pragma unroll
        for (int c = 0; c < kHN; ++c) o_acc[c] = fmaf(o_acc[c], alpha, __uint_as_float(a0[c]));
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_a(s_free + 8 * set);  // the S columns may now receive the next tile's scores
    }

    // ---- merge: the halves' partial sums, then the two sets (log-sum-exp); write the context rows.  Every MMA and
    // every TMA load of the CTA has completed once both sets are past their last o_full wait, so the K/V ring can
    // carry the exchange.
    const uint32_t xsum = sbase + kOffSum + ii * 4;  // + (set * 2 + half) * kBM * 4
    ptx::sts_f32(xsum + (set * 2 + half) * (kBM * 4), l_run);
    asm volatile("bar.sync 1, 512;" ::: "memory");
    const uint32_t xrow = sbase + kOffKV + ii * kXPitch * 4;
    if (set == 1) {
      if (half == 0) ptx::sts_f32(xrow, m_run);
#pragma unroll
      for (int c = 0; c < kHN / 4; ++c)
        ptx::sts128(xrow + 16 + half * 128 + 16 * c, __float_as_uint(o_acc[4 * c]), __float_as_uint(o_acc[4 * c + 1]),
                    __float_as_uint(o_acc[4 * c + 2]), __float_as_uint(o_acc[4 * c + 3]));
    }
    asm volatile("bar.sync 1, 512;" ::: "memory");
    if (set == 0) {
      const float l0 = ptx::lds_f32(xsum) + ptx::lds_f32(xsum + kBM * 4);
      const float l1 = ptx::lds_f32(xsum + 2 * kBM * 4) + ptx::lds_f32(xsum + 3 * kBM * 4);
      const float m1 = ptx::lds_f32(xrow);
      const float m = fmaxf(m_run, m1);  // set 0 always owns key tile 0, so m is finite
      const float w0 = fast_exp2((m_run - m) * scale), w1 = fast_exp2((m1 - m) * scale);
      const float l = l0 * w0 + l1 * w1;
      const float inv = (i < len && l > 0.f) ? 1.f / l : 0.f;  // padded query rows -> zeros
      const float c0 = w0 * inv, c1 = w1 * inv;
      if (i < S) {
        uint4* o = reinterpret_cast<uint4*>(p.ctx + (rb + i) * p.Dp + h * kDK + half * 32);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float4 x0 = lds_f32x4(xrow + 16 + half * 128 + 32 * c), x1 = lds_f32x4(xrow + 32 + half * 128 + 32 * c);
          uint4 u;
          u.x = ptx::pack_bf16x2(fmaf(o_acc[8 * c + 0], c0, x0.x * c1), fmaf(o_acc[8 * c + 1], c0, x0.y * c1));
          u.y = ptx::pack_bf16x2(fmaf(o_acc[8 * c + 2], c0, x0.z * c1), fmaf(o_acc[8 * c + 3], c0, x0.w * c1));
          u.z = ptx::pack_bf16x2(fmaf(o_acc[8 * c + 4], c0, x1.x * c1), fmaf(o_acc[8 * c + 5], c0, x1.y * c1));
          u.w = ptx::pack_bf16x2(fmaf(o_acc[8 * c + 6], c0, x1.z * c1), fmaf(o_acc[8 * c + 7], c0, x1.w * c1));
          o[c] = u;
        }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace

int launch_attn_tc5(const AttnDesc& a, cudaStream_t st, std::string* err) {
  if ((a.tiles == nullptr && a.B <= 0) || a.T <= 0) return 0;
  if (a.dkp != kDK) {
    if (err) *err = "attn_tc5: padded head dim must be 64";
    return -1;
  }
  const int Dp = a.H * a.dkp;
  const long long rows = a.tiles != nullptr ? a.rows : static_cast<long long>(a.B) * a.T;
  CUtensorMap tmKV, tmP;
  {
    uint64_t dims[2] = {static_cast<uint64_t>(4 * Dp), static_cast<uint64_t>(rows)};
    uint64_t strides[1] = {static_cast<uint64_t>(4 * Dp) * 2};
    uint32_t boxk[2] = {kDK, kBN};
    if (!encode_tmap_bf16(&tmKV, a.qkv, 2, dims, strides, boxk, err)) return -1;
  }
  {
    uint64_t dims[2] = {static_cast<uint64_t>(Dp), static_cast<uint64_t>(2 * a.T - 1)};
    uint64_t strides[1] = {static_cast<uint64_t>(a.ld_pos) * 2};
    uint32_t box[2] = {kDK, 64};
    if (!encode_tmap_bf16(&tmP, a.pos, 2, dims, strides, box, err)) return -1;
  }
  static bool configured[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!configured[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(rel_attn_tc5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal);
    if (e != cudaSuccess) {
      if (err) *err = std::string("cudaFuncSetAttribute(attn_tc5): ") + cudaGetErrorString(e);
      return static_cast<int>(e);
    }
    configured[dev & 63] = true;
  }
  Attn5Params p;
  p.qkv = reinterpret_cast<const bf16*>(a.qkv);
  p.lens = a.lens;
  p.ctx = reinterpret_cast<bf16*>(a.ctx);
  p.T = a.T;
  p.Dp = Dp;
  p.tiles = a.tiles;
  p.scale_log2 = 1.4426950408889634f / sqrtf(static_cast<float>(a.dk));
  p.flags = getenv("CFB_ATTN5_FLAGS") ? atoi(getenv("CFB_ATTN5_FLAGS")) : 0;
  dim3 grid((a.T + kBM - 1) / kBM, a.H, a.B);
  if (a.tiles != nullptr) grid = dim3(a.n_tiles, a.H, 1);
  cudaError_t e = launch_pdl(rel_attn_tc5_kernel, grid, dim3(kThreads), kSmemTotal, st, tmKV, tmP, p);
  if (e != cudaSuccess) {
    if (err) *err = std::string("attn_tc5 launch: ") + cudaGetErrorString(e);
    return static_cast<int>(e);
  }
  return 0;
}

}  // namespace cfb
