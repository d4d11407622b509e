// Fused relative-position attention for sm_100a (flash-style; no T x T matrix ever reaches HBM).
//
//   ctx[i,:] = softmax_j( ((q_i+u).k_j + (q_i+v).p_{T-1+j-i}) / sqrt(dk) ) v_j        (multi_head_attention.py:195-210)
//
// One CTA per (128-query tile, head, sequence); 384 threads = 3 warpgroups (registers re-balanced with setmaxnreg):
//   warp 0      TMA producer: Q+u, Q+v once; per 64-key tile K and V (3-stage ring) and the 64 NEW rows of the
//               192-row band of linear_pos(pos_emb) the tile can touch (rows T-1+j0-i0-127 .. +191): consecutive key
//               tiles share two thirds of their band, so the band lives in a 5-block ring of 64-row blocks
//   warps 1, 2  MMA issuers (tcgen05, cta_group::1), one per softmax set: S = (Q+u) K^T (128x64) and
//               G = (Q+v) Pband^T (128x192) into the set's TMEM buffer, later O_part = P V (128x64) into the S
//               columns of the same buffer.  Within a set the order S/G -> P V -> next S/G is a chain, so each
//               issuer simply blocks on its own barriers while the other set's issuer proceeds
//   warps 4..11 two softmax sets of four warps (one query row per thread).  Set s owns key tiles s, s+2, ... with its
//               own TMEM buffer, probability tile and running (max, sum, O); the sets ping-pong so one set's exp /
//               shift work overlaps the other's MMAs, and are merged (log-sum-exp) at the end.
//               rel_shift is an index remap -- row ii needs G[ii][127-ii+jj] -- done as a warp-uniform TMEM column
//               offset plus a per-lane offset applied through a private shared-memory row (16-byte stores, 4-byte
//               loads, conflict-free pitch); online softmax in fp32 (exp2); probabilities written as bf16 into a
//               128-byte-swizzled smem tile for the PV MMA.
// Keys j >= len[b] are masked to -inf (the reference's -10000 underflows to exactly 0 for valid rows); query rows
// i >= len[b] are written as zeros (multi_head_attention.py:104-113).
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"

namespace cfb {
namespace {

constexpr int kBM = 128;   // queries per CTA
constexpr int kBN = 64;    // keys per tile
constexpr int kDK = 64;    // padded head dim
constexpr int kBand = 192; // >= kBM + kBN - 1, multiple of 64
constexpr int kThreads = 384;  // warpgroup 0: TMA + MMA (+2 idle warps); warpgroups 1, 2: softmax sets 0, 1
constexpr int kQBytes = kBM * kDK * 2;     // 16 KB each for Q+u, Q+v
constexpr int kKBytes = kBN * kDK * 2;     // 8 KB
constexpr int kBandBytes = kBand * kDK * 2;  // 24 KB
constexpr int kKVStages = 3;               // K + V tiles, 16 KB per stage
constexpr int kKVBytes = 2 * kKBytes;
constexpr int kBandBlocks = 5;             // ring of 64-row band blocks (8 KB each); a tile reads 3 consecutive ones
constexpr int kBlockBytes = 64 * kDK * 2;
constexpr int kPBytes = kBM * kBN * 2;     // 16 KB probabilities per set
constexpr int kShiftPitch = 68;            // words per private shift row: 63-column window + pad; == 4 (mod 32), even
constexpr int kShiftBytes = 32 * kShiftPitch * 4;  // per softmax warp
constexpr int kOffKV = 2 * kQBytes;
constexpr int kOffBand = kOffKV + kKVStages * kKVBytes;
constexpr int kOffP = kOffBand + kBandBlocks * kBlockBytes;
constexpr int kOffShift = kOffP + 2 * kPBytes;
constexpr int kOffBar = kOffShift + 8 * kShiftBytes;
constexpr int kSmemTotal = kOffBar + 256 + 1024;
constexpr uint32_t kTmemCols = 512;  // two buffers (one per set) of [S 64 | G 192]

struct AttnParams {
  const int32_t* lens;
  bf16* ctx;
  int T, Dp;
  float scale_log2;  // log2(e) / sqrt(dk)
  int debug;         // CFB_ATTN_DEBUG ablation bits (timing experiments only; results are wrong when non-zero)
  long long* trace;  // debug bit 8: clock64 trace of CTA (0,0,0): [0..] softmax warp 4 lane 0, [512..] issuer of set 0
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
    if (warp >= 4 && warp < 8) {
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
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint64_t* q_full = bars + 0;
  uint64_t* sg_full = bars + 1;   // [2] per softmax set: S and G of the set's current key tile are in TMEM
  uint64_t* s_free = bars + 3;    // [2] O_part folded: the S columns may be overwritten
  uint64_t* g_free = bars + 5;    // [2] G window loaded: the G columns may be overwritten (next tile's G issued early)
  uint64_t* p_ready = bars + 7;   // [2]
  uint64_t* o_full = bars + 9;    // [2]
  uint64_t* kv_full = bars + 11;                  // [kKVStages]
  uint64_t* kv_empty = kv_full + kKVStages;       // [kKVStages]
  uint64_t* band_full = kv_empty + kKVStages;     // [kBandBlocks]
  uint64_t* band_empty = band_full + kBandBlocks; // [kBandBlocks]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(band_empty + kBandBlocks);

  const int n_kt = (len + kBN - 1) / kBN;

  if (warp == 0) {
    if (lane == 0) {
      ptx::prefetch_tmap(&tmQ);
      ptx::prefetch_tmap(&tmKV);
      ptx::prefetch_tmap(&tmP);
      ptx::mbar_init(q_full, 1);
      for (int s = 0; s < kKVStages; ++s) {
        ptx::mbar_init(&kv_full[s], 1);
        ptx::mbar_init(&kv_empty[s], 1);
      }
      for (int s = 0; s < kBandBlocks; ++s) {
        ptx::mbar_init(&band_full[s], 1);
        ptx::mbar_init(&band_empty[s], 1);
      }
      for (int s = 0; s < 2; ++s) {
        ptx::mbar_init(&sg_full[s], 1);
        ptx::mbar_init(&s_free[s], 4);   // one elected arrival per softmax warp
        ptx::mbar_init(&g_free[s], 4);
        ptx::mbar_init(&p_ready[s], 4);
        ptx::mbar_init(&o_full[s], 1);
      }
      ptx::fence_mbar_init();
    }
    __syncwarp();
    ptx::tmem_alloc(tmem_slot, kTmemCols);
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 72;" ::: "memory");
  if (warp == 0) {
    // ---------------------------------------------------------------------------------- TMA producer
    if (lane == 0) {
      const int row_q = b * T + i0;
      ptx::mbar_arrive_expect_tx(q_full, 2 * kQBytes);
      ptx::tma_load_2d(sQu, &tmQ, q_full, h * kDK, row_q);
      ptx::tma_load_2d(sQv, &tmQ, q_full, p.Dp + h * kDK, row_q);
      const int r0 = T - 1 - i0 - (kBM - 1);  // band row of G column 0 for key tile 0 (may be < 0: TMA zero-fills)
      auto load_band_block = [&](int g) {     // block g = band rows r0 + 64 g .. + 63, first needed by key tile g - 2
        const int slot = g % kBandBlocks, use = g / kBandBlocks;
        ptx::mbar_wait(&band_empty[slot], (use & 1) ^ 1);
        ptx::mbar_arrive_expect_tx(&band_full[slot], kBlockBytes);
        ptx::tma_load_2d(smem + kOffBand + slot * kBlockBytes, &tmP, &band_full[slot], h * kDK, r0 + 64 * g);
      };
      load_band_block(0);
      load_band_block(1);
      for (int kt = 0; kt < n_kt; ++kt) {
        const int st = kt % kKVStages, use = kt / kKVStages;
        ptx::mbar_wait(&kv_empty[st], (use & 1) ^ 1);
        ptx::mbar_arrive_expect_tx(&kv_full[st], kKVBytes);
        uint8_t* dst = smem + kOffKV + st * kKVBytes;
        ptx::tma_load_2d(dst, &tmKV, &kv_full[st], 2 * p.Dp + h * kDK, b * T + kt * kBN);
        ptx::tma_load_2d(dst + kKBytes, &tmKV, &kv_full[st], 3 * p.Dp + h * kDK, b * T + kt * kBN);
        load_band_block(kt + 2);
      }
    }
  } else if (warp == 1 || warp == 2) {
    // ---------------------------------------------------------------------------------- MMA issuer of set (warp-1)
    if (lane == 0) {
      const int s = warp - 1;
      constexpr uint32_t idesc_s = ptx::make_idesc_bf16(kBM, kBN, 0, 0);
      constexpr uint32_t idesc_g192 = ptx::make_idesc_bf16(kBM, 192, 0, 0);
      constexpr uint32_t idesc_g128 = ptx::make_idesc_bf16(kBM, 128, 0, 0);
      constexpr uint32_t idesc_g64 = ptx::make_idesc_bf16(kBM, 64, 0, 0);
      constexpr uint32_t idesc_o = ptx::make_idesc_bf16(kBM, kDK, 0, 1);  // B = V is MN-major (keys x dk rows)
      const uint64_t dQu = ptx::make_sdesc_sw128(ptx::smem_u32(sQu), 16, 1024);
      const uint64_t dQv = ptx::make_sdesc_sw128(ptx::smem_u32(sQv), 16, 1024);
      const uint64_t dP = ptx::make_sdesc_sw128(ptx::smem_u32(smem + kOffP + s * kPBytes), 16, 1024);
      const uint32_t band_base = ptx::smem_u32(smem + kOffBand);
      const uint32_t tS = tmem_base + s * 256;
      ptx::mbar_wait(q_full, 0);
      // G = (Q+v) band^T over ring blocks kt, kt+1, kt+2: one 192-wide MMA group, or two when the ring wraps.
      // Issued one tile ahead (as soon as the set has loaded the previous G window), so only the short S MMA and
      // the P V MMA sit on the set's critical path.
      auto issue_g = [&](int kt) {
        for (int g = kt; g <= kt + 2; ++g) ptx::mbar_wait(&band_full[g % kBandBlocks], (g / kBandBlocks) & 1);
        ptx::tc_fence_after();
        const int s0 = kt % kBandBlocks, s1 = (kt + 1) % kBandBlocks, s2 = (kt + 2) % kBandBlocks;
        const uint64_t d0 = ptx::make_sdesc_sw128(band_base + s0 * kBlockBytes, 16, 1024);
        if (s1 == s0 + 1 && s2 == s1 + 1) {
#pragma unroll
          for (int k = 0; k < kDK / 16; ++k) ptx::umma_bf16(tS + kBN, dQv + 2 * k, d0 + 2 * k, idesc_g192, k != 0);
        } else if (s1 != s0 + 1) {  // wrap after the first block
          const uint64_t d1 = ptx::make_sdesc_sw128(band_base + s1 * kBlockBytes, 16, 1024);
#pragma unroll
          for (int k = 0; k < kDK / 16; ++k) ptx::umma_bf16(tS + kBN, dQv + 2 * k, d0 + 2 * k, idesc_g64, k != 0);
#pragma unroll
          for (int k = 0; k < kDK / 16; ++k) ptx::umma_bf16(tS + kBN + 64, dQv + 2 * k, d1 + 2 * k, idesc_g128, k != 0);
        } else {  // wrap after the second block
          const uint64_t d2 = ptx::make_sdesc_sw128(band_base + s2 * kBlockBytes, 16, 1024);
#pragma unroll
          for (int k = 0; k < kDK / 16; ++k) ptx::umma_bf16(tS + kBN, dQv + 2 * k, d0 + 2 * k, idesc_g128, k != 0);
#pragma unroll
          for (int k = 0; k < kDK / 16; ++k) ptx::umma_bf16(tS + kBN + 128, dQv + 2 * k, d2 + 2 * k, idesc_g64, k != 0);
        }
        ptx::tc_commit(&band_empty[s0]);  // block kt is not read by any later tile
      };
      if (s < n_kt) issue_g(s);
      int it = 0;
      for (int kt = s; kt < n_kt; kt += 2, ++it) {
        const int kvs = kt % kKVStages;
        const bool tr = (p.debug & 8) && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && s == 0 && it < 60;
        if (tr) p.trace[512 + it * 8 + 0] = clock64();
        // ---- S of key tile kt (its G is already in flight or done)
        ptx::mbar_wait(&kv_full[kvs], (kt / kKVStages) & 1);
        ptx::mbar_wait(&s_free[s], (it & 1) ^ 1);
        ptx::tc_fence_after();
        if (tr) p.trace[512 + it * 8 + 1] = clock64();
        const uint32_t st = ptx::smem_u32(smem + kOffKV + kvs * kKVBytes);
        const uint64_t dK = ptx::make_sdesc_sw128(st, 16, 1024);
#pragma unroll
        for (int k = 0; k < kDK / 16; ++k) ptx::umma_bf16(tS, dQu + 2 * k, dK + 2 * k, idesc_s, k != 0);
        ptx::tc_commit(&sg_full[s]);
        if (tr) p.trace[512 + it * 8 + 2] = clock64();
        // ---- G of the set's next tile, once this tile's G window has been read
        if (kt + 2 < n_kt) {
          ptx::mbar_wait(&g_free[s], it & 1);
          issue_g(kt + 2);
        }
        if (tr) p.trace[512 + it * 8 + 3] = clock64();
        // ---- O_part = P V once the set has written its probabilities
        ptx::mbar_wait(&p_ready[s], it & 1);
        ptx::tc_fence_after();
        if (tr) p.trace[512 + it * 8 + 4] = clock64();
        // V tile: 64 keys (K of this MMA) x 64 dk (N), 128-byte rows along N -> MN-major, 8-key groups 1 KB apart
        const uint64_t dV = ptx::make_sdesc_sw128(st + kKBytes, 1024, 1024);
#pragma unroll
        for (int k = 0; k < kBN / 16; ++k)
          ptx::umma_bf16(tS, dP + 2 * k, dV + static_cast<uint64_t>(k) * (2048 >> 4), idesc_o, k != 0);
        ptx::tc_commit(&o_full[s]);
        ptx::tc_commit(&kv_empty[kvs]);
        if (tr) p.trace[512 + it * 8 + 5] = clock64();
      }
    }
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 216;" ::: "memory");
    // ---------------------------------------------------------------------------------- softmax warps
    const int quarter = warp & 3;
    const int set = (warp - 4) >> 2;
    const int ii = quarter * 32 + lane;  // query row inside the tile == TMEM lane
    const int i = i0 + ii;
    const uint32_t tS = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + set * 256;
    const int g_base = kBN + (96 - 32 * quarter);  // warp-uniform part of the rel_shift column offset
    const int sh = 31 - lane;                      // per-lane part, applied through the private smem row
    const uint32_t shift_row = ptx::smem_u32(smem + kOffShift + (warp - 4) * kShiftBytes) + lane * kShiftPitch * 4;
    const uint32_t prow = ptx::smem_u32(smem + kOffP + set * kPBytes) + ii * 128;
    float o_acc[kDK];
#pragma unroll
    for (int c = 0; c < kDK; ++c) o_acc[c] = 0.f;
    float m_run = -INFINITY, l_run = 0.f, alpha_prev = 1.f;

    auto fold_o_part = [&]() {  // o_acc = o_acc * alpha_prev + O_part (the S columns of this set's buffer)
      uint32_t a0[32], a1[32];
      ptx::tmem_ld_x32(tS, a0);
      ptx::tmem_ld_x32(tS + 32, a1);
      ptx::tc_wait_ld();
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        o_acc[c] = fmaf(o_acc[c], alpha_prev, __uint_as_float(a0[c]));
        o_acc[32 + c] = fmaf(o_acc[32 + c], alpha_prev, __uint_as_float(a1[c]));
      }
    };

    int it = 0;
    for (int kt = set; kt < n_kt; kt += 2, ++it) {
      const int j0 = kt * kBN;
      if (it > 0) {
        const bool tr2 = (p.debug & 8) && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && warp == 4 && lane == 0 && it < 60;
        ptx::mbar_wait(&o_full[set], (it - 1) & 1);
        ptx::tc_fence_after();
        if (tr2) p.trace[it * 8 + 5] = clock64();
        fold_o_part();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&s_free[set]);  // the S columns may now receive this tile's scores
        if (tr2) p.trace[it * 8 + 6] = clock64();
      }
      const bool tr = (p.debug & 8) && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && warp == 4 && lane == 0 && it < 60;
      if (tr) p.trace[it * 8 + 0] = clock64();
      ptx::mbar_wait(&sg_full[set], it & 1);
      ptx::tc_fence_after();
      if (tr) p.trace[it * 8 + 1] = clock64();
      float sv[kBN];
      {
        // S (64 columns) and the first half of the G window in flight together; the second half of the window is
        // requested before the first half is consumed, so its TMEM latency hides behind the shared-memory shift.
        uint32_t s0r[32], s1r[32], w0[32], w1[32];
        ptx::tmem_ld_x32(tS, s0r);
        ptx::tmem_ld_x32(tS + 32, s1r);
        ptx::tmem_ld_x32(tS + g_base, w0);
        ptx::tmem_ld_x32(tS + g_base + 32, w1);
        ptx::tc_wait_ld();
        if (!(p.debug & 4)) {
#pragma unroll
          for (int v4 = 0; v4 < 8; ++v4) {
            ptx::sts128(shift_row + v4 * 16, w0[4 * v4], w0[4 * v4 + 1], w0[4 * v4 + 2], w0[4 * v4 + 3]);
            ptx::sts128(shift_row + 128 + v4 * 16, w1[4 * v4], w1[4 * v4 + 1], w1[4 * v4 + 2], w1[4 * v4 + 3]);
          }
        }
        ptx::tmem_ld_x32(tS + g_base + 64, w0);  // window of keys 32..63 = G columns g_base+32 (still in w1) .. +95
        {
          float g[32];
          ptx::lds_f32x32(shift_row + sh * 4, g);
#pragma unroll
          for (int c = 0; c < 32; ++c) sv[c] = (__uint_as_float(s0r[c]) + g[c]) * p.scale_log2;
        }
        ptx::tc_wait_ld();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&g_free[set]);  // every TMEM read of this tile's S / G has completed
        if (!(p.debug & 4)) {
#pragma unroll
          for (int v4 = 0; v4 < 8; ++v4) {
            ptx::sts128(shift_row + v4 * 16, w1[4 * v4], w1[4 * v4 + 1], w1[4 * v4 + 2], w1[4 * v4 + 3]);
            ptx::sts128(shift_row + 128 + v4 * 16, w0[4 * v4], w0[4 * v4 + 1], w0[4 * v4 + 2], w0[4 * v4 + 3]);
          }
        }
        {
          float g[32];
          ptx::lds_f32x32(shift_row + sh * 4, g);
#pragma unroll
          for (int c = 0; c < 32; ++c) sv[32 + c] = (__uint_as_float(s1r[c]) + g[c]) * p.scale_log2;
        }
      }
      if (tr) p.trace[it * 8 + 2] = clock64();
      if (j0 + kBN > len) {  // only the last key tile can contain masked keys
#pragma unroll
        for (int c = 0; c < kBN; ++c)
          if (j0 + c >= len) sv[c] = -INFINITY;
      }
      float mx4[4] = {sv[0], sv[1], sv[2], sv[3]};  // four independent chains instead of one 64-deep one
#pragma unroll
      for (int c = 4; c < kBN; ++c) mx4[c & 3] = fmaxf(mx4[c & 3], sv[c]);
      const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
      const float m_new = fmaxf(m_run, mx);  // finite: every tile holds at least one key j < len
      const float alpha = fast_exp2(m_run - m_new);
      float rs4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int c = 0; c < kBN; ++c) {
        sv[c] = (p.debug & 1) ? (sv[c] - m_new) : fast_exp2(sv[c] - m_new);
        rs4[c & 3] += sv[c];
      }
      const float rsum = (rs4[0] + rs4[1]) + (rs4[2] + rs4[3]);
      l_run = fmaf(l_run, alpha, rsum);
      m_run = m_new;
      alpha_prev = alpha;
      if (tr) p.trace[it * 8 + 3] = clock64();
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        uint4 u;
        u.x = ptx::pack_bf16x2(sv[8 * c + 0], sv[8 * c + 1]);
        u.y = ptx::pack_bf16x2(sv[8 * c + 2], sv[8 * c + 3]);
        u.z = ptx::pack_bf16x2(sv[8 * c + 4], sv[8 * c + 5]);
        u.w = ptx::pack_bf16x2(sv[8 * c + 6], sv[8 * c + 7]);
        ptx::sts128(prow + ((c ^ (ii & 7)) << 4), u.x, u.y, u.z, u.w);  // 128-byte swizzle, K-major
      }
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&p_ready[set]);
      if (tr) p.trace[it * 8 + 4] = clock64();
    }
    if (it > 0) {  // drain the last tile of this set
      ptx::mbar_wait(&o_full[set], (it - 1) & 1);
      ptx::tc_fence_after();
      fold_o_part();
    }
    // ---- merge the two sets (log-sum-exp) and write the context rows
    const uint32_t xrow = ptx::smem_u32(smem + kOffShift + (quarter + 4) * kShiftBytes) + lane * kShiftPitch * 4;
    if (set == 1) {
      ptx::sts_f32(xrow, m_run);
      ptx::sts_f32(xrow + 4, l_run);
#pragma unroll
      for (int c = 0; c < kDK; ++c) ptx::sts_f32(xrow + 8 + 4 * c, o_acc[c]);
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (set == 0) {
      const float m1 = ptx::lds_f32(xrow), l1 = ptx::lds_f32(xrow + 4);
      const float m = fmaxf(m_run, m1);  // set 0 always owns key tile 0, so m is finite
      const float w0 = fast_exp2(m_run - m), w1 = fast_exp2(m1 - m);
      const float l = l_run * w0 + l1 * w1;
      const float inv = (i < len && l > 0.f) ? 1.f / l : 0.f;  // padded query rows -> zeros
      float o1a[32], o1b[32];
      ptx::lds_f32x32(xrow + 8, o1a);
      ptx::lds_f32x32(xrow + 8 + 128, o1b);
      if (i < T) {
        uint4* o = reinterpret_cast<uint4*>(p.ctx + (static_cast<long long>(b) * T + i) * p.Dp + h * kDK);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float r[8];
#pragma unroll
          for (int e = 0; e < 8; ++e)
            r[e] = (o_acc[8 * c + e] * w0 + (c < 4 ? o1a[8 * c + e] : o1b[8 * (c - 4) + e]) * w1) * inv;
          uint4 u;
          u.x = ptx::pack_bf16x2(r[0], r[1]);
          u.y = ptx::pack_bf16x2(r[2], r[3]);
          u.z = ptx::pack_bf16x2(r[4], r[5]);
          u.w = ptx::pack_bf16x2(r[6], r[7]);
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

long long* g_attn_trace = nullptr;
}  // namespace

// debug: copies the clock trace of the last traced launch to the host (1024 values)
extern "C" __attribute__((visibility("default"))) int cfb_debug_attn_trace(long long* host_out) {
  if (!g_attn_trace) return 1;
  cudaDeviceSynchronize();
  return cudaMemcpy(host_out, g_attn_trace, 1024 * sizeof(long long), cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : 2;
}

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
  p.scale_log2 = 1.4426950408889634f / sqrtf(static_cast<float>(a.dk));
  {
    const char* dbg = getenv("CFB_ATTN_DEBUG");
    p.debug = dbg ? atoi(dbg) : 0;
    p.trace = nullptr;
    if (p.debug & 8) {
      static long long* trace_buf = nullptr;
      if (!trace_buf) {
        cudaMalloc(&trace_buf, 1024 * sizeof(long long));
        cudaMemset(trace_buf, 0, 1024 * sizeof(long long));
      }
      p.trace = trace_buf;
      g_attn_trace = trace_buf;
    }
  }
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
