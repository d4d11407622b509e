// tcgen05 GEMM whose A operand is produced by a LayerNorm prologue and stays resident in TENSOR MEMORY:
//
//   [y = LN1(x); x_out = y]                 optional: norm_out of the previous layer (conformer_modules.py:120)
//   a = bf16(LN2(y or x))                   the block's input LayerNorm (:98, :103, :112, :116)
//   D = a W^T  with the fused epilogues of epilogue.cuh (Swish / QKV split / GLU + mask)
//
// One CTA per 128-row block (grid = row blocks; a second wave simply follows the first).  The eight epilogue warps
// first normalise the block's rows: one warp per row, the fp32 rows arrive through a per-warp ring of six 1-D bulk
// copies (cp.async.bulk + mbarrier: 96 KB in flight per SM -- with plain loads one row per warp was in flight and the
// prologue took 20 us), statistics exactly like layernorm_kernel (same lane -> column mapping, same summation order:
// results are bit-identical to LayerNorm kernel + GEMM kernel), 32 rows = one TMEM lane quarter at a time into a small
// shared-memory chunk; the two warps that may address that lane quarter then move the chunk into TMEM columns
// [0, d/2) with tcgen05.st (row = lane, two bf16 per column): the UMMA A operand (tcgen05.mma with A in tensor memory).
// The block then runs ALL its N tiles (128 wide, accumulators double-buffered in TMEM columns [256, 512)) against that
// resident operand: only W streams through the TMA ring (16 KB per k-block, 9 stages), the bf16 copy of LN(x) never
// exists in HBM or shared memory, and the stand-alone LayerNorm launch disappears.  The row ring aliases W stages
// 3..8, which the producer leaves alone until the operand is complete.
// (gemm_lna.cu kept A in shared memory instead: at d = 512 that left room for two W stages only, which starved the
// tensor pipe -- measured 63 us vs 48 us for LayerNorm + GEMM.)
//   warp 0      TMA producer (W boxes only)      warp 1      MMA issuer
//   warps 2..9  LayerNorm prologue, then the epilogue of every N tile
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "epilogue.cuh"
#include "ptx.cuh"

namespace cfb {
namespace {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kBN = 128;
constexpr int kEpiWarps = 8;
constexpr int kThreads = 64 + 32 * kEpiWarps;
constexpr int kMaxStages = 10;
constexpr int kBBytes = kBN * kBlockK * 2;       // one W stage: 16 KB
constexpr uint32_t kColAcc = 256;                // TMEM: A in [0, 256), two 128-column accumulators behind it
constexpr int kStagingBytes = 68 * 1024;         // epilogue: 2 output boxes per warp (64 KB); prologue: 2 row chunks
constexpr int kMaxVec = 4;                       // float4 per lane: d <= 512

struct LnaParams {
  int num_m_blocks, num_n_tiles, num_k_blocks, stages;
  int d;                 // K = LayerNorm width (d % 4 == 0, d <= 512)
  int off_staging, off_bar, smem_needed;
  int chunk_pitch;       // bytes per row of a prologue chunk: d * 2 + 16 (16-byte reads by row stay conflict-free)
  const float* x;        // (M, d) fp32
  long long ldx;
  const float* gamma1;   // optional first LayerNorm (norm_out): y = LN1(x) is written back to x_out
  const float* beta1;
  float* x_out;
  const float* gamma2;   // the LayerNorm whose output is the A operand
  const float* beta2;
  long long* trace;  // CFB_LNT_TRACE=1: clock marks of CTA 0 ([0] start, [1..4] quarter q normalised, [5] A complete,
                     // 16 + 4 nt + {0,1,2}: epilogue warp 2 (wait start, acc ready, tile done); 128 + 4 nt + {0,1}: MMA
                     // issuer (acc_empty ok, committed); 256 + i: producer after issuing k-block i (first 64))
  EpiParams ep;
};
#define LNT_MARK(slot)                                                     \
  do {                                                                     \
    if (p.trace != nullptr && blockIdx.x == 0) p.trace[slot] = clock64();  \
  } while (0)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float4 lds128f(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
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

constexpr int kXDepth = 6;    // fp32 rows in flight per prologue warp (1-D bulk copies)
constexpr int kXPitch = 2048; // ring slot: one row of up to 512 floats
constexpr int kWEarly = 3;    // W stages the producer may fill before the operand is complete (the ring aliases the rest)
constexpr int kStagesFixed = 9;
static_assert(kEpiWarps * kXDepth * kXPitch == (kStagesFixed - kWEarly) * kBBytes, "row ring = W stages 3..8");

__device__ __forceinline__ void bulk_load_row(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

template <int EPI>
__global__ void __launch_bounds__(kThreads, 1)
gemm_lnt_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmO, const LnaParams p) {
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
  const uint32_t a_ready = acc_empty + 16;            // the resident A tile of the row block is written
  const uint32_t tmem_slot = a_ready + 8;
  const uint32_t x_bar = a_ready + 16;                // [kEpiWarps][kXDepth]
  const uint32_t bias_s = x_bar + 8 * kEpiWarps * kXDepth;  // float [2][kBN]
  uint32_t tid;
  asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid));
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int mb = blockIdx.x;

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
      for (int s = 0; s < kEpiWarps * kXDepth; ++s) ptx::mbar_init_a(x_bar + 8 * s, 1);
      ptx::fence_mbar_init();
    }
    __syncwarp();
  } else if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");
  pdl_launch_dependents();
  pdl_wait();
  if (tid == 0) LNT_MARK(0);

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer: W only
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      bool ring_free = false;
      for (int nt = 0; nt < p.num_n_tiles; ++nt) {
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          if (!ring_free && stage >= kWEarly) {  // stages >= kWEarly hold the prologue's row ring until A is complete
            ptx::mbar_wait_a(a_ready, 0);
            ring_free = true;
          }
          ptx::mbar_wait_a(empty_bar + 8 * stage, phase ^ 1);
          ptx::mbar_arrive_expect_tx_a(full_bar + 8 * stage, kBBytes);
          ptx::tma_load_2d_a(sbase + stage * kBBytes, &tmW, full_bar + 8 * stage, kb * kBlockK, nt * kBN);
          if (nt * p.num_k_blocks + kb < 64) LNT_MARK(256 + nt * p.num_k_blocks + kb);
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = ptx::make_idesc_bf16(kBlockM, kBN, 0, 0);
    int stage = 0;
    uint32_t phase = 0;
    ptx::mbar_wait_a(a_ready, 0);
    ptx::tc_fence_after();
    for (int nt = 0; nt < p.num_n_tiles; ++nt) {
      const int buf = nt & 1;
      ptx::mbar_wait_a(acc_empty + 8 * buf, ((nt >> 1) & 1) ^ 1);
      ptx::tc_fence_after();
      if (lane == 0 && nt < 24) LNT_MARK(128 + 4 * nt);
      const uint32_t d_tmem = tmem_base + kColAcc + buf * kBN;
      for (int kb = 0; kb < p.num_k_blocks; ++kb) {
        ptx::mbar_wait_a(full_bar + 8 * stage, phase);
        ptx::tc_fence_after();
        const uint32_t ta = tmem_base + kb * 32;  // 64 K-columns of A = 32 TMEM columns, 8 per K = 16 step
        const uint64_t db = ptx::make_sdesc_sw128(sbase + stage * kBBytes, 16, 1024);
        if (ptx::elect_one()) {
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k)
            ptx::umma_bf16_ts(d_tmem, ta + 8 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          ptx::tc_commit_a(empty_bar + 8 * stage);
          if (kb == p.num_k_blocks - 1) {
            ptx::tc_commit_a(acc_full + 8 * buf);
            if (nt < 24) LNT_MARK(128 + 4 * nt + 1);
          }
        }
        __syncwarp();
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1;
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
    const uint32_t sbuf0 = sbase + p.off_staging + ew * 8192;  // two output boxes per warp
    uint32_t box_counter = 0;
    const uint32_t swz = static_cast<uint32_t>(lane & 7);
    const int d = p.d;
    const int nvec = d >> 2;
    const float inv_d = 1.0f / static_cast<float>(d);
    {
      // ---- LayerNorm prologue, one TMEM lane quarter (32 rows) at a time: every warp normalises 4 rows of the
      // quarter into the shared-memory chunk (double-buffered), then the two warps that own that lane quarter move
      // the chunk into TMEM (thread = row) while everybody else goes on with the next quarter.
      const uint32_t chunk0 = sbase + p.off_staging;
      const uint32_t chunk_bytes = 32u * static_cast<uint32_t>(p.chunk_pitch);
      const int words = d >> 1;  // packed bf16 pairs per row = TMEM columns of A
      const uint32_t ring = sbase + kWEarly * kBBytes + ew * (kXDepth * kXPitch);
      const uint32_t my_bar = x_bar + 8 * (ew * kXDepth);
      // float4 index of slot k (as layernorm_kernel): pair k / 2 covers float4 [64 (k/2) + 2 lane, + 1]
      auto vidx = [&](int k) { return 64 * (k >> 1) + 2 * lane + (k & 1); };
      // this warp's i-th row (i = 0..15): quarter i / 4, row (i % 4) * 8 + ew of the quarter
      auto row_of = [&](int i) { return (i >> 2) * 32 + (i & 3) * kEpiWarps + ew; };
      auto issue = [&](int i) {
        const long long row = static_cast<long long>(mb) * kBlockM + row_of(i);
        if (lane == 0 && row < p.ep.M) {
          const uint32_t slot = static_cast<uint32_t>(i % kXDepth);
          ptx::mbar_arrive_expect_tx_a(my_bar + 8 * slot, static_cast<uint32_t>(d) * 4u);
          bulk_load_row(ring + slot * kXPitch, p.x + row * p.ldx, static_cast<uint32_t>(d) * 4u, my_bar + 8 * slot);
        }
      };
#pragma unroll
      for (int i = 0; i < kXDepth; ++i) issue(i);
      float4 g2[kMaxVec], b2[kMaxVec];
#pragma unroll
      for (int k = 0; k < kMaxVec; ++k)
        if (vidx(k) < nvec) {
          g2[k] = __ldg(reinterpret_cast<const float4*>(p.gamma2) + vidx(k));
          b2[k] = __ldg(reinterpret_cast<const float4*>(p.beta2) + vidx(k));
        }
      auto copy_quarter = [&](int q) {
        // this warp's lane quarter: row `lane` of chunk q -> TMEM lane 32 q + lane, this warp's half of the columns
        const uint32_t src = chunk0 + (q & 1) * chunk_bytes + lane * p.chunk_pitch;
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        const int w_half = (words + 63) / 64 * 32;  // columns per half, a multiple of 32
        for (int c0 = half * w_half; c0 < words && c0 < (half + 1) * w_half; c0 += 32) {
          uint32_t v[32];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 t = lds128f(src + (c0 + 4 * j) * 4);
            v[4 * j] = __float_as_uint(t.x), v[4 * j + 1] = __float_as_uint(t.y);
            v[4 * j + 2] = __float_as_uint(t.z), v[4 * j + 3] = __float_as_uint(t.w);
          }
          ptx::tmem_st_x32(t_row + c0, v);
        }
      };
#pragma unroll 1
      for (int i = 0; i < kBlockM / kEpiWarps; ++i) {
        const int q = i >> 2;
        const int r = row_of(i);
        const long long row = static_cast<long long>(mb) * kBlockM + r;
        const uint32_t dst = chunk0 + (q & 1) * chunk_bytes + (r & 31) * p.chunk_pitch;
        if (row < p.ep.M) {
          const uint32_t slot = static_cast<uint32_t>(i % kXDepth);
          const bool trr = tid == 64 && i < 8;
          if (trr) LNT_MARK(400 + 8 * i);
          ptx::mbar_wait_a(my_bar + 8 * slot, static_cast<uint32_t>(i / kXDepth) & 1u);
          if (trr) LNT_MARK(400 + 8 * i + 1);
          float4 v[kMaxVec];
#pragma unroll
          for (int k = 0; k < kMaxVec; ++k)
            if (vidx(k) < nvec) v[k] = lds128f(ring + slot * kXPitch + vidx(k) * 16);
          auto stats = [&](float& mean, float& rstd) {
            float s = 0.f;
#pragma unroll
            for (int k = 0; k < kMaxVec; ++k)
              if (vidx(k) < nvec) s += (v[k].x + v[k].y) + (v[k].z + v[k].w);
            mean = warp_sum(s) * inv_d;
            float qq = 0.f;
#pragma unroll
            for (int k = 0; k < kMaxVec; ++k)
              if (vidx(k) < nvec) {
                v[k].x -= mean, v[k].y -= mean, v[k].z -= mean, v[k].w -= mean;
                qq += (v[k].x * v[k].x + v[k].y * v[k].y) + (v[k].z * v[k].z + v[k].w * v[k].w);
              }
            rstd = 1.0f / sqrtf(warp_sum(qq) * inv_d + 1e-5f);
          };
          float mean, rstd;
          stats(mean, rstd);
          if (trr) LNT_MARK(400 + 8 * i + 2);
          // the warp-wide sums consumed every lane's loads: the ring slot may be refilled
          if (i + kXDepth < kBlockM / kEpiWarps) issue(i + kXDepth);
          if (trr) LNT_MARK(400 + 8 * i + 3);
          if (p.gamma1 != nullptr) {
            // y = LN1(x) goes back to the fp32 residual stream and is normalised again for the GEMM operand
            float4* yo = reinterpret_cast<float4*>(p.x_out + row * p.ldx);
#pragma unroll
            for (int k = 0; k < kMaxVec; ++k) {
              const int c = vidx(k);
              if (c < nvec) {
                const float4 g = __ldg(reinterpret_cast<const float4*>(p.gamma1) + c);
                const float4 b = __ldg(reinterpret_cast<const float4*>(p.beta1) + c);
                v[k].x = fmaf(v[k].x * rstd, g.x, b.x), v[k].y = fmaf(v[k].y * rstd, g.y, b.y);
                v[k].z = fmaf(v[k].z * rstd, g.z, b.z), v[k].w = fmaf(v[k].w * rstd, g.w, b.w);
                yo[c] = v[k];
              }
            }
            stats(mean, rstd);
          }
#pragma unroll
          for (int k = 0; k < kMaxVec; ++k) {
            const int c = vidx(k);  // float4 index: columns 4 c .. 4 c + 3 -> packed words 2 c, 2 c + 1
            if (c < nvec)
              sts64(dst + c * 8, ptx::pack_bf16x2(fmaf(v[k].x * rstd, g2[k].x, b2[k].x), fmaf(v[k].y * rstd, g2[k].y, b2[k].y)),
                    ptx::pack_bf16x2(fmaf(v[k].z * rstd, g2[k].z, b2[k].z), fmaf(v[k].w * rstd, g2[k].w, b2[k].w)));
          }
          if (trr) LNT_MARK(400 + 8 * i + 4);
        } else {  // tail rows: a zero operand row
#pragma unroll
          for (int k = 0; k < kMaxVec; ++k)
            if (vidx(k) < nvec) sts64(dst + vidx(k) * 8, 0u, 0u);
        }
        if ((i & 3) == 3) {
          // chunk q is complete; chunk (q - 1)'s copy warps passed this barrier only after finishing their copy, so
          // the buffer that quarter q + 1 is about to overwrite is free
          asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");
          if (tid == 64) LNT_MARK(1 + q);
          if (quarter == q) copy_quarter(q);
        }
      }
      ptx::tc_wait_st();
      ptx::tc_fence_before();
      ptx::mbar_arrive_a(a_ready);
      if (tid == 64) LNT_MARK(5);
      // the epilogue staging boxes alias the chunk buffers: nobody may start storing output boxes before the last copy
      asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");
    }

    // ---- epilogue of every N tile of this row block
    const long long out_row = static_cast<long long>(mb) * kBlockM + row_in_tile;
    const bool row_ok = out_row < p.ep.M;
    const int row0 = mb * kBlockM + quarter * 32;
    for (int nt = 0; nt < p.num_n_tiles; ++nt) {
      const int buf = nt & 1;
      const bool trc = tid == 64 && nt < 24;
      if (trc) LNT_MARK(16 + 4 * nt);
      {  // this tile's bias -> shared memory while the accumulator is still being computed
        const int e = static_cast<int>(tid) - 64;
        if (e < kBN) {
          const int col = nt * kBN + e;
          ptx::sts_f32(bias_s + (buf * kBN + e) * 4, (p.ep.bias != nullptr && col < p.ep.N) ? __ldg(p.ep.bias + col) : 0.f);
        }
        asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");
      }
      ptx::mbar_wait_a(acc_full + 8 * buf, (nt >> 1) & 1);
      ptx::tc_fence_after();
      if (trc) LNT_MARK(16 + 4 * nt + 1);
      const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + kColAcc + buf * kBN;
#pragma unroll 1
      for (int box = half; box < kBoxes; box += kEpiWarps / 4) {
        const int acc_col0 = nt * kBN + box * kAccPerBox;
        if (acc_col0 >= p.ep.N) break;  // warp-uniform: nothing of this box is inside the matrix
        const int n_pass = (EPI == EPI_QKV && acc_col0 < p.ep.qkv_dp) ? 2 : 1;
#pragma unroll 1
        for (int pass = 0; pass < n_pass; ++pass) {
          const uint32_t sbuf = sbuf0 + (box_counter & 1u) * 4096;
          const uint32_t srow = sbuf + lane * 128;
          ++box_counter;
          if (lane == 0) ptx::bulk_wait_read<1>();  // the store that last read this buffer has drained it
          __syncwarp();
#pragma unroll
          for (int ch = 0; ch < kChunks; ++ch) {
            uint32_t v[32];
            ptx::tmem_ld_x32(t_addr + box * kAccPerBox + ch * 32, v);
            ptx::tc_wait_ld();
            float acc[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[j] = __uint_as_float(v[j]);
            if (row_ok) {
              if (EPI == EPI_QKV && pass == 1) {
                epi_compute<EPI, kFast>(p.ep, out_row, acc_col0 + ch * 32, acc, pass);  // q + v: bias2 from global
              } else {
                float b[32];
                const uint32_t bs = bias_s + (buf * kBN + box * kAccPerBox + ch * 32) * 4;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const float4 t = lds128f(bs + j * 16);
                  b[4 * j] = t.x, b[4 * j + 1] = t.y, b[4 * j + 2] = t.z, b[4 * j + 3] = t.w;
                }
                epi_math<EPI, kFast>(p.ep, out_row, acc, b);
              }
            }
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
      if (trc) LNT_MARK(16 + 4 * nt + 2);
    }
    if (lane == 0) ptx::bulk_wait_read<0>();  // the staging boxes must outlive their reads
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
  auto kern = gemm_lnt_kernel<EPI>;
  static bool configured[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!configured[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) {
      if (err) *err = std::string("cudaFuncSetAttribute(gemm_lnt): ") + cudaGetErrorString(e);
      return static_cast<int>(e);
    }
    configured[dev & 63] = true;
  }
  cudaError_t e = launch_pdl(kern, dim3(grid), dim3(kThreads), smem, st, tmW, tmO, p);
  if (e != cudaSuccess) {
    if (err) *err = std::string("gemm_lnt launch: ") + cudaGetErrorString(e);
    return static_cast<int>(e);
  }
  return 0;
}

}  // namespace

long long* g_lnt_trace = nullptr;

int launch_gemm_lnt(const GemmLnaDesc& g, cudaStream_t st, std::string* err) {
  if (g.M <= 0 || g.N <= 0) return 0;
  if (g.d % 64 != 0 || g.d > 512 || (g.ldw % 8) || (g.ldx % 4)) {
    if (err) *err = "gemm_lnt: d must be a multiple of 64 and <= 512";
    return -1;
  }
  if (g.epi != EPI_SWISH && g.epi != EPI_QKV && g.epi != EPI_GLU && g.epi != EPI_LINEAR) {
    if (err) *err = "gemm_lnt: unsupported epilogue";
    return -1;
  }
  LnaParams p{};
  p.d = g.d;
  p.num_m_blocks = (g.M + kBlockM - 1) / kBlockM;
  p.num_n_tiles = (g.N + kBN - 1) / kBN;
  p.num_k_blocks = (g.d + kBlockK - 1) / kBlockK;
  p.chunk_pitch = g.d * 2 + 16;
  if (2 * 32 * p.chunk_pitch > kStagingBytes) {
    if (err) *err = "gemm_lnt: chunk buffers do not fit";
    return -1;
  }
  constexpr int kBarBytes = 2048;  // ring barriers (20 + 5 + 48) x 8 B, TMEM slot, bias float[2][128]
  const int stages = kStagesFixed;
  static_assert(kStagesFixed * kBBytes + kStagingBytes + kBarBytes + 1024 <= 227 * 1024, "shared memory budget");
  p.stages = stages;
  p.off_staging = stages * kBBytes;
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
  p.trace = nullptr;
  if (getenv("CFB_LNT_TRACE")) {
    if (!g_lnt_trace) cudaMalloc(&g_lnt_trace, 512 * sizeof(long long));
    cudaMemsetAsync(g_lnt_trace, 0, 512 * sizeof(long long), st);
    p.trace = g_lnt_trace;
  }

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
      if (err) *err = "gemm_lnt: output leading dimension must be a multiple of 16 bytes";
      return -1;
    }
    uint64_t dims[2] = {cols, static_cast<uint64_t>(g.M)};
    uint64_t strides[1] = {static_cast<uint64_t>(g.ep.ldo) * 2};
    uint32_t box[2] = {64u, 32u};
    if (!encode_tmap_bf16(&tmO, g.ep.out, 2, dims, strides, box, err)) return -1;
  }
  const int grid = p.num_m_blocks;  // one CTA per row block; later waves follow as SMs free up
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

// debug: clock marks of the last traced launch (512 values; CFB_LNT_TRACE=1)
extern "C" __attribute__((visibility("default"))) int cfb_debug_lnt_trace(long long* host_out) {
  if (!cfb::g_lnt_trace) return 1;
  cudaDeviceSynchronize();
  return cudaMemcpy(host_out, cfb::g_lnt_trace, 512 * sizeof(long long), cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : 2;
}
