// Residual GEMM with the FOLLOWING LayerNorm(s) in its epilogue, row-complete across a cluster of two CTAs:
//
//   v = x + alpha * (A W^T + bias)                         conformer_modules.py:98-118 (residual updates)
//   y = LN1(v)  if gamma1 else v                            (norm_out, :120)
//   x <- y   (fp32 residual stream, in place)
//   a <- LN2(y)  (bf16: the next block's normalised input -- norm_self_att :103, norm_conv :112,
//                 norm_feed_forward1 of the next layer :98)
//
// Why a cluster: the residual stream has d_model = 512 columns = the whole tensor memory of one SM as fp32
// accumulators, so a single CTA cannot double-buffer the accumulator and its main loop, residual read and output
// write run back to back (measured in round 1: slower than GEMM + stand-alone LayerNorm).  Here each CTA of a pair
// owns 256 columns of the same 128 rows (2 x 256 TMEM columns: the epilogue of row block i overlaps the main loop of
// row block i+1) and the row statistics are combined through distributed shared memory: every thread owns (row,
// 128 columns), computes (mean, centred second moment) of its quarter row chunk by chunk (Chan's pairwise update),
// writes the pair into BOTH CTAs' exchange arrays (st.shared::cluster), and the four quarters are merged in a fixed
// order -- deterministic, and a row's result depends on nothing but the row.  d_model <= 256 runs without a cluster.
//
//   warp 0      TMA producer (A 128 x 64, W 256 x 64, 128-byte swizzle, 3 stages)
//   warp 1      MMA issuer (tcgen05.mma cta_group::1, 128 x 256 x 16)
//   warps 2..9  epilogue: thread = (row, 128 columns).  Pass 1: x chunk (TMA box) + accumulator -> v back into TMEM,
//               statistics.  [Pass 1b: y = LN1(v) -> TMEM, statistics of y.]  Pass 2: x and a leave through the two
//               4 KB boxes of the warp (swizzled staging, TMA stores).
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"

namespace cfb {
namespace {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kBN = 256;  // accumulator columns per CTA
constexpr int kEpiWarps = 8;
constexpr int kThreads = 64 + 32 * kEpiWarps;
constexpr int kStages = 4;
constexpr int kABytes = kBlockM * kBlockK * 2;
constexpr int kBBytes = kBN * kBlockK * 2;
constexpr int kStageBytes = kABytes + kBBytes;
constexpr int kOffVec = kStages * kStageBytes;                  // float [5][kBN]: alpha*bias, gamma1, beta1, gamma2, beta2
constexpr int kOffStats = kOffVec + 5 * kBN * 4;                // float2 [2][4][128]
constexpr int kOffBar = kOffStats + 2 * 4 * 128 * 8;
constexpr int kNeeded = kOffBar + 256;
constexpr int kTotal = kNeeded + 1024;
static_assert(kTotal <= 227 * 1024, "shared memory budget");

struct LncParams {
  int num_tiles;  // row blocks
  int num_k_blocks;
  int M, N;
  float alpha;
  float* x;       // (M x N) fp32 stream, in place
  long long ldx;
  bf16* out;      // (M x N) bf16
  long long ldo;
  const float* bias;
  const float* g1;
  const float* b1;
  const float* g2;
  const float* b2;
  int dbg;           // CFB_LNC_DEBUG: 1 = epilogue does nothing (bare main loop), timing experiments only
  long long* trace;  // CFB_LNC_TRACE=1: clock64 marks of CTA 0, epilogue warp 0 / MMA issuer (timing experiments)
};

__device__ __forceinline__ float2 lds64f(uint32_t addr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr) : "memory");
  return v;
}
// TMEM <-> registers, shape 16x256b.x4: 16 lanes x 32 columns, 16 registers per thread (layout in the epilogue comment)
__device__ __forceinline__ void tmem_ld_16x256b_x4(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x4.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st_16x256b_x4(uint32_t taddr, const uint32_t* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.16x256b.x4.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ uint32_t mapa_peer(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void sts64_cluster(uint32_t cluster_addr, float a, float b) {
  asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(cluster_addr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void sts64(uint32_t addr, float a, float b) {
  asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void mbar_arrive_release_cluster_local(uint32_t bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_release_cluster_remote(uint32_t cluster_bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
}
__device__ __forceinline__ void mbar_wait_acquire_cluster(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  long long t0 = 0;
  uint32_t spins = 0;
  while (true) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) break;
    if (++spins == 1024u) t0 = clock64();
    if (spins > 1024u && (spins & 1023u) == 0 && clock64() - t0 > CFB_MBAR_TIMEOUT_CYCLES) __trap();
  }
}


template <int CL, bool DUAL, bool FULL>
__global__ void __launch_bounds__(kThreads, 1)
gemm_lnc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                const LncParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  {
    uint32_t dyn;
    asm volatile("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn));
    if (sbase - ptx::smem_u32(smem_raw) + kNeeded > dyn) __trap();  // alignment slack did not fit
  }
  const uint32_t bar0 = sbase + kOffBar;
  const uint32_t full_bar = bar0;                        // [kStages]
  const uint32_t empty_bar = full_bar + 8 * kStages;     // [kStages]
  const uint32_t acc_full = empty_bar + 8 * kStages;     // [2]
  const uint32_t acc_empty = acc_full + 16;              // [2]
  const uint32_t st_bar = acc_empty + 16;                // [2]
  const uint32_t tmem_slot = st_bar + 16;
  static_assert(16 * kStages + 48 + 4 <= 256, "barrier block");
  const uint32_t vec = sbase + kOffVec;
  const uint32_t stats = sbase + kOffStats;
  uint32_t tid;
  asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid));
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const uint32_t rank = (CL == 2) ? ptx::cluster_ctarank() : 0u;
  const int cluster_id = static_cast<int>(blockIdx.x) / CL;
  const int n_clusters = static_cast<int>(gridDim.x) / CL;

  if (warp == 0) {
    if (lane == 0) {
      ptx::prefetch_tmap(&tmA);
      ptx::prefetch_tmap(&tmW);
      for (int s = 0; s < kStages; ++s) {
        ptx::mbar_init_a(full_bar + 8 * s, 1);
        ptx::mbar_init_a(empty_bar + 8 * s, 1);
      }
      for (int b = 0; b < 2; ++b) {
        ptx::mbar_init_a(acc_full + 8 * b, 1);
        ptx::mbar_init_a(acc_empty + 8 * b, 32 * kEpiWarps);
        ptx::mbar_init_a(st_bar + 8 * b, kEpiWarps * CL);
      }
      ptx::fence_mbar_init();
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // per-column vectors of this CTA's 256 columns (weights: not produced by the previous kernel)
  for (int e = static_cast<int>(tid); e < kBN; e += kThreads) {
    const int col = static_cast<int>(rank) * kBN + e;
    const bool ok = col < p.N;
    ptx::sts_f32(vec + (0 * kBN + e) * 4, (ok && p.bias != nullptr) ? p.alpha * __ldg(p.bias + col) : 0.f);
    ptx::sts_f32(vec + (1 * kBN + e) * 4, (DUAL && ok) ? __ldg(p.g1 + col) : 0.f);
    ptx::sts_f32(vec + (2 * kBN + e) * 4, (DUAL && ok) ? __ldg(p.b1 + col) : 0.f);
    ptx::sts_f32(vec + (3 * kBN + e) * 4, ok ? __ldg(p.g2 + col) : 0.f);
    ptx::sts_f32(vec + (4 * kBN + e) * 4, ok ? __ldg(p.b2 + col) : 0.f);
  }
  ptx::tc_fence_before();
  if constexpr (CL == 2) ptx::cluster_sync_all();  // the partner's barriers exist before anything is sent to them
  else __syncthreads();
  ptx::tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = cluster_id; tile < p.num_tiles; tile += n_clusters) {
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          ptx::mbar_wait_a(empty_bar + 8 * stage, phase ^ 1);
          ptx::mbar_arrive_expect_tx_a(full_bar + 8 * stage, kStageBytes);
          const uint32_t sa = sbase + stage * kStageBytes;
          ptx::tma_load_2d_a(sa, &tmA, full_bar + 8 * stage, kb * kBlockK, tile * kBlockM);
          ptx::tma_load_2d_a(sa + kABytes, &tmW, full_bar + 8 * stage, kb * kBlockK, static_cast<int>(rank) * kBN);
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_bf16(kBlockM, kBN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = cluster_id; tile < p.num_tiles; tile += n_clusters, ++it) {
        const int buf = it & 1;
        const bool trm = p.trace != nullptr && blockIdx.x == 0 && it < 4;
        if (trm) p.trace[32 + it * 4] = clock64();
        ptx::mbar_wait_a(acc_empty + 8 * buf, ((it >> 1) & 1) ^ 1);
        ptx::tc_fence_after();
        if (trm) p.trace[32 + it * 4 + 1] = clock64();
        const uint32_t d_tmem = tmem_base + buf * kBN;
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          ptx::mbar_wait_a(full_bar + 8 * stage, phase);
          ptx::tc_fence_after();
          const uint32_t sa = sbase + stage * kStageBytes;
          const uint64_t da = ptx::make_sdesc_sw128(sa, 16, 1024);
          const uint64_t db = ptx::make_sdesc_sw128(sa + kABytes, 16, 1024);
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k)
            ptx::umma_bf16(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          ptx::tc_commit_a(empty_bar + 8 * stage);
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        ptx::tc_commit_a(acc_full + 8 * buf);
        if (trm) p.trace[32 + it * 4 + 2] = clock64();
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps
    // Fragment layout of tcgen05.ld/st.16x256b.x4 (16 lanes x 32 columns per instruction): thread t holds, for
    // k = 0..3, j = 0..1, e = 0..1, register 4k + 2j + e = (lane t/4 + 8j, column 8k + 2(t%4) + e).  Four neighbouring
    // threads own 32 contiguous bytes of a row, so the residual stream is read and written straight from / to global
    // memory in whole sectors (LDG.64 / STG.64) -- no staging boxes, no TMA latency in the epilogue.
    const int ew = warp - 2;
    const int quarter = warp & 3;  // TMEM lanes this warp may access: [32*quarter, +32)
    const int hh = ew >> 2;        // which 128 of the CTA's 256 columns
    const int q4 = lane & 3, r8 = lane >> 2;
    const int part = static_cast<int>(rank) * 2 + hh;  // quarter of the row this warp works on
    const int ccol0 = hh * 128;
    const int gcol0 = static_cast<int>(rank) * kBN + ccol0;
    const int natoms = max(0, min(128, p.N - gcol0)) >> 3;  // valid 8-column atoms of this warp's 128 columns
    const int nch = (natoms + 3) >> 2;
    const float inv_n = 1.0f / static_cast<float>(p.N);
    uint32_t peer_stats = 0, peer_bar = 0;
    if constexpr (CL == 2) {
      peer_stats = mapa_peer(stats, rank ^ 1u);
      peer_bar = mapa_peer(st_bar, rank ^ 1u);
    }
    // tile-invariant weights of the pairwise (count, mean, M2) updates: chunk c brings nb values per row and thread
    float inv_nb[4], wc[4], nawc[4];
    {
      float na = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float nb = 2.f * static_cast<float>(max(0, min(4, natoms - 4 * c)));
        inv_nb[c] = nb > 0.f ? 1.0f / nb : 0.f;
        wc[c] = nb > 0.f ? nb / (na + nb) : 0.f;
        nawc[c] = na * wc[c];
        na += nb;
      }
    }
    const float n_thr = 2.f * static_cast<float>(natoms);  // values per row and thread
    // quarters of a row (128 columns each) in merge order
    float wp[4], nawp[4];
    {
      float na = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float nb = static_cast<float>(max(0, min(128, p.N - 128 * k)));
        wp[k] = nb > 0.f ? nb / (na + nb) : 0.f;
        nawp[k] = na * wp[k];
        na += nb;
      }
    }
    const uint32_t vec_thr = vec + static_cast<uint32_t>(ccol0 + 2 * q4) * 4;  // + (array * kBN + 32 c + 8 k) * 4
    float* const xg = p.x + gcol0 + 2 * q4;
    bf16* const ag = p.out + gcol0 + 2 * q4;

    // the four rows of this thread: ri = 2 * h2 + j -> tile row 32 quarter + 16 h2 + 8 j + r8
    auto row_of = [&](int ri) { return quarter * 32 + 16 * (ri >> 1) + 8 * (ri & 1) + r8; };

    auto exchange = [&](int slot, uint32_t parity, float (&mean)[4], float (&m2)[4], float (&rstd)[4]) {
      // the four threads of a row segment first: equal counts, symmetric form (all four end up with the same bits)
#pragma unroll
      for (int ri = 0; ri < 4; ++ri) {
#pragma unroll
        for (int sh = 1; sh <= 2; sh <<= 1) {
          const float mo = __shfl_xor_sync(0xffffffffu, mean[ri], sh);
          const float qo = __shfl_xor_sync(0xffffffffu, m2[ri], sh);
          const float dlt = mean[ri] - mo;
          mean[ri] = 0.5f * (mean[ri] + mo);
          m2[ri] = (m2[ri] + qo) + dlt * dlt * (0.5f * n_thr * static_cast<float>(sh));
        }
      }
      const uint32_t base = static_cast<uint32_t>(slot) * (4 * 128 * 8);
      if (q4 == 0) {
#pragma unroll
        for (int ri = 0; ri < 4; ++ri) {
          const uint32_t off = base + static_cast<uint32_t>(part * 128 + row_of(ri)) * 8;
          sts64(stats + off, mean[ri], m2[ri]);
          if constexpr (CL == 2) sts64_cluster(peer_stats + off, mean[ri], m2[ri]);
        }
      }
      __syncwarp();
      if (lane == 0) {
        mbar_arrive_release_cluster_local(st_bar + 8 * slot);
        if constexpr (CL == 2) mbar_arrive_release_cluster_remote(peer_bar + 8 * slot);
      }
      mbar_wait_acquire_cluster(st_bar + 8 * slot, parity);
#pragma unroll
      for (int ri = 0; ri < 4; ++ri) {
        float m = 0.f, q = 0.f;
#pragma unroll
        for (int k = 0; k < 2 * CL; ++k) {
          if (128 * k < p.N) {
            const float2 sv = lds64f(stats + base + static_cast<uint32_t>(k * 128 + row_of(ri)) * 8);
            if (k == 0) {
              m = sv.x, q = sv.y;
            } else {
              const float dlt = sv.x - m;
              m = fmaf(dlt, wp[k], m);
              q = q + sv.y + dlt * dlt * nawp[k];
            }
          }
        }
        mean[ri] = m;
        rstd[ri] = 1.0f / sqrtf(q * inv_n + 1e-5f);
      }
    };

    // per-chunk statistics of the 8 values a thread holds per row, merged into the running (mean, M2)
    auto chunk_stats = [&](int c, const float (&val)[32], float (&mean)[4], float (&m2)[4]) {
#pragma unroll
      for (int ri = 0; ri < 4; ++ri) {
        const int h2 = ri >> 1, j = ri & 1;
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {  // invalid atoms hold zeros
          s0 += val[16 * h2 + 4 * k + 2 * j];
          s1 += val[16 * h2 + 4 * k + 2 * j + 1];
        }
        const float cm = (s0 + s1) * inv_nb[c];
        float q0 = 0.f, q1 = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (FULL || 4 * c + k < natoms) {
            const float d0 = val[16 * h2 + 4 * k + 2 * j] - cm;
            const float d1 = val[16 * h2 + 4 * k + 2 * j + 1] - cm;
            q0 = fmaf(d0, d0, q0);
            q1 = fmaf(d1, d1, q1);
          }
        }
        if (c == 0) {
          mean[ri] = cm, m2[ri] = q0 + q1;
        } else {
          const float dlt = cm - mean[ri];
          mean[ri] = fmaf(dlt, wc[c], mean[ri]);
          m2[ri] = m2[ri] + (q0 + q1) + dlt * dlt * nawc[c];
        }
      }
    };

    int it = 0;
    const bool trc = p.trace != nullptr && blockIdx.x == 0 && ew == 0 && lane == 0;
#define LNC_TR(k) do { if (trc && it < 4) p.trace[it * 8 + (k)] = clock64(); } while (0)
    for (int tile = cluster_id; tile < p.num_tiles; tile += n_clusters, ++it) {
      const int buf = it & 1;
      LNC_TR(0);
      bool row_ok[4];
      long long row_off[4];
#pragma unroll
      for (int ri = 0; ri < 4; ++ri) {
        const long long row = static_cast<long long>(tile) * kBlockM + row_of(ri);
        row_ok[ri] = row < p.M;
        row_off[ri] = row;
      }
      // the residual values of chunk c: 16 x LDG.64, [ri * 4 + k]
      auto load_x = [&](int c, float2 (&xr)[16]) {
#pragma unroll
        for (int ri = 0; ri < 4; ++ri) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const bool ok = row_ok[ri] && (FULL || 4 * c + k < natoms);
            xr[ri * 4 + k] = ok ? __ldcg(reinterpret_cast<const float2*>(xg + row_off[ri] * p.ldx + 32 * c + 8 * k))
                                : make_float2(0.f, 0.f);
          }
        }
      };
      float2 xr[16];
      if (nch > 0) load_x(0, xr);  // in flight while the accumulator is still being computed
      ptx::mbar_wait_a(acc_full + 8 * buf, (it >> 1) & 1);
      ptx::tc_fence_after();
      LNC_TR(1);
      if (p.dbg == 1) {
        ptx::tc_fence_before();
        ptx::mbar_arrive_a(acc_empty + 8 * buf);
        LNC_TR(7);
        continue;
      }
      const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + buf * kBN + ccol0;

      // ---- pass 1: v = x + alpha * acc + alpha * bias -> back into TMEM (and, without LN1, into the stream)
      float mean[4], m2[4], rstd[4];
#pragma unroll
      for (int ri = 0; ri < 4; ++ri) mean[ri] = 0.f, m2[ri] = 0.f, rstd[ri] = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (c < nch) {
          uint32_t v[32];
          tmem_ld_16x256b_x4(t_addr + 32 * c, v);
          tmem_ld_16x256b_x4(t_addr + (16u << 16) + 32 * c, v + 16);
          float2 xn[16];
          if (c + 1 < nch) load_x(c + 1, xn);
          float2 ab[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) ab[k] = lds64f(vec_thr + static_cast<uint32_t>(0 * kBN + 32 * c + 8 * k) * 4);
          ptx::tc_wait_ld();
          float val[32];
#pragma unroll
          for (int ri = 0; ri < 4; ++ri) {
            const int h2 = ri >> 1, j = ri & 1;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int r = 16 * h2 + 4 * k + 2 * j;
              // x + fma(alpha, acc, alpha * bias): the arithmetic of the reduce-add epilogue (epilogue.cuh, EPI_RESID)
              const bool ok = FULL || 4 * c + k < natoms;
              val[r] = ok ? xr[ri * 4 + k].x + fmaf(p.alpha, __uint_as_float(v[r]), ab[k].x) : 0.f;
              val[r + 1] = ok ? xr[ri * 4 + k].y + fmaf(p.alpha, __uint_as_float(v[r + 1]), ab[k].y) : 0.f;
              v[r] = __float_as_uint(val[r]);
              v[r + 1] = __float_as_uint(val[r + 1]);
              if (!DUAL && ok && row_ok[ri])
                __stcg(reinterpret_cast<float2*>(xg + row_off[ri] * p.ldx + 32 * c + 8 * k), make_float2(val[r], val[r + 1]));
            }
          }
          tmem_st_16x256b_x4(t_addr + 32 * c, v);
          tmem_st_16x256b_x4(t_addr + (16u << 16) + 32 * c, v + 16);
          chunk_stats(c, val, mean, m2);
#pragma unroll
          for (int i = 0; i < 16; ++i) xr[i] = xn[i];
        }
      }
      ptx::tc_wait_st();
      LNC_TR(2);
      if constexpr (DUAL) exchange(0, static_cast<uint32_t>(it & 1), mean, m2, rstd);
      else exchange(buf, static_cast<uint32_t>((it >> 1) & 1), mean, m2, rstd);
      LNC_TR(3);

      if constexpr (DUAL) {
        // ---- pass 1b: y = LN1(v) replaces v in TMEM and goes to the stream; statistics of y
        float mean2[4], m22[4];
#pragma unroll
        for (int ri = 0; ri < 4; ++ri) mean2[ri] = 0.f, m22[ri] = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (c < nch) {
            uint32_t v[32];
            tmem_ld_16x256b_x4(t_addr + 32 * c, v);
            tmem_ld_16x256b_x4(t_addr + (16u << 16) + 32 * c, v + 16);
            float2 gg[4], be[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              gg[k] = lds64f(vec_thr + static_cast<uint32_t>(1 * kBN + 32 * c + 8 * k) * 4);
              be[k] = lds64f(vec_thr + static_cast<uint32_t>(2 * kBN + 32 * c + 8 * k) * 4);
            }
            ptx::tc_wait_ld();
            float val[32];
#pragma unroll
            for (int ri = 0; ri < 4; ++ri) {
              const int h2 = ri >> 1, j = ri & 1;
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const int r = 16 * h2 + 4 * k + 2 * j;
                const bool ok = FULL || 4 * c + k < natoms;
                val[r] = ok ? fmaf((__uint_as_float(v[r]) - mean[ri]) * rstd[ri], gg[k].x, be[k].x) : 0.f;
                val[r + 1] = ok ? fmaf((__uint_as_float(v[r + 1]) - mean[ri]) * rstd[ri], gg[k].y, be[k].y) : 0.f;
                v[r] = __float_as_uint(val[r]);
                v[r + 1] = __float_as_uint(val[r + 1]);
                if (ok && row_ok[ri])
                  __stcg(reinterpret_cast<float2*>(xg + row_off[ri] * p.ldx + 32 * c + 8 * k), make_float2(val[r], val[r + 1]));
              }
            }
            tmem_st_16x256b_x4(t_addr + 32 * c, v);
            tmem_st_16x256b_x4(t_addr + (16u << 16) + 32 * c, v + 16);
            chunk_stats(c, val, mean2, m22);
          }
        }
        ptx::tc_wait_st();
        exchange(1, static_cast<uint32_t>(it & 1), mean2, m22, rstd);
#pragma unroll
        for (int ri = 0; ri < 4; ++ri) mean[ri] = mean2[ri];
      }
      LNC_TR(4);

      // ---- pass 2: the normalised bf16 operand of the next GEMM
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (c < nch) {
          uint32_t v[32];
          tmem_ld_16x256b_x4(t_addr + 32 * c, v);
          tmem_ld_16x256b_x4(t_addr + (16u << 16) + 32 * c, v + 16);
          float2 gg[4], be[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            gg[k] = lds64f(vec_thr + static_cast<uint32_t>(3 * kBN + 32 * c + 8 * k) * 4);
            be[k] = lds64f(vec_thr + static_cast<uint32_t>(4 * kBN + 32 * c + 8 * k) * 4);
          }
          ptx::tc_wait_ld();
#pragma unroll
          for (int ri = 0; ri < 4; ++ri) {
            const int h2 = ri >> 1, j = ri & 1;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int r = 16 * h2 + 4 * k + 2 * j;
              const bool ok = FULL || 4 * c + k < natoms;
              const float z0 = fmaf((__uint_as_float(v[r]) - mean[ri]) * rstd[ri], gg[k].x, be[k].x);
              const float z1 = fmaf((__uint_as_float(v[r + 1]) - mean[ri]) * rstd[ri], gg[k].y, be[k].y);
              if (ok && row_ok[ri])
                *reinterpret_cast<uint32_t*>(ag + row_off[ri] * p.ldo + 32 * c + 8 * k) = ptx::pack_bf16x2(z0, z1);
            }
          }
        }
      }
      LNC_TR(5);
      ptx::tc_fence_before();
      ptx::mbar_arrive_a(acc_empty + 8 * buf);
      LNC_TR(7);
    }
#undef LNC_TR
  }

  // neither CTA may retire while its partner can still write its exchange arrays or signal its barriers
  ptx::tc_fence_before();
  if constexpr (CL == 2) ptx::cluster_sync_all();
  else __syncthreads();
  if (warp == 0) {
    ptx::tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

template <int CL, bool DUAL, bool FULL>
int launch_instance(const CUtensorMap& tmA, const CUtensorMap& tmW, const LncParams& p, cudaStream_t st, std::string* err) {
  auto kern = gemm_lnc_kernel<CL, DUAL, FULL>;
  static bool configured[64] = {};
  static int max_clusters[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (CL == 2) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 2;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = kTotal;
  cfg.stream = st;
  cfg.attrs = attr;
  if (!configured[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kTotal);
    if (e != cudaSuccess) {
      if (err) *err = std::string("cudaFuncSetAttribute(gemm_lnc): ") + cudaGetErrorString(e);
      return static_cast<int>(e);
    }
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
    int clusters = sms / CL;
    if (CL == 2) {
      cfg.gridDim = dim3((sms / 2) * 2);
      cfg.numAttrs = na;
      int occ = 0;
      if (cudaOccupancyMaxActiveClusters(&occ, kern, &cfg) == cudaSuccess && occ > 0) clusters = occ < clusters ? occ : clusters;
      else cudaGetLastError();
    }
    max_clusters[dev & 63] = clusters < 1 ? 1 : clusters;
    configured[dev & 63] = true;
  }
  const int clusters = p.num_tiles < max_clusters[dev & 63] ? p.num_tiles : max_clusters[dev & 63];
  cfg.gridDim = dim3(CL * clusters);
  if (pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.numAttrs = na;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tmA, tmW, p);
  if (e != cudaSuccess) {
    if (err) *err = std::string("gemm_lnc launch: ") + cudaGetErrorString(e);
    return static_cast<int>(e);
  }
  return 0;
}

long long* g_lnc_trace = nullptr;
}  // namespace

bool gemm_lnc_supported(int M, int N, int K) { return M > 0 && N >= 8 && N <= 2 * kBN && N % 8 == 0 && K % 8 == 0; }

int launch_gemm_lnc(const GemmLnDesc& g, cudaStream_t st, std::string* err) {
  if (g.M <= 0) return 0;
  if (!gemm_lnc_supported(g.M, g.N, g.K) || (g.lda % 8) || (g.ldw % 8) || (g.ldx % 4) || (g.ldo % 8) || g.x == nullptr ||
      g.out_bf16 == nullptr || g.gamma2 == nullptr || g.beta2 == nullptr || ((g.gamma1 == nullptr) != (g.beta1 == nullptr))) {
    if (err) *err = "gemm_lnc: unsupported shape or missing argument (N <= 512, N % 8 == 0, K % 8 == 0, 16-byte row strides)";
    return -1;
  }
  LncParams p{};
  p.num_tiles = (g.M + kBlockM - 1) / kBlockM;
  p.num_k_blocks = (g.K + kBlockK - 1) / kBlockK;
  p.M = g.M;
  p.N = g.N;
  p.alpha = g.alpha;
  p.bias = g.bias;
  p.g1 = g.gamma1;
  p.b1 = g.beta1;
  p.g2 = g.gamma2;
  p.b2 = g.beta2;
  p.trace = nullptr;
  p.dbg = getenv("CFB_LNC_DEBUG") ? atoi(getenv("CFB_LNC_DEBUG")) : 0;
  if (getenv("CFB_LNC_TRACE")) {
    if (!g_lnc_trace) cudaMalloc(&g_lnc_trace, 64 * sizeof(long long));
    cudaMemsetAsync(g_lnc_trace, 0, 64 * sizeof(long long), st);
    p.trace = g_lnc_trace;
  }
  p.x = g.x;
  p.ldx = g.ldx;
  p.out = reinterpret_cast<bf16*>(g.out_bf16);
  p.ldo = g.ldo;
  CUtensorMap tmA, tmW;
  {
    uint64_t dims[2] = {static_cast<uint64_t>(g.K), static_cast<uint64_t>(g.M)};
    uint64_t strides[1] = {static_cast<uint64_t>(g.lda) * 2};
    uint32_t box[2] = {kBlockK, kBlockM};
    if (!encode_tmap_bf16(&tmA, g.A, 2, dims, strides, box, err)) return -1;
  }
  {
    uint64_t dims[2] = {static_cast<uint64_t>(g.K), static_cast<uint64_t>(g.N)};
    uint64_t strides[1] = {static_cast<uint64_t>(g.ldw) * 2};
    uint32_t box[2] = {kBlockK, kBN};
    if (!encode_tmap_bf16(&tmW, g.W, 2, dims, strides, box, err)) return -1;
  }
  const bool dual = g.gamma1 != nullptr;
  const int cl = g.N > kBN ? 2 : 1;
  const bool full = g.N == cl * kBN;
#define CFB_LNC_GO(CL_, D_, F_) return launch_instance<CL_, D_, F_>(tmA, tmW, p, st, err)
  if (cl == 2) {
    if (dual) { if (full) CFB_LNC_GO(2, true, true); else CFB_LNC_GO(2, true, false); }
    else { if (full) CFB_LNC_GO(2, false, true); else CFB_LNC_GO(2, false, false); }
  } else {
    if (dual) { if (full) CFB_LNC_GO(1, true, true); else CFB_LNC_GO(1, true, false); }
    else { if (full) CFB_LNC_GO(1, false, true); else CFB_LNC_GO(1, false, false); }
  }
#undef CFB_LNC_GO
}

long long* g_lnc_trace_view() { return g_lnc_trace; }

}  // namespace cfb

// debug: clock marks of the last traced launch (64 values)
extern "C" __attribute__((visibility("default"))) int cfb_debug_lnc_trace(long long* host_out) {
  if (!cfb::g_lnc_trace_view()) return 1;
  cudaDeviceSynchronize();
  return cudaMemcpy(host_out, cfb::g_lnc_trace_view(), 64 * sizeof(long long), cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : 2;
}
