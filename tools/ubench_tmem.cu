// Micro-benchmark: cost of the pieces of the attention softmax warp's per-tile work on sm_100a, in isolation.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I conformer-nemo_b200/csrc tools/ubench_tmem.cu -o /tmp/ubench_tmem
// Prints cycles per iteration for: tcgen05.ld x32 streams (4 / 8 warps), the smem shift, ex2 chains.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#include "ptx.cuh"

using namespace cfb;

constexpr int kIters = 64;

// mode 0: NL tcgen05.ld.x32 per iteration, then wait::ld
// mode 1: 32 STS.128 + 64 LDS.32 per iteration (shift pattern, pitch 68)
// mode 2: 64 ex2 + 64 fadd per iteration
// mode 3: 64 cvt pack + 8 STS.128 (P store)
template <int MODE, int NL>
__global__ void __launch_bounds__(384, 1) ubench(long long* out, float* sink, int nwarps_active) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) ptx::tmem_alloc(&tmem_slot, 512);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tb = tmem_slot;
  float acc = 0.f;
  long long t0 = 0, t1 = 0;
  if (warp >= 4 && warp < 4 + nwarps_active) {
    const int quarter = warp & 3;
    const int set = (warp - 4) >> 2;
    const uint32_t tS = tb + (static_cast<uint32_t>(quarter * 32) << 16) + set * 256;
    const uint32_t shift_row = ptx::smem_u32(smem + (warp - 4) * (32 * 68 * 4)) + lane * 68 * 4;
    const int sh = 31 - lane;
    __syncwarp();
    t0 = clock64();
    for (int it = 0; it < kIters; ++it) {
      if constexpr (MODE == 0) {
        uint32_t v[NL][32];
#pragma unroll
        for (int l = 0; l < NL; ++l) ptx::tmem_ld_x32(tS + 32 * l, v[l]);
        ptx::tc_wait_ld();
#pragma unroll
        for (int l = 0; l < NL; ++l)
#pragma unroll
          for (int c = 0; c < 32; c += 8) acc += __uint_as_float(v[l][c]);
      } else if constexpr (MODE == 1) {
        uint32_t w[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) w[c] = __float_as_uint(acc + c);
#pragma unroll
        for (int rep = 0; rep < 2; ++rep) {
#pragma unroll
          for (int v4 = 0; v4 < 8; ++v4) {
            ptx::sts128(shift_row + v4 * 16, w[4 * v4], w[4 * v4 + 1], w[4 * v4 + 2], w[4 * v4 + 3]);
            ptx::sts128(shift_row + 128 + v4 * 16, w[4 * v4], w[4 * v4 + 1], w[4 * v4 + 2], w[4 * v4 + 3]);
          }
          float g[32];
          ptx::lds_f32x32(shift_row + sh * 4, g);
#pragma unroll
          for (int c = 0; c < 32; ++c) acc += g[c];
        }
      } else if constexpr (MODE == 2) {
        float s[4] = {0, 0, 0, 0};
#pragma unroll
        for (int c = 0; c < 64; ++c) {
          float y;
          asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(acc - c));
          s[c & 3] += y;
        }
        acc = (s[0] + s[1]) + (s[2] + s[3]);
      } else if constexpr (MODE == 3) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          uint32_t x = ptx::pack_bf16x2(acc + c, acc - c), y = ptx::pack_bf16x2(acc + 2 * c, acc - 2 * c);
          ptx::sts128(shift_row + ((c ^ (lane & 7)) << 4), x, y, x ^ 1, y ^ 1);
        }
        ptx::fence_proxy_async_smem();
        acc += 1.f;
      }
    }
    t1 = clock64();
  }
  if (lane == 0 && warp >= 4) out[blockIdx.x * 8 + (warp - 4)] = t1 - t0;
  sink[blockIdx.x * 384 + threadIdx.x] = acc;
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tb, 512);
  }
}

template <int MODE, int NL>
void run(const char* name, int nwarps) {
  long long* out;
  float* sink;
  cudaMalloc(&out, 148 * 8 * sizeof(long long));
  cudaMalloc(&sink, 148 * 384 * sizeof(float));
  cudaMemset(out, 0, 148 * 8 * sizeof(long long));
  const int smem = 8 * 32 * 68 * 4 + 2048;
  cudaFuncSetAttribute(ubench<MODE, NL>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  ubench<MODE, NL><<<148, 384, smem>>>(out, sink, nwarps);
  ubench<MODE, NL><<<148, 384, smem>>>(out, sink, nwarps);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[8];
  cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < nwarps; ++i) mx = h[i] > mx ? h[i] : mx;
  printf("%-28s warps=%d  %8.1f cyc/iter  (%s)\n", name, nwarps, double(mx) / kIters, cudaGetErrorString(e));
  cudaFree(out);
  cudaFree(sink);
}

int main() {
  for (int nw : {4, 8}) {
    run<0, 1>("tmem_ld x32 x1 + wait", nw);
    run<0, 2>("tmem_ld x32 x2 + wait", nw);
    run<0, 4>("tmem_ld x32 x4 + wait", nw);
    run<1, 0>("shift 32 STS128 + 64 LDS", nw);
    run<2, 0>("64 ex2 + sum", nw);
    run<3, 0>("P pack + 8 STS128 + fence", nw);
  }
  return 0;
}
