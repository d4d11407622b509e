// Probe: does tcgen05.mma kind::f16 accept bf16 operands with an fp16 accumulator (idesc c_format = 0), and how do fp16
// accumulators sit in tensor memory (what tcgen05.ld.32x32b returns with and without .pack::16b)?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I conformer-nemo_b200/csrc tools/ubench_f16acc.cu -o tools/ubench_f16acc.bin
// One CTA, D (128 x 64) = A (128 x 64) B^T (64 x 64), small exactly representable values; prints a few results of the fp32- and
// the fp16-accumulator forms beside the expected sums.
#include <cstdio>
#include <cstdlib>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "ptx.cuh"

using namespace cfb;

__host__ __device__ inline float aval(int r, int c) { return static_cast<float>(((r + c) % 7) - 3) * 0.5f; }
__host__ __device__ inline float bval(int n, int c) { return static_cast<float>(((n * 3 + c) % 5) - 2) * 0.25f; }

__device__ inline uint32_t sw128_off(int row, int col) {  // K-major 128-byte-swizzled tile, 64 bf16 per row
  return static_cast<uint32_t>((row >> 3) * 1024 + (row & 7) * 128 + ((((col >> 3) ^ (row & 7)) & 7) << 4) + (col & 7) * 2);
}

__global__ void __launch_bounds__(128, 1) probe(uint32_t* out_f32, uint32_t* out_f16_raw, uint32_t* out_f16_packed, int f16acc) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint32_t tmem_slot;
  __shared__ uint64_t bar;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* sa = smem;            // 128 x 64 bf16 = 16 KB
  uint8_t* sb = smem + 16384;    // 64 x 64 bf16 = 8 KB
  for (int i = threadIdx.x; i < 128 * 64; i += 128) {
    const int r = i / 64, c = i % 64;
    if (f16acc == 2) *reinterpret_cast<__half*>(sa + sw128_off(r, c)) = __float2half(aval(r, c));
    else *reinterpret_cast<__nv_bfloat16*>(sa + sw128_off(r, c)) = __float2bfloat16(aval(r, c));
  }
  for (int i = threadIdx.x; i < 64 * 64; i += 128) {
    const int n = i / 64, c = i % 64;
    if (f16acc == 2) *reinterpret_cast<__half*>(sb + sw128_off(n, c)) = __float2half(bval(n, c));
    else *reinterpret_cast<__nv_bfloat16*>(sb + sw128_off(n, c)) = __float2bfloat16(bval(n, c));
  }
  if (threadIdx.x == 0) {
    ptx::mbar_init(&bar, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 0) ptx::tmem_alloc(&tmem_slot, 64);
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tb = tmem_slot;
  if (threadIdx.x == 0) {
    // [4,6) c_format (0 = F16, 1 = F32)  [7,10) a_format  [10,13) b_format (0 = F16, 1 = BF16)  [17,23) N >> 3  [24,29) M >> 4
    const uint32_t idesc = f16acc == 2 ? ((64u >> 3) << 17) | ((128u >> 4) << 24)
                           : f16acc ? ptx::make_idesc_bf16_f16acc(128, 64, 0, 0) : ptx::make_idesc_bf16(128, 64, 0, 0);
    const uint64_t da = ptx::make_sdesc_sw128(ptx::smem_u32(sa), 16, 1024);
    const uint64_t db = ptx::make_sdesc_sw128(ptx::smem_u32(sb), 16, 1024);
    for (int k = 0; k < 4; ++k) ptx::umma_bf16(tb, da + 2 * k, db + 2 * k, idesc, k != 0);
    ptx::tc_commit(&bar);
  }
  ptx::mbar_wait(&bar, 0);
  ptx::tc_fence_after();
  const uint32_t t_lane = tb + (static_cast<uint32_t>(warp * 32) << 16);
  const int row = warp * 32 + lane;
  uint32_t v[32];
  for (int h = 0; h < 2; ++h) {
    ptx::tmem_ld_x32(t_lane + 32 * h, v);
    ptx::tc_wait_ld();
    for (int j = 0; j < 32; ++j) (f16acc ? out_f16_raw : out_f32)[row * 64 + 32 * h + j] = v[j];
  }
  if (f16acc) {
    uint32_t w[16];
    for (int h = 0; h < 2; ++h) {
      ptx::tmem_ld_x16_pack16(t_lane + 32 * h, w);
      ptx::tc_wait_ld();
      for (int j = 0; j < 16; ++j) out_f16_packed[row * 32 + 16 * h + j] = w[j];
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tb, 64);
}

int main() {
  uint32_t *d32, *draw, *dpk;
  cudaMalloc(&d32, 128 * 64 * 4);
  cudaMalloc(&draw, 128 * 64 * 4);
  cudaMalloc(&dpk, 128 * 32 * 4);
  cudaMemset(draw, 0xff, 128 * 64 * 4);
  cudaMemset(dpk, 0xff, 128 * 32 * 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
  probe<<<1, 128, 32768>>>(d32, draw, dpk, 0);
  cudaError_t e = cudaDeviceSynchronize();
  printf("fp32 accumulator run: %s\n", cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  const int mode = getenv("F16_OPERANDS") ? 2 : 1;
  probe<<<1, 128, 32768>>>(d32, draw, dpk, mode);
  e = cudaDeviceSynchronize();
  printf("fp16 accumulator run (%s operands): %s\n", mode == 2 ? "fp16" : "bf16", cudaGetErrorString(e));
  if (e != cudaSuccess) return 2;
  static uint32_t h32[128 * 64], hraw[128 * 64], hpk[128 * 32];
  cudaMemcpy(h32, d32, sizeof(h32), cudaMemcpyDeviceToHost);
  cudaMemcpy(hraw, draw, sizeof(hraw), cudaMemcpyDeviceToHost);
  cudaMemcpy(hpk, dpk, sizeof(hpk), cudaMemcpyDeviceToHost);
  int bad32 = 0, badraw = 0, badpk = 0;
  for (int r = 0; r < 128; ++r)
    for (int n = 0; n < 64; ++n) {
      float want = 0.f;
      for (int c = 0; c < 64; ++c) want += aval(r, c) * bval(n, c);
      float got32;
      memcpy(&got32, &h32[r * 64 + n], 4);
      bad32 += got32 != want;
      const __half_raw lo = {static_cast<unsigned short>(hraw[r * 64 + n] & 0xffff)};
      badraw += __half2float(__half(lo)) != want;
      const uint32_t word = hpk[r * 32 + n / 2];
      const __half_raw pk = {static_cast<unsigned short>((n & 1) ? (word >> 16) : (word & 0xffff))};
      badpk += __half2float(__half(pk)) != want;
      if (r == 5 && n < 6)
        printf("  D[5][%d]: expected %g | fp32 acc %g | fp16 acc raw column word 0x%08x (low half %g) | packed half %g\n", n, want,
               got32, hraw[r * 64 + n], __half2float(__half(lo)), __half2float(__half(pk)));
    }
  printf("mismatches: fp32 accumulator %d, fp16 accumulator read as one value per column %d, read with .pack::16b %d (of 8192)\n", bad32,
         badraw, badpk);
  return 0;
}
