// mma_rate.cu — how many cycles does one tcgen05.mma (bf16, M=128, K=16) take for N = 64/128/256
// when issued back to back from resident shared-memory tiles (no TMA in the loop)?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I rewriting_b200/csrc -o gpurun_out/mma_rate tools/cuda/mma_rate.cu
#include <cstdio>
#include "rw_common.cuh"
namespace rw { void set_last_error(const char*, ...) {} int check_cuda(cudaError_t, const char*) { return 0; } }
using namespace rw;

template <int N>
__global__ void __launch_bounds__(128, 1) rate_kernel(long long* out, int iters, int distinct) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tbase;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) tmem_alloc<512>(&tbase);
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
    const uint32_t sa = smem_u32(smem);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      // `distinct` different 16 KB A tiles and N*128 B tiles so that operands are not trivially cached
      const uint32_t a_off = (i % distinct) * 16384;
      const uint32_t b_off = 65536 + (i % distinct) * (N * 128);
      const uint64_t da = make_smem_desc(sa + a_off, 16, 1024, kSwizzle128B);
      const uint64_t db = make_smem_desc(sa + b_off, 16, 1024, kSwizzle128B);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) umma_bf16(tbase, da + 2 * kk, db + 2 * kk, idesc, 1u);
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc<512>(tbase); }
}

template <int N> void run(int grid) {
  long long* d; cudaMalloc(&d, grid * sizeof(long long));
  const int smem = 200 * 1024;
  cudaFuncSetAttribute(rate_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 4096;
  for (int distinct : {1, 2}) {
    rate_kernel<N><<<grid, 128, smem>>>(d, iters, distinct);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[256]; cudaMemcpy(h, d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
    long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("N=%d grid=%d distinct=%d: %s  %.1f cycles per MMA (M128,K16)  -> %.0f%% of the 128*N/256 floor\n", N, grid,
           distinct, cudaGetErrorString(e), (double)mx / (iters * 4), 100.0 * (128.0 * N / 256.0) / ((double)mx / (iters * 4)));
  }
  cudaFree(d);
}
int main() {
  for (int grid : {1, 148}) { run<64>(grid); run<128>(grid); run<144>(grid); run<160>(grid); run<192>(grid); run<224>(grid); run<256>(grid); }
  return 0;
}
