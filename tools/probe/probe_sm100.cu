// Hardware probes behind the fused up-conv epilogue design: what a thread receives
// from tcgen05.ld.16x256b at lane offsets 0 / 16, how a strided (elementStrides) TMA load lays
// rows out under the 128-byte swizzle, what a swizzled TMA store with a 32-byte inner box reads
// from shared memory, and stmatrix.  Build:  tools/probe/build.sh ; run on the GPU box.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../rewriting_b200/csrc/rw_common.cuh"

using namespace rw;

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);
static PFN_encodeTiled encode_fn() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  return reinterpret_cast<PFN_encodeTiled>(fn);
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

// ------------------------------------------------------------------ probe 1: tcgen05.ld.16x256b
__global__ void probe_tmem(uint32_t* out) {
  __shared__ uint32_t tbase;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) tmem_alloc<32>(&tbase);
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n");
  const uint32_t base = tbase;
  // every lane L of quarter `warp` writes value (128-lane index)*100 + column, columns 0..7
  uint32_t v[8];
  for (int c = 0; c < 8; ++c) v[c] = (warp * 32 + lane) * 100 + c;
  const uint32_t a = base + (static_cast<uint32_t>(warp * 32) << 16);
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};\n" ::"r"(a), "r"(v[0]),
               "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]));
  asm volatile("tcgen05.wait::st.sync.aligned;\n");
  __syncthreads();
  for (int half = 0; half < 2; ++half) {
    uint32_t r[4];
    const uint32_t la = base + (static_cast<uint32_t>(warp * 32 + half * 16) << 16);
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(la));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n");
    for (int i = 0; i < 4; ++i) out[((warp * 2 + half) * 32 + lane) * 4 + i] = r[i];
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc<32>(base);
}

// ------------------------------------------------------------------ probe 2: strided TMA load
__global__ void probe_tma_load(const __grid_constant__ CUtensorMap map, uint32_t* out, int x0) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  for (int i = threadIdx.x; i < 2048 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0xdeadbeefu;
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bar, 1024);
    tma_load_2d(smem, &map, &bar, 0, x0);
  }
  {
    int spins = 0;
    while (!mbar_try_wait(&bar, 0) && ++spins < 200000) __nanosleep(100);
    if (threadIdx.x == 0) out[600] = spins;
  }
  for (int i = threadIdx.x; i < 2048 / 4; i += blockDim.x) out[i] = reinterpret_cast<uint32_t*>(smem)[i];
}

// ------------------------------------------------------------------ probe 3: swizzled TMA store, 32-byte rows
__global__ void probe_tma_store(const __grid_constant__ CUtensorMap map, int c0, int x0) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  // smem 16-byte chunk k (0..127) holds bf16 values k*8 .. k*8+7
  for (int i = threadIdx.x; i < 1024; i += blockDim.x)
    reinterpret_cast<__nv_bfloat16*>(smem)[i] = __float2bfloat16(static_cast<float>(i / 8 * 8 + (i & 7)));
  fence_proxy_async_smem();
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];\n" ::"l"(&map),
                 "r"(c0), "r"(x0), "r"(smem_u32(smem))
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory");
  }
}

// ------------------------------------------------------------------ probe 4: stmatrix x4
__global__ void probe_stmatrix(uint32_t* out) {
  __shared__ __align__(16) uint32_t sm[4 * 8 * 4];   // 4 matrices x 8 rows x 16 B
  const int lane = threadIdx.x;
  uint32_t r[4];
  for (int m = 0; m < 4; ++m) r[m] = (m << 16) | lane;     // matrix m, thread lane
  // lane i gives the address of row i%8 of matrix i/8; rows laid out [m][row] dense
  const uint32_t addr = smem_u32(sm) + ((lane >> 3) * 8 + (lane & 7)) * 16;
  asm volatile("stmatrix.sync.aligned.m8n8.x4.shared.b16 [%0], {%1,%2,%3,%4};\n" ::"r"(addr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3])
               : "memory");
  __syncwarp();
  for (int i = lane; i < 128; i += 32) out[i] = sm[i];
}

int main(int argc, char** argv) {
  setvbuf(stdout, nullptr, _IONBF, 0);
  const int which = argc > 1 ? atoi(argv[1]) : 0;
  PFN_encodeTiled enc = encode_fn();
  // ---- 1
  if (which == 0 || which == 1) {
    uint32_t* d;
    CK(cudaMalloc(&d, 4 * 2 * 32 * 4 * 4));
    probe_tmem<<<1, 128>>>(d);
    CK(cudaDeviceSynchronize());
    std::vector<uint32_t> h(4 * 2 * 32 * 4);
    CK(cudaMemcpy(h.data(), d, h.size() * 4, cudaMemcpyDeviceToHost));
    printf("== tcgen05.ld.16x256b.x1: value = lane*100+col\n");
    for (int w = 0; w < 2; ++w)
      for (int half = 0; half < 2; ++half) {
        printf("warp %d lane-offset %d:", w, half * 16);
        for (int t = 0; t < 10; ++t) {
          const uint32_t* r = &h[((w * 2 + half) * 32 + t) * 4];
          printf(" t%d[%u %u %u %u]", t, r[0], r[1], r[2], r[3]);
        }
        printf("\n");
      }
  }
  // ---- 2: tensor [X=64 rows][C=64 bf16]
  if (which == 0 || which == 2) {
    const int X = 64, C = 64;
    std::vector<__nv_bfloat16> h(X * C);
    for (int x = 0; x < X; ++x)
      for (int c = 0; c < C; ++c) h[x * C + c] = __float2bfloat16(static_cast<float>((c & 7) == 0 ? x : c / 8));
    __nv_bfloat16* d;
    CK(cudaMalloc(&d, h.size() * 2));
    CK(cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
    uint32_t* o;
    CK(cudaMalloc(&o, 4096));
    for (int variant = 0; variant < 2; ++variant) {
      CUtensorMap m;
      cuuint64_t gdim[2] = {(cuuint64_t)C, (cuuint64_t)X};
      cuuint64_t gstr[1] = {(cuuint64_t)C * 2};
      cuuint32_t box[2] = {64, variant == 0 ? 8u : 32u};
      cuuint32_t es[2] = {1, 4};
      CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      printf("== strided TMA load, box x-extent %u elementStride 4: encode -> %d\n", box[1], (int)r);
      if (r != CUDA_SUCCESS) continue;
      CK(cudaMemset(o, 0, 2048));
      probe_tma_load<<<1, 128, 8192>>>(m, o, 1);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("  launch failed: %s\n", cudaGetErrorString(e)); return 1; }
      std::vector<uint32_t> ho(1024);
      CK(cudaMemcpy(ho.data(), o, 4096, cudaMemcpyDeviceToHost));
      printf("  wait spins %u (200000 = the barrier never completed)\n", ho[600]);
      // print, per 128-byte smem row, the x of its first element after undoing the 128B swizzle
      for (int row = 0; row < 16; ++row) {
        printf("  smem row %2d:", row);
        for (int ch = 0; ch < 8; ++ch) {
          uint32_t w = ho[row * 32 + ch * 4];
          if (w == 0xdeadbeefu) { printf(" ----"); continue; }
          const int xv = (int)__bfloat162float(reinterpret_cast<__nv_bfloat16*>(&w)[0]);
          const int cv = (int)__bfloat162float(reinterpret_cast<__nv_bfloat16*>(&w)[1]);
          printf(" x%d:k%d", xv, cv);
        }
        printf("\n");
      }
    }
  }
  // ---- 3: global [X=64 px][C=128 bf16]; store box {16 ch, 32 px}, SWIZZLE_128B (and NONE)
  for (int sw = 0; sw < 3 && (which == 0 || which == 3); ++sw) {
    const int X = 64, C = 128;
    __nv_bfloat16* d;
    CK(cudaMalloc(&d, X * C * 2));
    CK(cudaMemset(d, 0xff, X * C * 2));
    CUtensorMap m;
    cuuint64_t gdim[2] = {(cuuint64_t)C, (cuuint64_t)X};
    cuuint64_t gstr[1] = {(cuuint64_t)C * 2};
    cuuint32_t box[2] = {16, 32};
    cuuint32_t es[2] = {1, 1};
    CUtensorMapSwizzle swz = sw == 0 ? CU_TENSOR_MAP_SWIZZLE_NONE : sw == 1 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_128B;
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     swz, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("== TMA store box {16ch, 32px}, swizzle mode %d: encode -> %d\n", sw, (int)r);
    if (r != CUDA_SUCCESS) continue;
    probe_tma_store<<<1, 128, 4096>>>(m, 32, 8);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("  launch failed: %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<__nv_bfloat16> h(X * C);
    CK(cudaMemcpy(h.data(), d, X * C * 2, cudaMemcpyDeviceToHost));
    // for px 8..23 print which smem 16-byte chunk landed in channels 32..39 and 40..47
    for (int px = 8; px < 24; ++px) {
      const int a = (int)__bfloat162float(h[px * C + 32]) / 8, b = (int)__bfloat162float(h[px * C + 40]) / 8;
      printf("  px %2d <- smem chunks %3d %3d   (dense would be %3d %3d)\n", px, a, b, (px - 8) * 2, (px - 8) * 2 + 1);
    }
    cudaFree(d);
  }
  // ---- 4
  if (which == 0 || which == 4) {
    uint32_t* d;
    CK(cudaMalloc(&d, 512));
    probe_stmatrix<<<1, 32>>>(d);
    CK(cudaDeviceSynchronize());
    uint32_t h[128];
    CK(cudaMemcpy(h, d, 512, cudaMemcpyDeviceToHost));
    printf("== stmatrix.x4: smem word [matrix][row][4 words] = (matrix<<16 | source lane)\n");
    for (int m = 0; m < 2; ++m)
      for (int row = 0; row < 8; ++row) {
        printf("  m%d row%d:", m, row);
        for (int w = 0; w < 4; ++w) printf(" m%u/lane%u", h[(m * 8 + row) * 4 + w] >> 16, h[(m * 8 + row) * 4 + w] & 0xffff);
        printf("\n");
      }
  }
  return 0;
}
