// upconv_tc.cu — the upsampling StyledConv of the generation fast path as ONE kernel:
//
//   conv_transpose2d(k, scale*W, stride 2)  ->  blur 4x4 (pad 1,1)  ->  * demod  ->
//   + noise_w * noise + bias  ->  leaky-ReLU * sqrt(2)  ->  * next style  ->  bf16 hi/lo planes
//
// (reference chain: DemodulatedConv2dF models.py:313-329 -> BlurF :275-281 / upfirdn2d_kernel.cu
//  :52-137 -> NoiseInjectionF :535-546 -> FusedLeakyReLUF, fused_bias_act_kernel.cu:27-47 ->
//  the next layer's ApplyStyle :616-620).  The fp32 conv_transpose output `t` (1.04 GB at layer
//  13, batch 32) is never written: round 1 wrote it channels-last and read it back in a separate
//  SIMT blur kernel (2.3x the algorithmic DRAM traffic of the layer pair).
//
// Formulation ("scatter" polyphase).  For an INPUT pixel p and tap (u, v) let
//     P_uv[p, o] = sum_i k[p, i] * (scale*W)[o, i, u, v]
// (no shifted operands at all: one A tile serves all nine taps).  The conv_transpose output is
//     t[2y+u, 2x+v] += P_uv[(y, x)]
// so a GEMM tile is M = 128 input pixels x N = 144 = 9 taps x 16 output channels, K = Cin:
// tcgen05.mma M128 N144 K16 runs at 76.7 cycles (94 % of the 128*N/256 floor; N = 128 tiles
// reach 85 %, profiles/r2_mma_rate.txt), and every accumulation chain is only Cin/16 * 3 long
// (the 3-term bf16 split), so no chunk promotion is needed against the tensor core's truncating
// fp32 accumulate (see conv_tc.cu).
//
// A tile is ONE image row segment per image: lane l <-> (image l / W, x = l % W), W <= 128 a
// power of two, 128 / W images per tile.  A CTA marches down the rows of its images; the 4x4 FIR
// needs t rows 2y-3 .. 2y+1 to emit output rows 2y-2, 2y-1 after step y, all of which depend on
// P at rows <= y of the SAME lane (vertical direction) and of the two neighbouring lanes
// (horizontal direction): vertical state lives in registers (three horizontally filtered rows +
// the u = 2 taps of the previous row), horizontal neighbours come from warp shuffles (and a
// 2 KB shared-memory mailbox across warp boundaries).  Nothing is recomputed except two warm-up
// rows per row band.
//
// Warp roles (384 threads = 3 warpgroups): warps 0..7 = epilogue — lane quarter q = warp % 4
// (hardware restriction of tcgen05.ld), channel half h = warp / 4 (8 of the tile's 16 output
// channels); warp 8 = TMA producer, warp 9 = MMA issuer (+TMEM alloc), warps 10-11 idle.  The
// third warpgroup gives its registers back (setmaxnreg.dec 40) so that the epilogue warps can
// hold their ~200 live values (vertical window, carried taps, per-channel constants) without
// spilling (setmaxnreg.inc 232): ncu of the 320-thread version showed the step time set by
// long-scoreboard stalls on spill reloads with only two epilogue warps per scheduler.
#include <cstring>

#include "rw_common.cuh"
#include "rw_kernels.h"

namespace rw {

namespace {

constexpr int UM = 128;                 // input pixels per tile
constexpr int UNC = 16;                 // output channels per tile
constexpr int UN = 9 * UNC;             // GEMM N = 144
constexpr int UBK = 64;
constexpr int UK = 16;
constexpr int kUStages = 3;
constexpr int kUThreads = 384;
constexpr int kUTmaWarp = 8, kUMmaWarp = 9;
constexpr int kUABytes = UM * UBK * 2;  // one plane of A: 16 KB
constexpr int kUBBytes = UN * UBK * 2;  // one plane of B: 18 KB
constexpr int kUStageBytes = 2 * kUABytes + 2 * kUBBytes;    // 68 KB
constexpr int kUAccStride = 256;        // TMEM columns between the two accumulators
constexpr int kUMailFloats = 2 * 2 * 4 * 2 * 32;             // [buf][half][quarter][side][32]
constexpr int kUPairBytes = 4 * 2 * 4 * 32 * 16;               // [quarter][sender half][piece][lane][16 B]
constexpr int kUSmemTotal = kUStages * kUStageBytes + kUMailFloats * 4 + kUPairBytes + 1024 + 256;

struct UBarriers {
  uint64_t full[kUStages];
  uint64_t empty[kUStages];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
};

// 32 lanes x 8 consecutive fp32 columns -> 8 registers per thread
__device__ __forceinline__ void tmem_ld_32x8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar,
                                            int32_t c0, int32_t c1, int32_t c2, int32_t c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];\n" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

template <int N>
__device__ __forceinline__ void reg_alloc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(N));
}
template <int N>
__device__ __forceinline__ void reg_dealloc() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(N));
}
// producer-side wait: back off between polls so the spin does not take issue slots from the
// epilogue warps that share the scheduler
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    __nanosleep(40);
    if (++spins > RW_SPIN_LIMIT) __trap();
  }
}

// vector accesses to the mailbox by 32-bit shared-window address: an edge lane moves its 32
// values with 8 x 128-bit instructions.  (As 64 scalar loads through the generic `mail` pointer
// inside a one-lane divergent block, this fix-up took 2 980 of a step's 7 800 cycles:
// tools/prof_upconv.py.)
__device__ __forceinline__ void sts_v4(uint32_t a, float x, float y, float z, float w) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"r"(a), "f"(x), "f"(y), "f"(z), "f"(w)
               : "memory");
}
__device__ __forceinline__ void sts_v4u(uint32_t a, const uint32_t (&w)[4]) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(a), "r"(w[0]), "r"(w[1]), "r"(w[2]),
               "r"(w[3])
               : "memory");
}
__device__ __forceinline__ void lds_v4u(uint32_t a, uint32_t (&w)[4]) {
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];\n"
               : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3])
               : "r"(a)
               : "memory");
}
// one full 32-byte sector per thread (sm_100: 256-bit global store)
__device__ __forceinline__ void stg_256(void* dst, const uint32_t (&a)[4], const uint32_t (&b)[4]) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"l"(dst), "r"(a[0]),
               "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3])
               : "memory");
}
__device__ __forceinline__ void lds_v4(uint32_t a, float& x, float& y, float& z, float& w) {
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];\n"
               : "=f"(x), "=f"(y), "=f"(z), "=f"(w)
               : "r"(a)
               : "memory");
}

// waits of the two control warps: back off between polls — they share their schedulers with
// epilogue warps, and ncu showed ~16 % of the kernel's issued instructions in their spin loops
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity, uint32_t ns) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    __nanosleep(ns);
    if (++spins > RW_SPIN_LIMIT) __trap();
  }
}

__device__ __forceinline__ void named_bar_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(count) : "memory");
}

struct UpItem {
  int bg, band, cg;
  int y_first, y_emit, y_end;   // steps y_first .. y_end-1; rows are emitted for y >= y_emit
};

__device__ __forceinline__ UpItem decode_item(int item, const UpFusedParams& p) {
  UpItem it;
  it.cg = item % p.ncg;
  const int r = item / p.ncg;
  it.band = r % p.nbands;
  it.bg = r / p.nbands;
  // bands partition the emitting steps [1, H + 1)
  const int ya = 1 + (it.band * p.H) / p.nbands;
  const int yb = 1 + ((it.band + 1) * p.H) / p.nbands;
  it.y_emit = ya;
  it.y_first = ya - 2 < 0 ? 0 : ya - 2;
  it.y_end = yb;
  return it;
}

// PROF = true: bring-up variant that accumulates, per epilogue warp, the cycles spent in each phase
// of a step (tools/debug_upconv.py prof) into p.debug_prof; the product launches PROF = false.
template <bool PROF>
__global__ void __launch_bounds__(kUThreads, 1)
upconv_fused_kernel(const __grid_constant__ CUtensorMap map_a_hi,
                    const __grid_constant__ CUtensorMap map_a_lo,
                    const __grid_constant__ CUtensorMap map_w_hi,
                    const __grid_constant__ CUtensorMap map_w_lo, const UpFusedParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  float* mail = reinterpret_cast<float*>(smem + kUStages * kUStageBytes);
  uint8_t* pair_area = smem + kUStages * kUStageBytes + kUMailFloats * 4;
  UBarriers* bars = reinterpret_cast<UBarriers*>(pair_area + kUPairBytes);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int kb_count = p.Cin / UBK;
  const int G = UM / p.W;                 // images per tile

  if (warp == kUTmaWarp && lane == 0) {
    tma_prefetch_desc(&map_a_hi);
    tma_prefetch_desc(&map_a_lo);
    tma_prefetch_desc(&map_w_hi);
    tma_prefetch_desc(&map_w_lo);
    for (int s = 0; s < kUStages; ++s) {
      mbar_init(&bars->full[s], 1);
      mbar_init(&bars->empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bars->tmem_full[s], 1);
      mbar_init(&bars->tmem_empty[s], 8);
    }
    fence_mbar_init();
  }
  if (warp == kUMmaWarp) tmem_alloc<512>(&bars->tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp >= 8) {
    reg_dealloc<40>();
  if (warp == kUTmaWarp) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < p.nitems; item += gridDim.x) {
        const UpItem it = decode_item(item, p);
        const int b0 = it.bg * G;
        const int wrow = it.cg * UN;
        for (int y = it.y_first; y < it.y_end; ++y) {
          for (int kb = 0; kb < kb_count; ++kb) {
            mbar_wait_relaxed(&bars->empty[stage], phase ^ 1u);
            uint8_t* st = smem + stage * kUStageBytes;
            mbar_expect_tx(&bars->full[stage], kUStageBytes);
            tma_load_4d(st, &map_a_hi, &bars->full[stage], kb * UBK, 0, y, b0);
            tma_load_4d(st + kUABytes, &map_a_lo, &bars->full[stage], kb * UBK, 0, y, b0);
            tma_load_2d(st + 2 * kUABytes, &map_w_hi, &bars->full[stage], kb * UBK, wrow);
            tma_load_2d(st + 2 * kUABytes + kUBBytes, &map_w_lo, &bars->full[stage], kb * UBK,
                        wrow);
            if (++stage == kUStages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == kUMmaWarp) {
    // ------------------------------ MMA issuer --------------------------------
    constexpr uint32_t idesc = make_idesc_bf16(UM, UN, 0, 0);
    int stage = 0;
    uint32_t phase = 0;
    uint32_t step = 0;
    for (int item = blockIdx.x; item < p.nitems; item += gridDim.x) {
      const UpItem it = decode_item(item, p);
      for (int y = it.y_first; y < it.y_end; ++y, ++step) {
        const int as = step & 1u;
        const uint32_t aphase = (step >> 1) & 1u;
        mbar_wait_relaxed(&bars->tmem_empty[as], aphase ^ 1u, 32);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * kUAccStride;
        for (int kb = 0; kb < kb_count; ++kb) {
          mbar_wait(&bars->full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * kUStageBytes);
          const uint64_t da_hi = make_smem_desc(sa, 16, 1024, kSwizzle128B);
          const uint64_t da_lo = make_smem_desc(sa + kUABytes, 16, 1024, kSwizzle128B);
          const uint64_t db_hi = make_smem_desc(sa + 2 * kUABytes, 16, 1024, kSwizzle128B);
          const uint64_t db_lo = make_smem_desc(sa + 2 * kUABytes + kUBBytes, 16, 1024, kSwizzle128B);
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < UBK / UK; ++kk) {
              const uint64_t adv = static_cast<uint64_t>((kk * UK * 2) >> 4);
              umma_bf16(tmem_d, da_lo + adv, db_hi + adv, idesc, (kb | kk) != 0);
              umma_bf16(tmem_d, da_hi + adv, db_lo + adv, idesc, 1u);
              umma_bf16(tmem_d, da_hi + adv, db_hi + adv, idesc, 1u);
            }
            umma_commit(&bars->empty[stage]);
            if (kb + 1 == kb_count) umma_commit(&bars->tmem_full[as]);
          }
          __syncwarp();
          if (++stage == kUStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  }
  } else {
    // ------------------------------ epilogue ----------------------------------
    reg_alloc<232>();
    const int q = warp & 3;
    const int h = warp >> 2;
    const int l = q * 32 + lane;               // tile lane = TMEM lane
    const int W = p.W;
    const int x = l & (W - 1);
    const int img_in_tile = l / W;
    const int Ho = 2 * p.H, Wo = 2 * W;
    const bool first_x = (x == 0), last_x = (x == W - 1);
    const bool cross = W > 32;                 // x-neighbours can live in another warp
    // flipped blur kernel (upfirdn2d correlates with the flipped kernel), rank one:
    //   kf[a][b] = kv[a] * kh[b],  kv[a] = kf[a][0],  kh[b] = kf[0][b] / kf[0][0]
    float kv[4], kh[4];
    {
      const float k00 = __ldg(p.k4 + 15);
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        kv[a] = __ldg(p.k4 + 15 - a * 4);
        kh[a] = __ldg(p.k4 + 15 - a) / k00;
      }
    }
    const float nw = __ldg(p.noise_w);
    uint32_t step = 0;
    // mailbox slots of this warp, as shared-window byte addresses (buffer 0; +kMailBufBytes for
    // odd steps): slot (h, q, side) = 32 floats [row E/O][ro|le, od][8 channels]
    constexpr uint32_t kMailBufBytes = 2 * 4 * 2 * 32 * 4;
    const uint32_t mail_s = smem_u32(mail) + h * (4 * 2 * 32 * 4);
    const uint32_t post_r = mail_s + (q * 2 + 1) * 128, post_l = mail_s + (q * 2 + 0) * 128;
    const uint32_t read_l = mail_s + ((q - 1) * 2 + 1) * 128, read_r = mail_s + ((q + 1) * 2 + 0) * 128;
    // partner exchange (the two warps of a lane quarter own 8 channels each = half of every
    // 32-byte sector of the output planes): half 0 stores the `hi` plane, half 1 the `lo` plane;
    // each sends the 16-byte pieces of the OTHER plane to its partner through shared memory and
    // writes whole sectors with 256-bit stores.  [piece][lane][16 B] keeps the STS/LDS conflict-free.
    const uint32_t pair_s = smem_u32(pair_area) + q * (2 * 4 * 32 * 16);
    const uint32_t pair_send = pair_s + h * (4 * 32 * 16) + lane * 16;
    const uint32_t pair_recv = pair_s + (h ^ 1) * (4 * 32 * 16) + lane * 16;
    long long prof_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // wait, tmem, combine+post+barrier, shuffles, fixups, hf, emit, steps

    for (int item = blockIdx.x; item < p.nitems; item += gridDim.x) {
      const UpItem it = decode_item(item, p);
      const int b = it.bg * G + img_in_tile;
      const bool img_ok = b < p.B;
      const int c0 = it.cg * UNC + h * 8;      // first of this thread's 8 output channels
      float dm[8], bs[8], ns[8];
      {
        const size_t o = static_cast<size_t>(img_ok ? b : 0) * p.Cout + c0;
#pragma unroll
        for (int j4 = 0; j4 < 2; ++j4) {
          const float4 d4 = __ldg(reinterpret_cast<const float4*>(p.demod + o) + j4);
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + c0) + j4);
          const float4 n4 = __ldg(reinterpret_cast<const float4*>(p.next_scale + o) + j4);
          dm[4 * j4] = d4.x; dm[4 * j4 + 1] = d4.y; dm[4 * j4 + 2] = d4.z; dm[4 * j4 + 3] = d4.w;
          bs[4 * j4] = b4.x; bs[4 * j4 + 1] = b4.y; bs[4 * j4 + 2] = b4.z; bs[4 * j4 + 3] = b4.w;
          ns[4 * j4] = n4.x; ns[4 * j4 + 1] = n4.y; ns[4 * j4 + 2] = n4.z; ns[4 * j4 + 3] = n4.w;
        }
      }
      // vertical state: u = 2 taps of the previous input row, and the horizontally filtered
      // t rows 2y-3 (w0), 2y-2 (w1), 2y-1 (w2); each [2 output columns][8 channels]
      float c20[8], c21[8], c22[8];
      float w0[2][8], w1[2][8], w2[2][8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        c20[j] = c21[j] = c22[j] = 0.f;
        w0[0][j] = w0[1][j] = w1[0][j] = w1[1][j] = w2[0][j] = w2[1][j] = 0.f;
      }
      const float* nrow_base = p.noise + static_cast<size_t>(img_ok ? b : 0) * p.noise_bstride + 2 * x;
      __nv_bfloat16* out_hi = static_cast<__nv_bfloat16*>(p.next_hi);
      __nv_bfloat16* out_lo = static_cast<__nv_bfloat16*>(p.next_lo);
      const size_t img_row0 = static_cast<size_t>(img_ok ? b : 0) * (Ho + 1);

      for (int y = it.y_first; y < it.y_end; ++y, ++step) {
        const bool emit = (y >= it.y_emit) && img_ok;
        // noise of the two output rows of this step (issued before the TMEM wait)
        float2 nz0 = make_float2(0.f, 0.f), nz1 = make_float2(0.f, 0.f);
        if (emit) {
          nz0 = __ldg(reinterpret_cast<const float2*>(nrow_base + static_cast<size_t>(2 * y - 2) * Wo));
          nz1 = __ldg(reinterpret_cast<const float2*>(nrow_base + static_cast<size_t>(2 * y - 1) * Wo));
        }
        long long tq[8];
        if constexpr (PROF) tq[0] = clock64();
        const int as = step & 1u;
        const uint32_t aphase = (step >> 1) & 1u;
        mbar_wait(&bars->tmem_full[as], aphase);
        tc_fence_after();
        if constexpr (PROF) tq[1] = clock64();
        const uint32_t tcol = tmem_base + static_cast<uint32_t>(as * kUAccStride + h * 8) +
                              (static_cast<uint32_t>(q * 32) << 16);
        float P[9][8];
#pragma unroll
        for (int t = 0; t < 9; ++t) tmem_ld_32x8(tcol + t * UNC, P[t]);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->tmem_empty[as]);
        if constexpr (PROF) tq[2] = clock64();
        if (p.debug_p != nullptr && img_ok && y < p.H && y >= it.y_emit - 1) {
          float* dp = p.debug_p + ((static_cast<size_t>(b) * p.H + y) * W + x) * 9 * p.Cout + c0;
#pragma unroll
          for (int t = 0; t < 9; ++t)
#pragma unroll
            for (int j = 0; j < 8; ++j) dp[t * p.Cout + j] = P[t][j];
        }

        // t rows E = 2y, O = 2y+1 in lane-local pieces (see the header comment):
        //   E.e[x] = leE[x] + rE[x-1], E.o[x] = oE[x];   O.e[x] = leO[x] + rO[x-1], O.o[x] = oO[x]
        float le[2][8], ro[2][8], od[2][8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          le[0][j] = P[0][j] + c20[j];
          od[0][j] = P[1][j] + c21[j];
          ro[0][j] = P[2][j] + c22[j];
          le[1][j] = P[3][j];
          od[1][j] = P[4][j];
          ro[1][j] = P[5][j];
          c20[j] = P[6][j];
          c21[j] = P[7][j];
          c22[j] = P[8][j];
        }
        // mailbox across warp boundaries: lane 31 posts (ro, od) for its right neighbour,
        // lane 0 posts (le, od) for its left neighbour
        const uint32_t mbo = (step & 1u) * kMailBufBytes;
        if (cross) {
          if (lane == 31) {
#pragma unroll
            for (int r = 0; r < 2; ++r) {
              sts_v4(post_r + mbo + r * 64, ro[r][0], ro[r][1], ro[r][2], ro[r][3]);
              sts_v4(post_r + mbo + r * 64 + 16, ro[r][4], ro[r][5], ro[r][6], ro[r][7]);
              sts_v4(post_r + mbo + r * 64 + 32, od[r][0], od[r][1], od[r][2], od[r][3]);
              sts_v4(post_r + mbo + r * 64 + 48, od[r][4], od[r][5], od[r][6], od[r][7]);
            }
          }
          if (lane == 0) {
#pragma unroll
            for (int r = 0; r < 2; ++r) {
              sts_v4(post_l + mbo + r * 64, le[r][0], le[r][1], le[r][2], le[r][3]);
              sts_v4(post_l + mbo + r * 64 + 16, le[r][4], le[r][5], le[r][6], le[r][7]);
              sts_v4(post_l + mbo + r * 64 + 32, od[r][0], od[r][1], od[r][2], od[r][3]);
              sts_v4(post_l + mbo + r * 64 + 48, od[r][4], od[r][5], od[r][6], od[r][7]);
            }
          }
          named_bar_sync(1 + h, 128);
        }
        if constexpr (PROF) tq[3] = clock64();
        // neighbour pieces: left (ro, od) of lane x-1, right (le, od) of lane x+1
        float l_ro[2][8], l_od[2][8], r_le[2][8], r_od[2][8];
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            l_ro[r][j] = __shfl_up_sync(0xffffffffu, ro[r][j], 1);
            l_od[r][j] = __shfl_up_sync(0xffffffffu, od[r][j], 1);
            r_le[r][j] = __shfl_down_sync(0xffffffffu, le[r][j], 1);
            r_od[r][j] = __shfl_down_sync(0xffffffffu, od[r][j], 1);
          }
        if constexpr (PROF) tq[6] = clock64();
        if (first_x) {                            // image edge: nothing to the left
#pragma unroll
          for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int j = 0; j < 8; ++j) l_ro[r][j] = l_od[r][j] = 0.f;
        } else if (cross && lane == 0) {          // left neighbour lives in the previous warp
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            lds_v4(read_l + mbo + r * 64, l_ro[r][0], l_ro[r][1], l_ro[r][2], l_ro[r][3]);
            lds_v4(read_l + mbo + r * 64 + 16, l_ro[r][4], l_ro[r][5], l_ro[r][6], l_ro[r][7]);
            lds_v4(read_l + mbo + r * 64 + 32, l_od[r][0], l_od[r][1], l_od[r][2], l_od[r][3]);
            lds_v4(read_l + mbo + r * 64 + 48, l_od[r][4], l_od[r][5], l_od[r][6], l_od[r][7]);
          }
        }
        if (last_x) {
#pragma unroll
          for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int j = 0; j < 8; ++j) r_le[r][j] = r_od[r][j] = 0.f;
        } else if (cross && lane == 31) {
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            lds_v4(read_r + mbo + r * 64, r_le[r][0], r_le[r][1], r_le[r][2], r_le[r][3]);
            lds_v4(read_r + mbo + r * 64 + 16, r_le[r][4], r_le[r][5], r_le[r][6], r_le[r][7]);
            lds_v4(read_r + mbo + r * 64 + 32, r_od[r][0], r_od[r][1], r_od[r][2], r_od[r][3]);
            lds_v4(read_r + mbo + r * 64 + 48, r_od[r][4], r_od[r][5], r_od[r][6], r_od[r][7]);
          }
        }
        if constexpr (PROF) tq[7] = clock64();
        float hf[2][2][8];                       // [t row E/O][output column 2x / 2x+1][channel]
#pragma unroll
        for (int r = 0; r < 2; ++r) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float e0 = le[r][j] + l_ro[r][j];       // t col 2x
            const float e1 = r_le[r][j] + ro[r][j];       // t col 2x+2
            const float o0 = od[r][j];                    // t col 2x+1
            hf[r][0][j] = fmaf(kh[3], e1, fmaf(kh[2], o0, fmaf(kh[1], e0, kh[0] * l_od[r][j])));
            hf[r][1][j] = fmaf(kh[3], r_od[r][j], fmaf(kh[2], e1, fmaf(kh[1], o0, kh[0] * e0)));
          }
        }
        if constexpr (PROF) tq[4] = clock64();
        if (y >= it.y_emit) {                      // warp- and pair-uniform
          // output rows Y0 = 2y-2 (t rows 2y-3..2y) and Y1 = 2y-1 (t rows 2y-2..2y+1)
          uint32_t keep[4][4];                     // [yi*2+xi]: this half's 8 channels of ITS plane
#pragma unroll
          for (int yi = 0; yi < 2; ++yi) {
            const float2 nz = yi == 0 ? nz0 : nz1;
#pragma unroll
            for (int xi = 0; xi < 2; ++xi) {
              uint32_t hw[4], lw[4];
              const float nzv = nw * (xi == 0 ? nz.x : nz.y);
#pragma unroll
              for (int j2 = 0; j2 < 4; ++j2) {
                float kk[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                  const int j = 2 * j2 + e;
                  float v;
                  if (yi == 0)
                    v = fmaf(kv[3], hf[0][xi][j],
                             fmaf(kv[2], w2[xi][j], fmaf(kv[1], w1[xi][j], kv[0] * w0[xi][j])));
                  else
                    v = fmaf(kv[3], hf[1][xi][j],
                             fmaf(kv[2], hf[0][xi][j], fmaf(kv[1], w2[xi][j], kv[0] * w1[xi][j])));
                  v = (v * dm[j] + nzv) + bs[j];
                  v = fmaxf(v, 0.2f * v) * 1.4142135623730951f;
                  kk[e] = ns[j] * v;
                }
                const __nv_bfloat162 hh = __floats2bfloat162_rn(kk[0], kk[1]);
                const uint32_t hu = *reinterpret_cast<const uint32_t*>(&hh);
                const __nv_bfloat162 ll = __floats2bfloat162_rn(
                    kk[0] - __uint_as_float(hu << 16), kk[1] - __uint_as_float(hu & 0xffff0000u));
                hw[j2] = hu;
                lw[j2] = *reinterpret_cast<const uint32_t*>(&ll);
              }
              const int pc = yi * 2 + xi;
              if (h == 0) {
                sts_v4u(pair_send + pc * (32 * 16), lw);
#pragma unroll
                for (int i = 0; i < 4; ++i) keep[pc][i] = hw[i];
              } else {
                sts_v4u(pair_send + pc * (32 * 16), hw);
#pragma unroll
                for (int i = 0; i < 4; ++i) keep[pc][i] = lw[i];
              }
            }
          }
          named_bar_sync(3 + q, 64);               // both halves of this lane quarter have posted
          __nv_bfloat16* plane = (h == 0) ? out_hi : out_lo;
#pragma unroll
          for (int pc = 0; pc < 4; ++pc) {
            uint32_t other[4];
            lds_v4u(pair_recv + pc * (32 * 16), other);
            const int Y = 2 * y - 2 + (pc >> 1);
            const size_t off = ((img_row0 + Y) * (Wo + 1) + 2 * x + (pc & 1)) * p.Cout + it.cg * UNC;
            if (img_ok && (!PROF || p.debug_nostore == 0 || keep[pc][0] == 0x12345678u)) {
              if (h == 0) stg_256(plane + off, keep[pc], other);     // channels 0-7 | 8-15
              else stg_256(plane + off, other, keep[pc]);
            }
          }
          named_bar_sync(3 + q, 64);               // the exchange area may be overwritten
          if (img_ok && last_x) {                  // zero pad column of the output grid
#pragma unroll
            for (int yi = 0; yi < 2; ++yi) {
              const size_t off = ((img_row0 + 2 * y - 2 + yi) * (Wo + 1) + Wo) * p.Cout + c0;
              *reinterpret_cast<uint4*>(out_hi + off) = make_uint4(0u, 0u, 0u, 0u);
              *reinterpret_cast<uint4*>(out_lo + off) = make_uint4(0u, 0u, 0u, 0u);
            }
          }
        }
        if (emit) {
          if (y == p.H) {                          // last step of the image: zero pad row
            const size_t prow = (img_row0 + Ho) * (Wo + 1);
#pragma unroll
            for (int xi = 0; xi < 3; ++xi) {
              if (xi < 2 || last_x) {
                const size_t off = (prow + 2 * x + xi) * p.Cout + c0;
                *reinterpret_cast<uint4*>(out_hi + off) = make_uint4(0u, 0u, 0u, 0u);
                *reinterpret_cast<uint4*>(out_lo + off) = make_uint4(0u, 0u, 0u, 0u);
              }
            }
          }
        }
        if constexpr (PROF) {
          tq[5] = clock64();
          prof_acc[0] += tq[1] - tq[0];
          prof_acc[1] += tq[2] - tq[1];
          prof_acc[2] += tq[3] - tq[2];
          prof_acc[3] += tq[6] - tq[3];
          prof_acc[4] += tq[7] - tq[6];
          prof_acc[5] += tq[4] - tq[7];
          prof_acc[6] += tq[5] - tq[4];
          prof_acc[7] += 1;
        }
        // slide the vertical window: rows 2y-1, 2y, 2y+1 become 2(y+1)-3 .. 2(y+1)-1
#pragma unroll
        for (int xi = 0; xi < 2; ++xi)
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            w0[xi][j] = w2[xi][j];
            w1[xi][j] = hf[0][xi][j];
            w2[xi][j] = hf[1][xi][j];
          }
      }
    }
    if constexpr (PROF) {
      if (lane == 0 && p.debug_prof != nullptr) {
        long long* dst = p.debug_prof + (static_cast<size_t>(blockIdx.x) * 8 + warp) * 8;
#pragma unroll
        for (int i = 0; i < 8; ++i) dst[i] = prof_acc[i];
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kUMmaWarp) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace

int make_tmap_4d_bf16(CUtensorMap* out, const void* base, const uint64_t dims[4],
                      const uint64_t strides_bytes[3], const uint32_t box[4]);

int upconv_fused_launch(const UpFusedParams& pin, const void* a_hi, const void* a_lo,
                        const void* w_hi, const void* w_lo, cudaStream_t stream) {
  UpFusedParams p = pin;
  const int W = p.W, H = p.H;
  if (W < 4 || W > UM || (W & (W - 1)) != 0 || H < 1 || p.Cin % UBK != 0 || p.Cout % UNC != 0 ||
      p.B < 1) {
    set_last_error("upconv_fused: unsupported shape B=%d Cin=%d Cout=%d H=%d W=%d", p.B, p.Cin,
                   p.Cout, H, W);
    return RW_ERR_BAD_ARG;
  }
  const int G = UM / W;
  const int nbg = (p.B + G - 1) / G;
  p.ncg = p.Cout / UNC;
  const int sms = device_sm_count();
  // row bands: more items balance the static round-robin better, every band costs two warm-up
  // rows; pick the cheapest makespan
  int best = 1;
  long long best_cost = -1;
  for (int nb = 1; nb <= 32 && H / nb >= 4; nb *= 2) {
    const long long items = static_cast<long long>(nbg) * p.ncg * nb;
    const long long waves = (items + sms - 1) / sms;
    const long long cost = waves * ((H + nb - 1) / nb + 2);
    if (best_cost < 0 || cost < best_cost) {
      best_cost = cost;
      best = nb;
    }
  }
  p.nbands = best;
  p.nitems = nbg * p.ncg * p.nbands;

  CUtensorMap ma_hi, ma_lo, mw_hi, mw_lo;
  const uint64_t dims[4] = {static_cast<uint64_t>(p.Cin), static_cast<uint64_t>(W + 1),
                            static_cast<uint64_t>(H + 1), static_cast<uint64_t>(p.B)};
  const uint64_t str[3] = {static_cast<uint64_t>(p.Cin) * 2,
                           static_cast<uint64_t>(W + 1) * p.Cin * 2,
                           static_cast<uint64_t>(H + 1) * (W + 1) * p.Cin * 2};
  const uint32_t box[4] = {UBK, static_cast<uint32_t>(W), 1u, static_cast<uint32_t>(G)};
  int rc;
  if ((rc = make_tmap_4d_bf16(&ma_hi, a_hi, dims, str, box))) return rc;
  if ((rc = make_tmap_4d_bf16(&ma_lo, a_lo, dims, str, box))) return rc;
  const uint64_t wrows = static_cast<uint64_t>(p.ncg) * UN;
  if ((rc = make_tmap_2d_bf16(&mw_hi, w_hi, p.Cin, wrows, static_cast<uint64_t>(p.Cin) * 2, UBK, UN)))
    return rc;
  if ((rc = make_tmap_2d_bf16(&mw_lo, w_lo, p.Cin, wrows, static_cast<uint64_t>(p.Cin) * 2, UBK, UN)))
    return rc;
  const int grid = p.nitems < sms ? p.nitems : sms;
  if (p.debug_prof != nullptr) {
    static bool attr_p = false;
    if (!attr_p) {
      rc = check_cuda(cudaFuncSetAttribute(upconv_fused_kernel<true>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, kUSmemTotal),
                      "upconv_fused smem attr");
      if (rc) return rc;
      attr_p = true;
    }
    upconv_fused_kernel<true><<<grid, kUThreads, kUSmemTotal, stream>>>(ma_hi, ma_lo, mw_hi, mw_lo, p);
    return check_cuda(cudaGetLastError(), "upconv_fused launch");
  }
  static bool attr_set = false;
  if (!attr_set) {
    rc = check_cuda(cudaFuncSetAttribute(upconv_fused_kernel<false>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, kUSmemTotal),
                    "upconv_fused smem attr");
    if (rc) return rc;
    attr_set = true;
  }
  upconv_fused_kernel<false><<<grid, kUThreads, kUSmemTotal, stream>>>(ma_hi, ma_lo, mw_hi, mw_lo, p);
  return check_cuda(cudaGetLastError(), "upconv_fused launch");
}

}  // namespace rw
