// gram_tc.cu — tcgen05 "col-GEMM": contraction over pixel rows.
//
//   out[m, n] = sum_r A[r + shift_a, m] * B[r + shift_b, n]
//
// With A == B == key planes this is the key second moment  mom2 += sum_n a_n a_n^T
// that the reference accumulates with one rank-1 addbmm per row
// (utils/runningstats.py:1086-1097, 1181-1190 — 10 240 batched 512x1 @ 1x512
// products per call).  With A = output-gradient planes and B = shifted key
// planes it is the conv weight gradient (autograd of models.py:313-329).
//
// Operands are the bf16 hi/lo planes [rows][C] (channels contiguous), i.e. the
// contraction index is the *slow* one: both operands are MN-major for the
// tensor core.  TMA boxes of 64 channels x 64 rows land as 128-byte swizzled
// rows; the UMMA descriptor walks 8-row groups with SBO and 64-channel blocks
// with LBO.  Three MMAs per k-step (hi*hi + lo*hi + hi*lo), fp32 accumulate in
// TMEM.  Row ranges are split across CTAs; partial tiles go to a workspace that
// a second kernel reduces in a fixed order (bit-reproducible, unlike atomics).
#include "rw_common.cuh"
#include "rw_kernels.h"

namespace rw {

namespace {

constexpr int TM = 128;
constexpr int TN = 128;
constexpr int RB = 64;            // rows (contraction) per pipeline stage
constexpr int UMMA_K = 16;
constexpr int kStages = 3;
constexpr int kNumThreads = 192;
// fp32 accumulation in the tensor core truncates (see conv_tc.cu): a TMEM accumulator holds
// at most kChunkRB row-blocks (1024 rows = 192 accumulations); chunks are summed in fp32
// registers (round-to-nearest) by the epilogue warps.
constexpr int kChunkRB = 16;
constexpr int kNumAcc = 4;
constexpr int kBlockBytes = 64 * RB * 2;       // one 64-channel x RB-row box
constexpr int kPlaneBytes = (TM / 64) * kBlockBytes;
constexpr int kStageBytes = 4 * kPlaneBytes;   // A_hi, A_lo, B_hi, B_lo
constexpr int kSmemTotal = kStages * kStageBytes + 1024 + 256;

struct Barriers {
  uint64_t full[kStages];
  uint64_t empty[kStages];
  uint64_t tmem_full[kNumAcc];
  uint64_t tmem_empty[kNumAcc];
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(kNumThreads, 1)
gram_tc_kernel(const __grid_constant__ CUtensorMap map_a_hi,
               const __grid_constant__ CUtensorMap map_a_lo,
               const __grid_constant__ CUtensorMap map_b_hi,
               const __grid_constant__ CUtensorMap map_b_lo, const GramTcParams p,
               const int lbo_bytes, const int sbo_bytes) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  Barriers* bars = reinterpret_cast<Barriers*>(smem + kStages * kStageBytes);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // tile decode (optionally upper-triangular tiles only)
  const int nt_count = p.Cn / TN;
  int mt, nt;
  if (p.upper_only) {
    int t = blockIdx.x;
    mt = 0;
    while (t >= nt_count - mt) { t -= nt_count - mt; ++mt; }
    nt = mt + t;
  } else {
    mt = blockIdx.x / nt_count;
    nt = blockIdx.x % nt_count;
  }
  const int m0 = mt * TM;
  const int n0 = nt * TN;
  const int tap = blockIdx.z;
  const int shift_b = p.shift_b + p.tap_shift_b[tap];
  const int shift_a = p.shift_a + p.tap_shift_a[tap];
  const int acol = p.tap_acol[tap];
  const int col_ofs = p.tap_col_ofs[tap];

  const int total_rb = (p.rows + RB - 1) / RB;
  const int split = blockIdx.y;
  const int rb_per = (total_rb + p.splits - 1) / p.splits;
  const int rb_begin = split * rb_per;
  int rb_end = rb_begin + rb_per;
  if (rb_end > total_rb) rb_end = total_rb;
  const int num_rb = rb_end > rb_begin ? rb_end - rb_begin : 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a_hi);
    tma_prefetch_desc(&map_a_lo);
    tma_prefetch_desc(&map_b_hi);
    tma_prefetch_desc(&map_b_lo);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&bars->full[s], 1);
      mbar_init(&bars->empty[s], 1);
    }
    for (int s = 0; s < kNumAcc; ++s) {
      mbar_init(&bars->tmem_full[s], 1);
      mbar_init(&bars->tmem_empty[s], 4);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<kNumAcc * TN>(&bars->tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int rb = rb_begin; rb < rb_begin + num_rb; ++rb) {
        mbar_wait(&bars->empty[stage], phase ^ 1u);
        uint8_t* st = smem + stage * kStageBytes;
        mbar_expect_tx(&bars->full[stage], kStageBytes);
        const int ra = rb * RB + shift_a;
        const int rbb = rb * RB + shift_b;
#pragma unroll
        for (int j = 0; j < TM / 64; ++j) {
          tma_load_2d(st + j * kBlockBytes, &map_a_hi, &bars->full[stage], acol + m0 + j * 64, ra);
          tma_load_2d(st + kPlaneBytes + j * kBlockBytes, &map_a_lo, &bars->full[stage],
                      acol + m0 + j * 64, ra);
          tma_load_2d(st + 2 * kPlaneBytes + j * kBlockBytes, &map_b_hi, &bars->full[stage],
                      n0 + j * 64, rbb);
          tma_load_2d(st + 3 * kPlaneBytes + j * kBlockBytes, &map_b_lo, &bars->full[stage],
                      n0 + j * 64, rbb);
        }
        if (++stage == kStages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_bf16(TM, TN, 1, 1);
    // warp-uniform loop, one elected lane issues (descriptors stay in uniform registers)
    {
      int stage = 0;
      uint32_t phase = 0;
      uint32_t chunk = 0;
      for (int i0 = 0; i0 < num_rb; i0 += kChunkRB, ++chunk) {
        const int as = chunk % kNumAcc;
        const uint32_t aphase = (chunk / kNumAcc) & 1u;
        mbar_wait(&bars->tmem_empty[as], aphase ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * TN;
        const int i_end = (i0 + kChunkRB < num_rb) ? i0 + kChunkRB : num_rb;
        for (int i = i0; i < i_end; ++i) {
          mbar_wait(&bars->full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * kStageBytes);
          const uint64_t da_hi = make_smem_desc(sa, lbo_bytes, sbo_bytes, kSwizzle128B);
          const uint64_t da_lo =
              make_smem_desc(sa + kPlaneBytes, lbo_bytes, sbo_bytes, kSwizzle128B);
          const uint64_t db_hi =
              make_smem_desc(sa + 2 * kPlaneBytes, lbo_bytes, sbo_bytes, kSwizzle128B);
          const uint64_t db_lo =
              make_smem_desc(sa + 3 * kPlaneBytes, lbo_bytes, sbo_bytes, kSwizzle128B);
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < RB / UMMA_K; ++kk) {
              // 16 rows = two 8-row swizzle groups of 1024 B
              const uint64_t adv = static_cast<uint64_t>((kk * UMMA_K * 128) >> 4);
              umma_bf16(tmem_d, da_lo + adv, db_hi + adv, idesc, ((i - i0) | kk) != 0);
              umma_bf16(tmem_d, da_hi + adv, db_lo + adv, idesc, 1u);
              umma_bf16(tmem_d, da_hi + adv, db_hi + adv, idesc, 1u);
            }
            umma_commit(&bars->empty[stage]);
            if (i + 1 == i_end) umma_commit(&bars->tmem_full[as]);
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else {
    const int q = warp & 3;
    const int row = m0 + q * 32 + lane;
    float* dst = p.partial + (static_cast<size_t>(split) * p.Cm + row) * p.ldp + col_ofs + n0;
    float acc[TN];
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[j] = 0.f;
    uint32_t chunk = 0;
    for (int i0 = 0; i0 < num_rb; i0 += kChunkRB, ++chunk) {
      const int as = chunk % kNumAcc;
      const uint32_t aphase = (chunk / kNumAcc) & 1u;
      mbar_wait(&bars->tmem_full[as], aphase);
      tc_fence_after();
#pragma unroll
      for (int c0 = 0; c0 < TN; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + static_cast<uint32_t>(as * TN + c0) +
                          (static_cast<uint32_t>(q * 32) << 16), v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[c0 + j] += __uint_as_float(v[j]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->tmem_empty[as]);
    }
#pragma unroll
    for (int j = 0; j < TN; j += 4) {
      *reinterpret_cast<float4*>(dst + j) = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<kNumAcc * TN>(tmem_base);
  }
}

// out[m, n] (= | +=) sum_s partial[s][m][n], 32 x 32 tiles, float reads along n (coalesced).
// mirror_upper: only tiles on or above the diagonal are computed; an off-diagonal tile is also
// stored transposed through shared memory, so both the reads and the two stores are full
// 128-byte rows (the first version read the lower triangle column-wise: 32 lines per load).
__global__ void __launch_bounds__(256)
reduce_partials_kernel(const float* __restrict__ partial, int splits, int M, int N, long long ldp,
                       float* __restrict__ out, long long ldo, int accumulate, int mirror_upper) {
  __shared__ float tile[32][33];
  const int tn = blockIdx.x, tm = blockIdx.y;
  if (mirror_upper && tm > tn) return;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 32 x 8 threads
  const size_t plane = static_cast<size_t>(M) * ldp;
  const bool diag = mirror_upper && tm == tn;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int lm = ty + r * 8;
    const int m = tm * 32 + lm, n = tn * 32 + tx;
    float acc = 0.f;
    if (m < M && n < N && !(diag && lm > tx)) {
      const float* src = partial + static_cast<size_t>(m) * ldp + n;
      for (int s = 0; s < splits; ++s) acc += src[s * plane];
      float* o = out + static_cast<size_t>(m) * ldo + n;
      *o = accumulate ? (*o + acc) : acc;
    }
    tile[lm][tx] = acc;
  }
  if (!mirror_upper) return;
  if (diag) {
    // diagonal tile: the strictly-lower entries are the transposed upper ones (bit-identical,
    // so the accumulated matrix stays exactly symmetric)
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int lm = ty + r * 8;
      const int m = tm * 32 + lm, n = tn * 32 + tx;
      if (lm > tx && m < M && n < N) {
        const float v = tile[tx][lm];
        float* o = out + static_cast<size_t>(m) * ldo + n;
        *o = accumulate ? (*o + v) : v;
      }
    }
    return;
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int ln = ty + r * 8;                  // row of the mirrored tile = column of this one
    const int m2 = tn * 32 + ln, n2 = tm * 32 + tx;
    if (m2 < N && n2 < M) {
      float* o = out + static_cast<size_t>(m2) * ldo + n2;
      const float v = tile[tx][ln];
      *o = accumulate ? (*o + v) : v;
    }
  }
}

}  // namespace

// test hook: descriptor geometry can be overridden (tools/umma_probe) to pin the
// MN-major LBO/SBO convention on hardware.
static int g_gram_lbo = kBlockBytes;
static int g_gram_sbo = 1024;
void gram_tc_set_desc(int lbo, int sbo) { g_gram_lbo = lbo; g_gram_sbo = sbo; }

int gram_tc_launch(const GramTcParams& p, const void* a_hi, const void* a_lo, const void* b_hi,
                   const void* b_lo, cudaStream_t stream) {
  if (p.Cm % TM != 0 || p.Cn % TN != 0 || p.rows <= 0 || p.splits < 1 || p.ntaps < 1 || p.ntaps > 9) {
    set_last_error("gram_tc: unsupported shape Cm=%d Cn=%d rows=%d splits=%d", p.Cm, p.Cn, p.rows,
                   p.splits);
    return RW_ERR_BAD_ARG;
  }
  if (p.upper_only && (p.Cm != p.Cn)) {
    set_last_error("gram_tc: upper_only needs a square output");
    return RW_ERR_BAD_ARG;
  }
  // rows r >= p.rows must contribute zero: clip the A operand's row extent at the
  // contraction range so the TMA zero-fills past it (B may then hold anything).
  int max_sa = p.tap_shift_a[0];
  for (int t = 1; t < p.ntaps; ++t) max_sa = p.tap_shift_a[t] > max_sa ? p.tap_shift_a[t] : max_sa;
  long long a_extent = static_cast<long long>(p.rows) + p.shift_a + max_sa;
  if (a_extent > p.rows_a) a_extent = p.rows_a;
  if (a_extent < 1) a_extent = 1;
  const int a_cols = p.a_cols > 0 ? p.a_cols : p.Cm;
  const long long b_extent = p.rows_b;
  CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo;
  int rc;
  if ((rc = make_tmap_2d_bf16(&ma_hi, a_hi, a_cols, a_extent, (uint64_t)a_cols * 2, 64, RB))) return rc;
  if ((rc = make_tmap_2d_bf16(&ma_lo, a_lo, a_cols, a_extent, (uint64_t)a_cols * 2, 64, RB))) return rc;
  if ((rc = make_tmap_2d_bf16(&mb_hi, b_hi, p.Cn, b_extent, (uint64_t)p.Cn * 2, 64, RB))) return rc;
  if ((rc = make_tmap_2d_bf16(&mb_lo, b_lo, p.Cn, b_extent, (uint64_t)p.Cn * 2, 64, RB))) return rc;

  static bool attr_set = false;
  if (!attr_set) {
    rc = check_cuda(cudaFuncSetAttribute(gram_tc_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal),
                    "gram_tc smem attr");
    if (rc) return rc;
    attr_set = true;
  }
  const int mt = p.Cm / TM, nt = p.Cn / TN;
  const int tiles = p.upper_only ? mt * (mt + 1) / 2 : mt * nt;
  dim3 grid(tiles, p.splits, p.ntaps);
  gram_tc_kernel<<<grid, kNumThreads, kSmemTotal, stream>>>(ma_hi, ma_lo, mb_hi, mb_lo, p,
                                                            g_gram_lbo, g_gram_sbo);
  return check_cuda(cudaGetLastError(), "gram_tc launch");
}

int reduce_partials_launch(const float* partial, int splits, int M, int N, long long ldp,
                           float* out, long long ldo, int accumulate, int mirror_upper,
                           cudaStream_t stream) {
  dim3 grid((N + 31) / 32, (M + 31) / 32);
  reduce_partials_kernel<<<grid, 256, 0, stream>>>(partial, splits, M, N, ldp, out, ldo, accumulate,
                                                   mirror_upper);
  return check_cuda(cudaGetLastError(), "reduce_partials launch");
}

}  // namespace rw
