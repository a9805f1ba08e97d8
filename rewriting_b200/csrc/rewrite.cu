// rewrite.cu — the rank-r projected-gradient weight edit.
//
// Reference: ProgressiveGanRewriter.insert (rewrite/ganrewrite.py:254-298) and
// projected_conv (ganrewrite.py:806-813).  Per iteration the reference runs
//   loss = L1(v*, target_model(k*)) ; backward ; Adam.step ;
//   every `piter` its:  W <- W_ortho + P_d(W)
// as ~60 separate framework kernels over a 9.4 MB weight.
//
// Observation that shapes this kernel: with the key detached, *every* quantity
// of one iteration is local to one output channel o — t[o,:], demod[o], the L1
// gradient, dW[o,:,:,:], the Adam moments and the projection all touch only
// row o of W.  So the whole loop needs no grid-wide synchronisation: a CTA owns
// a few output channels, keeps W[o] (18 KB) in shared memory across iterations,
// streams m/v through L2 and writes one partial loss per (iteration, channel).
//
// Key crop layout: kpT [B][h+2][w+2][Cin] fp32 (zero border, channels-last) so
// that lanes <-> input channels gives coalesced 128-byte loads.
#include "rw_common.cuh"
#include "rw_kernels.h"

namespace rw {

namespace {

constexpr int kMaxW = 16;      // crop width handled by the register tile
constexpr int kMaxRank = 32;
constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
// the insert loop runs one output channel per CTA, 4 warps, 4 CTAs per SM: all 512 channels
// of a layer are resident at once (592 CTA slots) and their latency-bound phases interleave
constexpr int kLoopThreads = 128;
constexpr int kLoopWarps = kLoopThreads / 32;

// out[o,i,t] = base[o,i,t] + sign * sum_r d[r,i] * (sum_j W[o,j,t] d[r,j])
// one CTA per output channel o; row o of W is contiguous (Cin*taps floats).
__global__ void __launch_bounds__(kThreads)
project_rank_kernel(const float* __restrict__ w, const float* __restrict__ base,
                    const float* __restrict__ d, int rank, int Cin, int taps, float sign,
                    float* __restrict__ out) {
  extern __shared__ float sm[];
  float* ws = sm;                      // [Cin*taps]
  float* lam = sm + Cin * taps;        // [rank*taps]
  const int o = blockIdx.x;
  const int n = Cin * taps;
  const float* wrow = w + static_cast<size_t>(o) * n;
  for (int i = threadIdx.x; i < n; i += blockDim.x) ws[i] = wrow[i];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int rt = warp; rt < rank * taps; rt += kWarps) {
    const int r = rt / taps, t = rt - r * taps;
    float acc = 0.f;
    for (int j = lane; j < Cin; j += 32) acc = fmaf(ws[j * taps + t], __ldg(d + r * Cin + j), acc);
#pragma unroll
    for (int off = 16; off; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (lane == 0) lam[rt] = acc;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int ci = i / taps, t = i - ci * taps;
    float p = 0.f;
    for (int r = 0; r < rank; ++r) p = fmaf(lam[r * taps + t], __ldg(d + r * Cin + ci), p);
    const float b = base ? base[static_cast<size_t>(o) * n + i] : 0.f;
    out[static_cast<size_t>(o) * n + i] = b + sign * p;
  }
}

// ---------------------------------------------------------------------------
// fused insert loop
// ---------------------------------------------------------------------------
struct LoopSmem {
  // dynamic: W[Cin*9] | t[P] | gd[P] | red[...] ...
};

__global__ void __launch_bounds__(kLoopThreads, 4)
insert_loop_kernel(const InsertLoopParams p, const float* __restrict__ kpT) {
  extern __shared__ float sm[];
  const int Cin = p.Cin, h = p.h, w = p.w, B = p.B;
  const int P = B * h * w;
  const int wp = w + 2;
  const int nW = Cin * 9;
  float* Ws = sm;                 // [Cin*9]   current weight row
  float* tS = Ws + nW;            // [P]       raw conv output t
  float* gdS = tS + P;            // [P]       g * demod (wgrad coefficient)
  float* dWS = gdS + P;           // [Cin*9]   only used for projections (aliased scratch)
  float* lam = dWS + nW;          // [kMaxRank*9]
  float* sc_b = lam + kMaxRank * 9;   // [B] demod, [B] coeff, loss, misc (64 floats)
  float* demodS = sc_b;
  float* coefS = sc_b + 16;
  float* lossS = sc_b + 32;       // [kLoopWarps]
  float* GS = sc_b + 40;          // [kLoopWarps*? ] per-warp partial G[b] -> B<=... stored [kLoopWarps][B<=2]?

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float sc = rsqrtf(static_cast<float>(Cin * 9));
  const int nch = Cin / 32;       // channels per lane
  const float inv_numel = 1.0f / static_cast<float>(static_cast<long long>(B) * p.Cout * h * w);

  for (int o = blockIdx.x; o < p.Cout; o += gridDim.x) {
    float* Wg = p.W + static_cast<size_t>(o) * nW;
    float* mg = p.m + static_cast<size_t>(o) * nW;
    float* vg = p.v + static_cast<size_t>(o) * nW;
    for (int i = threadIdx.x; i < nW; i += kLoopThreads) Ws[i] = Wg[i];
    __syncthreads();

    for (int step = 0; step < p.nsteps; ++step) {
      const int it = p.it0 + step;
      // ---- demod[b] = rsqrt(sum_i style^2 * sum_uv (sc W)^2 + 1e-8), every warp redundantly
      if (warp < B) {
        const int b = warp;
        float acc = 0.f;
        for (int j = 0; j < nch; ++j) {
          const int i = lane + 32 * j;
          float ss = 0.f;
#pragma unroll
          for (int t = 0; t < 9; ++t) {
            const float v = sc * Ws[i * 9 + t];
            ss = fmaf(v, v, ss);
          }
          const float s = __ldg(p.style + b * Cin + i);
          acc = fmaf(s * s, ss, acc);
        }
#pragma unroll
        for (int off = 16; off; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
        if (lane == 0) demodS[b] = rsqrtf(acc + 1e-8f);
      }
      // ---- forward conv on the crop: warp <-> row units (b, y), lane <-> channel
      for (int u = warp; u < B * h; u += kLoopWarps) {
        const int b = u / h, y = u - b * h;
        float acc[kMaxW];
#pragma unroll
        for (int x = 0; x < kMaxW; ++x) acc[x] = 0.f;
        for (int j = 0; j < nch; ++j) {
          const int i = lane + 32 * j;
          float wr[9];
#pragma unroll
          for (int t = 0; t < 9; ++t) wr[t] = Ws[i * 9 + t];
#pragma unroll
          for (int r = 0; r < 3; ++r) {
            const float* krow = kpT + ((static_cast<size_t>(b) * (h + 2) + y + r) * wp) * Cin + i;
            float kv[kMaxW + 2];
#pragma unroll
            for (int x = 0; x < kMaxW + 2; ++x)
              kv[x] = (x < wp) ? __ldg(krow + static_cast<size_t>(x) * Cin) : 0.f;
#pragma unroll
            for (int x = 0; x < kMaxW; ++x) {
              acc[x] = fmaf(wr[r * 3 + 0], kv[x], acc[x]);
              acc[x] = fmaf(wr[r * 3 + 1], kv[x + 1], acc[x]);
              acc[x] = fmaf(wr[r * 3 + 2], kv[x + 2], acc[x]);
            }
          }
        }
#pragma unroll
        for (int x = 0; x < kMaxW; ++x) {
          float a = acc[x];
#pragma unroll
          for (int off = 16; off; off >>= 1) a += __shfl_xor_sync(0xffffffffu, a, off);
          if (lane == x && x < w) tS[u * w + x] = sc * a;
        }
      }
      __syncthreads();
      // ---- loss / output gradient, one thread per pixel; block-reduce loss and G[b]
      float lsum = 0.f;
      float gsum[4] = {0.f, 0.f, 0.f, 0.f};
      for (int q = threadIdx.x; q < P; q += kLoopThreads) {
        const int b = q / (h * w);
        const int pp = q - b * h * w;
        const float t = tS[q];
        float yv = t * demodS[b];
        float gate = 1.f;
        if (p.has_noise_act) {
          if (p.noise) yv += p.noise_w * __ldg(p.noise + b * h * w + pp);
          yv += __ldg(p.bias + o);
          gate = (yv > 0.f) ? 1.4142135623730951f : 0.2f * 1.4142135623730951f;
          yv = (yv > 0.f ? yv : 0.2f * yv) * 1.4142135623730951f;
        }
        const float tgt = __ldg(p.target + (static_cast<size_t>(b) * p.Cout + o) * h * w + pp);
        const float diff = yv - tgt;
        lsum += fabsf(diff);
        const float sgn = (diff > 0.f) ? 1.f : ((diff < 0.f) ? -1.f : 0.f);
        const float g = sgn * inv_numel * gate;
        gdS[q] = g * demodS[b];
#pragma unroll
        for (int bb = 0; bb < 4; ++bb)
          if (bb == b) gsum[bb] += g * t;
      }
#pragma unroll
      for (int off = 16; off; off >>= 1) {
        lsum += __shfl_xor_sync(0xffffffffu, lsum, off);
#pragma unroll
        for (int bb = 0; bb < 4; ++bb) gsum[bb] += __shfl_xor_sync(0xffffffffu, gsum[bb], off);
      }
      if (lane == 0) {
        lossS[warp] = lsum;
#pragma unroll
        for (int bb = 0; bb < 4; ++bb) GS[warp * 4 + bb] = gsum[bb];
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        float l = 0.f;
        for (int wv = 0; wv < kLoopWarps; ++wv) l += lossS[wv];
        p.loss_out[static_cast<size_t>(step) * p.Cout + o] = l;
      }
      if (threadIdx.x < B) {
        float G = 0.f;
        for (int wv = 0; wv < kLoopWarps; ++wv) G += GS[wv * 4 + threadIdx.x];
        const float dm = demodS[threadIdx.x];
        coefS[threadIdx.x] = G * dm * dm * dm;
      }
      __syncthreads();

      // ---- weight gradient: warp <-> channel pair, lane <-> channel; 9 accumulators each
      // Adam bias corrections as torch.optim.Adam computes them (python doubles)
      const double stepd = static_cast<double>(it + 1);
      const double bc1 = 1.0 - pow(static_cast<double>(p.beta1), stepd);
      const double bc2 = 1.0 - pow(static_cast<double>(p.beta2), stepd);
      const float step_size = static_cast<float>(static_cast<double>(p.lr) / bc1);
      const float bc2_sqrt = static_cast<float>(sqrt(bc2));
      const float one_m_b1 = 1.0f - p.beta1;
      const float one_m_b2 = 1.0f - p.beta2;

      for (int j = warp; j < nch; j += kLoopWarps) {
        const int i = lane + 32 * j;
        float acc[9];
#pragma unroll
        for (int t = 0; t < 9; ++t) acc[t] = 0.f;
        for (int b = 0; b < B; ++b) {
          for (int y = 0; y < h; ++y) {
            float gv[kMaxW];
#pragma unroll
            for (int x = 0; x < kMaxW; ++x) gv[x] = (x < w) ? gdS[(b * h + y) * w + x] : 0.f;
#pragma unroll
            for (int r = 0; r < 3; ++r) {
              const float* krow = kpT + ((static_cast<size_t>(b) * (h + 2) + y + r) * wp) * Cin + i;
              float kv[kMaxW + 2];
#pragma unroll
              for (int x = 0; x < kMaxW + 2; ++x)
                kv[x] = (x < wp) ? __ldg(krow + static_cast<size_t>(x) * Cin) : 0.f;
#pragma unroll
              for (int x = 0; x < kMaxW; ++x) {
                acc[r * 3 + 0] = fmaf(gv[x], kv[x], acc[r * 3 + 0]);
                acc[r * 3 + 1] = fmaf(gv[x], kv[x + 1], acc[r * 3 + 1]);
                acc[r * 3 + 2] = fmaf(gv[x], kv[x + 2], acc[r * 3 + 2]);
              }
            }
          }
        }
        // demod term: - sc^2 * W * sum_b coef[b] * style[b,i]^2
        float cs = 0.f;
        for (int b = 0; b < B; ++b) {
          const float s = __ldg(p.style + b * Cin + i);
          cs = fmaf(coefS[b], s * s, cs);
        }
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const float wv = Ws[i * 9 + t];
          const float g = sc * acc[t] - (sc * sc) * wv * cs;
          dWS[i * 9 + t] = g;
        }
      }
      __syncthreads();
      // ---- optional gradient projection onto span(d)   (ganrewrite.py:285-286)
      if (p.project_gradient) {
        for (int rt = warp; rt < p.rank * 9; rt += kLoopWarps) {
          const int r = rt / 9, t = rt - r * 9;
          float a = 0.f;
          for (int i = lane; i < Cin; i += 32) a = fmaf(dWS[i * 9 + t], __ldg(p.d + r * Cin + i), a);
#pragma unroll
          for (int off = 16; off; off >>= 1) a += __shfl_xor_sync(0xffffffffu, a, off);
          if (lane == 0) lam[rt] = a;
        }
        __syncthreads();
        for (int e = threadIdx.x; e < nW; e += kLoopThreads) {
          const int i = e / 9, t = e - i * 9;
          float pr = 0.f;
          for (int r = 0; r < p.rank; ++r) pr = fmaf(lam[r * 9 + t], __ldg(p.d + r * Cin + i), pr);
          dWS[e] = pr;
        }
        __syncthreads();
      }
      // ---- Adam (torch.optim.Adam, amsgrad=False, weight_decay=0)
      for (int e = threadIdx.x; e < nW; e += kLoopThreads) {
        const float g = dWS[e];
        float mm = mg[e], vv = vg[e];
        mm = mm + (g - mm) * one_m_b1;                 // exp_avg.lerp_(grad, 1-beta1)
        vv = vv * p.beta2 + one_m_b2 * g * g;          // mul_(beta2).addcmul_(g, g, 1-beta2)
        mg[e] = mm;
        vg[e] = vv;
        const float denom = sqrtf(vv) / bc2_sqrt + p.eps;
        Ws[e] = Ws[e] - step_size * (mm / denom);
      }
      __syncthreads();
      // ---- periodic projection  W <- W_ortho + P_d(W)   (ganrewrite.py:291-294)
      if (p.w_ortho != nullptr && (it % p.piter == 0 || it == p.niter_total - 1)) {
        for (int rt = warp; rt < p.rank * 9; rt += kLoopWarps) {
          const int r = rt / 9, t = rt - r * 9;
          float a = 0.f;
          for (int i = lane; i < Cin; i += 32) a = fmaf(Ws[i * 9 + t], __ldg(p.d + r * Cin + i), a);
#pragma unroll
          for (int off = 16; off; off >>= 1) a += __shfl_xor_sync(0xffffffffu, a, off);
          if (lane == 0) lam[rt] = a;
        }
        __syncthreads();
        const float* wo = p.w_ortho + static_cast<size_t>(o) * nW;
        for (int e = threadIdx.x; e < nW; e += kLoopThreads) {
          const int i = e / 9, t = e - i * 9;
          float pr = 0.f;
          for (int r = 0; r < p.rank; ++r) pr = fmaf(lam[r * 9 + t], __ldg(p.d + r * Cin + i), pr);
          Ws[e] = __ldg(wo + e) + pr;
        }
        __syncthreads();
      }
    }
    for (int i = threadIdx.x; i < nW; i += kLoopThreads) Wg[i] = Ws[i];
    __syncthreads();
  }
}

}  // namespace

}  // namespace rw

namespace rw {

int project_rank_launch_signed(const float* w, const float* base, const float* d, int rank,
                               int Cout, int Cin, int taps, float sign, float* out,
                               cudaStream_t stream) {
  if (rank < 1 || rank > 64) {
    set_last_error("project_rank: rank=%d out of range [1,64]", rank);
    return RW_ERR_BAD_ARG;
  }
  const size_t smem = (static_cast<size_t>(Cin) * taps + static_cast<size_t>(rank) * taps) * 4;
  if (smem > 200 * 1024) {
    set_last_error("project_rank: row too large for shared memory (%zu B)", smem);
    return RW_ERR_UNSUPPORTED;
  }
  static size_t attr = 0;
  if (smem > 48 * 1024 && smem > attr) {
    int rc = check_cuda(cudaFuncSetAttribute(project_rank_kernel,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             static_cast<int>(smem)),
                        "project_rank smem attr");
    if (rc) return rc;
    attr = smem;
  }
  project_rank_kernel<<<Cout, kThreads, smem, stream>>>(w, base, d, rank, Cin, taps, sign, out);
  return check_cuda(cudaGetLastError(), "project_rank launch");
}

int insert_loop_launch(const InsertLoopParams& p, cudaStream_t stream) {
  if (p.w > kMaxW || p.B > 4 || p.B < 1 || p.Cin % 32 != 0 || p.rank > kMaxRank ||
      static_cast<long long>(p.B) * p.h * p.w > 4096) {
    set_last_error("insert_loop: unsupported crop B=%d h=%d w=%d Cin=%d rank=%d", p.B, p.h, p.w,
                   p.Cin, p.rank);
    return RW_ERR_UNSUPPORTED;
  }
  const int P = p.B * p.h * p.w;
  const int nW = p.Cin * 9;
  const size_t smem = (static_cast<size_t>(2 * nW) + 2 * P + kMaxRank * 9 + 128) * sizeof(float);
  if (smem > 220 * 1024) {
    set_last_error("insert_loop: shared memory %zu B too large", smem);
    return RW_ERR_UNSUPPORTED;
  }
  static size_t attr = 0;
  if (smem > attr) {
    int rc = check_cuda(cudaFuncSetAttribute(insert_loop_kernel,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             static_cast<int>(smem)),
                        "insert_loop smem attr");
    if (rc) return rc;
    attr = smem;
  }
  const int grid = p.Cout;     // one output channel per CTA
  insert_loop_kernel<<<grid, kLoopThreads, smem, stream>>>(p, p.key);
  return check_cuda(cudaGetLastError(), "insert_loop launch");
}

}  // namespace rw
