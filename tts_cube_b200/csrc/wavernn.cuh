// Path W4: the autoregressive WaveRNN sample loop (reference cube/networks/modules.py:478-497) as ONE persistent
// cooperative kernel.  The loop is latency-bound (two GRU cells, two linears and a sampling head per audio sample;
// ~5 MFLOP and ~10 dependent steps), so the design goal is to remove every launch and every weight fetch from the
// critical path:
//   * weights are STATIONARY in shared memory, sliced over the grid by hidden unit (H=512 over 128 CTAs: 4 units,
//     i.e. 12 GRU rows per layer + 2 pre-output rows per CTA; ~210 KB of smem per CTA in total with the batch state);
//   * the input-to-hidden product of the first GRU for everything except the fed-back sample (80 mel + 21 low-res
//     conditioning channels) does not depend on the recurrence: it is precomputed for all T by the tiled conv kernel;
//   * per sample: 3 grid-wide barriers (after h1, after h2, after the pre-output layer); the small head (<= 32
//     logits) and the sampling are computed redundantly by every CTA so the fed-back sample needs no barrier;
//   * random draws are kernel INPUTS ([T][B][K]), which makes the loop replayable against the oracle.
// Batch items (the 20 time-folded chunks of CubenetVocoder._inference_batch) advance in lock-step.
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace cube {
namespace wrnn {

namespace cg = cooperative_groups;

enum { HEAD_MOL = 0, HEAD_GM = 1, HEAD_MULAW = 2, HEAD_RAW = 3 };
constexpr int PRE = 256;          // _preoutput width (cube/networks/modules.py:430)
constexpr int THREADS = 256;

struct WrnnParams {
  int H, L, B, T, S, head, U, P;   // U = hidden units per CTA, P = pre-output rows per CTA
  // weights (global; copied to smem once)
  const float* gx;                 // [B][3H][T]: W_ih1[:, :ic-1] . cond + b_ih1 for every step
  const float* w_last;             // [3H]      column of W_ih1 that multiplies the fed-back sample
  const float* whh1; const float* bhh1;                       // [3H][H], [3H]
  const float* wih2; const float* bih2; const float* whh2; const float* bhh2;   // layer 2 (L == 2)
  const float* wpre; const float* bpre;                       // [256][H], [256]
  const float* wout; const float* bout;                       // [S][256], [S]
  // state / scratch (global)
  float* hbuf;                     // [2][L][B][H]  ping-pong hidden state (zero-initialised)
  float* prebuf;                   // [B][256]
  float* logits;                   // [B][S]        (only when S > 32)
  // draws and output
  const float* draws;              // MOL: [T][B][nr_mix + 1] (nr_mix uniforms for the mixture pick, then 1 for the
                                   // logistic); GM: [T][B][1] normals; MULAW/RAW: [T][B][S] uniforms
  float* x_out;                    // [B][T]
  float log_scale_min;
};

__device__ __forceinline__ float sigm(float x) { return 1.f / (1.f + expf(-x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// dot(w[0..n), x[0..n)) with the 32 lanes striding over n (both in shared memory)
__device__ __forceinline__ float dot_ws(const float* w, const float* x, int n, int lane) {
  float a = 0.f;
  for (int k = lane; k < n; k += 32) a = fmaf(w[k], x[k], a);
  return warp_sum(a);
}

__device__ __forceinline__ void copy_g2s(float* dst, const float* src, int n, int tid) {
  for (int i = tid; i < n; i += THREADS) dst[i] = src[i];
}

__global__ void __launch_bounds__(THREADS, 1) wavernn_kernel(const WrnnParams p) {
  extern __shared__ __align__(16) float sm[];
  cg::grid_group grid = cg::this_grid();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int H = p.H, B = p.B, U = p.U, P = p.P, S = p.S;
  const int u0 = blockIdx.x * U;                     // first hidden unit of this CTA
  const int nu = max(0, min(U, H - u0));
  const int p0 = blockIdx.x * P;                     // first pre-output row of this CTA
  const int np = max(0, min(P, PRE - p0));
  const bool small_head = S <= 32;
  const int s0 = small_head ? 0 : blockIdx.x * ((S + gridDim.x - 1) / gridDim.x);
  const int ns = small_head ? S : max(0, min((S + (int)gridDim.x - 1) / (int)gridDim.x, S - s0));

  // ---- shared-memory carve-up ----
  float* W1 = sm;                          // [3][U][H]
  float* wl = W1 + 3 * U * H;              // [3][U]  w_last
  float* b1 = wl + 3 * U;                  // [3][U]  b_hh1
  float* W2i = b1 + 3 * U;                 // [3][U][H]   (L == 2)
  float* W2h = W2i + (p.L == 2 ? 3 * U * H : 0);
  float* b2i = W2h + (p.L == 2 ? 3 * U * H : 0);
  float* b2h = b2i + (p.L == 2 ? 3 * U : 0);
  float* Wp = b2h + (p.L == 2 ? 3 * U : 0);   // [P][H]
  float* bp = Wp + P * H;                  // [P]
  float* Wo = bp + P;                      // [ns][256]
  float* bo = Wo + (small_head ? S : (S + gridDim.x - 1) / gridDim.x) * PRE;   // [ns]
  float* xin = bo + (small_head ? S : (S + gridDim.x - 1) / gridDim.x);        // [B][H]   layer input / state staging
  float* hin = xin + B * H;                // [B][H]   (layer 2: previous h2)
  float* prev = hin + (p.L == 2 ? B * H : 0);   // [B][256]
  float* lg = prev + B * PRE;              // [B][S] logits (small head) or staging
  float* lastx = lg + B * max(S, 1);       // [B]

  // ---- weights -> smem (once) ----
  for (int g = 0; g < 3; ++g)
    for (int u = 0; u < nu; ++u) {
      const int row = g * H + u0 + u;
      copy_g2s(W1 + (g * U + u) * H, p.whh1 + (size_t)row * H, H, tid);
      if (p.L == 2) {
        copy_g2s(W2i + (g * U + u) * H, p.wih2 + (size_t)row * H, H, tid);
        copy_g2s(W2h + (g * U + u) * H, p.whh2 + (size_t)row * H, H, tid);
      }
      if (tid == 0) {
        wl[g * U + u] = p.w_last[row];
        b1[g * U + u] = p.bhh1[row];
        if (p.L == 2) { b2i[g * U + u] = p.bih2[row]; b2h[g * U + u] = p.bhh2[row]; }
      }
    }
  for (int r = 0; r < np; ++r) {
    copy_g2s(Wp + r * H, p.wpre + (size_t)(p0 + r) * H, H, tid);
    if (tid == 0) bp[r] = p.bpre[p0 + r];
  }
  for (int r = 0; r < ns; ++r) {
    copy_g2s(Wo + r * PRE, p.wout + (size_t)(s0 + r) * PRE, PRE, tid);
    if (tid == 0) bo[r] = p.bout[s0 + r];
  }
  for (int i = tid; i < B; i += THREADS) lastx[i] = 0.f;   // last_x = 0 (modules.py:472-473)
  __syncthreads();

  const size_t hstride = (size_t)p.L * B * H;      // one ping-pong half of hbuf
  int cur = 0;
  for (int t = 0; t < p.T; ++t) {
    const int nxt = cur ^ 1;
    // ================= layer 1 =================
    copy_g2s(xin, p.hbuf + cur * hstride, B * H, tid);           // h1(t-1)
    __syncthreads();
    for (int task = warp; task < nu * B; task += THREADS / 32) {
      const int u = task / B, b = task - u * B;
      const float* hb = xin + b * H;
      const float dr = dot_ws(W1 + (0 * U + u) * H, hb, H, lane);
      const float dz = dot_ws(W1 + (1 * U + u) * H, hb, H, lane);
      const float dn = dot_ws(W1 + (2 * U + u) * H, hb, H, lane);
      if (lane == 0) {
        const float lx = lastx[b];
        const float* gx = p.gx + ((size_t)b * 3 * H + u0 + u) * p.T + t;
        const float gr = gx[0] + wl[0 * U + u] * lx, gz = gx[(size_t)H * p.T] + wl[1 * U + u] * lx,
                    gn = gx[(size_t)2 * H * p.T] + wl[2 * U + u] * lx;
        const float r = sigm(gr + dr + b1[0 * U + u]);
        const float z = sigm(gz + dz + b1[1 * U + u]);
        const float n = tanhf(gn + r * (dn + b1[2 * U + u]));
        p.hbuf[nxt * hstride + (size_t)b * H + u0 + u] = (1.f - z) * n + z * hb[u0 + u];
      }
    }
    grid.sync();
    // ================= layer 2 =================
    if (p.L == 2) {
      copy_g2s(xin, p.hbuf + nxt * hstride, B * H, tid);                       // h1(t)
      copy_g2s(hin, p.hbuf + cur * hstride + (size_t)B * H, B * H, tid);       // h2(t-1)
      __syncthreads();
      for (int task = warp; task < nu * B; task += THREADS / 32) {
        const int u = task / B, b = task - u * B;
        const float *xb = xin + b * H, *hb = hin + b * H;
        const float ir = dot_ws(W2i + (0 * U + u) * H, xb, H, lane), hr = dot_ws(W2h + (0 * U + u) * H, hb, H, lane);
        const float iz = dot_ws(W2i + (1 * U + u) * H, xb, H, lane), hz = dot_ws(W2h + (1 * U + u) * H, hb, H, lane);
        const float in_ = dot_ws(W2i + (2 * U + u) * H, xb, H, lane), hn = dot_ws(W2h + (2 * U + u) * H, hb, H, lane);
        if (lane == 0) {
          const float r = sigm(ir + b2i[0 * U + u] + hr + b2h[0 * U + u]);
          const float z = sigm(iz + b2i[1 * U + u] + hz + b2h[1 * U + u]);
          const float n = tanhf(in_ + b2i[2 * U + u] + r * (hn + b2h[2 * U + u]));
          p.hbuf[nxt * hstride + (size_t)B * H + (size_t)b * H + u0 + u] = (1.f - z) * n + z * hb[u0 + u];
        }
      }
      grid.sync();
    }
    // ================= pre-output: tanh(W h + b) =================
    copy_g2s(xin, p.hbuf + nxt * hstride + (size_t)(p.L - 1) * B * H, B * H, tid);   // top layer's h(t)
    __syncthreads();
    for (int task = warp; task < np * B; task += THREADS / 32) {
      const int r = task / B, b = task - r * B;
      const float d = dot_ws(Wp + r * H, xin + b * H, H, lane);
      if (lane == 0) p.prebuf[(size_t)b * PRE + p0 + r] = tanhf(d + bp[r]);
    }
    grid.sync();
    // ================= output layer + sampling head =================
    copy_g2s(prev, p.prebuf, B * PRE, tid);
    __syncthreads();
    for (int task = warp; task < ns * B; task += THREADS / 32) {
      const int r = task / B, b = task - r * B;
      const float d = dot_ws(Wo + r * PRE, prev + b * PRE, PRE, lane) + bo[r];
      if (lane == 0) {
        if (small_head) lg[b * S + r] = d;
        else p.logits[(size_t)b * S + s0 + r] = d;
      }
    }
    if (!small_head) {
      grid.sync();
      copy_g2s(lg, p.logits, B * S, tid);
    }
    __syncthreads();
    if (p.head == HEAD_MOL || p.head == HEAD_GM) {
      if (tid < B) {
        const int b = tid;
        const float* y = lg + b * S;
        float x;
        if (p.head == HEAD_MOL) {       // cube/networks/loss.py:176-199
          const int nm = S / 3;
          const float* u = p.draws + ((size_t)t * B + b) * (nm + 1);
          int best = 0;
          float bs = -__builtin_inff();
          for (int k = 0; k < nm; ++k) {
            const float s = y[k] - logf(-logf(u[k]));
            if (s > bs) { bs = s; best = k; }
          }
          const float ls = fmaxf(y[2 * nm + best], p.log_scale_min);
          const float uu = u[nm];
          x = fminf(fmaxf(y[nm + best] + expf(ls) * (logf(uu) - logf(1.f - uu)), -1.f), 1.f);
        } else {                        // cube/networks/loss.py:50-52
          x = y[0] + __fmul_rn(p.draws[(size_t)t * B + b], 0.8f) * expf(y[1]);
        }
        lastx[b] = x;
        if (blockIdx.x == 0) p.x_out[(size_t)b * p.T + t] = x;
      }
    } else {                            // categorical heads: Gumbel-max over S logits, one warp per batch item
      for (int b = warp; b < B; b += THREADS / 32) {
        const float* y = lg + b * S;
        const float* u = p.draws + ((size_t)t * B + b) * S;
        float bs = -__builtin_inff();
        int bi = 0x7fffffff;
        for (int k = lane; k < S; k += 32) {
          const float s = y[k] - logf(-logf(u[k]));
          if (s > bs) { bs = s; bi = k; }
        }
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) {
          const float os = __shfl_xor_sync(0xffffffffu, bs, o);
          const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
          if (os > bs || (os == bs && oi < bi)) { bs = os; bi = oi; }
        }
        if (lane == 0) {
          float x;
          if (p.head == HEAD_MULAW) x = c_mulaw_decode[bi & 255];
          else x = __fmul_rn(__fsub_rn(__fdiv_rn((float)bi, 255.0f), 0.5f), 2.0f);
          lastx[b] = x;
          if (blockIdx.x == 0) p.x_out[(size_t)b * p.T + t] = x;
        }
      }
    }
    __syncthreads();
    cur = nxt;
  }
}

// conditioning tensor [B][C][T] channel-first for the precomputed input product:
//   c < 80: UpsampleNetR(mel) (repeat)           cube/networks/modules.py:378-389, mel is [B][F][80] time-major
//   80 <= c < 100: UpsampleNetR(lowres features) lowres [B][20][Tl]
//   c == 100: UpsampleNetI(x_low) (linear, align_corners=False)  modules.py:346-354
__global__ void wavernn_cond_kernel(const float* __restrict__ mel, const float* __restrict__ lowf, const float* __restrict__ xlow,
                                    float* __restrict__ cond, int B, int F, int nmel, int Tl, int up, int upl, int T, int C) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int c = blockIdx.y, b = blockIdx.z;
  if (t >= T) return;
  float v;
  if (c < nmel) {
    v = mel[((size_t)b * F + min(t / up, F - 1)) * nmel + c];
  } else if (c < nmel + 20) {
    v = lowf[((size_t)b * 20 + (c - nmel)) * Tl + min(t / upl, Tl - 1)];
  } else {
    float src = ((float)t + 0.5f) / (float)upl - 0.5f;
    src = fmaxf(src, 0.f);
    const int i0 = min((int)src, Tl - 1), i1 = min(i0 + 1, Tl - 1);
    const float w = src - (float)i0;
    v = (1.f - w) * xlow[(size_t)b * Tl + i0] + w * xlow[(size_t)b * Tl + i1];
  }
  cond[((size_t)b * C + c) * T + t] = v;
}

}  // namespace wrnn
}  // namespace cube
