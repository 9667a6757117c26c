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
  int DG;                          // rows of the dot-product scratch: max(6U, P, output rows per CTA)
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

__host__ __device__ constexpr int al4(int x) { return (x + 3) & ~3; }
__device__ __forceinline__ float sigm(float x) { return 1.f / (1.f + expf(-x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ void copy_g2s(float* dst, const float* src, int n, int tid) {
  for (int i = tid; i < n; i += THREADS) dst[i] = src[i];
}

// global [B][K] -> shared, TRANSPOSED to [K][B]: in the dot products below consecutive lanes are consecutive batch
// items, so they read consecutive shared-memory words while the weight element is a warp broadcast
__device__ __forceinline__ void copy_g2s_T(float* dst, const float* src, int B, int K, int tid) {
  for (int b = 0; b < B; ++b)
    for (int k = tid; k < K; k += THREADS) dst[k * B + b] = src[b * K + k];   // coalesced reads, no division
}

// out[r*B + b] = dot(W[r][0..K), x[b][0..K)) for r < nrows, b < B; W row-major in smem, xT = [K][B] in smem.
// One thread per (row, batch item) - no shuffles; when there are fewer than 256 such pairs the K range is split over
// KS thread groups and the partial sums are combined through `part` (KS*nrows*B floats).
__device__ __forceinline__ void dots(const float* W, int nrows, const float* xT, int B, int K, float* out, float* part, int tid) {
  const int ntask = nrows * B;
  if (ntask == 0) { __syncthreads(); __syncthreads(); return; }
  int KS = 1;
  while (KS < 16 && ntask * KS * 2 <= THREADS) KS *= 2;
  const int kl = ((K + KS - 1) / KS + 7) & ~7;          // slices start on 32-byte boundaries (float4 weight loads)
  for (int idx = tid; idx < ntask * KS; idx += THREADS) {
    const int ks = idx / ntask, task = idx - ks * ntask;
    const int r = task / B, b = task - r * B;
    const int k0 = ks * kl, k1 = min(K, k0 + kl);
    const float* w = W + r * K;
    // 8 independent loads per step and 4 accumulators: the loop is shared-memory-latency bound otherwise
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    const float* xp = xT + b;
    int k = k0;
    for (; k + 7 < k1; k += 8) {
      const float4 wa = *reinterpret_cast<const float4*>(w + k), wb = *reinterpret_cast<const float4*>(w + k + 4);
      const float x0 = xp[k * B], x1 = xp[(k + 1) * B], x2 = xp[(k + 2) * B], x3 = xp[(k + 3) * B];
      const float x4 = xp[(k + 4) * B], x5 = xp[(k + 5) * B], x6 = xp[(k + 6) * B], x7 = xp[(k + 7) * B];
      a0 = fmaf(wa.x, x0, a0); a1 = fmaf(wa.y, x1, a1); a2 = fmaf(wa.z, x2, a2); a3 = fmaf(wa.w, x3, a3);
      a0 = fmaf(wb.x, x4, a0); a1 = fmaf(wb.y, x5, a1); a2 = fmaf(wb.z, x6, a2); a3 = fmaf(wb.w, x7, a3);
    }
    for (; k < k1; ++k) a0 = fmaf(w[k], xp[k * B], a0);
    (KS == 1 ? out : part)[idx] = (a0 + a1) + (a2 + a3);
  }
  __syncthreads();
  if (KS > 1) {
    for (int task = tid; task < ntask; task += THREADS) {
      float a = 0.f;
      for (int ks = 0; ks < KS; ++ks) a += part[ks * ntask + task];
      out[task] = a;
    }
  }
  __syncthreads();
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

  // ---- shared-memory carve-up (every array starts on a 16-byte boundary; mirrored by wrnn_geom() on the host) ----
  const int so_rows = small_head ? S : (S + gridDim.x - 1) / gridDim.x;
  float* W1 = sm;                          // [3][U][H]
  float* wl = W1 + 3 * U * H;              // [3][U]  w_last
  float* b1 = wl + al4(3 * U);             // [3][U]  b_hh1
  float* W2i = b1 + al4(3 * U);            // [3][U][H]   (L == 2)
  float* W2h = W2i + (p.L == 2 ? 3 * U * H : 0);
  float* b2i = W2h + (p.L == 2 ? 3 * U * H : 0);
  float* b2h = b2i + (p.L == 2 ? al4(3 * U) : 0);
  float* Wp = b2h + (p.L == 2 ? al4(3 * U) : 0);   // [P][H]
  float* bp = Wp + P * H;                  // [P]
  float* Wo = bp + al4(P);                 // [so_rows][256]
  float* bo = Wo + so_rows * PRE;          // [so_rows]
  float* xin = bo + al4(so_rows);          // [H][B]  layer input (transposed)
  float* hin = xin + al4(B * H);           // [H][B]  (layer 2: previous h2, transposed)
  float* prev = hin + (p.L == 2 ? al4(B * H) : 0);   // [256][B]
  float* lg = prev + al4(B * PRE);         // [S][B] logits
  float* lastx = lg + al4(B * max(S, 1));  // [B]
  float* dg = lastx + al4(B);              // [DG][B] dot products (GRU gates: rows [0,3U) input, [3U,6U) hidden)
  float* part = dg + al4(p.DG * B);        // [THREADS] K-split partial sums

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
    copy_g2s_T(xin, p.hbuf + cur * hstride, B, H, tid);           // h1(t-1)
    __syncthreads();
    if (nu == U) {
      dots(W1, 3 * U, xin, B, H, dg, part, tid);            // rows [g*U + u]
    } else {         // ragged last CTA: its rows are not contiguous across gates - gate by gate
      for (int g = 0; g < 3; ++g) dots(W1 + g * U * H, nu, xin, B, H, dg + g * U * B, part, tid);
    }
    for (int task = tid; task < nu * B; task += THREADS) {
      const int u = task / B, b = task - u * B;
      const float lx = lastx[b];
      const float* gx = p.gx + ((size_t)b * 3 * H + u0 + u) * p.T + t;
      const float gr = gx[0] + wl[0 * U + u] * lx, gz = gx[(size_t)H * p.T] + wl[1 * U + u] * lx,
                  gn = gx[(size_t)2 * H * p.T] + wl[2 * U + u] * lx;
      const float r = sigm(gr + dg[(0 * U + u) * B + b] + b1[0 * U + u]);
      const float z = sigm(gz + dg[(1 * U + u) * B + b] + b1[1 * U + u]);
      const float n = tanhf(gn + r * (dg[(2 * U + u) * B + b] + b1[2 * U + u]));
      p.hbuf[nxt * hstride + (size_t)b * H + u0 + u] = (1.f - z) * n + z * xin[(u0 + u) * B + b];
    }
    grid.sync();
    // ================= layer 2 =================
    if (p.L == 2) {
      copy_g2s_T(xin, p.hbuf + nxt * hstride, B, H, tid);                       // h1(t)
      copy_g2s_T(hin, p.hbuf + cur * hstride + (size_t)B * H, B, H, tid);       // h2(t-1)
      __syncthreads();
      if (nu == U) {
        dots(W2i, 3 * U, xin, B, H, dg, part, tid);
        dots(W2h, 3 * U, hin, B, H, dg + 3 * U * B, part, tid);
      } else {
        for (int g = 0; g < 3; ++g) {
          dots(W2i + g * U * H, nu, xin, B, H, dg + g * U * B, part, tid);
          dots(W2h + g * U * H, nu, hin, B, H, dg + (3 + g) * U * B, part, tid);
        }
      }
      for (int task = tid; task < nu * B; task += THREADS) {
        const int u = task / B, b = task - u * B;
        const float* di = dg;
        const float* dh = dg + 3 * U * B;
        const float r = sigm(di[(0 * U + u) * B + b] + b2i[0 * U + u] + dh[(0 * U + u) * B + b] + b2h[0 * U + u]);
        const float z = sigm(di[(1 * U + u) * B + b] + b2i[1 * U + u] + dh[(1 * U + u) * B + b] + b2h[1 * U + u]);
        const float n = tanhf(di[(2 * U + u) * B + b] + b2i[2 * U + u] + r * (dh[(2 * U + u) * B + b] + b2h[2 * U + u]));
        p.hbuf[nxt * hstride + (size_t)B * H + (size_t)b * H + u0 + u] = (1.f - z) * n + z * hin[(u0 + u) * B + b];
      }
      grid.sync();
    }
    // ================= pre-output: tanh(W h + b) =================
    copy_g2s_T(xin, p.hbuf + nxt * hstride + (size_t)(p.L - 1) * B * H, B, H, tid);   // top layer's h(t)
    __syncthreads();
    dots(Wp, np, xin, B, H, dg, part, tid);
    for (int task = tid; task < np * B; task += THREADS) {
      const int r = task / B, b = task - r * B;
      p.prebuf[(size_t)b * PRE + p0 + r] = tanhf(dg[task] + bp[r]);
    }
    grid.sync();
    // ================= output layer + sampling head =================
    copy_g2s_T(prev, p.prebuf, B, PRE, tid);
    __syncthreads();
    if (small_head) {
      dots(Wo, ns, prev, B, PRE, lg, part, tid);          // lg[s*B + b]
    } else {
      dots(Wo, ns, prev, B, PRE, dg, part, tid);          // ns <= ceil(256/G) rows: fits dg
      for (int task = tid; task < ns * B; task += THREADS) {
        const int r = task / B, b = task - r * B;
        p.logits[(size_t)b * S + s0 + r] = dg[task] + bo[r];
      }
      grid.sync();
      for (int i = tid; i < B * S; i += THREADS) {         // [B][S] -> lg[s*B + b]
        const int b = i / S, k = i - b * S;
        lg[k * B + b] = p.logits[i];
      }
      __syncthreads();
    }
    if (p.head == HEAD_MOL || p.head == HEAD_GM) {
      if (tid < B) {
        const int b = tid;
        float y[32];
        for (int k = 0; k < S; ++k) y[k] = lg[k * B + b] + bo[k];     // small head: bias added here
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
        const float* u = p.draws + ((size_t)t * B + b) * S;
        float bs = -__builtin_inff();
        int bi = 0x7fffffff;
        for (int k = lane; k < S; k += 32) {
          const float s = lg[k * B + b] - logf(-logf(u[k]));
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
