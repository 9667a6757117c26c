// Device log-mel spectrogram (the feature contract on the input side of the vocoders).
//   reference A: hifigan/meldataset.py:50-74  mel_spectrogram()  - reflect pad (n_fft-hop)/2, torch.stft(center=False,
//                periodic hann), sqrt(re^2 + im^2 + 1e-9), mel_basis @ spec, log(clamp(., 1e-5))          -> [B, M, F]
//   reference B: cube/io_utils/vocoder.py:54-62,78-98  MelVocoder.melspectrogram() - optional pre-emphasis
//                (lfilter [1,-0.97]), librosa.stft(center=True: reflect pad n_fft/2, hann), |.|, mel basis,
//                log10(max(1e-5, .)), transposed                                                            -> [F, M]
// Both are  frames x (window * DFT basis)  -> magnitude -> mel basis -> log:  two small dense contractions
// (2.1 MFLOP per frame = 8 KFLOP per audio sample, against 1-25 MFLOP per sample for the vocoders), so this is a plain
// fp32 FFMA kernel: one CTA = 32 frames x all bins, the signal span of the tile staged once in shared memory with the
// padding rule applied at staging time, the windowed cos/sin tables ([n][bin], window folded in, built in double on the
// host) streamed from L2 as float4 with lanes along bins, magnitudes parked in shared memory, mel + log in the same
// launch.  fp32 accumulation over n_fft terms keeps the log-mel within ~1e-5 of the reference (tests: 2e-4).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cube {
namespace mel {

constexpr int FT = 32;        // frames per CTA
constexpr int BT = 128;       // bins per pass (32 lanes x 4)
constexpr int THREADS = 256;  // 8 warps x 4 frames

struct MelParams {
  const float* wav;        // [B][Tmax]
  const int* n_samples;    // [B] valid samples per utterance, or null (= Tmax)
  const float* cosT;       // [n_fft][KB]  window[n] * cos(2 pi k n / n_fft), zero for k >= n_bins
  const float* sinT;       // [n_fft][KB]
  const float* basis;      // [n_mels][n_bins]
  const int* k_lo;         // [n_mels] first / one-past-last bin with a non-zero weight (triangular filters)
  const int* k_hi;
  float* out;              // layout 0: [B][n_mels][Fmax]; layout 1: [B][Fmax][n_mels]
  int B, Tmax, Fmax;
  int n_fft, hop, n_bins, KB, n_mels;
  int pad_left, pad_right;
  int layout, log10_out, pad_mode;   // pad_mode 0: reflect at the utterance's own ends, 1: zeros
  float mag_eps, floor_val, pad_value, preemph;
};

// frames an utterance of L samples yields (torch.stft on the reflect-padded signal; reflect needs pad < L)
__host__ __device__ inline int n_frames_of(int L, int n_fft, int hop, int pl, int pr, int pad_mode = 0) {
  if (pad_mode == 0 && (L <= pl || L <= pr)) return 0;
  if (L < 1 || L + pl + pr < n_fft) return 0;
  return 1 + (L + pl + pr - n_fft) / hop;
}

__global__ void __launch_bounds__(THREADS) melspec_kernel(const MelParams p) {
  extern __shared__ float smem[];
  const int span = (FT - 1) * p.hop + p.n_fft;
  const int span4 = (span + 3) & ~3;
  float* sig = smem;                        // [span4]
  const int mpitch = p.KB + 1;              // odd pitch: lanes = frames read it conflict-free in the mel pass
  float* mag = smem + span4;                // [FT][mpitch]
  const int b = blockIdx.y, f0 = blockIdx.x * FT;
  const int L = p.n_samples ? min(p.n_samples[b], p.Tmax) : p.Tmax;
  const int Fb = n_frames_of(L, p.n_fft, p.hop, p.pad_left, p.pad_right, p.pad_mode);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (f0 < Fb) {
    // ---- stage the signal span of this tile: reflect padding at the utterance's own ends, pre-emphasis ----
    const float* x = p.wav + (size_t)b * p.Tmax;
    const int g0 = f0 * p.hop - p.pad_left;
    for (int i = threadIdx.x; i < span4; i += THREADS) {
      int g = g0 + i;
      if (p.pad_mode == 0) {
        if (g < 0) g = -g;
        if (g >= L) g = 2 * (L - 1) - g;
      }
      float v = 0.f;
      if (i < span && g >= 0 && g < L) {
        v = x[g];
        if (p.preemph != 0.f && g > 0) v -= p.preemph * x[g - 1];
      }
      sig[i] = v;
    }
    __syncthreads();
    // ---- DFT: thread = 4 frames (warp) x 4 bins (lane), bins in passes of 128 ----
    for (int k0 = 0; k0 < p.KB; k0 += BT) {
      float re[4][4], im[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) re[i][j] = im[i][j] = 0.f;
      const float* cp = p.cosT + k0 + lane * 4;
      const float* sp = p.sinT + k0 + lane * 4;
      const float* s0 = sig + (warp * 4) * p.hop;
      for (int n = 0; n < p.n_fft; n += 4) {
        float xv[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {   // hop and n are multiples of 4: 16-byte aligned warp-broadcast reads
          const float4 t = *reinterpret_cast<const float4*>(s0 + i * p.hop + n);
          xv[i][0] = t.x; xv[i][1] = t.y; xv[i][2] = t.z; xv[i][3] = t.w;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float4 c = __ldg(reinterpret_cast<const float4*>(cp + (size_t)(n + u) * p.KB));
          const float4 s = __ldg(reinterpret_cast<const float4*>(sp + (size_t)(n + u) * p.KB));
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float a = xv[i][u];
            re[i][0] = fmaf(a, c.x, re[i][0]); re[i][1] = fmaf(a, c.y, re[i][1]);
            re[i][2] = fmaf(a, c.z, re[i][2]); re[i][3] = fmaf(a, c.w, re[i][3]);
            im[i][0] = fmaf(a, s.x, im[i][0]); im[i][1] = fmaf(a, s.y, im[i][1]);
            im[i][2] = fmaf(a, s.z, im[i][2]); im[i][3] = fmaf(a, s.w, im[i][3]);
          }
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
          mag[(warp * 4 + i) * mpitch + k0 + lane * 4 + j] = sqrtf(re[i][j] * re[i][j] + im[i][j] * im[i][j] + p.mag_eps);
    }
    __syncthreads();
  }
  // ---- mel basis + log: lane = frame, warp = mel rows warp, warp+8, ... ----
  const int f = f0 + lane;
  if (f >= p.Fmax) return;
  for (int m = warp; m < p.n_mels; m += THREADS / 32) {
    float v = p.pad_value;
    if (f < Fb) {
      const float* bm = p.basis + (size_t)m * p.n_bins;
      const float* mg = mag + lane * mpitch;
      float acc = 0.f;
      const int k1 = p.k_hi[m];
      for (int k = p.k_lo[m]; k < k1; ++k) acc = fmaf(__ldg(bm + k), mg[k], acc);
      acc = fmaxf(acc, p.floor_val);
      v = p.log10_out ? log10f(acc) : logf(acc);
    }
    if (p.layout == 0) p.out[((size_t)b * p.n_mels + m) * p.Fmax + f] = v;
    else p.out[((size_t)b * p.Fmax + f) * p.n_mels + m] = v;
  }
}

inline size_t smem_bytes(int n_fft, int hop, int KB) {
  const int span4 = ((FT - 1) * hop + n_fft + 3) & ~3;
  return (size_t)(span4 + FT * (KB + 1)) * sizeof(float);
}

}  // namespace mel
}  // namespace cube
