// Output-head kernels (reference cube/networks/loss.py).  Element-wise, HBM-bound: one pass,
// coalesced, grid-stride.  The mu-law tables live in __constant__ memory and are GENERATED from
// the reference's float32 torch path (oracle/make_goldens.py -> mulaw_tables.inc) so the integer
// codes are bit-exact with it for every float32 input.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cube {

#include "mulaw_tables.inc"  // static const float MULAW_EDGES_H[255], MULAW_DECODE_H[256]

__constant__ float c_mulaw_edges[256];   // [255] = +inf sentinel
__constant__ float c_mulaw_decode[256];

inline cudaError_t upload_mulaw_tables() {
  float e[256];
  for (int i = 0; i < 255; ++i) e[i] = MULAW_EDGES_H[i];
  e[255] = __builtin_inff();
  cudaError_t err = cudaMemcpyToSymbol(c_mulaw_edges, e, sizeof(e));
  if (err != cudaSuccess) return err;
  return cudaMemcpyToSymbol(c_mulaw_decode, MULAW_DECODE_H, sizeof(float) * 256);
}

// code(x) = #{k : edge_k <= x}  (edges ascending)  == MULAWOutput.encode(x), loss.py:236-254
__global__ void mulaw_encode_kernel(const float* __restrict__ x, long long* __restrict__ q, long long n) {
  __shared__ float e[256];
  e[threadIdx.x] = c_mulaw_edges[threadIdx.x];
  __syncthreads();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = x[i];
    int lo = 0;  // number of edges <= v, binary search over 255 entries (8 steps)
#pragma unroll
    for (int step = 128; step >= 1; step >>= 1) {
      const int k = lo + step;
      if (k <= 255 && e[k - 1] <= v) lo = k;
    }
    q[i] = lo;
  }
}

__global__ void mulaw_decode_kernel(const long long* __restrict__ q, float* __restrict__ x, long long n) {
  __shared__ float d[256];
  d[threadIdx.x] = c_mulaw_decode[threadIdx.x];
  __syncthreads();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long c = q[i];
    if (c >= 0 && c < 256) {
      x[i] = d[c];
    } else {  // outside the codebook: evaluate the reference formula (loss.py:264-268)
      const float y = ((float)c / 255.f) * 2.f - 1.f;
      const float a = (expf(fabsf(y) * log1pf(255.f)) - 1.f) / 255.f;
      x[i] = y > 0.f ? a : (y < 0.f ? -a : 0.f);
    }
  }
}

// RAWOutput.encode: clip(((x+1)/2)*255, 0, 255).long()   (loss.py:293-295) - IEEE ops, bit-exact
__global__ void raw_encode_kernel(const float* __restrict__ x, long long* __restrict__ q, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float v = __fmul_rn(__fdiv_rn(__fadd_rn(x[i], 1.0f), 2.0f), 255.0f);
    v = fminf(fmaxf(v, 0.f), 255.f);
    q[i] = (long long)v;
  }
}

// RAWOutput.decode: ((q/255) - 0.5) * 2   (loss.py:297-299)
__global__ void raw_decode_kernel(const long long* __restrict__ q, float* __restrict__ x, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    x[i] = __fmul_rn(__fsub_rn(__fdiv_rn((float)q[i], 255.0f), 0.5f), 2.0f);
  }
}

// MOLOutput.sample (loss.py:176-199): Gumbel-max over the mixture logits, select mean/log-scale,
// logistic inverse-CDF, clamp.  One thread per sample; y rows are 3*nr_mix contiguous floats.
__global__ void mol_sample_kernel(const float* __restrict__ y, const float* __restrict__ u_mix,
                                  const float* __restrict__ u_x, float* __restrict__ x, long long n,
                                  int nr_mix, float log_scale_min, float temperature) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float* yr = y + i * 3 * nr_mix;
    const float* ur = u_mix + i * nr_mix;
    int best = 0;
    float bs = -__builtin_inff();
    for (int k = 0; k < nr_mix; ++k) {
      const float s = yr[k] - logf(-logf(ur[k] * temperature));
      if (s > bs) { bs = s; best = k; }  // first maximum wins, as torch.max does
    }
    const float mean = yr[nr_mix + best];
    const float ls = fmaxf(yr[2 * nr_mix + best], log_scale_min);
    const float u = u_x[i];
    float v = mean + expf(ls) * (logf(u) - logf(1.f - u));
    x[i] = fminf(fmaxf(v, -1.f), 1.f);
  }
}

// GaussianOutput.sample (loss.py:50-52): mean + (eps*0.8)*exp(log_std)
__global__ void gaussian_sample_kernel(const float* __restrict__ y, const float* __restrict__ eps,
                                       float* __restrict__ x, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float2 ml = reinterpret_cast<const float2*>(y)[i];
    x[i] = ml.x + __fmul_rn(eps[i], 0.8f) * expf(ml.y);
  }
}

// Categorical sample in Gumbel-max form: one warp per row of C logits, shuffle arg-max reduction.
__global__ void categorical_sample_kernel(const float* __restrict__ logits, const float* __restrict__ u,
                                          long long* __restrict__ idx, long long n, int C) {
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = warp; r < n; r += nwarps) {
    float bs = -__builtin_inff();
    int bi = 0x7fffffff;
    for (int k = lane; k < C; k += 32) {
      const float s = logits[r * C + k] - logf(-logf(u[r * C + k]));
      if (s > bs) { bs = s; bi = k; }
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
      const float os = __shfl_xor_sync(0xffffffffu, bs, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (os > bs || (os == bs && oi < bi)) { bs = os; bi = oi; }
    }
    if (lane == 0) idx[r] = bi;
  }
}

// cube/api.py:65  np.asarray(audio * 32767, dtype=np.int16): fp32 multiply, truncate toward zero
__global__ void wav_to_int16_kernel(const float* __restrict__ w, int16_t* __restrict__ o, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    o[i] = (int16_t)(int)__fmul_rn(w[i], 32767.0f);
  }
}

}  // namespace cube
