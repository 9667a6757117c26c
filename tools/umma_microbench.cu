// Microbenchmark: cycles per tcgen05.mma for the operand shapes the vocoder kernels use (shared-memory A and B, fp32
// accumulate in TMEM), one CTA (or CTA pair) per SM, back-to-back issue by one thread, completion via tcgen05.commit.
// Answers the question DESIGN.md 5.3 rests on: what does a narrow (N = 32/64) MMA cost when its math is 16-32 cycles?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/umma_microbench tools/umma_microbench.cu && /tmp/umma_microbench
#include <cstdio>
#include <cuda_runtime.h>
#include "../tts_cube_b200/csrc/tc_conv.cuh"

using namespace cube::tc;

enum { K_F16 = 0, K_F8_SW32 = 1, K_F8_SW64 = 2 };

// kind, N, pair (cta_group::2, M = 256) ; R MMAs per measured batch
template <int KIND, int N, bool PAIR>
__global__ void __launch_bounds__(128, 1) bench_kernel(unsigned long long* out, int R) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* a_s = smem;                 // 16 KB: [128 rows][64 B] (+ slack)
  uint8_t* b_s = smem + 16384;         // 32 KB: [256 rows][64 B]
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t crank = PAIR ? cluster_ctarank() : 0u;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) { if constexpr (PAIR) tmem_alloc2(&tmem_slot, 512); else tmem_alloc(&tmem_slot, 512); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  constexpr int M = PAIR ? 256 : 128;
  unsigned long long best = ~0ull;
  for (int rep = 0; rep < 4; ++rep) {
    long long t0 = 0;
    if (threadIdx.x == 0 && crank == 0) {
      t0 = clock64();
      for (int i = 0; i < R; ++i) {
        const uint32_t ko = (i & 1) * 32;                      // alternate the two K steps of a 64-byte row like the real kernels
        if constexpr (KIND == K_F16) {
          const uint64_t da = make_desc(smem_u32(a_s) + ko), db = make_desc(smem_u32(b_s) + ko);
          if constexpr (PAIR) umma_f16_2(tmem, da, db, make_idesc(N, M), i > 0); else umma_f16(tmem, da, db, make_idesc(N, M), i > 0);
        } else if constexpr (KIND == K_F8_SW32) {              // 32-byte rows, SWIZZLE_32B (what the Q8 passes read)
          const uint64_t da = make_desc32(smem_u32(a_s)), db = make_desc32(smem_u32(b_s));
          if constexpr (PAIR) {
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(da),
                         "l"(db), "r"(make_idesc_f8(N, M, 0)), "r"((uint32_t)(i > 0)) : "memory");
          } else {
            umma_f8(tmem, da, db, make_idesc_f8(N, M, 0), i > 0);
          }
        } else {                                               // 8-bit operands in 64-byte rows (SWIZZLE_64B), K = 32 = half a row
          const uint64_t da = make_desc(smem_u32(a_s) + ko), db = make_desc(smem_u32(b_s) + ko);
          if constexpr (PAIR) {
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(da),
                         "l"(db), "r"(make_idesc_f8(N, M, 0)), "r"((uint32_t)(i > 0)) : "memory");
          } else {
            umma_f8(tmem, da, db, make_idesc_f8(N, M, 0), i > 0);
          }
        }
      }
      if constexpr (PAIR) umma_commit_2(&bar); else umma_commit(&bar);
    }
    if (threadIdx.x == 0) {
      mbar_wait(&bar, rep & 1);
      if (crank == 0) {
        const unsigned long long dt = (unsigned long long)(clock64() - t0);
        if (dt < best) best = dt;
      }
    }
    __syncthreads();
    if constexpr (PAIR) cluster_sync_all();
  }
  if (threadIdx.x == 0 && crank == 0) atomicMax(out, best);      // slowest CTA of the grid
  tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();
  if (threadIdx.x < 32) { tc_fence_after(); if constexpr (PAIR) tmem_dealloc2(tmem, 512); else tmem_dealloc(tmem, 512); }
}

template <int KIND, int N, bool PAIR>
static void run(const char* name, int sms, double a_bytes, double b_bytes, double math_cycles) {
  unsigned long long* d;
  cudaMalloc(&d, 8);
  const int R = 512;
  const size_t smem = 16384 + 32768 + 2048;
  cudaFuncSetAttribute(bench_kernel<KIND, N, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int grid_sms : {1, sms}) {
    cudaMemset(d, 0, 8);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(PAIR ? 2 * ((grid_sms + 1) / 2) : grid_sms);
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    if (PAIR) {
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
    }
    unsigned long long* dp = d;
    int r = R;
    cudaError_t e = cudaLaunchKernelEx(&cfg, bench_kernel<KIND, N, PAIR>, dp, r);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    unsigned long long h = 0;
    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { printf("%-44s grid %3d: %s\n", name, (int)cfg.gridDim.x, cudaGetErrorString(e)); cudaGetLastError(); continue; }
    const double cyc = (double)h / R;
    printf("%-44s grid %3d: %7.1f cycles/MMA  (math %.0f; A %.0f B + B %.0f B per CTA -> %.1f B/clk operand fetch)\n", name, (int)cfg.gridDim.x, cyc,
           math_cycles, a_bytes, b_bytes, (a_bytes + b_bytes) / cyc);
  }
  cudaFree(d);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  printf("%s, %d SMs, clock %d kHz; 512 back-to-back MMAs, slowest CTA, best of 4\n", p.name, p.multiProcessorCount, p.clockRate);
  const int S = p.multiProcessorCount;
  // fp16, M = 128 (one CTA): K = 16 -> A = 128 rows x 32 B = 4 KB, B = N rows x 32 B
  run<K_F16, 256, false>("f16 M128 N256 K16 (SW64 rows)", S, 4096, 8192, 128);
  run<K_F16, 128, false>("f16 M128 N128 K16", S, 4096, 4096, 64);
  run<K_F16, 64, false>("f16 M128 N64  K16", S, 4096, 2048, 32);
  run<K_F16, 32, false>("f16 M128 N32  K16", S, 4096, 1024, 16);
  // 8-bit, K = 32: A = 128 x 32 B, B = N x 32 B
  run<K_F8_SW32, 256, false>("f8f6f4 M128 N256 K32, 32-byte rows (SW32)", S, 4096, 8192, 128);
  run<K_F8_SW64, 256, false>("f8f6f4 M128 N256 K32, 64-byte rows (SW64)", S, 4096, 8192, 128);
  // CTA pair, M = 256: each CTA fetches its 128 rows of A and HALF of B
  run<K_F16, 256, true>("f16 cta_group::2 M256 N256 K16", S, 4096, 4096, 128);
  run<K_F16, 128, true>("f16 cta_group::2 M256 N128 K16", S, 4096, 2048, 64);
  run<K_F8_SW32, 256, true>("f8f6f4 cta_group::2 M256 N256 K32 (SW32)", S, 4096, 4096, 128);
  run<K_F8_SW64, 256, true>("f8f6f4 cta_group::2 M256 N256 K32 (SW64)", S, 4096, 4096, 128);
  return 0;
}
