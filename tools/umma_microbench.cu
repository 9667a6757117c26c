// Microbenchmark: cycles per tcgen05.mma for the operand shapes the vocoder kernels use (shared-memory A and B, fp32
// accumulate in TMEM), one CTA (or CTA pair) per SM, back-to-back issue by one thread, completion via tcgen05.commit.
// Answers the question DESIGN.md 5.3 rests on: what does a narrow (N = 32/64) MMA cost when its math is 16-32 cycles?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/umma_microbench tools/umma_microbench.cu && /tmp/umma_microbench
#include <cstdio>
#include <cuda_runtime.h>
#include "../tts_cube_b200/csrc/tc_conv.cuh"

using namespace cube::tc;

enum { K_F16 = 0, K_F8_SW32 = 1, K_F8_SW64 = 2 };

// kind, N, pair (cta_group::2, M = 256) ; R MMAs per measured batch.
// a_step: bytes between the A tiles of consecutive MMAs (0 = the same tile every time; 8192 = cycle through 4 tiles like the
// sub-tiles of a real kernel - rules out any reuse of an unchanged operand); a_off: byte offset of the A start address
// (192 = 3 rows: a conv tap's row-offset descriptor, not aligned to the 8-row swizzle atom); b_step likewise for B.
template <int KIND, int N, bool PAIR>
__global__ void __launch_bounds__(128, 1) bench_kernel(unsigned long long* out, int R, int a_step, int a_off, int b_step, int b_off = 0) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* a_s = smem;                 // 48 KB: 4 tiles of [128 rows][64 B] + slack for row offsets
  uint8_t* b_s = smem + 49152 + b_off; // 64 KB: up to 2 tiles of [256 rows][64 B] (+ slack); b_off moves it (placement sweep)
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t crank = PAIR ? cluster_ctarank() : 0u;
  for (int i = threadIdx.x; i < (49152 + 65536) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
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
        const uint32_t ao = (uint32_t)((i >> 1) & 3) * a_step + a_off, bo = (uint32_t)((i >> 3) & 1) * b_step;
        if constexpr (KIND == K_F16) {
          const uint64_t da = make_desc(smem_u32(a_s) + ao + ko), db = make_desc(smem_u32(b_s) + bo + ko);
          if constexpr (PAIR) umma_f16_2(tmem, da, db, make_idesc(N, M), i > 0); else umma_f16(tmem, da, db, make_idesc(N, M), i > 0);
        } else if constexpr (KIND == K_F8_SW32) {              // 32-byte rows, SWIZZLE_32B (what the Q8 passes read)
          const uint64_t da = make_desc32(smem_u32(a_s) + (ao >> 1)), db = make_desc32(smem_u32(b_s) + (bo >> 1));
          if constexpr (PAIR) {
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(da),
                         "l"(db), "r"(make_idesc_f8(N, M, 0)), "r"((uint32_t)(i > 0)) : "memory");
          } else {
            umma_f8(tmem, da, db, make_idesc_f8(N, M, 0), i > 0);
          }
        } else {                                               // 8-bit operands in 64-byte rows (SWIZZLE_64B), K = 32 = half a row
          const uint64_t da = make_desc(smem_u32(a_s) + ao + ko), db = make_desc(smem_u32(b_s) + bo + ko);
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
static void run(const char* name, int sms, double a_bytes, double b_bytes, double math_cycles, int a_step = 0, int a_off = 0, int b_step = 0,
                int b_off = 0, bool only_full = false) {
  unsigned long long* d;
  cudaMalloc(&d, 8);
  const int R = 512;
  const size_t smem = 49152 + 65536 + 2048 + 110 * 1024;
  cudaFuncSetAttribute(bench_kernel<KIND, N, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int grid_sms : {1, sms}) {
    if (only_full && grid_sms == 1) continue;
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
    cudaError_t e = cudaLaunchKernelEx(&cfg, bench_kernel<KIND, N, PAIR>, dp, r, a_step, a_off, b_step, b_off);
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

// LEAN issue loop: the 8 descriptor pairs are computed before the clock starts and the loop body is nothing but 8 tcgen05.mma
// (accumulate = 1): separates what the TENSOR CORE needs per MMA from what ONE ISSUING THREAD needs to build descriptors.
// nthreads_issue = 2: a second warp issues the same stream into another accumulator (columns 256..) at the same time.
template <int N>
__global__ void __launch_bounds__(128, 1) lean_kernel(unsigned long long* out, int R, int two_issuers) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* a_s = smem;
  uint8_t* b_s = smem + 49152;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  for (int i = threadIdx.x; i < (49152 + 65536) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
  if (threadIdx.x == 0) { mbar_init(&bar, two_issuers ? 2 : 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc(&tmem_slot, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  unsigned long long best = ~0ull;
  const bool issuer = threadIdx.x == 0 || (two_issuers && threadIdx.x == 32);
  const uint32_t d = tmem + (threadIdx.x == 32 ? 256 : 0);
  uint64_t da[8], db[8];
  for (int j = 0; j < 8; ++j) {
    da[j] = make_desc(smem_u32(a_s) + (uint32_t)((j >> 1) & 3) * 8192 + (j & 1) * 32 + (threadIdx.x == 32 ? 32768 : 0));
    db[j] = make_desc(smem_u32(b_s) + (j & 1) * 32);
  }
  constexpr uint32_t idesc = make_idesc(N, 128);
  for (int rep = 0; rep < 4; ++rep) {
    __syncthreads();
    long long t0 = clock64();
    if (issuer) {
      for (int i = 0; i < R; i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(da[j]),
                       "l"(db[j]), "r"(idesc) : "memory");
      }
      umma_commit(&bar);
    }
    if (threadIdx.x == 0) {
      mbar_wait(&bar, rep & 1);
      const unsigned long long dt = (unsigned long long)(clock64() - t0);
      if (dt < best) best = dt;
    }
  }
  if (threadIdx.x == 0) atomicMax(out, best);
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

template <int N>
static void run_lean(const char* name, int sms, int two) {
  unsigned long long* d;
  cudaMalloc(&d, 8);
  cudaMemset(d, 0, 8);
  const int R = 512;
  const size_t smem = 49152 + 65536 + 2048;
  cudaFuncSetAttribute(lean_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  lean_kernel<N><<<sms, 128, smem>>>(d, R, two);
  cudaError_t e = cudaDeviceSynchronize();
  unsigned long long h = 0;
  cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) { printf("%-52s %s\n", name, cudaGetErrorString(e)); return; }
  printf("%-52s grid %3d: %7.1f cycles per MMA per issuer (%d issuer%s: %.1f cycles per MMA for the SM)\n", name, sms, (double)h / R, two ? 2 : 1,
         two ? "s" : "", (double)h / R / (two ? 2 : 1));
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
  // the same shapes with a DIFFERENT A tile for consecutive MMAs (sub-tiles of a real kernel), and a tap-style row offset
  printf("-- A cycles through 4 tiles (8 KB apart), B through 2\n");
  run<K_F16, 32, false>("f16 M128 N32  K16, A varies", S, 4096, 1024, 16, 8192, 0, 0);
  run<K_F16, 32, false>("f16 M128 N32  K16, A and B vary", S, 4096, 1024, 16, 8192, 0, 4096);
  run<K_F16, 64, false>("f16 M128 N64  K16, A varies", S, 4096, 2048, 32, 8192, 0, 0);
  run<K_F16, 256, false>("f16 M128 N256 K16, A and B vary", S, 4096, 8192, 128, 8192, 0, 32768);
  run<K_F16, 256, true>("f16 cta_group::2 M256 N256 K16, A, B vary", S, 4096, 4096, 128, 8192, 0, 16384);
  printf("-- A start address 3 rows (192 B) into the tile: a conv tap's row-offset descriptor\n");
  run<K_F16, 32, false>("f16 M128 N32  K16, A varies, +3 rows", S, 4096, 1024, 16, 8192, 192, 0);
  run<K_F16, 64, false>("f16 M128 N64  K16, A varies, +3 rows", S, 4096, 2048, 32, 8192, 192, 0);
  run<K_F16, 256, false>("f16 M128 N256 K16, A varies, +3 rows", S, 4096, 8192, 128, 8192, 192, 0);
  run<K_F16, 32, false>("f16 M128 N32  K16, A varies, +8 rows", S, 4096, 1024, 16, 8192, 512, 0);
  printf("-- lean issue loop (descriptors precomputed, 8 MMAs unrolled), one and two issuing warps\n");
  run_lean<32>("f16 M128 N32  K16, lean", S, 0);
  run_lean<32>("f16 M128 N32  K16, lean, 2 issuers", S, 1);
  run_lean<64>("f16 M128 N64  K16, lean", S, 0);
  run_lean<64>("f16 M128 N64  K16, lean, 2 issuers", S, 1);
  run_lean<128>("f16 M128 N128 K16, lean", S, 0);
  run_lean<128>("f16 M128 N128 K16, lean, 2 issuers", S, 1);
  run_lean<256>("f16 M128 N256 K16, lean", S, 0);
  run_lean<256>("f16 M128 N256 K16, lean, 2 issuers", S, 1);
  if (0) printf("-- placement sweep: A at byte a_off, B at byte 49152 + b_off of the (1 KB aligned) dynamic shared memory\n");
  char nm[96];
  if (0) for (int boff : {-49152 + 4096, -49152 + 8192, -49152 + 16384, -49152 + 24576, -32768 + 16384, 0, 8192, 16384, 32768, 49152, 65536, 81920, 98304, 110592}) {
    snprintf(nm, sizeof(nm), "f16 N32, A at 0, B at %6d", 49152 + boff);
    run<K_F16, 32, false>(nm, S, 4096, 1024, 16, 0, 0, 0, boff, true);
  }
  if (0) for (int aoff : {1024, 2048, 4096, 8192, 16384, 32768}) {
    snprintf(nm, sizeof(nm), "f16 N32, A at %6d, B at 49152", aoff);
    run<K_F16, 32, false>(nm, S, 4096, 1024, 16, 0, aoff, 0, 0, true);
  }
  if (0) for (int boff : {-49152 + 16384, 0, 65536}) {
    snprintf(nm, sizeof(nm), "f16 N128, A at 0, B at %6d", 49152 + boff);
    run<K_F16, 128, false>(nm, S, 4096, 4096, 64, 0, 0, 0, boff, true);
    snprintf(nm, sizeof(nm), "f16 N64 [hi|lo], A at 0, B at %6d", 49152 + boff);
    run<K_F16, 64, false>(nm, S, 4096, 2048, 32, 0, 0, 0, boff, true);
  }
  return 0;
}
