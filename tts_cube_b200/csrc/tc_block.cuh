// One ClariNet residual block in ONE tcgen05 kernel (sm_100a):
//     f|g = W1 . [h(t-2d), h(t-d), h(t), c(t)]      GEMM1  [128 rows x 512]   K = 3*128 + 80
//     o   = tanh(f) * sigmoid(g)                     stays ON CHIP (TMEM -> registers -> swizzled smem A tiles)
//     r|s = W2 . o                                   GEMM2  [128 rows x 256]   K = 256
//     h'  = (h + r) * sqrt(.5)  (fp16 planes, PING-PONG buffers: other tiles still read h through their taps),
//     skip (+)= s  (fp32)
// The unfused pair (tc_conv_kernel GATE + RESSKIP) writes o (1.8 GB per block at B=8 x 10 s) to HBM and reads it
// back, and its tensor-bound half (gate) and HBM-bound half (res/skip) run back to back; here o never leaves the
// SM and the residual/skip traffic of tile i hides under the MMAs of tile i+1.
//
// TMEM: two 256-column regions X, Y.  Per tile (parity p swaps the roles): GEMM1 n-tile 0 -> R0, GEMM1 n-tile 1 ->
// R1, GEMM2 (both K halves) -> R0 again once its gate epilogue has drained it.  The next tile starts in this tile's
// R1, which is free long before this tile's res/skip epilogue runs.
// smem: a ring of stages (A hi/lo 8 KB each by TMA + this CTA's weight rows hi/lo by bulk copy; GEMM2 stages carry weights
// only), a 64 KB buffer for one K-half of o as four K-major SWIZZLE_64B A tiles per plane, 6 KB of folded (de-scale, bias)
// pairs.  Roles as in tc_conv_kernel: warp 0 TMA, warp 1 MMA, warps 2-17 epilogue.
#pragma once
#include "tc_conv.cuh"

namespace cube {
namespace tc {

// One kernel text, two tilings (template parameter PAIR):
//   PAIR = false: one CTA per 128-row tile, tcgen05 cta_group::1 (M = 128); 3 ring stages of 48 KB (A 16 KB + weights 32 KB).
//   PAIR = true : a 2-CTA cluster per 256-row tile, cta_group::2 (M = 256).  Each CTA stages its own 128 rows of A (and
//                 keeps its own rows of o, h, skip: the epilogues are the same code) but only HALF of every weight tile -
//                 the pair's tensor cores exchange the halves - so a stage is 32 KB (4 stages) and the L2 -> shared-memory
//                 bytes per output row drop by a third.  The leader's (rank 0) MMA thread issues every MMA and owns the
//                 accumulator / o-buffer hand-shakes: the peer's epilogue warps arrive remotely (relaxed, cluster scope -
//                 the release form compiles to MEMBAR.ALL.GPU) on the leader's barriers, tcgen05.commit multicasts
//                 completions to both CTAs.
//   AONCE (PAIR only): a GEMM1 stage carries the A chunk ONCE and the weight halves of BOTH 256-column n-tiles (48 KB, 3 stages);
//                 the two n-tiles accumulate side by side in the two TMEM regions.  A is no longer loaded twice per tile
//                 (L2 -> shared memory bytes per tile 1.09 -> 0.85 MB, 47 instead of 64 B per MMA cycle, 3 k instead of 2 k cycles of
//                 latency covered by the ring); the price is that both regions are busy until the gate epilogues have drained
//                 them, so the next tile's GEMM1 cannot start under this tile's epilogues.
template <bool PAIR, bool AONCE = false> struct BlkCfg {
  static_assert(PAIR || !AONCE, "AONCE is a mode of the CTA-pair tiling");
  static constexpr int NST = AONCE ? 3 : (PAIR ? 4 : 3);                              // ring stages
  static constexpr int B_BYTES = (PAIR ? 128 : 256) * BK * 2;                         // this CTA's weight rows, one fp16 plane
  static constexpr int STG = 2 * A_TILE_BYTES + (AONCE ? 4 : 2) * B_BYTES;            // 32 / 48 KB
  static constexpr int O_BYTES = 2 * 4 * A_TILE_BYTES;                                // 4 chunks x (hi, lo) = 64 KB
  static constexpr int SMEM = NST * STG + O_BYTES + 1024 + 512 + 768 * 8;
  static constexpr int TILE_ROWS = PAIR ? 2 * BM : BM;
};
constexpr int BLK_O_BYTES = BlkCfg<false>::O_BYTES;
constexpr int BLK_SMEM = BlkCfg<false>::SMEM;
constexpr int PAIR_SMEM = BlkCfg<true>::SMEM;
constexpr int AONCE_SMEM = BlkCfg<true, true>::SMEM;
static_assert(BK == 32, "tc_block_kernel is written for 32-channel K chunks (SWIZZLE_64B)");

struct BlockParams {
  CUtensorMap tmH, tmC;          // h16 [2B][T][128], c16 [2B][T][80]            (boxes of 32 ch x 128 rows)
  CUtensorMap tmHin32, tmHout32; // block input / output residual stream, boxes of 32 ch x 32 rows (one epilogue warp)
  const __half* W1;              // gate images   [2 n-tiles][nch1][2][256*32]
  const __half* W2;              // res/skip images [1][8][2][256*32]
  const float* inv1; const float* bias1;   // [512]  (n-tile major: [nt*256 + col])
  const float* inv2; const float* bias2;   // [256]
  int taps, dil, off0;           // h segment: taps at rows t + off0 + j*dil
  int h_chunks, c_chunks, c_last_ksteps;   // K chunks per tap of h (4), of c (3), K steps in c's last chunk (1)
  int B, T, t_tiles;
  const int* lens;
  const __half* h_in16;          // residual stream of the block INPUT (same tensor tmH maps; other tiles' gate taps read it,
                                 // so it must not be updated in place: the unfused pair had a kernel boundary in between)
  __half* h_out16;               // residual stream of the block OUTPUT (ping-pong buffer)
  float* skip; int skip_set; __half* skip16; float scale;
  // ---- Q8 build only: GEMM1's two correction passes on 8-bit operands ----
  CUtensorMap tmH8, tmC8;        // uint8 planes [2B][T][C]: plane 0 = e4m3(a_hi), plane 1 = e5m2(16 a_lo); boxes 32 B x 128 rows
  const uint8_t* W1q;            // gate images [2 n-tiles][nch1][32 KB]: w_hi fp16 (SW64) | e4m3(w_lo) | e4m3(w_hi/16) (SW32)
  uint8_t* h8_out;               // 8-bit planes of the block OUTPUT (the next block's A operand)
  // ---- L2 prefetch of the NEXT tile's A rows by the (mostly waiting) epilogue warps: the residual stream and the conditioning
  // do not fit in L2 between launches, so a ring stage's A boxes come from DRAM while its weights come from L2; pulled into
  // L2 a tile ahead (plain prefetch.global.L2 through the LSU: nothing queues in the TMA engine, unlike the bulk prefetch
  // that was tried in round 1), both arrive with L2 latency.  0 = off.
  int prefetch_next;
  const __half* c_in16;          // conditioning planes [2][B][T][c_ch] (what tmC maps)
  const uint8_t* h8_in;          // Q8: 8-bit planes of the block input (what tmH8 maps), conditioning (tmC8)
  const uint8_t* c8_in;
  int c_ch;                      // channels of the conditioning planes (80)
  unsigned long long* stats;     // STATS build only: wait-cycle counters of CTA 0 (CUBE_BLOCK_STATS=1), 24 slots
  int lean;                      // 1: the MMA thread runs the short instruction stream (see "lean issue" in the kernel)
};

// wait on an mbarrier, charging the cycles to a counter slot in the instrumented build
#define BLK_WAIT(bar, par, slot)                                   \
  do {                                                             \
    if (STATS) {                                                   \
      const long long _t0 = clock64();                             \
      mbar_wait(bar, par);                                         \
      st_acc[slot] += clock64() - _t0;                             \
    } else {                                                       \
      mbar_wait(bar, par);                                         \
    }                                                              \
  } while (0)

__device__ __forceinline__ void umma_f8_2(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
// epilogue -> leader's MMA thread: local arrive on the leader, remote (relaxed, cluster scope) from the peer
__device__ __forceinline__ void pair_arrive(uint64_t* bar, uint32_t crank) {
  if (crank == 0) mbar_arrive(bar); else mbar_arrive_remote(bar, 0);
}

// the MMA / commit of the tiling
template <bool PAIR>
__device__ __forceinline__ void blk_mma_f16(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  if constexpr (PAIR) umma_f16_2(d, da, db, idesc, acc); else umma_f16(d, da, db, idesc, acc);
}
template <bool PAIR>
__device__ __forceinline__ void blk_mma_f8(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  if constexpr (PAIR) umma_f8_2(d, da, db, idesc, acc); else umma_f8(d, da, db, idesc, acc);
}
template <bool PAIR>
__device__ __forceinline__ void blk_commit(uint64_t* bar) {
  if constexpr (PAIR) umma_commit_2(bar); else umma_commit(bar);
}

// PAIR: the peer's MMA warp relays "my stage has landed" (wait own barrier -> remote arrive on the leader's `pfull`).  Letting the
// peer's loads complete on the LEADER's barrier directly was tried in round 2 and hangs: a 1-D cp.async.bulk (the weight
// images) may only signal an mbarrier of the CTA whose shared memory it fills (only the .cta_group::2 TENSOR form may
// signal the peer CTA's barrier).
// STATS: instrumented build (CUBE_BLOCK_STATS=1): CTA 0's cycles per barrier wait.
// 18 warps: the SM sub-partitions hold 5,5,4,4 of them, so 16384/5 -> 96 registers per thread is the hardware cap
template <bool PAIR, bool Q8, bool STATS = false, bool AONCE = false>
__global__ void __launch_bounds__(NUM_THREADS, 1) tc_block_kernel(const __grid_constant__ BlockParams p) {
  long long st_acc[STATS ? 8 : 1] = {0};
  const long long st_begin = STATS ? clock64() : 0;
  (void)st_acc; (void)st_begin;
  constexpr int BN = 256;
  constexpr int B_BYTES = BlkCfg<PAIR, AONCE>::B_BYTES;   // this CTA's weight rows (all 256, or its HALF of the pair's tile), one fp16 plane
  constexpr int NST = BlkCfg<PAIR, AONCE>::NST;
  constexpr int STG = BlkCfg<PAIR, AONCE>::STG;
  constexpr int NPASS = AONCE ? 1 : 2;             // GEMM1 passes over the K chunks (one per n-tile, or one for both)
  constexpr int NTP = AONCE ? 2 : 1;               // n-tiles per pass
  constexpr int TILE_ROWS = BlkCfg<PAIR>::TILE_ROWS;
  constexpr int NARR = PAIR ? 2 * NUM_EPI_WARPS : NUM_EPI_WARPS;      // epilogue warps that arrive on the leader's hand-shake barriers
  const uint32_t crank = PAIR ? cluster_ctarank() : 0u;               // 0 = leader: issues the MMAs
  const int tile0 = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int tile_step = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* o_smem = smem + NST * STG;      // [4 chunks][hi 8 KB | lo 8 KB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(o_smem + BlkCfg<PAIR, AONCE>::O_BYTES);
  uint64_t* full = bars;                       // [4] this CTA's stage has landed
  uint64_t* empty = bars + NST;                // [4] (commit multicast: both CTAs)
  uint64_t* pfull = bars + 2 * NST;            // [4] leader only: the PEER's stage has landed (relayed by its MMA warp)
  uint64_t* acc_full = bars + 3 * NST;         // [2] region holds a finished accumulator        (commit multicast -> both epilogues)
  uint64_t* acc_free = acc_full + 2;           // [2] leader only: region drained by BOTH CTAs   (2 x 16 warps)
  uint64_t* o_full = acc_free + 2;             // [1] leader only: o half staged in BOTH CTAs    (2 x 16 warps)
  uint64_t* o_free = o_full + 1;               // [1] GEMM2 has read the o half                  (MMA -> epilogue)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_free + 1);
  uint64_t* hbar = o_free + 2;                 // [16] per epilogue warp: its residual rows have landed in shared memory
  float2* sb1 = reinterpret_cast<float2*>(reinterpret_cast<uint8_t*>(bars) + 512);   // [512] gate: folded exp2 factors
  float2* sb2 = sb1 + 512;                                                           // [256] res/skip

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = p.t_tiles * p.B;
  const int nch1 = p.taps * p.h_chunks + p.c_chunks;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.tmH);
    prefetch_tmap(&p.tmC);
    for (int s = 0; s < NST; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); mbar_init(&pfull[s], 1); }
    for (int r = 0; r < 2; ++r) { mbar_init(&acc_full[r], 1); mbar_init(&acc_free[r], NARR); }
    mbar_init(o_full, NARR);
    mbar_init(o_free, 1);
    for (int w = 0; w < NUM_EPI_WARPS; ++w) mbar_init(&hbar[w], 1);
    prefetch_tmap(&p.tmHin32);
    prefetch_tmap(&p.tmHout32);
    if (Q8) { prefetch_tmap(&p.tmH8); prefetch_tmap(&p.tmC8); }
    fence_barrier_init();
  }
  if (warp == 1) { if constexpr (PAIR) tmem_alloc2(tmem_slot, 512); else tmem_alloc(tmem_slot, 512); }
  {  // folded per-column constants, once per CTA (this CTA owns all 512 + 256 columns)
    constexpr float LOG2E = 1.4426950408889634f;
    for (int i = threadIdx.x; i < 512; i += NUM_THREADS) {
      const float k = (i & 255) < 128 ? 2.f * LOG2E : -LOG2E;     // filter cols -> e^{2f}, gate cols -> e^{-g}
      sb1[i] = make_float2(__ldg(p.inv1 + i) * k, __ldg(p.bias1 + i) * k);
    }
    for (int i = threadIdx.x; i < 256; i += NUM_THREADS) sb2[i] = make_float2(__ldg(p.inv2 + i), __ldg(p.bias2 + i));
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();    // both CTAs' barriers are initialised before any remote arrive
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = tile0; tile < total_tiles; tile += tile_step) {
        const int tt = tile % p.t_tiles, b = tile / p.t_tiles;
        const int t0 = tt * TILE_ROWS + (int)crank * BM;       // this CTA's 128 rows of the tile
        for (int pass = 0; pass < NPASS; ++pass) {             // GEMM1: one 256-column n-tile per pass, or (AONCE) both in one pass
          int chunk = 0;
          for (int tap = 0; tap <= p.taps; ++tap) {            // tap == p.taps: the conditioning segment
            const bool cond = tap == p.taps;
            const int row = cond ? t0 : t0 + p.off0 + tap * p.dil;
            const int ncc = cond ? p.c_chunks : p.h_chunks;
            for (int cc = 0; cc < ncc; ++cc, ++chunk, ++it) {
              const int st = it % NST;
              BLK_WAIT(&empty[st], ((it / NST) & 1) ^ 1, 0);
              uint8_t* sb = smem + st * STG;
              const uint32_t fbar = smem_u32(&full[st]);
              mbar_expect_tx(&full[st], STG);
              const CUtensorMap* tm = cond ? &p.tmC : &p.tmH;
              tma_load_3d_bar(sb, tm, fbar, cc * BK, row, b);
              if constexpr (Q8) {   // a_hi fp16 (8 KB) | e4m3(a_hi) (4 KB) | e5m2(16 a_lo) (4 KB)
                const CUtensorMap* tm8 = cond ? &p.tmC8 : &p.tmH8;
                tma_load_3d_bar(sb + A_TILE_BYTES, tm8, fbar, cc * BK, row, b);
                tma_load_3d_bar(sb + A_TILE_BYTES + A_TILE_BYTES / 2, tm8, fbar, cc * BK, row, p.B + b);
              } else {
                tma_load_3d_bar(sb + A_TILE_BYTES, tm, fbar, cc * BK, row, p.B + b);
              }
#pragma unroll
              for (int k = 0; k < NTP; ++k) {                  // this CTA's weight rows of the n-tile(s) of this pass
                const int nt = AONCE ? k : pass;
                uint8_t* wb = sb + 2 * A_TILE_BYTES + k * 2 * B_BYTES;
                if constexpr (Q8) {
                  // rows [128 crank, +128) of each of the three sub-images of the 32 KB chunk image
                  const uint8_t* wq = p.W1q + ((size_t)nt * nch1 + chunk) * 32768;
                  bulk_load_bar(wb, wq + crank * B_BYTES, B_BYTES, fbar);
                  bulk_load_bar(wb + B_BYTES, wq + 16384 + crank * (B_BYTES / 2), B_BYTES / 2, fbar);
                  bulk_load_bar(wb + B_BYTES + B_BYTES / 2, wq + 24576 + crank * (B_BYTES / 2), B_BYTES / 2, fbar);
                } else {
                  const __half* wc = p.W1 + ((size_t)nt * nch1 + chunk) * 2 * (BN * BK) + (size_t)crank * (BN / 2) * BK /* crank = 0 without PAIR */;
                  bulk_load_bar(wb, wc, B_BYTES, fbar);
                  bulk_load_bar(wb + B_BYTES, wc + BN * BK, B_BYTES, fbar);
                }
              }
            }
          }
        }
        for (int ch = 0; ch < 8; ++ch, ++it) {                 // GEMM2: weights only (its A operand is o, on chip)
          const int st = it % NST;
          BLK_WAIT(&empty[st], ((it / NST) & 1) ^ 1, 1);
          uint8_t* sb = smem + st * STG;
          const uint32_t fbar = smem_u32(&full[st]);
          mbar_expect_tx(&full[st], 2 * B_BYTES);
          const __half* wc = p.W2 + (size_t)ch * 2 * (BN * BK) + (size_t)crank * (BN / 2) * BK /* crank = 0 without PAIR */;
          bulk_load_bar(sb + 2 * A_TILE_BYTES, wc, B_BYTES, fbar);
          bulk_load_bar(sb + 2 * A_TILE_BYTES + B_BYTES, wc + BN * BK, B_BYTES, fbar);
        }
      }
      if (STATS && p.stats && blockIdx.x == 0) {
        atomicAdd(p.stats + 0, (unsigned long long)st_acc[0]);
        atomicAdd(p.stats + 1, (unsigned long long)st_acc[1]);
        atomicAdd(p.stats + 2, (unsigned long long)(clock64() - st_begin));
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    if (crank == 0 && p.lean) {
      // Lean issue.  ONE thread issues every MMA of the CTA (pair), and its instruction stream - not the tensor core - set the
      // pace: ncu shows the issuing warp busy all the time (hardly any barrier wait) at ~6 cycles per dependent instruction, and
      // the generic loop below spends ~23 instructions per MMA (two descriptors built from addresses: shift, mask, or, 64-bit
      // pack; the per-thread-to-uniform-register hand-over ELECT / 5 x R2UR.BROADCAST / branch in front of every MMA; a run-time
      // K-step loop; modulo ring indices) = ~175 cycles per MMA against the 128.5 the tensor core needs.
      // Here the WHOLE warp runs the loop: waits and counters are warp-uniform, so offsets and descriptors are computed on the
      // uniform datapath and the MMAs of a stage issue back to back from uniform registers; one elected lane issues them and
      // the commits.  ONE descriptor per swizzle mode for the start of shared memory, every operand = that + a byte offset
      // (desc_add), K steps unrolled, ring stage / phase as counters.  Same MMAs, same order, same barriers as the generic loop.
      constexpr uint32_t idesc = make_idesc(BN, TILE_ROWS);
      constexpr uint32_t idesc8_0 = make_idesc_f8(BN, TILE_ROWS, 0), idesc8_1 = make_idesc_f8(BN, TILE_ROWS, 1);
      (void)idesc8_0; (void)idesc8_1;
      const uint64_t D16 = make_desc(smem_u32(smem)), D32 = make_desc32(smem_u32(smem));
      (void)D32;
      constexpr uint32_t O_OFF = NST * STG;                     // o_smem - smem
      constexpr uint32_t A8 = A_TILE_BYTES, AH = A_TILE_BYTES / 2, BH = B_BYTES / 2;
      (void)AH; (void)BH;
      const bool leader = elect_one();
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      uint32_t st = 0, ph = 0, titer = 0;                       // ring stage, parity of its current phase
      uint32_t free_ph[2] = {0, 0};
      uint32_t ofull_ph = 0;
      for (int tile = tile0; tile < total_tiles; tile += tile_step, ++titer) {
        const int r0 = titer & 1, r1 = r0 ^ 1;
        for (int pass = 0; pass < NPASS; ++pass) {
#pragma unroll
          for (int k = 0; k < NTP; ++k) {
            const int rg = (AONCE ? k : pass) == 0 ? r0 : r1;
            BLK_WAIT(&acc_free[rg], (free_ph[rg] & 1) ^ 1, AONCE ? k : pass);
            ++free_ph[rg];
          }
          tc_fence_after();
          uint32_t accumulate = 0;
          for (int ch = 0; ch < nch1; ++ch) {                   // the producer's order: taps of h, then the conditioning
            BLK_WAIT(&full[st], ph, 2);
            if constexpr (PAIR) BLK_WAIT(&pfull[st], ph, 2);
            tc_fence_after();
            const uint32_t so = st * STG;
            const int ksteps = ch != nch1 - 1 ? BK / 16 : p.c_last_ksteps;
            if (leader) {
#pragma unroll
              for (int k = 0; k < NTP; ++k) {
                const int rg = (AONCE ? k : pass) == 0 ? r0 : r1;
                const uint32_t d_tmem = tmem_u + rg * BN;
                const uint32_t bo = so + 2 * A_TILE_BYTES + k * 2 * B_BYTES;
                if constexpr (Q8) {
                  blk_mma_f16<PAIR>(d_tmem, desc_add(D16, so), desc_add(D16, bo), idesc, accumulate);
                  if (ksteps > 1) blk_mma_f16<PAIR>(d_tmem, desc_add(D16, so + 32), desc_add(D16, bo + 32), idesc, 1);
                  blk_mma_f8<PAIR>(d_tmem, desc_add(D32, so + A8), desc_add(D32, bo + B_BYTES), idesc8_0, 1);
                  blk_mma_f8<PAIR>(d_tmem, desc_add(D32, so + A8 + AH), desc_add(D32, bo + B_BYTES + BH), idesc8_1, 1);
                } else {
#pragma unroll
                  for (int ks = 0; ks < BK / 16; ++ks) {
                    if (ks < ksteps) {
                      const uint32_t ko = ks * 32;
                      blk_mma_f16<PAIR>(d_tmem, desc_add(D16, so + ko), desc_add(D16, bo + ko), idesc, ks ? 1u : accumulate);
                      blk_mma_f16<PAIR>(d_tmem, desc_add(D16, so + ko), desc_add(D16, bo + B_BYTES + ko), idesc, 1);
                      blk_mma_f16<PAIR>(d_tmem, desc_add(D16, so + A8 + ko), desc_add(D16, bo + ko), idesc, 1);
                    }
                  }
                }
              }
              blk_commit<PAIR>(&empty[st]);
            }
            accumulate = 1;
            if (++st == NST) { st = 0; ph ^= 1; }
          }
          if (leader) {
#pragma unroll
            for (int k = 0; k < NTP; ++k) blk_commit<PAIR>(&acc_full[(AONCE ? k : pass) == 0 ? r0 : r1]);
          }
        }
        {  // GEMM2 into r0 (drained by the gate epilogue of n-tile 0): K half kh uses the o half the epilogue staged
          BLK_WAIT(&acc_free[r0], (free_ph[r0] & 1) ^ 1, 3);
          ++free_ph[r0];
          const uint32_t d_tmem = tmem_u + r0 * BN;
          uint32_t accumulate = 0;
          for (int kh = 0; kh < 2; ++kh) {
            BLK_WAIT(o_full, ofull_ph & 1, 4 + kh);
            ++ofull_ph;
            tc_fence_after();
#pragma unroll 1
            for (int c4 = 0; c4 < 4; ++c4) {
              BLK_WAIT(&full[st], ph, 6);
              if constexpr (PAIR) BLK_WAIT(&pfull[st], ph, 6);
              tc_fence_after();
              const uint32_t ao = O_OFF + c4 * 2 * A_TILE_BYTES, bo = st * STG + 2 * A_TILE_BYTES;
              if (leader) {
#pragma unroll
                for (int ks = 0; ks < BK / 16; ++ks) {
                  const uint32_t ko = ks * 32;
                  blk_mma_f16<PAIR>(d_tmem, desc_add(D16, ao + ko), desc_add(D16, bo + ko), idesc, ks ? 1u : accumulate);
                  blk_mma_f16<PAIR>(d_tmem, desc_add(D16, ao + ko), desc_add(D16, bo + B_BYTES + ko), idesc, 1);
                  blk_mma_f16<PAIR>(d_tmem, desc_add(D16, ao + A8 + ko), desc_add(D16, bo + ko), idesc, 1);
                }
                blk_commit<PAIR>(&empty[st]);
              }
              accumulate = 1;
              if (++st == NST) { st = 0; ph ^= 1; }
            }
            if (leader) blk_commit<PAIR>(o_free);          // the o half may be overwritten
          }
          if (leader) blk_commit<PAIR>(&acc_full[r0]);     // r|s accumulator complete
        }
      }
      if (STATS && p.stats && blockIdx.x == 0 && leader) {
        for (int i = 0; i < 7; ++i) atomicAdd(p.stats + 3 + i, (unsigned long long)st_acc[i]);
        atomicAdd(p.stats + 10, (unsigned long long)(clock64() - st_begin));
      }
    } else if (lane == 0 && crank != 0) {             // (crank != 0 only exists with PAIR)
      // peer CTA: it issues no MMA; this thread relays "my stage has landed" to the leader
      uint32_t it = 0;
      for (int tile = tile0; tile < total_tiles; tile += tile_step)
        for (int n = NPASS * nch1 + 8; n > 0; --n, ++it) {
          const int st = it % NST;
          mbar_wait(&full[st], (it / NST) & 1);
          mbar_arrive_remote(&pfull[st], 0);
        }
    } else if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BN, TILE_ROWS);    // M = 128, or 256 across the pair
      uint32_t it = 0, titer = 0;
      uint32_t free_ph[2] = {0, 0};        // completed-phase counters of acc_free[r] this thread has consumed
      uint32_t ofull_ph = 0;
      for (int tile = tile0; tile < total_tiles; tile += tile_step, ++titer) {
        const int r0 = titer & 1, r1 = r0 ^ 1;      // region roles of this tile
        for (int pass = 0; pass < NPASS; ++pass) {
          // the region(s) must have been drained by their previous user (first use of each region: passes at once)
#pragma unroll
          for (int k = 0; k < NTP; ++k) {
            const int rg = (AONCE ? k : pass) == 0 ? r0 : r1;
            BLK_WAIT(&acc_free[rg], (free_ph[rg] & 1) ^ 1, AONCE ? k : pass);
            ++free_ph[rg];
          }
          tc_fence_after();
          uint32_t accumulate = 0;
          for (int tap = 0; tap <= p.taps; ++tap) {
            const bool cond = tap == p.taps;
            const int ncc = cond ? p.c_chunks : p.h_chunks;
            for (int cc = 0; cc < ncc; ++cc, ++it) {
              const int st = it % NST;
              BLK_WAIT(&full[st], (it / NST) & 1, 2);
              if constexpr (PAIR) BLK_WAIT(&pfull[st], (it / NST) & 1, 2);
              tc_fence_after();
              const uint32_t a_hi = smem_u32(smem + st * STG), a_lo = a_hi + A_TILE_BYTES;
              const int ksteps = (cond && cc == ncc - 1) ? p.c_last_ksteps : (BK / 16);
#pragma unroll
              for (int k = 0; k < NTP; ++k) {
                const int rg = (AONCE ? k : pass) == 0 ? r0 : r1;
                const uint32_t d_tmem = tmem_base + rg * BN;
                const uint32_t b_hi = a_hi + 2 * A_TILE_BYTES + k * 2 * B_BYTES, b_lo = b_hi + B_BYTES;
                uint32_t acc_k = accumulate;
                if constexpr (Q8) {
                  // hi*hi in fp16 (K = 16 per MMA), then the two 2^-11-weight corrections as ONE 8-bit MMA each (K = 32):
                  // e4m3(a_hi) x e4m3(w_lo)  and  e5m2(16 a_lo) x e4m3(w_hi / 16)  -> 4 MMAs per chunk instead of 6
                  for (int ks = 0; ks < ksteps; ++ks) {
                    blk_mma_f16<PAIR>(d_tmem, make_desc(a_hi + ks * 32), make_desc(b_hi + ks * 32), idesc, acc_k);
                    acc_k = 1;
                  }
                  const uint32_t a8_hi = a_hi + A_TILE_BYTES, a8_lo = a8_hi + A_TILE_BYTES / 2;
                  const uint32_t b8_lo = b_hi + B_BYTES, b8_hi = b8_lo + B_BYTES / 2;
                  blk_mma_f8<PAIR>(d_tmem, make_desc32(a8_hi), make_desc32(b8_lo), make_idesc_f8(BN, TILE_ROWS, 0), 1);
                  blk_mma_f8<PAIR>(d_tmem, make_desc32(a8_lo), make_desc32(b8_hi), make_idesc_f8(BN, TILE_ROWS, 1), 1);
                } else {
                  for (int ks = 0; ks < ksteps; ++ks) {
                    const uint32_t ko = ks * 32;
                    blk_mma_f16<PAIR>(d_tmem, make_desc(a_hi + ko), make_desc(b_hi + ko), idesc, acc_k);
                    blk_mma_f16<PAIR>(d_tmem, make_desc(a_hi + ko), make_desc(b_lo + ko), idesc, 1);
                    blk_mma_f16<PAIR>(d_tmem, make_desc(a_lo + ko), make_desc(b_hi + ko), idesc, 1);
                    acc_k = 1;
                  }
                }
              }
              accumulate = 1;
              blk_commit<PAIR>(&empty[st]);
            }
          }
#pragma unroll
          for (int k = 0; k < NTP; ++k) blk_commit<PAIR>(&acc_full[(AONCE ? k : pass) == 0 ? r0 : r1]);
        }
        // GEMM2 into r0 (drained by the gate epilogue of n-tile 0): K half kh uses the o half the epilogue staged
        {
          BLK_WAIT(&acc_free[r0], (free_ph[r0] & 1) ^ 1, 3);
          ++free_ph[r0];
          const uint32_t d_tmem = tmem_base + r0 * BN;
          uint32_t accumulate = 0;
          for (int kh = 0; kh < 2; ++kh) {
            BLK_WAIT(o_full, ofull_ph & 1, 4 + kh);
            ++ofull_ph;
            tc_fence_after();
            for (int c4 = 0; c4 < 4; ++c4, ++it) {
              const int st = it % NST;
              BLK_WAIT(&full[st], (it / NST) & 1, 6);
              if constexpr (PAIR) BLK_WAIT(&pfull[st], (it / NST) & 1, 6);
              tc_fence_after();
              const uint32_t a_hi = smem_u32(o_smem + c4 * 2 * A_TILE_BYTES), a_lo = a_hi + A_TILE_BYTES;
              const uint32_t b_hi = smem_u32(smem + st * STG) + 2 * A_TILE_BYTES, b_lo = b_hi + B_BYTES;
              for (int ks = 0; ks < BK / 16; ++ks) {
                const uint32_t ko = ks * 32;
                blk_mma_f16<PAIR>(d_tmem, make_desc(a_hi + ko), make_desc(b_hi + ko), idesc, accumulate);
                blk_mma_f16<PAIR>(d_tmem, make_desc(a_hi + ko), make_desc(b_lo + ko), idesc, 1);
                blk_mma_f16<PAIR>(d_tmem, make_desc(a_lo + ko), make_desc(b_hi + ko), idesc, 1);
                accumulate = 1;
              }
              blk_commit<PAIR>(&empty[st]);
            }
            blk_commit<PAIR>(o_free);          // the o half may be overwritten
          }
          blk_commit<PAIR>(&acc_full[r0]);     // r|s accumulator complete
        }
      }
      if (STATS && p.stats && blockIdx.x == 0) {
        for (int i = 0; i < 7; ++i) atomicAdd(p.stats + 3 + i, (unsigned long long)st_acc[i]);
        atomicAdd(p.stats + 10, (unsigned long long)(clock64() - st_begin));
      }
    }
  } else {
    // =========================== epilogue (warps 2..17) ===========================
    const int q = warp & 3;                 // TMEM lane quarter
    const int grp = (warp - 2) >> 2;        // 0..3
    const int row = q * 32 + lane;
    uint32_t titer = 0;
    uint32_t full_ph[2] = {0, 0};           // uses of acc_full[r] consumed so far
    uint32_t ofree_ph = 0;
    for (int tile = tile0; tile < total_tiles; tile += tile_step, ++titer) {
      const int tt = tile % p.t_tiles, b = tile / p.t_tiles;
      const int t = tt * TILE_ROWS + (int)crank * BM + row;
      const int r0 = titer & 1, r1 = r0 ^ 1;
      const int len = p.lens ? min(p.lens[b], p.T) : p.T;
      const bool in_range = t < p.T, valid = t < len;
      if (p.prefetch_next && tile + tile_step < total_tiles) {
        // the A rows of this CTA's NEXT tile -> L2: per tap 128 rows x (h: 2 planes; 256 B fp16 / 128 B 8-bit per row),
        // one 128-byte line per thread and step; the conditioning rows of the tile likewise
        const int ntile = tile + tile_step;
        const int ntt = ntile % p.t_tiles, nb = ntile / p.t_tiles;
        const int nt0 = ntt * TILE_ROWS + (int)crank * BM;
        const int e = (int)threadIdx.x - 64;               // 0..511
        const int rr = e & 127, part = e >> 7;             // row of the box, which quarter of the row's bytes
        const size_t hplane = (size_t)p.B * p.T * 128;
        for (int tap = 0; tap < p.taps; ++tap) {
          const int row = nt0 + p.off0 + tap * p.dil + rr;
          if (row >= 0 && row < p.T) {
            const size_t ro = (size_t)nb * p.T + row;
            if constexpr (Q8) {      // fp16 hi plane: 2 lines; 8-bit planes: 1 line each
              if (part < 2) prefetch_l2(p.h_in16 + ro * 128 + part * 64);
              else prefetch_l2(p.h8_in + (size_t)(part - 2) * hplane + ro * 128);
            } else {                 // fp16 hi and lo planes: 2 lines each
              prefetch_l2(p.h_in16 + (size_t)(part >> 1) * hplane + ro * 128 + (part & 1) * 64);
            }
          }
        }
        const int crow = nt0 + rr;
        if (crow < p.T && p.c_in16) {
          const size_t ro = (size_t)nb * p.T + crow;
          const size_t cplane = (size_t)p.B * p.T * p.c_ch;
          if constexpr (Q8) {
            if (part < 2) prefetch_l2(p.c_in16 + ro * p.c_ch + part * 64);                       // 160 B of fp16: two lines
            else if (p.c8_in) prefetch_l2(p.c8_in + (size_t)(part - 2) * cplane + ro * p.c_ch);  // 80 B per 8-bit plane
          } else {
            prefetch_l2(p.c_in16 + (size_t)(part >> 1) * cplane + ro * p.c_ch + (part & 1) * 64);
          }
        }
      }
      // a flow's last block reads the running skip total: pull this warp's slice into L2 while the MMAs run
      if (in_range && p.skip16 && !p.skip_set && (lane & 7) == 0) {
        const float* s0 = p.skip + ((size_t)b * 128 + grp * 32) * p.T + t;
#pragma unroll 8
        for (int j = 0; j < 32; ++j) prefetch_l2(s0 + (size_t)j * p.T);
      }
      // ---------------- gate epilogue of n-tile 0 (region r0) and n-tile 1 (region r1) ----------------
      for (int nt = 0; nt < 2; ++nt) {
        const int rg = nt == 0 ? r0 : r1;
        BLK_WAIT(&acc_full[rg], full_ph[rg] & 1, nt);
        ++full_ph[rg];
        const long long g_t0 = STATS ? clock64() : 0;
        tc_fence_after();
        const uint32_t taddr = tmem_base + rg * BN + ((uint32_t)(q * 32) << 16);
        const float2* sb = sb1 + nt * 256;
        // this warp: output channels [32*grp, +32) of this n-tile = K chunk `grp` of the o half
        uint8_t* otile = o_smem + grp * 2 * A_TILE_BYTES;
        uint32_t hi2[2][8], lo2[2][8];
#pragma unroll
        for (int ci = 0; ci < 2; ++ci) {
          const int cc = ci * 16;
          uint32_t f[16], g[16];
          tmem_ld16(taddr + grp * 32 + cc, f);
          tmem_ld16(taddr + 128 + grp * 32 + cc, g);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; j += 2) {
            float o[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const float2 sf = sb[grp * 32 + cc + j + u], sg = sb[128 + grp * 32 + cc + j + u];
              const float a = fminf(fmaxf(fmaf(__uint_as_float(f[j + u]), sf.x, sf.y), -40.f), 40.f);
              const float e = fminf(fmaf(__uint_as_float(g[j + u]), sg.x, sg.y), 60.f);
              const float E1 = ex2_fast(a), E2 = ex2_fast(e);
              o[u] = valid ? (E1 - 1.f) * rcp_fast((E1 + 1.f) * (1.f + E2)) : 0.f;
            }
            split16x2(o[0], o[1], hi2[ci][j >> 1], lo2[ci][j >> 1]);
          }
        }
        // the accumulator region is drained: hand it back before waiting for the o buffer
        tc_fence_before();
        __syncwarp();
        if (lane == 0) pair_arrive(&acc_free[rg], crank);
        // the o buffer is free once GEMM2 has consumed the previous half (first half ever: passes at once)
        if (STATS) st_acc[5] += clock64() - g_t0;       // gate math (TMEM load .. region handed back)
        BLK_WAIT(o_free, (ofree_ph & 1) ^ 1, 2 + nt);
        ++ofree_ph;
        if (nt == 0) {   // the residual rows the previous tile stored from this warp's piece of the o buffer have been read
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          __syncwarp();
        }
        // K-major SWIZZLE_64B A tile [128 rows][32 ch]: 16-byte piece c16 of row r lives at piece c16 ^ ((r>>1)&3)
#pragma unroll
        for (int ci = 0; ci < 2; ++ci) {
#pragma unroll
          for (int v = 0; v < 2; ++v) {
            const int c16 = ci * 2 + v;
            const uint32_t off = row * 64 + ((c16 ^ ((row >> 1) & 3)) << 4);
            *reinterpret_cast<uint4*>(otile + off) = make_uint4(hi2[ci][4 * v], hi2[ci][4 * v + 1], hi2[ci][4 * v + 2], hi2[ci][4 * v + 3]);
            *reinterpret_cast<uint4*>(otile + A_TILE_BYTES + off) = make_uint4(lo2[ci][4 * v], lo2[ci][4 * v + 1], lo2[ci][4 * v + 2], lo2[ci][4 * v + 3]);
          }
        }
        fence_proxy_async();                 // generic-proxy smem writes -> visible to the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) pair_arrive(o_full, crank);
      }
      // ---------------- res/skip epilogue (GEMM2 result in region r0) ----------------
      // Every warp owns 32 residual channels [32 grp, +32) and the 32 skip channels of the same index, for its 32 rows.
      // Measured with CUBE_BLOCK_STATS=1, the read-modify-write of these two streams (not the MMAs) paced the tile
      // loop: per-row 16-byte global accesses cost 32 L1 wavefronts per instruction, and the next tile's gate
      // epilogues queue behind them.  So:
      //  * residual: the warp's [32 rows][32 ch] hi/lo boxes travel by TMA through the warp's own 2 x 2 KB pieces of
      //    the o buffer (idle once GEMM2 has read it) - load, update in place in the swizzled layout, store;
      //  * skip (+)= s is a fire-and-forget red.global.add.f32, one 128-byte line per warp instruction (one thread
      //    per element and launch, so no contention, and the same single round-to-nearest fp32 add a load/add/store
      //    would do - REDG flushes subnormals, a < 1.2e-38 difference - without a load on the critical path).
      {
        const uint32_t taddr_rs = tmem_base + r0 * BN + ((uint32_t)(q * 32) << 16) + grp * 32;
        const size_t plane = (size_t)p.B * p.T * 128;
        float* sp0 = p.skip + ((size_t)b * 128 + grp * 32) * p.T + t;
        const bool last_block = p.skip16 != nullptr;
        const int ew = warp - 2;
        uint8_t* piece_hi = o_smem + grp * 2 * A_TILE_BYTES + q * 2048;      // rows [32 q, +32) of K chunk grp
        uint8_t* piece_lo = piece_hi + A_TILE_BYTES;
        const int tw = tt * TILE_ROWS + (int)crank * BM + q * 32;            // first time step of this warp's rows
        float old[32];             // only a flow's last block needs the running skip total (relu -> fp16 planes)
        if (last_block) {
#pragma unroll
          for (int j = 0; j < 32; ++j) old[j] = (in_range && !p.skip_set) ? __ldcs(sp0 + (size_t)j * p.T) : 0.f;
        }
        BLK_WAIT(&acc_full[r0], full_ph[r0] & 1, 4);
        ++full_ph[r0];
        const long long r_t0 = STATS ? clock64() : 0;
        tc_fence_after();
        if (lane == 0) {           // GEMM2 is complete: the o buffer is idle, fetch the residual rows into this warp's pieces
          mbar_expect_tx(&hbar[ew], 2 * 2048);
          tma_load_3d(piece_hi, &p.tmHin32, &hbar[ew], grp * 32, tw, b);
          tma_load_3d(piece_lo, &p.tmHin32, &hbar[ew], grp * 32, tw, p.B + b);
        }
        // ---- skip columns [128 + 32 grp, +32) ----
#pragma unroll
        for (int ci = 0; ci < 2; ++ci) {
          const int cc = grp * 32 + ci * 16;
          uint32_t racc[16];
          tmem_ld16(taddr_rs + 128 + ci * 16, racc);
          tmem_ld_wait();
          if (last_block) {
            uint32_t hi2[8], lo2[8];
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
              const float2 s0 = sb2[128 + cc + j], s1 = sb2[128 + cc + j + 1];
              const float y0 = old[ci * 16 + j] + fmaf(__uint_as_float(racc[j]), s0.x, s0.y);
              const float y1 = old[ci * 16 + j + 1] + fmaf(__uint_as_float(racc[j + 1]), s1.x, s1.y);
              split16x2(valid ? fmaxf(y0, 0.f) : 0.f, valid ? fmaxf(y1, 0.f) : 0.f, hi2[j >> 1], lo2[j >> 1]);
            }
            if (in_range) {
              __half* srow = p.skip16 + ((size_t)b * p.T + t) * 128 + cc;
#pragma unroll
              for (int v = 0; v < 2; ++v) {
                reinterpret_cast<uint4*>(srow)[v] = make_uint4(hi2[4 * v], hi2[4 * v + 1], hi2[4 * v + 2], hi2[4 * v + 3]);
                reinterpret_cast<uint4*>(srow + plane)[v] = make_uint4(lo2[4 * v], lo2[4 * v + 1], lo2[4 * v + 2], lo2[4 * v + 3]);
              }
            }
          } else if (in_range) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float2 s2 = sb2[128 + cc + j];
              const float y = valid ? fmaf(__uint_as_float(racc[j]), s2.x, s2.y) : 0.f;
              float* dst = sp0 + (size_t)(ci * 16 + j) * p.T;
              if (p.skip_set) __stcs(dst, y);
              else asm volatile("red.global.add.f32 [%0], %1;" ::"l"(dst), "f"(y) : "memory");
            }
          }
        }
        // ---- residual columns [32 grp, +32): h_out = (h_in + r) * sqrt(.5), in place in the swizzled pieces ----
        mbar_wait(&hbar[ew], titer & 1);
#pragma unroll
        for (int ci = 0; ci < 2; ++ci) {
          const int cc = grp * 32 + ci * 16;
          uint32_t racc[16];
          tmem_ld16(taddr_rs + ci * 16, racc);
          // K-major SWIZZLE_64B: 16-byte piece c16 of row r lives at piece c16 ^ ((r >> 1) & 3) of its 64-byte row
          const uint32_t off0 = lane * 64 + (((ci * 2) ^ ((lane >> 1) & 3)) << 4);
          const uint32_t off1 = lane * 64 + (((ci * 2 + 1) ^ ((lane >> 1) & 3)) << 4);
          uint4 hv[2], lv[2];
          hv[0] = *reinterpret_cast<const uint4*>(piece_hi + off0);
          hv[1] = *reinterpret_cast<const uint4*>(piece_hi + off1);
          lv[0] = *reinterpret_cast<const uint4*>(piece_lo + off0);
          lv[1] = *reinterpret_cast<const uint4*>(piece_lo + off1);
          tmem_ld_wait();
          if (ci == 1) {           // last TMEM read of this warp: hand the region back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) pair_arrive(&acc_free[r0], crank);
          }
          const uint32_t* hp = reinterpret_cast<const uint32_t*>(hv);
          const uint32_t* lp = reinterpret_cast<const uint32_t*>(lv);
          uint32_t hi2[8], lo2[8];
#pragma unroll
          for (int j = 0; j < 16; j += 2) {
            const float2 s0 = sb2[cc + j], s1 = sb2[cc + j + 1];
            const float2 oh = unpack_h2(hp[j >> 1]), ol = unpack_h2(lp[j >> 1]);
            const float v0 = fmaf(__uint_as_float(racc[j]), s0.x, s0.y), v1 = fmaf(__uint_as_float(racc[j + 1]), s1.x, s1.y);
            const float n0 = valid ? ((oh.x + ol.x) + v0) * p.scale : 0.f;
            const float n1 = valid ? ((oh.y + ol.y) + v1) * p.scale : 0.f;
            split16x2(n0, n1, hi2[j >> 1], lo2[j >> 1]);
          }
          if constexpr (Q8) {      // the next block's 8-bit A planes of these 16 channels (16 bytes per plane and row)
            if (in_range) {
              uint32_t qh[8], ql[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) q8_pair(hi2[j], lo2[j], qh[j], ql[j]);
              uint8_t* d8 = p.h8_out + ((size_t)b * p.T + t) * 128 + cc;
              *reinterpret_cast<uint4*>(d8) = make_uint4(qh[0] | (qh[1] << 16), qh[2] | (qh[3] << 16), qh[4] | (qh[5] << 16), qh[6] | (qh[7] << 16));
              *reinterpret_cast<uint4*>(d8 + plane) = make_uint4(ql[0] | (ql[1] << 16), ql[2] | (ql[3] << 16), ql[4] | (ql[5] << 16), ql[6] | (ql[7] << 16));
            }
          }
          *reinterpret_cast<uint4*>(piece_hi + off0) = make_uint4(hi2[0], hi2[1], hi2[2], hi2[3]);
          *reinterpret_cast<uint4*>(piece_hi + off1) = make_uint4(hi2[4], hi2[5], hi2[6], hi2[7]);
          *reinterpret_cast<uint4*>(piece_lo + off0) = make_uint4(lo2[0], lo2[1], lo2[2], lo2[3]);
          *reinterpret_cast<uint4*>(piece_lo + off1) = make_uint4(lo2[4], lo2[5], lo2[6], lo2[7]);
        }
        fence_proxy_async();       // generic-proxy writes -> visible to the TMA engine
        __syncwarp();
        if (lane == 0) {           // rows beyond T are clipped by the tensor map
          tma_store_3d(&p.tmHout32, piece_hi, grp * 32, tw, b);
          tma_store_3d(&p.tmHout32, piece_lo, grp * 32, tw, p.B + b);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        if (STATS) st_acc[6] += clock64() - r_t0;     // res/skip epilogue after the accumulator arrived
      }
    }
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // shared memory stays valid until the bulk engine has read it
    if (STATS && p.stats && blockIdx.x == 0 && lane == 0 && (warp == 2 || warp == 10)) {
      unsigned long long* o = p.stats + (warp == 2 ? 11 : 19);
      for (int i = 0; i < 8; ++i) atomicAdd(o + i, (unsigned long long)st_acc[i]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();    // no CTA of the pair exits while the other may still signal it
  if (warp == 1) {
    tc_fence_after();
    if constexpr (PAIR) tmem_dealloc2(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace tc
}  // namespace cube
