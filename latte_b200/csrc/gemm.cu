// tcgen05 GEMM with fused epilogues: out = epi(A[M,K] @ W[N,K]^T + bias)
//
// Replaces the nn.Linear calls inside a Latte TransformerBlock (reference models/latte.py:50 qkv, :75 proj,
// timm Mlp fc1/fc2 reached from :171,:180) together with the elementwise work that follows them:
//   EPI_BIAS           qkv projection               -> 16-bit [M,N]
//   EPI_BIAS_GELU      fc1 + GELU(tanh)             -> 16-bit [M,N]            (latte.py:169)
//   EPI_GATE_RESIDUAL  proj / fc2 + adaLN gate + residual add, fp32 residual stream in place
//                      (latte.py:179-180), optionally + temp_embed rows        (latte.py:357-358)
//
// What bounds this kernel on B200 is the operand bytes each SM must keep in flight (latency x consumption rate), so
// CTAs run as PAIRS (cluster of 2, tcgen05 cta_group::2): one 256 x BN tile per pair, M = 256 MMAs issued by the leader
// CTA, each CTA staging its own 128 A rows and only HALF of the W tile (the pair's MMA reads both halves).  Per SM that
// is 64 B per MMA cycle at BN = 256 instead of 96, and the freed shared memory buys a 6-deep TMA pipeline.
//   * both CTAs' TMA loads complete on the LEADER's "full" barrier (cta_group::2 loads, expect_tx = 2 stages); the peer's
//     producer adds one plain remote arrive per stage (NOT .release.cluster: that fence cost 1.5k cycles per k-block);
//   * the leader's tcgen05.commit is multicast to both CTAs' "empty" and "accumulator full" barriers;
//   * both CTAs' epilogue warps release an accumulator on the leader's "accumulator empty" barrier (8 arrivals).
//
// Structure (one persistent CTA per SM, 224 threads, pair-tiles visited n-fastest so concurrent clusters share A panels in L2):
//   warp 0      TMA producer: A tile 128x64 and W half-tile (BN/2)x64 (128B-swizzled) per pipeline stage
//   warp 1      TMEM allocator; in the leader CTA also the single-thread tcgen05.mma issuer (M=256, N=BN, K=16)
//   warps 2..5  epilogue warpgroup 0 (thread = accumulator row = TMEM lane)
//   warps 6..9  epilogue warpgroup 1 (16-bit output modes: the two groups take alternate 64-column chunks; with one
//               warp per scheduler the epilogue is instruction-latency bound, two warps per scheduler hide it)
//   warp 10     spare
// Two accumulator buffers in TMEM (2*BN columns) let the epilogue of tile i overlap the mainloop of tile i+1.
//
// Epilogue data movement is shaped so that every global access is a whole 128-byte line:
//   16-bit outputs:  registers -> smem staging tile (rows of 128 B, 16-byte chunks XOR-swizzled by row%8, conflict-free)
//                    -> re-read with 8 threads per row -> coalesced 16-byte stores.
//   residual:        x is never loaded: the delta gate*(acc+bias) goes to a swizzled smem tile and a TMA reduce-add
//                    (cp.reduce.async.bulk.tensor ... .add, fp32) folds it into the residual stream in L2.
#include <cstdio>
#include "common.h"
#include "ptx.cuh"

#include <cstdlib>

namespace b200 {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 x 16-bit = one 128-byte swizzle row
constexpr int kThreads = 352;   // warps: 0 TMA, 1 MMA/forwarder, 2-5 epilogue WG0, 6-9 epilogue WG1, 10 residual loader
constexpr int kSlotBytes = 128 * 128;  // one epilogue tile: 128 rows x 128 bytes

template <int BN, int EPI>
struct Cfg {
  static constexpr bool RESID = EPI == B200_EPI_GATE_RESIDUAL;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = (BN / 2) * BK * 2;                // this CTA's half of the W tile
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  // residual: 2 staging tiles per epilogue warpgroup (TMA reduce-add sources); else: 1 per warpgroup
  static constexpr int XSLOTS = RESID ? 4 : 2;
  static constexpr int STAGES = RESID ? (BN >= 256 ? 5 : (BN >= 192 ? 5 : 6)) : (BN >= 256 ? 6 : (BN >= 192 ? 6 : 8));
  static constexpr int TMEM_COLS = (2 * BN <= 256) ? 256 : 512;
  static constexpr int BAR_BYTES = 512;
  static constexpr int EPI_OFF = STAGES * STAGE_BYTES;            // 1 KiB aligned (TMA 128B-swizzle boxes live here)
  static constexpr int BAR_OFF = EPI_OFF + XSLOTS * kSlotBytes;
  static constexpr int SMEM_BYTES = BAR_OFF + BAR_BYTES + 1024;   // +1024: manual 1 KiB alignment of the base
  static_assert(STAGE_BYTES % 1024 == 0, "stage must keep 1 KiB alignment");
  static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB per-CTA shared memory limit");
  static_assert(3 * STAGES + 4 + 2 * XSLOTS + 1 <= BAR_BYTES / 8, "barrier area too small");
};

struct GemmDev {
  int M, N, K;
  int num_m, num_n;
  const float* bias;
  void* out16;
  const float* gate;
  long long gate_bs;
  int rows_per_batch;
  const float* row_add;
  int row_add_div, row_add_period;
  const uint16_t* add16;
  uint16_t* out16b;    // EPI_BIAS_GELU_BOTH: second output, gelu(out16)
  // implicit-GEMM convolution: see GemmArgs
  int conv_taps, conv_cblk, conv_h, conv_w, conv_bw, conv_bh;
  int conv_dx[9], conv_dy[9], conv_dz[9];
  int w_const;  // W is a weight (not produced by a kernel of the step): its first tiles may be fetched before griddepcontrol.wait
  int mn_major; // bit 0: A is stored [K, M]; bit 1: W is stored [K, N] (the contraction dimension is the ROW index): such an
                // operand is fetched as 64-wide MN chunks x 64 k-rows and multiplied through an MN-major UMMA descriptor.
                // wgrad sets both (dW = dY^T X), dgrad only bit 1 (dX = dY W with W in its [out, in] layout).
  int streamk;  // residual epilogue only: split the last partial wave of tiles along K across all pairs (see TileSched)
  // stream-K ordering flags (caller's workspace), one per (streamed tile, CTA rank, epilogue warpgroup): "k-blocks of the
  // tile already added into x".  All zero between launches: the segment that completes a tile resets its flag, so the
  // protocol holds no host-side state and a captured CUDA graph replays it unchanged.
  unsigned long long* sk_flags;
  int dbg;  // timing experiments only (results are wrong when set): bit0 = no operand TMA, bit1 = no MMA issue, bit2 = no 16-bit epilogue
};

__device__ __forceinline__ void flag_release(unsigned long long* f, unsigned long long v) {
  asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(f), "l"(v) : "memory");
}
__device__ __forceinline__ void flag_wait(const unsigned long long* f, unsigned long long want) {
  const long long t0 = clock64();
  while (true) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
    if (v == want) return;
    if (clock64() - t0 > 4000000000LL) {   // ~2 s: a protocol bug must not hang the GPU
      printf("gemm: stream-K flag wait timed out (block %d)\n", blockIdx.x);
      __trap();
    }
    __nanosleep(64);
  }
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// Work schedule of one CTA pair.  Data-parallel part: pair k owns tiles k, k + P, ... for the `full_waves` complete waves.
// Stream-K part (residual epilogue only, where partial sums can be reduce-added into x): the tiles of the last, partial
// wave (and of the full wave before it) are flattened into (tile, k-block) units and every pair takes an equal contiguous share, so no SM idles while a
// few pairs finish a whole extra tile.  A segment is (tile, [kb0, kb1)); bias / row_add go with the segment that has kb0 = 0.
// The partial sums of a tile are added into x in k order (a global flag per tile carries "k-blocks added so far"; the
// epilogue of a kb0 > 0 segment waits for flag == kb0), so the result is bit-reproducible although several pairs add to it.
struct TileSched {
  int pair, P, num_tiles, num_kb, full_waves, wave;
  long long u, u_end;
  __host__ __device__ TileSched(int pair_, int P_, int num_tiles_, int num_kb_, bool streamk)
      : pair(pair_), P(P_), num_tiles(num_tiles_), num_kb(num_kb_), wave(0) {
    if (streamk) {
      // stream the partial wave TOGETHER WITH the last full wave: every pair's share is then at least one tile long, so
      // a tile is cut at most once (two segments) and the ordered adds never form a chain of waiting pairs
      full_waves = num_tiles / P;
      if (full_waves > 0) --full_waves;
      const long long units = static_cast<long long>(num_tiles - full_waves * P) * num_kb;
      u = units * pair / P;
      u_end = units * (pair + 1) / P;
    } else {
      full_waves = (num_tiles + P - 1) / P;
      u = u_end = 0;
    }
  }
  __host__ __device__ bool next(int& tile, int& kb0, int& kb1) {
    if (wave < full_waves) {
      tile = wave * P + pair;
      ++wave;
      kb0 = 0;
      kb1 = num_kb;
      return tile < num_tiles;     // only the last data-parallel wave can run past the end
    }
    if (u >= u_end) return false;
    // streamed share, walked from its END: the segment that starts a tile (kb0 = 0) is done first and the one that
    // continues a tile begun by the previous pair (kb0 > 0) last, by which time that pair's part has long been added
    const int t = static_cast<int>((u_end - 1) / num_kb);
    const long long t0 = static_cast<long long>(t) * num_kb;
    kb1 = static_cast<int>(u_end - t0);
    kb0 = u > t0 ? static_cast<int>(u - t0) : 0;
    tile = full_waves * P + t;
    u_end = t0 + kb0;
    return true;
  }
};

// d/dx of the tanh-form GELU with the same one-MUFU tanh
__device__ __forceinline__ float gelu_tanh_grad(float x) {
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  const float x2 = x * x;
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(k0 * fmaf(k1 * x2, x, x)));
  const float sech2 = fmaf(-t, t, 1.0f);
  return fmaf(0.5f * x * sech2, k0 * fmaf(3.0f * k1, x2, 1.0f), fmaf(0.5f, t, 0.5f));
}

__device__ __forceinline__ float gelu_tanh(float x) {
  // 0.5 x (1 + tanh(u)),  u = sqrt(2/pi) (x + 0.044715 x^3).  ONE MUFU op per element (tanh.approx, rel. error 2^-11,
  // the same size as the 16-bit rounding of the result): with ex2 + rcp (two MUFU ops) the fc1 epilogue was bound by the
  // 16/clk/SM MUFU pipe (r01 experiment: 4.9k cycles per 128x256 tile).
  const float u = 0.7978845608028654f * fmaf(0.044715f * x * x, x, x);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
  const float hx = 0.5f * x;
  return fmaf(hx, t, hx);
}

template <int BN, int EPI, bool BF16>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const __grid_constant__ CUtensorMap tmX, const GemmDev p) {
  using C = Cfg<BN, EPI>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* epi_smem = smem + C::EPI_OFF;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::BAR_OFF);
  uint64_t* full = bars;
  uint64_t* empty = bars + C::STAGES;
  uint64_t* pfull = bars + 2 * C::STAGES;   // leader only: "the peer's stage has landed"
  uint64_t* tfull = bars + 3 * C::STAGES;
  uint64_t* tempty = tfull + 2;
  uint64_t* xfull = tempty + 2;
  uint64_t* xempty = xfull + C::XSLOTS;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(xempty + C::XSLOTS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if constexpr (C::RESID) tma_prefetch_desc(&tmX);
    for (int i = 0; i < C::STAGES; ++i) {
      mbar_init(&full[i], 2);    // (leader's copy is the live one) leader's expect_tx arrival + the peer's arrival
      mbar_init(&pfull[i], 1);   // unused
      mbar_init(&empty[i], 1);   // leader's tcgen05.commit, multicast to both CTAs
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);   // leader's tcgen05.commit, multicast to both CTAs
      mbar_init(&tempty[i], 16); // (leader's copy) one arrival per epilogue warp of BOTH CTAs
    }
    for (int i = 0; i < C::XSLOTS; ++i) {
      mbar_init(&xfull[i], 1);
      mbar_init(&xempty[i], 1);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc_pair(tmem_slot, C::TMEM_COLS);
  tc_fence_before();
  __syncthreads();      // CTA-scope order for the allocator's write of tmem_slot (what compute-sanitizer's racecheck models;
                        // the cluster barrier below already implies it)
  cluster_sync_all();   // barriers of both CTAs are initialised before any multicast / remote arrival can target them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();   // the next kernel may begin its prologue as SMs drain
  // griddepcontrol.wait (the predecessor's outputs become visible) is executed by each role that touches global memory --
  // the producer only AFTER it has put the first W tiles in flight when W is a constant weight (see below)

  // pair-tile schedule: cluster k owns pair-tiles k, k + #clusters, ...; a pair-tile is two M-adjacent 128-row tiles of
  // one N column; this CTA takes row-tile 2 * pair_m + rank (it may lie past M: zero-filled loads, clipped stores)
  const uint32_t rank = cluster_ctarank();
  const int num_pair_m = (p.num_m + 1) / 2;
  const int num_tiles = num_pair_m * p.num_n;
  const int my_pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;
  const int num_kb = p.K / BK;
  const bool streamk = C::RESID && p.streamk != 0;
  const bool leader = rank == 0;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (operands)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      // W tiles of the first k-blocks are fetched BEFORE the grid dependency resolves when W is a weight (never written by
      // a kernel of the step): they come from DRAM (1.35 GB of weights stream through L2 every step), the A tiles from the
      // L2 the predecessor just filled, so this takes the longer of the two latencies off the start of every GEMM.
      int prefetched = 0;
      if (p.w_const && !(p.dbg & 1)) {
        TileSched s0(my_pair, num_pairs, num_tiles, num_kb, streamk);
        int t0, k0, k1;
        if (s0.next(t0, k0, k1)) {
          const int n_blk = t0 % p.num_n;
          const int n_live = ((p.N - n_blk * BN) < BN && !(p.dbg & 64)) ? (p.N - n_blk * BN) : BN;
          const int w_row0 = n_blk * BN + static_cast<int>(rank) * (n_live / 2);
          prefetched = (k1 - k0) < C::STAGES ? (k1 - k0) : C::STAGES;
          for (int i = 0; i < prefetched; ++i) {
            const uint32_t full_leader = mapa_u32(&full[i], 0);
            if (leader) mbar_arrive_expect_tx(&full[i], 2 * C::STAGE_BYTES);
            else mbar_arrive_cluster(full_leader);
            tma_load_2d_pair(smem + i * C::STAGE_BYTES + C::A_BYTES, &tmB, full_leader, (k0 + i) * BK, w_row0);
          }
        }
      }
      pdl_wait();
      TileSched sched(my_pair, num_pairs, num_tiles, num_kb, streamk);
      int tile, kb0, kb1;
      while (sched.next(tile, kb0, kb1)) {
        const int m_blk = 2 * (tile / p.num_n) + static_cast<int>(rank), n_blk = tile % p.num_n;
        // narrow edge tile (N not a multiple of BN): the pair's MMA shrinks to the live columns (see the MMA warp), so the two
        // halves of the W tile are the two halves of the LIVE columns
        const int n_live = ((p.N - n_blk * BN) < BN && !(p.dbg & 64)) ? (p.N - n_blk * BN) : BN;
        const int w_row0 = n_blk * BN + static_cast<int>(rank) * (n_live / 2);
        for (int kb = kb0; kb < kb1; ++kb) {
          const bool w_in_flight = prefetched > 0;     // this stage's expect_tx / arrival and W load were issued above
          if (w_in_flight) --prefetched;
          else mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sa = smem + stage * C::STAGE_BYTES;
          const uint32_t full_leader = mapa_u32(&full[stage], 0);
          if (p.dbg & 1) {
            if (leader) mbar_arrive(&full[stage]); else mbar_arrive_cluster(full_leader);
          } else {
            // both CTAs' bytes complete on the LEADER's barrier (cta_group::2 loads may signal the pair leader)
            if (!w_in_flight) {
              if (leader) mbar_arrive_expect_tx(&full[stage], 2 * C::STAGE_BYTES);
              else mbar_arrive_cluster(full_leader);
            }
            if (p.conv_taps > 0) {
              // implicit GEMM: this k-block is channels [cb*64, +64) of filter tap `tap`; the A tile is the tile's
              // conv_bh x conv_bw pixel patch shifted by the tap offset (borders zero-filled by TMA)
              const int tap = kb / p.conv_cblk, cb = kb % p.conv_cblk;
              const int pix0 = m_blk * BM;
              const int hw = p.conv_h * p.conv_w;
              const int img = pix0 / hw, rem = pix0 % hw;
              tma_load_4d_pair(sa, &tmA, full_leader, cb * BK, rem % p.conv_w + p.conv_dx[tap], rem / p.conv_w + p.conv_dy[tap], img + p.conv_dz[tap]);
            } else if (p.mn_major & 1) {
              // operand stored [K][MN]: one box = 64 MN elements (a 128-byte swizzle row) x 64 k-rows = 8 KiB, the canonical
              // MN-major SW128 chunk; chunks past the live columns / rows are zero-filled by TMA and still count their bytes
#pragma unroll
              for (int j = 0; j < BM / 64; ++j) tma_load_2d_pair(sa + j * 8192, &tmA, full_leader, m_blk * BM + j * 64, kb * BK);
            } else {
              tma_load_2d_pair(sa, &tmA, full_leader, kb * BK, m_blk * BM);
            }
            if (p.mn_major & 2) {
#pragma unroll
              for (int j = 0; j < BN / 128; ++j)
                tma_load_2d_pair(sa + C::A_BYTES + j * 8192, &tmB, full_leader, w_row0 + j * 64, kb * BK);
            } else if (!w_in_flight) {
              tma_load_2d_pair(sa + C::A_BYTES, &tmB, full_leader, kb * BK, w_row0);
            }
          }
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (one thread of the leader CTA)
    if (leader && lane == 0) {
      constexpr uint32_t idesc_full = umma_idesc_f16(BF16, 2 * BM, BN, false, false);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      TileSched sched(my_pair, num_pairs, num_tiles, num_kb, streamk);
      int tile, kb0, kb1;
      while (sched.next(tile, kb0, kb1)) {
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        // edge tile: only the live columns (a multiple of 32) are multiplied -- no tensor work / energy on zero padding
        const int n0 = (tile % p.num_n) * BN;
        const int n_live = ((p.N - n0) < BN && !(p.dbg & 64)) ? (p.N - n0) : BN;   // dbg bit 6: A/B switch (full-width edge tiles)
        const bool a_mn = (p.mn_major & 1) != 0, b_mn = (p.mn_major & 2) != 0;
        const uint32_t idesc = (n_live == BN && !p.mn_major) ? idesc_full
                                                             : umma_idesc_f16(BF16, 2 * BM, static_cast<uint32_t>(n_live), a_mn, b_mn);
        // K-major: 8-row groups 1024 B apart, a k-step of 16 elements = 32 B inside the swizzle row.  MN-major: 64-wide MN
        // chunks 8192 B apart (leading offset), 8-k-row groups 1024 B apart, a k-step of 16 k-rows = 2048 B.
        const uint32_t lbo_a = a_mn ? 8192u : 0u, kstep_a = a_mn ? 2048u : 32u;
        const uint32_t lbo_b = b_mn ? 8192u : 0u, kstep_b = b_mn ? 2048u : 32u;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full[stage], phase);            // both CTAs' operands landed
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * C::STAGE_BYTES);
          const uint64_t da = umma_smem_desc(sa, lbo_a, 1024, UMMA_LAYOUT_SW128);
          const uint64_t db = umma_smem_desc(sa + C::A_BYTES, lbo_b, 1024, UMMA_LAYOUT_SW128);
          if (!(p.dbg & 2))
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma_f16_ss_pair(d_tmem, umma_desc_advance(da, k * kstep_a), umma_desc_advance(db, k * kstep_b), idesc,
                             (kb != kb0 || k != 0) ? 1u : 0u);
          umma_commit_pair(&empty[stage], 0x3);  // slot reusable in BOTH CTAs once these MMAs have read it
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit_pair(&tfull[acc], 0x3);      // accumulator complete -> both CTAs' epilogues
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp == 10) {
    // spare warp (kept so the warp-role layout is the same for every epilogue mode)
  } else {
    // ------------------------------------------------------------------ epilogue warps 2..9
    pdl_wait();                               // bias / gate / shortcut / residual stream: visible from here
    const int q = warp & 3;                   // TMEM lane quarter this warp may read
    const int row_a = q * 32 + lane;          // accumulator row of this thread
    const int wg = warp >= 6 ? 1 : 0;         // epilogue warpgroup
    const int te = threadIdx.x - 64 - wg * 128;  // 0..127 within the warpgroup
    int acc = 0;
    uint32_t acc_phase = 0;
    const uint32_t tempty_leader[2] = {mapa_u32(&tempty[0], 0), mapa_u32(&tempty[1], 0)};

    if constexpr (C::RESID) {
      // x += gate * (acc + bias) (+ row_add) WITHOUT loading x: the delta tile goes registers -> smem (128-byte rows,
      // 16-byte chunks XOR-swizzled to match the tensor map) -> TMA reduce-add into the fp32 residual stream.  Each
      // element receives one add per GEMM -- or, for stream-K tiles, one per segment in a fixed (k) order -- so results
      // are deterministic.  The two warpgroups take alternate 32-column chunks (owner = running chunk counter & 1, or
      // the chunk's own parity for stream-K so that a column always belongs to the same warpgroup) and double-buffer
      // their own two staging tiles.
      uint32_t cc = 0, mine = 0;
      const int bar_a = 1 + 2 * wg, bar_b = 2 + 2 * wg;
      TileSched sched(my_pair, num_pairs, num_tiles, num_kb, streamk);
      int tile, kb0, kb1;
      while (sched.next(tile, kb0, kb1)) {
        const bool first_seg = kb0 == 0;       // bias and row_add are added once per output element
        const int m0 = (2 * (tile / p.num_n) + static_cast<int>(rank)) * BM, n0 = (tile % p.num_n) * BN;
        const uint32_t t_row = tmem_base + acc * BN + (static_cast<uint32_t>(q * 32) << 16);
        const int row = m0 + row_a;
        const int row_c = row < p.M ? row : p.M - 1;  // rows past M produce deltas that the TMA store clips
        const float* gate_row = p.gate + static_cast<long long>(row_c / p.rows_per_batch) * p.gate_bs;
        const float* add_row = (p.row_add && first_seg) ? p.row_add + static_cast<size_t>((row_c / p.row_add_div) % p.row_add_period) * p.N : nullptr;
        constexpr int NCH = BN / 32;
        const int live = (p.N - n0) / 32 < NCH ? (p.N - n0) / 32 : NCH;   // chunks inside N (N % 32 == 0)
        const uint32_t cbase = streamk ? 0u : cc;
        int last_mine = -1;
        for (int c = 0; c < live; ++c)
          if (((cbase + c) & 1) == static_cast<uint32_t>(wg)) last_mine = c;
        // stream-K ordering: this warpgroup's adds for the tile follow those of the segment that ends at kb0
        const bool partial = kb0 > 0 || kb1 < num_kb;
        unsigned long long* flag = nullptr;
        if (partial) flag = p.sk_flags + (static_cast<size_t>(tile - sched.full_waves * num_pairs) * 2 + rank) * 2 + wg;
        bool must_wait = kb0 > 0;
        mbar_wait(&tfull[acc], acc_phase);
        tc_fence_after();
        if (last_mine < 0) {                   // a one-chunk edge tile owned by the other warpgroup
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(tempty_leader[acc]);
        }
#pragma unroll 1
        for (int c = 0; c < live; ++c) {
          if (((cbase + c) & 1) != static_cast<uint32_t>(wg)) continue;
          const int col0 = n0 + c * 32;
          uint32_t v[32];
          tmem_ld_32x32b_x32(t_row + c * 32, v);
          tmem_ld_wait();
          if (c == last_mine) {                // all of this warp's TMEM reads of the tile are done
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(tempty_leader[acc]);
          }
          if (p.dbg & 8) continue;     // timing experiment: accumulator read only
          uint8_t* slot = epi_smem + (wg * 2 + (mine & 1)) * kSlotBytes;
          ++mine;
          uint8_t* drow = slot + row_a * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p.bias && first_seg) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col0) + j);
            const float4 gt = __ldg(reinterpret_cast<const float4*>(gate_row + col0) + j);
            float4 dv;
            dv.x = gt.x * (__uint_as_float(v[4 * j + 0]) + b4.x);
            dv.y = gt.y * (__uint_as_float(v[4 * j + 1]) + b4.y);
            dv.z = gt.z * (__uint_as_float(v[4 * j + 2]) + b4.z);
            dv.w = gt.w * (__uint_as_float(v[4 * j + 3]) + b4.w);
            if (add_row) {
              const float4 ra = __ldg(reinterpret_cast<const float4*>(add_row + col0) + j);
              dv.x += ra.x; dv.y += ra.y; dv.z += ra.z; dv.w += ra.w;
            }
            *reinterpret_cast<float4*>(drow + ((j ^ (row_a & 7)) << 4)) = dv;
          }
          fence_proxy_async_smem();                     // generic-proxy smem writes -> visible to the TMA engine
          asm volatile("bar.sync %0, 128;" ::"r"(bar_a) : "memory");
          if (te == 0) {
            if (must_wait) {
              flag_wait(flag, static_cast<unsigned long long>(kb0));
              // nobody else waits on this flag; the segment that completes the tile leaves it zero for the next launch
              if (kb1 == num_kb) flag_release(flag, 0ull);
              fence_proxy_async_all();
            }
            tma_reduce_add_2d(&tmX, slot, col0, m0);
            tma_store_commit();
            tma_store_wait_read<1>();                   // the reduce issued one chunk ago (other tile) has read its smem
          }
          must_wait = false;
          asm volatile("bar.sync %0, 128;" ::"r"(bar_b) : "memory");   // ... so the other staging tile may be rewritten
        }
        if (partial && kb1 < num_kb && last_mine >= 0 && te == 0 && !(p.dbg & 8)) {
          tma_store_wait_all<0>();                      // this segment's adds have been performed ...
          __threadfence();
          flag_release(flag, static_cast<unsigned long long>(kb1));   // ... the next segment may add
        }
        cc += live;
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
      if (te == 0) tma_store_wait_all<0>();             // all residual updates have landed before the CTA retires
    } else {
      const int row_b0 = te >> 3;             // phase-B row within a group of 16
      const int ch_b = te & 7;                // phase-B 16-byte chunk within the 128-byte row segment
      uint8_t* buf = epi_smem + wg * kSlotBytes;   // one staging tile per warpgroup
      const int bar_a = 1 + 2 * wg, bar_b = 2 + 2 * wg;
      uint32_t cc = 0;                        // running chunk counter: chunk cc belongs to warpgroup (cc & 1)
      TileSched sched(my_pair, num_pairs, num_tiles, num_kb, false);
      int tile, kb0, kb1;
      while (sched.next(tile, kb0, kb1)) {
        const int m0 = (2 * (tile / p.num_n) + static_cast<int>(rank)) * BM, n0 = (tile % p.num_n) * BN;
        const uint32_t t_row = tmem_base + acc * BN + (static_cast<uint32_t>(q * 32) << 16);
        constexpr int NCH = BN / 64;           // 64 16-bit columns = 128 bytes per row per chunk
        // this warpgroup's last chunk of the tile (after it the accumulator is no longer read by us)
        int last_mine = -1;
#pragma unroll
        for (int c = 0; c < NCH; ++c)
          if (((cc + c) & 1) == static_cast<uint32_t>(wg)) last_mine = c;
        mbar_wait(&tfull[acc], acc_phase);
        tc_fence_after();
        if (last_mine < 0) {                   // cannot happen for NCH >= 2, kept for safety
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(tempty_leader[acc]);
        }
#pragma unroll 1
        for (int c = 0; c < NCH; ++c) {
          if (((cc + c) & 1) != static_cast<uint32_t>(wg)) continue;
          uint32_t v0[32], v1[32];
          tmem_ld_32x32b_x32(t_row + c * 64, v0);
          tmem_ld_32x32b_x32(t_row + c * 64 + 32, v1);
          tmem_ld_wait();
          if (c == last_mine) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(tempty_leader[acc]);
          }
          if (p.dbg & 4) continue;   // timing experiment: accumulator read only
          const int col0 = n0 + c * 64;
          uint8_t* srow = buf + row_a * 128;
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            float f[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(hh ? v1[j] : v0[j]);
            const int cb = col0 + hh * 32;
            if (p.bias && cb < p.N) {
              const float4* b4 = reinterpret_cast<const float4*>(p.bias + cb);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 b = __ldg(b4 + j);
                f[4 * j + 0] += b.x; f[4 * j + 1] += b.y; f[4 * j + 2] += b.z; f[4 * j + 3] += b.w;
              }
            }
            if constexpr (EPI == B200_EPI_BIAS_GELU) {
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] = gelu_tanh(f[j]);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 o;
              o.x = pack2<BF16>(f[8 * j + 0], f[8 * j + 1]);
              o.y = pack2<BF16>(f[8 * j + 2], f[8 * j + 3]);
              o.z = pack2<BF16>(f[8 * j + 4], f[8 * j + 5]);
              o.w = pack2<BF16>(f[8 * j + 6], f[8 * j + 7]);
              *reinterpret_cast<uint4*>(srow + (((hh * 4 + j) ^ (row_a & 7)) << 4)) = o;
            }
          }
          asm volatile("bar.sync %0, 128;" ::"r"(bar_a) : "memory");   // staging tile complete
          const int col = col0 + ch_b * 8;
          if (col < p.N) {
#pragma unroll
            for (int it = 0; it < 8; ++it) {
              const int rl = it * 16 + row_b0;
              const int row = m0 + rl;
              if (row < p.M) {
                uint4 val = *reinterpret_cast<const uint4*>(buf + rl * 128 + ((ch_b ^ (rl & 7)) << 4));
                if constexpr (EPI == B200_EPI_BIAS_MUL16) {   // gated feed-forward: (h wi_1^T) * gelu(h wi_0^T), the second factor read back in 16 bits
                  const uint4 rs = __ldg(reinterpret_cast<const uint4*>(p.add16 + static_cast<size_t>(row) * p.N + col));
                  const float2 a0 = unpack2<BF16>(val.x), a1 = unpack2<BF16>(val.y), a2 = unpack2<BF16>(val.z), a3 = unpack2<BF16>(val.w);
                  const float2 r0 = unpack2<BF16>(rs.x), r1 = unpack2<BF16>(rs.y), r2 = unpack2<BF16>(rs.z), r3 = unpack2<BF16>(rs.w);
                  val.x = pack2<BF16>(a0.x * r0.x, a0.y * r0.y);
                  val.y = pack2<BF16>(a1.x * r1.x, a1.y * r1.y);
                  val.z = pack2<BF16>(a2.x * r2.x, a2.y * r2.y);
                  val.w = pack2<BF16>(a3.x * r3.x, a3.y * r3.y);
                }
                if constexpr (EPI == B200_EPI_MUL_GELUGRAD16) {   // training, dgrad of fc2: du = da * gelu'(u), u (fc1's pre-activation) read back in 16 bits
                  const uint4 rs = __ldg(reinterpret_cast<const uint4*>(p.add16 + static_cast<size_t>(row) * p.N + col));
                  const float2 a0 = unpack2<BF16>(val.x), a1 = unpack2<BF16>(val.y), a2 = unpack2<BF16>(val.z), a3 = unpack2<BF16>(val.w);
                  const float2 r0 = unpack2<BF16>(rs.x), r1 = unpack2<BF16>(rs.y), r2 = unpack2<BF16>(rs.z), r3 = unpack2<BF16>(rs.w);
                  val.x = pack2<BF16>(a0.x * gelu_tanh_grad(r0.x), a0.y * gelu_tanh_grad(r0.y));
                  val.y = pack2<BF16>(a1.x * gelu_tanh_grad(r1.x), a1.y * gelu_tanh_grad(r1.y));
                  val.z = pack2<BF16>(a2.x * gelu_tanh_grad(r2.x), a2.y * gelu_tanh_grad(r2.y));
                  val.w = pack2<BF16>(a3.x * gelu_tanh_grad(r3.x), a3.y * gelu_tanh_grad(r3.y));
                }
                if constexpr (EPI == B200_EPI_BIAS_GELU_BOTH) {   // training, fc1: keep the pre-activation u (out16) AND write gelu(u) (out16b)
                  const float2 a0 = unpack2<BF16>(val.x), a1 = unpack2<BF16>(val.y), a2 = unpack2<BF16>(val.z), a3 = unpack2<BF16>(val.w);
                  uint4 gl;
                  gl.x = pack2<BF16>(gelu_tanh(a0.x), gelu_tanh(a0.y));
                  gl.y = pack2<BF16>(gelu_tanh(a1.x), gelu_tanh(a1.y));
                  gl.z = pack2<BF16>(gelu_tanh(a2.x), gelu_tanh(a2.y));
                  gl.w = pack2<BF16>(gelu_tanh(a3.x), gelu_tanh(a3.y));
                  *reinterpret_cast<uint4*>(p.out16b + static_cast<size_t>(row) * p.N + col) = gl;
                }
                if constexpr (EPI == B200_EPI_BIAS_ADD16) {   // + shortcut, both already rounded to 16 bits like the reference
                  const uint4 rs = __ldg(reinterpret_cast<const uint4*>(p.add16 + static_cast<size_t>(row) * p.N + col));
                  const float2 a0 = unpack2<BF16>(val.x), a1 = unpack2<BF16>(val.y), a2 = unpack2<BF16>(val.z), a3 = unpack2<BF16>(val.w);
                  const float2 r0 = unpack2<BF16>(rs.x), r1 = unpack2<BF16>(rs.y), r2 = unpack2<BF16>(rs.z), r3 = unpack2<BF16>(rs.w);
                  val.x = pack2<BF16>(a0.x + r0.x, a0.y + r0.y);
                  val.y = pack2<BF16>(a1.x + r1.x, a1.y + r1.y);
                  val.z = pack2<BF16>(a2.x + r2.x, a2.y + r2.y);
                  val.w = pack2<BF16>(a3.x + r3.x, a3.y + r3.y);
                }
                *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.out16) + static_cast<size_t>(row) * p.N + col) = val;
              }
            }
          }
          asm volatile("bar.sync %0, 128;" ::"r"(bar_b) : "memory");   // staging tile drained: safe to refill
        }
        cc += NCH;
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();   // the peer may still multicast into our smem / arrive on our barriers until it is done too
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, C::TMEM_COLS);
  }
}

template <int BN, int EPI, bool BF16>
int launch_one(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmX, const GemmDev& p, int grid,
               cudaStream_t stream) {
  using C = Cfg<BN, EPI>;
  auto kern = gemm_kernel<BN, EPI, BF16>;
  B200_SET_SMEM_ONCE(kern, C::SMEM_BYTES);
  B200_CHECK_CUDA(launch_pdl(kern, dim3(grid), dim3(kThreads), C::SMEM_BYTES, stream, tmA, tmB, tmX, p));
  return B200_OK;
}

template <int BN, bool BF16>
int launch_epi(int epi, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmX, const GemmDev& p, int grid,
               cudaStream_t s) {
  switch (epi) {
    case B200_EPI_BIAS: return launch_one<BN, B200_EPI_BIAS, BF16>(tmA, tmB, tmX, p, grid, s);
    case B200_EPI_BIAS_GELU: return launch_one<BN, B200_EPI_BIAS_GELU, BF16>(tmA, tmB, tmX, p, grid, s);
    case B200_EPI_GATE_RESIDUAL: return launch_one<BN, B200_EPI_GATE_RESIDUAL, BF16>(tmA, tmB, tmX, p, grid, s);
    case B200_EPI_BIAS_ADD16: return launch_one<BN, B200_EPI_BIAS_ADD16, BF16>(tmA, tmB, tmX, p, grid, s);
    case B200_EPI_BIAS_MUL16: return launch_one<BN, B200_EPI_BIAS_MUL16, BF16>(tmA, tmB, tmX, p, grid, s);
    case B200_EPI_BIAS_GELU_BOTH: return launch_one<BN, B200_EPI_BIAS_GELU_BOTH, BF16>(tmA, tmB, tmX, p, grid, s);
    case B200_EPI_MUL_GELUGRAD16: return launch_one<BN, B200_EPI_MUL_GELUGRAD16, BF16>(tmA, tmB, tmX, p, grid, s);
  }
  set_error("gemm: unknown epilogue %d", epi);
  return B200_ERR_UNSUPPORTED;
}

template <int BN>
int launch_bn(int bf16, int epi, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmX, const GemmDev& p,
              int grid, cudaStream_t s) {
  return bf16 ? launch_epi<BN, true>(epi, tmA, tmB, tmX, p, grid, s) : launch_epi<BN, false>(epi, tmA, tmB, tmX, p, grid, s);
}

constexpr int kStreamKMinKb = 32;          // K / 64 below which the split costs more than the idle tail it removes
constexpr int kSkFlags = B200_GEMM_SK_FLAGS;   // flags (u64) in the caller's buffer: 4 per streamed tile

int streamk_min_kb() {
  static const int v = env_int("B200_GEMM_SK_MINKB", kStreamKMinKb);
  return v;
}


int pick_block_n(int M, int N, int K, bool resid, int sms) {
  // minimise waves x per-tile time.  The kernel is bound by L2->SM operand bytes, not MMA cycles, so a tile costs
  // ~ (A bytes + W/2 bytes per k-block per CTA) = 128 + BN/2 rather than BN (measured: r01 microbench, profiles/).
  // Where the last wave is streamed along K (residual epilogue, long K) there is no wave rounding.
  if (N <= 128) return 128;   // narrow outputs (e.g. the VAE's 3-channel conv_out padded to 32): smallest tile that covers N
  static const bool no_sk = env_int("B200_GEMM_NO_STREAMK", 0) != 0;
  const bool sk = resid && !no_sk && K / BK >= streamk_min_kb();
  const int cand[3] = {256, 192, 128};
  int best = 128;
  double best_cost = 1e300;
  for (int i = 0; i < 3; ++i) {
    const long long tiles = static_cast<long long>((M + BM - 1) / BM) * ((N + cand[i] - 1) / cand[i]);
    const double waves = (sk && tiles > sms) ? static_cast<double>(tiles) / sms : static_cast<double>((tiles + sms - 1) / sms);
    const double cost = waves * (128 + cand[i] / 2);
    if (cost < best_cost - 1e-9) { best_cost = cost; best = cand[i]; }
  }
  return best;
}

// The scheduling decisions of one launch (shared by launch_gemm and the schedule dump the CPU tests read).
struct GemmPlan { int bn, pairs, pair_tiles, num_kb, streamk; };
GemmPlan plan_gemm(int M, int N, int K, bool resid, int block_n, int sms, bool split_small = false) {
  static const bool no_sk = env_int("B200_GEMM_NO_STREAMK", 0) != 0;
  GemmPlan g;
  g.bn = block_n ? block_n : pick_block_n(M, N, K, resid, sms);
  const int num_m = (M + BM - 1) / BM, num_n = (N + g.bn - 1) / g.bn;
  g.pair_tiles = ((num_m + 1) / 2) * num_n;
  const int grid = 2 * g.pair_tiles < sms ? 2 * g.pair_tiles : (sms & ~1);   // whole clusters of 2
  g.pairs = grid / 2;
  g.num_kb = K / BK;
  g.streamk = (resid && !no_sk && g.pair_tiles > g.pairs && g.pair_tiles % g.pairs != 0 && g.num_kb >= streamk_min_kb()) ? 1 : 0;
  // weight gradients: a small output (fewer tiles than CTA pairs) under a very long contraction (K = tokens).  All tiles are
  // streamed: the (tile, k-block) units are cut into one equal share per pair, a tile's partial sums are reduce-added in k
  // order through the same flags (a chain of waits only ever points from pair p+1 to pair p, and every pair is resident).
  if (split_small && resid && !no_sk && g.pair_tiles <= sms / 2 && g.pair_tiles % (sms / 2) != 0 &&
      static_cast<long long>(g.pair_tiles) * g.num_kb >= static_cast<long long>(sms / 2) * streamk_min_kb()) {
    g.pairs = sms / 2;
    g.streamk = 1;
  }
  return g;
}

}  // namespace

int gemm_schedule(int M, int N, int K, int epilogue, int block_n, int sms, int* bn_out, int* pairs_out, int* streamk_out,
                  int* segments, int max_segments, int wgrad) {
  B200_REQUIRE(M > 0 && N > 0 && K > 0 && K % BK == 0 && sms >= 2, B200_ERR_SHAPE, "gemm_schedule: bad arguments");
  B200_REQUIRE(block_n == 0 || block_n == 128 || block_n == 192 || block_n == 256, B200_ERR_UNSUPPORTED, "gemm_schedule: block_n %d", block_n);
  if (wgrad && block_n == 0) {               // the tile-width rule launch_gemm applies to weight gradients (mn_major == 3)
    block_n = pick_block_n(M, N, K, true, sms);
    if (block_n == 192) block_n = (N % 256 == 0 || N > 1024) ? 256 : 128;
    if (N >= 256) block_n = 256;
  }
  const GemmPlan g = plan_gemm(M, N, K, epilogue == B200_EPI_GATE_RESIDUAL, block_n, sms, wgrad != 0);
  if (bn_out) *bn_out = g.bn;
  if (pairs_out) *pairs_out = g.pairs;
  if (streamk_out) *streamk_out = g.streamk;
  int n = 0;
  for (int pair = 0; pair < g.pairs; ++pair) {
    TileSched sched(pair, g.pairs, g.pair_tiles, g.num_kb, g.streamk != 0);
    int tile, kb0, kb1;
    while (sched.next(tile, kb0, kb1)) {       // in the order the pair executes them
      if (segments && n < max_segments) {
        segments[4 * n + 0] = pair; segments[4 * n + 1] = tile; segments[4 * n + 2] = kb0; segments[4 * n + 3] = kb1;
      }
      ++n;
    }
  }
  return n;
}

int launch_gemm(const GemmArgs& a, cudaStream_t stream) {
  B200_REQUIRE(a.M > 0 && a.N > 0 && a.K > 0, B200_ERR_SHAPE, "gemm: bad shape M=%d N=%d K=%d", a.M, a.N, a.K);
  B200_REQUIRE(a.K % BK == 0, B200_ERR_SHAPE, "gemm: K=%d must be a multiple of %d", a.K, BK);
  B200_REQUIRE(a.N % 32 == 0, B200_ERR_SHAPE, "gemm: N=%d must be a multiple of 32", a.N);
  B200_REQUIRE(a.epilogue == B200_EPI_BIAS || a.epilogue == B200_EPI_BIAS_GELU || a.epilogue == B200_EPI_GATE_RESIDUAL ||
                   a.epilogue == B200_EPI_BIAS_ADD16 || a.epilogue == B200_EPI_BIAS_MUL16 || a.epilogue == B200_EPI_BIAS_GELU_BOTH ||
                   a.epilogue == B200_EPI_MUL_GELUGRAD16,
               B200_ERR_UNSUPPORTED, "gemm: unknown epilogue %d", a.epilogue);
  B200_REQUIRE((a.epilogue != B200_EPI_BIAS_ADD16 && a.epilogue != B200_EPI_BIAS_MUL16 && a.epilogue != B200_EPI_MUL_GELUGRAD16) ||
                   (a.add16 && (reinterpret_cast<uintptr_t>(a.add16) & 15) == 0), B200_ERR_ALIGN, "gemm: add16 tensor missing or unaligned");
  B200_REQUIRE(a.epilogue != B200_EPI_BIAS_GELU_BOTH || (a.out16b && (reinterpret_cast<uintptr_t>(a.out16b) & 15) == 0), B200_ERR_ALIGN,
               "gemm: second output missing or unaligned");
  B200_REQUIRE((a.epilogue != B200_EPI_BIAS_GELU_BOTH && a.epilogue != B200_EPI_MUL_GELUGRAD16) || a.N % 8 == 0, B200_ERR_SHAPE, "gemm: N %% 8");
  B200_REQUIRE((reinterpret_cast<uintptr_t>(a.A) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.W) & 15) == 0,
               B200_ERR_ALIGN, "gemm: A and W must be 16-byte aligned");
  const bool resid = a.epilogue == B200_EPI_GATE_RESIDUAL;
  if (resid) {
    B200_REQUIRE(a.resid && a.gate && a.rows_per_batch > 0, B200_ERR_SHAPE, "gemm: gated-residual epilogue needs resid, gate, rows_per_batch");
    B200_REQUIRE((reinterpret_cast<uintptr_t>(a.resid) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.gate) & 15) == 0 &&
                     (a.gate_batch_stride % 4) == 0,
                 B200_ERR_ALIGN, "gemm: resid/gate must be 16-byte aligned");
    B200_REQUIRE(!a.row_add || (reinterpret_cast<uintptr_t>(a.row_add) & 15) == 0, B200_ERR_ALIGN, "gemm: row_add must be 16-byte aligned");
  } else {
    B200_REQUIRE(a.out16 && (reinterpret_cast<uintptr_t>(a.out16) & 15) == 0, B200_ERR_ALIGN, "gemm: out16 must be 16-byte aligned");
  }
  B200_REQUIRE(!a.bias || (reinterpret_cast<uintptr_t>(a.bias) & 15) == 0, B200_ERR_ALIGN, "gemm: bias must be 16-byte aligned");
  B200_TRY(check_arch());
  int sms = 0;
  B200_TRY(device_sm_count(&sms));

  B200_REQUIRE(a.block_n == 0 || a.block_n == 128 || a.block_n == 192 || a.block_n == 256, B200_ERR_UNSUPPORTED,
               "gemm: block_n must be 128, 192 or 256 (got %d)", a.block_n);
  int block_n = a.block_n;
  if (a.mn_major) {
    B200_REQUIRE(a.mn_major >= 1 && a.mn_major <= 3 && a.conv_taps == 0 && (!(a.mn_major & 1) || a.M % 8 == 0) &&
                     (!(a.mn_major & 2) || a.N % 128 == 0) && (a.block_n == 0 || a.block_n == 128 || a.block_n == 256),
                 B200_ERR_UNSUPPORTED, "gemm (transposed operands): M %% 8 == 0, N %% 128 == 0, block_n 128 or 256 (M=%d N=%d)", a.M, a.N);
    if (block_n == 0) {
      block_n = pick_block_n(a.M, a.N, a.K, a.epilogue == B200_EPI_GATE_RESIDUAL, sms);
      if (block_n == 192) block_n = (a.N % 256 == 0 || a.N > 1024) ? 256 : 128;     // W chunks are 64 wide per CTA: 128 or 256 only
      // weight gradients: 256-wide tiles throughout -- the cost model's 128 for small outputs (proj: 1152 x 1152) measured slower
      // once the tiles are split along K anyway (88.5 vs 75.4 us); B200_WGRAD_BN=128|256 forces a width for A/B runs
      static const int wgrad_bn = env_int("B200_WGRAD_BN", 0);
      if (a.mn_major == 3) block_n = (wgrad_bn == 128 || wgrad_bn == 256) ? wgrad_bn : (a.N >= 256 ? 256 : block_n);
    }
  }
  const GemmPlan plan = plan_gemm(a.M, a.N, a.K, a.epilogue == B200_EPI_GATE_RESIDUAL, block_n, sms, a.mn_major == 3);
  const int bn = plan.bn;

  CUtensorMap tmA, tmB, tmX;
  int conv_bw = 0, conv_bh = 0;
  {
    if (a.conv_taps > 0) {
      B200_REQUIRE(a.conv_taps <= 9 && a.conv_c % BK == 0 && a.K == a.conv_taps * a.conv_c &&
                       a.M == a.conv_n * a.conv_h * a.conv_w,
                   B200_ERR_SHAPE, "conv: inconsistent geometry (taps %d, C %d, K %d, M %d)", a.conv_taps, a.conv_c, a.K, a.M);
      conv_bw = a.conv_w >= BM ? BM : a.conv_w;
      conv_bh = BM / conv_bw;
      B200_REQUIRE(BM % conv_bw == 0 && a.conv_w % conv_bw == 0 && a.conv_h % conv_bh == 0, B200_ERR_UNSUPPORTED,
                   "conv: %dx%d feature map cannot be tiled by 128-pixel patches", a.conv_h, a.conv_w);
      const uint64_t dimsA[4] = {static_cast<uint64_t>(a.conv_c), static_cast<uint64_t>(a.conv_w), static_cast<uint64_t>(a.conv_h),
                                 static_cast<uint64_t>(a.conv_n)};
      const uint64_t strA[3] = {static_cast<uint64_t>(a.conv_c) * 2, static_cast<uint64_t>(a.conv_c) * 2 * a.conv_w,
                                static_cast<uint64_t>(a.conv_c) * 2 * a.conv_w * a.conv_h};
      const uint32_t boxA[4] = {BK, static_cast<uint32_t>(conv_bw), static_cast<uint32_t>(conv_bh), 1};
      B200_TRY(make_tmap_16bit(&tmA, a.A, 4, dimsA, strA, boxA, TMAP_SW_128));
    } else if (a.mn_major & 1) {
      const uint64_t dimsA[2] = {static_cast<uint64_t>(a.M), static_cast<uint64_t>(a.K)};      // stored [K][M]
      const uint64_t strA[1] = {static_cast<uint64_t>(a.M) * 2};
      const uint32_t boxA[2] = {64, BK};
      B200_TRY(make_tmap_16bit(&tmA, a.A, 2, dimsA, strA, boxA, TMAP_SW_128));
    } else {
      const uint64_t dimsA[2] = {static_cast<uint64_t>(a.K), static_cast<uint64_t>(a.M)};
      const uint64_t strA[1] = {static_cast<uint64_t>(a.K) * 2};
      const uint32_t boxA[2] = {BK, BM};
      B200_TRY(make_tmap_16bit(&tmA, a.A, 2, dimsA, strA, boxA, TMAP_SW_128));
    }
    if (a.mn_major & 2) {
      const uint64_t dimsB[2] = {static_cast<uint64_t>(a.N), static_cast<uint64_t>(a.K)};      // stored [K][N]
      const uint64_t strB[1] = {static_cast<uint64_t>(a.N) * 2};
      const uint32_t boxB[2] = {64, BK};
      B200_TRY(make_tmap_16bit(&tmB, a.W, 2, dimsB, strB, boxB, TMAP_SW_128));
    } else {
      const uint64_t dimsB[2] = {static_cast<uint64_t>(a.K), static_cast<uint64_t>(a.N)};
      const uint64_t strB[1] = {static_cast<uint64_t>(a.K) * 2};
      const uint32_t boxB[2] = {BK, static_cast<uint32_t>(bn / 2)};   // each CTA of the pair fetches half and multicasts it
      B200_TRY(make_tmap_16bit(&tmB, a.W, 2, dimsB, strB, boxB, TMAP_SW_128));
    }
    if (resid) {
      const uint64_t dimsX[2] = {static_cast<uint64_t>(a.N), static_cast<uint64_t>(a.M)};
      const uint64_t strX[1] = {static_cast<uint64_t>(a.N) * 4};
      const uint32_t boxX[2] = {32, BM};
      B200_TRY(make_tmap(&tmX, a.resid, 4, 2, dimsX, strX, boxX, TMAP_SW_128));
    } else {
      tmX = tmA;  // unused
    }
  }
  GemmDev p;
  p.M = a.M; p.N = a.N; p.K = a.K;
  p.num_m = (a.M + BM - 1) / BM;
  p.num_n = (a.N + bn - 1) / bn;
  p.bias = a.bias;
  p.out16 = a.out16;
  p.gate = a.gate;
  p.gate_bs = a.gate_batch_stride;
  p.rows_per_batch = a.rows_per_batch > 0 ? a.rows_per_batch : 1;
  p.row_add = a.row_add;
  p.row_add_div = a.row_add_div > 0 ? a.row_add_div : 1;
  p.row_add_period = a.row_add_period > 0 ? a.row_add_period : 1;
  p.add16 = static_cast<const uint16_t*>(a.add16);
  p.out16b = static_cast<uint16_t*>(a.out16b);
  p.conv_taps = a.conv_taps;
  p.conv_cblk = a.conv_taps > 0 ? a.conv_c / BK : 0;
  p.conv_h = a.conv_h; p.conv_w = a.conv_w; p.conv_bw = conv_bw; p.conv_bh = conv_bh;
  for (int i = 0; i < 9; ++i) { p.conv_dx[i] = a.conv_dx[i]; p.conv_dy[i] = a.conv_dy[i]; p.conv_dz[i] = a.conv_dz[i]; }
  static const int dbg = env_int("B200_GEMM_DBG", 0);
  p.dbg = dbg;
  static const int no_wprefetch = env_int("B200_GEMM_NO_WPREFETCH", 0);     // A/B switch
  p.w_const = (a.w_const && !no_wprefetch && a.conv_taps == 0 && !a.mn_major) ? 1 : 0;
  p.mn_major = a.mn_major;
  const int pair_tiles = plan.pair_tiles;
  const int grid = 2 * plan.pairs;
  {
    // stream-K over the last waves: only where partial sums can be reduce-added (residual epilogue) and K is long enough
    // to be worth splitting (plan_gemm); B200_GEMM_NO_STREAMK=1 restores the one-add-per-element schedule.
    const int pairs = plan.pairs;
    // the ordering flags live in the CALLER's buffer (zeroed once; every launch leaves it zeroed): without one the
    // schedule stays data-parallel
    const int streamed = pair_tiles % pairs + pairs;  // tiles of the partial wave and of the full wave before it
    p.streamk = (plan.streamk && a.sk_flags != nullptr && streamed * 4 <= kSkFlags) ? 1 : 0;
    p.sk_flags = p.streamk ? a.sk_flags : nullptr;
  }
  switch (bn) {
    case 128: return launch_bn<128>(a.bf16, a.epilogue, tmA, tmB, tmX, p, grid, stream);
    case 192: return launch_bn<192>(a.bf16, a.epilogue, tmA, tmB, tmX, p, grid, stream);
    default: return launch_bn<256>(a.bf16, a.epilogue, tmA, tmB, tmX, p, grid, stream);
  }
}

}  // namespace b200
