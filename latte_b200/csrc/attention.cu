// Fused spatial / temporal attention on tcgen05:  out = softmax(q k^T / sqrt(hd)) v  per head.
//
// Replaces the 'math' branch of Attention.forward between its two Linears (reference models/latte.py:50-70:
// reshape/permute/.contiguous(), q@k^T, *scale, softmax, @v, transpose/reshape) AND the two einops
// rearranges that regroup tokens between spatial and temporal blocks (latte.py:355,368).  The hidden state and
// the qkv buffer stay in ONE layout, rows = (b, f, n); the regrouping is done by the TMA box that fetches a tile:
//   spatial  (temporal=0): a tile is 128 consecutive tokens of one frame; keys = the frame's N tokens.
//            3-D map {hd, 3H, T}; Q box {64|16, 1, 128}, K/V box {64|16, 1, N}.
//            N < 128: 128/N whole sequences are packed per tile with a block-diagonal mask (r/N == c/N).
//   temporal (temporal=1): a tile is G = 128/F neighbouring tokens x all F frames, fetched with a 4-D map
//            {hd, 3H, N, B*F} and box {64|16, 1, G, F} -> tile row r = f*G + g; the G sequences are masked
//            block-diagonally (r % G == c % G).  Wasted tensor FLOPs (x G) are free; HBM traffic is minimal:
//            every q/k/v element is read once, every output written once, no transpose pass.
// head_dim 72 (XL) is not a multiple of the UMMA K=16 / N=16 granules: the first 64 dims use 128B-swizzled
// tiles, dims [64,80) a second 32B-swizzled tile whose out-of-range columns TMA zero-fills.
//
// CTA = 160 threads: warps 0-3 softmax/epilogue (thread = tile row = TMEM lane), warp 4 = TMA + MMA issuer.
//   S = Q K^T -> TMEM cols [0,Lk)  ->  registers: mask, max, exp2, sum -> P (16-bit) to smem (K-major SW128)
//   O = P V   -> TMEM cols [256,336) -> registers: * 1/sum -> global.
#include <cstdlib>
#include "common.h"
#include "ptx.cuh"

namespace b200 {

namespace {

constexpr int kThreads = 160;
constexpr int kOCol = 0;  // O (<= 80 columns) reuses the first S columns: S is dead once every row has written its P

enum { MODE_FULL = 0, MODE_PACKED = 1, MODE_TEMPORAL = 2, MODE_CROSS = 3 };

struct AttnDev {
  void* out;
  int T;            // total rows
  int D;            // heads * head_dim
  int heads, hd;
  int tokens;       // N
  int frames;       // F
  int group;        // PACKED: N (tokens per sequence); TEMPORAL: G = 128 / F  (power of two)
  int gshift;       // log2(group)
  int k_head0, v_head0;  // index of head 0 of K / V in the (hd, heads-like, rows) view of the K/V buffer
  int kv_rows_per_batch; // CROSS: rows of the K/V buffer per sample (= text length L, the number of valid keys)
  int q_rows_per_batch;  // CROSS: query rows per sample (F * N)
  int Lk;           // keys per tile: N (FULL, <= 256) or 128
  int tiles_per_seq;  // FULL: N / 128; TEMPORAL: N / G
  float scale_log2; // hd^-0.5 * log2(e)
  int dbg;          // B200_ATTN_DBG bits (pipelined kernel only): 1 no TMA, 2 no softmax math, 4 no output stores, 8 no MMA
};

// Shared-memory plan (bytes, every piece 1 KiB aligned), sized by the key count Lk so that several CTAs fit per SM:
//   region A: [Q main 128x128B][K main Lk x128B][Q tail 128x32B][K tail Lk x32B]   -- later overwritten by P (128 x Lk 16-bit)
//   region V: [V main Lk x128B][V tail Lk x32B]
// Lk=256: 64 KiB + 40 KiB -> 2 CTAs/SM (TMEM 256 cols each); Lk=128: 40 KiB + 20 KiB -> 3 CTAs/SM (TMEM 128 cols each).
struct SmemPlan {
  int q_main, k_main, q_tail, k_tail, p, v_main, v_tail, bars, total;
};
__host__ __device__ inline SmemPlan make_plan(int Lk, bool tail) {
  SmemPlan s;
  s.q_main = 0;
  s.k_main = 128 * 128;
  s.q_tail = s.k_main + Lk * 128;
  s.k_tail = s.q_tail + (tail ? 128 * 32 : 0);
  const int qk_end = s.k_tail + (tail ? Lk * 32 : 0);
  const int p_bytes = 128 * Lk * 2;
  s.p = 0;
  const int region_a = qk_end > p_bytes ? qk_end : p_bytes;
  s.v_main = region_a;
  s.v_tail = s.v_main + Lk * 128;
  s.bars = s.v_tail + (tail ? Lk * 32 : 0);
  s.total = s.bars + 128 + 1024;
  return s;
}

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// group is a power of two: PACKED compares sequence ids (index >> shift), TEMPORAL compares token ids (index & mask)
template <int MODE>
__device__ __forceinline__ int row_key(int r, int gshift) {
  if constexpr (MODE == MODE_PACKED) return r >> gshift;
  if constexpr (MODE == MODE_TEMPORAL) return r & ((1 << gshift) - 1);
  return 0;
}
template <int MODE>
__device__ __forceinline__ bool key_valid(int rkey, int col, int gshift) {
  if constexpr (MODE == MODE_FULL) return true;
  if constexpr (MODE == MODE_CROSS) return col < rkey;   // rkey carries the number of valid keys (text tokens)
  return row_key<MODE>(col, gshift) == rkey;
}

template <bool BF16, bool TAIL, int MODE>
__global__ void __launch_bounds__(kThreads, 2)
attn_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmQt,
            const __grid_constant__ CUtensorMap tmKV, const __grid_constant__ CUtensorMap tmKVt, const AttnDev p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const SmemPlan sp = make_plan(p.Lk, TAIL);
  const int SQ_MAIN = sp.q_main, SK_MAIN = sp.k_main, SQ_TAIL = sp.q_tail, SK_TAIL = sp.k_tail, SP = sp.p,
            SV_MAIN = sp.v_main, SV_TAIL = sp.v_tail;
  const uint32_t kTmemCols = static_cast<uint32_t>(p.Lk);  // 128 or 256 (power of two >= 32)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + sp.bars);
  uint64_t* bar_qk = bars + 0;
  uint64_t* bar_v = bars + 1;
  uint64_t* bar_s = bars + 2;
  uint64_t* bar_p = bars + 3;
  uint64_t* bar_o = bars + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tile = blockIdx.x;
  const int head = blockIdx.y;
  const int Lk = p.Lk;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
    mbar_init(bar_qk, 1);
    mbar_init(bar_v, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_p, 128);
    mbar_init(bar_o, 1);
    fence_mbar_init();
  }
  if (warp == 4) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();   // qkv (written by the preceding GEMM) is visible from here

  if (warp == 4) {
    if (lane == 0) {
      // ---------------------------------------------------------------- TMA loads
      const uint32_t qk_bytes = 128 * 128 + Lk * 128 + (TAIL ? (128 * 32 + Lk * 32) : 0);
      const uint32_t v_bytes = Lk * 128 + (TAIL ? Lk * 32 : 0);
      mbar_arrive_expect_tx(bar_qk, qk_bytes);
      if constexpr (MODE == MODE_TEMPORAL) {
        const int b = tile / p.tiles_per_seq;
        const int n0 = (tile % p.tiles_per_seq) * p.group;
        const int bf0 = b * p.frames;
        tma_load_4d(smem + SQ_MAIN, &tmQ, bar_qk, 0, head, n0, bf0);
        tma_load_4d(smem + SK_MAIN, &tmKV, bar_qk, 0, p.k_head0 + head, n0, bf0);
        if constexpr (TAIL) {
          tma_load_4d(smem + SQ_TAIL, &tmQt, bar_qk, 64, head, n0, bf0);
          tma_load_4d(smem + SK_TAIL, &tmKVt, bar_qk, 64, p.k_head0 + head, n0, bf0);
        }
        mbar_arrive_expect_tx(bar_v, v_bytes);
        tma_load_4d(smem + SV_MAIN, &tmKV, bar_v, 0, p.v_head0 + head, n0, bf0);
        if constexpr (TAIL) tma_load_4d(smem + SV_TAIL, &tmKVt, bar_v, 64, p.v_head0 + head, n0, bf0);
      } else {
        int q_row0, kv_row0;
        if constexpr (MODE == MODE_FULL) {
          const int s = tile / p.tiles_per_seq;
          kv_row0 = s * p.tokens;
          q_row0 = kv_row0 + (tile % p.tiles_per_seq) * 128;
        } else if constexpr (MODE == MODE_CROSS) {
          q_row0 = tile * 128;                                         // 128 consecutive query tokens of one sample
          kv_row0 = (q_row0 / p.q_rows_per_batch) * p.kv_rows_per_batch;  // that sample's text tokens (<= 128 keys)
        } else {
          q_row0 = kv_row0 = tile * 128;
        }
        tma_load_3d(smem + SQ_MAIN, &tmQ, bar_qk, 0, head, q_row0);
        tma_load_3d(smem + SK_MAIN, &tmKV, bar_qk, 0, p.k_head0 + head, kv_row0);
        if constexpr (TAIL) {
          tma_load_3d(smem + SQ_TAIL, &tmQt, bar_qk, 64, head, q_row0);
          tma_load_3d(smem + SK_TAIL, &tmKVt, bar_qk, 64, p.k_head0 + head, kv_row0);
        }
        mbar_arrive_expect_tx(bar_v, v_bytes);
        tma_load_3d(smem + SV_MAIN, &tmKV, bar_v, 0, p.v_head0 + head, kv_row0);
        if constexpr (TAIL) tma_load_3d(smem + SV_TAIL, &tmKVt, bar_v, 64, p.v_head0 + head, kv_row0);
      }

      // ---------------------------------------------------------------- S = Q K^T   (M=128, N=Lk, K=hd padded to 16)
      mbar_wait(bar_qk, 0);
      tc_fence_after();
      const uint32_t idesc_s = umma_idesc_f16(BF16, 128, static_cast<uint32_t>(Lk), false, false);
      const uint64_t dq = umma_smem_desc(smem_u32(smem + SQ_MAIN), 0, 1024, UMMA_LAYOUT_SW128);
      const uint64_t dk = umma_smem_desc(smem_u32(smem + SK_MAIN), 0, 1024, UMMA_LAYOUT_SW128);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_f16_ss(tmem_base, umma_desc_advance(dq, k * 32), umma_desc_advance(dk, k * 32), idesc_s, k > 0 ? 1u : 0u);
      if constexpr (TAIL) {
        const uint64_t dqt = umma_smem_desc(smem_u32(smem + SQ_TAIL), 0, 256, UMMA_LAYOUT_SW32);
        const uint64_t dkt = umma_smem_desc(smem_u32(smem + SK_TAIL), 0, 256, UMMA_LAYOUT_SW32);
        umma_f16_ss(tmem_base, dqt, dkt, idesc_s, 1u);
      }
      umma_commit(bar_s);

      // ---------------------------------------------------------------- O = P V   (M=128, N=64 (+16), K=Lk)
      mbar_wait(bar_v, 0);
      mbar_wait(bar_p, 0);
      tc_fence_after();
      const uint32_t idesc_o = umma_idesc_f16(BF16, 128, 64, false, true);   // B = V is MN-major ([key][hd] in smem)
      const uint32_t idesc_ot = umma_idesc_f16(BF16, 128, 16, false, true);
      const uint64_t dv = umma_smem_desc(smem_u32(smem + SV_MAIN), static_cast<uint32_t>(Lk) * 128, 1024, UMMA_LAYOUT_SW128);
      const uint64_t dvt = umma_smem_desc(smem_u32(smem + SV_TAIL), static_cast<uint32_t>(Lk) * 32, 256, UMMA_LAYOUT_SW32);
      const int ksteps = Lk / 16;
      for (int k = 0; k < ksteps; ++k) {
        const uint64_t dp = umma_smem_desc(smem_u32(smem + SP + (k >> 2) * (128 * 128)) + (k & 3) * 32, 0, 1024, UMMA_LAYOUT_SW128);
        umma_f16_ss(tmem_base + kOCol, dp, umma_desc_advance(dv, k * 16 * 128), idesc_o, k > 0 ? 1u : 0u);
        if constexpr (TAIL)
          umma_f16_ss(tmem_base + kOCol + 64, dp, umma_desc_advance(dvt, k * 16 * 32), idesc_ot, k > 0 ? 1u : 0u);
      }
      umma_commit(bar_o);
    }
  } else {
    // ------------------------------------------------------------------ softmax + epilogue: thread = tile row
    const int r = warp * 32 + lane;
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    const int nchunks = Lk / 32;
    const int rkey = MODE == MODE_CROSS ? p.kv_rows_per_batch : row_key<MODE>(r, p.gshift);
    mbar_wait(bar_s, 0);
    tc_fence_after();

    float mx = -INFINITY;
    for (int c = 0; c < nchunks; ++c) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(t_row + c * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (key_valid<MODE>(rkey, c * 32 + j, p.gshift)) mx = fmaxf(mx, __uint_as_float(v[j]));
    }
    const float mscaled = mx * p.scale_log2;
    float sum = 0.f;
    uint8_t* prow = smem + SP + r * 128;
    const int sw = r & 7;
    for (int c = 0; c < nchunks; ++c) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(t_row + c * 32, v);
      tmem_ld_wait();
      float e[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float x = ex2(fmaf(__uint_as_float(v[j]), p.scale_log2, -mscaled));
        e[j] = key_valid<MODE>(rkey, c * 32 + j, p.gshift) ? x : 0.f;
        sum += e[j];
      }
      uint8_t* atom = prow + (c >> 1) * (128 * 128);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint4 o;
        o.x = pack2<BF16>(e[8 * i + 0], e[8 * i + 1]);
        o.y = pack2<BF16>(e[8 * i + 2], e[8 * i + 3]);
        o.z = pack2<BF16>(e[8 * i + 4], e[8 * i + 5]);
        o.w = pack2<BF16>(e[8 * i + 6], e[8 * i + 7]);
        const int chunk = ((c & 1) * 4 + i) ^ sw;  // 128B swizzle: 16-byte chunk index XOR (row % 8)
        *reinterpret_cast<uint4*>(atom + chunk * 16) = o;
      }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    mbar_arrive(bar_p);

    // output row of this thread
    long long out_row;
    bool row_ok = true;
    if constexpr (MODE == MODE_TEMPORAL) {
      const int b = tile / p.tiles_per_seq;
      const int n0 = (tile % p.tiles_per_seq) * p.group;
      out_row = (static_cast<long long>(b) * p.frames + r / p.group) * p.tokens + n0 + r % p.group;
    } else if constexpr (MODE == MODE_FULL) {
      out_row = static_cast<long long>(tile / p.tiles_per_seq) * p.tokens + (tile % p.tiles_per_seq) * 128 + r;
    } else {
      out_row = static_cast<long long>(tile) * 128 + r;
      row_ok = out_row < p.T;
    }
    uint16_t* optr = reinterpret_cast<uint16_t*>(p.out) + out_row * p.D + head * p.hd;
    const float inv = 1.0f / sum;

    mbar_wait(bar_o, 0);
    tc_fence_after();
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(t_row + kOCol + c * 32, v);
      tmem_ld_wait();
      if (row_ok) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 o;
          o.x = pack2<BF16>(__uint_as_float(v[8 * i + 0]) * inv, __uint_as_float(v[8 * i + 1]) * inv);
          o.y = pack2<BF16>(__uint_as_float(v[8 * i + 2]) * inv, __uint_as_float(v[8 * i + 3]) * inv);
          o.z = pack2<BF16>(__uint_as_float(v[8 * i + 4]) * inv, __uint_as_float(v[8 * i + 5]) * inv);
          o.w = pack2<BF16>(__uint_as_float(v[8 * i + 6]) * inv, __uint_as_float(v[8 * i + 7]) * inv);
          *reinterpret_cast<uint4*>(optr + c * 32 + i * 8) = o;
        }
      }
    }
    if constexpr (TAIL) {
      uint32_t v[16];
      tmem_ld_32x32b_x16(t_row + kOCol + 64, v);
      tmem_ld_wait();
      if (row_ok) {
        const int tail8 = (p.hd - 64) / 8;  // 1 (hd 72) or 2 (hd 80)
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          if (i >= tail8) break;
          uint4 o;
          o.x = pack2<BF16>(__uint_as_float(v[8 * i + 0]) * inv, __uint_as_float(v[8 * i + 1]) * inv);
          o.y = pack2<BF16>(__uint_as_float(v[8 * i + 2]) * inv, __uint_as_float(v[8 * i + 3]) * inv);
          o.z = pack2<BF16>(__uint_as_float(v[8 * i + 4]) * inv, __uint_as_float(v[8 * i + 5]) * inv);
          o.w = pack2<BF16>(__uint_as_float(v[8 * i + 6]) * inv, __uint_as_float(v[8 * i + 7]) * inv);
          *reinterpret_cast<uint4*>(optr + 64 + i * 8) = o;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Persistent, software-pipelined variant (the one the launchers use): one CTA per SM walks a list of (tile, head) work
// items with everything double-buffered, so the tensor pipe, the TMA engine and the two softmax warpgroups overlap
// across consecutive tiles instead of running back to back inside one short-lived CTA:
//   smem : 2 x {Q, K, V (+ head_dim tails)}             TMEM : 2 x 256 columns
//   per buffer: S = Q K^T -> cols [0,Lk) fp32;  P (16-bit, packed two per column) overwrites cols [0,Lk/2) in place
//               (tcgen05.st from the thread that owns the row);  O = P V (A operand from TMEM) -> cols [Lk/2, Lk/2+80).
//   warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM alloc), warps 2-5 / 6-9 = softmax+epilogue for even / odd tiles.
//   MMA issue order  S_0, S_1, PV_0, S_2, PV_1, ...: S_{i+1} runs on the tensor pipe while tile i is in its softmax.
// P never touches shared memory, which is what makes room for the second Q/K/V buffer.
constexpr int kPipeThreads = 320;

struct PipePlan {
  int q_main, k_main, v_main, q_tail, k_tail, v_tail, buf_bytes, bars, total;
};
__host__ __device__ inline PipePlan make_pipe_plan(int Lk, bool tail) {
  PipePlan s;
  s.q_main = 0;
  s.k_main = 128 * 128;
  s.v_main = s.k_main + Lk * 128;
  s.q_tail = s.v_main + Lk * 128;
  s.k_tail = s.q_tail + (tail ? 128 * 32 : 0);
  s.v_tail = s.k_tail + (tail ? Lk * 32 : 0);
  s.buf_bytes = s.v_tail + (tail ? Lk * 32 : 0);
  s.bars = 2 * s.buf_bytes;
  s.total = s.bars + 256 + 1024;
  return s;
}

template <bool BF16, bool TAIL, int MODE>
__global__ void __launch_bounds__(kPipeThreads, 1)
attn_pipe_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmQt,
                 const __grid_constant__ CUtensorMap tmKV, const __grid_constant__ CUtensorMap tmKVt, const AttnDev p,
                 const int total_items) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const PipePlan sp = make_pipe_plan(p.Lk, TAIL);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + sp.bars);
  uint64_t* qk_full = bars + 0;   // [2] TMA: Q and K of the buffer have landed
  uint64_t* v_full = bars + 2;    // [2] TMA: V has landed
  uint64_t* s_full = bars + 4;    // [2] MMA: S is complete in TMEM          (also: the Q/K smem of the buffer is free)
  uint64_t* p_full = bars + 6;    // [2] softmax (128 arrivals): P is in TMEM
  uint64_t* o_full = bars + 8;    // [2] MMA: O is complete in TMEM          (also: the V smem of the buffer is free)
  uint64_t* o_free = bars + 10;   // [2] epilogue (128 arrivals): O has been read, the TMEM buffer may be overwritten
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int Lk = p.Lk;
  const int H = p.heads;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
    for (int b = 0; b < 2; ++b) {
      mbar_init(qk_full + b, 1);
      mbar_init(v_full + b, 1);
      mbar_init(s_full + b, 1);
      mbar_init(p_full + b, 128);
      mbar_init(o_full + b, 1);
      mbar_init(o_free + b, 128);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();   // qkv (written by the preceding GEMM) is visible from here

  const int first = blockIdx.x, step = gridDim.x;

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------------------------------------------ TMA producer
      const uint32_t qk_bytes = 128 * 128 + Lk * 128 + (TAIL ? (128 * 32 + Lk * 32) : 0);
      const uint32_t v_bytes = Lk * 128 + (TAIL ? Lk * 32 : 0);
      int i = 0;
      for (int item = first; item < total_items; item += step, ++i) {
        const int b = i & 1;
        const uint32_t par = (i >> 1) & 1;
        uint8_t* buf = smem + b * sp.buf_bytes;
        const int head = item % H;
        const int tile = item / H;
        if (i >= 2) mbar_wait(s_full + b, par ^ 1);    // S of the tile that used this buffer is done: Q/K smem is free
        if (p.dbg & 1) {
          mbar_arrive(qk_full + b);
          if (i >= 2) mbar_wait(o_full + b, par ^ 1);
          mbar_arrive(v_full + b);
          continue;
        }
        mbar_arrive_expect_tx(qk_full + b, qk_bytes);
        if constexpr (MODE == MODE_TEMPORAL) {
          const int bb = tile / p.tiles_per_seq;
          const int n0 = (tile % p.tiles_per_seq) * p.group;
          const int bf0 = bb * p.frames;
          tma_load_4d(buf + sp.q_main, &tmQ, qk_full + b, 0, head, n0, bf0);
          tma_load_4d(buf + sp.k_main, &tmKV, qk_full + b, 0, p.k_head0 + head, n0, bf0);
          if constexpr (TAIL) {
            tma_load_4d(buf + sp.q_tail, &tmQt, qk_full + b, 64, head, n0, bf0);
            tma_load_4d(buf + sp.k_tail, &tmKVt, qk_full + b, 64, p.k_head0 + head, n0, bf0);
          }
          if (i >= 2) mbar_wait(o_full + b, par ^ 1);  // PV of the tile that used this buffer is done: V smem is free
          mbar_arrive_expect_tx(v_full + b, v_bytes);
          tma_load_4d(buf + sp.v_main, &tmKV, v_full + b, 0, p.v_head0 + head, n0, bf0);
          if constexpr (TAIL) tma_load_4d(buf + sp.v_tail, &tmKVt, v_full + b, 64, p.v_head0 + head, n0, bf0);
        } else {
          int q_row0, kv_row0;
          if constexpr (MODE == MODE_FULL) {
            const int s = tile / p.tiles_per_seq;
            kv_row0 = s * p.tokens;
            q_row0 = kv_row0 + (tile % p.tiles_per_seq) * 128;
          } else if constexpr (MODE == MODE_CROSS) {
            q_row0 = tile * 128;
            kv_row0 = (q_row0 / p.q_rows_per_batch) * p.kv_rows_per_batch;
          } else {
            q_row0 = kv_row0 = tile * 128;
          }
          tma_load_3d(buf + sp.q_main, &tmQ, qk_full + b, 0, head, q_row0);
          tma_load_3d(buf + sp.k_main, &tmKV, qk_full + b, 0, p.k_head0 + head, kv_row0);
          if constexpr (TAIL) {
            tma_load_3d(buf + sp.q_tail, &tmQt, qk_full + b, 64, head, q_row0);
            tma_load_3d(buf + sp.k_tail, &tmKVt, qk_full + b, 64, p.k_head0 + head, kv_row0);
          }
          if (i >= 2) mbar_wait(o_full + b, par ^ 1);
          mbar_arrive_expect_tx(v_full + b, v_bytes);
          tma_load_3d(buf + sp.v_main, &tmKV, v_full + b, 0, p.v_head0 + head, kv_row0);
          if constexpr (TAIL) tma_load_3d(buf + sp.v_tail, &tmKVt, v_full + b, 64, p.v_head0 + head, kv_row0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ------------------------------------------------------------------ MMA issuer
      const int n_items = first < total_items ? (total_items - first + step - 1) / step : 0;
      const uint32_t idesc_s = umma_idesc_f16(BF16, 128, static_cast<uint32_t>(Lk), false, false);
      const uint32_t idesc_o = umma_idesc_f16(BF16, 128, 64, false, true);   // B = V is MN-major ([key][hd] in smem)
      const uint32_t idesc_ot = umma_idesc_f16(BF16, 128, 16, false, true);
      const int ksteps = Lk / 16;
      auto issue_s = [&](int i) {
        const int b = i & 1;
        const uint32_t par = (i >> 1) & 1;
        uint8_t* buf = smem + b * sp.buf_bytes;
        mbar_wait(qk_full + b, par);
        if (i >= 2) mbar_wait(o_free + b, par ^ 1);     // the epilogue of the previous user has drained the TMEM buffer
        tc_fence_after();
        const uint32_t tS = tmem_base + b * 256;
        const uint64_t dq = umma_smem_desc(smem_u32(buf + sp.q_main), 0, 1024, UMMA_LAYOUT_SW128);
        const uint64_t dk = umma_smem_desc(smem_u32(buf + sp.k_main), 0, 1024, UMMA_LAYOUT_SW128);
        if (!(p.dbg & 8)) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16_ss(tS, umma_desc_advance(dq, k * 32), umma_desc_advance(dk, k * 32), idesc_s, k > 0 ? 1u : 0u);
          if constexpr (TAIL) {
            const uint64_t dqt = umma_smem_desc(smem_u32(buf + sp.q_tail), 0, 256, UMMA_LAYOUT_SW32);
            const uint64_t dkt = umma_smem_desc(smem_u32(buf + sp.k_tail), 0, 256, UMMA_LAYOUT_SW32);
            umma_f16_ss(tS, dqt, dkt, idesc_s, 1u);
          }
        }
        umma_commit(s_full + b);
      };
      if (n_items > 0) issue_s(0);
      for (int i = 0; i < n_items; ++i) {
        if (i + 1 < n_items) issue_s(i + 1);
        const int b = i & 1;
        const uint32_t par = (i >> 1) & 1;
        uint8_t* buf = smem + b * sp.buf_bytes;
        mbar_wait(v_full + b, par);
        mbar_wait(p_full + b, par);
        tc_fence_after();
        const uint32_t tP = tmem_base + b * 256;
        const uint32_t tO = tP + (Lk >> 1);
        const uint64_t dv = umma_smem_desc(smem_u32(buf + sp.v_main), static_cast<uint32_t>(Lk) * 128, 1024, UMMA_LAYOUT_SW128);
        const uint64_t dvt = umma_smem_desc(smem_u32(buf + sp.v_tail), static_cast<uint32_t>(Lk) * 32, 256, UMMA_LAYOUT_SW32);
        for (int k = 0; k < ksteps && !(p.dbg & 8); ++k) {
          umma_f16_ts(tO, tP + k * 8, umma_desc_advance(dv, k * 16 * 128), idesc_o, k > 0 ? 1u : 0u);
          if constexpr (TAIL) umma_f16_ts(tO + 64, tP + k * 8, umma_desc_advance(dvt, k * 16 * 32), idesc_ot, k > 0 ? 1u : 0u);
        }
        umma_commit(o_full + b);
      }
    }
  } else {
    // -------------------------------------------------------------------- softmax + epilogue: thread = tile row
    const int wg = (warp - 2) >> 2;            // warpgroup 0 takes the even work items of this CTA, 1 the odd ones
    const int q = warp & 3;                    // TMEM lane quarter this warp may access
    const int r = q * 32 + lane;
    const int nchunks = Lk / 32;
    const int rkey = MODE == MODE_CROSS ? p.kv_rows_per_batch : row_key<MODE>(r, p.gshift);
    const uint32_t t_row = tmem_base + wg * 256 + (static_cast<uint32_t>(q * 32) << 16);
    const uint32_t t_o = t_row + (Lk >> 1);
    int i = wg;
    for (int item = first + wg * step; item < total_items; item += 2 * step, i += 2) {
      const uint32_t par = (i >> 1) & 1;
      const int head = item % H;
      const int tile = item / H;
      mbar_wait(s_full + wg, par);
      tc_fence_after();

      const int nch = (p.dbg & 2) ? 0 : nchunks;   // Lk / 32: 4 or 8
      // pass 1: row maximum.  Two 32-column TMEM loads are in flight per wait (the loads, not the math, set the pace).
      float mx = -INFINITY;
      for (int c = 0; c < nch; c += 2) {
        uint32_t va[32], vb[32];
        tmem_ld_32x32b_x32(t_row + c * 32, va);
        tmem_ld_32x32b_x32(t_row + (c + 1) * 32, vb);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          if (key_valid<MODE>(rkey, c * 32 + j, p.gshift)) mx = fmaxf(mx, __uint_as_float(va[j]));
          if (key_valid<MODE>(rkey, (c + 1) * 32 + j, p.gshift)) mx = fmaxf(mx, __uint_as_float(vb[j]));
        }
      }
      const float mscaled = mx * p.scale_log2;
      float sum = 0.f;
      // pass 2: P = exp2(s * scale - max), written over the S columns already consumed.  The load of chunk c+1 is in flight
      // while chunk c is exponentiated (P chunk c lands in S chunk c/2 <= c, never in a chunk still being loaded).
      auto emit = [&](const uint32_t (&v)[32], int c) {
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float e0 = ex2(fmaf(__uint_as_float(v[2 * j]), p.scale_log2, -mscaled));
          float e1 = ex2(fmaf(__uint_as_float(v[2 * j + 1]), p.scale_log2, -mscaled));
          if (!key_valid<MODE>(rkey, c * 32 + 2 * j, p.gshift)) e0 = 0.f;
          if (!key_valid<MODE>(rkey, c * 32 + 2 * j + 1, p.gshift)) e1 = 0.f;
          sum += e0 + e1;
          pk[j] = pack2<BF16>(e0, e1);
        }
        tmem_st_32x32b_x16(t_row + c * 16, pk);
      };
      if (nch > 0) {
        uint32_t va[32], vb[32];
        tmem_ld_32x32b_x32(t_row, va);
        tmem_ld_wait();
        for (int c = 0; c < nch; c += 2) {
          tmem_ld_32x32b_x32(t_row + (c + 1) * 32, vb);
          emit(va, c);
          tmem_ld_wait();
          if (c + 2 < nch) tmem_ld_32x32b_x32(t_row + (c + 2) * 32, va);
          emit(vb, c + 1);
          tmem_ld_wait();
        }
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(p_full + wg);

      long long out_row;
      bool row_ok = true;
      if constexpr (MODE == MODE_TEMPORAL) {
        const int bb = tile / p.tiles_per_seq;
        const int n0 = (tile % p.tiles_per_seq) * p.group;
        out_row = (static_cast<long long>(bb) * p.frames + r / p.group) * p.tokens + n0 + r % p.group;
      } else if constexpr (MODE == MODE_FULL) {
        out_row = static_cast<long long>(tile / p.tiles_per_seq) * p.tokens + (tile % p.tiles_per_seq) * 128 + r;
      } else {
        out_row = static_cast<long long>(tile) * 128 + r;
        row_ok = out_row < p.T;
      }
      uint16_t* optr = reinterpret_cast<uint16_t*>(p.out) + out_row * p.D + head * p.hd;
      const float inv = 1.0f / sum;

      mbar_wait(o_full + wg, par);
      tc_fence_after();
      uint32_t o0[32], o1[32], o2[16];
      tmem_ld_32x32b_x32(t_o, o0);
      tmem_ld_32x32b_x32(t_o + 32, o1);
      if constexpr (TAIL) tmem_ld_32x32b_x16(t_o + 64, o2);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(o_free + wg);                   // O is in registers: the MMA warp may start S of the tile after next
      if (row_ok && !(p.dbg & 4)) {
#pragma unroll
        for (int i4 = 0; i4 < 4; ++i4) {
          uint4 o;
          o.x = pack2<BF16>(__uint_as_float(o0[8 * i4 + 0]) * inv, __uint_as_float(o0[8 * i4 + 1]) * inv);
          o.y = pack2<BF16>(__uint_as_float(o0[8 * i4 + 2]) * inv, __uint_as_float(o0[8 * i4 + 3]) * inv);
          o.z = pack2<BF16>(__uint_as_float(o0[8 * i4 + 4]) * inv, __uint_as_float(o0[8 * i4 + 5]) * inv);
          o.w = pack2<BF16>(__uint_as_float(o0[8 * i4 + 6]) * inv, __uint_as_float(o0[8 * i4 + 7]) * inv);
          *reinterpret_cast<uint4*>(optr + i4 * 8) = o;
        }
#pragma unroll
        for (int i4 = 0; i4 < 4; ++i4) {
          uint4 o;
          o.x = pack2<BF16>(__uint_as_float(o1[8 * i4 + 0]) * inv, __uint_as_float(o1[8 * i4 + 1]) * inv);
          o.y = pack2<BF16>(__uint_as_float(o1[8 * i4 + 2]) * inv, __uint_as_float(o1[8 * i4 + 3]) * inv);
          o.z = pack2<BF16>(__uint_as_float(o1[8 * i4 + 4]) * inv, __uint_as_float(o1[8 * i4 + 5]) * inv);
          o.w = pack2<BF16>(__uint_as_float(o1[8 * i4 + 6]) * inv, __uint_as_float(o1[8 * i4 + 7]) * inv);
          *reinterpret_cast<uint4*>(optr + 32 + i4 * 8) = o;
        }
        if constexpr (TAIL) {
          const int tail8 = (p.hd - 64) / 8;  // 1 (hd 72) or 2 (hd 80)
#pragma unroll
          for (int i4 = 0; i4 < 2; ++i4) {
            if (i4 >= tail8) break;
            uint4 o;
            o.x = pack2<BF16>(__uint_as_float(o2[8 * i4 + 0]) * inv, __uint_as_float(o2[8 * i4 + 1]) * inv);
            o.y = pack2<BF16>(__uint_as_float(o2[8 * i4 + 2]) * inv, __uint_as_float(o2[8 * i4 + 3]) * inv);
            o.z = pack2<BF16>(__uint_as_float(o2[8 * i4 + 4]) * inv, __uint_as_float(o2[8 * i4 + 5]) * inv);
            o.w = pack2<BF16>(__uint_as_float(o2[8 * i4 + 6]) * inv, __uint_as_float(o2[8 * i4 + 7]) * inv);
            *reinterpret_cast<uint4*>(optr + 64 + i4 * 8) = o;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <bool BF16, bool TAIL, int MODE>
int launch_pipe(const CUtensorMap* m, const AttnDev& p, dim3 grid, cudaStream_t stream) {
  auto kern = attn_pipe_kernel<BF16, TAIL, MODE>;
  static bool attr_set = false;
  if (!attr_set) {
    B200_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, make_pipe_plan(256, true).total));
    attr_set = true;
  }
  const int smem_bytes = make_pipe_plan(p.Lk, TAIL).total;
  const int total = static_cast<int>(grid.x * grid.y);
  int sms = 148;
  B200_TRY(device_sm_count(&sms));
  const int ctas = total < sms ? total : sms;
  B200_CHECK_CUDA(launch_pdl(kern, dim3(ctas), dim3(kPipeThreads), static_cast<size_t>(smem_bytes), stream, m[0], m[1], m[2], m[3], p, total));
  return B200_OK;
}

template <bool BF16, bool TAIL, int MODE>
int launch_mode(const CUtensorMap* m, const AttnDev& p, dim3 grid, cudaStream_t stream) {
  static const bool use_old = getenv("B200_ATTN_OLD") != nullptr;   // A/B switch: the one-tile-per-CTA kernel
  if (!use_old) return launch_pipe<BF16, TAIL, MODE>(m, p, grid, stream);
  auto kern = attn_kernel<BF16, TAIL, MODE>;
  static bool attr_set = false;
  if (!attr_set) {
    B200_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, make_plan(256, true).total));
    attr_set = true;
  }
  const int smem_bytes = make_plan(p.Lk, TAIL).total;
  B200_CHECK_CUDA(launch_pdl(kern, grid, dim3(kThreads), static_cast<size_t>(smem_bytes), stream, m[0], m[1], m[2], m[3], p));
  return B200_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// Long spatial sequences (N = 512 / 1024 tokens per frame: LatteT2V at 512 px): the keys are streamed in chunks of 256
// with the online-softmax recurrence.  S_c = Q K_c^T lands in TMEM columns [0,256); the softmax warps take the chunk
// maximum, rescale the running output O (TMEM columns [256,336), tcgen05.ld -> * alpha -> tcgen05.st) and the running
// sum, write P_c (16-bit, K-major SW128) to smem; O += P_c V_c accumulates in TMEM.  K_{c+1} is fetched as soon as
// S_c has been computed, V_{c+1} as soon as O += P_c V_c has been computed.
constexpr int L_SQ_MAIN = 0;
constexpr int L_SK_MAIN = L_SQ_MAIN + 128 * 128;
constexpr int L_SV_MAIN = L_SK_MAIN + 256 * 128;
constexpr int L_SP = L_SV_MAIN + 256 * 128;
constexpr int L_SQ_TAIL = L_SP + 128 * 256 * 2;
constexpr int L_SK_TAIL = L_SQ_TAIL + 128 * 32;
constexpr int L_SV_TAIL = L_SK_TAIL + 256 * 32;
constexpr int L_SBARS = L_SV_TAIL + 256 * 32;
constexpr int L_SMEM_BYTES = L_SBARS + 128 + 1024;
constexpr int L_OCOL = 256;

template <bool BF16, bool TAIL>
__global__ void __launch_bounds__(kThreads, 1)
attn_long_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmQt,
                 const __grid_constant__ CUtensorMap tmKV, const __grid_constant__ CUtensorMap tmKVt, const AttnDev p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L_SBARS);
  uint64_t* bar_q = bars + 0;
  uint64_t* bar_k = bars + 1;
  uint64_t* bar_v = bars + 2;
  uint64_t* bar_s = bars + 3;
  uint64_t* bar_p = bars + 4;
  uint64_t* bar_o = bars + 5;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tile = blockIdx.x;
  const int head = blockIdx.y;
  const int nchunk = p.tokens / 256;
  const int seq = tile / p.tiles_per_seq;
  const int kv_row0 = seq * p.tokens;
  const int q_row0 = kv_row0 + (tile % p.tiles_per_seq) * 128;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
    mbar_init(bar_q, 1);
    mbar_init(bar_k, 1);
    mbar_init(bar_v, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_p, 128);
    mbar_init(bar_o, 1);
    fence_mbar_init();
  }
  if (warp == 4) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 4) {
    if (lane == 0) {
      const uint32_t k_bytes = 256 * 128 + (TAIL ? 256 * 32 : 0);
      auto load_k = [&](int c) {
        mbar_arrive_expect_tx(bar_k, k_bytes);
        tma_load_3d(smem + L_SK_MAIN, &tmKV, bar_k, 0, p.k_head0 + head, kv_row0 + c * 256);
        if constexpr (TAIL) tma_load_3d(smem + L_SK_TAIL, &tmKVt, bar_k, 64, p.k_head0 + head, kv_row0 + c * 256);
      };
      auto load_v = [&](int c) {
        mbar_arrive_expect_tx(bar_v, k_bytes);
        tma_load_3d(smem + L_SV_MAIN, &tmKV, bar_v, 0, p.v_head0 + head, kv_row0 + c * 256);
        if constexpr (TAIL) tma_load_3d(smem + L_SV_TAIL, &tmKVt, bar_v, 64, p.v_head0 + head, kv_row0 + c * 256);
      };
      mbar_arrive_expect_tx(bar_q, 128 * 128 + (TAIL ? 128 * 32 : 0));
      tma_load_3d(smem + L_SQ_MAIN, &tmQ, bar_q, 0, head, q_row0);
      if constexpr (TAIL) tma_load_3d(smem + L_SQ_TAIL, &tmQt, bar_q, 64, head, q_row0);
      load_k(0);
      load_v(0);

      const uint32_t idesc_s = umma_idesc_f16(BF16, 128, 256, false, false);
      const uint32_t idesc_o = umma_idesc_f16(BF16, 128, 64, false, true);
      const uint32_t idesc_ot = umma_idesc_f16(BF16, 128, 16, false, true);
      const uint64_t dq = umma_smem_desc(smem_u32(smem + L_SQ_MAIN), 0, 1024, UMMA_LAYOUT_SW128);
      const uint64_t dk = umma_smem_desc(smem_u32(smem + L_SK_MAIN), 0, 1024, UMMA_LAYOUT_SW128);
      const uint64_t dqt = umma_smem_desc(smem_u32(smem + L_SQ_TAIL), 0, 256, UMMA_LAYOUT_SW32);
      const uint64_t dkt = umma_smem_desc(smem_u32(smem + L_SK_TAIL), 0, 256, UMMA_LAYOUT_SW32);
      const uint64_t dv = umma_smem_desc(smem_u32(smem + L_SV_MAIN), 256 * 128, 1024, UMMA_LAYOUT_SW128);
      const uint64_t dvt = umma_smem_desc(smem_u32(smem + L_SV_TAIL), 256 * 32, 256, UMMA_LAYOUT_SW32);
      mbar_wait(bar_q, 0);
      for (int c = 0; c < nchunk; ++c) {
        const uint32_t par = c & 1;
        mbar_wait(bar_k, par);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16_ss(tmem_base, umma_desc_advance(dq, k * 32), umma_desc_advance(dk, k * 32), idesc_s, k > 0 ? 1u : 0u);
        if constexpr (TAIL) umma_f16_ss(tmem_base, dqt, dkt, idesc_s, 1u);
        umma_commit(bar_s);
        mbar_wait(bar_s, par);                 // S_c computed: the K buffer is free
        if (c + 1 < nchunk) load_k(c + 1);
        mbar_wait(bar_v, par);
        mbar_wait(bar_p, par);                 // P_c written, O rescaled
        tc_fence_after();
        for (int k = 0; k < 16; ++k) {
          const uint64_t dp = umma_smem_desc(smem_u32(smem + L_SP + (k >> 2) * (128 * 128)) + (k & 3) * 32, 0, 1024, UMMA_LAYOUT_SW128);
          const uint32_t accum = (c > 0 || k > 0) ? 1u : 0u;
          umma_f16_ss(tmem_base + L_OCOL, dp, umma_desc_advance(dv, k * 16 * 128), idesc_o, accum);
          if constexpr (TAIL) umma_f16_ss(tmem_base + L_OCOL + 64, dp, umma_desc_advance(dvt, k * 16 * 32), idesc_ot, accum);
        }
        umma_commit(bar_o);
        if (c + 1 < nchunk) {
          mbar_wait(bar_o, par);               // O += P_c V_c computed: the V buffer (and P) are free
          load_v(c + 1);
        }
      }
    }
  } else {
    const int r = warp * 32 + lane;
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    float m_run = -INFINITY, l_run = 0.f;
    uint8_t* prow = smem + L_SP + r * 128;
    const int sw = r & 7;
    for (int c = 0; c < nchunk; ++c) {
      const uint32_t par = c & 1;
      mbar_wait(bar_s, par);
      tc_fence_after();
      float cmax = -INFINITY;
      for (int j8 = 0; j8 < 8; ++j8) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(t_row + j8 * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) cmax = fmaxf(cmax, __uint_as_float(v[j]));
      }
      const float m_new = fmaxf(m_run, cmax);
      const float alpha = ex2((m_run - m_new) * p.scale_log2);   // 0 for the first chunk (m_run = -inf)
      l_run *= alpha;
      m_run = m_new;
      if (c > 0) {
        mbar_wait(bar_o, (c - 1) & 1);          // O += P_{c-1} V_{c-1} has completed: O and the P buffer are ours again
        tc_fence_after();
#pragma unroll 1
        for (int cc = 0; cc < (TAIL ? 5 : 4); ++cc) {
          uint32_t v[16];
          tmem_ld_32x32b_x16(t_row + L_OCOL + cc * 16, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) * alpha);
          tmem_st_32x32b_x16(t_row + L_OCOL + cc * 16, v);
        }
        tmem_st_wait();
      }
      const float mscaled = m_new * p.scale_log2;
      for (int j8 = 0; j8 < 8; ++j8) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(t_row + j8 * 32, v);
        tmem_ld_wait();
        float e[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          e[j] = ex2(fmaf(__uint_as_float(v[j]), p.scale_log2, -mscaled));
          l_run += e[j];
        }
        uint8_t* atom = prow + (j8 >> 1) * (128 * 128);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 o;
          o.x = pack2<BF16>(e[8 * i + 0], e[8 * i + 1]);
          o.y = pack2<BF16>(e[8 * i + 2], e[8 * i + 3]);
          o.z = pack2<BF16>(e[8 * i + 4], e[8 * i + 5]);
          o.w = pack2<BF16>(e[8 * i + 6], e[8 * i + 7]);
          const int chunk = ((j8 & 1) * 4 + i) ^ sw;
          *reinterpret_cast<uint4*>(atom + chunk * 16) = o;
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(bar_p);
    }
    const long long out_row = q_row0 + r;
    uint16_t* optr = reinterpret_cast<uint16_t*>(p.out) + out_row * p.D + head * p.hd;
    const float inv = 1.0f / l_run;
    mbar_wait(bar_o, (nchunk - 1) & 1);
    tc_fence_after();
#pragma unroll 1
    for (int c2 = 0; c2 < 2; ++c2) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(t_row + L_OCOL + c2 * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint4 o;
        o.x = pack2<BF16>(__uint_as_float(v[8 * i + 0]) * inv, __uint_as_float(v[8 * i + 1]) * inv);
        o.y = pack2<BF16>(__uint_as_float(v[8 * i + 2]) * inv, __uint_as_float(v[8 * i + 3]) * inv);
        o.z = pack2<BF16>(__uint_as_float(v[8 * i + 4]) * inv, __uint_as_float(v[8 * i + 5]) * inv);
        o.w = pack2<BF16>(__uint_as_float(v[8 * i + 6]) * inv, __uint_as_float(v[8 * i + 7]) * inv);
        *reinterpret_cast<uint4*>(optr + c2 * 32 + i * 8) = o;
      }
    }
    if constexpr (TAIL) {
      uint32_t v[16];
      tmem_ld_32x32b_x16(t_row + L_OCOL + 64, v);
      tmem_ld_wait();
      const int tail8 = (p.hd - 64) / 8;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        if (i >= tail8) break;
        uint4 o;
        o.x = pack2<BF16>(__uint_as_float(v[8 * i + 0]) * inv, __uint_as_float(v[8 * i + 1]) * inv);
        o.y = pack2<BF16>(__uint_as_float(v[8 * i + 2]) * inv, __uint_as_float(v[8 * i + 3]) * inv);
        o.z = pack2<BF16>(__uint_as_float(v[8 * i + 4]) * inv, __uint_as_float(v[8 * i + 5]) * inv);
        o.w = pack2<BF16>(__uint_as_float(v[8 * i + 6]) * inv, __uint_as_float(v[8 * i + 7]) * inv);
        *reinterpret_cast<uint4*>(optr + 64 + i * 8) = o;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <bool BF16, bool TAIL>
int launch_long(const CUtensorMap* m, const AttnDev& p, dim3 grid, cudaStream_t stream) {
  auto kern = attn_long_kernel<BF16, TAIL>;
  static bool attr_set = false;
  if (!attr_set) {
    B200_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L_SMEM_BYTES));
    attr_set = true;
  }
  B200_CHECK_CUDA(launch_pdl(kern, grid, dim3(kThreads), static_cast<size_t>(L_SMEM_BYTES), stream, m[0], m[1], m[2], m[3], p));
  return B200_OK;
}

template <bool BF16, bool TAIL>
int launch_tail(int mode, const CUtensorMap* m, const AttnDev& p, dim3 grid, cudaStream_t s) {
  switch (mode) {
    case MODE_FULL: return launch_mode<BF16, TAIL, MODE_FULL>(m, p, grid, s);
    case MODE_PACKED: return launch_mode<BF16, TAIL, MODE_PACKED>(m, p, grid, s);
    case MODE_CROSS: return launch_mode<BF16, TAIL, MODE_CROSS>(m, p, grid, s);
    default: return launch_mode<BF16, TAIL, MODE_TEMPORAL>(m, p, grid, s);
  }
}

}  // namespace

int launch_attention(const AttnArgs& a, cudaStream_t stream) {
  B200_REQUIRE(a.batch > 0 && a.frames > 0 && a.tokens > 0 && a.heads > 0, B200_ERR_SHAPE, "attention: bad shape");
  B200_REQUIRE(a.head_dim == 64 || a.head_dim == 72 || a.head_dim == 80, B200_ERR_UNSUPPORTED,
               "attention: head_dim %d unsupported (64, 72, 80)", a.head_dim);
  B200_REQUIRE((reinterpret_cast<uintptr_t>(a.qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.out) & 15) == 0,
               B200_ERR_ALIGN, "attention: qkv/out must be 16-byte aligned");
  B200_TRY(check_arch());
  const int H = a.heads, hd = a.head_dim, D = H * hd;
  const long long T = static_cast<long long>(a.batch) * a.frames * a.tokens;
  const bool tail = hd > 64;

  AttnDev p;
  p.out = a.out;
  p.T = static_cast<int>(T);
  p.D = D;
  p.heads = H;
  p.hd = hd;
  p.tokens = a.tokens;
  p.frames = a.frames;
  p.scale_log2 = (1.0f / sqrtf(static_cast<float>(hd))) * 1.4426950408889634f;
  { const char* e = getenv("B200_ATTN_DBG"); p.dbg = e ? atoi(e) : 0; }

  p.group = 1;
  p.gshift = 0;
  p.k_head0 = H;
  p.v_head0 = 2 * H;
  p.kv_rows_per_batch = 0;
  p.q_rows_per_batch = 1;
  auto ilog2 = [](int v) { int s = 0; while ((1 << s) < v) ++s; return s; };

  CUtensorMap maps[4];
  int mode;
  dim3 grid;
  if (!a.temporal) {
    const uint64_t dims[3] = {static_cast<uint64_t>(hd), static_cast<uint64_t>(3 * H), static_cast<uint64_t>(T)};
    const uint64_t str[2] = {static_cast<uint64_t>(hd) * 2, static_cast<uint64_t>(3 * D) * 2};
    uint32_t kv_rows;
    if (a.tokens > 256) {
      B200_REQUIRE(a.tokens % 256 == 0, B200_ERR_UNSUPPORTED, "attention: spatial sequence length %d must be a multiple of 256", a.tokens);
      p.Lk = 256;
      p.tiles_per_seq = a.tokens / 128;
      const uint32_t boxQ[3] = {64, 1, 128}, boxQt[3] = {16, 1, 128};
      const uint32_t boxK[3] = {64, 1, 256}, boxKt[3] = {16, 1, 256};
      B200_TRY(make_tmap_16bit(&maps[0], a.qkv, 3, dims, str, boxQ, TMAP_SW_128));
      B200_TRY(make_tmap_16bit(&maps[2], a.qkv, 3, dims, str, boxK, TMAP_SW_128));
      if (tail) {
        B200_TRY(make_tmap_16bit(&maps[1], a.qkv, 3, dims, str, boxQt, TMAP_SW_32));
        B200_TRY(make_tmap_16bit(&maps[3], a.qkv, 3, dims, str, boxKt, TMAP_SW_32));
      } else {
        maps[1] = maps[0];
        maps[3] = maps[2];
      }
      const dim3 lgrid(a.batch * a.frames * p.tiles_per_seq, H);
      if (a.bf16) return tail ? launch_long<true, true>(maps, p, lgrid, stream) : launch_long<true, false>(maps, p, lgrid, stream);
      return tail ? launch_long<false, true>(maps, p, lgrid, stream) : launch_long<false, false>(maps, p, lgrid, stream);
    }
    if (a.tokens >= 128) {
      B200_REQUIRE(a.tokens == 128 || a.tokens == 256, B200_ERR_UNSUPPORTED,
                   "attention: spatial sequence length %d unsupported (<=64 power of two, 128, 256, multiples of 256)", a.tokens);
      mode = MODE_FULL;
      p.Lk = a.tokens;
      p.tiles_per_seq = a.tokens / 128;
      kv_rows = a.tokens;
      grid = dim3(a.batch * a.frames * p.tiles_per_seq, H);
    } else {
      B200_REQUIRE(128 % a.tokens == 0, B200_ERR_UNSUPPORTED, "attention: spatial sequence length %d must divide 128", a.tokens);
      mode = MODE_PACKED;
      p.Lk = 128;
      p.group = a.tokens;
      p.gshift = ilog2(a.tokens);
      p.tiles_per_seq = 1;
      kv_rows = 128;
      grid = dim3(static_cast<unsigned>((T + 127) / 128), H);
    }
    const uint32_t boxQ[3] = {64, 1, 128}, boxQt[3] = {16, 1, 128};
    const uint32_t boxK[3] = {64, 1, kv_rows}, boxKt[3] = {16, 1, kv_rows};
    B200_TRY(make_tmap_16bit(&maps[0], a.qkv, 3, dims, str, boxQ, TMAP_SW_128));
    B200_TRY(make_tmap_16bit(&maps[2], a.qkv, 3, dims, str, boxK, TMAP_SW_128));
    if (tail) {
      B200_TRY(make_tmap_16bit(&maps[1], a.qkv, 3, dims, str, boxQt, TMAP_SW_32));
      B200_TRY(make_tmap_16bit(&maps[3], a.qkv, 3, dims, str, boxKt, TMAP_SW_32));
    } else {
      maps[1] = maps[0];
      maps[3] = maps[2];
    }
  } else {
    const int F = a.frames;
    B200_REQUIRE(F >= 4 && F <= 128 && 128 % F == 0, B200_ERR_UNSUPPORTED,
                 "attention: temporal length %d unsupported (power of two in [4,128])", F);
    const int G = 128 / F;
    B200_REQUIRE(a.tokens % G == 0, B200_ERR_UNSUPPORTED, "attention: tokens %d must be a multiple of %d", a.tokens, G);
    mode = MODE_TEMPORAL;
    p.Lk = 128;
    p.group = G;
    p.gshift = ilog2(G);
    p.tiles_per_seq = a.tokens / G;
    grid = dim3(a.batch * p.tiles_per_seq, H);
    const uint64_t dims[4] = {static_cast<uint64_t>(hd), static_cast<uint64_t>(3 * H), static_cast<uint64_t>(a.tokens),
                              static_cast<uint64_t>(a.batch) * F};
    const uint64_t str[3] = {static_cast<uint64_t>(hd) * 2, static_cast<uint64_t>(3 * D) * 2,
                             static_cast<uint64_t>(3 * D) * 2 * a.tokens};
    const uint32_t box[4] = {64, 1, static_cast<uint32_t>(G), static_cast<uint32_t>(F)};
    const uint32_t boxt[4] = {16, 1, static_cast<uint32_t>(G), static_cast<uint32_t>(F)};
    B200_TRY(make_tmap_16bit(&maps[0], a.qkv, 4, dims, str, box, TMAP_SW_128));
    maps[2] = maps[0];
    if (tail) {
      B200_TRY(make_tmap_16bit(&maps[1], a.qkv, 4, dims, str, boxt, TMAP_SW_32));
      maps[3] = maps[1];
    } else {
      maps[1] = maps[0];
      maps[3] = maps[0];
    }
  }
  if (a.bf16) return tail ? launch_tail<true, true>(mode, maps, p, grid, stream) : launch_tail<true, false>(mode, maps, p, grid, stream);
  return tail ? launch_tail<false, true>(mode, maps, p, grid, stream) : launch_tail<false, false>(mode, maps, p, grid, stream);
}

int launch_cross_attention(const CrossAttnArgs& a, cudaStream_t stream) {
  B200_REQUIRE(a.batch > 0 && a.q_rows_per_batch > 0 && a.heads > 0, B200_ERR_SHAPE, "cross attention: bad shape");
  B200_REQUIRE(a.head_dim == 64 || a.head_dim == 72 || a.head_dim == 80, B200_ERR_UNSUPPORTED,
               "cross attention: head_dim %d unsupported (64, 72, 80)", a.head_dim);
  B200_REQUIRE(a.kv_len >= 1 && a.kv_len <= 128, B200_ERR_UNSUPPORTED, "cross attention: %d keys per sample (1..128 built)", a.kv_len);
  B200_REQUIRE(a.q_rows_per_batch % 128 == 0, B200_ERR_UNSUPPORTED, "cross attention: query rows per sample %d must be a multiple of 128",
               a.q_rows_per_batch);
  B200_REQUIRE(((reinterpret_cast<uintptr_t>(a.q) | reinterpret_cast<uintptr_t>(a.kv) | reinterpret_cast<uintptr_t>(a.out)) & 15) == 0,
               B200_ERR_ALIGN, "cross attention: q/kv/out must be 16-byte aligned");
  B200_TRY(check_arch());
  const int H = a.heads, hd = a.head_dim, D = H * hd;
  const long long T = static_cast<long long>(a.batch) * a.q_rows_per_batch;
  const long long R = static_cast<long long>(a.batch) * a.kv_len;
  const bool tail = hd > 64;
  AttnDev p{};
  p.out = a.out; p.T = static_cast<int>(T); p.D = D; p.heads = H; p.hd = hd;
  p.tokens = a.q_rows_per_batch; p.frames = 1; p.group = 1; p.gshift = 0; p.Lk = 128; p.tiles_per_seq = 1;
  p.scale_log2 = (1.0f / sqrtf(static_cast<float>(hd))) * 1.4426950408889634f;
  p.k_head0 = 0; p.v_head0 = H; p.kv_rows_per_batch = a.kv_len; p.q_rows_per_batch = a.q_rows_per_batch;
  CUtensorMap maps[4];
  const uint64_t qdims[3] = {static_cast<uint64_t>(hd), static_cast<uint64_t>(H), static_cast<uint64_t>(T)};
  const uint64_t qstr[2] = {static_cast<uint64_t>(hd) * 2, static_cast<uint64_t>(a.q_row_stride) * 2};
  const uint64_t kdims[3] = {static_cast<uint64_t>(hd), static_cast<uint64_t>(2 * H), static_cast<uint64_t>(R)};
  const uint64_t kstr[2] = {static_cast<uint64_t>(hd) * 2, static_cast<uint64_t>(a.kv_row_stride) * 2};
  const uint32_t box[3] = {64, 1, 128}, boxt[3] = {16, 1, 128};
  B200_REQUIRE(a.q_row_stride >= D && a.kv_row_stride >= 2 * D && a.q_row_stride % 8 == 0 && a.kv_row_stride % 8 == 0, B200_ERR_SHAPE, "cross attention: bad row strides");
  B200_TRY(make_tmap_16bit(&maps[0], a.q, 3, qdims, qstr, box, TMAP_SW_128));
  B200_TRY(make_tmap_16bit(&maps[2], a.kv, 3, kdims, kstr, box, TMAP_SW_128));
  if (tail) {
    B200_TRY(make_tmap_16bit(&maps[1], a.q, 3, qdims, qstr, boxt, TMAP_SW_32));
    B200_TRY(make_tmap_16bit(&maps[3], a.kv, 3, kdims, kstr, boxt, TMAP_SW_32));
  } else {
    maps[1] = maps[0];
    maps[3] = maps[2];
  }
  const dim3 grid(static_cast<unsigned>(T / 128), H);
  if (a.bf16) return tail ? launch_tail<true, true>(MODE_CROSS, maps, p, grid, stream) : launch_tail<true, false>(MODE_CROSS, maps, p, grid, stream);
  return tail ? launch_tail<false, true>(MODE_CROSS, maps, p, grid, stream) : launch_tail<false, false>(MODE_CROSS, maps, p, grid, stream);
}

}  // namespace b200
