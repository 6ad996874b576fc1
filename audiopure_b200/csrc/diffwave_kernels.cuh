// DiffWave epsilon-network kernels for sm_100a.
//
// Data layout in HBM:
//   (bf16 build described; in the tf32 build the same tensors hold fp32 words, see Mode<>)
//   h      : [B][L][256] bf16, channels-last -- a time step is one 512-byte row, so every conv tap of a
//            128-step tile is one TMA box at row offset l0 + (tap-1)*dilation; the 3-D tensor map (c, l, b)
//            zero-fills rows outside [0, L), which IS the conv's zero padding and keeps taps from bleeding
//            across clips.  h_n already contains the "+ fc_t(emb)" shift of layer n (WaveNet.py:82-84).
//   gate   : [layers][B][L][256] bf16 -- 2 * tanh*sigmoid output of every layer, kept so that the 36 skip
//            projections become ONE K = layers*256 GEMM in the tail kernel (fp32 accumulation in TMEM)
//            instead of a 16 MB/clip fp32 read-modify-write per layer.
//
// Orientation of every GEMM: M = 128 time steps (TMEM lanes), N = 256 output channels (TMEM columns),
// K = input channels (x taps).  A = activations, B = weights, both K-major, 128-byte swizzle.
#pragma once

#include "sm100.cuh"

namespace ap {

constexpr int kC = 256;          // residual / skip / gate channels (the only width the kernels support)
constexpr int kTileT = 128;      // time steps per tile = UMMA M
constexpr uint32_t kABytes = kTileT * 128;  // [128 rows x 128 bytes]: one K step of activations (64 bf16 / 32 tf32)
constexpr int kThreads = 384;    // warp 0: TMA, warp 1: MMA + TMEM alloc, warp 2: x loads (layer kernel), warp 3: idle, warps 4-11: epilogue
constexpr int kEpiWarp0 = 4;
constexpr int kEpiThreads = 256;
constexpr uint32_t kTmemCols = 512;
constexpr float kSqrtHalf = 0.70710678118654752440f;


// A compile-time integer that converts to int in device code (std::integral_constant's conversion is host-only).
template <int V>
struct IntC {
  __device__ constexpr operator int() const { return V; }
};

__device__ __forceinline__ uint8_t* align1024(uint8_t* p) {
  return reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(p) + 1023) & ~static_cast<uintptr_t>(1023));
}

// ---------------------------------------------------------------------------------------------------
// Philox4x32-10, counter-based: noise is a pure function of (seed, stream, element index), so how a batch
// is sharded over GPUs or chunked over launches never changes it.
// ---------------------------------------------------------------------------------------------------
__host__ __device__ inline void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                               uint32_t k1, uint32_t (&out)[4]) {
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = static_cast<uint64_t>(0xD2511F53u) * c0;
    const uint64_t p1 = static_cast<uint64_t>(0xCD9E8D57u) * c2;
    const uint32_t n0 = static_cast<uint32_t>(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n1 = static_cast<uint32_t>(p1);
    const uint32_t n2 = static_cast<uint32_t>(p0 >> 32) ^ c3 ^ k1;
    const uint32_t n3 = static_cast<uint32_t>(p0);
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// Standard normal for element `idx` of stream (stream_hi, stream_lo) under `seed` (Box-Muller on one
// Philox block per 4 consecutive elements).
__device__ __forceinline__ float philox_normal(uint64_t seed, uint32_t stream_lo, uint32_t stream_hi, uint64_t idx) {
  uint32_t r[4];
  const uint64_t blk = idx >> 2;
  philox4x32_10(static_cast<uint32_t>(blk), static_cast<uint32_t>(blk >> 32), stream_lo, stream_hi,
                static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32), r);
  const int sel = static_cast<int>(idx & 3);
  const uint32_t a = r[(sel >> 1) * 2], b = r[(sel >> 1) * 2 + 1];
  const float u1 = (static_cast<float>(a >> 8) + 1.0f) * (1.0f / 16777216.0f);  // (0, 1]
  const float u2 = static_cast<float>(b >> 8) * (1.0f / 16777216.0f);           // [0, 1)
  const float rad = sqrtf(-2.0f * logf(u1));
  float s, c;
  sincospif(2.0f * u2, &s, &c);
  return rad * ((sel & 1) ? s : c);
}

// The four normals of Philox block `blk` (elements 4*blk .. 4*blk+3 of the stream): the same values philox_normal
// returns one at a time, for kernels that walk a row in float4 steps.
__device__ __forceinline__ void philox_normal4(uint64_t seed, uint32_t stream_lo, uint32_t stream_hi, uint64_t blk,
                                               float (&out)[4]) {
  uint32_t r[4];
  philox4x32_10(static_cast<uint32_t>(blk), static_cast<uint32_t>(blk >> 32), stream_lo, stream_hi,
                static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32), r);
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const float u1 = (static_cast<float>(r[2 * h] >> 8) + 1.0f) * (1.0f / 16777216.0f);
    const float u2 = static_cast<float>(r[2 * h + 1] >> 8) * (1.0f / 16777216.0f);
    const float rad = sqrtf(-2.0f * logf(u1));
    float sn, cs;
    sincospif(2.0f * u2, &sn, &cs);
    out[2 * h] = rad * cs;
    out[2 * h + 1] = rad * sn;
  }
}

// ---------------------------------------------------------------------------------------------------
// K0: init conv (1 -> 256, k = 1, weight-norm folded) + ReLU + layer-0 step shift, fp32 [B][L] -> bf16
// [B][L][256].  WaveNet.py:147,168 then :82-84 of block 0.  HBM-bound: 4 B in, 512 B out per time step.
// ---------------------------------------------------------------------------------------------------
template <bool kTf32>
__global__ void __launch_bounds__(256) prologue_kernel(const float* __restrict__ x, const float* __restrict__ w0,
                                                       const float* __restrict__ b0,
                                                       const float* __restrict__ part0,
                                                       void* __restrict__ h, long long rows, uint32_t round_bias) {
  const int cg = threadIdx.x & 31;  // this thread's 8 channels
  float w[8], b[8], p[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    w[j] = w0[cg * 8 + j];
    b[j] = b0[cg * 8 + j];
    p[j] = part0[cg * 8 + j];
  }
  const long long stride = static_cast<long long>(gridDim.x) * (blockDim.x >> 5);
  for (long long row = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5); row < rows;
       row += stride) {
    const float xv = __ldg(x + row);
    if constexpr (kTf32) {
      uint4 o[2];
      uint32_t* ow = reinterpret_cast<uint32_t*>(o);
#pragma unroll
      for (int j = 0; j < 8; ++j) ow[j] = __float_as_uint(fmaxf(fmaf(w[j], xv, b[j]), 0.f) + p[j]) + round_bias;
      uint4* dst = reinterpret_cast<uint4*>(static_cast<float*>(h) + row * kC) + 2 * cg;
      dst[0] = o[0];
      dst[1] = o[1];
    } else {
      uint4 o;
      uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float v0 = fmaxf(fmaf(w[2 * j], xv, b[2 * j]), 0.f) + p[2 * j];
        const float v1 = fmaxf(fmaf(w[2 * j + 1], xv, b[2 * j + 1]), 0.f) + p[2 * j + 1];
        ow[j] = pack_bf16x2(v0, v1);
      }
      reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(h) + row * kC)[cg] = o;
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// Shared plumbing of the two tensor-core kernels.
//
// Both run as CTA PAIRS: a cluster of two CTAs (the two SMs of a TPC) per pair of 128-step tiles, tcgen05
// cta_group::2 (M = 256).  Each CTA stages its own 128 activation rows and HALF of the weight rows, the
// leader CTA (cluster rank 0) issues the MMAs for both, each CTA's TMEM receives its own tile's accumulators
// and each CTA runs its own epilogue.  Versus one CTA per tile this halves the weight bytes written into
// each SM's shared memory and the weight bytes the tensor core reads back out of it -- shared-memory
// bandwidth (128 B/clk/SM), not the tensor pipe, is what caps the single-CTA form (measured 0.95 ms vs
// 0.87 ms per layer launch at B = 64 before the epilogue rework).
// ---------------------------------------------------------------------------------------------------
struct Tc {
  static constexpr uint32_t kBRows = 128;           // weight rows staged per CTA per K step (half of N = 256)
  static constexpr uint32_t kBBytes = kBRows * 128;
  static constexpr uint32_t kStageBytes = kABytes + kBBytes;
  static constexpr uint32_t kEpiWarps = kEpiThreads / 32;
};

// The two precisions the tensor-core kernels are instantiated for.  bf16: operands bf16 in HBM and shared memory
// (kind::f16).  tf32: operands are fp32 words (kind::tf32, the tensor core reads sign, exponent and 10 mantissa
// bits); everything that is 64 bf16 wide in the bf16 build -- a 128-byte swizzled operand row, a TMA box, a K
// step -- is 32 fp32 wide, so there are twice as many K steps and sub-tiles and the [128 x 256] operand tile is
// 128 KB instead of 64 KB (which leaves room for a 3-stage ring only).
template <bool kTf32>
struct Mode : Tc {
  static constexpr int kElem = kTf32 ? 4 : 2;
  static constexpr int kSubK = 128 / kElem;         // channels per 128-byte operand row (one K step)
  static constexpr int kSubs = kC / kSubK;          // [128 x kSubK] sub-tiles per 256 channels
  static constexpr int kSubsPer64 = 64 / kSubK;     // sub-tiles per 64-channel epilogue chunk
  static constexpr uint32_t kTileBytes = kTileT * kC * kElem;  // a full [128 x 256] operand tile
  static constexpr uint32_t kIdesc = kTf32 ? umma_idesc_tf32(256, 256) : umma_idesc_bf16(256, 256);
  // TMA -> MMA ring depth.  The ring is latency-sensitive (3 -> 4 stages: -7 % layer, -11 % tail time), so
  // everything else in shared memory is squeezed into ONE operand tile per kernel to afford 5 stages (bf16).
#ifndef AP_LAYER_STAGES
#define AP_LAYER_STAGES 5
#endif
  // tf32 layer kernel: the operand tile holds HALF of the 256 channels at a time (kSplit): the gate of chunk 0 is
  // consumed by the first half of GEMM2 before chunk 1's gate overwrites it, and the residual input is re-loaded and
  // turned into h_next in two halves.  That frees 64 KB: 5 ring stages instead of 3 (the ring is latency-sensitive).
  static constexpr bool kSplit = kTf32;
#ifdef AP_TF32_NO_SPLIT  // A/B variant: the r01 layout (full 128 KB operand tile, 3 stages)
#undef AP_TF32_NO_SPLIT
#define AP_TF32_NO_SPLIT 1
#else
#define AP_TF32_NO_SPLIT 0
#endif
  static constexpr bool kSplitTile = kSplit && !AP_TF32_NO_SPLIT;
  static constexpr int kTileSubs = kSplitTile ? kSubs / 2 : kSubs;   // sub-tiles resident in the layer kernel's operand tile
  static constexpr uint32_t kLayerTileBytes = kTileSubs * kABytes;
  static constexpr int kLayerStages = kTf32 ? (kSplitTile ? 5 : 3) : AP_LAYER_STAGES;
  // tail: 4 -> 5 stages (possible since the bias tables moved to the constant bank): 4.48 -> 3.68 ms per launch
  static constexpr int kTailStages = kTf32 ? 3 : 5;
  static constexpr uint32_t kLayerSmem = kLayerStages * kStageBytes + kLayerTileBytes + 32 * 8 + 1024;
  static constexpr uint32_t kTailSmem = kTailStages * kStageBytes + kTileBytes + 2 * 128 * 4 + 32 * 8 + 1024;
  static_assert(kLayerStages <= 6 && kTailStages <= 6, "barrier slots");
  static_assert(kLayerSmem <= 232448 && kTailSmem <= 232448, "over the 227 KB per-CTA shared-memory limit");
};

// 32 consecutive channels [32g, 32g+32) of time row `row` of a [128 x 256] K-major SW128 operand tile.
// tf32: `round_bias` (0x1000 when the tensor core truncates the low 13 mantissa bits, see ap_create) is added to
// the fp32 bit pattern on the way in, so that truncation rounds to nearest, and taken off again on the way out,
// so the residual stream itself stays exact fp32.
// Shared-memory accesses of the epilogues go through 32-bit shared-window addresses (st.shared / ld.shared): the
// generic-pointer form costs 64-bit address arithmetic per access.
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}

template <bool kTf32>
__device__ __forceinline__ void tile_store32(uint32_t tile, int row, int g, const float (&v)[32], uint32_t round_bias) {
  if constexpr (kTf32) {
    const uint32_t sub = tile + g * kABytes;
#pragma unroll
    for (int c = 0; c < 8; ++c)
      sts128(sub + sw128_offset(row, c), __float_as_uint(v[4 * c]) + round_bias, __float_as_uint(v[4 * c + 1]) + round_bias,
             __float_as_uint(v[4 * c + 2]) + round_bias, __float_as_uint(v[4 * c + 3]) + round_bias);
  } else {
    const uint32_t sub = tile + (g >> 1) * kABytes;
    const int q0 = (g & 1) * 4;
#pragma unroll
    for (int c = 0; c < 4; ++c)
      sts128(sub + sw128_offset(row, q0 + c), pack_bf16x2(v[8 * c], v[8 * c + 1]), pack_bf16x2(v[8 * c + 2], v[8 * c + 3]),
             pack_bf16x2(v[8 * c + 4], v[8 * c + 5]), pack_bf16x2(v[8 * c + 6], v[8 * c + 7]));
  }
}
template <bool kTf32>
__device__ __forceinline__ void tile_load32(uint32_t tile, int row, int g, float (&v)[32], uint32_t round_bias) {
  if constexpr (kTf32) {
    const uint32_t sub = tile + g * kABytes;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const uint4 u = lds128(sub + sw128_offset(row, c));
      v[4 * c] = __uint_as_float(u.x - round_bias);
      v[4 * c + 1] = __uint_as_float(u.y - round_bias);
      v[4 * c + 2] = __uint_as_float(u.z - round_bias);
      v[4 * c + 3] = __uint_as_float(u.w - round_bias);
    }
  } else {
    const uint32_t sub = tile + (g >> 1) * kABytes;
    const int q0 = (g & 1) * 4;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const uint4 u = lds128(sub + sw128_offset(row, q0 + c));
      v[8 * c] = bf16_lo(u.x);
      v[8 * c + 1] = bf16_hi(u.x);
      v[8 * c + 2] = bf16_lo(u.y);
      v[8 * c + 3] = bf16_hi(u.y);
      v[8 * c + 4] = bf16_lo(u.z);
      v[8 * c + 5] = bf16_hi(u.z);
      v[8 * c + 6] = bf16_lo(u.w);
      v[8 * c + 7] = bf16_hi(u.w);
    }
  }
}

// Work distribution shared by all warp roles: unit u covers tiles 2u and 2u+1; this CTA takes tile 2u + rank.
// A pair whose second tile does not exist gives that CTA an all-out-of-bounds tile (TMA zero-fills its loads
// and clips its stores) so that both CTAs walk identical barrier sequences.
struct TileCoord {
  int b, l0;
  bool valid;
};
__device__ __forceinline__ TileCoord tile_coord(int tile, int num_tiles, int tiles_per_clip) {
  TileCoord t;
  t.valid = tile < num_tiles;
  if (t.valid) {
    t.b = tile / tiles_per_clip;
    t.l0 = (tile - t.b * tiles_per_clip) * kTileT;
  } else {
    t.b = 0;
    t.l0 = tiles_per_clip * kTileT;
  }
  return t;
}

// A conv tap whose 128-row window lies entirely in the zero padding contributes nothing: with dilation 2048 that
// is a third of GEMM1 for a quarter of a clip's tiles.  The K steps of such a tap are skipped when the window is
// dead for BOTH tiles of the pair (they share the MMA), by the TMA and MMA warps alike.
__device__ __forceinline__ bool tap_window_live(const TileCoord& t, int tap, int dilation, int L) {
  const int lo = t.l0 + (tap - 1) * dilation;
  return t.valid && lo + kTileT > 0 && lo < L;
}
__device__ __forceinline__ uint32_t live_tap_mask(int unit, int num_tiles, int tiles_per_clip, int dilation, int L) {
  const TileCoord t0 = tile_coord(2 * unit, num_tiles, tiles_per_clip);
  const TileCoord t1 = tile_coord(2 * unit + 1, num_tiles, tiles_per_clip);
  uint32_t m = 0;
#pragma unroll
  for (int tap = 0; tap < 3; ++tap)
    if (tap_window_live(t0, tap, dilation, L) || tap_window_live(t1, tap, dilation, L)) m |= 1u << tap;
  return m;
}

// ---------------------------------------------------------------------------------------------------
// K1: one residual layer (WaveNet.py:75-97), fully fused, persistent over pairs of 128-step tiles.
//
//   GEMM1  D1[128 x 512] = sum_{tap,c} h[l + (tap-1)d][c] * W1[o][c][tap]     (K = 768)
//          issued as two N = 256 chunks; chunk c holds gate channels [128c, 128c+128): TMEM columns
//          [0,128) are their tanh rows, [128,256) their sigmoid rows (W1 rows are permuted at pack time).
//   gate   g2 = 2 tanh(D1t + b) * sigmoid(2 (D1s + b))   (sigmoid rows packed pre-halved; TWICE the gate is kept, its 1/2
//          is folded into the res / skip weights: see gate_act)
//          -> bf16 -> shared memory (K-major, SW128) AND, by TMA
//          store from that same shared tile, to gate[layer] in HBM for the tail's skip GEMM.
//   GEMM2  D2[128 x 256] = g2 * (1/2 sqrt(.5) W_res)^T                             (K = 256)
//   out    h_next = sqrt(.5) * h + D2 + (sqrt(.5) b_res + part_{n+1})   (the residual term is the shifted
//          input -- SURVEY section 0 fact 1 -- and the next layer's shift is folded in here).
//
// TMEM (512 columns): two 256-column buffers X, Y.  For tile parity p: chunk0 -> bufA, chunk1 -> bufB,
// D2 -> bufA again (its gate half was drained by then), with (bufA, bufB) = (X, Y) for even tiles and
// (Y, X) for odd tiles, so the MMA warp runs ahead of the epilogue by one chunk at all times.
//
// Epilogue data movement is all TMA and everything lives in the ONE 64 KB operand tile: the epilogue writes the
// gate there (GEMM2's A operand, also TMA-stored to HBM); once GEMM2 has consumed it, warp 2 TMA-loads the
// layer input x over it (residual term), the epilogue turns x into h_next IN PLACE and TMA-stores it.  Every
// epilogue warp (q, hh) owns rows [32q, 32q+32) of the 64-channel sub-tiles {hh, 2+hh} for all three uses
// and stores its own [32 x 64] boxes, so the eight warps never need a CTA-wide barrier.  (Per-thread 16-byte
// global accesses at a 512-byte stride cost 0.91 -> 0.56 ms per launch in an ablation; see profiles/.)
// ---------------------------------------------------------------------------------------------------
// Ablation switches (AP_DEBUG env; profiles/r01_ablation.md, r02_ablation.md) change results and exist only in builds
// made with -DAP_ENABLE_ABLATION (python -m audiopure_b200.build --ablation); the product library compiles them out.
#ifdef AP_ENABLE_ABLATION
#define AP_ABL(args, bit) (((args).debug & (bit)) != 0)
#else
#define AP_ABL(args, bit) false
#endif

struct LayerArgs {
  int L, tiles_per_clip, num_tiles;
  int dilation, layer;
  int write_h;  // 0 for the last layer (its residual output is never consumed)
  int debug;    // ablation switches: 2 no MUFU, 4 no h_next store, 8 no gate store, 16 chunk 1 re-uses stale activation stages (no second A stream)
  uint32_t round_bias;  // tf32 only: see tile_store32
};
struct LayerBias {  // passed by value: lives in the constant bank, read with warp-uniform indices
  float b1[512];    // conv bias, permuted like W1's rows
  float c2[256];    // sqrt(.5)*b_res + part_{n+1}(t)
};

template <bool kTf32>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
layer_kernel(const __grid_constant__ CUtensorMap tm_h, const __grid_constant__ CUtensorMap tm_w1,
             const __grid_constant__ CUtensorMap tm_w2, const __grid_constant__ CUtensorMap tm_gate_st,
             const __grid_constant__ CUtensorMap tm_h_st, const __grid_constant__ LayerBias bias,
             const LayerArgs a) {
  using T = Mode<kTf32>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  uint8_t* stage_base = smem;
  constexpr int kStages = T::kLayerStages;
  constexpr int kSubs = T::kSubs, kSubK = T::kSubK;
  uint8_t* gate_s = smem + kStages * T::kStageBytes;  // the operand tile: gate, then x, then h_next
  uint64_t* bars = reinterpret_cast<uint64_t*>(gate_s + T::kLayerTileBytes);
  uint64_t* full = bars;             // [kStages] TMA -> MMA            (leader's copy is the live one)
  uint64_t* empty = bars + 6;        // [kStages] MMA -> TMA            (per CTA, multicast commit)
  uint64_t* d1_full = bars + 12;     // [2] chunk accumulator ready     MMA -> epilogue (per CTA)
  uint64_t* gate_ready = bars + 14;  // [2] gate half in smem, chunk's TMEM drained   epilogue -> MMA (leader)
  uint64_t* d2_full = bars + 16;     //     residual accumulator ready  MMA -> epilogue (per CTA)
  uint64_t* d2_empty = bars + 17;    //     residual accumulator drained epilogue -> MMA (leader)
  uint64_t* tile_free = bars + 18;   //     gate tile dead (GEMM2 + gate stores done)  epilogue -> x producer
  uint64_t* x_full = bars + 19;      // [kSubs <= 8] x sub-tile k landed in the operand tile    x producer -> epilogue
  uint64_t* g2a_done = bars + 27;    //     split tile: first half of GEMM2 has read the gate of chunk 0   MMA -> epilogue (per CTA)
  uint64_t* half_free = bars + 28;   //     split tile: first half of h_next stored, tile reusable          epilogue -> x producer
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 29);
  constexpr bool kSplitTile = T::kSplitTile;
  constexpr int kTileSubs = T::kTileSubs;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int unit0 = static_cast<int>(blockIdx.x >> 1);
  const int units = static_cast<int>(gridDim.x >> 1);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full[s], 2);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&d1_full[s], 1);
      mbar_init(&gate_ready[s], 2 * T::kEpiWarps);
    }
    for (int s = 0; s < kSubs; ++s) mbar_init(&x_full[s], 1);
    mbar_init(d2_full, 1);
    mbar_init(d2_empty, 2 * T::kEpiWarps);
    mbar_init(tile_free, T::kEpiWarps);
    mbar_init(g2a_done, 1);
    mbar_init(half_free, T::kEpiWarps);
    fence_mbar_init();
    tma_prefetch_desc(&tm_h);
    tma_prefetch_desc(&tm_w1);
    tma_prefetch_desc(&tm_w2);
    tma_prefetch_desc(&tm_gate_st);
    tma_prefetch_desc(&tm_h_st);
  }
  if (warp == 1) {
    tmem_alloc_pair(tmem_ptr, kTmemCols);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_launch_dependents();  // the next launch may set itself up while this one runs ...
  pdl_wait();               // ... and nothing below touches global memory before the previous launch is complete

  if (warp == 0) {
    // ======================= TMA producer (whole warp walks the loop; one elected lane issues) ==========
    const uint32_t full0 = mapa_u32(&full[0], 0);
    const int brow = static_cast<int>(rank * T::kBRows);
    uint32_t it = 0;
    for (int u = unit0; 2 * u < a.num_tiles; u += units) {
      const TileCoord tc = tile_coord(2 * u + rank, a.num_tiles, a.tiles_per_clip);
      const uint32_t live = live_tap_mask(u, a.num_tiles, a.tiles_per_clip, a.dilation, a.L);
      for (int c = 0; c < 2; ++c) {
        for (int ks = 0; ks < 3 * kSubs; ++ks) {
          const int tap = ks / kSubs;
          if (!((live >> tap) & 1)) continue;
          const int s = it % kStages;
          mbar_wait(&empty[s], ((it / kStages) & 1) ^ 1, 1);
          if (elect_one()) {
            const bool skip_a = AP_ABL(a, 16) && c == 1;  // timing bound only: chunk 1 multiplies whatever the stage holds
            mbar_arrive_expect_tx_cluster(full0 + 8 * s, skip_a ? T::kBBytes : T::kStageBytes);
            uint8_t* sa = stage_base + s * T::kStageBytes;
            if (!skip_a)
              tma_load_3d_pair(sa, &tm_h, full0 + 8 * s, (ks % kSubs) * kSubK, tc.l0 + (tap - 1) * a.dilation, tc.b);
            tma_load_2d_pair(sa + kABytes, &tm_w1, full0 + 8 * s, ks * kSubK, a.layer * 512 + c * 256 + brow);
          }
          __syncwarp();
          ++it;
        }
      }
      for (int ks = 0; ks < kSubs; ++ks, ++it) {
        const int s = it % kStages;
        mbar_wait(&empty[s], ((it / kStages) & 1) ^ 1, 2);
        if (elect_one()) {
          mbar_arrive_expect_tx_cluster(full0 + 8 * s, T::kBBytes);
          tma_load_2d_pair(stage_base + s * T::kStageBytes + kABytes, &tm_w2, full0 + 8 * s, ks * kSubK,
                           a.layer * 256 + brow);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ======================= MMA issuer (leader CTA; whole warp walks the loop, one lane issues) ========
    if (rank == 0) {
      const uint64_t desc0 = umma_desc_sw128(smem_u32(stage_base));  // descriptors differ only in the address field
      const uint64_t gdesc0 = umma_desc_sw128(smem_u32(gate_s));
      uint32_t it = 0;
      int i = 0;
      for (int u = unit0; 2 * u < a.num_tiles; u += units, ++i) {
        const uint32_t p = i & 1;
        const uint32_t bufA = tmem_base + (p ? 256u : 0u), bufB = tmem_base + (p ? 0u : 256u);
        const uint32_t live = live_tap_mask(u, a.num_tiles, a.tiles_per_clip, a.dilation, a.L);
        const int last_ks = kSubs * (31 - __clz(live)) + kSubs - 1;  // the centre tap is live for any real tile
        for (int c = 0; c < 2; ++c) {
          if (c == 1 && i > 0) {  // bufB held the previous tile's residual accumulator
            mbar_wait(d2_empty, (i - 1) & 1, 3);
            tc_fence_after();
          }
          const uint32_t d = c ? bufB : bufA;
          uint32_t acc = 0;  // the first MMA issued into the buffer overwrites it
          for (int ks = 0; ks < 3 * kSubs; ++ks) {
            if (!((live >> (ks / kSubs)) & 1)) continue;
            const int s = it % kStages;
            mbar_wait(&full[s], (it / kStages) & 1, 4);
            tc_fence_after();
            if (elect_one()) {
              const uint64_t da = desc0 + static_cast<uint64_t>((s * T::kStageBytes) >> 4);
              const uint64_t db = da + static_cast<uint64_t>(kABytes >> 4);
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_pair<kTf32>(d, da + 2 * k, db + 2 * k, T::kIdesc, acc | k);
              umma_commit_pair(&empty[s]);
              if (ks == last_ks) umma_commit_pair(&d1_full[c]);
            }
            __syncwarp();
            acc = 1;
            ++it;
          }
        }
        for (int ks = 0; ks < kSubs; ++ks, ++it) {
          if (ks == 0 || ks == kSubs / 2) {  // K 0..127 needs gate half 0 (and bufA drained), K 128..255 half 1
            mbar_wait(&gate_ready[ks ? 1 : 0], p, 5);
            tc_fence_after();
          }
          const int s = it % kStages;
          mbar_wait(&full[s], (it / kStages) & 1, 6);
          tc_fence_after();
          if (elect_one()) {
            const int asub = kSplitTile ? (ks & (kTileSubs - 1)) : ks;  // split tile: both gate halves live in sub-tiles 0..3
            const uint64_t da = gdesc0 + static_cast<uint64_t>((asub * kABytes) >> 4);
            const uint64_t db = desc0 + static_cast<uint64_t>((s * T::kStageBytes + kABytes) >> 4);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_pair<kTf32>(bufA, da + 2 * k, db + 2 * k, T::kIdesc, (ks | k) != 0);
            umma_commit_pair(&empty[s]);
            if (kSplitTile && ks == kSubs / 2 - 1) umma_commit_pair(g2a_done);  // chunk 0's gate may be overwritten
            if (ks == kSubs - 1) umma_commit_pair(d2_full);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 2) {
    // ======================= x producer: the layer input again, for the residual term ==================
    // Once the epilogue reports the gate tile dead (GEMM2 and the gate stores have read it), load the tile's
    // own 128 input rows over it, one barrier per sub-tile.
    int i = 0;
    for (int u = unit0; 2 * u < a.num_tiles; u += units, ++i) {
      const TileCoord tc = tile_coord(2 * u + rank, a.num_tiles, a.tiles_per_clip);
      mbar_wait(tile_free, i & 1, 9);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < kTileSubs; ++k) {
          mbar_arrive_expect_tx(&x_full[k], kABytes);
          tma_load_3d(gate_s + k * kABytes, &tm_h, &x_full[k], k * kSubK, tc.l0, tc.b);
        }
      }
      __syncwarp();
      if constexpr (kSplitTile) {  // second half of the channels, once the first half of h_next has left the tile
        mbar_wait(half_free, i & 1, 24);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < kTileSubs; ++k) {
            mbar_arrive_expect_tx(&x_full[k], kABytes);
            tma_load_3d(gate_s + k * kABytes, &tm_h, &x_full[k], (kTileSubs + k) * kSubK, tc.l0, tc.b);
          }
        }
        __syncwarp();
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ======================= epilogue (8 warps, each owns rows [32q,+32) of 64-channel chunks {hh, 2+hh}) ====
    // bf16: the body is instantiated once per warpgroup (hh = 0 / 1) with the chunk and half loops unrolled, so that
    // every bias index is a compile-time constant and the adds read the constant bank through warp-uniform loads
    // instead of per-thread LDC (-2.9 % per launch, profiles/r02_ablation.md).  tf32 keeps ONE rolled copy.
    const int q = warp & 3;   // TMEM lane quarter this warp may read
    auto epilogue = [&](auto hh_c) {
    const int hh = hh_c;  // warpgroup: which 64-channel chunks / which half of a gate chunk (a constant in the bf16 build)
    const int row = q * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t gate_ready0 = mapa_u32(&gate_ready[0], 0);
    const uint32_t d2_empty_l = mapa_u32(d2_empty, 0);
    const uint32_t gate_sa = smem_u32(gate_s);
    int i = 0;
    for (int u = unit0; 2 * u < a.num_tiles; u += units, ++i) {
      const uint32_t p = i & 1;
      const TileCoord tc = tile_coord(2 * u + rank, a.num_tiles, a.tiles_per_clip);
      const int b = tc.b, l0 = tc.l0;
      const uint32_t bufA = tmem_base + (p ? 256u : 0u), bufB = tmem_base + (p ? 0u : 256u);

      // this warp's regions are rewritten below: its previous h_next stores must have finished reading them
      if (lane == 0) tma_store_wait_read();
      __syncwarp();

      // ---- gate: two chunks of 128 gate channels; this warp does channels [64hh, 64hh+64) of each ----
      // (bf16: unrolled, so the bias indices are constants; tf32: its accurate tanhf / exp2f gate is ~10x the code,
      // four unrolled copies of it overflow the instruction cache -- measured 0.86 -> 1.10 ms per launch)
#pragma unroll(kTf32 ? 1 : 2)
      for (int c = 0; c < 2; ++c) {
        mbar_wait(&d1_full[c], p, 7);
        tc_fence_after();
        if (kSplitTile && c == 1) {
          // chunk 1's gate goes where chunk 0's is: the first half of GEMM2 and this warp's gate stores must have read it
          mbar_wait(g2a_done, p, 23);
          if (lane == 0) tma_store_wait_read();
          __syncwarp();
        }
        const uint32_t buf = (c ? bufB : bufA) + lane_addr;
#pragma unroll
        for (int itn = 0; itn < 2; ++itn) {
          const int j0 = hh * 64 + itn * 32;  // gate channel within the chunk
          uint32_t rt[32], rs[32];
          tmem_ld32(buf + j0, rt);
          tmem_ld32(buf + 128 + j0, rs);
          tmem_ld_wait();
          const float* bt = bias.b1 + c * 256 + j0;
          const float* bs = bt + 128;
          float o[32];
          if (AP_ABL(a, 2)) {
#pragma unroll
            for (int j = 0; j < 32; ++j) o[j] = __uint_as_float(rt[j]) + bt[j] + __uint_as_float(rs[j]) + bs[j];
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              o[j] = gate_act<kTf32>(__uint_as_float(rt[j]) + bt[j], __uint_as_float(rs[j]) + bs[j]);
          }
          const int gg = 4 * c + 2 * hh + itn;  // 32-channel group of the gate
          tile_store32<kTf32>(gate_sa, row, kSplitTile ? (gg & (kTileSubs - 1)) : gg, o, a.round_bias);
        }
        tc_fence_before();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive_cluster(gate_ready0 + 8 * c);
          // this warp's [32 x 64-channel] part of the gate tile -> HBM (operand of the tail's skip GEMM)
          if (!AP_ABL(a, 8)) {
#pragma unroll
            for (int s = 0; s < T::kSubsPer64; ++s) {
              const int sub = (2 * c + hh) * T::kSubsPer64 + s;
              const int ssub = kSplitTile ? (sub & (kTileSubs - 1)) : sub;
              tma_store_4d(&tm_gate_st, gate_s + ssub * kABytes + q * 32 * 128, sub * kSubK, l0 + q * 32, b, a.layer);
            }
            tma_store_commit();
          }
        }
      }

      // ---- residual output: h_next = sqrt(.5) x + D2 + c2, computed in place over x in the operand tile ----
      mbar_wait(d2_full, p, 8);  // also: the MMAs have finished reading the gate tile
      tc_fence_after();
      if (lane == 0) {
        tma_store_wait_read();   // ... and so have this warp's gate stores: the tile may be overwritten with x
        mbar_arrive(tile_free);
      }
      __syncwarp();
#pragma unroll(kTf32 ? 1 : 2)
      for (int kk = 0; kk < 2; ++kk) {
        const int k = hh + 2 * kk;  // 64-channel chunk
        uint32_t r0[32], r1[32];
        tmem_ld32(bufA + lane_addr + k * 64, r0);
        tmem_ld32(bufA + lane_addr + k * 64 + 32, r1);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int g = 2 * k + half;  // 32-channel group
          // split tile: channels 0-127 (kk = 0) and 128-255 (kk = 1) pass through sub-tiles 0..3 one after the other;
          // x_full[s] then completes twice per tile, phases 2i and 2i + 1
          const int gs = kSplitTile ? (g & (kTileSubs - 1)) : g;
          if (half == 0 || kTf32) mbar_wait(&x_full[kSplitTile ? gs : g * 32 / kSubK], kSplitTile ? (kk & 1) : p, 10);
          if (half == 0) tmem_ld_wait();
          const uint32_t* r = half ? r1 : r0;
          const float* cc = bias.c2 + g * 32;
          float v[32];
          tile_load32<kTf32>(gate_sa, row, gs, v, a.round_bias);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaf(v[j], kSqrtHalf, __uint_as_float(r[j]) + cc[j]);
          tile_store32<kTf32>(gate_sa, row, gs, v, a.round_bias);
        }
        if (kk == 1) {  // all of this warp's accumulator columns are in registers / consumed
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(d2_empty_l);
        }
        if constexpr (kSplitTile) {  // this half of h_next leaves now; after kk = 0 the tile is handed back for x's second half
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            if (a.write_h && !AP_ABL(a, 4)) {
#pragma unroll
              for (int s = 0; s < T::kSubsPer64; ++s) {
                const int sub = k * T::kSubsPer64 + s;
                tma_store_3d(&tm_h_st, gate_s + (sub & (kTileSubs - 1)) * kABytes + q * 32 * 128, sub * kSubK, l0 + q * 32, b);
              }
              tma_store_commit();
            }
            if (kk == 0) {
              tma_store_wait_read();
              mbar_arrive(half_free);
            }
          }
          __syncwarp();
        }
      }
      if constexpr (!kSplitTile) {
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0 && a.write_h && !AP_ABL(a, 4)) {
#pragma unroll
          for (int kk = 0; kk < 2; ++kk)
#pragma unroll
            for (int s = 0; s < T::kSubsPer64; ++s) {
              const int sub = (hh + 2 * kk) * T::kSubsPer64 + s;
              tma_store_3d(&tm_h_st, gate_s + sub * kABytes + q * 32 * 128, sub * kSubK, l0 + q * 32, b);
            }
          tma_store_commit();
        }
      }
    }
    };
    if constexpr (kTf32)
      epilogue((warp - kEpiWarp0) >> 2);  // one copy: the accurate gate is large, see above
    else if (((warp - kEpiWarp0) >> 2) == 0)
      epilogue(IntC<0>{});
    else
      epilogue(IntC<1>{});
    if (lane == 0) tma_store_wait_all();
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------------
// K2: skip GEMM + output head + reverse-step update, persistent over pairs of 128-step tiles.
//
//   GEMMs  S[128 x 256]  = sum_n (2 gate_n) * (1/2 sqrt(1/N) W_skip,n)^T   (K = N*256; WaveNet.py:95,133,135)
//   head   y  = relu(bf16(S + bias) * W_f^T + b_f)                     (GEMM, K = 256; WaveNet.py:160-161)
//          eps = y . w_o + b_o                                          (256 -> 1 dot; WaveNet.py:162)
//   update x_out = ca * x_in + cb * eps + cc * z                        (diffwave_ddpm.py:159,99-102 /
//          diffwave_sde.py Euler-Maruyama / one-shot, as coefficients), z injected or Philox.
//
// TMEM: two 256-column buffers; tile i accumulates S in buffer i&1, the head GEMM of tile i overwrites the
// same buffer once the epilogue has turned S into the bf16 smem operand, and is issued in the middle of
// tile i+1's K loop so the tensor pipe never waits for the epilogue.
// ---------------------------------------------------------------------------------------------------
struct TailArgs {
  float bo;
  const float* x_in;    // [B][L]
  const float* z;       // [B][L] injected noise or nullptr
  float* x_out;         // [B][L] or nullptr
  float* eps_out;       // [B][L] or nullptr
  float ca, cb, cc;
  unsigned long long seed;   // Philox key (used when z == nullptr and cc != 0)
  uint32_t stream_lo;        // Philox stream: purpose/step id
  long long elem_offset;     // global element index of x_in[0] (keeps noise independent of sharding)
  int B, L, tiles_per_clip, num_tiles, num_layers;
  uint32_t round_bias;       // tf32 only: see tile_store32
};
struct TailBias {  // passed by value: constant bank, warp-uniform indices
  float bs[256];   // sqrt(1/N) * sum_n b_skip,n
  float bf[256];   // final_conv[0] bias
  float wo[256];   // final_conv[2] (256 -> 1) weight
};

template <bool kTf32>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
tail_kernel(const __grid_constant__ CUtensorMap tm_gate, const __grid_constant__ CUtensorMap tm_ws,
            const __grid_constant__ CUtensorMap tm_wf, const __grid_constant__ TailBias bias, const TailArgs a) {
  using T = Mode<kTf32>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  uint8_t* stage_base = smem;
  constexpr int kStages = T::kTailStages;
  constexpr int kSubs = T::kSubs, kSubK = T::kSubK;
  uint8_t* s_tile = smem + kStages * T::kStageBytes;
  float* partial = reinterpret_cast<float*>(s_tile + T::kTileBytes);  // [2 parities][128 rows]: upper-half dot products
  uint64_t* bars = reinterpret_cast<uint64_t*>(partial + 256);
  uint64_t* full = bars;           // [kStages <= 6]                   (leader)
  uint64_t* empty = bars + 6;      // [kStages <= 6]                   (per CTA)
  uint64_t* d_full = bars + 12;    // [2] skip accumulator ready       MMA -> epilogue (per CTA)
  uint64_t* d_empty = bars + 14;   // [2] buffer fully consumed        epilogue -> MMA (leader)
  uint64_t* s_ready = bars + 16;   //     operand-precision skip tile in smem   epilogue -> MMA (leader)
  uint64_t* d3_full = bars + 17;   // [2] head accumulator ready       MMA -> epilogue (per CTA)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 19);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int unit0 = static_cast<int>(blockIdx.x >> 1);
  const int units = static_cast<int>(gridDim.x >> 1);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full[s], 2);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&d_full[s], 1);
      mbar_init(&d_empty[s], 2 * T::kEpiWarps);
      mbar_init(&d3_full[s], 1);
    }
    mbar_init(s_ready, 2 * T::kEpiWarps);
    fence_mbar_init();
    tma_prefetch_desc(&tm_gate);
    tma_prefetch_desc(&tm_ws);
    tma_prefetch_desc(&tm_wf);
  }
  if (warp == 1) {
    tmem_alloc_pair(tmem_ptr, kTmemCols);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_launch_dependents();
  pdl_wait();

  const int total_ks = a.num_layers * kSubs;
  const int J = total_ks / 2 < 4 * kSubs ? total_ks / 2 : 4 * kSubs;  // where the previous tile's head GEMM is slotted in

  if (warp == 0) {
    const uint32_t full0 = mapa_u32(&full[0], 0);
    const int brow = static_cast<int>(rank * T::kBRows);
    uint32_t it = 0;
    int i = 0;
    auto load_wf = [&]() {
      for (int ks = 0; ks < kSubs; ++ks, ++it) {
        const int s = it % kStages;
        mbar_wait(&empty[s], ((it / kStages) & 1) ^ 1, 11);
        if (elect_one()) {
          mbar_arrive_expect_tx_cluster(full0 + 8 * s, T::kBBytes);
          tma_load_2d_pair(stage_base + s * T::kStageBytes + kABytes, &tm_wf, full0 + 8 * s, ks * kSubK, brow);
        }
        __syncwarp();
      }
    };
    for (int u = unit0; 2 * u < a.num_tiles; u += units, ++i) {
      const TileCoord tc = tile_coord(2 * u + rank, a.num_tiles, a.tiles_per_clip);
      for (int ks = 0; ks < total_ks; ++ks, ++it) {
        if (ks == J && i > 0) load_wf();
        const int s = it % kStages;
        mbar_wait(&empty[s], ((it / kStages) & 1) ^ 1, 12);
        if (elect_one()) {
          mbar_arrive_expect_tx_cluster(full0 + 8 * s, T::kStageBytes);
          uint8_t* sa = stage_base + s * T::kStageBytes;
          tma_load_4d_pair(sa, &tm_gate, full0 + 8 * s, (ks % kSubs) * kSubK, tc.l0, tc.b, ks / kSubs);
          tma_load_2d_pair(sa + kABytes, &tm_ws, full0 + 8 * s, ks * kSubK, brow);
        }
        __syncwarp();
      }
    }
    if (i > 0) load_wf();
  } else if (warp == 1) {
    if (rank == 0) {
      const uint64_t desc0 = umma_desc_sw128(smem_u32(stage_base));
      const uint64_t sdesc0 = umma_desc_sw128(smem_u32(s_tile));
      uint32_t it = 0;
      int i = 0;
      auto head_gemm = [&](int ip) {
        const uint32_t d = tmem_base + ((ip & 1) ? 256u : 0u);
        mbar_wait(s_ready, ip & 1, 13);
        tc_fence_after();
        for (int ks = 0; ks < kSubs; ++ks, ++it) {
          const int s = it % kStages;
          mbar_wait(&full[s], (it / kStages) & 1, 14);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t da = sdesc0 + static_cast<uint64_t>((ks * kABytes) >> 4);
            const uint64_t db = desc0 + static_cast<uint64_t>((s * T::kStageBytes + kABytes) >> 4);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_pair<kTf32>(d, da + 2 * k, db + 2 * k, T::kIdesc, (ks | k) != 0);
            umma_commit_pair(&empty[s]);
            if (ks == kSubs - 1) umma_commit_pair(&d3_full[ip & 1]);
          }
          __syncwarp();
        }
      };
      for (int u = unit0; 2 * u < a.num_tiles; u += units, ++i) {
        const uint32_t p = i & 1, uu = i >> 1;
        const uint32_t d = tmem_base + (p ? 256u : 0u);
        if (uu >= 1) {
          mbar_wait(&d_empty[p], (uu - 1) & 1, 15);
          tc_fence_after();
        }
        for (int ks = 0; ks < total_ks; ++ks, ++it) {
          if (ks == J && i > 0) head_gemm(i - 1);
          const int s = it % kStages;
          mbar_wait(&full[s], (it / kStages) & 1, 16);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t da = desc0 + static_cast<uint64_t>((s * T::kStageBytes) >> 4);
            const uint64_t db = da + static_cast<uint64_t>(kABytes >> 4);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_pair<kTf32>(d, da + 2 * k, db + 2 * k, T::kIdesc, (ks | k) != 0);
            umma_commit_pair(&empty[s]);
            if (ks == total_ks - 1) umma_commit_pair(&d_full[p]);
          }
          __syncwarp();
        }
      }
      if (i > 0) head_gemm(i - 1);
    }
  } else if (warp >= kEpiWarp0) {
    const int e = warp - kEpiWarp0;
    const int q = warp & 3;
    const int hh = e >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t s_ready_l = mapa_u32(s_ready, 0);
    const uint32_t d_empty0 = mapa_u32(&d_empty[0], 0);
    int i = 0;
    for (int u = unit0; 2 * u < a.num_tiles; u += units, ++i) {
      const uint32_t p = i & 1, uu = i >> 1;
      const TileCoord tc = tile_coord(2 * u + rank, a.num_tiles, a.tiles_per_clip);
      const int b = tc.b, l0 = tc.l0;
      const uint32_t buf = tmem_base + (p ? 256u : 0u) + lane_addr;
      const bool row_ok = (l0 + row) < a.L;

      // ---- skip sum -> operand tile (bf16 / tf32) ----
      mbar_wait(&d_full[p], uu & 1, 17);
      tc_fence_after();
#pragma unroll 1
      for (int itn = 0; itn < 4; ++itn) {
        const int j0 = hh * 128 + itn * 32;
        uint32_t r[32];
        tmem_ld32(buf + j0, r);
        tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) + bias.bs[j0 + j];
        tile_store32<kTf32>(smem_u32(s_tile), row, j0 >> 5, v, a.round_bias);
      }
      tc_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(s_ready_l);

      // ---- head: relu, 256 -> 1 dot ----
      mbar_wait(&d3_full[p], uu & 1, 18);
      tc_fence_after();
      float acc = 0.f;
#pragma unroll 1
      for (int itn = 0; itn < 4; ++itn) {
        const int j0 = hh * 128 + itn * 32;
        uint32_t r[32];
        tmem_ld32(buf + j0, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j)
          acc = fmaf(fmaxf(__uint_as_float(r[j]) + bias.bf[j0 + j], 0.f), bias.wo[j0 + j], acc);
      }
      if (hh == 1) partial[p * 128 + row] = acc;
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(d_empty0 + 8 * p);
      named_bar_sync(1, kEpiThreads);
      if (hh == 0 && row_ok) {
        const float eps = acc + partial[p * 128 + row] + a.bo;
        const size_t idx = static_cast<size_t>(b) * a.L + l0 + row;
        if (a.eps_out) a.eps_out[idx] = eps;
        if (a.x_out) {
          float v = fmaf(a.ca, a.x_in[idx], a.cb * eps);
          if (a.cc != 0.f) {
            const float zz = a.z ? a.z[idx]
                                 : philox_normal(a.seed, a.stream_lo, 0u,
                                                 static_cast<uint64_t>(a.elem_offset) + idx);
            v = fmaf(a.cc, zz, v);
          }
          a.x_out[idx] = v;
        }
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------------
// Bring-up / regression kernel (also ap_create's probe of how the tensor core narrows fp32 operands to tf32):
// D[128 x 256] = A[128 x K] * B[256 x K]^T through exactly the TMA box,
// swizzle, descriptor and TMEM-load conventions the two kernels above rely on.  One CTA, 128 threads.
// ---------------------------------------------------------------------------------------------------
template <bool kTf32>
__global__ void __launch_bounds__(128, 1)
debug_gemm_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, float* d,
                  int K) {
  constexpr int kSubK = Mode<kTf32>::kSubK;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  constexpr uint32_t kStageBytes = kABytes + 256 * 128;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStageBytes);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_ptr, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = kTf32 ? umma_idesc_tf32(128, 256) : umma_idesc_bf16(128, 256);
    for (int ks = 0; ks < K / kSubK; ++ks) {
      mbar_arrive_expect_tx(&bars[0], kStageBytes);
      tma_load_2d(smem, &tm_a, &bars[0], ks * kSubK, 0);
      tma_load_2d(smem + kABytes, &tm_b, &bars[0], ks * kSubK, 0);
      mbar_wait(&bars[0], ks & 1, 21);
      tc_fence_after();
      const uint64_t da = umma_desc_sw128(smem_u32(smem)), db = umma_desc_sw128(smem_u32(smem + kABytes));
      for (int k = 0; k < 4; ++k) {
        if constexpr (kTf32)
          umma_tf32(tmem_base, umma_desc_advance_k(da, k), umma_desc_advance_k(db, k), idesc, (ks | k) != 0);
        else
          umma_bf16(tmem_base, umma_desc_advance_k(da, k), umma_desc_advance_k(db, k), idesc, (ks | k) != 0);
      }
      umma_commit(&bars[1]);
      mbar_wait(&bars[1], ks & 1, 22);
    }
  }
  __syncthreads();
  tc_fence_after();
  const int row = warp * 32 + lane;
  for (int j0 = 0; j0 < 256; j0 += 32) {
    uint32_t r[32];
    tmem_ld32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + j0, r);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) d[row * 256 + j0 + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// ---------------------------------------------------------------------------------------------------
// Small HBM-bound elementwise kernels.
// ---------------------------------------------------------------------------------------------------
// y = a*x + b*z ; z injected or Philox(seed, stream, elem_offset + i).  diffwave_ddpm.py:66-67, diffwave_sde.py:185-191.
__global__ void __launch_bounds__(256) axpbz_kernel(const float* __restrict__ x, const float* __restrict__ z,
                                                    float* __restrict__ y, float ca, float cb, long long n,
                                                    unsigned long long seed, uint32_t stream_lo,
                                                    long long elem_offset) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float zz = z ? z[i] : philox_normal(seed, stream_lo, 0u, static_cast<uint64_t>(elem_offset + i));
    y[i] = fmaf(ca, x[i], cb * zz);
  }
}

// Smoothing draws (certified_robust.py:46-54) for a batch of rows that may span several clips.  The work list of a
// certify call is the clip-major flattening of (clip, draw): row r of this launch is item flat = flat0 + r,
// clip = flat / per_clip, draw = first_draw + flat % per_clip, and
//   out[r][l] = scale * (x[clip][l] + sigma * z),   z = zinj[flat][l] (injected, indexed from the start of the call's
//   [clips][per_clip][L] tensor) or Philox keyed on (seed, clip_key0 + clip, draw, l).
// Because the key is (clip, draw), how the list is cut into batches or sharded over GPUs never changes a draw.
// HBM-bound: 4 B written per element (x stays in L2); float4 path when L is a multiple of 4.
struct SmoothArgs {
  const float* x;      // [clips][L]
  const float* zinj;   // or nullptr
  float* out;          // [n_rows][L]
  int L, n_rows;
  long long flat0, per_clip, first_draw;
  float sigma, scale;
  unsigned long long seed;
  uint32_t clip_key0;
};
constexpr uint32_t kSmoothStream = 0x534D4F4Fu;  // 'SMOO'

__global__ void __launch_bounds__(256) smooth_inputs_kernel(const SmoothArgs a) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long tid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if ((a.L & 3) == 0) {
    const int L4 = a.L >> 2;
    const long long n4 = static_cast<long long>(a.n_rows) * L4;
    for (long long i = tid; i < n4; i += stride) {
      const long long r = i / L4;
      const int l = static_cast<int>(i - r * L4) << 2;
      const long long flat = a.flat0 + r;
      const long long clip = flat / a.per_clip;
      const long long draw = a.first_draw + (flat - clip * a.per_clip);
      float zz[4];
      if (a.zinj) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(a.zinj + flat * a.L + l));
        zz[0] = v.x; zz[1] = v.y; zz[2] = v.z; zz[3] = v.w;
      } else {
        philox_normal4(a.seed, kSmoothStream, a.clip_key0 + static_cast<uint32_t>(clip),
                       (static_cast<uint64_t>(draw) * static_cast<uint64_t>(a.L) + l) >> 2, zz);
      }
      const float4 xv = __ldg(reinterpret_cast<const float4*>(a.x + clip * a.L + l));
      float4 o;
      o.x = a.scale * fmaf(a.sigma, zz[0], xv.x);
      o.y = a.scale * fmaf(a.sigma, zz[1], xv.y);
      o.z = a.scale * fmaf(a.sigma, zz[2], xv.z);
      o.w = a.scale * fmaf(a.sigma, zz[3], xv.w);
      reinterpret_cast<float4*>(a.out)[i] = o;
    }
  } else {
    const long long n = static_cast<long long>(a.n_rows) * a.L;
    for (long long i = tid; i < n; i += stride) {
      const long long r = i / a.L;
      const int l = static_cast<int>(i - r * a.L);
      const long long flat = a.flat0 + r;
      const long long clip = flat / a.per_clip;
      const long long draw = a.first_draw + (flat - clip * a.per_clip);
      const float zz = a.zinj ? a.zinj[flat * a.L + l]
                              : philox_normal(a.seed, kSmoothStream, a.clip_key0 + static_cast<uint32_t>(clip),
                                              static_cast<uint64_t>(draw) * static_cast<uint64_t>(a.L) + l);
      a.out[i] = a.scale * fmaf(a.sigma, zz, a.x[clip * a.L + l]);
    }
  }
}

// NES black-box gradient estimation (robustness_eval/_NES.py:15-55), antithetic sampling.  For audio a and sample j
// of a draw batch of S samples (S even): noise_j = +z_j for j < S/2, -z_{j-S/2} otherwise (_NES.py:19-21), and
//   eval[a][j][l] = x[a][l] + sigma * noise_j[l]                                            (_NES.py:24)
// with an optional leading clean row per audio (`lead` = 1 for the first draw batch, _NES.py:22-23).
// z_j: injected zinj[a][j][l] ([audios][S/2][L]) or Philox keyed on (seed, 'NESG', audio_key0 + a, (draw0 + j) * L + l).
struct NesArgs {
  const float* x;     // [audios][L]
  const float* zinj;  // [audios][S/2][L] or nullptr
  float* out;         // [audios][lead + S][L]
  const float* loss;  // [audios][loss_stride]: losses of the S perturbed samples start at loss_off   (grad kernel)
  float* grad;        // [audios][L], accumulated into                                                (grad kernel)
  int L, audios, S, lead, loss_stride, loss_off;
  float sigma, grad_scale;
  unsigned long long seed;
  uint32_t audio_key0;
  long long draw0;
};
constexpr uint32_t kNesStream = 0x4E455347u;  // 'NESG'

__device__ __forceinline__ float nes_z(const NesArgs& a, int audio, int j, int l) {
  return a.zinj ? a.zinj[(static_cast<size_t>(audio) * (a.S / 2) + j) * a.L + l]
                : philox_normal(a.seed, kNesStream, a.audio_key0 + static_cast<uint32_t>(audio),
                                static_cast<uint64_t>(a.draw0 + j) * static_cast<uint64_t>(a.L) + l);
}

__global__ void __launch_bounds__(256) nes_inputs_kernel(const NesArgs a) {
  const int half = a.S / 2;
  const long long n = static_cast<long long>(a.audios) * half * a.L;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int l = static_cast<int>(i % a.L);
    const long long aj = i / a.L;
    const int j = static_cast<int>(aj % half), audio = static_cast<int>(aj / half);
    const float xv = a.x[static_cast<size_t>(audio) * a.L + l];
    const float d = a.sigma * nes_z(a, audio, j, l);
    float* base = a.out + static_cast<size_t>(audio) * (a.lead + a.S) * a.L;
    base[static_cast<size_t>(a.lead + j) * a.L + l] = d + xv;          // noise * sigma + x (_NES.py:24)
    base[static_cast<size_t>(a.lead + half + j) * a.L + l] = -d + xv;
    if (a.lead && j == 0) base[l] = xv;
  }
}

// grad[a][l] += grad_scale * sum_j loss[a][j] * noise_j[l] = grad_scale * sum_{j < S/2} (loss_j - loss_{j+S/2}) z_j[l]
// (_NES.py:44-48,52: mean over the S samples, / sigma / num_batches folded into grad_scale).  The noise is
// re-generated from its Philox key instead of being kept in HBM (S * L * 4 bytes per audio).
__global__ void __launch_bounds__(256) nes_grad_kernel(const NesArgs a) {
  const int half = a.S / 2;
  const long long n = static_cast<long long>(a.audios) * a.L;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int l = static_cast<int>(i % a.L), audio = static_cast<int>(i / a.L);
    const float* ls = a.loss + static_cast<size_t>(audio) * a.loss_stride + a.loss_off;
    float acc = 0.f;
    for (int j = 0; j < half; ++j) acc = fmaf(ls[j] - ls[half + j], nes_z(a, audio, j, l), acc);
    a.grad[i] += a.grad_scale * acc;
  }
}

// Consumer-side epilogue (resnext.py:56-64 after batch-norm folding): y = relu?(y + bias[c] (+ res)) in
// place over a channels-last bf16 activation [rows][C], 8 channels (16 bytes) per thread.  HBM/L2-bound: one
// read-modify-write pass instead of torch's separate bias-add, residual-add and clamp passes.
__global__ void __launch_bounds__(256) bias_act_kernel(uint4* __restrict__ y, const float* __restrict__ bias,
                                                       const uint4* __restrict__ res, long long n_vec, int c_vec,
                                                       int relu) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n_vec; i += stride) {
    const int c0 = static_cast<int>(i % c_vec) * 8;
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + c0));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + c0 + 4));
    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    uint4 v = y[i];
    uint32_t* vw = reinterpret_cast<uint32_t*>(&v);
    uint4 r = make_uint4(0u, 0u, 0u, 0u);
    if (res) r = res[i];
    const uint32_t* rw = reinterpret_cast<const uint32_t*>(&r);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float lo = bf16_lo(vw[j]) + bb[2 * j] + bf16_lo(rw[j]);
      float hi = bf16_hi(vw[j]) + bb[2 * j + 1] + bf16_hi(rw[j]);
      if (relu) {
        lo = fmaxf(lo, 0.f);
        hi = fmaxf(hi, 0.f);
      }
      vw[j] = pack_bf16x2(lo, hi);
    }
    y[i] = v;
  }
}

// certified_robust.py:58-67 for the flattened (clip, draw) work list of smooth_inputs_kernel:
// counts[sel][clip][c] += #rows whose argmax is c (lowest index wins ties, like torch.max), where row r is item
// flat0 + r, clip = flat / per_clip and sel = 1 for draws >= n_split (the estimation pass of certify,
// certified_robust.py:89-93) and 0 otherwise (the selection pass, :84-87).  Integer counts: order-independent.
__global__ void __launch_bounds__(256) vote_counts_kernel(const float* __restrict__ logits, int rows, int K,
                                                          long long flat0, long long per_clip, long long n_split,
                                                          int n_clips, unsigned long long* __restrict__ counts) {
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += gridDim.x * blockDim.x) {
    const float* p = logits + static_cast<size_t>(r) * K;
    int best = 0;
    float bv = p[0];
    for (int c = 1; c < K; ++c) {
      const float v = p[c];
      if (v > bv) {
        bv = v;
        best = c;
      }
    }
    const long long flat = flat0 + r;
    const long long clip = flat / per_clip;
    const long long sel = (flat - clip * per_clip) >= n_split ? 1 : 0;
    atomicAdd(counts + (sel * n_clips + clip) * K + best, 1ull);
  }
}

}  // namespace ap
