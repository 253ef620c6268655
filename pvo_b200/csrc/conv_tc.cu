// Implicit-GEMM convolution on 5th-generation tensor cores: TMA -> shared memory -> tcgen05.mma -> TMEM -> epilogue.
//
// The dense convolutions of the update operator in half precision (what the reference's inference autocasts to,
// VO_Module/droid_slam/factor_graph.py:227,242): ConvGRU's three 448->128 3x3 convs (modules/gru.py:10-12, 59% of
// the operator's FLOPs), the encoders' 3x3 layers and the heads' first layers (droid_net.py:172-214).  The reference
// runs each as a cuDNN call followed by separate bias / activation / gate kernels; here bias, ReLU and the GRU gate
// arithmetic are the epilogue, and torch.cat of the GRU input is virtual (one TMA tensor map per input segment).
//
// GEMM view  D[pixel, cout] = sum_{tap, c} X[pixel + tap, c] * W[cout, tap, c]
//   M tile  = 128 pixels = an 8 x 16 patch of one image (TMA box {64 ch, 16, 8, 1}; out-of-image taps and rows are
//             zero-filled by TMA, which is the convolution's zero padding)
//   N tile  = BN output channels (64 / 128 / 256), K block = 64 channels of one tap (128-byte swizzled rows)
//   UMMA 128 x BN x 16, fp16 operands, fp32 accumulators in TMEM, double-buffered (2 x BN columns)
// Warp roles (320 threads, persistent over tiles): warp 0 = TMA producer, warp 1 = MMA issuer + TMEM owner,
// warps 2-9 = epilogue (two per TMEM lane quarter, each draining half of the accumulator columns).  A pipeline stage holds
// 1-2 k-blocks of {A 16 KB, B BN*128 B}; BN <= 128 runs two CTAs per SM.  BN = 256 bias tiles transpose their output
// through shared memory so that stores are row-contiguous.
// Roofline: tensor pipe.  Measured bounds, in the order they bite: per-k-block issue overhead on
// short-K layers, the row rate of the 4-D A-operand TMA boxes (~0.35 us per k-block per SM, so only N = 256 tiles approach
// the MMA peak), row-strided epilogue stores on 1x1 channel expansions.
#include <cstdlib>

#include "tc_common.cuh"

namespace {

using namespace tc;

constexpr int kBM = 128, kBK = 64, kTileW = 16, kTileH = 8;
constexpr int kThreads = 320;      // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue (two per TMEM lane quarter)
constexpr int kMaxSeg = 4;

struct TcParams {
  CUtensorMap tmA[kMaxSeg];
  CUtensorMap tmB;
  int seg_blocks[kMaxSeg];   // 64-channel blocks per segment
  int nseg;
  int blocks_per_tap;        // sum of seg_blocks
  int taps, KW, padH, padW;
  int H, W, nimg;
  int tiles_x, tiles_y, n_tiles, total_tiles;
  int Cout;
  int halo;                  // 3x3 halo-reuse kernel: 1 = 16 x 8 pixel tiles, 2 = two stacked (32 x 8); tmA boxes are the haloed patches
};

// ---- epilogue helpers.  The epilogue warps are on the critical path of every short-K layer (measured: 19 instructions
// per output element made the 1x1 gate conv 10x slower than its operand stream), so everything here is vectorised:
// 16-byte bias / state loads, packed half conversions, MUFU tanh for the gates.
__device__ __forceinline__ float fast_tanh(float x) { float t; asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(x)); return t; }
__device__ __forceinline__ float fast_sigmoid(float x) { return fmaf(fast_tanh(0.5f * x), 0.5f, 0.5f); }   // = 1/(1+e^-x)

__device__ __forceinline__ void ld16f(const float* p, float* o) {          // 16 consecutive floats, 16-byte aligned
#pragma unroll
  for (int q = 0; q < 4; q++) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p) + q);
    o[4 * q] = t.x; o[4 * q + 1] = t.y; o[4 * q + 2] = t.z; o[4 * q + 3] = t.w;
  }
}
template <bool LDG>
__device__ __forceinline__ void ld16h(const __half* p, float* o) {         // 16 consecutive halves -> floats
#pragma unroll
  for (int q = 0; q < 2; q++) {
    const uint4 t = LDG ? __ldg(reinterpret_cast<const uint4*>(p) + q) : *(reinterpret_cast<const uint4*>(p) + q);
    const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[k]));
      o[8 * q + 2 * k] = f.x; o[8 * q + 2 * k + 1] = f.y;
    }
  }
}
__device__ __forceinline__ void st16h(__half* p, const float* x) {         // 16 floats -> 16 halves, two 16-byte stores
#pragma unroll
  for (int q = 0; q < 2; q++) {
    uint32_t w[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const __half2 h = __floats2half2_rn(x[8 * q + 2 * k], x[8 * q + 2 * k + 1]);
      w[k] = *reinterpret_cast<const uint32_t*>(&h);
    }
    *(reinterpret_cast<uint4*>(p) + q) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

// one step of the cross-lane transpose-reduce: lanes with `BIT` set keep columns [W, 2W), the others [0, W)
template <int W, int BIT>
__device__ __forceinline__ void xr_step(float* x, int lane) {
  const bool up = lane & BIT;
#pragma unroll
  for (int j = 0; j < W; j++) {
    const float send = up ? x[j] : x[j + W];
    const float recv = __shfl_xor_sync(0xffffffffu, send, BIT);
    x[j] = (up ? x[j + W] : x[j]) + recv;
  }
}

// ---- epilogues: store16(m, e, n, v, ok, tile, q): 16 consecutive output channels n.. of pixel m (image/edge e) ------
struct EpBias {          // out = act(acc + bias); with a residual: relu(res + act(acc + bias))  (ResNet bottleneck tail)
  static constexpr bool kStaged = true;   // BN = 256 tiles transpose the output through shared memory (see the epilogue loop)
  const float* bias; __half* out; int ostride; int ooff; int Cout; int relu; const __half* res; int rstride;
  // 16 columns of one pixel -> 8 packed half2 words (bias, activation, residual applied); full 16-column chunks only
  __device__ __forceinline__ void apply16(long m, int n, const float* v, bool ok, uint32_t* w) const {
    float b[16], x[16];
    ld16f(bias + n, b);
#pragma unroll
    for (int j = 0; j < 16; j++) { x[j] = v[j] + b[j]; if (relu) x[j] = fmaxf(x[j], 0.0f); }
    if (res && ok) {
      float r[16];
      ld16h<true>(res + m * rstride + n, r);
#pragma unroll
      for (int j = 0; j < 16; j++) x[j] = fmaxf(x[j] + r[j], 0.0f);
    }
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const __half2 h = __floats2half2_rn(x[2 * k], x[2 * k + 1]);
      w[k] = *reinterpret_cast<const uint32_t*>(&h);
    }
  }
  __device__ __forceinline__ void store16(long m, long, int n, const float* v, bool ok, int, int) const {
    if (n >= Cout || !ok) return;
    __half* o = out + m * ostride + ooff + n;
    if (n + 16 <= Cout) {
      float b[16], x[16];
      ld16f(bias + n, b);
#pragma unroll
      for (int j = 0; j < 16; j++) { x[j] = v[j] + b[j]; if (relu) x[j] = fmaxf(x[j], 0.0f); }
      if (res) {
        float r[16];
        ld16h<true>(res + m * rstride + n, r);
#pragma unroll
        for (int j = 0; j < 16; j++) x[j] = fmaxf(x[j] + r[j], 0.0f);
      }
      st16h(o, x);
    } else {
#pragma unroll
      for (int j = 0; j < 16; j++) {
        if (n + j < Cout) {
          float x = v[j] + __ldg(bias + n + j);
          if (relu) x = fmaxf(x, 0.0f);
          if (res) x = fmaxf(x + __half2float(res[m * rstride + n + j]), 0.0f);
          o[j] = __float2half_rn(x);
        }
      }
    }
  }
};

struct EpGloSum {
  static constexpr bool kStaged = false;        // gru.py:25-26: sum over the tile's pixels of sigmoid(w(net)) * net -> partial[tile][quarter][128]
  const float* bias; const __half* net; int nstride; float* partial; int Cout;
  __device__ __forceinline__ void store16(long m, long, int n, const float* v, bool ok, int tile, int q) const {
    if (n >= Cout) return;                       // warp-uniform
    float b[16], h[16], x[16];
    ld16f(bias + n, b);
    if (ok) ld16h<true>(net + m * nstride + n, h);
#pragma unroll
    for (int j = 0; j < 16; j++) x[j] = ok ? fast_sigmoid(v[j] + b[j]) * h[j] : 0.0f;
    // transpose-reduce over the 32 lanes: 8+4+2+1 exchanges halve the values held per lane, then one more step
    const int lane = threadIdx.x & 31;
    xr_step<8, 16>(x, lane);
    xr_step<4, 8>(x, lane);
    xr_step<2, 4>(x, lane);
    xr_step<1, 2>(x, lane);
    x[0] += __shfl_xor_sync(0xffffffffu, x[0], 1);            // lanes 2k, 2k+1 hold the same column
    // column held by this lane pair: bit4 -> +8, bit3 -> +4, bit2 -> +2, bit1 -> +1
    const int col = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
    if (!(lane & 1)) partial[((long)tile * 4 + q) * Cout + n + col] = x[0];
  }
};

struct EpZR {
  static constexpr bool kStaged = false;            // gru.py:28-29 (see conv_simt.cu EpiZR)
  const float* bias; const float* glo; const __half* net; int nstride; __half* z; __half* rnet;
  __device__ __forceinline__ void store16(long m, long e, int n, const float* v, bool ok, int, int) const {
    if (n >= 256 || !ok) return;
    const bool isr = n >= 128;
    const int c0 = isr ? n - 128 : n;
    float b[16], g[16], x[16];
    ld16f(bias + n, b);
    ld16f(glo + e * 384 + n, g);
#pragma unroll
    for (int j = 0; j < 16; j++) x[j] = fast_sigmoid(v[j] + b[j] + g[j]);
    if (isr) {
      float h[16];
      ld16h<true>(net + m * nstride + c0, h);
#pragma unroll
      for (int j = 0; j < 16; j++) x[j] *= h[j];
    }
    st16h((isr ? rnet : z) + m * 128 + c0, x);
  }
};

struct EpQ {
  static constexpr bool kStaged = false;             // gru.py:30-32, state update in place
  const float* bias; const float* glo; const __half* z; __half* net; int nstride;
  __device__ __forceinline__ void store16(long m, long e, int n, const float* v, bool ok, int, int) const {
    if (n >= 128 || !ok) return;
    float b[16], g[16], zz[16], h[16], x[16];
    ld16f(bias + n, b);
    ld16f(glo + e * 384 + 256 + n, g);
    ld16h<true>(z + m * 128 + n, zz);
    ld16h<false>(net + m * nstride + n, h);
#pragma unroll
    for (int j = 0; j < 16; j++) {
      const float q = fast_tanh(v[j] + b[j] + g[j]);
      x[j] = fmaf(zz[j], q - h[j], h[j]);                   // (1-z) h + z q
    }
    st16h(net + m * nstride + n, x);
  }
};

constexpr int kMaxKb = 512;   // k-blocks per tile the on-chip k-block table holds (taps x 64-channel blocks)

// A pipeline stage carries kG consecutive k-blocks (each: A 128x64 + B BNx64 halves), so one barrier round trip
// (producer wait / expect_tx, consumer wait / commit) is paid per kG k-blocks.  Measured (tools/diag_conv.py): with one
// k-block per stage, short-K layers ran at ~0.5 us per k-block whether or not they loaded or multiplied anything --
// the single-thread producer / issuer loops and their barrier operations were the bound, not bytes or FLOPs.
template <int BN> struct Cfg {
  static constexpr int kStageA = kBM * kBK * 2;                 // 16 KB
  static constexpr int kStageB = BN * kBK * 2;
  static constexpr int kKb = kStageA + kStageB;                 // one k-block of operands
  static constexpr int kG = (BN == 64) ? 2 : 1;
  static constexpr int kStage = kG * kKb;
  static constexpr int kStages = (BN == 256) ? 4 : (BN == 128 ? 3 : 2);
  // BN <= 128 tiles are bound by the per-k-block issue overhead of the single producer / issuer threads, not by the
  // tensor pipe: two co-resident CTAs per SM (96 KB of stages and <= 256 TMEM columns each) double that throughput.
  static constexpr int kCtasPerSm = (BN == 256) ? 1 : 2;
  static constexpr int kTmemCols = (2 * BN < 32) ? 32 : 2 * BN; // 128 / 256 / 512: powers of two
  static constexpr int kStaging = (BN == 256) ? 8 * 2048 : 0;   // 8 epilogue warps x (32 rows x 64 B): output transpose for coalesced stores
  static constexpr int kSmem = kStages * kStage + 1024 /*align*/ + 256 /*barriers*/ + kMaxKb * 4 /*k-block table*/ + kStaging;
};

template <int BN, typename Epi>
__global__ void __launch_bounds__(kThreads, Cfg<BN>::kCtasPerSm) conv_tc_kernel(const __grid_constant__ TcParams p, const Epi epi) {
  using C = Cfg<BN>;
  extern __shared__ unsigned char smem_raw[];
  // 1024-byte aligned operand ring (SWIZZLE_128B atoms are 8 rows x 128 B)
  unsigned char* ring = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + C::kStages * C::kStage);
  uint64_t* full = bars;                       // [kStages]  TMA -> MMA
  uint64_t* empty = bars + C::kStages;         // [kStages]  MMA -> TMA
  uint64_t* tfull = bars + 2 * C::kStages;     // [2]        MMA -> epilogue
  uint64_t* tempty = tfull + 2;                // [2]        epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  int* kbt = reinterpret_cast<int*>(bars) + 64;    // [nkb] k-block table: cb | seg << 12 | (dy + 8) << 16 | (dx + 8) << 20

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb = p.taps * p.blocks_per_tap;

  for (int kb = threadIdx.x; kb < nkb; kb += blockDim.x) {      // (tap, segment, channel block) of every k-block, once per CTA
    const int tap = kb / p.blocks_per_tap;
    int rem = kb - tap * p.blocks_per_tap, seg = 0;
    while (rem >= p.seg_blocks[seg]) { rem -= p.seg_blocks[seg]; seg++; }
    kbt[kb] = rem | seg << 12 | (tap / p.KW - p.padH + 8) << 16 | (tap % p.KW - p.padW + 8) << 20;
  }
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.nseg; s++) prefetch_tmap(&p.tmA[s]);
    prefetch_tmap(&p.tmB);
    for (int s = 0; s < C::kStages; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int s = 0; s < 2; s++) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], 8); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, C::kTmemCols);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================================================================== TMA producer (warp-uniform; one elected lane issues)
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int nt = tile % p.n_tiles;
      const int mt = tile / p.n_tiles;
      const int img = mt / (p.tiles_x * p.tiles_y);
      const int r = mt - img * (p.tiles_x * p.tiles_y);
      const int y0 = (r / p.tiles_x) * kTileH, x0 = (r % p.tiles_x) * kTileW;
      for (int kb0 = 0; kb0 < nkb; kb0 += C::kG) {
        const int g = min(C::kG, nkb - kb0);
        mbar_wait(&empty[stage], phase ^ 1);
        if (elect_one()) {
          unsigned char* sa = ring + stage * C::kStage;
          mbar_arrive_expect_tx(&full[stage], g * C::kKb);
#pragma unroll
          for (int j = 0; j < C::kG; j++) {
            if (j < g) {
              const int e = kbt[kb0 + j];
              tma_load_4d(&p.tmA[(e >> 12) & 3], &full[stage], sa + j * C::kKb, (e & 0xfff) * kBK, x0 + ((e >> 20) & 15) - 8,
                          y0 + ((e >> 16) & 15) - 8, img);
              tma_load_2d(&p.tmB, &full[stage], sa + j * C::kKb + C::kStageA, (kb0 + j) * kBK, nt * BN);
            }
          }
        }
        __syncwarp();
        if (++stage == C::kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    int stage = 0;
    uint32_t phase = 0;
    int as = 0;
    uint32_t aphase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      mbar_wait(&tempty[as], aphase ^ 1);
      fence_after_sync();
      const uint32_t idesc = idesc_f16(kBM, min(BN, (p.Cout - (tile % p.n_tiles) * BN + 15) & ~15));     // narrow last N tile: see conv_tc_halo_kernel
      const uint32_t tmem_d = tmem_base + as * BN;
      for (int kb0 = 0; kb0 < nkb; kb0 += C::kG) {
        const int g = min(C::kG, nkb - kb0);
        mbar_wait(&full[stage], phase);
        fence_after_sync();
        if (elect_one()) {
          const uint32_t sa = smem_u32(ring + stage * C::kStage);
#pragma unroll
          for (int j = 0; j < C::kG; j++) {
            if (j < g) {
              const uint64_t da = smem_desc_sw128(sa + j * C::kKb), db = smem_desc_sw128(sa + j * C::kKb + C::kStageA);
#pragma unroll
              for (int k = 0; k < kBK / 16; k++)     // advance 32 B (= 16 halves) inside the swizzle atom: +2 in >>4 units
                umma_f16(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb0 | j | k) != 0);
            }
          }
          umma_commit(&empty[stage]);                // frees the smem stage when these MMAs retire
          if (kb0 + C::kG >= nkb) umma_commit(&tfull[as]);
        }
        __syncwarp();
        if (++stage == C::kStages) { stage = 0; phase ^= 1; }
      }
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
  } else {
    // ===================================================================== epilogue (warps 2..9)
    const int q = warp & 3;                          // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;                // which half of the accumulator columns this warp drains
    int as = 0;
    uint32_t aphase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int nt = tile % p.n_tiles;
      const int mt = tile / p.n_tiles;
      const int img = mt / (p.tiles_x * p.tiles_y);
      const int r = mt - img * (p.tiles_x * p.tiles_y);
      const int row = q * 32 + lane;                 // accumulator row = pixel within the 8x16 patch
      const int y = (r / p.tiles_x) * kTileH + row / kTileW, x = (r % p.tiles_x) * kTileW + row % kTileW;
      const bool ok = y < p.H && x < p.W;
      const long m = ((long)img * p.H + y) * p.W + x;
      mbar_wait(&tfull[as], aphase);
      fence_after_sync();
      const uint32_t taddr = tmem_base + as * BN + ((uint32_t)(q * 32) << 16) + half * (BN / 2);
#pragma unroll 1
      for (int c = 0; c < BN / 2; c += 32) {         // two 16-column loads in flight per wait
        uint32_t v0[16], v1[16];
        tmem_ld16(taddr + c, v0);
        tmem_ld16(taddr + c + 16, v1);
        tmem_ld_wait();
        const int n = nt * BN + half * (BN / 2) + c;
        if constexpr (Epi::kStaged && BN == 256) {
          // Each thread owns one pixel (row), so a direct store touches 32 different rows per instruction (32 L1
          // wavefronts).  Transpose 32 rows x 32 columns through shared memory instead: 4 lanes then write 64 contiguous
          // bytes of one row, 8 rows per instruction.  Chunk position is XOR-swizzled by the row pair: conflict-free both ways.
          if (n + 32 <= epi.Cout) {
            unsigned char* stg = reinterpret_cast<unsigned char*>(kbt + kMaxKb) + (warp - 2) * 2048;
            uint32_t w0[8], w1[8];
            epi.apply16(m, n, reinterpret_cast<const float*>(v0), ok, w0);
            epi.apply16(m, n + 16, reinterpret_cast<const float*>(v1), ok, w1);
            const int sw = (lane >> 1) & 3;
            uint4* rowp = reinterpret_cast<uint4*>(stg + lane * 64);
            rowp[0 ^ sw] = make_uint4(w0[0], w0[1], w0[2], w0[3]);
            rowp[1 ^ sw] = make_uint4(w0[4], w0[5], w0[6], w0[7]);
            rowp[2 ^ sw] = make_uint4(w1[0], w1[1], w1[2], w1[3]);
            rowp[3 ^ sw] = make_uint4(w1[4], w1[5], w1[6], w1[7]);
            __syncwarp();
#pragma unroll
            for (int it = 0; it < 4; it++) {
              const int rr = it * 8 + (lane >> 2), ch = lane & 3;
              const uint4 val = *reinterpret_cast<const uint4*>(stg + rr * 64 + ((ch ^ ((rr >> 1) & 3)) << 4));
              const int pr = q * 32 + rr;
              const int yy = (r / p.tiles_x) * kTileH + pr / kTileW, xx = (r % p.tiles_x) * kTileW + pr % kTileW;
              if (yy < p.H && xx < p.W)
                *reinterpret_cast<uint4*>(epi.out + (((long)img * p.H + yy) * p.W + xx) * epi.ostride + epi.ooff + n + ch * 8) = val;
            }
            __syncwarp();
            continue;
          }
        }
        epi.store16(m, img, n, reinterpret_cast<const float*>(v0), ok, mt, q);
        epi.store16(m, img, n + 16, reinterpret_cast<const float*>(v1), ok, mt, q);
      }
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[as]);
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, C::kTmemCols);
}

template <int BN, typename Epi>
int launch_tc(const TcParams& p, const Epi& epi, cudaStream_t s) {
  using C = Cfg<BN>;
  static bool set = false;
  if (!set) {
    if (cudaFuncSetAttribute(conv_tc_kernel<BN, Epi>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmem) != cudaSuccess)
      return PVO_ERR_CUDA;
    set = true;
  }
  const int slots = PVO_SM_COUNT * C::kCtasPerSm;
  const int grid = p.total_tiles < slots ? p.total_tiles : slots;
  conv_tc_kernel<BN, Epi><<<grid, kThreads, C::kSmem, s>>>(p, epi);
  return pvo_check_launch();
}

// =====================================================================================================================
// 3x3 convolutions with HALO REUSE.  The generic kernel above fetches one 8x16-pixel A box per (tap, channel block): nine TMA
// boxes of 128 pixel rows for every 64 input channels.  Here the M tile is 16 rows x 8 pixels and ONE (16+2) x (8+2) haloed
// patch per channel block (180 pixel rows, 23 KB) serves all nine taps: tap (dy, dx) is the same patch read through a UMMA
// descriptor that starts (dy*10 + dx) rows in, with the 8-row groups (one tile row each) 10 rows = 1280 B apart.  That
// addressing is legal because the 128-byte swizzle depends only on shared-memory address bits (tc_common.cuh, umma_probe.cu).
// A-operand TMA rows per channel block: 180 instead of 1152; the weight tiles stream through their own ring, one tap each.
// MT = 2 (N <= 128 layers): the CTA tile is TWO stacked M tiles (32 rows x 8 pixels, one 34 x 10 patch) that share every weight
// tile -- two MMAs per k-step into two accumulators -- which halves the B-operand stream that bounds those layers after halo reuse.
constexpr int kHTW = 8, kPW = kHTW + 2;

template <int BN, int MT> struct HCfg {
  static constexpr int kTH = 16 * MT, kPH = kTH + 2;            // tile rows, haloed patch rows
  static constexpr int kPatchBytes = kPW * kPH * 128;           // 23040 / 43520
  static constexpr int kPatchSlot = (kPatchBytes + 1023) / 1024 * 1024;
  static constexpr int kStageB = BN * kBK * 2;
  static constexpr int kAStages = 2;
  static constexpr int kBStages = (BN == 256) ? 4 : (MT == 2 ? 4 : (BN == 128 ? 3 : 4));
  static constexpr int kCtasPerSm = (BN == 256 || MT == 2) ? 1 : 2;
  static constexpr int kAccCols = MT * BN;                      // TMEM columns of one accumulator buffer
  static constexpr int kTmemCols = (2 * kAccCols < 32) ? 32 : 2 * kAccCols;
  static constexpr int kStaging = (BN == 256) ? 8 * 2048 : 0;
  static constexpr int kRing = kAStages * kPatchSlot + kBStages * kStageB;
  static constexpr int kSmem = kRing + 1024 /*align*/ + 256 /*barriers*/ + kMaxKb * 4 /*channel-block table*/ + kStaging;
};

template <int BN, int MT, typename Epi>
__global__ void __launch_bounds__(kThreads, HCfg<BN, MT>::kCtasPerSm) conv_tc_halo_kernel(const __grid_constant__ TcParams p, const Epi epi) {
  using C = HCfg<BN, MT>;
  static_assert(MT == 1 || BN <= 128, "two stacked M tiles need 2 x 2 x BN TMEM columns");
  extern __shared__ unsigned char smem_raw[];
  unsigned char* ringA = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* ringB = ringA + C::kAStages * C::kPatchSlot;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ringA + C::kRing);
  uint64_t* afull = bars;                              // [kAStages]
  uint64_t* aempty = afull + C::kAStages;              // [kAStages]
  uint64_t* bfull = aempty + C::kAStages;              // [kBStages]
  uint64_t* bempty = bfull + C::kBStages;              // [kBStages]
  uint64_t* tfull = bempty + C::kBStages;              // [2]
  uint64_t* tempty = tfull + 2;                        // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  int* cbt = reinterpret_cast<int*>(bars) + 64;        // [blocks_per_tap] channel-block table: cb | seg << 12

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ncb = p.blocks_per_tap;
  for (int i = threadIdx.x; i < ncb; i += blockDim.x) {
    int rem = i, seg = 0;
    while (rem >= p.seg_blocks[seg]) { rem -= p.seg_blocks[seg]; seg++; }
    cbt[i] = rem | seg << 12;
  }
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.nseg; s++) prefetch_tmap(&p.tmA[s]);
    prefetch_tmap(&p.tmB);
    for (int s = 0; s < C::kAStages; s++) { mbar_init(&afull[s], 1); mbar_init(&aempty[s], 1); }
    for (int s = 0; s < C::kBStages; s++) { mbar_init(&bfull[s], 1); mbar_init(&bempty[s], 1); }
    for (int s = 0; s < 2; s++) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], 8); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, C::kTmemCols);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================================================================== TMA producer
    int sa = 0, sb = 0;
    uint32_t pa = 0, pb = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int nt = tile % p.n_tiles;
      const int mt = tile / p.n_tiles;
      const int img = mt / (p.tiles_x * p.tiles_y);
      const int r = mt - img * (p.tiles_x * p.tiles_y);
      const int y0 = (r / p.tiles_x) * C::kTH, x0 = (r % p.tiles_x) * kHTW;
      for (int cbg = 0; cbg < ncb; cbg++) {
        const int e = cbt[cbg];
        mbar_wait(&aempty[sa], pa ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&afull[sa], C::kPatchBytes);
          tma_load_4d(&p.tmA[(e >> 12) & 3], &afull[sa], ringA + sa * C::kPatchSlot, (e & 0xfff) * kBK, x0 - 1, y0 - 1, img);
        }
        __syncwarp();
        if (++sa == C::kAStages) { sa = 0; pa ^= 1; }
        for (int tap = 0; tap < 9; tap++) {
          mbar_wait(&bempty[sb], pb ^ 1);
          if (elect_one()) {
            mbar_arrive_expect_tx(&bfull[sb], C::kStageB);
            tma_load_2d(&p.tmB, &bfull[sb], ringB + sb * C::kStageB, (tap * ncb + cbg) * kBK, nt * BN);
          }
          __syncwarp();
          if (++sb == C::kBStages) { sb = 0; pb ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    int sa = 0, sb = 0, as = 0;
    uint32_t pa = 0, pb = 0, aphase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      mbar_wait(&tempty[as], aphase ^ 1);
      fence_after_sync();
      // the last N tile of a layer whose width is not a multiple of BN multiplies only the columns that exist (heads: 640 = 256 + 256
      // + 128): UMMA N is a field of the instruction descriptor, the weight rows past Cout are TMA zero fill that nobody reads
      const uint32_t idesc = idesc_f16(kBM, min(BN, (p.Cout - (tile % p.n_tiles) * BN + 15) & ~15));
      const uint32_t tmem_d = tmem_base + as * C::kAccCols;
      for (int cbg = 0; cbg < ncb; cbg++) {
        mbar_wait(&afull[sa], pa);
        const uint32_t patch = smem_u32(ringA + sa * C::kPatchSlot);
        for (int tap = 0; tap < 9; tap++) {
          mbar_wait(&bfull[sb], pb);
          fence_after_sync();
          if (elect_one()) {
            const uint64_t da = smem_desc_sw128_sbo(patch + ((tap / 3) * kPW + tap % 3) * 128, kPW * 128);
            const uint64_t db = smem_desc_sw128(smem_u32(ringB + sb * C::kStageB));
#pragma unroll
            for (int k = 0; k < kBK / 16; k++) {
              umma_f16(tmem_d, da + 2 * k, db + 2 * k, idesc, (cbg | tap | k) != 0);
              if (MT == 2)                            // lower 16 tile rows: same weights, patch rows 16 further down
                umma_f16(tmem_d + BN, da + 2 * k + ((16 * kPW * 128) >> 4), db + 2 * k, idesc, (cbg | tap | k) != 0);
            }
            umma_commit(&bempty[sb]);
            if (tap == 8) {
              umma_commit(&aempty[sa]);              // the patch is free once all nine taps' MMAs retire
              if (cbg == ncb - 1) umma_commit(&tfull[as]);
            }
          }
          __syncwarp();
          if (++sb == C::kBStages) { sb = 0; pb ^= 1; }
        }
        if (++sa == C::kAStages) { sa = 0; pa ^= 1; }
      }
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
  } else {
    // ===================================================================== epilogue (warps 2..9)
    // MT = 1: two warps per TMEM lane quarter split the BN columns; MT = 2: they split the two stacked M tiles instead.
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    constexpr int kCols = (MT == 2) ? BN : BN / 2;   // columns this warp drains
    int as = 0;
    uint32_t aphase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int nt = tile % p.n_tiles;
      const int mt = tile / p.n_tiles;
      const int img = mt / (p.tiles_x * p.tiles_y);
      const int r = mt - img * (p.tiles_x * p.tiles_y);
      const int row = q * 32 + lane;
      const int ty0 = (r / p.tiles_x) * C::kTH + (MT == 2 ? half * 16 : 0), tx0 = (r % p.tiles_x) * kHTW;
      const int y = ty0 + row / kHTW, x = tx0 + row % kHTW;
      const bool ok = y < p.H && x < p.W;
      const long m = ((long)img * p.H + y) * p.W + x;
      if (lane == 0) mbar_wait(&tfull[as], aphase);
      __syncwarp();
      fence_after_sync();
      const uint32_t taddr = tmem_base + as * C::kAccCols + ((uint32_t)(q * 32) << 16) + (MT == 2 ? half * BN : half * (BN / 2));
#pragma unroll 1
      for (int c = 0; c < kCols; c += 32) {
        uint32_t v0[16], v1[16];
        tmem_ld16(taddr + c, v0);
        tmem_ld16(taddr + c + 16, v1);
        tmem_ld_wait();
        const int n = nt * BN + (MT == 2 ? 0 : half * (BN / 2)) + c;
        if constexpr (Epi::kStaged && BN == 256) {
          if (n + 32 <= epi.Cout) {                   // transposed store through shared memory (see conv_tc_kernel)
            unsigned char* stg = reinterpret_cast<unsigned char*>(cbt + kMaxKb) + (warp - 2) * 2048;
            uint32_t w0[8], w1[8];
            epi.apply16(m, n, reinterpret_cast<const float*>(v0), ok, w0);
            epi.apply16(m, n + 16, reinterpret_cast<const float*>(v1), ok, w1);
            const int sw = (lane >> 1) & 3;
            uint4* rowp = reinterpret_cast<uint4*>(stg + lane * 64);
            rowp[0 ^ sw] = make_uint4(w0[0], w0[1], w0[2], w0[3]);
            rowp[1 ^ sw] = make_uint4(w0[4], w0[5], w0[6], w0[7]);
            rowp[2 ^ sw] = make_uint4(w1[0], w1[1], w1[2], w1[3]);
            rowp[3 ^ sw] = make_uint4(w1[4], w1[5], w1[6], w1[7]);
            __syncwarp();
#pragma unroll
            for (int it = 0; it < 4; it++) {
              const int rr = it * 8 + (lane >> 2), ch = lane & 3;
              const uint4 val = *reinterpret_cast<const uint4*>(stg + rr * 64 + ((ch ^ ((rr >> 1) & 3)) << 4));
              const int pr = q * 32 + rr;
              const int yy = ty0 + pr / kHTW, xx = tx0 + pr % kHTW;
              if (yy < p.H && xx < p.W)
                *reinterpret_cast<uint4*>(epi.out + (((long)img * p.H + yy) * p.W + xx) * epi.ostride + epi.ooff + n + ch * 8) = val;
            }
            __syncwarp();
            continue;
          }
        }
        epi.store16(m, img, n, reinterpret_cast<const float*>(v0), ok, mt, q);
        epi.store16(m, img, n + 16, reinterpret_cast<const float*>(v1), ok, mt, q);
      }
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[as]);
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, C::kTmemCols);
}

template <int BN, int MT, typename Epi>
int launch_tc_halo(const TcParams& p, const Epi& epi, cudaStream_t s) {
  using C = HCfg<BN, MT>;
  static bool set = false;
  if (!set) {
    if (cudaFuncSetAttribute(conv_tc_halo_kernel<BN, MT, Epi>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmem) != cudaSuccess)
      return PVO_ERR_CUDA;
    set = true;
  }
  const int slots = PVO_SM_COUNT * C::kCtasPerSm;
  const int grid = p.total_tiles < slots ? p.total_tiles : slots;
  conv_tc_halo_kernel<BN, MT, Epi><<<grid, kThreads, C::kSmem, s>>>(p, epi);
  return pvo_check_launch();
}

template <typename Epi>
int dispatch_bn(const TcParams& p0, const Epi& epi, int bn, cudaStream_t s) {
  TcParams p = p0;
  p.n_tiles = (p.Cout + bn - 1) / bn;
  p.total_tiles = p.nimg * p.tiles_x * p.tiles_y * p.n_tiles;
  if (p.halo == 2) return launch_tc_halo<128, 2>(p, epi, s);
  if (p.halo) {
    if (bn == 256) return launch_tc_halo<256, 1>(p, epi, s);
    if (bn == 128) return launch_tc_halo<128, 1>(p, epi, s);
    return launch_tc_halo<64, 1>(p, epi, s);
  }
  if (bn == 256) return launch_tc<256>(p, epi, s);
  if (bn == 128) return launch_tc<128>(p, epi, s);
  return launch_tc<64>(p, epi, s);
}

}  // namespace

// Segments: pixel-major fp16 tensors (nimg*H*W, stride) whose channel slices are concatenated into the conv input;
// seg_channels[s] need not be a multiple of 64 (the tail is zero-filled by TMA; the packed weights carry zeros there).
// Wh: fp16 [Cout][taps * sum_s ceil64(seg_channels[s])], K ordered (tap, segment, channel) -- pack_conv_weight_tc.
// epilogues 0/1 accept an optional residual in aux0 (row stride aux_stride): out = relu(residual + act(conv)).
// epilogue: 0 bias, 1 bias+ReLU, 2 per-tile sums of sigmoid(.)*net, 3 GRU z|r, 4 GRU q + state update (same meaning and aux pointers as pvo_conv2d).
extern "C" int pvo_conv2d_tc(const void* const* seg_ptr, const int* seg_channels, const int* seg_stride, int nseg,
                             const void* Wh, const float* bias, int Cout, int KH, int KW, int nimg, int H, int Wd,
                             int epilogue, void* out, int ostride, int ooff, void* out2, const void* aux0, int aux_stride,
                             const float* aux1, void* stream) {
  if (nseg < 1 || nseg > kMaxSeg || Cout < 1 || nimg < 0 || !(KH & 1) || !(KW & 1) || KH > 15 || KW > 15) return PVO_ERR_INVALID;
  if (nimg == 0) return PVO_OK;
  TcParams p;
  static const bool no_halo = getenv("PVO_TC_NO_HALO") != nullptr;     // A/B switches for measurements
  static const bool no_m2 = getenv("PVO_TC_NO_M2") != nullptr;
  p.halo = (KH == 3 && KW == 3 && !no_halo) ? 1 : 0;
  p.nseg = nseg;
  p.blocks_per_tap = 0;
  for (int s = 0; s < kMaxSeg; s++) p.seg_blocks[s] = 0;
  for (int s = 0; s < nseg; s++) {
    if (seg_channels[s] % 8 || (seg_stride[s] * 2) % 16 || ((uintptr_t)seg_ptr[s]) % 16) return PVO_ERR_INVALID;
    p.seg_blocks[s] = (seg_channels[s] + kBK - 1) / kBK;
    p.blocks_per_tap += p.seg_blocks[s];
  }
  p.taps = KH * KW; p.KW = KW; p.padH = KH / 2; p.padW = KW / 2;
  p.H = H; p.W = Wd; p.nimg = nimg;
  p.tiles_x = p.halo ? (Wd + kHTW - 1) / kHTW : (Wd + kTileW - 1) / kTileW;
  p.tiles_y = p.halo ? (H + 15) / 16 : (H + kTileH - 1) / kTileH;
  p.Cout = Cout;
  if (p.taps * p.blocks_per_tap > kMaxKb) return PVO_ERR_UNSUPPORTED;
  const long Ktot = (long)p.taps * p.blocks_per_tap * kBK;
  int bn = Cout >= 256 ? 256 : (Cout >= 128 ? 128 : 64);
  // keep at least ~2 waves of tiles: prefer the narrower tile when the wide one leaves SMs idle
  {
    const long mt = (long)nimg * p.tiles_x * p.tiles_y;
    while (bn > 64 && mt * ((Cout + bn - 1) / bn) < 2 * PVO_SM_COUNT) bn >>= 1;
  }
  // N = 128 halo layers: stack two M tiles per CTA (they share the weight stream) when that still leaves >= 3 waves.
  // (Measured: q conv 0.172 -> 0.156 ms; N = 64 layers are faster with two independent CTAs per SM and stay on MT = 1;
  //  running the N = 256 z|r conv as 128-wide stacked tiles instead was slower, 0.244 -> 0.287 ms.)
  if (p.halo && bn == 128 && !no_m2) {
    const long t2 = (long)nimg * p.tiles_x * ((H + 31) / 32) * ((Cout + bn - 1) / bn);
    if (t2 >= 3 * PVO_SM_COUNT) { p.halo = 2; p.tiles_y = (H + 31) / 32; }
  }
  for (int s = 0; s < nseg; s++) {
    const uint64_t dims[4] = {(uint64_t)seg_channels[s], (uint64_t)Wd, (uint64_t)H, (uint64_t)nimg};
    const uint64_t strides[3] = {(uint64_t)seg_stride[s] * 2, (uint64_t)seg_stride[s] * 2 * Wd, (uint64_t)seg_stride[s] * 2 * Wd * H};
    const uint32_t box[4] = {kBK, (uint32_t)(p.halo ? kPW : kTileW), (uint32_t)(p.halo ? 16 * p.halo + 2 : kTileH), 1};
    const int rc = make_tmap_f16(&p.tmA[s], seg_ptr[s], 4, dims, strides, box);
    if (rc != PVO_OK) return rc;
  }
  for (int s = nseg; s < kMaxSeg; s++) p.tmA[s] = p.tmA[0];
  {
    const uint64_t dims[2] = {(uint64_t)Ktot, (uint64_t)Cout};
    const uint64_t strides[1] = {(uint64_t)Ktot * 2};
    const uint32_t box[2] = {kBK, (uint32_t)bn};
    const int rc = make_tmap_f16(&p.tmB, Wh, 2, dims, strides, box);
    if (rc != PVO_OK) return rc;
  }
  cudaStream_t s = (cudaStream_t)stream;
  switch (epilogue) {
    case 0: case 1:
      if ((ostride * 2) % 16 || (ooff * 2) % 16) return PVO_ERR_INVALID;
      return dispatch_bn(p, EpBias{bias, (__half*)out, ostride, ooff, Cout, epilogue, (const __half*)aux0, aux_stride}, bn, s);
    case 2:      // out = float partial sums [nimg * tiles][4][Cout] (reduced by pvo_glo_reduce_matvec)
      if (Cout % 32) return PVO_ERR_INVALID;
      return dispatch_bn(p, EpGloSum{bias, (const __half*)aux0, aux_stride, (float*)out, Cout}, bn, s);
    case 3:
      if (Cout != 256) return PVO_ERR_INVALID;
      return dispatch_bn(p, EpZR{bias, aux1, (const __half*)aux0, aux_stride, (__half*)out, (__half*)out2}, bn, s);
    case 4:
      if (Cout != 128) return PVO_ERR_INVALID;
      return dispatch_bn(p, EpQ{bias, aux1, (const __half*)aux0, (__half*)out, ostride}, bn, s);
    default:
      return PVO_ERR_UNSUPPORTED;
  }
}
