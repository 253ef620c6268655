// All-pairs correlation volume + 4-level pooling pyramid on tcgen05 tensor cores, one pass.
//
// Replaces CorrBlock.__init__ (VO_Module/droid_slam/modules/corr.py:23-38,63-71): torch.matmul on (E,HW,128)x(E,128,HW)
// half operands followed by three F.avg_pool2d passes that re-read and re-write the volume.  Here each edge's
// C = F1 . F2^T is produced tile by tile in TMEM and the epilogue emits ALL FOUR levels from registers: level 0 is
// written once and never re-read (61 MB/edge written, 0 read back -- the reference moves ~2x that).
//
//   work item = (edge, 128 source pixels);  A = F1 rows [128 x 128] loaded once per item (2 SW128 K blocks)
//   chunk     = 2 rows of image 2 = N = 2*W columns (one UMMA N);  B = F2 rows [2W x 128], TMA ring of 3 stages
//   accumulators: 2 x (2W) TMEM columns (double buffered) so the MMA of chunk c+1 overlaps the epilogue of chunk c
//   epilogue thread = one source pixel (TMEM lane): rounds to half (level 0), pools 2x2 of the ROUNDED values in
//   float in raster order and rounds again (levels 1..3), exactly the cascade avg_pool2d applies.
// Inputs are pixel-major (E, HW, 128) fp16 ALREADY divided by 4 (the reference's `fmap / 4.0`, corr.py:66-67).
// HBM-bound on the level-0 write: algorithmic bytes = 2 * (1 + 1/4 + 1/16 + 1/64) * HW^2 per edge.
#include "tc_common.cuh"

namespace {

using namespace tc;

constexpr int kThreads = 192;
constexpr int kStagesB = 3;

struct VolParams {
  CUtensorMap tmA, tmB;
  __half* lvl[4];
  const int* slots;          // TILED: edge e is written to plane block slots[e] of the level tensors (NULL: e)
  int E, H, HW, HW2, mtiles, items, nchunks;
  int wt;                    // true width of image 2 (<= W); columns wt..W-1 of the B operand are zero features
};

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t v[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr));
}

__device__ __forceinline__ uint32_t h2u(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }
__device__ __forceinline__ __half2 u2h2(uint32_t u) { return *reinterpret_cast<__half2*>(&u); }

__device__ __forceinline__ __half pool4(__half a, __half b, __half c, __half d) {
  float s = __half2float(a);
  s += __half2float(b);
  s += __half2float(c);
  s += __half2float(d);
  return __float2half_rn(s * 0.25f);
}

// TILED: every (edge, source pixel) plane is written as tiles of 4 rows x 8 columns (64 B), tile-row-major, zero padded to whole
// tiles -- the layout corr_lookup_tiled.cu gathers from: level l plane = [ceil(h_l/4)][ceil(w_l/8)][4][8] halves.  Level 0 leaves
// the epilogue as whole 32-byte sectors (two rows of one tile per store pair).
template <int W, bool TILED>
__global__ void __launch_bounds__(kThreads, 1) corr_volume_tc_kernel(const __grid_constant__ VolParams p) {
  constexpr int N = 2 * W;                       // UMMA N = two rows of image 2
  constexpr int kA = 2 * 128 * 128;              // A: two K blocks of [128 rows][64 halves]
  constexpr int kBblk = N * 128;                 // one K block of B
  constexpr int kStage = 2 * kBblk;
  constexpr int W1 = W / 2, W2 = W / 4, W3 = W / 8;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* smA = base;
  unsigned char* smB = base + kA;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smB + kStagesB * kStage);
  uint64_t* full = bars;                  // [3]
  uint64_t* empty = bars + kStagesB;      // [3]
  uint64_t* afull = empty + kStagesB;     // [1]
  uint64_t* aempty = afull + 1;           // [1]
  uint64_t* tfull = aempty + 1;           // [2]
  uint64_t* tempty = tfull + 2;           // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.tmA);
    prefetch_tmap(&p.tmB);
    for (int s = 0; s < kStagesB; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(afull, 1);
    mbar_init(aempty, 1);
    for (int s = 0; s < 2; s++) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], 4); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  constexpr int kAcc = (N <= 256) ? 256 : 0;     // column offset of the second accumulator

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0, aph = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
        const int e = item / p.mtiles, mt = item - e * p.mtiles;
        mbar_wait(aempty, aph ^ 1);
        mbar_arrive_expect_tx(afull, kA);
        tma_load_2d(&p.tmA, afull, smA, 0, e * p.HW + mt * 128);      // rows past this edge belong to the next one:
        tma_load_2d(&p.tmA, afull, smA + kA / 2, 64, e * p.HW + mt * 128);   // harmless, those lanes are never stored
        aph ^= 1;
        for (int c = 0; c < p.nchunks; c++) {
          mbar_wait(&empty[stage], phase ^ 1);
          unsigned char* sb = smB + stage * kStage;
          mbar_arrive_expect_tx(&full[stage], kStage);
          tma_load_2d(&p.tmB, &full[stage], sb, 0, e * p.HW2 + c * N);
          tma_load_2d(&p.tmB, &full[stage], sb + kBblk, 64, e * p.HW2 + c * N);
          if (++stage == kStagesB) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = idesc_f16(128, N);
    int stage = 0, as = 0;
    uint32_t phase = 0, aphase = 0, aph = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
      mbar_wait(afull, aph);
      aph ^= 1;
      for (int c = 0; c < p.nchunks; c++) {
        mbar_wait(&tempty[as], aphase ^ 1);
        mbar_wait(&full[stage], phase);
        fence_after_sync();
        if (elect_one()) {
          const uint32_t sa = smem_u32(smA), sb = smem_u32(smB + stage * kStage);
          const uint32_t tmem_d = tmem_base + as * kAcc;
#pragma unroll
          for (int kb = 0; kb < 2; kb++) {
            const uint64_t da = smem_desc_sw128(sa + kb * (kA / 2)), db = smem_desc_sw128(sb + kb * kBblk);
#pragma unroll
            for (int k = 0; k < 4; k++) umma_f16(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
          }
          umma_commit(&empty[stage]);
          umma_commit(&tfull[as]);
          if (c == p.nchunks - 1) umma_commit(aempty);
        }
        __syncwarp();
        if (++stage == kStagesB) { stage = 0; phase ^= 1; }
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
    }
  } else {
    const int q = warp & 3;
    int as = 0;
    uint32_t aphase = 0;
    const int H = p.H, H1 = H / 2, H2 = H1 / 2, H3 = H2 / 2;
    // image 2 narrower than the template width (its feature map was zero padded to W columns): level 0 is exact by itself (zero
    // features give zero correlations); the pooled levels must end at (wt >> l) like avg_pool2d's floor, so their columns past
    // that are forced to zero before they are stored or pooled further
    const bool masked = p.wt != W;
    const int wt1 = p.wt >> 1, wt2 = p.wt >> 2, wt3 = p.wt >> 3;
    constexpr int TC0 = W / 8, TC1 = (W1 + 7) / 8, TC2 = (W2 + 7) / 8, TC3 = (W3 + 7) / 8;
    const int TR0 = (H + 3) / 4, TR1 = (H1 + 3) / 4, TR2 = (H2 + 3) / 4, TR3 = (H3 + 3) / 4;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
      const int e = item / p.mtiles, mt = item - e * p.mtiles;
      const int p1 = mt * 128 + q * 32 + lane;
      const bool ok = p1 < p.HW;
      const long rowi = (long)((TILED && p.slots) ? __ldg(p.slots + e) : e) * p.HW + p1;
      __half* o0 = p.lvl[0] + rowi * (TILED ? (long)TR0 * TC0 * 32 : (long)H * W);
      __half* o1 = p.lvl[1] + rowi * (TILED ? (long)TR1 * TC1 * 32 : (long)H1 * W1);
      __half* o2 = p.lvl[2] + rowi * (TILED ? (long)TR2 * TC2 * 32 : (long)H2 * W2);
      __half* o3 = p.lvl[3] + rowi * (TILED ? (long)TR3 * TC3 * 32 : (long)H3 * W3);
      // pooled rows are kept packed (two halves per register); every loop below is fully unrolled so they stay in
      // registers.  pair x of a level-k row = columns (2x, 2x+1).
      __half2 l1prev[W1 / 2], l1cur[W1 / 2], l2prev[W2 / 2 + 1], l2cur[W2 / 2 + 1];
      for (int c = 0; c < p.nchunks; c++) {
        mbar_wait(&tfull[as], aphase);
        fence_after_sync();
        const uint32_t taddr = tmem_base + as * kAcc + ((uint32_t)(q * 32) << 16);
        const bool row1 = 2 * c + 1 < H;
#pragma unroll
        for (int j = 0; j < W / 8; j++) {
          uint32_t a[8], b[8];
          tmem_ld8(taddr + j * 8, a);
          tmem_ld8(taddr + W + j * 8, b);
          tmem_ld_wait();
          __half2 pa[4], pb[4];
#pragma unroll
          for (int t = 0; t < 4; t++) {
            pa[t] = __floats2half2_rn(__uint_as_float(a[2 * t]), __uint_as_float(a[2 * t + 1]));
            pb[t] = __floats2half2_rn(__uint_as_float(b[2 * t]), __uint_as_float(b[2 * t + 1]));
          }
          if (ok) {
            if (TILED) {     // rows 2c, 2c+1 = sub-rows (2(c&1), 2(c&1)+1) of tile (c>>1, j): 32 contiguous bytes = one sector
              __half* t = o0 + ((long)(c >> 1) * TC0 + j) * 32 + (c & 1) * 16;
              *reinterpret_cast<uint4*>(t) = make_uint4(h2u(pa[0]), h2u(pa[1]), h2u(pa[2]), h2u(pa[3]));
              *reinterpret_cast<uint4*>(t + 8) = row1 ? make_uint4(h2u(pb[0]), h2u(pb[1]), h2u(pb[2]), h2u(pb[3])) : make_uint4(0u, 0u, 0u, 0u);
            } else {
              *reinterpret_cast<uint4*>(o0 + (long)(2 * c) * W + j * 8) = make_uint4(h2u(pa[0]), h2u(pa[1]), h2u(pa[2]), h2u(pa[3]));
              if (row1)
                *reinterpret_cast<uint4*>(o0 + (long)(2 * c + 1) * W + j * 8) = make_uint4(h2u(pb[0]), h2u(pb[1]), h2u(pb[2]), h2u(pb[3]));
            }
          }
#pragma unroll
          for (int t = 0; t < 2; t++)
            l1cur[2 * j + t] = __halves2half2(pool4(__low2half(pa[2 * t]), __high2half(pa[2 * t]), __low2half(pb[2 * t]), __high2half(pb[2 * t])),
                                              pool4(__low2half(pa[2 * t + 1]), __high2half(pa[2 * t + 1]), __low2half(pb[2 * t + 1]), __high2half(pb[2 * t + 1])));
        }
        if (masked) {
#pragma unroll
          for (int x = 0; x < W1 / 2; x++)
            l1cur[x] = u2h2(h2u(l1cur[x]) & ((2 * x + 1 < wt1) ? 0xffffffffu : ((2 * x < wt1) ? 0x0000ffffu : 0u)));
        }
        // the accumulator is drained: hand it back before the (register-only) pooling tail
        fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[as]);
        if (++as == 2) { as = 0; aphase ^= 1; }
        if (ok && c < H1) {
          if (TILED) {
#pragma unroll
            for (int x = 0; x < TC1; x++)
              *reinterpret_cast<uint4*>(o1 + ((long)(c >> 2) * TC1 + x) * 32 + (c & 3) * 8) =
                  make_uint4(h2u(l1cur[4 * x]), (4 * x + 1 < W1 / 2) ? h2u(l1cur[4 * x + 1]) : 0u,
                             (4 * x + 2 < W1 / 2) ? h2u(l1cur[4 * x + 2]) : 0u, (4 * x + 3 < W1 / 2) ? h2u(l1cur[4 * x + 3]) : 0u);
          } else {
#pragma unroll
            for (int x = 0; x < W1 / 4; x++)
              *reinterpret_cast<uint2*>(o1 + (long)c * W1 + 4 * x) = make_uint2(h2u(l1cur[2 * x]), h2u(l1cur[2 * x + 1]));
          }
        }
        if (c & 1) {
          const int r2 = c >> 1;
#pragma unroll
          for (int x = 0; x < W2 / 2; x++)
            l2cur[x] = __halves2half2(pool4(__low2half(l1prev[2 * x]), __high2half(l1prev[2 * x]), __low2half(l1cur[2 * x]), __high2half(l1cur[2 * x])),
                                      pool4(__low2half(l1prev[2 * x + 1]), __high2half(l1prev[2 * x + 1]), __low2half(l1cur[2 * x + 1]), __high2half(l1cur[2 * x + 1])));
          if (masked) {
#pragma unroll
            for (int x = 0; x < W2 / 2; x++)
              l2cur[x] = u2h2(h2u(l2cur[x]) & ((2 * x + 1 < wt2) ? 0xffffffffu : ((2 * x < wt2) ? 0x0000ffffu : 0u)));
          }
          if (ok && r2 < H2) {
            if (TILED) {
#pragma unroll
              for (int t = 0; t < TC2; t++)
                *reinterpret_cast<uint4*>(o2 + ((long)(r2 >> 2) * TC2 + t) * 32 + (r2 & 3) * 8) =
                    make_uint4(h2u(l2cur[4 * t]), (4 * t + 1 < W2 / 2) ? h2u(l2cur[4 * t + 1]) : 0u,
                               (4 * t + 2 < W2 / 2) ? h2u(l2cur[4 * t + 2]) : 0u, (4 * t + 3 < W2 / 2) ? h2u(l2cur[4 * t + 3]) : 0u);
            } else {
#pragma unroll
              for (int x = 0; x < W2 / 2; x++) *reinterpret_cast<uint32_t*>(o2 + (long)r2 * W2 + 2 * x) = h2u(l2cur[x]);
            }
          }
          if (r2 & 1) {
            const int r3 = r2 >> 1;
            if (ok && r3 < H3) {
              if (TILED) {
                __half v3[TC3 * 8];
#pragma unroll
                for (int x = 0; x < TC3 * 8; x++)
                  v3[x] = (x < W3 && x < wt3) ? pool4(__low2half(l2prev[x < W3 ? x : 0]), __high2half(l2prev[x < W3 ? x : 0]), __low2half(l2cur[x < W3 ? x : 0]),
                                         __high2half(l2cur[x < W3 ? x : 0]))
                                 : __float2half_rn(0.0f);
#pragma unroll
                for (int t = 0; t < TC3; t++)
                  *reinterpret_cast<uint4*>(o3 + ((long)(r3 >> 2) * TC3 + t) * 32 + (r3 & 3) * 8) =
                      make_uint4(h2u(__halves2half2(v3[8 * t], v3[8 * t + 1])), h2u(__halves2half2(v3[8 * t + 2], v3[8 * t + 3])),
                                 h2u(__halves2half2(v3[8 * t + 4], v3[8 * t + 5])), h2u(__halves2half2(v3[8 * t + 6], v3[8 * t + 7])));
              } else {
#pragma unroll
                for (int x = 0; x < W3; x++)
                  o3[(long)r3 * W3 + x] = pool4(__low2half(l2prev[x]), __high2half(l2prev[x]), __low2half(l2cur[x]), __high2half(l2cur[x]));
              }
            }
          } else {
#pragma unroll
            for (int x = 0; x < W2 / 2; x++) l2prev[x] = l2cur[x];
          }
        } else {
#pragma unroll
          for (int x = 0; x < W1 / 2; x++) l1prev[x] = l1cur[x];
        }
      }
      if (TILED && ok) {      // rows past the last written one, up to the end of the last tile row, are zeros in every level
        const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
        for (int r = 2 * p.nchunks; r < 4 * TR0; r++)
          for (int t = 0; t < TC0; t++) *reinterpret_cast<uint4*>(o0 + ((long)(r >> 2) * TC0 + t) * 32 + (r & 3) * 8) = z4;
        for (int r = H1; r < 4 * TR1; r++)
          for (int t = 0; t < TC1; t++) *reinterpret_cast<uint4*>(o1 + ((long)(r >> 2) * TC1 + t) * 32 + (r & 3) * 8) = z4;
        for (int r = H2; r < 4 * TR2; r++)
          for (int t = 0; t < TC2; t++) *reinterpret_cast<uint4*>(o2 + ((long)(r >> 2) * TC2 + t) * 32 + (r & 3) * 8) = z4;
        for (int r = H3; r < 4 * TR3; r++)
          for (int t = 0; t < TC3; t++) *reinterpret_cast<uint4*>(o3 + ((long)(r >> 2) * TC3 + t) * 32 + (r & 3) * 8) = z4;
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

template <int W, bool TILED>
int launch_vol(const VolParams& p, cudaStream_t s) {
  constexpr int smem = 2 * 128 * 128 + kStagesB * 2 * (2 * W) * 128 + 1024 + 256;
  static bool set = false;
  if (!set) {
    if (cudaFuncSetAttribute(corr_volume_tc_kernel<W, TILED>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
      return PVO_ERR_CUDA;
    set = true;
  }
  const int grid = p.items < PVO_SM_COUNT ? p.items : PVO_SM_COUNT;
  corr_volume_tc_kernel<W, TILED><<<grid, kThreads, smem, s>>>(p);
  return pvo_check_launch();
}

}  // namespace

// fmap1, fmap2: (E, h*w, 128) fp16 pixel-major, pre-divided by 4.  levels_out[l]: (E, h, w, h>>l, w>>l) fp16 (planar) or tiled.
// Supported: num_levels == 4, h >= 8, template widths {40, 80} (planar and tiled) and {104, 128} (tiled; narrower images are zero
// padded to the template width: N = 2 * wpad must be a legal UMMA N and the pooling rows stay register resident);
// anything else returns PVO_ERR_UNSUPPORTED and the caller uses pvo_corr_volume.
static int corr_volume_f16_impl(const void* fmap1, const void* fmap2, void* const* levels_out, const int* slots, int E, int h, int w,
                                int wpad, int num_levels, bool tiled, void* stream) {
  if (E < 0 || h <= 0 || w <= 0) return PVO_ERR_INVALID;
  if (num_levels != 4 || h < 8) return PVO_ERR_UNSUPPORTED;
  if (!(wpad == 40 || wpad == 80 || wpad == 104 || wpad == 128) || w > wpad || (!tiled && w != wpad)) return PVO_ERR_UNSUPPORTED;
  if (E == 0) return PVO_OK;
  VolParams p;
  const int HW = h * w, HW2 = h * wpad;
  const uint64_t dimsA[2] = {128, (uint64_t)E * HW}, dimsB[2] = {128, (uint64_t)E * HW2};
  const uint64_t strides[1] = {256};
  const uint32_t boxA[2] = {64, 128};
  const uint32_t boxB[2] = {64, (uint32_t)(2 * wpad)};
  int rc = tc::make_tmap_f16(&p.tmA, fmap1, 2, dimsA, strides, boxA);
  if (rc != PVO_OK) return rc;
  rc = tc::make_tmap_f16(&p.tmB, fmap2, 2, dimsB, strides, boxB);
  if (rc != PVO_OK) return rc;
  for (int l = 0; l < 4; l++) p.lvl[l] = (__half*)levels_out[l];
  p.slots = slots;
  p.E = E; p.H = h; p.HW = HW; p.HW2 = HW2; p.wt = w;
  p.mtiles = (HW + 127) / 128;
  p.items = E * p.mtiles;
  p.nchunks = (h + 1) / 2;
  cudaStream_t s = (cudaStream_t)stream;
  if (tiled) {
    switch (wpad) {
      case 40: return launch_vol<40, true>(p, s);
      case 80: return launch_vol<80, true>(p, s);
      case 104: return launch_vol<104, true>(p, s);
      default: return launch_vol<128, true>(p, s);
    }
  }
  return wpad == 80 ? launch_vol<80, false>(p, s) : (wpad == 40 ? launch_vol<40, false>(p, s) : PVO_ERR_UNSUPPORTED);
}

extern "C" int pvo_corr_volume_f16(const void* fmap1, const void* fmap2, void* const* levels_out, int E, int h, int w,
                                   int num_levels, void* stream) {
  return corr_volume_f16_impl(fmap1, fmap2, levels_out, nullptr, E, h, w, w, num_levels, false, stream);
}

// levels_out[l]: (capacity, h*w, ceil((h>>l)/4), ceil((wpad>>l)/8), 32) fp16 -- the tiled private layout of corr_lookup_tiled.cu.
// fmap1 (E, h*w, 128); fmap2 (E, h*wpad, 128) with zero features in columns w..wpad-1 (wpad in {40, 80, 104, 128}: the UMMA N is
// 2*wpad).  slots (E ints, or NULL): edge e is written to plane block slots[e] of the level tensors.
extern "C" int pvo_corr_volume_f16_tiled(const void* fmap1, const void* fmap2, void* const* levels_out, const int* slots, int E,
                                         int h, int w, int wpad, int num_levels, void* stream) {
  return corr_volume_f16_impl(fmap1, fmap2, levels_out, slots, E, h, w, wpad, num_levels, true, stream);
}
