// Correlation window lookup over the TILED private volume layout (production half path).
//
// Same arithmetic, bit for bit, as corr_lookup.cu / the reference's corr_index_forward_kernel
// (VO_Module/src/correlation_kernels.cu:19-70) for radius 3 on half volumes; what changes is where the bytes come from.
//
// Why a new layout.  In the reference layout every source pixel owns a row-major (h2, w2) plane; an 8x8 window is 8 pieces of
// 16 bytes, each from a different 160-byte (level 0) row: every piece costs its own 64-byte DRAM access -- ncu measured 433 MB of
// DRAM traffic per config-2 launch against 162 MB algorithmic -- and 19 % / 37 % / 65 % / 100 % of the windows of levels 0..3
// overlap the plane border and took the per-tap generic path.
// Here each plane is stored as tiles of 4 rows x 8 columns (64 bytes = one DRAM access, two 32-byte sectors), tile-row-major, zero
// padded to whole tiles: (E, h1*w1, ceil(h2/4), ceil(w2/8), 4, 8).  The 8 rows of a window at any alignment lie in <= 2 tile
// columns and <= 3 tile rows; a window row is one 16-byte piece per tile column, so a lane fetches exactly the 8 x 2 pieces it
// needs (only one column when the window is column aligned), neighbouring pieces share sectors, every byte of a fetched sector
// belongs to the plane, and out-of-plane pieces are never fetched: they are zero-filled by cp.async's src-size operand, which is
// arithmetically identical to the reference skipping the tap (the running sum starts at +0 and +0 * w = +0 leaves it unchanged).
// ONE code path serves interior, border and fully-outside windows.
//
// Kernel: block = 32 pixel-edges x 4 levels (warp <-> level, lane <-> pixel), 6 blocks per SM.  Each lane stages its 16 pieces
// (256 B) into its own shared-memory slot with 16-byte cp.async.cg (L1 bypass: nothing is reused), window-relative, so the blend
// loop indexes rows statically; a resident block keeps 32 KB of gathers in flight without holding them in registers.  Rows come
// back with two conflict-free LDS.128 (slot stride 272 B), are realigned with selects + funnel shifts, blended in packed half
// (non-fused HMUL2/HADD2, the reference's rounding order), and the 7 outputs of a window row overwrite the slot row that was just
// consumed (one STS.128).  After a block barrier the 196 outputs of each pixel are gathered from the four level slots through a
// channel->offset table and written as fully coalesced 128-byte lines straight into the update operator's pixel-major input.
#include "common.cuh"

namespace {

struct TLevels {
  const __half* vol[4];
  int tr[4];   // tile rows  = ceil(h2 / 4)
  int tc[4];   // tile cols  = ceil(w2 / 8)  (w2 = the padded width the volume was built with)
  const int* slots;   // edge -> plane block of the level tensors (NULL: identity).  FactorGraph keeps its volumes in a pool so that
                      // adding / retiring edges never copies them.
};

constexpr int kSlot = 272;            // bytes per lane slot: 8 window rows x 32 B, +16 so that LDS.128 / STS.128 are conflict free
constexpr int kLevelBytes = 32 * kSlot;

__device__ __forceinline__ int safe_floor_t(float x) {
  float f = floorf(x);
  f = fminf(fmaxf(f, -16777216.0f), 16777216.0f);
  return (f == f) ? (int)f : -16777216;
}

__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}

__device__ __forceinline__ __half2 u2h(uint32_t u) { return *reinterpret_cast<__half2*>(&u); }
__device__ __forceinline__ uint32_t h2u_(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }

// packed words of the 8 taps of window row j (slot row j = 16 halves starting at tile column tcx0), taps start at half `ox` (0..7)
__device__ __forceinline__ void slot_row(const unsigned char* slot, int j, int ox, uint32_t r[4]) {
  const uint4 A = *reinterpret_cast<const uint4*>(slot + j * 32);
  const uint4 B = *reinterpret_cast<const uint4*>(slot + j * 32 + 16);
  const bool s2 = ox & 4, s1 = ox & 2;
  const uint32_t d0 = s2 ? A.z : A.x, d1 = s2 ? A.w : A.y, d2 = s2 ? B.x : A.z, d3 = s2 ? B.y : A.w, d4 = s2 ? B.z : B.x,
                 d5 = s2 ? B.w : B.y;
  const uint32_t e0 = s1 ? d1 : d0, e1 = s1 ? d2 : d1, e2 = s1 ? d3 : d2, e3 = s1 ? d4 : d3, e4 = s1 ? d5 : d4;
  const uint32_t sh = (ox & 1) * 16;
  r[0] = __funnelshift_r(e0, e1, sh);
  r[1] = __funnelshift_r(e1, e2, sh);
  r[2] = __funnelshift_r(e2, e3, sh);
  r[3] = __funnelshift_r(e3, e4, sh);
}

//   LAYOUT 0: planar  out[e][l*49 + a*7 + b][pix]      LAYOUT 1: pixel-major out[e*HW + pix][Cs], pad channels zero
template <int LAYOUT>
__global__ void __launch_bounds__(128, 6) corr_lookup_tiled_kernel(const __grid_constant__ TLevels lv, const float2* __restrict__ coords,
                                                                   __half* __restrict__ out, long npix_total, int HW, int Cs) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, l = threadIdx.x >> 5;
  const long p = (long)blockIdx.x * 32 + lane;
  const bool active = p < npix_total;
  unsigned char* slot = smem_raw + l * kLevelBytes + lane * kSlot;
  uint16_t* tab = reinterpret_cast<uint16_t*>(smem_raw + 4 * kLevelBytes);      // channel -> byte offset of its value (pixel 0)
  if (LAYOUT == 1) {
    for (int c = threadIdx.x; c < 196; c += blockDim.x) {
      const int cl = c / 49, k = c - cl * 49, a = k / 7, b = k - a * 7;
      tab[c] = (uint16_t)(cl * kLevelBytes + b * 32 + a * 2);
    }
  }
  if (active) {
    const float2 c = __ldg(coords + p);
    const float sc = 1.0f / (float)(1 << l);                        // coords / 2**l is exact in float
    const float x0 = c.x * sc, y0 = c.y * sc;
    const int fx = safe_floor_t(x0), fy = safe_floor_t(y0);
    const int TR = lv.tr[l], TC = lv.tc[l];
    const int wx0 = fx - 3, wy0 = fy - 3;
    const int tcx0 = wx0 >> 3, ox = wx0 & 7;
    long pb = p;
    if (lv.slots) {
      const long e = p / HW;
      pb = (long)__ldg(lv.slots + e) * HW + (p - e * HW);
    }
    const __half* plane = lv.vol[l] + pb * ((long)TR * TC * 32);
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(slot);
    const bool vx0 = (unsigned)tcx0 < (unsigned)TC;
    const bool vx1 = ox != 0 && (unsigned)(tcx0 + 1) < (unsigned)TC;      // a column-aligned window lies in one tile column
    bool any = false;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const int y = wy0 + j;
      const bool vy = (unsigned)y < (unsigned)(4 * TR);                   // rows h2 .. 4*TR-1 are stored zeros
      const __half* src = plane + ((long)(y >> 2) * TC + tcx0) * 32 + (y & 3) * 8;
      const bool v0 = vy && vx0, v1 = vy && vx1;
      any |= v0 | v1;
      cp_async16_zfill(sbase + j * 32, v0 ? src : plane, v0 ? 16 : 0);
      cp_async16_zfill(sbase + j * 32 + 16, v1 ? src + 32 : plane, v1 ? 16 : 0);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    const float dx = __fsub_rn(x0, floorf(x0)), dy = __fsub_rn(y0, floorf(y0));
    const __half2 wA = __half2half2(__float2half_rn(__fmul_rn(__fsub_rn(1.0f, dx), __fsub_rn(1.0f, dy))));
    const __half2 wB = __half2half2(__float2half_rn(__fmul_rn(__fsub_rn(1.0f, dx), dy)));
    const __half2 wC = __half2half2(__float2half_rn(__fmul_rn(dx, __fsub_rn(1.0f, dy))));
    const __half2 wD = __half2half2(__float2half_rn(__fmul_rn(dx, dy)));
    const __half2 zero = __float2half2_rn(0.0f);
    __half* o = nullptr;
    if (LAYOUT == 0) {
      const long e = p / HW, pix = p - e * HW;
      o = out + (e * 196L + l * 49) * HW + pix;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (any) {            // (a window with no piece inside the plane -- incl. NaN / inf coordinates -- is all zeros, weights unused)
      uint32_t cur[4], nxt[4];
      slot_row(slot, 0, ox, cur);
#pragma unroll
      for (int b = 0; b < 7; b++) {
        slot_row(slot, b + 1, ox, nxt);
        uint32_t res[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const uint32_t qc = (k < 3) ? __funnelshift_r(cur[k], cur[k + 1], 16) : (cur[3] >> 16);
          const uint32_t qn = (k < 3) ? __funnelshift_r(nxt[k], nxt[k + 1], 16) : (nxt[3] >> 16);
          __half2 acc = __hadd2_rn(zero, __hmul2_rn(u2h(cur[k]), wA));
          acc = __hadd2_rn(acc, __hmul2_rn(u2h(nxt[k]), wB));
          acc = __hadd2_rn(acc, __hmul2_rn(u2h(qc), wC));
          acc = __hadd2_rn(acc, __hmul2_rn(u2h(qn), wD));
          res[k] = h2u_(acc);                   // outputs a = 2k (low), 2k+1 (high) of window row b; a = 7 is not an output
        }
        if (LAYOUT == 0) {
#pragma unroll
          for (int k = 0; k < 4; k++) {
            o[(long)((2 * k) * 7 + b) * HW] = __low2half(u2h(res[k]));
            if (k < 3) o[(long)((2 * k + 1) * 7 + b) * HW] = __high2half(u2h(res[k]));
          }
        } else {          // slot row b was consumed one iteration ago: it now holds this row's 7 outputs
          *reinterpret_cast<uint4*>(slot + b * 32) = make_uint4(res[0], res[1], res[2], res[3]);
        }
#pragma unroll
        for (int k = 0; k < 4; k++) cur[k] = nxt[k];
      }
    } else if (LAYOUT == 0) {
#pragma unroll 7
      for (int q = 0; q < 49; q++) o[(long)q * HW] = __float2half_rn(0.0f);
    } else {
#pragma unroll
      for (int b = 0; b < 7; b++) *reinterpret_cast<uint4*>(slot + b * 32) = make_uint4(0u, 0u, 0u, 0u);
    }
  }
  if (LAYOUT == 1) {
    __syncthreads();
    const int wpp = Cs / 2;                                        // 32-bit words per pixel record
    const long remaining = npix_total - (long)blockIdx.x * 32;
    const int npx = remaining < 32 ? (int)remaining : 32;
    uint32_t* g32 = reinterpret_cast<uint32_t*>(out) + (long)blockIdx.x * 32 * wpp;
    // lane <-> word of the pixel record (2 channels), warp <-> pixels l, l+4, ...: the table lookups are hoisted out of the pixel loop
    for (int w0 = 0; w0 < wpp; w0 += 32) {
      const int w = w0 + lane;
      const bool inw = w < wpp;
      const int c0 = 2 * w, c1 = 2 * w + 1;
      const bool r0 = inw && c0 < 196, r1 = inw && c1 < 196;
      const int off0 = r0 ? tab[c0] : 0, off1 = r1 ? tab[c1] : 0;
      for (int pp = l; pp < npx; pp += 4) {
        const unsigned char* base = smem_raw + pp * kSlot;
        const uint32_t lo = r0 ? *reinterpret_cast<const uint16_t*>(base + off0) : 0u;
        const uint32_t hi = r1 ? *reinterpret_cast<const uint16_t*>(base + off1) : 0u;
        if (inw) g32[(long)pp * wpp + w] = lo | (hi << 16);
      }
    }
  }
}

// tiled -> planar (API compatibility: CorrBlock.corr_pyramid in the reference's layout, on demand)
__global__ void untile_kernel(const __half* __restrict__ src, __half* __restrict__ dst, long nplanes, int h2, int w2, int TR, int TC) {
  const long total = nplanes * h2 * w2;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long pl = i / ((long)h2 * w2);
    const int rem = (int)(i - pl * ((long)h2 * w2));
    const int y = rem / w2, x = rem - y * w2;
    dst[i] = src[(pl * TR * TC + (long)(y >> 2) * TC + (x >> 3)) * 32 + (y & 3) * 8 + (x & 7)];
  }
}

// planar -> tiled (zero padded); used when a caller hands CorrBlock a planar pyramid and by the tests
__global__ void tile_kernel(const __half* __restrict__ src, __half* __restrict__ dst, long nplanes, int h2, int w2, int TR, int TC) {
  const long total = nplanes * TR * TC * 32;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long t = i >> 5;
    const int e = (int)(i & 31);
    const long pl = t / ((long)TR * TC);
    const int tt = (int)(t - pl * ((long)TR * TC));
    const int y = (tt / TC) * 4 + (e >> 3), x = (tt % TC) * 8 + (e & 7);
    dst[i] = (y < h2 && x < w2) ? src[pl * h2 * w2 + (long)y * w2 + x] : __float2half_rn(0.0f);
  }
}

}  // namespace

// levels[l]: (E, h*w, ceil(level_h[l]/4), ceil(level_w[l]/8), 32) float16 (see the header); coords (E,h,w,2) float32.
extern "C" int pvo_corr_lookup_tiled(const void* const* levels, const int* level_h, const int* level_w, int num_levels,
                                     const int* slots, const float* coords, void* out, int E, int h, int w, int radius, int layout,
                                     int out_cstride, void* stream) {
  if (num_levels != 4 || radius != 3) return PVO_ERR_UNSUPPORTED;
  if (E < 0 || h <= 0 || w <= 0) return PVO_ERR_INVALID;
  if (layout == 1 && (out_cstride < 196 || (out_cstride & 1))) return PVO_ERR_INVALID;
  if (E == 0) return PVO_OK;
  TLevels lv;
  lv.slots = slots;
  for (int i = 0; i < 4; i++) {
    lv.vol[i] = (const __half*)levels[i];
    lv.tr[i] = (level_h[i] + 3) / 4;
    lv.tc[i] = (level_w[i] + 7) / 8;
  }
  const long total = (long)E * h * w;
  const int grid = ceil_div(total, 32);
  cudaStream_t s = (cudaStream_t)stream;
  const size_t smem = 4 * kLevelBytes + 512;
  static bool set[2] = {false, false};
  if (!set[layout & 1]) {
    cudaError_t e = (layout == 1)
        ? cudaFuncSetAttribute(corr_lookup_tiled_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
        : cudaFuncSetAttribute(corr_lookup_tiled_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return PVO_ERR_CUDA;
    set[layout & 1] = true;
  }
  if (layout == 1) {
    corr_lookup_tiled_kernel<1><<<grid, 128, smem, s>>>(lv, (const float2*)coords, (__half*)out, total, h * w, out_cstride);
  } else {
    corr_lookup_tiled_kernel<0><<<grid, 128, smem, s>>>(lv, (const float2*)coords, (__half*)out, total, h * w, out_cstride);
  }
  return pvo_check_launch();
}

extern "C" int pvo_corr_untile(const void* tiled, void* planar, long nplanes, int h2, int w2, int w2pad, void* stream) {
  if (nplanes < 0 || h2 <= 0 || w2 <= 0 || w2pad < w2) return PVO_ERR_INVALID;
  if (nplanes == 0) return PVO_OK;
  untile_kernel<<<PVO_SM_COUNT * 8, 256, 0, (cudaStream_t)stream>>>((const __half*)tiled, (__half*)planar, nplanes, h2, w2, (h2 + 3) / 4,
                                                                    (w2pad + 7) / 8);
  return pvo_check_launch();
}

extern "C" int pvo_corr_tile(const void* planar, void* tiled, long nplanes, int h2, int w2, int w2pad, void* stream) {
  if (nplanes < 0 || h2 <= 0 || w2 <= 0 || w2pad < w2) return PVO_ERR_INVALID;
  if (nplanes == 0) return PVO_OK;
  tile_kernel<<<PVO_SM_COUNT * 8, 256, 0, (cudaStream_t)stream>>>((const __half*)planar, (__half*)tiled, nplanes, h2, w2, (h2 + 3) / 4,
                                                                  (w2pad + 7) / 8);
  return pvo_check_launch();
}
