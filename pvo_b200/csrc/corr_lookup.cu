// Correlation-volume window lookup (bilinear (2r+1)^2 window per pixel per pyramid level).
//
// Replaces, bit for bit, the reference's corr_index_forward_kernel
// (VO_Module/src/correlation_kernels.cu:19-70, host wrapper :126-155) and fuses the 4-level loop + torch.cat of
// CorrBlock.__call__ (VO_Module/droid_slam/modules/corr.py:40-50).
//
// Reference algorithm: one thread per pixel walks the (2r+2)^2 taps (x-offset i outer, y-offset j inner) and
// does up to four read-modify-writes per tap into a zero-initialised global output.  Per OUTPUT (a,b) that is
//   acc = 0; acc += s(a,b)*w[(1-dx)(1-dy)]; acc += s(a,b+1)*w[(1-dx)dy]; acc += s(a+1,b)*w[dx(1-dy)];
//   acc += s(a+1,b+1)*w[dx dy]          (out-of-bounds taps are SKIPPED, not added as zero)
// with, for float, each `acc += s*w` contracted to one FMA by nvcc (-fmad=true) and, for c10::Half, the
// weight rounded to half, the product rounded to half and the sum rounded to half.  Half products/sums of two
// halves computed via float and rounded once equal the native HMUL/HADD results (the float intermediate is
// exact or far from a rounding tie), so the half path below uses non-fused __hmul2_rn/__hadd2_rn.
//
// B200 design: each output is a pure gather -> computed in registers and written exactly once (no zero-fill,
// no RMW).  The 8 taps of a window row are contiguous in the volume's last dim: the half path fetches a row
// with two aligned 16-byte loads + a funnel-shift realign instead of 8 two-byte loads (8x fewer L1
// wavefronts; the kernel is L1/L2-gather bound, not FLOP bound).  All pyramid levels of all edges go in ONE
// launch; output is written either in the reference's planar (E, L*49, h, w) layout or pixel-major
// (E, h*w, Cs) for the tensor-core encoder that consumes it (smem-staged, fully coalesced 128-B stores).
#include "common.cuh"

namespace {

struct Levels {
  const void* vol[8];
  int h2[8];
  int w2[8];
};

// ---------------------------------------------------------------------------------------------------
// generic (any radius, any position): per-output 4-tap gather in the reference's accumulation order
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool inb(int y, int x, int H, int W) { return y >= 0 && y < H && x >= 0 && x < W; }

__device__ __forceinline__ float out_generic(const float* v, int h2, int w2, int x1, int y1,
                                             float wA, float wB, float wC, float wD) {
  float acc = 0.0f;
  if (inb(y1, x1, h2, w2)) acc = __fmaf_rn(__ldg(v + (long)y1 * w2 + x1), wA, acc);
  if (inb(y1 + 1, x1, h2, w2)) acc = __fmaf_rn(__ldg(v + (long)(y1 + 1) * w2 + x1), wB, acc);
  if (inb(y1, x1 + 1, h2, w2)) acc = __fmaf_rn(__ldg(v + (long)y1 * w2 + x1 + 1), wC, acc);
  if (inb(y1 + 1, x1 + 1, h2, w2)) acc = __fmaf_rn(__ldg(v + (long)(y1 + 1) * w2 + x1 + 1), wD, acc);
  return acc;
}

__device__ __forceinline__ __half out_generic(const __half* v, int h2, int w2, int x1, int y1,
                                              __half wA, __half wB, __half wC, __half wD) {
  __half acc = __float2half_rn(0.0f);
  if (inb(y1, x1, h2, w2)) acc = __hadd_rn(acc, __hmul_rn(__ldg(v + (long)y1 * w2 + x1), wA));
  if (inb(y1 + 1, x1, h2, w2)) acc = __hadd_rn(acc, __hmul_rn(__ldg(v + (long)(y1 + 1) * w2 + x1), wB));
  if (inb(y1, x1 + 1, h2, w2)) acc = __hadd_rn(acc, __hmul_rn(__ldg(v + (long)y1 * w2 + x1 + 1), wC));
  if (inb(y1 + 1, x1 + 1, h2, w2)) acc = __hadd_rn(acc, __hmul_rn(__ldg(v + (long)(y1 + 1) * w2 + x1 + 1), wD));
  return acc;
}

template <typename T> struct Wt;
template <> struct Wt<float> { static __device__ __forceinline__ float cvt(float w) { return w; } };
template <> struct Wt<__half> { static __device__ __forceinline__ __half cvt(float w) { return __float2half_rn(w); } };

__device__ __forceinline__ int safe_floor(float x) {
  // floorf then saturate: any coordinate beyond +-2^24 has every tap out of bounds -> zero output either way
  float f = floorf(x);
  f = fminf(fmaxf(f, -16777216.0f), 16777216.0f);
  return (f == f) ? (int)f : -16777216;
}

// ---------------------------------------------------------------------------------------------------
// radius-3 fast paths (window fully inside the level): rows fetched whole, outputs in registers
// ---------------------------------------------------------------------------------------------------
// 8 consecutive halves starting at p (2-byte aligned) -> 4 packed words, via two aligned 16-B loads.
__device__ __forceinline__ void load_row8(const __half* p, uint32_t r[4]) {
  const uintptr_t a = reinterpret_cast<uintptr_t>(p);
  const uint32_t o = (uint32_t)(a & 15u);
  const uint4* q = reinterpret_cast<const uint4*>(a - o);
  const uint4 A = __ldg(q);
  uint4 B = A;
  if (o) B = __ldg(q + 1);
  const uint32_t ws = o >> 2, bs = (o & 3u) * 8u;
  const bool s2 = ws & 2u, s1 = ws & 1u;
  const uint32_t d0 = s2 ? A.z : A.x, d1 = s2 ? A.w : A.y, d2 = s2 ? B.x : A.z, d3 = s2 ? B.y : A.w,
                 d4 = s2 ? B.z : B.x, d5 = s2 ? B.w : B.y;
  const uint32_t e0 = s1 ? d1 : d0, e1 = s1 ? d2 : d1, e2 = s1 ? d3 : d2, e3 = s1 ? d4 : d3, e4 = s1 ? d5 : d4;
  r[0] = __funnelshift_r(e0, e1, bs);
  r[1] = __funnelshift_r(e1, e2, bs);
  r[2] = __funnelshift_r(e2, e3, bs);
  r[3] = __funnelshift_r(e3, e4, bs);
}

__device__ __forceinline__ __half2 as_h2(uint32_t u) { return *reinterpret_cast<__half2*>(&u); }

template <typename Store>
__device__ __forceinline__ void window_r3_fast(const __half* v, int w2, int fx, int fy, float dx, float dy,
                                               bool scalar_loads, Store st) {
  const __half2 wA = __half2half2(__float2half_rn(__fmul_rn(__fsub_rn(1.0f, dx), __fsub_rn(1.0f, dy))));
  const __half2 wB = __half2half2(__float2half_rn(__fmul_rn(__fsub_rn(1.0f, dx), dy)));
  const __half2 wC = __half2half2(__float2half_rn(__fmul_rn(dx, __fsub_rn(1.0f, dy))));
  const __half2 wD = __half2half2(__float2half_rn(__fmul_rn(dx, dy)));
  const __half2 zero = __float2half2_rn(0.0f);
  const __half* row = v + (long)(fy - 3) * w2 + (fx - 3);
  uint32_t cur[4], nxt[4];
  auto fetch = [&](const __half* p, uint32_t r[4]) {
    if (scalar_loads) {
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const uint32_t lo = __half_as_ushort(__ldg(p + 2 * k)), hi = __half_as_ushort(__ldg(p + 2 * k + 1));
        r[k] = lo | (hi << 16);
      }
    } else {
      load_row8(p, r);
    }
  };
  fetch(row, cur);
#pragma unroll
  for (int b = 0; b < 7; b++) {
    fetch(row + (long)(b + 1) * w2, nxt);
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const uint32_t qc = (k < 3) ? __funnelshift_r(cur[k], cur[k + 1], 16) : (cur[3] >> 16);
      const uint32_t qn = (k < 3) ? __funnelshift_r(nxt[k], nxt[k + 1], 16) : (nxt[3] >> 16);
      __half2 acc = __hadd2_rn(zero, __hmul2_rn(as_h2(cur[k]), wA));
      acc = __hadd2_rn(acc, __hmul2_rn(as_h2(nxt[k]), wB));
      acc = __hadd2_rn(acc, __hmul2_rn(as_h2(qc), wC));
      acc = __hadd2_rn(acc, __hmul2_rn(as_h2(qn), wD));
      st(2 * k, b, __low2half(acc));
      if (k < 3) st(2 * k + 1, b, __high2half(acc));
    }
#pragma unroll
    for (int k = 0; k < 4; k++) cur[k] = nxt[k];
  }
}

template <typename Store>
__device__ __forceinline__ void window_r3_fast(const float* v, int w2, int fx, int fy, float dx, float dy,
                                               bool, Store st) {
  const float wA = __fmul_rn(__fsub_rn(1.0f, dx), __fsub_rn(1.0f, dy));
  const float wB = __fmul_rn(__fsub_rn(1.0f, dx), dy);
  const float wC = __fmul_rn(dx, __fsub_rn(1.0f, dy));
  const float wD = __fmul_rn(dx, dy);
  const float* row = v + (long)(fy - 3) * w2 + (fx - 3);
  float cur[8], nxt[8];
#pragma unroll
  for (int k = 0; k < 8; k++) cur[k] = __ldg(row + k);
#pragma unroll
  for (int b = 0; b < 7; b++) {
#pragma unroll
    for (int k = 0; k < 8; k++) nxt[k] = __ldg(row + (long)(b + 1) * w2 + k);
#pragma unroll
    for (int a = 0; a < 7; a++) {
      float acc = __fmaf_rn(cur[a], wA, 0.0f);
      acc = __fmaf_rn(nxt[a], wB, acc);
      acc = __fmaf_rn(cur[a + 1], wC, acc);
      acc = __fmaf_rn(nxt[a + 1], wD, acc);
      st(a, b, acc);
    }
#pragma unroll
    for (int k = 0; k < 8; k++) cur[k] = nxt[k];
  }
}

template <typename T, typename Store>
__device__ __forceinline__ void window_any(const T* v, int h2, int w2, float x0, float y0, int r,
                                           bool scalar_loads, Store st) {
  const int fx = safe_floor(x0), fy = safe_floor(y0);
  const float dx = __fsub_rn(x0, floorf(x0)), dy = __fsub_rn(y0, floorf(y0));
  if (r == 3 && fx - 3 >= 0 && fy - 3 >= 0 && fx + 4 < w2 && fy + 4 < h2) {
    window_r3_fast(v, w2, fx, fy, dx, dy, scalar_loads, st);
    return;
  }
  const auto wA = Wt<T>::cvt(__fmul_rn(__fsub_rn(1.0f, dx), __fsub_rn(1.0f, dy)));
  const auto wB = Wt<T>::cvt(__fmul_rn(__fsub_rn(1.0f, dx), dy));
  const auto wC = Wt<T>::cvt(__fmul_rn(dx, __fsub_rn(1.0f, dy)));
  const auto wD = Wt<T>::cvt(__fmul_rn(dx, dy));
  const int rd = 2 * r + 1;
  for (int b = 0; b < rd; b++)
    for (int a = 0; a < rd; a++) st(a, b, out_generic(v, h2, w2, fx - r + a, fy - r + b, wA, wB, wC, wD));
}

// ---------------------------------------------------------------------------------------------------
// kernel 1: droid_backends.corr_index_forward -- one level, planar coords (N,2,h1,w1), out (N,rd,rd,h1,w1)
// ---------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(128) corr_index_fwd_kernel(const T* __restrict__ vol, const float* __restrict__ coords,
                                                             T* __restrict__ out, int N, int h1, int w1, int h2, int w2,
                                                             int r, int scalar_loads) {
  const long plane = (long)h1 * w1;
  const long p = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= (long)N * plane) return;
  const int n = (int)(p / plane);
  const long pix = p - (long)n * plane;
  const float x0 = __ldg(coords + ((long)n * 2 + 0) * plane + pix);
  const float y0 = __ldg(coords + ((long)n * 2 + 1) * plane + pix);
  const int rd = 2 * r + 1;
  const T* v = vol + p * (long)h2 * w2;
  T* o = out + (long)n * rd * rd * plane + pix;
  window_any<T>(v, h2, w2, x0, y0, r, scalar_loads != 0,
                [&](int a, int b, T val) { o[(long)(a * rd + b) * plane] = val; });
}

// ---------------------------------------------------------------------------------------------------
// kernel 2: fused CorrBlock.__call__ -- L levels, interleaved coords (E,h,w,2), radius 3
//   LAYOUT 0: planar  out[e][l*49 + a*7 + b][pix]                 (reference layout)
//   LAYOUT 1: pixel-major out[e*HW + pix][Cs], channels l*49+a*7+b, [L*49, Cs) zero-filled
// block = 32 pixels x L levels (warp <-> level)
// ---------------------------------------------------------------------------------------------------
template <typename T, int LAYOUT>
__global__ void __launch_bounds__(256) corr_lookup_fused_kernel(const __grid_constant__ Levels lv, const float2* __restrict__ coords,
                                                                T* __restrict__ out, long npix_total, int HW, int L,
                                                                int Cs, int scalar_loads) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, l = threadIdx.x >> 5;
  const long p = (long)blockIdx.x * 32 + lane;
  const bool active = p < npix_total;
  const int sstride = (LAYOUT == 1) ? (int)(((Cs * sizeof(T) / 4) | 1) * (4 / sizeof(T))) : 0;  // odd #words
  T* sm = reinterpret_cast<T*>(smem_raw);
  if (LAYOUT == 1) {
    // zero the pad channels of this tile once (they sit past the last level's 49 outputs)
    for (int i = threadIdx.x; i < 32 * (Cs - L * 49); i += blockDim.x) {
      const int pp = i / (Cs - L * 49), c = L * 49 + i % (Cs - L * 49);
      sm[pp * sstride + c] = T(0.0f);
    }
  }
  if (active) {
    const float2 c = __ldg(coords + p);
    const float sc = 1.0f / (float)(1 << l);                 // coords / 2**l is exact in float
    const float x0 = c.x * sc, y0 = c.y * sc;
    const int h2 = lv.h2[l], w2 = lv.w2[l];
    const T* v = reinterpret_cast<const T*>(lv.vol[l]) + p * (long)h2 * w2;
    if (LAYOUT == 0) {
      const long e = p / HW;
      const long pix = p - e * HW;
      T* o = out + (e * (long)(L * 49) + l * 49) * HW + pix;
      window_any<T>(v, h2, w2, x0, y0, 3, scalar_loads != 0,
                    [&](int a, int b, T val) { o[(long)(a * 7 + b) * HW] = val; });
    } else {
      T* o = sm + lane * sstride + l * 49;
      window_any<T>(v, h2, w2, x0, y0, 3, scalar_loads != 0, [&](int a, int b, T val) { o[a * 7 + b] = val; });
    }
  }
  if (LAYOUT == 1) {
    __syncthreads();
    const int wpp = (int)(Cs * sizeof(T) / 4);               // 32-bit words per pixel record
    const int swpp = (int)(sstride * sizeof(T) / 4);
    const uint32_t* s32 = reinterpret_cast<const uint32_t*>(smem_raw);
    uint32_t* g32 = reinterpret_cast<uint32_t*>(out) + (long)blockIdx.x * 32 * wpp;
    const long remaining = npix_total - (long)blockIdx.x * 32;
    const int npx = remaining < 32 ? (int)remaining : 32;
    for (int i = threadIdx.x; i < npx * wpp; i += blockDim.x) {
      const int pp = i / wpp, w = i - pp * wpp;
      g32[i] = s32[pp * swpp + w];
    }
  }
}

}  // namespace

// ===================================================================================================
// C ABI
// ===================================================================================================
extern "C" int pvo_corr_index_forward(const void* volume, const float* coords, void* out, int N, int h1, int w1,
                                      int h2, int w2, int radius, int dtype, int flags, void* stream) {
  if (N < 0 || h1 <= 0 || w1 <= 0 || h2 <= 0 || w2 <= 0 || radius < 0) return PVO_ERR_INVALID;
  if (N == 0) return PVO_OK;
  const long total = (long)N * h1 * w1;
  const int grid = ceil_div(total, 128);
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == PVO_DTYPE_F32)
    corr_index_fwd_kernel<float><<<grid, 128, 0, s>>>((const float*)volume, coords, (float*)out, N, h1, w1, h2, w2,
                                                       radius, flags & 1);
  else if (dtype == PVO_DTYPE_F16)
    corr_index_fwd_kernel<__half><<<grid, 128, 0, s>>>((const __half*)volume, coords, (__half*)out, N, h1, w1, h2, w2,
                                                        radius, flags & 1);
  else
    return PVO_ERR_UNSUPPORTED;
  return pvo_check_launch();
}

extern "C" int pvo_corr_lookup_pyramid(const void* const* levels, const int* level_h, const int* level_w,
                                       int num_levels, const float* coords, void* out, int E, int h, int w,
                                       int radius, int dtype, int layout, int out_cstride, int flags, void* stream) {
  if (num_levels < 1 || num_levels > 8 || E < 0 || h <= 0 || w <= 0) return PVO_ERR_INVALID;
  if (radius != 3) return PVO_ERR_UNSUPPORTED;
  if (E == 0) return PVO_OK;
  Levels lv;
  for (int i = 0; i < num_levels; i++) {
    lv.vol[i] = levels[i];
    lv.h2[i] = level_h[i];
    lv.w2[i] = level_w[i];
  }
  const long total = (long)E * h * w;
  const int grid = ceil_div(total, 32);
  const int block = 32 * num_levels;
  cudaStream_t s = (cudaStream_t)stream;
  const size_t esz = dtype == PVO_DTYPE_F32 ? 4 : 2;
  if (layout == 1) {
    if (out_cstride < num_levels * 49 || (out_cstride * esz) % 4) return PVO_ERR_INVALID;
  }
  const size_t smem = layout == 1 ? 32 * ((((size_t)out_cstride * esz / 4) | 1) * 4) : 0;
#define LAUNCH(T, LAY)                                                                                          \
  corr_lookup_fused_kernel<T, LAY><<<grid, block, smem, s>>>(lv, (const float2*)coords, (T*)out, total, h * w, \
                                                             num_levels, out_cstride, flags & 1)
  if (dtype == PVO_DTYPE_F32) {
    if (layout == 0) LAUNCH(float, 0); else LAUNCH(float, 1);
  } else if (dtype == PVO_DTYPE_F16) {
    if (layout == 0) LAUNCH(__half, 0); else LAUNCH(__half, 1);
  } else {
    return PVO_ERR_UNSUPPORTED;
  }
#undef LAUNCH
  return pvo_check_launch();
}
