// Remaining droid_backends correlation entry points: the lookup's backward pass and the on-the-fly ("alt")
// correlation used by the global-BA path.
//
//   corr_index_backward : VO_Module/src/correlation_kernels.cu:73-124,157-185  (training only)
//   altcorr_forward     : VO_Module/src/altcorr_kernel.cu:27-149,290-319      (FactorGraph.update_lowmem)
#include "common.cuh"

namespace {

__device__ __forceinline__ bool inb(int y, int x, int H, int W) { return y >= 0 && y < H && x >= 0 && x < W; }

template <typename T> struct Ar;
template <> struct Ar<float> {
  static __device__ __forceinline__ float w(float v) { return v; }
  static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
  static __device__ __forceinline__ float fma_(float a, float b, float c) { return __fmaf_rn(a, b, c); }
  static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
  static __device__ __forceinline__ float zero() { return 0.0f; }
};
template <> struct Ar<__half> {
  static __device__ __forceinline__ __half w(float v) { return __float2half_rn(v); }
  static __device__ __forceinline__ __half fma_(__half a, __half b, __half c) { return __hadd_rn(c, __hmul_rn(a, b)); }
  static __device__ __forceinline__ __half add(__half a, __half b) { return __hadd_rn(a, b); }
  static __device__ __forceinline__ __half zero() { return __float2half_rn(0.0f); }
};

// one thread per source pixel owns its (h2,w2) slice of volume_grad: g accumulates the <= 4 contributions in the
// reference's order, then one += into the (zero-initialised) slice -- same arithmetic as the reference kernel.
template <typename T>
__global__ void __launch_bounds__(128) corr_index_bwd_kernel(const float* __restrict__ coords, const T* __restrict__ cg,
                                                             T* __restrict__ vg, int N, int h1, int w1, int h2, int w2, int r) {
  const long plane = (long)h1 * w1;
  const long p = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= (long)N * plane) return;
  const int n = (int)(p / plane);
  const long pix = p - (long)n * plane;
  const float x0 = coords[((long)n * 2 + 0) * plane + pix], y0 = coords[((long)n * 2 + 1) * plane + pix];
  const float dx = x0 - floorf(x0), dy = y0 - floorf(y0);
  const int fx = (int)fminf(fmaxf(floorf(x0), -16777216.f), 16777216.f), fy = (int)fminf(fmaxf(floorf(y0), -16777216.f), 16777216.f);
  const int rd = 2 * r + 1;
  const T* g = cg + (long)n * rd * rd * plane + pix;
  T* v = vg + p * (long)h2 * w2;
  const auto wDD = Ar<T>::w(dx * dy), wDN = Ar<T>::w(dx * (1.0f - dy)), wND = Ar<T>::w((1.0f - dx) * dy),
             wNN = Ar<T>::w((1.0f - dx) * (1.0f - dy));
  for (int i = 0; i < rd + 1; i++)
    for (int j = 0; j < rd + 1; j++) {
      const int x1 = fx - r + i, y1 = fy - r + j;
      if (!inb(y1, x1, h2, w2)) continue;
      T acc = Ar<T>::zero();
      if (i > 0 && j > 0) acc = Ar<T>::fma_(g[(long)((i - 1) * rd + (j - 1)) * plane], wDD, acc);
      if (i > 0 && j < rd) acc = Ar<T>::fma_(g[(long)((i - 1) * rd + j) * plane], wDN, acc);
      if (i < rd && j > 0) acc = Ar<T>::fma_(g[(long)(i * rd + (j - 1)) * plane], wND, acc);
      if (i < rd && j < rd) acc = Ar<T>::fma_(g[(long)(i * rd + j) * plane], wNN, acc);
      v[(long)y1 * w2 + x1] = Ar<T>::add(v[(long)y1 * w2 + x1], acc);
    }
}

// altcorr forward (AltCorrBlock, the global-BA path `update_lowmem`): the 8x8 dot products <f1[pixel], f2[tap]> of every pixel, blended
// bilinearly into the 7x7 outputs (VO_Module/src/altcorr_kernel.cu:27-149).  The reference (and round 1 here) gives every pixel to one
// thread, which walks 64 taps x C channels with row loads that no two lanes share.  Here a WARP owns a pixel: the lanes split the
// channels (one float4 each at C = 128), so every tap is one coalesced 512-byte row read; the 64 per-lane partial sums are
// transpose-reduced across the warp (62 shuffles instead of 64 x 5) so that lane L ends with taps 2L, 2L+1, which go through 256 B of
// shared memory to the lanes that blend the 49 outputs.  A block is an 8x8 pixel tile of one (batch, sample) plane: neighbouring
// pixels' windows overlap, so most tap rows of a tile are L1 hits.  float32 throughout (the reference forces .float(), corr.py:120).
template <int WW, int BIT>
__device__ __forceinline__ void xr64_step(float* x, int lane) {
  const bool up = lane & BIT;
#pragma unroll
  for (int j = 0; j < WW; j++) {
    const float send = up ? x[j] : x[j + WW];
    const float recv = __shfl_xor_sync(0xffffffffu, send, BIT);
    x[j] = (up ? x[j + WW] : x[j]) + recv;
  }
}

__global__ void __launch_bounds__(256, 2) altcorr_fwd_kernel(const float* __restrict__ f1, const float* __restrict__ f2,
                                                          const float* __restrict__ coords, float* __restrict__ out,
                                                          int B, int S, int H1, int W1, int H2, int W2, int C, int tiles_x, int tiles_y) {
  __shared__ float dots[8][64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long plane = (long)H1 * W1;
  const int tile = blockIdx.x % (tiles_x * tiles_y);
  const long bs = blockIdx.x / (tiles_x * tiles_y);
  const int b = (int)(bs / S);
  const int y = (tile / tiles_x) * 8 + warp;
  if (y >= H1) return;
  for (int k = 0; k < 8; k++) {
    const int x = (tile % tiles_x) * 8 + k;
    if (x >= W1) break;                                                  // warp-uniform
    const long pix = (long)y * W1 + x;
    const long t = bs * plane + pix;
    const float x2 = coords[t * 2], y2 = coords[t * 2 + 1];
    const float dx = x2 - floorf(x2), dy = y2 - floorf(y2);
    const int fx = (int)fminf(fmaxf(floorf(x2), -16777216.f), 16777216.f), fy = (int)fminf(fmaxf(floorf(y2), -16777216.f), 16777216.f);
    const float* a = f1 + ((long)b * plane + pix) * C;
#pragma unroll 1
    for (int part = 0; part < 4; part++) {                                 // two window rows (16 taps) at a time: 16 partial sums per lane
      float acc[16];
#pragma unroll
      for (int q = 0; q < 16; q++) acc[q] = 0.0f;
      for (int c0 = lane * 4; c0 < C; c0 += 128) {
        const float4 u = __ldg(reinterpret_cast<const float4*>(a + c0));
#pragma unroll
        for (int iy = 0; iy < 2; iy++) {
#pragma unroll
          for (int ix = 0; ix < 8; ix++) {
            const int hh = fy - 3 + part * 2 + iy, ww = fx - 3 + ix;
            const bool v_ok = inb(hh, ww, H2, W2);                         // warp-uniform
            const float4 v = v_ok ? __ldg(reinterpret_cast<const float4*>(f2 + (((long)b * H2 + hh) * W2 + ww) * C + c0))
                                  : make_float4(0.f, 0.f, 0.f, 0.f);
            float s = acc[iy * 8 + ix];
            s = fmaf(u.x, v.x, s); s = fmaf(u.y, v.y, s); s = fmaf(u.z, v.z, s); s = fmaf(u.w, v.w, s);
            acc[iy * 8 + ix] = s;
          }
        }
      }
      xr64_step<8, 16>(acc, lane);
      xr64_step<4, 8>(acc, lane);
      xr64_step<2, 4>(acc, lane);
      xr64_step<1, 2>(acc, lane);
      acc[0] += __shfl_xor_sync(0xffffffffu, acc[0], 1);                   // lanes 2k, 2k+1 now both hold tap k of this part
      if (!(lane & 1)) dots[warp][part * 16 + (lane >> 1)] = acc[0];       // tap = iy * 8 + ix, fully summed over the channels
    }
    __syncwarp();
    float* o = out + bs * 49 * plane + pix;
    for (int q = lane; q < 49; q += 32) {
      const int ox = q / 7, oy = q - ox * 7;                                               // channel = y-offset + 7 * x-offset
      const float* d = &dots[warp][oy * 8 + ox];
      float r = d[0] * ((1 - dy) * (1 - dx));
      r += d[1] * ((1 - dy) * dx);
      r += d[8] * (dy * (1 - dx));
      r += d[9] * (dy * dx);
      o[(long)q * plane] = r;
    }
    __syncwarp();                                                          // dots[warp] is rewritten by the next pixel
  }
}

// altcorr backward (training path): one warp per (b, source pixel), lane <-> 4 channels.  For every sample n and tap (iy, ix) of the
// 8x8 support: g = sum of the <= 4 output gradients the tap contributes to, each times its bilinear weight
// (altcorr_kernel.cu:233-249; channel = y-offset + 7 * x-offset); fmap1_grad[pixel] += g * f2[tap] (kept in registers over all taps
// and samples, written once), fmap2_grad[tap] += g * f1[pixel] (atomics, as the reference).  coords get no gradient (:340 zeros).
__global__ void __launch_bounds__(128) altcorr_bwd_kernel(const float* __restrict__ f1, const float* __restrict__ f2,
                                                          const float* __restrict__ coords, const float* __restrict__ cg,
                                                          float* __restrict__ g1, float* __restrict__ g2, int B, int S, int H1, int W1,
                                                          int H2, int W2, int C) {
  const long plane = (long)H1 * W1;
  const long wid = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (wid >= (long)B * plane) return;
  const int b = (int)(wid / plane);
  const long pix = wid - (long)b * plane;
  const float* a = f1 + ((long)b * plane + pix) * C;
  for (int c0 = lane * 4; c0 < C; c0 += 128) {
    const float4 u = *reinterpret_cast<const float4*>(a + c0);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int n = 0; n < S; n++) {
      const long t = ((long)b * S + n) * plane + pix;
      const float x2 = coords[t * 2], y2 = coords[t * 2 + 1];
      const float dx = x2 - floorf(x2), dy = y2 - floorf(y2);
      const int fx = (int)fminf(fmaxf(floorf(x2), -16777216.f), 16777216.f), fy = (int)fminf(fmaxf(floorf(y2), -16777216.f), 16777216.f);
      const float* gp = cg + ((long)b * S + n) * 49 * plane + pix;
      for (int iy = 0; iy < 8; iy++)
        for (int ix = 0; ix < 8; ix++) {
          const int hh = fy - 3 + iy, ww = fx - 3 + ix;
          if (!inb(hh, ww, H2, W2)) continue;          // the reference's f2 = 0 and no atomicAdd for such taps
          float g = 0.0f;
          if (iy > 0 && ix > 0) g += gp[(long)((iy - 1) + 7 * (ix - 1)) * plane] * dy * dx;
          if (iy > 0 && ix < 7) g += gp[(long)((iy - 1) + 7 * ix) * plane] * dy * (1 - dx);
          if (iy < 7 && ix > 0) g += gp[(long)(iy + 7 * (ix - 1)) * plane] * (1 - dy) * dx;
          if (iy < 7 && ix < 7) g += gp[(long)(iy + 7 * ix) * plane] * (1 - dy) * (1 - dx);
          const long off = (((long)b * H2 + hh) * W2 + ww) * C + c0;
          const float4 v = *reinterpret_cast<const float4*>(f2 + off);
          acc.x = fmaf(g, v.x, acc.x); acc.y = fmaf(g, v.y, acc.y); acc.z = fmaf(g, v.z, acc.z); acc.w = fmaf(g, v.w, acc.w);
          atomicAdd(g2 + off, g * u.x); atomicAdd(g2 + off + 1, g * u.y); atomicAdd(g2 + off + 2, g * u.z); atomicAdd(g2 + off + 3, g * u.w);
        }
    }
    *reinterpret_cast<float4*>(g1 + ((long)b * plane + pix) * C + c0) = acc;
  }
}

}  // namespace

extern "C" int pvo_version(void) { return 101; }

extern "C" int pvo_altcorr_backward(const float* fmap1, const float* fmap2, const float* coords, const float* corr_grad,
                                    float* fmap1_grad, float* fmap2_grad, int B, int S, int H1, int W1, int H2, int W2, int C,
                                    int radius, void* stream) {
  if (B < 0 || S < 0 || C % 4) return PVO_ERR_INVALID;
  if (radius != 3) return PVO_ERR_UNSUPPORTED;
  if (B == 0) return PVO_OK;
  const long warps = (long)B * H1 * W1;
  altcorr_bwd_kernel<<<ceil_div(warps * 32, 128), 128, 0, (cudaStream_t)stream>>>(fmap1, fmap2, coords, corr_grad, fmap1_grad, fmap2_grad, B,
                                                                                   S, H1, W1, H2, W2, C);
  return pvo_check_launch();
}

extern "C" int pvo_corr_index_backward(const float* coords, const void* corr_grad, void* volume_grad, int N, int h1,
                                       int w1, int h2, int w2, int radius, int dtype, void* stream) {
  if (N < 0 || h1 <= 0 || w1 <= 0 || h2 <= 0 || w2 <= 0 || radius < 0) return PVO_ERR_INVALID;
  if (N == 0) return PVO_OK;
  const int grid = ceil_div((long)N * h1 * w1, 128);
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == PVO_DTYPE_F32)
    corr_index_bwd_kernel<float><<<grid, 128, 0, s>>>(coords, (const float*)corr_grad, (float*)volume_grad, N, h1, w1, h2, w2, radius);
  else if (dtype == PVO_DTYPE_F16)
    corr_index_bwd_kernel<__half><<<grid, 128, 0, s>>>(coords, (const __half*)corr_grad, (__half*)volume_grad, N, h1, w1, h2, w2, radius);
  else
    return PVO_ERR_UNSUPPORTED;
  return pvo_check_launch();
}

extern "C" int pvo_altcorr_forward(const float* fmap1, const float* fmap2, const float* coords, float* out, int B, int S,
                                   int H1, int W1, int H2, int W2, int C, int radius, void* stream) {
  if (B < 0 || S < 0 || C % 4) return PVO_ERR_INVALID;
  if (radius != 3) return PVO_ERR_UNSUPPORTED;
  if (B == 0 || S == 0) return PVO_OK;
  const int tx = (W1 + 7) / 8, ty = (H1 + 7) / 8;
  const long blocks = (long)B * S * tx * ty;
  if (blocks > 0x7fffffffL) return PVO_ERR_UNSUPPORTED;
  altcorr_fwd_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(fmap1, fmap2, coords, out, B, S, H1, W1, H2, W2, C, tx, ty);
  return pvo_check_launch();
}
