// 8x upsampling of the 1/8-resolution fields (depth, flow, masks) to image resolution.
//
//   pvo_cvx_upsample    droid_net.py:23-37  cvx_upsample: every output pixel of the 8x8 block of a coarse pixel is a convex combination
//                       (softmax over 9 logits of the 576-channel mask) of the 3x3 coarse neighbourhood (zero padded, F.unfold).
//                       The reference materialises softmax(mask) and unfold(data) (2 x 9 x 64 values per coarse pixel) and multiplies
//                       them; here one block per coarse row segment stages the 576 x 32 logits through shared memory (coalesced
//                       reads along x), and each thread produces one output pixel (all `dim` channels), written coalesced.
//   pvo_upsample_inter  droid_net.py:54-61 / droid.py:106 / factor_graph.py:431  F.interpolate(scale_factor=8, bilinear,
//                       align_corners=True) on a channels-last field.
#include "common.cuh"

namespace {

constexpr int kXT = 16;            // coarse pixels per block along x

template <typename TM>
__global__ void __launch_bounds__(256) cvx_upsample_kernel(const float* __restrict__ data, const TM* __restrict__ mask,
                                                           float* __restrict__ out, int N, int ht, int wd, int dim) {
  __shared__ float sm[576][kXT + 1];
  const int n = blockIdx.z, y = blockIdx.y, x0 = blockIdx.x * kXT;
  const long plane = (long)ht * wd;
  for (int i = threadIdx.x; i < 576 * kXT; i += blockDim.x) {
    const int c = i / kXT, xx = i - c * kXT;
    sm[c][xx] = (x0 + xx < wd) ? (float)mask[((long)n * 576 + c) * plane + (long)y * wd + x0 + xx] : 0.0f;
  }
  __syncthreads();
  // 64 sub-pixels x kXT coarse pixels per block; thread order (sy, xx, sx) so that a warp writes 32 consecutive output pixels
  for (int i = threadIdx.x; i < 64 * kXT; i += blockDim.x) {
    const int sx = i & 7, xx = (i >> 3) % kXT, sy = i / (8 * kXT);
    const int x = x0 + xx;
    if (x >= wd) continue;
    float w[9], mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < 9; k++) { w[k] = sm[k * 64 + sy * 8 + sx][xx]; mx = fmaxf(mx, w[k]); }
    float den = 0.0f;
#pragma unroll
    for (int k = 0; k < 9; k++) { w[k] = expf(w[k] - mx); den += w[k]; }
    float* o = out + (((long)n * 8 * ht + y * 8 + sy) * 8 * wd + x * 8 + sx) * dim;
    for (int d = 0; d < dim; d++) {
      float acc = 0.0f;
#pragma unroll
      for (int k = 0; k < 9; k++) {
        const int yy = y + k / 3 - 1, xq = x + k % 3 - 1;
        const float v = (yy >= 0 && yy < ht && xq >= 0 && xq < wd) ? data[(((long)n * ht + yy) * wd + xq) * dim + d] : 0.0f;
        acc += (w[k] / den) * v;
      }
      o[d] = acc;
    }
  }
}

__global__ void __launch_bounds__(256) upsample_inter_kernel(const float* __restrict__ in, float* __restrict__ out, int N, int ht, int wd,
                                                             int dim) {
  const int Ho = 8 * ht, Wo = 8 * wd;
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)N * Ho * Wo) return;
  const int xo = (int)(i % Wo);
  const long t = i / Wo;
  const int yo = (int)(t % Ho);
  const long n = t / Ho;
  // align_corners=True: src = dst * (in - 1) / (out - 1)
  const float ry = Ho > 1 ? (float)(ht - 1) / (float)(Ho - 1) : 0.0f, rx = Wo > 1 ? (float)(wd - 1) / (float)(Wo - 1) : 0.0f;
  const float sy = ry * yo, sx = rx * xo;
  const int y0 = (int)sy, x0 = (int)sx;
  const int y1 = min(y0 + 1, ht - 1), x1 = min(x0 + 1, wd - 1);
  const float ly = sy - y0, lx = sx - x0;
  const float* b = in + n * (long)ht * wd * dim;
  float* o = out + i * dim;
  for (int d = 0; d < dim; d++) {
    const float v00 = b[((long)y0 * wd + x0) * dim + d], v01 = b[((long)y0 * wd + x1) * dim + d];
    const float v10 = b[((long)y1 * wd + x0) * dim + d], v11 = b[((long)y1 * wd + x1) * dim + d];
    o[d] = (1.f - ly) * ((1.f - lx) * v00 + lx * v01) + ly * ((1.f - lx) * v10 + lx * v11);
  }
}

}  // namespace

// data (N,ht,wd,dim) float32; mask (N,576,ht,wd) float32 | float16; out (N,8ht,8wd,dim) float32
extern "C" int pvo_cvx_upsample(const float* data, const void* mask, float* out, int N, int ht, int wd, int dim, int mask_dtype,
                                void* stream) {
  if (N < 0 || ht <= 0 || wd <= 0 || dim <= 0) return PVO_ERR_INVALID;
  if (N == 0) return PVO_OK;
  dim3 grid(ceil_div(wd, kXT), ht, N);
  cudaStream_t s = (cudaStream_t)stream;
  if (mask_dtype == PVO_DTYPE_F32) cvx_upsample_kernel<float><<<grid, 256, 0, s>>>(data, (const float*)mask, out, N, ht, wd, dim);
  else if (mask_dtype == PVO_DTYPE_F16) cvx_upsample_kernel<__half><<<grid, 256, 0, s>>>(data, (const __half*)mask, out, N, ht, wd, dim);
  else return PVO_ERR_UNSUPPORTED;
  return pvo_check_launch();
}

// in (N,ht,wd,dim) float32 -> out (N,8ht,8wd,dim) float32
extern "C" int pvo_upsample_inter(const float* in, float* out, int N, int ht, int wd, int dim, void* stream) {
  if (N < 0 || ht <= 0 || wd <= 0 || dim <= 0) return PVO_ERR_INVALID;
  if (N == 0) return PVO_OK;
  const long total = (long)N * 64 * ht * wd;
  upsample_inter_kernel<<<ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(in, out, N, ht, wd, dim);
  return pvo_check_launch();
}
