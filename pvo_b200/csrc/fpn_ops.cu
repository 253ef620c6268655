// Memory-bound helper kernels of the Panoptic-FPN forward (pixel-major / NHWC activations): everything in the dense
// part of the network that is not a convolution.  References under VPS_Module/detectron2/:
//   max_pool 3x3/s2 after the stem            modeling/backbone/resnet.py:356-359
//   stride-2 sampling for the strided 1x1s    modeling/backbone/resnet.py:152-160 (STRIDE_IN_1X1, defaults.py:489)
//   top-down nearest x2 + lateral sum         modeling/backbone/fpn.py:139-147
//   GroupNorm(32) + ReLU, bilinear x2         modeling/meta_arch/semantic_seg.py:186-207
#include "common.cuh"

namespace {

// 16-byte channel vectors: 8 halves or 4 floats per thread (C % VE == 0 on every call site; pointers 16-B aligned)
template <typename T> struct Vec;
template <> struct Vec<float> {
  static constexpr int VE = 4;
  static __device__ __forceinline__ void ld(const float* p, float (&v)[4]) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  static __device__ __forceinline__ void st(float* p, const float (&v)[4]) { *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); }
};
template <> struct Vec<__half> {
  static constexpr int VE = 8;
  static __device__ __forceinline__ void ld(const __half* p, float (&v)[8]) {
    const uint4 t = __ldg(reinterpret_cast<const uint4*>(p));
    const __half2* h = reinterpret_cast<const __half2*>(&t);
#pragma unroll
    for (int j = 0; j < 4; j++) { const float2 f = __half22float2(h[j]); v[2 * j] = f.x; v[2 * j + 1] = f.y; }
  }
  static __device__ __forceinline__ void st(__half* p, const float (&v)[8]) {
    uint4 t;
    __half2* h = reinterpret_cast<__half2*>(&t);
#pragma unroll
    for (int j = 0; j < 4; j++) h[j] = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
    *reinterpret_cast<uint4*>(p) = t;
  }
};

// out[n][yo][xo][c] = max over the 3x3 window (stride 2, pad 1) of in
template <typename T>
__global__ void __launch_bounds__(256) maxpool3s2_kernel(const T* __restrict__ in, T* __restrict__ out, int N, int H, int W, int C) {
  constexpr int VE = Vec<T>::VE;
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1, CV = C / VE;
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)N * Ho * Wo * CV) return;
  const int c = (int)(i % CV) * VE;
  long p = i / CV;
  const int xo = (int)(p % Wo); p /= Wo;
  const int yo = (int)(p % Ho);
  const long n = p / Ho;
  float m[VE];
#pragma unroll
  for (int j = 0; j < VE; j++) m[j] = -INFINITY;
  for (int dy = 0; dy < 3; dy++)
    for (int dx = 0; dx < 3; dx++) {
      const int y = 2 * yo - 1 + dy, x = 2 * xo - 1 + dx;
      if (y >= 0 && y < H && x >= 0 && x < W) {
        float v[VE];
        Vec<T>::ld(in + ((n * H + y) * W + x) * C + c, v);
#pragma unroll
        for (int j = 0; j < VE; j++) m[j] = fmaxf(m[j], v[j]);
      }
    }
  Vec<T>::st(out + ((n * Ho + yo) * Wo + xo) * C + c, m);
}

// out[n][yo][xo][:] = in[n][2yo][2xo][:]   (16-byte vectors)
__global__ void __launch_bounds__(256) subsample2_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int N, int H, int W, int CV) {
  const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)N * Ho * Wo * CV) return;
  const int c = (int)(i % CV);
  long p = i / CV;
  const int xo = (int)(p % Wo); p /= Wo;
  const int yo = (int)(p % Ho);
  const long n = p / Ho;
  out[i] = __ldg(in + ((n * H + 2 * yo) * W + 2 * xo) * CV + c);
}

// lat[n][y][x][c] += top[n][y/2][x/2][c]   (FPN top-down pathway, nearest-neighbour x2)
template <typename T>
__global__ void __launch_bounds__(256) upsample_add_kernel(T* __restrict__ lat, const T* __restrict__ top, int N, int H, int W, int C,
                                                           int Ht, int Wt) {
  constexpr int VE = Vec<T>::VE;
  const int CV = C / VE;
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)N * H * W * CV) return;
  const int c = (int)(i % CV) * VE;
  long p = i / CV;
  const int x = (int)(p % W); p /= W;
  const int y = (int)(p % H);
  const long n = p / H;
  const int yt = min(y >> 1, Ht - 1), xt = min(x >> 1, Wt - 1);
  float a[VE], b[VE];
  T* o = lat + ((n * H + y) * W + x) * C + c;
  Vec<T>::ld(o, a);
  Vec<T>::ld(top + ((n * Ht + yt) * Wt + xt) * C + c, b);
#pragma unroll
  for (int j = 0; j < VE; j++) a[j] += b[j];
  Vec<T>::st(o, a);
}

// F.interpolate(scale_factor=S, mode='bilinear', align_corners=False); optional accumulate into out
template <typename T>
__global__ void __launch_bounds__(256) upsample_bilinear_kernel(const T* __restrict__ in, T* __restrict__ out, int N, int H, int W,
                                                                int C, int S, int accumulate) {
  constexpr int VE = Vec<T>::VE;
  const int Ho = H * S, Wo = W * S, CV = C / VE;
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)N * Ho * Wo * CV) return;
  const int c = (int)(i % CV) * VE;
  long p = i / CV;
  const int xo = (int)(p % Wo); p /= Wo;
  const int yo = (int)(p % Ho);
  const long n = p / Ho;
  const float sy = fmaxf((yo + 0.5f) / S - 0.5f, 0.0f), sx = fmaxf((xo + 0.5f) / S - 0.5f, 0.0f);
  const int y0 = (int)sy, x0 = (int)sx;
  const int y1 = min(y0 + 1, H - 1), x1 = min(x0 + 1, W - 1);
  const float ly = sy - y0, lx = sx - x0;
  const T* b = in + n * (long)H * W * C + c;
  float v00[VE], v01[VE], v10[VE], v11[VE], r[VE];
  Vec<T>::ld(b + ((long)y0 * W + x0) * C, v00);
  Vec<T>::ld(b + ((long)y0 * W + x1) * C, v01);
  Vec<T>::ld(b + ((long)y1 * W + x0) * C, v10);
  Vec<T>::ld(b + ((long)y1 * W + x1) * C, v11);
  T* o = out + ((n * Ho + yo) * Wo + xo) * C + c;
  if (accumulate) Vec<T>::ld(o, r);
#pragma unroll
  for (int j = 0; j < VE; j++) {
    const float v = (1.f - ly) * ((1.f - lx) * v00[j] + lx * v01[j]) + ly * ((1.f - lx) * v10[j] + lx * v11[j]);
    r[j] = accumulate ? r[j] + v : v;
  }
  Vec<T>::st(o, r);
}

// GroupNorm statistics: part[n][chunk][c][2] = {sum, sum of squares} over the chunk's pixels; 256 threads = (C/VE) channel
// vectors x row lanes, reduced through shared memory
template <typename T>
__global__ void __launch_bounds__(256) gn_stats_partial_kernel(const T* __restrict__ in, float* __restrict__ part, int HW, int C, int chunks) {
  constexpr int VE = Vec<T>::VE;
  extern __shared__ float red[];                 // [rows][C][2]
  const int e = blockIdx.x, ch = blockIdx.y, CV = C / VE;
  const int rows = blockDim.x / CV;              // row lanes (host guarantees CV <= 256)
  const int cv = threadIdx.x % CV, rl = threadIdx.x / CV;
  const int per = (HW + chunks - 1) / chunks;
  const int p0 = ch * per, p1 = min(HW, p0 + per);
  float s[VE], q[VE];
#pragma unroll
  for (int j = 0; j < VE; j++) { s[j] = 0.0f; q[j] = 0.0f; }
  if (rl < rows)
    for (int p = p0 + rl; p < p1; p += rows) {
      float v[VE];
      Vec<T>::ld(in + ((long)e * HW + p) * C + cv * VE, v);
#pragma unroll
      for (int j = 0; j < VE; j++) { s[j] += v[j]; q[j] = fmaf(v[j], v[j], q[j]); }
    }
  if (rl < rows) {
#pragma unroll
    for (int j = 0; j < VE; j++) {
      red[((long)rl * C + cv * VE + j) * 2] = s[j];
      red[((long)rl * C + cv * VE + j) * 2 + 1] = q[j];
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float ts = 0.0f, tq = 0.0f;
    for (int r = 0; r < rows; r++) { ts += red[((long)r * C + c) * 2]; tq += red[((long)r * C + c) * 2 + 1]; }
    part[(((long)e * chunks + ch) * C + c) * 2] = ts;
    part[(((long)e * chunks + ch) * C + c) * 2 + 1] = tq;
  }
}

// part[n][chunk][c][2] -> stats[n][c][2] = {group mean, group rstd}
__global__ void __launch_bounds__(128) gn_stats_final_kernel(const float* __restrict__ part, float* __restrict__ stats, int HW, int C,
                                                             int gs, int chunks, float eps) {
  const int n = blockIdx.x;
  for (int g = threadIdx.x; g < C / gs; g += blockDim.x) {
    double s = 0.0, q = 0.0;
    for (int k = 0; k < chunks; k++)
      for (int j = 0; j < gs; j++) {
        s += part[(((long)n * chunks + k) * C + g * gs + j) * 2];
        q += part[(((long)n * chunks + k) * C + g * gs + j) * 2 + 1];
      }
    const double cnt = (double)HW * gs;
    const double mean = s / cnt;
    const double var = fmax(q / cnt - mean * mean, 0.0);
    for (int j = 0; j < gs; j++) {
      stats[((long)n * C + g * gs + j) * 2] = (float)mean;
      stats[((long)n * C + g * gs + j) * 2 + 1] = (float)(1.0 / sqrt(var + (double)eps));
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) gn_apply_kernel(const T* __restrict__ in, const float* __restrict__ stats, const float* __restrict__ gamma,
                                                       const float* __restrict__ beta, T* __restrict__ out, int HW, int C, long total, int relu) {
  constexpr int VE = Vec<T>::VE;
  const long i = ((long)blockIdx.x * blockDim.x + threadIdx.x) * VE;
  if (i >= total) return;
  const int c = (int)(i % C);
  const long n = i / ((long)HW * C);
  float v[VE];
  Vec<T>::ld(in + i, v);
#pragma unroll
  for (int j = 0; j < VE; j++) {
    const float2 st = __ldg(reinterpret_cast<const float2*>(stats) + n * C + c + j);
    v[j] = (v[j] - st.x) * st.y * __ldg(gamma + c + j) + __ldg(beta + c + j);
    if (relu) v[j] = fmaxf(v[j], 0.0f);
  }
  Vec<T>::st(out + i, v);
}

// Stem input for the tensor-core path.  The 7x7 / stride-2 / 3-channel stem (resnet.py:340-354) is rewritten as a 5x1 stride-1
// convolution over a 2x2 space-to-depth, x-expanded copy of the image:
//   E[n][Yb][X][tx*12 + dy*6 + dx*3 + c] = x[n][2*Yb + dy][2*(X - 2 + tx) + dx][c]   (tx < 4; channels 48..63 zero)
// so that out[Y][X] = sum_{ty, j} W2[ty][j] * E[Y + ty - 2][X][j] with W2[ty][(tx,dy,dx,c)] = w[ky = 2ty + dy - 1][kx = 2tx + dx - 1][c].
__global__ void __launch_bounds__(256) stem_expand_kernel(const __half* __restrict__ x, __half* __restrict__ E, int N, int H, int W) {
  const int Ho = H / 2, Wo = W / 2;
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)N * Ho * Wo * 8) return;
  const int q = (int)(i & 7);
  long p = i >> 3;
  const int X = (int)(p % Wo); p /= Wo;
  const int Yb = (int)(p % Ho);
  const long n = p / Ho;
  __align__(16) __half v[8];
#pragma unroll
  for (int e = 0; e < 8; e++) {
    const int j = q * 8 + e;
    __half val = __float2half_rn(0.0f);
    if (j < 48) {
      const int tx = j / 12, r = j % 12, dy = r / 6, dx = (r % 6) / 3, c = r % 3;
      const int yy = 2 * Yb + dy, xx = 2 * (X - 2 + tx) + dx;
      if (xx >= 0 && xx < W) val = __ldg(x + ((n * H + yy) * W + xx) * 3 + c);
    }
    v[e] = val;
  }
  *reinterpret_cast<uint4*>(E + i * 8) = *reinterpret_cast<const uint4*>(v);
}

}  // namespace

#define DISPATCH_T(dtype, CALL_F32, CALL_F16) \
  if ((dtype) == PVO_DTYPE_F32) { CALL_F32; } else if ((dtype) == PVO_DTYPE_F16) { CALL_F16; } else return PVO_ERR_UNSUPPORTED;

#define VE_OF(dtype) ((dtype) == PVO_DTYPE_F32 ? 4 : 8)

extern "C" int pvo_maxpool3s2(const void* in, void* out, int N, int H, int W, int C, int dtype, void* stream) {
  if (C % VE_OF(dtype)) return PVO_ERR_INVALID;
  const long total = (long)N * ((H - 1) / 2 + 1) * ((W - 1) / 2 + 1) * (C / VE_OF(dtype));
  if (total <= 0) return PVO_OK;
  cudaStream_t s = (cudaStream_t)stream;
  DISPATCH_T(dtype, (maxpool3s2_kernel<float><<<ceil_div(total, 256), 256, 0, s>>>((const float*)in, (float*)out, N, H, W, C)),
             (maxpool3s2_kernel<__half><<<ceil_div(total, 256), 256, 0, s>>>((const __half*)in, (__half*)out, N, H, W, C)))
  return pvo_check_launch();
}

extern "C" int pvo_subsample2(const void* in, void* out, int N, int H, int W, int C, int dtype, void* stream) {
  const size_t esz = dtype == PVO_DTYPE_F32 ? 4 : 2;
  if ((C * esz) % 16) return PVO_ERR_INVALID;
  const int CV = (int)(C * esz / 16);
  const long total = (long)N * ((H + 1) / 2) * ((W + 1) / 2) * CV;
  if (total <= 0) return PVO_OK;
  subsample2_kernel<<<ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>((const uint4*)in, (uint4*)out, N, H, W, CV);
  return pvo_check_launch();
}

extern "C" int pvo_upsample_add(void* lat, const void* top, int N, int H, int W, int C, int Ht, int Wt, int dtype, void* stream) {
  if (C % VE_OF(dtype)) return PVO_ERR_INVALID;
  const long total = (long)N * H * W * (C / VE_OF(dtype));
  if (total <= 0) return PVO_OK;
  cudaStream_t s = (cudaStream_t)stream;
  DISPATCH_T(dtype, (upsample_add_kernel<float><<<ceil_div(total, 256), 256, 0, s>>>((float*)lat, (const float*)top, N, H, W, C, Ht, Wt)),
             (upsample_add_kernel<__half><<<ceil_div(total, 256), 256, 0, s>>>((__half*)lat, (const __half*)top, N, H, W, C, Ht, Wt)))
  return pvo_check_launch();
}

extern "C" int pvo_upsample_bilinear(const void* in, void* out, int N, int H, int W, int C, int scale, int accumulate, int dtype,
                                     void* stream) {
  if (C % VE_OF(dtype)) return PVO_ERR_INVALID;
  const long total = (long)N * H * scale * W * scale * (C / VE_OF(dtype));
  if (total <= 0) return PVO_OK;
  cudaStream_t s = (cudaStream_t)stream;
  DISPATCH_T(dtype, (upsample_bilinear_kernel<float><<<ceil_div(total, 256), 256, 0, s>>>((const float*)in, (float*)out, N, H, W, C, scale, accumulate)),
             (upsample_bilinear_kernel<__half><<<ceil_div(total, 256), 256, 0, s>>>((const __half*)in, (__half*)out, N, H, W, C, scale, accumulate)))
  return pvo_check_launch();
}

// GroupNorm(num_groups = C / group_size, eps) with affine, fused ReLU; in == out allowed.  scratch: N*64*C*2 + N*C*2 floats
extern "C" int pvo_group_norm(const void* in, void* out, const float* gamma, const float* beta, float* scratch, int N, int HW, int C,
                              int group_size, float eps, int relu, int dtype, void* stream) {
  if (N == 0) return PVO_OK;
  const int ve = VE_OF(dtype);
  if (!scratch || C % group_size || C % ve || C / ve > 256) return PVO_ERR_INVALID;
  constexpr int chunks = 64;
  cudaStream_t s = (cudaStream_t)stream;
  float* part = scratch;
  float* stats = scratch + (size_t)N * chunks * C * 2;
  dim3 grid(N, chunks);
  const long total = (long)N * HW * C;
  const size_t smem = (size_t)(256 / (C / ve)) * C * 2 * sizeof(float);
  if (smem > 48 * 1024) return PVO_ERR_UNSUPPORTED;
  if (dtype == PVO_DTYPE_F32) gn_stats_partial_kernel<float><<<grid, 256, smem, s>>>((const float*)in, part, HW, C, chunks);
  else gn_stats_partial_kernel<__half><<<grid, 256, smem, s>>>((const __half*)in, part, HW, C, chunks);
  gn_stats_final_kernel<<<N, 128, 0, s>>>(part, stats, HW, C, group_size, chunks, eps);
  DISPATCH_T(dtype, (gn_apply_kernel<float><<<ceil_div(total / ve, 256), 256, 0, s>>>((const float*)in, stats, gamma, beta, (float*)out, HW, C, total, relu)),
             (gn_apply_kernel<__half><<<ceil_div(total / ve, 256), 256, 0, s>>>((const __half*)in, stats, gamma, beta, (__half*)out, HW, C, total, relu)))
  return pvo_check_launch();
}

/* x (N,H,W,3) half, H and W even -> E (N,H/2,W/2,64) half (see stem_expand_kernel) */
extern "C" int pvo_stem_expand(const void* x, void* E, int N, int H, int W, void* stream) {
  if ((H | W) & 1) return PVO_ERR_INVALID;
  const long total = (long)N * (H / 2) * (W / 2) * 8;
  if (total <= 0) return PVO_OK;
  stem_expand_kernel<<<ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>((const __half*)x, (__half*)E, N, H, W);
  return pvo_check_launch();
}
