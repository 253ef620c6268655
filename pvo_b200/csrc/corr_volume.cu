// All-pairs correlation volume + pooling pyramid on CUDA cores (any dtype), and the pooling kernel shared with the
// tensor-core build.
//
// Replaces CorrBlock.corr + the avg_pool2d pyramid of CorrBlock.__init__
// (VO_Module/droid_slam/modules/corr.py:23-38,63-71): corr[e][p1][p2] = <fmap1[e][:,p1]/4, fmap2[e][:,p2]/4>.
// This SIMT version is the float32 path (training-style volumes) and the fallback; the float16 inference volume is
// built by the tcgen05 kernel in corr_volume_tc.cu.  Feature maps are pixel-major (E, HW, C) so both GEMM operands
// are K-major.  float16: operands /4 rounded to half (as the reference), fp32 accumulate, rounded once to half -- what a half GEMM gives.
#include "common.cuh"

namespace {

template <typename T> __device__ __forceinline__ float ld(const T* p);
template <> __device__ __forceinline__ float ld<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float ld<__half>(const __half* p) { return __half2float(__ldg(p)); }
// x/4 rounded to the storage type, as `fmap / 4.0` on a half tensor does (corr.py:66-67); exact unless subnormal
template <typename T> __device__ __forceinline__ float quarter(float x);
template <> __device__ __forceinline__ float quarter<float>(float x) { return x * 0.25f; }
template <> __device__ __forceinline__ float quarter<__half>(float x) { return __half2float(__float2half_rn(x * 0.25f)); }
template <typename T> __device__ __forceinline__ void st(T* p, float v);
template <> __device__ __forceinline__ void st<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void st<__half>(__half* p, float v) { *p = __float2half_rn(v); }

// C[e][m][n] = scale * sum_k A[e][m][k] * B[e][n][k];  tile 128x128x16, 256 threads, 8x8 per thread
template <typename T>
__global__ void __launch_bounds__(256) gemm_nt_kernel(const T* __restrict__ A, const T* __restrict__ B, T* __restrict__ C,
                                                      int M, int N, int K, float scale) {
  constexpr int BM = 128, BN = 128, BK = 16;
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN + 4];
  const int e = blockIdx.z;
  A += (long)e * M * K;
  B += (long)e * N * K;
  C += (long)e * M * N;
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int lr = tid >> 1, lk = (tid & 1) * 8;
  float ra[8], rb[8];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int q = 0; q < 8; q++) {
      const int k = k0 + lk + q;
      ra[q] = (m0 + lr < M && k < K) ? quarter<T>(ld(A + (long)(m0 + lr) * K + k)) : 0.0f;
      rb[q] = (n0 + lr < N && k < K) ? quarter<T>(ld(B + (long)(n0 + lr) * K + k)) : 0.0f;
    }
  };
  auto stash = [&](int buf) {
#pragma unroll
    for (int q = 0; q < 8; q++) { As[buf][lk + q][lr] = ra[q]; Bs[buf][lk + q][lr] = rb[q]; }
  };
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int j = 0; j < 8; j++) acc[i][j] = 0.0f;
  const int ty = tid >> 4, tx = tid & 15;
  const int nk = (K + BK - 1) / BK;
  fetch(0); stash(0);
  __syncthreads();
  for (int t = 0; t < nk; t++) {
    const int buf = t & 1;
    if (t + 1 < nk) fetch((t + 1) * BK);
#pragma unroll
    for (int kk = 0; kk < BK; kk++) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 8 + 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (t + 1 < nk) stash(buf ^ 1);
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const int m = m0 + ty * 8 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + j - 4);
      if (n < N) st(C + (long)m * N + n, acc[i][j] * scale);
    }
  }
}

// F.avg_pool2d(x, 2, stride=2) on (R, h, w) planes: float accumulate in raster order, * 0.25, round to T
template <typename T>
__global__ void __launch_bounds__(256) avgpool2_kernel(const T* __restrict__ in, T* __restrict__ out, long R, int h, int w) {
  const int ho = h / 2, wo = w / 2;
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R * ho * wo) return;
  const long r = i / (ho * wo);
  const int rem = (int)(i - r * ho * wo);
  const int y = rem / wo, x = rem - y * wo;
  const T* p = in + r * (long)h * w + (long)(2 * y) * w + 2 * x;
  float s = ld(p);
  s += ld(p + 1);
  s += ld(p + w);
  s += ld(p + w + 1);
  st(out + i, s * 0.25f);
}

}  // namespace

// fmap1, fmap2: (E, h*w, C) pixel-major.  levels_out[l]: (E, h, w, h>>l, w>>l).
extern "C" int pvo_corr_volume(const void* fmap1, const void* fmap2, void* const* levels_out, int E, int h, int w, int C,
                               int num_levels, int dtype, void* stream) {
  if (E < 0 || h <= 0 || w <= 0 || C <= 0 || num_levels < 1) return PVO_ERR_INVALID;
  if (E == 0) return PVO_OK;
  const int HW = h * w;
  cudaStream_t s = (cudaStream_t)stream;
  dim3 grid(ceil_div(HW, 128), ceil_div(HW, 128), E);
  if (dtype == PVO_DTYPE_F32)
    gemm_nt_kernel<float><<<grid, 256, 0, s>>>((const float*)fmap1, (const float*)fmap2, (float*)levels_out[0], HW, HW, C, 1.0f);
  else if (dtype == PVO_DTYPE_F16)
    gemm_nt_kernel<__half><<<grid, 256, 0, s>>>((const __half*)fmap1, (const __half*)fmap2, (__half*)levels_out[0], HW, HW, C, 1.0f);
  else
    return PVO_ERR_UNSUPPORTED;
  int hh = h, ww = w;
  for (int l = 1; l < num_levels; l++) {
    const long R = (long)E * HW;
    const long n = R * (hh / 2) * (ww / 2);
    if (n > 0) {
      if (dtype == PVO_DTYPE_F32)
        avgpool2_kernel<float><<<ceil_div(n, 256), 256, 0, s>>>((const float*)levels_out[l - 1], (float*)levels_out[l], R, hh, ww);
      else
        avgpool2_kernel<__half><<<ceil_div(n, 256), 256, 0, s>>>((const __half*)levels_out[l - 1], (__half*)levels_out[l], R, hh, ww);
    }
    hh /= 2; ww /= 2;
  }
  return pvo_check_launch();
}
