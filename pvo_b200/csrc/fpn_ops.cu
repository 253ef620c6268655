// Memory-bound helper kernels of the Panoptic-FPN forward (pixel-major / NHWC activations): everything in the dense
// part of the network that is not a convolution.  References under VPS_Module/detectron2/:
//   max_pool 3x3/s2 after the stem            modeling/backbone/resnet.py:356-359
//   stride-2 sampling for the strided 1x1s    modeling/backbone/resnet.py:152-160 (STRIDE_IN_1X1, defaults.py:489)
//   top-down nearest x2 + lateral sum         modeling/backbone/fpn.py:139-147
//   GroupNorm(32) + ReLU, bilinear x2         modeling/meta_arch/semantic_seg.py:186-207
//   panoptic merge                            modeling/meta_arch/panoptic_fpn.py:497-588 (a Python loop with .item() syncs
//                                             per instance / per label in the reference; one block per image here)
#include "common.cuh"

namespace {

template <typename T> __device__ __forceinline__ float ldv(const T* p);
template <> __device__ __forceinline__ float ldv<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float ldv<__half>(const __half* p) { return __half2float(__ldg(p)); }
template <typename T> __device__ __forceinline__ void stv(T* p, float v);
template <> __device__ __forceinline__ void stv<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void stv<__half>(__half* p, float v) { *p = __float2half_rn(v); }

// out[n][yo][xo][c] = max over the 3x3 window (stride 2, pad 1) of in
template <typename T>
__global__ void __launch_bounds__(256) maxpool3s2_kernel(const T* __restrict__ in, T* __restrict__ out, int N, int H, int W, int C) {
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)N * Ho * Wo * C) return;
  const int c = (int)(i % C);
  long p = i / C;
  const int xo = (int)(p % Wo); p /= Wo;
  const int yo = (int)(p % Ho);
  const long n = p / Ho;
  float m = -INFINITY;
  for (int dy = 0; dy < 3; dy++)
    for (int dx = 0; dx < 3; dx++) {
      const int y = 2 * yo - 1 + dy, x = 2 * xo - 1 + dx;
      if (y >= 0 && y < H && x >= 0 && x < W) m = fmaxf(m, ldv(in + ((n * H + y) * W + x) * C + c));
    }
  stv(out + i, m);
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
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)N * H * W * C) return;
  const int c = (int)(i % C);
  long p = i / C;
  const int x = (int)(p % W); p /= W;
  const int y = (int)(p % H);
  const long n = p / H;
  const int yt = min(y >> 1, Ht - 1), xt = min(x >> 1, Wt - 1);
  stv(lat + i, ldv(lat + i) + ldv(top + ((n * Ht + yt) * Wt + xt) * C + c));
}

// F.interpolate(scale_factor=S, mode='bilinear', align_corners=False); optional accumulate into out
template <typename T>
__global__ void __launch_bounds__(256) upsample_bilinear_kernel(const T* __restrict__ in, T* __restrict__ out, int N, int H, int W,
                                                                int C, int S, int accumulate) {
  const int Ho = H * S, Wo = W * S;
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)N * Ho * Wo * C) return;
  const int c = (int)(i % C);
  long p = i / C;
  const int xo = (int)(p % Wo); p /= Wo;
  const int yo = (int)(p % Ho);
  const long n = p / Ho;
  const float sy = fmaxf((yo + 0.5f) / S - 0.5f, 0.0f), sx = fmaxf((xo + 0.5f) / S - 0.5f, 0.0f);
  const int y0 = (int)sy, x0 = (int)sx;
  const int y1 = min(y0 + 1, H - 1), x1 = min(x0 + 1, W - 1);
  const float ly = sy - y0, lx = sx - x0;
  const T* b = in + n * (long)H * W * C + c;
  float v = (1.f - ly) * ((1.f - lx) * ldv(b + ((long)y0 * W + x0) * C) + lx * ldv(b + ((long)y0 * W + x1) * C)) +
            ly * ((1.f - lx) * ldv(b + ((long)y1 * W + x0) * C) + lx * ldv(b + ((long)y1 * W + x1) * C));
  if (accumulate) v += ldv(out + i);
  stv(out + i, v);
}

// GroupNorm statistics from per-(image, channel) partial sums produced by the instance-norm partial kernel layout:
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
__global__ void __launch_bounds__(64) gn_stats_partial_kernel(const T* __restrict__ in, float* __restrict__ part, int HW, int C, int chunks) {
  const int e = blockIdx.x, ch = blockIdx.y;
  const int per = (HW + chunks - 1) / chunks;
  const int p0 = ch * per, p1 = min(HW, p0 + per);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.0f, q = 0.0f;
    for (int p = p0; p < p1; p++) {
      const float v = ldv(in + ((long)e * HW + p) * C + c);
      s += v;
      q = fmaf(v, v, q);
    }
    part[(((long)e * chunks + ch) * C + c) * 2] = s;
    part[(((long)e * chunks + ch) * C + c) * 2 + 1] = q;
  }
}
template <typename T>
__global__ void __launch_bounds__(256) gn_apply_kernel(const T* __restrict__ in, const float* __restrict__ stats, const float* __restrict__ gamma,
                                                       const float* __restrict__ beta, T* __restrict__ out, int HW, int C, long total, int relu) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % C);
  const long n = i / ((long)HW * C);
  float v = (ldv(in + i) - stats[(n * C + c) * 2]) * stats[(n * C + c) * 2 + 1] * gamma[c] + beta[c];
  if (relu) v = fmaxf(v, 0.0f);
  stv(out + i, v);
}

// ---- panoptic merge: one block per image.  Instances are visited in descending score order (order[] prepared by the
// caller); every decision is block-uniform, so the sequential dependency of the reference loop is kept exactly.
// info_out[n][k] = {segment id, category, isthing, area, instance index}; counts_out[n] = number of segments.
__device__ __forceinline__ int block_sum_int(int v, int* sh) {
  v = __reduce_add_sync(0xffffffffu, v);
  const int w = threadIdx.x >> 5;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[w] = v;
  __syncthreads();
  int t = 0;
  for (int k = 0; k < (int)(blockDim.x >> 5); k++) t += sh[k];
  return t;
}

__global__ void __launch_bounds__(1024) panoptic_combine_kernel(const unsigned char* __restrict__ masks /*[R][HW] bool*/, const float* __restrict__ scores,
                                                                const long* __restrict__ classes, const long* __restrict__ order,
                                                                const int* __restrict__ inst_off /*[N+1]*/, const long* __restrict__ sem /*[N][HW] argmax*/,
                                                                int* __restrict__ pan /*[N][HW]*/, int* __restrict__ info_out, int* __restrict__ counts_out,
                                                                int HW, int num_sem, float overlap_thr, int stuff_area_thr, float score_thr,
                                                                int thing_cat0, int max_info) {
  __shared__ int sh[32];
  const int n = blockIdx.x;
  int* P = pan + (long)n * HW;
  for (int i = threadIdx.x; i < HW; i += blockDim.x) P[i] = 0;
  __syncthreads();
  int cur = 0;
  const int r0 = inst_off[n], r1 = inst_off[n + 1];
  for (int k = r0; k < r1; k++) {
    const int r = r0 + (int)order[k];                       // order[] holds per-image ranks sorted by descending score
    if (scores[r] < score_thr) break;
    const unsigned char* M = masks + (long)r * HW;
    int area = 0, inter = 0;
    for (int i = threadIdx.x; i < HW; i += blockDim.x) {
      const int m = M[i] != 0;
      area += m;
      inter += m && (P[i] > 0);
    }
    area = block_sum_int(area, sh);
    inter = block_sum_int(inter, sh);
    if (area == 0) continue;
    if ((double)inter / (double)area > (double)overlap_thr) continue;       // the reference divides Python numbers (doubles)
    cur++;
    const int cat = thing_cat0 + (int)classes[r];
    const int id = cat * 10000 + cur;
    for (int i = threadIdx.x; i < HW; i += blockDim.x)
      if (M[i] && P[i] == 0) P[i] = id;
    if (threadIdx.x == 0 && cur <= max_info) {
      int* o = info_out + ((long)n * max_info + cur - 1) * 5;
      o[0] = id; o[1] = cat; o[2] = 1; o[3] = area - inter; o[4] = r - r0;
    }
    __syncthreads();
  }
  const long* S = sem + (long)n * HW;
  for (int lab = 1; lab < num_sem; lab++) {                 // label 0 is the "thing" placeholder class (:560-561)
    int area = 0, present = 0;
    for (int i = threadIdx.x; i < HW; i += blockDim.x) {
      const int is = S[i] == lab;
      present |= is;
      area += is && (P[i] == 0);
    }
    area = block_sum_int(area, sh);
    present = block_sum_int(present, sh);
    if (!present || area < stuff_area_thr) continue;
    cur++;
    const int id = lab * 10000;                             // stuff_id2cat is the identity on VKITTI trainIds 1..11
    for (int i = threadIdx.x; i < HW; i += blockDim.x)
      if (S[i] == lab && P[i] == 0) P[i] = id;
    if (threadIdx.x == 0 && cur <= max_info) {
      int* o = info_out + ((long)n * max_info + cur - 1) * 5;
      o[0] = id; o[1] = lab; o[2] = 0; o[3] = area; o[4] = -1;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) counts_out[n] = cur;
}

}  // namespace

#define DISPATCH_T(dtype, CALL_F32, CALL_F16) \
  if ((dtype) == PVO_DTYPE_F32) { CALL_F32; } else if ((dtype) == PVO_DTYPE_F16) { CALL_F16; } else return PVO_ERR_UNSUPPORTED;

extern "C" int pvo_maxpool3s2(const void* in, void* out, int N, int H, int W, int C, int dtype, void* stream) {
  const long total = (long)N * ((H - 1) / 2 + 1) * ((W - 1) / 2 + 1) * C;
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
  const long total = (long)N * H * W * C;
  if (total <= 0) return PVO_OK;
  cudaStream_t s = (cudaStream_t)stream;
  DISPATCH_T(dtype, (upsample_add_kernel<float><<<ceil_div(total, 256), 256, 0, s>>>((float*)lat, (const float*)top, N, H, W, C, Ht, Wt)),
             (upsample_add_kernel<__half><<<ceil_div(total, 256), 256, 0, s>>>((__half*)lat, (const __half*)top, N, H, W, C, Ht, Wt)))
  return pvo_check_launch();
}

extern "C" int pvo_upsample_bilinear(const void* in, void* out, int N, int H, int W, int C, int scale, int accumulate, int dtype,
                                     void* stream) {
  const long total = (long)N * H * scale * W * scale * C;
  if (total <= 0) return PVO_OK;
  cudaStream_t s = (cudaStream_t)stream;
  DISPATCH_T(dtype, (upsample_bilinear_kernel<float><<<ceil_div(total, 256), 256, 0, s>>>((const float*)in, (float*)out, N, H, W, C, scale, accumulate)),
             (upsample_bilinear_kernel<__half><<<ceil_div(total, 256), 256, 0, s>>>((const __half*)in, (__half*)out, N, H, W, C, scale, accumulate)))
  return pvo_check_launch();
}

// GroupNorm(num_groups = C / group_size, eps) with affine, fused ReLU; in == out allowed.  scratch: N*32*C*2 + N*C*2 floats
extern "C" int pvo_group_norm(const void* in, void* out, const float* gamma, const float* beta, float* scratch, int N, int HW, int C,
                              int group_size, float eps, int relu, int dtype, void* stream) {
  if (N == 0) return PVO_OK;
  if (!scratch || C % group_size) return PVO_ERR_INVALID;
  constexpr int chunks = 32;
  cudaStream_t s = (cudaStream_t)stream;
  float* part = scratch;
  float* stats = scratch + (size_t)N * chunks * C * 2;
  dim3 grid(N, chunks);
  const long total = (long)N * HW * C;
  if (dtype == PVO_DTYPE_F32) gn_stats_partial_kernel<float><<<grid, 64, 0, s>>>((const float*)in, part, HW, C, chunks);
  else gn_stats_partial_kernel<__half><<<grid, 64, 0, s>>>((const __half*)in, part, HW, C, chunks);
  gn_stats_final_kernel<<<N, 128, 0, s>>>(part, stats, HW, C, group_size, chunks, eps);
  DISPATCH_T(dtype, (gn_apply_kernel<float><<<ceil_div(total, 256), 256, 0, s>>>((const float*)in, stats, gamma, beta, (float*)out, HW, C, total, relu)),
             (gn_apply_kernel<__half><<<ceil_div(total, 256), 256, 0, s>>>((const __half*)in, stats, gamma, beta, (__half*)out, HW, C, total, relu)))
  return pvo_check_launch();
}

// combine_semantic_and_instance_outputs for a batch: masks (R,HW) uint8 (all images, image n owns rows inst_off[n]..inst_off[n+1]),
// order = per-image rank lists (descending score) concatenated, sem (N,HW) int64 argmax labels.  Outputs: pan (N,HW) int32,
// info (N,max_info,5) int32, counts (N) int32.
extern "C" int pvo_panoptic_combine(const void* masks, const float* scores, const long* classes, const long* order, const int* inst_off,
                                    const long* sem, int* pan, int* info_out, int* counts_out, int N, int HW, int num_sem,
                                    float overlap_thr, int stuff_area_thr, float score_thr, int thing_cat0, int max_info, void* stream) {
  if (N == 0) return PVO_OK;
  panoptic_combine_kernel<<<N, 1024, 0, (cudaStream_t)stream>>>((const unsigned char*)masks, scores, classes, order, inst_off, sem, pan,
                                                                 info_out, counts_out, HW, num_sem, overlap_thr, stuff_area_thr, score_thr,
                                                                 thing_cat0, max_info);
  return pvo_check_launch();
}
