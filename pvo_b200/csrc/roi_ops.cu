// Region ops of the Panoptic-FPN ROI heads on pixel-major (NHWC) pyramid features.  References under VPS_Module/detectron2/:
//   ROIPooler level assignment + ROIAlignV2    modeling/poolers.py:21-60,189-235, layers/roi_align.py:49-58
//                                              (torchvision roi_align, aligned=True, sampling_ratio=0)
//   batched NMS (one group per image x level / image x class)   layers/nms.py:9-22 -> torchvision nms
// The reference pools NCHW features with one thread per output element; here one warp-sized slice of channels is
// contiguous, so each bilinear tap is a coalesced 512-byte read and the output is written pixel-major -- directly the A
// operand of the tensor-core FC / conv that follows (no permute, no fp16 box coordinates).
#include "common.cuh"

namespace {

struct Pyr {
  const void* f[4];
  int H[4], W[4];
};

template <typename T> struct V2;
template <> struct V2<__half> {
  static __device__ __forceinline__ float2 ld(const __half* p) { return __half22float2(__ldg(reinterpret_cast<const __half2*>(p))); }
  static __device__ __forceinline__ void st(__half* p, float2 v) { *reinterpret_cast<__half2*>(p) = __floats2half2_rn(v.x, v.y); }
};
template <> struct V2<float> {
  static __device__ __forceinline__ float2 ld(const float* p) { return __ldg(reinterpret_cast<const float2*>(p)); }
  static __device__ __forceinline__ void st(float* p, float2 v) { *reinterpret_cast<float2*>(p) = v; }
};

// one block per (roi, output bin); thread t owns channels 2t, 2t+1
template <typename T>
__global__ void __launch_bounds__(128) roi_align_fpn_kernel(const __grid_constant__ Pyr pyr, int C, const float* __restrict__ rois, int S,
                                                            T* __restrict__ out) {
  const int r = blockIdx.x / (S * S);
  const int bin = blockIdx.x % (S * S);
  const int ph = bin / S, pw = bin % S;
  const float* roi = rois + (long)r * 5;
  const int n = (int)roi[0];
  const float x0 = roi[1], y0 = roi[2], x1 = roi[3], y1 = roi[4];
  // poolers.py assign_boxes_to_levels: floor(4 + log2(sqrt(area) / 224 + 1e-8)) clamped to [2, 5]
  const float size = sqrtf(fmaxf((x1 - x0) * (y1 - y0), 0.0f));
  int lvl = (int)floorf(4.0f + log2f(size / 224.0f + 1e-8f));
  lvl = min(max(lvl, 2), 5) - 2;
  const float scale = 1.0f / (float)(4 << lvl);
  const int H = pyr.H[lvl], W = pyr.W[lvl];
  const T* F = reinterpret_cast<const T*>(pyr.f[lvl]) + (long)n * H * W * C;
  const float sw = x0 * scale - 0.5f, sh = y0 * scale - 0.5f;
  const float rw = x1 * scale - 0.5f - sw, rh = y1 * scale - 0.5f - sh;
  const float bw = rw / (float)S, bh = rh / (float)S;
  const int gh = (int)ceilf(rh / (float)S), gw = (int)ceilf(rw / (float)S);
  const float count = fmaxf((float)(gh * gw), 1.0f);
  for (int c = 2 * threadIdx.x; c < C; c += 2 * blockDim.x) {
    float2 acc = make_float2(0.0f, 0.0f);
    for (int iy = 0; iy < gh; iy++) {
      float y = sh + ph * bh + (iy + 0.5f) * bh / (float)gh;
      for (int ix = 0; ix < gw; ix++) {
        float x = sw + pw * bw + (ix + 0.5f) * bw / (float)gw;
        float yy = y;
        if (yy < -1.0f || yy > (float)H || x < -1.0f || x > (float)W) continue;
        if (yy <= 0.0f) yy = 0.0f;
        if (x <= 0.0f) x = 0.0f;
        int yl = (int)yy, xl = (int)x, yh, xh;
        if (yl >= H - 1) { yh = yl = H - 1; yy = (float)yl; } else yh = yl + 1;
        if (xl >= W - 1) { xh = xl = W - 1; x = (float)xl; } else xh = xl + 1;
        const float ly = yy - yl, lx = x - xl, hy = 1.0f - ly, hx = 1.0f - lx;
        const float w1 = hy * hx, w2 = hy * lx, w3 = ly * hx, w4 = ly * lx;
        const float2 v1 = V2<T>::ld(F + ((long)yl * W + xl) * C + c), v2 = V2<T>::ld(F + ((long)yl * W + xh) * C + c);
        const float2 v3 = V2<T>::ld(F + ((long)yh * W + xl) * C + c), v4 = V2<T>::ld(F + ((long)yh * W + xh) * C + c);
        acc.x += w1 * v1.x + w2 * v2.x + w3 * v3.x + w4 * v4.x;
        acc.y += w1 * v1.y + w2 * v2.y + w3 * v3.y + w4 * v4.y;
      }
    }
    acc.x /= count; acc.y /= count;
    V2<T>::st(out + ((long)r * S * S + bin) * C + c, acc);
  }
}

// torchvision nms_kernel's devIoU, same operation order
__device__ __forceinline__ bool iou_over(const float4 a, const float4 b, float thr) {
  const float left = fmaxf(a.x, b.x), right = fminf(a.z, b.z);
  const float top = fmaxf(a.y, b.y), bottom = fminf(a.w, b.w);
  const float width = fmaxf(right - left, 0.0f), height = fmaxf(bottom - top, 0.0f);
  const float inter = width * height;
  const float sa = (a.z - a.x) * (a.w - a.y), sb = (b.z - b.x) * (b.w - b.y);
  return inter / (sa + sb - inter) > thr;
}

// One block per group; boxes of a group are sorted by descending score.  Phase 1: for every row i, the bitmask of
// later boxes j > i it would suppress (n <= 1024 -> 32 words per row, in shared memory).  Phase 2: warp 0 sweeps the rows in
// score order carrying the "removed" bitmask in registers (lane w owns word w) -- the sequential greedy NMS, on chip.
__global__ void __launch_bounds__(1024) nms_groups_kernel(const float4* __restrict__ boxes, const unsigned char* __restrict__ valid, int n,
                                                          float thr, unsigned char* __restrict__ keep) {
  extern __shared__ unsigned int sm[];
  float4* B = reinterpret_cast<float4*>(sm);                  // [n]
  unsigned int* M = sm + 4 * n;                               // [n][32]
  const long g = blockIdx.x;
  for (int i = threadIdx.x; i < n; i += blockDim.x) B[i] = boxes[g * n + i];
  __syncthreads();
  const int nw = (n + 31) >> 5;
  {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int t = warp; t < n * nw; t += nwarps) {             // one (row, word) per warp step: lane <-> candidate j
      const int i = t / nw, w = t % nw, j0 = w << 5;
      unsigned int bits = 0;
      if (j0 + 31 > i) {                                      // warp-uniform
        const int j = j0 + lane;
        const bool sup = j > i && j < n && iou_over(B[i], B[min(j, n - 1)], thr);
        bits = __ballot_sync(0xffffffffu, sup);
      }
      if (lane == 0) M[i * 32 + w] = bits;
    }
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    unsigned int removed = 0;                                 // word `lane` of the removed set
    for (int w = 0; w < nw; w++) {
      // invalid entries (empty boxes, padding, below the score threshold) start out removed
      unsigned int inv = 0;
      const int j = (w << 5) + lane;
      const bool ok = j < n && valid[g * n + j];
      inv = __ballot_sync(0xffffffffu, !ok);
      if (lane == w) removed |= inv;
    }
    for (int i = 0; i < n; i++) {
      const unsigned int word = __shfl_sync(0xffffffffu, removed, i >> 5);
      const bool alive = !((word >> (i & 31)) & 1u);
      if (alive && lane < nw) removed |= M[i * 32 + lane];
      if (lane == 0) keep[g * n + i] = alive ? 1 : 0;
    }
  }
}

}  // namespace

extern "C" int pvo_roi_align_fpn(const void* const* feats, const int* Hs, const int* Ws, int C, const float* rois, int R, int S, void* out,
                                 int dtype, void* stream) {
  if (R == 0) return PVO_OK;
  if (C % 2 || S < 1) return PVO_ERR_INVALID;
  Pyr p;
  for (int l = 0; l < 4; l++) { p.f[l] = feats[l]; p.H[l] = Hs[l]; p.W[l] = Ws[l]; }
  cudaStream_t s = (cudaStream_t)stream;
  const unsigned grid = (unsigned)((long)R * S * S);
  if (dtype == PVO_DTYPE_F16) roi_align_fpn_kernel<__half><<<grid, 128, 0, s>>>(p, C, rois, S, (__half*)out);
  else if (dtype == PVO_DTYPE_F32) roi_align_fpn_kernel<float><<<grid, 128, 0, s>>>(p, C, rois, S, (float*)out);
  else return PVO_ERR_UNSUPPORTED;
  return pvo_check_launch();
}

extern "C" int pvo_nms_groups(const float* boxes, const unsigned char* valid, int G, int n, float thr, unsigned char* keep, void* stream) {
  if (G == 0 || n == 0) return PVO_OK;
  if (n > 1024) return PVO_ERR_UNSUPPORTED;
  const size_t smem = (size_t)n * 16 + (size_t)n * 32 * 4;
  cudaFuncSetAttribute(nms_groups_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  nms_groups_kernel<<<G, 1024, smem, (cudaStream_t)stream>>>((const float4*)boxes, valid, n, thr, keep);
  return pvo_check_launch();
}
