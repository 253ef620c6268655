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

// one block per (roi, output row ph); warp w computes the bins pw = w, w + 4, ...; lane l owns channels [VE*l, VE*l + VE)
// of every 32*VE-channel slab, so each bilinear tap is one 512-byte coalesced read and neighbouring bins hit in L1.
template <typename T>
__global__ void __launch_bounds__(128) roi_align_fpn_kernel(const __grid_constant__ Pyr pyr, int C, const float* __restrict__ rois, int S,
                                                            T* __restrict__ out) {
  constexpr int VE = Vec<T>::VE;
  const int r = blockIdx.x / S, ph = blockIdx.x % S;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
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
  for (int pw = warp; pw < S; pw += 4) {
    for (int c = VE * lane; c < C; c += 32 * VE) {
      float acc[VE];
#pragma unroll
      for (int j = 0; j < VE; j++) acc[j] = 0.0f;
      for (int iy = 0; iy < gh; iy++) {
        const float y = sh + ph * bh + (iy + 0.5f) * bh / (float)gh;
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
          float v1[VE], v2[VE], v3[VE], v4[VE];
          Vec<T>::ld(F + ((long)yl * W + xl) * C + c, v1);
          Vec<T>::ld(F + ((long)yl * W + xh) * C + c, v2);
          Vec<T>::ld(F + ((long)yh * W + xl) * C + c, v3);
          Vec<T>::ld(F + ((long)yh * W + xh) * C + c, v4);
#pragma unroll
          for (int j = 0; j < VE; j++) acc[j] += w1 * v1[j] + w2 * v2[j] + w3 * v3[j] + w4 * v4[j];
        }
      }
#pragma unroll
      for (int j = 0; j < VE; j++) acc[j] /= count;
      Vec<T>::st(out + (((long)r * S + ph) * S + pw) * C + c, acc);
    }
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

// Greedy NMS in two launches.  (1) mask: for every row i the bitmask of later boxes j > i it would suppress, one warp per
// (row, 32-box word), spread over the whole GPU.  (2) sweep: one block per group stages its n x 32-word mask in shared
// memory and warp 0 walks the rows in score order carrying the "removed" set in registers (lane w owns word w).
__global__ void __launch_bounds__(256) nms_mask_kernel(const float4* __restrict__ boxes, int n, float thr, unsigned int* __restrict__ mask) {
  const long g = blockIdx.x;
  const int i = blockIdx.y * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (i >= n) return;
  const int nw = (n + 31) >> 5;
  const float4 a = __ldg(boxes + g * n + i);
  unsigned int* row = mask + (g * n + i) * 32;
  for (int w = 0; w < nw; w++) {
    unsigned int bits = 0;
    const int j = (w << 5) + lane;
    if ((w << 5) + 31 > i) {                                  // warp-uniform
      const bool sup = j > i && j < n && iou_over(a, __ldg(boxes + g * n + min(j, n - 1)), thr);
      bits = __ballot_sync(0xffffffffu, sup);
    }
    if (lane == 0) row[w] = bits;
  }
}

__global__ void __launch_bounds__(1024) nms_sweep_kernel(const unsigned int* __restrict__ mask, const unsigned char* __restrict__ valid, int n,
                                                         unsigned char* __restrict__ keep) {
  extern __shared__ unsigned int M[];                         // [n][32]
  const long g = blockIdx.x;
  const int nw = (n + 31) >> 5;
  for (int t = threadIdx.x; t < n * 32; t += blockDim.x)
    if ((t & 31) < nw) M[t] = mask[g * n * 32 + t];
  __syncthreads();
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    unsigned int removed = 0;                                 // word `lane` of the removed set
    for (int w = 0; w < nw; w++) {
      // invalid entries (empty boxes, padding, below the score threshold) start out removed
      const int j = (w << 5) + lane;
      const bool ok = j < n && valid[g * n + j];
      const unsigned int inv = __ballot_sync(0xffffffffu, !ok);
      if (lane == w) removed |= inv;
    }
    unsigned int mine = 0;                                    // keep bits of word `lane`
    for (int i = 0; i < n; i++) {
      const unsigned int word = __shfl_sync(0xffffffffu, removed, i >> 5);
      const bool alive = !((word >> (i & 31)) & 1u);
      if (alive && lane < nw) removed |= M[i * 32 + lane];
      if (alive && lane == (i >> 5)) mine |= 1u << (i & 31);
    }
    for (int w = 0; w < nw; w++) {
      const unsigned int bits = __shfl_sync(0xffffffffu, mine, w);
      const int j = (w << 5) + lane;
      if (j < n) keep[g * n + j] = (bits >> lane) & 1u;
    }
  }
}

// ---- box decoding, fused.  Box2BoxTransform.apply_deltas (modeling/box_regression.py:76-116) evaluated with the same
// individually rounded fp32 operations as the reference's tensor expressions (no FMA contraction), then Boxes.clip and the
// non-empty test (structures/boxes.py:185-215).
struct Box4 { float x0, y0, x1, y1; };
__device__ __forceinline__ Box4 apply_deltas(const float4 b, float dx, float dy, float dw, float dh, float wx, float wy, float ww, float wh) {
  const float kClamp = 4.135166556742356f;                    // log(1000 / 16)
  const float widths = __fsub_rn(b.z, b.x), heights = __fsub_rn(b.w, b.y);
  const float cx = __fadd_rn(b.x, __fmul_rn(0.5f, widths)), cy = __fadd_rn(b.y, __fmul_rn(0.5f, heights));
  dx = __fdiv_rn(dx, wx); dy = __fdiv_rn(dy, wy);
  dw = fminf(__fdiv_rn(dw, ww), kClamp); dh = fminf(__fdiv_rn(dh, wh), kClamp);
  const float pcx = __fadd_rn(__fmul_rn(dx, widths), cx), pcy = __fadd_rn(__fmul_rn(dy, heights), cy);
  const float pw = __fmul_rn(expf(dw), widths), ph = __fmul_rn(expf(dh), heights);
  Box4 o;
  o.x0 = __fsub_rn(pcx, __fmul_rn(0.5f, pw)); o.y0 = __fsub_rn(pcy, __fmul_rn(0.5f, ph));
  o.x1 = __fadd_rn(pcx, __fmul_rn(0.5f, pw)); o.y1 = __fadd_rn(pcy, __fmul_rn(0.5f, ph));
  return o;
}
__device__ __forceinline__ float clampf(float v, float hi) { return fminf(fmaxf(v, 0.0f), hi); }   // NaN -> 0..hi, flagged invalid before

struct RpnLevels {
  const __half* out[5];     // (N, h, w, 16): objectness x3 | deltas 3x4 | pad
  const long* idx[5];       // (N, k) top-k indices into (h, w, 3)
  int h[5], w[5], k[5], stride[5];
  float cell[5][3][4];      // cell anchors (anchor_generator.py generate_cell_anchors)
};

// grid (ceil(P/128), 5, N): one thread per (image, level, rank) -> boxes (N,5,P,4), scores (N,5,P), valid (N,5,P)
__global__ void __launch_bounds__(128) rpn_decode_kernel(const __grid_constant__ RpnLevels L, const float* __restrict__ limits /*(N,4)*/, int P,
                                                         float4* __restrict__ boxes, float* __restrict__ scores, unsigned char* __restrict__ valid) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x, l = blockIdx.y, n = blockIdx.z;
  if (r >= P) return;
  const long o = ((long)n * 5 + l) * P + r;
  if (r >= L.k[l]) {
    boxes[o] = make_float4(0.f, 0.f, 0.f, 0.f);
    scores[o] = -INFINITY;
    valid[o] = 0;
    return;
  }
  const long id = L.idx[l][(long)n * L.k[l] + r];
  const int a = (int)(id % 3);
  const long cell = id / 3;
  const int x = (int)(cell % L.w[l]), y = (int)(cell / L.w[l]);
  const __half* q = L.out[l] + ((long)n * L.h[l] * L.w[l] + cell) * 16;
  const float sx = (float)(x * L.stride[l]), sy = (float)(y * L.stride[l]);
  const float4 anc = make_float4(__fadd_rn(sx, L.cell[l][a][0]), __fadd_rn(sy, L.cell[l][a][1]), __fadd_rn(sx, L.cell[l][a][2]),
                                 __fadd_rn(sy, L.cell[l][a][3]));
  const float score = __half2float(q[a]);
  Box4 b = apply_deltas(anc, __half2float(q[3 + 4 * a]), __half2float(q[4 + 4 * a]), __half2float(q[5 + 4 * a]), __half2float(q[6 + 4 * a]),
                        1.f, 1.f, 1.f, 1.f);
  bool ok = isfinite(b.x0) && isfinite(b.y0) && isfinite(b.x1) && isfinite(b.y1) && isfinite(score);
  const float4 lim = __ldg(reinterpret_cast<const float4*>(limits) + n);
  b.x0 = clampf(b.x0, lim.x); b.y0 = clampf(b.y0, lim.y); b.x1 = clampf(b.x1, lim.z); b.y1 = clampf(b.y1, lim.w);
  ok = ok && (b.x1 - b.x0) > 0.0f && (b.y1 - b.y0) > 0.0f;    // min_box_size = 0
  boxes[o] = make_float4(b.x0, b.y0, b.x1, b.y1);
  scores[o] = score;
  valid[o] = ok ? 1 : 0;
}

// FastRCNNOutputLayers.predict_boxes / predict_probs (fast_rcnn.py:400-440): softmax over K+1 logits, per-class decode with
// weights (10,10,5,5), clip.  One thread per (roi, class) -> class-major padded scores (N,K,P) / boxes (N,K,P,4) (zero-initialised).
__global__ void __launch_bounds__(128) det_decode_kernel(const float* __restrict__ cls, const float* __restrict__ reg, const float4* __restrict__ props,
                                                         const int* __restrict__ offs /*(N+1)*/, const float* __restrict__ limits, int R, int N,
                                                         int P, int K, float* __restrict__ scores, float4* __restrict__ boxes) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= R * K) return;
  const int r = t / K, k = t % K;
  int n = 0;
  while (n + 1 < N && r >= offs[n + 1]) n++;
  const int rank = r - offs[n];
  const float* c = cls + (long)r * (K + 1);
  float mx = c[0];
  for (int j = 1; j <= K; j++) mx = fmaxf(mx, c[j]);
  float sum = 0.0f;
  for (int j = 0; j <= K; j++) sum += expf(c[j] - mx);
  const float score = expf(c[k] - mx) / sum;
  bool ok = true;
  for (int j = 0; j <= K; j++) ok = ok && isfinite(c[j]);
  const float* d = reg + (long)r * 4 * K;
  for (int j = 0; j < 4 * K; j++) ok = ok && isfinite(d[j]);                 // the reference drops the row if ANY of its boxes is non-finite
  Box4 b = apply_deltas(__ldg(props + r), d[4 * k], d[4 * k + 1], d[4 * k + 2], d[4 * k + 3], 10.f, 10.f, 5.f, 5.f);
  if (ok) {                                                                  // ... which needs every class's decode
    const float4 pb = __ldg(props + r);
    for (int j = 0; j < K; j++) {
      const Box4 bj = apply_deltas(pb, d[4 * j], d[4 * j + 1], d[4 * j + 2], d[4 * j + 3], 10.f, 10.f, 5.f, 5.f);
      ok = ok && isfinite(bj.x0) && isfinite(bj.y0) && isfinite(bj.x1) && isfinite(bj.y1);
    }
  }
  const float4 lim = __ldg(reinterpret_cast<const float4*>(limits) + n);
  const long o = ((long)n * K + k) * P + rank;
  scores[o] = ok ? score : 0.0f;
  boxes[o] = make_float4(clampf(b.x0, lim.x), clampf(b.y0, lim.y), clampf(b.x1, lim.z), clampf(b.y1, lim.w));
}

}  // namespace

extern "C" int pvo_rpn_decode(const void* const* rpn_out, const long* const* topk_idx, const int* hs, const int* ws, const int* ks,
                              const int* strides, const float* cell_anchors /*host (5,3,4)*/, const float* limits, int N, int P,
                              float* boxes, float* scores, unsigned char* valid, void* stream) {
  if (N == 0 || P == 0) return PVO_OK;
  RpnLevels L;
  for (int l = 0; l < 5; l++) {
    L.out[l] = (const __half*)rpn_out[l]; L.idx[l] = topk_idx[l];
    L.h[l] = hs[l]; L.w[l] = ws[l]; L.k[l] = ks[l]; L.stride[l] = strides[l];
    for (int a = 0; a < 3; a++)
      for (int c = 0; c < 4; c++) L.cell[l][a][c] = cell_anchors[(l * 3 + a) * 4 + c];
  }
  rpn_decode_kernel<<<dim3((P + 127) / 128, 5, N), 128, 0, (cudaStream_t)stream>>>(L, limits, P, (float4*)boxes, scores, valid);
  return pvo_check_launch();
}

extern "C" int pvo_det_decode(const float* cls, const float* reg, const float* props, const int* offsets, const float* limits, int R, int N,
                              int P, int K, float* scores, float* boxes, void* stream) {
  if (R == 0) return PVO_OK;
  if (K < 1 || K > 16) return PVO_ERR_INVALID;
  det_decode_kernel<<<(R * K + 127) / 128, 128, 0, (cudaStream_t)stream>>>(cls, reg, (const float4*)props, offsets, limits, R, N, P, K, scores,
                                                                            (float4*)boxes);
  return pvo_check_launch();
}

extern "C" int pvo_roi_align_fpn(const void* const* feats, const int* Hs, const int* Ws, int C, const float* rois, int R, int S, void* out,
                                 int dtype, void* stream) {
  if (R == 0) return PVO_OK;
  if (C % (dtype == PVO_DTYPE_F32 ? 4 : 8) || S < 1) return PVO_ERR_INVALID;
  Pyr p;
  for (int l = 0; l < 4; l++) { p.f[l] = feats[l]; p.H[l] = Hs[l]; p.W[l] = Ws[l]; }
  cudaStream_t s = (cudaStream_t)stream;
  const unsigned grid = (unsigned)((long)R * S);
  if (dtype == PVO_DTYPE_F16) roi_align_fpn_kernel<__half><<<grid, 128, 0, s>>>(p, C, rois, S, (__half*)out);
  else if (dtype == PVO_DTYPE_F32) roi_align_fpn_kernel<float><<<grid, 128, 0, s>>>(p, C, rois, S, (float*)out);
  else return PVO_ERR_UNSUPPORTED;
  return pvo_check_launch();
}

// scratch: G * n * 32 uint32 (the suppression bitmasks)
extern "C" int pvo_nms_groups(const float* boxes, const unsigned char* valid, int G, int n, float thr, unsigned char* keep,
                              unsigned int* scratch, void* stream) {
  if (G == 0 || n == 0) return PVO_OK;
  if (n > 1024 || !scratch) return PVO_ERR_UNSUPPORTED;
  cudaStream_t s = (cudaStream_t)stream;
  nms_mask_kernel<<<dim3(G, (n + 7) / 8), 256, 0, s>>>((const float4*)boxes, n, thr, scratch);
  const size_t smem = (size_t)n * 32 * 4;
  cudaFuncSetAttribute(nms_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  nms_sweep_kernel<<<G, 1024, smem, s>>>(scratch, valid, n, keep);
  return pvo_check_launch();
}
