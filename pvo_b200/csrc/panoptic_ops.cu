// Mask pasting and the panoptic merge of the Panoptic-FPN post-processing.  References under VPS_Module/detectron2/:
//   paste_masks_in_image / _do_paste_mask      layers/mask_ops.py:17-152 (F.grid_sample of every 28x28 mask over the image)
//   combine_semantic_and_instance_outputs      modeling/meta_arch/panoptic_fpn.py:497-588 (a Python loop with .item()
//                                              syncs per instance and per label)
// Here: one kernel pastes all masks (only pixels inside a detection's box are sampled, the rest are written as zeros),
// and one thread-block CLUSTER per image (8 CTAs, each owning a strip of rows) runs the sequential merge on chip: per
// instance one pass over the box region for {area, overlap}, one cluster barrier + distributed-shared-memory exchange
// for the accept decision, one pass to paint.  No host synchronisation.
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int kCluster = 8;
constexpr int kMaxSem = 64;

// ---- paste: out[r][y][x] = bilinear(prob[r], grid(y, x)) >= thr, the arithmetic of mask_ops.py:46-71 + at::grid_sampler_2d
// (bilinear, zeros padding, align_corners=False), operation for operation in fp32.
__device__ __forceinline__ float unnormalize(float c, int size) { return fmaf(__fadd_rn(c, 1.0f), (float)size, -1.0f) / 2.0f; }

__device__ __forceinline__ bool paste_px(const float* __restrict__ P, int S, float gx, float gy, float thr) {
  const float ix = unnormalize(gx, S), iy = unnormalize(gy, S);
  const float fx = floorf(ix), fy = floorf(iy);
  const int x0 = (int)fx, y0 = (int)fy, x1 = x0 + 1, y1 = y0 + 1;
  const float nw = (fx + 1.0f - ix) * (fy + 1.0f - iy), ne = (ix - fx) * (fy + 1.0f - iy);
  const float sw = (fx + 1.0f - ix) * (iy - fy), se = (ix - fx) * (iy - fy);
  float acc = 0.0f;
  const bool xin0 = x0 >= 0 && x0 < S, xin1 = x1 >= 0 && x1 < S, yin0 = y0 >= 0 && y0 < S, yin1 = y1 >= 0 && y1 < S;
  if (yin0 && xin0) acc = fmaf(P[y0 * S + x0], nw, acc);
  if (yin0 && xin1) acc = fmaf(P[y0 * S + x1], ne, acc);
  if (yin1 && xin0) acc = fmaf(P[y1 * S + x0], sw, acc);
  if (yin1 && xin1) acc = fmaf(P[y1 * S + x1], se, acc);
  return acc >= thr;
}

// grid (ceil(H*W/4 / 256), R); thread <-> 4 consecutive pixels of a row (W % 4 == 0), one 32-bit store
__global__ void __launch_bounds__(256) paste_masks_kernel(const float* __restrict__ prob /*[R][S][S]*/, const float* __restrict__ boxes /*[R][4]*/,
                                                          int S, int H, int W, float thr, unsigned char* __restrict__ out /*[R][H][W]*/) {
  extern __shared__ float P[];
  const int r = blockIdx.y;
  for (int i = threadIdx.x; i < S * S; i += blockDim.x) P[i] = prob[(long)r * S * S + i];
  __syncthreads();
  const float bx0 = boxes[r * 4 + 0], by0 = boxes[r * 4 + 1], bx1 = boxes[r * 4 + 2], by1 = boxes[r * 4 + 3];
  const float bw = __fsub_rn(bx1, bx0), bh = __fsub_rn(by1, by0);
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= H * W / 4) return;
  const int y = (q * 4) / W, x = (q * 4) % W;
  unsigned int bits = 0;
  // a pixel centre outside the box samples beyond the border texels' half-cell: the interpolated value is < 0.5 * max
  const float cy = (float)y + 0.5f;
  if (cy >= by0 - 1.0f && cy <= by1 + 1.0f && (float)x + 4.5f >= bx0 - 1.0f && (float)x + 0.5f <= bx1 + 1.0f) {
    const float gy = __fsub_rn(__fmul_rn(__fdiv_rn(__fsub_rn(cy, by0), bh), 2.0f), 1.0f);
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const float cx = (float)(x + k) + 0.5f;
      const float gx = __fsub_rn(__fmul_rn(__fdiv_rn(__fsub_rn(cx, bx0), bw), 2.0f), 1.0f);
      if (paste_px(P, S, gx, gy, thr)) bits |= 1u << (8 * k);
    }
  }
  reinterpret_cast<unsigned int*>(out + (long)r * H * W)[q] = bits;
}

// the same paste, one pixel per thread: image widths that are not a multiple of 4 (a packed word would straddle two rows)
__global__ void __launch_bounds__(256) paste_masks_px_kernel(const float* __restrict__ prob, const float* __restrict__ boxes, int S, int H, int W,
                                                             float thr, unsigned char* __restrict__ out) {
  extern __shared__ float P[];
  const int r = blockIdx.y;
  for (int i = threadIdx.x; i < S * S; i += blockDim.x) P[i] = prob[(long)r * S * S + i];
  __syncthreads();
  const float bx0 = boxes[r * 4 + 0], by0 = boxes[r * 4 + 1], bx1 = boxes[r * 4 + 2], by1 = boxes[r * 4 + 3];
  const float bw = __fsub_rn(bx1, bx0), bh = __fsub_rn(by1, by0);
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= H * W) return;
  const int y = q / W, x = q - y * W;
  const float cy = (float)y + 0.5f, cx = (float)x + 0.5f;
  bool on = false;
  if (cy >= by0 - 1.0f && cy <= by1 + 1.0f && cx >= bx0 - 1.0f && cx <= bx1 + 1.0f) {
    const float gy = __fsub_rn(__fmul_rn(__fdiv_rn(__fsub_rn(cy, by0), bh), 2.0f), 1.0f);
    const float gx = __fsub_rn(__fmul_rn(__fdiv_rn(__fsub_rn(cx, bx0), bw), 2.0f), 1.0f);
    on = paste_px(P, S, gx, gy, thr);
  }
  out[(long)r * H * W + q] = on ? 1 : 0;
}

// ---- merge
// info_out[n][k] = {segment id, category, isthing, area, instance index}; counts_out[n] = number of segments.
__global__ void __cluster_dims__(kCluster, 1, 1) __launch_bounds__(1024)
    panoptic_combine_kernel(const unsigned char* __restrict__ masks /*[R][H*W] bool*/, const float* __restrict__ boxes /*[R][4] or null*/,
                            const float* __restrict__ scores, const long* __restrict__ classes, const long* __restrict__ order,
                            const int* __restrict__ inst_off /*[N+1]*/, const long* __restrict__ sem /*[N][H*W] argmax*/,
                            int* __restrict__ pan /*[N][H*W]*/, int* __restrict__ info_out, int* __restrict__ counts_out, int H, int W,
                            int num_sem, float overlap_thr, int stuff_area_thr, float score_thr, int thing_cat0, int max_info) {
  cg::cluster_group cluster = cg::this_cluster();
  __shared__ int sh[64];
  __shared__ int slot[2][2];                     // {area, overlap} of this CTA's strip, double buffered by instance parity
  __shared__ int hist[kMaxSem], hist_all[kMaxSem];
  const int rank = (int)cluster.block_rank();
  const int n = blockIdx.x / kCluster;
  const int rows = (H + kCluster - 1) / kCluster;
  const int ys = min(rank * rows, H), ye = min(ys + rows, H);            // this CTA's strip of rows
  const long HW = (long)H * W;
  int* P = pan + n * HW;
  for (long i = (long)ys * W + threadIdx.x; i < (long)ye * W; i += blockDim.x) P[i] = 0;
  __syncthreads();
  int cur = 0, it = 0;
  const int r0 = inst_off[n], r1 = inst_off[n + 1];
  const bool vec = (W & 3) == 0;                            // 4 pixels per "word" when rows are 4-aligned, else 1
  const int step = vec ? 4 : 1;
  // An instance's mask words for this strip live in registers: they are fetched while the PREVIOUS instance is being
  // reduced / decided / painted (the fetch does not depend on any decision), and reused for the paint pass.
  constexpr int kBuf = 10;
  struct Region { const unsigned char* M; int r, x0, y0, rw, cnt; bool live; };
  auto region = [&](int k) {
    Region g;
    g.live = k < r1;
    g.r = g.live ? r0 + (int)order[k] : 0;                  // order[] holds per-image ranks sorted by descending score
    g.live = g.live && !(scores[g.r] < score_thr);          // the reference loop breaks at the first score below the threshold
    g.M = masks + (long)g.r * HW;
    int y0 = ys, y1 = ye, x0 = 0, x1 = W;                   // region that can hold mask pixels (pixel centres within the box)
    if (boxes && g.live) {
      x0 = max(0, (int)floorf(boxes[g.r * 4 + 0] - 1.5f)); y0 = max(ys, (int)floorf(boxes[g.r * 4 + 1] - 1.5f));
      x1 = min(W, (int)ceilf(boxes[g.r * 4 + 2] + 1.5f)); y1 = min(ye, (int)ceilf(boxes[g.r * 4 + 3] + 1.5f));
    }
    if (vec) { x0 &= ~3; x1 = min(W, (x1 + 3) & ~3); }
    g.x0 = x0; g.y0 = y0;
    g.rw = max(x1 - x0, 0) / step;
    g.cnt = g.live ? g.rw * max(y1 - y0, 0) : 0;
    return g;
  };
  auto pixel = [&](const Region& g, int i) { return (long)(g.y0 + i / g.rw) * W + g.x0 + (i % g.rw) * step; };
  auto word = [&](const Region& g, int i) -> unsigned int {
    const long px = pixel(g, i);
    return vec ? __ldg(reinterpret_cast<const unsigned int*>(g.M + px)) : (unsigned int)g.M[px];
  };
  auto count = [&](unsigned int m4, long px, int& area, int& inter) {
    if (!m4) return;
    if (vec) {
      const int4 p4 = *reinterpret_cast<const int4*>(P + px);
      const int b0 = (m4 & 0xffu) != 0, b1 = (m4 & 0xff00u) != 0, b2 = (m4 & 0xff0000u) != 0, b3 = (m4 & 0xff000000u) != 0;
      area += b0 + b1 + b2 + b3;
      inter += (b0 && p4.x > 0) + (b1 && p4.y > 0) + (b2 && p4.z > 0) + (b3 && p4.w > 0);
    } else {
      area += 1;
      inter += P[px] > 0;
    }
  };
  auto paint = [&](unsigned int m4, long px, int id) {
    if (!m4) return;
    if (vec) {
      int4 p4 = *reinterpret_cast<const int4*>(P + px);
      if ((m4 & 0xffu) && p4.x == 0) p4.x = id;
      if ((m4 & 0xff00u) && p4.y == 0) p4.y = id;
      if ((m4 & 0xff0000u) && p4.z == 0) p4.z = id;
      if ((m4 & 0xff000000u) && p4.w == 0) p4.w = id;
      *reinterpret_cast<int4*>(P + px) = p4;
    } else if (P[px] == 0) {
      P[px] = id;
    }
  };
  int k = r0;
  Region nx = region(k);
  unsigned int nbuf[kBuf];
#pragma unroll
  for (int j = 0; j < kBuf; j++) { const int i = threadIdx.x + j * 1024; nbuf[j] = i < nx.cnt ? word(nx, i) : 0u; }
  while (nx.live) {
    const Region g = nx;
    unsigned int buf[kBuf];
#pragma unroll
    for (int j = 0; j < kBuf; j++) buf[j] = nbuf[j];
    nx = region(++k);
#pragma unroll
    for (int j = 0; j < kBuf; j++) { const int i = threadIdx.x + j * 1024; nbuf[j] = i < nx.cnt ? word(nx, i) : 0u; }
    int area = 0, inter = 0;
#pragma unroll
    for (int j = 0; j < kBuf; j++) { const int i = threadIdx.x + j * 1024; if (i < g.cnt) count(buf[j], pixel(g, i), area, inter); }
    for (int i = threadIdx.x + kBuf * 1024; i < g.cnt; i += 1024) count(word(g, i), pixel(g, i), area, inter);   // oversized regions
    area = __reduce_add_sync(0xffffffffu, area);
    inter = __reduce_add_sync(0xffffffffu, inter);
    if ((threadIdx.x & 31) == 0) { sh[(threadIdx.x >> 5) * 2] = area; sh[(threadIdx.x >> 5) * 2 + 1] = inter; }
    __syncthreads();
    if (threadIdx.x == 0) {
      int ta = 0, ti = 0;
      for (int w = 0; w < 32; w++) { ta += sh[2 * w]; ti += sh[2 * w + 1]; }
      slot[it & 1][0] = ta; slot[it & 1][1] = ti;
    }
    cluster.sync();
    area = 0; inter = 0;
    for (int c = 0; c < kCluster; c++) {
      const int* rs = cluster.map_shared_rank(&slot[it & 1][0], c);
      area += rs[0]; inter += rs[1];
    }
    it++;
    if (area == 0) continue;
    if ((double)inter / (double)area > (double)overlap_thr) continue;       // the reference divides Python numbers (doubles)
    cur++;
    const int cat = thing_cat0 + (int)classes[g.r];
    const int id = cat * 10000 + cur;
#pragma unroll
    for (int j = 0; j < kBuf; j++) { const int i = threadIdx.x + j * 1024; if (i < g.cnt) paint(buf[j], pixel(g, i), id); }
    for (int i = threadIdx.x + kBuf * 1024; i < g.cnt; i += 1024) paint(word(g, i), pixel(g, i), id);
    if (rank == 0 && threadIdx.x == 0 && cur <= max_info) {
      int* o = info_out + ((long)n * max_info + cur - 1) * 5;
      o[0] = id; o[1] = cat; o[2] = 1; o[3] = area - inter; o[4] = g.r - r0;
    }
    __syncthreads();
  }
  // stuff: per label the area not taken by things (label 0 is the "thing" placeholder class, :560-561)
  const long* S = sem + n * HW;
  for (int i = threadIdx.x; i < kMaxSem; i += blockDim.x) hist[i] = 0;
  __syncthreads();
  for (long i = (long)ys * W + threadIdx.x; i < (long)ye * W; i += blockDim.x) {
    const int lab = (int)S[i];
    if (lab > 0 && lab < num_sem && P[i] == 0) atomicAdd(&hist[lab], 1);
  }
  cluster.sync();
  for (int lab = threadIdx.x; lab < num_sem; lab += blockDim.x) {
    int t = 0;
    for (int c = 0; c < kCluster; c++) t += cluster.map_shared_rank(&hist[0], c)[lab];
    hist_all[lab] = t;
  }
  __syncthreads();
  for (long i = (long)ys * W + threadIdx.x; i < (long)ye * W; i += blockDim.x) {
    const int lab = (int)S[i];
    if (lab > 0 && lab < num_sem && P[i] == 0 && hist_all[lab] >= stuff_area_thr) P[i] = lab * 10000;   // stuff_id2cat: identity on trainIds
  }
  if (rank == 0 && threadIdx.x == 0) {
    for (int lab = 1; lab < num_sem; lab++) {
      if (hist_all[lab] < stuff_area_thr || hist_all[lab] == 0) continue;
      cur++;
      if (cur <= max_info) {
        int* o = info_out + ((long)n * max_info + cur - 1) * 5;
        o[0] = lab * 10000; o[1] = lab; o[2] = 0; o[3] = hist_all[lab]; o[4] = -1;
      }
    }
    counts_out[n] = cur;
  }
  cluster.sync();                                // keep every CTA's shared memory alive until all remote reads are done
}

}  // namespace

// paste_masks_in_image for R detections of images of one size: prob (R,S,S) float32 probabilities of the predicted class,
// boxes (R,4) float32, out (R,H,W) uint8 (0/1).  Any width (4 pixels per thread when W % 4 == 0).
extern "C" int pvo_paste_masks(const float* prob, const float* boxes, int R, int S, int H, int W, float thr, void* out, void* stream) {
  if (R == 0) return PVO_OK;
  if (S < 1 || S > 64) return PVO_ERR_INVALID;
  if (W % 4 == 0) {
    dim3 grid(ceil_div((long)H * W / 4, 256), R);
    paste_masks_kernel<<<grid, 256, S * S * sizeof(float), (cudaStream_t)stream>>>(prob, boxes, S, H, W, thr, (unsigned char*)out);
  } else {
    dim3 grid(ceil_div((long)H * W, 256), R);
    paste_masks_px_kernel<<<grid, 256, S * S * sizeof(float), (cudaStream_t)stream>>>(prob, boxes, S, H, W, thr, (unsigned char*)out);
  }
  return pvo_check_launch();
}

// combine_semantic_and_instance_outputs for a batch: masks (R,H*W) uint8 (all images; image n owns rows inst_off[n]..inst_off[n+1]),
// boxes (R,4) or NULL (restricts each instance's scan to its box; masks pasted by pvo_paste_masks are zero outside it),
// order = per-image rank lists (descending score) concatenated, sem (N,H*W) int64 argmax labels.
// Outputs: pan (N,H*W) int32, info (N,max_info,5) int32, counts (N) int32.
extern "C" int pvo_panoptic_combine(const void* masks, const float* boxes, const float* scores, const long* classes, const long* order,
                                    const int* inst_off, const long* sem, int* pan, int* info_out, int* counts_out, int N, int H, int W,
                                    int num_sem, float overlap_thr, int stuff_area_thr, float score_thr, int thing_cat0, int max_info,
                                    void* stream) {
  if (N == 0) return PVO_OK;
  if (num_sem > kMaxSem) return PVO_ERR_UNSUPPORTED;
  panoptic_combine_kernel<<<N * kCluster, 1024, 0, (cudaStream_t)stream>>>((const unsigned char*)masks, boxes, scores, classes, order, inst_off,
                                                                            sem, pan, info_out, counts_out, H, W, num_sem, overlap_thr,
                                                                            stuff_area_thr, score_thr, thing_cat0, max_info);
  return pvo_check_launch();
}
