// Post-processing of the update operator's outputs inside one frame-graph iteration, including the panoptic
// segment filter.
//
// Replaces FactorGraph.update lines :250-283,:303-304 (VO_Module/droid_slam/factor_graph.py):
//   target_cam = coords1 + delta[...,0:2]
//   raw_mask  += delta_m ;  bin_mask = sigmoid(raw_mask) >= 0.5
//   [segm_filter, :256-276]  for every (edge, panoptic label != 0): if #dynamic / #pixels > thresh the whole segment
//                            is forced dynamic.  The reference does this on the HOST every iteration (two D2H copies,
//                            two np.unique over E*HW keys, a Python loop per label); here it is integer counting in
//                            a per-edge open-addressing table in global memory, no host round trip.
//   delta_dy   = delta[...,2:4] * (1 - bin_mask)
//   weight     = sigmoid(weight_logit + 10 * (1 - bin_mask))
//   full_flow  = coords1 + delta_dy - coords0
// head_out is the fused output of the four 128->2 heads: (M, 8) = [delta | delta_dy | weight | delta_mask].
#include "common.cuh"

namespace {

constexpr int kSlots = 1024;   // per-edge table slots (key, total, dynamic) -- a frame holds O(100) panoptic ids

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ uint32_t hash_label(int label) {
  uint32_t h = (uint32_t)label * 2654435761u;
  return (h >> 16) & (kSlots - 1);
}

// pass A: raw_mask += delta_m; count (edge,label) totals and dynamic pixels
__global__ void __launch_bounds__(256) post_count_kernel(const float* __restrict__ head, const int* __restrict__ segm,
                                                         float* __restrict__ raw_mask, int* __restrict__ table, int HW,
                                                         long M) {
  const long m = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  const float r0 = raw_mask[2 * m] + head[8 * m + 6], r1 = raw_mask[2 * m + 1] + head[8 * m + 7];
  raw_mask[2 * m] = r0;
  raw_mask[2 * m + 1] = r1;
  const bool dyn = !(sigmoidf_(r0) >= 0.5f) || !(sigmoidf_(r1) >= 0.5f);
  const int label = segm[m];
  const long e = m / HW;
  int* tab = table + e * (3L * kSlots);
  uint32_t s = hash_label(label);
  const int key = label + 1;                       // 0 marks an empty slot
  for (int probe = 0; probe < kSlots; probe++) {
    const int old = atomicCAS(&tab[3 * s], 0, key);
    if (old == 0 || old == key) {
      atomicAdd(&tab[3 * s + 1], 1);
      if (dyn) atomicAdd(&tab[3 * s + 2], 1);
      return;
    }
    s = (s + 1) & (kSlots - 1);
  }
}

// pass B (or the only pass when segm == nullptr)
template <bool SEG>
__global__ void __launch_bounds__(256) post_apply_kernel(const float2* __restrict__ coords1, const float* __restrict__ head,
                                                         const int* __restrict__ segm, const int* __restrict__ table,
                                                         double thresh, float2* __restrict__ target,
                                                         float2* __restrict__ weight, float2* __restrict__ raw_mask,
                                                         float2* __restrict__ delta_dy, float2* __restrict__ full_flow,
                                                         int HW, int wd, long M) {
  const long m = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  const float4 h0 = reinterpret_cast<const float4*>(head)[2 * m];       // delta, delta_dy
  const float4 h1 = reinterpret_cast<const float4*>(head)[2 * m + 1];   // weight logits, delta_m
  const float2 c1 = coords1[m];
  float2 rm = raw_mask[m];
  if (!SEG) {
    rm.x += h1.z; rm.y += h1.w;
    raw_mask[m] = rm;
  }
  bool b0 = sigmoidf_(rm.x) >= 0.5f, b1 = sigmoidf_(rm.y) >= 0.5f;
  if (SEG) {
    const int label = segm[m];
    if (label != 0) {                              // factor_graph.py:268: label % 1e6 == 0 is skipped
      const long e = m / HW;
      const int* tab = table + e * (3L * kSlots);
      uint32_t s = hash_label(label);
      const int key = label + 1;
      for (int probe = 0; probe < kSlots; probe++) {
        const int k = tab[3 * s];
        if (k == key) {
          const int tot = tab[3 * s + 1], dy = tab[3 * s + 2];
          if (dy > 0 && (double)dy / (double)tot > thresh) { b0 = false; b1 = false; }
          break;
        }
        if (k == 0) break;
        s = (s + 1) & (kSlots - 1);
      }
    }
  }
  const float m0 = b0 ? 1.0f : 0.0f, m1 = b1 ? 1.0f : 0.0f;
  target[m] = make_float2(c1.x + h0.x, c1.y + h0.y);
  const float2 dd = make_float2(h0.z * (1.0f - m0), h0.w * (1.0f - m1));
  delta_dy[m] = dd;
  weight[m] = make_float2(sigmoidf_(h1.x + (1.0f - m0) * 10.0f), sigmoidf_(h1.y + (1.0f - m1) * 10.0f));
  if (full_flow) {
    const int pix = (int)(m % HW);
    const int y = pix / wd, x = pix - y * wd;
    full_flow[m] = make_float2(c1.x + dd.x - (float)x, c1.y + dd.y - (float)y);
  }
}

}  // namespace

extern "C" long pvo_update_post_workspace_ints(int E) { return 3L * kSlots * (E > 0 ? E : 1); }

// coords1 (E,ht,wd,2); head_out (E*ht*wd, 8) float32; segm (E,ht,wd) int32 or NULL (filter off);
// target / weight / raw_mask / delta_dy / full_flow: (E,ht,wd,2) float32; raw_mask is updated in place.
extern "C" int pvo_update_post(const float* coords1, const float* head_out, const int* segm, double thresh,
                               float* target, float* weight, float* raw_mask, float* delta_dy, float* full_flow,
                               int* seg_workspace, long seg_workspace_ints, int E, int ht, int wd, void* stream) {
  if (E < 0 || ht <= 0 || wd <= 0) return PVO_ERR_INVALID;
  if (E == 0) return PVO_OK;
  const int HW = ht * wd;
  const long M = (long)E * HW;
  cudaStream_t s = (cudaStream_t)stream;
  const int grid = ceil_div(M, 256);
  if (segm) {
    if (!seg_workspace || seg_workspace_ints < 3L * kSlots * E) return PVO_ERR_WORKSPACE;
    cudaMemsetAsync(seg_workspace, 0, sizeof(int) * 3L * kSlots * E, s);
    post_count_kernel<<<grid, 256, 0, s>>>(head_out, segm, raw_mask, seg_workspace, HW, M);
    post_apply_kernel<true><<<grid, 256, 0, s>>>((const float2*)coords1, head_out, segm, seg_workspace, thresh,
                                                 (float2*)target, (float2*)weight, (float2*)raw_mask, (float2*)delta_dy,
                                                 (float2*)full_flow, HW, wd, M);
  } else {
    post_apply_kernel<false><<<grid, 256, 0, s>>>((const float2*)coords1, head_out, nullptr, nullptr, thresh,
                                                  (float2*)target, (float2*)weight, (float2*)raw_mask, (float2*)delta_dy,
                                                  (float2*)full_flow, HW, wd, M);
  }
  return pvo_check_launch();
}
