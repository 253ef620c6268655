// Reprojection of every pixel of frame ii[e] into frame jj[e], plus the 8 motion-feature channels the update
// operator's flow encoder consumes.
//
// Replaces DepthVideo.reproject (VO_Module/droid_slam/depth_video.py:154-163) ->
// geom/projective_ops.py:102-130 (iproj :21-40, Gij = G_j * G_i^-1 via lietorch, actp :72-75, proj :43-55,
// valid :117) and the motion-feature construction + clamp of FactorGraph.update (factor_graph.py:231-237):
//   motn = clamp([target-coords0 | target-coords0+delta_dy | target-coords1 | raw_mask], -64, 64)
// The reference runs ~30 small torch kernels for this; here it is one pass: 40 B read + (8..24) B written per
// pixel per edge, HBM-bound.  frame_distance (droid_kernels.cu:497-636) shares the geometry and lives here too.
#include "se3.cuh"

namespace {

template <typename MT>
__global__ void __launch_bounds__(256) reproject_kernel(
    const float* __restrict__ poses, const float* __restrict__ disps, const float* __restrict__ intrinsics,
    const long* __restrict__ ii, const long* __restrict__ jj, float2* __restrict__ coords, float* __restrict__ valid,
    const float2* __restrict__ target, const float2* __restrict__ delta_dy, const float2* __restrict__ raw_mask,
    MT* __restrict__ motn, int E, int ht, int wd) {
  const int e = blockIdx.y;
  const int HW = ht * wd;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int ix = (int)ii[e], jx = (int)jj[e];
  Vec3 tij; Quat qij;
  lt_rel_pose(poses + 7 * ix, poses + 7 * jx, tij, qij);
  if (k >= HW) return;
  const float fxi = intrinsics[4 * ix + 0], fyi = intrinsics[4 * ix + 1], cxi = intrinsics[4 * ix + 2],
              cyi = intrinsics[4 * ix + 3];
  const float fxj = intrinsics[4 * jx + 0], fyj = intrinsics[4 * jx + 1], cxj = intrinsics[4 * jx + 2],
              cyj = intrinsics[4 * jx + 3];
  const int i = k / wd, j = k - i * wd;
  const float u = (float)j, v = (float)i;
  const float d = disps[(long)ix * HW + k];
  const Vec3 X0{(u - cxi) / fxi, (v - cyi) / fyi, 1.0f};
  const Vec3 r = q_rot(qij, X0);
  const float X = r.x + tij.x * d, Y = r.y + tij.y * d, Zr = r.z + tij.z * d;
  const float Z = (Zr < 0.1f) ? 1.0f : Zr;                 // proj(): Z < 0.5*MIN_DEPTH -> 1
  const float dinv = 1.0f / Z;
  const float x1 = fxj * (X * dinv) + cxj;
  const float y1 = fyj * (Y * dinv) + cyj;
  const long p = (long)e * HW + k;
  coords[p] = make_float2(x1, y1);
  if (valid) valid[p] = (Zr > 0.2f) ? 1.0f : 0.0f;         // X0.z == 1 > MIN_DEPTH always
  if (motn) {
    const float2 t = target[p], dd = delta_dy[p], rm = raw_mask[p];
    float m[8] = {t.x - u, t.y - v, t.x - u + dd.x, t.y - v + dd.y, t.x - x1, t.y - y1, rm.x, rm.y};
    MT* o = motn + p * 8;
#pragma unroll
    for (int c = 0; c < 8; c++) o[c] = (MT)fminf(fmaxf(m[c], -64.0f), 64.0f);
  }
}

// frame_distance_kernel (droid_kernels.cu:497-636): mean induced-flow magnitude, beta-mix of full SE3 and
// translation-only flow; 1000 when fewer than 75% of the weighted pixels are in front of the camera.
__global__ void __launch_bounds__(256) frame_distance_kernel(
    const float* __restrict__ poses, const float* __restrict__ disps, const float* __restrict__ intrinsics,
    const long* __restrict__ ii, const long* __restrict__ jj, float* __restrict__ dist, int ht, int wd, float beta) {
  const int e = blockIdx.x;
  const int HW = ht * wd;
  const int ix = (int)ii[e], jx = (int)jj[e];
  const float fx = intrinsics[0], fy = intrinsics[1], cx = intrinsics[2], cy = intrinsics[3];
  Vec3 tij; Quat qij;
  dk_rel_pose(poses + 7 * ix, poses + 7 * jx, tij, qij);
  float accum = 0.f, valid = 0.f, total = 0.f;
  for (int k = threadIdx.x; k < HW; k += blockDim.x) {
    const int i = k / wd, j = k - i * wd;
    const float u = (float)j, v = (float)i;
    const float d = disps[(long)ix * HW + k];
    const Vec3 X0{(u - cx) / fx, (v - cy) / fy, 1.0f};
    const Vec3 r = q_rot(qij, X0);
    float X = r.x + d * tij.x, Y = r.y + d * tij.y, Z = r.z + d * tij.z;
    float du = fx * (X / Z) + cx - u, dv = fy * (Y / Z) + cy - v;
    total += beta;
    if (Z > 0.25f) { accum += beta * sqrtf(du * du + dv * dv); valid += beta; }
    X = X0.x + d * tij.x; Y = X0.y + d * tij.y; Z = 1.0f + d * tij.z;
    du = fx * (X / Z) + cx - u; dv = fy * (Y / Z) + cy - v;
    total += (1 - beta);
    if (Z > 0.25f) { accum += (1 - beta) * sqrtf(du * du + dv * dv); valid += (1 - beta); }
  }
  __shared__ float red[3][8];
  accum = warp_sum(accum); valid = warp_sum(valid); total = warp_sum(total);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { red[0][warp] = accum; red[1][warp] = valid; red[2][warp] = total; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0, vv = 0, t = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) { a += red[0][w]; vv += red[1][w]; t += red[2][w]; }
    dist[e] = (vv / (t + 1e-8f) < 0.75f) ? 1000.0f : a / vv;
  }
}

}  // namespace

extern "C" int pvo_reproject(const float* poses, const float* disps, const float* intrinsics, const long* ii,
                             const long* jj, float* coords, float* valid, const float* target,
                             const float* delta_dy, const float* raw_mask, void* motn, int motn_dtype, int E, int ht,
                             int wd, void* stream) {
  if (E < 0 || ht <= 0 || wd <= 0) return PVO_ERR_INVALID;
  if (E == 0) return PVO_OK;
  if (motn && (!target || !delta_dy || !raw_mask)) return PVO_ERR_INVALID;
  dim3 grid(ceil_div((long)ht * wd, 256), E);
  cudaStream_t s = (cudaStream_t)stream;
  if (motn_dtype == PVO_DTYPE_F16)
    reproject_kernel<__half><<<grid, 256, 0, s>>>(poses, disps, intrinsics, ii, jj, (float2*)coords, valid,
                                                  (const float2*)target, (const float2*)delta_dy,
                                                  (const float2*)raw_mask, (__half*)motn, E, ht, wd);
  else
    reproject_kernel<float><<<grid, 256, 0, s>>>(poses, disps, intrinsics, ii, jj, (float2*)coords, valid,
                                                 (const float2*)target, (const float2*)delta_dy,
                                                 (const float2*)raw_mask, (float*)motn, E, ht, wd);
  return pvo_check_launch();
}

extern "C" int pvo_frame_distance(const float* poses, const float* disps, const float* intrinsics, const long* ii,
                                  const long* jj, float* dist, int n, int ht, int wd, float beta, void* stream) {
  if (n < 0 || ht <= 0 || wd <= 0) return PVO_ERR_INVALID;
  if (n == 0) return PVO_OK;
  frame_distance_kernel<<<n, 256, 0, (cudaStream_t)stream>>>(poses, disps, intrinsics, ii, jj, dist, ht, wd, beta);
  return pvo_check_launch();
}
