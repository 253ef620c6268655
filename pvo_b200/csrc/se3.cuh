// SE3 helpers shared by the geometry kernels.  Pose layout (tx,ty,tz,qx,qy,qz,qw) as in lietorch.
//
// Two flavours, because the reference has two:
//  * "lt_*"  follow lietorch (thirdparty/lietorch/lietorch/include/so3.h:31-60, se3.h:36-56): the quaternion is
//            normalised on EVERY construction (so3.h:31-37).  Used by reproject (DepthVideo.reproject ->
//            geom/projective_ops.py:102-130).
//  * "dk_*"  follow VO_Module/src/droid_kernels.cu:58-176,856-874: quaternions used as stored.  Used by the BA.
#pragma once
#include "common.cuh"

struct Quat { float x, y, z, w; };
struct Vec3 { float x, y, z; };

__device__ __forceinline__ Quat q_normalize(Quat q) {
  const float n = sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
  return Quat{q.x / n, q.y / n, q.z / n, q.w / n};
}

__device__ __forceinline__ Quat q_mul(Quat a, Quat b) {
  return Quat{a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
              a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x,
              a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w,
              a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z};
}

// p + w*uv + v x uv, uv = 2 (v x p)   (so3.h:54-60, droid_kernels.cu:58-69)
__device__ __forceinline__ Vec3 q_rot(Quat q, Vec3 p) {
  Vec3 uv{2.0f * (q.y * p.z - q.z * p.y), 2.0f * (q.z * p.x - q.x * p.z), 2.0f * (q.x * p.y - q.y * p.x)};
  return Vec3{p.x + q.w * uv.x + (q.y * uv.z - q.z * uv.y),
              p.y + q.w * uv.y + (q.z * uv.x - q.x * uv.z),
              p.z + q.w * uv.z + (q.x * uv.y - q.y * uv.x)};
}

// lietorch: Gij = G_j * G_i^{-1}, every constructor normalises.
__device__ __forceinline__ void lt_rel_pose(const float* Pi, const float* Pj, Vec3& tij, Quat& qij) {
  const Quat qi = q_normalize(Quat{Pi[3], Pi[4], Pi[5], Pi[6]});
  const Quat qi_inv = q_normalize(Quat{-qi.x, -qi.y, -qi.z, qi.w});       // SO3(conjugate) renormalises
  const Vec3 r = q_rot(qi_inv, Vec3{Pi[0], Pi[1], Pi[2]});
  const Vec3 ti_inv{-r.x, -r.y, -r.z};
  // Mul op re-constructs both operands from raw data (normalise again), then normalises the product
  const Quat a = q_normalize(Quat{Pj[3], Pj[4], Pj[5], Pj[6]});
  const Quat b = q_normalize(qi_inv);
  const Vec3 rt = q_rot(a, ti_inv);
  tij = Vec3{Pj[0] + rt.x, Pj[1] + rt.y, Pj[2] + rt.z};
  qij = q_normalize(q_normalize(q_mul(a, b)));                            // product, then act4's constructor
}

// droid_kernels.cu:96-107  qij = qj * conj(qi) (no normalisation), tij = tj - R(qij) ti
__device__ __forceinline__ void dk_rel_pose(const float* Pi, const float* Pj, Vec3& tij, Quat& qij) {
  const Quat qi{Pi[3], Pi[4], Pi[5], Pi[6]}, qj{Pj[3], Pj[4], Pj[5], Pj[6]};
  qij = Quat{-qj.w * qi.x + qj.x * qi.w - qj.y * qi.z + qj.z * qi.y,
             -qj.w * qi.y + qj.y * qi.w - qj.z * qi.x + qj.x * qi.z,
             -qj.w * qi.z + qj.z * qi.w - qj.x * qi.y + qj.y * qi.x,
             qj.w * qi.w + qj.x * qi.x + qj.y * qi.y + qj.z * qi.z};
  const Vec3 r = q_rot(qij, Vec3{Pi[0], Pi[1], Pi[2]});
  tij = Vec3{Pj[0] - r.x, Pj[1] - r.y, Pj[2] - r.z};
}

// droid_kernels.cu:80-94: Y = [R^T X_t ; R^T X_r + R^T (X_t x t)]
__device__ __forceinline__ void dk_adj(Vec3 t, Quat q, const float X[6], float Y[6]) {
  const Quat qi{-q.x, -q.y, -q.z, q.w};
  const Vec3 a = q_rot(qi, Vec3{X[0], X[1], X[2]});
  const Vec3 b = q_rot(qi, Vec3{X[3], X[4], X[5]});
  const Vec3 u{t.z * X[1] - t.y * X[2], t.x * X[2] - t.z * X[0], t.y * X[0] - t.x * X[1]};
  const Vec3 c = q_rot(qi, u);
  Y[0] = a.x; Y[1] = a.y; Y[2] = a.z;
  Y[3] = b.x + c.x; Y[4] = b.y + c.y; Y[5] = b.z + c.z;
}

// droid_kernels.cu:110-176 with xi[5] where the reference reads the out-of-bounds xi[45] (:154)
__device__ __forceinline__ void dk_exp(const float xi[6], Vec3& t, Quat& q) {
  const float px = xi[3], py = xi[4], pz = xi[5];
  const float theta_sq = px * px + py * py + pz * pz;
  const float theta = sqrtf(theta_sq);
  float imag, real;
  if (theta_sq < 1e-8f) {
    const float theta_p4 = theta_sq * theta_sq;
    imag = 0.5f - (1.0f / 48.0f) * theta_sq + (1.0f / 3840.0f) * theta_p4;
    real = 1.0f - (1.0f / 8.0f) * theta_sq + (1.0f / 384.0f) * theta_p4;
  } else {
    imag = sinf(0.5f * theta) / theta;
    real = cosf(0.5f * theta);
  }
  q = Quat{imag * px, imag * py, imag * pz, real};
  t = Vec3{xi[0], xi[1], xi[2]};
  if (theta > 1e-4f) {
    const float a = (1.0f - cosf(theta)) / theta_sq;
    const float b = (theta - sinf(theta)) / (theta * theta_sq);
    const Vec3 c1{py * xi[2] - pz * xi[1], pz * xi[0] - px * xi[2], px * xi[1] - py * xi[0]};
    const Vec3 c2{py * c1.z - pz * c1.y, pz * c1.x - px * c1.z, px * c1.y - py * c1.x};
    t.x += a * c1.x + b * c2.x;
    t.y += a * c1.y + b * c2.y;
    t.z += a * c1.z + b * c2.z;
  }
}

// droid_kernels.cu:856-874  Exp(xi) * T
__device__ __forceinline__ void dk_retr(const float xi[6], float P[7]) {
  Vec3 dt; Quat dq;
  dk_exp(xi, dt, dq);
  const Quat q{P[3], P[4], P[5], P[6]};
  const Quat q1{dq.w * q.x + dq.x * q.w + dq.y * q.z - dq.z * q.y,
                dq.w * q.y + dq.y * q.w + dq.z * q.x - dq.x * q.z,
                dq.w * q.z + dq.z * q.w + dq.x * q.y - dq.y * q.x,
                dq.w * q.w - dq.x * q.x - dq.y * q.y - dq.z * q.z};
  const Vec3 r = q_rot(dq, Vec3{P[0], P[1], P[2]});
  P[0] = r.x + dt.x; P[1] = r.y + dt.y; P[2] = r.z + dt.z;
  P[3] = q1.x; P[4] = q1.y; P[5] = q1.z; P[6] = q1.w;
}

// lietorch retraction Exp(xi) * T (groups.py:153-156; so3.h:153-175 Exp, so3.h:177-196 left Jacobian, se3.h:134-142),
// quaternion renormalised as every lietorch constructor does.
__device__ __forceinline__ void lt_retr(const float xi[6], float P[7]) {
  const float px = xi[3], py = xi[4], pz = xi[5];
  const float theta2 = px * px + py * py + pz * pz;
  const float theta = sqrtf(theta2);
  float imag, real, c1, c2;
  if (theta < 1e-6f) {
    const float theta4 = theta2 * theta2;
    imag = 0.5f - (1.0f / 48.0f) * theta2 + (1.0f / 3840.0f) * theta4;
    real = 1.0f - (1.0f / 8.0f) * theta2 + (1.0f / 384.0f) * theta4;
    c1 = 0.5f - (1.0f / 24.0f) * theta2;
    c2 = 1.0f / 6.0f - (1.0f / 120.0f) * theta2;
  } else {
    imag = sinf(0.5f * theta) / theta;
    real = cosf(0.5f * theta);
    c1 = (1.0f - cosf(theta)) / theta2;
    c2 = (theta - sinf(theta)) / (theta2 * theta);
  }
  const Quat dq = q_normalize(Quat{imag * px, imag * py, imag * pz, real});
  // t = (I + c1 Phi + c2 Phi^2) tau
  const Vec3 tau{xi[0], xi[1], xi[2]};
  const Vec3 a{py * tau.z - pz * tau.y, pz * tau.x - px * tau.z, px * tau.y - py * tau.x};
  const Vec3 b{py * a.z - pz * a.y, pz * a.x - px * a.z, px * a.y - py * a.x};
  const Vec3 dt{tau.x + c1 * a.x + c2 * b.x, tau.y + c1 * a.y + c2 * b.y, tau.z + c1 * a.z + c2 * b.z};
  const Quat q = q_normalize(Quat{P[3], P[4], P[5], P[6]});
  const Quat q1 = q_normalize(q_mul(dq, q));
  const Vec3 r = q_rot(dq, Vec3{P[0], P[1], P[2]});
  P[0] = r.x + dt.x; P[1] = r.y + dt.y; P[2] = r.z + dt.z;
  P[3] = q1.x; P[4] = q1.y; P[5] = q1.z; P[6] = q1.w;
}
