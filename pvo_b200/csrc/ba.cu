// Dense bundle adjustment (Gauss-Newton / LM step on poses + per-pixel inverse depth via Schur complement),
// entirely on the device.
//
// Replaces droid_backends.ba -> ba_cuda (VO_Module/src/droid.cpp:87-114, VO_Module/src/droid_kernels.cu:1293-1410)
// and the kernels / host code it drives:
//   projective_transform_kernel :177-403   accum_kernel + accum_cuda (host CSR, D2H/H2D) :833-853,927-977
//   EEt6x6 / Ev6x1 / EvT6x1     :980-1094  SparseBlock (Eigen SimplicialLLT on the CPU, float64) :1096-1198
//   schur_block (host O(P^2 deg^2) loops) :1201-1290   pose_retr / disp_retr :856-925
// Semantics kept (SURVEY.md section 8(a) addendum): MIN_DEPTH 0.25 gating of weight and 1/Z; quaternions used as
// stored; Q = 1/(C + eta); LM damping applied to the SCHUR COMPLEMENT diag += ep + lm*diag; failed factorisation
// => dx = 0; EvT's "pose index <= 0 contributes nothing to dz" quirk (:1084); in-place update of poses/disps.
// Deviation: expSE3 reads xi[5] where the reference reads the out-of-bounds xi[45] (:154, undefined behaviour).
//
// B200 design.  The reference launches one block per EDGE (36 blocks on 148 SMs), materialises Eii/Eij/Cii/bz for
// every edge, then bounces through the host >= 6 times per iteration (argsort CSR, Eigen triplets, sparse LLT).
// Here work is organised by SOURCE FRAME: block (k, s) owns 256 pixels of depth frame kx[k] and walks that frame's
// out-edges, so C, w and the self-row Ei are complete in registers without any cross-block accumulation, the
// per-edge 12x12 Hessian is reduced with warp shuffles, and the grid is K * ceil(HW/256) blocks (152 at config 2).
// The reduced camera system (6P x 6P) is accumulated in float64 with atomics straight into one dense matrix, the
// Schur products are formed per frame from an smem tile (6x6 pose blocks in registers), and a single-block float64 band
// LDL^T solves it on the device -- from both ends at once when the window is large (ba_solve_kernel).  No host
// synchronisation anywhere: the whole call is CUDA-graph capturable.
// Algorithmic HBM bytes per GN iteration: 76 B * HW * E -> latency-bound at config 2.
#include "se3.cuh"

namespace {

constexpr float kMinDepth = 0.25f;
constexpr int kPix = 256;       // pixels (= threads) per block
constexpr int kParts = 2;       // replicas of the reduced system (see Bufs::sys)

struct Plan {          // lives in the integer part of the workspace
  int* frame2k;        // [nframes] frame index -> row in kx, or -1
  int* kx;             // [Kcap]    sorted unique(ts U ii)
  int* src_ptr;        // [Kcap+1]  CSR over edges grouped by source frame (rows follow kx)
  int* src_idx;        // [E]       edge ids, ascending within a row
  int* meta;           // [0]=K on device, [1]=error flag, [2]=max out-degree
};

struct Bufs {
  float* Eij;          // [E][6][HW]
  float* Ei;           // [K][6][HW]
  float* Q;            // [K][HW]
  float* w;            // [K][HW]
  double* sys;         // [kParts][n*n + 2n]: per part {H [n][n] reduced camera system | b [n] | dA [n]}, n = 6P.
                       // Blocks accumulate into part (blockIdx.x % kParts) to spread same-address fp64 atomics; measured on
                       // B200 the L2 absorbs them (1 vs 16 replicas: no change in the Schur pass) while every extra
                       // replica costs the single-block solve a latency-bound read, so 2 are kept.
  float* dx;           // [P][6]
};

// ---------------------------------------------------------------------------------------------------
// plan: kx = unique(arange(t0,t1) U ii), edges grouped by source frame.  One block.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) ba_plan_kernel(const long* __restrict__ ii, int E, int nframes, int t0, int t1,
                                                       int Kcap, Plan pl, double* sys, int n) {
  __shared__ int s_warp[32];
  __shared__ int s_total;
  const int tid = threadIdx.x;
  for (int f = tid; f < nframes; f += blockDim.x) pl.frame2k[f] = (f >= t0 && f < t1) ? 1 : 0;
  __syncthreads();
  for (int e = tid; e < E; e += blockDim.x) {
    const long f = ii[e];
    if (f >= 0 && f < nframes) pl.frame2k[f] = 1;
  }
  __syncthreads();
  // block-wide exclusive scan over nframes flags, chunked per thread
  const int chunk = (nframes + blockDim.x - 1) / blockDim.x;
  const int lo = min(tid * chunk, nframes), hi = min(lo + chunk, nframes);
  int cnt = 0;
  for (int f = lo; f < hi; f++) cnt += pl.frame2k[f];
  int inc = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, inc, o);
    if ((tid & 31) >= o) inc += v;
  }
  if ((tid & 31) == 31) s_warp[tid >> 5] = inc;
  __syncthreads();
  if (tid < 32) {
    int v = (tid < (int)(blockDim.x >> 5)) ? s_warp[tid] : 0;
    int w = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, w, o);
      if (tid >= o) w += u;
    }
    s_warp[tid] = w - v;
    if (tid == 31) s_total = w;
  }
  __syncthreads();
  int base = s_warp[tid >> 5] + inc - cnt;
  for (int f = lo; f < hi; f++) {
    if (pl.frame2k[f]) {
      if (base < Kcap) pl.kx[base] = f;
      pl.frame2k[f] = base++;
    } else {
      pl.frame2k[f] = -1;
    }
  }
  __syncthreads();
  const int K = s_total;
  if (tid == 0) {
    pl.meta[0] = K;
    pl.meta[1] = (K > Kcap) ? 1 : 0;
    pl.meta[2] = 0;
  }
  __syncthreads();
  // CSR rows: thread per source frame, edges kept in ascending id order (deterministic accumulation order)
  const int Kc = min(K, Kcap);
  for (int k = tid; k < Kc; k += blockDim.x) {
    int c = 0;
    const int f = pl.kx[k];
    for (int e = 0; e < E; e++) c += (ii[e] == f);
    pl.src_ptr[k + 1] = c;                       // counts first, scanned below
    atomicMax(&pl.meta[2], c);
  }
  __syncthreads();
  if (tid == 0) {
    pl.src_ptr[0] = 0;
    for (int k = 0; k < Kc; k++) pl.src_ptr[k + 1] += pl.src_ptr[k];
  }
  __syncthreads();
  for (int k = tid; k < Kc; k += blockDim.x) {
    int o = pl.src_ptr[k];
    const int f = pl.kx[k];
    for (int e = 0; e < E; e++)
      if (ii[e] == f) pl.src_idx[o++] = e;
  }
  if (sys)
    for (long i = tid; i < (long)kParts * ((long)n * n + 2 * n); i += blockDim.x) sys[i] = 0.0;
}

// ---------------------------------------------------------------------------------------------------
// pass 1: residuals, Jacobians, per-edge Hessian blocks, per-frame C / w / Ei, per-edge Eij
// ---------------------------------------------------------------------------------------------------
// one step of a cross-lane transpose-reduce: lanes with BIT set keep values [W, 2W), the others [0, W); after the
// steps (48,16) (24,8) (12,4) (6,2) (3,1) every lane holds 3 fully reduced values of a 96-vector
template <int W, int BIT>
__device__ __forceinline__ void xr_step(float* x, int lane) {
  const bool up = lane & BIT;
#pragma unroll
  for (int j = 0; j < W; j++) {
    const float send = up ? x[j] : x[j + W];
    const float recv = __shfl_xor_sync(0xffffffffu, send, BIT);
    x[j] = (up ? x[j + W] : x[j]) + recv;
  }
}

struct EdgeJac {
  float Ju[12], Jv[12];     // [Ji | Jj] for the x and y residual rows
  float Jzu, Jzv, wu, wv, ru, rv;
};

__device__ __forceinline__ void edge_jacobian(const float* Pi, const float* Pj, float fx, float fy, float cx, float cy,
                                              float u, float v, float disp, float tu, float tv, float wgu, float wgv,
                                              int sem, EdgeJac& J) {
  Vec3 tij; Quat qij;
  if (sem) lt_rel_pose(Pi, Pj, tij, qij); else dk_rel_pose(Pi, Pj, tij, qij);
  const Vec3 X0{(u - cx) / fx, (v - cy) / fy, 1.0f};
  const Vec3 r = q_rot(qij, X0);
  const float x = r.x + disp * tij.x, y = r.y + disp * tij.y, z = r.z + disp * tij.z, h = disp;
  // sem 0 (droid_kernels.cu:26,284-291): Z < 0.25 zeroes weight and 1/Z.
  // sem 1 (geom/projective_ops.py:6,48,117): valid = Z > 0.2; Z < 0.1 -> 1 before the divide.
  const bool bad = sem ? !(z > 0.2f) : (z < kMinDepth);
  const float d = sem ? (1.0f / ((z < 0.1f) ? 1.0f : z)) : (bad ? 0.0f : 1.0f / z);
  const float d2 = d * d;
  J.wu = bad ? 0.0f : .001f * wgu;
  J.wv = bad ? 0.0f : .001f * wgv;
  J.ru = tu - (fx * d * x + cx);
  J.rv = tv - (fy * d * y + cy);
  float* Jj = &J.Ju[6];
  Jj[0] = fx * (h * d); Jj[1] = fx * 0; Jj[2] = fx * (-x * h * d2);
  Jj[3] = fx * (-x * y * d2); Jj[4] = fx * (1 + x * x * d2); Jj[5] = fx * (-y * d);
  J.Jzu = fx * (tij.x * d - tij.z * (x * d2));
  dk_adj(tij, qij, Jj, &J.Ju[0]);
#pragma unroll
  for (int n = 0; n < 6; n++) J.Ju[n] = -J.Ju[n];
  Jj = &J.Jv[6];
  Jj[0] = fy * 0; Jj[1] = fy * (h * d); Jj[2] = fy * (-y * h * d2);
  Jj[3] = fy * (-1 - y * y * d2); Jj[4] = fy * (x * y * d2); Jj[5] = fy * (x * d);
  J.Jzv = fy * (tij.y * d - tij.z * (y * d2));
  dk_adj(tij, qij, Jj, &J.Jv[0]);
#pragma unroll
  for (int n = 0; n < 6; n++) J.Jv[n] = -J.Jv[n];
}

__global__ void __launch_bounds__(kPix) ba_pass1_kernel(
    const float* __restrict__ poses, const float* __restrict__ disps, const float* __restrict__ intrinsics,
    const float* __restrict__ targets, const float* __restrict__ weights, const float* __restrict__ eta,
    const long* __restrict__ ii, const long* __restrict__ jj, Plan pl, Bufs bf, int E, int HW, int wd, int t0, int P,
    int interleaved, int motion_only, int sem) {
  const int K = pl.meta[0];
  const int k = blockIdx.y;
  if (k >= K) return;
  const int f = pl.kx[k];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int pix = blockIdx.x * kPix + tid;
  const bool active = pix < HW;
  const float fx = intrinsics[0], fy = intrinsics[1], cx = intrinsics[2], cy = intrinsics[3];
  const int i = active ? pix / wd : 0, j = active ? pix - i * wd : 0;
  const float disp = active ? disps[(long)f * HW + pix] : 1.0f;
  float Ei[6] = {0, 0, 0, 0, 0, 0};
  float C = 0.f, wacc = 0.f;
  __shared__ float red[kPix / 32][90];
  const int n = 6 * P;
  double* Hp = bf.sys + (long)(blockIdx.x % kParts) * ((long)n * n + 2 * n);
  double* bp = Hp + (long)n * n;
  double* dAp = bp + n;
  for (int it = pl.src_ptr[k]; it < pl.src_ptr[k + 1]; it++) {
    const int e = pl.src_idx[it];
    const int jx = (int)jj[e];
    float tu = 0, tv = 0, wgu = 0, wgv = 0;
    if (active) {
      if (interleaved) {
        const float2 t = reinterpret_cast<const float2*>(targets)[(long)e * HW + pix];
        const float2 w = reinterpret_cast<const float2*>(weights)[(long)e * HW + pix];
        tu = t.x; tv = t.y; wgu = w.x; wgv = w.y;
      } else {
        tu = targets[((long)e * 2 + 0) * HW + pix]; tv = targets[((long)e * 2 + 1) * HW + pix];
        wgu = weights[((long)e * 2 + 0) * HW + pix]; wgv = weights[((long)e * 2 + 1) * HW + pix];
      }
    }
    EdgeJac J;
    edge_jacobian(poses + 7 * f, poses + 7 * jx, fx, fy, cx, cy, (float)j, (float)i, disp, tu, tv, wgu, wgv, sem, J);
    if (!active) { J.wu = 0.f; J.wv = 0.f; }
    if (!motion_only) {
      const float au = J.wu * J.Jzu, av = J.wv * J.Jzv;
      if (active) {
#pragma unroll
        for (int m = 0; m < 6; m++) bf.Eij[((long)e * 6 + m) * HW + pix] = au * J.Ju[6 + m] + av * J.Jv[6 + m];
      }
#pragma unroll
      for (int m = 0; m < 6; m++) Ei[m] += au * J.Ju[m] + av * J.Jv[m];
      C += au * J.Jzu + av * J.Jzv;
      wacc += au * J.ru + av * J.rv;
    }
    // 12x12 lower triangle (78) + vi, vj (12) = 90 per-pixel values (padded to 96): cross-lane transpose-reduce
    // (48+24+12+6+3 = 93 shuffles instead of 90 x 5), then the 8 warps combine through shared memory
    {
      float h[96];
      int l = 0;
#pragma unroll
      for (int a = 0; a < 12; a++) {
#pragma unroll
        for (int b = 0; b <= a; b++) h[l++] = J.wu * J.Ju[a] * J.Ju[b] + J.wv * J.Jv[a] * J.Jv[b];
      }
#pragma unroll
      for (int a = 0; a < 12; a++) h[78 + a] = J.wu * J.ru * J.Ju[a] + J.wv * J.rv * J.Jv[a];
#pragma unroll
      for (int a = 90; a < 96; a++) h[a] = 0.0f;
      xr_step<48, 16>(h, lane);
      xr_step<24, 8>(h, lane);
      xr_step<12, 4>(h, lane);
      xr_step<6, 2>(h, lane);
      xr_step<3, 1>(h, lane);
      const int col = ((lane >> 4) & 1) * 48 + ((lane >> 3) & 1) * 24 + ((lane >> 2) & 1) * 12 + ((lane >> 1) & 1) * 6 + (lane & 1) * 3;
#pragma unroll
      for (int q = 0; q < 3; q++)
        if (col + q < 90) red[warp][col + q] = h[q];
    }
    __syncthreads();
    if (tid < 90) {
      double s = 0.0;
#pragma unroll
      for (int w = 0; w < kPix / 32; w++) s += (double)red[w][tid];
      const int pi = f - t0, pj = jx - t0;
      if (tid < 78) {
        int a = 0, rem = tid;                      // invert l = a(a+1)/2 + b
        while (rem > a) { rem -= a + 1; a++; }
        const int bcol = rem;
        const int pa = a < 6 ? pi : pj, pb = bcol < 6 ? pi : pj;
        if (pa >= 0 && pa < P && pb >= 0 && pb < P) {
          const int row = 6 * pa + (a % 6), col = 6 * pb + (bcol % 6);
          atomicAdd(&Hp[(long)row * n + col], s);
          if (row != col) atomicAdd(&Hp[(long)col * n + row], s);
          else if (sem) atomicAdd(&dAp[row], s);
        }
      } else {
        const int a = tid - 78;
        const int pa = a < 6 ? pi : pj;
        if (pa >= 0 && pa < P) atomicAdd(&bp[6 * pa + (a % 6)], s);
      }
    }
    __syncthreads();
  }
  if (active && !motion_only) {
#pragma unroll
    for (int m = 0; m < 6; m++) bf.Ei[((long)k * 6 + m) * HW + pix] = Ei[m];
    bf.Q[(long)k * HW + pix] = sem ? 1.0f / (C + eta[(long)k * HW + pix] + 1e-7f)      // geom/ba.py:91
                                   : 1.0f / (C + eta[(long)k * HW + pix]);             // droid_kernels.cu:1376
    bf.w[(long)k * HW + pix] = wacc;
  }
}

// ---------------------------------------------------------------------------------------------------
// pass 2: Schur complement.  For depth frame k the E-rows are: the frame's own pose (Ei) and the target pose of
// every out-edge (Eij), restricted to poses inside [t0,t1).  S += sum_pix Erow_a Q Erow_b^T for ALL row pairs
// (exactly the pair enumeration of schur_block :1236-1251), b -= sum_pix Q w Erow.
// ---------------------------------------------------------------------------------------------------
constexpr int kPix2 = 128;      // pixels per block in the Schur pass (more, smaller blocks: 2+ resident per SM)
// Register-tiled: a task is one 6x6 pose block (row ra, row rb <= ra) of the frame's Gram matrix, held as 36 accumulators by each
// of 8 lanes that split the pixels (lane q takes pixels q, q+8, ...: consecutive shared-memory rows, no bank conflicts); three
// shuffle rounds fold the 8 partial blocks and the lanes share out the float64 atomics.  One shared-memory read per three
// multiply-adds -- the scalar-pair version this replaces (one (a, b) entry per thread: three reads per multiply-add, 10.7 k load
// wavefronts per block) was the largest kernel of the BA at 38 us.  Products are formed exactly as before, (E_a Q) E_b in float32.
__global__ void __launch_bounds__(256, 3) ba_pass2_kernel(const long* __restrict__ jj, Plan pl, Bufs bf, int HW, int t0,
                                                          int P, int smem_floats) {       // 3 blocks per SM: the 304 blocks of config 2 in one wave
  extern __shared__ float sm[];
  const int K = pl.meta[0];
  const int k = blockIdx.y;
  if (k >= K) return;
  const int f = pl.kx[k];
  const int tid = threadIdx.x;
  const int deg = pl.src_ptr[k + 1] - pl.src_ptr[k];
  if (deg == 0) return;                              // no out-edges here (e.g. a frame another rank owns): E_k = 0, nothing to eliminate
  const int R = deg + 1, R6 = 6 * R;                 // row 0 = self, rows 1.. = out-edges (CSR order)
  // smem: rowpose[R] ints | Wt[PC] | tile E [PC][R6] | tile E Q [PC][R6]
  int* rowpose = reinterpret_cast<int*>(sm);
  const int hdr = (R + 3) & ~3;
  int PC = (smem_floats - hdr) / (2 * R6 + 2);
  PC = min(kPix2, PC & ~7);
  if (PC <= 0) { if (tid == 0) pl.meta[1] = 2; return; }
  float* Wt = sm + hdr;
  float* tile = Wt + ((PC + 1) & ~1);                // 8-byte aligned rows (R6 is even)
  float* tileq = tile + PC * R6;
  for (int r = tid; r < R; r += 256) {
    const int pose = (r == 0) ? f - t0 : (int)jj[pl.src_idx[pl.src_ptr[k] + r - 1]] - t0;
    rowpose[r] = (pose >= 0 && pose < P) ? pose : -1;
  }
  const int n = 6 * P;
  double* Hp = bf.sys + (long)(blockIdx.x % kParts) * ((long)n * n + 2 * n);
  double* bp = Hp + (long)n * n;
  const int pix0 = blockIdx.x * kPix2;
  const int q = tid & 7, ntask = R * (R + 1) / 2;
  for (int c0 = 0; c0 < kPix2; c0 += PC) {
    __syncthreads();
    const int np = min(PC, HW - pix0 - c0);
    if (np <= 0) break;
    // load both tiles (coalesced along pixels); rows outside the window and pixels past the end are zero
    // warp per tile row, lanes along the pixels, four loads of E and four of Q in flight per lane (a one-element-per-trip copy loop
    // pays one L2 round trip per element: 21 trips per thread at config 2)
    const float* Qk = bf.Q + (long)k * HW + pix0 + c0;
    for (int row = tid >> 5; row < R6; row += 8) {
      const int r = row / 6, m = row - 6 * r;
      const float* src = (r == 0) ? bf.Ei + ((long)k * 6 + m) * HW : bf.Eij + ((long)pl.src_idx[pl.src_ptr[k] + r - 1] * 6 + m) * HW;
      src += pix0 + c0;
      const bool live = rowpose[r] >= 0;
      for (int pb0 = 0; pb0 < PC; pb0 += 128) {
        float v[4], qv[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int p = pb0 + 32 * u + (tid & 31);
          const bool ok = live && p < np;
          v[u] = ok ? __ldg(src + p) : 0.f;
          qv[u] = ok ? __ldg(Qk + p) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int p = pb0 + 32 * u + (tid & 31);
          if (p < PC) {
            tile[p * R6 + row] = v[u];
            tileq[p * R6 + row] = v[u] * qv[u];
          }
        }
      }
    }
    for (int p = tid; p < PC; p += 256) Wt[p] = p < np ? bf.w[(long)k * HW + pix0 + c0 + p] : 0.f;
    __syncthreads();
    for (int task = tid >> 3; task < ntask; task += 32) {              // all 8 lanes of a task stay together (ntask is block-uniform)
      int ra = (int)((sqrtf(8.0f * task + 1.0f) - 1.0f) * 0.5f);       // task = ra (ra + 1) / 2 + rb, rb <= ra
      while (ra * (ra + 1) / 2 > task) ra--;
      while ((ra + 1) * (ra + 2) / 2 <= task) ra++;
      const int rb = task - ra * (ra + 1) / 2;
      const int pa = rowpose[ra], pb = rowpose[rb];
      if (pa < 0 || pb < 0) continue;                                  // uniform over the 8 lanes
      float acc[6][6], bw[6];
#pragma unroll
      for (int i = 0; i < 6; i++) {
        bw[i] = 0.f;
#pragma unroll
        for (int j = 0; j < 6; j++) acc[i][j] = 0.f;
      }
      const bool diag = ra == rb;
      for (int p = q; p < PC; p += 8) {
        const float2* ea = reinterpret_cast<const float2*>(tileq + p * R6 + 6 * ra);
        const float2* eb = reinterpret_cast<const float2*>(tile + p * R6 + 6 * rb);
        const float2 a0 = ea[0], a1 = ea[1], a2 = ea[2], b0 = eb[0], b1 = eb[1], b2 = eb[2];
        const float a[6] = {a0.x, a0.y, a1.x, a1.y, a2.x, a2.y}, b[6] = {b0.x, b0.y, b1.x, b1.y, b2.x, b2.y};
#pragma unroll
        for (int i = 0; i < 6; i++)
#pragma unroll
          for (int j = 0; j < 6; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        if (diag) {
          const float wv = Wt[p];
#pragma unroll
          for (int i = 0; i < 6; i++) bw[i] = fmaf(wv, a[i], bw[i]);
        }
      }
      const unsigned gmask = 0xFFu << (tid & 24);                      // the task's 8 lanes: tasks of one warp may diverge (skipped rows, loop tail)
#pragma unroll
      for (int o = 1; o < 8; o <<= 1) {
#pragma unroll
        for (int i = 0; i < 6; i++) {
          bw[i] += __shfl_xor_sync(gmask, bw[i], o);
#pragma unroll
          for (int j = 0; j < 6; j++) acc[i][j] += __shfl_xor_sync(gmask, acc[i][j], o);
        }
      }
      // every lane holds the block; lane q emits entries q, q + 8, ...  (i, j) = system indices of (row a, row b), a >= b as scalars
#pragma unroll
      for (int e = 0; e < 36; e++) {
        if ((e & 7) != q) continue;
        const int ma = e / 6, mb = e % 6;
        if (diag && mb > ma) continue;
        const double sv = -(double)acc[ma][mb];
        const long i = 6 * pa + ma, j = 6 * pb + mb;
        if (diag && ma == mb) {
          atomicAdd(&Hp[i * n + j], sv);
        } else if (i == j) {                                           // two rows of the same pose (duplicate edges)
          atomicAdd(&Hp[i * n + j], 2.0 * sv);
        } else {
          atomicAdd(&Hp[i * n + j], sv);
          atomicAdd(&Hp[j * n + i], sv);
        }
      }
      if (diag && q < 6) {
        float bsel = bw[0];
#pragma unroll
        for (int i = 1; i < 6; i++) if (q == i) bsel = bw[i];
        atomicAdd(&bp[6 * pa + q], -(double)bsel);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// solve: sum the kParts replicas, LM damping, blocked (6x6 = one pose) right-looking LDL^T + blocked triangular
// solves in float64 (SPD test: every pivot d_j > 0, same acceptance as the reference's LLT).  One block; two barriers per block column.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void damp(double* A, int ld, const double* dA, int P, float lm, float ep, int sem, int moba, int tid, int T) {
  const int n = 6 * P;
  if (moba) {
    // geom/chol.py:32-44 block_solve: `H + (ep + lm*H) * eye(6)` broadcasts over the (N,N,6,6) block tensor, i.e. the
    // diagonal of EVERY 6x6 block -- off-diagonal pose blocks included -- is damped.  Kept as the reference does it.
    for (int idx = tid; idx < P * P * 6; idx += T) {
      const int pq = idx / 6, a = idx - 6 * pq;
      const long off = (long)(6 * (pq / P) + a) * ld + 6 * (pq % P) + a;
      A[off] += (double)ep + (double)lm * A[off];
    }
  } else {
    // sem 0: damp the Schur complement (droid_kernels.cu:1175-1176); sem 1: damp H before subtracting E Q E^T (chol.py:55-62)
    for (int i = tid; i < n; i += T) A[(long)i * ld + i] += (double)ep + (double)lm * (sem ? dA[i] : A[(long)i * ld + i]);
  }
}

// sums parts 1.. into part 0 and clears them (multi-GPU: part 0 is what gets all-reduced; large n: global-memory solve)
__global__ void __launch_bounds__(256) ba_combine_kernel(Bufs bf, int P) {
  const long psz = 36L * P * P + 12L * P;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < psz; i += (long)gridDim.x * blockDim.x) {
    double s = bf.sys[i];
    for (int g = 1; g < kParts; g++) { s += bf.sys[g * psz + i]; bf.sys[g * psz + i] = 0.0; }
    bf.sys[i] = s;
  }
}

// Band-packed exchange (frame graphs with a known block bandwidth): what the factorisation reads is the lower band only -- n (band + 6)
// doubles instead of n^2 (P = 63, radius 3: 133 KB instead of 1.1 MB).  Folds the replicas' band into `packed` ([n][band + 6] rows in the
// SysMat<true> layout | b [n] | dA [n]) and clears EVERY entry of every replica (pass 1 / pass 2 write both triangles), so the sharded
// path all-reduces the packed block and nobody needs ba_zero_kernel afterwards.  An in-band entry is read and cleared by the same
// thread; out-of-band entries are only written.
__global__ void __launch_bounds__(256) ba_combine_band_kernel(Bufs bf, double* __restrict__ packed, int P, int band) {
  const int n = 6 * P, W = band + 6;
  const long psz = (long)n * n + 2 * n;
  const long stride = (long)gridDim.x * blockDim.x, t0 = (long)blockIdx.x * blockDim.x + threadIdx.x;
  for (long i = t0; i < (long)n * W + 2 * n; i += stride) {
    double s = 0.0;
    if (i < (long)n * W) {
      const int r = (int)(i / W), c = (int)(i % W) + max(0, 6 * (r / 6) - band);
      if (c <= r) {
        const long d = (long)r * n + c;
#pragma unroll
        for (int g = 0; g < kParts; g++) { s += bf.sys[g * psz + d]; bf.sys[g * psz + d] = 0.0; }
      }
    } else {
      const long d = (long)n * n + (i - (long)n * W);
#pragma unroll
      for (int g = 0; g < kParts; g++) { s += bf.sys[g * psz + d]; bf.sys[g * psz + d] = 0.0; }
    }
    packed[i] = s;
  }
  for (long i = t0; i < (long)n * n; i += stride) {
    const int r = (int)(i / n), c = (int)(i % n);
    if (c > r || c < 6 * (r / 6) - band) {
#pragma unroll
      for (int g = 0; g < kParts; g++) bf.sys[g * psz + i] = 0.0;
    }
  }
}

// Dense (row-major, leading dimension ld) or band-packed view of the symmetric system's lower triangle.  Band-packed:
// row r keeps columns [6*(r/6) - band, r]: exactly the entries the band-limited factorisation touches.
template <bool BAND> struct SysMat {
  double* p; int ld; int band; int r0;                     // r0: first resident row (band-packed windows; 0 otherwise)
  __device__ __forceinline__ long off(int r) const {       // element (r, c) lives at p[off(r) + c]
    if (BAND) return (long)(r - r0) * ld - max(0, 6 * (r / 6) - band);
    return (long)r * ld;
  }
  __device__ __forceinline__ double& operator()(int r, int c) const { return p[off(r) + c]; }
};

// 1 / d for the pivots.  IEEE double division is a ~200-cycle dependent software sequence and sits six times on the critical path of
// every block column; the hardware reciprocal seed (relative error 2^-23) plus two Newton steps (-> 2^-92, i.e. last-bit accurate) is
// ~70 cycles.  d > 0 and far from the subnormal range here (the diagonal carries the damping ep + lm d).
__device__ __forceinline__ double fast_rcp(double d) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  r = fma(fma(-d, r, 1.0), r, r);
  r = fma(fma(-d, r, 1.0), r, r);
  return r;
}

// 6x6 diagonal block at c0: A_kk = L D L^T in place (unit lower L below the diagonal, d on it), 1/d -> dinv, and the
// block's share of the forward substitution (y_k <- L_kk^-1 y_k).  One thread, right-looking: column j costs one reciprocal (hardware
// seed + two Newton steps; float64 div / sqrt are ~100-cycle software sequences), five multiplies and the rank-1 update of what is left.
// Used by the generic path (systems that stay in global memory); shared-memory systems use the warp form below.
template <class M>
__device__ __forceinline__ void ldl_diag_block(M A, double* y, double* dinv, int c0, int* s_fail) {
  double D[6][6], z[6];
  double* row0 = A.p + A.off(c0) + c0;                      // rows c0 .. c0+5 share the band start: row i begins i * ld further
#pragma unroll
  for (int i = 0; i < 6; i++) {
    z[i] = y[c0 + i];
#pragma unroll
    for (int j = 0; j < 6; j++) D[i][j] = (j <= i) ? row0[(long)i * A.ld + j] : 0.0;
  }
  bool bad = false;
#pragma unroll
  for (int j = 0; j < 6; j++) {
    const double d = D[j][j];
    bad = bad || !(d > 0.0);                                // does not feed the chain: a failed factorisation zeroes dx, garbage after it is harmless
    const double iv = fast_rcp(d);
    dinv[c0 + j] = iv;
    double l[6];
#pragma unroll
    for (int i = 0; i < 6; i++) l[i] = (i > j) ? D[i][j] * iv : 0.0;       // D[i][j] holds L_ij d_j until the end of this step
#pragma unroll
    for (int i = 0; i < 6; i++) {
      if (i > j) {
        z[i] = fma(-l[i], z[j], z[i]);
#pragma unroll
        for (int c = 0; c < 6; c++)
          if (c > j && c <= i) D[i][c] = fma(-l[i], D[c][j], D[i][c]);
      }
    }
#pragma unroll
    for (int i = 0; i < 6; i++)
      if (i > j) D[i][j] = l[i];
  }
  if (bad) *s_fail = 1;
#pragma unroll
  for (int i = 0; i < 6; i++) {
    y[c0 + i] = z[i];
#pragma unroll
    for (int j = 0; j < 6; j++)
      if (j <= i) row0[(long)i * A.ld + j] = D[i][j];
  }
}

// The same 6x6 step by ONE WARP: lane <-> one entry of the block's lower triangle (lanes 0..20, row-major) or of its right-hand side
// (lanes 21..26, treated as a seventh matrix row: the forward substitution is the same elimination).  `v` is the lane's entry, already
// holding every update from earlier block columns -- it comes straight from the trailing-update registers.  Per pivot: broadcast d,
// one reciprocal (all lanes), two lane-indexed shuffles, one multiply, one FMA: ~14 instructions instead of ~60 in the single-thread
// form, whose ~370 instructions per block ran at 3-4 cycles each on the serial warp (section 4d of DESIGN.md); the dependent chain
// (shuffle, rcp, mul, fma per pivot) is ~90 cycles.  dk receives what the panel rows need: L_jt d_t [6][6] | 1/d [6] | z_k [6].
template <class M>
__device__ __forceinline__ void ldl_diag_block_warp(M A, double* y, double* dinv, int c0, int* s_fail, double* dk, int lane, int bi, int bj,
                                                    double v) {
  const bool tri = lane < 21, rhs = lane >= 21 && lane < 27;
  const int i = tri ? bi : 6, c = tri ? bj : (rhs ? lane - 21 : 7);        // c = 7: idle lane (never updated, never stored)
  const int baseA = tri ? bi * (bi + 1) / 2 : 21, baseB = (c < 6) ? c * (c + 1) / 2 : 0;
  double myiv = 0.0;
  bool bad = false;
#pragma unroll
  for (int j = 0; j < 6; j++) {
    const double d = __shfl_sync(0xffffffffu, v, j * (j + 1) / 2 + j);
    bad = bad || !(d > 0.0);
    const double iv = fast_rcp(d);
    const double va = __shfl_sync(0xffffffffu, v, min(baseA + j, 31));      // entry (i, j), still L_ij d_j
    const double vb = __shfl_sync(0xffffffffu, v, min(baseB + j, 31));      // entry (c, j)
    if (c == j) myiv = iv;
    if (c > j && c < 6) v = fma(-(va * iv), vb, v);
  }
  if (bad && lane == 0) *s_fail = 1;
  if (tri) {
    double* e = A.p + A.off(c0) + (long)i * A.ld + c0 + c;
    if (i == c) {
      *e = v;                                              // d_i
      dinv[c0 + i] = myiv;
      dk[36 + i] = myiv;
    } else {
      dk[i * 6 + c] = v;                                   // L_ic d_c for the panel recurrence
      *e = v * myiv;                                       // L_ic
    }
  } else if (rhs) {
    y[c0 + c] = v;
    dk[42 + c] = v;
  }
}

// A = L D L^T restricted to the block band, then L z = b, D w = z, L^T x = w in place in y.  dinv: scratch [n].
// Schedule per 6-wide block column k (two barriers, and the serial 6x6 step overlaps the trailing update):
//   [panel: L_rk for the band rows + their forward-substitution update] | [trailing update by warps 1.. while warp 0
//   updates, factors and forward-solves the NEXT diagonal block].
// Forward elimination of block columns kb0 .. kb1-1 (every row they touch -- up to 6 kb1 + band + 5 -- must be resident in A), then,
// separately, the D scaling and the backward substitution over a range of block columns: a system whose band does not fit shared
// memory is processed window by window (ba_solve_kernel), everything else in one call (ldl_solve).
//
// Measured on the 8-GPU window (P = 63, band 36; tools/prof_solve.py, cycles of thread 0 per block column): panel 760, next diagonal
// block 540 + 1 680, and then 1 300 more WAITING for the trailing update -- its loads walked rows four apart in a 42-double-pitch
// matrix (8-way bank conflicts) and recomputed L d per entry.  Now the panel leaves its two results, L and L d, transposed in a
// scratch ([6][band] each: consecutive rows in consecutive banks), the trailing update is one lower-triangle entry per thread
// with lanes along a matrix row, and matrices in shared memory get an odd pitch.
// Cycle probes for tools/prof_solve.py (a separate build with -DPVO_SOLVE_PROF; never compiled into the product library): thread 0 adds
// the cycles since its previous probe to slot i, so the slots partition thread 0's time through the solve.
#ifdef PVO_SOLVE_PROF
__device__ long long g_solve_prof[32];
__shared__ long long s_solve_prof[32];                      // accumulated in shared memory (a global read-modify-write would stall thread 0 ~500 cycles per probe)
#define PROF_DECL long long t_prof = clock64()
#define PROF(i) do { if (threadIdx.x == 0) { const long long t_now = clock64(); s_solve_prof[i] += t_now - t_prof; t_prof = clock64(); } } while (0)
#define PROF_BEGIN do { if (threadIdx.x == 0) for (int q = 0; q < 32; q++) s_solve_prof[q] = 0; } while (0)
#define PROF_END do { if (threadIdx.x == 0) for (int q = 0; q < 32; q++) g_solve_prof[q] += s_solve_prof[q]; } while (0)
#else
#define PROF_DECL
#define PROF(i)
#define PROF_BEGIN
#define PROF_END
#endif

// A thread GROUP (tid in [0, T), T a multiple of 32, its own named barrier `bar`) runs these: the whole block (bar 0 == __syncthreads) or
// one half of it when two eliminations proceed side by side (two-sided factorisation in ba_solve_kernel).
__device__ __forceinline__ void group_sync(int bar, int T) { asm volatile("bar.sync %0, %1;" ::"r"(bar), "r"(T) : "memory"); }

__host__ __device__ __forceinline__ int ldl_scratch_doubles(int band) { return 12 * band + 48; }

// scratch: ldl_scratch_doubles(band) doubles of shared memory owned by this group (panel L [6][band] | L d [6][band] | diagonal-block
//   digest [48]), or nullptr: generic accessor-based path (systems that live in global memory).  It carries the digest of the last factored
//   diagonal block from one call to the next (windowed factorisation).
// factor_first: factor the diagonal block of column kb0 on entry (false when the previous call already did, as its "next" block);
// factor_last: factor the diagonal block after column kb1 - 1 (false when other contributions are still to be added to it).
// Which warps of a group of T threads take the trailing update when warp `fw` factors the diagonal blocks: all the others.  (Warps w and
// w + 4 share an SM sub-partition; keeping the factoring sub-partition free of trailing work was measured: the single-thread 6x6
// factorisation got 17 % faster, the trailing update -- then the longer of the two -- 40 % slower; P = 63 solve 82 vs 79.5 us.)
__device__ __forceinline__ int default_trailing_mask(int T, int fw = 0) { return ((1 << (T >> 5)) - 1) & ~(1 << fw); }

// fw: the group-relative warp that factors the diagonal blocks; tmask: the group-relative warps that do the trailing update (0 = default)
template <class M>
__device__ void ldl_forward(M A, double* y, double* dinv, int n, int band, int tid, int T, int* s_fail, int* s_rowoff, double* scratch,
                            int kb0, int kb1, int row_lo, int row_hi, int bar = 0, bool factor_first = true, bool factor_last = true,
                            int fw = 0, int tmask = 0) {
  const int lane = tid & 31, warp = tid >> 5;
  if (!tmask) tmask = default_trailing_mask(T, fw);
  const int Tt = 32 * __popc(tmask);                                              // threads on the trailing update
  const int ttid = ((tmask >> warp) & 1) ? 32 * __popc(tmask & ((1 << warp) - 1)) + lane : -1;      // this thread's rank among them
  const bool fast = scratch != nullptr && n <= 1024 && A.off(row_hi - 1) + n < 0x7fffffffL && A.off(row_lo) > -0x7fffffffL;
  const int PS = band;
  double* Lp = scratch;                                    // [6][PS]: L_rj of the current panel, row-contiguous per j
  double* LDp = scratch + 6 * PS;                          // [6][PS]: L_rj d_j
  double* dk = fast ? scratch + 12 * PS : nullptr;
  if (fast)                                                // row r of the matrix starts at A.p + s_rowoff[r]
    for (int r = row_lo + tid; r < row_hi; r += T) s_rowoff[r] = (int)A.off(r);
  int bi = 0, bj = lane;                                   // lower-triangle entry of a 6x6 block updated by this lane (lanes < 21)
  while (bj > bi) { bj -= bi + 1; bi++; }
  group_sync(bar, T);
  if (factor_first && warp == fw) {
    if (fast) {
      const int f0 = 6 * kb0;
      double v = 0.0;
      if (lane < 21) v = A.p[s_rowoff[f0 + bi] + f0 + bj];
      else if (lane < 27) v = y[f0 + lane - 21];
      ldl_diag_block_warp(A, y, dinv, f0, s_fail, dk, lane, bi, bj, v);
    } else if (lane == 0) {
      ldl_diag_block(A, y, dinv, 6 * kb0, s_fail);
    }
  }
  group_sync(bar, T);
  PROF_DECL;
  for (int kb = kb0; kb < kb1; kb++) {
    const int c0 = 6 * kb;
    const int rend = min(n, c0 + 6 + band);
    const int m = rend - c0 - 6;                           // band rows below the diagonal block (a multiple of 6)
    if (fast) {                                            // panel: L_rk = A_rk L_kk^-T D_k^-1; y_r -= L_rk z_k
      if (tid < m) {
        double ldk[6][6], di[6], zk[6];                    // the diagonal block's digest: L_jt d_t, 1/d_j, z_k
#pragma unroll
        for (int j = 0; j < 6; j++) {
          di[j] = dk[36 + j];
          zk[j] = dk[42 + j];
#pragma unroll
          for (int t = 0; t < 6; t++) ldk[j][t] = (t < j) ? dk[j * 6 + t] : 0.0;
        }
        for (int rr = tid; rr < m; rr += T) {
          const int r = c0 + 6 + rr;
          double* Ar = A.p + s_rowoff[r] + c0;
          double a[6], x[6];
#pragma unroll
          for (int j = 0; j < 6; j++) a[j] = Ar[j];
          double ya = y[r], yb2 = 0.0;
#pragma unroll
          for (int j = 0; j < 6; j++) {
            double v = a[j];                               // L_rj d_j = A_rj - sum_{t<j} L_rt (d_t L_jt)
#pragma unroll
            for (int t = 0; t < 6; t++) if (t < j) v = fma(-x[t], ldk[j][t], v);
            LDp[j * PS + rr] = v;
            x[j] = v * di[j];
            Lp[j * PS + rr] = x[j];
            Ar[j] = x[j];
            if (j & 1) yb2 = fma(x[j], zk[j], yb2); else ya = fma(-x[j], zk[j], ya);
          }
          y[r] = ya - yb2;
        }
      }
    } else {
      for (int r = c0 + 6 + tid; r < rend; r += T) {
        double x[6];
        double yr = y[r];
#pragma unroll
        for (int j = 0; j < 6; j++) {
          double v = A(r, c0 + j);
#pragma unroll
          for (int t = 0; t < 6; t++) if (t < j) v -= x[t] * A((c0 + t), c0 + t) * A((c0 + j), c0 + t);
          x[j] = v * dinv[c0 + j];
          yr -= x[j] * y[c0 + j];
        }
#pragma unroll
        for (int j = 0; j < 6; j++) A(r, c0 + j) = x[j];
        y[r] = yr;
      }
    }
    PROF(0);                                                // panel (thread 0's row)
    group_sync(bar, T);
    PROF(1);                                                // wait for the slowest panel row
    if (m > 0) {                                            // trailing update (lower triangle): A_ic -= sum_t L_it d_t L_ct
      if (warp == fw) {
        // the next diagonal block first (21 entries, one per lane), then its factorisation -- same lanes, the entries stay in
        // registers -- while the other warps do the rest of the trailing update
        const bool factor_next = factor_last || kb + 1 < kb1;
        if (fast) {
          double v = 0.0;
          if (lane < 21) {
            double* Ai = A.p + s_rowoff[c0 + 6 + bi] + c0 + 6;
            double va = Ai[bj], vb = 0.0;
#pragma unroll
            for (int t = 0; t < 3; t++) {
              va = fma(-LDp[t * PS + bi], Lp[t * PS + bj], va);
              vb = fma(LDp[(t + 3) * PS + bi], Lp[(t + 3) * PS + bj], vb);
            }
            v = va - vb;
            if (!factor_next) Ai[bj] = v;
          } else if (lane < 27) {
            v = y[c0 + 6 + lane - 21];
          }
          PROF(2);                                          // next diagonal block: 21-entry update
          if (factor_next) ldl_diag_block_warp(A, y, dinv, c0 + 6, s_fail, dk, lane, bi, bj, v);     // the updated entries never leave the registers
          PROF(3);                                          // its factorisation
        } else {
          if (lane < 21) {
            const int i = c0 + 6 + bi, c = c0 + 6 + bj;
            double vn = A(i, c);
#pragma unroll
            for (int t = 0; t < 6; t++) vn -= A(i, c0 + t) * A((c0 + t), c0 + t) * A(c, c0 + t);
            A(i, c) = vn;
          }
          __syncwarp();
          if (lane == 0 && factor_next) ldl_diag_block(A, y, dinv, c0 + 6, s_fail);
          __syncwarp();
        }
      } else if (fast && ttid >= 0) {
        // rows 6 .. m-1 of the trailing block (the first six belong to warp 0), lower triangle: row 6+p (7+p entries) and row m-1-p
        // (m-p entries) together always hold m+7 entries, so R/2 row pairs x (m+7) entries enumerate the triangle without idle lanes
        const int R = m - 6, width = m + 7, total = (R >> 1) * width;
        int idx = ttid;
        int p = idx / width, k = idx - p * width;
        for (; idx < total; idx += Tt) {
          const bool first = k <= 6 + p;
          const int i = first ? 6 + p : m - 1 - p, c = first ? k : k - 7 - p;
          double* Ai = A.p + s_rowoff[c0 + 6 + i] + c0 + 6;
          double va = Ai[c], vb = 0.0;
#pragma unroll
          for (int t = 0; t < 3; t++) {
            va = fma(-LDp[t * PS + i], Lp[t * PS + c], va);
            vb = fma(LDp[(t + 3) * PS + i], Lp[(t + 3) * PS + c], vb);
          }
          Ai[c] = va - vb;
          k += Tt;
          while (k >= width) { k -= width; p++; }
        }
      } else if (!fast && ttid >= 0) {
        for (int idx = ttid; idx < m * m; idx += Tt) {
          const int i = c0 + 6 + idx / m, c = c0 + 6 + idx % m;
          if (c > i || i < c0 + 12) continue;              // rows of the next diagonal block belong to warp 0
          double v = A(i, c);
#pragma unroll
          for (int t = 0; t < 6; t++) v -= A(i, c0 + t) * A((c0 + t), c0 + t) * A(c, c0 + t);
          A(i, c) = v;
        }
      }
    }
    group_sync(bar, T);
    PROF(4);                                                // wait for the trailing update
  }
}

// L^T x = w for block columns kb_hi down to kb_lo, right-looking, one warp, no block barriers: once x_k is known it is pushed into the
// right-hand side of the band columns to its left (lanes along a matrix row: no reductions, no bank conflicts).  The rows below the
// range -- solved by an earlier call: the separator, or the previous window -- are pushed first.  A column outside [kb_lo, kb_hi] is
// never written, so every (row block, column) contribution is applied by exactly one call.
template <class M>
__device__ void ldl_backward(M A, double* y, int n, int band, int tid, int kb_hi, int kb_lo, int bar = 0, int T = 0, int fw = 0) {
  const int lane = tid & 31, warp = tid >> 5;
  if (warp == fw && kb_hi >= kb_lo) {
    const int clo = 6 * kb_lo;
    auto push = [&](int rb, int cb) {                      // y_c -= sum_j L(rb + j, c) x_{rb + j} for columns max(clo, rb - band) <= c < cb
      const double* row0 = A.p + A.off(rb);                // the six rows of a block share the band start: row j begins j * ld further
      double x[6];
#pragma unroll
      for (int j = 0; j < 6; j++) x[j] = y[rb + j];
      for (int c = max(clo, rb - band) + lane; c < cb; c += 64) {      // two columns per trip: their loads overlap
        const int c2 = c + 32;
        const bool two = c2 < cb;
        double va = y[c], vb = 0.0, wa = two ? y[c2] : 0.0, wb = 0.0;
#pragma unroll
        for (int j = 0; j < 3; j++) {
          va = fma(-row0[(long)j * A.ld + c], x[j], va);
          vb = fma(row0[(long)(j + 3) * A.ld + c], x[j + 3], vb);
          if (two) {
            wa = fma(-row0[(long)j * A.ld + c2], x[j], wa);
            wb = fma(row0[(long)(j + 3) * A.ld + c2], x[j + 3], wb);
          }
        }
        y[c] = va - vb;
        if (two) y[c2] = wa - wb;
      }
    };
    const int ctop = 6 * (kb_hi + 1), rtop = min(n, ctop + band);
    for (int rb = ctop; rb < rtop; rb += 6) { push(rb, ctop); __syncwarp(); }
    for (int kb = kb_hi; kb >= kb_lo; kb--) {
      const int c0 = 6 * kb;
      if (lane == 0) {                                     // unit upper-triangular 6x6 (L_kk^T) by back substitution
        const double* r0p = A.p + A.off(c0) + c0;
        double x[6], l[6][6];
#pragma unroll
        for (int t = 0; t < 6; t++)
#pragma unroll
          for (int j = 0; j < 6; j++) l[t][j] = (t > j) ? r0p[(long)t * A.ld + j] : 0.0;
#pragma unroll
        for (int j = 5; j >= 0; j--) {
          double v = y[c0 + j];
#pragma unroll
          for (int t = 0; t < 6; t++) if (t > j) v = fma(-l[t][j], x[t], v);
          x[j] = v;
        }
#pragma unroll
        for (int j = 0; j < 6; j++) y[c0 + j] = x[j];
      }
      __syncwarp();
      push(c0, c0);
      __syncwarp();
    }
  }
  if (T) group_sync(bar, T); else __syncthreads();
}

template <class M>
__device__ void ldl_solve(M A, double* y, double* dinv, int P, int n, int band, int tid, int T, int* s_fail, int* s_rowoff, double* scratch) {
  ldl_forward(A, y, dinv, n, band, tid, T, s_fail, s_rowoff, scratch, 0, P, 0, n);
  if (*s_fail) return;
  for (int i = tid; i < n; i += T) y[i] *= dinv[i];        // D w = z
  __syncthreads();
  ldl_backward(A, y, n, band, tid, P - 1, 0);
}

constexpr int kSolveSmem = 208 * 1024;      // dynamic shared memory of ba_solve_kernel<false>
constexpr int kWindowBytes = 190 * 1024;    // ... of which a sliding window of band rows + y + 1/d may take this much (the rest: panel scratch)

// Can a window of band rows (odd pitch band + 7) plus y and 1/d live in shared memory?  Shared by the kernel and by packed_band().
__host__ __device__ __forceinline__ bool window_fits(int n, int band) {
  return sizeof(double) * ((size_t)(band + 7) * (band + 18) + 2 * (size_t)n) <= (size_t)kWindowBytes;
}

// f(i, src[i]) for i in [i0, i1), eight loads in flight per thread: a dependent-latency copy loop (one L2 round trip per element and
// thread) was 16 us of the P = 63 solve
template <class F>
__device__ __forceinline__ void for_each_loaded(const double* __restrict__ src, long i0, long i1, int tid, int T, F f) {
  for (long base = i0; base < i1; base += 8L * T) {
    double v[8];
#pragma unroll
    for (int u = 0; u < 8; u++) {
      const long i = base + (long)u * T + tid;
      v[u] = i < i1 ? __ldg(src + i) : 0.0;
    }
#pragma unroll
    for (int u = 0; u < 8; u++) {
      const long i = base + (long)u * T + tid;
      if (i < i1) f(i, v[u]);
    }
  }
}

template <bool IN_SMEM>
__global__ void __launch_bounds__(512) ba_solve_kernel(Bufs bf, int P, float lm, float ep, int sem, int moba, int bw_hint, double* packed) {
  extern __shared__ double smd[];
  __shared__ int s_fail;
  __shared__ int s_rowoff[1024];
  const int n = 6 * P;
  const int tid = threadIdx.x, T = blockDim.x;
  const int warp = tid >> 5, lane = tid & 31, nw = T >> 5;
  const long psz = (long)n * n + 2 * n;
  const int ld = IN_SMEM ? n + 1 : n;
  if (!IN_SMEM && packed) {
    // band-packed system from ba_combine_band_kernel (all-reduced in between on the sharded path): [n][W] | b | dA.  The band is
    // the caller's (bw_hint >= 0, not the block-diagonal-damped MoBA variant); no dense matrix is touched.  In shared memory the rows
    // get the odd pitch Ws: threads that walk down a column (panel rows, the back substitution) then hit distinct banks.
    const int band = 6 * min(bw_hint, P - 1), W = band + 6, Ws = W + 1, SCR = ldl_scratch_doubles(band);
    const double* pb = packed + (long)n * W;                  // b, then dA
    PROF_BEGIN;
    PROF_DECL;
    if (tid == 0) s_fail = 0;
    auto damped = [&](double v, int r) { return v + (double)ep + (double)lm * (sem ? pb[n + r] : v); };     // LM damping of a diagonal entry (see damp())
    const bool whole = sizeof(double) * ((size_t)n * Ws + 2 * n + SCR) <= (size_t)kSolveSmem;
    // Two-sided factorisation.  The 63 block columns of the 8-GPU window are 63 SEQUENTIAL steps of a few hundred dependent
    // instructions each (profiles/r02_solve63_source.csv: the kernel is the sum of its serial pieces).  A band matrix splits into
    // top part T | separator S (one band width) | bottom part B with no T-B coupling, so T is eliminated downwards by one half of
    // the block while the other half eliminates B UPWARDS -- run as the same forward code on the index-reversed matrix -- each adding
    // its Schur complement into S; then S is factored, and the back substitutions of T and B again run side by side.  Sequential
    // depth |T| + |S| = 29 + 6 block columns instead of 63.  Every rank executes the same schedule, so replicas stay bitwise equal.
    const int bwb = band / 6, Tb = (P - bwb + 1) / 2, Bb = P - bwb - Tb;
    const bool two_sided = whole && T == 512 && Bb >= 2 && n + band <= 1024 &&
                           sizeof(double) * ((size_t)(n + band) * (Ws + 2) + 2 * SCR) <= (size_t)kSolveSmem;
    if (two_sided) {
      const int n1 = 6 * (Tb + bwb), nB = 6 * Bb, n2 = nB + band;      // matrix 1: rows of T and S; matrix 2 (reversed indices): B then S
      double* A1 = smd;
      double* A2 = A1 + (long)n1 * Ws;
      double* y1 = A2 + (long)n2 * Ws;
      double* y2 = y1 + n1;
      double* d1 = y2 + n2;
      double* d2 = d1 + n1;
      double* sc1 = d2 + n2;
      double* sc2 = sc1 + SCR;
      for (long i = tid; i < (long)n2 * Ws; i += T) A2[i] = 0.0;       // upper triangle, pad column and the S x S block (it collects B's Schur complement)
      for (int i = tid; i < n1; i += T) y1[i] = pb[i];
      for (int l = tid; l < n2; l += T) y2[l] = l < nB ? pb[n - 1 - l] : 0.0;
      for_each_loaded(packed, 0, (long)n1 * W, tid, T, [&](long i, double v) {        // matrix 1 = the first n1 packed rows (damped diagonal)
        const int r = (int)(i / W), cc = (int)(i - (long)r * W);
        if (max(0, 6 * (r / 6) - band) + cc == r) v = damped(v, r);
        A1[(long)r * Ws + cc] = v;
      });
      __syncthreads();
      for_each_loaded(packed, (long)(6 * Tb + band) * W, (long)n * W, tid, T, [&](long i, double v) {      // matrix 2: element (R, C) of the rows of B lands at (n-1-C, n-1-R)
        const int R = (int)(i / W), C = max(0, 6 * (R / 6) - band) + (int)(i - (long)R * W);
        if (C > R) return;
        if (C == R) v = damped(v, R);
        const int r = n - 1 - C, c = n - 1 - R;
        A2[(long)r * Ws + c - max(0, 6 * (r / 6) - band)] = v;
      });
      __syncthreads();
      PROF(9);                                                  // load of both halves
      const int g = tid >> 8, gt = tid & 255;
      const SysMat<true> M1{A1, Ws, band, 0}, M2{A2, Ws, band, 0};
      // warps 0 (top half) and 9 (bottom half) factor: different SM sub-partitions
      if (g == 0) ldl_forward(M1, y1, d1, n1, band, gt, 256, &s_fail, s_rowoff, sc1, 0, Tb, 0, n1, 1, true, false, 0);
      else        ldl_forward(M2, y2, d2, n2, band, gt, 256, &s_fail, s_rowoff + n1, sc2, 0, Bb, 0, n2, 2, true, false, 1);
      __syncthreads();
      PROF(10);                                                 // both eliminations
      for (int idx = tid; idx < band * band; idx += T) {        // S += B's Schur complement (matrix 2 indexes rows in reverse)
        const int a = idx / band, b = idx % band;
        if (b <= a) M1(6 * Tb + a, 6 * Tb + b) += M2(n - 1 - (6 * Tb + b), n - 1 - (6 * Tb + a));
      }
      for (int a = tid; a < band; a += T) y1[6 * Tb + a] += y2[n - 1 - (6 * Tb + a)];
      __syncthreads();
      PROF(11);                                                 // merge
      ldl_forward(M1, y1, d1, n1, band, tid, T, &s_fail, s_rowoff, sc1, Tb, Tb + bwb, 0, n1, 0, true, true);
      PROF(12);                                                 // separator
      if (!s_fail) {
        for (int i = tid; i < n1; i += T) y1[i] *= d1[i];       // D w = z
        for (int l = tid; l < nB; l += T) y2[l] *= d2[l];
        __syncthreads();
        ldl_backward(M1, y1, n1, band, tid, Tb + bwb - 1, Tb);  // x_S
        for (int l = nB + tid; l < n2; l += T) y2[l] = y1[n - 1 - l];
        __syncthreads();
        if (g == 0) ldl_backward(M1, y1, n1, band, gt, Tb - 1, 0, 1, 256);
        else        ldl_backward(M2, y2, n2, band, gt, Bb - 1, 0, 2, 256, 1);
      }
      __syncthreads();
      PROF(13);                                                 // back substitution
      PROF_END;
      const bool failed2 = s_fail != 0;
      for (int i = tid; i < n; i += T) bf.dx[i] = failed2 ? 0.0f : (float)(i < n1 ? y1[i] : y2[n - 1 - i]);
      return;
    }
    if (whole) {
      double* Ab = smd;
      double* yb = smd + (long)n * Ws;
      double* db = yb + n;
      for (int i = tid; i < n; i += T) yb[i] = pb[i];
      for_each_loaded(packed, 0, (long)n * W, tid, T, [&](long i, double v) {
        const int r = (int)(i / W), cc = (int)(i - (long)r * W);
        if (max(0, 6 * (r / 6) - band) + cc == r) v = damped(v, r);
        Ab[(long)r * Ws + cc] = v;
      });
      __syncthreads();
      ldl_solve(SysMat<true>{Ab, Ws, band, 0}, yb, db, P, n, band, tid, T, &s_fail, s_rowoff, db + n);
      __syncthreads();
      const bool failed = s_fail != 0;
      for (int i = tid; i < n; i += T) bf.dx[i] = failed ? 0.0f : (float)yb[i];
      return;
    }
    // The whole band does not fit (BASELINE configs[4]: 63 poses, block bandwidth 16: 378 x 102 doubles = 308 KB), but a WINDOW of
    // rows does: the elimination of block column k touches rows 6k .. 6k + band + 5 only.  Keep `rows` consecutive band rows resident,
    // eliminate as many block columns as they cover, write the rows back (the packed block in global memory is the backing store),
    // slide.  The backward substitution walks the windows in reverse.  Measured before this existed: the factorisation ran out of L2
    // and the BA of the banded-8 graph took 2.1 ms (of which ~1.6 ms solve).
    for (int i = tid; i < n; i += T) {                         // rows are re-read after they were factored: damp in place, once
      double* d = packed + (long)i * W + (i - max(0, 6 * (i / 6) - band));
      *d = damped(*d, i);
    }
    const int rows = min(n, (int)((kWindowBytes - sizeof(double) * 2 * n) / (sizeof(double) * Ws)) / 6 * 6);    // resident rows (>= band + 12)
    const int kcount = (rows - band - 6) / 6 + 1;              // block columns one window can eliminate
    double* Ab = smd;
    double* yb = smd + (long)rows * Ws;
    double* db = yb + n;
    double* scratch = smd + kWindowBytes / sizeof(double);
    for (int i = tid; i < n; i += T) yb[i] = pb[i];
    auto load = [&](int r0, int r1) {
      for (int r = r0 + warp; r < r1; r += nw)
        for (int cc = lane; cc < W; cc += 32) Ab[(long)(r - r0) * Ws + cc] = packed[(long)r * W + cc];
    };
    int k0 = 0;
    for (;; k0 += kcount) {                                    // forward, window by window
      const int r0 = 6 * k0, k1 = min(P, k0 + kcount), r1 = min(n, r0 + rows);
      __syncthreads();
      load(r0, r1);
      __syncthreads();
      ldl_forward(SysMat<true>{Ab, Ws, band, r0}, yb, db, n, band, tid, T, &s_fail, s_rowoff, scratch, k0, k1, r0, r1, 0, k0 == 0, true);
      if (k1 == P) break;
      __syncthreads();
      for (int r = r0 + warp; r < r1; r += nw)
        for (int cc = lane; cc < W; cc += 32) packed[(long)r * W + cc] = Ab[(long)(r - r0) * Ws + cc];
    }
    __syncthreads();
    if (!s_fail) {
      for (int i = tid; i < n; i += T) yb[i] *= db[i];         // D w = z
      __syncthreads();
      for (int kw = k0;; kw -= kcount) {                       // backward: the last window is still resident
        const int r0 = 6 * kw, r1 = min(n, r0 + rows);
        if (kw != k0) {
          __syncthreads();
          load(r0, r1);
          __syncthreads();
        }
        ldl_backward(SysMat<true>{Ab, Ws, band, r0}, yb, n, band, tid, min(P, kw + kcount) - 1, kw);
        if (kw == 0) break;
      }
    }
    __syncthreads();
    const bool failed = s_fail != 0;
    for (int i = tid; i < n; i += T) bf.dx[i] = failed ? 0.0f : (float)yb[i];
    return;
  }
  // ---- dense system: in shared memory (IN_SMEM: up to 26 poses) or, without a caller-supplied bandwidth / for the MoBA variant, in
  // global memory (parts folded into part 0 by ba_combine_kernel beforehand)
  double* A = IN_SMEM ? smd : bf.sys;
  double* y = IN_SMEM ? smd + (long)n * ld : bf.sys + (long)n * n;
  double* dA = IN_SMEM ? y + n : bf.sys + (long)n * n + n;
  if (IN_SMEM) {
    static_assert(kParts == 2, "the in-shared-memory solve folds exactly two replicas");
    const double* part1 = bf.sys + psz;
    for_each_loaded(bf.sys, 0, psz, tid, T, [&](long i, double v) {
      v += __ldcg(part1 + i);
      if (i < (long)n * n) A[(i / n) * ld + (i % n)] = v; else y[i - (long)n * n] = v;      // y | dA are contiguous
    });
  }
  if (tid == 0) s_fail = 0;
  __syncthreads();
  damp(A, ld, dA, P, lm, ep, sem, moba, tid, T);
  __syncthreads();
  // Block bandwidth of the system: frame graphs connect |i - j| <= r, so the Schur complement couples poses at most 2r
  // frames apart and the factorisation never fills outside that band.  Everything below works on rows c0+6 .. rend only:
  // O(n b^2) instead of O(n^3) -- what keeps the redundant per-rank solve cheap when 8 GPUs share a 63-pose window.
  // bw_hint >= 0: the caller knows the graph (block bandwidth <= 2 max|ii - jj|, FactorGraph passes it); -1: scan the matrix.
  __shared__ int s_bw;
  if (tid == 0) s_bw = (bw_hint >= 0 && !moba) ? min(bw_hint, P - 1) : 0;
  __syncthreads();
  for (int b = tid; b < ((bw_hint >= 0 && !moba) ? 0 : P * P); b += T) {
    const int bi = b / P, bj = b % P;
    if (bj >= bi || bi - bj <= s_bw) continue;
    bool nz = false;
    for (int e = 0; e < 36 && !nz; e++) nz = A[(long)(6 * bi + e / 6) * ld + 6 * bj + e % 6] != 0.0;
    if (nz) atomicMax(&s_bw, bi - bj);
  }
  __syncthreads();
  const int band = 6 * s_bw;
  const int SCR = ldl_scratch_doubles(band);
  double* dinv = dA;                                       // dA is dead after damp(): reuse it for 1/d_j
  if (IN_SMEM) {
    ldl_solve(SysMat<false>{A, ld, band, 0}, y, dinv, P, n, band, tid, T, &s_fail, s_rowoff, dA + n);      // the host sized the scratch for band = n
  } else if (sizeof(double) * ((size_t)n * (band + 7) + 2 * n + SCR) <= (size_t)kSolveSmem) {
    // too large for a dense copy, but the band (found by the scan) fits: pack it into shared memory
    const int Ws = band + 7;
    double* Ab = smd;
    double* yb = smd + (long)n * Ws;
    double* db = yb + n;
    for (int r = warp; r < n; r += nw) {
      const int cs = max(0, 6 * (r / 6) - band);
      for (int cc = lane; cc < Ws; cc += 32) Ab[(long)r * Ws + cc] = (cs + cc <= r) ? A[(long)r * ld + cs + cc] : 0.0;
    }
    for (int i = tid; i < n; i += T) yb[i] = y[i];
    __syncthreads();
    ldl_solve(SysMat<true>{Ab, Ws, band, 0}, yb, db, P, n, band, tid, T, &s_fail, s_rowoff, db + n);
    __syncthreads();
    for (int i = tid; i < n; i += T) y[i] = yb[i];
    __syncthreads();
  } else {
    // wide or unknown structure in global memory (rare: the MoBA variant of a large window): generic path, panel scratch when it fits
    ldl_solve(SysMat<false>{A, ld, band, 0}, y, dinv, P, n, band, tid, T, &s_fail, s_rowoff,
              (n <= 1024 && sizeof(double) * SCR <= (size_t)kSolveSmem) ? smd : nullptr);
  }
  const bool fail = s_fail != 0;
  for (int i = tid; i < n; i += T) bf.dx[i] = fail ? 0.0f : (float)y[i];
  __syncthreads();
  // leave every replica zeroed for the next iteration / next call (large systems: ba_zero_kernel, many blocks)
  if (IN_SMEM)
    for (long i = tid; i < (long)kParts * psz; i += T) bf.sys[i] = 0.0;
}

__global__ void __launch_bounds__(256) ba_zero_kernel(Bufs bf, int P) {
  const long tot = (long)kParts * (36L * P * P + 12L * P);
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += (long)gridDim.x * blockDim.x) bf.sys[i] = 0.0;
}

// ---------------------------------------------------------------------------------------------------
// pass 3: dz = Q (w - sum_rows E^T dx) [rows with pose index <= 0 skipped, EvT6x1_kernel :1084], disps += dz,
// poses <- Exp(dx) * poses
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kPix) ba_pass3_kernel(float* __restrict__ poses, float* __restrict__ disps,
                                                        const long* __restrict__ jj, Plan pl, Bufs bf,
                                                        float* __restrict__ dz_out, int HW, int t0, int P,
                                                        int motion_only, int sem) {
  const int K = pl.meta[0];
  const int k = blockIdx.y;
  const int tid = threadIdx.x;
  if (blockIdx.x == 0 && k == 0) {
    for (int p = tid; p < P; p += kPix) {
      float xi[6], T[7];
#pragma unroll
      for (int m = 0; m < 6; m++) xi[m] = bf.dx[6 * p + m];
#pragma unroll
      for (int m = 0; m < 7; m++) T[m] = poses[7 * (t0 + p) + m];
      if (sem) lt_retr(xi, T); else dk_retr(xi, T);
#pragma unroll
      for (int m = 0; m < 7; m++) poses[7 * (t0 + p) + m] = T[m];
    }
  }
  if (motion_only || k >= K) return;
  const int f = pl.kx[k];
  const int pix = blockIdx.x * kPix + tid;
  if (pix >= HW) return;
  float acc = 0.f;
  const int pmin = sem ? 0 : 1;          // EvT6x1_kernel :1084 drops pose index <= 0; chol.py:69 uses every pose
  {
    const int p = f - t0;
    if (p >= pmin && p < P) {
#pragma unroll
      for (int m = 0; m < 6; m++) acc += bf.Ei[((long)k * 6 + m) * HW + pix] * bf.dx[6 * p + m];
    }
  }
  for (int it = pl.src_ptr[k]; it < pl.src_ptr[k + 1]; it++) {
    const int e = pl.src_idx[it];
    const int p = (int)jj[e] - t0;
    if (p >= pmin && p < P) {
      float dw = 0.f;
#pragma unroll
      for (int m = 0; m < 6; m++) dw += bf.Eij[((long)e * 6 + m) * HW + pix] * bf.dx[6 * p + m];
      acc += dw;
    }
  }
  const float dz = bf.Q[(long)k * HW + pix] * (bf.w[(long)k * HW + pix] - acc);
  float dnew = disps[(long)f * HW + pix] + dz;
  if (sem) { dnew = (dnew > 10.0f) ? 0.0f : dnew; dnew = fmaxf(dnew, 0.0f); }      // geom/ba.py:103-104
  disps[(long)f * HW + pix] = dnew;
  if (dz_out) dz_out[(long)k * HW + pix] = dz;
}

__global__ void clamp_min_kernel(float* x, long n, float lo) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] = fmaxf(x[i], lo);
}

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct Layout {
  size_t off_frame2k, off_kx, off_ptr, off_idx, off_meta, off_Eij, off_Ei, off_Q, off_w, off_sys, off_packed, off_dx, total;
};

Layout make_layout(int E, int nframes, int P, int K, int HW) {
  Layout L;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o = align_up(o + bytes, 256); return r; };
  L.off_frame2k = take(sizeof(int) * (size_t)nframes);
  L.off_kx = take(sizeof(int) * (size_t)K);
  L.off_ptr = take(sizeof(int) * (size_t)(K + 1));
  L.off_idx = take(sizeof(int) * (size_t)(E > 0 ? E : 1));
  L.off_meta = take(sizeof(int) * 8);
  L.off_Eij = take(sizeof(float) * 6 * (size_t)E * HW);
  L.off_Ei = take(sizeof(float) * 6 * (size_t)K * HW);
  L.off_Q = take(sizeof(float) * (size_t)K * HW);
  L.off_w = take(sizeof(float) * (size_t)K * HW);
  L.off_sys = take(sizeof(double) * kParts * (36 * (size_t)P * P + 12 * (size_t)P));
  L.off_packed = take(sizeof(double) * (36 * (size_t)P * P + 12 * (size_t)P));      // band-packed system (<= one dense part)
  L.off_dx = take(sizeof(float) * 6 * (size_t)P);
  L.total = o;
  return L;
}

// The band-packed exchange applies when the caller states the block bandwidth (flags bits 20-27), the system is too large for the dense
// shared-memory solve, it is not the MoBA variant (whose damping fills every block), and a band window fits shared memory.
// Returns the band in scalars (6 x block bandwidth), or -1.
int packed_band(int P, int flags) {
  const int n = 6 * P;
  const int bw_hint = ((flags >> 20) & 0xFF) - 1;
  if (bw_hint < 0) return -1;
  if (sizeof(double) * ((size_t)n * (n + 1) + 2 * n) <= 200 * 1024) return -1;        // dense solve in shared memory
  if (n > 1024) return -1;
  const int band = 6 * (bw_hint < P - 1 ? bw_hint : P - 1);
  if (!window_fits(n, band)) return -1;
  return band;
}

}  // namespace

extern "C" long pvo_ba_workspace_bytes(int E, int nframes, int P, int K, int HW) {
  if (E < 0 || nframes <= 0 || P <= 0 || K <= 0 || HW <= 0) return -1;
  return (long)make_layout(E, nframes, P, K, HW).total;
}

// poses (nframes,7) and disps (nframes,ht,wd) are updated IN PLACE.  targets/weights: (E,2,ht,wd) when
// interleaved == 0 (the reference's layout, factor_graph.py:297-300) or (E,ht,wd,2) when interleaved == 1.
// eta: (K,ht,wd) with K = |unique(arange(t0,t1) U ii)|, rows ordered by frame index.  dx_out (P,6), dz_out (K,HW)
// may be null.  flags bit0: skip the plan kernel (workspace already holds the plan for this ii/jj/t0/t1).
extern "C" int pvo_ba(float* poses, float* disps, const float* intrinsics, const float* targets, const float* weights,
                      const float* eta, const long* ii, const long* jj, int E, int nframes, int ht, int wd, int t0,
                      int t1, int K, int iterations, float lm, float ep, int motion_only, int interleaved,
                      float* dx_out, float* dz_out, void* workspace, long workspace_bytes, int flags, void* stream) {
  // flags bits 8-19: optional hint = max out-degree of the graph (0 = unknown); only sizes pass 2's shared memory
  const int P = t1 - t0, HW = ht * wd;
  if (E <= 0 || P <= 0 || K <= 0 || nframes <= 0 || t0 < 0 || t1 > nframes || HW <= 0) return PVO_ERR_INVALID;
  const Layout L = make_layout(E, nframes, P, K, HW);
  if (!workspace || workspace_bytes < (long)L.total) return PVO_ERR_WORKSPACE;
  char* ws = (char*)workspace;
  Plan pl{(int*)(ws + L.off_frame2k), (int*)(ws + L.off_kx), (int*)(ws + L.off_ptr), (int*)(ws + L.off_idx),
          (int*)(ws + L.off_meta)};
  Bufs bf{(float*)(ws + L.off_Eij), (float*)(ws + L.off_Ei), (float*)(ws + L.off_Q), (float*)(ws + L.off_w),
          (double*)(ws + L.off_sys), (float*)(ws + L.off_dx)};
  cudaStream_t s = (cudaStream_t)stream;
  const int n = 6 * P;
  // flags bit 4: the caller guarantees that the system replicas are already zero (a workspace nobody else uses: every call leaves
  // them zeroed) -- skips the 2.3 MB single-block clear at P = 63
  if (!(flags & 1)) ba_plan_kernel<<<1, 1024, 0, s>>>(ii, E, nframes, t0, t1, K, pl, (flags & 16) ? nullptr : bf.sys, n);
  const dim3 grid(ceil_div(HW, kPix), K);
  static bool attr_set = false;
  const int deg_hint = (flags >> 8) & 0xFFF;
  const int bw_hint = ((flags >> 20) & 0xFF) - 1;          // bits 20-27: 1 + block bandwidth of the pose system (0 = unknown: scan)
  int pass2_smem = 96 * 1024;   // out-degree unknown (the reference's own FactorGraph passes no hint): two blocks per SM; frames of higher degree take several pixel chunks
  if (deg_hint > 0) {        // two tiles [128][6(deg+1)] + w + row table; smaller request => several blocks per SM, one wave
    const long need = 4L * (128L * (12 * (deg_hint + 1) + 2) + ((deg_hint + 1 + 3) & ~3)) + 64;
    if (need < pass2_smem) pass2_smem = (int)need;
  }
  const bool solve_in_smem = sizeof(double) * ((size_t)n * (n + 1) + 2 * n) <= 200 * 1024;
  const size_t solve_smem = sizeof(double) * ((size_t)n * (n + 1) + 2 * n + ldl_scratch_doubles(n));     // + panel scratch for any band
  if (!attr_set) {
    cudaFuncSetAttribute(ba_pass2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    cudaFuncSetAttribute(ba_solve_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 216 * 1024);
    cudaFuncSetAttribute(ba_solve_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSolveSmem);
    attr_set = true;
  }
  // flags bits 1-2 select a phase for the multi-GPU path (pvo_b200/dist.py): 0 = whole iteration(s),
  // 1 = assemble only (pass 1 + Schur, replicas folded into part 0 which is then all-reduced),
  // 2 = solve + back-substitution + retraction only.
  const int phase = (flags >> 1) & 3;
  const int sem = (flags >> 3) & 1;      // 0: droid_backends.ba semantics, 1: geom/ba.py BA()/MoBA() semantics
  const int pband = (sem && motion_only) ? -1 : packed_band(P, flags);
  double* packed = (double*)(ws + L.off_packed);
  for (int it = 0; it < iterations; it++) {
    if (phase != 2) {
      ba_pass1_kernel<<<grid, kPix, 0, s>>>(poses, disps, intrinsics, targets, weights, eta, ii, jj, pl, bf, E, HW, wd,
                                            t0, P, interleaved, motion_only, sem);
      if (!motion_only)
        ba_pass2_kernel<<<dim3(ceil_div(HW, kPix2), K), 256, pass2_smem, s>>>(jj, pl, bf, HW, t0, P, pass2_smem / 4);
    }
    const int comb_blocks = ceil_div(36L * P * P + 12L * P, 2048) < 148 ? (ceil_div(36L * P * P + 12L * P, 2048) < 8 ? 8 : ceil_div(36L * P * P + 12L * P, 2048)) : 148;
    if (pband >= 0) {
      // band-packed: fold + clear -> [all-reduce of the packed block between the phases] -> solve straight from it
      if (phase != 2) ba_combine_band_kernel<<<comb_blocks, 256, 0, s>>>(bf, packed, P, pband);
      if (phase == 1) break;
      ba_solve_kernel<false><<<1, 512, kSolveSmem, s>>>(bf, P, lm, ep, sem, 0, bw_hint, packed);
    } else if (phase == 1) {
      ba_combine_kernel<<<comb_blocks, 256, 0, s>>>(bf, P);
      break;
    } else if (solve_in_smem) {
      // small windows are latency bound (128 threads); from ~a dozen poses on the trailing updates (dense windows of the frontend: O(n^2) entries
      // per block column) want the whole SM
      ba_solve_kernel<true><<<1, P <= 10 ? 128 : 512, solve_smem, s>>>(bf, P, lm, ep, sem, sem && motion_only, bw_hint, nullptr);
    } else {
      if (phase != 2) ba_combine_kernel<<<comb_blocks, 256, 0, s>>>(bf, P);      // phase 2: phase 1 already folded the replicas
      ba_solve_kernel<false><<<1, 512, kSolveSmem, s>>>(bf, P, lm, ep, sem, sem && motion_only, bw_hint, nullptr);   // band-packed copy in smem when it fits
      ba_zero_kernel<<<comb_blocks, 256, 0, s>>>(bf, P);
    }
    const bool last = it == iterations - 1;
    if (last && dx_out)
      cudaMemcpyAsync(dx_out, bf.dx, sizeof(float) * n, cudaMemcpyDeviceToDevice, s);
    ba_pass3_kernel<<<grid, kPix, 0, s>>>(poses, disps, jj, pl, bf, last ? dz_out : nullptr, HW, t0, P, motion_only, sem);
    if (phase == 2) break;
  }
  return pvo_check_launch();
}

// torch.unique(ii) + grouping of edges by source frame, for GraphAgg's scatter_mean (droid_net.py:83-88)
extern "C" int pvo_group_edges(const long* ii, int E, int nframes, int t0, int t1, int Kcap, int* frame2k, int* kx,
                               int* ptr, int* idx, int* meta, void* stream) {
  if (E < 0 || nframes <= 0 || Kcap <= 0) return PVO_ERR_INVALID;
  Plan pl{frame2k, kx, ptr, idx, meta};
  ba_plan_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(ii, E, nframes, t0, t1, Kcap, pl, nullptr, 0);
  return pvo_check_launch();
}

// DepthVideo.ba's trailing disps.clamp_(min=0.001) (depth_video.py:214)
extern "C" int pvo_clamp_min(float* x, long n, float lo, void* stream) {
  if (n <= 0) return PVO_OK;
  clamp_min_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(x, n, lo);
  return pvo_check_launch();
}

// byte range of the reduced camera system {H (6P x 6P doubles), b (6P doubles)} inside the workspace: the only data
// exchanged between GPUs when edges are sharded by source keyframe
extern "C" int pvo_ba_system_range(int E, int nframes, int P, int K, int HW, long* offset, long* bytes) {
  if (E < 0 || nframes <= 0 || P <= 0 || K <= 0 || HW <= 0) return PVO_ERR_INVALID;
  const Layout L = make_layout(E > 0 ? E : 1, nframes, P, K, HW);
  *offset = (long)L.off_sys;                                  // part 0 (ba_combine_kernel folds the other parts into it)
  *bytes = (long)(sizeof(double) * (36 * (size_t)P * P + 12 * (size_t)P));
  return PVO_OK;
}

// What the ranks exchange for a pvo_ba call with these flags / motion_only: the band-packed block {[n][band + 6] | b | dA} when
// packed_band() applies (P = 63, radius 3: 133 KB), else part 0 of the dense system as pvo_ba_system_range reports it.
extern "C" int pvo_ba_exchange_range(int E, int nframes, int P, int K, int HW, int flags, int motion_only, long* offset, long* bytes) {
  if (E < 0 || nframes <= 0 || P <= 0 || K <= 0 || HW <= 0) return PVO_ERR_INVALID;
  const int sem = (flags >> 3) & 1;
  const int pband = (sem && motion_only) ? -1 : packed_band(P, flags);
  if (pband < 0) return pvo_ba_system_range(E, nframes, P, K, HW, offset, bytes);
  const Layout L = make_layout(E > 0 ? E : 1, nframes, P, K, HW);
  *offset = (long)L.off_packed;
  *bytes = (long)(sizeof(double) * ((size_t)6 * P * (pband + 6) + 12 * (size_t)P));
  return PVO_OK;
}

#ifdef PVO_SOLVE_PROF
extern "C" int pvo_solve_prof_read(long long* host32, int reset) {
  if (cudaMemcpyFromSymbol(host32, g_solve_prof, sizeof(long long) * 32) != cudaSuccess) return PVO_ERR_CUDA;
  if (reset) { long long z[32] = {}; cudaMemcpyToSymbol(g_solve_prof, z, sizeof(z)); }
  return PVO_OK;
}
#endif

extern "C" int pvo_ba_read_meta(const void* workspace, int E, int nframes, int P, int K, int HW, int* meta_host3) {
  const Layout L = make_layout(E, nframes, P, K, HW);
  return cudaMemcpy(meta_host3, (const char*)workspace + L.off_meta, 3 * sizeof(int), cudaMemcpyDeviceToHost) ==
                 cudaSuccess ? PVO_OK : PVO_ERR_CUDA;
}
