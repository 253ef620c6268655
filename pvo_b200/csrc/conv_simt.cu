// Implicit-GEMM convolution on CUDA cores (FP32 accumulate) with fused epilogues, pixel-major (NHWC) activations.
//
// This is the float32 PARITY path of the update operator -- every nn.Conv2d of DynamicUpdateModule / ConvGRU /
// GraphAgg (VO_Module/droid_slam/droid_net.py:167-225,64-95; modules/gru.py:6-32) -- and the path for the small or
// oddly shaped layers in half mode.  The dense 3x3 layers in half mode run on tcgen05 (conv_tc.cu).
// The reference calls cuDNN once per conv and then separate elementwise kernels for bias/ReLU/sigmoid/tanh, the
// torch.cat of the GRU inputs and the gate arithmetic; here the concatenation is virtual (input = up to 4 channel
// segments with their own strides) and the gate arithmetic is the epilogue.
//
//   GEMM view: M = pixels (all edges), N = Cout, K = KH*KW*Cin, k = tap*Cin + c.
//   Tile 128 x BN x 16, 256 threads, 8 x (BN/16) outputs per thread, double-buffered smem.
#include "common.cuh"

namespace {

struct Seg {
  const void* ptr;   // element (m, c) at ptr[m*stride + c]
  int channels;
  int stride;
};

struct ConvP {
  Seg seg[4];
  int nseg;
  int Cin;           // sum of segment channels
  const float* W;    // [KH*KW*Cin][Wstride], Wstride % 4 == 0, zero padded
  int Wstride;
  int Cout;
  int KH, KW, padH, padW;
  int H, Wd;         // INPUT spatial size
  int Ho, Wo;        // OUTPUT spatial size; M = nimg * Ho * Wo
  int stride;
  long M;
};

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

template <typename T> __device__ __forceinline__ float ldf(const T* p);
template <> __device__ __forceinline__ float ldf<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float ldf<__half>(const __half* p) { return __half2float(__ldg(p)); }
template <typename T> __device__ __forceinline__ void stf(T* p, float v);
template <> __device__ __forceinline__ void stf<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void stf<__half>(__half* p, float v) { *p = __float2half_rn(v); }

// ------------------------------------------------------------------------------------------------ epilogues
// store<V>(m, n, acc): V consecutive output channels n..n+V-1 of pixel m.
template <typename TO>
struct EpiBias {      // out = act(acc + bias); with a residual: relu(res + act(acc + bias))  (extractor.py:47-55)
  const float* bias; TO* out; int ostride; int ooff; int Cout; int relu; const TO* res; int rstride;
  template <int V> __device__ __forceinline__ void store(long m, int n, const float* a) const {
#pragma unroll
    for (int j = 0; j < V; j++) {
      if (n + j < Cout) {
        float v = a[j] + bias[n + j];
        if (relu) v = fmaxf(v, 0.0f);
        if (res) v = fmaxf(v + ldf(res + m * rstride + n + j), 0.0f);
        stf(out + m * ostride + ooff + n + j, v);
      }
    }
  }
};

template <typename TO>
struct EpiGlo {       // gru.py:25: sigmoid(w(net)) * net       (the spatial mean is taken by colmean_kernel)
  const float* bias; const TO* net; int nstride; TO* out; int Cout;
  template <int V> __device__ __forceinline__ void store(long m, int n, const float* a) const {
#pragma unroll
    for (int j = 0; j < V; j++)
      if (n + j < Cout) stf(out + m * Cout + n + j, sigmoidf_(a[j] + bias[n + j]) * ldf(net + m * nstride + n + j));
  }
};

template <typename TO>
struct EpiZR {        // gru.py:28-29: z | r = sigmoid(conv([net,inp]) + conv_glo(glo)); also emits r*net
  const float* bias; const float* glo /*[E][384] = gz|gr|gq*/; const TO* net; int nstride; TO* z; TO* rnet; int HW;
  template <int V> __device__ __forceinline__ void store(long m, int n, const float* a) const {
    const long e = m / HW;
#pragma unroll
    for (int j = 0; j < V; j++) {
      const int c = n + j;
      if (c >= 256) continue;
      const float s = sigmoidf_(a[j] + bias[c] + glo[e * 384 + c]);
      if (c < 128) stf(z + m * 128 + c, s);
      else stf(rnet + m * 128 + (c - 128), s * ldf(net + m * nstride + (c - 128)));
    }
  }
};

template <typename TO>
struct EpiQ {         // gru.py:30-32: q = tanh(conv([r*net,inp]) + glo); net = (1-z)*net + z*q   (in place)
  const float* bias; const float* glo; const TO* z; TO* net; int nstride; int HW;
  template <int V> __device__ __forceinline__ void store(long m, int n, const float* a) const {
    const long e = m / HW;
#pragma unroll
    for (int j = 0; j < V; j++) {
      const int c = n + j;
      if (c >= 128) continue;
      const float q = tanhf(a[j] + bias[c] + glo[e * 384 + 256 + c]);
      const float zz = ldf(z + m * 128 + c);
      const float h = ldf(net + m * nstride + c);
      stf(net + m * nstride + c, (1.0f - zz) * h + zz * q);
    }
  }
};

// ------------------------------------------------------------------------------------------------ main kernel
template <typename TI, int BN, bool VEC, typename Epi>
__global__ void __launch_bounds__(256) conv_simt_kernel(const __grid_constant__ ConvP p, const Epi epi) {
  constexpr int BM = 128, BK = 16, TN = BN / 16;
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN];
  const int tid = threadIdx.x;
  const long m0 = (long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int Ktot = p.KH * p.KW * p.Cin;
  const int HW = p.H * p.Wd, HWo = p.Ho * p.Wo;
  // A-tile ownership: output pixel row am, 8 consecutive k starting at ak
  const int am = tid >> 1, ak = (tid & 1) * 8;
  const long m = m0 + am;
  const bool mvalid = m < p.M;
  const long img = mvalid ? m / HWo : 0;
  const int rem = mvalid ? (int)(m - img * HWo) : 0;
  const int py = (rem / p.Wo) * p.stride, px = (rem - (rem / p.Wo) * p.Wo) * p.stride;

  float ra[8];
  float4 rb[(BK * BN / 4 + 255) / 256];

  auto fetch = [&](int k0) {
    // ---- A
    if (VEC) {
      const int kb = k0 + ak;
      bool ok = mvalid && kb < Ktot;
      const int tap = kb / p.Cin;
      int c = kb - tap * p.Cin;
      const int dy = tap / p.KW, dxx = tap - dy * p.KW;
      const int iy = py + dy - p.padH, ix = px + dxx - p.padW;
      ok = ok && iy >= 0 && iy < p.H && ix >= 0 && ix < p.Wd;
      if (ok) {
        int s = 0;
        while (s < p.nseg - 1 && c >= p.seg[s].channels) { c -= p.seg[s].channels; s++; }
        const long pm = img * HW + (long)iy * p.Wd + ix;
        const TI* src = reinterpret_cast<const TI*>(p.seg[s].ptr) + pm * p.seg[s].stride + c;
        if (sizeof(TI) == 4) {
          const float4 v0 = __ldg(reinterpret_cast<const float4*>(src));
          const float4 v1 = __ldg(reinterpret_cast<const float4*>(src) + 1);
          ra[0] = v0.x; ra[1] = v0.y; ra[2] = v0.z; ra[3] = v0.w;
          ra[4] = v1.x; ra[5] = v1.y; ra[6] = v1.z; ra[7] = v1.w;
        } else {
          const uint4 v = __ldg(reinterpret_cast<const uint4*>(src));
          const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
          for (int q = 0; q < 4; q++) {
            const float2 f = __half22float2(h[q]);
            ra[2 * q] = f.x; ra[2 * q + 1] = f.y;
          }
        }
      } else {
#pragma unroll
        for (int q = 0; q < 8; q++) ra[q] = 0.0f;
      }
    } else {
#pragma unroll
      for (int q = 0; q < 8; q++) {
        const int k = k0 + ak + q;
        float v = 0.0f;
        if (mvalid && k < Ktot) {
          const int tap = k / p.Cin;
          int c = k - tap * p.Cin;
          const int dy = tap / p.KW, dxx = tap - dy * p.KW;
          const int iy = py + dy - p.padH, ix = px + dxx - p.padW;
          if (iy >= 0 && iy < p.H && ix >= 0 && ix < p.Wd) {
            int s = 0;
            while (s < p.nseg - 1 && c >= p.seg[s].channels) { c -= p.seg[s].channels; s++; }
            const long pm = img * HW + (long)iy * p.Wd + ix;
            v = ldf(reinterpret_cast<const TI*>(p.seg[s].ptr) + pm * p.seg[s].stride + c);
          }
        }
        ra[q] = v;
      }
    }
    // ---- B
#pragma unroll
    for (int i = 0; i < (BK * BN / 4 + 255) / 256; i++) {
      const int idx = tid + i * 256;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (idx < BK * BN / 4) {
        const int bk = idx / (BN / 4), bn = (idx - bk * (BN / 4)) * 4;
        const int k = k0 + bk;
        if (k < Ktot && n0 + bn < p.Wstride) v = __ldg(reinterpret_cast<const float4*>(p.W + (long)k * p.Wstride + n0 + bn));
      }
      rb[i] = v;
    }
  };
  auto stash = [&](int buf) {
#pragma unroll
    for (int q = 0; q < 8; q++) As[buf][ak + q][am] = ra[q];
#pragma unroll
    for (int i = 0; i < (BK * BN / 4 + 255) / 256; i++) {
      const int idx = tid + i * 256;
      if (idx < BK * BN / 4) {
        const int bk = idx / (BN / 4), bn = (idx - bk * (BN / 4)) * 4;
        *reinterpret_cast<float4*>(&Bs[buf][bk][bn]) = rb[i];
      }
    }
  };

  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int j = 0; j < TN; j++) acc[i][j] = 0.0f;

  const int ty = tid >> 4, tx = tid & 15;
  const int nk = (Ktot + BK - 1) / BK;
  fetch(0);
  stash(0);
  __syncthreads();
  for (int t = 0; t < nk; t++) {
    const int buf = t & 1;
    if (t + 1 < nk) fetch((t + 1) * BK);
#pragma unroll
    for (int kk = 0; kk < BK; kk++) {
      float a[8], b[TN];
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 8 + 4]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
      if (TN == 8) {
        const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
        const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][BN / 2 + tx * 4]);
        b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
        b[4 % TN] = b1.x; b[5 % TN] = b1.y; b[6 % TN] = b1.z; b[7 % TN] = b1.w;
      } else if (TN == 4) {
        const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
        b[0] = b0.x; b[1 % TN] = b0.y; b[2 % TN] = b0.z; b[3 % TN] = b0.w;
      } else {
        const float2 b0 = *reinterpret_cast<const float2*>(&Bs[buf][kk][tx * 2]);
        b[0] = b0.x; b[1 % TN] = b0.y;
      }
#pragma unroll
      for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < TN; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (t + 1 < nk) stash(buf ^ 1);
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const long mo = m0 + ty * 8 + i;
    if (mo >= p.M) continue;
    if (TN == 8) {
      epi.template store<4>(mo, n0 + tx * 4, &acc[i][0]);
      epi.template store<4>(mo, n0 + BN / 2 + tx * 4, &acc[i][4 % TN]);
    } else {
      epi.template store<TN>(mo, n0 + tx * TN, &acc[i][0]);
    }
  }
}

// ------------------------------------------------------------------------------------------------ small-N grouped 3x3
// out[m][g*CG + o] = post(bias + sum_{tap,c} in[m@tap][g*cin_g + c] * W[g][tap][c][o]),  CG <= 2: the four 128->2 head
// convolutions (droid_net.py:189-214, one group per head) and GraphAgg's 128->1 eta head (:72-75).
// post: 0 none, 1 = 0.01 * softplus (droid_net.py:95 with nn.Softplus defaults beta=1, threshold=20).
// HBM-bound (each input read once): block = one 8x16 pixel patch of one group; the 10x18 halo patch is staged in
// shared memory (pixel stride padded by 16 B so the per-pixel 16-byte reads are bank-conflict free), weights in smem.
template <typename TI, int CG>
__global__ void __launch_bounds__(128) conv3x3_small_kernel(const TI* __restrict__ in, int istride, int cin_g,
                                                            const float* __restrict__ W, const float* __restrict__ bias,
                                                            float* __restrict__ out, int ostride, int G, int H, int Wd,
                                                            int tiles_x, int tiles_y, int post) {
  constexpr int TH = 8, TW = 16, PH = TH + 2, PW = TW + 2;
  constexpr int CH = 64;                                        // channels staged per pass (keeps 6 blocks / SM resident)
  extern __shared__ __align__(16) unsigned char sm_raw[];
  constexpr int pstride = CH + 16 / (int)sizeof(TI);            // elements per patch pixel (+16 B pad: conflict-free LDS.128)
  TI* patch = reinterpret_cast<TI*>(sm_raw);
  float* wsm = reinterpret_cast<float*>(sm_raw + (((size_t)PH * PW * pstride * sizeof(TI) + 15) & ~size_t(15)));
  const int g = blockIdx.y;
  const int tile = blockIdx.x;
  const int img = tile / (tiles_x * tiles_y);
  const int r = tile - img * (tiles_x * tiles_y);
  const int y0 = (r / tiles_x) * TH, x0 = (r % tiles_x) * TW;
  const int tid = threadIdx.x;
  for (int i = tid; i < 9 * cin_g * CG; i += 128) wsm[i] = W[(long)g * 9 * cin_g * CG + i];
  const int ty = tid / TW, tx = tid % TW;
  float acc[CG][4];                                             // 4 independent chains per output (ILP)
#pragma unroll
  for (int o = 0; o < CG; o++)
#pragma unroll
    for (int k = 0; k < 4; k++) acc[o][k] = 0.0f;
  constexpr int cpp = CH * (int)sizeof(TI) / 16;                // 16-byte chunks per patch pixel
  for (int cb = 0; cb < cin_g; cb += CH) {
    __syncthreads();
    for (int i = tid; i < PH * PW * cpp; i += 128) {
      const int pp = i / cpp, ch = i - pp * cpp;
      const int py = y0 - 1 + pp / PW, px = x0 - 1 + pp % PW;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (py >= 0 && py < H && px >= 0 && px < Wd && cb + ch * (16 / (int)sizeof(TI)) < cin_g)
        v = __ldg(reinterpret_cast<const uint4*>(in + ((long)(img * H + py) * Wd + px) * istride + g * cin_g + cb) + ch);
      *reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(patch + (long)pp * pstride) + 16 * ch) = v;
    }
    __syncthreads();
#pragma unroll
    for (int tap = 0; tap < 9; tap++) {
      const TI* src = patch + (long)((ty + tap / 3) * PW + tx + tap % 3) * pstride;
      const float* w = wsm + (tap * cin_g + cb) * CG;
#pragma unroll
      for (int c = 0; c < CH; c += 8) {
        float v[8];
        if (sizeof(TI) == 4) {
          const float4 v0 = *reinterpret_cast<const float4*>(src + c);
          const float4 v1 = *reinterpret_cast<const float4*>(src + c + 4);
          v[0] = v0.x; v[1] = v0.y; v[2] = v0.z; v[3] = v0.w; v[4] = v1.x; v[5] = v1.y; v[6] = v1.z; v[7] = v1.w;
        } else {
          const uint4 u = *reinterpret_cast<const uint4*>(src + c);
          const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
          for (int q = 0; q < 4; q++) { const float2 f = __half22float2(h[q]); v[2 * q] = f.x; v[2 * q + 1] = f.y; }
        }
#pragma unroll
        for (int q = 0; q < 8; q++) {
          if (CG == 2) {
            const float2 ww = *reinterpret_cast<const float2*>(w + (c + q) * 2);
            acc[0][q & 3] = fmaf(v[q], ww.x, acc[0][q & 3]);
            acc[CG - 1][q & 3] = fmaf(v[q], ww.y, acc[CG - 1][q & 3]);
          } else {
            acc[0][q & 3] = fmaf(v[q], w[c + q], acc[0][q & 3]);
          }
        }
      }
    }
  }
  const int y = y0 + ty, x = x0 + tx;
  if (y < H && x < Wd) {
    const long m = (long)(img * H + y) * Wd + x;
#pragma unroll
    for (int o = 0; o < CG; o++) {
      float v = ((acc[o][0] + acc[o][1]) + (acc[o][2] + acc[o][3])) + bias[g * CG + o];
      if (post == 1) v = 0.01f * ((v > 20.0f) ? v : log1pf(expf(v)));
      out[m * ostride + g * CG + o] = v;
    }
  }
}

// Half-precision variant of the grouped head convolution on the legacy tensor-core path (mma.sync m16n8k16, fp32
// accumulate): M = 16 consecutive pixels of a tile row, K = 16 channels of one tap, N = 8 (the group's 2 outputs + 6
// zero columns).  Tiny N makes tcgen05 pointless here (its smallest N is 16 and the operand stream would dominate);
// the CUDA-core version above is FMA/LDS bound at ~4x the time of this one.  A comes from the same staged halo patch
// via ldmatrix, B fragments are pre-arranged in shared memory.
__device__ __forceinline__ void ldmatrix_x4(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, const void* p) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(a));
}
__device__ __forceinline__ void hmma16816(float c[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(128) conv3x3_heads_mma_kernel(const __half* __restrict__ in, int istride, int cin_g,
                                                                const uint2* __restrict__ Wfrag, const float* __restrict__ bias,
                                                                float* __restrict__ out, int ostride, int H, int Wd,
                                                                int tiles_x, int tiles_y) {
  constexpr int TH = 8, TW = 16, PH = TH + 2, PW = TW + 2, CH = 64, CG = 2;
  constexpr int pstride = CH + 8;                               // halves per patch pixel (144 B: ldmatrix conflict-free)
  extern __shared__ __align__(16) unsigned char sm_raw[];
  __half* patch = reinterpret_cast<__half*>(sm_raw);
  uint2* wfrag = reinterpret_cast<uint2*>(sm_raw + (((size_t)PH * PW * pstride * 2 + 15) & ~size_t(15)));   // [9][cin_g/16][32]
  const int g = blockIdx.y;
  const int tile = blockIdx.x;
  const int img = tile / (tiles_x * tiles_y);
  const int r = tile - img * (tiles_x * tiles_y);
  const int y0 = (r / tiles_x) * TH, x0 = (r % tiles_x) * TW;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nchunk = cin_g / 16;
  // B fragments (packed once on the host, nn.pack_heads_mma): lane holds B[k = (lane%4)*2 + {0,1}][n = lane/4] and the
  // same for k+8 as two half2; only n < 2 is non-zero
  for (int i = tid; i < 9 * nchunk * 32; i += 128) wfrag[i] = __ldg(Wfrag + (long)g * 9 * nchunk * 32 + i);
  float acc[2][4];                                               // two 16-pixel rows per warp (tile rows 2*warp, 2*warp+1)
#pragma unroll
  for (int t = 0; t < 2; t++)
#pragma unroll
    for (int q = 0; q < 4; q++) acc[t][q] = 0.0f;
  constexpr int cpp = CH * 2 / 16;
  for (int cb = 0; cb < cin_g; cb += CH) {
    __syncthreads();
    for (int i = tid; i < PH * PW * cpp; i += 128) {
      const int pp = i / cpp, ch = i - pp * cpp;
      const int py = y0 - 1 + pp / PW, px = x0 - 1 + pp % PW;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (py >= 0 && py < H && px >= 0 && px < Wd)
        v = __ldg(reinterpret_cast<const uint4*>(in + ((long)(img * H + py) * Wd + px) * istride + g * cin_g + cb) + ch);
      *reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(patch + (long)pp * pstride) + 16 * ch) = v;
    }
    __syncthreads();
#pragma unroll
    for (int tap = 0; tap < 9; tap++) {
#pragma unroll
      for (int c = 0; c < CH; c += 16) {
        const uint2 b = wfrag[(tap * nchunk + (cb + c) / 16) * 32 + lane];
#pragma unroll
        for (int t = 0; t < 2; t++) {
          const int ty = 2 * warp + t;
          // ldmatrix x4: lanes 0-15 address pixel rows 0..15 at channel c, lanes 16-31 the same rows at channel c+8
          const __half* ap = patch + (long)((ty + tap / 3) * PW + (lane & 15) + tap % 3) * pstride + c + ((lane >> 4) << 3);
          uint32_t a0, a1, a2, a3;
          ldmatrix_x4(a0, a1, a2, a3, ap);
          hmma16816(acc[t], a0, a1, a2, a3, b.x, b.y);
        }
      }
    }
  }
  // C fragment: c0,c1 = (row lane/4, cols (lane%4)*2, +1), c2,c3 = row + 8: the two real outputs sit in lanes with lane%4 == 0
  if ((lane & 3) == 0) {
#pragma unroll
    for (int t = 0; t < 2; t++) {
      const int y = y0 + 2 * warp + t;
#pragma unroll
      for (int hrow = 0; hrow < 2; hrow++) {
        const int x = x0 + (lane >> 2) + 8 * hrow;
        if (y < H && x < Wd) {
          float* o = out + ((long)(img * H + y) * Wd + x) * ostride + g * CG;
          o[0] = acc[t][2 * hrow] + bias[g * CG];
          o[1] = acc[t][2 * hrow + 1] + bias[g * CG + 1];
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ small helpers
// mean over the HW pixels of each image: in [E][HW][C] -> out [E][C]   (gru.py:26).  Two deterministic stages:
// kColChunks partial sums per image (coalesced rows, 2 channels per thread), then a fixed-order final sum.
constexpr int kColChunks = 32;
template <typename TI>
__global__ void __launch_bounds__(64) colsum_partial_kernel(const TI* __restrict__ in, float* __restrict__ part, int HW, int C) {
  const int e = blockIdx.x, ch = blockIdx.y;
  const int per = (HW + kColChunks - 1) / kColChunks;
  const int p0 = ch * per, p1 = min(HW, p0 + per);
  for (int c = 2 * threadIdx.x; c < C; c += 2 * blockDim.x) {
    float s0 = 0.0f, s1 = 0.0f;
    for (int p = p0; p < p1; p++) {
      const TI* q = in + ((long)e * HW + p) * C + c;
      s0 += ldf(q);
      s1 += ldf(q + 1);
    }
    part[((long)e * kColChunks + ch) * C + c] = s0;
    part[((long)e * kColChunks + ch) * C + c + 1] = s1;
  }
}
__global__ void __launch_bounds__(128) colsum_final_kernel(const float* __restrict__ part, float* __restrict__ out, int HW, int C) {
  const int e = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.0f;
#pragma unroll 4
    for (int k = 0; k < kColChunks; k++) s += part[((long)e * kColChunks + k) * C + c];
    out[(long)e * C + c] = s / (float)HW;
  }
}

// nn.InstanceNorm2d (affine=False, eps=1e-5, biased variance; extractor.py:28-32,129-130) on pixel-major data, fused
// with the ReLU that always follows it and, for the last norm of a residual block, the skip addition + ReLU.
// stats: per (image, channel) sum / sum of squares in two deterministic stages (reuses the colsum layout).
template <typename TI>
__global__ void __launch_bounds__(64) in_stats_partial_kernel(const TI* __restrict__ in, float* __restrict__ part, int HW, int C) {
  const int e = blockIdx.x, ch = blockIdx.y;
  const int per = (HW + kColChunks - 1) / kColChunks;
  const int p0 = ch * per, p1 = min(HW, p0 + per);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.0f, q = 0.0f;
    for (int p = p0; p < p1; p++) {
      const float v = ldf(in + ((long)e * HW + p) * C + c);
      s += v;
      q = fmaf(v, v, q);
    }
    part[(((long)e * kColChunks + ch) * C + c) * 2] = s;
    part[(((long)e * kColChunks + ch) * C + c) * 2 + 1] = q;
  }
}
__global__ void __launch_bounds__(128) in_stats_final_kernel(const float* __restrict__ part, float* __restrict__ stats, int HW, int C) {
  const int e = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    double s = 0.0, q = 0.0;
    for (int k = 0; k < kColChunks; k++) {
      s += part[(((long)e * kColChunks + k) * C + c) * 2];
      q += part[(((long)e * kColChunks + k) * C + c) * 2 + 1];
    }
    const double mean = s / HW;
    const double var = fmax(q / HW - mean * mean, 0.0);
    stats[((long)e * C + c) * 2] = (float)mean;
    stats[((long)e * C + c) * 2 + 1] = (float)(1.0 / sqrt(var + 1e-5));
  }
}
template <typename TI>
__global__ void __launch_bounds__(256) in_apply_kernel(const TI* __restrict__ in, const float* __restrict__ stats, const TI* __restrict__ res,
                                                       TI* __restrict__ out, int HW, int C, long total, int relu) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % C);
  const long e = i / ((long)HW * C);
  float v = (ldf(in + i) - stats[(e * C + c) * 2]) * stats[(e * C + c) * 2 + 1];
  if (relu) v = fmaxf(v, 0.0f);
  if (res) v = fmaxf(v + ldf(res + i), 0.0f);
  stf(out + i, v);
}

// flow-encoder input for the tensor-core path: the 7x7 conv over 8 motion channels (droid_net.py:179) is run as a
// 7x1 conv over 64 channels, where channel (dxi*8 + c) of pixel (y,x) holds motion channel c of pixel (y, x+dxi-3)
// (zero outside the image) and channels 56..63 are zero.
__global__ void __launch_bounds__(256) expand_x7_kernel(const __half* __restrict__ motn, __half* __restrict__ out, int Wd, long M) {
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long m = t >> 3;
  const int k = (int)(t & 7);
  if (m >= M) return;
  const int x = (int)(m % Wd);
  uint4 v = make_uint4(0, 0, 0, 0);
  const int xs = x + k - 3;
  if (k < 7 && xs >= 0 && xs < Wd) v = __ldg(reinterpret_cast<const uint4*>(motn) + (m + k - 3));
  reinterpret_cast<uint4*>(out)[m * 8 + k] = v;
}

// the three 1x1 "glo" convolutions on the pooled vector (gru.py:28-30): out[e][g*128+o] = b_g[o] + W_g[o][:] . glo[e]
__global__ void __launch_bounds__(384) glo_matvec_kernel(const float* __restrict__ glo, const float* __restrict__ W /*[3][128][128] (o,c)*/,
                                                         const float* __restrict__ b /*[3][128]*/, float* __restrict__ out) {
  const int e = blockIdx.x;
  __shared__ float g[128];
  if (threadIdx.x < 128) g[threadIdx.x] = glo[(long)e * 128 + threadIdx.x];
  __syncthreads();
  const int row = threadIdx.x;                          // 0..383
  const float* w = W + (long)row * 128;
  float s = b[row];
  for (int c = 0; c < 128; c++) s = fmaf(w[c], g[c], s);
  out[(long)e * 384 + row] = s;
}

// tensor-core path: the gate conv's epilogue leaves per-tile sums in partial[e][tiles][4][128]; finish the spatial
// mean (fixed order => deterministic) and apply the three glo mat-vecs in one launch.
__global__ void __launch_bounds__(384) glo_reduce_matvec_kernel(const float* __restrict__ partial, int parts_per_img, float inv_hw,
                                                                const float* __restrict__ W, const float* __restrict__ b,
                                                                float* __restrict__ out) {
  const int e = blockIdx.x;
  __shared__ __align__(16) float g[128];
  __shared__ float part3[3][128];
  {
    const int c = threadIdx.x & 127, grp = threadIdx.x >> 7;          // 3 groups of 128 threads split the partials
    const float* p = partial + (long)e * parts_per_img * 128 + c;
    float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
    int k = grp;
    for (; k + 9 < parts_per_img; k += 12) {
      s0 += p[(long)k * 128]; s1 += p[(long)(k + 3) * 128]; s2 += p[(long)(k + 6) * 128]; s3 += p[(long)(k + 9) * 128];
    }
    for (; k < parts_per_img; k += 3) s0 += p[(long)k * 128];
    part3[grp][c] = (s0 + s1) + (s2 + s3);
  }
  __syncthreads();
  if (threadIdx.x < 128) g[threadIdx.x] = ((part3[0][threadIdx.x] + part3[1][threadIdx.x]) + part3[2][threadIdx.x]) * inv_hw;
  __syncthreads();
  // 384 x 128 mat-vec: warp w owns rows [32w, 32w+32); a row is one coalesced 512-byte read (float4 per lane)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float4 gv = *reinterpret_cast<const float4*>(&g[4 * lane]);
#pragma unroll 4
  for (int r = 0; r < 32; r++) {
    const int row = warp * 32 + r;
    const float4 wv = __ldg(reinterpret_cast<const float4*>(W + (long)row * 128) + lane);
    float s = wv.x * gv.x + wv.y * gv.y + wv.z * gv.z + wv.w * gv.w;
    s = warp_sum(s);
    if (lane == 0) out[(long)e * 384 + row] = s + b[row];
  }
}

// scatter_mean over edges sharing a source frame (droid_net.py:83-88): out[k][pix][c] = mean_{e in row k} in[e][pix][c]
template <typename T>
__global__ void __launch_bounds__(256) segment_mean_kernel(const T* __restrict__ in, int istride, int ioff,
                                                           T* __restrict__ out, const int* __restrict__ ptr,
                                                           const int* __restrict__ idx, const int* __restrict__ meta,
                                                           int HW, int C) {
  constexpr int V = 16 / (int)sizeof(T);                     // channels per 16-byte access
  const int k = blockIdx.y;
  if (k >= meta[0]) return;
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int cv = C / V;
  if (i >= (long)HW * cv) return;
  const long pix = i / cv;
  const int c = (int)(i - pix * cv) * V;
  float s[V];
#pragma unroll
  for (int q = 0; q < V; q++) s[q] = 0.0f;
  const int b = ptr[k], e_ = ptr[k + 1];
  for (int it = b; it < e_; it++) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(in + ((long)idx[it] * HW + pix) * istride + ioff + c));
    if (sizeof(T) == 2) {
      const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
      for (int q = 0; q < 4; q++) { const float2 f = __half22float2(h[q]); s[2 * q] += f.x; s[(2 * q + 1) % V] += f.y; }
    } else {
      const float* f = reinterpret_cast<const float*>(&u);
#pragma unroll
      for (int q = 0; q < 4; q++) s[q % V] += f[q];
    }
  }
  const float inv = 1.0f / (float)((e_ - b) > 0 ? (e_ - b) : 1);
  T* o = out + ((long)k * HW + pix) * C + c;
#pragma unroll
  for (int q = 0; q < V; q++) stf(o + q, s[q] * inv);
}

// layout / dtype conversion between the reference's planar (N,C,H,W) tensors and pixel-major (N,HW,Cs) buffers
template <typename TS, typename TD>
__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(const TS* __restrict__ src, TD* __restrict__ dst, int C, int HW,
                                                           int dstride, int doff, long total) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {
    const int c = c0 + r, p = p0 + tx;
    tile[r][tx] = (c < C && p < HW) ? ldf(src + ((long)n * C + c) * HW + p) : 0.0f;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int p = p0 + r, c = c0 + tx;
    if (c < C && p < HW) stf(dst + ((long)n * HW + p) * dstride + doff + c, tile[tx][r]);
  }
}

template <typename TS, typename TD>
__global__ void __launch_bounds__(256) nhwc_to_nchw_kernel(const TS* __restrict__ src, int sstride, int soff, TD* __restrict__ dst,
                                                           int C, int HW) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {
    const int p = p0 + r, c = c0 + tx;
    tile[r][tx] = (c < C && p < HW) ? ldf(src + ((long)n * HW + p) * sstride + soff + c) : 0.0f;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int c = c0 + r, p = p0 + tx;
    if (c < C && p < HW) stf(dst + ((long)n * C + c) * HW + p, tile[tx][r]);
  }
}

template <typename TI, typename Epi>
int launch_conv(const ConvP& p, const Epi& epi, bool vec, cudaStream_t s) {
  const int gm = ceil_div(p.M, 128);
  if (p.Cout > 64) {
    dim3 grid(gm, ceil_div(p.Cout, 128));
    if (vec) conv_simt_kernel<TI, 128, true, Epi><<<grid, 256, 0, s>>>(p, epi);
    else conv_simt_kernel<TI, 128, false, Epi><<<grid, 256, 0, s>>>(p, epi);
  } else if (p.Cout > 32) {
    dim3 grid(gm, 1);
    if (vec) conv_simt_kernel<TI, 64, true, Epi><<<grid, 256, 0, s>>>(p, epi);
    else conv_simt_kernel<TI, 64, false, Epi><<<grid, 256, 0, s>>>(p, epi);
  } else {
    dim3 grid(gm, 1);
    if (vec) conv_simt_kernel<TI, 32, true, Epi><<<grid, 256, 0, s>>>(p, epi);
    else conv_simt_kernel<TI, 32, false, Epi><<<grid, 256, 0, s>>>(p, epi);
  }
  return pvo_check_launch();
}

}  // namespace

// ===================================================================================================
// C ABI
// ===================================================================================================
// Generic convolution.  Activations are pixel-major: element (pixel m, channel c) of input segment s lives at
// seg_ptr[s][m*seg_stride[s] + c]; the concatenation of the segments is the conv input (replaces torch.cat,
// gru.py:20-21,30).  Weights are pre-packed [KH*KW*Cin][wstride] float32 (pack_conv_weight in pvo_b200/nn.py).
// epilogue: 0 bias, 1 bias+ReLU, 2 sigmoid(.)*net ("glo" gate), 3 GRU z|r, 4 GRU q + state update.
// stride: 1 or 2 (extractor.py's downsampling convs); output size = (H + 2*(K/2) - K)/stride + 1.
// epilogues 0/1: aux0 (optional) = residual tensor (stride aux_stride): out = relu(residual + act(conv)).
// aux pointers by epilogue:  2: aux0 = net (stride aux_stride)        3: aux0 = net, aux1 = glo[E][384], out = z, out2 = r*net
//                            4: aux0 = z, aux1 = glo[E][384], out = net (in place, stride ostride)
extern "C" int pvo_conv2d(const void* const* seg_ptr, const int* seg_channels, const int* seg_stride, int nseg,
                          const float* W, int wstride, const float* bias, int Cout, int KH, int KW, int pad, int nimg,
                          int H, int Wd, int dtype, int epilogue, void* out, int ostride, int ooff, void* out2,
                          const void* aux0, int aux_stride, const float* aux1, int stride, void* stream) {
  if (nseg < 1 || nseg > 4 || Cout < 1 || wstride % 4 || nimg < 0 || stride < 1) return PVO_ERR_INVALID;
  if (nimg == 0) return PVO_OK;
  ConvP p;
  p.nseg = nseg;
  p.Cin = 0;
  const size_t esz = dtype == PVO_DTYPE_F32 ? 4 : 2;
  bool vec = true;
  for (int s = 0; s < nseg; s++) {
    p.seg[s] = Seg{seg_ptr[s], seg_channels[s], seg_stride[s]};
    p.Cin += seg_channels[s];
    if (seg_channels[s] % 8) vec = false;
    if ((seg_stride[s] * esz) % 16 || ((uintptr_t)seg_ptr[s]) % 16) vec = false;
  }
  for (int s = nseg; s < 4; s++) p.seg[s] = Seg{nullptr, 0, 0};
  p.W = W; p.Wstride = wstride; p.Cout = Cout; p.KH = KH; p.KW = KW; p.padH = KH / 2; p.padW = KW / 2; (void)pad;   /* 'same' padding is the only mode the path uses */ p.H = H; p.Wd = Wd;
  p.stride = stride;
  p.Ho = (H + 2 * (KH / 2) - KH) / stride + 1;
  p.Wo = (Wd + 2 * (KW / 2) - KW) / stride + 1;
  p.M = (long)nimg * p.Ho * p.Wo;
  cudaStream_t s = (cudaStream_t)stream;
  const int HW = p.Ho * p.Wo;
#define DISPATCH(TI)                                                                                               \
  switch (epilogue) {                                                                                              \
    case 0: case 1: return launch_conv<TI>(p, EpiBias<TI>{bias, (TI*)out, ostride, ooff, Cout, epilogue, (const TI*)aux0, aux_stride}, vec, s); \
    case 2: return launch_conv<TI>(p, EpiGlo<TI>{bias, (const TI*)aux0, aux_stride, (TI*)out, Cout}, vec, s);      \
    case 3: return launch_conv<TI>(p, EpiZR<TI>{bias, aux1, (const TI*)aux0, aux_stride, (TI*)out, (TI*)out2, HW}, vec, s); \
    case 4: return launch_conv<TI>(p, EpiQ<TI>{bias, aux1, (const TI*)aux0, (TI*)out, ostride, HW}, vec, s);       \
    default: return PVO_ERR_UNSUPPORTED;                                                                           \
  }
  if (dtype == PVO_DTYPE_F32) { DISPATCH(float) }
  else if (dtype == PVO_DTYPE_F16) { DISPATCH(__half) }
#undef DISPATCH
  return PVO_ERR_UNSUPPORTED;
}

// Wfrag (optional, half mode, CG == 2): the weights pre-arranged as mma.sync B fragments [G][9][cin_g/16][32] x {half2,half2}
extern "C" int pvo_conv3x3_small(const void* in, int istride, int cin_g, const float* W, const float* bias, float* out,
                                 int ostride, int G, int CG, int nimg, int H, int Wd, int dtype, int post, const void* Wfrag,
                                 void* stream) {
  if (cin_g % 8 || (CG != 1 && CG != 2) || G < 1) return PVO_ERR_INVALID;
  if (nimg == 0) return PVO_OK;
  const size_t esz = dtype == PVO_DTYPE_F32 ? 4 : 2;
  if ((istride * esz) % 16 || cin_g % 64 || ((uintptr_t)in) % 16) return PVO_ERR_INVALID;
  const int tiles_x = (Wd + 15) / 16, tiles_y = (H + 7) / 8;
  const size_t patch = ((size_t)10 * 18 * (64 * esz + 16) + 15) & ~size_t(15);      // 64 channels staged per pass
  const size_t smem = patch + sizeof(float) * 9 * cin_g * CG;
  dim3 grid(nimg * tiles_x * tiles_y, G);
  cudaStream_t s = (cudaStream_t)stream;
#define L(TI, CGV)                                                                                                  \
  {                                                                                                                 \
    cudaFuncSetAttribute(conv3x3_small_kernel<TI, CGV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);    \
    conv3x3_small_kernel<TI, CGV><<<grid, 128, smem, s>>>((const TI*)in, istride, cin_g, W, bias, out, ostride, G, H, Wd, \
                                                          tiles_x, tiles_y, post);                                  \
  }
  if (dtype == PVO_DTYPE_F16 && CG == 2 && post == 0 && Wfrag) {
    const size_t sm2 = (((size_t)10 * 18 * (64 + 8) * 2 + 15) & ~size_t(15)) + sizeof(uint2) * 9 * (cin_g / 16) * 32;
    cudaFuncSetAttribute(conv3x3_heads_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2);
    conv3x3_heads_mma_kernel<<<grid, 128, sm2, s>>>((const __half*)in, istride, cin_g, (const uint2*)Wfrag, bias, out, ostride, H, Wd, tiles_x, tiles_y);
  }
  else if (dtype == PVO_DTYPE_F32) { if (CG == 1) L(float, 1) else L(float, 2) }
  else if (dtype == PVO_DTYPE_F16) { if (CG == 1) L(__half, 1) else L(__half, 2) }
  else return PVO_ERR_UNSUPPORTED;
#undef L
  return pvo_check_launch();
}

// scratch: E * 32 * C floats
extern "C" int pvo_colmean(const void* in, float* out, float* scratch, int E, int HW, int C, int dtype, void* stream) {
  if (E == 0) return PVO_OK;
  if (C % 2 || !scratch) return PVO_ERR_INVALID;
  cudaStream_t s = (cudaStream_t)stream;
  dim3 grid(E, kColChunks);
  if (dtype == PVO_DTYPE_F32) colsum_partial_kernel<float><<<grid, 64, 0, s>>>((const float*)in, scratch, HW, C);
  else colsum_partial_kernel<__half><<<grid, 64, 0, s>>>((const __half*)in, scratch, HW, C);
  colsum_final_kernel<<<E, 128, 0, s>>>(scratch, out, HW, C);
  return pvo_check_launch();
}

// in, out, res: (N, HW, C) pixel-major (contiguous); scratch: N*32*C*2 + N*C*2 floats.  out = [relu](IN(in)); with res:
// out = relu(res + [relu](IN(in))).  in == out is allowed.
extern "C" int pvo_instance_norm(const void* in, void* out, const void* res, float* scratch, int N, int HW, int C, int relu,
                                 int dtype, void* stream) {
  if (N == 0) return PVO_OK;
  if (!scratch) return PVO_ERR_INVALID;
  cudaStream_t s = (cudaStream_t)stream;
  float* part = scratch;
  float* stats = scratch + (size_t)N * kColChunks * C * 2;
  dim3 grid(N, kColChunks);
  const long total = (long)N * HW * C;
  if (dtype == PVO_DTYPE_F32) {
    in_stats_partial_kernel<float><<<grid, 64, 0, s>>>((const float*)in, part, HW, C);
    in_stats_final_kernel<<<N, 128, 0, s>>>(part, stats, HW, C);
    in_apply_kernel<float><<<ceil_div(total, 256), 256, 0, s>>>((const float*)in, stats, (const float*)res, (float*)out, HW, C, total, relu);
  } else {
    in_stats_partial_kernel<__half><<<grid, 64, 0, s>>>((const __half*)in, part, HW, C);
    in_stats_final_kernel<<<N, 128, 0, s>>>(part, stats, HW, C);
    in_apply_kernel<__half><<<ceil_div(total, 256), 256, 0, s>>>((const __half*)in, stats, (const __half*)res, (__half*)out, HW, C, total, relu);
  }
  return pvo_check_launch();
}

extern "C" int pvo_expand_x7(const void* motn, void* out, int nimg, int H, int Wd, void* stream) {
  if (nimg == 0) return PVO_OK;
  const long M = (long)nimg * H * Wd;
  expand_x7_kernel<<<ceil_div(M * 8, 256), 256, 0, (cudaStream_t)stream>>>((const __half*)motn, (__half*)out, Wd, M);
  return pvo_check_launch();
}

extern "C" int pvo_glo_matvec(const float* glo, const float* W, const float* b, float* out, int E, void* stream) {
  if (E == 0) return PVO_OK;
  glo_matvec_kernel<<<E, 384, 0, (cudaStream_t)stream>>>(glo, W, b, out);
  return pvo_check_launch();
}

extern "C" int pvo_glo_reduce_matvec(const float* partial, int parts_per_img, int HW, const float* W, const float* b, float* out,
                                     int E, void* stream) {
  if (E == 0) return PVO_OK;
  glo_reduce_matvec_kernel<<<E, 384, 0, (cudaStream_t)stream>>>(partial, parts_per_img, 1.0f / (float)HW, W, b, out);
  return pvo_check_launch();
}

extern "C" int pvo_segment_mean(const void* in, int istride, int ioff, void* out, const int* ptr, const int* idx,
                                const int* meta, int Kmax, int HW, int C, int dtype, void* stream) {
  if (Kmax == 0) return PVO_OK;
  const size_t esz = dtype == PVO_DTYPE_F32 ? 4 : 2;
  if ((C * esz) % 16 || (istride * esz) % 16 || (ioff * esz) % 16) return PVO_ERR_INVALID;
  dim3 grid(ceil_div((long)HW * C * esz / 16, 256), Kmax);
  if (dtype == PVO_DTYPE_F32)
    segment_mean_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)in, istride, ioff, (float*)out, ptr, idx, meta, HW, C);
  else
    segment_mean_kernel<__half><<<grid, 256, 0, (cudaStream_t)stream>>>((const __half*)in, istride, ioff, (__half*)out, ptr, idx, meta, HW, C);
  return pvo_check_launch();
}

// src (N,C,H,W) -> dst pixel-major with channel stride/offset; sdtype/ddtype independently f32/f16
extern "C" int pvo_nchw_to_nhwc(const void* src, int sdtype, void* dst, int ddtype, int N, int C, int HW, int dstride,
                                int doff, void* stream) {
  if (N == 0) return PVO_OK;
  dim3 grid(ceil_div(HW, 32), ceil_div(C, 32), N);
  cudaStream_t s = (cudaStream_t)stream;
  const long total = (long)N * C * HW;
  if (sdtype == 0 && ddtype == 0) nchw_to_nhwc_kernel<float, float><<<grid, 256, 0, s>>>((const float*)src, (float*)dst, C, HW, dstride, doff, total);
  else if (sdtype == 0 && ddtype == 1) nchw_to_nhwc_kernel<float, __half><<<grid, 256, 0, s>>>((const float*)src, (__half*)dst, C, HW, dstride, doff, total);
  else if (sdtype == 1 && ddtype == 0) nchw_to_nhwc_kernel<__half, float><<<grid, 256, 0, s>>>((const __half*)src, (float*)dst, C, HW, dstride, doff, total);
  else nchw_to_nhwc_kernel<__half, __half><<<grid, 256, 0, s>>>((const __half*)src, (__half*)dst, C, HW, dstride, doff, total);
  return pvo_check_launch();
}

extern "C" int pvo_nhwc_to_nchw(const void* src, int sdtype, int sstride, int soff, void* dst, int ddtype, int N, int C,
                                int HW, void* stream) {
  if (N == 0) return PVO_OK;
  dim3 grid(ceil_div(HW, 32), ceil_div(C, 32), N);
  cudaStream_t s = (cudaStream_t)stream;
  if (sdtype == 0 && ddtype == 0) nhwc_to_nchw_kernel<float, float><<<grid, 256, 0, s>>>((const float*)src, sstride, soff, (float*)dst, C, HW);
  else if (sdtype == 0 && ddtype == 1) nhwc_to_nchw_kernel<float, __half><<<grid, 256, 0, s>>>((const float*)src, sstride, soff, (__half*)dst, C, HW);
  else if (sdtype == 1 && ddtype == 0) nhwc_to_nchw_kernel<__half, float><<<grid, 256, 0, s>>>((const __half*)src, sstride, soff, (float*)dst, C, HW);
  else nhwc_to_nchw_kernel<__half, __half><<<grid, 256, 0, s>>>((const __half*)src, sstride, soff, (__half*)dst, C, HW);
  return pvo_check_launch();
}
