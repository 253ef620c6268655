// Latency micro-benchmarks that bound the single-thread parts of the BA solve (ba.cu): dependent DFMA / DMUL chains, the reciprocal
// sequence, shared-memory 64-bit load-to-use, and a 256-thread named barrier.  One warp, lane 0 times with clock64().
//   nvcc -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC tools/ubench_fp64.cu -o tools/_prof/libubench.so
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double seed, long long* out, double* sink) {
  __shared__ double sm[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = seed + i;
  __syncthreads();
  long long t0, t1;
  double a = seed, b = seed * 0.5, c = 1.0 + seed;
  if (threadIdx.x == 0) {
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; i++) a = fma(a, b, c);
    t1 = clock64();
    out[0] = t1 - t0;                       // 64 dependent DFMA
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; i++) a = a * b;
    t1 = clock64();
    out[1] = t1 - t0;                       // 64 dependent DMUL
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 16; i++) {
      double r;
      asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
      r = fma(fma(-a, r, 1.0), r, r);
      r = fma(fma(-a, r, 1.0), r, r);
      a = r + c;
    }
    t1 = clock64();
    out[2] = t1 - t0;                       // 16 x (rcp + 2 Newton + add)
    int idx = ((int)seed) & 1023;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 32; i++) idx = ((int)sm[idx]) & 1023;
    t1 = clock64();
    out[3] = t1 - t0;                       // 32 dependent LDS.64 + F2I + AND
    float fa = (float)seed, fb = 0.5f;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; i++) fa = fmaf(fa, fb, 1.0f);
    t1 = clock64();
    out[4] = t1 - t0;                       // 64 dependent FFMA
    a += fa + idx;
    // two independent DFMA chains interleaved
    double a2 = seed * 3.0;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; i++) { a = fma(a, b, c); a2 = fma(a2, b, c); }
    t1 = clock64();
    out[5] = t1 - t0;                       // 2 x 64 DFMA, two chains
    a += a2;
    double a3 = seed * 5.0, a4 = seed * 7.0;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; i++) { a = fma(a, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c); a4 = fma(a4, b, c); }
    t1 = clock64();
    out[7] = t1 - t0;                       // 4 x 64 DFMA, four chains
    a += a2 + a3 + a4;
  }
  __syncthreads();
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 32; i++) asm volatile("bar.sync 1, 256;" ::: "memory");
  t1 = clock64();
  if (threadIdx.x == 0) { out[6] = t1 - t0; sink[0] = a; }     // 32 barriers, 256 threads
}
extern "C" int ubench_run() {
  long long* d; double* s;
  cudaMalloc(&d, 64); cudaMalloc(&s, 8);
  long long h[8];
  for (int rep = 0; rep < 2; rep++) { k<<<1, 256>>>(1.25, d, s); cudaDeviceSynchronize(); }
  cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
  printf("dependent DFMA        %.1f cycles each\n", h[0] / 64.0);
  printf("dependent DMUL        %.1f cycles each\n", h[1] / 64.0);
  printf("rcp + 2 Newton + add  %.1f cycles each\n", h[2] / 16.0);
  printf("LDS.64 -> F2I -> AND  %.1f cycles each\n", h[3] / 32.0);
  printf("dependent FFMA        %.1f cycles each\n", h[4] / 64.0);
  printf("2 DFMA chains         %.1f cycles per pair\n", h[5] / 64.0);
  printf("4 DFMA chains         %.1f cycles per quad\n", h[7] / 64.0);
  printf("bar.sync 256 threads  %.1f cycles each\n", h[6] / 32.0);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}
