/* CPU restatement of the reference's correlation-lookup kernels -- ORACLE, test infrastructure only.
 *
 * Restates, operation for operation (same loop nest, same read-modify-write order, same rounding):
 *   VO_Module/src/correlation_kernels.cu:19-70   corr_index_forward_kernel  (float and c10::Half)
 *   VO_Module/src/correlation_kernels.cu:126-155 host wrapper (zero-initialised output)
 * and, to floating-point tolerance (summation order over channels differs):
 *   VO_Module/src/altcorr_kernel.cu:27-149       altcorr_forward_kernel
 *
 * Rounding model (SURVEY.md section 8(c) "K1 bit-exactness recipe"):
 *   float : nvcc's default -fmad=true contracts `corr += s * w` into one fused multiply-add -> fmaf().
 *   half  : c10::Half operators up-cast to float, compute, and round back to half after EVERY operation
 *           (torch/headeronly/util/Half.h): w_h = half(w); p = half(float(s)*float(w_h));
 *           acc = half(float(acc)+float(p)).
 * The weight products (dx*dy etc.) are plain float multiplies of float subtractions (no contraction
 * possible).  Built with -ffp-contract=off so the compiler adds no fusion of its own.
 *
 * Parity status: PINNED only on the GPU box, against the reference's own kernel compiled into
 * oracle/_ref/ (tests/test_corr_gpu.py::test_reference_kernel_matches_oracle). */
#include <math.h>
#include <stdint.h>
#include <string.h>

typedef _Float16 half_t;

static inline int inb(int h, int w, int H, int W) { return h >= 0 && h < H && w >= 0 && w < W; }

/* volume [N][h1][w1][h2][w2], coords [N][2][h1][w1] (x then y), out [N][rd][rd][h1][w1] */
void corr_index_forward_f32(const float* vol, const float* coords, float* out,
                            int N, int h1, int w1, int h2, int w2, int r)
{
  const int rd = 2 * r + 1;
  const long plane = (long)h1 * w1;
  memset(out, 0, sizeof(float) * (size_t)N * rd * rd * plane);
#pragma omp parallel for collapse(2) schedule(static)
  for (int n = 0; n < N; n++)
    for (int y = 0; y < h1; y++)
      for (int x = 0; x < w1; x++) {
        const float x0 = coords[((long)n * 2 + 0) * plane + (long)y * w1 + x];
        const float y0 = coords[((long)n * 2 + 1) * plane + (long)y * w1 + x];
        const float dx = x0 - floorf(x0);
        const float dy = y0 - floorf(y0);
        const float* v = vol + (((long)n * h1 + y) * w1 + x) * (long)h2 * w2;
        float* o = out + (long)n * rd * rd * plane + (long)y * w1 + x;
#define O(i, j) o[((long)(i) * rd + (j)) * plane]
        for (int i = 0; i < rd + 1; i++)
          for (int j = 0; j < rd + 1; j++) {
            const int x1 = (int)floorf(x0) - r + i;
            const int y1 = (int)floorf(y0) - r + j;
            if (!inb(y1, x1, h2, w2)) continue;
            const float s = v[(long)y1 * w2 + x1];
            if (i > 0 && j > 0)   O(i - 1, j - 1) = fmaf(s, dx * dy, O(i - 1, j - 1));
            if (i > 0 && j < rd)  O(i - 1, j)     = fmaf(s, dx * (1.0f - dy), O(i - 1, j));
            if (i < rd && j > 0)  O(i, j - 1)     = fmaf(s, (1.0f - dx) * dy, O(i, j - 1));
            if (i < rd && j < rd) O(i, j)         = fmaf(s, (1.0f - dx) * (1.0f - dy), O(i, j));
          }
#undef O
      }
}

static inline half_t hmul(half_t a, half_t b) { return (half_t)((float)a * (float)b); }
static inline half_t hadd(half_t a, half_t b) { return (half_t)((float)a + (float)b); }

void corr_index_forward_f16(const uint16_t* vol_, const float* coords, uint16_t* out_,
                            int N, int h1, int w1, int h2, int w2, int r)
{
  const half_t* vol = (const half_t*)vol_;
  half_t* out = (half_t*)out_;
  const int rd = 2 * r + 1;
  const long plane = (long)h1 * w1;
  memset(out, 0, sizeof(half_t) * (size_t)N * rd * rd * plane);
#pragma omp parallel for collapse(2) schedule(static)
  for (int n = 0; n < N; n++)
    for (int y = 0; y < h1; y++)
      for (int x = 0; x < w1; x++) {
        const float x0 = coords[((long)n * 2 + 0) * plane + (long)y * w1 + x];
        const float y0 = coords[((long)n * 2 + 1) * plane + (long)y * w1 + x];
        const float dx = x0 - floorf(x0);
        const float dy = y0 - floorf(y0);
        const half_t* v = vol + (((long)n * h1 + y) * w1 + x) * (long)h2 * w2;
        half_t* o = out + (long)n * rd * rd * plane + (long)y * w1 + x;
#define O(i, j) o[((long)(i) * rd + (j)) * plane]
        for (int i = 0; i < rd + 1; i++)
          for (int j = 0; j < rd + 1; j++) {
            const int x1 = (int)floorf(x0) - r + i;
            const int y1 = (int)floorf(y0) - r + j;
            if (!inb(y1, x1, h2, w2)) continue;
            const half_t s = v[(long)y1 * w2 + x1];
            if (i > 0 && j > 0)   O(i - 1, j - 1) = hadd(O(i - 1, j - 1), hmul(s, (half_t)(dx * dy)));
            if (i > 0 && j < rd)  O(i - 1, j)     = hadd(O(i - 1, j), hmul(s, (half_t)(dx * (1.0f - dy))));
            if (i < rd && j > 0)  O(i, j - 1)     = hadd(O(i, j - 1), hmul(s, (half_t)((1.0f - dx) * dy)));
            if (i < rd && j < rd) O(i, j)         = hadd(O(i, j), hmul(s, (half_t)((1.0f - dx) * (1.0f - dy))));
          }
#undef O
      }
}

/* fmap1 [B][H1][W1][C], fmap2 [B][H2][W2][C], coords [B][S][H1][W1][2], out [B][S][rd*rd][H1][W1];
 * channel = iy + rd*ix (altcorr_kernel.cu:102-105).  Dot product accumulated in 32-channel slabs as the
 * reference does, each slab splatted separately. */
void altcorr_forward_f32(const float* f1, const float* f2, const float* coords, float* out,
                         int B, int S, int H1, int W1, int H2, int W2, int C, int r)
{
  const int rd = 2 * r + 1;
  const long plane = (long)H1 * W1;
  memset(out, 0, sizeof(float) * (size_t)B * S * rd * rd * plane);
#pragma omp parallel for collapse(2) schedule(static)
  for (int b = 0; b < B; b++)
    for (int h = 0; h < H1; h++)
      for (int w = 0; w < W1; w++)
        for (int c0 = 0; c0 < C; c0 += 32)
          for (int n = 0; n < S; n++) {
            const float* cp = coords + ((((long)b * S + n) * H1 + h) * W1 + w) * 2;
            const float x2 = cp[0], y2 = cp[1];
            const float dx = x2 - floorf(x2), dy = y2 - floorf(y2);
            const float* a = f1 + (((long)b * H1 + h) * W1 + w) * C + c0;
            float* o = out + ((long)b * S + n) * rd * rd * plane + (long)h * W1 + w;
            for (int iy = 0; iy < rd + 1; iy++)
              for (int ix = 0; ix < rd + 1; ix++) {
                const int hh = (int)floorf(y2) - r + iy, ww = (int)floorf(x2) - r + ix;
                float s = 0.0f;
                if (inb(hh, ww, H2, W2)) {
                  const float* bb = f2 + (((long)b * H2 + hh) * W2 + ww) * C + c0;
                  const int cn = (C - c0 < 32) ? C - c0 : 32;
                  for (int k = 0; k < cn; k++) s = fmaf(a[k], bb[k], s);
                }
                if (iy > 0 && ix > 0)   o[((iy - 1) + rd * (ix - 1)) * plane] += s * (dy * dx);
                if (iy > 0 && ix < rd)  o[((iy - 1) + rd * ix) * plane]       += s * (dy * (1 - dx));
                if (iy < rd && ix > 0)  o[(iy + rd * (ix - 1)) * plane]       += s * ((1 - dy) * dx);
                if (iy < rd && ix < rd) o[(iy + rd * ix) * plane]             += s * ((1 - dy) * (1 - dx));
              }
          }
}
