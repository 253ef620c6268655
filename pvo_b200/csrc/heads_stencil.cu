// Second convolution of the four update-operator heads (delta, delta_dy, weight, delta_mask: 3x3, 128 -> 2 each,
// droid_net.py:195-214) on the tensor-core path.  A 3x3 conv with 2 outputs is re-associated as
//   out[p][o] = bias[o] + sum_tap T[p + tap][tap][o],     T[q][tap][o] = sum_c W[o][c][tap] * Hb[q][c]
// i.e. one 1x1 convolution 128 -> 18 per head (all four heads in one block-diagonal 512 -> 4x32 tcgen05 launch,
// pvo_conv2d_tc) followed by this 9-tap gather-sum.  The direct 3x3 form reads its 512-channel input through a halo tile
// per block (0.12 ms at config 2); this form reads it once at tensor-core speed and the stencil touches 256 B per pixel.
#include "common.cuh"

namespace {

// one thread per (pixel, head): T is [M][128] half with head g's taps at columns g*32 + tap*2 + {0,1}
__global__ void __launch_bounds__(256) heads_stencil_kernel(const __half* __restrict__ T, const float* __restrict__ bias, float* __restrict__ out,
                                                            int ostride, long M, int H, int W) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * 4) return;
  const int g = (int)(i & 3);
  const long m = i >> 2;
  const int x = (int)(m % W), y = (int)((m / W) % H);
  float a0 = bias[2 * g], a1 = bias[2 * g + 1];
#pragma unroll
  for (int tap = 0; tap < 9; tap++) {
    const int dy = tap / 3 - 1, dx = tap % 3 - 1;
    if (y + dy < 0 || y + dy >= H || x + dx < 0 || x + dx >= W) continue;
    const __half2 v = __ldg(reinterpret_cast<const __half2*>(T + (m + (long)dy * W + dx) * 128 + g * 32 + tap * 2));
    const float2 f = __half22float2(v);
    a0 += f.x;
    a1 += f.y;
  }
  out[m * ostride + 2 * g] = a0;
  out[m * ostride + 2 * g + 1] = a1;
}

}  // namespace

// T: [nimg*H*W][128] half (pvo_conv2d_tc output of the block-diagonal 1x1 conv); bias: [8]; out: [nimg*H*W][ostride] float32,
// columns 0..7 = the four heads' two outputs (same layout as pvo_conv3x3_small with G = 4, CG = 2).
extern "C" int pvo_heads_stencil(const void* T, const float* bias, float* out, int ostride, int nimg, int H, int W, void* stream) {
  const long M = (long)nimg * H * W;
  if (M == 0) return PVO_OK;
  heads_stencil_kernel<<<ceil_div(M * 4, 256), 256, 0, (cudaStream_t)stream>>>((const __half*)T, bias, out, ostride, M, H, W);
  return pvo_check_launch();
}
