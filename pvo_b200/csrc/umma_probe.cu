// Hardware probe (test infrastructure, not on the product path): how does tcgen05.mma address a SWIZZLE_128B K-major A operand
// whose descriptor start address is NOT 1024-byte aligned and whose 8-row groups are not 1024 B apart?  The answer decides
// whether a 3x3 convolution can read its nine taps out of ONE haloed input patch in shared memory.
//   G   [rows_g][64] half, loaded by one 2-D TMA box (SW128) to a 1024-aligned patch -> patch row r at +128*r, swizzled by r & 7
//   A   = 128 rows: group g (8 rows) starts at patch + 128*row_off + g*sbo_bytes   (descriptor: start, SBO, base_offset)
//   B   [64][64] half (normal layout),  D[128][64] = A * B^T read back as float
#include "tc_common.cuh"

namespace {
using namespace tc;

__global__ void __launch_bounds__(128) umma_probe_kernel(const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmB,
                                                         int rows_g, int row_off, int sbo_bytes, int base_offset, float* __restrict__ D) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* patch = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* bt = patch + 64 * 1024;                       // B tile, 1024-aligned
  uint64_t* bar = reinterpret_cast<uint64_t*>(bt + 8192);
  uint64_t* done = bar + 1;
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(done, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(slot, 64);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar, rows_g * 128 + 8192);
    for (int r0 = 0; r0 < rows_g; r0 += 256) tma_load_2d(&tmG, bar, patch + r0 * 128, 0, r0);
    tma_load_2d(&tmB, bar, bt, 0, 0);
    mbar_wait(bar, 0);
    fence_after_sync();
    uint64_t da = 0;
    const uint32_t sa = smem_u32(patch) + row_off * 128;
    da |= (uint64_t)((sa & 0x3FFFF) >> 4);
    da |= (uint64_t)1 << 16;
    da |= (uint64_t)(sbo_bytes >> 4) << 32;
    da |= (uint64_t)1 << 46;
    da |= (uint64_t)(base_offset & 7) << 49;
    da |= (uint64_t)2 << 61;
    const uint64_t db = smem_desc_sw128(smem_u32(bt));
    constexpr uint32_t idesc = idesc_f16(128, 64);
    for (int k = 0; k < 4; k++) umma_f16(tmem, da + 2 * k, db + 2 * k, idesc, k != 0);
    umma_commit(done);
  }
  __syncthreads();
  if (warp < 4) {
    if (lane == 0) mbar_wait(done, 0);
    __syncwarp();
    fence_after_sync();
    for (int c = 0; c < 64; c += 16) {
      uint32_t v[16];
      tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c, v);
      tmem_ld_wait();
      for (int j = 0; j < 16; j++) D[(warp * 32 + lane) * 64 + c + j] = __uint_as_float(v[j]);
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 64);
}
}  // namespace

// G: device [rows_g][64] half (rows_g <= 512, multiple of 8), B: device [64][64] half, D: device [128][64] float
extern "C" int pvo_umma_probe(const void* G, const void* B, int rows_g, int row_off, int sbo_bytes, int base_offset, float* D, void* stream) {
  if (rows_g > 512 || rows_g % 8) return PVO_ERR_INVALID;
  CUtensorMap tmG, tmB;
  {
    const uint64_t dims[2] = {64, (uint64_t)rows_g};
    const uint64_t strides[1] = {128};
    const uint32_t box[2] = {64, (uint32_t)(rows_g < 256 ? rows_g : 256)};
    if (make_tmap_f16(&tmG, G, 2, dims, strides, box) != PVO_OK) return PVO_ERR_INVALID;
  }
  {
    const uint64_t dims[2] = {64, 64};
    const uint64_t strides[1] = {128};
    const uint32_t box[2] = {64, 64};
    if (make_tmap_f16(&tmB, B, 2, dims, strides, box) != PVO_OK) return PVO_ERR_INVALID;
  }
  const int smem = 64 * 1024 + 8192 + 64 + 1024;
  cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  umma_probe_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(tmG, tmB, rows_g, row_off, sbo_bytes, base_offset, D);
  return pvo_check_launch();
}
