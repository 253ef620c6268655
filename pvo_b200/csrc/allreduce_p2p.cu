// One-shot all-reduce (sum, fp64) of the BA reduced camera system over NVLink / NVSwitch peer memory.
//
// The sharded frame graph (pvo_b200/dist.py) needs, per Gauss-Newton iteration, the sum over ranks
// of a (36 P^2 + 6 P)-element system: 15 KB at P = 7, 1.1 MB at P = 63.  For messages this small a ring / tree collective
// is all latency, so every rank PUSHES its partial straight into a staging slot in every peer's memory (symmetric
// allocation, plain st.global over NVLink), raises a flag there, and then sums the slots that arrived in its own memory:
//   block b of rank r:  for each peer p: stage_p[parity][r][chunk b] = partial[chunk b];  fence;  flag_p[r][b] = epoch
//                       wait until flag_local[q][b] == epoch for every rank q;  partial[chunk b] = sum_q stage_local[parity][q][chunk b]
// One launch, no host involvement, capturable in a CUDA graph (the epoch lives in device memory and advances by itself).
// The epoch is ONE counter per rank, advanced once per launch by the last block to finish (not one per block: the block count
// follows the message size, so per-block counters would drift apart when P changes between calls and with them the staging
// parity).  Two staging parities suffice: a rank can be at most one all-reduce ahead of a peer, because finishing call c+1 needs
// every block of that peer to have pushed for c+1, which the peer does only after its call c has completed (stream order).
#include "common.cuh"

namespace {

constexpr int kMaxRanks = 8;

struct Peers {
  double* stage[kMaxRanks];   // each: [2][world][n_max] doubles, in rank i's memory
  int* flags[kMaxRanks];      // each: [world][kMaxBlocks] ints, in rank i's memory
};
constexpr int kMaxBlocks = 256;

__device__ __forceinline__ void st_release_sys(int* p, int v) { asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ int ld_acquire_sys(const int* p) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(256) allreduce_oneshot_kernel(double* __restrict__ data, const __grid_constant__ Peers peers, int* __restrict__ epochs,
                                                                int rank, int world, long n, long n_max) {
  __shared__ int s_epoch;
  const int b = blockIdx.x;
  if (threadIdx.x == 0) s_epoch = *reinterpret_cast<volatile int*>(epochs) + 1;      // epochs[0]: calls completed, epochs[1]: blocks done
  __syncthreads();
  const int epoch = s_epoch;
  const int parity = epoch & 1;
  const long per = (n + gridDim.x - 1) / gridDim.x;
  const long i0 = (long)b * per, i1 = min(n, i0 + per);
  // push my chunk into every rank's staging slot for me (my own included: one code path, one summation order everywhere)
  for (int p = 0; p < world; p++) {
    double* dst = peers.stage[p] + ((long)parity * world + rank) * n_max;
    for (long i = i0 + threadIdx.x; i < i1; i += blockDim.x) dst[i] = data[i];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < world) st_release_sys(peers.flags[threadIdx.x] + rank * kMaxBlocks + b, epoch);
  // wait for every rank's chunk b to land here
  if (threadIdx.x < world) {
    const int* f = peers.flags[rank] + threadIdx.x * kMaxBlocks + b;
    long spins = 0;
    while (ld_acquire_sys(f) < epoch) {
      if (++spins > (1L << 28)) __trap();            // a lost peer must not hang the GPU
    }
  }
  __syncthreads();
  const double* mine = peers.stage[rank] + (long)parity * world * n_max;
  for (long i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
    double s = 0.0;
    for (int q = 0; q < world; q++) s += __ldcg(mine + (long)q * n_max + i);      // rank order: bitwise identical on every rank
    data[i] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(epochs + 1, 1) == (int)gridDim.x - 1) {      // last block of this launch: every block has read epochs[0]
      epochs[1] = 0;
      __threadfence();
      *reinterpret_cast<volatile int*>(epochs) = epoch;
    }
  }
}

}  // namespace

// data: n doubles, summed in place across `world` ranks.  stage_ptrs[i] / flag_ptrs[i]: rank i's symmetric staging buffer
// ([2][world][n_max] doubles) and flag array ([world][256] ints, zero-initialised once).  epochs: local [256] ints, zeroed once.
extern "C" int pvo_allreduce_oneshot(double* data, void* const* stage_ptrs, void* const* flag_ptrs, int* epochs, int rank, int world, long n,
                                     long n_max, void* stream) {
  if (world < 1 || world > kMaxRanks || n > n_max || rank < 0 || rank >= world) return PVO_ERR_INVALID;
  if (n == 0 || world == 1) return PVO_OK;
  Peers p;
  for (int i = 0; i < kMaxRanks; i++) {
    p.stage[i] = i < world ? (double*)stage_ptrs[i] : nullptr;
    p.flags[i] = i < world ? (int*)flag_ptrs[i] : nullptr;
  }
  int blocks = (int)((n + 1023) / 1024);
  blocks = blocks < 1 ? 1 : (blocks > 128 ? 128 : blocks);       // every block must be co-resident with its peers' blocks
  allreduce_oneshot_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(data, p, epochs, rank, world, n, n_max);
  return pvo_check_launch();
}
