// Micro-benchmark: issue rate of tcgen05.mma kind::f16 (SS: both operands in shared memory, SWIZZLE_128B K-major tiles,
// fp32 accumulate in TMEM) for the instruction patterns the 3xFP16 kernels use.  One CTA per SM, no operand loads (the
// tiles hold zeros), one thread issues `groups` repetitions of a pattern and commits; cycles are clock64 deltas.
//   pattern 0: N=128, one accumulator           1: N=128, two accumulators alternating     2: N=128, four accumulators
//   pattern 3: N=256, one accumulator           4: N=256, two accumulators
//   pattern 5: 3xFP16 triplet as shipped  (dc <- al*bh, dc <- ah*bl, dm <- ah*bh), N=128
//   pattern 6: fused pair                 ([dm|dc] <- ah*[bh;bl] as ONE N=256 MMA, then dc <- al*bh N=128)
//   pattern 7: triplet reordered          (dc <- al*bh, dm <- ah*bh, dc <- ah*bl)
//   pattern 8: N=240 triplet (k_mp_gru)   9: N=64 triplet      10: N=96 triplet
//   pattern 11: triplet N=128 with distinct smem tiles per operand role but A advanced along K (4 x K16 per 64-wide k-block)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t sdesc(uint32_t addr) {
  return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ uint32_t idesc(int M, int N) { return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
               "l"(a), "l"(b), "r"(id), "r"(acc)
               : "memory");
}

__global__ void __launch_bounds__(128, 1) k_rate(int pattern, int groups, long long *cycles, long long *nanos) {
  extern __shared__ uint8_t raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  // A_hi 16 KB | A_lo 16 KB | B_hi 32 KB | B_lo 32 KB (B_lo directly after B_hi when N = 128: offset 16 KB)
  for (int i = threadIdx.x; i < (96 * 1024) / 16; i += blockDim.x) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
  uint64_t *bar = reinterpret_cast<uint64_t *>(smem + 96 * 1024);
  uint32_t *slot = reinterpret_cast<uint32_t *>(bar + 1);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = *slot;
  if (threadIdx.x == 0) {
    const uint32_t base = smem_u32(smem);
    const uint64_t ah = sdesc(base), al = sdesc(base + 16384), bh = sdesc(base + 32768);
    long long t0 = clock64(), n0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(n0));
    for (int g = 0; g < groups; ++g) {
      const uint64_t o = (uint64_t)((g & 3) * 2);
      switch (pattern) {
        case 0: { const uint32_t id = idesc(128, 128); mma(tm, ah + o, bh + o, id, 1); mma(tm, ah + o, bh + o, id, 1); mma(tm, ah + o, bh + o, id, 1); break; }
        case 1: { const uint32_t id = idesc(128, 128); mma(tm, ah + o, bh + o, id, 1); mma(tm + 128, ah + o, bh + o, id, 1); mma(tm, ah + o, bh + o, id, 1); mma(tm + 128, ah + o, bh + o, id, 1); break; }
        case 2: { const uint32_t id = idesc(128, 128); for (int q = 0; q < 4; ++q) mma(tm + 128 * q, ah + o, bh + o, id, 1); break; }
        case 3: { const uint32_t id = idesc(128, 256); mma(tm, ah + o, bh + o, id, 1); mma(tm, ah + o, bh + o, id, 1); break; }
        case 4: { const uint32_t id = idesc(128, 256); mma(tm, ah + o, bh + o, id, 1); mma(tm + 256, ah + o, bh + o, id, 1); break; }
        case 5: { const uint32_t id = idesc(128, 128); const uint64_t bl = sdesc(base + 32768 + 16384);
                  mma(tm + 128, al + o, bh + o, id, 1); mma(tm + 128, ah + o, bl + o, id, 1); mma(tm, ah + o, bh + o, id, 1); break; }
        case 6: { const uint32_t id2 = idesc(128, 256), id1 = idesc(128, 128);
                  mma(tm, ah + o, bh + o, id2, 1); mma(tm + 128, al + o, bh + o, id1, 1); break; }
        case 7: { const uint32_t id = idesc(128, 128); const uint64_t bl = sdesc(base + 32768 + 16384);
                  mma(tm + 128, al + o, bh + o, id, 1); mma(tm, ah + o, bh + o, id, 1); mma(tm + 128, ah + o, bl + o, id, 1); break; }
        case 8: { const uint32_t id = idesc(128, 240); const uint64_t bl = sdesc(base + 32768 + 30720);
                  mma(tm + 256, al + o, bh + o, id, 1); mma(tm + 256, ah + o, bl + o, id, 1); mma(tm, ah + o, bh + o, id, 1); break; }
        case 9: { const uint32_t id = idesc(128, 64); const uint64_t bl = sdesc(base + 32768 + 8192);
                  mma(tm + 64, al + o, bh + o, id, 1); mma(tm + 64, ah + o, bl + o, id, 1); mma(tm, ah + o, bh + o, id, 1); break; }
        case 10: { const uint32_t id = idesc(128, 96); const uint64_t bl = sdesc(base + 32768 + 12288);
                  mma(tm + 96, al + o, bh + o, id, 1); mma(tm + 96, ah + o, bl + o, id, 1); mma(tm, ah + o, bh + o, id, 1); break; }
        default: break;
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    uint32_t done = 0;
    while (!done)
      asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}"
                   : "=r"(done) : "r"(smem_u32(bar)), "r"(0) : "memory");
    long long t1 = clock64(), n1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(n1));
    cycles[blockIdx.x] = t1 - t0; nanos[blockIdx.x] = n1 - n0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
}

int main(int argc, char **argv) {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int smem = 96 * 1024 + 1024 + 64;
  cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  long long *dc, *dn;
  cudaMalloc(&dc, sms * 8); cudaMalloc(&dn, sms * 8);
  long long hc[256], hn[256];
  const char *names[] = {"N128 x3, one accumulator", "N128 x4, two accumulators", "N128 x4, four accumulators", "N256 x2, one accumulator",
                         "N256 x2, two accumulators", "3xFP16 triplet N128 (shipped)", "fused N256 + N128", "triplet reordered N128",
                         "triplet N240", "triplet N64", "triplet N96"};
  const int per_group[] = {3, 4, 4, 2, 2, 3, 2, 3, 3, 3, 3};
  const double mac_per_group[] = {3 * 128. * 128 * 16, 4 * 128. * 128 * 16, 4 * 128. * 128 * 16, 2 * 128. * 256 * 16, 2 * 128. * 256 * 16,
                                  3 * 128. * 128 * 16, 3 * 128. * 128 * 16, 3 * 128. * 128 * 16, 3 * 128. * 240 * 16, 3 * 128. * 64 * 16, 3 * 128. * 96 * 16};
  for (int grid : {1, sms}) {
    printf("--- %d CTA(s)\n", grid);
    for (int pat = 0; pat <= 10; ++pat) {
      const int groups = 20000;
      k_rate<<<grid, 128, smem>>>(pat, 2000, dc, dn);     // warm-up
      k_rate<<<grid, 128, smem>>>(pat, groups, dc, dn);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("pattern %d: %s\n", pat, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(hc, dc, grid * 8, cudaMemcpyDeviceToHost); cudaMemcpy(hn, dn, grid * 8, cudaMemcpyDeviceToHost);
      double c = 0, n = 0;
      for (int i = 0; i < grid; ++i) { c += hc[i]; n += hn[i]; }
      c /= grid; n /= grid;
      const double cyc_per_mma = c / ((double)groups * per_group[pat]);
      const double tf = 2.0 * mac_per_group[pat] * groups * grid / (n * 1e-9) / 1e12;
      printf("pattern %2d %-32s: %7.1f cycles/MMA, %6.1f cycles/group, clock %.2f GHz, %7.1f TFLOP/s (MMA passes), %.0f%% of 8192 flop/clk/SM\n",
             pat, names[pat], cyc_per_mma, c / groups, c / n, tf, 100.0 * 2.0 * mac_per_group[pat] / (c / groups) / 8192.0);
    }
  }
  return 0;
}
