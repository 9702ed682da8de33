// Micro-benchmark: what slows a tcgen05.mma stream inside a real pipeline?  One CTA per SM, warp 1 lane 0 issues k-blocks
// of 4 x (N=256 + N=128) kind::f16 SS MMAs (the fused 3xFP16 pattern, 128-wide tile) while optional "disturbances" run:
//   bit 0: tcgen05.commit to an mbarrier after every k-block (nobody waits)
//   bit 1: tcgen05.fence::after_thread_sync before every k-block
//   bit 2: warps 2..9 spin on mbarrier.try_wait (a barrier that only completes at the end)
//   bit 3: warps 2..9 stream tcgen05.ld (32x32b.x16) + wait::ld from TMEM columns the MMAs do not touch
//   bit 4: the issuing thread does one mbarrier.try_wait on an already completed barrier per k-block
//   bit 5: warps 2..9 stream 16-byte shared-memory stores into an unrelated 64 KB region (stand-in for TMA writes)
//   bit 6: warps 2..9 stream tcgen05.ld from the SAME columns the MMAs accumulate into
//   bit 7: operands rotate through a ring of three A stages and three B stages (fresh shared-memory addresses every k-block)
//   bit 8: the accumulators ping-pong between two TMEM buffer pairs every 4 k-blocks, first MMA of a chunk overwrites (acc = 0)
//   bit 9: shared memory holds random fp16 data instead of zeros
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
__device__ __forceinline__ uint32_t try_wait(uint64_t *bar, uint32_t parity) {
  uint32_t done;
  asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}"
               : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return done;
}

__global__ void __launch_bounds__(320, 1) k_mix(int mode, int kblocks, long long *cycles) {
  extern __shared__ uint8_t raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  for (int i = threadIdx.x; i < (192 * 1024) / 16; i += blockDim.x) {
    uint32_t h = (uint32_t)i * 2654435761u;
    // random fp16 pairs with small exponents (|x| < 2): sign/exponent bits masked
    const uint32_t v = (mode & 512) ? ((h & 0x83FF83FFu) | 0x38003800u) : 0u;
    reinterpret_cast<uint4 *>(smem)[i] = make_uint4(v, v ^ 0x00110011u, v ^ 0x01010101u, v);
  }
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + 192 * 1024);   // [0] end, [1] done-at-start, [2..5] commit ring
  uint32_t *slot = reinterpret_cast<uint32_t *>(bars + 8);
  volatile int *stop = reinterpret_cast<volatile int *>(bars + 9);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 6; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bars + i)));
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bars + 1)) : "memory");
    *stop = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = *slot;
  if (warp == 1) {
    if (lane == 0) {
      const uint32_t base = smem_u32(smem);
      const uint32_t id2 = idesc(128, 256), id1 = idesc(128, 128);
      const long long t0 = clock64();
      for (int kb = 0; kb < kblocks; ++kb) {
        if (mode & 16) { while (!try_wait(bars + 1, 0)) {} }
        if (mode & 2) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t st = (mode & 128) ? (uint32_t)(kb % 3) : 0u;
        const uint64_t ah = sdesc(base + st * 32768), al = sdesc(base + st * 32768 + 16384), bh = sdesc(base + 98304 + st * 32768);
        const uint32_t d = tm + ((mode & 256) ? (uint32_t)(((kb >> 2) & 1) * 256) : 0u);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const uint64_t o = (uint64_t)(kk * 2);
          const uint32_t acc = ((mode & 256) && (kb & 3) == 0 && kk == 0) ? 0u : 1u;
          mma(d, ah + o, bh + o, id2, acc);
          mma(d + 128, al + o, bh + o, id1, 1);
        }
        if (mode & 1) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bars + 2 + (kb & 3))) : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bars)) : "memory");
      while (!try_wait(bars, 0)) {}
      cycles[blockIdx.x] = clock64() - t0;
      *stop = 1;
    }
  } else if (warp >= 2) {
    const uint32_t taddr = tm + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t sink = 0;
    if (mode & 4) {
      while (!try_wait(bars, 0)) {}
    } else if (mode & (8 | 64)) {
      const uint32_t col = (mode & 64) ? 0u : 256u;
      while (!*stop) {
        uint32_t r[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                       "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                     : "r"(taddr + col + (uint32_t)(((warp - 2) >> 2) * 64)));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        sink += r[0] + r[15];
      }
    } else if (mode & 32) {
      uint8_t *dst = smem + 128 * 1024;
      uint32_t i = threadIdx.x;
      while (!*stop) {
        *reinterpret_cast<uint4 *>(dst + ((i * 16) & 0xFFFF)) = make_uint4(i, i, i, i);
        i += 256;
      }
    }
    if (sink == 0x12345678) cycles[0] = sink;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
}

int main() {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int smem = 192 * 1024 + 1024 + 256;
  cudaFuncSetAttribute(k_mix, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  long long *dc;
  cudaMalloc(&dc, sms * 8);
  long long hc[256];
  const int modes[] = {0, 128, 256, 512, 128 | 512, 128 | 256 | 512, 128 | 256 | 512 | 1 | 2 | 16, 128 | 256 | 512 | 1 | 2 | 16 | 64, 128 | 256 | 512 | 1 | 2 | 16 | 32, 128 | 256 | 512 | 1 | 2 | 16 | 32 | 64};
  for (int mode : modes) {
    const int kblocks = 4000;
    k_mix<<<sms, 320, smem>>>(mode, 400, dc);
    k_mix<<<sms, 320, smem>>>(mode, kblocks, dc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(hc, dc, sms * 8, cudaMemcpyDeviceToHost);
    double c = 0;
    for (int i = 0; i < sms; ++i) c += hc[i];
    c /= sms;
    printf("mode %4d (%s%s%s%s%s%s%s%s%s%s): %7.1f cycles per k-block of 8 MMAs (tensor floor 768)\n", mode, mode & 1 ? "commit " : "", mode & 2 ? "fence " : "",
           mode & 4 ? "spinners " : "", mode & 8 ? "tmem-ld-other " : "", mode & 16 ? "issuer-try_wait " : "", mode & 32 ? "smem-stores " : "",
           mode & 64 ? "tmem-ld-same " : "", mode & 128 ? "rotate-stages " : "", mode & 256 ? "tmem-pingpong " : "", mode & 512 ? "random-data " : "", c / kblocks);
  }
  return 0;
}
