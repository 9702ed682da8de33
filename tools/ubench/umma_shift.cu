// Micro-experiment: can a tcgen05 SWIZZLE_128B K-major A descriptor start at an arbitrary 128-byte ROW of a larger
// swizzled buffer (start address not 1024-byte aligned), with a stride between 8-row groups (SBO) that is not 1024?
// This is what an implicit-GEMM convolution needs to read its nine shifted taps out of ONE halo tile in shared memory.
//   buffer: R rows x 64 halfs (128 B per row), written the way TMA writes a SWIZZLE_128B box into a 1024-aligned buffer:
//           16-byte chunk c of row r lives at r*128 + ((c ^ (r & 7)) << 4)
//   A(m, k) = buffer[(m / 8) * sbo_rows + (m % 8) + shift][k],   D = A * B^T   (M = 128, N = 16, K = 64)
// Modes: base_offset field (descriptor bits 49..51) = 0, or (start_address >> 7) & 7.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o umma_shift umma_shift.cu ; run: ./umma_shift
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

constexpr int R = 512, N = 16, M = 128, K = 64;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128, 1) k_shift(const __half *__restrict__ buf, const __half *__restrict__ bmat, int shift, int sbo_rows,
                                                  int use_base_offset, float *__restrict__ out) {
  extern __shared__ uint8_t raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  uint8_t *sa = smem;                       // R * 128 bytes
  uint8_t *sb = smem + R * 128;             // N * 128 bytes
  uint64_t *bar = reinterpret_cast<uint64_t *>(sb + N * 128);
  uint32_t *slot = reinterpret_cast<uint32_t *>(bar + 1);
  for (int i = threadIdx.x; i < R * 8; i += blockDim.x) {
    const int r = i >> 3, c = i & 7;
    *reinterpret_cast<uint4 *>(sa + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4 *>(buf + (size_t)r * 64 + c * 8);
  }
  for (int i = threadIdx.x; i < N * 8; i += blockDim.x) {
    const int r = i >> 3, c = i & 7;
    *reinterpret_cast<uint4 *>(sb + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4 *>(bmat + (size_t)r * 64 + c * 8);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    const uint32_t a_start = smem_u32(sa) + (uint32_t)shift * 128u;
    uint64_t adesc = (uint64_t)((a_start & 0x3FFFF) >> 4) | ((uint64_t)((sbo_rows * 128) >> 4) << 32) | (1ull << 46) | (2ull << 61);
    if (use_base_offset) adesc |= (uint64_t)((a_start >> 7) & 7) << 49;
    const uint64_t bdesc = (uint64_t)((smem_u32(sb) & 0x3FFFF) >> 4) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    for (int kk = 0; kk < K / 16; ++kk) {
      const uint32_t acc = kk ? 1u : 0u;
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem),
          "l"(adesc + (uint64_t)(kk * 2)), "l"(bdesc + (uint64_t)(kk * 2)), "r"(idesc), "r"(acc)
          : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
  }
  {
    uint32_t done = 0;
    while (!done)
      asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}"
                   : "=r"(done)
                   : "r"(smem_u32(bar)), "r"(0)
                   : "memory");
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t r[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(tmem + ((uint32_t)(warp * 32) << 16)));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int j = 0; j < 16; ++j) out[(warp * 32 + lane) * N + j] = __uint_as_float(r[j]);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tmem) : "memory");
}

int main() {
  std::vector<__half> hb((size_t)R * 64), hw((size_t)N * 64);
  std::vector<float> fb((size_t)R * 64), fw((size_t)N * 64);
  srand(1);
  for (size_t i = 0; i < hb.size(); ++i) { fb[i] = (float)((rand() % 17) - 8); hb[i] = __float2half(fb[i]); }
  for (size_t i = 0; i < hw.size(); ++i) { fw[i] = (float)((rand() % 9) - 4); hw[i] = __float2half(fw[i]); }
  __half *dbuf, *dw; float *dout;
  cudaMalloc(&dbuf, hb.size() * 2); cudaMalloc(&dw, hw.size() * 2); cudaMalloc(&dout, M * N * 4);
  cudaMemcpy(dbuf, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dw, hw.data(), hw.size() * 2, cudaMemcpyHostToDevice);
  const int smem = R * 128 + N * 128 + 1024 + 64;
  cudaFuncSetAttribute(k_shift, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int shifts[] = {0, 1, 2, 3, 5, 7, 8, 10, 11, 12, 20, 21, 22, 33};
  const int sbos[] = {8, 10, 16, 18, 24};
  std::vector<float> out(M * N);
  for (int sbo : sbos)
    for (int bo = 0; bo < 2; ++bo) {
      printf("sbo_rows=%2d base_offset=%s :", sbo, bo ? "set " : "zero");
      for (int sh : shifts) {
        cudaMemset(dout, 0, M * N * 4);
        k_shift<<<1, 128, smem>>>(dbuf, dw, sh, sbo, bo, dout);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf(" shift %d: CUDA error %s\n", sh, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(out.data(), dout, M * N * 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int m = 0; m < M; ++m) {
          const int row = (m / 8) * sbo + (m % 8) + sh;
          for (int n = 0; n < N; ++n) {
            float ref = 0.f;
            for (int k = 0; k < K; ++k) ref += fb[(size_t)row * 64 + k] * fw[(size_t)n * 64 + k];
            if (ref != out[m * N + n]) ++bad;
          }
        }
        printf(" s%d:%s", sh, bad ? "BAD" : "ok");
      }
      printf("\n");
    }
  return 0;
}
