// tc_probe2.cu -- timing of tcgen05.mma issue patterns (M=128, N=16, K=8 tf32, A in TMEM):
//   how many cycles do 75 MMAs take when they accumulate into 1 / 2 / 4 / 8 different TMEM tiles?
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
               :: "r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ uint32_t elect_one_sync() {
  uint32_t pred = 0, laneid = 0;
  asm volatile("{\n.reg .b32 %%rx;\n.reg .pred %%px;\n     elect.sync %%rx|%%px, %2;\n@%%px mov.s32 %1, 1;\n     mov.s32 %0, %%rx;\n}\n" : "+r"(laneid), "+r"(pred) : "r"(0xFFFFFFFF));
  return pred;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}\n" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
template <int N>
__global__ void __launch_bounds__(128) k(long long* out, int n_acc, int n_mma, int reps) {
  __shared__ __align__(128) float sB[N * 200];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_s;
  for (int i = threadIdx.x; i < N * 200; i += 128) sB[i] = 0.001f * (i % 7);
  if (threadIdx.x == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar)));
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(smem_u32(&tmem_s)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_s;
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  long long t_issue = 0, t_total = 0;
  for (int r = 0; r < reps; ++r) {
    __syncthreads();
    const long long t0 = clock64();
    if (threadIdx.x < 32) {
      if (elect_one_sync()) {
        uint64_t d = make_desc(smem_u32(sB), N * 16, 128);
        uint32_t acol = 256;
        for (int i = 0; i < n_mma; ++i) {
          mma_ts(tmem + (i % n_acc) * N, tmem + acol, d, idesc, i >= n_acc);
          d += 2 * N; acol += 8; if (acol >= 256 + 8 * 24) { acol = 256; d = make_desc(smem_u32(sB), N * 16, 128); }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar)) : "memory");
      }
      __syncwarp();
    }
    const long long t1 = clock64();
    mbar_wait(&bar, r & 1);
    const long long t2 = clock64();
    if (threadIdx.x == 0) { t_issue += t1 - t0; t_total += t2 - t0; }
  }
  if (threadIdx.x == 0) { out[0] = t_issue / reps; out[1] = t_total / reps; }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmem));
}
int main() {
  long long* d; CK(cudaMalloc(&d, 16));
  long long h[2];
  int accs[] = {1, 2, 3, 4, 6, 8};
  for (int n_mma : {25, 75}) for (int na : accs) {
    k<16><<<1, 128>>>(d, na, n_mma, 20); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost));
    printf("N=16 n_mma=%d n_acc=%d : issue %lld cyc, issue+complete %lld cyc (%.1f / MMA)\n", n_mma, na, h[0], h[1], (double)h[1] / n_mma);
  }
  for (int na : {1, 4}) {
    k<32><<<1, 128>>>(d, na, 75, 20); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost));
    printf("N=32 n_mma=75 n_acc=%d : issue %lld cyc, issue+complete %lld cyc (%.1f / MMA)\n", na, h[0], h[1], (double)h[1] / 75);
  }
  return 0;
}
