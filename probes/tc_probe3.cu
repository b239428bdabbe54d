// tc_probe3.cu -- SS tcgen05.mma kind::tf32 with BOTH operands MN-major (the layout of a K-slow GEMM:
// A stored [K][M], B stored [K][N]).  D[128,N] = sum_k A[k][m] * B[k][n].
// smem canonical MN-major, no swizzle: element (mn, k) at (mn/4)*SBO + (k/8)*LBO + (k%8)*16 + (mn%4)*4 bytes.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
               :: "r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}\n" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
// variant 0: LBO = k-group stride, SBO = mn-group stride (as documented for INTERLEAVE MN-major); variant 1: swapped
template <int N>
__global__ void __launch_bounds__(128) k(const float* A, const float* B, float* D, int K, int variant) {
  extern __shared__ __align__(1024) uint8_t raw[];
  float* sA = (float*)raw;              // [K/8][128/4][8][4]
  float* sB = sA + 128 * K;             // [K/8][N/4][8][4]
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 128 * K; i += 128) {
    const int kk = i / 128, m = i % 128;
    int idx;
    if (variant == 2) { const int row = kk % 8, chunk = (m / 4) % 8; idx = (kk / 8) * (128 * 8) + (m / 32) * 256 + row * 32 + ((chunk ^ row) * 4) + (m % 4); }
    else if (variant == 4) idx = (kk / 4) * (128 * 4) + m * 4 + (kk % 4);
    else idx = (kk / 8) * (128 * 8) + (m / 4) * 32 + (kk % 8) * 4 + (m % 4);
    sA[idx] = A[kk * 128 + m];
  }
  for (int i = tid; i < N * K; i += 128) {
    const int kk = i / N, n = i % N;
    int idx;
    if (variant == 2) { const int row = kk % 8, chunk = (n / 4) % 8; idx = (kk / 8) * (N * 8) + (n / 32) * 256 + row * 32 + ((chunk ^ row) * 4) + (n % 4); }
    else if (variant == 3 || variant == 4) idx = (kk / 4) * (N * 4) + n * 4 + (kk % 4);     // K-major interleave [k/4][n][4]
    else idx = (kk / 8) * (N * 8) + (n / 4) * 32 + (kk % 8) * 4 + (n % 4);
    sB[idx] = B[kk * N + n];
  }
  if (tid == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar)));
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" :: "r"(smem_u32(&tmem_s)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_s;
  if (tid == 0) {
    // kind::tf32, fp32 accum, A MN-major (bit 15), B MN-major (bit 16)
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((variant == 4 ? 0u : 1u) << 15) | (((variant == 3 || variant == 4) ? 0u : 1u) << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    for (int ks = 0; ks < K / 8; ++ks) {
      const uint32_t a_addr = smem_u32(sA) + ks * 128 * 8 * 4, b_addr = smem_u32(sB) + ks * N * 8 * 4;
      const uint32_t mn_stride = 128;               // bytes between groups of 4 mn-elements
      const uint32_t a_kstride = 128 * 8 * 4, b_kstride = N * 8 * 4;   // bytes between groups of 8 k (unused within one MMA)
      uint64_t ad, bd;
      if (variant == 0) { ad = make_desc(a_addr, a_kstride, mn_stride); bd = make_desc(b_addr, b_kstride, mn_stride); }
      else if (variant == 2) {   // SW128: LBO = stride between 32-element MN groups (1024 B), SBO = stride between 8-row K groups
        ad = make_desc(a_addr, 1024, a_kstride) | ((uint64_t)2 << 61);
        bd = make_desc(b_addr, 1024, b_kstride) | ((uint64_t)2 << 61);
      } else if (variant == 4) {
        ad = make_desc(smem_u32(sA) + ks * 2 * (128 * 16), 128 * 16, 128);
        bd = make_desc(smem_u32(sB) + ks * 2 * (N * 16), N * 16, 128);
      } else if (variant == 3) {
        ad = make_desc(a_addr, a_kstride, mn_stride);
        bd = make_desc(smem_u32(sB) + ks * 2 * (N * 16), N * 16, 128);
      }
      else { ad = make_desc(a_addr, mn_stride, a_kstride); bd = make_desc(b_addr, mn_stride, b_kstride); }
      mma_ss(tmem, ad, bd, idesc, ks > 0);
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar)) : "memory");
  }
  mbar_wait(&bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int m = warp * 32 + lane;
  for (int c0 = 0; c0 < N; c0 += 8) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(tmem + ((uint32_t)(warp * 32) << 16) + c0));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 8; ++j) D[m * N + c0 + j] = __uint_as_float(r[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" :: "r"(tmem));
}
template <int N> void run(int K, int variant) {
  std::vector<float> A(128 * K), B(N * K), D(128 * N), ref(128 * N);
  srand(3 + K);
  for (auto& v : A) v = (float)((rand() % 17) - 8) * 0.25f;
  for (auto& v : B) v = (float)((rand() % 13) - 6) * 0.5f;
  for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) { double s = 0; for (int kk = 0; kk < K; ++kk) s += (double)A[kk * 128 + m] * B[kk * N + n]; ref[m * N + n] = (float)s; }
  float *dA, *dB, *dD;
  CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0xFF, D.size() * 4));
  size_t smem = (size_t)(128 + N) * K * 4 + 2048;
  CK(cudaFuncSetAttribute(k<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k<N><<<1, 128, smem>>>(dA, dB, dD, K, variant);
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
  int bad = 0; double maxerr = 0;
  for (int i = 0; i < 128 * N; ++i) { double e = fabs((double)D[i] - ref[i]); if (!(e == e)) e = 1e30; if (e > maxerr) maxerr = e; if (e > 1e-3) bad++; }
  printf("MN-major N=%d K=%d variant=%d: max err %.3e mismatches %d  D[0..2]=%g %g %g  ref=%g %g %g  D[N]=%g ref=%g\n", N, K, variant, maxerr, bad, D[0], D[1], D[2], ref[0], ref[1], ref[2], D[N], ref[N]);
}
int main() {
  run<32>(8, 4); run<32>(32, 4); run<32>(8, 0); run<32>(8, 2); run<32>(8, 3);
  run<32>(32, 2); run<32>(32, 3);
  run<160>(32, 2); run<160>(32, 3);
  return 0;
}
