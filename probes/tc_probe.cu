// tc_probe.cu -- standalone validation of the tcgen05 building blocks used by the recurrent kernel:
//   (1) SS: D[128,N] = A[128,K] * B[N,K]^T, A and B in shared memory (K-major, no swizzle)
//   (2) TS: same with A staged in TMEM by tcgen05.st
//   (3) 3xTF32 split accuracy on random fp32 data
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tc_probe tc_probe.cu
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, SWIZZLE_NONE canonical layout: [k/4][mn][k%4] floats; core matrix = 8 rows x 16 B
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // version = 1 (sm100)
  return d;                // layout_type = 0 (no swizzle), base_offset = 0
}

__device__ __forceinline__ uint32_t make_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
               :: "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
               :: "r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
               "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
               "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}\n" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}

// mode 0: SS, 1: TS, 2: TS 3xTF32
template <int N>
__global__ void __launch_bounds__(128) probe_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                     float* __restrict__ D, int K, int mode) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  float* sA = reinterpret_cast<float*>(smem_raw);            // [K/4][128][4]
  float* sB = sA + 128 * K;                                  // [K/4][N][4]   (hi)
  float* sBlo = sB + N * K;                                  // lo part for 3xTF32
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  for (int i = tid; i < 128 * K; i += 128) {
    const int m = i / K, k = i % K;
    sA[(k / 4) * (128 * 4) + m * 4 + (k % 4)] = A[i];
  }
  for (int i = tid; i < N * K; i += 128) {
    const int n = i / K, k = i % K;
    float v = B[i];
    float hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    if (mode != 2) hi = v;
    sB[(k / 4) * (N * 4) + n * 4 + (k % 4)] = hi;
    sBlo[(k / 4) * (N * 4) + n * 4 + (k % 4)] = v - hi;
  }
  if (tid == 0) mbar_init(&bar, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic smem writes -> async proxy (UMMA)
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  const uint32_t tD = tmem;             // columns [0, N)
  const uint32_t tAhi = tmem + 32;      // columns [32, 32+K)
  const uint32_t tAlo = tmem + 32 + 224;

  if (mode >= 1) {
    // stage A in TMEM: thread (warp w, lane i) owns TMEM lane 32w+i = row m; columns = k
    const int m = warp * 32 + lane;
    for (int k0 = 0; k0 < K; k0 += 8) {
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float v = A[m * K + k0 + j];
        const float h = (mode == 2) ? __uint_as_float(__float_as_uint(v) & 0xFFFFE000u) : v;
        hi[j] = __float_as_uint(h);
        lo[j] = __float_as_uint(v - h);
      }
      const uint32_t addr_hi = tAhi + ((uint32_t)(warp * 32) << 16) + k0;
      const uint32_t addr_lo = tAlo + ((uint32_t)(warp * 32) << 16) + k0;
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                   :: "r"(addr_hi), "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]), "r"(hi[4]), "r"(hi[5]), "r"(hi[6]), "r"(hi[7]) : "memory");
      if (mode == 2)
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                     :: "r"(addr_lo), "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]), "r"(lo[4]), "r"(lo[5]), "r"(lo[6]), "r"(lo[7]) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }

  if (tid == 0) {
    const uint32_t idesc = make_idesc_tf32(128, N);
    uint32_t acc = 0;
    for (int ks = 0; ks < K / 8; ++ks) {
      const uint64_t bdesc = make_desc(smem_u32(sB) + ks * 2 * (N * 16), N * 16, 128);
      if (mode == 0) {
        const uint64_t adesc = make_desc(smem_u32(sA) + ks * 2 * (128 * 16), 128 * 16, 128);
        mma_ss(tD, adesc, bdesc, idesc, acc);
      } else if (mode == 1) {
        mma_ts(tD, tAhi + ks * 8, bdesc, idesc, acc);
      } else {
        const uint64_t bdesc_lo = make_desc(smem_u32(sBlo) + ks * 2 * (N * 16), N * 16, 128);
        mma_ts(tD, tAhi + ks * 8, bdesc, idesc, acc);
        mma_ts(tD, tAhi + ks * 8, bdesc_lo, idesc, 1);
        mma_ts(tD, tAlo + ks * 8, bdesc, idesc, 1);
      }
      acc = 1;
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  {
    const int m = warp * 32 + lane;
    for (int c0 = 0; c0 < N; c0 += 8) {
      uint32_t r[8];
      const uint32_t addr = tD + ((uint32_t)(warp * 32) << 16) + c0;
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(addr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int j = 0; j < 8; ++j) D[m * N + c0 + j] = __uint_as_float(r[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512));
}

template <int N>
void run(int K, int mode, bool exact) {
  std::vector<float> A(128 * K), B(N * K), D(128 * N), ref(128 * N);
  srand(7 + K + mode);
  for (auto& v : A) v = exact ? (float)((rand() % 17) - 8) * 0.25f : ((float)rand() / RAND_MAX - 0.5f);
  for (auto& v : B) v = exact ? (float)((rand() % 13) - 6) * 0.5f : ((float)rand() / RAND_MAX - 0.5f);
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < N; ++n) {
      double s = 0;
      for (int k = 0; k < K; ++k) s += (double)A[m * K + k] * B[n * K + k];
      ref[m * N + n] = (float)s;
    }
  float *dA, *dB, *dD;
  CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0xFF, D.size() * 4));
  size_t smem = (size_t)(128 * K + 2 * N * K) * 4 + 128;
  CK(cudaFuncSetAttribute(probe_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  probe_kernel<N><<<1, 128, smem>>>(dA, dB, dD, K, mode);
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
  double maxerr = 0, maxref = 0;
  int bad = 0;
  for (int i = 0; i < 128 * N; ++i) {
    double e = fabs((double)D[i] - ref[i]);
    if (!(e == e)) e = 1e30;
    if (e > maxerr) maxerr = e;
    if (fabs(ref[i]) > maxref) maxref = fabs(ref[i]);
    if (e > 1e-3 * (1 + fabs(ref[i]))) bad++;
  }
  printf("N=%3d K=%3d mode=%d %s: max|err|=%.3e (max|ref|=%.3f) mismatches=%d  D[0..3]=%g %g %g %g ref=%g %g %g %g\n", N, K, mode,
         exact ? "exact " : "random", maxerr, maxref, bad, D[0], D[1], D[2], D[3], ref[0], ref[1], ref[2], ref[3]);
  cudaFree(dA); cudaFree(dB); cudaFree(dD);
}

int main() {
  run<16>(8, 0, true);
  run<16>(16, 0, true);
  run<16>(200, 0, true);
  run<32>(64, 0, true);
  run<16>(8, 1, true);
  run<16>(200, 1, true);
  run<32>(200, 1, true);
  run<16>(200, 0, false);
  run<16>(200, 1, false);
  run<16>(200, 2, false);
  run<32>(104, 2, false);
  return 0;
}
