// gemm.cu -- fp32 GEMM used by every non-recurrent GEMM-shaped stage of the path:
//   output projection   logits[B,N]   = h_T[B,H] * W_out^T          (rnn_one_hot.py:65, rnn_margin.py:103)
//   its two gradients   dW_out^T[N,H] = dlogits^T * h_T ; dh_T = dlogits * W_out^T
//   layer>=1 / embedding input GEMM  Xg[T*B,G*H] = in[T*B,I] * W_in  (Lasagne precompute_input)
//   BPTT weight gradient dW_hid[H,G*H] = sum_t h_{t-1}^T * dgates_t  (one K = T*B GEMM after the scan)
//
// C[M,N] = alpha * op(A) * op(B) + beta * C, row-major, fp32 in / fp32 accumulate (SBR_MATH_FP32).
// 128x128x8 CTA tile, 8x8 register micro-tile, register-staged global->shared prefetch; split-K
// over gridDim.z (fp32 atomics) so that tall-K / small-MN problems still fill 148 SMs.
#include <algorithm>

#include "common.cuh"

namespace {

constexpr int BM = 128, BN = 128, BK = 8, PAD = 4;

template <bool TA, bool TB>
__global__ void __launch_bounds__(256) sgemm_kernel(int M, int N, int K, const float* __restrict__ A, int lda,
                                                     const float* __restrict__ B, int ldb, float* __restrict__ C,
                                                     int ldc, float alpha, int k_per_split, int accumulate) {
  __shared__ __align__(16) float As[BK][BM + PAD];
  __shared__ __align__(16) float Bs[BK][BN + PAD];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int k_begin = blockIdx.z * k_per_split;
  const int k_end = min(K, k_begin + k_per_split);

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float ra[4], rb[4];
  auto load_tiles = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + i * 256;
      int mm, kk;
      if (TA) { kk = idx / BM; mm = idx % BM; } else { mm = idx / BK; kk = idx % BK; }
      const int gm = m0 + mm, gk = k0 + kk;
      float v = 0.f;
      if (gm < M && gk < k_end) v = TA ? A[(int64_t)gk * lda + gm] : A[(int64_t)gm * lda + gk];
      ra[i] = v;
      int nn, kb;
      if (TB) { nn = idx / BK; kb = idx % BK; } else { kb = idx / BN; nn = idx % BN; }
      const int gn = n0 + nn, gkb = k0 + kb;
      float w = 0.f;
      if (gn < N && gkb < k_end) w = TB ? B[(int64_t)gn * ldb + gkb] : B[(int64_t)gkb * ldb + gn];
      rb[i] = w;
    }
  };
  auto store_tiles = [&]() {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + i * 256;
      int mm, kk;
      if (TA) { kk = idx / BM; mm = idx % BM; } else { mm = idx / BK; kk = idx % BK; }
      As[kk][mm] = ra[i];
      int nn, kb;
      if (TB) { nn = idx / BK; kb = idx % BK; } else { kb = idx / BN; nn = idx % BN; }
      Bs[kb][nn] = rb[i];
    }
  };

  if (k_begin < k_end) {
    load_tiles(k_begin);
    store_tiles();
    __syncthreads();
    for (int k0 = k_begin; k0 < k_end; k0 += BK) {
      const bool more = k0 + BK < k_end;
      if (more) load_tiles(k0 + BK);
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
        const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][64 + ty * 4]);
        const float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
        const float4 b1 = *reinterpret_cast<const float4*>(&Bs[kk][64 + tx * 4]);
        const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      __syncthreads();
      if (more) {
        store_tiles();
        __syncthreads();
      }
    }
  }

  const bool atomic_out = gridDim.z > 1;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int gm = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int gn = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (gn >= N) continue;
      float* c = C + (int64_t)gm * ldc + gn;
      const float v = alpha * acc[i][j];
      if (atomic_out) atomicAdd(c, v);
      else if (accumulate) *c += v;
      else *c = v;
    }
  }
}


// ---- specialised C += A^T B for K-slow operands (A stored [K, lda], B stored [K, ldb]): the BPTT weight
// gradients dW = sum_rows in[row,:]^T dXg[row,:].  Both tiles are copied global->shared with 16-byte
// cp.async exactly in the layout the FMA loop reads (no transposition), 3-stage pipeline.
constexpr int TN_BK = 16, TN_STAGES = 3;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, int src_bytes) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(d), "l"(gsrc), "r"(src_bytes) : "memory");
}

__global__ void __launch_bounds__(256) sgemm_tn_kernel(int M, int N, int K, const float* __restrict__ A, int lda,
                                                        const float* __restrict__ B, int ldb, float* __restrict__ C,
                                                        int ldc, float alpha, int k_per_split, int accumulate) {
  extern __shared__ __align__(16) float tn_smem[];
  float (*As)[TN_BK][BM] = reinterpret_cast<float (*)[TN_BK][BM]>(tn_smem);
  float (*Bs)[TN_BK][BN] = reinterpret_cast<float (*)[TN_BK][BN]>(tn_smem + TN_STAGES * TN_BK * BM);
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int k_begin = blockIdx.z * k_per_split;
  const int k_end = min(K, k_begin + k_per_split);
  const int n_tiles = (k_end - k_begin + TN_BK - 1) / TN_BK;

  auto issue = [&](int tile, int stage) {
    const int k0 = k_begin + tile * TN_BK;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int idx = tid + i * 256;          // 512 float4 per operand tile
      const int kk = idx >> 5, c4 = (idx & 31) * 4;
      const int gk = k0 + kk;
      const bool kin = gk < k_end;
      const bool ain = kin && (m0 + c4 < M);
      const bool bin = kin && (n0 + c4 < N);
      cp_async16(&As[stage][kk][c4], ain ? A + (int64_t)gk * lda + m0 + c4 : A, ain ? 16 : 0);
      cp_async16(&Bs[stage][kk][c4], bin ? B + (int64_t)gk * ldb + n0 + c4 : B, bin ? 16 : 0);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  for (int s = 0; s < TN_STAGES - 1; ++s) {
    if (s < n_tiles) issue(s, s);
    else asm volatile("cp.async.commit_group;" ::: "memory");
  }
  for (int tile = 0; tile < n_tiles; ++tile) {
    asm volatile("cp.async.wait_group %0;" :: "n"(TN_STAGES - 2) : "memory");
    __syncthreads();
    const int nxt = tile + TN_STAGES - 1;
    if (nxt < n_tiles) issue(nxt, nxt % TN_STAGES);
    else asm volatile("cp.async.commit_group;" ::: "memory");
    const int st = tile % TN_STAGES;
#pragma unroll
    for (int kk = 0; kk < TN_BK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[st][kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[st][kk][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[st][kk][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[st][kk][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
  }
  const bool atomic_out = gridDim.z > 1;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int gm = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int gn = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (gn >= N) continue;
      float* c = C + (int64_t)gm * ldc + gn;
      const float v = alpha * acc[i][j];
      if (atomic_out) atomicAdd(c, v);
      else if (accumulate) *c += v;
      else *c = v;
    }
  }
}

// ---- tensor-core variant of the same K-slow product: warp-level mma.sync m16n8k8 TF32 with the 3xTF32 split
// (hi = tf32(x), lo = x - hi; a_hi*b_hi + a_hi*b_lo + a_lo*b_hi, fp32 accumulate), fragments read straight from
// the cp.async-filled [k][m] / [k][n] tiles (row pitch +8 floats: conflict-free).  8 warps as 2 (M) x 4 (N),
// warp tile 64 x 32.  Operands need no transposition or descriptor; the tcgen05 version of this GEMM is the
// round-2 item.
constexpr int MM_BK = 16, MM_STAGES = 3, MM_PITCH = 128 + 8;

__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
  const float l = x - __uint_as_float(hi);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(l));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

__global__ void __launch_bounds__(256) sgemm_tn_mma_kernel(int M, int N, int K, const float* __restrict__ A, int lda,
                                                            const float* __restrict__ B, int ldb, float* __restrict__ C,
                                                            int ldc, float alpha, int k_per_split, int accumulate) {
  extern __shared__ __align__(16) float mm_smem[];
  float (*As)[MM_BK][MM_PITCH] = reinterpret_cast<float (*)[MM_BK][MM_PITCH]>(mm_smem);
  float (*Bs)[MM_BK][MM_PITCH] = reinterpret_cast<float (*)[MM_BK][MM_PITCH]>(mm_smem + MM_STAGES * MM_BK * MM_PITCH);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = (warp >> 2) * 64, wn = (warp & 3) * 32;       // warp tile origin inside the CTA tile
  const int g = lane >> 2, tq = lane & 3;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int k_begin = blockIdx.z * k_per_split;
  const int k_end = min(K, k_begin + k_per_split);
  const int n_tiles = (k_end - k_begin + MM_BK - 1) / MM_BK;

  auto issue = [&](int tile, int stage) {
    const int k0 = k_begin + tile * MM_BK;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int idx = tid + i * 256;
      const int kk = idx >> 5, c4 = (idx & 31) * 4;
      const int gk = k0 + kk;
      const bool kin = gk < k_end;
      const bool ain = kin && (m0 + c4 < M);
      const bool bin = kin && (n0 + c4 < N);
      cp_async16(&As[stage][kk][c4], ain ? A + (int64_t)gk * lda + m0 + c4 : A, ain ? 16 : 0);
      cp_async16(&Bs[stage][kk][c4], bin ? B + (int64_t)gk * ldb + n0 + c4 : B, bin ? 16 : 0);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  float acc[4][4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[i][j][r] = 0.f;

  for (int s2 = 0; s2 < MM_STAGES - 1; ++s2) {
    if (s2 < n_tiles) issue(s2, s2);
    else asm volatile("cp.async.commit_group;" ::: "memory");
  }
  for (int tile = 0; tile < n_tiles; ++tile) {
    asm volatile("cp.async.wait_group %0;" :: "n"(MM_STAGES - 2) : "memory");
    __syncthreads();
    const int nxt = tile + MM_STAGES - 1;
    if (nxt < n_tiles) issue(nxt, nxt % MM_STAGES);
    else asm volatile("cp.async.commit_group;" ::: "memory");
    const int st = tile % MM_STAGES;
#pragma unroll
    for (int k8 = 0; k8 < MM_BK; k8 += 8) {
      uint32_t bh[4][2], bl[4][2];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        split_tf32(Bs[st][k8 + tq][wn + j * 8 + g], bh[j][0], bl[j][0]);
        split_tf32(Bs[st][k8 + tq + 4][wn + j * 8 + g], bh[j][1], bl[j][1]);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint32_t ah[4], al[4];
        const int mrow = wm + i * 16 + g;
        split_tf32(As[st][k8 + tq][mrow], ah[0], al[0]);
        split_tf32(As[st][k8 + tq][mrow + 8], ah[1], al[1]);
        split_tf32(As[st][k8 + tq + 4][mrow], ah[2], al[2]);
        split_tf32(As[st][k8 + tq + 4][mrow + 8], ah[3], al[3]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          mma_tf32(acc[i][j], al, bh[j]);      // small terms first
          mma_tf32(acc[i][j], ah, bl[j]);
          mma_tf32(acc[i][j], ah, bh[j]);
        }
      }
    }
  }
  const bool atomic_out = gridDim.z > 1;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int gm = m0 + wm + i * 16 + g + (r >= 2 ? 8 : 0);
        const int gn = n0 + wn + j * 8 + 2 * tq + (r & 1);
        if (gm < M && gn < N) {
          float* c = C + (int64_t)gm * ldc + gn;
          const float v = alpha * acc[i][j][r];
          if (atomic_out) atomicAdd(c, v);
          else if (accumulate) *c += v;
          else *c = v;
        }
      }
}

__global__ void zero_matrix_kernel(float* C, int M, int N, int ldc) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)M * N) return;
  C[(i / N) * ldc + (i % N)] = 0.f;
}

}  // namespace

int launch_gemm(sbr_model* m, bool ta, bool tb, int M, int N, int K, const float* A, int lda, const float* B,
                int ldb, float* C, int ldc, float alpha, float beta) {
  if (M <= 0 || N <= 0) return 0;
  if (beta != 0.f && beta != 1.f) {
    sbr_set_error(m, SBR_E_ARG, "gemm: beta must be 0 or 1");
    return SBR_E_ARG;
  }
  {
    const int rc = launch_gemm_tc(m, ta, tb, M, N, K, A, lda, B, ldb, C, ldc, alpha, beta, nullptr);   // tcgen05 3xTF32
    if (rc <= 0) return rc;
  }
  const int tiles = cdiv(M, BM) * cdiv(N, BN);
  int splits = 1;
  if (K > 0) {
    splits = std::max(1, std::min((2 * m->n_sm) / tiles, cdiv(K, 64)));   // <= 2 CTAs per SM: a single wave
    if (tiles >= m->n_sm) splits = 1;
  }
  int kps = K > 0 ? (int)round_up(cdiv(K, splits), 16) : 16;
  splits = K > 0 ? cdiv(K, kps) : 1;
  if (splits > 1 && beta == 0.f) {
    if (ldc == N) {
      cudaError_t e = cudaMemsetAsync(C, 0, (size_t)M * N * sizeof(float), m->stream);
      if (e != cudaSuccess) { sbr_set_error(m, SBR_E_CUDA, "memset: %s", cudaGetErrorString(e)); return SBR_E_CUDA; }
    } else {
      zero_matrix_kernel<<<cdiv((int64_t)M * N, 256), 256, 0, m->stream>>>(C, M, N, ldc);
      KERNEL_CHECK(m);
    }
  }
  const dim3 grid(cdiv(N, BN), cdiv(M, BM), splits);
  const int accumulate = beta == 1.f ? 1 : 0;
  const bool al16 = ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B)) & 15) == 0;
  if (ta && !tb && al16 && M % 4 == 0 && N % 4 == 0 && lda % 4 == 0 && ldb % 4 == 0 && K >= 64) {
    static const bool use_ffma = getenv("SBR_GEMM_FFMA") != nullptr;
    if (!use_ffma) {
      const size_t smem = (size_t)2 * MM_STAGES * MM_BK * MM_PITCH * sizeof(float);
      static std::vector<int> attr_devs;      // per-device attribute
      if (std::find(attr_devs.begin(), attr_devs.end(), m->dev) == attr_devs.end()) {
        cudaFuncSetAttribute(sgemm_tn_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_devs.push_back(m->dev);
      }
      sgemm_tn_mma_kernel<<<grid, 256, smem, m->stream>>>(M, N, K, A, lda, B, ldb, C, ldc, alpha, kps, accumulate);
      KERNEL_CHECK(m);
      return 0;
    }
    const size_t smem = (size_t)TN_STAGES * TN_BK * (BM + BN) * sizeof(float);
    static std::vector<int> attr_devs2;
    if (std::find(attr_devs2.begin(), attr_devs2.end(), m->dev) == attr_devs2.end()) {
      cudaFuncSetAttribute(sgemm_tn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      attr_devs2.push_back(m->dev);
    }
    sgemm_tn_kernel<<<grid, 256, smem, m->stream>>>(M, N, K, A, lda, B, ldb, C, ldc, alpha, kps, accumulate);
    KERNEL_CHECK(m);
    return 0;
  }
  if (!ta && !tb) sgemm_kernel<false, false><<<grid, 256, 0, m->stream>>>(M, N, K, A, lda, B, ldb, C, ldc, alpha, kps, accumulate);
  else if (!ta && tb) sgemm_kernel<false, true><<<grid, 256, 0, m->stream>>>(M, N, K, A, lda, B, ldb, C, ldc, alpha, kps, accumulate);
  else if (ta && !tb) sgemm_kernel<true, false><<<grid, 256, 0, m->stream>>>(M, N, K, A, lda, B, ldb, C, ldc, alpha, kps, accumulate);
  else sgemm_kernel<true, true><<<grid, 256, 0, m->stream>>>(M, N, K, A, lda, B, ldb, C, ldc, alpha, kps, accumulate);
  KERNEL_CHECK(m);
  return 0;
}

__global__ void bias_rows_fill_kernel(float* __restrict__ out, const float* __restrict__ bias, int64_t rows, int cols, int64_t ld) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * cols) return;
  out[(i / cols) * ld + (i % cols)] = bias[i % cols];
}

int launch_gemm_bias(sbr_model* m, bool tb, int M, int N, int K, const float* A, int lda, const float* B, int ldb,
                     float* C, int ldc, const float* bias) {
  if (M <= 0 || N <= 0) return 0;
  {
    const int rc = launch_gemm_tc(m, false, tb, M, N, K, A, lda, B, ldb, C, ldc, 1.f, 0.f, bias);   // bias fused in the epilogue
    if (rc <= 0) return rc;
  }
  bias_rows_fill_kernel<<<cdiv((int64_t)M * N, 256), 256, 0, m->stream>>>(C, bias, M, N, ldc);
  KERNEL_CHECK(m);
  return launch_gemm(m, false, tb, M, N, K, A, lda, B, ldb, C, ldc, 1.f, 1.f);
}
