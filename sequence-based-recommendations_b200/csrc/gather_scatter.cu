// gather_scatter.cu -- stage 1 of the hot path and its transpose.
//
//   gather : Xg[t*B+b, :] = sum_k W_in[X[b,t,k], :] + bias        (valid steps only)
//            reference: W_in_stacked[input, :].sum(axis=-2) + b_stacked
//            (neural_networks/sparse_lstm.py:368, :755, :1111)
//   scatter: dW_in[X[b,t,k], :] += dXg[t*B+b, :]                   (duplicates accumulate)
//            reference: gradient of the AdvancedSubtensor1 above (theano.grad, rnn_base.py:183)
//
// Both are HBM-bound row copies.  Rows are G*H floats (0.6-8 KB) so one warp moves one row with
// 128-bit accesses; the grid is a multiple of the SM count.  The scatter is duplicate-heavy: a
// reference mini-batch is B nested prefixes of ONE user's sequence (rnn_base.py:396-415), so at a
// given timestep up to B rows carry the same item id.  A warp therefore owns 32 consecutive time-major
// (t, b) rows x 128 columns, sums runs of equal ids in registers (run-length merge with warp shuffles)
// and issues ONE 16-byte red.global.add.v4.f32 per run (warp-aggregated atomics) instead of one per row.
#include <stdlib.h>
#include <utility>
#include <vector>

#include "common.cuh"

namespace {

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
// one 16-byte vector reduction (sm_90+) instead of four scalar REDs: 4x fewer atomic operations at the L2
__device__ __forceinline__ void red_add_v4(float* p, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// one warp per (t, b) row; VEC4 requires ncols % 4 == 0 and 16B-aligned bases
template <bool VEC4>
__global__ void __launch_bounds__(256) gather_rows_kernel(const int32_t* __restrict__ X, const int32_t* __restrict__ len,
                                                           const float* __restrict__ W, const float* __restrict__ bias,
                                                           float* __restrict__ out, int B, int T, int K, int ncols,
                                                           int n_rows, int n_table) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  for (int row = blockIdx.x * warps_per_block + (threadIdx.x >> 5); row < n_rows; row += gridDim.x * warps_per_block) {
    const int t = row / B, b = row - t * B;
    if (t >= len[b]) continue;  // padded step: never read downstream
    const int32_t* ids = X + ((int64_t)b * T + t) * K;
    float* o = out + (int64_t)row * ncols;
    if (VEC4) {
      for (int c = lane * 4; c < ncols; c += 128) {
        float4 acc = ld4(bias + c);
        for (int k = 0; k < K; ++k) {
          int id = __ldg(ids + k);
          id = min(max(id, 0), n_table - 1);
          float4 w = ld4(W + (int64_t)id * ncols + c);
          acc.x += w.x; acc.y += w.y; acc.z += w.z; acc.w += w.w;
        }
        st4(o + c, acc);
      }
    } else {
      for (int c = lane; c < ncols; c += 32) {
        float acc = bias[c];
        for (int k = 0; k < K; ++k) {
          int id = __ldg(ids + k);
          id = min(max(id, 0), n_table - 1);
          acc += W[(int64_t)id * ncols + c];
        }
        o[c] = acc;
      }
    }
  }
}

// TMA-staged gather (opt-in, SBR_GATHER_TMA=1; rows must be multiples of 16 bytes): the table rows of a (t, b) entry are
// fetched into shared memory by the bulk-copy engine -- ONE cp.async.bulk of ncols*4 bytes per id, completion on an
// mbarrier -- while the warp sums and writes the previous entry; each warp runs its own two-stage ring, so a whole
// row per stage is in flight instead of the one or two 16-byte loads per lane of the register path.
//   stage layout: [K][ncols] floats; warp w of the block owns stages 2w, 2w+1 and barriers 2w, 2w+1.
constexpr int GT_WARPS = 4;
__device__ __forceinline__ uint32_t gs_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(GT_WARPS * 32) gather_rows_tma_kernel(const int32_t* __restrict__ X, const int32_t* __restrict__ len,
                                                                         const float* __restrict__ W, const float* __restrict__ bias,
                                                                         float* __restrict__ out, int B, int T, int K, int ncols,
                                                                         int n_rows, int n_table) {
  extern __shared__ __align__(128) float gt_smem[];
  __shared__ __align__(8) uint64_t bars[GT_WARPS * 2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int stage_floats = K * ncols;
  float* st0 = gt_smem + (size_t)warp * 2 * stage_floats;
  uint64_t* bar = bars + warp * 2;
  if (lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(gs_smem_u32(&bar[0])) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(gs_smem_u32(&bar[1])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const int stride = gridDim.x * GT_WARPS;
  auto next_valid = [&](int r) {           // first row >= r (on this warp's stride) whose step lies inside its sequence
    for (; r < n_rows; r += stride) {
      const int t = r / B, b = r - t * B;
      if (t < len[b]) return r;
    }
    return -1;
  };
  auto issue = [&](int r, int s) {         // lane 0: K bulk copies of one table row each
    const int t = r / B, b = r - t * B;
    const int32_t* ids = X + ((int64_t)b * T + t) * K;
    const uint32_t bytes = (uint32_t)ncols * 4u;
    const uint32_t bar_a = gs_smem_u32(&bar[s]);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar_a), "r"(bytes * (uint32_t)K) : "memory");
    for (int k = 0; k < K; ++k) {
      int id = __ldg(ids + k);
      id = min(max(id, 0), n_table - 1);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   :: "r"(gs_smem_u32(st0 + (size_t)s * stage_floats + (size_t)k * ncols)), "l"(W + (int64_t)id * ncols), "r"(bytes), "r"(bar_a)
                   : "memory");
    }
  };
  int r = next_valid(blockIdx.x * GT_WARPS + warp);
  int s = 0;
  uint32_t phase[2] = {0u, 0u};
  if (r >= 0 && lane == 0) issue(r, 0);
  while (r >= 0) {
    const int rn = next_valid(r + stride);
    if (rn >= 0) {
      // the other stage was read (generic proxy) one iteration ago: order those reads before the engine overwrites it
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) issue(rn, s ^ 1);
    }
    {
      const uint32_t bar_a = gs_smem_u32(&bar[s]);
      asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}\n"
                   :: "r"(bar_a), "r"(phase[s]) : "memory");
      phase[s] ^= 1u;
    }
    const float* src = st0 + (size_t)s * stage_floats;
    float* o = out + (int64_t)r * ncols;
    for (int c = lane * 4; c < ncols; c += 128) {
      float4 acc = ld4(bias + c);
      for (int k = 0; k < K; ++k) {
        const float4 w = *reinterpret_cast<const float4*>(src + (size_t)k * ncols + c);
        acc.x += w.x; acc.y += w.y; acc.z += w.z; acc.w += w.w;
      }
      st4(o + c, acc);
    }
    s ^= 1;
    r = rn;
  }
}

// Warp-aggregated scatter-add.  A warp owns 32 consecutive time-major entries e = t*B + b (same t,
// consecutive b => same id under the nested-prefix batch layout).
template <bool VEC4>
__global__ void __launch_bounds__(256) scatter_add_rows_kernel(const int32_t* __restrict__ X,
                                                                const int32_t* __restrict__ len,
                                                                const float* __restrict__ dOut, float* __restrict__ dW,
                                                                int B, int T, int K, int ncols, int n_rows,
                                                                int n_chunks) {
  // one warp = 32 consecutive (t, b) rows x one 128-column chunk (VEC4) / 32-column chunk.  Runs of equal ids
  // (nested-prefix batches put the same item in the whole warp) are summed in registers and leave as ONE reduction;
  // distinct ids leave as independent load -> RED pairs, eight loads in flight per lane.
  const int lane = threadIdx.x & 31;
  const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int grp = w / n_chunks;
  if (grp * 32 >= n_rows) return;
  const int chunk = w - grp * n_chunks;
  const int e = grp * 32 + lane;
  int t = 0, b = 0;
  bool valid = e < n_rows;
  if (valid) {
    t = e / B;
    b = e - t * B;
    valid = t < len[b];
  }
  const float* src0 = dOut + (int64_t)grp * 32 * ncols;
  for (int k = 0; k < K; ++k) {
    const int id = valid ? X[((int64_t)b * T + t) * K + k] : -1;
    if (VEC4) {
      const int c = chunk * 128 + lane * 4;
      const bool act = c < ncols;
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
      int cur = -1;
#pragma unroll
      for (int r0 = 0; r0 < 32; r0 += 8) {
        float4 v[8];
        int rid[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          rid[i] = __shfl_sync(0xffffffffu, id, r0 + i);
          v[i] = (rid[i] >= 0 && act) ? __ldg(reinterpret_cast<const float4*>(src0 + (int64_t)(r0 + i) * ncols + c))
                                      : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (rid[i] < 0) continue;
          if (rid[i] != cur) {
            if (cur >= 0 && act) red_add_v4(dW + (int64_t)cur * ncols + c, s);
            s = make_float4(0.f, 0.f, 0.f, 0.f);
            cur = rid[i];
          }
          s.x += v[i].x; s.y += v[i].y; s.z += v[i].z; s.w += v[i].w;
        }
      }
      if (cur >= 0 && act) red_add_v4(dW + (int64_t)cur * ncols + c, s);
    } else {
      const int c = chunk * 32 + lane;
      const bool act = c < ncols;
      float s = 0.f;
      int cur = -1;
      for (int r = 0; r < 32; ++r) {
        const int rid = __shfl_sync(0xffffffffu, id, r);
        if (rid < 0) continue;
        const float v = act ? src0[(int64_t)r * ncols + c] : 0.f;
        if (rid != cur) {
          if (cur >= 0 && act) atomicAdd(dW + (int64_t)cur * ncols + c, s);
          s = 0.f;
          cur = rid;
        }
        s += v;
      }
      if (cur >= 0 && act) atomicAdd(dW + (int64_t)cur * ncols + c, s);
    }
  }
}

// out[c] += sum_r A[r*ld + c]   (A is zero on masked rows, so no mask is needed)
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ A, int rows, int cols, int ld,
                                                      int rows_per_block, float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(rows, r0 + rows_per_block);
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  int r = r0;
  for (; r + 3 < r1; r += 4) {
    s0 += A[(int64_t)r * ld + c];
    s1 += A[(int64_t)(r + 1) * ld + c];
    s2 += A[(int64_t)(r + 2) * ld + c];
    s3 += A[(int64_t)(r + 3) * ld + c];
  }
  for (; r < r1; ++r) s0 += A[(int64_t)r * ld + c];
  atomicAdd(out + c, (s0 + s1) + (s2 + s3));
}

// rows of an item-major table (output embeddings): out_rows[i,:] = table[ids[i],:], out_bias[i] = bias[ids[i]]
__global__ void __launch_bounds__(256) gather_table_rows_kernel(const float* __restrict__ table,
                                                                 const float* __restrict__ bias,
                                                                 const int32_t* __restrict__ ids, int n_ids, int ncols,
                                                                 float* __restrict__ out_rows,
                                                                 float* __restrict__ out_bias) {
  const int lane = threadIdx.x & 31;
  const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= n_ids) return;
  const int id = ids[w];
  for (int c = lane; c < ncols; c += 32) out_rows[(int64_t)w * ncols + c] = table[(int64_t)id * ncols + c];
  if (lane == 0 && out_bias) out_bias[w] = bias[id];
}

__global__ void __launch_bounds__(256) scatter_table_rows_kernel(const float* __restrict__ rows,
                                                                  const float* __restrict__ brow,
                                                                  const int32_t* __restrict__ ids, int n_ids,
                                                                  int ncols, float* __restrict__ table_grad,
                                                                  float* __restrict__ bias_grad) {
  const int lane = threadIdx.x & 31;
  const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= n_ids) return;
  const int id = ids[w];
  for (int c = lane; c < ncols; c += 32) atomicAdd(table_grad + (int64_t)id * ncols + c, rows[(int64_t)w * ncols + c]);
  if (lane == 0 && bias_grad) atomicAdd(bias_grad + id, brow[w]);
}

// out[c*rows + r] = in[r*ld_in + c]
__global__ void transpose_kernel(const float* __restrict__ in, int rows, int cols, int ld_in, float* __restrict__ out) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < rows && c < cols) ? in[(int64_t)r * ld_in + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (r < rows && c < cols) out[(int64_t)c * rows + r] = tile[threadIdx.x][i];
  }
}

// EmbeddingLayer (recurrent_layers.py:47-50): out[t*B+b, k*E + e] = table[X[b,t,k], e]
__global__ void __launch_bounds__(256) embed_gather_kernel(const int32_t* __restrict__ X, const int32_t* __restrict__ len,
                                                            const float* __restrict__ table, float* __restrict__ out,
                                                            int B, int T, int K, int E, int n_rows) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n_rows) return;
  const int t = row / B, b = row - t * B;
  const bool valid = t < len[b];
  for (int k = 0; k < K; ++k) {
    const int id = valid ? X[((int64_t)b * T + t) * K + k] : 0;
    for (int e = lane; e < E; e += 32)
      out[(int64_t)row * K * E + k * E + e] = valid ? table[(int64_t)id * E + e] : 0.f;
  }
}

__global__ void __launch_bounds__(256) embed_scatter_kernel(const int32_t* __restrict__ X, const int32_t* __restrict__ len,
                                                             const float* __restrict__ dOut, float* __restrict__ dTable,
                                                             int B, int T, int K, int E, int n_rows) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n_rows) return;
  const int t = row / B, b = row - t * B;
  if (t >= len[b]) return;
  for (int k = 0; k < K; ++k) {
    const int id = X[((int64_t)b * T + t) * K + k];
    for (int e = lane; e < E; e += 32) atomicAdd(dTable + (int64_t)id * E + e, dOut[(int64_t)row * K * E + k * E + e]);
  }
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

int launch_gather_rows(sbr_model* m, const int32_t* X, const int32_t* len, const float* W, const float* bias,
                       float* out, int B, int T, int K, int ncols, int t_max, int n_rows_table) {
  const int n_rows = t_max * B;
  if (n_rows == 0) return 0;
  const int wpb = 8;
  int grid = cdiv(n_rows, wpb);
  grid = std::min(grid, m->n_sm * 16);
  const bool v4 = (ncols % 4 == 0) && aligned16(W) && aligned16(bias) && aligned16(out);
  const size_t tma_smem = (size_t)GT_WARPS * 2 * K * ncols * sizeof(float);
  // measured on the C2 batches (9 122 valid rows of 3.2 KB): 19.2 us staged vs 16.9 us for the register path, so the
  // staged kernel is opt-in (SBR_GATHER_TMA=1)
  const bool use_tma = getenv("SBR_GATHER_TMA") != nullptr;
  if (v4 && use_tma && tma_smem <= 200 * 1024) {
    static std::vector<std::pair<int, size_t>> attr;      // (device, bytes) already opted in
    bool have = false;
    for (auto& e : attr) have = have || (e.first == m->dev && e.second >= tma_smem);
    if (!have) {
      CU_TRY(m, cudaFuncSetAttribute(gather_rows_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024)));
      attr.push_back({m->dev, (size_t)200 * 1024});
    }
    const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (200 * 1024) / tma_smem));
    const int g = std::min(cdiv(n_rows, GT_WARPS), m->n_sm * per_sm);
    gather_rows_tma_kernel<<<g, GT_WARPS * 32, tma_smem, m->stream>>>(X, len, W, bias, out, B, T, K, ncols, n_rows, n_rows_table);
    KERNEL_CHECK(m);
    return 0;
  }
  if (v4)
    gather_rows_kernel<true><<<grid, wpb * 32, 0, m->stream>>>(X, len, W, bias, out, B, T, K, ncols, n_rows, n_rows_table);
  else
    gather_rows_kernel<false><<<grid, wpb * 32, 0, m->stream>>>(X, len, W, bias, out, B, T, K, ncols, n_rows, n_rows_table);
  KERNEL_CHECK(m);
  return 0;
}

int launch_scatter_add_rows(sbr_model* m, const int32_t* X, const int32_t* len, const float* dOut, float* dW,
                            int B, int T, int K, int ncols, int t_max) {
  const int n_rows = t_max * B;
  if (n_rows == 0) return 0;
  const int wpb = 8;
  const bool v4 = (ncols % 4 == 0) && aligned16(dOut) && aligned16(dW);
  const int n_chunks = cdiv(ncols, v4 ? 128 : 32);
  const long long warps = (long long)cdiv(n_rows, 32) * n_chunks;
  const int grid = (int)((warps + wpb - 1) / wpb);
  if (v4)
    scatter_add_rows_kernel<true><<<grid, wpb * 32, 0, m->stream>>>(X, len, dOut, dW, B, T, K, ncols, n_rows, n_chunks);
  else
    scatter_add_rows_kernel<false><<<grid, wpb * 32, 0, m->stream>>>(X, len, dOut, dW, B, T, K, ncols, n_rows, n_chunks);
  KERNEL_CHECK(m);
  return 0;
}

int launch_colsum(sbr_model* m, const float* A, int rows, int cols, int ld, float* out) {
  if (rows == 0 || cols == 0) return 0;
  const int gx = cdiv(cols, 256);
  int gy = std::max(1, std::min(cdiv(rows, 64), (m->n_sm * 4) / gx));
  const int rpb = cdiv(rows, gy);
  gy = cdiv(rows, rpb);
  colsum_kernel<<<dim3(gx, gy), 256, 0, m->stream>>>(A, rows, cols, ld, rpb, out);
  KERNEL_CHECK(m);
  return 0;
}

int launch_gather_table_rows(sbr_model* m, const float* table, const float* bias, const int32_t* ids, int n_ids,
                             int ncols, float* out_rows, float* out_bias) {
  if (n_ids == 0) return 0;
  gather_table_rows_kernel<<<cdiv(n_ids, 8), 256, 0, m->stream>>>(table, bias, ids, n_ids, ncols, out_rows, out_bias);
  KERNEL_CHECK(m);
  return 0;
}

int launch_scatter_table_rows(sbr_model* m, const float* rows, const float* brow, const int32_t* ids, int n_ids,
                              int ncols, float* table_grad, float* bias_grad) {
  if (n_ids == 0) return 0;
  scatter_table_rows_kernel<<<cdiv(n_ids, 8), 256, 0, m->stream>>>(rows, brow, ids, n_ids, ncols, table_grad, bias_grad);
  KERNEL_CHECK(m);
  return 0;
}

int launch_transpose(sbr_model* m, const float* in, int rows, int cols, int ld_in, float* out) {
  transpose_kernel<<<dim3(cdiv(cols, 32), cdiv(rows, 32)), dim3(32, 8), 0, m->stream>>>(in, rows, cols, ld_in, out);
  KERNEL_CHECK(m);
  return 0;
}

int launch_embed_gather(sbr_model* m, const int32_t* X, const int32_t* len, const float* table, float* out,
                        int B, int T, int K, int E, int t_max) {
  const int n_rows = t_max * B;
  if (n_rows == 0) return 0;
  embed_gather_kernel<<<cdiv(n_rows, 8), 256, 0, m->stream>>>(X, len, table, out, B, T, K, E, n_rows);
  KERNEL_CHECK(m);
  return 0;
}

int launch_embed_scatter(sbr_model* m, const int32_t* X, const int32_t* len, const float* dOut, float* dTable,
                         int B, int T, int K, int E, int t_max) {
  const int n_rows = t_max * B;
  if (n_rows == 0) return 0;
  embed_scatter_kernel<<<cdiv(n_rows, 8), 256, 0, m->stream>>>(X, len, dOut, dTable, B, T, K, E, n_rows);
  KERNEL_CHECK(m);
  return 0;
}


// ------------------------------------------------------------------------------------------------
// Bidirectional stacks (recurrent_layers.py:72-78).  A backwards layer over left-aligned rows equals a forward layer
// over the rows with their valid prefix reversed, and its output un-reversed the same way is the aligned output
// (DESIGN.md section 3.5 states the identity; the CPU tests check it), so the backwards layers reuse
// the forward-only scan kernels; these kernels move data between the two coordinate systems.  Time-major rows
// (row = t*B + b); position t of row b maps to len_b - 1 - t.
// ------------------------------------------------------------------------------------------------
namespace {
__global__ void reverse_ids_kernel(const int32_t* __restrict__ X, const int32_t* __restrict__ len, int32_t* __restrict__ Xr,
                                   int B, int T, int K) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)B * T * K) return;
  const int k = (int)(i % K), t = (int)((i / K) % T), b = (int)(i / ((int64_t)K * T));
  const int l = min(len[b], T);
  Xr[i] = t < l ? X[((int64_t)b * T + (l - 1 - t)) * K + k] : 0;
}
// cat_al[(t,b)] = [hs_f(t,b) | hs_b(l-1-t,b)],  cat_rv[(t,b)] = [hs_f(l-1-t,b) | hs_b(t,b)];  zero beyond the row's length
__global__ void bi_concat_kernel(const float* __restrict__ hs_f, const float* __restrict__ hs_b, const int32_t* __restrict__ len,
                                 float* __restrict__ cat_al, float* __restrict__ cat_rv, int B, int t_max, int H) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)t_max * B * H) return;
  const int k = (int)(i % H);
  const int64_t row = i / H;
  const int b = (int)(row % B), t = (int)(row / B);
  const int l = min(len[b], t_max);
  float f_al = 0.f, b_al = 0.f, f_rv = 0.f, b_rv = 0.f;
  if (t < l) {
    const int64_t mir = (int64_t)(l - 1 - t) * B + b;
    f_al = hs_f[row * H + k]; b_rv = hs_b[row * H + k];
    f_rv = hs_f[mir * H + k]; b_al = hs_b[mir * H + k];
  }
  cat_al[row * 2 * H + k] = f_al; cat_al[row * 2 * H + H + k] = b_al;
  cat_rv[row * 2 * H + k] = f_rv; cat_rv[row * 2 * H + H + k] = b_rv;
}
// dhs_f(t,b) = dcat_al(t,b)[0:H] + dcat_rv(l-1-t,b)[0:H];  dhs_b(t,b) = dcat_al(l-1-t,b)[H:2H] + dcat_rv(t,b)[H:2H]
__global__ void bi_split_kernel(const float* __restrict__ dcat_al, const float* __restrict__ dcat_rv, const int32_t* __restrict__ len,
                                float* __restrict__ dhs_f, float* __restrict__ dhs_b, int B, int t_max, int H) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)t_max * B * H) return;
  const int k = (int)(i % H);
  const int64_t row = i / H;
  const int b = (int)(row % B), t = (int)(row / B);
  const int l = min(len[b], t_max);
  float df = 0.f, db = 0.f;
  if (t < l) {
    const int64_t mir = (int64_t)(l - 1 - t) * B + b;
    df = dcat_al[row * 2 * H + k] + dcat_rv[mir * 2 * H + k];
    db = dcat_al[mir * 2 * H + H + k] + dcat_rv[row * 2 * H + H + k];
  }
  dhs_f[row * H + k] = df;
  dhs_b[row * H + k] = db;
}
}  // namespace

int launch_reverse_ids(sbr_model* m, const int32_t* X, const int32_t* len, int32_t* X_rev, int B, int T, int K) {
  const int64_t n = (int64_t)B * T * K;
  if (n == 0) return 0;
  reverse_ids_kernel<<<cdiv(n, 256), 256, 0, m->stream>>>(X, len, X_rev, B, T, K);
  KERNEL_CHECK(m);
  return 0;
}
int launch_bi_concat(sbr_model* m, const float* hs_f, const float* hs_b, const int32_t* len, float* cat_al, float* cat_rv,
                     int B, int t_max, int H) {
  const int64_t n = (int64_t)t_max * B * H;
  if (n == 0) return 0;
  bi_concat_kernel<<<cdiv(n, 256), 256, 0, m->stream>>>(hs_f, hs_b, len, cat_al, cat_rv, B, t_max, H);
  KERNEL_CHECK(m);
  return 0;
}
int launch_bi_split(sbr_model* m, const float* dcat_al, const float* dcat_rv, const int32_t* len, float* dhs_f, float* dhs_b,
                    int B, int t_max, int H) {
  const int64_t n = (int64_t)t_max * B * H;
  if (n == 0) return 0;
  bi_split_kernel<<<cdiv(n, 256), 256, 0, m->stream>>>(dcat_al, dcat_rv, len, dhs_f, dhs_b, B, t_max, H);
  KERNEL_CHECK(m);
  return 0;
}
