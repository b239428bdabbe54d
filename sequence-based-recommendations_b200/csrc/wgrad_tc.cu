// wgrad_tc.cu -- the BPTT weight gradient on the 5th-generation tensor cores:
//
//     dW_hid[h, c] += sum over rows (t, b) of  h_{t-1}[row, h] * da_t[row, c]          (sparse_lstm.py:383,766 via
//                                                                                        theano.grad, rnn_base.py:183)
//
// a [H x G*H] output with a K = T*B contraction.  kind::tf32 only accepts K-major operands (MN-major descriptors
// yield zero accumulators, probes/tc_probe3.cu), so the scan kernels (rnn_tc.cu) emit both operands already
// K-major, already split for 3xTF32 (hi | lo), and already tiled the way the MMA wants them:
//
//     hT[part][h/128][row/4][h%128][row%4]      aT[part][col/128][row/4][col%128][row%4]
//
// i.e. for one 128-wide tile, 8 consecutive row-quads (32 rows of K) are ONE contiguous 16 KB block in the
// canonical no-swizzle K-major layout (8x16 B core matrices, SBO 128 B, LBO 2 KB).  A stage is therefore four
// cp.async.bulk copies (A_hi, A_lo, B_hi, B_lo) signalled on an mbarrier -- no tensor map, no transposition.
//
// CTA = one 128 x 128 output tile x one K split.  Warp 4 lane 0 produces (bulk copies, 3-stage ring), warp 5 lane 0
// issues the MMAs (D1 = A_hi B_hi in TMEM columns 0..127, D2 = A_hi B_lo + A_lo B_hi in 128..255) and releases a
// stage with tcgen05.commit; warps 0-3 drain TMEM and add the tile into the gradient arena (fp32 RED).
#include "common.cuh"

namespace {

constexpr int WG_STAGES = 3;
constexpr int WG_QPS = 8;                        // row-quads (4 rows each) per stage = 32 rows of K
constexpr int WG_PART_BYTES = WG_QPS * 128 * 16;  // 16 KB per operand part per stage
constexpr int WG_NT = 192;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
               :: "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
               "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
               "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}\n" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint32_t elect_one_sync() {
  uint32_t pred = 0, laneid = 0;
  asm volatile("{\n.reg .b32 %%rx;\n.reg .pred %%px;\n     elect.sync %%rx|%%px, %2;\n@%%px mov.s32 %1, 1;\n     mov.s32 %0, %%rx;\n}\n"
               : "+r"(laneid), "+r"(pred) : "r"(0xFFFFFFFF));
  return pred;
}
__device__ __forceinline__ void tmem_ld16(uint32_t addr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(addr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

struct WgArgs {
  const float* hT; const float* aT;
  long long hT_part, hT_tile, aT_part, aT_tile;
  float* dW;
  int ldw, H, GH;
  int rq_total;      // row-quads of K (rows / 4), even
  const int* list;   // list[0] = n, list[1..n] = the stages (32 consecutive rows of K) that hold at least one valid row
};

// Stages of the contraction worth visiting.  The scans write exact zeros into aT for masked (t, b) rows, so a stage
// whose 32 rows are all beyond their sequences contributes nothing: with the nested-prefix batches of the reference
// (mean length ~ a third of max_length) that is most of K.  Order does not matter for a sum.
__global__ void wgrad_stage_list_kernel(const int32_t* __restrict__ len, int B, int rows, int* __restrict__ list) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s * 32 >= rows) return;
  bool any = false;
  for (int r = s * 32; r < min(rows, s * 32 + 32) && !any; ++r) {
    const int t = r / B, b = r - t * B;
    any = t < len[b];
  }
  if (any) list[1 + atomicAdd(list, 1)] = s;
}

__global__ void __launch_bounds__(WG_NT, 1) wgrad_tc_kernel(const WgArgs a) {
  extern __shared__ __align__(1024) uint8_t wg_smem[];
  // stage s: [A_hi | A_lo | B_hi | B_lo], 16 KB each
  __shared__ __align__(8) uint64_t full[WG_STAGES];
  __shared__ __align__(8) uint64_t empty[WG_STAGES];
  __shared__ __align__(8) uint64_t done;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nt = blockIdx.x, mt = blockIdx.y;
  // this CTA's share of the stage list
  const int n_list = a.list[0];
  const int per = (n_list + (int)gridDim.z - 1) / (int)gridDim.z;
  const int i0 = min(n_list, (int)blockIdx.z * per);
  const int n_stages = min(n_list, i0 + per) - i0;
  const int* stages = a.list + 1 + i0;

  if (tid == 0) {
    for (int s = 0; s < WG_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(&done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;

  if (n_stages > 0) {
    if (warp == 4) {
      // ---- producer: four bulk copies per stage
      if (elect_one_sync()) {
        const float* gA[2] = {a.hT + mt * a.hT_tile, a.hT + a.hT_part + mt * a.hT_tile};
        const float* gB[2] = {a.aT + nt * a.aT_tile, a.aT + a.aT_part + nt * a.aT_tile};
        for (int i = 0; i < n_stages; ++i) {
          const int s = i % WG_STAGES;
          if (i >= WG_STAGES) mbar_wait(&empty[s], ((i / WG_STAGES) - 1) & 1);
          const int rq = __ldg(stages + i) * WG_QPS;
          const int nq = min(WG_QPS, a.rq_total - rq);
          const uint32_t bytes = (uint32_t)nq * 128 * 16;
          uint8_t* st = wg_smem + (size_t)s * 4 * WG_PART_BYTES;
          mbar_arrive_expect_tx(&full[s], 4 * bytes);
          const long long off = (long long)rq * 512;   // floats per row-quad inside a tile
          bulk_load(st, gA[0] + off, bytes, &full[s]);
          bulk_load(st + WG_PART_BYTES, gA[1] + off, bytes, &full[s]);
          bulk_load(st + 2 * WG_PART_BYTES, gB[0] + off, bytes, &full[s]);
          bulk_load(st + 3 * WG_PART_BYTES, gB[1] + off, bytes, &full[s]);
        }
      }
      __syncwarp();
    } else if (warp == 5) {
      // ---- MMA issuer
      if (elect_one_sync()) {
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        uint32_t acc = 0;
        for (int i = 0; i < n_stages; ++i) {
          const int s = i % WG_STAGES;
          mbar_wait(&full[s], (i / WG_STAGES) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const int rq = __ldg(stages + i) * WG_QPS;
          const int nq = min(WG_QPS, a.rq_total - rq);
          const uint32_t base = smem_u32(wg_smem + (size_t)s * 4 * WG_PART_BYTES);
          for (int kc = 0; kc < nq / 2; ++kc) {     // one MMA k-chunk = 8 rows = 2 row-quads = 4 KB
            const uint32_t o = (uint32_t)kc * 4096;
            const uint64_t ahi = make_desc(base + o, 2048, 128);
            const uint64_t alo = make_desc(base + WG_PART_BYTES + o, 2048, 128);
            const uint64_t bhi = make_desc(base + 2 * WG_PART_BYTES + o, 2048, 128);
            const uint64_t blo = make_desc(base + 3 * WG_PART_BYTES + o, 2048, 128);
            mma_ss(tmem, ahi, bhi, idesc, acc);
            mma_ss(tmem + 128, ahi, blo, idesc, acc);
            mma_ss(tmem + 128, alo, bhi, idesc, 1);
            acc = 1;
          }
          umma_commit(&empty[s]);                   // the stage is free once these MMAs have read it
        }
        umma_commit(&done);
      }
      __syncwarp();
    }
  }

  // ---- epilogue: warps 0-3 drain their TMEM lane quadrant and add the tile into the gradient arena
  if (warp < 4 && n_stages > 0) {
    mbar_wait(&done, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int h = mt * 128 + warp * 32 + lane;
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
    for (int c0 = 0; c0 < 128; c0 += 16) {
      float v[16], w[16];
      tmem_ld16(tmem + lane_off + c0, v);
      tmem_ld16(tmem + 128 + lane_off + c0, w);
      if (h < a.H) {
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          const int col = nt * 128 + c0 + i;
          float* dst = a.dW + (long long)h * a.ldw + col;
          if (col + 3 < a.GH && (a.ldw & 3) == 0) {
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};"
                         :: "l"(dst), "f"(v[i] + w[i]), "f"(v[i + 1] + w[i + 1]), "f"(v[i + 2] + w[i + 2]), "f"(v[i + 3] + w[i + 3]) : "memory");
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (col + e < a.GH) atomicAdd(dst + e, v[i + e] + w[i + e]);
          }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(256));
}

}  // namespace

int launch_wgrad_tc(sbr_model* m, const LayerDesc& L, int rows, float* dW, int ldw) {
  if (rows <= 0) return 0;
  if (rows % 8 != 0 || !L.hT || !L.aT) {
    sbr_set_error(m, SBR_E_ARG, "wgrad_tc: rows must be a multiple of 8 and the K-major copies must exist");
    return SBR_E_ARG;
  }
  WgArgs a{};
  a.hT = L.hT; a.aT = L.aT;
  a.hT_part = L.hT_part; a.hT_tile = L.hT_tile; a.aT_part = L.aT_part; a.aT_tile = L.aT_tile;
  a.dW = dW; a.ldw = ldw; a.H = L.H; a.GH = L.G * L.H;
  a.rq_total = rows / 4;
  if (!m->wg_list || m->wg_list_rows != rows) {
    sbr_set_error(m, SBR_E_ARG, "wgrad_tc: the stage list of this batch is missing");
    return SBR_E_ARG;
  }
  a.list = m->wg_list;
  const int mts = cdiv(L.H, 128), nts = cdiv(L.G * L.H, 128);
  const int splits = std::max(1, std::min(m->n_sm / (mts * nts), cdiv(rows, 32)));
  const size_t smem = (size_t)WG_STAGES * 4 * WG_PART_BYTES + 1024;
  static std::vector<int> attr_devs;      // the opt-in shared-memory limit is a per-device attribute
  if (std::find(attr_devs.begin(), attr_devs.end(), m->dev) == attr_devs.end()) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { sbr_set_error(m, SBR_E_CUDA, "wgrad_tc attr: %s", cudaGetErrorString(e)); return SBR_E_CUDA; }
    attr_devs.push_back(m->dev);
  }
  wgrad_tc_kernel<<<dim3(nts, mts, splits), WG_NT, smem, m->stream>>>(a);
  KERNEL_CHECK(m);
  return 0;
}

// once per batch (every layer of the stack shares the lengths): which 32-row stages of K = t_max * B hold valid rows
int launch_wgrad_stage_list(sbr_model* m, const int32_t* len, int B, int rows) {
  m->wg_list_rows = 0;
  if (!m->wg_list || rows <= 0) return 0;
  CU_TRY(m, cudaMemsetAsync(m->wg_list, 0, sizeof(int), m->stream));
  const int n = cdiv(rows, 32);
  wgrad_stage_list_kernel<<<cdiv(n, 128), 128, 0, m->stream>>>(len, B, rows, m->wg_list);
  KERNEL_CHECK(m);
  m->wg_list_rows = rows;
  return 0;
}
