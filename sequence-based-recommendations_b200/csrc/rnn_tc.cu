// rnn_tc.cu -- stage 2 on the 5th-generation tensor cores: the recurrent scan (and its BPTT) with the
// per-step gate GEMM on tcgen05.mma, fp32-accurate through the 3xTF32 split.
//
// Same reference semantics as rnn_cluster.cu (sparse_lstm.py:377-425, :764-805, :1120-1152) and the
// same ownership: a cluster of C CTAs owns a tile of 16 batch rows for all T steps; CTA r owns the
// hidden units [r*Hs, (r+1)*Hs) of every gate.  What changes is how a step is computed:
//
//   * the CTA's slice of W_hid is the **A operand, resident in TMEM** for the whole scan
//     (tcgen05.st once): row m = 4*j + g (unit j, gate g) of a 128-row tile, K along the columns,
//     stored twice -- hi = top 19 bits (what kind::tf32 reads), lo = w - hi;
//   * h_{t-1} (forward) / da_t (backward) is the **B operand in shared memory** [k/4][16 rows][4]
//     (K-major, no swizzle: 8x16B core matrices, SBO 128 B, LBO 256 B), also as hi/lo;
//   * D[128 x 16] += A_hi*B_hi + A_hi*B_lo + A_lo*B_hi, fp32 accumulators in TMEM, issued by one thread;
//     tcgen05.commit -> mbarrier -> tcgen05.ld (each thread = one gate row, 16 batch columns);
//   * the fused gate math runs on (row, 4 consecutive units) per thread, so every state exchange
//     is 16-byte: h_t quads are stored straight into the shared memory of all CTAs of the cluster
//     (DSMEM, fp32) and announced with one remote mbarrier.arrive(release.cluster) per warp --
//     no cluster-wide barrier in the loop; the receiver splits hi/lo locally.
//
// The backward keeps W_hid^T-style tiles in TMEM (rows = hidden index k, columns = own gate columns):
// dh_{t-1}[b][k] partial = sum over OWN gate columns of da[b][gj] W_hid[k][gj] (split-K), reduce-
// scattered to the owners through DSMEM like rnn_cluster.cu.
//
// Applicability: H % 4 == 0, H <= 240 (hi+lo copies of the K extent must fit in 512 TMEM columns),
// otherwise launch_rnn_* falls back to the FFMA cluster kernels.
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int TC_NT = 128;   // 4 warps: TMEM lane quadrants 0..3
constexpr int TC_BT = 16;    // batch rows per cluster tile = MMA N
constexpr int GSM_LD = 132;  // padded row of the gate staging buffer

struct TcArgs {
  const float* Xg; const float* W_hid; const float* W_hidT; const float* peep; const float* h_init; const float* c_init;
  const int32_t* len;
  float* hs; float* cs; float* act; float* h_last;
  const float* dh_last; const float* dhs; float* dXg; float* dac; float* g_peep; float* g_h_init; float* g_c_init;
  float clip;
  int B, H, Hs, Kp, t_max;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// top 19 bits of v, rounded to nearest: exactly what kind::tf32 reads; lo = v - hi is exact in fp32
__device__ __forceinline__ float tf32_hi(float v) { return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xFFFFE000u); }
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }
__device__ __forceinline__ float clipf_(float x, float c) { return c > 0.f ? fminf(fmaxf(x, -c), c) : x; }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version 1 (sm_100); layout SWIZZLE_NONE
  return d;
}
// kind::tf32, fp32 accumulate, K-major A and B
__device__ __forceinline__ uint32_t make_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
               :: "r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait_cta(uint64_t* bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
               "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
               "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}\n" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
// wait with cluster-scope acquire: pairs with the remote arrive.release.cluster of the producers
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
               "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
               "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}\n" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" :: "r"(cluster_addr) : "memory");
}
__device__ __forceinline__ uint32_t map_to_rank(uint32_t cta_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(cta_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t addr, float4 v) {
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" :: "r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
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
__device__ __forceinline__ void tmem_st8(uint32_t addr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               :: "r"(addr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
#define TC_FENCE_BEFORE() asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory")
#define TC_FENCE_AFTER() asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory")
#define PROXY_FENCE_SMEM() asm volatile("fence.proxy.async.shared::cta;" ::: "memory")

// B-operand layout of a [16 rows] x [Kp] matrix: float index of element (row b, k)
__device__ __forceinline__ int bidx(int b, int k) { return (k >> 2) * (TC_BT * 4) + b * 4 + (k & 3); }

// 3xTF32: D1[128x16] = A_hi*B_hi ; D2[128x16] = A_hi*B_lo + A_lo*B_hi over KS k-chunks of 8.  The
// accumulator add of the tensor core truncates, so the large term and the 2^-11-times-smaller
// correction terms are kept in separate TMEM accumulators (D2 = D1 + 16 columns) and summed in fp32
// by the epilogue: the chain into the large accumulator is KS adds instead of 3*KS.
__device__ __forceinline__ void issue_3xtf32(uint32_t tD, uint32_t tAhi, uint32_t tAlo, const float* Bhi, const float* Blo,
                                             int KS, uint32_t idesc) {
  const uint32_t bhi = smem_u32(Bhi), blo = smem_u32(Blo);
  uint32_t acc = 0;
  for (int ks = 0; ks < KS; ++ks) {
    const uint64_t dhi = make_desc(bhi + ks * 2 * (TC_BT * 16), TC_BT * 16, 128);
    const uint64_t dlo = make_desc(blo + ks * 2 * (TC_BT * 16), TC_BT * 16, 128);
    mma_ts(tD, tAhi + ks * 8, dhi, idesc, acc);
    mma_ts(tD + TC_BT, tAhi + ks * 8, dlo, idesc, acc);
    mma_ts(tD + TC_BT, tAlo + ks * 8, dhi, idesc, 1);
    acc = 1;
  }
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
template <int G>
__global__ void __launch_bounds__(TC_NT, 1) rnn_fwd_tc_kernel(const TcArgs a) {
  cg::cluster_group cluster = cg::this_cluster();
  const int C = cluster.num_blocks();
  const int rank = cluster.block_rank();
  const int tile = blockIdx.x / C;
  const int b0 = tile * TC_BT;
  const int H = a.H, Hs = a.Hs, Kp = a.Kp, GH = G * H, B = a.B;
  const int KS = Kp / 8;
  const int j0 = rank * Hs;
  const int nj = max(0, min(Hs, H - j0));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  extern __shared__ __align__(128) float smem[];
  float* hraw = smem;                            // [2][Kp/4][16][4] fp32 h_{t-1}, written by every CTA of the cluster
  float* Bhi = hraw + 2 * Kp * TC_BT;            // [Kp/4][16][4]
  float* Blo = Bhi + Kp * TC_BT;
  float* gsm = Blo + Kp * TC_BT;                 // [16][GSM_LD] gate pre-activations, row b, column m = 4j+g
  __shared__ __align__(8) uint64_t raw_full[2];
  __shared__ __align__(8) uint64_t mma_done;
  __shared__ uint32_t tmem_base_s;
  __shared__ int lens_s[TC_BT];
  __shared__ int t_end_s;

  if (tid < TC_BT) lens_s[tid] = (b0 + tid < B) ? min(a.len[b0 + tid], a.t_max) : 0;
  if (tid == 0) {
    mbar_init(&raw_full[0], C * (TC_NT / 32));
    mbar_init(&raw_full[1], C * (TC_NT / 32));
    mbar_init(&mma_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // h_{-1} = learned init, broadcast over the rows: straight into the split B operand
  for (int i = tid; i < Kp * TC_BT; i += TC_NT) {
    const int kc = i / (TC_BT * 4), rem = i - kc * (TC_BT * 4);
    const int k = kc * 4 + (rem & 3);
    const float v = k < H ? a.h_init[k] : 0.f;
    const float hi = tf32_hi(v);
    Bhi[i] = hi;
    Blo[i] = v - hi;
    hraw[i] = v;               // buffer 0 doubles as "h_prev" of step 0
    hraw[Kp * TC_BT + i] = 0.f;
  }
  TC_FENCE_BEFORE();
  __syncthreads();
  TC_FENCE_AFTER();
  if (tid == 0) {
    int mx = 0;
    for (int b = 0; b < TC_BT; ++b) mx = max(mx, lens_s[b]);
    t_end_s = mx;
  }
  const uint32_t tmem = tmem_base_s;
  const uint32_t tD = tmem;                 // 16 columns
  const uint32_t tAhi = tmem + 32;          // Kp columns
  const uint32_t tAlo = tmem + 32 + Kp;     // Kp columns (32 + 2*Kp <= 512)

  // ---- A operand: row m = tid = 4*j + g  <->  W_hid[:, g*H + j0 + j]; K along the TMEM columns
  {
    const int j = tid >> 2, g = tid & 3;
    const bool live = (g < G) && (j < nj);
    const float* src = a.W_hidT + (int64_t)(g * H + j0 + j) * H;   // k contiguous
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
    for (int k0 = 0; k0 < Kp; k0 += 8) {
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float v = (live && k0 + i < H) ? __ldg(src + k0 + i) : 0.f;
        const float h = tf32_hi(v);
        hi[i] = __float_as_uint(h);
        lo[i] = __float_as_uint(v - h);
      }
      tmem_st8(tAhi + lane_off + k0, hi);
      tmem_st8(tAlo + lane_off + k0, lo);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }

  // ---- per-thread ownership for the gate math: row b = tid/8, units 4*jq .. 4*jq+3 of the slice
  const int eb = tid >> 3, jq = tid & 7;
  const int ju = 4 * jq;
  const bool own = ju < nj;                       // nj is a multiple of 4 (H % 4 == 0, Hs % 4 == 0)
  const bool row_ok = b0 + eb < B;
  float cst[4] = {0.f, 0.f, 0.f, 0.f};
  float wci[4], wcf[4], wco[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    wci[u] = wcf[u] = wco[u] = 0.f;
    if (G == 4 && own) {
      wci[u] = a.peep[j0 + ju + u];
      wcf[u] = a.peep[H + j0 + ju + u];
      wco[u] = a.peep[2 * H + j0 + ju + u];
      cst[u] = a.c_init[j0 + ju + u];
    }
  }
  if (own && row_ok) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      a.hs[(int64_t)(b0 + eb) * H + j0 + ju + u] = a.h_init[j0 + ju + u];
      if (G == 4) a.cs[(int64_t)(b0 + eb) * H + j0 + ju + u] = a.c_init[j0 + ju + u];
    }
  }
  float xc[G][4], xn[G][4];
#pragma unroll
  for (int g = 0; g < G; ++g)
#pragma unroll
    for (int u = 0; u < 4; ++u) xc[g][u] = xn[g][u] = 0.f;
  auto load_x = [&](int t, float (&x)[G][4]) {
    if (own && t < lens_s[eb]) {
      const float* src = a.Xg + ((int64_t)t * B + b0 + eb) * GH + j0 + ju;
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(src + g * H));
        x[g][0] = v.x; x[g][1] = v.y; x[g][2] = v.z; x[g][3] = v.w;
      }
    }
  };
  PROXY_FENCE_SMEM();       // Bhi/Blo were written through the generic proxy
  TC_FENCE_BEFORE();
  __syncthreads();
  TC_FENCE_AFTER();
  const int t_end = t_end_s;
  if (t_end > 0) load_x(0, xc);
  cluster.sync();           // barriers initialised and buffers zeroed everywhere before remote traffic

  const uint32_t idesc = make_idesc_tf32(128, TC_BT);
  const uint32_t hraw_addr = smem_u32(hraw);
  const uint32_t bar_addr[2] = {smem_u32(&raw_full[0]), smem_u32(&raw_full[1])};

  for (int t = 0; t < t_end; ++t) {
    const int cur = t & 1, nxt = cur ^ 1;
    if (t + 1 < t_end) load_x(t + 1, xn);
    const float* hprev = hraw + cur * Kp * TC_BT;
    if (t > 0) {
      // h_{t-1} has landed from every CTA of the cluster; split it into the hi/lo B operand
      const int use = (t - (cur == 0 ? 2 : 1)) >> 1;
      mbar_wait_cluster(&raw_full[cur], use & 1);
      const float4* src = reinterpret_cast<const float4*>(hprev);
      float4* dhi = reinterpret_cast<float4*>(Bhi);
      float4* dlo = reinterpret_cast<float4*>(Blo);
      for (int i = tid; i < Kp * TC_BT / 4; i += TC_NT) {
        const float4 v = src[i];
        float4 h, l;
        h.x = tf32_hi(v.x); h.y = tf32_hi(v.y); h.z = tf32_hi(v.z); h.w = tf32_hi(v.w);
        l.x = v.x - h.x; l.y = v.y - h.y; l.z = v.z - h.z; l.w = v.w - h.w;
        dhi[i] = h;
        dlo[i] = l;
      }
      PROXY_FENCE_SMEM();
      TC_FENCE_BEFORE();
      __syncthreads();
      TC_FENCE_AFTER();
    }
    if (tid == 0) {
      issue_3xtf32(tD, tAhi, tAlo, Bhi, Blo, KS, idesc);
      umma_commit(&mma_done);
    }
    mbar_wait_cta(&mma_done, t & 1);
    TC_FENCE_AFTER();
    {
      float v[16], w[16];
      tmem_ld16(tD + ((uint32_t)(warp * 32) << 16), v);
      tmem_ld16(tD + TC_BT + ((uint32_t)(warp * 32) << 16), w);
#pragma unroll
      for (int b = 0; b < 16; ++b) gsm[b * GSM_LD + tid] = v[b] + w[b];
    }
    TC_FENCE_BEFORE();
    __syncthreads();

    if (own) {
      const bool active = t < lens_s[eb];
      const float4 hp4 = *reinterpret_cast<const float4*>(hprev + bidx(eb, j0 + ju));
      const float hp[4] = {hp4.x, hp4.y, hp4.z, hp4.w};
      float hn[4];
      float sv[4][4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float4 p4 = *reinterpret_cast<const float4*>(gsm + eb * GSM_LD + 4 * (ju + u));
        const float pre[4] = {p4.x, p4.y, p4.z, p4.w};
        float xg[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int g = 0; g < G; ++g) xg[g] = xc[g][u];
        hn[u] = hp[u];
        sv[u][0] = sv[u][1] = sv[u][2] = sv[u][3] = 0.f;
        if (active) {
          if constexpr (G == 4) {
            const float c_prev = cst[u];
            const float ig = sigmoidf_(xg[0] + pre[0] + c_prev * wci[u]);
            const float fg = sigmoidf_(xg[1] + pre[1] + c_prev * wcf[u]);
            const float gg = tanhf(xg[2] + pre[2]);
            const float c_new = fg * c_prev + ig * gg;
            const float og = sigmoidf_(xg[3] + pre[3] + c_new * wco[u]);
            hn[u] = og * tanhf(c_new);
            cst[u] = c_new;
            sv[u][0] = ig; sv[u][1] = fg; sv[u][2] = gg; sv[u][3] = og;
          } else if constexpr (G == 3) {
            const float r = sigmoidf_(pre[0] + xg[0]);
            const float uu = sigmoidf_(pre[1] + xg[1]);
            const float ac = pre[2];
            const float cand = tanhf(xg[2] + r * ac);
            hn[u] = (1.f - uu) * hp[u] + uu * cand;
            sv[u][0] = r; sv[u][1] = uu; sv[u][2] = cand; sv[u][3] = ac;
          } else {
            hn[u] = tanhf(xg[0] + pre[0]);
          }
        }
      }
      // publish h_t[b, j0+ju .. +3] (fp32, one 16-byte store) into every CTA of the cluster
      const float4 hv = make_float4(hn[0], hn[1], hn[2], hn[3]);
      const uint32_t off = hraw_addr + (uint32_t)(nxt * Kp * TC_BT + bidx(eb, j0 + ju)) * 4u;
      for (int rr = 0; rr < C; ++rr) st_cluster_v4(map_to_rank(off, rr), hv);
      if (row_ok) {
        const int64_t row1 = (int64_t)(t + 1) * B + b0 + eb;
        *reinterpret_cast<float4*>(a.hs + row1 * H + j0 + ju) = hv;
        if (G == 4) *reinterpret_cast<float4*>(a.cs + row1 * H + j0 + ju) = make_float4(cst[0], cst[1], cst[2], cst[3]);
        if (active && G > 1) {
          float* ap = a.act + ((int64_t)t * B + b0 + eb) * 4 * H + j0 + ju;
#pragma unroll
          for (int g = 0; g < 4; ++g)
            *reinterpret_cast<float4*>(ap + g * H) = make_float4(sv[0][g], sv[1][g], sv[2][g], sv[3][g]);
        }
      }
    }
#pragma unroll
    for (int g = 0; g < G; ++g)
#pragma unroll
      for (int u = 0; u < 4; ++u) xc[g][u] = xn[g][u];
    // announce: one release-arrive per warp on the next buffer's barrier of every CTA
    __syncwarp();
    if (lane == 0)
      for (int rr = 0; rr < C; ++rr) mbar_arrive_remote(map_to_rank(bar_addr[nxt], rr));
  }

  // final state: wait for the last exchange so every CTA can read the full h (only own slice is written)
  if (t_end > 0) {
    const int fin = t_end & 1;
    const int use = (t_end - (fin == 0 ? 2 : 1)) >> 1;
    mbar_wait_cluster(&raw_full[fin], use & 1);
  }
  if (a.h_last && own && row_ok) {
    const int fin = t_end & 1;
    const float4 hv = *reinterpret_cast<const float4*>(hraw + fin * Kp * TC_BT + bidx(eb, j0 + ju));
    *reinterpret_cast<float4*>(a.h_last + (int64_t)(b0 + eb) * H + j0 + ju) = hv;
  }
  TC_FENCE_BEFORE();
  cluster.sync();   // nobody exits while peers may still write to / arrive on this CTA's shared memory
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512));
}

// ------------------------------------------------------------------------------------------------
// launch plumbing
// ------------------------------------------------------------------------------------------------
struct TcPlan { int C, Hs, Kp; size_t smem; bool ok; };

TcPlan tc_plan(int G, int H) {
  TcPlan p{};
  p.ok = false;
  if (getenv("SBR_DISABLE_TC")) return p;
  if (H % 4 != 0 || H < 8) return p;
  p.Kp = (int)round_up(H, 8);
  if (32 + 2 * p.Kp > 512) return p;
  for (int C = 8; C >= 1; C >>= 1) {
    const int Hs = (int)round_up(cdiv(H, C), 4);
    if (Hs > 32) break;
    if (Hs * (C - 1) < H) { p.C = C; p.Hs = Hs; p.ok = true; break; }
  }
  if (!p.ok) return p;
  size_t f = (size_t)4 * p.Kp * TC_BT + (size_t)TC_BT * GSM_LD;
  p.smem = std::max<size_t>(f * sizeof(float), 120 * 1024);   // > half an SM: one CTA (one TMEM allocation) per SM
  return p;
}

template <typename Kern>
int launch_tc(sbr_model* m, Kern kern, const TcPlan& p, int n_tiles, const TcArgs& args) {
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);
  if (e != cudaSuccess) { sbr_set_error(m, SBR_E_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return SBR_E_CUDA; }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(p.C * n_tiles, 1, 1);
  cfg.blockDim = dim3(TC_NT, 1, 1);
  cfg.dynamicSmemBytes = p.smem;
  cfg.stream = m->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = p.C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, kern, args);
  if (e != cudaSuccess) { sbr_set_error(m, SBR_E_CUDA, "tcgen05 scan launch (C=%d) failed: %s", p.C, cudaGetErrorString(e)); return SBR_E_CUDA; }
  m->launches++;
  return 0;
}

}  // namespace

// returns 1 when the tensor-core path does not apply (caller falls back), 0 on success, <0 on error
int launch_rnn_forward_tc(sbr_model* m, const LayerDesc& L, const int32_t* len, int B, int t_max, float* h_last) {
  const TcPlan p = tc_plan(L.G, L.H);
  if (!p.ok) return 1;
  int rc = launch_transpose(m, m->params + L.W_hid, L.H, L.G * L.H, L.G * L.H, m->WhidT);
  if (rc) return rc;
  TcArgs a{};
  a.Xg = L.Xg; a.W_hid = m->params + L.W_hid; a.W_hidT = m->WhidT;
  a.peep = m->params + L.peep; a.h_init = m->params + L.h_init; a.c_init = m->params + L.c_init;
  a.len = len; a.hs = L.hs; a.cs = L.cs; a.act = L.act; a.h_last = h_last;
  a.clip = m->cfg.grad_clip; a.B = B; a.H = L.H; a.Hs = p.Hs; a.Kp = p.Kp; a.t_max = t_max;
  const int n_tiles = cdiv(B, TC_BT);
  if (L.G == 4) return launch_tc(m, rnn_fwd_tc_kernel<4>, p, n_tiles, a);
  if (L.G == 3) return launch_tc(m, rnn_fwd_tc_kernel<3>, p, n_tiles, a);
  return launch_tc(m, rnn_fwd_tc_kernel<1>, p, n_tiles, a);
}
