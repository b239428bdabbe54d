// optim.cu -- fused dense parameter update over the flat arena.
//
// Reference: neural_networks/update_manager.py:24-82 -> lasagne.updates.{adam, adagrad, adadelta,
// rmsprop, nesterov_momentum} (Lasagne master; restated in SURVEY.md Appendix A.5).  Lasagne
// updates EVERY element of every parameter on every step -- including embedding rows whose
// gradient is zero, because the Adam moments keep decaying -- so the update is one dense,
// HBM-bound streaming pass: 16 B read (p, g, m, v) + 12 B written (p, m, v) per parameter, plus
// the 4 B that re-zero the gradient arena for the next step (fused here instead of a memset).
// One launch, 128-bit accesses, grid = a multiple of 148 SMs.
#include "common.cuh"

namespace {

struct OptArgs {
  float* p;
  float* g;
  float* a;
  float* b;
  int64_t n;
  int kind;
  float lr, rho, b1, b2, a_t;   // a_t: Adam's bias-corrected step size for this t
  int zero_grad;
};

__device__ __forceinline__ void update_one(const OptArgs& o, float& p, float g, float& a, float& b) {
  switch (o.kind) {
    case SBR_UPD_ADAM: {
      a = o.b1 * a + (1.f - o.b1) * g;
      b = o.b2 * b + (1.f - o.b2) * g * g;
      p -= o.a_t * a / (sqrtf(b) + 1e-8f);
      break;
    }
    case SBR_UPD_ADAGRAD: {
      a += g * g;
      p -= o.lr * g / sqrtf(a + 1e-6f);
      break;
    }
    case SBR_UPD_RMSPROP: {
      a = o.rho * a + (1.f - o.rho) * g * g;
      p -= o.lr * g / sqrtf(a + 1e-6f);
      break;
    }
    case SBR_UPD_ADADELTA: {
      a = o.rho * a + (1.f - o.rho) * g * g;
      const float upd = g * sqrtf(b + 1e-6f) / sqrtf(a + 1e-6f);
      p -= o.lr * upd;
      b = o.rho * b + (1.f - o.rho) * upd * upd;
      break;
    }
    default: {  // nesterov momentum
      a = o.rho * a - o.lr * g;
      p += o.rho * a - o.lr * g;
      break;
    }
  }
}

__global__ void __launch_bounds__(256) optimizer_kernel(const OptArgs o) {
  const int64_t n4 = o.n >> 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  float4* p4 = reinterpret_cast<float4*>(o.p);
  float4* g4 = reinterpret_cast<float4*>(o.g);
  float4* a4 = reinterpret_cast<float4*>(o.a);
  float4* b4 = reinterpret_cast<float4*>(o.b);
  const bool two = (o.kind == SBR_UPD_ADAM || o.kind == SBR_UPD_ADADELTA);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 p = p4[i], g = g4[i], a = a4[i];
    float4 b = two ? b4[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    update_one(o, p.x, g.x, a.x, b.x);
    update_one(o, p.y, g.y, a.y, b.y);
    update_one(o, p.z, g.z, a.z, b.z);
    update_one(o, p.w, g.w, a.w, b.w);
    p4[i] = p;
    a4[i] = a;
    if (two) b4[i] = b;
    if (o.zero_grad) g4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

}  // namespace

int launch_optimizer(sbr_model* m) {
  OptArgs o{};
  o.p = m->params; o.g = m->grads; o.a = m->opt_a; o.b = m->opt_b;
  o.n = m->P_pad;   // P_pad is a multiple of 4; padding and the cost slot hold p = g = 0 and stay 0
  o.kind = m->cfg.updater;
  o.lr = m->cfg.lr; o.rho = m->cfg.rho; o.b1 = m->cfg.beta1; o.b2 = m->cfg.beta2;
  m->opt_t += 1;
  // a_t = lr * sqrt(1 - b2^t) / (1 - b1^t), evaluated in fp32 like the reference's floatX graph
  const float t = (float)m->opt_t;
  o.a_t = o.lr * sqrtf(1.f - powf(o.b2, t)) / (1.f - powf(o.b1, t));
  o.zero_grad = 1;
  const int64_t n4 = o.n >> 2;
  int grid = (int)std::min<int64_t>((n4 + 255) / 256, (int64_t)m->n_sm * 8);
  grid = std::max(grid, 1);
  optimizer_kernel<<<grid, 256, 0, m->stream>>>(o);
  KERNEL_CHECK(m);
  return 0;
}
