// Fused SVI tail: constraint transforms, their chain rule and an Adam / AdamW update
// on the flat parameter buffer, replacing the reference's per-parameter Python loop
// (gdrf/train_script.py:325-327 -> pyro.optim wrapping torch.optim; PyroParam constraints of
// gdrf/models/sparse_gdrf.py:79-112 and abstract_gdrf.py:79-84).
//
// Flat layout (same order as the gradient of gdrf_elbo_step):
//   [ u_scale_tril K*M*M | u_loc K*M | phi K*V | Z M*D | variance 1 | lengthscale ls_dim | noise 1 | scale_mixture 0/1 ]
// unconstrained -> constrained:  lower_cholesky (strict lower free, exp on the diagonal, zero above), identity,
// row softmax (stacked simplex), sigmoid (interval(0,1) per dimension) and exp (positive).
#pragma once
#include "common.cuh"

namespace gdrf {

struct FlatLayout {
  long long oS, oU, oP, oZ, oV, oL, oN, total;
  int K, M, V, D, ls_dim;
};

__host__ __device__ inline FlatLayout make_layout(int K, int M, int V, int D, int ls_dim, int has_alpha = 0) {
  FlatLayout f;
  f.K = K; f.M = M; f.V = V; f.D = D; f.ls_dim = ls_dim;
  f.oS = 0;
  f.oU = f.oS + (long long)K * M * M;
  f.oP = f.oU + (long long)K * M;
  f.oZ = f.oP + (long long)K * V;
  f.oV = f.oZ + (long long)M * D;
  f.oL = f.oV + 1;
  f.oN = f.oL + ls_dim;
  f.total = f.oN + 1 + has_alpha;   // RationalQuadratic's scale_mixture (positive, like the other kernel parameters)
  return f;
}

// phi = row softmax(u); one block per topic row.  Also usable alone.
__global__ void k_softmax_rows(const float* __restrict__ u, float* __restrict__ out, int V) {
  __shared__ float scratch[32];
  __shared__ float bc;
  const float* row = u + (long long)blockIdx.x * V;
  float mx = -INFINITY;
  for (int v = threadIdx.x; v < V; v += blockDim.x) mx = fmaxf(mx, row[v]);
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    float m = scratch[0];
    for (int w = 1; w < (blockDim.x + 31) / 32; ++w) m = fmaxf(m, scratch[w]);
    bc = m;
  }
  __syncthreads();
  mx = bc;
  float s = 0.f;
  for (int v = threadIdx.x; v < V; v += blockDim.x) s += expf(row[v] - mx);
  s = block_sum(s, scratch);
  if (threadIdx.x == 0) bc = s;
  __syncthreads();
  const float inv = 1.f / bc;
  for (int v = threadIdx.x; v < V; v += blockDim.x) out[(long long)blockIdx.x * V + v] = expf(row[v] - mx) * inv;
}

// every block of the flat buffer except phi
__global__ void k_constrain(FlatLayout f, const float* __restrict__ u, float* __restrict__ c, int learn_z) {
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < f.total;
       t += (long long)gridDim.x * blockDim.x) {
    if (t < f.oU) {
      const long long e = t % ((long long)f.M * f.M);
      const int i = (int)(e / f.M), j = (int)(e % f.M);
      c[t] = (i > j) ? u[t] : (i == j ? expf(u[t]) : 0.f);
    } else if (t < f.oP) {
      c[t] = u[t];
    } else if (t < f.oZ) {
      // phi: k_softmax_rows
    } else if (t < f.oV) {
      c[t] = learn_z ? 1.f / (1.f + expf(-u[t])) : u[t];   // fixed inducing points are stored constrained
    } else {
      c[t] = expf(u[t]);
    }
  }
}

// d_k = sum_v phi_kv g_kv  (softmax chain rule), one block per topic row
__global__ void k_phi_rowdot(const float* __restrict__ phi, const float* __restrict__ g, int V, float* __restrict__ out) {
  __shared__ float scratch[32];
  float s = 0.f;
  for (int v = threadIdx.x; v < V; v += blockDim.x)
    s = fmaf(phi[(long long)blockIdx.x * V + v], g[(long long)blockIdx.x * V + v], s);
  s = block_sum(s, scratch);
  if (threadIdx.x == 0) out[blockIdx.x] = s;
}

struct AdamHyper {
  float lr, beta1, beta2, eps, weight_decay;   // weight_decay > 0: decoupled (AdamW)
  float bc1, bc2;                               // 1 - beta1^t, 1 - beta2^t
  float grad_scale;                             // d loss / d ELBO  (= -1 / N for the reference's loss)
  float clip;                                   // > 0: pyro.optim.ClippedAdam (element-wise clamp, L2 weight decay,
                                                //      denom = sqrt(v) + eps, step = lr sqrt(bc2) / bc1); 0: Adam / AdamW
};

// grad is d ELBO / d constrained (the flat output of gdrf_elbo_step); one fused pass: chain rule to the
// unconstrained parameter, Adam moments, update.
__global__ void k_adam(FlatLayout f, float* __restrict__ u, const float* __restrict__ c, const float* __restrict__ grad,
                       const float* __restrict__ phi_dot, float* __restrict__ m, float* __restrict__ v, AdamHyper h,
                       int learn_z) {
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < f.total;
       t += (long long)gridDim.x * blockDim.x) {
    float g = h.grad_scale * grad[t];
    if (t < f.oU) {
      const long long e = t % ((long long)f.M * f.M);
      const int i = (int)(e / f.M), j = (int)(e % f.M);
      if (i < j) continue;                 // above the diagonal: not a parameter
      if (i == j) g *= c[t];
    } else if (t < f.oP) {
      // identity
    } else if (t < f.oZ) {
      const long long e = t - f.oP;
      g = c[t] * (g - h.grad_scale * phi_dot[e / f.V]);
    } else if (t < f.oV) {
      if (!learn_z) continue;
      g *= c[t] * (1.f - c[t]);
    } else {
      g *= c[t];
    }
    float x = u[t];
    if (h.clip > 0.f) {
      // pyro.optim.ClippedAdam (scripts/mvco.py:135): clamp, then L2 decay folded into the gradient
      g = fminf(fmaxf(g, -h.clip), h.clip);
      if (h.weight_decay != 0.f) g = fmaf(h.weight_decay, x, g);
      const float mc = h.beta1 * m[t] + (1.f - h.beta1) * g;
      const float vc = h.beta2 * v[t] + (1.f - h.beta2) * g * g;
      m[t] = mc;
      v[t] = vc;
      u[t] = x - (h.lr * sqrtf(h.bc2) / h.bc1) * mc / (sqrtf(vc) + h.eps);
      continue;
    }
    if (h.weight_decay > 0.f) x -= h.lr * h.weight_decay * x;
    const float mm = h.beta1 * m[t] + (1.f - h.beta1) * g;
    const float vv = h.beta2 * v[t] + (1.f - h.beta2) * g * g;
    m[t] = mm;
    v[t] = vv;
    const float denom = sqrtf(vv) / sqrtf(h.bc2) + h.eps;
    u[t] = x - (h.lr / h.bc1) * mm / denom;
  }
}

}  // namespace gdrf
