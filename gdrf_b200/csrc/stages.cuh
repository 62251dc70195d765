// Per-observation (HBM-bound) stages of the ELBO and its gradient, and the operand packers.
// Reference semantics restated in SURVEY.md section 8(c); citations are to /root/reference files.
#pragma once
#include "common.cuh"

namespace gdrf {

struct Hyper {
  const float* variance;     // [1]
  const float* lengthscale;  // [ls_dim]
  const float* noise;        // [1]
  const float* alpha;        // [1] RationalQuadratic scale_mixture (NULL otherwise)
  int ls_dim, kid, D;
};

// slots of the fp64 accumulator block (zeroed at the start of every step)
enum { ACC_LP_MU = 0, ACC_LQ = 1, ACC_LL = 2, ACC_LP_PHI = 3, ACC_DNOISE = 4, ACC_DVAR = 5, ACC_DLS = 6,
       ACC_DALPHA = ACC_DLS + MAX_D, ACC_HEAD = ACC_DLS + MAX_D + 2 };

// compile-time (input dimension, kernel id) dispatch for the two kernels that evaluate k(x, z) element by element;
// DT = 0 keeps the run-time dimension loop (D > 3)
#define GDRF_DISPATCH_D_(D_, KID_, F_)                                                  \
  switch ((D_) <= 3 ? (D_) : 0) {                                                          \
    case 1: F_(1, KID_); break;                                                            \
    case 2: F_(2, KID_); break;                                                            \
    case 3: F_(3, KID_); break;                                                            \
    default: F_(0, KID_); break;                                                           \
  }
#define GDRF_DISPATCH_DK(D_, KID_, F_)                                                     \
  switch (KID_) {                                                                          \
    case 0: GDRF_DISPATCH_D_(D_, 0, F_); break;                                            \
    case 1: GDRF_DISPATCH_D_(D_, 1, F_); break;                                            \
    case 2: GDRF_DISPATCH_D_(D_, 2, F_); break;                                            \
    case 3: GDRF_DISPATCH_D_(D_, 3, F_); break;                                            \
    default: GDRF_DISPATCH_D_(D_, 4, F_); break;                                           \
  }

// ---------------------------------------------------------------------------------------------
// K_xz = k(xs, Z) for one 128 x 64 block per CTA, written as three bf16 planes.
// Direct differences (more accurate than the reference's |x|^2 - 2xz + |z|^2 expansion,
// pyro Isotropy._square_scaled_dist); padding rows / columns are written as zeros.
// 8 lanes per 128-byte row; a thread keeps its 8 inducing points in registers.
// ---------------------------------------------------------------------------------------------
template <int DT, int KID>
__global__ void __launch_bounds__(256) k_kxz_planes(const float* __restrict__ xs, int nc, const float* __restrict__ Z,
                                                    int M, Hyper hp, PlaneMat kxz) {
  constexpr int DM = DT ? DT : MAX_D;
  const int D = DT ? DT : hp.D;
  const int cb = blockIdx.x, rt = blockIdx.y;
  const int g = threadIdx.x & 7;
  // fp64 throughout: the scaled distance is an argument of exp(), which turns its relative rounding error into
  // |r2 / 2| times as much in K_xz, and the whitening W = Kxz L^-T amplifies that again by its cancellation
  double il[DM], zl[8][DM];
#pragma unroll
  for (int d = 0; d < DM; ++d)
    if (d < D) il[d] = 1.0 / (double)hp.lengthscale[hp.ls_dim == 1 ? 0 : d];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int col = cb * 64 + g * 8 + j;
#pragma unroll
    for (int d = 0; d < DM; ++d)
      if (d < D) zl[j][d] = (col < M) ? (double)Z[col * D + d] : 0.0;
  }
  const double var = hp.variance[0];
  const double alpha = (KID == KERNEL_RQ) ? (double)hp.alpha[0] : 1.0;
  const int cvalid = min(8, M - (cb * 64 + g * 8));     // columns of this thread inside M (may be <= 0)
#pragma unroll
  for (int w = 0; w < 4; ++w) {
    const int r = w * 32 + (threadIdx.x >> 3);
    const int n = rt * 128 + r;
    double x[DM];
#pragma unroll
    for (int d = 0; d < DM; ++d)
      if (d < D) x[d] = (n < nc) ? (double)xs[(long long)n * D + d] : 0.0;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      double r2 = 0.0;
#pragma unroll
      for (int d = 0; d < DM; ++d)
        if (d < D) {
          const double t = (x[d] - zl[j][d]) * il[d];
          r2 = fma(t, t, r2);
        }
      double k, dk;
      kernel_eval<double>(KID, r2, k, dk, alpha);
      v[j] = (n < nc && j < cvalid) ? (float)(var * k) : 0.f;
    }
    uint4 pk[3];
    split8<3>(v, pk);
#pragma unroll
    for (int pl = 0; pl < 3; ++pl)
      *reinterpret_cast<uint4*>(kxz.elem(pl, n, cb * 64 + g * 8)) = pk[pl];
  }
}

// ---------------------------------------------------------------------------------------------
// ST[(k, j), i] = S_k[i, j] for i >= j (u_scale_tril, sparse_gdrf.py:100-110); zero elsewhere.
//   F16 = true : two fp16 planes (22-bit operands) twice -- `st16` with the rows of every 256-row tile permuted for the
//                forward contraction's narrow MMAs (policies.cuh, G2) and `st16n` in natural order for dW (G3)
//   F16 = false: three bf16 planes (`st`; any fp32 range) for the 24-bit forward and the bf16 backward
// max |S| goes to ps[PS_SMAX]: the prologue turns it into the "leaves the fp16 range" status.
// grid (Mp/64 [j tile], Mp/64 [i tile], K), 256 threads; transposes through shared memory.
// ---------------------------------------------------------------------------------------------
template <bool F16>
__global__ void __launch_bounds__(256) k_pack_st(const float* __restrict__ S, int K, int M, int Mp, PlaneMat st,
                                                 PlaneMat st16, PlaneMat st16n, unsigned* __restrict__ ps) {
  __shared__ float tile[64][65];
  const int jt = blockIdx.x, it = blockIdx.y, k = blockIdx.z;
  float mx = 0.f;
  for (int t = threadIdx.x; t < 64 * 64; t += 256) {
    const int ii = t >> 6, jj = t & 63;
    const int i = it * 64 + ii, j = jt * 64 + jj;
    const float v = (i < M && j < M && i >= j) ? S[((long long)k * M + i) * M + j] : 0.f;
    tile[ii][jj] = v;
    mx = fmaxf(mx, fabsf(v));
  }
  __syncthreads();
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0 && mx > 0.f) atomic_max_abs(ps + PS_SMAX, mx);
  for (int t = threadIdx.x; t < 64 * 8; t += 256) {
    const int jj = t >> 3, g = t & 7;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = tile[g * 8 + e][jj];
    const int row = k * Mp + jt * 64 + jj;
    if (!F16) {
      uint4 pk[3];
      split8<3>(v, pk);
#pragma unroll
      for (int pl = 0; pl < 3; ++pl) *reinterpret_cast<uint4*>(st.elem(pl, row, it * 64 + g * 8)) = pk[pl];
    } else {
      uint4 hk[2];
      split8h<2>(v, hk);
      // permuted copy: rows of every 256-row tile in the order [g3a g2a g1a g0a | g0b g1b g2b g3b] (policies.cuh, G2)
      const int jl = (jt * 64 + jj) & 255, X = jl >> 6, w32 = jl & 31;
      const int pos = (jl & 32) ? 128 + 32 * X + w32 : 32 * (3 - X) + w32;
      const int row16 = row - jl + pos;
#pragma unroll
      for (int pl = 0; pl < 2; ++pl) {
        *reinterpret_cast<uint4*>(st16.elem(pl, row16, it * 64 + g * 8)) = hk[pl];
        *reinterpret_cast<uint4*>(st16n.elem(pl, row, it * 64 + g * 8)) = hk[pl];
      }
    }
  }
}

// LINV planes (3 bf16 for the whitening G1; 2 fp16 for dKxz, G4) from the fp64 inverse Cholesky factor [Mp][Mp];
// identity padding is dropped.  max |Linv| goes to ps[PS_LINVMAX].
__global__ void __launch_bounds__(256) k_pack_linv(const double* __restrict__ Linv, int M, int Mp, PlaneMat linv,
                                                   PlaneMat linv16, unsigned* __restrict__ ps) {
  const int cb = blockIdx.x, rt = blockIdx.y;
  float mx = 0.f;
  for (int t = threadIdx.x; t < 128 * 8; t += 256) {
    const int r = t >> 3, g = t & 7;
    const int m = rt * 128 + r;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int i = cb * 64 + g * 8 + e;
      v[e] = (m < M && i <= m) ? (float)Linv[(long long)m * Mp + i] : 0.f;
      mx = fmaxf(mx, fabsf(v[e]));
    }
    uint4 pk[3];
    split8<3>(v, pk);
#pragma unroll
    for (int pl = 0; pl < 3; ++pl) *reinterpret_cast<uint4*>(linv.elem(pl, m, cb * 64 + g * 8)) = pk[pl];
    uint4 hk[2];
    split8h<2>(v, hk);
#pragma unroll
    for (int pl = 0; pl < 2; ++pl) *reinterpret_cast<uint4*>(linv16.elem(pl, m, cb * 64 + g * 8)) = hk[pl];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0 && mx > 0.f) atomic_max_abs(ps + PS_LINVMAX, mx);
}

// U16 [256 rows (k; zero beyond K) x Mp] = fp16 pair of s_u * u_loc for GF; s_u from ps[PS_UMAX], 1 / s_u -> ps[PS_SU_INV]
__global__ void __launch_bounds__(256) k_pack_u(const float* __restrict__ u, int K, int M, PlaneMat u16,
                                                unsigned* __restrict__ ps) {
  const int cb = blockIdx.x, rt = blockIdx.y;
  float inv;
  const float s_u = pow2_scale(__uint_as_float(ps[PS_UMAX]), &inv);
  if (cb == 0 && rt == 0 && threadIdx.x == 0) reinterpret_cast<float*>(ps)[PS_SU_INV] = inv;
  for (int t = threadIdx.x; t < 128 * 8; t += 256) {
    const int r = t >> 3, g = t & 7;
    const int k = rt * 128 + r;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int m = cb * 64 + g * 8 + e;
      v[e] = (k < K && m < M) ? s_u * u[(long long)k * M + m] : 0.f;
    }
    uint4 hk[2];
    split8h<2>(v, hk);
#pragma unroll
    for (int pl = 0; pl < 2; ++pl) *reinterpret_cast<uint4*>(u16.elem(pl, k, cb * 64 + g * 8)) = hk[pl];
  }
}

// max |x| over a small array (u_loc) into a per-step slot
__global__ void k_absmax(const float* __restrict__ x, long long n, unsigned* __restrict__ slot) {
  float mx = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    mx = fmaxf(mx, fabsf(x[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0 && mx > 0.f) atomic_max_abs(slot, mx);
}

// ---------------------------------------------------------------------------------------------
// f_loc[k, n] = sum_m W[n, m] u_loc[k, m]   (pyro conditional: loc = W @ v_2D; mean function is zero,
// abstract_gdrf.py:17-18).  grid (RT, ceil(K/KT)), 128 threads = rows; W rebuilt from its 3 planes.
// ---------------------------------------------------------------------------------------------
template <int KT>   // topics per CTA
__global__ void __launch_bounds__(128) k_floc(PlaneMat w, const float* __restrict__ u, int K, int M, int MB,
                                              double* __restrict__ floc, int ncp) {
  __shared__ __align__(16) float us[KT][64];
  const int rt = blockIdx.x, kg = blockIdx.y;
  const int n = rt * 128 + threadIdx.x;
  double acc[KT];     // fp32 dot products over 64 columns, summed across blocks in fp64
#pragma unroll
  for (int i = 0; i < KT; ++i) acc[i] = 0.0;
  for (int mb = 0; mb < MB; ++mb) {
    __syncthreads();
    for (int t = threadIdx.x; t < KT * 64; t += 128) {
      const int kk = t >> 6, c = t & 63;
      const int k = kg * KT + kk, m = mb * 64 + c;
      us[kk][c] = (k < K && m < M) ? u[(long long)k * M + m] : 0.f;
    }
    __syncthreads();
    float part[KT];
#pragma unroll
    for (int i = 0; i < KT; ++i) part[i] = 0.f;
#pragma unroll 2
    for (int g = 0; g < 8; ++g) {
      uint4 pk[3];
#pragma unroll
      for (int pl = 0; pl < 3; ++pl) pk[pl] = *reinterpret_cast<const uint4*>(w.elem(pl, n, mb * 64 + g * 8));
      float wv[8];
      join8<3>(pk, wv);
#pragma unroll
      for (int kk = 0; kk < KT; ++kk) {
        const float4 a = *reinterpret_cast<const float4*>(&us[kk][g * 8]);       // warp-wide broadcast loads
        const float4 b = *reinterpret_cast<const float4*>(&us[kk][g * 8 + 4]);
        float t = fmaf(wv[0], a.x, part[kk]);
        t = fmaf(wv[1], a.y, t);
        t = fmaf(wv[2], a.z, t);
        t = fmaf(wv[3], a.w, t);
        t = fmaf(wv[4], b.x, t);
        t = fmaf(wv[5], b.y, t);
        t = fmaf(wv[6], b.z, t);
        part[kk] = fmaf(wv[7], b.w, t);
      }
    }
#pragma unroll
    for (int i = 0; i < KT; ++i) acc[i] += (double)part[i];
  }
#pragma unroll
  for (int kk = 0; kk < KT; ++kk) {
    const int k = kg * KT + kk;
    if (k < K) floc[(long long)k * ncp + n] = acc[kk];
  }
}

__global__ void k_phisum(const float* __restrict__ phi, int K, int V, float* __restrict__ phisum) {
  __shared__ float scratch[32];
  const int k = blockIdx.x;
  float s = 0.f;
  for (int v = threadIdx.x; v < V; v += blockDim.x) s += phi[(long long)k * V + v];
  s = block_sum(s, scratch);
  if (threadIdx.x == 0) phisum[k] = s;
}

// ---------------------------------------------------------------------------------------------
// Per observation, before the likelihood pass (sparse_gdrf.py:354-361, 403-405):
//   f_var = var0 + q,  mu = f_loc + f_var * eps  (the variance is the Normal's scale),
//   log q(mu), log p(mu) with scale f_var + noise,  theta = softmax_k(mu),  s = sum_k theta_k rowsum(phi)_k
// 8 lanes per observation (lane j owns topics j, j + 8, ...; K <= 8 KQ), 256 threads = 32 observations.
// ---------------------------------------------------------------------------------------------
template <int KQ>
__global__ void __launch_bounds__(256) k_obs_prepare(int nc, int ncp, int K, long long n0, long long n_stride,
                                                     const double* __restrict__ floc, const double* __restrict__ q,
                                                     const double* __restrict__ wsq, const float* __restrict__ eps,
                                                     Hyper hp, const float* __restrict__ phisum,
                                                     float* __restrict__ fvar, float* __restrict__ theta,
                                                     float* __restrict__ srow, double* __restrict__ acc,
                                                     double inv_p) {
  // fp64 throughout: mu carries f_var * eps with f_var = O(variance), and d ll / d mu = O(counts), so fp32
  // rounding of mu (1e-5 absolute) alone would cost 1e-4 ... 1e-3 relative in the gradients.
  __shared__ double scratch[32];
  const int sub = threadIdx.x & 7;
  const int n = blockIdx.x * 32 + (threadIdx.x >> 3);
  const bool live = n < nc;
  double lq = 0.0, lp = 0.0;
  double m[KQ];
  const double var = hp.variance[0], noise = hp.noise[0];
  const double var0 = live ? fmax(var - wsq[n], 0.0) : 0.0;
  const double HALF_LOG_2PI = 0.91893853320467274178;
  double mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < KQ; ++i) {
    const int k = sub + 8 * i;
    m[i] = -INFINITY;
    if (live && k < K) {
      const long long o = (long long)k * ncp + n;
      const double fv = var0 + q[o];
      const double e = eps[(long long)k * n_stride + n0 + n];
      const double d = fv * e;
      m[i] = floc[o] + d;
      const double sp = fv + noise;
      lq += -log(fv) - HALF_LOG_2PI - 0.5 * e * e;
      const double z = d / sp;
      lp += -log(sp) - HALF_LOG_2PI - 0.5 * z * z;
      fvar[o] = (float)fv;
      mx = fmax(mx, m[i]);
    }
  }
#pragma unroll
  for (int o = 1; o < 8; o <<= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  double den = 0.0;
#pragma unroll
  for (int i = 0; i < KQ; ++i) {
    m[i] = (live && sub + 8 * i < K) ? exp(m[i] - mx) : 0.0;
    den += m[i];
  }
#pragma unroll
  for (int o = 1; o < 8; o <<= 1) den += __shfl_xor_sync(0xffffffffu, den, o);
  const double inv = live ? 1.0 / den : 0.0;
  double sr = 0.0;
#pragma unroll
  for (int i = 0; i < KQ; ++i) {
    const int k = sub + 8 * i;
    if (live && k < K) {
      const double t = m[i] * inv;
      theta[(long long)k * ncp + n] = (float)t;
      sr += t * (double)phisum[k];
    }
  }
#pragma unroll
  for (int o = 1; o < 8; o <<= 1) sr += __shfl_xor_sync(0xffffffffu, sr, o);
  if (live && sub == 0) srow[n] = (float)sr;
  lq = block_sum(lq, scratch);
  lp = block_sum(lp, scratch);
  if (threadIdx.x == 0) {      // inv_p = 1 / particles: the ELBO is the mean over the guide's draws
    atomicAdd(&acc[ACC_LQ], lq * inv_p);
    atomicAdd(&acc[ACC_LP_MU], lp * inv_p);
  }
}

// ---------------------------------------------------------------------------------------------
// Fused mixture + multinomial log-likelihood, forward and backward in one pass over ws
// (sparse_gdrf.py:361-372 -> torch.distributions.Multinomial.log_prob):
//   p = theta phi;  phat = p / sum_v p;  logit = log(clamp(phat, eps32, 1 - eps32))
//   ll_n = lgamma(n_n + 1) - sum_v lgamma(w + 1) + sum_v w logit
// The N x V probability matrix never leaves the SM.  Backward, with a_v = [eps32 <= phat <= 1-eps32],
// r_v = w_v a_v / p_v and A_n = sum_v w_v a_v:
//   d ll / d theta_k = sum_v phi_kv r_v - (A_n / s_n) rowsum(phi)_k          (G1 accumulates the first term)
//   d ll / d phi_kv  = sum_n theta_nk r_nv - sum_n theta_nk A_n / s_n        (second term: k_obs_finalize)
// 512 threads; observation tiles of 32; V processed in chunks of VJ*32 columns so that phi's chunk,
// the r tile and the d-phi register accumulators fit on chip.
// ---------------------------------------------------------------------------------------------
constexpr int LK_THREADS = 512;
constexpr int LK_TN = 32;
constexpr int LK_TS = LK_TN + 4;   // row stride of the observation-minor tiles (float4 access, conflict-free)

__host__ __device__ inline size_t lk_round4(size_t x) { return (x + 3) & ~(size_t)3; }
// shared-memory floats: phi chunk [K][VC+1] | r tile [VC][TS] | theta tile [K][TS] | G1 tile [TN][K+1] | 4 x [TN]
__host__ __device__ inline size_t lk_smem_floats(int K, int VC) {
  return lk_round4((size_t)K * (VC + 1)) + (size_t)VC * LK_TS + (size_t)K * LK_TS + lk_round4((size_t)LK_TN * (K + 1)) +
         4 * LK_TN;
}

template <int KPW, int VJ>
__global__ void __launch_bounds__(LK_THREADS, 1)
    k_likelihood(int nc, int ncp, int K, int V, const int* __restrict__ ws, const float* __restrict__ theta,
                 const float* __restrict__ srow, const float* __restrict__ phi, float* __restrict__ g1 /*[ncp][K]*/,
                 float* __restrict__ arow, float* __restrict__ cnt, double* __restrict__ dphi_acc /*[K][V]*/,
                 double* __restrict__ acc, double inv_p) {
  constexpr int VC = VJ * 32;
  extern __shared__ __align__(16) float lk_smem[];
  float* phis = lk_smem;                                    // [K][VC + 1]
  float* rt = phis + lk_round4((size_t)K * (VC + 1));       // [VC][TS]   r[v][obs]
  float* th = rt + VC * LK_TS;                              // [K][TS]    theta[k][obs]
  float* g1s = th + K * LK_TS;                              // [TN][K + 1]
  float* sinv = g1s + lk_round4((size_t)LK_TN * (K + 1));   // [TN]
  float* sh_ll = sinv + LK_TN;                              // [TN] per-observation partial sums of this V chunk
  float* sh_a = sh_ll + LK_TN;
  float* sh_c = sh_a + LK_TN;
  __shared__ double scratch[32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = warp & 7, vh = warp >> 3;                 // 4 observations x half of the chunk's columns per warp
  const int i0 = grp * 4;
  const int ntiles = (nc + LK_TN - 1) / LK_TN;
  const float EPS32 = 1.1920928955078125e-07f;
  double ll_local = 0.0;

  for (int v0 = 0; v0 < V; v0 += VC) {
    __syncthreads();
    for (int t = threadIdx.x; t < K * VC; t += LK_THREADS) {
      const int k = t / VC, v = t - k * VC;
      phis[k * (VC + 1) + v] = (v0 + v < V) ? phi[(long long)k * V + v0 + v] : 0.f;
    }
    float dacc[KPW][VJ];
#pragma unroll
    for (int a = 0; a < KPW; ++a)
#pragma unroll
      for (int j = 0; j < VJ; ++j) dacc[a][j] = 0.f;

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int nb = tile * LK_TN;
      __syncthreads();
      for (int t = threadIdx.x; t < LK_TN * K; t += LK_THREADS) {
        const int k = t / LK_TN, i = t - k * LK_TN;
        th[k * LK_TS + i] = (nb + i < nc) ? theta[(long long)k * ncp + nb + i] : 0.f;
      }
      for (int t = threadIdx.x; t < LK_TN * (K + 1); t += LK_THREADS) g1s[t] = 0.f;
      if (threadIdx.x < LK_TN) {
        sinv[threadIdx.x] = (nb + threadIdx.x < nc) ? 1.f / srow[nb + threadIdx.x] : 0.f;
        sh_ll[threadIdx.x] = 0.f;
        sh_a[threadIdx.x] = 0.f;
        sh_c[threadIdx.x] = 0.f;
      }
      __syncthreads();
      // ---- (b) p = theta phi for 4 observations at once, log-likelihood, r = w a / p ----
      {
        float llp[4] = {0.f, 0.f, 0.f, 0.f}, ap[4] = {0.f, 0.f, 0.f, 0.f}, cp[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
        for (int j = vh * (VJ / 2); j < (vh + 1) * (VJ / 2); ++j) {
          const int v = lane + 32 * j;
          int c[4];
          bool any = false;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int n = nb + i0 + q;
            c[q] = (n < nc && v0 + v < V) ? ws[(long long)n * V + v0 + v] : 0;
            any |= (c[q] != 0);
          }
          float4 r4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (any) {
            float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f;
            for (int k = 0; k < K; ++k) {
              const float f = phis[k * (VC + 1) + v];
              const float4 t4 = *reinterpret_cast<const float4*>(th + k * LK_TS + i0);
              p0 = fmaf(t4.x, f, p0);
              p1 = fmaf(t4.y, f, p1);
              p2 = fmaf(t4.z, f, p2);
              p3 = fmaf(t4.w, f, p3);
            }
            const float pq[4] = {p0, p1, p2, p3};
            float rq[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              rq[q] = 0.f;
              if (c[q] != 0) {
                const float ph = pq[q] * sinv[i0 + q];
                const float pc = fminf(fmaxf(ph, EPS32), 1.f - EPS32);
                const float cf = (float)c[q];
                llp[q] += cf * logf(pc) - (c[q] > 1 ? lgammaf(cf + 1.f) : 0.f);   // logf, not __logf: phat may sit at the 1 - eps clamp
                cp[q] += cf;
                if (ph >= EPS32 && ph <= 1.f - EPS32) {
                  ap[q] += cf;
                  rq[q] = cf / pq[q];
                }
              }
            }
            r4 = make_float4(rq[0], rq[1], rq[2], rq[3]);
          }
          *reinterpret_cast<float4*>(rt + v * LK_TS + i0) = r4;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float a = warp_sum(llp[q]), b = warp_sum(ap[q]), d = warp_sum(cp[q]);
          if (lane == 0) {
            atomicAdd(&sh_ll[i0 + q], a);
            atomicAdd(&sh_a[i0 + q], b);
            atomicAdd(&sh_c[i0 + q], d);
          }
        }
      }
      __syncthreads();
      if (threadIdx.x < LK_TN && nb + threadIdx.x < nc) {
        const int n = nb + threadIdx.x;
        ll_local += (double)sh_ll[threadIdx.x];
        if (v0 == 0) {
          arow[n] = sh_a[threadIdx.x];
          cnt[n] = sh_c[threadIdx.x];
        } else {
          arow[n] += sh_a[threadIdx.x];
          cnt[n] += sh_c[threadIdx.x];
        }
      }
      // ---- (c) G1[n][k] += sum_v phi[k][v] r[n][v] : lanes = k, warp = (4 observations, half of the columns) ----
      for (int k = lane; k < K; k += 32) {
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        const float* pk = phis + k * (VC + 1);
#pragma unroll 4
        for (int v = vh * (VC / 2); v < (vh + 1) * (VC / 2); ++v) {
          const float f = pk[v];
          const float4 r4 = *reinterpret_cast<const float4*>(rt + v * LK_TS + i0);
          a0 = fmaf(f, r4.x, a0);
          a1 = fmaf(f, r4.y, a1);
          a2 = fmaf(f, r4.z, a2);
          a3 = fmaf(f, r4.w, a3);
        }
        atomicAdd(&g1s[(i0 + 0) * (K + 1) + k], a0);
        atomicAdd(&g1s[(i0 + 1) * (K + 1) + k], a1);
        atomicAdd(&g1s[(i0 + 2) * (K + 1) + k], a2);
        atomicAdd(&g1s[(i0 + 3) * (K + 1) + k], a3);
      }
      // ---- (d) dphi[k][v] += sum_n theta[n][k] r[n][v] : warps = k groups, lanes = v, 4 observations per load ----
      {
        const int kbase = warp * KPW;
#pragma unroll 2
        for (int n0 = 0; n0 < LK_TN; n0 += 4) {
          float4 t4[KPW];
#pragma unroll
          for (int a = 0; a < KPW; ++a)
            t4[a] = (kbase + a < K) ? *reinterpret_cast<const float4*>(th + (kbase + a) * LK_TS + n0)
                                    : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int j = 0; j < VJ; ++j) {
            const float4 r4 = *reinterpret_cast<const float4*>(rt + (lane + 32 * j) * LK_TS + n0);
#pragma unroll
            for (int a = 0; a < KPW; ++a)
              dacc[a][j] = fmaf(t4[a].x, r4.x, fmaf(t4[a].y, r4.y, fmaf(t4[a].z, r4.z, fmaf(t4[a].w, r4.w, dacc[a][j]))));
          }
        }
      }
      __syncthreads();
      for (int t = threadIdx.x; t < LK_TN * K; t += LK_THREADS) {
        const int i = t / K, k = t - i * K;
        const int n = nb + i;
        if (n < nc) {
          const float a = g1s[i * (K + 1) + k];
          if (v0 == 0) g1[(long long)n * K + k] = a;
          else g1[(long long)n * K + k] += a;
        }
      }
    }
#pragma unroll
    for (int a = 0; a < KPW; ++a) {
      const int k = warp * KPW + a;
      if (k < K) {
#pragma unroll
        for (int j = 0; j < VJ; ++j) {
          const int v = v0 + lane + 32 * j;
          if (v < V && dacc[a][j] != 0.f) atomicAdd(&dphi_acc[(long long)k * V + v], (double)dacc[a][j] * inv_p);
        }
      }
    }
  }
  ll_local = block_sum(ll_local, scratch);
  if (threadIdx.x == 0) atomicAdd(&acc[ACC_LL], ll_local * inv_p);
}

// ---------------------------------------------------------------------------------------------
// Per observation, after the likelihood pass: softmax backward and the gradients w.r.t. the
// marginal mean / variance, noise and (directly) the kernel variance.
//   g_theta_k = G1[n][k] - (A_n / s_n) rowsum(phi)_k ;  g_mu_k = theta_k (g_theta_k - sum_j theta_j g_theta_j)
//   dELBO/df_var = -1/sp - eps^2 f_var noise / sp^3 + 1/f_var + g_mu eps      (sp = f_var + noise)
//   dELBO/dnoise = -1/sp + eps^2 f_var^2 / sp^3
//   gv0_n = sum_k dELBO/df_var  where the clamp var0 = max(variance - |W_n|^2, 0) is inactive
// Outputs: g_loc = g_mu, g2 = 2 dELBO/df_var (row scale of R), gv0; padding rows are zeroed.
// Particles (Trace_ELBO(num_particles=P, vectorize_particles=True), train_script.py:330-335): the marginal moments do
// not depend on the draw and every contraction of the backward is linear in (g_loc, g2, gv0), so the P passes of the
// per-observation chain accumulate the MEAN of those weights here (first / last / inv_p) and the contractions run once.
// 8 lanes per observation (lane j owns topics j, j + 8, ...; K <= 8 KQ), 256 threads = 32 observations.
// ---------------------------------------------------------------------------------------------
template <int KQ>
__global__ void __launch_bounds__(256) k_obs_finalize(int nc, int ncp, int K, long long n0, long long n_stride,
                                                      const float* __restrict__ theta, const float* __restrict__ srow,
                                                      const float* __restrict__ g1, const float* __restrict__ arow,
                                                      const float* __restrict__ cnt, const float* __restrict__ fvar,
                                                      const double* __restrict__ wsq, const float* __restrict__ eps,
                                                      Hyper hp, const float* __restrict__ phisum,
                                                      float* __restrict__ g_loc, float* __restrict__ g2,
                                                      float* __restrict__ gv0, double* __restrict__ ck,
                                                      double* __restrict__ acc, int npad,
                                                      unsigned* __restrict__ cs, float inv_p, int first, int last) {
  __shared__ double scratch[32];
  __shared__ float cks[8 * KQ];
  __shared__ unsigned smax[3];
  if (threadIdx.x < 3) smax[threadIdx.x] = 0u;
  float mx_g2 = 0.f, mx_gl = 0.f;
  const int sub = threadIdx.x & 7;
  const int n = blockIdx.x * 32 + (threadIdx.x >> 3);
  for (int t = threadIdx.x; t < 8 * KQ; t += 256) cks[t] = 0.f;
  __syncthreads();
  double dnoise = 0.0, dvar = 0.0, llc = 0.0;
  const bool live = n < nc;          // the 8 lanes of an observation agree; shuffles stay outside any branch
  const float noise = hp.noise[0], var = hp.variance[0];
  const float ratio = live ? arow[n] / srow[n] : 0.f;
  float gt[KQ], th[KQ];
  float dot = 0.f;
#pragma unroll
  for (int i = 0; i < KQ; ++i) {
    const int k = sub + 8 * i;
    gt[i] = th[i] = 0.f;
    if (live && k < K) {
      gt[i] = g1[(long long)n * K + k] - ratio * phisum[k];
      th[i] = theta[(long long)k * ncp + n];
      dot = fmaf(th[i], gt[i], dot);
    }
  }
#pragma unroll
  for (int o = 1; o < 8; o <<= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
  float gsum = 0.f, dn = 0.f;
#pragma unroll
  for (int i = 0; i < KQ; ++i) {
    const int k = sub + 8 * i;
    if (k < K && n < npad) {
      const long long o = (long long)k * ncp + n;
      float gm = 0.f, gv = 0.f;
      if (live) {
        gm = th[i] * (gt[i] - dot);
        const float fv = fvar[o];
        const float e = eps[(long long)k * n_stride + n0 + n];
        const float sp = fv + noise;
        const float isp = 1.f / sp;
        const float e2 = e * e;
        gv = -isp - e2 * fv * noise * isp * isp * isp + 1.f / fv + gm * e;
        dn += -isp + e2 * fv * fv * isp * isp * isp;
        gsum += gv;
        // c_k = sum_n theta[n][k] A_n / s_n  (the renormalisation term of d ll / d phi)
        atomicAdd(&cks[k], inv_p * th[i] * ratio);
      }
      const float gl_new = (first ? 0.f : g_loc[o]) + inv_p * gm;      // padding rows are zeroed
      const float g2_new = (first ? 0.f : g2[o]) + inv_p * 2.f * gv;
      g_loc[o] = gl_new;
      g2[o] = g2_new;
      mx_g2 = fmaxf(mx_g2, fabsf(g2_new));
      mx_gl = fmaxf(mx_gl, fabsf(gl_new));
    }
  }
#pragma unroll
  for (int o = 1; o < 8; o <<= 1) gsum += __shfl_xor_sync(0xffffffffu, gsum, o);
  dnoise = (double)dn * (double)inv_p;
  if (sub == 0 && n < npad) {
    float g0 = 0.f;
    if (live) {
      const bool clamp_open = ((double)var - wsq[n]) >= 0.0;
      g0 = clamp_open ? gsum : 0.f;
      dvar = (double)g0 * (double)inv_p;
      llc = lgamma((double)cnt[n] + 1.0) * (double)inv_p;
    }
    const float g0_new = (first ? 0.f : gv0[n]) + inv_p * g0;
    gv0[n] = g0_new;
    if (g0_new != 0.f) atomicMax(&smax[2], __float_as_uint(fabsf(g0_new)));
  }
  // chunk-wide maxima of the backward weights: they size the power-of-two scales of the fp16 operand planes
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mx_g2 = fmaxf(mx_g2, __shfl_xor_sync(0xffffffffu, mx_g2, o));
    mx_gl = fmaxf(mx_gl, __shfl_xor_sync(0xffffffffu, mx_gl, o));
  }
  if ((threadIdx.x & 31) == 0) {
    if (mx_g2 > 0.f) atomicMax(&smax[0], __float_as_uint(mx_g2));
    if (mx_gl > 0.f) atomicMax(&smax[1], __float_as_uint(mx_gl));
  }
  __syncthreads();
  if (last && threadIdx.x < 3 && smax[threadIdx.x] != 0u) atomicMax(cs + CS_G2MAX + threadIdx.x, smax[threadIdx.x]);
  for (int k = threadIdx.x; k < K; k += 256)
    if (cks[k] != 0.f) atomicAdd(&ck[k], (double)cks[k]);
  dnoise = block_sum(dnoise, scratch);
  dvar = block_sum(dvar, scratch);
  llc = block_sum(llc, scratch);
  if (threadIdx.x == 0) {
    atomicAdd(&acc[ACC_DNOISE], dnoise);
    atomicAdd(&acc[ACC_DVAR], dvar);
    atomicAdd(&acc[ACC_LL], llc);
  }
}

// ---------------------------------------------------------------------------------------------
// Backward weights of one chunk from UPSTREAM gradients of the marginal moments (gdrf_moments_vjp: the vector-Jacobian
// product of SparseGDRF.forward, sparse_gdrf.py:277-319) instead of from the ELBO's per-observation chain:
//   g_loc = d/d f_loc,  g2 = 2 d/d f_var,  gv0_n = sum_k d/d f_var[k, n] where var0 = max(variance - |W_n|^2, 0) is open
// (f_var = var0 + q), the direct d/d variance, and the chunk maxima that size the fp16 operand scales.  Same layout and
// thread map as k_obs_finalize (8 lanes per observation); padding rows are zeroed.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_vjp_weights(int nc, int ncp, int K, long long n0, long long n_stride,
                                                     const float* __restrict__ up_loc, const float* __restrict__ up_var,
                                                     const double* __restrict__ wsq, Hyper hp,
                                                     float* __restrict__ g_loc, float* __restrict__ g2,
                                                     float* __restrict__ gv0, double* __restrict__ acc, int npad,
                                                     unsigned* __restrict__ cs) {
  __shared__ double scratch[32];
  __shared__ unsigned smax[3];
  if (threadIdx.x < 3) smax[threadIdx.x] = 0u;
  __syncthreads();
  const int sub = threadIdx.x & 7;
  const int n = blockIdx.x * 32 + (threadIdx.x >> 3);
  const bool live = n < nc;
  float mx_g2 = 0.f, mx_gl = 0.f, gsum = 0.f;
  if (n < npad)
    for (int k = sub; k < K; k += 8) {
      const long long o = (long long)k * ncp + n, u = (long long)k * n_stride + n0 + n;
      const float gl = live ? up_loc[u] : 0.f;
      const float gv = (live && up_var) ? up_var[u] : 0.f;
      g_loc[o] = gl;
      g2[o] = 2.f * gv;
      gsum += gv;
      mx_g2 = fmaxf(mx_g2, fabsf(2.f * gv));
      mx_gl = fmaxf(mx_gl, fabsf(gl));
    }
#pragma unroll
  for (int o = 1; o < 8; o <<= 1) gsum += __shfl_xor_sync(0xffffffffu, gsum, o);
  double dvar = 0.0;
  if (sub == 0 && n < npad) {
    float g0 = 0.f;
    if (live && ((double)hp.variance[0] - wsq[n]) >= 0.0) g0 = gsum;
    dvar = (double)g0;
    gv0[n] = g0;
    if (g0 != 0.f) atomicMax(&smax[2], __float_as_uint(fabsf(g0)));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mx_g2 = fmaxf(mx_g2, __shfl_xor_sync(0xffffffffu, mx_g2, o));
    mx_gl = fmaxf(mx_gl, __shfl_xor_sync(0xffffffffu, mx_gl, o));
  }
  if ((threadIdx.x & 31) == 0) {
    if (mx_g2 > 0.f) atomicMax(&smax[0], __float_as_uint(mx_g2));
    if (mx_gl > 0.f) atomicMax(&smax[1], __float_as_uint(mx_gl));
  }
  __syncthreads();
  if (threadIdx.x < 3 && smax[threadIdx.x] != 0u) atomicMax(cs + CS_G2MAX + threadIdx.x, smax[threadIdx.x]);
  dvar = block_sum(dvar, scratch);
  if (threadIdx.x == 0) atomicAdd(&acc[ACC_DVAR], dvar);
}

// ---------------------------------------------------------------------------------------------
// du_loc[k, m] += sum_n g_loc[k, n] W[n, m]     (adjoint of f_loc = W u_loc^T)
// One CTA per (64-column block, slab of row tiles).  Each 128-row tile of W is rebuilt from its planes into shared
// memory; a thread owns a 4-column x 8-topic register tile and a slice of the rows (4 rows per step: one 16-byte
// load of g per topic serves 4 rows).  256 threads = 16 column groups x KGP topic groups x (16 / KGP) row slices.
// dynamic shared memory: W tile [128][68] | g tile [8 KGP][128]   (reused for the cross-slice reduction)
// ---------------------------------------------------------------------------------------------
template <int KGP>   // topic groups of 8 (power of two, K <= 8 KGP <= 128)
__global__ void __launch_bounds__(256) k_du(PlaneMat w, const float* __restrict__ g_loc, int K, int M, int RT, int ncp,
                                            int tiles_per_cta, double* __restrict__ du_acc) {
  constexpr int RS = 16 / KGP;            // row slices
  constexpr int WS = 68;                  // W tile row stride (floats)
  extern __shared__ __align__(16) float du_smem[];
  float* wsm = du_smem;                   // [128][WS]
  float* gsm = du_smem + 128 * WS;        // [8 * KGP][128]
  const int cb = blockIdx.x;
  const int cg = threadIdx.x & 15, tq = threadIdx.x >> 4;
  const int kq = tq % KGP, slice = tq / KGP;
  float a[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) a[i][j] = 0.f;
  const int rt0 = blockIdx.y * tiles_per_cta;
  const int rt1 = min(RT, rt0 + tiles_per_cta);
  for (int rt = rt0; rt < rt1; ++rt) {
    __syncthreads();
    for (int t = threadIdx.x; t < 128 * 8; t += 256) {
      const int r = t >> 3, g = t & 7;
      uint4 pk[3];
#pragma unroll
      for (int pl = 0; pl < 3; ++pl) pk[pl] = *reinterpret_cast<const uint4*>(w.elem(pl, rt * 128 + r, cb * 64 + g * 8));
      float v[8];
      join8<3>(pk, v);
      *reinterpret_cast<float4*>(wsm + r * WS + g * 8) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(wsm + r * WS + g * 8 + 4) = make_float4(v[4], v[5], v[6], v[7]);
    }
    for (int t = threadIdx.x; t < 8 * KGP * 128; t += 256) {
      const int k = t >> 7, r = t & 127;
      gsm[t] = (k < K) ? g_loc[(long long)k * ncp + rt * 128 + r] : 0.f;
    }
    __syncthreads();
#pragma unroll 2
    for (int r0 = 4 * slice; r0 < 128; r0 += 4 * RS) {
      float4 wv[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) wv[q] = *reinterpret_cast<const float4*>(wsm + (r0 + q) * WS + cg * 4);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 gv = *reinterpret_cast<const float4*>(gsm + (kq * 8 + i) * 128 + r0);
        a[i][0] = fmaf(gv.x, wv[0].x, fmaf(gv.y, wv[1].x, fmaf(gv.z, wv[2].x, fmaf(gv.w, wv[3].x, a[i][0]))));
        a[i][1] = fmaf(gv.x, wv[0].y, fmaf(gv.y, wv[1].y, fmaf(gv.z, wv[2].y, fmaf(gv.w, wv[3].y, a[i][1]))));
        a[i][2] = fmaf(gv.x, wv[0].z, fmaf(gv.y, wv[1].z, fmaf(gv.z, wv[2].z, fmaf(gv.w, wv[3].z, a[i][2]))));
        a[i][3] = fmaf(gv.x, wv[0].w, fmaf(gv.y, wv[1].w, fmaf(gv.z, wv[2].w, fmaf(gv.w, wv[3].w, a[i][3]))));
      }
    }
  }
  // cross-slice reduction through shared memory: red[slice][k][64]
  __syncthreads();
  float* red = du_smem;
#pragma unroll
  for (int i = 0; i < 8; ++i)
    *reinterpret_cast<float4*>(red + ((slice * 8 * KGP) + kq * 8 + i) * 64 + cg * 4) =
        make_float4(a[i][0], a[i][1], a[i][2], a[i][3]);
  __syncthreads();
  for (int t = threadIdx.x; t < 8 * KGP * 64; t += 256) {
    const int k = t >> 6, c = t & 63;
    const int m = cb * 64 + c;
    if (k < K && m < M) {
      float v = 0.f;
#pragma unroll
      for (int sl = 0; sl < RS; ++sl) v += red[(sl * 8 * KGP + k) * 64 + c];
      atomicAdd(&du_acc[(long long)k * M + m], (double)v);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// WG[n, (k, m)] = s_g * g2[k, n] * W[n, m]  (two planes, fp16 or bf16): the per-topic row-weighted copies of W that
// turn dS_k = W^T diag(g2_k) T_k into a plain contraction.  s_g is the power-of-two scale that puts
// max |g2| * sqrt(variance) (>= max |WG|: |W_n| <= sqrt(k(x, x))) at 2^13..2^14; G6's epilogue multiplies by 1 / s_g
// (cs[CS_SG_INV]).  grid (MB, RT), 256 threads; W's block is rebuilt once and rescaled K times.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_scale_w(PlaneMat w, const float* __restrict__ g2, int K, int MB, int ncp,
                                                 PlaneMat wg, int fmt, const float* __restrict__ variance,
                                                 unsigned* __restrict__ cs) {
  const int cb = blockIdx.x, rt = blockIdx.y;
  const int g = threadIdx.x & 7;            // 8 lanes cover the eight 16-byte chunks of one 128-byte row
  float inv_s;
  const float s_g = pow2_scale(__uint_as_float(cs[CS_G2MAX]) * sqrtf(variance[0]), &inv_s);
  if (cb == 0 && rt == 0 && threadIdx.x == 0) reinterpret_cast<float*>(cs)[CS_SG_INV] = inv_s;
#pragma unroll 1
  for (int it = 0; it < 4; ++it) {
    const int r = it * 32 + (threadIdx.x >> 3);
    const int n = rt * 128 + r;
    uint4 pk[3];
#pragma unroll
    for (int pl = 0; pl < 3; ++pl) pk[pl] = *reinterpret_cast<const uint4*>(w.elem(pl, n, cb * 64 + g * 8));
    float wj[8];
    join8<3>(pk, wj);
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
      const float sc = s_g * g2[(long long)k * ncp + n];
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = sc * wj[j];
      uint4 out[2];
      split8x2(fmt, v, out);
#pragma unroll
      for (int pl = 0; pl < 2; ++pl)      // streaming store: 2.4 GB per chunk, read back only after it left L2
        __stcs(reinterpret_cast<uint4*>(wg.elem(pl, n, (k * MB + cb) * 64 + g * 8)), out[pl]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// dWtot = s_d * (dW (from G3) + sum_k g_loc[k, n] u_loc[k, m] - 2 gv0[n] W[n, m])   -> 2 planes (fp16 or bf16)
// s_d: power-of-two scale from the bound max|dW| + K max|g_loc| max|u_loc| + 2 max|gv0| sqrt(variance); the epilogues of
// G4 / G5 multiply by 1 / s_d (cs[CS_SD_INV]).
// grid (MB, RT), 256 threads: 8 lanes per 128-byte row, a thread owns 4 rows (r, r + 32, r + 64, r + 96) x 8 columns;
// the rank-K update runs first on a 4 x 8 register tile (per topic: one 16-byte load of g for the 4 rows, two of u).
// dynamic shared memory: u block [K][72] | g tile [K][128] with row r stored at (r & 31) * 4 + (r >> 5)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_dw_finalize(PlaneMat w, const float* __restrict__ dw, int Mp,
                                                     const float* __restrict__ g_loc, const float* __restrict__ gv0,
                                                     const float* __restrict__ u, int K, int M, int ncp, PlaneMat dwt,
                                                     int fmt, const float* __restrict__ variance,
                                                     unsigned* __restrict__ cs, const unsigned* __restrict__ ps,
                                                     int gl_blocks) {
  extern __shared__ __align__(16) float dwf_smem[];
  float inv_sd;
  const float s_d = pow2_scale(__uint_as_float(cs[CS_DWMAX]) +
                                   (float)K * __uint_as_float(cs[CS_GLOCMAX]) * __uint_as_float(ps[PS_UMAX]) +
                                   2.f * __uint_as_float(cs[CS_GV0MAX]) * sqrtf(variance[0]),
                               &inv_sd);
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) reinterpret_cast<float*>(cs)[CS_SD_INV] = inv_sd;
  float* us = dwf_smem;            // [K][72]: 64 columns, the upper 32 shifted by 4 words (bank spread)
  float* gs = us + K * 72;         // [K][128]
  const int cb = blockIdx.x, rt = blockIdx.y;
  for (int t = threadIdx.x; t < K * 64; t += 256) {
    const int k = t >> 6, c = t & 63;
    const int m = cb * 64 + c;
    us[k * 72 + c + (c >> 5) * 4] = (m < M) ? u[(long long)k * M + m] : 0.f;
  }
  for (int t = threadIdx.x; t < K * 128; t += 256) {
    const int k = t >> 7, pos = t & 127;
    const int r = (pos >> 2) + 32 * (pos & 3);
    gs[t] = g_loc[(long long)k * ncp + rt * 128 + r];
  }
  __syncthreads();
  if (gl_blocks > 0 && cb == 0) {
    // s_l * g_loc as gl_blocks extra 64-column blocks [Mp + 64 b, ...) of the dWtot planes (columns = topics): the A
    // operand of du_loc = g_loc W, contracted by G5 together with C5
    float inv_sl;
    const float s_l = pow2_scale(__uint_as_float(cs[CS_GLOCMAX]), &inv_sl);
    if (rt == 0 && threadIdx.x == 0) reinterpret_cast<float*>(cs)[CS_SL_INV] = inv_sl;
    for (int t = threadIdx.x; t < 128 * 8 * gl_blocks; t += 256) {
      const int r = (t >> 3) & 127, gq = t & 7, b = t >> 10;
      const int pos = (r & 31) * 4 + (r >> 5);
      float o8[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int k = b * 64 + gq * 8 + e;
        o8[e] = (k < K) ? s_l * gs[k * 128 + pos] : 0.f;
      }
      uint4 out[2];
      split8x2(fmt, o8, out);
#pragma unroll
      for (int pl = 0; pl < 2; ++pl) *reinterpret_cast<uint4*>(dwt.elem(pl, rt * 128 + r, Mp + b * 64 + gq * 8)) = out[pl];
    }
  }
  const int rr = threadIdx.x >> 3, g = threadIdx.x & 7;
  const int col = cb * 64 + g * 8;
  float v[4][8];
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int j = 0; j < 8; ++j) v[q][j] = 0.f;
  const float* up = us + g * 8 + (g >> 2) * 4;
#pragma unroll 4
  for (int k = 0; k < K; ++k) {
    const float4 gk = *reinterpret_cast<const float4*>(gs + k * 128 + rr * 4);
    const float4 u0 = *reinterpret_cast<const float4*>(up + k * 72);
    const float4 u1 = *reinterpret_cast<const float4*>(up + k * 72 + 4);
    const float uu[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
    const float gq[4] = {gk.x, gk.y, gk.z, gk.w};
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int j = 0; j < 8; ++j) v[q][j] = fmaf(gq[q], uu[j], v[q][j]);
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int n = rt * 128 + rr + 32 * q;
    uint4 pk[3];
#pragma unroll
    for (int pl = 0; pl < 3; ++pl) pk[pl] = *reinterpret_cast<const uint4*>(w.elem(pl, n, col));
    float wj[8];
    join8<3>(pk, wj);
    const float4* src = reinterpret_cast<const float4*>(dw + (long long)n * Mp + col);
    const float4 d0 = src[0], d1 = src[1];
    const float dd[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
    const float m2g = -2.f * gv0[n];
    float o8[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o8[j] = s_d * (fmaf(m2g, wj[j], dd[j]) + v[q][j]);
    uint4 out[2];      // G4 and G5 consume dWtot as two 16-bit planes
    split8x2(fmt, o8, out);
#pragma unroll
    for (int pl = 0; pl < 2; ++pl) *reinterpret_cast<uint4*>(dwt.elem(pl, n, col)) = out[pl];
  }
}

// ---------------------------------------------------------------------------------------------
// Chain dKxz into the kernel hyper-parameters and the inducing points:
//   K_xz[n,i] = variance * f(r2),  r2 = sum_d ((x_nd - z_id) / l_d)^2 = sum_d t_d^2
//   dvariance += dKxz f ;  dl_d += dKxz variance f' (-2 t_d^2 / l_d) ;  dZ[i,d] += dKxz variance f' (-2 t_d / l_d)
// grid (Mp/128, row splits), 128 threads = columns i; the constant factors (variance, -2 / l_d) are applied once
// per thread after the row loop.
// ---------------------------------------------------------------------------------------------
template <int DT, int KID>
__global__ void __launch_bounds__(128) k_kxz_backward(const float* __restrict__ dkxz, int Mp, const float* __restrict__ xs,
                                                      int nc, const float* __restrict__ Z, int M, Hyper hp,
                                                      int rows_per_cta, double* __restrict__ dz_acc,
                                                      double* __restrict__ acc) {
  constexpr int DM = DT ? DT : MAX_D;
  const int D = DT ? DT : hp.D;
  __shared__ double scratch[32];
  __shared__ float xsh[128][DM];
  const int i = blockIdx.x * 128 + threadIdx.x;
  const int r0 = blockIdx.y * rows_per_cta;
  const int r1 = min(nc, r0 + rows_per_cta);
  float z[DM], il[DM];
#pragma unroll
  for (int d = 0; d < DM; ++d)
    if (d < D) {
      il[d] = 1.f / hp.lengthscale[hp.ls_dim == 1 ? 0 : d];
      z[d] = (i < M) ? Z[i * D + d] : 0.f;
    }
  // fp64 running sums: these are sums over all observations of large terms of both signs
  double dz[DM], dl[DM];
#pragma unroll
  for (int d = 0; d < DM; ++d) dz[d] = dl[d] = 0.0;
  double dv = 0.0, da = 0.0;
  const float alpha = (KID == KERNEL_RQ) ? hp.alpha[0] : 1.f;
  const float* col = dkxz + i;
  for (int rb = r0; rb < r1; rb += 128) {
    __syncthreads();
    for (int t = threadIdx.x; t < 128 * D; t += 128) {
      const int r = t / D, d = t - r * D;
      xsh[r][d] = (rb + r < r1) ? xs[(long long)(rb + r) * D + d] : 0.f;
    }
    __syncthreads();
    const int rn = min(128, r1 - rb);
    if (i < M) {
#pragma unroll 4
      for (int r = 0; r < rn; ++r) {
        const float g = col[(long long)(rb + r) * Mp];
        float t[DM];
        float r2 = 0.f;
#pragma unroll
        for (int d = 0; d < DM; ++d)
          if (d < D) {
            t[d] = (xsh[r][d] - z[d]) * il[d];
            r2 = fmaf(t[d], t[d], r2);
          }
        float k, dk, dka = 0.f;
        kernel_eval<float>(KID, r2, k, dk, alpha, &dka);
        dv += (double)(g * k);
        if (KID == KERNEL_RQ) da += (double)(g * dka);
        const float h = g * dk;
#pragma unroll
        for (int d = 0; d < DM; ++d)
          if (d < D) {
            const float ht = h * t[d];
            dz[d] += (double)ht;
            dl[d] += (double)(ht * t[d]);
          }
      }
    }
  }
  const double c = -2.0 * (double)hp.variance[0];
  if (i < M)
#pragma unroll
    for (int d = 0; d < DM; ++d)
      if (d < D) atomicAdd(&dz_acc[i * D + d], c * (double)il[d] * dz[d]);
  double dvs = block_sum(dv, scratch);
  if (threadIdx.x == 0) atomicAdd(&acc[ACC_DVAR], dvs);
  if (KID == KERNEL_RQ) {
    double das = block_sum((double)hp.variance[0] * da, scratch);
    if (threadIdx.x == 0) atomicAdd(&acc[ACC_DALPHA], das);
  }
  if (hp.ls_dim == 1) {
    double s = 0.0;
#pragma unroll
    for (int d = 0; d < DM; ++d)
      if (d < D) s += c * (double)il[d] * dl[d];
    double t = block_sum(s, scratch);
    if (threadIdx.x == 0) atomicAdd(&acc[ACC_DLS], t);
  } else {
    for (int d = 0; d < D; ++d) {
      double t = block_sum(c * (double)il[d] * dl[d], scratch);
      if (threadIdx.x == 0) atomicAdd(&acc[ACC_DLS + d], t);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Dirichlet prior on phi (sparse_gdrf.py:358-360) and the final gradient assembly into the flat
// fp32 buffer  [dS K*M*M | du K*M | dphi K*V | dZ M*D | dvariance | dlengthscale ls_dim | dnoise].
// dS is accumulated in place by G6; everything else comes from the fp64 accumulators.
// ---------------------------------------------------------------------------------------------
__global__ void k_prior(const float* __restrict__ phi, const float* __restrict__ beta, int K, int V,
                        double* __restrict__ acc) {
  __shared__ double scratch[32];
  const int k = blockIdx.x;
  double sb = 0.0, t = 0.0;
  for (int v = threadIdx.x; v < V; v += blockDim.x) {
    const double b = beta[(long long)k * V + v];
    sb += b;
    t += (b - 1.0) * log((double)phi[(long long)k * V + v]) - lgamma(b);
  }
  sb = block_sum(sb, scratch);
  t = block_sum(t, scratch);
  if (threadIdx.x == 0) atomicAdd(&acc[ACC_LP_PHI], lgamma(sb) + t);
}

__global__ void k_assemble(int K, int M, int V, int D, int ls_dim, int has_alpha, int include_prior,
                           const float* __restrict__ phi,
                           const float* __restrict__ beta, const double* __restrict__ acc,
                           const double* __restrict__ ck, const double* __restrict__ du_acc,
                           const double* __restrict__ dphi_acc, const double* __restrict__ dz_acc,
                           float* __restrict__ grad) {
  const long long oS = 0, oU = oS + (long long)K * M * M, oP = oU + (long long)K * M, oZ = oP + (long long)K * V,
                  oV = oZ + (long long)M * D, oL = oV + 1, oN = oL + ls_dim;
  const long long total = (long long)K * M + (long long)K * V + (long long)M * D + 2 + ls_dim + has_alpha;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    long long idx = t;
    if (idx < (long long)K * M) { grad[oU + idx] = (float)du_acc[idx]; continue; }
    idx -= (long long)K * M;
    if (idx < (long long)K * V) {
      const int k = (int)(idx / V);
      double g = dphi_acc[idx] - ck[k];
      if (include_prior) g += ((double)beta[idx] - 1.0) / (double)phi[idx];
      grad[oP + idx] = (float)g;
      continue;
    }
    idx -= (long long)K * V;
    if (idx < (long long)M * D) { grad[oZ + idx] = (float)dz_acc[idx]; continue; }
    idx -= (long long)M * D;
    if (idx == 0) { grad[oV] = (float)acc[ACC_DVAR]; continue; }
    idx -= 1;
    if (idx < ls_dim) { grad[oL + idx] = (float)acc[ACC_DLS + idx]; continue; }
    idx -= ls_dim;
    if (idx == 0) { grad[oN] = (float)acc[ACC_DNOISE]; continue; }
    grad[oN + 1] = (float)acc[ACC_DALPHA];
  }
}

}  // namespace gdrf
