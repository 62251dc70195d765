// Small dense M x M linear algebra on the device (M <= 4096): Kuu, blocked Cholesky with a
// not-positive-definite flag (mirrors the try/except of gdrf/models/utils.py:27-40), triangular inverse,
// fp64 GEMMs and the Cholesky / kernel adjoints.  All matrices are row-major [Mp][Mp] with Mp a multiple
// of 256; the padding carries an identity so no kernel needs bounds checks.
#pragma once
#include "common.cuh"
#include "stages.cuh"

namespace gdrf {

constexpr int NB = 32;   // block size of the blocked factorisations

// ---------------------------------------------------------------------------------------------
// Kuu = k(Z, Z) + jitter * I.
//   T = double: direct differences in fp64 (the values used downstream)
//   T = float : the reference's own arithmetic -- |x|^2 - 2 x.z + |z|^2 expansion in fp32, clamp at 0, and
//               the jitter added as njitter+1 successive fp32 increments jitter*10^i (utils.py:33) -- used
//               only to decide whether the reference's fp32 Cholesky would have failed.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void k_kuu(const float* __restrict__ Z, int M, int Mp, Hyper hp, double jitter, int njitter,
                      T* __restrict__ Kuu) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.y;
  if (j >= Mp) return;
  T val;
  if (i >= M || j >= M) {
    val = (i == j) ? T(1) : T(0);
  } else {
    const int D = hp.D;
    T r2 = 0;
    if (sizeof(T) == 8) {
      for (int d = 0; d < D; ++d) {
        const T l = (T)hp.lengthscale[hp.ls_dim == 1 ? 0 : d];
        const T t = ((T)Z[i * D + d] - (T)Z[j * D + d]) / l;
        r2 += t * t;
      }
    } else {
      T x2 = 0, z2 = 0, xz = 0;
      for (int d = 0; d < D; ++d) {
        const T l = (T)hp.lengthscale[hp.ls_dim == 1 ? 0 : d];
        const T a = (T)Z[i * D + d] / l, b = (T)Z[j * D + d] / l;
        x2 += a * a;
        z2 += b * b;
        xz += a * b;
      }
      r2 = x2 - T(2) * xz + z2;
      r2 = r2 < T(0) ? T(0) : r2;
    }
    T k, dk;
    kernel_eval<T>(hp.kid, r2, k, dk);
    val = (T)hp.variance[0] * k;
    if (i == j) {
      if (sizeof(T) == 8) {
        double tot = 0.0, inc = jitter;
        for (int t = 0; t <= njitter; ++t, inc *= 10.0) tot += inc;
        val += (T)tot;
      } else {
        double inc = jitter;
        for (int t = 0; t <= njitter; ++t, inc *= 10.0) val += (T)inc;
      }
    }
  }
  Kuu[(long long)i * Mp + j] = val;
}

// ---------------------------------------------------------------------------------------------
// Blocked right-looking Cholesky (lower), in place.  status: 0 ok, c+1 = pivot of column c not > 0.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(NB * NB) k_chol_diag(T* __restrict__ A, int Mp, int kb, int* __restrict__ status) {
  __shared__ T a[NB][NB + 1];
  __shared__ int bad;
  const int c = threadIdx.x & (NB - 1), r = threadIdx.x / NB;
  T* blk = A + ((long long)kb * NB) * Mp + kb * NB;
  a[r][c] = blk[(long long)r * Mp + c];
  if (threadIdx.x == 0) bad = 0;
  __syncthreads();
  for (int j = 0; j < NB; ++j) {
    if (threadIdx.x == 0) {
      const T d = a[j][j];
      if (!(d > T(0))) {
        if (bad == 0) bad = kb * NB + j + 1;
        a[j][j] = T(1);
      } else {
        a[j][j] = sqrt(d);
      }
    }
    __syncthreads();
    if (c == j && r > j) a[r][j] /= a[j][j];
    __syncthreads();
    if (c > j && r >= c) a[r][c] -= a[r][j] * a[c][j];
    __syncthreads();
  }
  blk[(long long)r * Mp + c] = (c <= r) ? a[r][c] : T(0);
  if (threadIdx.x == 0 && bad != 0) atomicCAS(status, 0, bad);
}

// panel: A[ib][kb] <- A[ib][kb] * L_kk^-T   (one warp per 32-row block, thread = row)
template <typename T>
__global__ void __launch_bounds__(NB) k_chol_panel(T* __restrict__ A, int Mp, int kb) {
  __shared__ T l[NB][NB + 1];
  const int ib = kb + 1 + blockIdx.x;
  const T* lk = A + ((long long)kb * NB) * Mp + kb * NB;
  for (int t = threadIdx.x; t < NB * NB; t += NB) l[t / NB][t % NB] = lk[(long long)(t / NB) * Mp + (t % NB)];
  __syncthreads();
  T* row = A + ((long long)ib * NB + threadIdx.x) * Mp + kb * NB;
  T x[NB];
#pragma unroll
  for (int c = 0; c < NB; ++c) x[c] = row[c];
#pragma unroll
  for (int c = 0; c < NB; ++c) {
    T s = x[c];
#pragma unroll
    for (int t = 0; t < NB; ++t)
      if (t < c) s -= x[t] * l[c][t];
    x[c] = s / l[c][c];
  }
#pragma unroll
  for (int c = 0; c < NB; ++c) row[c] = x[c];
}

// trailing update: A[ib][jb] -= L[ib][kb] L[jb][kb]^T for ib >= jb > kb; also zeroes blocks above the diagonal
template <typename T>
__global__ void __launch_bounds__(NB * NB) k_chol_update(T* __restrict__ A, int Mp, int kb) {
  const int ib = kb + 1 + blockIdx.y, jb = kb + 1 + blockIdx.x;
  if (jb > ib) return;
  __shared__ T li[NB][NB + 1], lj[NB][NB + 1];
  const int c = threadIdx.x & (NB - 1), r = threadIdx.x / NB;
  li[r][c] = A[((long long)ib * NB + r) * Mp + kb * NB + c];
  lj[r][c] = A[((long long)jb * NB + r) * Mp + kb * NB + c];
  __syncthreads();
  T s = 0;
#pragma unroll
  for (int t = 0; t < NB; ++t) s += li[r][t] * lj[c][t];
  A[((long long)ib * NB + r) * Mp + jb * NB + c] -= s;
}

template <typename T>
__global__ void k_zero_upper(T* __restrict__ A, int Mp) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
  if (j < Mp && j > i) A[(long long)i * Mp + j] = T(0);
}

template <typename T>
inline void cholesky_inplace(T* A, int Mp, int* status, cudaStream_t st) {
  const int nblk = Mp / NB;
  for (int kb = 0; kb < nblk; ++kb) {
    k_chol_diag<T><<<1, NB * NB, 0, st>>>(A, Mp, kb, status);
    const int rem = nblk - kb - 1;
    if (rem > 0) {
      k_chol_panel<T><<<rem, NB, 0, st>>>(A, Mp, kb);
      k_chol_update<T><<<dim3(rem, rem), NB * NB, 0, st>>>(A, Mp, kb);
    }
  }
  k_zero_upper<T><<<dim3(ceil_div(Mp, 256), Mp), 256, 0, st>>>(A, Mp);
}

// ---------------------------------------------------------------------------------------------
// X = L^-1 (lower), blocked forward substitution by block rows.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NB) k_trinv_diag(const double* __restrict__ L, double* __restrict__ X, int Mp) {
  __shared__ double l[NB][NB + 1];
  const int b = blockIdx.x;
  const double* lb = L + ((long long)b * NB) * Mp + b * NB;
  for (int t = threadIdx.x; t < NB * NB; t += NB) l[t / NB][t % NB] = lb[(long long)(t / NB) * Mp + (t % NB)];
  __syncthreads();
  const int c = threadIdx.x;   // column of the inverse
  double x[NB];
#pragma unroll
  for (int r = 0; r < NB; ++r) {
    double s = (r == c) ? 1.0 : 0.0;
#pragma unroll
    for (int t = 0; t < NB; ++t)
      if (t < r) s -= l[r][t] * x[t];
    x[r] = s / l[r][r];
  }
  double* xb = X + ((long long)b * NB) * Mp + b * NB;
#pragma unroll
  for (int r = 0; r < NB; ++r) xb[(long long)r * Mp + c] = (r >= c) ? x[r] : 0.0;
}

// X[ib][jb] = -X[ib][ib] * sum_{t=jb}^{ib-1} L[ib][t] X[t][jb]     grid = ib blocks (jb = blockIdx.x)
__global__ void __launch_bounds__(NB * NB) k_trinv_row(const double* __restrict__ L, double* __restrict__ X, int Mp,
                                                       int ib) {
  __shared__ double a[NB][NB + 1], b[NB][NB + 1];
  const int jb = blockIdx.x;
  const int c = threadIdx.x & (NB - 1), r = threadIdx.x / NB;
  double s = 0.0;
  for (int t = jb; t < ib; ++t) {
    __syncthreads();
    a[r][c] = L[((long long)ib * NB + r) * Mp + t * NB + c];
    b[r][c] = X[((long long)t * NB + r) * Mp + jb * NB + c];
    __syncthreads();
#pragma unroll
    for (int e = 0; e < NB; ++e) s += a[r][e] * b[e][c];
  }
  __syncthreads();
  a[r][c] = X[((long long)ib * NB + r) * Mp + ib * NB + c];
  b[r][c] = s;
  __syncthreads();
  double o = 0.0;
#pragma unroll
  for (int e = 0; e < NB; ++e) o += a[r][e] * b[e][c];
  X[((long long)ib * NB + r) * Mp + jb * NB + c] = -o;
}

inline void tri_inverse(const double* L, double* X, int Mp, cudaStream_t st) {
  cudaMemsetAsync(X, 0, sizeof(double) * (size_t)Mp * Mp, st);
  const int nblk = Mp / NB;
  k_trinv_diag<<<nblk, NB, 0, st>>>(L, X, Mp);
  for (int ib = 1; ib < nblk; ++ib) k_trinv_row<<<ib, NB * NB, 0, st>>>(L, X, Mp, ib);
}

// ---------------------------------------------------------------------------------------------
// C = op(A) * op(B), square [Mp][Mp] fp64, 64 x 64 tile per CTA, 4 x 4 per thread.
// ---------------------------------------------------------------------------------------------
template <bool TA, bool TB>
__global__ void __launch_bounds__(256) k_dgemm(const double* __restrict__ A, const double* __restrict__ B,
                                               double* __restrict__ C, int Mp) {
  __shared__ double as[16][65], bs[16][65];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int i0 = blockIdx.y * 64, j0 = blockIdx.x * 64;
  double acc[4][4] = {};
  for (int k0 = 0; k0 < Mp; k0 += 16) {
    __syncthreads();
    for (int t = threadIdx.x; t < 16 * 64; t += 256) {
      const int kk = t >> 6, e = t & 63;
      as[kk][e] = TA ? A[(long long)(k0 + kk) * Mp + i0 + e] : A[(long long)(i0 + e) * Mp + k0 + kk];
      bs[kk][e] = TB ? B[(long long)(j0 + e) * Mp + k0 + kk] : B[(long long)(k0 + kk) * Mp + j0 + e];
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      double a[4], b[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        a[u] = as[kk][ty * 4 + u];
        b[u] = bs[kk][tx * 4 + u];
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) acc[u][v] += a[u] * b[v];
    }
  }
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int v = 0; v < 4; ++v) C[(long long)(i0 + ty * 4 + u) * Mp + j0 + tx * 4 + v] = acc[u][v];
}

template <bool TA, bool TB>
inline void dgemm(const double* A, const double* B, double* C, int Mp, cudaStream_t st) {
  k_dgemm<TA, TB><<<dim3(Mp / 64, Mp / 64), 256, 0, st>>>(A, B, C, Mp);
}

// mode 0: X <- -tril(X)      mode 1: X <- tril(X) with the diagonal halved  (the Phi of the Cholesky adjoint)
__global__ void k_tril_op(double* __restrict__ X, int Mp, int mode) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
  if (j >= Mp) return;
  double v = X[(long long)i * Mp + j];
  if (j > i) v = 0.0;
  else if (mode == 0) v = -v;
  else if (j == i) v *= 0.5;
  X[(long long)i * Mp + j] = v;
}

// ---------------------------------------------------------------------------------------------
// Kuu adjoint -> (Z, lengthscale, variance).  GK is the un-symmetrised L^-T Phi L^-1; the symmetric part
// sym = (GK + GK^T)/2 is the gradient w.r.t. every entry of Kuu (torch cholesky_backward).
// one thread per row a, 128 threads per CTA.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_kuu_backward(const double* __restrict__ GK, int Mp, const float* __restrict__ Z,
                                                      int M, Hyper hp, double* __restrict__ dz_acc,
                                                      double* __restrict__ acc) {
  __shared__ double scratch[32];
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  const int D = hp.D;
  double dz[MAX_D], dl[MAX_D], dv = 0.0;
  for (int d = 0; d < D; ++d) dz[d] = dl[d] = 0.0;
  if (a < M) {
    const double var = hp.variance[0];
    double za[MAX_D], il[MAX_D];
    for (int d = 0; d < D; ++d) {
      za[d] = Z[a * D + d];
      il[d] = 1.0 / (double)hp.lengthscale[hp.ls_dim == 1 ? 0 : d];
    }
    for (int j = 0; j < M; ++j) {
      const double g = 0.5 * (GK[(long long)a * Mp + j] + GK[(long long)j * Mp + a]);
      double diff[MAX_D], r2 = 0.0;
      for (int d = 0; d < D; ++d) {
        diff[d] = (za[d] - (double)Z[j * D + d]) * il[d];
        r2 += diff[d] * diff[d];
      }
      double k, dk;
      kernel_eval<double>(hp.kid, r2, k, dk);
      dv += g * k;
      const double h = g * var * dk;
      for (int d = 0; d < D; ++d) {
        dz[d] += 4.0 * h * diff[d] * il[d];          // both arguments of k(z_a, z_j) move with z_a
        dl[d] += -2.0 * h * diff[d] * diff[d] * il[d];
      }
    }
    for (int d = 0; d < D; ++d) atomicAdd(&dz_acc[a * D + d], dz[d]);
  }
  dv = block_sum(dv, scratch);
  if (threadIdx.x == 0) atomicAdd(&acc[ACC_DVAR], dv);
  if (hp.ls_dim == 1) {
    double s = 0.0;
    for (int d = 0; d < D; ++d) s += dl[d];
    s = block_sum(s, scratch);
    if (threadIdx.x == 0) atomicAdd(&acc[ACC_DLS], s);
  } else {
    for (int d = 0; d < D; ++d) {
      double s = block_sum(dl[d], scratch);
      if (threadIdx.x == 0) atomicAdd(&acc[ACC_DLS + d], s);
    }
  }
}

}  // namespace gdrf
