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
__device__ __forceinline__ void kuu_body(const float* __restrict__ Z, int M, int Mp, const Hyper& hp, double jitter,
                                         int njitter, T* __restrict__ Kuu) {
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
    kernel_eval<T>(hp.kid, r2, k, dk, hp.kid == KERNEL_RQ ? (T)hp.alpha[0] : T(1));
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
template <typename T>
__global__ void k_kuu(const float* __restrict__ Z, int M, int Mp, Hyper hp, double jitter, int njitter,
                      T* __restrict__ Kuu) {
  kuu_body<T>(Z, M, Mp, hp, jitter, njitter, Kuu);
}

// ---------------------------------------------------------------------------------------------
// Blocked right-looking Cholesky (lower), out of place: A holds Kuu and receives the trailing updates, L the
// factor, Dinv the inverses of L's 32 x 32 diagonal blocks.  Two launches per block column:
//   k_chol_diag : 32 warps, one thread per element of the block, columns exchanged through shared memory; also
//                 inverts the triangular block.  status: 0 ok, c+1 = pivot of column c not > 0 (the
//                 RuntimeError of torch.linalg.cholesky that gdrf/models/utils.py:31-37 catches).
//   k_chol_step : block (ib, jb), ib >= jb > kb: panel tiles P_i = A[ib][kb] Dinv^T, P_j likewise, then
//                 A[ib][jb] -= P_i P_j^T; the jb == kb+1 blocks also store L[ib][kb] = P_i; block (kb+1, kb+1) goes on
//                 to factorise its tile (look-ahead), so a block column costs ONE launch.
// ---------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void chol_diag_core(T a, T (*sl)[NB + 1], T* __restrict__ L, T* __restrict__ Dinv, int Mp,
                                               int kb, int* __restrict__ status) {
  // 32 warps, one thread per element (input and output map: r = warp, c = lane).  Three single-warp formulations (all
  // shuffles and fully unrolled; all shared memory; row in registers) took 29-35 us per block column: one warp issuing
  // ~14 K mostly dependent instructions.  Here a column step is one block barrier and one multiply-add per thread, and a
  // column of the inverse is one warp's forward substitution.  The factorisation's arithmetic and its order are
  // unchanged (right-looking: element (r, c) receives -L[r][j] L[c][j] for j = 0 .. c-1 in order).
  // sl: [NB][NB + 1] shared scratch of the caller for L (zero above the diagonal); first written after a block barrier
  __shared__ T scol[2][NB];        // the column being eliminated (double-buffered: one block barrier per column)
  __shared__ T sinv[NB];           // 1 / L[j][j]
  const int r = threadIdx.x >> 5, c = threadIdx.x & 31;
  // The elimination runs on the TRANSPOSED thread map (thread = element (row = lane, column = warp)): column j then
  // lives in ONE warp, which alone takes the pivot's reciprocal square root (every thread doing so -- 32 warps x ~40 fp64
  // instructions per column on a 64-lane fp64 pipe -- was what a column step cost: ~1200 clk), scales its column and
  // publishes it; the other warps wait at the single barrier of the step and apply the rank-1 update.
  __syncthreads();                 // the caller may still be reading sl (chol_trail_body's ai)
  sl[r][c] = a;
  __syncthreads();
  const int rr = c, cc = r;        // lane, warp
  T at = sl[rr][cc];
#pragma unroll 1
  for (int j = 0; j < NB; ++j) {
    if (cc == j) {                 // warp-uniform
      T d = __shfl_sync(0xffffffffu, at, j);
      if (!(d > T(0))) {
        if (rr == 0) atomicCAS(status, 0, kb * NB + j + 1);
        d = T(1);
      }
      // one reciprocal square root per column instead of two square roots and a division (each a long dependent
      // chain in fp64); a Newton step puts sqrt(d) = d * inv back within an ulp
      const T inv = rsqrt(d);
      T sq = d * inv;
      sq = fma(T(0.5) * inv, fma(-sq, sq, d), sq);
      const T lj = (rr == j) ? sq : (rr > j ? at * inv : T(0));
      at = lj;
      scol[j & 1][rr] = lj;
      if (rr == j) sinv[j] = inv;
    }
    __syncthreads();
    if (cc > j && rr >= cc) at -= scol[j & 1][rr] * scol[j & 1][cc];
  }
  __syncthreads();                 // sl was read by everyone long ago; scol/sinv complete
  sl[rr][cc] = (cc <= rr) ? at : T(0);
  __syncthreads();
  L[((long long)kb * NB + r) * Mp + kb * NB + c] = sl[r][c];
  // the inverse X = L^-1, one column per warp on the transposed map: forward substitution L x = e_cc in its
  // right-looking form -- lane i publishes x_i = b_i / L[i][i] with one shuffle and every lane below it updates its
  // right-hand side b -= L[rr][i] x_i (a shuffle and a multiply-add per step instead of a five-level shuffle tree)
  T b = (rr == cc) ? T(1) : T(0), x = T(0);
#pragma unroll 1
  for (int i = cc; i < NB; ++i) {
    const T xi = __shfl_sync(0xffffffffu, b, i) * sinv[i];
    if (rr == i) x = xi;
    if (rr > i) b -= sl[rr][i] * xi;
  }
  Dinv[((long long)kb * NB + rr) * NB + cc] = x;
}

template <typename T>
__device__ __forceinline__ void chol_diag_body(const T* __restrict__ A, T* __restrict__ L, T* __restrict__ Dinv,
                                               int Mp, int kb, int* __restrict__ status) {
  __shared__ T sl[NB][NB + 1];
  const int r = threadIdx.x >> 5, c = threadIdx.x & 31;
  chol_diag_core<T>(A[((long long)kb * NB + r) * Mp + kb * NB + c], sl, L, Dinv, Mp, kb, status);
}

template <typename T>
__global__ void __launch_bounds__(NB * NB) k_chol_diag(const T* __restrict__ A, T* __restrict__ L, T* __restrict__ Dinv,
                                                  int Mp, int kb, int* __restrict__ status) {
  chol_diag_body<T>(A, L, Dinv, Mp, kb, status);
}

// the fp64 factorisation (values) and the fp32 one (the reference's "did it fail" decision) side by side:
// block 0 = double, block 1 = float
__global__ void __launch_bounds__(NB * NB) k_chol_diag_both(const double* __restrict__ A, double* __restrict__ L,
                                                       double* __restrict__ Dinv, const float* __restrict__ Af,
                                                       float* __restrict__ Lf, float* __restrict__ Dinvf, int Mp, int kb,
                                                       int* __restrict__ status) {
  if (blockIdx.x == 0) chol_diag_body<double>(A, L, Dinv, Mp, kb, status);
  else chol_diag_body<float>(Af, Lf, Dinvf, Mp, kb, status);
}

// LOOKAHEAD: the block that updates the next diagonal tile (kb+1, kb+1) goes straight on to factorise it (the updated
// element is in a register), while the other blocks of the launch finish the trailing update: one launch per block
// column instead of two, and the 20 us factorisation hides the 15 us update instead of following it.
template <typename T, bool LOOKAHEAD>
__device__ __forceinline__ void chol_trail_body(T* __restrict__ A, T* __restrict__ L, T* __restrict__ Dinv,
                                                int Mp, int kb, int* __restrict__ status) {
  const int ib = kb + 1 + blockIdx.y, jb = kb + 1 + blockIdx.x;
  if (jb > ib) return;
  __shared__ T ai[NB][NB + 1], aj[NB][NB + 1], dd[NB][NB + 1];
  const int c = threadIdx.x & (NB - 1), r = threadIdx.x / NB;
  ai[r][c] = A[((long long)ib * NB + r) * Mp + kb * NB + c];
  aj[r][c] = A[((long long)jb * NB + r) * Mp + kb * NB + c];
  dd[r][c] = Dinv[((long long)kb * NB + r) * NB + c];
  __syncthreads();
  T pi = 0, pj = 0;
#pragma unroll
  for (int t = 0; t < NB; ++t) {
    pi += ai[r][t] * dd[c][t];
    pj += aj[r][t] * dd[c][t];
  }
  __syncthreads();
  ai[r][c] = pi;
  aj[r][c] = pj;
  if (jb == kb + 1) L[((long long)ib * NB + r) * Mp + kb * NB + c] = pi;
  __syncthreads();
  T s = 0;
#pragma unroll
  for (int t = 0; t < NB; ++t) s += ai[r][t] * aj[c][t];
  const T a_new = A[((long long)ib * NB + r) * Mp + jb * NB + c] - s;
  A[((long long)ib * NB + r) * Mp + jb * NB + c] = a_new;
  if (LOOKAHEAD && blockIdx.x == 0 && blockIdx.y == 0)      // block-uniform; ai is dead after the first barrier inside
    chol_diag_core<T>(a_new, ai, L, Dinv, Mp, kb + 1, status);
}

template <typename T>
__global__ void __launch_bounds__(NB * NB) k_chol_step(T* __restrict__ A, T* __restrict__ L, T* __restrict__ Dinv, int Mp,
                                                       int kb, int* __restrict__ status) {
  chol_trail_body<T, true>(A, L, Dinv, Mp, kb, status);
}

__global__ void __launch_bounds__(NB * NB) k_chol_step_both(double* __restrict__ A, double* __restrict__ L,
                                                            double* __restrict__ Dinv, float* __restrict__ Af,
                                                            float* __restrict__ Lf, float* __restrict__ Dinvf, int Mp,
                                                            int kb, int* __restrict__ status) {
  if (blockIdx.z == 0) chol_trail_body<double, true>(A, L, Dinv, Mp, kb, status);
  else chol_trail_body<float, true>(Af, Lf, Dinvf, Mp, kb, status);
}

// Batched status probe (jittercholesky's "does this level fail?", utils.py:31-37, for SEVERAL jitter levels at once):
// blockIdx.z / blockIdx.x selects the level; every level has its own fp32 Kuu, L, Dinv and status word, `stride`
// (matrices), `dstride` (Dinv) elements apart.  The levels are independent, so the batch is ONE launch chain -- latency-bound
// on the 32 x 32 diagonal factorisations for small M -- instead of one chain and one host read-back per level.
__global__ void __launch_bounds__(NB * NB) k_chol_diag_batch(const float* __restrict__ A, float* __restrict__ L,
                                                             float* __restrict__ Dinv, int Mp, int kb, long long stride,
                                                             long long dstride, int* __restrict__ status) {
  const long long b = blockIdx.x;
  chol_diag_body<float>(A + b * stride, L + b * stride, Dinv + b * dstride, Mp, kb, status + b);
}
__global__ void __launch_bounds__(NB * NB) k_chol_step_batch(float* __restrict__ A, float* __restrict__ L,
                                                             float* __restrict__ Dinv, int Mp, int kb, long long stride,
                                                             long long dstride, int* __restrict__ status) {
  const long long b = blockIdx.z;
  chol_trail_body<float, true>(A + b * stride, L + b * stride, Dinv + b * dstride, Mp, kb, status + b);
}
// Kuu in the reference's fp32 arithmetic for `count` consecutive jitter levels nj0, nj0 + 1, ... (blockIdx.z)
__global__ void k_kuu_batch(const float* __restrict__ Z, int M, int Mp, Hyper hp, double jitter, int nj0,
                            float* __restrict__ Kuu, long long stride) {
  kuu_body<float>(Z, M, Mp, hp, jitter, nj0 + (int)blockIdx.z, Kuu + (long long)blockIdx.z * stride);
}
inline void cholesky_batch(float* A, float* L, float* Dinv, int Mp, int count, int* status, cudaStream_t st) {
  const long long stride = (long long)Mp * Mp, dstride = (long long)Mp * NB;
  cudaMemsetAsync(L, 0, sizeof(float) * (size_t)stride * count, st);
  const int nblk = Mp / NB;
  k_chol_diag_batch<<<count, NB * NB, 0, st>>>(A, L, Dinv, Mp, 0, stride, dstride, status);
  for (int kb = 0; kb + 1 < nblk; ++kb) {
    const int rem = nblk - kb - 1;
    k_chol_step_batch<<<dim3(rem, rem, count), NB * NB, 0, st>>>(A, L, Dinv, Mp, kb, stride, dstride, status);
  }
}

// A: Kuu (destroyed), L: factor (zero above the diagonal), Dinv: [Mp/32][32][32]
template <typename T>
inline void cholesky(T* A, T* L, T* Dinv, int Mp, int* status, cudaStream_t st) {
  cudaMemsetAsync(L, 0, sizeof(T) * (size_t)Mp * Mp, st);
  const int nblk = Mp / NB;
  k_chol_diag<T><<<1, NB * NB, 0, st>>>(A, L, Dinv, Mp, 0, status);
  for (int kb = 0; kb + 1 < nblk; ++kb) {       // trailing update of block column kb + factorisation of tile kb + 1
    const int rem = nblk - kb - 1;
    k_chol_step<T><<<dim3(rem, rem), NB * NB, 0, st>>>(A, L, Dinv, Mp, kb, status);
  }
}

inline void cholesky_both(double* A, double* L, double* Dinv, float* Af, float* Lf, float* Dinvf, int Mp, int* status,
                          cudaStream_t st) {
  cudaMemsetAsync(L, 0, sizeof(double) * (size_t)Mp * Mp, st);
  cudaMemsetAsync(Lf, 0, sizeof(float) * (size_t)Mp * Mp, st);
  const int nblk = Mp / NB;
  k_chol_diag_both<<<2, NB * NB, 0, st>>>(A, L, Dinv, Af, Lf, Dinvf, Mp, 0, status);
  for (int kb = 0; kb + 1 < nblk; ++kb) {
    const int rem = nblk - kb - 1;
    k_chol_step_both<<<dim3(rem, rem, 2), NB * NB, 0, st>>>(A, L, Dinv, Af, Lf, Dinvf, Mp, kb, status);
  }
}

// ---------------------------------------------------------------------------------------------
// X = L^-1 (lower) by recursive doubling over the diagonal: the 32 x 32 diagonal blocks are already inverted (Dinv, from
// the factorisation); at level s = 32, 64, ... every 2s x 2s diagonal block [[L11, 0], [L21, L22]] gets its lower-left
// quarter X21 = -X22 (L21 X11) from the two s x s inverses of the level below.  Two batched launches per level,
// log2(Mp / 32) levels (10 launches at M = 1024) instead of the 2 Mp / 32 - 1 dependent launches of a block forward
// substitution.  Tm: [Mp][Mp] scratch for L21 X11.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NB * NB) k_trinv_diag(const double* __restrict__ Dinv, double* __restrict__ X, int Mp) {
  const int kb = blockIdx.x, c = threadIdx.x & (NB - 1), r = threadIdx.x / NB;
  X[((long long)kb * NB + r) * Mp + kb * NB + c] = Dinv[((long long)kb * NB + r) * NB + c];
}

// MODE 0: Tm21 = L21 X11 (X11 lower: k >= column);  MODE 1: X21 = -X22 Tm21 (X22 lower: k <= row)
template <int MODE>
__global__ void __launch_bounds__(NB * NB) k_trinv_level(const double* __restrict__ L, double* __restrict__ X,
                                                         double* __restrict__ Tm, int Mp, int s) {
  __shared__ double as[NB][NB + 1], bs[NB][NB + 1];
  const int c = threadIdx.x & (NB - 1), r = threadIdx.x / NB;
  const int tj = blockIdx.x, ti = blockIdx.y;
  // the pair merges the diagonal blocks [lo, mid) and [mid, hi); the last pair of a level may have a short (or no)
  // second block when Mp / 32 is not a power of two
  const long long lo = 2LL * s * blockIdx.z, mid = lo + s, hi = (lo + 2 * s < Mp) ? lo + 2 * s : Mp;
  if (mid + ti * NB >= hi) return;                 // block-uniform
  const long long row = mid + ti * NB + r, col = lo + tj * NB + c;
  const int k0 = MODE == 0 ? tj * NB : 0, k1 = MODE == 0 ? s : (ti + 1) * NB;
  const double* Am = MODE == 0 ? L : X;            // left factor: L21 / X22 (rows `row`)
  const long long acol0 = MODE == 0 ? lo : mid;
  const double* Bm = MODE == 0 ? X : Tm;           // right factor: X11 / Tm21 (columns `col`)
  const long long brow0 = MODE == 0 ? lo : mid;
  double acc = 0.0;
  for (int k = k0; k < k1; k += NB) {
    __syncthreads();
    as[r][c] = Am[row * Mp + acol0 + k + c];
    bs[r][c] = Bm[(brow0 + k + r) * Mp + col];
    __syncthreads();
#pragma unroll
    for (int t = 0; t < NB; ++t) acc += as[r][t] * bs[t][c];
  }
  if (MODE == 0) Tm[row * Mp + col] = acc;
  else X[row * Mp + col] = -acc;
}

inline int tri_inverse(const double* L, const double* Dinv, double* X, double* Tm, int Mp, cudaStream_t st) {
  cudaMemsetAsync(X, 0, sizeof(double) * (size_t)Mp * Mp, st);
  k_trinv_diag<<<Mp / NB, NB * NB, 0, st>>>(Dinv, X, Mp);
  int launches = 1;
  for (int s = NB; s < Mp; s *= 2) {
    const dim3 grid(s / NB, s / NB, (Mp + 2 * s - 1) / (2 * s));
    k_trinv_level<0><<<grid, NB * NB, 0, st>>>(L, X, Tm, Mp, s);
    k_trinv_level<1><<<grid, NB * NB, 0, st>>>(L, X, Tm, Mp, s);
    launches += 2;
  }
  return launches;
}

// ---------------------------------------------------------------------------------------------
// C = op(A) * op(B), square [Mp][Mp] fp64, 128 x 128 tile per CTA, 8 x 8 per thread.
// The Cholesky adjoint only multiplies triangular matrices: KSTART says where the contraction index can start for
// a tile (0: 0, 1: the tile's first row, 2: its first column, 3: the larger of the two) and LOWER_ONLY skips the
// tiles strictly above the diagonal (their values are discarded by the caller's tril).  The contraction range of a
// tile is cut into gridDim.z pieces that add into a zeroed C (a single CTA walking all of it is a 64-step chain of
// load -> barrier -> 256 dependent fp64 FMAs per thread, which is what the kernel's duration used to be).
// ---------------------------------------------------------------------------------------------
template <bool TA, bool TB, int KSTART, bool LOWER_ONLY>
__global__ void __launch_bounds__(256) k_dgemm(const double* __restrict__ A, const double* __restrict__ B,
                                               double* __restrict__ C, int Mp) {
  // 128 x 128 tile per CTA, 8 x 8 per thread (rows ty + 16 u, columns tx + 16 v: a warp reads 16 consecutive doubles of
  // each operand row, conflict-free), k in steps of 8.  One shared-memory load per four multiply-adds: the 64 x 64 /
  // 4 x 4 version (one per two) was bound by shared-memory bandwidth at ~5.7 TFLOP/s.
  constexpr int BM = 128, BK = 8, LD = BM + 2;
  __shared__ double as[BK][LD], bs[BK][LD];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int i0 = blockIdx.y * BM, j0 = blockIdx.x * BM;
  if (LOWER_ONLY && j0 > i0) return;
  const int kfirst = KSTART == 0 ? 0 : (KSTART == 1 ? i0 : (KSTART == 2 ? j0 : max(i0, j0)));
  const int steps = (Mp - kfirst) / BK, per = (steps + gridDim.z - 1) / gridDim.z;
  const int kbeg = kfirst + BK * per * blockIdx.z, kend = min(Mp, kbeg + BK * per);
  if (kbeg >= kend) return;
  double acc[8][8] = {};
  for (int k0 = kbeg; k0 < kend; k0 += BK) {
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int idx = threadIdx.x + 256 * q;
      if (TA) { const int e = idx & (BM - 1), kk = idx >> 7; as[kk][e] = A[(long long)(k0 + kk) * Mp + i0 + e]; }
      else    { const int kk = idx & (BK - 1), e = idx >> 3; as[kk][e] = A[(long long)(i0 + e) * Mp + k0 + kk]; }
      if (TB) { const int kk = idx & (BK - 1), e = idx >> 3; bs[kk][e] = B[(long long)(j0 + e) * Mp + k0 + kk]; }
      else    { const int e = idx & (BM - 1), kk = idx >> 7; bs[kk][e] = B[(long long)(k0 + kk) * Mp + j0 + e]; }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      double a[8], b[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        a[u] = as[kk][ty + 16 * u];
        b[u] = bs[kk][tx + 16 * u];
      }
#pragma unroll
      for (int u = 0; u < 8; ++u)
#pragma unroll
        for (int v = 0; v < 8; ++v) acc[u][v] = fma(a[u], b[v], acc[u][v]);
    }
  }
#pragma unroll
  for (int u = 0; u < 8; ++u)
#pragma unroll
    for (int v = 0; v < 8; ++v) atomicAdd(&C[(long long)(i0 + ty + 16 * u) * Mp + j0 + tx + 16 * v], acc[u][v]);
}

template <bool TA, bool TB, int KSTART = 0, bool LOWER_ONLY = false>
inline void dgemm(const double* A, const double* B, double* C, int Mp, cudaStream_t st) {
  cudaMemsetAsync(C, 0, sizeof(double) * (size_t)Mp * Mp, st);
  // split the contraction so that the tiles (half of them skipped when only the lower part is wanted) fill the SMs
  const int tiles = (Mp / 128) * (Mp / 128);
  int z = (2 * 148 + tiles - 1) / tiles;
  z = z < 1 ? 1 : (z > 16 ? 16 : z);
  k_dgemm<TA, TB, KSTART, LOWER_ONLY><<<dim3(Mp / 128, Mp / 128, z), 256, 0, st>>>(A, B, C, Mp);
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
  const int a = blockIdx.x;       // one block per row a; threads stride over j
  const int D = hp.D;
  double dz[MAX_D], dl[MAX_D], dv = 0.0, da = 0.0;
  const double alpha = hp.kid == KERNEL_RQ ? (double)hp.alpha[0] : 1.0;
  double za[MAX_D], il[MAX_D];
  for (int d = 0; d < D; ++d) {
    dz[d] = dl[d] = 0.0;
    za[d] = Z[a * D + d];
    il[d] = 1.0 / (double)hp.lengthscale[hp.ls_dim == 1 ? 0 : d];
  }
  const double var = hp.variance[0];
  for (int j = threadIdx.x; j < M; j += blockDim.x) {
    const double g = 0.5 * (GK[(long long)a * Mp + j] + GK[(long long)j * Mp + a]);
    double diff[MAX_D], r2 = 0.0;
    for (int d = 0; d < D; ++d) {
      diff[d] = (za[d] - (double)Z[j * D + d]) * il[d];
      r2 += diff[d] * diff[d];
    }
    double k, dk, dka = 0.0;
    kernel_eval<double>(hp.kid, r2, k, dk, alpha, &dka);
    dv += g * k;
    da += g * var * dka;
    const double h = g * var * dk;
    for (int d = 0; d < D; ++d) {
      dz[d] += 4.0 * h * diff[d] * il[d];          // both arguments of k(z_a, z_j) move with z_a
      dl[d] += -2.0 * h * diff[d] * diff[d] * il[d];
    }
  }
  for (int d = 0; d < D; ++d) {
    const double t = block_sum(dz[d], scratch);
    if (threadIdx.x == 0) atomicAdd(&dz_acc[a * D + d], t);
  }
  dv = block_sum(dv, scratch);
  if (threadIdx.x == 0) atomicAdd(&acc[ACC_DVAR], dv);
  if (hp.kid == KERNEL_RQ) {
    da = block_sum(da, scratch);
    if (threadIdx.x == 0) atomicAdd(&acc[ACC_DALPHA], da);
  }
  if (hp.ls_dim == 1) {
    double t = 0.0;
    for (int d = 0; d < D; ++d) t += dl[d];
    t = block_sum(t, scratch);
    if (threadIdx.x == 0) atomicAdd(&acc[ACC_DLS], t);
  } else {
    for (int d = 0; d < D; ++d) {
      const double t = block_sum(dl[d], scratch);
      if (threadIdx.x == 0) atomicAdd(&acc[ACC_DLS + d], t);
    }
  }
}

}  // namespace gdrf
