// Shared device/host helpers for the gdrf_b200 CUDA path (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace gdrf {

typedef __nv_bfloat16 bf16;

// ------------------------------------------------------------------------------------------
// "Tiled plane" operand format.
//
// Every matrix that feeds the tensor pipe is stored as P bf16 planes (x ~= p0 + p1 [+ p2], an
// error-compensated split of the fp32 value) and each plane is cut into 128-row x 64-column
// blocks of 16 KB.  Inside a block the bytes are laid out exactly as a SWIZZLE_128B UMMA
// shared-memory tile expects them (rows of 128 B, 8-row groups of 1 KB, 16-byte chunk index
// XORed with row%8), so a block -- or its upper / lower 64-row half -- goes HBM -> SMEM with one
// cp.async.bulk and no tensor map.  The same block serves as a K-major operand (contraction
// along its 64 columns) and as an MN-major operand (contraction along its rows).
// Blocks are ordered [row_tile][col_block].
// ------------------------------------------------------------------------------------------
constexpr int TILE_R = 128;
constexpr int TILE_C = 64;
constexpr int TILE_ELEMS = TILE_R * TILE_C;   // 8192 bf16 = 16 KB

__host__ __device__ __forceinline__ int tile_off(int r, int c) {
  return ((r >> 3) << 9) + ((r & 7) << 6) + ((((c >> 3) ^ (r & 7)) << 3) | (c & 7));
}

struct PlaneMat {
  bf16* base;            // plane 0
  long long plane_stride;  // elements between planes
  int row_tiles;         // rows / 128
  int col_blocks;        // cols / 64
  __host__ __device__ __forceinline__ long long block_off(int rt, int cb) const {
    return ((long long)rt * col_blocks + cb) * TILE_ELEMS;
  }
  __device__ __forceinline__ bf16* elem(int plane, int r, int c) const {
    return base + plane * plane_stride + block_off(r >> 7, c >> 6) + tile_off(r & 127, c & 63);
  }
};

// error-compensated split of fp32 values into up to three bf16 planes:
// 8 consecutive fp32 values -> one 16-byte packet per plane.
// Two values are converted at a time with the packed cvt.rn.bf16x2.f32 (F2FP.PACK_AB, full ALU rate); the scalar
// cvt.rn.bf16.f32 is an F2F on the quarter-rate conversion pipe and made the plane-writing epilogues XU-bound.
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
template <int P>
__device__ __forceinline__ void split8(const float* v, uint4 (&pk)[P]) {
  uint32_t w[P][4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float r0 = v[2 * i], r1 = v[2 * i + 1];
#pragma unroll
    for (int p = 0; p < P; ++p) {
      const uint32_t d = pack_bf16x2(r0, r1);
      w[p][i] = d;
      if (p + 1 < P) {
        r0 -= __uint_as_float(d << 16);
        r1 -= __uint_as_float(d & 0xffff0000u);
      }
    }
  }
#pragma unroll
  for (int p = 0; p < P; ++p) pk[p] = make_uint4(w[p][0], w[p][1], w[p][2], w[p][3]);
}

// 8 consecutive fp32 values -> one 16-byte packet per fp16 plane (x ~= h0 + h1, 22 significant bits for
// |x| in the normal fp16 range; the caller guarantees |x| < 65504)
template <int P>
__device__ __forceinline__ void split8h(const float* v, uint4 (&pk)[P]) {
  uint32_t w[P][4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float r0 = v[2 * i], r1 = v[2 * i + 1];
#pragma unroll
    for (int p = 0; p < P; ++p) {
      const uint32_t d = pack_f16x2(r0, r1);
      w[p][i] = d;
      if (p + 1 < P) {
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&d));
        r0 -= f.x;
        r1 -= f.y;
      }
    }
  }
#pragma unroll
  for (int p = 0; p < P; ++p) pk[p] = make_uint4(w[p][0], w[p][1], w[p][2], w[p][3]);
}

enum { FMT_F16 = 0, FMT_BF16 = 1 };   // tcgen05 kind::f16 operand formats (instruction descriptor encoding)

// two-plane split in the format chosen at run time (FMT_F16: 22 significant bits, |x| < 65504; FMT_BF16: 16 bits, any
// fp32 range) -- the backward operands follow the forward's format (GDRF_FLAG_FWD_BF16)
__device__ __forceinline__ void split8x2(int fmt, const float* v, uint4 (&pk)[2]) {
  if (fmt == FMT_F16) split8h<2>(v, pk);
  else split8<2>(v, pk);
}

// Power-of-two scale s with s * bound in [2^13, 2^14): operands written as fp16 planes are multiplied by s (exact) so
// that they sit well inside the fp16 range whatever the magnitude of the gradients flowing through them; the
// consuming contraction's epilogue multiplies by 1/s.  Pure bit arithmetic, so every kernel that recomputes it from the
// same bound gets the same value.  bound <= 0 (an all-zero chunk) gives s = 1.
__device__ __forceinline__ float pow2_scale(float bound, float* inv) {
  const int eb = (int)((__float_as_uint(bound) >> 23) & 0xff);
  if (!(bound > 0.f) || eb == 0 || eb == 255) { *inv = 1.f; return 1.f; }
  int se = 13 - (eb - 127);                 // 2^(eb-127) <= bound < 2^(eb-126)
  se = se < -100 ? -100 : (se > 100 ? 100 : se);
  *inv = __uint_as_float((uint32_t)(127 - se) << 23);
  return __uint_as_float((uint32_t)(127 + se) << 23);
}

// per-chunk scalars (zeroed together with q at the start of every chunk) and per-step scalars (zeroed by the prologue /
// at the start of a step): running maxima kept as float bit patterns, and the power-of-two operand scales derived from them
enum { CS_G2MAX = 0, CS_GLOCMAX = 1, CS_GV0MAX = 2, CS_DWMAX = 3, CS_SG_INV = 4, CS_SD_INV = 5, CS_SL_INV = 6, CS_COUNT = 64 };
enum { PS_SMAX = 0, PS_LINVMAX = 1, PS_UMAX = 2, PS_SU_INV = 3, PS_COUNT = 64 };

// running maximum of non-negative floats kept as their bit patterns (ordered like unsigned integers)
__device__ __forceinline__ void atomic_max_abs(unsigned* slot, float absval) {
  atomicMax(slot, __float_as_uint(absval));
}

__device__ __forceinline__ float plane_to_float(const bf16* p, int fmt) {
  return fmt == FMT_BF16 ? __bfloat162float(*p) : __half2float(*reinterpret_cast<const __half*>(p));
}

__device__ __forceinline__ float bf16_bits_to_float(unsigned short h) {
  return __uint_as_float(((uint32_t)h) << 16);
}

// sum of P planes for 8 consecutive elements (16-byte packets)
template <int P>
__device__ __forceinline__ void join8(const uint4 (&pk)[P], float* v) {
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = 0.f;
#pragma unroll
  for (int p = P - 1; p >= 0; --p) {   // small planes first
    uint32_t w[4] = {pk[p].x, pk[p].y, pk[p].z, pk[p].w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[2 * i] += bf16_bits_to_float((unsigned short)(w[i] & 0xffff));
      v[2 * i + 1] += bf16_bits_to_float((unsigned short)(w[i] >> 16));
    }
  }
}

// ------------------------------------------------------------------------------------------
// reductions
// ------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum; result valid in thread 0.  `scratch` holds >= 32 T.
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* scratch) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  T r = (threadIdx.x < nw) ? scratch[threadIdx.x] : T(0);
  if (warp == 0) r = warp_sum(r);
  return r;
}

__host__ __device__ __forceinline__ long long round_up_ll(long long a, long long b) { return (a + b - 1) / b * b; }
__host__ __device__ __forceinline__ int ceil_div(int a, int b) { return (a + b - 1) / b; }

// kernel ids shared with the host API (include/gdrf_b200.h)
enum { KERNEL_RBF = 0, KERNEL_MATERN32 = 1, KERNEL_MATERN52 = 2, KERNEL_EXPONENTIAL = 3, KERNEL_RQ = 4 };
constexpr int MAX_D = 8;

// k(r2)/variance and d k / d r2 / variance for the isotropic kernels
// (pyro.contrib.gp.kernels.{RBF,Matern32,Matern52,Exponential,RationalQuadratic}; r = sqrt(r2 + 1e-12) as in
// Isotropy._scaled_dist).  RationalQuadratic: (1 + r2 / (2 alpha))^(-alpha); *dk_dalpha is written for it only.
template <typename T>
__host__ __device__ __forceinline__ void kernel_eval(int kid, T r2, T& k, T& dk_dr2, T alpha = T(1),
                                                     T* dk_dalpha = nullptr) {
  if (kid == KERNEL_RBF) {
    k = exp(T(-0.5) * r2);
    dk_dr2 = T(-0.5) * k;
  } else if (kid == KERNEL_MATERN32) {
    T s = sqrt(T(3) * (r2 + T(1e-12)));
    T e = exp(-s);
    k = (T(1) + s) * e;
    dk_dr2 = T(-1.5) * e;
  } else if (kid == KERNEL_MATERN52) {
    T s = sqrt(T(5) * (r2 + T(1e-12)));
    T e = exp(-s);
    k = (T(1) + s + (T(5) / T(3)) * r2) * e;
    dk_dr2 = -(T(5) / T(6)) * (T(1) + s) * e;
  } else if (kid == KERNEL_EXPONENTIAL) {   // exp(-r)
    T r = sqrt(r2 + T(1e-12));
    k = exp(-r);
    dk_dr2 = T(-0.5) * k / r;
  } else {                                  // RationalQuadratic
    const T b = T(1) + (T(0.5) / alpha) * r2;
    const T lb = log(b);
    k = exp(-alpha * lb);
    dk_dr2 = T(-0.5) * k / b;
    if (dk_dalpha) *dk_dalpha = k * (r2 / (T(2) * alpha * b) - lb);
  }
}

}  // namespace gdrf
