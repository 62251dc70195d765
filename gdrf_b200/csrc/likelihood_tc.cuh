// Stage 4 on the tensor pipe: the theta-phi mixture, the count-weighted multinomial log-likelihood and its backward in
// ONE kernel with three chained tcgen05 contractions per 128-observation tile (sparse_gdrf.py:361-372 ->
// torch.distributions.Multinomial.log_prob; K <= 64 topics, any V):
//
//   P   [n, v] = sum_k theta[n, k] phi[k, v]                       (M = 128 observations, N = 128 categories, K = 64)
//   elementwise, P read from TMEM:  phat = p / s_n, logit = log(clamp(phat, eps32, 1 - eps32)),
//                                   ll_n += w logit - lgamma(w + 1),  r = w a / p  (a: clamp inactive)
//   G1  [n, k] += sum_v r[n, v] phi[k, v]                          (d ll / d theta before the renormalisation term)
//   dphi[k, v]  = sum_n theta[n, k] r[n, v]                        (this tile's share of d ll / d phi)
//
// theta and phi are error-compensated bf16 triples (probabilities span more exponents than fp16 has): P runs on 24-bit
// operands (6 MMAs per product); r is a bf16 pair and the two backward products run on 16-bit operands (3 MMAs per
// product: their rounding is unbiased noise on per-observation weights, 2^-17 relative, that averages over V and N).  theta and r are built by the kernel's own threads straight into the swizzled shared-memory
// tiles the MMAs read (fence.proxy.async in between); phi^T comes pre-packed (k_pack_phit) by cp.async.bulk.  The same
// theta tile is the K-major A of P and the MN-major A of dphi; the same phi^T tile is the K-major B of P and the
// MN-major B of G1; the same r tile is the K-major A of G1 and the MN-major B of dphi.  The N x V probability matrix
// lives in TMEM (128 columns at a time) and r in shared memory: neither reaches HBM.
//
// Warps: 0 = bulk copies + MMA issue (one lane), 1 = TMEM allocation, 2..17 = 16 elementwise warps (4 per TMEM lane
// quarter, 32 accumulator columns each).  One tile per CTA; V is walked in chunks of 128 categories.
#pragma once
#include "gemm_tc.cuh"
#include "stages.cuh"

namespace gdrf {

constexpr int LT_THREADS = 64 + 16 * 32;
constexpr int LT_VC = 128;                                     // categories per chunk
constexpr int LT_PLANE = 16384;                                // one 128 x 64 bf16 tile
constexpr int LT_OFF_THETA = 0;                                // 3 planes [128 n x 64 k]
constexpr int LT_OFF_PHIT = 3 * LT_PLANE;                      // 2 buffers x 3 planes [128 v x 64 k]
constexpr int LT_OFF_R = 9 * LT_PLANE;                         // 2 planes x 2 tiles [128 n x 64 v]
constexpr int LT_OFF_MISC = 13 * LT_PLANE;                     // barriers, per-row partial sums
constexpr int LT_LFACT = 4096;                                 // log-factorial table: lgamma(c + 1) for counts c < 4096
constexpr int LT_SMEM = LT_OFF_MISC + 64 + 3 * 512 * 4 + 1024; // barriers, partial sums, alignment slack

// phi^T as bf16 triples in the tiled layout: PHIT[v, k] = phi[k, v], rows padded to 128, 64 columns (zero beyond K)
__global__ void __launch_bounds__(256) k_pack_phit(const float* __restrict__ phi, int K, int V, PlaneMat phit) {
  const int rt = blockIdx.x;
  for (int t = threadIdx.x; t < 128 * 8; t += 256) {
    const int r = t >> 3, g = t & 7;
    const int v = rt * 128 + r;
    float x[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int k = g * 8 + e;
      x[e] = (v < V && k < K) ? phi[(long long)k * V + v] : 0.f;
    }
    uint4 pk[3];
    split8<3>(x, pk);
#pragma unroll
    for (int pl = 0; pl < 3; ++pl) *reinterpret_cast<uint4*>(phit.elem(pl, v, g * 8)) = pk[pl];
  }
}

// lgamma(c + 1) for integer counts: the kernel's lgammaf calls were divergent (every lane its own count) and a
// double-digit share of its instructions
__global__ void k_lfact_table(float* __restrict__ tab) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < LT_LFACT) tab[c] = (float)lgamma((double)c + 1.0);
}

// dphi_acc[k, v] += inv_p * sum_tiles dphi_part[tile, k, v]: thread = (k, v), slabs read coalesced; the tiles are cut into
// gridDim.y groups (enough CTAs to pull the slabs at HBM / L2 speed) whose sums meet in one fp64 atomic each
__global__ void __launch_bounds__(256) k_reduce_dphi(const float* __restrict__ part, int tiles, int K, int V, int Vp,
                                                     double* __restrict__ dphi_acc, double inv_p) {
  const long long i = blockIdx.x * 256LL + threadIdx.x;
  if (i >= (long long)K * Vp) return;
  const int k = (int)(i / Vp), v = (int)(i - (long long)k * Vp);
  if (v >= V) return;
  float s0 = 0.f, s1 = 0.f;
  int t = blockIdx.y;
  for (; t + (int)gridDim.y < tiles; t += 2 * gridDim.y) {
    s0 += part[(long long)t * K * Vp + i];
    s1 += part[(long long)(t + gridDim.y) * K * Vp + i];
  }
  if (t < tiles) s0 += part[(long long)t * K * Vp + i];
  atomicAdd(&dphi_acc[(long long)k * V + v], inv_p * ((double)s0 + (double)s1));
}

// counts beyond the table (rare): kept out of line so that the unrolled element loop stays small
__device__ __noinline__ float lt_lgamma_big(float cf) { return lgammaf(cf + 1.f); }

__device__ __forceinline__ void lt_umma(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, bool acc) {
  umma_bf16(d, da, db, idesc, acc ? 1u : 0u);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__global__ void __launch_bounds__(LT_THREADS, 1)
    k_likelihood_tc(int nc, int ncp, int K, int V, const int* __restrict__ ws, const float* __restrict__ theta,
                    const float* __restrict__ srow, PlaneMat phit, float* __restrict__ g1 /*[ncp][K]*/,
                    float* __restrict__ arow, float* __restrict__ cnt,
                    float* __restrict__ dphi_part /*[tiles][K][Vp], Vp = V rounded up to 128*/,
                    const float* __restrict__ lfact, double* __restrict__ acc, double inv_p) {
  extern __shared__ uint8_t lt_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)lt_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(smem + LT_OFF_MISC);
  uint64_t* bar_phit = bars;          // [2] phi^T chunk landed in buffer c & 1 (tx bytes)
  uint64_t* bar_p = bars + 2;         // P accumulator complete
  uint64_t* bar_r = bars + 3;         // r tile (and, first time, the theta tile) written by the 16 elementwise warps
  uint64_t* bar_g = bars + 4;         // G1 / dphi MMAs of the chunk retired
  uint32_t* tmem_slot = (uint32_t*)(bars + 5);
  float* row_ll = (float*)(smem + LT_OFF_MISC + 64);       // [4 column groups][128 rows] partial sums, added in a fixed order
  float* row_a = row_ll + 512;
  float* row_c = row_a + 512;
  __shared__ double scratch[32];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * 128;
  const int nch = (V + LT_VC - 1) / LT_VC;
  constexpr int TM_P = 0, TM_G1 = 128, TM_DPHI = 256;

  if (threadIdx.x == 0) {
    mbar_init(&bar_phit[0], 1);
    mbar_init(&bar_phit[1], 1);
    mbar_init(bar_p, 1);
    mbar_init(bar_r, 16);
    mbar_init(bar_g, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      const uint32_t s_theta = smem_u32(smem + LT_OFF_THETA);
      const uint32_t s_phit0 = smem_u32(smem + LT_OFF_PHIT), s_r = smem_u32(smem + LT_OFF_R);
      const uint32_t id_p = make_idesc(128, 128, false, false, FMT_BF16);     // P: A, B K-major
      const uint32_t id_g = make_idesc(128, 64, false, true, FMT_BF16);       // G1: A K-major (r), B MN-major (phi^T)
      const uint32_t id_d = make_idesc(128, 128, true, true, FMT_BF16);       // dphi: A MN-major (theta), B MN-major (r)
      auto load_phit = [&](int c) {       // phi^T chunk c: rows [128 c, 128 c + 128) of PHIT, three planes, buffer c & 1
        const int buf = c & 1;
        mbar_arrive_expect_tx(&bar_phit[buf], 3 * LT_PLANE);
#pragma unroll
        for (int pl = 0; pl < 3; ++pl)
          bulk_g2s(smem + LT_OFF_PHIT + (buf * 3 + pl) * LT_PLANE, phit.base + pl * phit.plane_stride + phit.block_off(c, 0),
                   LT_PLANE, &bar_phit[buf]);
      };
      auto issue_p = [&](int c) {         // P = theta phi^T : 4 k-steps x 6 products, corrections first
        const uint32_t s_phit = s_phit0 + (c & 1) * 3 * LT_PLANE;
        mbar_wait(&bar_phit[c & 1], (uint32_t)((c >> 1) & 1));
        tc_fence_after();
        bool first = true;
#pragma unroll
        for (int phase = 0; phase < 2; ++phase)
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
#pragma unroll
            for (int pa = 0; pa < 3; ++pa)
#pragma unroll
              for (int pb = 0; pb < 3; ++pb) {
                if (pa + pb > 2 || ((pa + pb == 0) != (phase == 1))) continue;
                const uint64_t da = make_smem_desc(s_theta + pa * LT_PLANE + ks * 32, 16, 1024);
                const uint64_t db = make_smem_desc(s_phit + pb * LT_PLANE + ks * 32, 16, 1024);
                lt_umma(tmem_base + TM_P, da, db, id_p, !first);
                first = false;
              }
        umma_commit(bar_p);
      };
      load_phit(0);
      if (nch > 1) load_phit(1);
      mbar_wait(bar_r, 0);                 // theta tile written
      issue_p(0);
      for (int c = 0; c < nch; ++c) {
        const uint32_t s_phit = s_phit0 + (c & 1) * 3 * LT_PLANE;
        // ---- wait for r(c) (P(c) is drained then) ----
        mbar_wait(bar_r, (uint32_t)((c + 1) & 1));
        tc_fence_after();
        // P of the next chunk first: the elementwise warps go straight on to it, while G1 / dphi of this chunk run
        // behind their arithmetic
        if (c + 1 < nch) issue_p(c + 1);
        // ---- G1 += r phi  and  dphi = theta^T r  (16-bit operands: bf16 pairs, 3 products) ----
        bool first = (c == 0);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)               // 16 categories per step: tile ks >> 2, 32 bytes per step inside
#pragma unroll
          for (int pa = 0; pa < 2; ++pa)
#pragma unroll
            for (int pb = 0; pb < 2; ++pb) {
              if (pa + pb > 1) continue;
              const uint64_t da = make_smem_desc(s_r + pa * 2 * LT_PLANE + (ks >> 2) * LT_PLANE + (ks & 3) * 32, 16, 1024);
              const uint64_t db = make_smem_desc(s_phit + pb * LT_PLANE + ks * 2048, 8192, 1024);
              lt_umma(tmem_base + TM_G1, da, db, id_g, !first);
              first = false;
            }
        first = true;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)               // 16 observations per step
#pragma unroll
          for (int pa = 0; pa < 2; ++pa)
#pragma unroll
            for (int pb = 0; pb < 2; ++pb) {
              if (pa + pb > 1) continue;
              // A = theta^T, MN-major with M = 128: the second 64-row group (accumulator rows 64..127, never read:
              // K <= 64) is pointed at the next theta plane -- any finite data -- instead of a tile of zeros
              const uint64_t da = make_smem_desc(s_theta + pa * LT_PLANE + ks * 2048, LT_PLANE, 1024);
              const uint64_t db = make_smem_desc(s_r + pb * 2 * LT_PLANE + ks * 2048, LT_PLANE, 1024);
              lt_umma(tmem_base + TM_DPHI, da, db, id_d, !first);
              first = false;
            }
        umma_commit(bar_g);
        // once G1(c) has retired its phi^T buffer takes chunk c + 2
        if (c + 2 < nch) {
          mbar_wait(bar_g, (uint32_t)(c & 1));
          load_phit(c + 2);
        }
      }
    }
    __syncwarp();
  } else if (warp >= 2) {
    // ------------------------------ elementwise warps ------------------------------
    const int ew = warp - 2;                       // 0..15
    const int quarter = warp & 3;                  // TMEM lane quarter this warp may read
    const int cg = ew >> 2;                        // column group 0..3: 32 of the chunk's 128 columns
    const int row = quarter * 32 + lane;           // observation row / topic row (dphi) owned by this thread
    const int n = n0 + row;
    const bool live = n < nc;
    const uint32_t tlane = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const float EPS32 = 1.1920928955078125e-07f;
    // theta tile: this thread packs topics [16 cg, 16 cg + 16) of its row
    {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float x[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int k = cg * 16 + h * 8 + e;
          x[e] = (live && k < K) ? theta[(long long)k * ncp + n] : 0.f;
        }
        uint4 pk[3];
        split8<3>(x, pk);
#pragma unroll
        for (int pl = 0; pl < 3; ++pl)
          *reinterpret_cast<uint4*>(smem + LT_OFF_THETA + pl * LT_PLANE + 2 * tile_off(row, cg * 16 + h * 8)) = pk[pl];
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_r);           // phase 0 of bar_r: operands of the first MMAs are in place
    }
    const float sinv = live ? 1.f / srow[n] : 0.f;
    float ll = 0.f, asum = 0.f, csum = 0.f;
    // dphi of chunk c (accumulator rows are topics): this tile's share of d ll / d phi goes to its own slab, which
    // k_reduce_dphi adds up (2.4 M fp64 atomics on 16 K addresses per chunk were a bottleneck)
    auto drain_dphi = [&](int c) {
      mbar_wait(bar_g, (uint32_t)(c & 1));
      tc_fence_after();
      if (quarter * 32 < K) {                       // warp-uniform: this lane quarter holds topics < K
        float d[32];
        tmem_ld32(tlane + TM_DPHI + cg * 32, d);
        if (row < K) {
          float4* dst = reinterpret_cast<float4*>(dphi_part + ((long long)blockIdx.x * K + row) * (nch * LT_VC) +
                                                  c * LT_VC + cg * 32);
#pragma unroll
          for (int q = 0; q < 8; ++q) dst[q] = make_float4(d[4 * q], d[4 * q + 1], d[4 * q + 2], d[4 * q + 3]);
        }
      }
      tc_fence_before();
    };
    for (int c = 0; c < nch; ++c) {
      const int v0 = c * LT_VC + cg * 32;
      // counts of this thread's 32 categories (one contiguous 128-byte run of its row)
      int w[32];
      if (live && v0 + 32 <= V && (V & 3) == 0) {
        const int4* src = reinterpret_cast<const int4*>(ws + (long long)n * V + v0);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int4 t4 = __ldg(src + q);
          w[4 * q] = t4.x; w[4 * q + 1] = t4.y; w[4 * q + 2] = t4.z; w[4 * q + 3] = t4.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) w[j] = (live && v0 + j < V) ? ws[(long long)n * V + v0 + j] : 0;
      }
      mbar_wait(bar_p, (uint32_t)(c & 1));
      tc_fence_after();
      float rr[32];
      {
        float p[32];
        tmem_ld32(tlane + TM_P + cg * 32, p);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float r = 0.f;
          if (w[j] != 0) {
            const float ph = p[j] * sinv;
            const float pc = fminf(fmaxf(ph, EPS32), 1.f - EPS32);
            const float cf = (float)w[j];
            ll += cf * logf(pc) - (w[j] < LT_LFACT ? __ldg(lfact + w[j]) : lt_lgamma_big(cf));
            csum += cf;
            if (ph >= EPS32 && ph <= 1.f - EPS32) {
              asum += cf;
              r = cf / p[j];
            }
          }
          rr[j] = r;
        }
      }
      // the MMAs that read r(c - 1) ran behind the arithmetic above: drain their dphi, then r(c) may overwrite r(c - 1)
      if (c > 0) drain_dphi(c - 1);
      // r -> a bf16 pair, tile (cg >> 1), columns 32 (cg & 1) ... + 31 of this thread's row
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        uint4 pk[2];
        split8<2>(&rr[8 * h], pk);
#pragma unroll
        for (int pl = 0; pl < 2; ++pl)
          *reinterpret_cast<uint4*>(smem + LT_OFF_R + pl * 2 * LT_PLANE + (cg >> 1) * LT_PLANE +
                                    2 * tile_off(row, (cg & 1) * 32 + 8 * h)) = pk[pl];
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_r);
    }
    drain_dphi(nch - 1);
    // ---- per-row sums over the 4 warps of a quarter; G1 out ----
    row_ll[cg * 128 + row] = ll;
    row_a[cg * 128 + row] = asum;
    row_c[cg * 128 + row] = csum;
    {
      float gq[16];
      tmem_ld16(tlane + TM_G1 + cg * 16, gq);      // the last bar_g wait above covers the final G1 MMAs
      if (live) {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (cg * 16 + j < K) g1[(long long)n * K + cg * 16 + j] = gq[j];
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  double ll_local = 0.0;
  if (threadIdx.x < 128) {
    const int n = n0 + threadIdx.x;
    if (n < nc) {
      const int r = threadIdx.x;
      ll_local = (double)((row_ll[r] + row_ll[128 + r]) + (row_ll[256 + r] + row_ll[384 + r]));
      arow[n] = (row_a[r] + row_a[128 + r]) + (row_a[256 + r] + row_a[384 + r]);
      cnt[n] = (row_c[r] + row_c[128 + r]) + (row_c[256 + r] + row_c[384 + r]);
    }
  }
  ll_local = block_sum(ll_local, scratch);
  if (threadIdx.x == 0) atomicAdd(&acc[ACC_LL], ll_local * inv_p);
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace gdrf
