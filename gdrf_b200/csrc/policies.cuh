// The six contractions of the ELBO + gradient, expressed as policies for gemm_tc_kernel.
//
// Notation: n observation, i/j/m inducing index, k topic.
//   Kxz = k(xs, Z)              W   = Kxz Linv^T            (Linv = chol(Kuu + jI)^-1, lower)
//   T_k = W S_k                 q_kn = sum_j T_k[n,j]^2     (S_k lower triangular; T is kept, 16-bit, for the backward)
//   g2  = 2 dELBO/df_var        WG_k = diag(g2[k,:]) W
//   dW  = sum_k diag(g2[k,:]) (T_k S_k^T) (+ mean / var0 terms added by k_dw_finalize)
//   dS_k = tril(WG_k^T T_k)     dKxz = dWtot Linv           C5 = dWtot^T W   (feeds the Cholesky adjoint)
//
// Operand storage (all "tiled planes", common.cuh), per chunk of RT*128 observation rows, Mp = M
// rounded up to 256, MB = Mp/64, JT = Mp/256:
//   KXZ, W, DWT : [rows n][cols inducing]            3 planes
//   LINV        : [rows m][cols i]                   3 planes   (zero above the diagonal / in padding)
//   ST          : [rows (k, j)][cols i] = S_k[i, j]  3 bf16 planes (+ 2 fp16 planes for the forward); zero for i < j
//   TP          : [rows n][cols (k, j)] = T_k[n, j]  2 planes   (written by the forward contraction's epilogue)
//   WG          : [rows n][cols (k, m)] = g2[k,n] W  2 planes   (k_scale_w)
#pragma once
#include "gemm_tc.cuh"

namespace gdrf {

__device__ __forceinline__ void store8(bf16* dst, const uint4& pk) { *reinterpret_cast<uint4*>(dst) = pk; }

// Two adjacent 8-element packets (columns col .. col+15, col a multiple of 16) of row r as ONE 32-byte store.
// The 128-byte swizzle only permutes 16-byte chunks with XOR (row & 7), so the two chunks of a 32-byte sector stay
// in the same sector (possibly swapped): a full-sector store needs no read-for-fill in L2 (a pair of 16-byte stores
// cost a DRAM read per written sector -- measured 2.3 GB per launch on the T store).
// `policy` != 0: an L2 cache-policy operand (evict_first, l2_evict_first_policy()) for multi-GB outputs that are not
// re-read before they fall out of L2 anyway (the T planes): without it they push the operand the same kernel re-reads
// (the W tile of a CTA pair, once per topic) out of L2.  Measured on the forward contraction (ncu, one chunk): DRAM
// reads 0.19 GB with the T store switched off, 3.37 GB with `st.global.cs` (the streaming hint does not reach L2's
// replacement), 2.55 GB with the cache-policy operand; the step 309.9 -> 307.2 ms.
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void store16(const PlaneMat& m, int plane, int r, int col, const uint4& lo, const uint4& hi,
                                        uint64_t policy = 0) {
  bf16* p = m.elem(plane, r, col);                       // address of the chunk holding columns col .. col+7
  const bool swapped = (r & 1) != 0;                     // chunk index parity flips with row parity
  bf16* sector = swapped ? p - 8 : p;
  const uint4& a = swapped ? hi : lo;
  const uint4& b = swapped ? lo : hi;
  if (policy != 0)
    asm volatile("st.global.L2::cache_hint.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8}, %9;" ::"l"(sector), "r"(a.x),
                 "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w), "l"(policy)
                 : "memory");
  else
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(sector), "r"(a.x), "r"(a.y), "r"(a.z),
                 "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w)
                 : "memory");
}

// sum of 32 squares: four independent fp32 partial sums of 8 terms, combined in fp64.  (All-fp64 was measured: the 32
// F2F conversions per call run on the quarter-rate pipe and made the forward contraction 20 % slower, for 3e-8 of
// relative accuracy that does not show behind the accumulation noise of T itself.)
__device__ __forceinline__ double sumsq32(const float (&v)[32]) {
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    s0 = fmaf(v[j], v[j], s0);
    s1 = fmaf(v[8 + j], v[8 + j], s1);
    s2 = fmaf(v[16 + j], v[16 + j], s2);
    s3 = fmaf(v[24 + j], v[24 + j], s3);
  }
  return ((double)s0 + (double)s1) + ((double)s2 + (double)s3);
}

// ---------------------------------------------------------------------------------------------
// G1:  W[n, m] = sum_{i <= m} Kxz[n, i] Linv[m, i]      epilogue: W planes + wsq[n] = sum_m W^2
// ---------------------------------------------------------------------------------------------
template <int BN_>
struct G1T {
  static constexpr int EPI_WARPS = 4;
  static constexpr int PA = 3, PB = 3, BN = BN_;     // 128: one CTA per tile; 256: CTA pair (gemm_tc2_kernel)
  static constexpr bool A_MN = false, B_MN = false;
  static constexpr int FMT = FMT_BF16;
  static constexpr int CB = BN / 64, PCS = BN / 128;
  struct Params {
    PlaneMat kxz, linv, w, w16;   // w16: fp16 2-plane copy of W for the forward row-norm contraction
    double* wsq;   // |W_n|^2, accumulated in fp64 (a 1e-6 error here is a 1e-3 error in the gradients)
    int RT, MB;
    int corr_kn;   // CTA pairs: tiles of at most this many k-blocks issue their correction products first (CorrFirst)
  };
  __device__ static bool corr_first(const Params& p, int kn) { return kn <= p.corr_kn; }
  __device__ static int num_items(const Params& p) { return p.RT; }
  __device__ static int num_subs(const Params& p, int) { return p.MB / CB; }
  __device__ static int k_iters(const Params& p, int, int sub) { return min(p.MB, CB * (sub + 1)); }
  __device__ static const bf16* a_src(const Params& p, int item, int, int kit, int pl, int) {
    return p.kxz.base + pl * p.kxz.plane_stride + p.kxz.block_off(item, kit);
  }
  __device__ static const bf16* b_src(const Params& p, int, int sub, int kit, int pl, int pc) {
    return p.linv.base + pl * p.linv.plane_stride + p.linv.block_off(sub * PCS + pc, kit);
  }
  struct Epi {
    double acc;
    __device__ void item_begin(const Params&, int, int) { acc = 0.0; }
    __device__ void sub_begin(const Params&, int, int, int) {}
    __device__ void chunk(const Params& p, int item, int sub, int row, int c0, const float (&v)[32]) {
      const int r = item * 128 + row;
#pragma unroll
      for (int g = 0; g < 4; g += 2) {
        uint4 pa[3], pb[3];
        split8<3>(&v[g * 8], pa);
        split8<3>(&v[g * 8 + 8], pb);
        const int col = sub * BN + c0 + g * 8;
#pragma unroll
        for (int pl = 0; pl < 3; ++pl) store16(p.w, pl, r, col, pa[pl], pb[pl]);
        uint4 ha[2], hb[2];
        split8h<2>(&v[g * 8], ha);
        split8h<2>(&v[g * 8 + 8], hb);
#pragma unroll
        for (int pl = 0; pl < 2; ++pl) store16(p.w16, pl, r, col, ha[pl], hb[pl]);
      }
      acc += sumsq32(v);
    }
    __device__ void sub_end(const Params&, int, int, int) {}
    __device__ void item_end(const Params& p, int item, int row) { p.wsq[item * 128 + row] = acc; }
  };
};
using G1 = G1T<128>;

// ---------------------------------------------------------------------------------------------
// GF:  f_loc[k, n] = sum_m W[n, m] u_loc[k, m]   (pyro conditional: loc = W @ v_2D; the mean function is zero,
// abstract_gdrf.py:17-18) on the tensor pipe: A = W16 (fp16 pair), B = U16 [256 rows (k, zero beyond K) x Mp] = the fp16
// pair of s_u * u_loc (s_u: power-of-two scale from max |u_loc|, so that any fp32 u_loc fits); the epilogue writes
// f_loc in fp64 times 1 / s_u.  One 256 x 256 pair tile per 256 observation rows; only the first K accumulator columns
// carry data (a 256-wide MMA costs what a narrower one does next to the operand fill).
// ---------------------------------------------------------------------------------------------
struct GF {
  static constexpr int EPI_WARPS = 4;
  static constexpr int FMT = FMT_F16;
  static constexpr int PA = 2, PB = 2, BN = 256;
  static constexpr bool A_MN = false, B_MN = false;
  struct Params {
    PlaneMat w, u;
    double* floc;              // [K][ncp]
    const float* inv_scale;    // 1 / s_u (ps[PS_SU_INV])
    int RT, MB, K, ncp;
  };
  __device__ static int num_items(const Params& p) { return p.RT; }
  __device__ static int num_subs(const Params&, int) { return 1; }
  __device__ static int k_iters(const Params& p, int, int) { return p.MB; }
  __device__ static bool corr_first(const Params&, int) { return true; }
  __device__ static const bf16* a_src(const Params& p, int item, int, int kit, int pl, int) {
    return p.w.base + pl * p.w.plane_stride + p.w.block_off(item, kit);
  }
  __device__ static const bf16* b_src(const Params& p, int, int, int kit, int pl, int pc) {
    return p.u.base + pl * p.u.plane_stride + p.u.block_off(pc, kit);
  }
  struct Epi {
    double inv;
    __device__ void item_begin(const Params& p, int, int) { inv = (double)p.inv_scale[0]; }
    __device__ void sub_begin(const Params&, int, int, int) {}
    __device__ void chunk(const Params& p, int item, int, int row, int c0, const float (&v)[32]) {
      if (c0 >= p.K) return;
      const long long n = (long long)item * 128 + row;
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (c0 + j < p.K) p.floc[(long long)(c0 + j) * p.ncp + n] = inv * (double)v[j];
    }
    __device__ void sub_end(const Params&, int, int, int) {}
    __device__ void item_end(const Params&, int, int) {}
  };
};

// ---------------------------------------------------------------------------------------------
// G2:  T[n, (k, j)] = sum_{i >= j} W[n, i] S_k[i, j]
//   epilogue: q[k, n] = sum_j T^2 (fp64) and, when the gradient is wanted, T itself as two bf16 planes (TP).
//   T is needed to fp32 accuracy here: f_var enters mu = f_loc + f_var * eps as a *scale* of O(variance) and
//   d ll / d mu is O(counts), so a 2^-17 relative error in q shows up as 1e-3 in the gradients.
// MODE 0: bf16 3 x 3 planes / 6 products, 128-wide tiles (24-bit operands; any fp32 range)
// MODE 2: fp16 2 x 2 planes / 3 products, 256-wide tiles (22-bit operands) -- the default
// MODE 3: MODE 2 with segmented accumulation (GDRF_FLAG_SEGMENTED_FWD; gemm_tc.cuh, SegK)
// ---------------------------------------------------------------------------------------------
template <int MODE>
struct G2 {
  // MODE 2: four warps per TMEM lane quarter (two in the other modes): the split + store of T is a latency-bound
  // chain per 32-column chunk, and the MMAs of a tile may only start when the epilogue of the tile before the last one
  // has drained its accumulator -- short tiles (the last column tiles of a topic) wait for it.  8 -> 16 warps:
  // 89.4 -> 86.5 ms per step at C4 (95 registers per thread, so 576 threads still fit the register file)
  static constexpr int EPI_WARPS = (MODE == 2) ? 16 : 8;
  static constexpr int FMT = (MODE >= 2) ? FMT_F16 : FMT_BF16;
  static constexpr int PA = (MODE == 0) ? 3 : 2, PB = (MODE == 0) ? 3 : 2, BN = (MODE == 0) ? 128 : 256;
  static constexpr bool A_MN = false, B_MN = false;
  static constexpr int CB = BN / 64;      // 64-column blocks per tile
  static constexpr int PCS = BN / 128;    // 128-row pieces of ST per tile
  struct Params {
    PlaneMat w, st, tp;
    double* q;         // [K][ncp]  (fp64: q feeds mu = f_loc + f_var * eps); zeroed by the caller, the two column
                       //           halves of a row add their partial sums
    int store_t;       // write TP (gradient pass)
    int RT, MB, K, NT, ncp;   // NT = Mp / BN column tiles per topic
    int varn;          // CTA pairs: 1 = narrow MMAs in the diagonal blocks, 0 = full width (GDRF_FLAG_FULL_WIDTH, A/B)
    int segk;          // CTA pairs: 1 = segmented accumulation (gemm_tc.cuh, SegK), 0 = one TMEM accumulation per tile
    int corr_kn;       // CTA pairs: tiles of at most this many k-blocks issue their correction products first (CorrFirst)
    // CTA pairs: items are (256-row pair tile, group of K / ksplit topics), two consecutive items = the two 128-row
    // halves of one pair tile.  A chunk with fewer than 74 pair tiles (the last chunk of a shard) then still fills
    // the machine: ksplit is chosen by the host so that ceil(pair tiles * ksplit / 74) / ksplit is smallest.
    // 0: one item per 128-row tile with all topics (single-CTA kernel and the checker).
    int ksplit;
  };
  __device__ static int num_items(const Params& p) { return p.ksplit ? ((p.RT + 1) >> 1) * 2 * p.ksplit : p.RT; }
  __device__ static int row_tile(const Params& p, int item) {
    return p.ksplit ? ((item >> 1) / p.ksplit) * 2 + (item & 1) : item;
  }
  // first (topic, column tile) index of the item
  __device__ static int sub0(const Params& p, int item) {
    return p.ksplit ? ((item >> 1) % p.ksplit) * (p.K / p.ksplit) * p.NT : 0;
  }
  __device__ static int num_subs(const Params& p, int) { return (p.ksplit ? p.K / p.ksplit : p.K) * p.NT; }
  __device__ static int k_iters(const Params& p, int, int sub) { return p.MB - CB * (sub % p.NT); }
  __device__ static const bf16* a_src(const Params& p, int item, int sub, int kit, int pl, int) {
    const int jt = sub % p.NT;
    return p.w.base + pl * p.w.plane_stride + p.w.block_off(row_tile(p, item), CB * jt + kit);
  }
  __device__ static const bf16* b_src(const Params& p, int item, int sub, int kit, int pl, int pc) {
    const int jt = sub % p.NT;
    return p.st.base + pl * p.st.plane_stride + p.st.block_off((sub0(p, item) + sub) * PCS + pc, CB * jt + kit);
  }
  // MODE 2, CTA pairs: the first four k-blocks of a column tile are its diagonal 256 x 256 block of S_k, where
  // k-block b only reaches the columns j < 64 (b + 1).  The fp16 ST planes are stored with the rows of every 256-row
  // tile in the order [g3a g2a g1a g0a | g0b g1b g2b g3b] (gXa / gXb: first / second 32 rows of 64-row group X,
  // k_pack_st), so the needed rows are the ones next to the middle: rank 0 stages the last N/2 rows of its block,
  // rank 1 the first N/2 of its, and the MMA covers accumulator columns [128 - N/2, 128 + N/2).  The k-blocks are
  // walked last-to-first so that the first MMA is full width; 15 % fewer MMA columns per tile row.
  static constexpr int VARN = (MODE >= 2) ? 1 : 0;
  static constexpr int SEGK = (MODE == 3) ? 1 : 0;   // its own instantiation: the default kernel keeps its registers
  __device__ static bool corr_first(const Params& p, int kn) { return kn <= p.corr_kn; }
  __device__ static int kblock(int kit, int kn) { return kn - 1 - kit; }
  __device__ static int ncols(const Params& p, int kb, int) {
    return (kb >= 4 || p.varn == 0) ? 256 : 64 * (kb + 1);
  }
  // operand bases of one (item, sub), decoded once by the producer (a single thread issues every copy of a k-block:
  // integer divisions per copy would make it the bottleneck)
  struct Ctx { const bf16 *a0, *b0; };
  __device__ static Ctx make_ctx(const Params& p, int item, int sub) {
    const int jt = sub % p.NT;
    Ctx c;
    c.a0 = p.w.base + p.w.block_off(row_tile(p, item), CB * jt);
    c.b0 = p.st.base + p.st.block_off((sub0(p, item) + sub) * PCS, CB * jt);
    return c;
  }
  __device__ static const bf16* a_at(const Params& p, const Ctx& c, int kb, int pl) {
    return c.a0 + pl * p.w.plane_stride + (long long)kb * TILE_ELEMS;
  }
  __device__ static void load_b(const Params& p, const Ctx& c, int kb, int pl, int rank, int ncols, uint8_t* dst,
                                uint64_t* bar) {
    // this CTA's 128-row block of ST (rows already permuted, k_pack_st): row block `rank` of the tile's two
    const bf16* blk = c.b0 + pl * p.st.plane_stride + ((long long)rank * p.st.col_blocks + kb) * TILE_ELEMS;
    const int rows = ncols >> 1;
    bulk_g2s(dst, rank == 0 ? blk + (128 - rows) * 64 : blk, rows * 128, bar);
  }
  // accumulator column -> column of the tile under that row order (identity for the other modes)
  __device__ static int tile_col(int c) {
    if (MODE < 2) return c;
    return c < 128 ? 64 * (3 - (c >> 5)) + (c & 31) : 64 * ((c - 128) >> 5) + 32 + (c & 31);
  }
  struct Epi {
    double qacc;
    int rt, s0;
    uint64_t pol;
    __device__ void item_begin(const Params& p, int item, int) {
      qacc = 0.0;
      rt = row_tile(p, item);
      s0 = sub0(p, item);
      pol = l2_evict_first_policy();
    }
    __device__ void sub_begin(const Params& p, int, int sub, int) {
      if (sub % p.NT == 0) qacc = 0.0;
    }
    __device__ void chunk(const Params& p, int, int sub, int row, int c0, const float (&v)[32]) {
      qacc += sumsq32(v);
      if (p.store_t && rt < p.RT) {
        const int r = rt * 128 + row;
        const int col0 = (s0 + sub) * BN + tile_col(c0);     // (k * NT + jt) * BN == k * Mp + jt * BN
#pragma unroll
        for (int g = 0; g < 4; g += 2) {
          uint4 pa[2], pb[2];       // T planes in the forward's format: fp16 pairs (22 bits) or bf16 pairs
          split8x2(FMT, &v[g * 8], pa);
          split8x2(FMT, &v[g * 8 + 8], pb);
#pragma unroll
          for (int pl = 0; pl < 2; ++pl) store16(p.tp, pl, r, col0 + g * 8, pa[pl], pb[pl], pol);
        }
      }
    }
    __device__ void sub_end(const Params& p, int, int sub, int row) {
      if ((sub % p.NT) == p.NT - 1 && rt < p.RT)
        atomicAdd(&p.q[(long long)((s0 + sub) / p.NT) * p.ncp + rt * 128 + row], qacc);
    }
    __device__ void item_end(const Params&, int, int) {}
  };
};

// ---------------------------------------------------------------------------------------------
// G3:  dW[n, i] = sum_k g2[k, n] * sum_{j <= i} T_k[n, j] S_k[i, j]
// One accumulation per (column tile, topic); the epilogue (8 warps: thread = row x column half) applies the
// per-observation, per-topic scale in fp32 and keeps the running sum over topics in 128 registers, so neither a
// rescaled copy of T nor a recompute of T is needed.  B = ST read MN-major.
// ---------------------------------------------------------------------------------------------
struct G3 {
  static constexpr int EPI_WARPS = 8;
  static constexpr int FMT = FMT_BF16;   // overridden at run time by Params::fmt
  static constexpr int PA = 2, PB = 2, BN = 256;
  static constexpr bool A_MN = false, B_MN = true;
  struct Params {
    PlaneMat tp, st;   // st: natural-order planes in the format of tp (fp16 pairs: ST16N; bf16 pairs: ST)
    const float* g2;   // [K][ncp]
    float* dw;         // [ncp][Mp] fp32
    unsigned* cs;      // per-chunk scalars: max |dW| is recorded in cs[CS_DWMAX] (sizes the scale of the dWtot planes)
    int fmt;           // FMT_F16 / FMT_BF16 of both operands
    int RT, MB, K, JT, Mp, ncp;
    int varn;          // CTA pairs only: narrow MMAs in the diagonal block + the column order that goes with them
    // CTA pairs: items are (column tile, 256-row pair tile), widest column tile first, two consecutive items = the two
    // 128-row halves of a pair tile; round-robin over the CTA pairs this is the same schedule as one item per row
    // tile when the chunk has 74 pair tiles, and it spreads a shorter chunk (the last one of a shard) over all SMs.
    // 0: one item per 128-row tile with all column tiles (single-CTA kernel and the checker).
    int isplit;
  };
  __device__ static int rt2(const Params& p) { return (p.RT + 1) >> 1; }
  __device__ static int num_items(const Params& p) { return p.isplit ? 2 * rt2(p) * p.JT : p.RT; }
  __device__ static int row_tile(const Params& p, int item) {
    return p.isplit ? ((item >> 1) % rt2(p)) * 2 + (item & 1) : item;
  }
  __device__ static int col_tile(const Params& p, int item, int sub) {
    return p.isplit ? p.JT - 1 - (item >> 1) / rt2(p) : sub / p.K;
  }
  __device__ static int num_subs(const Params& p, int) { return p.isplit ? p.K : p.JT * p.K; }   // sub = [it * K +] k
  __device__ static int k_iters(const Params& p, int item, int sub) { return 4 * (col_tile(p, item, sub) + 1); }
  __device__ static const bf16* a_src(const Params& p, int item, int sub, int kit, int pl, int) {
    const int k = sub % p.K;
    return p.tp.base + pl * p.tp.plane_stride + p.tp.block_off(row_tile(p, item), k * p.MB + kit);
  }
  __device__ static const bf16* b_src(const Params& p, int item, int sub, int kit, int pl, int pc) {
    const int it = col_tile(p, item, sub), k = sub % p.K;
    return p.st.base + pl * p.st.plane_stride + p.st.block_off(k * 2 * p.JT + (kit >> 1), it * 4 + pc) +
           (kit & 1) * 4096;
  }
  // CTA pairs (Params::varn): the LAST four k-blocks of a column tile are its diagonal block, where k-block b only
  // reaches the columns i >= 64 b.  The accumulator holds the four 64-column groups in the order [g0 g2 | g3 g1], so
  // the last two k-blocks run 128-wide MMAs on the middle of it (rank 0 stages g2, rank 1 g3); 10 % fewer MMA columns.
  static constexpr int VARN = 2;
  __device__ static int kblock(int kit, int) { return kit; }
  __device__ static int ncols(const Params& p, int kb, int kn) { return (p.varn && kb >= kn - 2) ? 128 : 256; }
  struct Ctx { const bf16 *a0, *b0; };
  __device__ static Ctx make_ctx(const Params& p, int item, int sub) {
    const int k = sub % p.K;
    Ctx c;
    c.a0 = p.tp.base + p.tp.block_off(row_tile(p, item), k * p.MB);
    c.b0 = p.st.base + p.st.block_off(k * 2 * p.JT, col_tile(p, item, sub) * 4);
    return c;
  }
  __device__ static const bf16* a_at(const Params& p, const Ctx& c, int kb, int pl) {
    return c.a0 + pl * p.tp.plane_stride + (long long)kb * TILE_ELEMS;
  }
  __device__ static void load_b(const Params& p, const Ctx& c, int kb, int pl, int rank, int ncols, uint8_t* dst,
                                uint64_t* bar) {
    // 64 (j) x 64 (i) pieces of ST: row block kb >> 1, half (kb & 1), column block = piece
    const bf16* row = c.b0 + pl * p.st.plane_stride + (long long)(kb >> 1) * p.st.col_blocks * TILE_ELEMS + (kb & 1) * 4096;
    if (ncols == 256) {
      const int p0 = p.varn ? (rank == 0 ? 0 : 3) : 2 * rank, p1 = p.varn ? (rank == 0 ? 2 : 1) : 2 * rank + 1;
      bulk_g2s(dst, row + p0 * TILE_ELEMS, 8192, bar);
      bulk_g2s(dst + 8192, row + p1 * TILE_ELEMS, 8192, bar);
    } else {
      bulk_g2s(dst, row + (rank == 0 ? 2 : 3) * TILE_ELEMS, 8192, bar);
    }
  }
  struct Epi {
    float acc[128];
    float scale;
    __device__ void item_begin(const Params&, int, int) {}
    __device__ void sub_begin(const Params& p, int item, int sub, int row) {
      const int k = sub % p.K;
      if (k == 0) {
#pragma unroll
        for (int j = 0; j < 128; ++j) acc[j] = 0.f;
      }
      scale = p.g2[(long long)k * p.ncp + row_tile(p, item) * 128 + row];   // no state beyond acc[]: the kernel is at
    }                                                                          // its register limit
    __device__ void chunk(const Params&, int, int, int, int c0, const float (&v)[32]) {
      // c0 is 0/32/64/96 (+128 for the upper-half warps); the unrolled switch keeps acc[] in registers
      switch ((c0 & 127) >> 5) {
        case 0:
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[j] = fmaf(scale, v[j], acc[j]);
          break;
        case 1:
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[32 + j] = fmaf(scale, v[j], acc[32 + j]);
          break;
        case 2:
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[64 + j] = fmaf(scale, v[j], acc[64 + j]);
          break;
        default:
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[96 + j] = fmaf(scale, v[j], acc[96 + j]);
          break;
      }
      last_c0 = c0;
    }
    int last_c0;
    __device__ void sub_end(const Params& p, int item, int sub, int row) {
      const int rt = row_tile(p, item);
      if (sub % p.K != p.K - 1 || rt >= p.RT) return;
      const int it = col_tile(p, item, sub);
      const int half = last_c0 >> 7;
      float* base = p.dw + (long long)(rt * 128 + row) * p.Mp + it * 256;
      {
        float mx = 0.f;
#pragma unroll
        for (int j = 0; j < 128; ++j) mx = fmaxf(mx, fabsf(acc[j]));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if ((row & 31) == 0 && mx > 0.f) atomic_max_abs(p.cs + CS_DWMAX, mx);
      }
      // 64-column groups held by this thread: natural order, or [g0 g2 | g3 g1] under varn
      float4* d0 = reinterpret_cast<float4*>(base + (p.varn ? (half ? 192 : 0) : half * 128));
      float4* d1 = reinterpret_cast<float4*>(base + (p.varn ? (half ? 64 : 128) : half * 128 + 64));
#pragma unroll
      for (int g = 0; g < 16; ++g) d0[g] = make_float4(acc[4 * g], acc[4 * g + 1], acc[4 * g + 2], acc[4 * g + 3]);
#pragma unroll
      for (int g = 0; g < 16; ++g)
        d1[g] = make_float4(acc[64 + 4 * g], acc[65 + 4 * g], acc[66 + 4 * g], acc[67 + 4 * g]);
    }
    __device__ void item_end(const Params&, int, int) {}
  };
};

// ---------------------------------------------------------------------------------------------
// G4:  dKxz[n, i] = sum_{m >= i} dWtot[n, m] Linv[m, i]          (B = LINV read MN-major)
// ---------------------------------------------------------------------------------------------
template <int BN_>
struct G4T {
  static constexpr int EPI_WARPS = 4;
  static constexpr int FMT = FMT_BF16;
  static constexpr int PA = 2, PB = 2, BN = BN_;   // 16-bit operands: see DESIGN.md (hyper-gradient probe)
  static constexpr bool A_MN = false, B_MN = true;
  static constexpr int CB = BN / 64;
  struct Params {
    PlaneMat dwt, linv;   // both in format `fmt` (fp16 pairs: dWtot scaled by s_d, LINV16; bf16 pairs: LINV planes 0, 1)
    float* dkxz;          // [ncp][Mp] fp32
    const float* inv_scale;   // 1 / s_d (cs[CS_SD_INV]), applied to the accumulator
    int fmt;
    int RT, MB, Mp;
  };
  __device__ static int num_items(const Params& p) { return p.RT; }
  __device__ static int num_subs(const Params& p, int) { return p.MB / CB; }
  __device__ static int k_iters(const Params& p, int, int sub) { return p.MB - CB * sub; }
  __device__ static const bf16* a_src(const Params& p, int item, int sub, int kit, int pl, int) {
    return p.dwt.base + pl * p.dwt.plane_stride + p.dwt.block_off(item, CB * sub + kit);
  }
  __device__ static const bf16* b_src(const Params& p, int, int sub, int kit, int pl, int pc) {
    const int mb = CB * sub + kit;
    return p.linv.base + pl * p.linv.plane_stride + p.linv.block_off(mb >> 1, sub * CB + pc) + (mb & 1) * 4096;
  }
  struct Epi {
    float inv;
    __device__ void item_begin(const Params& p, int, int) { inv = p.inv_scale[0]; }
    __device__ void sub_begin(const Params&, int, int, int) {}
    __device__ void chunk(const Params& p, int item, int sub, int row, int c0, const float (&v)[32]) {
      float4* dst = reinterpret_cast<float4*>(p.dkxz + (long long)(item * 128 + row) * p.Mp + sub * BN + c0);
#pragma unroll
      for (int g = 0; g < 8; ++g)
        dst[g] = make_float4(inv * v[4 * g], inv * v[4 * g + 1], inv * v[4 * g + 2], inv * v[4 * g + 3]);
    }
    __device__ void sub_end(const Params&, int, int, int) {}
    __device__ void item_end(const Params&, int, int) {}
  };
};
using G4 = G4T<128>;

// ---------------------------------------------------------------------------------------------
// G5:  C5[a, b] += sum_n dWtot[n, a] W[n, b]        (both operands MN-major; split over n; fp64 atomics)
// ---------------------------------------------------------------------------------------------
template <int BN_>
struct G5T {
  static constexpr int EPI_WARPS = 4;
  static constexpr int FMT = FMT_BF16;
  static constexpr int PA = 2, PB = 2, BN = BN_;   // 16-bit operands, like G4
  static constexpr bool A_MN = true, B_MN = true;
  static constexpr int CB = BN / 64;
  struct Params {
    PlaneMat dwt, w;      // both in format `fmt` (fp16 pairs: dWtot scaled by s_d, W16; bf16 pairs: W planes 0, 1)
    double* c5;           // [Mp][Mp]
    const float* inv_scale;   // 1 / s_d
    int fmt;
    int RT, MB, Mp, MT, splits, nb_per_split;
    // du_loc[k, m] += sum_n g_loc[k, n] W[n, m] rides along: k_dw_finalize writes s_l * g_loc as the column blocks
    // [Mp, Mp + 128) of the dWtot planes, and MT counts two more 128-row output tiles than Mp / 128 (a pair tile; the
    // second one is padding) whose rows are topics and go to du instead of C5
    double* du;           // [K][M] fp64 accumulator, or NULL (MT == MTW then)
    const float* inv_scale_l;   // 1 / s_l (cs[CS_SL_INV])
    int MTW, K, M;        // MTW = Mp / 128 row tiles of C5
  };
  // item = (split, column tile bt of BN columns, 128-row tile at), at fastest: with BN = 256 the items 2t, 2t + 1 are
  // the two row halves of one 256 x 256 pair tile
  __device__ static int per_split(const Params& p) { return p.MT * (p.Mp / BN); }
  __device__ static int num_items(const Params& p) { return per_split(p) * p.splits; }
  __device__ static int num_subs(const Params&, int) { return 1; }
  __device__ static int k_iters(const Params& p, int item, int) {
    const int s = item / per_split(p);
    return min(p.nb_per_split, 2 * p.RT - s * p.nb_per_split);
  }
  __device__ static const bf16* a_src(const Params& p, int item, int, int kit, int pl, int pc) {
    const int s = item / per_split(p), t = item - s * per_split(p), at = min(t % p.MT, p.MTW);   // padding tile: any data
    const int nb = s * p.nb_per_split + kit;
    return p.dwt.base + pl * p.dwt.plane_stride + p.dwt.block_off(nb >> 1, at * 2 + pc) + (nb & 1) * 4096;
  }
  __device__ static const bf16* b_src(const Params& p, int item, int, int kit, int pl, int pc) {
    const int s = item / per_split(p), t = item - s * per_split(p), bt = t / p.MT;
    const int nb = s * p.nb_per_split + kit;
    return p.w.base + pl * p.w.plane_stride + p.w.block_off(nb >> 1, bt * CB + pc) + (nb & 1) * 4096;
  }
  struct Epi {
    double inv;
    __device__ void item_begin(const Params& p, int, int) { inv = (double)p.inv_scale[0]; }
    __device__ void sub_begin(const Params&, int, int, int) {}
    __device__ void chunk(const Params& p, int item, int, int row, int c0, const float (&v)[32]) {
      const int t = item % per_split(p), at = t % p.MT, bt = t / p.MT;
      if (at >= p.MTW) {                      // topic rows: du_loc
        const int k = (at - p.MTW) * 128 + row, m0 = bt * BN + c0;
        if (k >= p.K) return;
        const double il = (double)p.inv_scale_l[0];
        double* dst = p.du + (long long)k * p.M + m0;
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (m0 + j < p.M) atomicAdd(dst + j, il * (double)v[j]);
        return;
      }
      double* dst = p.c5 + (long long)(at * 128 + row) * p.Mp + bt * BN + c0;
#pragma unroll
      for (int j = 0; j < 32; ++j) atomicAdd(dst + j, inv * (double)v[j]);
    }
    __device__ void sub_end(const Params&, int, int, int) {}
    __device__ void item_end(const Params&, int, int) {}
  };
};
using G5 = G5T<128>;

// ---------------------------------------------------------------------------------------------
// G6:  dS_k[i, j] += sum_n WG[n, (k, i)] T[n, (k, j)]   for j <= i   (both MN-major; tiles on/below the diagonal)
// ---------------------------------------------------------------------------------------------
struct G6 {
  static constexpr int EPI_WARPS = 4;
  static constexpr int FMT = FMT_BF16;
  static constexpr int PA = 2, PB = 2, BN = 256;
  static constexpr bool A_MN = true, B_MN = true;
  static constexpr int MAX_TILES = 512;
  struct Params {
    PlaneMat wg, tp;   // both in format `fmt`; WG carries the power-of-two scale s_g
    float* ds;   // [K][M][M] fp32, unpadded gradient accumulator
    const float* inv_scale;   // 1 / s_g (cs[CS_SG_INV]), applied to the accumulator
    int fmt;
    // items [0, n_whole) contract the whole observation range of the chunk and own their tile (plain += on dS);
    // each of the remaining (topic, tile) pairs is cut into tail_sp pieces of tail_per 64-observation blocks that add
    // atomically -- sized by the host so that the last round of the persistent CTAs / CTA pairs is full
    int RT, MB, K, M, ntile, n_whole, tail_sp, tail_per;
    unsigned char ta[MAX_TILES], tb[MAX_TILES];
  };
  struct Item {
    int k, t, nb0, cnt;
    bool whole;
  };
  __device__ static Item decode(const Params& p, int item) {
    Item it;
    int rem;
    if (item < p.n_whole) {
      rem = item; it.nb0 = 0; it.cnt = 2 * p.RT; it.whole = true;
    } else {
      const int q = item - p.n_whole, pi = q >> 1, h = q & 1;     // 2 t, 2 t + 1: the row halves of one pair tile
      const int sp = pi % p.tail_sp, bp = pi / p.tail_sp;
      rem = p.n_whole + 2 * bp + h;
      it.nb0 = sp * p.tail_per;
      it.cnt = min(p.tail_per, 2 * p.RT - it.nb0);
      it.whole = false;
    }
    it.k = rem / p.ntile;
    it.t = rem - it.k * p.ntile;
    return it;
  }
  __device__ static int num_items(const Params& p) { return p.n_whole + (p.K * p.ntile - p.n_whole) * p.tail_sp; }
  __device__ static int num_subs(const Params&, int) { return 1; }
  __device__ static int k_iters(const Params& p, int item, int) { return decode(p, item).cnt; }
  __device__ static const bf16* a_at(const Params& p, const Item& it, int kit, int pl, int pc) {
    const int nb = it.nb0 + kit;
    return p.wg.base + pl * p.wg.plane_stride + p.wg.block_off(nb >> 1, it.k * p.MB + p.ta[it.t] * 2 + pc) +
           (nb & 1) * 4096;
  }
  __device__ static const bf16* b_at(const Params& p, const Item& it, int kit, int pl, int pc) {
    const int nb = it.nb0 + kit;
    return p.tp.base + pl * p.tp.plane_stride + p.tp.block_off(nb >> 1, it.k * p.MB + p.tb[it.t] * 4 + pc) +
           (nb & 1) * 4096;
  }
  __device__ static const bf16* a_src(const Params& p, int item, int, int kit, int pl, int pc) {
    return a_at(p, decode(p, item), kit, pl, pc);
  }
  __device__ static const bf16* b_src(const Params& p, int item, int, int kit, int pl, int pc) {
    return b_at(p, decode(p, item), kit, pl, pc);
  }
  struct Epi {
    float inv;
    __device__ void item_begin(const Params& p, int, int) { inv = p.inv_scale[0]; }
    __device__ void sub_begin(const Params&, int, int, int) {}
    __device__ void chunk(const Params& p, int item, int, int row, int c0, const float (&v)[32]) {
      const Item it = decode(p, item);
      const int i = p.ta[it.t] * 128 + row;
      const int j0 = p.tb[it.t] * 256 + c0;
      if (i >= p.M) return;
      float* dst = p.ds + ((long long)it.k * p.M + i) * p.M;
      if (it.whole) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j0 + j <= i) dst[j0 + j] = fmaf(inv, v[j], dst[j0 + j]);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j0 + j <= i) atomicAdd(dst + j0 + j, inv * v[j]);
      }
    }
    __device__ void sub_end(const Params&, int, int, int) {}
    __device__ void item_end(const Params&, int, int) {}
  };
};

}  // namespace gdrf
