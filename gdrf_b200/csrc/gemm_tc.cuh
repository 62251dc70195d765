// Split-bf16 GEMM on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM), fed by
// cp.async.bulk (TMA bulk copies) of pre-swizzled 16 KB operand blocks, warp-specialised:
//
//   warp 0      one elected lane issues the bulk copies into an NSTAGES-deep SMEM ring
//   warp 1      allocates TMEM; one elected lane issues tcgen05.mma for every (plane_a, plane_b)
//               product of the error-compensated split and commits to the ring / accumulator barriers
//   warps 2..5  epilogue: tcgen05.ld the 128 x BN fp32 accumulator (double-buffered in TMEM) and run
//               the policy's fused epilogue (row norms, rescale + re-split, fp32 store, accumulate)
//
// A "policy" describes one of the contractions of the ELBO (see policies.cuh): where the operand
// blocks live, which k-blocks a triangular operand lets us skip, and what the epilogue does.  The
// same policy drives `gemm_ref_kernel`, a plain-FMA CUDA kernel used by the tests to check the tensor
// path element for element (it is a device-side checker, not a fallback: the product never selects it).
#pragma once
#include <type_traits>
#include <utility>
#include "common.cuh"

namespace gdrf {

// ------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must end in a trap (reported as a CUDA error), never in a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {     // ~2 s at 2 GHz: no legitimate wait in these kernels is that long
      printf("gdrf gemm: mbarrier timeout block %d thread %d\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], bf16 inputs, fp32 accumulate, one CTA.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// 32 lanes x 32 consecutive 32-bit columns -> 32 registers per thread
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// ------------------------------------------------------------------------------------------
// descriptors  (bit layouts: cute/arch/mma_sm100_desc.hpp SmemDescriptor / InstrDescriptor)
// ------------------------------------------------------------------------------------------
// SWIZZLE_128B shared-memory matrix descriptor.
//   K-major  operand: rows of 64 bf16 (128 B); 8-row groups 1024 B apart (SBO); LBO unused (1).
//   MN-major operand: 64 MN-elements contiguous (128 B) per k-row; 8-k-row groups 1024 B apart (SBO);
//                     successive 64-wide MN groups LBO bytes apart.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);              // [0,14)  start address >> 4
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;      // [16,30) leading byte offset >> 4
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;      // [32,46) stride byte offset >> 4
  d |= (uint64_t)1 << 46;                                // [46,48) descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                                // [61,64) layout type SWIZZLE_128B
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, bool a_mn, bool b_mn, int fmt) {
  return (1u << 4)                     // [4,6)   accumulator format F32
         | ((uint32_t)fmt << 7)        // [7,10)  A format: 0 = F16, 1 = BF16
         | ((uint32_t)fmt << 10)       // [10,13) B format
         | ((a_mn ? 1u : 0u) << 15)    // [15]    A major (0 = K, 1 = MN)
         | ((b_mn ? 1u : 0u) << 16)    // [16]    B major
         | ((uint32_t)(N >> 3) << 17)  // [17,23) N >> 3
         | ((uint32_t)(M >> 4) << 24); // [24,29) M >> 4
}

// operand format of a policy: the compile-time P::FMT, or Params::fmt when the policy's operands come in either format
template <class P, class = void>
struct FmtOf {
  __host__ __device__ static int get(const typename P::Params&) { return P::FMT; }
};
template <class P>
struct FmtOf<P, std::void_t<decltype(std::declval<typename P::Params>().fmt)>> {
  __host__ __device__ static int get(const typename P::Params& p) { return p.fmt; }
};

// (plane_a, plane_b) products kept by the issue loops: all pairs with pa + pb <= max(PA, PB) - 1

// warps: 0 = bulk-copy producer, 1 = MMA issuer / TMEM owner, 2.. = epilogue (P::EPI_WARPS of them: 4, 8 or 16)
template <class P>
constexpr int gemm_threads() { return 64 + 32 * P::EPI_WARPS; }
constexpr int GEMM_SMEM_BUDGET = 216 * 1024;

// Operand addresses in the producer loops.  A policy whose (item -> operand block) map is expensive declares
// `Item`, `decode(prm, item)` and `a_at / b_at(prm, Item, kit, plane, piece)`; the producer then decodes once per
// item instead of once per bulk copy (a single thread issues 8-12 copies per k-block).
template <class P, class = void>
struct OperandCtx {
  struct type { int item, sub; };
  __device__ static type make(const typename P::Params&, int item, int sub) { return type{item, sub}; }
  __device__ static const bf16* a(const typename P::Params& p, const type& c, int kit, int pl, int pc) {
    return P::a_src(p, c.item, c.sub, kit, pl, pc);
  }
  __device__ static const bf16* b(const typename P::Params& p, const type& c, int kit, int pl, int pc) {
    return P::b_src(p, c.item, c.sub, kit, pl, pc);
  }
};
template <class P>
struct OperandCtx<P, std::void_t<typename P::Item>> {
  using type = typename P::Item;
  __device__ static type make(const typename P::Params& p, int item, int) { return P::decode(p, item); }
  __device__ static const bf16* a(const typename P::Params& p, const type& c, int kit, int pl, int pc) {
    return P::a_at(p, c, kit, pl, pc);
  }
  __device__ static const bf16* b(const typename P::Params& p, const type& c, int kit, int pl, int pc) {
    return P::b_at(p, c, kit, pl, pc);
  }
};

template <class P>
struct GemmCfg {
  static constexpr int A_BYTES = 16384;                       // one plane of a 128 x 64 A stage
  static constexpr int B_BYTES = (P::BN / 128) * 16384;       // one plane of a BN x 64 B stage
  static constexpr int STAGE_BYTES = P::PA * A_BYTES + P::PB * B_BYTES;
  static constexpr int NSTAGES = (GEMM_SMEM_BUDGET / STAGE_BYTES) < 4 ? (GEMM_SMEM_BUDGET / STAGE_BYTES) : 4;
  static constexpr int SMEM_BYTES = NSTAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
  static constexpr int TMEM_COLS = (2 * P::BN <= 256) ? 256 : 512;
  static_assert(NSTAGES >= 2, "need at least a double-buffered ring");
  static_assert(P::BN == 128 || P::BN == 256, "BN");
};

// ------------------------------------------------------------------------------------------
// the tensor-core kernel
// ------------------------------------------------------------------------------------------
template <class P>
__global__ void __launch_bounds__(gemm_threads<P>(), 1) gemm_tc_kernel(const __grid_constant__ typename P::Params prm) {
  using Cfg = GemmCfg<P>;
  constexpr int NST = Cfg::NSTAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(smem + NST * Cfg::STAGE_BYTES);
  uint64_t* full_bar = bars;                 // [NST]
  uint64_t* empty_bar = bars + NST;          // [NST]
  uint64_t* tfull_bar = bars + 2 * NST;      // [2]
  uint64_t* tempty_bar = bars + 2 * NST + 2; // [2]
  uint32_t* tmem_slot = (uint32_t*)(bars + 2 * NST + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NST; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], P::EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int n_items = P::num_items(prm);

  if (warp == 0) {
    // ------------------------------ bulk-copy producer ------------------------------
    if (lane == 0) {
      uint32_t it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int nsub = P::num_subs(prm, item);
        for (int sub = 0; sub < nsub; ++sub) {
          const int kn = P::k_iters(prm, item, sub);
          const auto octx = OperandCtx<P>::make(prm, item, sub);
          for (int kit = 0; kit < kn; ++kit, ++it) {
            const int s = it % NST;
            const uint32_t ph = (it / NST) & 1;
            mbar_wait(&empty_bar[s], ph ^ 1);
            uint8_t* st = smem + s * Cfg::STAGE_BYTES;
            mbar_arrive_expect_tx(&full_bar[s], Cfg::STAGE_BYTES);
#pragma unroll
            for (int pl = 0; pl < P::PA; ++pl) {
              uint8_t* dst = st + pl * Cfg::A_BYTES;
              if (!P::A_MN) {
                bulk_g2s(dst, OperandCtx<P>::a(prm, octx, kit, pl, 0), 16384, &full_bar[s]);
              } else {
#pragma unroll
                for (int pc = 0; pc < 2; ++pc)
                  bulk_g2s(dst + pc * 8192, OperandCtx<P>::a(prm, octx, kit, pl, pc), 8192, &full_bar[s]);
              }
            }
#pragma unroll
            for (int pl = 0; pl < P::PB; ++pl) {
              uint8_t* dst = st + P::PA * Cfg::A_BYTES + pl * Cfg::B_BYTES;
              if (!P::B_MN) {
#pragma unroll
                for (int pc = 0; pc < P::BN / 128; ++pc)
                  bulk_g2s(dst + pc * 16384, OperandCtx<P>::b(prm, octx, kit, pl, pc), 16384, &full_bar[s]);
              } else {
#pragma unroll
                for (int pc = 0; pc < P::BN / 64; ++pc)
                  bulk_g2s(dst + pc * 8192, OperandCtx<P>::b(prm, octx, kit, pl, pc), 8192, &full_bar[s]);
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer ------------------------------
    if (lane == 0) {
      const uint32_t idesc = make_idesc(128, P::BN, P::A_MN, P::B_MN, FmtOf<P>::get(prm));
      constexpr int ORD = (P::PA > P::PB ? P::PA : P::PB) - 1;
      uint32_t it = 0, unit = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int nsub = P::num_subs(prm, item);
        for (int sub = 0; sub < nsub; ++sub, ++unit) {
          const int acc = unit & 1;
          const uint32_t aph = (unit >> 1) & 1;
          mbar_wait(&tempty_bar[acc], aph ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * P::BN;
          const int kn = P::k_iters(prm, item, sub);
          for (int kit = 0; kit < kn; ++kit, ++it) {
            const int s = it % NST;
            const uint32_t ph = (it / NST) & 1;
            mbar_wait(&full_bar[s], ph);
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + s * Cfg::STAGE_BYTES);
            const uint32_t sb = sa + P::PA * Cfg::A_BYTES;
            bool first = (kit == 0);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
              for (int pa = 0; pa < P::PA; ++pa) {
#pragma unroll
                for (int pb = 0; pb < P::PB; ++pb) {
                  if (pa + pb > ORD) continue;
                  const uint32_t a_addr = sa + pa * Cfg::A_BYTES + (P::A_MN ? ks * 2048 : ks * 32);
                  const uint32_t b_addr = sb + pb * Cfg::B_BYTES + (P::B_MN ? ks * 2048 : ks * 32);
                  const uint64_t da = make_smem_desc(a_addr, P::A_MN ? 8192 : 16, 1024);
                  const uint64_t db = make_smem_desc(b_addr, P::B_MN ? 8192 : 16, 1024);
                  umma_bf16(d_tmem, da, db, idesc, first ? 0u : 1u);
                  first = false;
                }
              }
            }
            umma_commit(&empty_bar[s]);        // frees the ring slot when these MMAs retire
          }
          umma_commit(&tfull_bar[acc]);        // accumulator complete -> epilogue
        }
      }
    }
  } else {
    // ------------------------------ epilogue (warps 2..) ------------------------------
    // a warp may only read the TMEM lane quarter warp%4; with 8 (16) epilogue warps every further group of four warps
    // takes the next half (quarter) of the accumulator columns
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;       // accumulator row owned by this thread
    constexpr int NCH = (P::BN / 32) / (P::EPI_WARPS / 4);
    const int c_begin = ((warp - 2) >> 2) * NCH;
    typename P::Epi epi;
    uint32_t unit = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int nsub = P::num_subs(prm, item);
      epi.item_begin(prm, item, row);
      for (int sub = 0; sub < nsub; ++sub, ++unit) {
        const int acc = unit & 1;
        const uint32_t aph = (unit >> 1) & 1;
        epi.sub_begin(prm, item, sub, row);
        mbar_wait(&tfull_bar[acc], aph);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * P::BN;
#pragma unroll 1
        for (int c = c_begin; c < c_begin + NCH; ++c) {
          float v[32];
          tmem_ld32(taddr + c * 32, v);
          if (c == c_begin + NCH - 1) {        // this warp's share is read: hand the TMEM stage back
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[acc]);
          }
          epi.chunk(prm, item, sub, row, c * 32, v);
        }
        epi.sub_end(prm, item, sub, row);
      }
      epi.item_end(prm, item, row);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2): two CTAs of a cluster on the two SMs of a TPC compute one 256 x 256 tile.
// Each CTA stages its own 128 rows of A and only HALF of B (128 of the 256 B rows), so the shared-memory fill
// per MMA drops from 96 KB to 64 KB per k-block -- the L2 -> SM fabric (about 42 B/clk/SM), not the tensor pipe,
// is what bounds the single-CTA kernel -- and the ring deepens to 3 stages.
//   * logical item of CTA `rank` in cluster work item `item2` is 2 * item2 + rank (policies order their items so)
//   * rank 0 issues every tcgen05.mma.cta_group::2; completion is multicast to both CTAs' barriers
//   * rank 1 relays "my stage is full" and "my epilogue drained the accumulator" to rank 0's barriers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_u32(bar)),
      "r"(cta)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t addr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma2_commit_mc(uint64_t* bar) {   // arrive on `bar` in both CTAs of the pair
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(mask)
               : "memory");
}

// Triangular operands inside a diagonal tile: a policy with `VARN != 0` tells the pair kernel, per k-block, how many
// of the 256 accumulator columns can be non-zero (`ncols`, a multiple of 64).  The issuer then runs a narrower MMA on
// the column window [128 - N/2, 128 + N/2) of the accumulator -- each CTA of the pair supplies the N/2 B rows that
// sit next to the centre, so the policy orders its columns with the most-needed groups in the middle -- and the
// producer stages only those rows (`load_b`).  `kblock` lets the policy walk the k-blocks so that the first one is
// full width (it initialises the whole accumulator).
template <class P, class = void>
struct VarN { static constexpr int value = 0; };
template <class P>
struct VarN<P, std::void_t<decltype(P::VARN)>> { static constexpr int value = P::VARN; };

template <class P, bool = (VarN<P>::value != 0)>
struct VarCtx {
  struct type {};
  __device__ static type make(const typename P::Params&, int, int) { return type{}; }
};
template <class P>
struct VarCtx<P, true> {
  using type = typename P::Ctx;
  __device__ static type make(const typename P::Params& p, int item, int sub) { return P::make_ctx(p, item, sub); }
};

// Segmented accumulation (policy declares `SEGK = 1` and Params::segk): every 64-deep k-block of a tile is
// accumulated in a FRESH TMEM buffer (the two buffers alternate per k-block instead of per tile), the epilogue warps
// drain each k-block and add it into fp32 registers with round-to-nearest, and the policy's epilogue runs on the register
// sums.  Inside a k-block the small correction products (hi*lo, lo*hi) are issued first, while the accumulator is still
// small, and the hi*hi products last.  Why: the tensor pipe's fp32 accumulation costs about one ulp of the running sum
// per MMA (measured: the marginal variance off by 2.3e-7 relative, random -- which the model turns into 2e-4 of
// gradient error because it uses that variance as the *scale* of the guide's draw, sparse_gdrf.py:403-405); with 12
// MMAs per k-block and a dense contraction that is 4 ulp of T.  Segmented, an MMA's error is an ulp of ONE k-block's
// share of T, and only 4 MMAs per k-block meet a non-negligible accumulator.
template <class P, class = void>
struct SegK { static constexpr int value = 0; };
template <class P>
struct SegK<P, std::void_t<decltype(P::SEGK)>> { static constexpr int value = P::SEGK; };
template <class P, bool = (SegK<P>::value != 0)>
struct SegOn { __device__ static bool get(const typename P::Params&) { return false; } };
template <class P>
struct SegOn<P, true> { __device__ static bool get(const typename P::Params& p) { return p.segk != 0; } };

// MMA order inside a 64-deep k-block.  Default: per 16-deep k-step, (hi, hi), (hi, lo), (lo, hi) back to back -- two of
// the three read the same A slice, which the tensor pipe reuses, and the kernels run close to the shared-memory
// bandwidth (TMA fill + operand reads), so this order is 15-20 % faster.  A policy with `corr_first(prm, kn)` may ask,
// per tile, for the small correction products of the whole k-block first and the hi * hi products last: an MMA costs
// about an ulp of the accumulator it adds into, and the accumulator is smallest before the k-block's main products
// arrive (measured on the forward row-norm contraction at M = 256: noise of the marginal variance 2.3e-7 -> 1.3e-7).
template <class P, class = void>
struct CorrFirst { __device__ static bool get(const typename P::Params&, int) { return false; } };
template <class P>
struct CorrFirst<P, std::void_t<decltype(&P::corr_first)>> {
  __device__ static bool get(const typename P::Params& p, int kn) { return P::corr_first(p, kn); }
};

// one k-block's MMAs (pair kernel); CORR: corrections first
template <class P, class Cfg, bool CORR>
__device__ __forceinline__ void issue_kblock2(uint32_t sa, uint32_t sb, uint32_t d, uint32_t idesc, bool first) {
  constexpr int ORD = (P::PA > P::PB ? P::PA : P::PB) - 1;
#pragma unroll
  for (int phase = 0; phase < (CORR ? 2 : 1); ++phase) {
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
      for (int pa = 0; pa < P::PA; ++pa) {
#pragma unroll
        for (int pb = 0; pb < P::PB; ++pb) {
          if (pa + pb > ORD) continue;
          if (CORR && ((pa + pb == 0) != (phase == 1))) continue;
          const uint32_t a_addr = sa + pa * Cfg::A_BYTES + (P::A_MN ? ks * 2048 : ks * 32);
          const uint32_t b_addr = sb + pb * Cfg::B_BYTES + (P::B_MN ? ks * 2048 : ks * 32);
          const uint64_t da = make_smem_desc(a_addr, P::A_MN ? 8192 : 16, 1024);
          const uint64_t db = make_smem_desc(b_addr, P::B_MN ? 8192 : 16, 1024);
          umma2_f16(d, da, db, idesc, first ? 0u : 1u);
          first = false;
        }
      }
    }
  }
}

template <class P>
struct Gemm2Cfg {
  static_assert(P::BN == 256, "pair kernel computes 256 x 256 tiles");
  static constexpr int A_BYTES = 16384;                  // one plane, this CTA's 128 rows x 64 k
  static constexpr int B_BYTES = 16384;                  // one plane, this CTA's half of B (128 rows x 64 k)
  static constexpr int STAGE_BYTES = P::PA * A_BYTES + P::PB * B_BYTES;
  static constexpr int NSTAGES = (GEMM_SMEM_BUDGET / STAGE_BYTES) < 4 ? (GEMM_SMEM_BUDGET / STAGE_BYTES) : 4;
  static constexpr int SMEM_BYTES = NSTAGES * STAGE_BYTES + 1024 + 256;
  static constexpr int TMEM_COLS = 512;
  static_assert(NSTAGES >= 2, "need at least a double-buffered ring");
};

template <class P>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(gemm_threads<P>(), 1)
    gemm_tc2_kernel(const __grid_constant__ typename P::Params prm, int n_items1) {
  using Cfg = Gemm2Cfg<P>;
  constexpr int NST = Cfg::NSTAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(smem + NST * Cfg::STAGE_BYTES);
  uint64_t* full_bar = bars;                     // [NST] this CTA's stage landed
  uint64_t* empty_bar = bars + NST;              // [NST] stage consumed (multicast commit from rank 0)
  uint64_t* peer_full_bar = bars + 2 * NST;      // [NST] (rank 0 only) rank 1's stage landed
  uint64_t* tfull_bar = bars + 3 * NST;          // [2]   accumulator complete (multicast commit)
  uint64_t* tempty_bar = bars + 3 * NST + 2;     // [2]   (rank 0 only) both epilogues drained the accumulator
  uint32_t* tmem_slot = (uint32_t*)(bars + 3 * NST + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1;
  const int n_clusters = gridDim.x >> 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NST; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
      mbar_init(&peer_full_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], 2 * P::EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc2(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int n_items2 = (n_items1 + 1) >> 1;
  constexpr int B_HALF_K = 1;                    // K-major: one 128-row piece per CTA
  constexpr int B_HALF_MN = 2;                   // MN-major: two 64-column pieces per CTA

  if (warp == 0) {
    // ------------------------------ bulk-copy producer (both CTAs) ------------------------------
    if (lane == 0) {
      uint32_t it = 0;
      for (int item2 = cluster_id; item2 < n_items2; item2 += n_clusters) {
        const int item = min(2 * item2 + (int)rank, n_items1 - 1);
        const int nsub = P::num_subs(prm, item);
        for (int sub = 0; sub < nsub; ++sub) {
          const int kn = P::k_iters(prm, item, sub);
          const auto octx = OperandCtx<P>::make(prm, item, sub);
          const auto vctx = VarCtx<P>::make(prm, item, sub);   // operand base addresses of this (item, sub), decoded once
          for (int kit = 0; kit < kn; ++kit, ++it) {
            const int s = it % NST;
            const uint32_t ph = (it / NST) & 1;
            mbar_wait(&empty_bar[s], ph ^ 1);
            uint8_t* st = smem + s * Cfg::STAGE_BYTES;
            if constexpr (VarN<P>::value != 0) {
              const int kb = P::kblock(kit, kn);
              const int ncols = P::ncols(prm, kb, kn);
              mbar_arrive_expect_tx(&full_bar[s], P::PA * Cfg::A_BYTES + P::PB * (ncols >> 1) * 128);
#pragma unroll
              for (int pl = 0; pl < P::PA; ++pl)
                bulk_g2s(st + pl * Cfg::A_BYTES, P::a_at(prm, vctx, kb, pl), 16384, &full_bar[s]);
#pragma unroll
              for (int pl = 0; pl < P::PB; ++pl)
                P::load_b(prm, vctx, kb, pl, (int)rank, ncols, st + P::PA * Cfg::A_BYTES + pl * Cfg::B_BYTES,
                          &full_bar[s]);
              continue;
            }
            mbar_arrive_expect_tx(&full_bar[s], Cfg::STAGE_BYTES);
#pragma unroll
            for (int pl = 0; pl < P::PA; ++pl) {
              uint8_t* dst = st + pl * Cfg::A_BYTES;
              if (!P::A_MN) {
                bulk_g2s(dst, OperandCtx<P>::a(prm, octx, kit, pl, 0), 16384, &full_bar[s]);
              } else {
#pragma unroll
                for (int pc = 0; pc < 2; ++pc)
                  bulk_g2s(dst + pc * 8192, OperandCtx<P>::a(prm, octx, kit, pl, pc), 8192, &full_bar[s]);
              }
            }
#pragma unroll
            for (int pl = 0; pl < P::PB; ++pl) {
              uint8_t* dst = st + P::PA * Cfg::A_BYTES + pl * Cfg::B_BYTES;
              if (!P::B_MN) {
                bulk_g2s(dst, OperandCtx<P>::b(prm, octx, kit, pl, (int)rank * B_HALF_K), 16384, &full_bar[s]);
              } else {
#pragma unroll
                for (int pc = 0; pc < B_HALF_MN; ++pc)
                  bulk_g2s(dst + pc * 8192, OperandCtx<P>::b(prm, octx, kit, pl, (int)rank * B_HALF_MN + pc), 8192,
                           &full_bar[s]);
              }
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      if (rank == 0) {
        // ------------------------------ MMA issuer (leader CTA) ------------------------------
        const int fmt = FmtOf<P>::get(prm);
        const uint32_t idesc = make_idesc(256, 256, P::A_MN, P::B_MN, fmt);
        constexpr int ORD = (P::PA > P::PB ? P::PA : P::PB) - 1;
        uint32_t it = 0, unit = 0;
        for (int item2 = cluster_id; item2 < n_items2; item2 += n_clusters) {
          const int item = min(2 * item2, n_items1 - 1);
          const int nsub = P::num_subs(prm, item);
          for (int sub = 0; sub < nsub; ++sub) {
            const int kn = P::k_iters(prm, item, sub);
            if constexpr (SegK<P>::value != 0) if (SegOn<P>::get(prm)) {
              // one TMEM buffer per k-block; corrections first, hi * hi last
              for (int kit = 0; kit < kn; ++kit, ++it, ++unit) {
                const int acc = unit & 1;
                const uint32_t aph = (unit >> 1) & 1;
                mbar_wait(&tempty_bar[acc], aph ^ 1);
                const int s = it % NST;
                const uint32_t ph = (it / NST) & 1;
                mbar_wait(&full_bar[s], ph);
                mbar_wait(&peer_full_bar[s], ph);
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + s * Cfg::STAGE_BYTES);
                const uint32_t sb = sa + P::PA * Cfg::A_BYTES;
                uint32_t idesc_k = idesc, d_k = tmem_base + acc * 256;
                if constexpr (VarN<P>::value != 0) {
                  const int ncols = P::ncols(prm, P::kblock(kit, kn), kn);
                  idesc_k = make_idesc(256, ncols, P::A_MN, P::B_MN, fmt);
                  d_k += (128 - (ncols >> 1));
                }
                issue_kblock2<P, Cfg, true>(sa, sb, d_k, idesc_k, true);
                umma2_commit_mc(&empty_bar[s]);
                umma2_commit_mc(&tfull_bar[acc]);
              }
              continue;
            }
            const int acc = unit & 1;
            const uint32_t aph = (unit >> 1) & 1;
            ++unit;
            mbar_wait(&tempty_bar[acc], aph ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * 256;
            const bool corr_first = CorrFirst<P>::get(prm, kn);
            for (int kit = 0; kit < kn; ++kit, ++it) {
              const int s = it % NST;
              const uint32_t ph = (it / NST) & 1;
              mbar_wait(&full_bar[s], ph);
              mbar_wait(&peer_full_bar[s], ph);
              tc_fence_after();
              const uint32_t sa = smem_u32(smem + s * Cfg::STAGE_BYTES);
              const uint32_t sb = sa + P::PA * Cfg::A_BYTES;
              bool first = (kit == 0);
              uint32_t idesc_k = idesc, d_k = d_tmem;
              if constexpr (VarN<P>::value != 0) {
                const int ncols = P::ncols(prm, P::kblock(kit, kn), kn);   // first k-block is full width by contract
                idesc_k = make_idesc(256, ncols, P::A_MN, P::B_MN, fmt);
                d_k = d_tmem + (128 - (ncols >> 1));
              }
              if (corr_first) issue_kblock2<P, Cfg, true>(sa, sb, d_k, idesc_k, first);
              else issue_kblock2<P, Cfg, false>(sa, sb, d_k, idesc_k, first);
              umma2_commit_mc(&empty_bar[s]);
            }
            umma2_commit_mc(&tfull_bar[acc]);
          }
        }
      } else {
        // ------------------------------ stage-full relay (rank 1 -> rank 0) ------------------------------
        uint32_t it = 0;
        for (int item2 = cluster_id; item2 < n_items2; item2 += n_clusters) {
          const int item = min(2 * item2 + 1, n_items1 - 1);
          const int nsub = P::num_subs(prm, item);
          for (int sub = 0; sub < nsub; ++sub) {
            const int kn = P::k_iters(prm, item, sub);
            for (int kit = 0; kit < kn; ++kit, ++it) {
              const int s = it % NST;
              const uint32_t ph = (it / NST) & 1;
              mbar_wait(&full_bar[s], ph);
              mbar_arrive_remote(&peer_full_bar[s], 0);
            }
          }
        }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------ epilogue (warps 2.., both CTAs) ------------------------------
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    constexpr int NCH = 8 / (P::EPI_WARPS / 4);
    const int c_begin = ((warp - 2) >> 2) * NCH;
    typename P::Epi epi;
    uint32_t unit = 0;
    for (int item2 = cluster_id; item2 < n_items2; item2 += n_clusters) {
      const int item_raw = 2 * item2 + (int)rank;
      const bool valid = item_raw < n_items1;
      const int item = valid ? item_raw : n_items1 - 1;
      const int nsub = P::num_subs(prm, item);
      if (valid) epi.item_begin(prm, item, row);
      for (int sub = 0; sub < nsub; ++sub) {
        if (valid) epi.sub_begin(prm, item, sub, row);
        if constexpr (SegK<P>::value != 0) if (SegOn<P>::get(prm)) {
          // segmented accumulation: drain every k-block's TMEM buffer into fp32 registers (round-to-nearest adds)
          static_assert(SegK<P>::value == 0 || NCH == 4, "segmented accumulation: 8 epilogue warps x 128 columns");
          float sum[4][32];
#pragma unroll
          for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int j = 0; j < 32; ++j) sum[c][j] = 0.f;
          const int kn = P::k_iters(prm, item, sub);
          for (int kit = 0; kit < kn; ++kit, ++unit) {
            const int acc = unit & 1;
            const uint32_t aph = (unit >> 1) & 1;
            int w0 = 0, w1 = 256;                              // accumulator columns this k-block's MMAs wrote
            if constexpr (VarN<P>::value != 0) {
              const int ncols = P::ncols(prm, P::kblock(kit, kn), kn);
              w0 = 128 - (ncols >> 1);
              w1 = 128 + (ncols >> 1);
            }
            mbar_wait(&tfull_bar[acc], aph);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * 256;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const int c0 = (c_begin + c) * 32;
              if (c0 >= w0 && c0 + 32 <= w1) {                 // warp-uniform
                float v[32];
                tmem_ld32(taddr + c0, v);
#pragma unroll
                for (int j = 0; j < 32; ++j) sum[c][j] += v[j];
              }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (rank == 0) mbar_arrive(&tempty_bar[acc]);
              else mbar_arrive_remote(&tempty_bar[acc], 0);
            }
          }
          if (valid) {
#pragma unroll
            for (int c = 0; c < 4; ++c) epi.chunk(prm, item, sub, row, (c_begin + c) * 32, sum[c]);
            epi.sub_end(prm, item, sub, row);
          }
          continue;
        }
        const int acc = unit & 1;
        const uint32_t aph = (unit >> 1) & 1;
        ++unit;
        mbar_wait(&tfull_bar[acc], aph);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * 256;
#pragma unroll 1
        for (int c = c_begin; c < c_begin + NCH; ++c) {
          float v[32];
          tmem_ld32(taddr + c * 32, v);
          if (c == c_begin + NCH - 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (rank == 0) mbar_arrive(&tempty_bar[acc]);
              else mbar_arrive_remote(&tempty_bar[acc], 0);
            }
          }
          if (valid) epi.chunk(prm, item, sub, row, c * 32, v);
        }
        if (valid) epi.sub_end(prm, item, sub, row);
      }
      if (valid) epi.item_end(prm, item, row);
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, Cfg::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------
// plain-FMA checker with the same policy interface (tests only; 128 threads, thread = row)
// ------------------------------------------------------------------------------------------
template <class P>
__global__ void __launch_bounds__(32 * P::EPI_WARPS) gemm_ref_kernel(const typename P::Params prm) {
  const int row = threadIdx.x & 127;
  constexpr int NCH = (P::BN / 32) / (P::EPI_WARPS / 4);
  const int c_begin = (threadIdx.x >> 7) * NCH;
  const int n_items = P::num_items(prm);
  const int fmt = FmtOf<P>::get(prm);
  typename P::Epi epi;
  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int nsub = P::num_subs(prm, item);
    epi.item_begin(prm, item, row);
    for (int sub = 0; sub < nsub; ++sub) {
      epi.sub_begin(prm, item, sub, row);
      const int kn = P::k_iters(prm, item, sub);
      for (int c = c_begin; c < c_begin + NCH; ++c) {
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0.f;
        for (int kit = 0; kit < kn; ++kit) {
          for (int kk = 0; kk < 64; ++kk) {
            float a = 0.f;
            for (int pl = P::PA - 1; pl >= 0; --pl) {
              const bf16* src = P::A_MN ? P::a_src(prm, item, sub, kit, pl, row >> 6) + tile_off(kk, row & 63)
                                        : P::a_src(prm, item, sub, kit, pl, 0) + tile_off(row, kk);
              a += plane_to_float(src, fmt);
            }
            for (int j = 0; j < 32; ++j) {
              const int col = c * 32 + j;
              float b = 0.f;
              for (int pl = P::PB - 1; pl >= 0; --pl) {
                const bf16* src = P::B_MN ? P::b_src(prm, item, sub, kit, pl, col >> 6) + tile_off(kk, col & 63)
                                          : P::b_src(prm, item, sub, kit, pl, col >> 7) + tile_off(col & 127, kk);
                b += plane_to_float(src, fmt);
              }
              v[j] = fmaf(a, b, v[j]);
            }
          }
        }
        epi.chunk(prm, item, sub, row, c * 32, v);
      }
      epi.sub_end(prm, item, sub, row);
    }
    epi.item_end(prm, item, row);
  }
}

template <class P>
inline cudaError_t launch_gemm(const typename P::Params& prm, int n_items, int num_sms, bool use_ref,
                               cudaStream_t stream) {
  if (n_items <= 0) return cudaSuccess;
  if (use_ref) {
    gemm_ref_kernel<P><<<n_items < 4096 ? n_items : 4096, 32 * P::EPI_WARPS, 0, stream>>>(prm);
    return cudaGetLastError();
  }
  using Cfg = GemmCfg<P>;
  cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
  if (e != cudaSuccess) return e;
  const int grid = n_items < num_sms ? n_items : num_sms;
  gemm_tc_kernel<P><<<grid, gemm_threads<P>(), Cfg::SMEM_BYTES, stream>>>(prm);
  return cudaGetLastError();
}

template <class P>
inline cudaError_t launch_gemm2(const typename P::Params& prm, int n_items1, int num_sms, cudaStream_t stream) {
  if (n_items1 <= 0) return cudaSuccess;
  using Cfg = Gemm2Cfg<P>;
  cudaError_t e = cudaFuncSetAttribute(gemm_tc2_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
  if (e != cudaSuccess) return e;
  const int n_items2 = (n_items1 + 1) / 2;
  int clusters = num_sms / 2;
  if (clusters > n_items2) clusters = n_items2;
  gemm_tc2_kernel<P><<<2 * clusters, gemm_threads<P>(), Cfg::SMEM_BYTES, stream>>>(prm, n_items1);
  return cudaGetLastError();
}

}  // namespace gdrf
