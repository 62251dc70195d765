// C ABI of gdrf_b200 (include/gdrf_b200.h): workspace planning and the per-step kernel schedule.
#include "../../include/gdrf_b200.h"

#include <atomic>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "gemm_tc.cuh"
#include "likelihood_tc.cuh"
#include "linalg.cuh"
#include "optimizer.cuh"
#include "policies.cuh"
#include "stages.cuh"
#include "streaming.cuh"

using namespace gdrf;

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, const char* a = "", long long b = 0) {
  snprintf(g_err, sizeof(g_err), fmt, a, b);
  return code;
}
#define CU(x)                                                                         \
  do {                                                                                \
    cudaError_t e_ = (x);                                                             \
    if (e_ != cudaSuccess) return fail(2, "CUDA error: %s (line %lld)", cudaGetErrorString(e_), __LINE__); \
  } while (0)
#define LAUNCH_CHECK() do { ++g_launches; CU(cudaGetLastError()); } while (0)

constexpr int DEFAULT_SMS = 148;

// ---- optional instrumentation (bench.py): kernel-launch counter and CUDA-event timing of the contractions ----
// The only process-global state of the library (include/gdrf_b200.h, "Instrumentation"): an atomic launch counter and,
// while profiling is enabled, a mutex-guarded list of event pairs that grows with the run (nothing is dropped).
std::atomic<long long> g_launches{0};
std::atomic<bool> g_profile{false};
enum { PK_G1 = 0, PK_G2F, PK_G2B, PK_G3, PK_G4, PK_G5, PK_G6, PK_COUNT };
struct ProfRec { int kind; cudaEvent_t a, b; };
std::mutex g_prof_mu;
std::vector<ProfRec> g_prof;
std::vector<cudaEvent_t> g_event_pool;

cudaEvent_t prof_event() {   // g_prof_mu held
  if (!g_event_pool.empty()) { cudaEvent_t e = g_event_pool.back(); g_event_pool.pop_back(); return e; }
  cudaEvent_t e; cudaEventCreate(&e); return e;
}
struct ProfScope {
  cudaEvent_t b = nullptr; cudaStream_t st;
  ProfScope(int kind, cudaStream_t s) : st(s) {
    if (g_profile.load(std::memory_order_relaxed)) {
      std::lock_guard<std::mutex> lk(g_prof_mu);
      ProfRec r{kind, prof_event(), prof_event()};
      g_prof.push_back(r);
      b = r.b;
      cudaEventRecord(r.a, st);
    }
  }
  ~ProfScope() { if (b) cudaEventRecord(b, st); }
};

struct Plan {
  // dims
  int D, M, Mp, MB, JT, MT, K, V;
  long long ncp;          // padded rows per chunk (stride of the [K][ncp] arrays)
  int chunk_rows;
  // persistent
  size_t acc, ck, du, dphi, dz, c5, phisum, ps;
  size_t L64, Linv64, tmpA, tmpB, dinv, dinv32, probe;
  size_t linv_pl, linv16_pl, st_pl, w16_pl, u16_pl, phit_pl, lfact;   // st_pl: 4 planes -- bf16 mode: ST (3); fp16 mode: ST16 permuted (2) | ST16N (2)
  // per chunk
  size_t kxz_pl, w_pl, tp_pl, wg_pl, dwf, dphi_part;   // dwt aliases kxz; dkxz aliases dw
  size_t srow, arow, cnt, gv0, floc, cs, wsq, q, fvar, theta, g_loc, g2, g1;   // cs | wsq | q are contiguous: one memset
  size_t total;
  long long zero_bytes;   // [acc .. c5] contiguous region cleared every step
};

size_t bump(size_t& off, size_t bytes) {
  size_t o = off;
  off += (bytes + 255) & ~(size_t)255;
  return o;
}

int make_plan(const gdrf_shape* s, Plan& p) {
  if (!s) return fail(1, "null shape%s");
  if (s->d < 1 || s->d > MAX_D) return fail(1, "d must be in [1, 8]%s");
  if (s->m < 1 || s->m > 4096) return fail(1, "m must be in [1, 4096]%s");
  if (s->k < 1 || s->k > 128) return fail(1, "k must be in [1, 128]%s");
  if (s->v < 1) return fail(1, "v must be positive%s");
  if (s->n_local < 0) return fail(1, "n_local must be non-negative%s");
  if (s->ls_dim != 1 && s->ls_dim != s->d) return fail(1, "ls_dim must be 1 or d%s");
  if (s->kernel_id < 0 || s->kernel_id > 4) return fail(1, "unknown kernel_id%s");
  if (s->chunk_rows < 0 || (s->chunk_rows % 256) != 0) return fail(1, "chunk_rows must be a multiple of 256%s");
  if (s->n_particles < 0 || s->n_particles > 1024) return fail(1, "n_particles must be in [0, 1024]%s");
  p.D = s->d; p.M = s->m; p.K = s->k; p.V = s->v;
  p.Mp = (int)round_up_ll(s->m, 256);
  p.MB = p.Mp / 64; p.JT = p.Mp / 256; p.MT = p.Mp / 128;
  long long chunk = s->chunk_rows ? s->chunk_rows : (long long)DEFAULT_SMS * 128;
  const long long nneed = round_up_ll(s->n_local > 0 ? s->n_local : 1, 256);
  if (chunk > nneed) chunk = nneed;
  // full chunks of one 256-row tile per CTA pair plus one short last chunk: G2 / G3 / G6 / G5 cut the items of a short
  // chunk so that it still fills the machine (policies.cuh: ksplit, isplit, G6's tail pieces), so its cost is
  // proportional to its rows -- equalised chunks would instead run every round with idle pairs
  p.chunk_rows = (int)chunk;
  p.ncp = chunk;
  const size_t Mp2 = (size_t)p.Mp * p.Mp;
  size_t off = 0;
  p.acc = bump(off, sizeof(double) * ACC_HEAD);
  p.ck = bump(off, sizeof(double) * p.K);
  p.du = bump(off, sizeof(double) * (size_t)p.K * p.M);
  p.dphi = bump(off, sizeof(double) * (size_t)p.K * p.V);
  p.dz = bump(off, sizeof(double) * (size_t)p.M * p.D);
  p.c5 = bump(off, sizeof(double) * Mp2);
  p.zero_bytes = (long long)off;
  p.phisum = bump(off, sizeof(float) * p.K);
  p.ps = bump(off, sizeof(unsigned) * PS_COUNT);
  p.L64 = bump(off, sizeof(double) * Mp2);
  p.Linv64 = bump(off, sizeof(double) * Mp2);
  p.tmpA = bump(off, sizeof(double) * Mp2);
  p.tmpB = bump(off, sizeof(double) * Mp2);
  p.dinv = bump(off, sizeof(double) * (size_t)p.Mp * NB);
  p.dinv32 = bump(off, sizeof(float) * (size_t)p.Mp * NB);
  p.probe = bump(off, sizeof(float) * GDRF_PROBE_MAX * (2 * Mp2 + (size_t)p.Mp * NB));   // gdrf_jitter_probe
  p.linv_pl = bump(off, sizeof(bf16) * 3 * Mp2);
  p.linv16_pl = bump(off, sizeof(bf16) * 2 * Mp2);
  p.u16_pl = bump(off, sizeof(bf16) * 2 * 256 * (size_t)p.Mp);
  p.phit_pl = bump(off, sizeof(bf16) * 3 * 64 * (size_t)round_up_ll(p.V, 128));
  p.lfact = bump(off, sizeof(float) * LT_LFACT);
  p.st_pl = bump(off, sizeof(bf16) * 4 * (size_t)p.K * Mp2);
  const size_t nm = (size_t)p.ncp * p.Mp;
  p.kxz_pl = bump(off, sizeof(bf16) * 3 * nm);
  p.w_pl = bump(off, sizeof(bf16) * 3 * nm);
  p.w16_pl = bump(off, sizeof(bf16) * 2 * nm);
  p.tp_pl = bump(off, sizeof(bf16) * 2 * nm * p.K);
  p.wg_pl = bump(off, sizeof(bf16) * 2 * nm * p.K);
  p.dwf = bump(off, sizeof(float) * nm);
  p.dphi_part = bump(off, p.K <= 64 ? sizeof(float) * (size_t)(p.ncp / 128) * p.K * round_up_ll(p.V, 128) : 0);
  p.srow = bump(off, sizeof(float) * p.ncp);
  p.arow = bump(off, sizeof(float) * p.ncp);
  p.cnt = bump(off, sizeof(float) * p.ncp);
  p.gv0 = bump(off, sizeof(float) * p.ncp);
  const size_t kn = (size_t)p.K * p.ncp;
  p.floc = bump(off, sizeof(double) * kn);
  p.cs = bump(off, sizeof(unsigned) * CS_COUNT);     // 256 bytes; wsq and q follow immediately: one memset per chunk
  p.wsq = bump(off, sizeof(double) * p.ncp);
  p.q = bump(off, sizeof(double) * kn);
  p.fvar = bump(off, sizeof(float) * kn);
  p.theta = bump(off, sizeof(float) * kn);
  p.g_loc = bump(off, sizeof(float) * kn);
  p.g2 = bump(off, sizeof(float) * kn);
  p.g1 = bump(off, sizeof(float) * kn);
  p.total = off;
  return 0;
}

template <typename T>
T* at(void* ws, size_t off) { return reinterpret_cast<T*>(reinterpret_cast<char*>(ws) + off); }

PlaneMat plane_mat(void* ws, size_t off, long long rows, long long cols) {
  PlaneMat m;
  m.base = at<bf16>(ws, off);
  m.row_tiles = (int)(rows / 128);
  m.col_blocks = (int)(cols / 64);
  m.plane_stride = rows * cols;
  return m;
}

// operand planes of S in the two formats (they share Plan::st_pl)
PlaneMat st_bf16(void* ws, const Plan& p) { return plane_mat(ws, p.st_pl, (long long)p.K * p.Mp, p.Mp); }
PlaneMat st_f16_perm(void* ws, const Plan& p) { return plane_mat(ws, p.st_pl, (long long)p.K * p.Mp, p.Mp); }
PlaneMat st_f16_nat(void* ws, const Plan& p) {
  return plane_mat(ws, p.st_pl + sizeof(bf16) * 2 * (size_t)p.K * p.Mp * p.Mp, (long long)p.K * p.Mp, p.Mp);
}

Hyper make_hyper(const gdrf_shape* s, const gdrf_inputs* in) {
  Hyper hp;
  hp.variance = in->variance; hp.lengthscale = in->lengthscale; hp.noise = in->noise; hp.alpha = in->scale_mixture;
  hp.ls_dim = s->ls_dim; hp.kid = s->kernel_id; hp.D = s->d;
  return hp;
}

int num_sms() {
  int dev = 0, n = DEFAULT_SMS;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  return n > 0 ? n : DEFAULT_SMS;
}

int check_device() {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return fail(3, "no CUDA device: gdrf_b200 has no CPU path%s");
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (major != 10) return fail(3, "gdrf_b200 is built for sm_100a only%s");
  return 0;
}

template <int KPW, int VJ>
int launch_likelihood(const Plan& p, int nc, const gdrf_inputs* in, long long n0, void* ws, int sms, cudaStream_t st,
                      double inv_p) {
  const int VC = VJ * 32;
  const size_t smem = sizeof(float) * lk_smem_floats(p.K, VC);
  CU(cudaFuncSetAttribute(k_likelihood<KPW, VJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int ntiles = (nc + LK_TN - 1) / LK_TN;
  const int grid = ntiles < sms ? ntiles : sms;
  k_likelihood<KPW, VJ><<<grid, LK_THREADS, smem, st>>>(
      nc, (int)p.ncp, p.K, p.V, in->ws + n0 * p.V, at<float>(ws, p.theta), at<float>(ws, p.srow), in->phi,
      at<float>(ws, p.g1), at<float>(ws, p.arow), at<float>(ws, p.cnt), at<double>(ws, p.dphi), at<double>(ws, p.acc),
      inv_p);
  LAUNCH_CHECK();
  return 0;
}

// stage 4 on the tensor pipe (likelihood_tc.cuh) for K <= 64 topics; the CUDA-core kernel otherwise / on request
bool likelihood_on_tensor_pipe(const gdrf_shape* s) { return s->k <= 64 && (s->flags & GDRF_FLAG_LIKELIHOOD_FMA) == 0; }

PlaneMat phit_mat(void* ws, const Plan& p) { return plane_mat(ws, p.phit_pl, round_up_ll(p.V, 128), 64); }

int dispatch_likelihood(const gdrf_shape* s, const Plan& p, int nc, const gdrf_inputs* in, long long n0, void* ws, int sms,
                        cudaStream_t st, double inv_p) {
  if (likelihood_on_tensor_pipe(s)) {
    CU(cudaFuncSetAttribute(k_likelihood_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, LT_SMEM));
    k_likelihood_tc<<<(nc + 127) / 128, LT_THREADS, LT_SMEM, st>>>(
        nc, (int)p.ncp, p.K, p.V, in->ws + n0 * p.V, at<float>(ws, p.theta), at<float>(ws, p.srow), phit_mat(ws, p),
        at<float>(ws, p.g1), at<float>(ws, p.arow), at<float>(ws, p.cnt), at<float>(ws, p.dphi_part),
        at<float>(ws, p.lfact), at<double>(ws, p.acc), inv_p);
    LAUNCH_CHECK();
    const int Vp = (int)round_up_ll(p.V, 128);
    k_reduce_dphi<<<dim3((unsigned)(((long long)p.K * Vp + 255) / 256), 8), 256, 0, st>>>(at<float>(ws, p.dphi_part), (nc + 127) / 128,
                                                                            p.K, p.V, Vp, at<double>(ws, p.dphi), inv_p);
    LAUNCH_CHECK();
    return 0;
  }
  const int kpw = (p.K + 15) / 16;
  const int vj_need = (p.V + 31) / 32;
#define LK(KPW, VJ) return launch_likelihood<KPW, VJ>(p, nc, in, n0, ws, sms, st, inv_p)
  if (kpw <= 1) { if (vj_need <= 4) LK(1, 4); if (vj_need <= 8) LK(1, 8); LK(1, 16); }
  if (kpw <= 2) { if (vj_need <= 4) LK(2, 4); if (vj_need <= 8) LK(2, 8); LK(2, 16); }
  if (kpw <= 4) { if (vj_need <= 4) LK(4, 4); if (vj_need <= 8) LK(4, 8); LK(4, 16); }
  if (vj_need <= 4) LK(8, 4);
  LK(8, 8);
#undef LK
}

// topics per lane for the 8-lanes-per-observation kernels (K <= 128)
#define GDRF_DISPATCH_KQ(K_, F_)       \
  do {                                 \
    if ((K_) <= 8) F_(1);              \
    else if ((K_) <= 16) F_(2);        \
    else if ((K_) <= 32) F_(4);        \
    else if ((K_) <= 64) F_(8);        \
    else F_(16);                       \
  } while (0)

int launch_du(const Plan& p, PlaneMat w, int RT, void* ws, int sms, cudaStream_t st) {
  int tiles_per_cta = (RT * p.MB + 4 * sms - 1) / (4 * sms);   // ~4 CTAs per SM: one CTA's tile rebuild hides under another's FMAs
  if (tiles_per_cta < 1) tiles_per_cta = 1;
  const dim3 grid(p.MB, (RT + tiles_per_cta - 1) / tiles_per_cta);
  const float* g = at<float>(ws, p.g_loc);
  double* du = at<double>(ws, p.du);
#define DU(KGP)                                                                                             \
  do {                                                                                                      \
    const size_t smem = sizeof(float) * (128 * 68 + 8 * KGP * 128);                                         \
    CU(cudaFuncSetAttribute(k_du<KGP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));            \
    k_du<KGP><<<grid, 256, smem, st>>>(w, g, p.K, p.M, RT, (int)p.ncp, tiles_per_cta, du);                  \
  } while (0)
  if (p.K <= 8) DU(1);
  else if (p.K <= 16) DU(2);
  else if (p.K <= 32) DU(4);
  else if (p.K <= 64) DU(8);
  else DU(16);
#undef DU
  LAUNCH_CHECK();
  return 0;
}

// G2 on CTA pairs: how many topic groups to cut a pair tile's work into so that `tiles` pair tiles fill `pairs` CTA
// pairs best: smallest ceil(tiles * s / pairs) / s over the divisors s <= 8 of K (ties: fewer, longer items)
int topic_split(int K, int tiles, int pairs) {
  int best = 1;
  double best_cost = 1e30;
  for (int sp = 1; sp <= 8 && sp <= K; ++sp) {
    if (K % sp) continue;
    const double cost = (double)((tiles * sp + pairs - 1) / pairs) / sp;
    if (cost < best_cost - 1e-9) { best_cost = cost; best = sp; }
  }
  return best;
}

// the four large contractions run on CTA pairs (cta_group::2) unless the checker or the single-CTA kernel is asked for
template <class P>
cudaError_t launch_big(const typename P::Params& g, int n_items, int sms, bool use_ref, bool single_cta, cudaStream_t st) {
  if (use_ref || single_cta) return launch_gemm<P>(g, n_items, sms, use_ref, st);
  return launch_gemm2<P>(g, n_items, sms, st);
}

// the forward contractions (whitening, row norms) issue the correction products of a k-block before its hi * hi products
// (gemm_tc.cuh, CorrFirst) in tiles of at most this many k-blocks: all of them, unless the A/B flag asks for the
// interleaved order
int corr_kn(const gdrf_shape* s) { return (s->flags & GDRF_FLAG_INTERLEAVED_MMAS) ? 0 : (1 << 30); }

// once per call, before the first chunk: the fp16 operand planes of u_loc for f_loc = W u_loc^T on the tensor pipe
int pack_u(const gdrf_shape* s, const gdrf_inputs* in, const Plan& p, void* ws, cudaStream_t st) {
  if (s->flags & GDRF_FLAG_FWD_BF16) return 0;      // bf16 mode keeps the CUDA-core k_floc
  unsigned* ps = at<unsigned>(ws, p.ps);
  CU(cudaMemsetAsync(ps + PS_UMAX, 0, sizeof(unsigned), st));
  k_absmax<<<32, 256, 0, st>>>(in->u_loc, (long long)p.K * p.M, ps + PS_UMAX);
  LAUNCH_CHECK();
  k_pack_u<<<dim3(p.MB, 2), 256, 0, st>>>(in->u_loc, p.K, p.M, plane_mat(ws, p.u16_pl, 256, p.Mp), ps);
  LAUNCH_CHECK();
  return 0;
}

// forward contraction chain of one chunk: Kxz -> W -> f_loc (and q when with_var)
int chunk_forward(const gdrf_shape* s, const gdrf_inputs* in, const Plan& p, void* ws, long long n0, int nc, int RT,
                  bool with_var, bool store_t, int sms, cudaStream_t st) {
  const Hyper hp = make_hyper(s, in);
  PlaneMat kxz = plane_mat(ws, p.kxz_pl, p.ncp, p.Mp);
  PlaneMat w = plane_mat(ws, p.w_pl, p.ncp, p.Mp);
  PlaneMat linv = plane_mat(ws, p.linv_pl, p.Mp, p.Mp);
  PlaneMat stm = st_bf16(ws, p);
  // per-chunk scalars, |W_n|^2 and (with_var) q are accumulated into: cleared together
  CU(cudaMemsetAsync(at<char>(ws, p.cs), 0,
                     with_var ? (p.q - p.cs) + sizeof(double) * (size_t)p.K * p.ncp : (p.q - p.cs), st));
#define GDRF_KXZ_PLANES(DT, KID) \
  k_kxz_planes<DT, KID><<<dim3(p.MB, RT), 256, 0, st>>>(in->xs + n0 * p.D, nc, in->z, p.M, hp, kxz)
  GDRF_DISPATCH_DK(p.D, hp.kid, GDRF_KXZ_PLANES);
#undef GDRF_KXZ_PLANES
  LAUNCH_CHECK();
  {
    auto fill = [&](auto& g) {
      g.kxz = kxz; g.linv = linv; g.w = w; g.w16 = plane_mat(ws, p.w16_pl, p.ncp, p.Mp); g.wsq = at<double>(ws, p.wsq);
      g.RT = RT; g.MB = p.MB; g.corr_kn = corr_kn(s);
    };
    ProfScope ps(PK_G1, st);
    ++g_launches;
    if ((s->flags & (GDRF_FLAG_REF_G1 | GDRF_FLAG_SINGLE_CTA)) != 0) {
      G1T<128>::Params g{}; fill(g);
      CU(launch_gemm<G1T<128>>(g, RT, sms, (s->flags & GDRF_FLAG_REF_G1) != 0, st));
    } else {
      G1T<256>::Params g{}; fill(g);
      CU(launch_gemm2<G1T<256>>(g, RT, sms, st));
    }
  }
  if ((s->flags & (GDRF_FLAG_FWD_BF16 | GDRF_FLAG_SINGLE_CTA | GDRF_FLAG_REF_G1)) == 0) {
    GF::Params g{};      // f_loc = W u_loc^T on the tensor pipe (fp16 pairs)
    g.w = plane_mat(ws, p.w16_pl, p.ncp, p.Mp); g.u = plane_mat(ws, p.u16_pl, 256, p.Mp);
    g.floc = at<double>(ws, p.floc); g.inv_scale = at<float>(ws, p.ps) + PS_SU_INV;
    g.RT = RT; g.MB = p.MB; g.K = p.K; g.ncp = (int)p.ncp;
    ++g_launches;
    CU(launch_gemm2<GF>(g, RT, sms, st));
  } else {
    k_floc<16><<<dim3(RT, (p.K + 15) / 16), 128, 0, st>>>(w, in->u_loc, p.K, p.M, p.MB, at<double>(ws, p.floc), (int)p.ncp);
    LAUNCH_CHECK();
  }
  if (with_var) {
    if (s->flags & GDRF_FLAG_FWD_BF16) {     // 24-bit operands, 6 products
      G2<0>::Params g{};
      g.w = w; g.st = stm; g.tp = plane_mat(ws, p.tp_pl, p.ncp, (long long)p.K * p.Mp);
      g.q = at<double>(ws, p.q); g.store_t = store_t ? 1 : 0; g.RT = RT; g.MB = p.MB; g.K = p.K; g.NT = p.Mp / G2<0>::BN; g.ncp = (int)p.ncp;
      { ProfScope ps(PK_G2F, st); ++g_launches; CU(launch_gemm<G2<0>>(g, RT, sms, (s->flags & GDRF_FLAG_REF_G2) != 0, st)); }
    } else {                                 // fp16 2 x 2 planes (22-bit operands), 3 products
      auto run = [&](auto tag) -> int {
        using P2 = decltype(tag);
        typename P2::Params g{};
        g.w = plane_mat(ws, p.w16_pl, p.ncp, p.Mp); g.st = st_f16_perm(ws, p);
        g.tp = plane_mat(ws, p.tp_pl, p.ncp, (long long)p.K * p.Mp);
        g.q = at<double>(ws, p.q); g.store_t = store_t ? 1 : 0; g.RT = RT; g.MB = p.MB; g.K = p.K; g.NT = p.Mp / P2::BN; g.ncp = (int)p.ncp;
        g.varn = (s->flags & GDRF_FLAG_FULL_WIDTH) ? 0 : 1;
        g.segk = P2::SEGK;
        g.corr_kn = corr_kn(s);
        const bool pairs = (s->flags & (GDRF_FLAG_REF_G2 | GDRF_FLAG_SINGLE_CTA)) == 0;
        g.ksplit = pairs ? topic_split(p.K, (RT + 1) / 2, sms / 2) : 0;
        const int n_items = pairs ? ((RT + 1) / 2) * 2 * g.ksplit : RT;
        ProfScope ps(PK_G2F, st);
        ++g_launches;
        CU(launch_big<P2>(g, n_items, sms, (s->flags & GDRF_FLAG_REF_G2) != 0, (s->flags & GDRF_FLAG_SINGLE_CTA) != 0, st));
        return 0;
      };
      if (s->flags & GDRF_FLAG_SEGMENTED_FWD) { if (int e = run(G2<3>{})) return e; }
      else { if (int e = run(G2<2>{})) return e; }
    }
  }
  return 0;
}

__global__ void k_scale(const float* __restrict__ g, long long n, const float* __restrict__ sdev, float shost,
                        float* __restrict__ dst) {
  const float s = (sdev ? sdev[0] : 1.f) * shost;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[i] = s * g[i];
}

template <typename T>
__global__ void k_export_floc(const double* __restrict__ floc, int ncp, int nc, T* __restrict__ out,
                              long long n_stride) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x, k = blockIdx.y;
  if (n < nc) out[(long long)k * n_stride + n] = (T)floc[(long long)k * ncp + n];
}

// f_var[k, n] = clamp(variance - |W_n|^2, 0) + q[k, n]   (pyro conditional, full_cov=False)
template <typename T>
__global__ void k_export_fvar(const double* __restrict__ q, const double* __restrict__ wsq,
                              const float* __restrict__ variance, int ncp, int nc, T* __restrict__ out,
                              long long n_stride) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x, k = blockIdx.y;
  if (n < nc) out[(long long)k * n_stride + n] = (T)(fmax((double)variance[0] - wsq[n], 0.0) + q[(long long)k * ncp + n]);
}

// status stays >0 on a failed factorisation; otherwise -1 flags that an operand of the fp16 planes may leave the fp16
// range: S itself, T = W S_k (|T| <= |W_n| |S_k[:, j]| <= sqrt(variance M) max|S|), Linv, or W (|W| <= sqrt(variance))
__global__ void k_merge_status(const unsigned* __restrict__ ps, const float* __restrict__ variance, int M,
                               int* __restrict__ status) {
  const float smax = __uint_as_float(ps[PS_SMAX]), lmax = __uint_as_float(ps[PS_LINVMAX]), var = variance[0];
  const bool ok = smax < 60000.f && smax * sqrtf(var * (float)M) < 60000.f && lmax < 60000.f && var < 3.0e9f;
  if (*status == 0 && !ok) *status = -1;
}

// terms as fp64, and -- when the gradient buffer is there -- once more behind the gradient as four (hi, lo) fp32 pairs,
// so that ONE fp32 all-reduce of the flat buffer carries the loss as well (hi + lo restores the value to fp32 accuracy of
// the summed hi parts: the reference's loss is an fp32 number)
__global__ void k_copy_terms(const double* __restrict__ acc, double* __restrict__ terms, float* __restrict__ tail) {
  if (threadIdx.x == 0) {
    const double t[4] = {acc[ACC_LP_MU], acc[ACC_LQ], acc[ACC_LL], acc[ACC_LP_PHI]};
    for (int i = 0; i < 4; ++i) {
      terms[i] = t[i];
      if (tail) {
        const float hi = (float)t[i];
        tail[2 * i] = hi;
        tail[2 * i + 1] = (float)(t[i] - (double)hi);
      }
    }
  }
}

// sum_n sum_v w log(softmax_k(f_loc)[n] . phi[:, v])   (abstract_gdrf.py:113-139); warp per observation
__global__ void __launch_bounds__(256) k_perplexity(int N, int K, int V, const float* __restrict__ floc,
                                                    const int* __restrict__ ws, const float* __restrict__ phi,
                                                    double* __restrict__ out) {
  extern __shared__ float pp_smem[];   // [8 warps][K]
  __shared__ double scratch[32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* th = pp_smem + warp * K;
  double num = 0.0, den = 0.0;
  for (long long n = (long long)blockIdx.x * 8 + warp; n < N; n += (long long)gridDim.x * 8) {
    float mx = -INFINITY;
    for (int k = lane; k < K; k += 32) mx = fmaxf(mx, floc[(long long)k * N + n]);
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
    for (int k = lane; k < K; k += 32) {
      const float e = __expf(floc[(long long)k * N + n] - mx);
      th[k] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    __syncwarp();
    const float inv = 1.f / sum;
    float a = 0.f, c = 0.f;
    for (int v = lane; v < V; v += 32) {
      const int w = ws[n * V + v];
      if (w != 0) {
        float pv = 0.f;
        for (int k = 0; k < K; ++k) pv = fmaf(th[k], phi[(long long)k * V + v], pv);
        a += (float)w * __logf(pv * inv);
        c += (float)w;
      }
    }
    num += (double)warp_sum(a);
    den += (double)warp_sum(c);
    __syncwarp();
  }
  if (lane != 0) num = den = 0.0;
  num = block_sum(num, scratch);
  den = block_sum(den, scratch);
  if (threadIdx.x == 0) {
    atomicAdd(&out[0], num);
    atomicAdd(&out[1], den);
  }
}

template <typename T>
int marginal_moments_impl(const gdrf_shape* s, const gdrf_inputs* in, T* out_floc, T* out_fvar, void* ws,
                                 size_t ws_bytes, gdrf_stream_t stream) {
  Plan p;
  if (int e = make_plan(s, p)) return e;
  if (int e = check_device()) return e;
  if (!in || !ws || !out_floc) return fail(1, "null pointer argument%s");
  if (s->kernel_id == KERNEL_RQ && !in->scale_mixture) return fail(1, "the RationalQuadratic kernel needs in->scale_mixture%s");
  if (out_fvar && !in->u_scale_tril) return fail(1, "the marginal variance needs u_scale_tril%s");
  if (ws_bytes < p.total) return fail(1, "workspace too small%s (need %lld bytes)", "", (long long)p.total);
  cudaStream_t st = (cudaStream_t)stream;
  const int sms = num_sms();
  if (int e = pack_u(s, in, p, ws, st)) return e;
  for (long long n0 = 0; n0 < s->n_local; n0 += p.chunk_rows) {
    const int nc = (int)((s->n_local - n0 < p.chunk_rows) ? (s->n_local - n0) : p.chunk_rows);
    const int RT = (nc + 127) / 128;
    if (int e = chunk_forward(s, in, p, ws, n0, nc, RT, out_fvar != nullptr, false, sms, st)) return e;
    k_export_floc<T><<<dim3((nc + 255) / 256, p.K), 256, 0, st>>>(at<double>(ws, p.floc), (int)p.ncp, nc,
                                                                out_floc + n0, (long long)s->n_local);
    LAUNCH_CHECK();
    if (out_fvar) {
      k_export_fvar<T><<<dim3((nc + 255) / 256, p.K), 256, 0, st>>>(at<double>(ws, p.q), at<double>(ws, p.wsq),
                                                                  in->variance, (int)p.ncp, nc, out_fvar + n0,
                                                                  (long long)s->n_local);
      LAUNCH_CHECK();
    }
  }
  return 0;
}

}  // namespace

extern "C" {

const char* gdrf_last_error(void) { return g_err; }

long long gdrf_launch_count(void) { return g_launches.load(); }

int gdrf_profile_enable(int on) {
  g_profile.store(on != 0);
  return 0;
}

// ms[7], launches[7]: summed CUDA-event time and launch count of G1, G2-forward, k_scale_w, G3, G4, G5, G6
// since the last read.  Synchronises the device.
int gdrf_profile_read(double* ms, long long* launches) {
  if (!ms || !launches) return fail(1, "null pointer argument%s");
  for (int i = 0; i < PK_COUNT; ++i) { ms[i] = 0.0; launches[i] = 0; }
  CU(cudaDeviceSynchronize());
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (const ProfRec& r : g_prof) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess) {
      ms[r.kind] += t;
      launches[r.kind] += 1;
    }
    g_event_pool.push_back(r.a);
    g_event_pool.push_back(r.b);
  }
  g_prof.clear();
  return 0;
}

int gdrf_constrain(const gdrf_shape* s, const float* theta_u, float* theta_c, int learn_z, gdrf_stream_t stream) {
  Plan p;
  if (int e = make_plan(s, p)) return e;
  if (int e = check_device()) return e;
  if (!theta_u || !theta_c) return fail(1, "null pointer argument%s");
  cudaStream_t st = (cudaStream_t)stream;
  const FlatLayout f = make_layout(p.K, p.M, p.V, p.D, s->ls_dim, s->kernel_id == KERNEL_RQ);
  long long blocks = (f.total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  k_constrain<<<(int)blocks, 256, 0, st>>>(f, theta_u, theta_c, learn_z);
  LAUNCH_CHECK();
  k_softmax_rows<<<p.K, 256, 0, st>>>(theta_u + f.oP, theta_c + f.oP, p.V);
  LAUNCH_CHECK();
  return 0;
}

static int adam_common(const gdrf_shape* s, float* theta_u, const float* theta_c, const float* grad, float* m, float* v,
                       float* row_scratch, float lr, float beta1, float beta2, float eps, float weight_decay, float clip,
                       int step, float grad_scale, int learn_z, gdrf_stream_t stream) {
  Plan p;
  if (int e = make_plan(s, p)) return e;
  if (int e = check_device()) return e;
  if (!theta_u || !theta_c || !grad || !m || !v || !row_scratch) return fail(1, "null pointer argument%s");
  if (step < 1) return fail(1, "step counts from 1%s");
  cudaStream_t st = (cudaStream_t)stream;
  const FlatLayout f = make_layout(p.K, p.M, p.V, p.D, s->ls_dim, s->kernel_id == KERNEL_RQ);
  k_phi_rowdot<<<p.K, 256, 0, st>>>(theta_c + f.oP, grad + f.oP, p.V, row_scratch);
  LAUNCH_CHECK();
  AdamHyper h;
  h.lr = lr; h.beta1 = beta1; h.beta2 = beta2; h.eps = eps; h.weight_decay = weight_decay;
  h.bc1 = 1.f - powf(beta1, (float)step);
  h.bc2 = 1.f - powf(beta2, (float)step);
  h.grad_scale = grad_scale;
  h.clip = clip;
  long long blocks = (f.total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  k_adam<<<(int)blocks, 256, 0, st>>>(f, theta_u, theta_c, grad, row_scratch, m, v, h, learn_z);
  LAUNCH_CHECK();
  return 0;
}

int gdrf_adam_step(const gdrf_shape* s, float* theta_u, const float* theta_c, const float* grad, float* m, float* v,
                   float* row_scratch, float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                   float grad_scale, int learn_z, gdrf_stream_t stream) {
  return adam_common(s, theta_u, theta_c, grad, m, v, row_scratch, lr, beta1, beta2, eps, weight_decay, 0.f, step,
                     grad_scale, learn_z, stream);
}

int gdrf_clipped_adam_step(const gdrf_shape* s, float* theta_u, const float* theta_c, const float* grad, float* m,
                           float* v, float* row_scratch, float lr, float beta1, float beta2, float eps,
                           float weight_decay, float clip_norm, int step, float grad_scale, int learn_z,
                           gdrf_stream_t stream) {
  if (!(clip_norm > 0.f)) return fail(1, "clip_norm must be positive%s");
  return adam_common(s, theta_u, theta_c, grad, m, v, row_scratch, lr, beta1, beta2, eps, weight_decay, clip_norm, step,
                     grad_scale, learn_z, stream);
}

int gdrf_gather_rows(const float* xs, const int32_t* ws, const int64_t* index, int64_t n_sel, int64_t n_rows, int32_t d,
                     int32_t v, float* xs_out, int32_t* ws_out, int* dev_status, gdrf_stream_t stream) {
  if (int e = check_device()) return e;
  if (n_sel < 0 || n_rows < 1) return fail(1, "n_sel must be >= 0 and n_rows >= 1%s");
  if (d < 1 || d > MAX_D) return fail(1, "d must be in [1, 8]%s");
  if (v < 1) return fail(1, "v must be positive%s");
  if (n_sel == 0) return 0;
  if (!xs || !ws || !index || !xs_out || !ws_out) return fail(1, "null pointer argument%s");
  cudaStream_t st = (cudaStream_t)stream;
  long long blocks = (n_sel + 7) / 8;                       // one warp per selected row, 8 warps per block
  if (blocks > DEFAULT_SMS * 8) blocks = DEFAULT_SMS * 8;
  const bool vec = (v % 4 == 0) && ((reinterpret_cast<uintptr_t>(ws) | reinterpret_cast<uintptr_t>(ws_out)) % 16 == 0);
  if (vec)
    k_gather_rows<true><<<(int)blocks, 256, 0, st>>>(xs, ws, (const long long*)index, n_sel, n_rows, d, v, xs_out, ws_out,
                                                     dev_status);
  else
    k_gather_rows<false><<<(int)blocks, 256, 0, st>>>(xs, ws, (const long long*)index, n_sel, n_rows, d, v, xs_out, ws_out,
                                                      dev_status);
  LAUNCH_CHECK();
  return 0;
}

const char* gdrf_build_info(void) {
  return "gdrf_b200 sm_100a: tcgen05 split-bf16 contractions (TMEM accumulators, cp.async.bulk operand ring)";
}

int gdrf_workspace_bytes(const gdrf_shape* shape, size_t* out_bytes) {
  Plan p;
  if (int e = make_plan(shape, p)) return e;
  if (!out_bytes) return fail(1, "null out pointer%s");
  *out_bytes = p.total;
  return 0;
}

int gdrf_grad_elems(const gdrf_shape* s, int64_t* out_elems) {
  Plan p;
  if (int e = make_plan(s, p)) return e;
  if (!out_elems) return fail(1, "null out pointer%s");
  *out_elems = (int64_t)s->k * s->m * s->m + (int64_t)s->k * s->m + (int64_t)s->k * s->v + (int64_t)s->m * s->d + 2 +
               s->ls_dim + (s->kernel_id == KERNEL_RQ ? 1 : 0);
  return 0;
}

int gdrf_prologue(const gdrf_shape* s, const gdrf_inputs* in, double jitter, int njitter, void* ws, size_t ws_bytes,
                  gdrf_stream_t stream, int* dev_status) {
  Plan p;
  if (int e = make_plan(s, p)) return e;
  if (int e = check_device()) return e;
  if (!in || !ws || !dev_status) return fail(1, "null pointer argument%s");
  if (s->kernel_id == KERNEL_RQ && !in->scale_mixture) return fail(1, "the RationalQuadratic kernel needs in->scale_mixture%s");
  if (ws_bytes < p.total) return fail(1, "workspace too small%s (need %lld bytes)", "", (long long)p.total);
  if (njitter < 0) return fail(1, "njitter must be >= 0%s");
  cudaStream_t st = (cudaStream_t)stream;
  const Hyper hp = make_hyper(s, in);
  CU(cudaMemsetAsync(dev_status, 0, sizeof(int), st));
  const dim3 g2d(ceil_div(p.Mp, 256), p.Mp);
  double* L = at<double>(ws, p.L64);
  double* Linv = at<double>(ws, p.Linv64);
  double* kuu = at<double>(ws, p.tmpA);
  if (s->flags & GDRF_FLAG_STATUS_ONLY) {
    // only the question jittercholesky asks at this level -- does the reference's fp32 factorisation fail? -- is
    // answered: the fp32 Kuu and its fp32 Cholesky, nothing else (no fp64 values, inverse or operand planes)
    float* k32 = at<float>(ws, p.tmpB);
    float* l32 = k32 + (size_t)p.Mp * p.Mp;
    k_kuu<float><<<g2d, 256, 0, st>>>(in->z, p.M, p.Mp, hp, jitter, njitter, k32);
    LAUNCH_CHECK();
    cholesky<float>(k32, l32, at<float>(ws, p.dinv32), p.Mp, dev_status, st);
    g_launches += p.Mp / NB - 1;
    LAUNCH_CHECK();
    return 0;
  }
  k_kuu<double><<<g2d, 256, 0, st>>>(in->z, p.M, p.Mp, hp, jitter, njitter, kuu);
  LAUNCH_CHECK();
  if (s->flags & GDRF_FLAG_CHOL_FP32_STATUS) {
    // the reference's own fp32 arithmetic decides whether this jitter level "fails" (utils.py:31-37); its
    // factorisation runs in the same launches as the fp64 one (values), in the two halves of tmpB
    float* k32 = at<float>(ws, p.tmpB);
    float* l32 = k32 + (size_t)p.Mp * p.Mp;
    k_kuu<float><<<g2d, 256, 0, st>>>(in->z, p.M, p.Mp, hp, jitter, njitter, k32);
    LAUNCH_CHECK();
    cholesky_both(kuu, L, at<double>(ws, p.dinv), k32, l32, at<float>(ws, p.dinv32), p.Mp, dev_status, st);
  } else {
    cholesky<double>(kuu, L, at<double>(ws, p.dinv), p.Mp, dev_status, st);
  }
  g_launches += p.Mp / NB - 1;          // + the one LAUNCH_CHECK counts
  LAUNCH_CHECK();
  // Kuu's memory (tmpA) is free once it is factorised: scratch of the recursive inverse
  g_launches += tri_inverse(L, at<double>(ws, p.dinv), Linv, kuu, p.Mp, st) - 1;
  LAUNCH_CHECK();
  const bool f16 = (s->flags & GDRF_FLAG_FWD_BF16) == 0;
  unsigned* ps = at<unsigned>(ws, p.ps);
  CU(cudaMemsetAsync(ps, 0, sizeof(unsigned) * PS_COUNT, st));
  PlaneMat linv = plane_mat(ws, p.linv_pl, p.Mp, p.Mp);
  k_pack_linv<<<dim3(p.MB, p.MT), 256, 0, st>>>(Linv, p.M, p.Mp, linv, plane_mat(ws, p.linv16_pl, p.Mp, p.Mp), ps);
  LAUNCH_CHECK();
  if (in->u_scale_tril) {
    // operand planes of S for this step: fp16 pairs (22 bits) unless the caller asked for bf16 (GDRF_FLAG_FWD_BF16);
    // in fp16 mode a value that may leave the fp16 range turns the status into -1, and the caller repeats the
    // prologue and runs the step with GDRF_FLAG_FWD_BF16
    const dim3 grid(p.MB, p.MB, p.K);
    if (f16) k_pack_st<true><<<grid, 256, 0, st>>>(in->u_scale_tril, p.K, p.M, p.Mp, st_bf16(ws, p), st_f16_perm(ws, p), st_f16_nat(ws, p), ps);
    else k_pack_st<false><<<grid, 256, 0, st>>>(in->u_scale_tril, p.K, p.M, p.Mp, st_bf16(ws, p), st_f16_perm(ws, p), st_f16_nat(ws, p), ps);
    LAUNCH_CHECK();
  }
  if (f16) {
    k_merge_status<<<1, 1, 0, st>>>(ps, in->variance, p.M, dev_status);
    LAUNCH_CHECK();
  }
  return 0;
}

int gdrf_jitter_probe(const gdrf_shape* s, const gdrf_inputs* in, double jitter, int njitter_first, int count, void* ws,
                      size_t ws_bytes, gdrf_stream_t stream, int* dev_status) {
  Plan p;
  if (int e = make_plan(s, p)) return e;
  if (int e = check_device()) return e;
  if (!in || !ws || !dev_status) return fail(1, "null pointer argument%s");
  if (s->kernel_id == KERNEL_RQ && !in->scale_mixture) return fail(1, "the RationalQuadratic kernel needs in->scale_mixture%s");
  if (ws_bytes < p.total) return fail(1, "workspace too small%s (need %lld bytes)", "", (long long)p.total);
  if (njitter_first < 0 || count < 1 || count > GDRF_PROBE_MAX) return fail(1, "njitter_first must be >= 0 and count in [1, GDRF_PROBE_MAX]%s");
  cudaStream_t st = (cudaStream_t)stream;
  const Hyper hp = make_hyper(s, in);
  const size_t Mp2 = (size_t)p.Mp * p.Mp;
  float* k32 = at<float>(ws, p.probe);
  float* l32 = k32 + GDRF_PROBE_MAX * Mp2;
  float* d32 = l32 + GDRF_PROBE_MAX * Mp2;
  CU(cudaMemsetAsync(dev_status, 0, sizeof(int) * count, st));
  k_kuu_batch<<<dim3(ceil_div(p.Mp, 256), p.Mp, count), 256, 0, st>>>(in->z, p.M, p.Mp, hp, jitter, njitter_first, k32,
                                                                     (long long)Mp2);
  LAUNCH_CHECK();
  cholesky_batch(k32, l32, d32, p.Mp, count, dev_status, st);
  g_launches += p.Mp / NB - 1;
  LAUNCH_CHECK();
  return 0;
}

// ---- shared by gdrf_elbo_step and gdrf_moments_vjp: everything that follows the per-observation weights ----
struct StepCtx {
  const gdrf_shape* s;
  const gdrf_inputs* in;
  Plan p;
  void* ws;
  float* grad;       // flat gradient buffer (dS accumulates here)
  int sms;
  cudaStream_t st;
  Hyper hp;
  bool fuse_du, f16;
  int fmt;
  PlaneMat kxz, dwt, w, linv, st_b, linv_b, w_b, tpm, wgm;
  unsigned *cs, *ps;
  const float* cs_f;
  float* dwf;
  double* acc;
  G6::Params g6;
};

int make_step_ctx(const gdrf_shape* s, const gdrf_inputs* in, const Plan& p, void* ws, float* grad, cudaStream_t st,
                  bool want_grad, StepCtx& c) {
  c.s = s; c.in = in; c.p = p; c.ws = ws; c.grad = grad; c.st = st;
  c.sms = num_sms();
  c.hp = make_hyper(s, in);
  const int Mp = p.Mp;
  c.kxz = plane_mat(ws, p.kxz_pl, p.ncp, Mp);
  // Kxz is dead once W exists; its planes are reused for dWtot (two planes) plus, in fp16 mode, 128 more columns that
  // carry g_loc for du_loc (G5)
  c.fuse_du = (s->flags & (GDRF_FLAG_FWD_BF16 | GDRF_FLAG_SINGLE_CTA | GDRF_FLAG_REF_G5)) == 0;
  c.dwt = c.fuse_du ? plane_mat(ws, p.kxz_pl, p.ncp, Mp + 128) : c.kxz;
  c.w = plane_mat(ws, p.w_pl, p.ncp, Mp);
  c.linv = plane_mat(ws, p.linv_pl, Mp, Mp);
  // 16-bit operand planes of the backward follow the forward's format: fp16 pairs (22 bits) by default
  c.f16 = (s->flags & GDRF_FLAG_FWD_BF16) == 0;
  c.fmt = c.f16 ? FMT_F16 : FMT_BF16;
  c.st_b = c.f16 ? st_f16_nat(ws, p) : st_bf16(ws, p);                       // B of dW (G3)
  c.linv_b = c.f16 ? plane_mat(ws, p.linv16_pl, Mp, Mp) : c.linv;            // B of dKxz (G4)
  c.w_b = c.f16 ? plane_mat(ws, p.w16_pl, p.ncp, Mp) : c.w;                  // B of C5 (G5)
  c.cs = at<unsigned>(ws, p.cs);
  c.ps = at<unsigned>(ws, p.ps);
  c.cs_f = at<float>(ws, p.cs);
  c.tpm = plane_mat(ws, p.tp_pl, p.ncp, (long long)p.K * Mp);
  c.wgm = plane_mat(ws, p.wg_pl, p.ncp, (long long)p.K * Mp);
  c.dwf = at<float>(ws, p.dwf);
  c.acc = at<double>(ws, p.acc);
  // tiles of dS on / below the diagonal (i tile of 128 rows, j tile of 256 columns)
  c.g6 = G6::Params{};
  if (want_grad) {
    int nt = 0;
    // pair order: entries 2t, 2t+1 are the 128-row tiles (2 a2, b), (2 a2 + 1, b) of one 256 x 256 pair tile
    for (int b = 0; b < p.JT; ++b)
      for (int a2 = b; a2 < p.JT; ++a2)
        if (256 * a2 < p.M && 256 * b < p.M)
          for (int h = 0; h < 2; ++h) {
            if (nt >= G6::MAX_TILES) return fail(1, "too many dS tiles%s");
            c.g6.ta[nt] = (unsigned char)(2 * a2 + h);
            c.g6.tb[nt] = (unsigned char)b;
            ++nt;
          }
    c.g6.ntile = nt;
  }
  return 0;
}

// once per call, before the first chunk, when a gradient is wanted: what the backward operand scales need of u_loc
int prepare_u(StepCtx& c, bool want_grad) {
  if (c.f16) return pack_u(c.s, c.in, c.p, c.ws, c.st);
  if (want_grad) {
    CU(cudaMemsetAsync(c.ps + PS_UMAX, 0, sizeof(unsigned), c.st));
    k_absmax<<<32, 256, 0, c.st>>>(c.in->u_loc, (long long)c.p.K * c.p.M, c.ps + PS_UMAX);
    LAUNCH_CHECK();
  }
  return 0;
}

// Backward of one chunk from the per-observation weights (g_loc = d/d f_loc, g2 = 2 d/d f_var, gv0 = the share of
// d/d f_var that reaches var0 = max(variance - |W_n|^2, 0), and their chunk maxima in cs[]) to dS, du_loc, C5, dZ,
// d lengthscale, d variance (the Kxz part).  xs_chunk: the inputs the chunk's forward ran on.
int chunk_backward(StepCtx& c, const float* xs_chunk, int nc, int RT) {
  const gdrf_shape* s = c.s;
  const gdrf_inputs* in = c.in;
  const Plan& p = c.p;
  void* ws = c.ws;
  cudaStream_t st = c.st;
  const int sms = c.sms, K = p.K, M = p.M, Mp = p.Mp, fmt = c.fmt;
  const bool fuse_du = c.fuse_du;
  const Hyper& hp = c.hp;
  unsigned *cs = c.cs, *ps = c.ps;
  const float* cs_f = c.cs_f;
  float* dwf = c.dwf;
  double* acc = c.acc;
  PlaneMat &w = c.w, &wgm = c.wgm, &tpm = c.tpm, &st_b = c.st_b, &dwt = c.dwt, &linv_b = c.linv_b, &w_b = c.w_b;
  G6::Params& g6 = c.g6;
  {
    ProfScope ps(PK_G2B, st);   // slot reused: the row-weighted copies of W
    k_scale_w<<<dim3(p.MB, RT), 256, 0, st>>>(w, at<float>(ws, p.g2), K, p.MB, (int)p.ncp, wgm, fmt, in->variance, cs);
    LAUNCH_CHECK();
  }
  {
    G3::Params g{};
    g.tp = tpm; g.st = st_b; g.g2 = at<float>(ws, p.g2); g.dw = dwf; g.cs = cs; g.fmt = fmt;
    g.RT = RT; g.MB = p.MB; g.K = K; g.JT = p.JT; g.Mp = Mp; g.ncp = (int)p.ncp;
    g.varn = (s->flags & (GDRF_FLAG_REF_G3 | GDRF_FLAG_SINGLE_CTA | GDRF_FLAG_FULL_WIDTH)) ? 0 : 1;
    g.isplit = (s->flags & (GDRF_FLAG_REF_G3 | GDRF_FLAG_SINGLE_CTA)) ? 0 : 1;
    const int n_items = g.isplit ? ((RT + 1) / 2) * 2 * p.JT : RT;
    { ProfScope ps(PK_G3, st); ++g_launches; CU(launch_big<G3>(g, n_items, sms, (s->flags & GDRF_FLAG_REF_G3) != 0, (s->flags & GDRF_FLAG_SINGLE_CTA) != 0, st)); }
  }
  {
    const size_t smem = sizeof(float) * (size_t)K * (72 + 128);
    CU(cudaFuncSetAttribute(k_dw_finalize, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_dw_finalize<<<dim3(p.MB, RT), 256, smem, st>>>(w, dwf, Mp, at<float>(ws, p.g_loc), at<float>(ws, p.gv0),
                                                     in->u_loc, K, M, (int)p.ncp, dwt, fmt, in->variance, cs, ps,
                                                     fuse_du ? (K + 63) / 64 : 0);
    LAUNCH_CHECK();
  }
  if (!fuse_du)
    if (int e = launch_du(p, w, RT, ws, sms, st)) return e;
  {
    g6.wg = wgm; g6.tp = tpm; g6.ds = c.grad; g6.RT = RT; g6.MB = p.MB; g6.K = K; g6.M = M;
    g6.fmt = fmt; g6.inv_scale = cs_f + CS_SG_INV;
    // (topic, tile) items in units of CTA pairs: as many whole rounds of the sms / 2 pairs as fit run unsplit; the
    // items of the last, partial round are cut along the observations so that this round is full as well
    const int base = K * g6.ntile, pairs = base / 2, clusters = sms / 2, NBt = 2 * RT;
    int whole_pairs = (pairs / clusters) * clusters, sp;
    if (pairs - whole_pairs == 0) sp = 1;
    else sp = clusters / (pairs - whole_pairs);
    if (sp < 1) sp = 1;
    if (sp > 8 && whole_pairs > 0) sp = 8;
    if (sp > NBt) sp = NBt;
    const int per = (NBt + sp - 1) / sp;
    sp = (NBt + per - 1) / per;
    if (sp == 1) whole_pairs = pairs;
    g6.n_whole = 2 * whole_pairs; g6.tail_sp = sp; g6.tail_per = per;
    const int n_items6 = g6.n_whole + (base - g6.n_whole) * sp;
    { ProfScope ps(PK_G6, st); ++g_launches; CU(launch_big<G6>(g6, n_items6, sms, (s->flags & GDRF_FLAG_REF_G6) != 0, (s->flags & GDRF_FLAG_SINGLE_CTA) != 0, st)); }
  }
  {
    auto fill = [&](auto& g) {
      g.dwt = dwt; g.linv = linv_b; g.dkxz = dwf; g.RT = RT; g.MB = p.MB; g.Mp = Mp;
      g.fmt = fmt; g.inv_scale = cs_f + CS_SD_INV;
    };
    ProfScope ps(PK_G4, st);
    ++g_launches;
    if ((s->flags & (GDRF_FLAG_REF_G4 | GDRF_FLAG_SINGLE_CTA)) != 0) {
      G4T<128>::Params g{}; fill(g);
      CU(launch_gemm<G4T<128>>(g, RT, sms, (s->flags & GDRF_FLAG_REF_G4) != 0, st));
    } else {
      G4T<256>::Params g{}; fill(g);
      CU(launch_gemm2<G4T<256>>(g, RT, sms, st));
    }
  }
  {
    // one 128-column x rows_per_cta block per CTA, a serial row loop per thread: ~8 CTAs per SM hide its latency.  With
    // few inducing points (M = 256: two column blocks) 128 rows per CTA left two CTAs per SM and the kernel at 55 us per
    // chunk whatever M; fewer rows per CTA then
    int rows_per_cta = 128;
    while (rows_per_cta > 32 && (long long)p.MT * ((nc + rows_per_cta - 1) / rows_per_cta) < 6LL * sms) rows_per_cta >>= 1;
#define GDRF_KXZ_BACKWARD(DT, KID)                                                                  \
k_kxz_backward<DT, KID><<<dim3(p.MT, (nc + rows_per_cta - 1) / rows_per_cta), 128, 0, st>>>(      \
    dwf, Mp, xs_chunk, nc, in->z, M, hp, rows_per_cta, at<double>(ws, p.dz), acc)
    GDRF_DISPATCH_DK(p.D, hp.kid, GDRF_KXZ_BACKWARD);
#undef GDRF_KXZ_BACKWARD
    LAUNCH_CHECK();
  }
  {
    const int NBt = 2 * RT;
    auto fill = [&](auto& g, int splits) {
      g.dwt = dwt; g.w = w_b; g.c5 = at<double>(ws, p.c5); g.RT = RT; g.MB = p.MB; g.Mp = Mp;
      g.MTW = p.MT; g.MT = p.MT + (fuse_du ? 2 : 0);
      g.du = fuse_du ? at<double>(ws, p.du) : nullptr; g.inv_scale_l = cs_f + CS_SL_INV; g.K = K; g.M = M;
      g.fmt = fmt; g.inv_scale = cs_f + CS_SD_INV;
      if (splits > NBt) splits = NBt;
      if (splits < 1) splits = 1;
      const int per = (NBt + splits - 1) / splits;
      g.splits = (NBt + per - 1) / per;
      g.nb_per_split = per;
    };
    ProfScope ps(PK_G5, st);
    ++g_launches;
    if ((s->flags & (GDRF_FLAG_REF_G5 | GDRF_FLAG_SINGLE_CTA)) != 0) {
      const int base = p.MT * p.MT;
      G5T<128>::Params g{}; fill(g, base >= 2 * sms ? 1 : (2 * sms + base - 1) / base);
      CU(launch_gemm<G5T<128>>(g, base * g.splits, sms, (s->flags & GDRF_FLAG_REF_G5) != 0, st));
    } else {
      // 256 x 256 pair tiles: as many splits of the observation range as fill the sms / 2 CTA pairs once
      const int base = (p.MT + (fuse_du ? 2 : 0)) * (Mp / 256), pairs = base / 2, clusters = sms / 2;
      G5T<256>::Params g{}; fill(g, pairs >= clusters ? 1 : clusters / pairs);
      CU(launch_gemm2<G5T<256>>(g, base * g.splits, sms, st));
    }
  }
  return 0;
}

// Per-step tail of the gradient: Cholesky adjoint of C5 -> Kuu adjoint -> (Z, lengthscale, variance), the Dirichlet
// prior, and the assembly of the small gradients behind dS in the flat buffer.
int step_epilogue(StepCtx& c, bool want_grad) {
  const gdrf_shape* s = c.s;
  const gdrf_inputs* in = c.in;
  const Plan& p = c.p;
  void* ws = c.ws;
  cudaStream_t st = c.st;
  const int K = p.K, M = p.M, Mp = p.Mp;
  const Hyper& hp = c.hp;
  double* acc = c.acc;
  if (want_grad) {
    // Cholesky adjoint (Murray 2016; torch cholesky_backward):  G_L = -tril(L^-T C5),
    // Phi = tril(L^T G_L) with halved diagonal, G_K = L^-T Phi L^-1, symmetrised inside k_kuu_backward.
    double* L = at<double>(ws, p.L64);
    double* Linv = at<double>(ws, p.Linv64);
    double* tA = at<double>(ws, p.tmpA);
    double* tB = at<double>(ws, p.tmpB);
    const dim3 g2d(ceil_div(Mp, 256), Mp);
    // every factor is triangular: (L^-T X)[i, j] sums over k >= i, a lower X over k >= j, X L^-1 over k >= j
    dgemm<true, false, 1, true>(Linv, at<double>(ws, p.c5), tA, Mp, st);     // lower part of L^-T C5
    k_tril_op<<<g2d, 256, 0, st>>>(tA, Mp, 0);
    dgemm<true, false, 1, true>(L, tA, tB, Mp, st);                          // lower part of L^T G_L
    k_tril_op<<<g2d, 256, 0, st>>>(tB, Mp, 1);
    dgemm<true, false, 3, false>(Linv, tB, tA, Mp, st);                      // L^-T Phi, Phi lower
    dgemm<false, false, 2, false>(tA, Linv, tB, Mp, st);                     // (...) L^-1
    g_launches += 5;   // 4 fp64 products + 2 tril ops, one counted by LAUNCH_CHECK
    LAUNCH_CHECK();
    k_kuu_backward<<<M, 128, 0, st>>>(tB, Mp, in->z, M, hp, at<double>(ws, p.dz), acc);
    LAUNCH_CHECK();
  }
  const int include_prior = (s->flags & GDRF_FLAG_INCLUDE_PRIOR) ? 1 : 0;
  if (include_prior) {
    k_prior<<<K, 128, 0, st>>>(in->phi, in->beta, K, p.V, acc);
    LAUNCH_CHECK();
  }
  if (want_grad) {
    const int has_alpha = s->kernel_id == KERNEL_RQ ? 1 : 0;
    const long long small = (long long)K * M + (long long)K * p.V + (long long)M * p.D + 2 + s->ls_dim + has_alpha;
    k_assemble<<<(int)((small + 255) / 256), 256, 0, st>>>(K, M, p.V, p.D, s->ls_dim, has_alpha, include_prior, in->phi, in->beta,
                                                           acc, at<double>(ws, p.ck), at<double>(ws, p.du),
                                                           at<double>(ws, p.dphi), at<double>(ws, p.dz), c.grad);
    LAUNCH_CHECK();
  }
  return 0;
}

int gdrf_elbo_step(const gdrf_shape* s, const gdrf_inputs* in, const gdrf_outputs* out, void* ws, size_t ws_bytes,
                   gdrf_stream_t stream) {
  Plan p;
  if (int e = make_plan(s, p)) return e;
  if (int e = check_device()) return e;
  if (!in || !out || !ws || !out->terms) return fail(1, "null pointer argument%s");
  if (s->kernel_id == KERNEL_RQ && !in->scale_mixture) return fail(1, "the RationalQuadratic kernel needs in->scale_mixture%s");
  if (ws_bytes < p.total) return fail(1, "workspace too small%s (need %lld bytes)", "", (long long)p.total);
  const bool want_grad = (s->flags & GDRF_FLAG_WANT_GRAD) != 0;
  if (want_grad && !out->grad) return fail(1, "GDRF_FLAG_WANT_GRAD needs out->grad%s");
  if (s->n_offset < 0 || s->n_offset + s->n_local > s->n_eps) return fail(1, "eps window out of range%s");
  cudaStream_t st = (cudaStream_t)stream;
  const int sms = num_sms();
  const int K = p.K, M = p.M;

  const int P = s->n_particles > 1 ? s->n_particles : 1;
  const double inv_p = 1.0 / P;
  const bool cont = (s->flags & GDRF_FLAG_CONTINUE) != 0;   // accumulators carry over from the previous call
  const bool partial = (s->flags & GDRF_FLAG_PARTIAL) != 0; // more sub-shards follow: skip the per-step epilogue
  if (!cont) {
    CU(cudaMemsetAsync(at<char>(ws, p.acc), 0, (size_t)p.zero_bytes, st));
    if (want_grad) CU(cudaMemsetAsync(out->grad, 0, sizeof(float) * (size_t)K * M * M, st));
  }
  StepCtx c;
  if (int e = make_step_ctx(s, in, p, ws, out->grad, st, want_grad, c)) return e;
  const Hyper& hp = c.hp;
  unsigned* cs = c.cs;
  double* acc = c.acc;

  if (!cont) {
    k_phisum<<<K, 128, 0, st>>>(in->phi, K, p.V, at<float>(ws, p.phisum));
    LAUNCH_CHECK();
    if (likelihood_on_tensor_pipe(s)) {
      k_pack_phit<<<(int)(round_up_ll(p.V, 128) / 128), 256, 0, st>>>(in->phi, K, p.V, phit_mat(ws, p));
      LAUNCH_CHECK();
      k_lfact_table<<<LT_LFACT / 256, 256, 0, st>>>(at<float>(ws, p.lfact));
      LAUNCH_CHECK();
    }
    if (int e = prepare_u(c, want_grad)) return e;
  }

  for (long long n0 = 0; n0 < s->n_local; n0 += p.chunk_rows) {
    const int nc = (int)((s->n_local - n0 < p.chunk_rows) ? (s->n_local - n0) : p.chunk_rows);
    const int RT = (nc + 127) / 128;
    if (int e = chunk_forward(s, in, p, ws, n0, nc, RT, true, want_grad, sms, st)) return e;
    // the per-observation chain once per particle (draw of the guide): the marginal moments above do not depend on the
    // draw; the ELBO terms and the backward weights (g_loc, g2, gv0) accumulate their mean over the particles
    for (int pi = 0; pi < P; ++pi) {
      const float* eps_p = in->eps + (long long)pi * K * s->n_eps;
#define GDRF_OBS_PREPARE(KQ)                                                                                   \
  k_obs_prepare<KQ><<<(nc + 31) / 32, 256, 0, st>>>(nc, (int)p.ncp, K, s->n_offset + n0, s->n_eps,                \
                                                    at<double>(ws, p.floc), at<double>(ws, p.q), at<double>(ws, p.wsq), \
                                                    eps_p, hp, at<float>(ws, p.phisum), at<float>(ws, p.fvar),     \
                                                    at<float>(ws, p.theta), at<float>(ws, p.srow), acc, inv_p)
      GDRF_DISPATCH_KQ(K, GDRF_OBS_PREPARE);
#undef GDRF_OBS_PREPARE
      LAUNCH_CHECK();
      if (int e = dispatch_likelihood(s, p, nc, in, n0, ws, sms, st, inv_p)) return e;
#define GDRF_OBS_FINALIZE(KQ)                                                                                  \
  k_obs_finalize<KQ><<<RT * 4, 256, 0, st>>>(nc, (int)p.ncp, K, s->n_offset + n0, s->n_eps, at<float>(ws, p.theta), \
                                             at<float>(ws, p.srow), at<float>(ws, p.g1), at<float>(ws, p.arow),    \
                                             at<float>(ws, p.cnt), at<float>(ws, p.fvar), at<double>(ws, p.wsq),   \
                                             eps_p, hp, at<float>(ws, p.phisum), at<float>(ws, p.g_loc),           \
                                             at<float>(ws, p.g2), at<float>(ws, p.gv0), at<double>(ws, p.ck), acc, \
                                             RT * 128, cs, (float)inv_p, pi == 0, pi == P - 1)
      GDRF_DISPATCH_KQ(K, GDRF_OBS_FINALIZE);
#undef GDRF_OBS_FINALIZE
      LAUNCH_CHECK();
    }
    if (!want_grad) continue;
    if (int e = chunk_backward(c, in->xs + n0 * p.D, nc, RT)) return e;
  }

  if (partial) return 0;
  if (int e = step_epilogue(c, want_grad)) return e;
  float* tail = nullptr;
  if (want_grad && (s->flags & GDRF_FLAG_TERMS_IN_GRAD)) {
    int64_t ge = 0;
    gdrf_grad_elems(s, &ge);
    tail = out->grad + ge;
  }
  k_copy_terms<<<1, 32, 0, st>>>(acc, out->terms, tail);
  LAUNCH_CHECK();
  return 0;
}

int gdrf_moments_vjp(const gdrf_shape* s, const gdrf_inputs* in, const float* up_floc, const float* up_fvar, float* grad,
                     void* ws, size_t ws_bytes, gdrf_stream_t stream) {
  Plan p;
  if (int e = make_plan(s, p)) return e;
  if (int e = check_device()) return e;
  if (!in || !ws || !up_floc || !grad) return fail(1, "null pointer argument%s");
  if (!in->u_scale_tril) return fail(1, "gdrf_moments_vjp needs u_scale_tril (packed by gdrf_prologue)%s");
  if (s->kernel_id == KERNEL_RQ && !in->scale_mixture) return fail(1, "the RationalQuadratic kernel needs in->scale_mixture%s");
  if (ws_bytes < p.total) return fail(1, "workspace too small%s (need %lld bytes)", "", (long long)p.total);
  cudaStream_t st = (cudaStream_t)stream;
  const int K = p.K, M = p.M;
  CU(cudaMemsetAsync(at<char>(ws, p.acc), 0, (size_t)p.zero_bytes, st));
  CU(cudaMemsetAsync(grad, 0, sizeof(float) * (size_t)K * M * M, st));
  gdrf_shape sh = *s;                       // no prior, no terms
  sh.flags = (s->flags | GDRF_FLAG_WANT_GRAD) & ~(GDRF_FLAG_INCLUDE_PRIOR | GDRF_FLAG_TERMS_IN_GRAD | GDRF_FLAG_CONTINUE | GDRF_FLAG_PARTIAL);
  StepCtx c;
  if (int e = make_step_ctx(&sh, in, p, ws, grad, st, true, c)) return e;
  if (int e = prepare_u(c, true)) return e;
  for (long long n0 = 0; n0 < s->n_local; n0 += p.chunk_rows) {
    const int nc = (int)((s->n_local - n0 < p.chunk_rows) ? (s->n_local - n0) : p.chunk_rows);
    const int RT = (nc + 127) / 128;
    if (int e = chunk_forward(&sh, in, p, ws, n0, nc, RT, true, true, c.sms, st)) return e;
    k_vjp_weights<<<RT * 4, 256, 0, st>>>(nc, (int)p.ncp, K, n0, s->n_local, up_floc, up_fvar, at<double>(ws, p.wsq), c.hp,
                                          at<float>(ws, p.g_loc), at<float>(ws, p.g2), at<float>(ws, p.gv0), c.acc,
                                          RT * 128, c.cs);
    LAUNCH_CHECK();
    if (int e = chunk_backward(c, in->xs + n0 * p.D, nc, RT)) return e;
  }
  return step_epilogue(c, true);
}

int gdrf_elbo_backward(const float* grad, int64_t elems, const float* scale_dev, float scale_host, float* dst,
                       gdrf_stream_t stream) {
  if (!grad || !dst || elems < 0) return fail(1, "bad argument%s");
  if (elems == 0) return 0;
  long long blocks = (elems + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  k_scale<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(grad, elems, scale_dev, scale_host, dst);
  LAUNCH_CHECK();
  return 0;
}

int gdrf_marginal_mean(const gdrf_shape* s, const gdrf_inputs* in, float* out_floc, void* ws, size_t ws_bytes,
                       gdrf_stream_t stream) {
  return gdrf_marginal_moments(s, in, out_floc, nullptr, ws, ws_bytes, stream);
}

int gdrf_marginal_moments(const gdrf_shape* s, const gdrf_inputs* in, float* out_floc, float* out_fvar, void* ws,
                          size_t ws_bytes, gdrf_stream_t stream) {
  return marginal_moments_impl<float>(s, in, out_floc, out_fvar, ws, ws_bytes, stream);
}

int gdrf_marginal_moments_f64(const gdrf_shape* s, const gdrf_inputs* in, double* out_floc, double* out_fvar, void* ws,
                              size_t ws_bytes, gdrf_stream_t stream) {
  return marginal_moments_impl<double>(s, in, out_floc, out_fvar, ws, ws_bytes, stream);
}

int gdrf_perplexity_terms(const gdrf_shape* s, const gdrf_inputs* in, const float* floc, double* out,
                          gdrf_stream_t stream) {
  Plan p;
  if (int e = make_plan(s, p)) return e;
  if (int e = check_device()) return e;
  if (!in || !floc || !out) return fail(1, "null pointer argument%s");
  cudaStream_t st = (cudaStream_t)stream;
  CU(cudaMemsetAsync(out, 0, 2 * sizeof(double), st));
  if (s->n_local == 0) return 0;
  long long blocks = (s->n_local + 7) / 8;
  if (blocks > 148 * 8) blocks = 148 * 8;
  k_perplexity<<<(int)blocks, 256, sizeof(float) * 8 * p.K, st>>>((int)s->n_local, p.K, p.V, floc, in->ws, in->phi, out);
  LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
