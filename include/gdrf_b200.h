/* gdrf_b200 -- C ABI of the B200-native sparse multinomial GDRF ELBO + gradient.
 *
 * This is the whole drop-in boundary: plain pointers and sizes, no C++ or torch types.  It replaces, for
 * the reference san-soucie/gdrf, everything that `pyro.infer.SVI.step` evaluates through
 *   gdrf/models/sparse_gdrf.py:322-373   SparseMultinomialGDRF.model
 *   gdrf/models/sparse_gdrf.py:375-409   SparseMultinomialGDRF.guide
 *   gdrf/models/utils.py:27-40           jittercholesky
 *   gdrf/models/sparse_gdrf.py:161-186   log_topic_probs (evaluation path)
 * (and the pyro.contrib.gp kernels / `conditional` / torch.distributions code those call).
 *
 * Ownership: the caller owns every buffer (inputs, outputs, workspace); the library allocates no device memory and
 * keeps no state between calls besides a thread-local error string and the optional instrumentation at the end of
 * this header (an atomic launch counter; a mutex-guarded list of CUDA events while profiling is enabled).  Calls on
 * different (workspace, stream) pairs are independent and may be issued from different host threads.  All pointers are DEVICE pointers unless
 * stated otherwise.  All work is enqueued on `stream`; no call synchronises the device.  The only value
 * the host has to read back between calls is the 4-byte Cholesky status (to reproduce the try/except
 * escalation of jittercholesky).  Every entry point returns 0 on success and a non-zero code otherwise
 * (gdrf_last_error() then describes it).  The library needs an sm_100a device; there is no CPU path.
 */
#ifndef GDRF_B200_H
#define GDRF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* gdrf_stream_t; /* == cudaStream_t */

enum { GDRF_KERNEL_RBF = 0, GDRF_KERNEL_MATERN32 = 1, GDRF_KERNEL_MATERN52 = 2, GDRF_KERNEL_EXPONENTIAL = 3,
       GDRF_KERNEL_RATIONALQUADRATIC = 4 };

enum {
  GDRF_FLAG_WANT_GRAD = 1,        /* gdrf_elbo_step also produces the flat gradient                         */
  GDRF_FLAG_INCLUDE_PRIOR = 2,    /* add the Dirichlet log-density of phi and its gradient (one rank only)   */
  GDRF_FLAG_CHOL_FP32_STATUS = 4, /* decide not-PD with an fp32 factorisation of the reference's fp32 Kuu    */
  GDRF_FLAG_STATUS_ONLY = 8,      /* gdrf_prologue: only decide whether this jitter level fails (the fp32 factorisation of
                                     the reference's fp32 Kuu); a caller escalating through failing levels runs the full
                                     prologue once, at the level that succeeds                                  */
  GDRF_FLAG_FWD_BF16 = 16,        /* every 16-bit operand plane is bf16 (any fp32 range): the forward row-norm
                                     contraction runs on 3 bf16 planes / 6 products (24-bit operands) and the
                                     backward contractions on bf16 pairs (16-bit), instead of fp16 pairs / 3 products
                                     (22-bit operands, |x| < 6e4) everywhere.  Pass it to gdrf_prologue AND
                                     gdrf_elbo_step when gdrf_prologue reported status -1                    */
  GDRF_FLAG_CONTINUE = 64,        /* gdrf_elbo_step: keep the accumulators of the previous call on this workspace
                                     (second and later sub-shards of one step streamed from host memory)       */
  GDRF_FLAG_PARTIAL = 128,        /* gdrf_elbo_step: more sub-shards follow; skip the per-step epilogue (Cholesky
                                     adjoint, prior, gradient assembly, terms)                                 */
  GDRF_FLAG_TERMS_IN_GRAD = 1 << 17, /* out->grad has 8 more floats behind gdrf_grad_elems(): the four terms again, as
                                     (hi, lo) fp32 pairs, so that one all-reduce of the flat buffer carries the loss */
  GDRF_FLAG_SINGLE_CTA = 32,      /* run the four large contractions on single CTAs (cta_group::1) instead of
                                     CTA pairs (cta_group::2); same results, used for A/B measurement         */
  GDRF_FLAG_FULL_WIDTH = 1 << 14, /* issue full 256-column MMAs in the diagonal blocks of the triangular operands too
                                     (default: narrower MMAs there, 10-15 % fewer MMA columns); same results up to
                                     fp32 summation order, used for A/B measurement                           */
  /* Accumulation order of the two forward contractions (whitening W = Kxz L^-T, row norms T = W S_k).  The tensor
     pipe's fp32 accumulation costs about one ulp of the running sum per MMA, and the model uses the marginal variance as
     a *scale* (sparse_gdrf.py:403-405), which turns 2e-7 of relative noise there into 2e-4 of gradient error.
     Default: inside every 64-deep k-block the small correction products of the split are issued before the hi*hi
     products (csrc/gemm_tc.cuh, CorrFirst) -- measured: noise of the marginal variance 2.1-2.4e-7 -> 1.3-1.8e-7, small
     problems 3-7x closer to fp64, no cost.                                                                         */
  GDRF_FLAG_INTERLEAVED_MMAS = 1 << 15,  /* (hi,hi), (hi,lo), (lo,hi) per 16-deep k-step, as in the backward
                                            contractions (A/B measurement)                                          */
  GDRF_FLAG_SEGMENTED_FWD = 1 << 16,     /* segmented accumulation (csrc/gemm_tc.cuh, SegK): a fresh TMEM accumulator
                                            per k-block, summed in fp32 registers: noise 1.1-1.4e-7, but bound by the
                                            TMEM read port -- forward contraction +45 %, step +9 %                   */
  GDRF_FLAG_LIKELIHOOD_FMA = 1 << 18, /* mixture + multinomial likelihood on CUDA cores (k_likelihood) instead of the three
                                     chained tcgen05 contractions of k_likelihood_tc (default for k <= 64); A/B + checker */
  /* test hooks: run contraction Gi (i = 1..6) through the plain-FMA checker kernel instead of tcgen05 */
  GDRF_FLAG_REF_G1 = 1 << 8, GDRF_FLAG_REF_G2 = 1 << 9, GDRF_FLAG_REF_G3 = 1 << 10,
  GDRF_FLAG_REF_G4 = 1 << 11, GDRF_FLAG_REF_G5 = 1 << 12, GDRF_FLAG_REF_G6 = 1 << 13,
  GDRF_FLAG_REF_ALL = 0x3f << 8
};

typedef struct gdrf_shape {
  int64_t n_local;    /* observations in xs / ws / eps handed to this call (this rank's shard)               */
  int64_t n_offset;   /* column of eps holding this shard's first observation                                */
  int64_t n_eps;      /* row stride of eps (= total observations when eps is the global [K, N] tensor)        */
  int32_t d;          /* input dimensions (<= 8)                                                              */
  int32_t m;          /* inducing points (<= 4096)                                                            */
  int32_t k;          /* topics (<= 128)                                                                      */
  int32_t v;          /* observation categories                                                               */
  int32_t kernel_id;  /* GDRF_KERNEL_*   (train_script.py:93-99 KERNEL_DICT, all five)                           */
  int32_t ls_dim;     /* 1 (isotropic lengthscale) or d                                                       */
  int32_t chunk_rows; /* observations streamed per pass, multiple of 256; 0 = library default                 */
  int32_t flags;      /* GDRF_FLAG_*                                                                          */
  int32_t n_particles;/* draws of the guide per observation (Trace_ELBO(num_particles=P, vectorize_particles=True),
                         train_script.py:330-335); 0 or 1 = one.  eps is then [n_particles, k, n_eps], and the terms /
                         gradient are the MEAN over the particles.  The contractions run once: only the
                         per-observation chain is repeated per particle.                                       */
} gdrf_shape;

/* Constrained parameter values and data, row-major, fp32 unless noted. */
typedef struct gdrf_inputs {
  const float* xs;           /* [n_local, d]   observation locations, already scaled to the unit cube         */
  const int32_t* ws;         /* [n_local, v]   category counts (train_script.py:268: int32)                   */
  const float* eps;          /* [k, n_eps]     fixed standard-normal draws of the guide's mu site
                                               ([n_particles, k, n_eps] with gdrf_shape::n_particles > 1)             */
  const float* z;            /* [m, d]         inducing points                                                */
  const float* variance;     /* [1]            kernel variance                                                */
  const float* lengthscale;  /* [ls_dim]       kernel lengthscale                                             */
  const float* u_loc;        /* [k, m]                                                                        */
  const float* u_scale_tril; /* [k, m, m]      lower triangular (entries above the diagonal are ignored)      */
  const float* noise;        /* [1]                                                                           */
  const float* phi;          /* [k, v]         word-topic matrix, rows on the simplex                         */
  const float* beta;         /* [k, v]         Dirichlet concentration                                        */
  const float* scale_mixture;/* [1]            RationalQuadratic's third hyper-parameter; NULL for the other kernels  */
} gdrf_inputs;

/* terms[0..3] = { log p(mu), log q(mu), log p(w | mu, phi), log p(phi) } summed over this shard (fp64);
 *   ELBO = terms[0] + terms[3] + terms[2] - terms[1];  the reference's loss is -ELBO / N (train_script.py:365).
 * grad  = d ELBO / d (constrained parameter), fp32, laid out
 *   [ u_scale_tril k*m*m | u_loc k*m | phi k*v | z m*d | variance 1 | lengthscale ls_dim | noise 1 | scale_mixture 1 ]
 *   (gdrf_grad_elems() floats; the last entry exists for GDRF_KERNEL_RATIONALQUADRATIC only); u_scale_tril's
 *   entries above the diagonal are 0.                                                                       */
typedef struct gdrf_outputs {
  double* terms; /* [4]                      */
  float* grad;   /* [gdrf_grad_elems(shape)] or NULL when GDRF_FLAG_WANT_GRAD is clear                        */
} gdrf_outputs;

/* Sizes.  HOST out-pointers. */
int gdrf_workspace_bytes(const gdrf_shape* shape, size_t* out_bytes);
int gdrf_grad_elems(const gdrf_shape* shape, int64_t* out_elems);

/* Kuu = k(Z,Z) + (sum_{i<=njitter} jitter*10^i) I, its Cholesky factor and inverse, packed for the tensor
 * pipe (also packs u_scale_tril when in->u_scale_tril is non-NULL); *dev_status (DEVICE int) = 0 when the
 * factorisation succeeded, 1 + the failing column when it did not, -1 when it succeeded but u_scale_tril, L^-1, the
 * kernel variance or the bound sqrt(variance m) max|u_scale_tril| on T = W S_k may leave the fp16 range (then call
 * gdrf_prologue again, and gdrf_elbo_step, with GDRF_FLAG_FWD_BF16; with that flag the status is never -1).
 * The caller loops njitter = 0, 1, ... < maxjitter exactly like jittercholesky (utils.py:31-39); after a first failure
 * it may ask gdrf_jitter_probe for the following levels (several at once; or one at a time with GDRF_FLAG_STATUS_ONLY)
 * and run the full prologue at the first that passes.                                                           */
int gdrf_prologue(const gdrf_shape* shape, const gdrf_inputs* in, double jitter, int njitter, void* workspace,
                  size_t workspace_bytes, gdrf_stream_t stream, int* dev_status);

/* jittercholesky's question "does level njitter fail?" (utils.py:31-37: the try/except around the fp32
 * torch.linalg.cholesky) answered for `count` <= GDRF_PROBE_MAX consecutive levels njitter_first, njitter_first + 1, ...
 * in ONE launch chain: dev_status[i] (DEVICE int[count]) = 0 when the reference's fp32 factorisation of
 * Kuu + (sum_{t <= njitter_first + i} jitter 10^t) I succeeds, 1 + the failing column otherwise.  The levels are
 * independent and the chain is latency-bound, so the batch costs little more than one level (eight levels: 0.71 ms at
 * m = 625, 1.33 ms at m = 1024).  Writes only its own scratch in the workspace: a completed gdrf_prologue on the same
 * workspace stays valid.                                                                                          */
#define GDRF_PROBE_MAX 8
int gdrf_jitter_probe(const gdrf_shape* shape, const gdrf_inputs* in, double jitter, int njitter_first, int count,
                      void* workspace, size_t workspace_bytes, gdrf_stream_t stream, int* dev_status);

/* One evaluation of the ELBO terms (and, with GDRF_FLAG_WANT_GRAD, of the full gradient) over this shard,
 * streamed in chunks of chunk_rows observations.  Requires a successful gdrf_prologue on the same
 * workspace with the same z / variance / lengthscale.  A shard may be fed in several calls (sub-shards with
 * the same chunk_rows, the same out->grad; GDRF_FLAG_PARTIAL on all but the last, GDRF_FLAG_CONTINUE on all but
 * the first) so that host-to-device copies of the next sub-shard overlap the current one.                    */
int gdrf_elbo_step(const gdrf_shape* shape, const gdrf_inputs* in, const gdrf_outputs* out, void* workspace,
                   size_t workspace_bytes, gdrf_stream_t stream);

/* Backward of the scalar op: dst[i] = scale_dev[0] * scale_host * grad[i] (autograd chain with the
 * upstream gradient kept on the device).                                                                    */
int gdrf_elbo_backward(const float* grad, int64_t elems, const float* scale_dev, float scale_host, float* dst,
                       gdrf_stream_t stream);

/* Evaluation path: marginal mean f_loc[k, n] = (W u_loc^T)^T (log_topic_probs, sparse_gdrf.py:161-186).
 * out_floc is [k, n_local] fp32.  Requires gdrf_prologue.                                                   */
int gdrf_marginal_mean(const gdrf_shape* shape, const gdrf_inputs* in, float* out_floc, void* workspace,
                       size_t workspace_bytes, gdrf_stream_t stream);

/* Mean and variance of the sparse-GP marginal at new inputs: SparseGDRF.forward(Xnew, full_cov=False)
 * (sparse_gdrf.py:277-319) = pyro conditional(...) -> (f_loc, f_var), both [k, n_local] fp32; out_fvar may be
 * NULL.  Requires gdrf_prologue (which packs u_scale_tril).                                                 */
int gdrf_marginal_moments(const gdrf_shape* shape, const gdrf_inputs* in, float* out_floc, float* out_fvar,
                          void* workspace, size_t workspace_bytes, gdrf_stream_t stream);

/* Vector-Jacobian product of gdrf_marginal_moments: given the upstream gradients up_floc, up_fvar ([k, n_local] fp32 each;
 * up_fvar may be NULL = zero) of a scalar with respect to (f_loc, f_var), writes that scalar's gradient with respect to
 * the constrained u_scale_tril, u_loc, Z, variance, lengthscale (and scale_mixture) into `grad`, laid out like
 * gdrf_outputs.grad (the phi and noise entries are zero).  This is what torch autograd computes when the reference
 * differentiates through SparseGDRF.forward / gp.util.conditional (sparse_gdrf.py:277-319, :334-344): the forward
 * contractions of a chunk are re-run (T is kept), the upstream gradients take the place of the ELBO's per-observation
 * weights, and the backward contractions, the Cholesky adjoint and the kernel adjoints are the ones gdrf_elbo_step
 * runs.  Requires gdrf_prologue (with u_scale_tril) on the same workspace.                                       */
int gdrf_moments_vjp(const gdrf_shape* shape, const gdrf_inputs* in, const float* up_floc, const float* up_fvar,
                     float* grad, void* workspace, size_t workspace_bytes, gdrf_stream_t stream);

/* The same moments as the library holds them internally (fp64 per-observation chain): [k, n_local] fp64.  The
 * reference's outputs are fp32; this variant exists so that a caller (and the parity tests) can see the marginal
 * variance -- which the model uses as a *scale*, sparse_gdrf.py:403-405 -- below fp32 resolution.               */
int gdrf_marginal_moments_f64(const gdrf_shape* shape, const gdrf_inputs* in, double* out_floc, double* out_fvar,
                              void* workspace, size_t workspace_bytes, gdrf_stream_t stream);

/* perplexity pieces (abstract_gdrf.py:137-139): out[0] = sum w log(word_probs), out[1] = sum w  (fp64),
 * from f_loc [k, n_local]; the N x V word-probability matrix is never materialised.                        */
int gdrf_perplexity_terms(const gdrf_shape* shape, const gdrf_inputs* in, const float* floc, double* out,
                          gdrf_stream_t stream);

/* Fused SVI tail (the step either side of the path: train_script.py:325-327 optimiser on PyroParam-constrained
 * parameters).  theta_u / theta_c are the flat unconstrained / constrained parameter buffers in the layout of
 * gdrf_outputs::grad.  gdrf_constrain maps theta_u -> theta_c (lower_cholesky, identity, row softmax, sigmoid [or
 * identity when learn_z == 0: fixed inducing points], exp).  gdrf_adam_step chains `grad` (d ELBO / d constrained,
 * as written by gdrf_elbo_step) to the unconstrained parameters with loss = grad_scale * ELBO (grad_scale = -1/N)
 * and applies one Adam update (decoupled weight decay when weight_decay > 0, i.e. AdamW); m, v: moment buffers of
 * gdrf_grad_elems floats, zero before step 1; row_scratch: k floats; step counts from 1.                     */
int gdrf_constrain(const gdrf_shape* shape, const float* theta_u, float* theta_c, int learn_z, gdrf_stream_t stream);
int gdrf_adam_step(const gdrf_shape* shape, float* theta_u, const float* theta_c, const float* grad, float* m,
                   float* v, float* row_scratch, float lr, float beta1, float beta2, float eps, float weight_decay,
                   int step, float grad_scale, int learn_z, gdrf_stream_t stream);

/* pyro.optim.ClippedAdam on the same flat buffers (scripts/mvco.py:135; OPTIMIZER_DICT "clippedadam",
 * train_script.py:75): the chained gradient is clamped to [-clip_norm, clip_norm] element-wise, weight_decay is L2
 * (added to the clamped gradient), denom = sqrt(v) + eps and step size = lr sqrt(1 - beta2^t) / (1 - beta1^t).  ClippedAdam's
 * learning-rate decay (lr *= lrd before every step) is applied by the caller: pass lr = lr0 * lrd^step.                */
int gdrf_clipped_adam_step(const gdrf_shape* shape, float* theta_u, const float* theta_c, const float* grad, float* m,
                           float* v, float* row_scratch, float lr, float beta1, float beta2, float eps,
                           float weight_decay, float clip_norm, int step, float grad_scale, int learn_z,
                           gdrf_stream_t stream);

/* Streaming / mini-batch inference (train_script.py:442-460: `xs[selection, ...]`, `ws[selection, ...]` with a multiset
 * `selection` drawn on the host every sub-epoch): xs_out[i] = xs[index[i]], ws_out[i] = ws[index[i]] for i < n_sel, with
 * the data set [n_rows, d] / [n_rows, v] resident on the device.  index: DEVICE int64, negative values count from the end
 * like numpy.  An out-of-range index zero-fills its row and, when dev_status (DEVICE int, caller-zeroed, may be NULL) is
 * given, records 1 + the first offending position there (the reference raises IndexError).                           */
int gdrf_gather_rows(const float* xs, const int32_t* ws, const int64_t* index, int64_t n_sel, int64_t n_rows, int32_t d,
                     int32_t v, float* xs_out, int32_t* ws_out, int* dev_status, gdrf_stream_t stream);

/* Instrumentation for bench.py -- the library's only process-global state: kernels launched by this process so far
 * (atomic counter); CUDA-event timing of the six contractions (ms[7] / launches[7] in the order G1, G2, k_scale_w,
 * G3, G4, G5, G6, summed over EVERY launch since the previous read -- the record list grows with the run; the read
 * synchronises the device and recycles the events).  Off by default.                                        */
long long gdrf_launch_count(void);
int gdrf_profile_enable(int on);
int gdrf_profile_read(double* ms, long long* launches);

const char* gdrf_last_error(void);
const char* gdrf_build_info(void);

#ifdef __cplusplus
}
#endif
#endif /* GDRF_B200_H */
