/* znll.h -- C ABI of the B200-native unbinned-likelihood hot path (libznll.so).
 *
 * This is the drop-in boundary for zfit's UnbinnedNLL / ExtendedUnbinnedNLL value + gradient.
 * The reference (zfit 0.28.0, pure Python on TensorFlow) has no FFI; its own plug-in slots for this
 * path are the BaseLoss override hooks.  Each entry point below names the reference interface it
 * replaces (file:line relative to the reference tree).  The Python host layer (zfit_b200/) binds
 * these with ctypes and calls them from  _loss_func / _value_gradient.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / CUDA types in signatures (streams are void*).
 *   - every function returns 0 on success, non-zero on a programming / CUDA error
 *     (text via znll_last_error()).  Numerical trouble (NaN/inf) is NEVER an error: it is
 *     returned in the outputs so that LossEval's NaN strategy keeps working
 *     (src/zfit/minimizers/evaluation.py:199-227).
 *   - all arithmetic is fp64.
 *
 * Model trees
 *   A channel's composite PDF is described as a pre-order array of znll_node.  The "slots" of a
 *   tree are the scalar arguments its nodes take, concatenated in pre-order (a node's own slots
 *   first, then its children's):
 *       ZNLL_GAUSS  mu, sigma                 (src/zfit/models/dist_tfp.py:187-268)
 *       ZNLL_EXP    lambda                    (src/zfit/models/basic.py:34-243)
 *       ZNLL_CB     mu, sigma, alpha, n       (src/zfit/models/physics.py:31-45,83-142,233-336)
 *       ZNLL_CHEB   c_0 .. c_degree           (src/zfit/models/polynomials.py:334-482)
 *       ZNLL_LEGENDRE, ZNLL_CHEB2  c_0 .. c_degree   (polynomials.py:181-332, 487-600; "next" row f.3)
 *       ZNLL_GCB    mu, sigmal, alphal, nl, sigmar, alphar, nr   (GeneralizedCB, physics.py:47-80,186-232;
 *                   DoubleCB is the same node with the sigma slot given twice)
 *       ZNLL_GET    mu, sigma, alpha          (GaussExpTail, physics.py:608-623,636-686,727-805)
 *       ZNLL_GGET   mu, sigmal, alphal, sigmar, alphar   (GeneralizedGaussExpTail, physics.py:625-634,689-724,819-923)
 *       ZNLL_BERNSTEIN  c_0 .. c_degree       (Bernstein, polynomials.py:872-1028; every coefficient is a slot)
 *       ZNLL_TGAUSS mu, sigma, low, high      (TruncatedGauss, dist_tfp.py:357-417: zero outside [low, high])
 *       ZNLL_HERMITE, ZNLL_LAGUERRE  c_0 .. c_degree   (Hermite, polynomials.py:752-869; Laguerre, :600-750)
 *       ZNLL_SUM    frac_0 .. frac_{k-1}      (src/zfit/models/functor.py:74-313; ALL k fractions)
 *       ZNLL_PROD   (none)                    (src/zfit/models/functor.py:331-483)
 *   The library differentiates with respect to slots; the host applies d(slot)/d(theta)
 *   (shared parameters, frac = yield/sum(yields), 1 - sum(fracs), user ComposedParameters).
 */
#ifndef ZNLL_H
#define ZNLL_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ZNLL_ABI_VERSION 1
#define ZNLL_MAX_COLS 8

enum znll_kind {
  ZNLL_GAUSS = 1,
  ZNLL_EXP = 2,
  ZNLL_CB = 3,
  ZNLL_CHEB = 4,
  ZNLL_LEGENDRE = 5,
  ZNLL_CHEB2 = 6,
  ZNLL_GCB = 7,
  ZNLL_GET = 8,
  ZNLL_GGET = 9,
  ZNLL_BERNSTEIN = 10,
  ZNLL_TGAUSS = 11,
  ZNLL_HERMITE = 12,
  ZNLL_LAGUERRE = 13,
  ZNLL_SUM = 16,
  ZNLL_PROD = 17
};

/* One node of a model tree, pre-order.  Leaf limits are fixed at plan time (they come from the
 * PDF's Space / norm, src/zfit/core/space.py:1246-2239), only slots change per evaluation. */
typedef struct znll_node {
  int32_t kind;       /* enum znll_kind */
  int32_t n_children; /* functors: number of direct children; leaves: 0 */
  int32_t column;     /* leaves: index of the bound data column this leaf reads */
  int32_t degree;     /* polynomials: degree d (slots c_0..c_d); else 0 */
  double norm_lo;     /* leaves: normalisation integral limits (BasePDF.normalization, basepdf.py:198-251) */
  double norm_hi;
  double space_lo;    /* polynomials: rescale limits (rescale_minus_plus_one, polynomials.py:32-43) */
  double space_hi;
  double shift;       /* ZNLL_EXP: numerics data shift s (basic.py:128-154): centre of the norm, or 0
                         inside a ProductPDF */
} znll_node;

/* flags for znll_eval / znll_launch */
#define ZNLL_WANT_GRAD 1u /* also return the slot gradient (loss.value_gradient, core/loss.py:697-763) */
#define ZNLL_KAHAN 2u     /* sumtype == "kahan" (core/loss.py:138-141); the kernels always use compensated
                             fp64 accumulation, the flag is accepted for interface parity */

typedef struct znll_plan znll_plan;

int znll_abi_version(void);
const char* znll_last_error(void);

/* Number of CUDA devices visible (0 => the product path cannot run; there is no CPU fallback). */
int znll_device_count(void);

/* Create / destroy a plan of n_channels simultaneous channels on CUDA device `device`.
 * Replaces: BaseLoss.__init__ holding model/data lists (core/loss.py:192-241). */
int znll_plan_create(int device, int n_channels, znll_plan** out);
void znll_plan_destroy(znll_plan* plan);

/* Describe channel `ch`'s composite PDF.  Replaces the traced pdf() dispatch chain
 * BasePDF.pdf -> _norm_pdf -> _call_pdf -> _fallback_pdf (core/basepdf.py:398-461).
 * Selects the fused sm_100a kernel instantiated for this tree shape (ahead-of-time catalogue,
 * else runtime instantiation of the same templates through NVRTC). */
int znll_plan_set_model(znll_plan* plan, int ch, const znll_node* nodes, int n_nodes);

/* Bind channel data that already lives in device memory (borrowed; the caller keeps it alive --
 * Data owns the event tensors for the life of the loss, core/loss.py:226).  cols_dev[i] is a
 * contiguous, 16-byte aligned fp64 column of n events (structure-of-arrays); w_dev may be NULL.
 * Replaces: Data / LightDataset column access (core/data.py:876-982,1772-1937). */
int znll_plan_bind_device(znll_plan* plan, int ch, int n_cols, const double* const* cols_dev,
                          const double* w_dev, int64_t n_events);

/* Same, from HOST buffers: the library allocates device columns, copies once, and owns them. */
int znll_plan_bind_host(znll_plan* plan, int ch, int n_cols, const double* const* cols_host,
                        const double* w_host, int64_t n_events);

/* Bind-time reordering for warp coherence (csrc/znll_prep.cu).  Events are static for a fit and
 * the NLL is a sum over them, so their order is free: a STABLE partition of the n events into 32 equal-width buckets of
 * column key_col over [lo, hi] makes every warp of the step kernel uniform with respect to a piecewise pdf (CrystalBall
 * core / tail, models/physics.py:31-45, where the reference evaluates both branches for every event through tf.where).
 * All n_cols columns and the weights (nullable, both or neither) move together; src and dst are device pointers to n
 * doubles each and must not alias.  Asynchronous on `stream`; deterministic (same input, same order). */
int znll_partition_events(int device, int n_cols, const double* const* src_cols, double* const* dst_cols,
                          const double* src_w, double* dst_w, int key_col, double lo, double hi, int64_t n,
                          void* stream);

/* The one-time inclusive cut of the events to the observable space, on the device, with a STABLE compaction (kept events
 * stay in their original order).  Replaces check_cut_data_weights / the boolean mask of Data (core/data.py:1203-1244:
 * `lower <= x <= upper` on every observable, weights cut alongside) for columns that already live in HBM.  An event is
 * kept iff lower[c] <= src_cols[c][i] <= upper[c] for every column c (NaN fails) and, when the optional pair is given,
 * less_a[i] < less_b[i] -- the accept step of accept-reject sampling (u < pdf, core/sample.py:440-505) compacts with the
 * same pass.  dst columns (and dst_w) need room for n events; *n_kept receives the survivor count.  lower / upper are
 * HOST arrays of n_cols doubles.  Synchronises `stream` before returning (the count is needed on the host). */
int znll_cut_events(int device, int n_cols, const double* const* src_cols, double* const* dst_cols, const double* src_w,
                    double* dst_w, const double* lower, const double* upper, const double* less_a /* nullable */,
                    const double* less_b /* nullable */, int64_t n, int64_t* n_kept, void* stream);
const char* znll_prep_last_error(void);

/* Sum of weights (or N) of the bound shard, computed on the device at bind time with the same
 * deterministic reduction as the NLL.  Data.samplesize, core/data.py:268-275.  Multi-GPU: the
 * host all-reduces it once and writes the global figure back with znll_plan_set_sumw. */
int znll_plan_get_sumw(znll_plan* plan, int ch, double* sumw);
int znll_plan_set_sumw(znll_plan* plan, int ch, double sumw);

/* Multi-GPU, one process per GPU: how the library sums a small HOST buffer over the ranks (in place, identical result on
 * every rank; returns 0 on success).  Only the passes that are not reduced inside the kernel need it -- today the
 * information matrix that seeds znll_migrad_theta's metric (znll_eval itself exchanges its accumulators over NVLink inside
 * the kernel, znll_plan_set_peers).  The host layer installs an all-gather + fixed-order sum over its communicator
 * (zfit_b200/distributed.py::all_reduce_sum_); NULL removes the hook. */
typedef int (*znll_reduce_t)(void* user, double* buf, int n);
int znll_plan_set_reduce_hook(znll_plan* plan, znll_reduce_t fn, void* user);

int znll_plan_n_slots(const znll_plan* plan, int ch);   /* slots of channel ch        */
int znll_plan_n_acc(const znll_plan* plan, int ch);     /* accumulators of channel ch */
int znll_plan_total_slots(const znll_plan* plan);       /* sum over channels          */
int znll_plan_total_acc(const znll_plan* plan);
/* name of the kernel instantiation chosen for channel ch, e.g. "Sum<CBall<0>,Expo<0>>" */
const char* znll_plan_kernel_name(const znll_plan* plan, int ch);
/* 1 if the instantiation came from the ahead-of-time catalogue, 2 if NVRTC built it */
int znll_plan_kernel_origin(const znll_plan* plan, int ch);
/* number of kernel launches issued through this plan so far (bench.py's gpu_launches) */
int64_t znll_plan_launch_count(const znll_plan* plan);

/* One evaluation of every channel: value_k = -sum_i (w_i*log(pdf_k(x_i)+1e-307) - log_offset)
 * and, with ZNLL_WANT_GRAD, d value_k / d slot for every slot of every channel.
 *   slots       [total_slots]   host, channel-major, pre-order
 *   log_offset  per-event constant subtracted after the weight multiply (core/loss.py:134-137);
 *               pass 0.0 for full=True
 *   values      [n_channels]    host out
 *   grad_slots  [total_slots]   host out (may be NULL without ZNLL_WANT_GRAD)
 *   stream      cudaStream_t as void* (NULL = the default stream); all work of a call is enqueued there
 * Synchronous with respect to its outputs.  Replaces _unbinned_nll_tf + _nll_calc_unbinned_tf
 * (core/loss.py:73-142) and autodiff_value_gradient (z/math.py:236-261). */
int znll_eval(znll_plan* plan, const double* slots, uint32_t flags, double log_offset,
              double* values, double* grad_slots, void* stream);

/* Split form for multi-GPU: launch leaves the raw accumulators of all channels in a device
 * buffer (znll_plan_acc_device) so the caller can all-reduce (sum) them over ranks on the same
 * stream; collect copies them back, synchronises, and applies the host epilogue. */
int znll_launch(znll_plan* plan, const double* slots, uint32_t flags, double log_offset, void* stream);
double* znll_plan_acc_device(znll_plan* plan);
int znll_collect(znll_plan* plan, uint32_t flags, double* values, double* grad_slots, void* stream);

/* Weighted-covariance pass: value, slot gradient and, per channel, the (n_slots+1)^2 matrix
 *      [ sum_i w_i^2 g_i g_i^T    sum_i w_i^2 g_i ]        g_i = d log pdf(x_i) / d slot  (normalisation included)
 *      [ (sum_i w_i^2 g_i)^T      sum_i w_i^2     ]
 * row-major, channels concatenated, in ONE extra pass over the events.  Replaces the per-event Jacobian of
 * covariance_with_weights (src/zfit/minimizers/errors.py:333-464), which costs P forward-mode passes there. */
int znll_eval_cov(znll_plan* plan, const double* slots, double log_offset, double* values, double* grad_slots,
                  double* cov, void* stream);

/* Exact Hessian in ONE fused pass.  values[k], d value / d slot and, per channel, the (n_slots)^2
 * matrix d2 value_k / d slot_a d slot_b (row-major, channels concatenated), value_k = -sum_i (w_i log pdf_k(x_i) - offset)
 * with the normalisation integrals' dependence on the slots included.  Replaces loss.hessian's second differentiation of
 * the whole graph (src/zfit/core/loss.py:765-878, src/zfit/z/math.py:288-352; numdifftools over loss.value by default).
 * Available (znll_plan_has_hess) for a leaf, a ProductPDF of leaves, a SumPDF of leaves and of ProductPDFs of leaves, every
 * leaf kind, from the ahead-of-time catalogue or built through NVRTC, as long as 2 + (n+1)(n+2)/2 + the leaves'
 * second-order blocks fit the reduction's 256 accumulators (n = n_slots, about 20); for any other tree the function
 * returns 3 and the caller differences the analytic gradient (2P passes).  A zero fraction yields inf / NaN. */
int znll_eval_hess(znll_plan* plan, const double* slots, double log_offset, double* values, double* grad_slots,
                   double* hess, void* stream);
int znll_plan_has_hess(const znll_plan* plan, int ch);

/* ---- theta-space evaluation: the host chain rule of loss.value_gradient behind the C ABI -------------------------------
 * Replaces, for the built-in parameter kinds, what BaseLoss.value_gradient does around the per-event sum on every call:
 * Parameter.value with its clip (core/parameter.py:548-557), the fractions a SumPDF derives from yields and the
 * remaining fraction (models/basefunctor.py:189-205, 228-244), the reverse sweep through them (z/math.py:236-261), and
 * ExtendedUnbinnedNLL's Poisson term per channel (core/loss.py:1301-1317).  theta is the vector of ALL independent
 * parameters the loss depends on (floating or not), in an order the caller fixes once; every slot is one rule over refs. */
typedef struct znll_ref {
  int32_t kind;  /* 0: the constant `value`;  1: theta[index] (clipped to its limits, identity gradient) */
  int32_t index;
  double value;
} znll_ref;
enum znll_rule_kind {
  ZNLL_RULE_VALUE = 0,     /* slot = ref[first]                                   (Parameter, ConstantParameter)      */
  ZNLL_RULE_FRACTION = 1,  /* slot = ref[first] / sum(ref[first+1 .. first+n-1])  (frac_i = yield_i / sum(yields))    */
  ZNLL_RULE_REMAINDER = 2, /* slot = 1 - sum(ref[first .. first+n-1])             (the last fraction of a SumPDF)     */
  ZNLL_RULE_SUM = 3        /* slot = sum(ref[first .. first+n-1])                 (the yield of an automatic extended sum) */
};
typedef struct znll_rule {
  int32_t kind, first, n, pad;
} znll_rule;
#define ZNLL_FULL 4u /* flag of znll_eval_theta: full=True, i.e. log_offset is False (no offset, Stirling term kept) */
/* lower / upper: clip bounds of theta (+-inf = none; either may be NULL).  slot_rules: one per slot of the plan, channel
 * major.  yield_rules: one per channel for an extended loss (the channel's yield), NULL otherwise. */
int znll_plan_set_theta_map(znll_plan* plan, int n_theta, const double* lower, const double* upper, const znll_ref* refs,
                            int n_refs, const znll_rule* slot_rules, const znll_rule* yield_rules);
/* value = sum_k value_k (+ the extended terms), grad_theta[n_theta] = d value / d theta (NULL without ZNLL_WANT_GRAD).
 * One fused pass per channel; with peers set the result is world-reduced like znll_eval's. */
int znll_eval_theta(znll_plan* plan, const double* theta, uint32_t flags, double log_offset, double* value,
                    double* grad_theta, void* stream);
/* Fused multi-GPU reduction (no reference equivalent; the reference has no multi-device path).  peer_bufs[r] is rank
 * r's exchange buffer of at least znll_plan_exchange_bytes() bytes, zero-initialised, peer-mapped into this process
 * (CUDA IPC / VMM symmetric memory; the host obtains it from torch.distributed._symmetric_memory).  Once set, the last
 * block of each evaluation's (last) kernel pushes the raw accumulators to all peers over NVLink, waits for theirs and
 * sums in fixed rank order, so znll_eval returns world-reduced, bit-identical results on every rank without any
 * collective library call on the step path.  All ranks must call znll_eval the same number of times.  world <= 1 clears. */
int znll_plan_set_peers(znll_plan* plan, int rank, int world, void* const* peer_bufs, size_t buf_bytes);
size_t znll_plan_exchange_bytes(const znll_plan* plan, int world);

/* Per-event normalised pdf values of channel ch on its bound events: out_host[n_events].
 * Replaces BasePDF.pdf (core/basepdf.py:398-429) for plotting / diagnostics; not on the fit's hot path. */
int znll_pdf(znll_plan* plan, int ch, const double* slots, double* out_host, void* stream);
/* Same values into a caller-owned DEVICE buffer of n_events doubles, launched on `stream` without synchronising: the
 * accept-reject step of on-device toy generation (replaces the pdf evaluation inside accept_reject_sample,
 * core/sample.py:148-566, for SamplerData.resample, core/data.py:1563-1599). */
int znll_pdf_device(znll_plan* plan, int ch, const double* slots, double* out_dev, void* stream);

/* ---- native MIGRAD-protocol driver (host only; no GPU, no CUDA types) --------------------------------------------
 * Stands in for iminuit's C++ Minuit2 behind zfit.minimize.Minuit (src/zfit/minimizers/minimizer_minuit.py:175-319):
 * Minuit's parameter transformations, DFP metric update, line search, EDM stop rule with the Dcovar bookkeeping,
 * HESSE from gradient differences, the driver's own numerical gradient when has_gradient = 0, and optionally a metric
 * seeded / refreshed from the objective's information matrix.  The objective is seen only through the callbacks; they
 * return 0 on success, non-zero aborts the minimisation (znll_migrad then returns 2). */
#define ZNLL_FCN_VALUE 1
#define ZNLL_FCN_GRADIENT 2
/* what: bit mask of ZNLL_FCN_*; value is written when VALUE is set, grad[n] when GRADIENT is set (x is external). */
typedef int (*znll_fcn_t)(void* user, const double* x, int what, double* value, double* grad);
/* info[n*n] (row-major): approximate Hessian of the objective at x, e.g. the summed score outer product of a likelihood;
 * non-zero return = not available (the driver then uses Minuit's diagonal seed). */
typedef int (*znll_info_t)(void* user, const double* x, double* info);
typedef struct znll_migrad_options {
  double edm_goal;        /* stop when EDM (1 + 3 Dcovar) < edm_goal (zfit passes its tol, minimizer_minuit.py:313-315) */
  int32_t maxfcn;         /* limit on value + gradient calls; <= 0: 10000 */
  int32_t hesse;          /* 1: strategy-1 HESSE verification while Dcovar > 0.05; 0: never */
  int32_t has_gradient;   /* 0: finite differences of the value inside the driver (Minuit's default) */
  int32_t has_value_gradient; /* 1: a VALUE|GRADIENT call costs one pass (line-search probes then ask for both) */
} znll_migrad_options;
typedef struct znll_migrad_result {
  double fval, edm;
  int32_t niter, nfcn, ngrad, converged, valid;
  char message[96];
} znll_migrad_result;
/* lower / upper: per-parameter limits, NaN or +-inf = none (either pointer may be NULL); cov_out[n*n] = inverse Hessian
 * estimate in external coordinates, grad_out[n] = external gradient at the result (both may be NULL). */
int znll_migrad(znll_fcn_t fcn, znll_info_t info, void* user, int n, const double* x0, const double* steps,
                const double* lower, const double* upper, const znll_migrad_options* options, double* x_out,
                double* cov_out, double* grad_out, znll_migrad_result* result);

/* A whole MIGRAD minimisation without leaving native code: znll_migrad driven by znll_eval_theta over the n_fit entries
 * fit_index[] of theta (the others stay at theta0), metric seeded from the information matrix of znll_eval_cov when
 * use_information != 0.  Outputs as znll_migrad's; theta_out[n_theta] holds the final point.  Returns 2 when an
 * evaluation produced a non-finite value (the caller falls back to the evaluator path and its NaN strategy). */
int znll_migrad_theta(znll_plan* plan, int n_fit, const int32_t* fit_index, const double* theta0, const double* steps,
                      const double* lower, const double* upper, const znll_migrad_options* options, uint32_t flags,
                      double log_offset, int use_information, double* theta_out, double* cov_out, double* grad_out,
                      znll_migrad_result* result, void* stream);

/* Host-only helpers (no GPU needed): normalisation integral of one leaf and d I / d slot, exactly
 * as the per-call prologue computes them.  Replaces the registered analytic integrals
 * (dist_tfp.py:113-120, basic.py:211-229, physics.py:83-142, polynomials.py:443-478). */
int znll_leaf_integral(const znll_node* leaf, const double* slots, double* integral, double* dintegral);

/* Test hooks, host only (no GPU): the per-call prologue (slots -> the kernel's constant block, in the layout
 * znll_device.cuh documents) and epilogue (raw accumulators [sum, Kahan compensation, slot sums...] -> channel value and
 * d value / d slot) that znll_eval runs around the kernel.  tests/hostsim compiles the device header for the host and
 * uses these two to check the kernels' per-event arithmetic against the oracle on a CPU; the product never runs a
 * kernel on the host.  kernel_name receives the tree's instantiation name, e.g. "Sum<CBall<0>,Expo<0>>". */
int znll_test_prologue(const znll_node* nodes, int n_nodes, const double* slots, double* consts, int max_consts,
                       int* n_consts, int* n_slots, char* kernel_name, int name_len);
int znll_test_epilogue(const znll_node* nodes, int n_nodes, const double* slots, const double* acc, double sumw,
                       double* value, double* grad_slots /* nullable */);

/* Test hook, host only: the host half of znll_eval_hess -- raw accumulators of the Hessian pass ([value pair | upper
 * triangle of sum w [g;1][g;1]^T | the leaves' second-order blocks], n_hess of the latter) -> value, d value / d slot,
 * d2 value / d slot^2 ((n_slots)^2, row-major). */
int znll_test_hess_epilogue(const znll_node* nodes, int n_nodes, const double* slots, const double* acc, int n_acc, int n_hess,
                            double* value, double* grad_slots, double* hess);

/* Test hook, host only (NVRTC compiles without a GPU): obtain the compiled image of a tree outside the catalogue the way
 * znll_plan_set_model does -- from the on-disk cache ($ZNLL_CACHE_DIR, default $XDG_CACHE_HOME/zfit_b200 or
 * ~/.cache/zfit_b200; "" / "off" disables) when use_cache != 0 and a valid file exists, else through NVRTC (and store it).
 * model: instantiation name, e.g. "Sum<Gauss<0>,Gauss<0>,Cheb<0,2>>".  Loading the image needs the driver and is not done. */
int znll_test_rtc_image(const char* model, int n_slots, int hess_nh, int use_cache, size_t* cubin_bytes, int* n_names,
                        int* from_cache);

/* Test hook: out[i] = f(in[i]) on the device for the kernel's own fp64 elementary functions
 * (kind 0: exp, 1: log, 2: reciprocal), so their accuracy can be checked against libm. */
int znll_test_math(int kind, int64_t n, const double* in_host, double* out_host);

/* Device micro-benchmark used for the roofline denominator: register-resident DFMA loop.
 * Returns achieved FP64 TFLOP/s over `ms` milliseconds of work (approx.). */
int znll_fp64_peak(int device, double ms, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* ZNLL_H */
