/*
 * anml_b200 -- C-ABI of the B200-native likelihood-evaluation path of anml.
 *
 * The reference (ihmeuw-msca/anml) is pure Python and has no FFI; its boundary
 * for this path is the Python class API.  Each entry point below names the
 * reference interface it replaces (file:line under /root/reference).  The
 * Python host layer in anml_b200/ binds these with ctypes
 * and raises the reference's exception types from the status codes.
 *
 * Conventions
 *   - every function returns an int status (ANML_OK == 0); on failure
 *     anml_last_error() holds a message for the calling thread;
 *   - plain pointers and sizes only; "host" pointers are ordinary process
 *     memory, "device" pointers are CUDA device memory on the handle's GPU;
 *   - all floating point is IEEE double; matrices are row-major;
 *   - handles are not thread-safe; one handle per rank / GPU;
 *   - calls are synchronous on return unless the name ends in _async.
 */
#ifndef ANML_B200_H
#define ANML_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ANML_B200_VERSION 200 /* 0.2.0 */

/* status codes */
enum {
    ANML_OK = 0,
    ANML_EINVAL = 1,       /* bad argument                      -> ValueError / TypeError */
    ANML_ECUDA = 2,        /* CUDA runtime / driver failure     -> RuntimeError           */
    ANML_ENOMEM = 3,       /* device or host allocation failed  -> MemoryError            */
    ANML_ESTATE = 4,       /* handle not fully configured       -> ValueError             */
    ANML_EUNSUPPORTED = 5, /* shape outside what is implemented -> NotImplementedError    */
    ANML_EDOMAIN = 6       /* Log/Logit argument out of domain  -> ValueError
                              (smoothmapping.py:98-99, :148-151)                          */
};

/* SmoothMapping subclasses, src/anml/parameter/smoothmapping.py:51-156 */
enum {
    ANML_LINK_IDENTITY = 0, /* :51-65   */
    ANML_LINK_EXP = 1,      /* :68-78   */
    ANML_LINK_LOG = 2,      /* :81-104  */
    ANML_LINK_EXPIT = 3,    /* :107-126 */
    ANML_LINK_LOGIT = 4     /* :129-156 */
};

/* per-row negative log-likelihood of the model layer (SURVEY.md 3.3 / row a19;
 * the consumer of Parameter.get_params that the reference leaves to callers) */
enum {
    ANML_FAMILY_GAUSSIAN = 0,        /* 1/2 ((obs-mu)/se)^2, data/example.py:7-8 */
    ANML_FAMILY_BINOMIAL = 1,        /* -(obs log p + (1-obs) log(1-p))          */
    ANML_FAMILY_POISSON = 2,         /* lam - obs log lam                        */
    ANML_FAMILY_GAUSSIAN_MEANVAR = 3 /* 1/2 log v + 1/2 (obs-mu)^2/v, 2 params   */
};

enum { ANML_WANT_OBJ = 1, ANML_WANT_GRAD = 2, ANML_WANT_HESS = 4 };

typedef struct anml_design anml_design; /* a Parameter's cached design matrix in HBM */
typedef struct anml_model anml_model;   /* obs + K parameters + priors -> obj/grad/Hessian */

const char* anml_last_error(void);
int anml_version(void);
int anml_device_count(int* count);
int anml_device_info(int device, int* sm_count, int64_t* total_mem_bytes, int* cc_major, int* cc_minor);

/* ---- raw device buffers (offset / obs columns owned by the Python layer) ---- */
int anml_dev_alloc(int device, int64_t bytes, void** out_dev);
int anml_dev_free(int device, void* dev);
int anml_dev_upload(int device, void* dst_dev, const void* src_host, int64_t bytes);
int anml_dev_download(int device, void* dst_host, const void* src_dev, int64_t bytes);
/* free the pinned / device staging lanes the uploads of `device` keep between calls (device < 0: every
 * device); they are rebuilt on demand */
int anml_upload_release(int device);

/* ---- design matrix: replaces Parameter.design_mat = np.hstack([...])
 *      (src/anml/parameter/main.py:160-162) with a row-major (n, p) FP64
 *      matrix resident in HBM, leading dimension rounded up to even -------- */
int anml_design_create(int device, int64_t n_rows, int n_cols, anml_design** out);
/* Same shape contract, but nothing is allocated yet: spline blocks set with device-resident abscissae
 * (anml_design_set_spline, x_on_device = 1, interval_out = NULL) are only RECORDED -- the abscissae are
 * borrowed and must outlive the design -- and the dense matrix is built the first time something needs
 * it (set_columns, download, params, shape(dev), validate, or an evaluation that streams it).  A
 * one-spline model on the banded evaluator (anml_model_enable_banded_spline) never does: BASELINE
 * config #3 then holds the 0.4 GB of abscissae instead of an 8.8 GB design it does not read. */
int anml_design_create_deferred(int device, int64_t n_rows, int n_cols, anml_design** out);
int anml_design_is_materialized(const anml_design* d, int* out);
/* adopt a caller-owned device matrix without copying (ld even, 16-byte aligned) */
int anml_design_wrap_device(int device, int64_t n_rows, int n_cols, double* dev, int64_t ld, anml_design** out);
int anml_design_destroy(anml_design* d);
int anml_design_shape(const anml_design* d, int64_t* n_rows, int* n_cols, int64_t* ld, double** dev);
/* Variable.get_design_mat (src/anml/variable/main.py:99-114): write `ncols`
 * columns starting at col0 from src[(row)*src_ld + c] */
int anml_design_set_columns(anml_design* d, int col0, int ncols, const double* src, int64_t src_ld, int src_on_device);
/* SplineVariable.get_design_mat -> XSpline.get_design_mat(x, order)
 * (src/anml/variable/spline.py:109-113, getter/prior.py:191-194): knot-interval
 * lookup + clamped B-spline basis built on device into columns
 * [col0, col0 + n_knots - 1 + degree).  interval_out (host, n int32, nullable)
 * receives the knot-interval index of every row. */
int anml_design_set_spline(anml_design* d, int col0, const double* x, int x_on_device, const double* knots,
                           int n_knots, int degree, int ldegree, int rdegree, int order, int32_t* interval_out);
/* NoNans / Positive validators (src/anml/data/validator.py:21-34) on the uploaded
 * columns: flags_out[c] (host, ncols ints) gets bit 0 if column col0+c holds a NaN,
 * bit 1 if it holds a value <= 0. */
int anml_design_validate_columns(const anml_design* d, int col0, int ncols, int* flags_out);
int anml_design_download(const anml_design* d, double* host_out, int64_t out_ld);
/* Parameter.get_params(x, order) (src/anml/parameter/main.py:265-280):
 * order 0 -> out (n), order 1 -> out (n,p), order 2 -> out (n,p,p), host.
 * offset_dev: device (n) or NULL (:270-271). */
int anml_design_params(const anml_design* d, const double* beta_host, const double* offset_dev, int link, int order,
                       double* out_host);

/* standalone spline basis (XSpline.get_design_mat for SplinePriorGetter, any n):
 * x host (n) -> out host (n, n_knots-1+degree) row-major, computed on `device` */
int anml_spline_design(int device, const double* x_host, int64_t n, const double* knots, int n_knots, int degree,
                       int ldegree, int rdegree, int order, double* out_host, int32_t* interval_out);

/* ---- knot placement from data (SplineGetter.get_spline, getter/spline.py:92-99) ----
 * 'rel_domain' knots need data.min()/max(), 'rel_freq' knots np.quantile(data, k): one pass /
 * one sort over the n abscissae, done on the device that already holds them.
 * minmax: NaN anywhere -> *has_nan = 1 (NumPy then returns nan for both).
 * order_stats: out[i] = sorted(x)[ranks[i]], 0 <= ranks[i] < n (the order statistics
 * np.quantile interpolates between; the interpolation itself stays on the host). */
int anml_vector_minmax(int device, const double* x, int64_t n, int x_on_device, double* out_min, double* out_max, int* has_nan);
int anml_vector_order_stats(int device, const double* x, int64_t n, int x_on_device, const int64_t* ranks, int n_ranks,
                            double* out);

/* ---- model: objective / gradient / Hessian ---------------- */
int anml_model_create(int device, int64_t n_rows, int family, int n_params, anml_model** out);
int anml_model_destroy(anml_model* m);
/* parameter k = f_k(offset_k + X_k beta_k); the design is borrowed, not copied */
int anml_model_set_param(anml_model* m, int k, const anml_design* design, int link, const double* offset_dev);
int anml_model_set_obs(anml_model* m, const double* obs, int on_device);
int anml_model_set_obs_se(anml_model* m, const double* obs_se /* NULL = 1.0 */, int on_device);
/* One-parameter model whose design is ONE B-spline basis (BASELINE config #3): evaluate the basis on the
 * fly from the abscissae instead of streaming the dense (n, nb) design (csrc/spline_eval.cuh).  Call after
 * anml_model_set_param with the dense design of the same spline (it stays the fallback and serves
 * get_params); x holds the n abscissae (host or device; not retained).  ANML_EUNSUPPORTED when an abscissa
 * lies outside the knot span or degree > 3: the model then keeps using the dense design. */
int anml_model_enable_banded_spline(anml_model* m, const double* x, int x_on_device, const double* knots, int n_knots, int degree);
/* optional per-row likelihood weights (>= 0; NULL = all ones): every row's NLL term, and
 * with it its gradient and Hessian contribution, is multiplied by its weight (trimming) */
int anml_model_set_weights(anml_model* m, const double* weights, int on_device);
/* combined Gaussian priors of Parameter k, as Parameter.prior_dict holds them
 * (src/anml/parameter/main.py:167-227): direct (mean, sd) of length p_k (sd may
 * be +inf) and linear (mat (m, p_k), mean (m), sd (m)), m may be 0. */
int anml_model_set_gaussian_prior(anml_model* m, int k, const double* mean, const double* sd, int n_lin,
                                  const double* lin_mat, const double* lin_mean, const double* lin_sd);
/* row-sharded runs add the prior on one rank only */
int anml_model_set_prior_enabled(anml_model* m, int enabled);
int anml_model_num_coefs(const anml_model* m, int* total);
/* One evaluation at x (host, length sum p_k).  `want` is a mask of ANML_WANT_*;
 * outputs are caller-owned host buffers (obj: 1, grad: P, hess: P*P row-major),
 * NULL where not wanted.  Log/Logit domain violations return ANML_EDOMAIN. */
int anml_model_eval(anml_model* m, const double* x_host, int want, double* obj, double* grad, double* hess);
/* Same, leaving the packed result [obj | grad(P) | hess(P*P) | domain flag] in
 * device memory (packed_dev, 2 + P + P*P doubles) on `stream` (a cudaStream_t;
 * NULL = the model's own stream, cudaStreamLegacy (0x1) = the legacy default
 * stream) without synchronising: the row-sharded caller all-reduces it over
 * NCCL before reading it. */
int anml_model_eval_async(anml_model* m, const double* x_host, int want, double* packed_dev, void* stream);
/* Ordering: an evaluation queued with anml_model_eval_async on a caller's stream is followed by an
 * event; the setters, anml_model_eval and anml_model_destroy wait for it, and an evaluation on a
 * different stream is ordered behind it on the device. */
/* CUDA-event time of the dominant kernel, accumulated over evaluations since
 * the last reset (what bench.py reports as roofline.achieved) */
int anml_model_kernel_time(anml_model* m, int reset, double* total_ms, int64_t* launches, int64_t* all_launches);
/* name and template arguments of the kernel that anml_model_kernel_time last timed, e.g.
 * "fused_hess_kernel<4, 2, 0, 0, 1, 32>" (bench.py looks its ncu DRAM traffic up under that key) */
int anml_model_main_kernel(const anml_model* m, char* out, int out_len);
/* the timed kernels of the last evaluation in launch order, ';'-separated (a mean / variance model launches three) */
int anml_model_last_kernels(const anml_model* m, char* out, int out_len);
/* Evaluation switches (all default 0 unless noted; development / A-B use, results stay within tolerance either way):
 *   "hess_i8" (default 1)       Hessian of an eligible model (at least 8192 rows on the device, one parameter, exactly 64 dense columns, binomial/expit or
 *                               gaussian/identity, no per-row weights) on the INT8 tensor cores
 *                               (csrc/fused_hess_i8.cuh), and block (1,1) of a mean (identity) / variance (exp) model over one
 *                               64-column design; 0 keeps them on the FP64 tensor cores (DMMA), 2 lifts the row minimum.  Environment
 *                               ANML_B200_HESS_I8=0 switches it off for the whole process.
 *   "hess_i8_rescan"            forget the per-column value ranges the INT8 path recorded for the design (taken again at
 *                               the next evaluation); needed only after rewriting a WRAPPED device matrix in place --
 *                               an evaluation that meets a value outside the recorded range fails with ANML_ESTATE
 *                               rather than returning a wrong Hessian, and rescans by itself
 *   "force_generic"             the general path (linpred / rowterms / xtr / wide_hessian) whatever the model shape
 *   "no_fast_math_paths"        the reference's literal link / family formulas instead of the canonical-pair shortcuts
 *   "no_warp_specialisation"    symmetric 8-warp Hessian kernel (fused_eval_kernel<.., true, ..>)
 *   "force_signed_weights"      sign handling of w < 0 even for canonical pairs
 *   "banded_register_prefetch"  register-prefetch variant of the banded spline evaluator */
int anml_model_set_option(anml_model* m, const char* name, int64_t value);

/* ---- row-sharded evaluation: all-reduce of the packed record over peer memory ----
 * The reference has no distributed layer; this is the exchange step the north star names ("only the
 * p-vector gradient and the p x p Hessian all-reduced over NVLink").  One process per GPU of one node.
 * Each rank creates a window (cudaMalloc), exports its cudaIpcMemHandle (anml_comm_handle_bytes bytes),
 * the host layer gathers the handles of all ranks in rank order (any transport: torch.distributed,
 * MPI, a file) and hands them to anml_comm_connect, then synchronises the ranks once.  After that
 *   anml_comm_allreduce_async(c, buf, n, stream)
 * sums buf (device, n <= max_elems doubles) over the ranks in place with ONE kernel launch on `stream`
 * (a cudaStream_t; NULL or cudaStreamLegacy = the legacy default stream): every element is pushed, value
 * and evaluation number in the same 16-byte store, into every peer's window through NVLink and summed in
 * rank order from the local window, so every rank gets the same bits.  Every rank must call it the same
 * number of times.  Meant for latency-bound records (max_elems <= 2^22; the Python layer uses NCCL above
 * 2^16 elements).  anml_comm_error reports the evaluation number at which a peer failed to arrive within
 * the timeout (ANML_B200_COMM_TIMEOUT_S, default 20 s; the element is then NaN), 0 if none. */
typedef struct anml_comm anml_comm;
int anml_comm_handle_bytes(int* out);
int anml_comm_create(int device, int rank, int world, int64_t max_elems, anml_comm** out);
int anml_comm_export(anml_comm* c, void* handle_out);
int anml_comm_connect(anml_comm* c, const void* all_handles);
int anml_comm_allreduce_async(anml_comm* c, double* buf_dev, int64_t n_elems, void* stream);
int anml_comm_error(anml_comm* c, int64_t* failed_epoch);
int anml_comm_destroy(anml_comm* c);

#ifdef __cplusplus
}
#endif
#endif /* ANML_B200_H */
