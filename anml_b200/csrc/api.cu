// C-ABI of libanml_b200 (declared in include/anml_b200.h): handles, TMA
// descriptors, kernel dispatch.  Host-side only; kernels live in the .cuh files.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>

#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/anml_b200.h"
#include "comm.cuh"
#include "fused_eval.cuh"
#include "finish.cuh"
#include "fused_hess.cuh"
#include "fused_hess_i8.cuh"
#include "generic.cuh"
#include "spline.cuh"
#include "spline_eval.cuh"
#include "upload.cuh"
#include "wide_hessian.cuh"

using namespace anml;

// ----------------------------------------------------------------------------
// errors
// ----------------------------------------------------------------------------
static thread_local std::string g_err;
static thread_local char g_kernel[160];  // signature of the kernel a launch wrapper just started (bench.py keys its ncu traffic on it)

static int fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CU_TRY(expr)                                                                                       \
    do {                                                                                                   \
        cudaError_t e__ = (expr);                                                                          \
        if (e__ != cudaSuccess) {                                                                          \
            const int code__ = (e__ == cudaErrorMemoryAllocation) ? ANML_ENOMEM : ANML_ECUDA;              \
            (void)cudaGetLastError();                                                                      \
            return fail(code__, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
        }                                                                                                  \
    } while (0)

#define TRY(expr)                 \
    do {                          \
        int rc__ = (expr);        \
        if (rc__ != ANML_OK) return rc__; \
    } while (0)

static int round_up(int v, int m) { return (v + m - 1) / m * m; }
static long long round_up_ll(long long v, long long m) { return (v + m - 1) / m * m; }

// ----------------------------------------------------------------------------
// handles
// ----------------------------------------------------------------------------
struct anml_design {
    int device = 0;
    long long n = 0;
    int p = 0;
    long long ld = 0;
    double* X = nullptr;
    bool owned = false;
    CUtensorMap tmap;    // box 16 rows x 16 cols, 128B swizzle
    CUtensorMap tmap32;  // box 32 rows x 16 cols, 128B swizzle
    CUtensorMap tmap3d;  // (16 cols, rows, 16-column groups), box 16 x 32 x p/16: one TMA per 32-row block (p % 16 == 0 only)
    bool has3d = false;
    CUtensorMap tmap64;    // box 64 rows x 16 cols: 64-row blocks of narrow designs (p <= 32)
    CUtensorMap tmap3d64;  // box 16 x 64 x p/16
    bool has64 = false, has3d64 = false;
    int sm_count = 0;
    unsigned long long version = 0;  // bumped whenever the library rewrites X (column maxima cached by a model go stale)
    // deferred build (anml_design_create_deferred): X is allocated, and the recorded spline blocks are
    // evaluated, the first time something needs the dense matrix.  A one-spline model on the banded
    // evaluator never does, so BASELINE config #3 no longer carries an 8.8 GB design it does not read.
    bool deferred = false;
    struct PendingSpline {
        int col0;
        const double* x_dev;  // borrowed: the caller keeps the abscissae alive as long as the design
        std::vector<double> knots;
        int degree, ldegree, rdegree, order;
    };
    std::vector<PendingSpline> pending;
};

struct PriorStore {
    bool set = false;
    int m = 0;
    double *mean = nullptr, *inv_var = nullptr, *lin_mat = nullptr, *lin_mean = nullptr, *lin_inv_var = nullptr,
           *lin_hess = nullptr;
};

struct ParamSlot {
    const anml_design* design = nullptr;
    int link = ANML_LINK_IDENTITY;
    const double* offset = nullptr;  // device, borrowed
    int off = 0;                     // first coefficient in x
    int pad_off = 0;                 // offset of the zero-padded copy inside d_beta
    PriorStore prior;
};

struct EventPair {
    cudaEvent_t a, b;
};

struct anml_model {
    int device = 0;
    long long n = 0, n_pad = 0;
    int family = 0, K = 0, P = 0;
    ParamSlot par[2];
    const double* obs = nullptr;
    const double* se = nullptr;
    const double* weights = nullptr;
    double* obs_owned = nullptr;
    double* se_owned = nullptr;
    double* weights_owned = nullptr;
    bool prior_enabled = true;
    bool force_generic = false, no_fast = false, no_specialise = false, banded_registers = false;
    bool force_signed = false;
    unsigned long long se_version = 0;  // bumped by anml_model_set_obs_se
    std::string last_kernels;           // timed kernels of the last evaluation, ';'-separated
    bool i8_force = false;  // option "hess_i8" = 2: also below I8_MIN_ROWS (tests)
    bool no_i8 = false;  // option "hess_i8" = 0: keep the Hessian of 64-column logistic / Gaussian models on DMMA
    // INT8-tensor-core Hessian (fused_hess_i8.cuh): column scales of the design they were taken from, flush scratch
    struct I8State {
        const anml_design* design = nullptr;
        const double* X = nullptr;
        long long n = 0;
        unsigned long long version = 0;
        bool usable = false;
        const double* se = nullptr;           // obs_se vector the bound on sqrt(w) = 1 / se was taken from
        unsigned long long se_version = 0;
        double sv_scale = 1.0;                // power of two with sqrt(w) sv_scale <= 1/2

        double* cs = nullptr;                 // [64] scales | [64] reciprocals
        unsigned long long* colmax = nullptr; // [64] bit patterns of the column maxima; [64]: max of a weight vector (wmax_kernel)
        double* scratch = nullptr;            // [sm_count][128][64]
    } i8;
    cudaStream_t stream = nullptr;
    int sm_count = 0;
    // evaluation buffers (allocated by ensure_buffers once P is known)
    bool ready = false;
    int beta_len = 0;
    double *d_beta = nullptr, *h_beta = nullptr;
    double *d_packed = nullptr, *h_packed = nullptr;
    double* d_partials = nullptr;
    long long partials_len = 0;
    int* d_status = nullptr;
    cudaEvent_t beta_copied = nullptr;  // guards reuse of h_beta across async evaluations
    bool beta_pending = false;
    // end of the last evaluation: anml_model_eval_async may run on a caller's stream, while the setters and
    // the synchronous evaluation use the model's own -- they order themselves behind this event
    cudaEvent_t eval_done = nullptr;
    bool eval_pending = false;
    cudaStream_t last_stream = nullptr;
    // general-path vectors
    double* vec[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // y0 y1 r0 r1 w00 w01 w11
    double* d_wide = nullptr;
    long long wide_len = 0;
    // banded spline evaluator (spline_eval.cuh): rows sorted by x, per-row vectors gathered into that order
    struct Banded {
        bool enabled = false, gathered = false;
        int nk = 0, degree = 0, nb = 0, rec = 0, grid = 0;
        long long rows_per_cta = 0;
        double *xs = nullptr, *t_dev = nullptr, *coef = nullptr, *partials = nullptr, *reduced = nullptr;
        long long* seg = nullptr;  // first sorted row of every knot interval
        int32_t* perm = nullptr;
        double *obs_s = nullptr, *se_s = nullptr, *off_s = nullptr, *w_s = nullptr;
    } banded;
    // timing of the dominant kernel
    std::vector<EventPair> ev_pool;
    size_t ev_used = 0;
    double ev_accum_ms = 0.0;
    long long main_launches = 0, all_launches = 0;
    std::string main_kernel;  // the last timed kernel, with its template arguments
};

// ----------------------------------------------------------------------------
// misc
// ----------------------------------------------------------------------------
extern "C" const char* anml_last_error(void) { return g_err.c_str(); }
extern "C" int anml_version(void) { return ANML_B200_VERSION; }

extern "C" int anml_device_count(int* count) {
    if (!count) return fail(ANML_EINVAL, "count is NULL");
    int c = 0;
    cudaError_t e = cudaGetDeviceCount(&c);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        *count = 0;
        return fail(ANML_ECUDA, "cudaGetDeviceCount failed: %s", cudaGetErrorString(e));
    }
    *count = c;
    return ANML_OK;
}

extern "C" int anml_device_info(int device, int* sm_count, int64_t* total_mem, int* cc_major, int* cc_minor) {
    cudaDeviceProp prop;
    CU_TRY(cudaGetDeviceProperties(&prop, device));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (total_mem) *total_mem = (int64_t)prop.totalGlobalMem;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    return ANML_OK;
}

extern "C" int anml_dev_alloc(int device, int64_t bytes, void** out) {
    if (!out || bytes < 0) return fail(ANML_EINVAL, "anml_dev_alloc: bad arguments");
    CU_TRY(cudaSetDevice(device));
    CU_TRY(cudaMalloc(out, bytes > 0 ? (size_t)bytes : 16));
    return ANML_OK;
}
extern "C" int anml_dev_free(int device, void* dev) {
    CU_TRY(cudaSetDevice(device));
    CU_TRY(cudaFree(dev));
    return ANML_OK;
}
extern "C" int anml_dev_upload(int device, void* dst, const void* src, int64_t bytes) {
    if (bytes < 0 || (bytes > 0 && (!dst || !src))) return fail(ANML_EINVAL, "anml_dev_upload: bad arguments");
    CU_TRY(cudaSetDevice(device));
    if (bytes % 8 == 0 && (reinterpret_cast<uintptr_t>(dst) & 7) == 0 && (reinterpret_cast<uintptr_t>(src) & 7) == 0) {
        CU_TRY(upload_vector(device, 0, static_cast<double*>(dst), static_cast<const double*>(src), bytes / 8));
        return ANML_OK;
    }
    CU_TRY(cudaMemcpy(dst, src, (size_t)bytes, cudaMemcpyHostToDevice));
    return ANML_OK;
}
extern "C" int anml_upload_release(int device) {
    if (device < 0) {
        for (int d = 0; d < UPLOAD_MAX_DEVICES; ++d) upload_release_device(d);
    } else {
        upload_release_device(device);
    }
    return ANML_OK;
}
extern "C" int anml_dev_download(int device, void* dst, const void* src, int64_t bytes) {
    CU_TRY(cudaSetDevice(device));
    CU_TRY(cudaMemcpy(dst, src, (size_t)bytes, cudaMemcpyDeviceToHost));
    return ANML_OK;
}

// ----------------------------------------------------------------------------
// TMA descriptor (driver entry point resolved through the runtime: no -lcuda)
// ----------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int encode_design_tmap(anml_design* d) {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        CU_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres));
        if (!sym || qres != cudaDriverEntryPointSuccess) return fail(ANML_ECUDA, "cuTensorMapEncodeTiled not available in this driver");
        fn = (EncodeTiledFn)sym;
    }
    memset(&d->tmap, 0, sizeof(d->tmap));
    memset(&d->tmap32, 0, sizeof(d->tmap32));
    if (d->n == 0) return ANML_OK;
    cuuint64_t gdim[2] = {(cuuint64_t)d->p, (cuuint64_t)d->n};
    cuuint64_t gstride[1] = {(cuuint64_t)d->ld * sizeof(double)};
    cuuint32_t box[2] = {16, 16};
    cuuint32_t estride[2] = {1, 1};
    CUresult r = fn(&d->tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, d->X, gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(ANML_ECUDA, "cuTensorMapEncodeTiled failed with CUresult %d (n=%lld p=%d ld=%lld)", (int)r, d->n, d->p, d->ld);
    cuuint32_t box32[2] = {16, 32};
    r = fn(&d->tmap32, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, d->X, gdim, gstride, box32, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(ANML_ECUDA, "cuTensorMapEncodeTiled (32-row box) failed with CUresult %d", (int)r);
    memset(&d->tmap3d, 0, sizeof(d->tmap3d));
    d->has3d = false;
    if (d->p % 16 == 0 && d->p <= 64) {
        // the whole 32-row block in one TMA instruction: dim2 walks the 16-column groups (stride 128 B),
        // so shared memory gets [group][row][16 cols] = the layout of the per-group boxes
        cuuint64_t gdim3[3] = {16, (cuuint64_t)d->n, (cuuint64_t)(d->p / 16)};
        cuuint64_t gstride3[2] = {(cuuint64_t)d->ld * sizeof(double), 128};
        cuuint32_t box3[3] = {16, 32, (cuuint32_t)(d->p / 16)};
        cuuint32_t estride3[3] = {1, 1, 1};
        r = fn(&d->tmap3d, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, d->X, gdim3, gstride3, box3, estride3, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        d->has3d = (r == CUDA_SUCCESS);
    }
    memset(&d->tmap64, 0, sizeof(d->tmap64));
    memset(&d->tmap3d64, 0, sizeof(d->tmap3d64));
    d->has64 = d->has3d64 = false;
    if (d->p <= 32) {
        cuuint32_t box64[2] = {16, 64};
        r = fn(&d->tmap64, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, d->X, gdim, gstride, box64, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        d->has64 = (r == CUDA_SUCCESS);
        if (d->has64 && d->p % 16 == 0) {
            cuuint64_t gdim3[3] = {16, (cuuint64_t)d->n, (cuuint64_t)(d->p / 16)};
            cuuint64_t gstride3[2] = {(cuuint64_t)d->ld * sizeof(double), 128};
            cuuint32_t box3[3] = {16, 64, (cuuint32_t)(d->p / 16)};
            cuuint32_t estride3[3] = {1, 1, 1};
            r = fn(&d->tmap3d64, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, d->X, gdim3, gstride3, box3, estride3, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            d->has3d64 = (r == CUDA_SUCCESS);
        }
    }
    return ANML_OK;
}

// ----------------------------------------------------------------------------
// design matrix
// ----------------------------------------------------------------------------
static int design_finish(anml_design* d) {
    cudaDeviceProp prop;
    CU_TRY(cudaGetDeviceProperties(&prop, d->device));
    if (prop.major < 10) return fail(ANML_EUNSUPPORTED, "anml_b200 needs an sm_100 (B200) device, found sm_%d%d", prop.major, prop.minor);
    d->sm_count = prop.multiProcessorCount;
    return encode_design_tmap(d);
}

static int run_spline(int device, int sm_count, const double* x_dev, long long n, const double* knots, int nk, int degree,
                      int ldegree, int rdegree, int order, double* out_dev, long long ld, int col0, int32_t* idx_dev);

// a deferred design becomes a real one: allocate, encode the tensor maps, run the recorded spline blocks
static int design_materialize(anml_design* d) {
    if (!d || !d->deferred || d->X) return ANML_OK;
    CU_TRY(cudaSetDevice(d->device));
    const size_t bytes = (size_t)(d->n > 0 ? d->n : 1) * d->ld * sizeof(double);
    cudaError_t e = cudaMalloc(&d->X, bytes);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        d->X = nullptr;
        return fail(ANML_ENOMEM, "cudaMalloc of %zu bytes for the design matrix failed: %s", bytes, cudaGetErrorString(e));
    }
    d->owned = true;
    e = cudaMemset(d->X, 0, bytes);
    int rc = (e == cudaSuccess) ? encode_design_tmap(d) : fail(ANML_ECUDA, "cudaMemset failed: %s", cudaGetErrorString(e));
    for (size_t i = 0; rc == ANML_OK && i < d->pending.size(); ++i) {
        const anml_design::PendingSpline& ps = d->pending[i];
        rc = run_spline(d->device, d->sm_count, ps.x_dev, d->n, ps.knots.data(), (int)ps.knots.size(), ps.degree, ps.ldegree, ps.rdegree,
                        ps.order, d->X, d->ld, ps.col0, nullptr);
    }
    if (rc != ANML_OK) {
        cudaFree(d->X);
        d->X = nullptr;
        return rc;
    }
    d->pending.clear();
    return ANML_OK;
}

extern "C" int anml_design_create_deferred(int device, int64_t n_rows, int n_cols, anml_design** out) {
    if (!out || n_rows < 0 || n_cols <= 0) return fail(ANML_EINVAL, "anml_design_create_deferred: need n_rows >= 0 and n_cols > 0");
    if (n_rows > 2147483000LL) return fail(ANML_EUNSUPPORTED, "anml_design_create_deferred: at most 2^31 rows per device");
    CU_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail(ANML_EUNSUPPORTED, "anml_b200 needs an sm_100 (B200) device, found sm_%d%d", prop.major, prop.minor);
    anml_design* d = new (std::nothrow) anml_design();
    if (!d) return fail(ANML_ENOMEM, "out of host memory");
    d->device = device;
    d->n = n_rows;
    d->p = n_cols;
    d->ld = round_up(n_cols, 2);
    d->deferred = true;
    d->sm_count = prop.multiProcessorCount;
    *out = d;
    return ANML_OK;
}

extern "C" int anml_design_is_materialized(const anml_design* d, int* out) {
    if (!d || !out) return fail(ANML_EINVAL, "NULL argument");
    *out = d->X != nullptr ? 1 : 0;
    return ANML_OK;
}

extern "C" int anml_design_create(int device, int64_t n_rows, int n_cols, anml_design** out) {
    if (!out || n_rows < 0 || n_cols <= 0) return fail(ANML_EINVAL, "anml_design_create: need n_rows >= 0 and n_cols > 0");
    if (n_rows > 2147483000LL) return fail(ANML_EUNSUPPORTED, "anml_design_create: at most 2^31 rows per device");
    CU_TRY(cudaSetDevice(device));
    anml_design* d = new (std::nothrow) anml_design();
    if (!d) return fail(ANML_ENOMEM, "out of host memory");
    d->device = device;
    d->n = n_rows;
    d->p = n_cols;
    d->ld = round_up(n_cols, 2);
    d->owned = true;
    const size_t bytes = (size_t)(n_rows > 0 ? n_rows : 1) * d->ld * sizeof(double);
    cudaError_t e = cudaMalloc(&d->X, bytes);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        delete d;
        return fail(ANML_ENOMEM, "cudaMalloc of %zu bytes for the design matrix failed: %s", bytes, cudaGetErrorString(e));
    }
    e = cudaMemset(d->X, 0, bytes);
    if (e != cudaSuccess) {
        cudaFree(d->X);
        delete d;
        return fail(ANML_ECUDA, "cudaMemset failed: %s", cudaGetErrorString(e));
    }
    int rc = design_finish(d);
    if (rc != ANML_OK) {
        cudaFree(d->X);
        delete d;
        return rc;
    }
    *out = d;
    return ANML_OK;
}

extern "C" int anml_design_wrap_device(int device, int64_t n_rows, int n_cols, double* dev, int64_t ld, anml_design** out) {
    if (!out || !dev || n_rows < 0 || n_cols <= 0 || ld < n_cols) return fail(ANML_EINVAL, "anml_design_wrap_device: bad shape");
    if ((ld & 1) || (reinterpret_cast<uintptr_t>(dev) & 15)) return fail(ANML_EINVAL, "anml_design_wrap_device: ld must be even and the pointer 16-byte aligned (TMA row stride)");
    if (n_rows > 2147483000LL) return fail(ANML_EUNSUPPORTED, "anml_design_wrap_device: at most 2^31 rows per device");
    CU_TRY(cudaSetDevice(device));
    anml_design* d = new (std::nothrow) anml_design();
    if (!d) return fail(ANML_ENOMEM, "out of host memory");
    d->device = device;
    d->n = n_rows;
    d->p = n_cols;
    d->ld = ld;
    d->X = dev;
    d->owned = false;
    int rc = design_finish(d);
    if (rc != ANML_OK) {
        delete d;
        return rc;
    }
    *out = d;
    return ANML_OK;
}

extern "C" int anml_design_destroy(anml_design* d) {
    if (!d) return ANML_OK;
    cudaSetDevice(d->device);
    if (d->owned && d->X) cudaFree(d->X);
    delete d;
    return ANML_OK;
}

extern "C" int anml_design_shape(const anml_design* d, int64_t* n_rows, int* n_cols, int64_t* ld, double** dev) {
    if (!d) return fail(ANML_EINVAL, "design is NULL");
    if (n_rows) *n_rows = d->n;
    if (n_cols) *n_cols = d->p;
    if (ld) *ld = d->ld;
    if (dev) {
        TRY(design_materialize(const_cast<anml_design*>(d)));
        *dev = d->X;
    }
    return ANML_OK;
}

static int grid_for(long long work, int threads, int sm_count, int per_sm = 8) {
    long long blocks = (work + threads - 1) / threads;
    long long cap = (long long)sm_count * per_sm;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

extern "C" int anml_design_set_columns(anml_design* d, int col0, int ncols, const double* src, int64_t src_ld, int src_on_device) {
    if (!d || !src) return fail(ANML_EINVAL, "anml_design_set_columns: NULL argument");
    if (col0 < 0 || ncols <= 0 || col0 + ncols > d->p || src_ld < ncols) return fail(ANML_EINVAL, "anml_design_set_columns: columns [%d, %d) outside the %d-column design or src_ld too small", col0, col0 + ncols, d->p);
    if (d->n == 0) return ANML_OK;
    CU_TRY(cudaSetDevice(d->device));
    TRY(design_materialize(d));
    d->version++;
    if (src_on_device) {
        scatter_columns_kernel<<<grid_for(d->n * ncols, 256, d->sm_count), 256>>>(d->X, d->n, d->ld, col0, ncols, src, src_ld);
        CU_TRY(cudaGetLastError());
        CU_TRY(cudaDeviceSynchronize());
        return ANML_OK;
    }
    // host source: pinned lanes filled by a few host threads, DMA + column scatter per lane (upload.cuh)
    const UploadJob job{d->X, d->ld, col0, d->n, ncols, src, src_ld};
    const cudaError_t e = upload_run(d->device, d->sm_count, job);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        return fail(ANML_ECUDA, "anml_design_set_columns: %s", cudaGetErrorString(e));
    }
    return ANML_OK;
}

extern "C" int anml_design_validate_columns(const anml_design* d, int col0, int ncols, int* flags_out) {
    if (!d || !flags_out) return fail(ANML_EINVAL, "anml_design_validate_columns: NULL argument");
    if (col0 < 0 || ncols <= 0 || col0 + ncols > d->p) return fail(ANML_EINVAL, "anml_design_validate_columns: columns out of range");
    for (int c = 0; c < ncols; ++c) flags_out[c] = 0;
    if (d->n == 0) return ANML_OK;
    CU_TRY(cudaSetDevice(d->device));
    TRY(design_materialize(const_cast<anml_design*>(d)));
    int* flags = nullptr;
    CU_TRY(cudaMalloc(&flags, (size_t)ncols * sizeof(int)));
    cudaError_t e = cudaMemset(flags, 0, (size_t)ncols * sizeof(int));
    if (e == cudaSuccess) {
        validate_columns_kernel<<<grid_for(d->n * ncols, 256, d->sm_count), 256>>>(d->X, d->n, d->ld, col0, ncols, flags);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(flags_out, flags, (size_t)ncols * sizeof(int), cudaMemcpyDeviceToHost);
    cudaFree(flags);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        return fail(ANML_ECUDA, "anml_design_validate_columns: %s", cudaGetErrorString(e));
    }
    return ANML_OK;
}

// ----------------------------------------------------------------------------
// knot placement from data (getter/spline.py:92-99)
// ----------------------------------------------------------------------------
static int device_sm_count(int device, int* out) {
    CU_TRY(cudaDeviceGetAttribute(out, cudaDevAttrMultiProcessorCount, device));
    return ANML_OK;
}

// x on the device: either the caller's pointer or an uploaded copy (freed by the caller through *owned)
static int vector_on_device(int device, int sm_count, const double* x, int64_t n, int on_device, const double** view, double** owned) {
    *owned = nullptr;
    if (on_device) {
        *view = x;
        return ANML_OK;
    }
    CU_TRY(cudaMalloc(owned, (size_t)n * sizeof(double)));
    const cudaError_t e = upload_vector(device, sm_count, *owned, x, n);
    if (e != cudaSuccess) {
        cudaFree(*owned);
        *owned = nullptr;
        (void)cudaGetLastError();
        return fail(ANML_ECUDA, "upload failed: %s", cudaGetErrorString(e));
    }
    *view = *owned;
    return ANML_OK;
}

extern "C" int anml_vector_minmax(int device, const double* x, int64_t n, int x_on_device, double* out_min, double* out_max, int* has_nan) {
    if (!x || !out_min || !out_max || !has_nan) return fail(ANML_EINVAL, "anml_vector_minmax: NULL argument");
    if (n <= 0) return fail(ANML_EINVAL, "anml_vector_minmax: zero-size array has no minimum");
    CU_TRY(cudaSetDevice(device));
    int sms = 0;
    TRY(device_sm_count(device, &sms));
    const double* xd = nullptr;
    double* owned = nullptr;
    TRY(vector_on_device(device, sms, x, n, x_on_device, &xd, &owned));
    const int grid = grid_for(n, 256, sms, 8);
    double* part = nullptr;
    cudaError_t e = cudaMalloc(&part, (size_t)grid * (2 * sizeof(double) + sizeof(int)));
    std::vector<double> hmin(grid), hmax(grid);
    std::vector<int> hnan(grid);
    if (e == cudaSuccess) {
        double* pmin = part;
        double* pmax = part + grid;
        int* pnan = reinterpret_cast<int*>(part + 2 * (size_t)grid);
        minmax_kernel<<<grid, 256>>>(xd, n, pmin, pmax, pnan);
        e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaMemcpy(hmin.data(), pmin, grid * sizeof(double), cudaMemcpyDeviceToHost);
        if (e == cudaSuccess) e = cudaMemcpy(hmax.data(), pmax, grid * sizeof(double), cudaMemcpyDeviceToHost);
        if (e == cudaSuccess) e = cudaMemcpy(hnan.data(), pnan, grid * sizeof(int), cudaMemcpyDeviceToHost);
    }
    if (part) cudaFree(part);
    if (owned) cudaFree(owned);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        return fail(ANML_ECUDA, "anml_vector_minmax: %s", cudaGetErrorString(e));
    }
    double lo = hmin[0], hi = hmax[0];
    int nan = hnan[0];
    for (int b = 1; b < grid; ++b) {
        lo = std::fmin(lo, hmin[b]);
        hi = std::fmax(hi, hmax[b]);
        nan |= hnan[b];
    }
    *out_min = lo;
    *out_max = hi;
    *has_nan = nan;
    return ANML_OK;
}

extern "C" int anml_vector_order_stats(int device, const double* x, int64_t n, int x_on_device, const int64_t* ranks, int n_ranks,
                                       double* out) {
    if (!x || !ranks || !out || n_ranks <= 0) return fail(ANML_EINVAL, "anml_vector_order_stats: bad arguments");
    if (n <= 0) return fail(ANML_EINVAL, "anml_vector_order_stats: empty vector");
    if (n > 2147483000LL) return fail(ANML_EUNSUPPORTED, "anml_vector_order_stats: at most 2^31 values");
    for (int i = 0; i < n_ranks; ++i)
        if (ranks[i] < 0 || ranks[i] >= n) return fail(ANML_EINVAL, "anml_vector_order_stats: rank %lld outside [0, %lld)", (long long)ranks[i], (long long)n);
    CU_TRY(cudaSetDevice(device));
    int sms = 0;
    TRY(device_sm_count(device, &sms));
    const double* xd = nullptr;
    double* owned = nullptr;
    TRY(vector_on_device(device, sms, x, n, x_on_device, &xd, &owned));
    double* sorted = nullptr;
    void* temp = nullptr;
    long long* ranks_dev = nullptr;
    double* out_dev = nullptr;
    size_t temp_bytes = 0;
    cudaError_t e = cub::DeviceRadixSort::SortKeys(nullptr, temp_bytes, xd, sorted, (int)n);
    if (e == cudaSuccess) e = cudaMalloc(&sorted, (size_t)n * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&temp, temp_bytes ? temp_bytes : 16);
    if (e == cudaSuccess) e = cudaMalloc(&ranks_dev, (size_t)n_ranks * sizeof(long long));
    if (e == cudaSuccess) e = cudaMalloc(&out_dev, (size_t)n_ranks * sizeof(double));
    if (e == cudaSuccess) e = cub::DeviceRadixSort::SortKeys(temp, temp_bytes, xd, sorted, (int)n);
    if (e == cudaSuccess) {
        static_assert(sizeof(long long) == sizeof(int64_t), "rank type");
        e = cudaMemcpy(ranks_dev, ranks, (size_t)n_ranks * sizeof(long long), cudaMemcpyHostToDevice);
    }
    if (e == cudaSuccess) {
        gather_kernel<<<(n_ranks + 127) / 128, 128>>>(sorted, ranks_dev, n_ranks, out_dev);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(out, out_dev, (size_t)n_ranks * sizeof(double), cudaMemcpyDeviceToHost);
    const int code = (e == cudaErrorMemoryAllocation) ? ANML_ENOMEM : ANML_ECUDA;
    if (sorted) cudaFree(sorted);
    if (temp) cudaFree(temp);
    if (ranks_dev) cudaFree(ranks_dev);
    if (out_dev) cudaFree(out_dev);
    if (owned) cudaFree(owned);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        return fail(code, "anml_vector_order_stats: %s", cudaGetErrorString(e));
    }
    return ANML_OK;
}

static int run_spline(int device, int sm_count, const double* x_dev, long long n, const double* knots, int nk, int degree,
                      int ldegree, int rdegree, int order, double* out_dev, long long ld, int col0, int32_t* idx_dev) {
    if (nk < 2 || nk > SPLINE_MAX_KNOTS) return fail(ANML_EINVAL, "spline needs between 2 and %d knots, got %d", SPLINE_MAX_KNOTS, nk);
    if (degree < 0 || degree > SPLINE_MAX_DEGREE) return fail(ANML_EUNSUPPORTED, "spline degree must be in [0, %d], got %d", SPLINE_MAX_DEGREE, degree);
    if (order < 0) return fail(ANML_EINVAL, "spline derivative order must be non-negative");
    for (int i = 1; i < nk; ++i)
        if (!(knots[i] >= knots[i - 1])) return fail(ANML_EINVAL, "spline knots must be sorted ascending");
    if (!(knots[nk - 1] > knots[0])) return fail(ANML_EINVAL, "spline knots must span a non-empty interval");
    if (ldegree < 0) ldegree = 0;
    if (rdegree < 0) rdegree = 0;
    if (ldegree > degree) ldegree = degree;
    if (rdegree > degree) rdegree = degree;
    const int nt = nk + 2 * degree;
    std::vector<double> t(nt);
    for (int i = 0; i < degree; ++i) t[i] = knots[0];
    for (int i = 0; i < nk; ++i) t[degree + i] = knots[i];
    for (int i = 0; i < degree; ++i) t[degree + nk + i] = knots[nk - 1];
    double* t_dev = nullptr;
    CU_TRY(cudaMalloc(&t_dev, nt * sizeof(double)));
    cudaError_t e = cudaMemcpy(t_dev, t.data(), nt * sizeof(double), cudaMemcpyHostToDevice);
    if (e == cudaSuccess && n > 0) {
        SplineArgs a;
        a.x = x_dev; a.n = n; a.t = t_dev; a.nk = nk; a.degree = degree; a.ldegree = ldegree; a.rdegree = rdegree;
        a.order = order; a.out = out_dev; a.ld = ld; a.col0 = col0; a.idx_out = idx_dev;
        const int nb = nk - 1 + degree;
        const size_t smem = ((size_t)((nt + 1) & ~1) + (size_t)256 * (nb + 1)) * sizeof(double);
        if (smem > 200 * 1024) {
            cudaFree(t_dev);
            return fail(ANML_EUNSUPPORTED, "spline with %d basis functions needs %zu bytes of shared memory per block", nb, smem);
        }
        const int grid = grid_for(n, 256, sm_count, 4);
#define SPLINE_CASE(D)                                                                                                   \
    case D:                                                                                                              \
        if (smem > 48 * 1024) e = cudaFuncSetAttribute(spline_design_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        if (e == cudaSuccess) spline_design_kernel<D><<<grid, 256, smem>>>(a);                                           \
        break;
        switch (degree) {
            SPLINE_CASE(0)
            SPLINE_CASE(1)
            SPLINE_CASE(2)
            SPLINE_CASE(3)
            SPLINE_CASE(4)
            SPLINE_CASE(5)
        }
#undef SPLINE_CASE
        e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
    }
    cudaFree(t_dev);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        return fail(ANML_ECUDA, "spline_design_kernel: %s", cudaGetErrorString(e));
    }
    (void)device;
    return ANML_OK;
}

extern "C" int anml_design_set_spline(anml_design* d, int col0, const double* x, int x_on_device, const double* knots, int n_knots,
                                      int degree, int ldegree, int rdegree, int order, int32_t* interval_out) {
    if (!d || !x || !knots) return fail(ANML_EINVAL, "anml_design_set_spline: NULL argument");
    const int nb = n_knots - 1 + degree;
    if (col0 < 0 || col0 + nb > d->p) return fail(ANML_EINVAL, "anml_design_set_spline: %d basis columns at %d do not fit a %d-column design", nb, col0, d->p);
    d->version++;
    CU_TRY(cudaSetDevice(d->device));
    if (d->deferred && !d->X && x_on_device && !interval_out) {
        // record the block; it is evaluated if and when the dense matrix is needed (design_materialize)
        if (n_knots < 2 || n_knots > SPLINE_MAX_KNOTS) return fail(ANML_EINVAL, "spline needs between 2 and %d knots, got %d", SPLINE_MAX_KNOTS, n_knots);
        if (degree < 0 || degree > SPLINE_MAX_DEGREE) return fail(ANML_EUNSUPPORTED, "spline degree must be in [0, %d], got %d", SPLINE_MAX_DEGREE, degree);
        if (order < 0) return fail(ANML_EINVAL, "spline derivative order must be non-negative");
        for (int i = 1; i < n_knots; ++i)
            if (!(knots[i] >= knots[i - 1])) return fail(ANML_EINVAL, "spline knots must be sorted ascending");
        if (!(knots[n_knots - 1] > knots[0])) return fail(ANML_EINVAL, "spline knots must span a non-empty interval");
        anml_design::PendingSpline ps;
        ps.col0 = col0; ps.x_dev = x; ps.knots.assign(knots, knots + n_knots);
        ps.degree = degree; ps.ldegree = ldegree; ps.rdegree = rdegree; ps.order = order;
        d->pending.push_back(ps);
        return ANML_OK;
    }
    TRY(design_materialize(d));
    double* x_tmp = nullptr;
    int32_t* idx_dev = nullptr;
    const size_t nn = (size_t)(d->n > 0 ? d->n : 1);
    if (!x_on_device) {
        CU_TRY(cudaMalloc(&x_tmp, nn * sizeof(double)));
        cudaError_t e = upload_vector(d->device, d->sm_count, x_tmp, x, d->n);
        if (e != cudaSuccess) {
            cudaFree(x_tmp);
            return fail(ANML_ECUDA, "upload of spline abscissae failed: %s", cudaGetErrorString(e));
        }
    }
    if (interval_out) {
        cudaError_t e = cudaMalloc(&idx_dev, nn * sizeof(int32_t));
        if (e != cudaSuccess) {
            if (x_tmp) cudaFree(x_tmp);
            (void)cudaGetLastError();
            return fail(ANML_ENOMEM, "cudaMalloc for interval indices failed");
        }
    }
    int rc = run_spline(d->device, d->sm_count, x_on_device ? x : x_tmp, d->n, knots, n_knots, degree, ldegree, rdegree, order,
                        d->X, d->ld, col0, idx_dev);
    if (rc == ANML_OK && interval_out && d->n > 0) {
        cudaError_t e = cudaMemcpy(interval_out, idx_dev, (size_t)d->n * sizeof(int32_t), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc = fail(ANML_ECUDA, "download of interval indices failed: %s", cudaGetErrorString(e));
    }
    if (x_tmp) cudaFree(x_tmp);
    if (idx_dev) cudaFree(idx_dev);
    return rc;
}

extern "C" int anml_design_download(const anml_design* d, double* host_out, int64_t out_ld) {
    if (!d || !host_out || out_ld < d->p) return fail(ANML_EINVAL, "anml_design_download: bad arguments");
    if (d->n == 0) return ANML_OK;
    CU_TRY(cudaSetDevice(d->device));
    TRY(design_materialize(const_cast<anml_design*>(d)));
    CU_TRY(cudaMemcpy2D(host_out, (size_t)out_ld * sizeof(double), d->X, (size_t)d->ld * sizeof(double), (size_t)d->p * sizeof(double),
                        (size_t)d->n, cudaMemcpyDeviceToHost));
    return ANML_OK;
}

extern "C" int anml_spline_design(int device, const double* x_host, int64_t n, const double* knots, int n_knots, int degree,
                                  int ldegree, int rdegree, int order, double* out_host, int32_t* interval_out) {
    if (!x_host || !knots || !out_host || n < 0) return fail(ANML_EINVAL, "anml_spline_design: bad arguments");
    const int nb = n_knots - 1 + degree;
    if (nb <= 0) return fail(ANML_EINVAL, "anml_spline_design: no basis functions");
    anml_design* d = nullptr;
    TRY(anml_design_create(device, n, nb, &d));
    int rc = anml_design_set_spline(d, 0, x_host, 0, knots, n_knots, degree, ldegree, rdegree, order, interval_out);
    if (rc == ANML_OK) rc = anml_design_download(d, out_host, nb);
    anml_design_destroy(d);
    return rc;
}

// y (optionally through the link at `order`) for all rows
static int launch_linpred(const anml_design* d, const double* beta_dev, const double* offset, int link, int order, double* out,
                          int* status, cudaStream_t st) {
    if (d->n == 0) return ANML_OK;
    const int npairs = (int)(d->ld / 2);
    int gs = 1;
    while (gs < npairs && gs < 32) gs <<= 1;
    const int threads = 256;
    const long long rows_per_block = (threads / 32) * (32 / gs);
    const int grid = grid_for(d->n * (threads / rows_per_block), threads, d->sm_count, 8);
    const size_t smem = (size_t)d->ld * sizeof(double);
    if (smem > 48 * 1024) return fail(ANML_EUNSUPPORTED, "design wider than 6144 columns is not supported");
#define LP(GS) linpred_kernel<GS><<<grid, threads, smem, st>>>(d->X, d->n, d->p, d->ld, beta_dev, offset, link, order, out, status)
    switch (gs) {
        case 1: LP(1); break;
        case 2: LP(2); break;
        case 4: LP(4); break;
        case 8: LP(8); break;
        case 16: LP(16); break;
        default: LP(32); break;
    }
#undef LP
    CU_TRY(cudaGetLastError());
    return ANML_OK;
}

extern "C" int anml_design_params(const anml_design* d, const double* beta_host, const double* offset_dev, int link, int order,
                                  double* out_host) {
    if (!d || !beta_host || !out_host) return fail(ANML_EINVAL, "anml_design_params: NULL argument");
    if (order < 0 || order > 2) return fail(ANML_EINVAL, "Order must be 0, 1 or 2.");
    if (link < 0 || link > 4) return fail(ANML_EINVAL, "unknown link %d", link);
    if (d->n == 0) return ANML_OK;
    CU_TRY(cudaSetDevice(d->device));
    TRY(design_materialize(const_cast<anml_design*>(d)));
    const long long n = d->n;
    const int p = d->p;
    long long out_elems = n;
    if (order == 1) out_elems = n * p;
    if (order == 2) out_elems = n * p * (long long)p;
    size_t free_b = 0, total_b = 0;
    CU_TRY(cudaMemGetInfo(&free_b, &total_b));
    if ((double)out_elems * 8.0 + (double)n * 8.0 > 0.9 * (double)free_b)
        return fail(ANML_ENOMEM, "get_params(order=%d) would materialise %.3g GB on the device; use the model evaluation instead", order, out_elems * 8.0 / 1e9);
    double *beta_dev = nullptr, *z = nullptr, *out = nullptr;
    int* status = nullptr;
    int rc = ANML_OK;
    cudaError_t e = cudaMalloc(&beta_dev, (size_t)p * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&z, (size_t)n * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&status, sizeof(int));
    if (e == cudaSuccess && order > 0) e = cudaMalloc(&out, (size_t)out_elems * sizeof(double));
    if (e == cudaSuccess) e = cudaMemset(status, 0, sizeof(int));
    if (e == cudaSuccess) e = cudaMemcpy(beta_dev, beta_host, (size_t)p * sizeof(double), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        rc = launch_linpred(d, beta_dev, offset_dev, link, order, z, status, 0);
        if (rc == ANML_OK && order == 1) jacobian_kernel<<<grid_for(out_elems, 256, d->sm_count), 256>>>(d->X, n, p, d->ld, z, out);
        if (rc == ANML_OK && order == 2) tensor2_kernel<<<grid_for(out_elems, 256, d->sm_count), 256>>>(d->X, n, p, d->ld, z, out);
        if (rc == ANML_OK) e = cudaGetLastError();
        if (rc == ANML_OK && e == cudaSuccess) e = cudaMemcpy(out_host, order == 0 ? z : out, (size_t)out_elems * sizeof(double), cudaMemcpyDeviceToHost);
        int flag = 0;
        if (rc == ANML_OK && e == cudaSuccess) e = cudaMemcpy(&flag, status, sizeof(int), cudaMemcpyDeviceToHost);
        if (rc == ANML_OK && e == cudaSuccess && flag) {
            rc = fail(ANML_EDOMAIN, link == ANML_LINK_LOG ? "All values for log function must be positive."
                                                          : "All values for logit function must be strictly between 0 and 1.");
        }
    }
    if (beta_dev) cudaFree(beta_dev);
    if (z) cudaFree(z);
    if (out) cudaFree(out);
    if (status) cudaFree(status);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        return fail(e == cudaErrorMemoryAllocation ? ANML_ENOMEM : ANML_ECUDA, "anml_design_params: %s", cudaGetErrorString(e));
    }
    return rc;
}

// ----------------------------------------------------------------------------
// model
// ----------------------------------------------------------------------------
extern "C" int anml_model_create(int device, int64_t n_rows, int family, int n_params, anml_model** out) {
    if (!out || n_rows < 0) return fail(ANML_EINVAL, "anml_model_create: bad arguments");
    if (family < 0 || family > 3) return fail(ANML_EINVAL, "unknown family %d", family);
    const int need = (family == ANML_FAMILY_GAUSSIAN_MEANVAR) ? 2 : 1;
    if (n_params != need) return fail(ANML_EINVAL, "family %d takes %d parameter(s), got %d", family, need, n_params);
    CU_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail(ANML_EUNSUPPORTED, "anml_b200 needs an sm_100 (B200) device, found sm_%d%d", prop.major, prop.minor);
    anml_model* m = new (std::nothrow) anml_model();
    if (!m) return fail(ANML_ENOMEM, "out of host memory");
    m->device = device;
    m->n = n_rows;
    m->n_pad = round_up_ll(n_rows > 0 ? n_rows : 1, 32);
    m->family = family;
    m->K = n_params;
    m->sm_count = prop.multiProcessorCount;
    cudaError_t e = cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        delete m;
        return fail(ANML_ECUDA, "cudaStreamCreate failed: %s", cudaGetErrorString(e));
    }
    *out = m;
    return ANML_OK;
}

static void free_prior(PriorStore& s) {
    double** ptrs[] = {&s.mean, &s.inv_var, &s.lin_mat, &s.lin_mean, &s.lin_inv_var, &s.lin_hess};
    for (double** p : ptrs) {
        if (*p) cudaFree(*p);
        *p = nullptr;
    }
    s.set = false;
    s.m = 0;
}

static void free_eval_buffers(anml_model* m) {
    if (m->d_beta) cudaFree(m->d_beta);
    if (m->h_beta) cudaFreeHost(m->h_beta);
    if (m->d_packed) cudaFree(m->d_packed);
    if (m->h_packed) cudaFreeHost(m->h_packed);
    if (m->d_partials) cudaFree(m->d_partials);
    if (m->d_status) cudaFree(m->d_status);
    if (m->d_wide) cudaFree(m->d_wide);
    for (double*& v : m->vec) {
        if (v) cudaFree(v);
        v = nullptr;
    }
    m->d_beta = m->h_beta = m->d_packed = m->h_packed = m->d_partials = m->d_wide = nullptr;
    m->d_status = nullptr;
    m->wide_len = 0;
    m->ready = false;
}

static void free_banded(anml_model* m) {
    anml_model::Banded& b = m->banded;
    double* bufs[] = {b.xs, b.t_dev, b.coef, b.partials, b.reduced, b.obs_s, b.se_s, b.off_s, b.w_s};
    for (double* q : bufs)
        if (q) cudaFree(q);
    if (b.perm) cudaFree(b.perm);
    if (b.seg) cudaFree(b.seg);
    b = anml_model::Banded();
}

// Nothing of this model is in flight any more: the last evaluation (possibly on a caller's stream,
// anml_model_eval_async) and the model's own stream have drained.  Setters call it before they touch
// buffers an evaluation reads.
static int model_quiesce(anml_model* m) {
    if (m->eval_pending && m->eval_done) {
        CU_TRY(cudaEventSynchronize(m->eval_done));
        m->eval_pending = false;
    }
    if (m->stream) CU_TRY(cudaStreamSynchronize(m->stream));
    return ANML_OK;
}

extern "C" int anml_model_destroy(anml_model* m) {
    if (!m) return ANML_OK;
    cudaSetDevice(m->device);
    (void)model_quiesce(m);
    if (m->eval_done) cudaEventDestroy(m->eval_done);
    if (m->i8.cs) cudaFree(m->i8.cs);
    if (m->i8.colmax) cudaFree(m->i8.colmax);
    if (m->i8.scratch) cudaFree(m->i8.scratch);
    free_eval_buffers(m);
    free_banded(m);
    for (int k = 0; k < 2; ++k) free_prior(m->par[k].prior);
    if (m->obs_owned) cudaFree(m->obs_owned);
    if (m->se_owned) cudaFree(m->se_owned);
    if (m->weights_owned) cudaFree(m->weights_owned);
    for (EventPair& ep : m->ev_pool) {
        cudaEventDestroy(ep.a);
        cudaEventDestroy(ep.b);
    }
    if (m->beta_copied) cudaEventDestroy(m->beta_copied);
    if (m->stream) cudaStreamDestroy(m->stream);
    delete m;
    return ANML_OK;
}

extern "C" int anml_model_set_param(anml_model* m, int k, const anml_design* design, int link, const double* offset_dev) {
    if (!m || !design) return fail(ANML_EINVAL, "anml_model_set_param: NULL argument");
    if (k < 0 || k >= m->K) return fail(ANML_EINVAL, "parameter index %d out of range (model has %d)", k, m->K);
    if (link < 0 || link > 4) return fail(ANML_EINVAL, "unknown link %d", link);
    if (design->n != m->n) return fail(ANML_EINVAL, "design has %lld rows, model has %lld", design->n, m->n);
    if (design->device != m->device) return fail(ANML_EINVAL, "design lives on device %d, model on %d", design->device, m->device);
    CU_TRY(cudaSetDevice(m->device));
    TRY(model_quiesce(m));
    if (m->par[k].design && m->par[k].design->p != design->p) {
        free_prior(m->par[k].prior);
        free_eval_buffers(m);
    }
    m->par[k].design = design;
    m->par[k].link = link;
    m->par[k].offset = offset_dev;
    free_banded(m);  // tied to the previous design; the caller enables it again after set_param
    return ANML_OK;
}

static int set_row_vector(anml_model* m, const double* src, int on_device, const double** view, double** owned) {
    CU_TRY(cudaSetDevice(m->device));
    TRY(model_quiesce(m));
    m->banded.gathered = false;  // the sorted copies are stale
    if (*owned) {
        CU_TRY(cudaFree(*owned));
        *owned = nullptr;
    }
    *view = nullptr;
    if (!src) return ANML_OK;
    if (on_device) {
        *view = src;  // borrowed
        return ANML_OK;
    }
    CU_TRY(cudaMalloc(owned, (size_t)m->n_pad * sizeof(double)));
    CU_TRY(cudaMemset(*owned, 0, (size_t)m->n_pad * sizeof(double)));
    CU_TRY(upload_vector(m->device, 0, *owned, src, m->n));
    *view = *owned;
    return ANML_OK;
}

extern "C" int anml_model_set_obs(anml_model* m, const double* obs, int on_device) {
    if (!m || !obs) return fail(ANML_EINVAL, "anml_model_set_obs: NULL argument");
    return set_row_vector(m, obs, on_device, &m->obs, &m->obs_owned);
}

extern "C" int anml_model_set_obs_se(anml_model* m, const double* se, int on_device) {
    if (!m) return fail(ANML_EINVAL, "anml_model_set_obs_se: NULL model");
    m->se_version++;
    return set_row_vector(m, se, on_device, &m->se, &m->se_owned);
}

extern "C" int anml_model_set_weights(anml_model* m, const double* weights, int on_device) {
    if (!m) return fail(ANML_EINVAL, "anml_model_set_weights: NULL model");
    return set_row_vector(m, weights, on_device, &m->weights, &m->weights_owned);
}

static int upload_vec(double** dst, const std::vector<double>& v) {
    CU_TRY(cudaMalloc(dst, (v.size() ? v.size() : 1) * sizeof(double)));
    if (!v.empty()) CU_TRY(cudaMemcpy(*dst, v.data(), v.size() * sizeof(double), cudaMemcpyHostToDevice));
    return ANML_OK;
}

extern "C" int anml_model_set_gaussian_prior(anml_model* m, int k, const double* mean, const double* sd, int n_lin,
                                             const double* lin_mat, const double* lin_mean, const double* lin_sd) {
    if (!m) return fail(ANML_EINVAL, "anml_model_set_gaussian_prior: NULL model");
    if (k < 0 || k >= m->K || !m->par[k].design) return fail(ANML_ESTATE, "set the design of parameter %d before its prior", k);
    if (n_lin < 0 || (n_lin > 0 && (!lin_mat || !lin_mean || !lin_sd))) return fail(ANML_EINVAL, "linear prior arrays missing");
    CU_TRY(cudaSetDevice(m->device));
    const int p = m->par[k].design->p;
    // validate and build everything on the host first: a rejected call leaves the installed prior in place
    std::vector<double> mu(p, 0.0), iv(p, 0.0);
    for (int j = 0; j < p; ++j) {
        if (mean) mu[j] = mean[j];
        if (sd) {
            if (!(sd[j] > 0.0)) return fail(ANML_EINVAL, "Gaussian prior standard deviations must be positive.");
            iv[j] = std::isinf(sd[j]) ? 0.0 : 1.0 / (sd[j] * sd[j]);
        }
    }
    std::vector<double> lmean(n_lin), liv(n_lin), lh((size_t)(n_lin > 0 ? p * p : 0), 0.0), lmat((size_t)n_lin * p);
    for (int i = 0; i < n_lin; ++i) {
        if (!(lin_sd[i] > 0.0)) return fail(ANML_EINVAL, "Gaussian prior standard deviations must be positive.");
        lmean[i] = lin_mean[i];
        liv[i] = std::isinf(lin_sd[i]) ? 0.0 : 1.0 / (lin_sd[i] * lin_sd[i]);
        for (int j = 0; j < p; ++j) lmat[(size_t)i * p + j] = lin_mat[(size_t)i * p + j];
    }
    // (M^T / sd^2) M is constant in x: GaussianPrior.hessian, prior/main.py:197
    for (int i = 0; i < n_lin; ++i)
        for (int a = 0; a < p; ++a) {
            const double ma = lmat[(size_t)i * p + a] * liv[i];
            if (ma == 0.0) continue;
            for (int b = 0; b <= a; ++b) lh[(size_t)a * p + b] += ma * lmat[(size_t)i * p + b];
        }
    for (int a = 0; a < p && n_lin > 0; ++a)  // mirror the lower triangle: exactly symmetric
        for (int b = 0; b < a; ++b) lh[(size_t)b * p + a] = lh[(size_t)a * p + b];
    PriorStore fresh;
    int rc = upload_vec(&fresh.mean, mu);
    if (rc == ANML_OK) rc = upload_vec(&fresh.inv_var, iv);
    if (rc == ANML_OK) rc = upload_vec(&fresh.lin_mat, lmat);
    if (rc == ANML_OK) rc = upload_vec(&fresh.lin_mean, lmean);
    if (rc == ANML_OK) rc = upload_vec(&fresh.lin_inv_var, liv);
    if (rc == ANML_OK) rc = upload_vec(&fresh.lin_hess, lh);
    if (rc != ANML_OK) {
        free_prior(fresh);
        return rc;
    }
    fresh.m = n_lin;
    fresh.set = true;
    TRY(model_quiesce(m));  // the old arrays may still be read by an evaluation in flight
    free_prior(m->par[k].prior);
    m->par[k].prior = fresh;
    return ANML_OK;
}

extern "C" int anml_model_set_prior_enabled(anml_model* m, int enabled) {
    if (!m) return fail(ANML_EINVAL, "NULL model");
    m->prior_enabled = enabled != 0;
    return ANML_OK;
}

extern "C" int anml_model_set_option(anml_model* m, const char* name, int64_t value) {
    if (!m || !name) return fail(ANML_EINVAL, "NULL argument");
    if (!strcmp(name, "force_generic")) m->force_generic = value != 0;
    else if (!strcmp(name, "no_fast_math_paths")) m->no_fast = value != 0;
    else if (!strcmp(name, "no_warp_specialisation")) m->no_specialise = value != 0;
    else if (!strcmp(name, "force_signed_weights")) m->force_signed = value != 0;
    else if (!strcmp(name, "banded_register_prefetch")) m->banded_registers = value != 0;
    else if (!strcmp(name, "hess_i8")) {
        m->no_i8 = value == 0;
        m->i8_force = value == 2;
    }
    else if (!strcmp(name, "hess_i8_rescan")) m->i8.design = nullptr;
    else return fail(ANML_EINVAL, "unknown option '%s'", name);
    return ANML_OK;
}

extern "C" int anml_model_num_coefs(const anml_model* m, int* total) {
    if (!m || !total) return fail(ANML_EINVAL, "NULL argument");
    int P = 0;
    for (int k = 0; k < m->K; ++k) {
        if (!m->par[k].design) return fail(ANML_ESTATE, "parameter %d has no design matrix", k);
        P += m->par[k].design->p;
    }
    *total = P;
    return ANML_OK;
}

static int fused_el(int nb16, bool hess) {
    const int nb8 = 2 * nb16, nt = nb8 * (nb8 + 1) / 2;
    return ((hess ? 2 * nt : 0) + nb8 + 1) * 32;
}

static int ensure_buffers_alloc(anml_model* m) {
    int P = 0;
    TRY(anml_model_num_coefs(m, &P));
    if (!m->obs) return fail(ANML_ESTATE, "observations not set");
    m->P = P;
    int at = 0, pad_at = P;
    for (int k = 0; k < m->K; ++k) {
        m->par[k].off = at;
        m->par[k].pad_off = pad_at;
        at += m->par[k].design->p;
        pad_at += round_up(m->par[k].design->p, 64);
    }
    m->beta_len = pad_at;
    const size_t packed = (size_t)2 + P + (size_t)P * P;
    CU_TRY(cudaMalloc(&m->d_beta, m->beta_len * sizeof(double)));
    CU_TRY(cudaMallocHost(&m->h_beta, m->beta_len * sizeof(double)));
    CU_TRY(cudaMalloc(&m->d_packed, packed * sizeof(double)));
    CU_TRY(cudaMallocHost(&m->h_packed, packed * sizeof(double)));
    CU_TRY(cudaMemset(m->d_packed, 0, packed * sizeof(double)));
    int max_ld = 0;
    for (int k = 0; k < m->K; ++k) max_ld = max_ld > (int)m->par[k].design->ld ? max_ld : (int)m->par[k].design->ld;
    long long plen = (long long)m->sm_count * fused_el(4, true);
    const long long xtr_len = (long long)m->sm_count * 4 * max_ld;
    if (xtr_len > plen) plen = xtr_len;
    if ((long long)m->sm_count * 8 > plen) plen = (long long)m->sm_count * 8;
    m->partials_len = plen;
    CU_TRY(cudaMalloc(&m->d_partials, (size_t)plen * sizeof(double)));
    CU_TRY(cudaMalloc(&m->d_status, sizeof(int)));
    CU_TRY(cudaMemset(m->d_status, 0, sizeof(int)));  // re-armed by the last launch of every evaluation
    if (!m->beta_copied) CU_TRY(cudaEventCreateWithFlags(&m->beta_copied, cudaEventDisableTiming));
    if (!m->eval_done) CU_TRY(cudaEventCreateWithFlags(&m->eval_done, cudaEventDisableTiming));
    return ANML_OK;
}

static int ensure_buffers(anml_model* m) {
    if (m->ready) return ANML_OK;
    const int rc = ensure_buffers_alloc(m);
    if (rc != ANML_OK) {
        const std::string msg = g_err;
        free_eval_buffers(m);  // no half-allocated set survives a failure (a retry would leak it)
        g_err = msg;
        return rc;
    }
    m->ready = true;
    return ANML_OK;
}

static int ensure_vec(anml_model* m, int which) {
    if (m->vec[which]) return ANML_OK;
    CU_TRY(cudaMalloc(&m->vec[which], (size_t)m->n_pad * sizeof(double)));
    CU_TRY(cudaMemset(m->vec[which], 0, (size_t)m->n_pad * sizeof(double)));
    return ANML_OK;
}

static EventPair* next_events(anml_model* m) {
    if (m->ev_used == m->ev_pool.size()) {
        if (m->ev_pool.size() >= 4096) return nullptr;  // bounded; beyond that launches go untimed
        EventPair ep;
        if (cudaEventCreate(&ep.a) != cudaSuccess) return nullptr;
        if (cudaEventCreate(&ep.b) != cudaSuccess) {
            cudaEventDestroy(ep.a);
            return nullptr;
        }
        m->ev_pool.push_back(ep);
    }
    return &m->ev_pool[m->ev_used++];
}

// cudaFuncSetAttribute is per device: remember which devices a kernel has been configured on
struct ConfiguredOn {
    std::atomic<unsigned long long> mask{0};
    bool has(int device) const { return device >= 0 && device < 64 && ((mask.load(std::memory_order_acquire) >> device) & 1ull); }
    void set(int device) {
        if (device >= 0 && device < 64) mask.fetch_or(1ull << device, std::memory_order_release);
    }
};

template <int NB16, bool HESS, int MODE>
static int launch_fused_t(const anml_design* d, const FusedArgs& a, int grid, cudaStream_t st) {
    using C = FusedCfg<NB16>;
    static ConfiguredOn configured;
    auto kern = fused_eval_kernel<NB16, HESS, MODE>;
    if (!configured.has(d->device)) {
        CU_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        configured.set(d->device);
    }
    kern<<<grid, C::THREADS, C::SMEM_BYTES, st>>>(d->tmap, a);
    CU_TRY(cudaGetLastError());
    snprintf(g_kernel, sizeof(g_kernel), "fused_eval_kernel<%d, %d, %d>", NB16, HESS ? 1 : 0, MODE);
    return ANML_OK;
}

template <int NB16, int NHELP, bool SIGNED, int MODE, bool TMA3D, int ROWS>
static int launch_hess_tt(const anml_design* d, const FusedArgs& a, int grid, cudaStream_t st) {
    using C = HessCfg<NB16, NHELP, ROWS>;
    static ConfiguredOn configured;
    auto kern = fused_hess_kernel<NB16, NHELP, SIGNED, MODE, TMA3D, ROWS>;
    if (!configured.has(d->device)) {
        CU_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        configured.set(d->device);
    }
    const CUtensorMap& tm = ROWS == 64 ? (TMA3D ? d->tmap3d64 : d->tmap64) : (TMA3D ? d->tmap3d : d->tmap32);
    kern<<<grid, C::THREADS, C::SMEM_BYTES, st>>>(tm, a);
    CU_TRY(cudaGetLastError());
    snprintf(g_kernel, sizeof(g_kernel), "fused_hess_kernel<%d, %d, %d, %d, %d, %d>", NB16, NHELP, SIGNED ? 1 : 0, MODE, TMA3D ? 1 : 0, ROWS);
    return ANML_OK;
}

template <int NB16, int NHELP, bool SIGNED, int MODE>
static int launch_hess_t(const anml_design* d, const FusedArgs& a, int grid, cudaStream_t st) {
    // narrow designs (p <= 32) run 64-row blocks; one 3-D TMA per block when the width is a
    // multiple of 16 (no partial column group)
    if (NB16 <= 2 && d->has64) {
        constexpr int R = NB16 <= 2 ? 64 : 32;  // (keeps the 64-row variants of the wide kernels uninstantiated)
        if (d->has3d64 && d->p == 16 * NB16) return launch_hess_tt<NB16, NHELP, SIGNED, MODE, true, R>(d, a, grid, st);
        return launch_hess_tt<NB16, NHELP, SIGNED, MODE, false, R>(d, a, grid, st);
    }
    if (d->has3d && d->p == 16 * NB16) return launch_hess_tt<NB16, NHELP, SIGNED, MODE, true, 32>(d, a, grid, st);
    return launch_hess_tt<NB16, NHELP, SIGNED, MODE, false, 32>(d, a, grid, st);
}

// canonical link/family pairs have Hessian weights w > 0 (no sign handling in the tensor warp)
static bool weights_positive(int family, int link, bool fast) {
    (void)fast;
    return (family == ANML_FAMILY_BINOMIAL && link == ANML_LINK_EXPIT) || (family == ANML_FAMILY_GAUSSIAN && link == ANML_LINK_IDENTITY) ||
           (family == ANML_FAMILY_POISSON && link == ANML_LINK_EXP);
}

static PriorDev prior_dev(const anml_model* m, int k) {
    PriorDev pd;
    memset(&pd, 0, sizeof(pd));
    const PriorStore& s = m->par[k].prior;
    if (!m->prior_enabled || !s.set) return pd;
    pd.on = 1; pd.p = m->par[k].design->p; pd.off = m->par[k].off; pd.m = s.m;
    pd.mean = s.mean; pd.inv_var = s.inv_var; pd.lin_mat = s.lin_mat; pd.lin_mean = s.lin_mean;
    pd.lin_inv_var = s.lin_inv_var; pd.lin_hess = s.lin_hess;
    return pd;
}

// What the one-launch epilogue (finish.cuh) does with the partial record of a fused block
struct FinishPlan {
    int row_off = 0, col_off = 0;
    bool write_grad = true;
    int obj_mode = 1;        // 0 leave, 1 write (+ prior objectives), 2 accumulate
    bool write_flag = false; // last launch of the evaluation: publish the domain flag
    int prior_k = -1;        // parameter whose prior gradient / Hessian terms ride on this block (-1: none)
    bool prior_obj = false;  // add every parameter's prior objective (with obj_mode 1)
};

static int launch_finish(anml_model* m, int grid, int nb16, bool hess, int p, const FinishPlan& plan, double* packed, cudaStream_t st) {
    FinishArgs f;
    memset(&f, 0, sizeof(f));
    f.partials = m->d_partials; f.ncta = grid; f.el = fused_el(nb16, hess);
    f.nb16 = nb16; f.hess = hess ? 1 : 0; f.p = p;
    f.packed = packed; f.P_total = m->P; f.row_off = plan.row_off; f.col_off = plan.col_off;
    f.write_grad = plan.write_grad ? 1 : 0; f.obj_mode = plan.obj_mode; f.write_flag = plan.write_flag ? 1 : 0;
    f.beta = m->d_beta; f.status = m->d_status;
    int max_m = 0;
    if (plan.prior_k >= 0) {
        f.prior = prior_dev(m, plan.prior_k);
        if (f.prior.on) max_m = f.prior.m;
    }
    if (plan.prior_obj && plan.obj_mode == 1) {
        for (int k = 0; k < m->K; ++k) {
            f.prior_obj[k] = prior_dev(m, k);
            if (f.prior_obj[k].on && f.prior_obj[k].m > max_m) max_m = f.prior_obj[k].m;
        }
        f.n_prior_obj = m->K;
    }
    const size_t smem = (size_t)(max_m > 0 ? max_m : 1) * sizeof(double);
    if (smem > 40 * 1024) return fail(ANML_EUNSUPPORTED, "more than 5120 linear prior rows on one parameter");
    finish_fused_kernel<<<f.el / 32, 128, smem, st>>>(f);
    CU_TRY(cudaGetLastError());
    m->all_launches += 1;
    return ANML_OK;
}

// Two-parameter model over ONE design with p <= 64, gradient-only evaluation: a single streaming pass
// computes both linear predictors, the row terms (r0, r1, ... -> device vectors) and the objective.
// (With the Hessian wanted, fused_hess_kernel MODE 3 does this inside the first Hessian launch.)
static int run_dual_terms(anml_model* m, double* packed, cudaStream_t st) {
    const anml_design* d = m->par[0].design;
    const int nb16 = (d->p + 15) / 16;
    const long long nblk = (d->n + 31) / 32;
    int grid = (int)((nblk + 7) / 8);
    if (grid > m->sm_count) grid = m->sm_count;
    if (grid < 1) grid = 1;
    FusedArgs a;
    memset(&a, 0, sizeof(a));
    a.n = m->n; a.p = d->p; a.link = m->par[0].link; a.link1 = m->par[1].link; a.family = m->family; a.fast = m->no_fast ? 0 : 1;
    a.beta = m->d_beta + m->par[0].pad_off; a.beta1 = m->d_beta + m->par[1].pad_off;
    a.obs = m->obs; a.offset = m->par[0].offset; a.offset1 = m->par[1].offset; a.weights = m->weights;
    a.out_r0 = m->vec[2]; a.out_r1 = m->vec[3]; a.out_w00 = m->vec[4]; a.out_w01 = m->vec[5]; a.out_w11 = m->vec[6];
    a.partials = m->d_partials; a.status = m->d_status;
    int rc = ANML_OK;
    switch (nb16) {
        case 1: rc = launch_fused_t<1, false, 2>(d, a, grid, st); break;
        case 2: rc = launch_fused_t<2, false, 2>(d, a, grid, st); break;
        case 3: rc = launch_fused_t<3, false, 2>(d, a, grid, st); break;
        default: rc = launch_fused_t<4, false, 2>(d, a, grid, st); break;
    }
    TRY(rc);
    m->all_launches += 1;
    FinishPlan plan;
    plan.write_grad = false; plan.obj_mode = 1; plan.prior_obj = true;  // objective (+ prior objectives) only
    return launch_finish(m, grid, nb16, false, d->p, plan, packed, st);
}

// fused kernel + one-launch epilogue for one parameter block.
// mode 0: terms from X, beta, obs; 1: (r, w) from a.rvec / a.wvec; 3: dual-predictor terms + block (0, 0)
static int run_fused_block(anml_model* m, const anml_design* d, FusedArgs a, bool hess, int mode, const FinishPlan& plan, double* packed,
                           bool timed, cudaStream_t st) {
    const int nb16 = (d->p + 15) / 16;
    if (nb16 > 4) return fail(ANML_EUNSUPPORTED, "fused path handles at most 64 columns");
    const long long nblk = (d->n + 31) / 32;
    const bool specialised = hess && !m->no_specialise && mode != 2;
    if (mode == 3 && !specialised) return fail(ANML_ESTATE, "dual-predictor mode needs the warp-specialised Hessian kernel");
    int grid = (int)((nblk + (specialised ? 3 : 7)) / (specialised ? 4 : 8));
    if (grid > m->sm_count) grid = m->sm_count;
    if (grid < 1) grid = 1;
    a.partials = m->d_partials;
    a.status = m->d_status;
    EventPair* ep = timed ? next_events(m) : nullptr;
    if (ep) CU_TRY(cudaEventRecord(ep->a, st));
    int rc = ANML_OK;
#define FUSED_CASE(NB)                                                                         \
    case NB:                                                                                   \
        if (mode == 0) rc = hess ? launch_fused_t<NB, true, 0>(d, a, grid, st) : launch_fused_t<NB, false, 0>(d, a, grid, st); \
        else rc = hess ? launch_fused_t<NB, true, 1>(d, a, grid, st) : launch_fused_t<NB, false, 1>(d, a, grid, st);           \
        break;
    if (specialised) {
        const bool sgn = !weights_positive(a.family, a.link, a.fast != 0) || m->force_signed;
// helper warps per team: 2 at full width (registers: setmaxnreg 216 / 144).  Designs of at most 16 columns
// are paced by the helpers' per-row work (3 accumulator tiles keep the tensor warp idle most of the time):
// p = 16, n = 5e7: 2 helpers 1.70 ms, 3 helpers 1.41, 4 helpers 1.27 (profiles/r2_ab_narrow.txt).  At 17-32
// columns more helpers lose (p = 22: 2.92 -> 2.97 -> 3.65 ms): the tensor warp sets the pace there.
#ifndef ANML_NARROW_NHELP
#define ANML_NARROW_NHELP 4
#endif
#define HESS_CASE(NB)                                                                              \
    case NB: {                                                                                     \
        constexpr int NH = (NB == 1) ? ANML_NARROW_NHELP : 2;                                      \
        if (mode == 3) rc = launch_hess_t<NB, NH, true, 3>(d, a, grid, st);                        \
        else if (mode == 1) rc = launch_hess_t<NB, NH, true, 1>(d, a, grid, st);                   \
        else rc = sgn ? launch_hess_t<NB, NH, true, 0>(d, a, grid, st) : launch_hess_t<NB, NH, false, 0>(d, a, grid, st); \
        break;                                                                                     \
    }
        switch (nb16) {
            HESS_CASE(1)
            HESS_CASE(2)
            HESS_CASE(3)
            HESS_CASE(4)
        }
#undef HESS_CASE
    } else
    switch (nb16) {
        FUSED_CASE(1)
        FUSED_CASE(2)
        FUSED_CASE(3)
        FUSED_CASE(4)
    }
#undef FUSED_CASE
    TRY(rc);
    if (ep) CU_TRY(cudaEventRecord(ep->b, st));
    m->main_kernel = g_kernel;
    m->last_kernels += (m->last_kernels.empty() ? "" : ";") + m->main_kernel;
    m->main_launches += 1;
    m->all_launches += 1;
    return launch_finish(m, grid, nb16, hess, d->p, plan, packed, st);
}

// ---- Hessian on the INT8 tensor cores (fused_hess_i8.cuh) -------------------------------------------------
// One-parameter models whose sqrt(w) has a bound known in advance: canonical logistic (w <= 1/4) and the
// Gaussian identity model (w = 1 / obs_se^2 <= 1 / min(obs_se)^2).  64 columns, dense row pitch, no per-row weights.
// Below this many rows per device the DMMA kernels run: the digit planes bound the ABSOLUTE error of every row's
// contribution (2^-47 of the square of the column's largest scaled value); over many rows these errors average out (1e-14 at
// the sizes this path is for), a few hundred rows of a heavy-tailed column do not give them the chance
// (tests/test_i8_split_cpu.py::test_precision_falls_quadratically_with_a_loose_scale), and a persistent kernel with a
// TMEM flush buys nothing at that size anyway.
constexpr long long I8_MIN_ROWS = 8192;

static bool i8_eligible(const anml_model* m, const anml_design* d, bool hess) {
    if (m->n < I8_MIN_ROWS && !m->i8_force) return false;
    if (!hess || m->no_i8 || m->no_fast || m->no_specialise || m->force_signed || m->force_generic) return false;
    if (m->K != 1 || d->p != 64 || d->ld != 64 || !d->has3d || m->weights) return false;
    if (const char* s = getenv("ANML_B200_HESS_I8"))
        if (s[0] == '0') return false;
    const int link = m->par[0].link;
    if (m->family == ANML_FAMILY_BINOMIAL && link == ANML_LINK_EXPIT) return true;
    if (m->family == ANML_FAMILY_GAUSSIAN && link == ANML_LINK_IDENTITY) return true;  // sqrt(w) = 1 / obs_se <= 1 / min(obs_se)
    return false;
}

// power of two s with sv_max * s <= 1/2
static double i8_scale_for(double sv_max) {
    int e = 0;
    const double f = std::frexp(sv_max, &e);  // sv_max = f 2^e, f in [1/2, 1)
    return std::ldexp(1.0, f == 0.5 ? -e : -(e + 1));
}

// column scales cs_j = 2^-(e_j + 1), 2^(e_j - 1) <= max_i |x_ij| < 2^(e_j): |x cs| < 1/2.  One pass over X, once per
// design (and again after the library rewrote it).  A NaN or infinite column keeps the model on the DMMA path.
static int i8_prepare(anml_model* m, const anml_design* d, cudaStream_t st) {
    anml_model::I8State& s = m->i8;
    if (s.design == d && s.X == d->X && s.n == d->n && s.version == d->version && s.se == m->se && s.se_version == m->se_version) return ANML_OK;
    if (!s.cs) CU_TRY(cudaMalloc(&s.cs, 128 * sizeof(double)));
    if (!s.colmax) CU_TRY(cudaMalloc(&s.colmax, 65 * sizeof(unsigned long long)));
    if (!s.scratch) CU_TRY(cudaMalloc(&s.scratch, (size_t)m->sm_count * 4 * 64 * 64 * sizeof(double)));
    CU_TRY(cudaMemsetAsync(s.colmax, 0, 64 * sizeof(unsigned long long), st));
    colmax64_kernel<<<m->sm_count * 4, 256, 0, st>>>(d->X, d->n, d->ld, s.colmax);
    CU_TRY(cudaGetLastError());
    unsigned long long bits[64];
    CU_TRY(cudaMemcpyAsync(bits, s.colmax, sizeof(bits), cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    double cs[128];
    bool ok = true;
    for (int j = 0; j < 64; ++j) {
        double mx;
        memcpy(&mx, &bits[j], sizeof(double));
        if (!std::isfinite(mx)) ok = false;
        int e = 0;
        if (ok && mx > 0.0) (void)std::frexp(mx, &e);  // mx = f 2^e, f in [1/2, 1)
        cs[j] = (ok && mx > 0.0) ? std::ldexp(1.0, -(e + 1)) : 1.0;
        cs[64 + j] = 1.0 / cs[j];
    }
    CU_TRY(cudaMemcpyAsync(s.cs, cs, sizeof(cs), cudaMemcpyHostToDevice, st));
    CU_TRY(cudaStreamSynchronize(st));
    // bound on sqrt(w): 1/2 for the logistic; 1 / min(obs_se) for the Gaussian (1 without obs_se)
    double sv_max = 0.5;
    if (m->family == ANML_FAMILY_GAUSSIAN) {
        sv_max = 1.0;
        if (m->se) {
            double lo = 0.0, hi = 0.0;
            int nan = 0;
            TRY(anml_vector_minmax(m->device, m->se, m->n, 1, &lo, &hi, &nan));
            if (nan || !(lo > 0.0) || !std::isfinite(hi)) ok = false;
            else sv_max = 1.0 / lo;
        }
    }
    s.sv_scale = 1.0;
    if (ok) {
        s.sv_scale = i8_scale_for(sv_max);  // sv_max sv_scale <= 1/2
        if (!std::isfinite(s.sv_scale) || s.sv_scale == 0.0 || !std::isfinite(1.0 / (s.sv_scale * s.sv_scale))) ok = false;
    }
    s.design = d; s.X = d->X; s.n = d->n; s.version = d->version; s.se = m->se; s.se_version = m->se_version; s.usable = ok;
    m->all_launches += 1;
    return ANML_OK;
}

template <int MODE>
static int launch_i8_t(const anml_design* d, const FusedArgs& a, const I8Extra& x, int grid, cudaStream_t st) {
    using C = I8Cfg;
    static ConfiguredOn configured;
    auto kern = fused_hess_i8_kernel<MODE>;
    if (!configured.has(d->device)) {
        CU_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        configured.set(d->device);
    }
    kern<<<grid, C::THREADS, C::SMEM_BYTES, st>>>(d->tmap3d, a, x);
    CU_TRY(cudaGetLastError());
    if (MODE == 0) snprintf(g_kernel, sizeof(g_kernel), "fused_hess_i8_kernel");
    else snprintf(g_kernel, sizeof(g_kernel), "fused_hess_i8_kernel<%d>", MODE);
    return ANML_OK;
}

// mode 0: one-parameter model; 1: term-fed block (w >= 0 in a.wvec, its maximum in i8.colmax[64])
static int run_fused_i8(anml_model* m, const anml_design* d, FusedArgs a, int mode, double sv_scale, const FinishPlan& plan, double* packed,
                        cudaStream_t st) {
    const long long nblk = (d->n + 31) / 32;
    int grid = nblk < m->sm_count ? (int)nblk : m->sm_count;
    if (grid < 1) grid = 1;
    a.partials = m->d_partials;
    a.status = m->d_status;
    I8Extra x;
    x.cs = m->i8.cs; x.scratch = m->i8.scratch;
    x.sv_scale = sv_scale;  // sqrt(w) sv_scale <= 1/2
    x.out_scale = 1.0 / (x.sv_scale * x.sv_scale);
    x.wmax_bits = mode == 1 ? m->i8.colmax + 64 : nullptr;  // term-fed block: the kernel derives its scale from max w
    EventPair* ep = next_events(m);
    if (ep) CU_TRY(cudaEventRecord(ep->a, st));
    int rc = ANML_OK;
    if (mode == 1) rc = launch_i8_t<1>(d, a, x, grid, st);
    else rc = launch_i8_t<0>(d, a, x, grid, st);
    TRY(rc);
    if (ep) CU_TRY(cudaEventRecord(ep->b, st));
    m->main_kernel = g_kernel;
    m->last_kernels += (m->last_kernels.empty() ? "" : ";") + m->main_kernel;
    m->main_launches += 1;
    m->all_launches += 1;
    return launch_finish(m, grid, 4, true, d->p, plan, packed, st);
}

// Gaussian prior terms as a separate launch: only the paths whose epilogue is not the fused one
// (designs wider than 64 columns, two different designs) still use it.
static int run_priors(anml_model* m, int want, double* packed, cudaStream_t st) {
    if (!m->prior_enabled) return ANML_OK;
    for (int k = 0; k < m->K; ++k) {
        const PriorStore& s = m->par[k].prior;
        if (!s.set) continue;
        PriorArgs a;
        a.p = m->par[k].design->p; a.off = m->par[k].off; a.P_total = m->P; a.m = s.m; a.want = want;
        a.beta = m->d_beta; a.mean = s.mean; a.inv_var = s.inv_var; a.lin_mat = s.lin_mat; a.lin_mean = s.lin_mean;
        a.lin_inv_var = s.lin_inv_var; a.lin_hess = s.lin_hess; a.packed = packed;
        const size_t smem = (size_t)(s.m > 0 ? s.m : 1) * sizeof(double);
        if (smem > 48 * 1024) return fail(ANML_EUNSUPPORTED, "more than 6144 linear prior rows on one parameter");
        add_priors_kernel<<<1, 256, smem, st>>>(a);
        CU_TRY(cudaGetLastError());
        m->all_launches += 1;
    }
    return ANML_OK;
}

static int eval_general(anml_model* m, int want, double* packed, cudaStream_t st);

// H block (row_off, col_off) = Xa^T diag(w) Xb for designs wider than 64 columns
// or two different designs (wide_hessian.cuh).
static int run_wide_hessian(anml_model* m, const anml_design* da, const anml_design* db, const double* w, int row_off, int col_off,
                            double* packed, cudaStream_t st) {
    const int nba = (da->p + 127) / 128, nbb = (db->p + 63) / 64;  // 128-column blocks of Xa, 64-column blocks of Xb
    const bool symmetric = (da == db) && (row_off == col_off);
    // symmetric: row bi of the block grid holds the 2*bi+2 blocks that touch or lie below the diagonal
    const int npairs = symmetric ? nba * (nba + 1) : nba * nbb;
    // Row chunks: about 12 waves of CTAs (the hardware scheduler then balances diagonal blocks, which do
    // less work), with the count nudged so that the last wave is as full as possible; at least 1024 rows each.
    long long S0 = (12LL * m->sm_count + npairs - 1) / npairs;
    const long long s_max = (m->n_pad + 1023) / 1024;
    if (S0 > s_max) S0 = s_max;
    if (S0 < 1) S0 = 1;
    long long S = S0;
    double best_fill = -1.0;
    for (long long c = (S0 > 3 ? S0 - 3 : 1); c <= S0 + 3 && c <= (s_max > 1 ? s_max : 1); ++c) {
        const long long ctas = c * npairs;
        const long long waves = (ctas + m->sm_count - 1) / m->sm_count;
        const double fill = (double)ctas / (double)(waves * m->sm_count);
        if (fill > best_fill + 1e-9) {
            best_fill = fill;
            S = c;
        }
    }
    long long chunk_rows = (m->n_pad + S - 1) / S;
    chunk_rows = round_up_ll(chunk_rows, 16);
    if (chunk_rows < 16) chunk_rows = 16;
    S = (m->n_pad + chunk_rows - 1) / chunk_rows;
    const long long need = S * npairs * (long long)WIDE_TILE_ELEMS;
    if (need > m->wide_len) {
        CU_TRY(cudaStreamSynchronize(st));
        if (m->d_wide) CU_TRY(cudaFree(m->d_wide));
        m->d_wide = nullptr;
        m->wide_len = 0;
        CU_TRY(cudaMalloc(&m->d_wide, (size_t)need * sizeof(double)));
        m->wide_len = need;
    }
    static ConfiguredOn configured;
    if (!configured.has(m->device)) {
        CU_TRY(cudaFuncSetAttribute(wide_hessian_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WIDE_SMEM_BYTES));
        configured.set(m->device);
    }
    WideArgs a;
    a.n_pad = m->n_pad; a.chunk_rows = chunk_rows; a.npairs = npairs; a.nbj = nbb; a.symmetric = symmetric ? 1 : 0;
    a.shared_panels = (da == db) ? 1 : 0; a.w = w; a.partials = m->d_wide;
    EventPair* ep = next_events(m);
    if (ep) CU_TRY(cudaEventRecord(ep->a, st));
    wide_hessian_kernel<<<(unsigned)(S * npairs), WIDE_THREADS, WIDE_SMEM_BYTES, st>>>(da->tmap, db->tmap, a);
    CU_TRY(cudaGetLastError());
    if (ep) CU_TRY(cudaEventRecord(ep->b, st));
    m->main_kernel = "wide_hessian_kernel";
    m->last_kernels += (m->last_kernels.empty() ? "" : ";") + m->main_kernel;
    const long long total = (long long)npairs * WIDE_TILE_ELEMS;
    wide_reduce_kernel<<<grid_for(total, 256, m->sm_count, 8), 256, 0, st>>>(m->d_wide, (int)S, npairs, nbb, symmetric ? 1 : 0, da->p, db->p,
                                                                            packed + 1 + m->P, m->P, row_off, col_off);
    CU_TRY(cudaGetLastError());
    m->main_launches += 1;
    m->all_launches += 2;
    return ANML_OK;
}

// ----------------------------------------------------------------------------
// banded spline evaluator (spline_eval.cuh)
// ----------------------------------------------------------------------------
constexpr int BANDED_CTAS_PER_SM = 6;  // 2-3 resident CTAs per SM; a few slices per slot even out the tail

// per-row vector gathered into the sorted order; two spare (zero) doubles so that the 16-byte loads of the
// last row pair stay inside the allocation
static int gather_sorted(anml_model* m, const double* src, double** dst, cudaStream_t st) {
    if (!src) {
        if (*dst) CU_TRY(cudaFree(*dst));
        *dst = nullptr;
        return ANML_OK;
    }
    if (!*dst) {
        CU_TRY(cudaMalloc(dst, (size_t)(m->n + 2) * sizeof(double)));
        CU_TRY(cudaMemsetAsync(*dst + m->n, 0, 2 * sizeof(double), st));
    }
    gather_rows_kernel<<<grid_for(m->n, 256, m->sm_count, 8), 256, 0, st>>>(*dst, src, m->banded.perm, m->n);
    CU_TRY(cudaGetLastError());
    return ANML_OK;
}

static int eval_banded(anml_model* m, int want, double* packed, cudaStream_t st) {
    anml_model::Banded& b = m->banded;
    if (!m->obs) return fail(ANML_ESTATE, "observations not set");
    if (!b.gathered) {
        TRY(gather_sorted(m, m->obs, &b.obs_s, st));
        TRY(gather_sorted(m, m->se, &b.se_s, st));
        TRY(gather_sorted(m, m->par[0].offset, &b.off_s, st));
        TRY(gather_sorted(m, m->weights, &b.w_s, st));
        b.gathered = true;
    }
    const bool hess = (want & ANML_WANT_HESS) != 0;
    BandedArgs a;
    a.xs = b.xs; a.obs = b.obs_s; a.se = b.se_s; a.offset = b.off_s; a.weights = b.w_s;
    a.n = m->n; a.knots = b.t_dev + b.degree; a.coef = b.coef; a.seg = b.seg; a.nk = b.nk; a.beta = m->d_beta;
    a.family = m->family; a.link = m->par[0].link; a.fast = m->no_fast ? 0 : 1;
    a.rows_per_cta = b.rows_per_cta; a.partials = b.partials; a.status = m->d_status;
    EventPair* ep = next_events(m);
    if (ep) CU_TRY(cudaEventRecord(ep->a, st));
    const bool extras = a.se || a.offset || a.weights;
    const int nv = 2 + (a.se ? 1 : 0) + (a.offset ? 1 : 0) + (a.weights ? 1 : 0);
    const size_t bulk_smem = (size_t)BULK_STAGES * nv * (extras ? BULK_ROWS_EXTRAS : BULK_ROWS_PLAIN) * sizeof(double);
    const bool bulk = !m->banded_registers;
    // the bulk-copy kernel (row vectors staged through shared memory by cp.async.bulk) is the default; the
    // register-prefetch kernel stays selectable ("banded_register_prefetch") for A/B runs
#define BANDED_ONE(KERN, D, H, E)                                                                                   \
    {                                                                                                               \
        if (bulk) {                                                                                                 \
            static ConfiguredOn configured;                                                                         \
            if (!configured.has(m->device)) {                                                                       \
                CU_TRY(cudaFuncSetAttribute(spline_banded_bulk_kernel<D, H, E>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                            (int)(BULK_STAGES * ((E) ? 5 * BULK_ROWS_EXTRAS : 2 * BULK_ROWS_PLAIN) * sizeof(double)))); \
                configured.set(m->device);                                                                          \
            }                                                                                                       \
            spline_banded_bulk_kernel<D, H, E><<<b.grid, 256, bulk_smem, st>>>(a);                                  \
        } else {                                                                                                    \
            spline_banded_eval_kernel<D, H, E><<<b.grid, 256, 0, st>>>(a);                                          \
        }                                                                                                           \
    }
#define BANDED_LAUNCH(D)                                  \
    if (extras) {                                         \
        if (hess) BANDED_ONE(_, D, true, true)            \
        else BANDED_ONE(_, D, false, true)                \
    } else {                                              \
        if (hess) BANDED_ONE(_, D, true, false)           \
        else BANDED_ONE(_, D, false, false)               \
    }
    switch (b.degree) {
        case 1: BANDED_LAUNCH(1) break;
        case 2: BANDED_LAUNCH(2) break;
        default: BANDED_LAUNCH(3) break;
    }
#undef BANDED_LAUNCH
#undef BANDED_ONE
    CU_TRY(cudaGetLastError());
    if (ep) CU_TRY(cudaEventRecord(ep->b, st));
    snprintf(g_kernel, sizeof(g_kernel), "%s<%d, %d, %d>", bulk ? "spline_banded_bulk_kernel" : "spline_banded_eval_kernel", b.degree, hess ? 1 : 0,
             extras ? 1 : 0);
    m->main_kernel = g_kernel;
    m->last_kernels += (m->last_kernels.empty() ? "" : ";") + m->main_kernel;
    banded_reduce_kernel<<<(b.rec + 31) / 32, dim3(32, 32), 0, st>>>(b.partials, b.grid, b.rec, b.reduced);
    CU_TRY(cudaGetLastError());
    BandedFinishArgs f;
    f.reduced = b.reduced; f.coef = b.coef; f.nint = b.nk - 1; f.deg = b.degree; f.hess = hess ? 1 : 0; f.P = m->P;
    f.packed = packed; f.beta = m->d_beta; f.prior = prior_dev(m, 0); f.status = m->d_status;
    const size_t smem = (size_t)(f.prior.on && f.prior.m > 0 ? f.prior.m : 1) * sizeof(double);
    if (smem > 40 * 1024) return fail(ANML_EUNSUPPORTED, "more than 5120 linear prior rows on one parameter");
    banded_finish_kernel<<<(1 + m->P + (hess ? m->P * m->P : 0) + 255) / 256, 256, smem, st>>>(f);
    CU_TRY(cudaGetLastError());
    m->main_launches += 1;
    m->all_launches += 3;
    return ANML_OK;
}

extern "C" int anml_model_enable_banded_spline(anml_model* m, const double* x, int x_on_device, const double* knots, int n_knots, int degree) {
    if (!m || !x || !knots) return fail(ANML_EINVAL, "anml_model_enable_banded_spline: NULL argument");
    if (m->K != 1 || !m->par[0].design) return fail(ANML_ESTATE, "the banded spline evaluator needs a one-parameter model with its design set");
    if (degree < 1 || degree > 3) return fail(ANML_EUNSUPPORTED, "banded spline evaluator: degree %d (supported: 1..3)", degree);
    if (n_knots < 2 || n_knots > SPLINE_MAX_KNOTS) return fail(ANML_EINVAL, "spline needs between 2 and %d knots", SPLINE_MAX_KNOTS);
    const int nb = n_knots - 1 + degree;
    if (m->par[0].design->p != nb) return fail(ANML_EINVAL, "design has %d columns, the spline %d bases", m->par[0].design->p, nb);
    if (m->n <= 0 || m->n > 2147483000LL) return fail(ANML_EUNSUPPORTED, "banded spline evaluator: unsupported row count");
    for (int i = 1; i < n_knots; ++i)
        if (!(knots[i] >= knots[i - 1])) return fail(ANML_EINVAL, "spline knots must be sorted ascending");
    CU_TRY(cudaSetDevice(m->device));
    TRY(model_quiesce(m));
    free_banded(m);
    anml_model::Banded& b = m->banded;
    const long long n = m->n;
    const double* xd = nullptr;
    double* owned = nullptr;
    TRY(vector_on_device(m->device, m->sm_count, x, n, x_on_device, &xd, &owned));
    int32_t* iota = nullptr;
    void* temp = nullptr;
    size_t temp_bytes = 0;
    cudaError_t e = cudaMalloc(&b.xs, (size_t)(n + 2) * sizeof(double));
    if (e == cudaSuccess) e = cudaMemset(b.xs + n, 0, 2 * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&b.perm, (size_t)n * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMalloc(&iota, (size_t)n * sizeof(int32_t));
    if (e == cudaSuccess) {
        iota_kernel<<<grid_for(n, 256, m->sm_count, 8), 256>>>(iota, n);
        e = cudaGetLastError();
    }
    // one-off at attach; CUB's radix sort is library code and off the evaluation path
    if (e == cudaSuccess) e = cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, xd, b.xs, iota, b.perm, (int)n);
    if (e == cudaSuccess) e = cudaMalloc(&temp, temp_bytes ? temp_bytes : 16);
    if (e == cudaSuccess) e = cub::DeviceRadixSort::SortPairs(temp, temp_bytes, xd, b.xs, iota, b.perm, (int)n);
    double ends[2] = {0.0, 0.0};
    if (e == cudaSuccess) e = cudaMemcpy(&ends[0], b.xs, sizeof(double), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(&ends[1], b.xs + (n - 1), sizeof(double), cudaMemcpyDeviceToHost);
    if (iota) cudaFree(iota);
    if (temp) cudaFree(temp);
    if (owned) cudaFree(owned);
    if (e != cudaSuccess) {
        const int code = (e == cudaErrorMemoryAllocation) ? ANML_ENOMEM : ANML_ECUDA;
        (void)cudaGetLastError();
        free_banded(m);
        return fail(code, "anml_model_enable_banded_spline: %s", cudaGetErrorString(e));
    }
    // NaN sorts last, so both ends inside the knot span means every abscissa is (no extrapolation on this path)
    if (!(ends[0] >= knots[0] && ends[1] <= knots[n_knots - 1])) {
        free_banded(m);
        return fail(ANML_EUNSUPPORTED, "banded spline evaluator: abscissae [%g, %g] leave the knot span [%g, %g]", ends[0], ends[1], knots[0],
                    knots[n_knots - 1]);
    }
    const int nt = n_knots + 2 * degree;
    std::vector<double> t(nt);
    for (int i = 0; i < degree; ++i) t[i] = knots[0];
    for (int i = 0; i < n_knots; ++i) t[degree + i] = knots[i];
    for (int i = 0; i < degree; ++i) t[degree + n_knots + i] = knots[n_knots - 1];
    const int nm = (degree + 1) + (2 * degree + 1);  // gradient + Hessian power moments per interval
    b.nk = n_knots; b.degree = degree; b.nb = nb; b.rec = (n_knots - 1) * nm + 1;
    // equal, contiguous, even-sized row slices; never more CTAs than 512-row slices
    long long grid = (long long)m->sm_count * BANDED_CTAS_PER_SM;
    const long long max_grid = (n + 511) / 512;
    if (grid > max_grid) grid = max_grid;
    if (grid < 1) grid = 1;
    long long rpc = (n + grid - 1) / grid;
    rpc += rpc & 1;
    b.rows_per_cta = rpc;
    b.grid = (int)((n + rpc - 1) / rpc);
    e = cudaMalloc(&b.t_dev, nt * sizeof(double));
    if (e == cudaSuccess) e = cudaMemcpy(b.t_dev, t.data(), nt * sizeof(double), cudaMemcpyHostToDevice);
    const int ncoef = (n_knots - 1) * (degree + 1) * (degree + 1);
    if (e == cudaSuccess) e = cudaMalloc(&b.coef, (size_t)ncoef * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&b.seg, (size_t)n_knots * sizeof(long long));
    if (e == cudaSuccess) {
        const int work = (n_knots - 1) * (degree + 1);
        switch (degree) {
            case 1: banded_coeff_kernel<1><<<(work + 127) / 128, 128>>>(b.t_dev, n_knots, b.coef); break;
            case 2: banded_coeff_kernel<2><<<(work + 127) / 128, 128>>>(b.t_dev, n_knots, b.coef); break;
            default: banded_coeff_kernel<3><<<(work + 127) / 128, 128>>>(b.t_dev, n_knots, b.coef); break;
        }
        banded_segments_kernel<<<(n_knots + 127) / 128, 128>>>(b.xs, n, b.t_dev + degree, n_knots, b.seg);
        e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
    }
    if (e == cudaSuccess) e = cudaMalloc(&b.partials, (size_t)b.grid * b.rec * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&b.reduced, (size_t)b.rec * sizeof(double));
    if (e != cudaSuccess) {
        const int code = (e == cudaErrorMemoryAllocation) ? ANML_ENOMEM : ANML_ECUDA;
        (void)cudaGetLastError();
        free_banded(m);
        return fail(code, "anml_model_enable_banded_spline: %s", cudaGetErrorString(e));
    }
    b.gathered = false;
    b.enabled = true;
    return ANML_OK;
}

static int eval_on_stream(anml_model* m, const double* x_host, int want, double* packed, cudaStream_t st) {
    TRY(ensure_buffers(m));
    for (int k = 0; k < m->K; ++k)
        if (m->par[k].design->n != m->n) return fail(ANML_ESTATE, "design of parameter %d changed shape", k);
    if (!packed) packed = m->d_packed;
    const bool hess = (want & ANML_WANT_HESS) != 0;
    const bool banded = m->banded.enabled && !m->force_generic;
    if (!banded)  // a deferred (spline) design is built the first time something streams it
        for (int k = 0; k < m->K; ++k) TRY(design_materialize(const_cast<anml_design*>(m->par[k].design)));
    // work of the previous evaluation may still be running on another stream (anml_model_eval_async)
    if (m->eval_pending && m->last_stream != st) CU_TRY(cudaStreamWaitEvent(st, m->eval_done, 0));
    // stage beta: [x (P) | zero-padded copy per parameter]
    if (m->beta_pending) {
        CU_TRY(cudaEventSynchronize(m->beta_copied));
        m->beta_pending = false;
    }
    memset(m->h_beta, 0, m->beta_len * sizeof(double));
    memcpy(m->h_beta, x_host, m->P * sizeof(double));
    for (int k = 0; k < m->K; ++k) memcpy(m->h_beta + m->par[k].pad_off, x_host + m->par[k].off, m->par[k].design->p * sizeof(double));
    CU_TRY(cudaMemcpyAsync(m->d_beta, m->h_beta, m->beta_len * sizeof(double), cudaMemcpyHostToDevice, st));
    CU_TRY(cudaEventRecord(m->beta_copied, st));
    m->beta_pending = true;
    // d_status is zero here: zeroed at allocation, and every evaluation's last launch re-arms it

    const anml_design* d0 = m->par[0].design;
    m->last_kernels.clear();
    if (banded) {
        TRY(eval_banded(m, want, packed, st));
    } else if (m->K == 1 && d0->p <= 64 && !m->force_generic) {
        FusedArgs a;
        memset(&a, 0, sizeof(a));
        a.n = m->n; a.p = d0->p; a.link = m->par[0].link; a.family = m->family; a.fast = m->no_fast ? 0 : 1;
        a.beta = m->d_beta + m->par[0].pad_off; a.obs = m->obs; a.se = m->se; a.offset = m->par[0].offset; a.weights = m->weights;
        FinishPlan plan;
        plan.write_flag = true; plan.prior_k = 0; plan.prior_obj = true;
        bool i8 = i8_eligible(m, d0, hess);
        if (i8) {
            TRY(i8_prepare(m, d0, st));
            i8 = m->i8.usable;
        }
        if (i8) TRY(run_fused_i8(m, d0, a, 0, m->i8.sv_scale, plan, packed, st));
        else TRY(run_fused_block(m, d0, a, hess, 0, plan, packed, true, st));
    } else {
        TRY(eval_general(m, want, packed, st));
    }
    CU_TRY(cudaEventRecord(m->eval_done, st));
    m->eval_pending = true;
    m->last_stream = st;
    return ANML_OK;
}

// General path: any K <= 2, any width.
//   * two parameters over ONE design of at most 64 columns (BASELINE config #4): fused_hess MODE 3 (both
//     predictors, row terms, objective, g0, H00) -> MODE 1 (r1, w11 -> g1, H11) -> MODE 1 (w01 -> H01):
//     X is read three times, priors and the domain flag ride on the three one-launch epilogues;
//   * otherwise linear predictors -> row terms -> per-block gradient / Hessian kernels.
static int eval_general(anml_model* m, int want, double* packed, cudaStream_t st) {
    const bool hess = (want & ANML_WANT_HESS) != 0;
    const int K = m->K;
    for (int k = 0; k < K; ++k) {
        TRY(ensure_vec(m, k));      // y_k
        TRY(ensure_vec(m, 2 + k));  // r_k
    }
    TRY(ensure_vec(m, 4));
    if (K == 2) {
        TRY(ensure_vec(m, 5));
        TRY(ensure_vec(m, 6));
    }
    const bool dual = (K == 2) && (m->par[0].design == m->par[1].design) && (m->par[0].design->p <= 64) && !m->force_generic;
    if (dual) {
        const anml_design* d = m->par[0].design;
        const int off0 = m->par[0].off, off1 = m->par[1].off;
        if (hess && !m->no_specialise) {
            FusedArgs a;
            memset(&a, 0, sizeof(a));
            a.n = m->n; a.p = d->p; a.link = m->par[0].link; a.link1 = m->par[1].link; a.family = m->family; a.fast = m->no_fast ? 0 : 1;
            a.beta = m->d_beta + m->par[0].pad_off; a.beta1 = m->d_beta + m->par[1].pad_off;
            a.obs = m->obs; a.offset = m->par[0].offset; a.offset1 = m->par[1].offset; a.weights = m->weights;
            a.out_r1 = m->vec[3]; a.out_w01 = m->vec[5]; a.out_w11 = m->vec[6];
            FinishPlan p0;
            p0.row_off = off0; p0.col_off = off0; p0.prior_k = 0; p0.prior_obj = true;
            TRY(run_fused_block(m, d, a, true, 3, p0, packed, true, st));
            // Block (1,1) = X^T diag(w11) X, w11 = res^2 / (2 v) >= 0 for the identity / exp pair: on the INT8 tensor cores
            // (fused_hess_i8.cuh MODE 1).  Its digit scale needs max w11, which the launch above has just left in vec[6]:
            // one 0.8 GB read, and the kernel takes the scale from device memory.  Block (0,0) stays on DMMA: sqrt(w00) =
            // e^{-y1/2} has no bound that is known before the pass that computes it, and a loose a-priori bound costs
            // precision quadratically (profiles/r2_ab_i8.txt).
            const bool i8 = !m->no_i8 && (m->n >= I8_MIN_ROWS || m->i8_force) && !m->no_fast && !m->force_signed && !m->weights && d->p == 64 &&
                            d->ld == 64 && d->has3d &&
                            m->par[0].link == ANML_LINK_IDENTITY && m->par[1].link == ANML_LINK_EXP &&
                            !(getenv("ANML_B200_HESS_I8") && getenv("ANML_B200_HESS_I8")[0] == '0');
            if (i8) {
                TRY(i8_prepare(m, d, st));
                if (m->i8.usable) {
                    CU_TRY(cudaMemsetAsync(m->i8.colmax + 64, 0, sizeof(unsigned long long), st));
                    wmax_kernel<<<m->sm_count * 8, 256, 0, st>>>(m->vec[6], m->n, m->i8.colmax + 64);
                    CU_TRY(cudaGetLastError());
                    m->all_launches += 1;
                    FusedArgs b;
                    memset(&b, 0, sizeof(b));
                    b.n = m->n; b.p = d->p; b.rvec = m->vec[3]; b.wvec = m->vec[6];
                    FinishPlan p1;
                    p1.row_off = off1; p1.col_off = off1; p1.prior_k = 1; p1.obj_mode = 0;
                    TRY(run_fused_i8(m, d, b, 1, 1.0, p1, packed, st));
                    FusedArgs c;
                    memset(&c, 0, sizeof(c));
                    c.n = m->n; c.p = d->p; c.rvec = nullptr; c.wvec = m->vec[5];
                    FinishPlan px;
                    px.row_off = off0; px.col_off = off1; px.write_grad = false; px.obj_mode = 0; px.write_flag = true;
                    TRY(run_fused_block(m, d, c, true, 1, px, packed, true, st));
                    return ANML_OK;
                }
            }
        } else {
            TRY(run_dual_terms(m, packed, st));  // objective + (r0, r1, w00, w01, w11)
            FusedArgs a;
            memset(&a, 0, sizeof(a));
            a.n = m->n; a.p = d->p; a.rvec = m->vec[2]; a.wvec = m->vec[4];
            FinishPlan p0;
            p0.row_off = off0; p0.col_off = off0; p0.prior_k = 0; p0.obj_mode = 0;
            TRY(run_fused_block(m, d, a, hess, 1, p0, packed, true, st));
        }
        {
            FusedArgs a;
            memset(&a, 0, sizeof(a));
            a.n = m->n; a.p = d->p; a.rvec = m->vec[3]; a.wvec = m->vec[6];
            FinishPlan p1;
            p1.row_off = off1; p1.col_off = off1; p1.prior_k = 1; p1.obj_mode = 0; p1.write_flag = !hess;
            TRY(run_fused_block(m, d, a, hess, 1, p1, packed, true, st));
        }
        if (hess) {
            FusedArgs a;
            memset(&a, 0, sizeof(a));
            a.n = m->n; a.p = d->p; a.rvec = nullptr; a.wvec = m->vec[5];
            FinishPlan px;
            px.row_off = off0; px.col_off = off1; px.write_grad = false; px.obj_mode = 0; px.write_flag = true;
            TRY(run_fused_block(m, d, a, true, 1, px, packed, true, st));
        }
        return ANML_OK;
    }

    for (int k = 0; k < K; ++k) {
        TRY(launch_linpred(m->par[k].design, m->d_beta + m->par[k].off, m->par[k].offset, 0, -1, m->vec[k], m->d_status, st));
        m->all_launches += 1;
    }
    RowTermsArgs ra;
    ra.n = m->n; ra.family = m->family; ra.nparams = K; ra.link0 = m->par[0].link; ra.link1 = (K == 2) ? m->par[1].link : 0;
    ra.fast = m->no_fast ? 0 : 1;
    ra.y0 = m->vec[0]; ra.y1 = m->vec[1]; ra.obs = m->obs; ra.se = m->se; ra.weights = m->weights;
    ra.r0 = m->vec[2]; ra.r1 = m->vec[3]; ra.w00 = m->vec[4]; ra.w01 = m->vec[5]; ra.w11 = m->vec[6];
    ra.obj_partials = m->d_partials; ra.status = m->d_status;
    const int rgrid = grid_for(m->n, 256, m->sm_count, 8);
    rowterms_kernel<<<rgrid, 256, 0, st>>>(ra);
    CU_TRY(cudaGetLastError());
    sum_partials_kernel<<<1, 256, 0, st>>>(m->d_partials, rgrid, packed, 0);
    CU_TRY(cudaGetLastError());
    m->all_launches += 2;

    for (int k = 0; k < K; ++k) {
        const anml_design* d = m->par[k].design;
        const double* wkk = (k == 0) ? m->vec[4] : m->vec[6];
        if (d->p <= 64) {
            FusedArgs a;
            memset(&a, 0, sizeof(a));
            a.n = m->n; a.p = d->p; a.rvec = m->vec[2 + k]; a.wvec = wkk;
            FinishPlan plan;  // priors are added by run_priors below on this path
            plan.row_off = m->par[k].off; plan.col_off = m->par[k].off; plan.obj_mode = 0;
            TRY(run_fused_block(m, d, a, hess, 1, plan, packed, true, st));
        } else {
            // gradient: streaming X^T r
            const int nchunks = (int)((d->ld + 63) / 64);
            int grid = (8 * m->sm_count + nchunks - 1) / nchunks;
            if ((long long)grid * 8 > d->n) grid = (int)((d->n + 7) / 8);
            if (grid < 1) grid = 1;
            if ((long long)grid * d->ld > m->partials_len) grid = (int)(m->partials_len / d->ld);
            xtr_kernel<<<dim3(grid, nchunks), 256, 0, st>>>(d->X, d->n, d->ld, m->vec[2 + k], m->d_partials);
            CU_TRY(cudaGetLastError());
            column_sum_kernel<<<(d->p + 255) / 256, 256, 0, st>>>(m->d_partials, grid, d->ld, d->p, packed + 1 + m->par[k].off);
            CU_TRY(cudaGetLastError());
            m->all_launches += 2;
            if (hess) TRY(run_wide_hessian(m, d, d, wkk, m->par[k].off, m->par[k].off, packed, st));
        }
    }
    if (K == 2 && hess) {
        const anml_design* da = m->par[0].design;
        const anml_design* db = m->par[1].design;
        if (da == db && da->p <= 64) {
            FusedArgs a;
            memset(&a, 0, sizeof(a));
            a.n = m->n; a.p = da->p; a.rvec = nullptr; a.wvec = m->vec[5];
            FinishPlan plan;
            plan.row_off = m->par[0].off; plan.col_off = m->par[1].off; plan.write_grad = false; plan.obj_mode = 0;
            TRY(run_fused_block(m, da, a, true, 1, plan, packed, true, st));
        } else {
            TRY(run_wide_hessian(m, da, db, m->vec[5], m->par[0].off, m->par[1].off, packed, st));
        }
    }
    TRY(run_priors(m, want, packed, st));
    set_flag_kernel<<<1, 1, 0, st>>>(packed + 1 + m->P + (size_t)m->P * m->P, m->d_status);
    CU_TRY(cudaGetLastError());
    m->all_launches += 1;
    return ANML_OK;
}

extern "C" int anml_model_eval_async(anml_model* m, const double* x_host, int want, double* packed_dev, void* stream) {
    if (!m || !x_host) return fail(ANML_EINVAL, "anml_model_eval_async: NULL argument");
    CU_TRY(cudaSetDevice(m->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : m->stream;
    // the epilogue kernels write the packed record straight into the caller's buffer
    return eval_on_stream(m, x_host, want, packed_dev, st);
}

extern "C" int anml_model_eval(anml_model* m, const double* x_host, int want, double* obj, double* grad, double* hess) {
    if (!m || !x_host) return fail(ANML_EINVAL, "anml_model_eval: NULL argument");
    if (((want & ANML_WANT_OBJ) && !obj) || ((want & ANML_WANT_GRAD) && !grad) || ((want & ANML_WANT_HESS) && !hess))
        return fail(ANML_EINVAL, "anml_model_eval: output buffer missing for a requested quantity");
    CU_TRY(cudaSetDevice(m->device));
    // the gradient is always produced (it rides on the same pass); the Hessian only on request
    TRY(eval_on_stream(m, x_host, want | ANML_WANT_OBJ | ANML_WANT_GRAD, nullptr, m->stream));
    const size_t P = m->P;
    const size_t flag_at = 1 + P + P * P;
    if (want & ANML_WANT_HESS) {
        CU_TRY(cudaMemcpyAsync(m->h_packed, m->d_packed, (flag_at + 1) * sizeof(double), cudaMemcpyDeviceToHost, m->stream));
    } else {
        CU_TRY(cudaMemcpyAsync(m->h_packed, m->d_packed, (1 + P) * sizeof(double), cudaMemcpyDeviceToHost, m->stream));
        CU_TRY(cudaMemcpyAsync(m->h_packed + flag_at, m->d_packed + flag_at, sizeof(double), cudaMemcpyDeviceToHost, m->stream));
    }
    CU_TRY(cudaStreamSynchronize(m->stream));
    m->eval_pending = false;
    if (m->h_packed[flag_at] >= 1024.0) {
        m->i8.design = nullptr;  // rescan the column maxima at the next evaluation
        return fail(ANML_ESTATE, "design or obs_se values outside the ranges recorded for the INT8 Hessian path (was a borrowed device array rewritten in "
                                 "place?); the ranges are taken again at the next evaluation");
    }
    if (m->h_packed[flag_at] != 0.0) return fail(ANML_EDOMAIN, "linear predictor outside the domain of a Log/Logit mapping");
    if (want & ANML_WANT_OBJ) *obj = m->h_packed[0];
    if (want & ANML_WANT_GRAD) memcpy(grad, m->h_packed + 1, P * sizeof(double));
    if (want & ANML_WANT_HESS) memcpy(hess, m->h_packed + 1 + P, P * P * sizeof(double));
    return ANML_OK;
}

extern "C" int anml_model_last_kernels(const anml_model* m, char* out, int out_len) {
    if (!m || !out || out_len <= 0) return fail(ANML_EINVAL, "anml_model_last_kernels: bad arguments");
    snprintf(out, (size_t)out_len, "%s", m->last_kernels.c_str());
    return ANML_OK;
}

extern "C" int anml_model_main_kernel(const anml_model* m, char* out, int out_len) {
    if (!m || !out || out_len <= 0) return fail(ANML_EINVAL, "anml_model_main_kernel: bad arguments");
    snprintf(out, (size_t)out_len, "%s", m->main_kernel.c_str());
    return ANML_OK;
}

// ----------------------------------------------------------------------------
// peer-memory all-reduce of the packed record (comm.cuh)
// ----------------------------------------------------------------------------
struct anml_comm {
    int device = 0, rank = 0, world = 1, sm_count = 0;
    long long max_elems = 0;
    unsigned int epoch = 0;
    unsigned char* window = nullptr;                    // local, cudaMalloc'ed, exported
    unsigned char* peer[COMM_MAX_WORLD] = {};           // mapped windows, peer[rank] == window
    bool connected = false;
    long long timeout_cycles = 0;
};

extern "C" int anml_comm_handle_bytes(int* out) {
    if (!out) return fail(ANML_EINVAL, "NULL argument");
    *out = (int)sizeof(cudaIpcMemHandle_t);
    return ANML_OK;
}

extern "C" int anml_comm_create(int device, int rank, int world, int64_t max_elems, anml_comm** out) {
    if (!out || world < 1 || world > COMM_MAX_WORLD || rank < 0 || rank >= world || max_elems <= 0)
        return fail(ANML_EINVAL, "anml_comm_create: need 0 <= rank < world <= %d and max_elems > 0", COMM_MAX_WORLD);
    if (max_elems > (1LL << 22)) return fail(ANML_EUNSUPPORTED, "anml_comm_create: the peer window is meant for latency-bound records (at most 2^22 doubles); use NCCL for larger ones");
    CU_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU_TRY(cudaGetDeviceProperties(&prop, device));
    anml_comm* c = new (std::nothrow) anml_comm();
    if (!c) return fail(ANML_ENOMEM, "out of host memory");
    c->device = device; c->rank = rank; c->world = world; c->max_elems = max_elems; c->sm_count = prop.multiProcessorCount;
    c->timeout_cycles = (long long)(20.0 * 1e3 * (double)prop.clockRate);  // ~20 s of SM clock (clockRate is in kHz)
    if (const char* s = getenv("ANML_B200_COMM_TIMEOUT_S")) {
        const double v = atof(s);
        if (v > 0.0) c->timeout_cycles = (long long)(v * 1e3 * (double)prop.clockRate);
    }
    // header | slot[2 parities][world sources][max_elems] of 16-byte entries
    const size_t bytes = (size_t)COMM_HEADER_BYTES + (size_t)2 * world * max_elems * 16;
    cudaError_t e = cudaMalloc(&c->window, bytes);
    if (e == cudaSuccess) e = cudaMemset(c->window, 0, bytes);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        if (c->window) cudaFree(c->window);
        delete c;
        return fail(e == cudaErrorMemoryAllocation ? ANML_ENOMEM : ANML_ECUDA, "anml_comm_create: %s", cudaGetErrorString(e));
    }
    c->peer[rank] = c->window;
    *out = c;
    return ANML_OK;
}

extern "C" int anml_comm_export(anml_comm* c, void* handle_out) {
    if (!c || !handle_out) return fail(ANML_EINVAL, "NULL argument");
    CU_TRY(cudaSetDevice(c->device));
    cudaIpcMemHandle_t h;
    CU_TRY(cudaIpcGetMemHandle(&h, c->window));
    memcpy(handle_out, &h, sizeof(h));
    return ANML_OK;
}

extern "C" int anml_comm_connect(anml_comm* c, const void* all_handles) {
    if (!c || !all_handles) return fail(ANML_EINVAL, "NULL argument");
    if (c->connected) return fail(ANML_ESTATE, "anml_comm_connect: already connected");
    CU_TRY(cudaSetDevice(c->device));
    const char* hs = static_cast<const char*>(all_handles);
    for (int j = 0; j < c->world; ++j) {
        if (j == c->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, hs + (size_t)j * sizeof(h), sizeof(h));
        void* ptr = nullptr;
        const cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            (void)cudaGetLastError();
            for (int q = 0; q < j; ++q)
                if (q != c->rank && c->peer[q]) {
                    cudaIpcCloseMemHandle(c->peer[q]);
                    c->peer[q] = nullptr;
                }
            return fail(ANML_ECUDA, "cudaIpcOpenMemHandle for rank %d failed: %s", j, cudaGetErrorString(e));
        }
        c->peer[j] = static_cast<unsigned char*>(ptr);
    }
    c->connected = true;
    return ANML_OK;
}

extern "C" int anml_comm_allreduce_async(anml_comm* c, double* buf_dev, int64_t n_elems, void* stream) {
    if (!c || !buf_dev) return fail(ANML_EINVAL, "NULL argument");
    if (!c->connected && c->world > 1) return fail(ANML_ESTATE, "anml_comm_allreduce_async: not connected");
    if (n_elems <= 0 || n_elems > c->max_elems) return fail(ANML_EINVAL, "anml_comm_allreduce_async: %lld elements, capacity %lld", (long long)n_elems, c->max_elems);
    CU_TRY(cudaSetDevice(c->device));
    CommArgs a;
    memset(&a, 0, sizeof(a));
    a.buf = buf_dev; a.n = n_elems; a.max_elems = c->max_elems; a.rank = c->rank; a.world = c->world;
    if (++c->epoch == 0) c->epoch = 2;  // never 0 (the window's initial state); keep the parity sequence alternating
    a.epoch = c->epoch; a.local = c->window; a.timeout_cycles = c->timeout_cycles;
    for (int j = 0; j < c->world; ++j) a.peer[j] = c->peer[j];
    long long grid = (n_elems + 255) / 256;  // one element per thread up to two waves, then a grid-stride loop
    if (grid > 2LL * c->sm_count) grid = 2LL * c->sm_count;
    if (grid < 1) grid = 1;
    comm_allreduce_kernel<<<(int)grid, 256, 0, (cudaStream_t)stream>>>(a);
    CU_TRY(cudaGetLastError());
    return ANML_OK;
}

extern "C" int anml_comm_error(anml_comm* c, int64_t* failed_epoch) {
    if (!c || !failed_epoch) return fail(ANML_EINVAL, "NULL argument");
    CU_TRY(cudaSetDevice(c->device));
    unsigned long long v = 0;
    CU_TRY(cudaMemcpy(&v, c->window, sizeof(v), cudaMemcpyDeviceToHost));
    *failed_epoch = (int64_t)v;
    return ANML_OK;
}

extern "C" int anml_comm_destroy(anml_comm* c) {
    if (!c) return ANML_OK;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    for (int j = 0; j < c->world; ++j)
        if (j != c->rank && c->peer[j]) cudaIpcCloseMemHandle(c->peer[j]);
    if (c->window) cudaFree(c->window);
    (void)cudaGetLastError();
    delete c;
    return ANML_OK;
}

extern "C" int anml_model_kernel_time(anml_model* m, int reset, double* total_ms, int64_t* launches, int64_t* all_launches) {
    if (!m) return fail(ANML_EINVAL, "NULL model");
    CU_TRY(cudaSetDevice(m->device));
    for (size_t i = 0; i < m->ev_used; ++i) {
        CU_TRY(cudaEventSynchronize(m->ev_pool[i].b));
        float ms = 0.f;
        CU_TRY(cudaEventElapsedTime(&ms, m->ev_pool[i].a, m->ev_pool[i].b));
        m->ev_accum_ms += ms;
    }
    m->ev_used = 0;
    if (total_ms) *total_ms = m->ev_accum_ms;
    if (launches) *launches = m->main_launches;
    if (all_launches) *all_launches = m->all_launches;
    if (reset) {
        m->ev_accum_ms = 0.0;
        m->main_launches = 0;
        m->all_launches = 0;
    }
    return ANML_OK;
}
