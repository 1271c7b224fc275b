// Pipelined host -> device upload: the ingest step in front of the path
// (replaces Component.attach + np.hstack, data/component.py:84-91,
// parameter/main.py:160-162).
//
// A plain cudaMemcpy from pageable memory goes through one driver-side staging thread
// (6-8 GB/s measured on this pool's boxes).  Here a few host threads each own a *lane*
// (16 MiB pinned buffer + 16 MiB device buffer + stream): a thread gathers its chunk of
// rows into pinned memory, queues the DMA and the column scatter on its stream, and moves
// on to its next chunk while the copy engine works.  The lanes never share a buffer, so
// the only synchronisation is each thread's cudaStreamSynchronize on its own stream.
#pragma once
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "generic.cuh"

namespace anml {

struct UploadLane {
    double* pinned = nullptr;
    double* dev = nullptr;
    cudaStream_t stream = nullptr;
};

struct UploadPool {
    std::mutex mu;  // one upload at a time per device
    std::vector<UploadLane> lanes;
};

constexpr long long UPLOAD_LANE_ELEMS = 1LL << 21;      // 16 MiB of doubles per lane
constexpr long long UPLOAD_INLINE_ELEMS = 1LL << 18;    // up to 2 MiB: calling thread only
constexpr int UPLOAD_MAX_DEVICES = 64;

static UploadPool g_upload_pools[UPLOAD_MAX_DEVICES];

static int upload_thread_count() {
    if (const char* s = getenv("ANML_B200_UPLOAD_THREADS")) {
        const int v = atoi(s);
        if (v >= 1) return v > 32 ? 32 : v;
    }
    const unsigned hc = std::thread::hardware_concurrency();
    int t = hc ? (int)(hc / 2) : 4;
    return t < 1 ? 1 : (t > 8 ? 8 : t);
}

static void upload_pool_release(UploadPool& pool) {
    for (auto& l : pool.lanes) {
        if (l.pinned) cudaFreeHost(l.pinned);
        if (l.dev) cudaFree(l.dev);
        if (l.stream) cudaStreamDestroy(l.stream);
    }
    pool.lanes.clear();
}

// caller holds pool.mu and has the device current
static cudaError_t upload_pool_ensure(UploadPool& pool, int lanes_wanted) {
    while ((int)pool.lanes.size() < lanes_wanted) {
        UploadLane l;
        cudaError_t e = cudaHostAlloc(&l.pinned, UPLOAD_LANE_ELEMS * sizeof(double), cudaHostAllocDefault);
        if (e == cudaSuccess) e = cudaMalloc(&l.dev, UPLOAD_LANE_ELEMS * sizeof(double));
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&l.stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) {
            if (l.pinned) cudaFreeHost(l.pinned);
            if (l.dev) cudaFree(l.dev);
            if (l.stream) cudaStreamDestroy(l.stream);
            if (!pool.lanes.empty()) {  // run with the lanes we have
                (void)cudaGetLastError();
                return cudaSuccess;
            }
            return e;
        }
        pool.lanes.push_back(l);
    }
    return cudaSuccess;
}

// Rows [0, n) of a host block (ncols wide, row stride src_ld) -> columns [col0, col0 + ncols) of the
// device matrix dst (row stride dst_ld).  A dense destination (col0 == 0, dst_ld == ncols) is
// written by the DMA itself; anything else goes through the lane's device buffer and a scatter.
struct UploadJob {
    double* dst;
    long long dst_ld;
    int col0;
    long long n;
    int ncols;
    const double* src;
    long long src_ld;
};

static void upload_gather(double* pinned, const UploadJob& j, long long r0, long long rows) {
    const double* s = j.src + r0 * j.src_ld;
    if (j.src_ld == j.ncols) {
        memcpy(pinned, s, (size_t)rows * j.ncols * sizeof(double));
    } else if (j.ncols == 1) {
        for (long long i = 0; i < rows; ++i) pinned[i] = s[i * j.src_ld];
    } else {
        for (long long i = 0; i < rows; ++i) memcpy(pinned + i * j.ncols, s + i * j.src_ld, (size_t)j.ncols * sizeof(double));
    }
}

static cudaError_t upload_run(int device, int sm_count, const UploadJob& job) {
    if (job.n <= 0 || job.ncols <= 0) return cudaSuccess;
    if (device < 0 || device >= UPLOAD_MAX_DEVICES) return cudaErrorInvalidDevice;
    UploadPool& pool = g_upload_pools[device];
    std::lock_guard<std::mutex> guard(pool.mu);
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return e;

    const long long max_rows = UPLOAD_LANE_ELEMS / job.ncols;
    if (max_rows < 1) return cudaErrorInvalidValue;  // wider than a lane: callers never pass that
    const bool small = job.n * job.ncols <= UPLOAD_INLINE_ELEMS;
    int nthreads = small ? 1 : upload_thread_count();
    // about four chunks per thread (dynamic scheduling evens out the tail), 1 MiB .. one lane each
    long long rows_per_chunk = (job.n + 4LL * nthreads - 1) / (4LL * nthreads);
    long long min_rows = (1LL << 17) / job.ncols;
    if (min_rows < 1) min_rows = 1;
    if (rows_per_chunk < min_rows) rows_per_chunk = min_rows;
    if (rows_per_chunk > max_rows) rows_per_chunk = max_rows;
    const long long nchunks = (job.n + rows_per_chunk - 1) / rows_per_chunk;
    if (nthreads > nchunks) nthreads = (int)nchunks;
    e = upload_pool_ensure(pool, nthreads);
    if (e != cudaSuccess) return e;
    if (nthreads > (int)pool.lanes.size()) nthreads = (int)pool.lanes.size();

    // The lanes' streams are non-blocking: they do not order themselves behind the legacy default stream,
    // where the callers' cudaMemset of a fresh destination runs.  Wait for that stream only -- a
    // device-wide synchronisation here would stall unrelated streams (another model evaluating on this
    // GPU); callers that re-upload into a buffer an evaluation may still read quiesce the model first.
    e = cudaStreamSynchronize(cudaStreamLegacy);
    if (e != cudaSuccess) return e;

    const bool dense = job.col0 == 0 && job.dst_ld == job.ncols;
    std::atomic<long long> next(0);
    std::atomic<int> first_error((int)cudaSuccess);
    auto work = [&](int lane_id) {
        UploadLane& lane = pool.lanes[lane_id];
        cudaError_t le = cudaSetDevice(device);
        while (le == cudaSuccess && first_error.load(std::memory_order_relaxed) == (int)cudaSuccess) {
            const long long c = next.fetch_add(1);
            if (c >= nchunks) break;
            const long long r0 = c * rows_per_chunk;
            const long long rows = job.n - r0 < rows_per_chunk ? job.n - r0 : rows_per_chunk;
            le = cudaStreamSynchronize(lane.stream);  // the lane's buffers are free again
            if (le != cudaSuccess) break;
            upload_gather(lane.pinned, job, r0, rows);
            const size_t bytes = (size_t)rows * job.ncols * sizeof(double);
            if (dense) {
                le = cudaMemcpyAsync(job.dst + r0 * job.ncols, lane.pinned, bytes, cudaMemcpyHostToDevice, lane.stream);
            } else {
                le = cudaMemcpyAsync(lane.dev, lane.pinned, bytes, cudaMemcpyHostToDevice, lane.stream);
                if (le == cudaSuccess) {
                    long long blocks = (rows * job.ncols + 255) / 256;
                    const long long cap = (long long)sm_count * 8;
                    if (blocks > cap) blocks = cap;
                    scatter_columns_kernel<<<(int)blocks, 256, 0, lane.stream>>>(job.dst + r0 * job.dst_ld, rows, job.dst_ld, job.col0,
                                                                                job.ncols, lane.dev, job.ncols);
                    le = cudaGetLastError();
                }
            }
        }
        const cudaError_t se = cudaStreamSynchronize(lane.stream);
        if (le == cudaSuccess) le = se;
        if (le != cudaSuccess) {
            int expected = (int)cudaSuccess;
            first_error.compare_exchange_strong(expected, (int)le);
        }
    };
    if (nthreads <= 1) {
        work(0);
    } else {
        std::vector<std::thread> threads;
        threads.reserve(nthreads - 1);
        try {
            for (int t = 1; t < nthreads; ++t) threads.emplace_back(work, t);
        } catch (...) {
            // thread creation refused (resource limits): the lanes that did start plus this thread share the chunks
        }
        work(0);
        for (auto& t : threads) t.join();
    }
    return (cudaError_t)first_error.load();
}

// free the lanes of one device (pinned + device staging buffers, streams); they are rebuilt on demand
static void upload_release_device(int device) {
    if (device < 0 || device >= UPLOAD_MAX_DEVICES) return;
    UploadPool& pool = g_upload_pools[device];
    std::lock_guard<std::mutex> guard(pool.mu);
    if (pool.lanes.empty()) return;
    if (cudaSetDevice(device) != cudaSuccess) {
        (void)cudaGetLastError();
        return;
    }
    upload_pool_release(pool);
}

// contiguous host vector -> contiguous device vector
static cudaError_t upload_vector(int device, int sm_count, double* dst, const double* src, long long n) {
    UploadJob j{dst, 1, 0, n, 1, src, 1};
    return upload_run(device, sm_count, j);
}

}  // namespace anml
