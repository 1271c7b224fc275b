// All-reduce of the packed record [obj | grad | hess | flag] of a row-sharded evaluation over PEER
// MEMORY: one launch per evaluation, no NCCL on the evaluation path.
//
// Every rank owns one cudaMalloc'ed window, exported with cudaIpcGetMemHandle and mapped by every
// other rank of the node (NVLink 5 / NVSwitch peer access):
//     window = [ flags: world x u64 | counter u64 | error u64 | pad ][ data[2][max_elems] ]
// Evaluation number e (1, 2, ...) uses data[e & 1]:
//   1. every CTA copies its slice of the local record into the rank's own data[e & 1];
//   2. the last CTA to finish (device-scope counter) makes the copy visible system-wide and writes e
//      into flags[rank] of EVERY rank's window (st.release.sys over NVLink);
//   3. every CTA waits until all `world` flags of its own window have reached e (ld.acquire.sys on
//      local memory), then sums its slice over the ranks 0, 1, .., world-1 -- remote slices are read
//      straight from the peers' windows -- in that fixed order, so every rank ends up with the same
//      bits, and writes the sum back into the record.
// Two data buffers suffice: a rank can start evaluation e+1 (writing data[(e+1) & 1]) while a slow peer
// still reads data[e & 1], but it cannot finish e+1 -- let alone start e+2 -- before that peer has
// published e+1, which it does only after its own launch e has completed.
// The grid is at most one CTA per SM so that all CTAs are co-resident while they wait.
// A peer that never arrives (crashed rank) ends the wait after ~20 s: the record is poisoned with NaN
// and the window's error word is set (anml_comm_error).
#pragma once
#include "common.cuh"

namespace anml {

constexpr int COMM_MAX_WORLD = 16;
constexpr int COMM_HEADER_U64 = 32;  // flags[16] | counter | error | pad -> 256 bytes

struct CommArgs {
    double* buf;                 // the local record, reduced in place
    long long n;                 // elements
    long long max_elems;         // capacity of one data buffer
    int rank, world;
    unsigned long long epoch;    // 1, 2, ...
    unsigned long long* local;   // this rank's window
    unsigned long long* peer[COMM_MAX_WORLD];  // every rank's window as mapped here (peer[rank] == local)
    long long timeout_cycles;
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ double ld_peer(const double* p) {  // L2-coherent load, never served from a stale L1 line
    double v;
    asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(256) comm_allreduce_kernel(const CommArgs a) {
    __shared__ int timed_out;
    const int par = (int)(a.epoch & 1ull);
    double* mine = reinterpret_cast<double*>(a.local + COMM_HEADER_U64) + (size_t)par * a.max_elems;
    const long long per = (a.n + gridDim.x - 1) / gridDim.x;
    const long long lo = (long long)blockIdx.x * per;
    long long hi = lo + per;
    if (hi > a.n) hi = a.n;
    if (threadIdx.x == 0) timed_out = 0;
    // 1. publish the local slice
    for (long long i = lo + threadIdx.x; i < hi; i += blockDim.x) mine[i] = a.buf[i];
    __syncthreads();
    // 2. last CTA of this rank: tell everybody
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned long long done = atomicAdd(a.local + COMM_MAX_WORLD, 1ull);
        if (done == (unsigned long long)gridDim.x - 1) {
            a.local[COMM_MAX_WORLD] = 0;  // every CTA of this launch has arrived; the next launch starts from zero
            __threadfence_system();
            for (int j = 0; j < a.world; ++j) st_release_sys(a.peer[j] + a.rank, a.epoch);
        }
    }
    // 3. wait for every rank's flag in the local window
    if (threadIdx.x < a.world) {
        const long long t0 = clock64();
        while (ld_acquire_sys(a.local + threadIdx.x) < a.epoch) {
            if (clock64() - t0 > a.timeout_cycles) {
                timed_out = 1;
                break;
            }
        }
    }
    __syncthreads();
    if (timed_out) {
        if (threadIdx.x == 0) a.local[COMM_MAX_WORLD + 1] = a.epoch;
        for (long long i = lo + threadIdx.x; i < hi; i += blockDim.x) a.buf[i] = __longlong_as_double(0x7ff8000000000000ll);
        return;
    }
    for (long long i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        double s = 0.0;
        for (int j = 0; j < a.world; ++j) {
            const double* src = reinterpret_cast<const double*>(a.peer[j] + COMM_HEADER_U64) + (size_t)par * a.max_elems;
            s += ld_peer(src + i);
        }
        a.buf[i] = s;
    }
}

}  // namespace anml
