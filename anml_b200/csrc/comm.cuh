// All-reduce of the packed record [obj | grad | hess | flag] of a row-sharded evaluation over PEER
// MEMORY: one launch per evaluation, no NCCL on the evaluation path.
//
// The record is small (33 KB at P = 64, 132 KB at P = 128), so the exchange is pure latency.  The
// protocol is the "low-latency" one: data and flag travel in the SAME store, there is no fence, no
// separate flag round trip and no grid-wide synchronisation.
//
// Every rank owns one cudaMalloc'ed window, exported with cudaIpcGetMemHandle and mapped by every other
// rank of the node (NVLink 5 / NVSwitch peer access):
//     window = [ header 256 B | slot[parity 0..1][source rank 0..world-1][max_elems] ]
// with 16-byte slot entries { lo32(value), epoch, hi32(value), epoch }.  Evaluation number e (1, 2, ...)
// uses parity e & 1.  A thread that owns element i
//   1. packs its local value with e and PUSHES it into slot[e & 1][rank][i] of every peer's window
//      (one 16-byte st.volatile per peer, straight over NVLink; each 8-byte half carries its own flag and
//      is written atomically);
//   2. polls slot[e & 1][j][i] of ITS OWN window (local memory) for every j != rank until both flags
//      read e, and sums the values in rank order 0, 1, .., world-1 (its own value from the register) --
//      every rank adds the same numbers in the same order and ends up with the same bits.
// Two parities suffice: a rank can push evaluation e+1 while a slow peer still reads e, but it cannot
// finish e+1 -- let alone start e+2, which reuses parity e & 1 -- before that peer has pushed e+1, which it
// does only after its own launch e has completed.  No CTA waits on another CTA of its own grid, so any
// grid size is safe.  A peer that never arrives (crashed rank) ends the wait after ~20 s: the element
// becomes NaN and the window's error word is set (anml_comm_error).
#pragma once
#include "common.cuh"

namespace anml {

constexpr int COMM_MAX_WORLD = 16;
constexpr int COMM_HEADER_BYTES = 256;  // [0] error word (u64)

struct CommArgs {
    double* buf;                 // the local record, reduced in place
    long long n;                 // elements
    long long max_elems;         // capacity of one slot
    int rank, world;
    unsigned int epoch;          // 1, 2, ... (wraps after 4e9 evaluations)
    unsigned char* local;        // this rank's window
    unsigned char* peer[COMM_MAX_WORLD];  // every rank's window as mapped here (peer[rank] == local)
    long long timeout_cycles;
};

__device__ __forceinline__ void st_slot(uint4* p, uint4 v) {
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 ld_slot(const uint4* p) {
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(256) comm_allreduce_kernel(const CommArgs a) {
    const size_t par_off = (size_t)(a.epoch & 1u) * a.world * a.max_elems;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    // 1. push
    for (long long i = i0; i < a.n; i += stride) {
        const double v = a.buf[i];
        const uint4 pk = make_uint4((unsigned)__double2loint(v), a.epoch, (unsigned)__double2hiint(v), a.epoch);
#pragma unroll
        for (int j = 0; j < COMM_MAX_WORLD; ++j) {
            if (j < a.world && j != a.rank) {
                uint4* dst = reinterpret_cast<uint4*>(a.peer[j] + COMM_HEADER_BYTES) + par_off + (size_t)a.rank * a.max_elems + i;
                st_slot(dst, pk);
            }
        }
    }
    // 2. poll the local window, sum in rank order
    const uint4* mine = reinterpret_cast<const uint4*>(a.local + COMM_HEADER_BYTES) + par_off;
    const long long t0 = clock64();
    bool timed_out = false;
    for (long long i = i0; i < a.n; i += stride) {
        const double own = a.buf[i];
        uint4 got[COMM_MAX_WORLD];
        unsigned pending = 0;
#pragma unroll
        for (int j = 0; j < COMM_MAX_WORLD; ++j)
            if (j < a.world && j != a.rank) pending |= 1u << j;
        while (pending && !timed_out) {
#pragma unroll
            for (int j = 0; j < COMM_MAX_WORLD; ++j) {
                if (pending & (1u << j)) {
                    got[j] = ld_slot(mine + (size_t)j * a.max_elems + i);
                    if (got[j].y == a.epoch && got[j].w == a.epoch) pending &= ~(1u << j);
                }
            }
            if (pending && clock64() - t0 > a.timeout_cycles) timed_out = true;
        }
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < COMM_MAX_WORLD; ++j) {
            if (j < a.world) s += (j == a.rank) ? own : __hiloint2double((int)got[j].z, (int)got[j].x);
        }
        a.buf[i] = timed_out ? __longlong_as_double(0x7ff8000000000000ll) : s;
    }
    if (timed_out) *reinterpret_cast<unsigned long long*>(a.local) = a.epoch;
}

}  // namespace anml
