// Microbenchmark: how fast can ONE warp per scheduler stream the Hessian's DMMA
// tile pattern (36 lower-triangle 8x8 tiles per 4-row k-step), and what do the
// B-operand scaling (DMUL), shared-memory fragment loads and a competing FP64
// helper warp on the same scheduler cost?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mstream_bench mstream_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ double2 lds128(unsigned addr) {
    double2 v; asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr)); return v;
}

// MODE bit0: DMUL scaling of B, bit1: fragments from shared memory, HELPERS: extra warps per scheduler running DFMA chains
template <int MODE, int HELPERS, int HCHAIN>
__global__ void __launch_bounds__(128 * (1 + HELPERS), 1) k_stream(double* out, int iters, double seed) {
    __shared__ __align__(16) double sm[4 * 1024];
    for (int i = threadIdx.x; i < 4 * 1024; i += blockDim.x) sm[i] = seed * i;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp < 4) {
        double acc[72];
#pragma unroll
        for (int i = 0; i < 72; i++) acc[i] = 0.0;
        double av[8], bv[8];
#pragma unroll
        for (int i = 0; i < 8; i++) { av[i] = seed + i + lane; bv[i] = seed - i; }
        const unsigned base = (unsigned)__cvta_generic_to_shared(sm) + warp * 8192 + lane * 16;
        double w = seed;
        for (int it = 0; it < iters; it++) {
            if (MODE & 2) {
#pragma unroll
                for (int j = 0; j < 4; j++) { double2 v = lds128(base + j * 2048 + (it & 3) * 512); av[2 * j] = v.x; av[2 * j + 1] = v.y; }
            }
            if (MODE & 1) {
#pragma unroll
                for (int j = 0; j < 8; j++) bv[j] = w * av[j];
            }
#pragma unroll
            for (int I = 0; I < 8; I++)
#pragma unroll
                for (int J = 0; J <= I; J++) { const int t = I * (I + 1) / 2 + J; dmma(acc[2 * t], acc[2 * t + 1], av[I], bv[J]); }
        }
        double s = 0;
#pragma unroll
        for (int i = 0; i < 72; i++) s += acc[i];
        if (s == 12345.678) out[0] = s;
    } else {
        // helper: HCHAIN dependent DFMA chains of length 16 per "block", paced to mimic ~260 FP64 instr per 288 DMMA
        double c[HCHAIN > 0 ? HCHAIN : 1];
#pragma unroll
        for (int i = 0; i < HCHAIN; i++) c[i] = seed * i;
        double a = seed + lane, b = seed - lane;
        for (int it = 0; it < iters / 8; it++) {
#pragma unroll
            for (int r = 0; r < 16; r++)
#pragma unroll
                for (int i = 0; i < HCHAIN; i++) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(c[i]) : "d"(a), "d"(b));
        }
        double s = 0;
#pragma unroll
        for (int i = 0; i < HCHAIN; i++) s += c[i];
        if (s == 12345.678) out[1] = s;
    }
}

template <typename F> float time_ms(F f) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    f(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 3; r++) { CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms; }
    return best;
}

template <int MODE, int HELPERS, int HCHAIN> void run(double* out, int sms) {
    const int iters = 40000;
    float ms = time_ms([&] { k_stream<MODE, HELPERS, HCHAIN><<<sms, 128 * (1 + HELPERS)>>>(out, iters, 1e-9); });
    CK(cudaGetLastError());
    double tf = (double)sms * 4 * iters * 36 * 512.0 / (ms * 1e-3) / 1e12;
    double helper_instr_per_kstep = HELPERS * HCHAIN * 16.0 / 8.0;
    printf("{\"bench\":\"mstream\",\"dmul\":%d,\"lds\":%d,\"helpers\":%d,\"helper_dfma_per_kstep\":%.0f,\"ms\":%.3f,\"dmma_tflops\":%.2f}\n",
           MODE & 1, (MODE >> 1) & 1, HELPERS, helper_instr_per_kstep, ms, tf);
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    double* out; CK(cudaMalloc(&out, 64));
    run<0, 0, 0>(out, sms);
    run<1, 0, 0>(out, sms);
    run<2, 0, 0>(out, sms);
    run<3, 0, 0>(out, sms);
    run<3, 1, 4>(out, sms);    //  8 helper DFMA per k-step
    run<3, 1, 8>(out, sms);    // 16
    run<3, 1, 16>(out, sms);   // 32  (~ what the real helpers issue)
    run<3, 2, 8>(out, sms);    // 32 split over two helper warps
    run<3, 2, 16>(out, sms);   // 64
    return 0;
}
