// Microbenchmark 2: which source shape gives the tensor warp a well-pipelined
// stream?  Per 4-row k-step: 4 LDS.128 (fragments) + 1 LDS.64 (w), 8 DMUL (B = w*a),
// optionally 8 DFMA (gradient), 36 DMMA.8x8x4 (lower-triangle tiles).
//   V0: naive (loads, scale, DMMAs in program order)
//   V1: double-buffered, next loads before the DMMAs, next scaling after them
//   V2: as V1 but next scaling spread between the DMMAs (asm volatile everywhere)
//   V3: V1 with two warps per scheduler, each doing half of the tiles (column split)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mstream2_bench mstream2_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma_nv(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ double2 lds128(unsigned addr) {
    double2 v; asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr)); return v;
}
__device__ __forceinline__ double lds64(unsigned addr) {
    double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr)); return v;
}
__device__ __forceinline__ void dmul_v(double& d, double a, double b) { asm volatile("mul.f64 %0, %1, %2;" : "=d"(d) : "d"(a), "d"(b)); }
__device__ __forceinline__ void dfma_v(double& d, double a, double b) { asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(d) : "d"(a), "d"(b)); }

__device__ __forceinline__ void load_frag(double* av, double& w, unsigned base, int k) {
    w = lds64((base & ~511u) + 8192 + (k & 31) * 8);
#pragma unroll
    for (int j = 0; j < 4; j++) { double2 v = lds128(base + j * 2048 + (k & 3) * 512); av[2 * j] = v.x; av[2 * j + 1] = v.y; }
}

template <int VARIANT, bool GRAD>
__global__ void __launch_bounds__(VARIANT == 3 ? 256 : 128, 1) k_stream(double* out, int iters, double seed) {
    extern __shared__ __align__(16) double sm[];
    for (int i = threadIdx.x; i < (4 * 9216 + 512) / 8; i += blockDim.x) sm[i] = seed * (i % 97);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned base = ((unsigned)__cvta_generic_to_shared(sm) + 511u & ~511u) + (warp & 3) * 9216 + lane * 16;
    double acc[72];
#pragma unroll
    for (int i = 0; i < 72; i++) acc[i] = 0.0;
    double gacc[8];
#pragma unroll
    for (int i = 0; i < 8; i++) gacc[i] = 0.0;

    if (VARIANT == 0) {
        for (int it = 0; it < iters; it++) {
            double av[8], bv[8], w;
            load_frag(av, w, base, it);
#pragma unroll
            for (int j = 0; j < 8; j++) bv[j] = w * av[j];
            if (GRAD) {
#pragma unroll
                for (int j = 0; j < 8; j++) gacc[j] = fma(w, av[j], gacc[j]);
            }
#pragma unroll
            for (int I = 0; I < 8; I++)
#pragma unroll
                for (int J = 0; J <= I; J++) { const int t = I * (I + 1) / 2 + J; dmma_nv(acc[2 * t], acc[2 * t + 1], av[I], bv[J]); }
        }
    } else if (VARIANT == 1 || VARIANT == 2) {
        double avA[8], bvA[8], avB[8], bvB[8], wA, wB;
        load_frag(avA, wA, base, 0);
#pragma unroll
        for (int j = 0; j < 8; j++) bvA[j] = wA * avA[j];
        for (int it = 0; it < iters; it += 2) {
            // ---- k-step on set A, prefetch into B
            load_frag(avB, wB, base, it + 1);
            if (VARIANT == 1) {
#pragma unroll
                for (int I = 0; I < 8; I++)
#pragma unroll
                    for (int J = 0; J <= I; J++) { const int t = I * (I + 1) / 2 + J; dmma(acc[2 * t], acc[2 * t + 1], avA[I], bvA[J]); }
                if (GRAD) {
#pragma unroll
                    for (int j = 0; j < 8; j++) dfma_v(gacc[j], wA, avA[j]);
                }
#pragma unroll
                for (int j = 0; j < 8; j++) dmul_v(bvB[j], wB, avB[j]);
            } else {
#pragma unroll
                for (int I = 0; I < 8; I++)
#pragma unroll
                    for (int J = 0; J <= I; J++) {
                        const int t = I * (I + 1) / 2 + J;
                        dmma(acc[2 * t], acc[2 * t + 1], avA[I], bvA[J]);
                        if ((t & 1) && (t >> 1) < 8 && GRAD) dfma_v(gacc[t >> 1], wA, avA[t >> 1]);
                        if ((t & 1) && (t >> 1) >= 8 && (t >> 1) < 16) dmul_v(bvB[(t >> 1) - 8], wB, avB[(t >> 1) - 8]);
                    }
            }
            // ---- k-step on set B, prefetch into A
            load_frag(avA, wA, base, it + 2);
            if (VARIANT == 1) {
#pragma unroll
                for (int I = 0; I < 8; I++)
#pragma unroll
                    for (int J = 0; J <= I; J++) { const int t = I * (I + 1) / 2 + J; dmma(acc[2 * t], acc[2 * t + 1], avB[I], bvB[J]); }
                if (GRAD) {
#pragma unroll
                    for (int j = 0; j < 8; j++) dfma_v(gacc[j], wB, avB[j]);
                }
#pragma unroll
                for (int j = 0; j < 8; j++) dmul_v(bvA[j], wA, avA[j]);
            } else {
#pragma unroll
                for (int I = 0; I < 8; I++)
#pragma unroll
                    for (int J = 0; J <= I; J++) {
                        const int t = I * (I + 1) / 2 + J;
                        dmma(acc[2 * t], acc[2 * t + 1], avB[I], bvB[J]);
                        if ((t & 1) && (t >> 1) < 8 && GRAD) dfma_v(gacc[t >> 1], wB, avB[t >> 1]);
                        if ((t & 1) && (t >> 1) >= 8 && (t >> 1) < 16) dmul_v(bvA[(t >> 1) - 8], wA, avA[(t >> 1) - 8]);
                    }
            }
        }
    } else {
        // two warps per scheduler; warp>>2 selects the column-block half (J in {0,7,1,6} vs {2,5,3,4})
        const bool first = (warp >> 2) == 0;
        for (int it = 0; it < iters; it++) {
            double av[8], bv[8], w;
            load_frag(av, w, base, it);
#pragma unroll
            for (int j = 0; j < 8; j++) bv[j] = w * av[j];
            if (GRAD && first) {
#pragma unroll
                for (int j = 0; j < 8; j++) gacc[j] = fma(w, av[j], gacc[j]);
            }
#pragma unroll
            for (int J = 0; J < 8; J++)
#pragma unroll
                for (int I = J; I < 8; I++) {
                    const int q = J < 7 - J ? J : 7 - J;
                    const int t = I * (I + 1) / 2 + J;
                    if (((q & 1) == 0) == first) dmma_nv(acc[2 * t], acc[2 * t + 1], av[I], bv[J]);
                }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 72; i++) s += acc[i];
#pragma unroll
    for (int i = 0; i < 8; i++) s += gacc[i];
    if (s == 12345.678) out[0] = s;
}

template <typename F> float time_ms(F f) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    f(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 3; r++) { CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms; }
    return best;
}

template <int VARIANT, bool GRAD> void run(double* out, int sms) {
    const int iters = 40000;
    const int threads = VARIANT == 3 ? 256 : 128;
    float ms = time_ms([&] { k_stream<VARIANT, GRAD><<<sms, threads, 4 * 9216 + 512>>>(out, iters, 1e-9); });
    CK(cudaGetLastError());
    double tf = (double)sms * 4 * iters * 36 * 512.0 / (ms * 1e-3) / 1e12;
    printf("{\"bench\":\"mstream2\",\"variant\":%d,\"grad\":%d,\"ms\":%.3f,\"dmma_tflops\":%.2f}\n", VARIANT, (int)GRAD, ms, tf);
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    double* out; CK(cudaMalloc(&out, 64));
    run<0, false>(out, sms); run<0, true>(out, sms);
    run<1, false>(out, sms); run<1, true>(out, sms);
    run<2, false>(out, sms); run<2, true>(out, sms);
    run<3, false>(out, sms); run<3, true>(out, sms);
    return 0;
}
