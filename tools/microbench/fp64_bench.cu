// Microbenchmarks that decide the Hessian kernel's shape on B200 (sm_100a):
//   dmma   : DMMA.8x8x4 issue rate with NACC independent accumulators per warp
//   dfma   : plain DFMA rate (same nominal peak on B200) for comparison
//   mixed  : DMMA with R DFMAs interleaved per DMMA (do they share a pipe?)
//   hbm    : read-only streaming bandwidth (LDG.128, persistent grid)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_bench fp64_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NACC, int NFMA>
__global__ void __launch_bounds__(256, 1) k_dmma(double* out, int iters, double seed) {
    double c0[NACC], c1[NACC];
    double f[NFMA > 0 ? NFMA : 1];
#pragma unroll
    for (int i = 0; i < NACC; i++) { c0[i] = seed * i; c1[i] = seed; }
#pragma unroll
    for (int i = 0; i < NFMA; i++) f[i] = seed + i;
    double a = seed + threadIdx.x, b = seed - threadIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) {
            dmma884(c0[i], c1[i], a, b);
            if (NFMA > 0) {
                // NFMA DFMAs per NACC DMMAs, spread evenly
                if ((i * NFMA) / NACC != ((i + 1) * NFMA) / NACC) {
                    int q = (i * NFMA) / NACC;
                    asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(f[q]) : "d"(a), "d"(b));
                }
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += c0[i] + c1[i];
#pragma unroll
    for (int i = 0; i < NFMA; i++) s += f[i];
    if (s == 12345.678) out[0] = s;
}

template <int NACC>
__global__ void __launch_bounds__(1024) k_dfma(double* out, int iters, double seed) {
    double c[NACC];
#pragma unroll
    for (int i = 0; i < NACC; i++) c[i] = seed * i;
    double a = seed + threadIdx.x, b = seed - threadIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(c[i]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += c[i];
    if (s == 12345.678) out[0] = s;
}

__global__ void __launch_bounds__(512) k_hbm(const double2* __restrict__ x, size_t n2, double* out) {
    double s = 0;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n2; i += 4 * stride) {
        double2 v0 = __ldg(x + i), v1 = __ldg(x + i + stride), v2 = __ldg(x + i + 2 * stride), v3 = __ldg(x + i + 3 * stride);
        s += v0.x + v0.y + v1.x + v1.y + v2.x + v2.y + v3.x + v3.y;
    }
    for (; i < n2; i += stride) { double2 v = __ldg(x + i); s += v.x + v.y; }
    if (s == 12345.678) out[0] = s;
}

template <typename F>
float time_ms(F f, int reps = 5) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    f(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    return best;
}

template <int NACC, int NFMA>
void run_dmma(double* out, int warps_per_cta, int ctas_per_sm, int sms) {
    int iters = 20000;
    int grid = sms * ctas_per_sm, block = warps_per_cta * 32;
    float ms = time_ms([&] { k_dmma<NACC, NFMA><<<grid, block>>>(out, iters, 1e-9); });
    CK(cudaGetLastError());
    double mmas = (double)grid * warps_per_cta * iters * NACC;
    double tf = mmas * 512.0 / (ms * 1e-3) / 1e12;
    double fma_tf = mmas / NACC * NFMA * 64.0 / (ms * 1e-3) / 1e12;
    printf("{\"bench\":\"dmma\",\"nacc\":%d,\"nfma_per_group\":%d,\"warps_per_sm\":%d,\"ms\":%.3f,\"dmma_tflops\":%.2f,\"dfma_tflops\":%.2f}\n",
           NACC, NFMA, warps_per_cta * ctas_per_sm, ms, tf, fma_tf);
}

template <int NACC>
void run_dfma(double* out, int warps_per_cta, int sms) {
    int iters = 20000;
    float ms = time_ms([&] { k_dfma<NACC><<<sms, warps_per_cta * 32>>>(out, iters, 1e-9); });
    double fl = (double)sms * warps_per_cta * iters * NACC * 64.0;
    printf("{\"bench\":\"dfma\",\"nacc\":%d,\"warps_per_sm\":%d,\"ms\":%.3f,\"tflops\":%.2f}\n", NACC, warps_per_cta, ms, fl / (ms * 1e-3) / 1e12);
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    printf("{\"device\":\"%s\",\"sms\":%d,\"clock_khz\":%d}\n", p.name, sms, p.clockRate);
    double* out; CK(cudaMalloc(&out, 64));
    // DMMA: accumulators per warp x warps per SM (CTAs of <= 8 warps, 255 regs allowed)
    run_dmma<1, 0>(out, 4, 1, sms);
    run_dmma<2, 0>(out, 4, 1, sms);
    run_dmma<4, 0>(out, 4, 1, sms);
    run_dmma<8, 0>(out, 4, 1, sms);
    run_dmma<36, 0>(out, 4, 1, sms);
    run_dmma<36, 0>(out, 8, 1, sms);
    run_dmma<8, 0>(out, 8, 2, sms);
    // mixed: 36 DMMA + k DFMA per group, 8 warps/SM
    run_dmma<36, 4>(out, 8, 1, sms);
    run_dmma<36, 9>(out, 8, 1, sms);
    run_dmma<36, 18>(out, 8, 1, sms);
    run_dmma<36, 36>(out, 8, 1, sms);
    run_dmma<36, 36>(out, 4, 1, sms);
    // DFMA
    run_dfma<8>(out, 4, sms);
    run_dfma<8>(out, 8, sms);
    run_dfma<8>(out, 16, sms);
    run_dfma<16>(out, 32, sms);
    // HBM read
    size_t bytes = (size_t)8 << 30;
    double2* x; CK(cudaMalloc(&x, bytes)); CK(cudaMemset(x, 0, bytes));
    for (int mult = 1; mult <= 8; mult *= 2) {
        float ms = time_ms([&] { k_hbm<<<sms * mult, 512>>>(x, bytes / 16, out); });
        printf("{\"bench\":\"hbm_read\",\"ctas_per_sm\":%d,\"ms\":%.3f,\"gbs\":%.1f}\n", mult, ms, bytes / (ms * 1e-3) / 1e9);
    }
    return 0;
}
