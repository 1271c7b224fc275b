// Microbenchmark 3: what exactly does a helper warp cost the tensor warp on the same
// scheduler?  Tensor warp: pure LDS + 36 DMMA per k-step (the shape of fused_hess_kernel's
// M warp).  One helper warp per scheduler runs one of:
//   0 nothing            1 dependent DFMA chain (ILP 1)    2 bursts of 16 independent DFMA
//   3 LDS.128+STS.128 traffic only (no FP64)               4 integer ALU only
//   5 DFMA chain with ILP 4                                 6 SHFL + DADD reduction pattern
// `per_kstep` = helper instructions issued per tensor-warp k-step (approx., by loop counts).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mstream3_bench mstream3_bench.cu
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
__device__ __forceinline__ void sts128(unsigned addr, double x, double y) {
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(x), "d"(y) : "memory");
}

template <int HMODE>
__global__ void __launch_bounds__(256, 1) k_stream(double* out, int iters, int helper_iters, double seed, volatile int* stop) {
    extern __shared__ __align__(16) double sm[];
    for (int i = threadIdx.x; i < (8 * 9216 + 512) / 8; i += blockDim.x) sm[i] = seed * (i % 97);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned base = (((unsigned)__cvta_generic_to_shared(sm) + 511u) & ~511u) + warp * 9216 + lane * 16;
    if (warp < 4) {
        double acc[72];
#pragma unroll
        for (int i = 0; i < 72; i++) acc[i] = 0.0;
        double avA[8], avB[8];
#pragma unroll
        for (int j = 0; j < 4; j++) { double2 v = lds128(base + j * 2048); avA[2 * j] = v.x; avA[2 * j + 1] = v.y; }
        for (int it = 0; it < iters; it += 2) {
#pragma unroll
            for (int j = 0; j < 4; j++) { double2 v = lds128(base + j * 2048 + ((it + 1) & 3) * 512); avB[2 * j] = v.x; avB[2 * j + 1] = v.y; }
#pragma unroll
            for (int I = 0; I < 8; I++)
#pragma unroll
                for (int J = 0; J <= I; J++) { const int t = I * (I + 1) / 2 + J; dmma(acc[2 * t], acc[2 * t + 1], avA[I], avA[J]); }
#pragma unroll
            for (int j = 0; j < 4; j++) { double2 v = lds128(base + j * 2048 + ((it + 2) & 3) * 512); avA[2 * j] = v.x; avA[2 * j + 1] = v.y; }
#pragma unroll
            for (int I = 0; I < 8; I++)
#pragma unroll
                for (int J = 0; J <= I; J++) { const int t = I * (I + 1) / 2 + J; dmma(acc[2 * t], acc[2 * t + 1], avB[I], avB[J]); }
        }
        double s = 0;
#pragma unroll
        for (int i = 0; i < 72; i++) s += acc[i];
        if (s == 12345.678) out[0] = s;
    } else if (HMODE != 0) {
        double a = seed + lane, b = 1.0 - seed, c[16];
        unsigned u = lane;
#pragma unroll
        for (int i = 0; i < 16; i++) c[i] = seed * i;
        for (int it = 0; it < helper_iters; it++) {
            if (HMODE == 1) {
#pragma unroll
                for (int r = 0; r < 16; r++) asm volatile("fma.rn.f64 %0, %1, %0, %2;" : "+d"(c[0]) : "d"(b), "d"(a));
            } else if (HMODE == 2) {
#pragma unroll
                for (int i = 0; i < 16; i++) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(c[i]) : "d"(a), "d"(b));
                __nanosleep(200);
            } else if (HMODE == 3) {
#pragma unroll
                for (int j = 0; j < 4; j++) { double2 v = lds128(base + j * 2048 + (it & 3) * 512); sts128(base + j * 2048 + ((it + 1) & 3) * 512, v.x, v.y); }
            } else if (HMODE == 4) {
#pragma unroll
                for (int r = 0; r < 16; r++) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(u) : "r"(u | 3), "r"(r));
            } else if (HMODE == 5) {
#pragma unroll
                for (int r = 0; r < 4; r++)
#pragma unroll
                    for (int i = 0; i < 4; i++) asm volatile("fma.rn.f64 %0, %1, %0, %2;" : "+d"(c[i]) : "d"(b), "d"(a));
            } else if (HMODE == 6) {
#pragma unroll
                for (int r = 0; r < 4; r++) { c[0] += __shfl_xor_sync(0xffffffffu, c[0], 16 >> r); }
            }
        }
        double s = u;
#pragma unroll
        for (int i = 0; i < 16; i++) s += c[i];
        if (s == 12345.678) out[1] = s;
    }
    (void)stop;
}

template <typename F> float time_ms(F f) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    f(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 3; r++) { CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms; }
    return best;
}

template <int HMODE> void run(double* out, int sms, int per_kstep_x16) {
    // helper executes `helper_iters` groups of 16 (or 8/4) instructions while the tensor warp does `iters` k-steps
    const int iters = 40000;
    const int helper_iters = (int)((long long)iters * per_kstep_x16 / 16);
    float ms = time_ms([&] { k_stream<HMODE><<<sms, 256, 8 * 9216 + 1024>>>(out, iters, helper_iters, 1e-9, nullptr); });
    CK(cudaGetLastError());
    double tf = (double)sms * 4 * iters * 36 * 512.0 / (ms * 1e-3) / 1e12;
    printf("{\"bench\":\"mstream3\",\"helper_mode\":%d,\"helper_instr_per_kstep\":%d,\"ms\":%.3f,\"dmma_tflops\":%.2f}\n", HMODE, per_kstep_x16, ms, tf);
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    double* out; CK(cudaMalloc(&out, 64));
    CK(cudaFuncSetAttribute(k_stream<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 9216 + 1024));
    CK(cudaFuncSetAttribute(k_stream<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 9216 + 1024));
    CK(cudaFuncSetAttribute(k_stream<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 9216 + 1024));
    CK(cudaFuncSetAttribute(k_stream<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 9216 + 1024));
    CK(cudaFuncSetAttribute(k_stream<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 9216 + 1024));
    CK(cudaFuncSetAttribute(k_stream<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 9216 + 1024));
    CK(cudaFuncSetAttribute(k_stream<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 9216 + 1024));
    run<0>(out, sms, 0);
    run<1>(out, sms, 16); run<1>(out, sms, 32);
    run<2>(out, sms, 16); run<2>(out, sms, 32);
    run<5>(out, sms, 16); run<5>(out, sms, 32);
    run<3>(out, sms, 8);  run<3>(out, sms, 16);
    run<4>(out, sms, 32); run<4>(out, sms, 64);
    run<6>(out, sms, 8);
    return 0;
}
