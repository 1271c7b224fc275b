// probe: which f64 mma.sync shapes assemble for sm_100a and what SASS they become
#include <cstdio>
__global__ void k884(double* out, const double* in) {
    double a = in[threadIdx.x], b = in[threadIdx.x + 32];
    double c0 = 0, c1 = 0;
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
    out[threadIdx.x] = c0 + c1;
}
__global__ void k1684(double* out, const double* in) {
    double a0 = in[threadIdx.x], a1 = in[threadIdx.x + 64], b = in[threadIdx.x + 32];
    double c0 = 0, c1 = 0, c2 = 0, c3 = 0;
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+d"(c0), "+d"(c1), "+d"(c2), "+d"(c3) : "d"(a0), "d"(a1), "d"(b));
    out[threadIdx.x] = c0 + c1 + c2 + c3;
}
__global__ void k1688(double* out, const double* in) {
    double a0 = in[threadIdx.x], a1 = in[threadIdx.x + 64], a2 = in[threadIdx.x + 96], a3 = in[threadIdx.x + 128];
    double b0 = in[threadIdx.x + 32], b1 = in[threadIdx.x + 33];
    double c0 = 0, c1 = 0, c2 = 0, c3 = 0;
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+d"(c0), "+d"(c1), "+d"(c2), "+d"(c3) : "d"(a0), "d"(a1), "d"(a2), "d"(a3), "d"(b0), "d"(b1));
    out[threadIdx.x] = c0 + c1 + c2 + c3;
}
__global__ void k16816(double* out, const double* in) {
    double a[8], b[4];
    for (int i = 0; i < 8; i++) a[i] = in[threadIdx.x + 32 * i];
    for (int i = 0; i < 4; i++) b[i] = in[threadIdx.x + 7 * i + 300];
    double c0 = 0, c1 = 0, c2 = 0, c3 = 0;
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                 : "+d"(c0), "+d"(c1), "+d"(c2), "+d"(c3)
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                   "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
    out[threadIdx.x] = c0 + c1 + c2 + c3;
}
