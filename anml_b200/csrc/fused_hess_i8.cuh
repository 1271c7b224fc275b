// Objective + gradient + Hessian of a one-parameter model with 64 design columns in ONE pass over X, with the
// Hessian contraction on the INT8 tensor cores (tcgen05.mma kind::i8, accumulators in TMEM) instead of DMMA.
//
// Why: H = X^T diag(w) X on DMMA costs 4608 pipe cycles per 32 rows and scheduler (fused_hess.cuh: 15.8 ms at
// n = 1e8 on one B200, 2x the HBM time of the same pass).  The INT8 tensor cores are ~120x faster per MAC and
// their int32 accumulation is EXACT, so the FP64 product can be rebuilt from exact integer pieces (Ozaki
// splitting; prototype and exactness check: tools/microbench/ozaki_i8.cu, profiles/r2_microbench_ozaki_i8.jsonl):
//   * x~ = sqrt(w) * x * cs_col (cs_col a power of two from the column's largest |x|, so that |x~| < 1/4) is
//     rounded ONCE to 48 fractional bits and written as six signed radix-256 digits,
//         x~ = sum_a d_a 2^(-8a),  a = 1..6,  d_a in [-128, 127];
//     one FMA does the scaling, the rounding and the digit bias: t = fma(x, sqrt(w) cs, 16 + B 2^-48) with
//     B = 0x808080808080 leaves the biased digits in the low six mantissa bytes of t (digit = byte ^ 0x80);
//   * the six digit planes of a 32-row block are int8 operands (K-major, 8x16-byte core matrices, no swizzle, K = row),
//     handled as three plane PAIRS {1,2}, {3,4}, {5,6} of 128 operand rows each (2 planes x 64 columns); four
//     tcgen05.mma (M = N = 128, K = 32) form {1,2}x{1,2}, {1,2}x{3,4}, {1,2}x{5,6} and {3,4}x{3,4} into four
//     128-column TMEM blocks: entry (m, n) of pair product (pp, pp') carries weight 2^-8(2(pp+pp') + 2 + x_m + x_n)
//     (x = plane parity inside the pair).  {3,4}x{1,2} and {5,6}x{1,2} are the transposes of two of those and are
//     added when the record is written; every other digit-plane product has weight <= 2^-64 2^14 and is dropped;
//   * one product of magnitude <= 2^14 meets an int32 accumulator per row: every EPOCH = 4095 blocks (131040 rows)
//     the accumulators are flushed, weighted, into an FP64 scratch matrix of the CTA.
// What is dropped: |x~ - fixed(x~)| <= 2^-49 per element and the products above: the Hessian agrees with the DMMA
// path to ~1e-14 of its largest entry (tests/test_gpu_fused.py), far inside the 1e-9 of BASELINE.json.  Integer sums
// do not depend on the order: results are bitwise reproducible.
//
// Roles (one persistent CTA per SM, 352 threads, 168 registers):
//   * 10 helper warps = 5 pairs; a pair owns whole 32-row blocks, each warp 16 rows of them (lanes 16..31 mirror
//     0..15 while the 16 rows' link / family terms are taken);
//   * 1 control warp: lane 0 requests blocks by TMA LOOKAHEAD ahead, and as each block's pair reports its digit
//     planes ready issues the four MMAs and commits them to the buffer's free[] barrier; it closes every epoch.
// Buffer life, operand layout and the in-place digit write are described above the kernel.
// The CTA's partial record has the layout of fused_hess_kernel (finish_fused_kernel is the shared epilogue).
#pragma once
#include "common.cuh"
#include "rowmath.cuh"
#include "fused_eval.cuh"

namespace anml {

struct I8Extra {
    const double* cs;       // [64] column scales (powers of two), then [64] their reciprocals
    double sv_scale;        // power of two folded into sqrt(w) (1 for the logistic w <= 1/4)
    double out_scale;       // 1 / sv_scale^2
    const unsigned long long* wmax_bits;  // MODE 1: bit pattern of max_i w_i (device, written by an earlier launch); the kernel
                                          // derives sv_scale from it instead of taking the host's
    double* scratch;        // [gridDim.x][2][2][64][64] FP64 flush target
};

struct I8Cfg {
#ifndef ANML_I8_PAIRS
#define ANML_I8_PAIRS 5
#endif
    static constexpr int PAIRS = ANML_I8_PAIRS;                            // helper warp pairs: a pair owns a 32-row block, each warp 16 rows of it
    static constexpr int NHELP = 2 * PAIRS;
    static constexpr int THREADS = 32 * (NHELP + 1);           // 352: 168 registers per thread, which the helpers need to stay out of local memory
                                                               // (7 pairs at 128 registers spill and lose 4 %, 6 pairs likewise, 4 pairs are too few: profiles/r2_ab_i8.txt)
    static constexpr int BUF_BYTES = 16384;                    // one block: 4 column groups x 32 rows x 128 B, later its digit planes
#ifndef ANML_I8_NBUF
#define ANML_I8_NBUF 12
#endif
    static constexpr int NBUF = ANML_I8_NBUF;
    static constexpr int LOOKAHEAD = 10;                       // blocks requested from HBM ahead of the MMA issue
    static constexpr int SBO = 1024, LBO = 144, PAIR_STRIDE = 288;
    static constexpr int EPOCH = 4095;                         // blocks between flushes: 2^14 * 32 * 4095 < 2^31 (one product per accumulator and row)
    static constexpr int NB8 = 8, NT = 36;
    static constexpr int EL = (2 * NT + NB8 + 1) * 32;
    static constexpr int TR_BYTES = 4 * 36 * 8;                // per helper: transpose scratch
    static constexpr int HT_BYTES = 256;                       // per helper: r and sqrt(w) of its 16 rows
    static constexpr int HREC_BYTES = (NB8 + 1) * 256;         // per helper: gradient / objective record at the end (in the idle ring)
    static constexpr int SMEM_BYTES = 1024 + NBUF * BUF_BYTES + 512 /*barriers, TMEM slot*/ + 2 * 64 * 8 /*beta, column scales*/ + NHELP * (TR_BYTES + HT_BYTES);
    static_assert(NHELP * HREC_BYTES <= NBUF * BUF_BYTES, "the helper records reuse the ring");
};

__device__ __forceinline__ uint64_t i8_smem_desc(uint32_t addr, int lbo, int sbo) {
    // K-major, SWIZZLE_NONE canonical layout ((8,n),2):((1,SBO),LBO) in 16-byte units; descriptor version 1 (sm_100)
    return (uint64_t)((addr >> 4) & 0x3fff) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
__host__ __device__ constexpr uint32_t i8_instr_desc(int M, int N) {
    // D = S32 (2 << 4) | A = INT8 (1 << 7) | B = INT8 (1 << 10) | both K-major | N >> 3 at bit 17 | M >> 4 at bit 24
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void i8_mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void i8_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void i8_tmem_ld32(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]),
          "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]),
          "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }

// 16 + B 2^-48, B = 0x808080808080: exactly representable (53 significant bits)
constexpr double I8_MAGIC = 16.0 + (double)0x808080808080ull / 281474976710656.0;

// Shared-memory life of a block buffer (16 KB):
//   1. TMA lands the block as 4 column groups x 32 rows x 128 B (the 3-D tensor map of fused_hess_kernel, 128-byte swizzle);
//   2. the pair's two warps (rows 0..15 / 16..31) read it for y = X beta, take the row terms, then group by group read
//      their rows again, meet at the pair's named barrier, and write the digit planes of that group IN PLACE: the 4 KB
//      of group j hold, for k = 2 x + ch (x = plane parity inside a plane pair, ch = which 8 of the 16 columns) and plane
//      pair pp = 0..2, a core-matrix pair at j 4096 + k 1024 + pp 288 (+144 for rows 16..31, the second K half);
//   3. seen through a descriptor with SBO = 1024 the buffer is, for every plane pair, a K-major 128 x 32 int8 operand
//      whose row m = 32 j + 16 x + 8 ch + r is digit plane 2 pp + 1 + x of column 16 j + 8 ch + r; four MMAs
//      (M = N = 128) form pair(pp) x pair(pp') for (pp, pp') = (0,0), (0,1), (0,2), (1,1) into four 128-column TMEM
//      blocks, where entry (m, n) carries weight 2^-8(2 (pp + pp') + 2 + x_m + x_n); (1,0) and (2,0) are the
//      transposes of (0,1) and (0,2) and are added when the record is written;
//   4. the MMAs' commit frees the buffer for the next TMA.
// MODE 0: row terms of a one-parameter model taken here (binomial/expit, gaussian/identity);
// MODE 1: term-fed block (BASELINE config #4, block (1,1)): r from a.rvec (nullptr: no gradient), w >= 0 from a.wvec, sqrt taken
//         here; no y pass, no row terms.  The bound on sqrt(w) is the exact maximum of the vector, left in device memory by
//         wmax_kernel: the scale is derived on the device, so it is as tight as for the logistic model.
template <int MODE>
__global__ void __launch_bounds__(I8Cfg::THREADS, 1) fused_hess_i8_kernel(const __grid_constant__ CUtensorMap tmap, const FusedArgs a, const I8Extra x) {
    using C = I8Cfg;
    constexpr int NBUF = C::NBUF, NB8 = C::NB8, NT = C::NT, NHELP = C::NHELP;

    extern __shared__ unsigned char smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t full_u32 = smem_base + NBUF * C::BUF_BYTES;     // full[NBUF] | ready[NBUF] | free[NBUF] | done
    const uint32_t ready_u32 = full_u32 + NBUF * 8;
    const uint32_t free_u32 = ready_u32 + NBUF * 8;
    const uint32_t done_u32 = free_u32 + NBUF * 8;
    const uint32_t slot_u32 = done_u32 + 8;                        // TMEM base address
    const uint32_t beta_u32 = full_u32 + 512;
    const uint32_t cs_u32 = beta_u32 + 64 * 8;
    const uint32_t tr_base = cs_u32 + 64 * 8;
    const uint32_t hterm_base = tr_base + NHELP * C::TR_BYTES;
    const uint32_t hrec_base = smem_base;  // after the last epoch the ring is idle
    static_assert((3 * NBUF + 1) * 8 + 16 <= 512, "barrier area");

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap);
        for (int i = 0; i < NBUF; ++i) {
            mbar_init(full_u32 + i * 8, 1);
            mbar_init(ready_u32 + i * 8, 2);
            mbar_init(free_u32 + i * 8, 1);
        }
        mbar_init(done_u32, 1);
        mbar_fence_init();
    }
    double sv_scale = x.sv_scale, out_scale = x.out_scale;
    if (MODE == 1 && x.wmax_bits) {
        // sqrt(w) <= sqrt(wmax) = f 2^e, f in [1/2, 1): sv_scale = 2^-(e+1) keeps sqrt(w) sv_scale < 1/2
        const double svm = sqrt(__longlong_as_double((long long)*x.wmax_bits));
        const int e = ((__double2hiint(svm) >> 20) & 0x7ff) - 1022;
        sv_scale = svm > 0.0 ? __longlong_as_double((long long)(1023 - (e + 1)) << 52) : 1.0;
        out_scale = 1.0 / (sv_scale * sv_scale);
    }
    if (threadIdx.x < 64) {
        sts64(beta_u32 + threadIdx.x * 8, MODE == 1 ? 0.0 : a.beta[threadIdx.x]);
        sts64(cs_u32 + threadIdx.x * 8, x.cs[threadIdx.x] * sv_scale);  // the global factor of sqrt(w) rides on the column scale
    }
    if (warp == NHELP) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_u32), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = lds32(slot_u32);

    // this CTA's blocks: global block blockIdx.x + q gridDim.x, q = 0 .. nq-1; pair pw owns q = pw (mod PAIRS)
    const long long nblk = (a.n + 31) / 32;
    const long long nq = nblk > blockIdx.x ? (nblk - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    double* scr = x.scratch + (size_t)blockIdx.x * 4 * 64 * 64;  // [S | T][plane parity][row][column]

    if (warp < NHELP) {
        // ================================ helper warps ================================
        const int pw = warp >> 1, h = warp & 1;
        double objacc = 0.0;
        LogProduct logprod;
        double gacc[NB8];
#pragma unroll
        for (int i = 0; i < NB8; ++i) gacc[i] = 0.0;
        bool bad_domain = false;
        uint32_t bad_range = 0;
        const uint32_t tr_u32 = tr_base + warp * C::TR_BYTES;
        const uint32_t ht_u32 = hterm_base + warp * C::HT_BYTES;
        const double sv_max = 0.5 / sv_scale;
        const int rl = lane & 15;                 // position 4s' + t' of the row whose terms this lane evaluates (lanes 16..31 mirror 0..15)
        const int my_row = 16 * h + 2 * (rl & 3) + ((rl >> 2) & 1) + 8 * (rl >> 3);
        // reading: k-step s holds rows 16h + 2t + (s & 1) + 8 (s >> 1) of column group j, 16-byte chunk g ^ (row & 7)
        // (128-byte swizzle): the 8 lanes of a quarter warp hit 8 distinct chunks
        uint32_t rd_off[4];
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            const int row = 16 * h + 2 * t + (s & 1) + 8 * (s >> 1);
            rd_off[s] = (uint32_t)(row * 128 + ((g ^ (row & 7)) << 4));
        }
        // writing: column 16j + 2g + e, rows k = 16h + 4t + (0..3): byte 4t of core-matrix row (2(g&3)+e) of slot ch = g>>2
        const uint32_t wr_off = (uint32_t)((g >> 2) * 1024 + h * C::LBO + (2 * (g & 3)) * 16 + 4 * t);
        bool first_flush = true;

        for (long long e0 = 0; e0 < nq; e0 += C::EPOCH) {
            const long long e1 = e0 + C::EPOCH < nq ? e0 + C::EPOCH : nq;
            for (long long q = e0 + ((pw - e0 % C::PAIRS) + C::PAIRS) % C::PAIRS; q < e1; q += C::PAIRS) {
                const long long row0 = ((long long)blockIdx.x + q * gridDim.x) * 32;
                const int buf = (int)(q % NBUF);
                const uint32_t bp = smem_base + buf * C::BUF_BYTES;
                const long long row = row0 + my_row;
                const bool valid = row < a.n;
                double rr = 0.0, sc = 0.0;
                if (MODE == 1) {
                    // ---- term-fed: r and w of the row come from the vectors of an earlier launch
                    if (valid) {
                        if (a.rvec) rr = a.rvec[row];
                        sc = sqrt(a.wvec[row]);
                        if (!(sc <= sv_max)) bad_range = 0xffff0000u;  // (also catches w < 0: NaN)
                    }
                    mbar_wait(full_u32 + buf * 8, (uint32_t)((q / NBUF) & 1));
                } else {
                    double off = 0.0, ob = 0.0, se = 1.0;
                    if (valid) {
                        if (a.offset) off = a.offset[row];
                        ob = a.obs[row];
                        if (a.se) se = a.se[row];
                    }
                    mbar_wait(full_u32 + buf * 8, (uint32_t)((q / NBUF) & 1));
                    // ---- pass 1: linear predictor of this warp's 16 rows
                    double part[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const double2 b = lds128(beta_u32 + (16 * j + 2 * g) * 8);
#pragma unroll
                        for (int s = 0; s < 4; ++s) {
                            const double2 v = lds128(bp + j * 4096 + rd_off[s]);
                            part[s] = fma(v.y, b.y, fma(v.x, b.x, part[s]));
                        }
                    }
                    // lane (g,t) stores part[s] at [s][g][t]; the lane at position 4s' + t' adds the 8 entries [s'][*][t']
#pragma unroll
                    for (int s = 0; s < 4; ++s) sts64(tr_u32 + (s * 36 + lane) * 8, part[s]);
                    __syncwarp();
                    double y;
                    {
                        const uint32_t base = tr_u32 + ((rl >> 2) * 36 + (rl & 3)) * 8;
                        const double q0 = lds64(base), q1 = lds64(base + 32), q2 = lds64(base + 64), q3 = lds64(base + 96);
                        const double q4 = lds64(base + 128), q5 = lds64(base + 160), q6 = lds64(base + 192), q7 = lds64(base + 224);
                        y = ((q0 + q1) + (q2 + q3)) + ((q4 + q5) + (q6 + q7));
                    }
                    if (valid) {
                        const RowTerms1 rt = row_terms1<true, true>(a.family, a.link, true, y + off, ob, se);
                        bad_domain |= !rt.ok;
                        if (lane < 16) {
                            objacc += rt.l;
                            logprod.mul(rt.ls);
                        }
                        rr = rt.r;
                        sc = rt.sw;
                        if (!(sc <= sv_max)) bad_range = 0xffff0000u;  // outside the digit range: flagged, the host reports it
                    }
                }
                if (lane < 16) {
                    sts64(ht_u32 + lane * 8, rr);
                    sts64(ht_u32 + 128 + lane * 8, sc);
                }
                __syncwarp();
                double rv[4], sv[4];
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    rv[s] = lds64(ht_u32 + (4 * s + t) * 8);
                    sv[s] = lds64(ht_u32 + 128 + (4 * s + t) * 8);
                }
                // ---- pass 2: gradient and digit planes, group by group, in place
                double2 cur[4], nxt[4];
#pragma unroll
                for (int s = 0; s < 4; ++s) cur[s] = lds128(bp + rd_off[s]);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (j < 3) {
#pragma unroll
                        for (int s = 0; s < 4; ++s) nxt[s] = lds128(bp + (j + 1) * 4096 + rd_off[s]);
                    }
                    // both warps of the pair hold their rows of groups <= j + 1 in registers: group j may be overwritten
                    asm volatile("bar.sync %0, 64;" ::"r"(1 + pw) : "memory");
                    const double2 csj = lds128(cs_u32 + (16 * j + 2 * g) * 8);
                    uint32_t dw[2][6];  // [column e][plane - 1]: the lane's four digits (rows k = 16h + 4t + 0..3) of that plane
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const double cse = e ? csj.y : csj.x;
                        uint32_t lo[4], hi[4];
#pragma unroll
                        for (int s = 0; s < 4; ++s) {
                            const double xe = e ? cur[s].y : cur[s].x;
                            gacc[2 * j + e] = fma(rv[s], xe, gacc[2 * j + e]);
                            const double tt = fma(xe, sv[s] * cse, I8_MAGIC);
                            lo[s] = (uint32_t)__double2loint(tt);
                            hi[s] = (uint32_t)__double2hiint(tt);
                            bad_range |= hi[s] ^ 0x40300000u;  // t must stay in [16, 17): exponent and top mantissa bits fixed
                        }
                        const uint32_t t0 = __byte_perm(lo[0], lo[1], 0x5140), t1 = __byte_perm(lo[0], lo[1], 0x7362);
                        const uint32_t t2 = __byte_perm(lo[2], lo[3], 0x5140), t3 = __byte_perm(lo[2], lo[3], 0x7362);
                        const uint32_t u0 = __byte_perm(hi[0], hi[1], 0x5140), u1 = __byte_perm(hi[2], hi[3], 0x5140);
                        dw[e][0] = __byte_perm(u0, u1, 0x7632) ^ 0x80808080u;  // byte 5 of the fixed-point value: plane 1
                        dw[e][1] = __byte_perm(u0, u1, 0x5410) ^ 0x80808080u;  // byte 4: plane 2
                        dw[e][2] = __byte_perm(t1, t3, 0x7632) ^ 0x80808080u;  // byte 3: plane 3
                        dw[e][3] = __byte_perm(t1, t3, 0x5410) ^ 0x80808080u;  // byte 2: plane 4
                        dw[e][4] = __byte_perm(t0, t2, 0x7632) ^ 0x80808080u;  // byte 1: plane 5
                        dw[e][5] = __byte_perm(t0, t2, 0x5410) ^ 0x80808080u;  // byte 0: plane 6
                    }
                    // plane a = 2 pp + 1 + x goes to slot 2x (+ ch, in wr_off) at pp * 288.  Lanes with ch = 1 store their odd
                    // column first: the two column halves of a store instruction then fall into different banks
                    const bool swap = (g >> 2) != 0;
#pragma unroll
                    for (int ee = 0; ee < 2; ++ee) {
                        const uint32_t cb = bp + j * 4096 + wr_off + ((ee != 0) != swap ? 16u : 0u);
#pragma unroll
                        for (int pl = 0; pl < 6; ++pl) {
                            const uint32_t v = ((ee != 0) != swap) ? dw[1][pl] : dw[0][pl];
                            sts32(cb + (pl & 1) * 2048 + (pl >> 1) * 288, v);
                        }
                    }
                    if (j < 3) {
#pragma unroll
                        for (int s = 0; s < 4; ++s) cur[s] = nxt[s];
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores before the tensor core's reads
                __syncwarp();
                if (lane == 0) mbar_arrive(ready_u32 + buf * 8);
            }
            // ---- epoch end: every MMA of the epoch is complete once the control warp passes barrier 8
            asm volatile("bar.sync 8, %0;" ::"n"(C::THREADS) : "memory");
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            {
                // this thread's TMEM lane m = 32 qd + lane: column group j = qd, plane parity xm, column 16 qd + 8 ch + r
                const int qd = warp & 3;
                const int nw = (NHELP - qd + 3) / 4;           // helper warps sharing this lane quarter
                const int xm = (lane >> 4) & 1;
                const int irow = 16 * qd + (lane & 15);        // 8 ch + r = lane & 15
                for (int cc = warp >> 2; cc < 4; cc += nw) {   // 32 TMEM columns = 16 design columns (both plane parities) at a time
                    // blocks 0 / 3 ({1,2}x{1,2}, {3,4}x{3,4}) are symmetric products; blocks 1 / 2 stand for themselves AND their transposes
                    double accS[16], accT[16];
#pragma unroll
                    for (int jj = 0; jj < 16; ++jj) accS[jj] = accT[jj] = 0.0;
#pragma unroll
                    for (int B = 0; B < 4; ++B) {
                        const int P = B == 3 ? 2 : B;  // plane pairs (pp, pp'): weight of (m, n) = 2^-8(2P + 2 + xm + xn)
                        const double w0 = __longlong_as_double((long long)(1023 - 8 * (2 * P + 2 + xm)) << 52);
                        const double w1 = w0 * 0.00390625;
                        uint32_t v[32];
                        i8_tmem_ld32(tmem + ((uint32_t)(32 * qd) << 16) + (uint32_t)(B * 128 + cc * 32), v);
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                        if (B == 0 || B == 3) {
#pragma unroll
                            for (int jj = 0; jj < 16; ++jj) accS[jj] = fma(w1, (double)(int)v[16 + jj], fma(w0, (double)(int)v[jj], accS[jj]));
                        } else {
#pragma unroll
                            for (int jj = 0; jj < 16; ++jj) accT[jj] = fma(w1, (double)(int)v[16 + jj], fma(w0, (double)(int)v[jj], accT[jj]));
                        }
                    }
                    double* dstS = scr + ((size_t)xm * 64 + irow) * 64 + cc * 16;
                    double* dstT = dstS + 2 * 64 * 64;
                    if (first_flush) {
#pragma unroll
                        for (int jj = 0; jj < 16; jj += 2) {
                            *reinterpret_cast<double2*>(dstS + jj) = make_double2(accS[jj], accS[jj + 1]);
                            *reinterpret_cast<double2*>(dstT + jj) = make_double2(accT[jj], accT[jj + 1]);
                        }
                    } else {
#pragma unroll
                        for (int jj = 0; jj < 16; jj += 2) {
                            double2 o = *reinterpret_cast<double2*>(dstS + jj);
                            o.x += accS[jj];
                            o.y += accS[jj + 1];
                            *reinterpret_cast<double2*>(dstS + jj) = o;
                            double2 u = *reinterpret_cast<double2*>(dstT + jj);
                            u.x += accT[jj];
                            u.y += accT[jj + 1];
                            *reinterpret_cast<double2*>(dstT + jj) = u;
                        }
                    }
                }
                first_flush = false;
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            asm volatile("bar.sync 9, %0;" ::"n"(C::THREADS) : "memory");  // the control warp may overwrite the accumulators
        }
        if (bad_domain) atomicOr(a.status, 1);
        if (bad_range & 0xffff0000u) atomicOr(a.status, 1024);
        {
            const uint32_t hb = hrec_base + warp * C::HREC_BYTES;
            __syncwarp();
#pragma unroll
            for (int jb = 0; jb < NB8; ++jb) sts64(hb + (jb * 32 + lane) * 8, gacc[jb]);
            sts64(hb + (NB8 * 32 + lane) * 8, lane < 16 ? objacc + logprod.log_value() : 0.0);
        }
    } else {
        // ================================ control warp: TMA requests and MMA issue ================================
        constexpr uint32_t IDESC = i8_instr_desc(128, 128);
        uint32_t done_phase = 0;
        auto request = [&](long long qq) {  // block qq -> buffer qq % NBUF (its previous occupant's MMAs must be complete)
            const int b = (int)(qq % NBUF);
            const long long use = qq / NBUF;
            if (use > 0) mbar_wait(free_u32 + b * 8, (uint32_t)((use - 1) & 1));
            const long long r0 = ((long long)blockIdx.x + qq * gridDim.x) * 32;
            mbar_arrive_expect_tx(full_u32 + b * 8, C::BUF_BYTES);
            tma_load_3d(smem_base + b * C::BUF_BYTES, &tmap, 0, (int)r0, 0, full_u32 + b * 8);
        };
        if (lane == 0) {
            const long long pre = nq < C::LOOKAHEAD ? nq : C::LOOKAHEAD;
            for (long long qq = 0; qq < pre; ++qq) request(qq);
        }
        for (long long e0 = 0; e0 < nq; e0 += C::EPOCH) {
            const long long e1 = e0 + C::EPOCH < nq ? e0 + C::EPOCH : nq;
            if (lane == 0) {
                for (long long q = e0; q < e1; ++q) {
                    if (q + C::LOOKAHEAD < nq) request(q + C::LOOKAHEAD);
                    const int b = (int)(q % NBUF);
                    mbar_wait(ready_u32 + b * 8, (uint32_t)((q / NBUF) & 1));
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t bp = smem_base + b * C::BUF_BYTES;
                    const uint64_t d0 = i8_smem_desc(bp, C::LBO, C::SBO), d1 = i8_smem_desc(bp + C::PAIR_STRIDE, C::LBO, C::SBO),
                                   d2 = i8_smem_desc(bp + 2 * C::PAIR_STRIDE, C::LBO, C::SBO);
                    const uint32_t accumulate = q > e0 ? 1u : 0u;
                    i8_mma(tmem + 0, d0, d0, IDESC, accumulate);    // planes {1,2} x {1,2}: symmetric by itself
                    i8_mma(tmem + 128, d0, d1, IDESC, accumulate);  // {1,2} x {3,4}: {3,4} x {1,2} is its transpose (added at the end)
                    i8_mma(tmem + 256, d0, d2, IDESC, accumulate);  // {1,2} x {5,6}: likewise
                    i8_mma(tmem + 384, d1, d1, IDESC, accumulate);  // {3,4} x {3,4}
                    i8_commit(free_u32 + b * 8);
                }
                i8_commit(done_u32);
                mbar_wait(done_u32, done_phase);
            }
            done_phase ^= 1u;
            __syncwarp();
            asm volatile("bar.sync 8, %0;" ::"n"(C::THREADS) : "memory");
            asm volatile("bar.sync 9, %0;" ::"n"(C::THREADS) : "memory");
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
    }
    __syncthreads();
    // ---- the CTA's partial record (layout of fused_hess_kernel): Hessian rows (tile, reg), gradient rows, objective row
    double* out = a.partials + (size_t)blockIdx.x * C::EL;
    const double* ics = x.cs + 64;
    for (int e = threadIdx.x; e < C::EL; e += C::THREADS) {
        double s = 0.0;
        if (e < 2 * NT * 32) {
            const int rowi = e >> 5, ln = e & 31;
            const int reg = rowi & 1, tile = rowi >> 1;
            int I = 0;
            while ((I + 1) * (I + 2) / 2 <= tile) ++I;
            const int J = tile - I * (I + 1) / 2;
            const int c1 = 16 * (I >> 1) + 2 * (ln >> 2) + (I & 1);
            const int c2 = 16 * (J >> 1) + 2 * (2 * (ln & 3) + reg) + (J & 1);
            if (nq > 0) {
                const double* T = scr + 2 * 64 * 64;
                const size_t ij = (size_t)c1 * 64 + c2, ji = (size_t)c2 * 64 + c1;
                s = ((scr[ij] + scr[4096 + ij]) + ((T[ij] + T[4096 + ij]) + (T[ji] + T[4096 + ji]))) * (ics[c1] * ics[c2] * out_scale);
            }
        } else {
            const int l = e - 2 * NT * 32;
#pragma unroll
            for (int w = 0; w < NHELP; ++w) s += lds64(hrec_base + w * C::HREC_BYTES + l * 8);
        }
        out[e] = s;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == NHELP) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

// largest entry of a non-negative vector, as a bit pattern (0 for an empty / all-zero vector; NaN or negative entries give a
// pattern above every finite positive double, which the consumer's range check then rejects)
__global__ void wmax_kernel(const double* __restrict__ w, long long n, unsigned long long* __restrict__ out) {
    unsigned long long m = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const unsigned long long b = (unsigned long long)__double_as_longlong(w[i]);
        m = b > m ? b : m;
    }
#pragma unroll
    for (int k = 16; k >= 1; k >>= 1) {
        const unsigned long long o = __shfl_xor_sync(FULL_MASK, m, k);
        m = o > m ? o : m;
    }
    if ((threadIdx.x & 31) == 0 && m) atomicMax(out, m);
}

// largest |x| per column (64 columns), as the bit pattern of a non-negative double (integer order = value order)
__global__ void colmax64_kernel(const double* __restrict__ X, long long n, long long ld, unsigned long long* __restrict__ out) {
    __shared__ unsigned long long red[256];
    const int c = threadIdx.x & 63, r0 = threadIdx.x >> 6;  // 256 threads: 4 rows at a time
    double m = 0.0;
    bool nan = false;
    for (long long r = (long long)blockIdx.x * 4 + r0; r < n; r += (long long)gridDim.x * 4) {
        const double v = fabs(X[r * ld + c]);
        nan |= !(v == v);
        m = fmax(m, v);
    }
    red[threadIdx.x] = nan ? 0x7ff8000000000000ull : (unsigned long long)__double_as_longlong(m);
    __syncthreads();
    if (threadIdx.x < 64) {
        unsigned long long v = red[threadIdx.x];
        for (int k = 1; k < 4; ++k) v = v > red[threadIdx.x + 64 * k] ? v : red[threadIdx.x + 64 * k];
        atomicMax(out + c, v);
    }
}

}  // namespace anml
