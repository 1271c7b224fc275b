// Warp-specialised fused evaluator: objective + gradient + Hessian in one pass
// over X for one Parameter with p <= 64 columns (the north-star configuration).
//
// Why a second kernel: in fused_eval_kernel every warp alternates between a
// latency-bound phase (linear predictor, exp/log) and a DMMA-bound phase; with
// the 255-register accumulator tile only two warps fit per scheduler and they
// phase-lock, leaving the shared FP64/tensor pipe ~25% idle (ncu, profiles/).
// Here the two phases run in different warps:
//
//   * one persistent CTA of 12 warps per SM = 4 "teams"; team q = tensor warp q
//     (M) + helper warps q+4 and q+8 (H0, H1).  All three sit on scheduler q, so
//     every scheduler owns one uninterrupted DMMA stream and two helpers.
//     setmaxnreg moves registers from the helpers (128) to the tensor warps (248).
//   * each team owns 32-row blocks b = gp, gp+GP, ... and a private ring of NBUF
//     block buffers (32 rows x 16*NB16 columns, NB16 TMA boxes of 32x16 doubles,
//     128-byte swizzle, zero fill outside the matrix).
//   * helper H(i&1) handles the team's block i: waits for the TMA bytes, y = X beta
//     on MMA-layout fragments + 8-lane transposing reduction, per-row link + family
//     terms (one row per lane), publishes (r, w)[32] through shared memory with an
//     mbarrier arrive.  A first version with one helper per tensor warp was
//     helper-bound: its FP64 instructions queue behind the DMMAs on the shared pipe.
//   * tensor warp M: waits for (r, w), streams the block from shared memory as
//     DMMA.8x8x4 fragments (A = X^T, B = w * X; 36 tiles on/below the diagonal for
//     p = 64), g += r_k x_k from the same fragments, then re-arms the TMA for the
//     buffer it just consumed.  The barrier of the next block is probed before the
//     last k-steps so the probe latency hides behind the DMMA stream.
//
// Ordering argument for the lock-free buffer reuse: M can finish block i only
// after the helper published (r, w) for block i, which happens after the helper
// has read block i; so when M re-issues TMA into that buffer nobody reads it any
// more, and the (r, w) slot of the buffer is likewise dead.
//
// Partial-record layout is the one of fused_eval_kernel (reduce_partials_kernel
// and assemble_kernel are shared).
#pragma once
#include "common.cuh"
#include "rowmath.cuh"
#include "fused_eval.cuh"

namespace anml {

template <int NB16, int NHELP = 2>
struct HessCfg {
    static constexpr int TEAMS = 4;
    static constexpr int HELPERS = NHELP;  // per team
    static constexpr int WARPS = TEAMS * (1 + HELPERS);
    static constexpr int THREADS = WARPS * 32;
    static constexpr int REGS_M = 248, REGS_H = 128;  // NHELP == 2: 4*32*248 + 8*32*128 == 384*168 (NHELP == 1: no transfer needed)
    static constexpr int NB8 = 2 * NB16;
    static constexpr int NT = NB8 * (NB8 + 1) / 2;
    static constexpr int BUF_BYTES = NB16 * 4096;  // NB16 boxes of 32 rows x 128 B
    static constexpr int NBUF = (49152 / BUF_BYTES) > 8 ? 8 : (49152 / BUF_BYTES);
    static constexpr int TEAM_BYTES = NBUF * BUF_BYTES;
    static constexpr int EL = (2 * NT + NB8 + 1) * 32;
    static constexpr int SMEM_BYTES = 1024 + TEAMS * TEAM_BYTES + TEAMS * NBUF * 64 * 8 /*r,w*/ + TEAMS * NBUF * 2 * 8 /*bars*/ +
                                      16 * NB16 * 8 /*beta*/ + TEAMS * HELPERS * 32 * 8 /*obj*/;
};

// Volatile forms pin the PTX order so the hand-made software pipeline below
// survives: every FP64 helper instruction of the tensor warp is issued between
// DMMAs that do not depend on it.
__device__ __forceinline__ void dmma884_v(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dfma_v(double& d, double a, double b) {
    asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(d) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmul_v(double& d, double a, double b) {
    asm volatile("mul.f64 %0, %1, %2;" : "=d"(d) : "d"(a), "d"(b));
}

// fragments of one 4-row k-step: av[jb] = X[row(t)][col(jb, g)], plus the row's r and w
template <int NB16>
__device__ __forceinline__ void load_kstep(double* av, double& rv, double& wv, uint32_t bp, uint32_t tb, int s, int g, int t) {
    // k-step s holds rows 16*(s>>2) + 8*((s>>1)&1) + (s&1) + 2t, t = 0..3
    const int xr = 2 * t + (s & 1);
    const int row = 16 * (s >> 2) + 8 * ((s >> 1) & 1) + xr;
    const uint32_t rp = bp + row * 128 + ((g ^ xr) << 4);
    rv = lds64(tb + (4 * s + t) * 8);
    wv = lds64(tb + 256 + (4 * s + t) * 8);
#pragma unroll
    for (int j = 0; j < NB16; ++j) {
        const double2 v = lds128(rp + j * 4096);
        av[2 * j] = v.x;
        av[2 * j + 1] = v.y;
    }
}

// One k-step of the tensor warp: NT DMMAs on (av0, bv0); between them, one helper
// instruction after every second DMMA: first the gradient FMAs of this k-step,
// then the B-operand scaling bv1 = wv1 * av1 of the NEXT k-step (its loads were
// issued before the first DMMA, so their latency hides behind the DMMAs).
template <int NB8>
__device__ __forceinline__ void mma_kstep(double* acc, double* gacc, const double* av0, const double* bv0, double rv0,
                                          const double* av1, double* bv1, double wv1) {
    constexpr int NT = NB8 * (NB8 + 1) / 2;
#pragma unroll
    for (int I = 0; I < NB8; ++I) {
#pragma unroll
        for (int J = 0; J <= I; ++J) {
            const int tile = I * (I + 1) / 2 + J;
            dmma884_v(acc[2 * tile], acc[2 * tile + 1], av0[I], bv0[J]);
            if (tile & 1) {
                const int q = tile >> 1;
                if (q < NB8) dfma_v(gacc[q], rv0, av0[q]);
                else if (q < 2 * NB8) dmul_v(bv1[q - NB8], wv1, av1[q - NB8]);
            }
        }
    }
#pragma unroll
    for (int q = NT / 2; q < 2 * NB8; ++q) {  // narrow designs: fewer DMMAs than helper instructions
        if (q < NB8) dfma_v(gacc[q], rv0, av0[q]);
        else dmul_v(bv1[q - NB8], wv1, av1[q - NB8]);
    }
}

template <int NB16, int NHELP>
__global__ void __launch_bounds__(128 * (1 + NHELP), 1) fused_hess_kernel(const __grid_constant__ CUtensorMap tmap32, const FusedArgs a) {
    using C = HessCfg<NB16, NHELP>;
    constexpr int NB8 = C::NB8, NT = C::NT, NBUF = C::NBUF, TEAMS = C::TEAMS;

    extern __shared__ unsigned char smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* smem = smem_raw + (smem_base - smem_u32(smem_raw));
    // layout: rings | (r,w) terms | barriers | beta | helper objective partials
    const uint32_t terms_base = smem_base + TEAMS * C::TEAM_BYTES;
    const uint32_t bars_base = terms_base + TEAMS * NBUF * 64 * 8;
    const uint32_t beta_base = bars_base + TEAMS * NBUF * 2 * 8;
    const uint32_t obj_base = beta_base + 16 * NB16 * 8;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const int team = warp & 3;
    const int role = warp >> 2;  // 0 = tensor warp, 1..2 = helpers

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap32);
        for (int i = 0; i < TEAMS * NBUF * 2; ++i) mbar_init(bars_base + i * 8, 1);
        mbar_fence_init();
    }
    if (threadIdx.x < 16 * NB16) sts64(beta_base + threadIdx.x * 8, a.beta[threadIdx.x]);
    __syncthreads();

    const uint32_t ring_u32 = smem_base + team * C::TEAM_BYTES;
    const uint32_t term_u32 = terms_base + team * NBUF * 64 * 8;        // [buf][0..31] = r, [32..63] = w
    const uint32_t full_u32 = bars_base + team * NBUF * 2 * 8;          // full[b]  : TMA bytes landed
    const uint32_t tbar_u32 = full_u32 + NBUF * 8;                      // terms[b] : (r, w) published

    const long long nblk = (a.n + 31) >> 5;
    const long long gp = (long long)blockIdx.x * TEAMS + team;
    const long long GP = (long long)gridDim.x * TEAMS;
    const long long nmine = (nblk > gp) ? (nblk - gp + GP - 1) / GP : 0;

    double* scratch = reinterpret_cast<double*>(smem + team * C::TEAM_BYTES);  // epilogue: EL doubles per team
    static_assert(C::EL * 8 <= C::TEAM_BYTES, "reduction scratch must fit in the team's ring");

    if (role == 0) {
        // ============================ tensor warp ============================
        if (NHELP == 2) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(C::REGS_M));
        auto issue = [&](long long blk_seq, int buf) {
            const long long row0 = (gp + blk_seq * GP) * 32;
            const uint32_t bar = full_u32 + buf * 8;
            const uint32_t dst = ring_u32 + buf * C::BUF_BYTES;
            mbar_arrive_expect_tx(bar, C::BUF_BYTES);
#pragma unroll
            for (int j = 0; j < NB16; ++j) tma_load_2d(dst + j * 4096, &tmap32, 16 * j, (int)row0, bar);
        };
        if (lane == 0) {
            const int pre = nmine < NBUF ? (int)nmine : NBUF;
            for (int i = 0; i < pre; ++i) issue(i, i);
        }
        double acc[2 * NT];
        double gacc[NB8];
#pragma unroll
        for (int i = 0; i < 2 * NT; ++i) acc[i] = 0.0;
#pragma unroll
        for (int i = 0; i < NB8; ++i) gacc[i] = 0.0;

        // Software pipeline over k-steps: while the DMMAs of k-step k run on register
        // set 0, the fragments of k-step k+1 are loaded and scaled into set 1.  Sets A/B
        // alternate (8 k-steps per block, even), so no register moves are needed.
        double avA[NB8], bvA[NB8], avB[NB8], bvB[NB8];
        double rvA = 0.0, rvB = 0.0, wvA = 0.0, wvB = 0.0;
        uint32_t phase_bits = 0;
        int buf = 0;
        if (nmine > 0) {
            mbar_wait(tbar_u32, 0);
            mbar_wait(full_u32, 0);
            load_kstep<NB16>(avA, rvA, wvA, ring_u32, term_u32, 0, g, t);
#pragma unroll
            for (int jb = 0; jb < NB8; ++jb) bvA[jb] = wvA * avA[jb];
        }
        for (long long i = 0; i < nmine; ++i) {
            phase_bits ^= 1u << buf;
            const uint32_t bp = ring_u32 + buf * C::BUF_BYTES;
            const uint32_t tb = term_u32 + buf * 512;
            const int nbuf = (buf + 1 == NBUF) ? 0 : buf + 1;
            const uint32_t npar = (phase_bits >> nbuf) & 1u;
            const bool has_next = i + 1 < nmine;
            bool nready = false;
#pragma unroll
            for (int s = 0; s < 8; s += 2) {
                // even k-step: DMMAs on set A, prefetch k-step s+1 into set B
                load_kstep<NB16>(avB, rvB, wvB, bp, tb, s + 1, g, t);
                mma_kstep<NB8>(acc, gacc, avA, bvA, rvA, avB, bvB, wvB);
                if (s == 4 && has_next) {
                    // probe the next block's barriers early; the answer returns behind DMMAs
                    nready = mbar_try(tbar_u32 + nbuf * 8, npar);
                    nready = mbar_try(full_u32 + nbuf * 8, npar) && nready;
                }
                // odd k-step: DMMAs on set B, prefetch k-step s+2 (or the next block's first) into set A
                if (s + 2 < 8) {
                    load_kstep<NB16>(avA, rvA, wvA, bp, tb, s + 2, g, t);
                } else {
                    // this block now lives entirely in registers: hand its buffer back to TMA
                    __syncwarp();
                    if (lane == 0 && i + NBUF < nmine) issue(i + NBUF, buf);
                    if (has_next && !nready) {
                        mbar_wait(tbar_u32 + nbuf * 8, npar);
                        mbar_wait(full_u32 + nbuf * 8, npar);
                    }
                    // (without a next block this reads stale but valid shared memory; the values are never used)
                    load_kstep<NB16>(avA, rvA, wvA, ring_u32 + nbuf * C::BUF_BYTES, term_u32 + nbuf * 512, 0, g, t);
                }
                mma_kstep<NB8>(acc, gacc, avB, bvB, rvB, avA, bvA, wvA);
            }
            buf = nbuf;
        }
        // every team is done with its ring: reuse it as reduction scratch
        asm volatile("bar.sync 1, %0;" ::"n"(C::THREADS) : "memory");
#pragma unroll
        for (int e = 0; e < 2 * NT; ++e) scratch[e * 32 + lane] = acc[e];
#pragma unroll
        for (int jb = 0; jb < NB8; ++jb) scratch[(2 * NT + jb) * 32 + lane] = gacc[jb];
    } else {
        // ============================ helper warps ===========================
        if (NHELP == 2) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(C::REGS_H));
        const int hid = role - 1;  // handles the team's blocks i with (i % HELPERS) == hid
        double objacc = 0.0;
        bool bad_domain = false;
        // lane (g,t) ends up owning the row of k-step s == g, sub-row t
        const int R = 16 * (g >> 2) + 8 * ((g >> 1) & 1) + (g & 1) + 2 * t;
        const uint32_t bsel = beta_base + 2 * g * 8;
        for (long long i = hid; i < nmine; i += C::HELPERS) {
            const int buf = (int)(i % NBUF);
            const uint32_t par = (uint32_t)((i / NBUF) & 1);
            mbar_wait(full_u32 + buf * 8, par);
            const uint32_t bp = ring_u32 + buf * C::BUF_BYTES;
            double part[8];
#pragma unroll
            for (int s = 0; s < 8; ++s) {
                const int xr = 2 * t + (s & 1);
                const int row = 16 * (s >> 2) + 8 * ((s >> 1) & 1) + xr;
                const uint32_t rp = bp + row * 128 + ((g ^ xr) << 4);
                double d = 0.0;
#pragma unroll
                for (int j = 0; j < NB16; ++j) {
                    const double2 v = lds128(rp + j * 4096);
                    const double2 b = lds128(bsel + j * 128);
                    d = fma(v.x, b.x, d);
                    d = fma(v.y, b.y, d);
                }
                part[s] = d;
            }
            double k4[4], k2[2], y;
            {
                const bool hi = (g & 4) != 0;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const double send = hi ? part[q] : part[q + 4];
                    const double keep = hi ? part[q + 4] : part[q];
                    k4[q] = keep + shfl_xor_d(send, 16);
                }
            }
            {
                const bool hi = (g & 2) != 0;
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const double send = hi ? k4[q] : k4[q + 2];
                    const double keep = hi ? k4[q + 2] : k4[q];
                    k2[q] = keep + shfl_xor_d(send, 8);
                }
            }
            {
                const bool hi = (g & 1) != 0;
                const double send = hi ? k2[0] : k2[1];
                const double keep = hi ? k2[1] : k2[0];
                y = keep + shfl_xor_d(send, 4);
            }
            const long long row = (gp + i * GP) * 32 + R;
            double rr = 0.0, ww = 0.0;
            if (row < a.n) {
                if (a.offset) y += a.offset[row];
                const double ob = a.obs[row];
                const double se = a.se ? a.se[row] : 1.0;
                const RowTerms1 rt = row_terms1(a.family, a.link, a.fast != 0, y, ob, se);
                objacc += rt.l;
                rr = rt.r;
                ww = rt.w;
                bad_domain |= !rt.ok;
            }
            // lane == 4*s + t of the row it owns
            sts64(term_u32 + buf * 512 + lane * 8, rr);
            sts64(term_u32 + buf * 512 + 256 + lane * 8, ww);
            __syncwarp();
            if (lane == 0) mbar_arrive(tbar_u32 + buf * 8);
        }
        if (bad_domain) atomicOr(a.status, 1);
        sts64(obj_base + ((team * C::HELPERS + hid) * 32 + lane) * 8, objacc);
        asm volatile("bar.sync 1, %0;" ::"n"(C::THREADS) : "memory");
    }
    __syncthreads();
    double* out = a.partials + (size_t)blockIdx.x * C::EL;
    for (int e = threadIdx.x; e < C::EL; e += C::THREADS) {
        double s = 0.0;
        if (e < (2 * NT + NB8) * 32) {
#pragma unroll
            for (int q = 0; q < TEAMS; ++q) s += reinterpret_cast<const double*>(smem + q * C::TEAM_BYTES)[e];
        } else {
            const int l = e - (2 * NT + NB8) * 32;
#pragma unroll
            for (int q = 0; q < TEAMS * C::HELPERS; ++q) s += lds64(obj_base + (q * 32 + l) * 8);
        }
        out[e] = s;
    }
}

}  // namespace anml
