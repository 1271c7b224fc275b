// Warp-specialised fused evaluator: objective + gradient + Hessian in one pass
// over X for one Parameter with p <= 64 columns (the north-star configuration).
//
// Measured facts that shape it (tools/microbench/, profiles/):
//   * DMMA.8x8x4 and scalar FP64 share one pipe; one warp per scheduler streaming
//     independent DMMAs reaches 36.5 TFLOP/s, TWO DMMA-issuing warps on a scheduler
//     drop to 21 TFLOP/s, and every FP64 instruction inside the DMMA stream costs
//     5-7 pipe cycles (vs ~1.5-3 when it comes from another warp);
//   * so the tensor warp must issue nothing but LDS + DMMA, and there must be exactly
//     one of it per scheduler.
//
// Structure: one persistent CTA of 12 warps per SM = 4 teams (one per scheduler):
// tensor warp M (warp q) + helper warps H0, H1 (warps q+4, q+8); setmaxnreg moves
// registers from helpers to M.  Each team owns row blocks b = gp, gp+GP, ... (32 rows,
// 64 for designs of at most 32 columns) and a private ring of NBUF block buffers
// ([column group][row][16 doubles], 128-byte swizzle, zero fill outside the matrix),
// filled by one 3-D TMA per block when p is a multiple of 16, else by one 2-D box per
// 16-column group.
//   * helper H(i&1), block i: waits for the TMA bytes; pass 1: y = X beta on
//     MMA-layout fragments, partial sums transposed through shared memory, then link +
//     family terms, one row per lane (the logistic log(1 + e^-y) is not taken per row:
//     rowmath.cuh LogProduct); pass 2 over the same block: g += r_k x_k and the block is
//     rescaled IN PLACE to sqrt(|w_k|) x_k; publishes sign(w) and then the sequence
//     flag ready[b] = i + 1.
//   * tensor warp M: X~ = sqrt(|w|) X makes H = X~^T X~, so A and B fragments are the
//     same registers: per 4-row k-step 4 LDS.128 and 36 DMMAs (tiles on/below the
//     diagonal at p = 64), software-pipelined over two fragment sets.  (Non-canonical
//     link/family pairs can have w < 0: then B = -A for that row, an integer XOR on
//     the sign bit -- still no FP64-pipe instruction.)  It reads ready[b] of the next
//     block with a plain LDS two k-steps early (an mbarrier wait there cost 3 %), and
//     when a block is entirely in registers it re-arms TMA for its buffer.
//
// Ordering: M starts block i only after the helper published ready[b] for it, which
// is after the helper's last access to the buffer (threadfence_block before the flag;
// M's first fragment address carries a dependency on the flag value); so when M
// re-issues TMA nobody touches the buffer (fence.proxy.async orders the helper's
// generic-proxy stores before the async-proxy refill).  Helpers observe every phase of
// every full[] barrier, also for blocks of the other helper (a parity wait cannot tell
// phases two apart).
//
// Partial-record layout is the one of fused_eval_kernel (finish_fused_kernel, finish.cuh,
// is the shared one-launch epilogue).
#pragma once
#include "common.cuh"
#include "rowmath.cuh"
#include "fused_eval.cuh"

namespace anml {

// ROWS = rows per block: 32, or 64 for narrow designs (p <= 32), where the per-block overhead of the
// tensor warp (TMA re-arm, flag hand-off: ~700 cycles) would otherwise rival its DMMA time
// (10 tiles x 8 k-steps x 16 cycles = 1280 at p = 32).  A 64-row block is two 32-row halves for the helpers.
template <int NB16, int NHELP = 2, int ROWS = 32>
struct HessCfg {
    static constexpr int TEAMS = 4;
    static constexpr int HELPERS = NHELP;  // per team
    static constexpr int WARPS = TEAMS * (1 + HELPERS);
    static constexpr int THREADS = WARPS * 32;
    // 384 threads x 168 registers at launch = 4*32*216 (tensor warps) + 8*32*144 (helpers)
    static constexpr int REGS_M = 216, REGS_H = 144;
    static constexpr int NB8 = 2 * NB16;
    static constexpr int NT = NB8 * (NB8 + 1) / 2;
    static constexpr int HALVES = ROWS / 32;
    static constexpr int KSTEPS = ROWS / 4;
    static constexpr int GROUP_BYTES = ROWS * 128;             // one 16-column group of the block
    static constexpr int BUF_BYTES = NB16 * GROUP_BYTES;       // NB16 boxes of ROWS rows x 128 B
    static constexpr int NBUF = (49152 / BUF_BYTES) > 8 ? 8 : (49152 / BUF_BYTES);
    static constexpr int TEAM_BYTES = NBUF * BUF_BYTES;
    static constexpr int EL = (2 * NT + NB8 + 1) * 32;
    static constexpr int SMEM_BYTES = 1024 + TEAMS * TEAM_BYTES + TEAMS * NBUF * ROWS * 4 /*sign masks*/ + TEAMS * NBUF * 2 * 8 /*full barriers, ready flags*/ +
                                      2 * 16 * NB16 * 8 /*beta, beta1*/ + TEAMS * HELPERS * (NB8 + 1) * 32 * 8 /*helper g, obj; transpose scratch*/ +
                                      TEAMS * HELPERS * 64 * 8 /*helper r, sqrt|w|*/;
};

// DMMA without `volatile`: ptxas may hoist the next k-step's LDS above it
__device__ __forceinline__ void dmma884_nv(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
// flag read used for synchronisation: "memory" keeps the compiler from caching or dropping it
__device__ __forceinline__ uint32_t lds32_sync(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, double x, double y) {
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(x), "d"(y) : "memory");
}
__device__ __forceinline__ double flip_sign(double v, uint32_t mask) {  // mask = 0 or 0x80000000
    return __hiloint2double(__double2hiint(v) ^ (int)mask, __double2loint(v));
}

// fragments of one 4-row k-step of the rescaled block: av[jb] = X~[row(t)][col(jb, g)]
template <int NB16, int GROUP_BYTES = 4096>
__device__ __forceinline__ void load_kstep(double* av, uint32_t bp, int s, int g, int t) {
    // k-step s holds rows 16*(s>>2) + 8*((s>>1)&1) + (s&1) + 2t, t = 0..3
    const int xr = 2 * t + (s & 1);
    const int row = 16 * (s >> 2) + 8 * ((s >> 1) & 1) + xr;
    const uint32_t rp = bp + row * 128 + ((g ^ xr) << 4);
#pragma unroll
    for (int j = 0; j < NB16; ++j) {
        const double2 v = lds128(rp + j * GROUP_BYTES);
        av[2 * j] = v.x;
        av[2 * j + 1] = v.y;
    }
}

// 36 (at NB8 = 8) DMMAs of one k-step: C[I][J] += a[I] (x) +-a[J]
template <int NB8, bool SIGNED>
__device__ __forceinline__ void mma_kstep(double* acc, const double* av, uint32_t sign_mask) {
    double bv[NB8];
#pragma unroll
    for (int J = 0; J < NB8; ++J) bv[J] = SIGNED ? flip_sign(av[J], sign_mask) : av[J];
#pragma unroll
    for (int I = 0; I < NB8; ++I) {
#pragma unroll
        for (int J = 0; J <= I; ++J) {
            const int tile = I * (I + 1) / 2 + J;
            dmma884_nv(acc[2 * tile], acc[2 * tile + 1], av[I], bv[J]);
        }
    }
}

// MODE 0: per-row terms computed here from X, beta, obs (one-parameter models);
// MODE 1: (r, w) read from a.rvec / a.wvec (blocks of the general path; rvec == nullptr: no gradient);
// MODE 3: two parameters over this one design (mean / variance model, BASELINE config #4): the helpers
//         form BOTH linear predictors, take the row terms of the pair once, accumulate the objective,
//         g0 = X^T r0 and H00 = X^T diag(w00) X here and leave (r1, w01, w11) in a.out_* for the two
//         MODE 1 launches that follow -- the separate dual-predictor pass of round 1 (a fourth read of
//         X) is gone.
// TMA3D: the tensor map is the 3-D view (16 cols, rows, column groups) and one instruction
// fetches the whole block (api.cu encodes it when p is a multiple of 16).
template <int NB16, int NHELP, bool SIGNED, int MODE, bool TMA3D, int ROWS>
__global__ void __launch_bounds__(128 * (1 + NHELP), 1) fused_hess_kernel(const __grid_constant__ CUtensorMap tmap32, const FusedArgs a) {
    using C = HessCfg<NB16, NHELP, ROWS>;
    constexpr int KSTEPS = C::KSTEPS, GB = C::GROUP_BYTES, SIGN_BYTES = ROWS * 4;
    constexpr int NB8 = C::NB8, NT = C::NT, NBUF = C::NBUF, TEAMS = C::TEAMS;

    extern __shared__ unsigned char smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* smem = smem_raw + (smem_base - smem_u32(smem_raw));
    // layout: rings | sign masks | barriers | beta | helper records (gradient rows, objective row)
    const uint32_t signs_base = smem_base + TEAMS * C::TEAM_BYTES;
    const uint32_t bars_base = signs_base + TEAMS * NBUF * SIGN_BYTES;
    const uint32_t beta_base = bars_base + TEAMS * NBUF * 2 * 8;
    const uint32_t hrec_base = beta_base + 2 * 16 * NB16 * 8;  // [beta | beta1 (MODE 3)]
    const uint32_t hterm_base = hrec_base + TEAMS * C::HELPERS * (NB8 + 1) * 256;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const int team = warp & 3;
    const int role = warp >> 2;  // 0 = tensor warp, 1.. = helpers

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap32);
        for (int q = 0; q < TEAMS; ++q)
            for (int i = 0; i < NBUF; ++i) {
                mbar_init(bars_base + (q * NBUF * 2 + i) * 8, 1);
                asm volatile("st.shared.u32 [%0], %1;" ::"r"(bars_base + (q * NBUF * 2 + NBUF + i) * 8), "r"(0) : "memory");
            }
        mbar_fence_init();
    }
    if (threadIdx.x < 16 * NB16) {
        sts64(beta_base + threadIdx.x * 8, (MODE == 0 || MODE == 3) ? a.beta[threadIdx.x] : 0.0);
        sts64(beta_base + (16 * NB16 + threadIdx.x) * 8, MODE == 3 ? a.beta1[threadIdx.x] : 0.0);
    }
    __syncthreads();

    const uint32_t ring_u32 = smem_base + team * C::TEAM_BYTES;
    const uint32_t sign_u32 = signs_base + team * NBUF * SIGN_BYTES;    // [buf][ROWS] sign bit of w per row
    const uint32_t full_u32 = bars_base + team * NBUF * 2 * 8;          // full[b]  : TMA bytes landed
    const uint32_t rdy_u32 = full_u32 + NBUF * 8;                       // ready[b] : u32 sequence flag, block rescaled + signs published

    const long long nblk = (a.n + ROWS - 1) / ROWS;
    const long long gp = (long long)blockIdx.x * TEAMS + team;
    const long long GP = (long long)gridDim.x * TEAMS;
    const long long nmine = (nblk > gp) ? (nblk - gp + GP - 1) / GP : 0;

    double* scratch = reinterpret_cast<double*>(smem + team * C::TEAM_BYTES);  // epilogue: 2*NT*32 doubles per team
    static_assert(2 * NT * 32 * 8 <= C::TEAM_BYTES, "reduction scratch must fit in the team's ring");

    if (role == 0) {
        // ============================ tensor warp ============================
        if (NHELP == 2) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(C::REGS_M));
        auto issue = [&](long long blk_seq, int buf) {
            const long long row0 = (gp + blk_seq * GP) * ROWS;
            const uint32_t bar = full_u32 + buf * 8;
            const uint32_t dst = ring_u32 + buf * C::BUF_BYTES;
            mbar_arrive_expect_tx(bar, C::BUF_BYTES);
            if (TMA3D) {
                tma_load_3d(dst, &tmap32, 0, (int)row0, 0, bar);
            } else {
#pragma unroll
                for (int j = 0; j < NB16; ++j) tma_load_2d(dst + j * GB, &tmap32, 16 * j, (int)row0, bar);
            }
        };
        if (lane == 0) {
            const int pre = nmine < NBUF ? (int)nmine : NBUF;
            for (int i = 0; i < pre; ++i) issue(i, i);
        }
        double acc[2 * NT];
#pragma unroll
        for (int i = 0; i < 2 * NT; ++i) acc[i] = 0.0;

        // Two fragment sets: while the DMMAs of k-step k run on one, k-step k+1 is loaded
        // into the other (8 k-steps per block, even, so the sets alternate without moves).
        double avA[NB8], avB[NB8];
        uint32_t sgA = 0, sgB = 0;
        uint32_t phase_bits = 0;
        int buf = 0;
        // The helpers publish "block i is rescaled" as a sequence number in shared memory
        // (ready[b] = i + 1).  The tensor warp must never execute a long-latency dependent
        // instruction (a stall longer than one DMMA is a pipe bubble), so it reads the flag
        // with a plain LDS two k-steps ahead, like a fragment, instead of an mbarrier wait.
        if (nmine > 0) {
            uint32_t f0;
            do { f0 = lds32_sync(rdy_u32); } while (f0 != 1u);
            // (f0 ^ 1) == 0: an address dependency, so no fragment load can be issued before the flag arrived
            load_kstep<NB16, GB>(avA, ring_u32 + (f0 ^ 1u), 0, g, t);
            if (SIGNED) sgA = lds32(sign_u32 + (f0 ^ 1u) + t * 4);
        }
        for (long long i = 0; i < nmine; ++i) {
            phase_bits ^= 1u << buf;
            const uint32_t bp = ring_u32 + buf * C::BUF_BYTES;
            const uint32_t sb = sign_u32 + buf * SIGN_BYTES;
            const int nbuf = (buf + 1 == NBUF) ? 0 : buf + 1;
            const bool has_next = i + 1 < nmine;
            const uint32_t want_flag = (uint32_t)(i + 2);  // block i+1 published
            uint32_t seen_flag = 0;
#pragma unroll
            for (int s = 0; s < KSTEPS; s += 2) {
                // even k-step on set A, prefetch k-step s+1 into set B
                load_kstep<NB16, GB>(avB, bp, s + 1, g, t);
                if (SIGNED) sgB = lds32(sb + (4 * (s + 1) + t) * 4);
                mma_kstep<NB8, SIGNED>(acc, avA, sgA);
                if (s == KSTEPS - 4) seen_flag = lds32_sync(rdy_u32 + nbuf * 8);  // consumed two k-steps later
                // odd k-step on set B, prefetch k-step s+2 (or the next block's first) into set A
                if (s + 2 < KSTEPS) {
                    load_kstep<NB16, GB>(avA, bp, s + 2, g, t);
                    if (SIGNED) sgA = lds32(sb + (4 * (s + 2) + t) * 4);
                } else {
                    // this block now lives entirely in registers: hand its buffer back to TMA
                    __syncwarp();
                    if (lane == 0 && i + NBUF < nmine) {
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // helper's in-place stores before the refill
                        issue(i + NBUF, buf);
                    }
                    if (has_next) {
                        while (seen_flag != want_flag) seen_flag = lds32_sync(rdy_u32 + nbuf * 8);
                    } else {
                        seen_flag = want_flag;
                    }
                    // (seen ^ want) == 0: address dependency on the flag (see prologue).  Without a next
                    // block this reads stale but valid shared memory; the values are never used.
                    const uint32_t dep = seen_flag ^ want_flag;
                    load_kstep<NB16, GB>(avA, ring_u32 + nbuf * C::BUF_BYTES + dep, 0, g, t);
                    if (SIGNED) sgA = lds32(sign_u32 + nbuf * SIGN_BYTES + dep + t * 4);
                }
                mma_kstep<NB8, SIGNED>(acc, avB, sgB);
            }
            buf = nbuf;
        }
        // every team is done with its ring: reuse it as reduction scratch
        asm volatile("bar.sync 1, %0;" ::"n"(C::THREADS) : "memory");
#pragma unroll
        for (int e = 0; e < 2 * NT; ++e) scratch[e * 32 + lane] = acc[e];
    } else {
        // ============================ helper warps ===========================
        if (NHELP == 2) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(C::REGS_H));
        const int hid = role - 1;  // handles the team's blocks i with (i % HELPERS) == hid
        double objacc = 0.0;
        LogProduct logprod;
        double gacc[NB8];
#pragma unroll
        for (int i = 0; i < NB8; ++i) gacc[i] = 0.0;
        bool bad_domain = false;
        const bool want_grad = (MODE != 1) || a.rvec != nullptr;
        // lane (g,t) ends up owning the row of k-step s == g, sub-row t
        const int R = 16 * (g >> 2) + 8 * ((g >> 1) & 1) + (g & 1) + 2 * t;
        const uint32_t tr_u32 = hrec_base + (team * C::HELPERS + hid) * (NB8 + 1) * 256;  // transpose scratch (epilogue record later)
        const uint32_t ht_u32 = hterm_base + (team * C::HELPERS + hid) * 512;             // [0..31] r, [32..63] sqrt|w| of the block
        // this lane's slice of beta (columns 16j + 2g, +1) stays in registers
        double2 bsl[NB16];
#pragma unroll
        for (int j = 0; j < NB16; ++j) bsl[j] = lds128(beta_base + 2 * g * 8 + j * 128);  // (MODE 3: beta1 is re-read from shared memory per k-step)
        uint32_t phase_bits = 0;
        int buf = -1;
        for (long long i = 0; i < nmine; ++i) {
            buf = (buf + 1 == NBUF) ? 0 : buf + 1;
            // observe EVERY phase of every full[] barrier (see header)
            mbar_wait(full_u32 + buf * 8, (phase_bits >> buf) & 1u);
            phase_bits ^= 1u << buf;
            if ((int)(i % C::HELPERS) != hid) continue;
#pragma unroll 1
            for (int hh = 0; hh < C::HALVES; ++hh) {  // a 64-row block is handled as two 32-row halves
            const long long row = (gp + i * GP) * ROWS + 32 * hh + R;
            const bool valid = row < a.n;
            double rr = 0.0, ww = 0.0, sc = 0.0;
            const uint32_t bp = ring_u32 + buf * C::BUF_BYTES + hh * 4096;
            if (MODE == 0) {
                // per-row global loads first: their latency hides behind the dot products
                double off = 0.0, ob = 0.0, se = 1.0;
                if (valid) {
                    if (a.offset) off = a.offset[row];
                    ob = a.obs[row];
                    if (a.se) se = a.se[row];
                }
                // ---- pass 1: linear predictor
                double part[8];
#pragma unroll
                for (int s = 0; s < 8; ++s) {
                    const int xr = 2 * t + (s & 1);
                    const int rowi = 16 * (s >> 2) + 8 * ((s >> 1) & 1) + xr;
                    const uint32_t rp = bp + rowi * 128 + ((g ^ xr) << 4);
                    double d = 0.0;
#pragma unroll
                    for (int j = 0; j < NB16; ++j) {
                        const double2 v = lds128(rp + j * GB);
                        d = fma(v.x, bsl[j].x, d);
                        d = fma(v.y, bsl[j].y, d);
                    }
                    part[s] = d;
                }
                // 8 partial sums per lane -> one row total per lane, through shared memory: lane (g,t)
                // stores part[s] at [s][g][t] (row pitch 36 doubles: conflict-free), then lane (s'=g, t)
                // adds the 8 entries [g][*][t].  24 instructions with ILP instead of ~60 dependent
                // shuffle/select instructions (every helper instruction costs the tensor warp cycles).
                double y;
                if (NB8 == 8) {
#pragma unroll
                    for (int s = 0; s < 8; ++s) sts64(tr_u32 + (s * 36 + lane) * 8, part[s]);
                    __syncwarp();
                    double q0 = lds64(tr_u32 + (g * 36 + 0 * 4 + t) * 8), q1 = lds64(tr_u32 + (g * 36 + 1 * 4 + t) * 8);
                    double q2 = lds64(tr_u32 + (g * 36 + 2 * 4 + t) * 8), q3 = lds64(tr_u32 + (g * 36 + 3 * 4 + t) * 8);
                    double q4 = lds64(tr_u32 + (g * 36 + 4 * 4 + t) * 8), q5 = lds64(tr_u32 + (g * 36 + 5 * 4 + t) * 8);
                    double q6 = lds64(tr_u32 + (g * 36 + 6 * 4 + t) * 8), q7 = lds64(tr_u32 + (g * 36 + 7 * 4 + t) * 8);
                    y = ((q0 + q1) + (q2 + q3)) + ((q4 + q5) + (q6 + q7));
                    __syncwarp();
                } else {
                    y = transpose_reduce8(part, g);
                }
                if (valid) {
                    if (a.weights) {  // per-row likelihood weight (>= 0): scales l, r, w
                        const RowTerms1 rt = row_terms1<true>(a.family, a.link, (a.fast & 1) != 0, y + off, ob, se);
                        bad_domain |= !rt.ok;
                        const double wt = a.weights[row];
                        objacc += wt * rt.l;
                        rr = wt * rt.r;
                        ww = wt * rt.w;
                        sc = sqrt(wt) * rt.sw;
                    } else {
                        // the logistic log(1 + e^-y) terms are multiplied up and logged once at the end
                        const RowTerms1 rt = row_terms1<true, true>(a.family, a.link, (a.fast & 1) != 0, y + off, ob, se);
                        bad_domain |= !rt.ok;
                        objacc += rt.l;
                        logprod.mul(rt.ls);
                        rr = rt.r;
                        ww = rt.w;
                        sc = rt.sw;
                    }
                }
            } else if (MODE == 3) {
                double off0 = 0.0, off1 = 0.0, ob = 0.0;
                if (valid) {
                    if (a.offset) off0 = a.offset[row];
                    if (a.offset1) off1 = a.offset1[row];
                    ob = a.obs[row];
                }
                // ---- pass 1: both linear predictors from one sweep over the block
                double part0[8], part1[8];
#pragma unroll
                for (int s = 0; s < 8; ++s) {
                    const int xr = 2 * t + (s & 1);
                    const int rowi = 16 * (s >> 2) + 8 * ((s >> 1) & 1) + xr;
                    const uint32_t rp = bp + rowi * 128 + ((g ^ xr) << 4);
                    double d0 = 0.0, d1 = 0.0;
#pragma unroll
                    for (int j = 0; j < NB16; ++j) {
                        const double2 v = lds128(rp + j * GB);
                        const double2 b1 = lds128(beta_base + (16 * NB16 + 2 * g) * 8 + j * 128);
                        d0 = fma(v.x, bsl[j].x, d0);
                        d0 = fma(v.y, bsl[j].y, d0);
                        d1 = fma(v.x, b1.x, d1);
                        d1 = fma(v.y, b1.y, d1);
                    }
                    part0[s] = d0;
                    part1[s] = d1;
                }
                const double y0 = transpose_reduce8(part0, g) + off0;
                const double y1 = transpose_reduce8(part1, g) + off1;
                if (valid) {
                    const RowTerms2 rt = row_terms2(a.link, a.link1, y0, y1, ob, a.fast != 0);
                    bad_domain |= !rt.ok;
                    const double wt = a.weights ? a.weights[row] : 1.0;
                    objacc += wt * rt.l;
                    rr = wt * rt.r0;
                    ww = wt * rt.w00;
                    sc = (rt.sw00 >= 0.0) ? (a.weights ? sqrt(wt) * rt.sw00 : rt.sw00) : sqrt(fabs(ww));
                    a.out_r1[row] = wt * rt.r1;
                    a.out_w01[row] = wt * rt.w01;
                    a.out_w11[row] = wt * rt.w11;
                }
            } else if (valid) {
                if (a.rvec) rr = a.rvec[row];
                ww = a.wvec[row];
                sc = sqrt(fabs(ww));
            }
            if (SIGNED) asm volatile("st.shared.u32 [%0], %1;" ::"r"(sign_u32 + buf * SIGN_BYTES + (32 * hh + lane) * 4), "r"(ww < 0.0 ? 0x80000000u : 0u) : "memory");
            // ---- pass 2: gradient, and rescale the block in place to sqrt(|w|) x
            // (r, sqrt|w|) of the 32 rows go through shared memory: 2 stores + broadcast loads
            // instead of 32 shuffles
            sts64(ht_u32 + lane * 8, rr);
            sts64(ht_u32 + 256 + lane * 8, sc);
            __syncwarp();
#pragma unroll
            for (int s = 0; s < 8; ++s) {
                const int xr = 2 * t + (s & 1);
                const int rowi = 16 * (s >> 2) + 8 * ((s >> 1) & 1) + xr;
                const uint32_t rp = bp + rowi * 128 + ((g ^ xr) << 4);
                const double rv = lds64(ht_u32 + (4 * s + t) * 8);
                const double sv = lds64(ht_u32 + 256 + (4 * s + t) * 8);
#pragma unroll
                for (int j = 0; j < NB16; ++j) {
                    const double2 v = lds128(rp + j * GB);
                    if (MODE != 1 || want_grad) {  // (MODE 1, cross block of two parameters: Hessian only)
                        gacc[2 * j] = fma(rv, v.x, gacc[2 * j]);
                        gacc[2 * j + 1] = fma(rv, v.y, gacc[2 * j + 1]);
                    }
                    sts128(rp + j * GB, sv * v.x, sv * v.y);
                }
            }
            __syncwarp();
            }  // halves
            if (lane == 0) {
                __threadfence_block();  // the block's stores before the flag
                asm volatile("st.shared.u32 [%0], %1;" ::"r"(rdy_u32 + buf * 8), "r"((uint32_t)(i + 1)) : "memory");
            }
        }
        if (bad_domain) atomicOr(a.status, 1);
        {
            const uint32_t hb = hrec_base + (team * C::HELPERS + hid) * (NB8 + 1) * 256;
#pragma unroll
            for (int jb = 0; jb < NB8; ++jb) sts64(hb + (jb * 32 + lane) * 8, gacc[jb]);
            sts64(hb + (NB8 * 32 + lane) * 8, objacc + logprod.log_value());
        }
        asm volatile("bar.sync 1, %0;" ::"n"(C::THREADS) : "memory");
    }
    __syncthreads();
    double* out = a.partials + (size_t)blockIdx.x * C::EL;
    for (int e = threadIdx.x; e < C::EL; e += C::THREADS) {
        double s = 0.0;
        if (e < 2 * NT * 32) {
#pragma unroll
            for (int q = 0; q < TEAMS; ++q) s += reinterpret_cast<const double*>(smem + q * C::TEAM_BYTES)[e];
        } else {
            const int l = e - 2 * NT * 32;  // helper record: NB8 gradient rows, then the objective row
#pragma unroll
            for (int q = 0; q < TEAMS * C::HELPERS; ++q) s += lds64(hrec_base + (q * (NB8 + 1) * 32 + l) * 8);
        }
        out[e] = s;
    }
}

}  // namespace anml
