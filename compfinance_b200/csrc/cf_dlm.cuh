// cf_dlm.cuh -- multi-asset displaced-lognormal model x {Autocall, Baskets, MultiStats}: value and AAD.
//
// Replaces MultiDisplaced::generatePath (mcMdlMultiDisplaced.h:653-717) + Autocall / Baskets /
// MultiStats::payoffs (mcPrdMulti.h:388-439, 274-286, 141-178) under the loops of mcBase.h:378-386 /
// 680-704, and the per-path tape sweep on the AAD side.  Adjoint equations: SURVEY.md Appendix A.3.
//
// One path per thread, RNG + generatePath + payoffs fused.  Per step the thread draws A Gaussians
// (step-major dimension i * A + k), correlates them with the lower Cholesky factor, and advances
// every asset with one of the four schemes.  The reverse sweep reads the spots, the Gaussians and
// the alive notional back from a per-thread history ([row][slot], coalesced) and accumulates the
// adjoints of every init() table.  Sums over the 32 paths of a warp go through a per-warp scratch
// block of rows [value][lane]: every lane writes its values, then lane r adds row r over the 32
// columns from a rotated start (conflict free, fixed order) -- one pass for the 3 A per-step table
// adjoints of a step instead of one shuffle tree per value.  The Cholesky adjoint
// cholBar[k][j] = sum over paths and steps of cwBar_k w_j is an outer product over the same rows:
// lane p owns the pairs p, p + 32, ... and keeps their sums in registers over the whole run.  Spot
// and alpha adjoints are thread-local and reduced once at the end.  All orders are fixed by lane /
// warp / block index: results are bit-reproducible.
#pragma once

#include "cf_kernels.cuh"

namespace cf {

struct LArgs {
    uint64_t first_path, n_paths;
    int      n_batches;
    uint32_t seed1, seed2;
    int      dim;                  // n_steps * n_assets
    const uint32_t* sobol_dir;     // [32][dim]
    const uint64_t* mrg_jump;
    // model (mcMdlMultiDisplaced.h:474-606)
    int      A, D, E, today;       // assets, steps, events; today = 1 when timeline point 0 is an event date
    const double*  spots;          // [A]
    const double*  chol;           // [A][A] lower
    double         cholv[136];     // the same lower triangle, row k at k (k + 1) / 2: kernel parameters live in the constant bank,
                                   // so the correlation sums read their coefficients as instruction operands, not through loads
    const double*  alphas;         // [A]
    const int32_t* dyn;            // [A] 0 lognormal 1 normal 2 surnormal 3 subnormal
    const double*  dynFwd;         // [D][A]
    const double*  drifts;         // [D][A]
    const double*  stds;           // [D][A]
    const double*  ff;             // [E][A] forward factor of the first forward maturity
    const double*  num;            // [E] or null (numeraire not requested: Sample default 1)
    // product
    int      n_payoffs, n_strikes;
    double   strike, ko, smooth, coupon, cpn_dt;
    const double* strikes;         // Baskets [n_strikes]
    const double* pweights;        // Baskets weights / Autocall references [A]
    const double* w;               // payoff weights of the aggregate [n_payoffs] (AAD)
    // outputs
    double*  partial;              // [grid][partial_stride]: payoff sums, (agg, table adjoints)
    int      partial_stride;
    double*  per_path_payoffs;
    double*  per_path_agg;
    double*  hist;                 // [D][2 A + 1][grid * kBlock]
};

// Layout of the table-adjoint vector of the displaced model (after the aggregate):
//   spots [A] | alphas [A] | chol [A][A] | dynFwd [D][A] | drifts [D][A] | stds [D][A] | numeraires [E] | ff [E][A]
__host__ __device__ inline int dlm_adj_size(int A, int D, int E) { return 2 * A + A * A + 3 * D * A + E + E * A; }
__host__ __device__ inline int dlm_step_tables(int A, int D, int E) { return 3 * D * A + E + E * A; }

struct LSmemSizes { size_t pay, tab, red, gq, tagq, dirlow, base, scr, total; };

__host__ __device__ inline LSmemSizes dlm_smem(int A, int D, int E, int nPay, int dim, bool sobol, bool aad)
{
    LSmemSizes s{};
    s.pay = align16(sizeof(double) * kWarps * size_t(nPay));
    s.tab = aad ? align16(sizeof(double) * kWarps * size_t(dlm_step_tables(A, D, E))) : 0;
    s.red = align16(sizeof(double) * kWarps);
    s.gq = align16(sizeof(double) * kWarps * kChunk * 32);
    s.tagq = align16(sizeof(uint16_t) * kWarps * kChunk * 32);
    s.dirlow = sobol ? align16(sizeof(uint32_t) * size_t(dim) * kLowBits) : 0;
    s.base = sobol ? align16(sizeof(uint32_t) * 2 * size_t(dim)) : 0;
    s.scr = aad ? align16(sizeof(double) * kWarps * size_t(3 * A) * 32) : 0;     // per warp: [3 A][32] rows
    s.total = s.pay + s.tab + s.red + s.gq + s.tagq + s.dirlow + s.base + s.scr;
    return s;
}

// Sum of one scratch row (32 doubles, one per lane of the warp that wrote it) from a rotated start: the lanes of
// a half-warp read 16 different 8-byte banks.  Four partial sums, fixed order.
__device__ __forceinline__ double dlm_row_sum(const double* row, int lane)
{
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
    for (int c = 0; c < 32; c += 4) {
        s0 += row[(lane + c) & 31]; s1 += row[(lane + c + 1) & 31];
        s2 += row[(lane + c + 2) & 31]; s3 += row[(lane + c + 3) & 31];
    }
    return (s0 + s1) + (s2 + s3);
}

// Rounds of 32 (k, j <= k) pairs of the Cholesky adjoint owned by a lane
template <int AMAX> struct DlmPairs { static constexpr int kRounds = (AMAX * (AMAX + 1) / 2 + 31) / 32; };

// AAD with more than 8 assets: one block per SM (up to 255 registers: spot, Gaussian and adjoint vectors stay in registers)
template <int AMAX, int PRD, bool AAD, int RNGK>
__global__ void __launch_bounds__(kBlock, (AAD && AMAX > 8) ? 1 : 2) dlm_kernel(const LArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int A = a.A, D = a.D, E = a.E;
    constexpr bool kSobol = (RNGK == CF_RNG_SOBOL);
    const LSmemSizes z = dlm_smem(A, D, E, a.n_payoffs, a.dim, kSobol, AAD);
    unsigned char* p = smem_raw;
    double* payRows = reinterpret_cast<double*>(p);     p += z.pay;       // [kWarps][n_payoffs]
    double* tabRows = reinterpret_cast<double*>(p);     p += z.tab;       // [kWarps][nStepTab]
    double* red = reinterpret_cast<double*>(p);         p += z.red;
    double* gq = reinterpret_cast<double*>(p);          p += z.gq;
    uint16_t* tagq = reinterpret_cast<uint16_t*>(p);    p += z.tagq;
    uint32_t* dirlow = reinterpret_cast<uint32_t*>(p);  p += z.dirlow;
    uint32_t* base = reinterpret_cast<uint32_t*>(p);    p += z.base;
    double* scrAll = reinterpret_cast<double*>(p);                        // [kWarps][3 A][32] (AAD)

    const int nPay = a.n_payoffs;
    const int nStepTab = dlm_step_tables(A, D, E);
    // offsets inside a warp row of step tables
    const int oFwd = 0, oDrift = D * A, oStd = 2 * D * A, oNum = 3 * D * A, oFf = 3 * D * A + E;
    for (int i = tid; i < kWarps * nPay; i += kBlock) payRows[i] = 0.0;
    if (AAD) for (int i = tid; i < kWarps * nStepTab; i += kBlock) tabRows[i] = 0.0;
    if (kSobol) sobol_load_low(dirlow, a.sobol_dir, a.dim);
    __syncthreads();
    double* myPay = payRows + size_t(warp) * nPay;
    double* myTab = tabRows + size_t(warp) * nStepTab;

    GaussGen<RNGK> gen;
    gen.gq = gq + size_t(warp) * kChunk * 32;
    gen.tagq = tagq + size_t(warp) * kChunk * 32;
    gen.dirlow = dirlow; gen.base = base; gen.dim = a.dim;

    const size_t nSlots = size_t(gridDim.x) * kBlock;
    const size_t slot = size_t(blockIdx.x) * kBlock + tid;
    const int hRows = 2 * A + 1;                       // per step: spots after the step [A], Gaussians [A], alive before the event
    double aggSum = 0.0;
    // per-asset table adjoints, thread-local over all paths of the thread
    double spotBar[AMAX], alphaBar[AMAX];
    // Cholesky adjoint: pair p = k (k + 1) / 2 + j of round q is owned by lane p - 32 q of every warp
    constexpr int kRounds = DlmPairs<AMAX>::kRounds;
    double cholAcc[kRounds];
    int pairK[kRounds], pairJ[kRounds];
    double* scr = scrAll + size_t(warp) * (3 * A) * 32;
    if (AAD) {
#pragma unroll
        for (int k = 0; k < AMAX; ++k) { spotBar[k] = 0.0; alphaBar[k] = 0.0; }
#pragma unroll
        for (int q = 0; q < kRounds; ++q) {
            const int pp = lane + 32 * q;
            int k = 0;
            while ((k + 1) * (k + 2) / 2 <= pp) ++k;
            pairK[q] = k; pairJ[q] = pp - k * (k + 1) / 2;      // k >= A: no such pair
            cholAcc[q] = 0.0;
        }
    }
    const double sm2 = 2.0 * a.smooth;

    for (int batch = blockIdx.x; batch < a.n_batches; batch += gridDim.x) {
        const uint64_t pidx = uint64_t(batch) * kBlock + tid;
        const bool valid = pidx < a.n_paths;
        const uint64_t pabs = a.first_path + pidx;
        gen.sign = 1.0;
        if (kSobol) {
            const uint32_t n0 = uint32_t(a.first_path + uint64_t(batch) * kBlock + 1);
            const uint32_t H0 = n0 >> kLowBits;
            __syncthreads();
            sobol_block_base(base, a.sobol_dir, a.dim, H0);
            __syncthreads();
            gen.sob.init(uint32_t(pabs + 1), H0);
        } else {
            gen.mrg.init(a.seed1, a.seed2, pabs >> 1, a.mrg_jump);
            gen.sign = (pabs & 1ull) ? -1.0 : 1.0;
        }

        // ---- payoff accumulation helpers
        double agg = 0.0;
        auto emit = [&](int k, double v) {                 // payoff k of this path
            const double s = warp_sum(valid ? v : 0.0);
            if (lane == 0) myPay[k] += s;
            if (AAD) agg += a.w[k] * v;
            if (valid && a.per_path_payoffs) a.per_path_payoffs[pidx * nPay + k] = v;
        };

        // ---- forward
        double S[AMAX], Fprev[AMAX];
#pragma unroll
        for (int k = 0; k < AMAX; ++k) { S[k] = (k < A) ? __ldg(a.spots + k) : 0.0; Fprev[k] = 0.0; }
        double alive = 1.0, pay = 0.0;                     // Autocall state
        int payIdx = 0;                                    // MultiStats running payoff index

        auto observe = [&](int e) {                        // event date e with the current spots
            const double num = a.num ? __ldg(a.num + e) : 1.0;
            double F[AMAX];
#pragma unroll
            for (int k = 0; k < AMAX; ++k) F[k] = (k < A) ? S[k] * __ldg(a.ff + e * A + k) : 0.0;
            if (PRD == CF_PRODUCT_AUTOCALL) {
                double worst = div_fast(F[0], __ldg(a.pweights));
#pragma unroll
                for (int k = 1; k < AMAX; ++k)
                    if (k < A) { const double pf = div_fast(F[k], __ldg(a.pweights + k)); if (pf < worst) worst = pf; }
                pay += div_z(alive * a.coupon * a.cpn_dt, num);
                if (e < E - 1) {
                    const double f = fmin(1.0, fmax(0.0, (a.ko + a.smooth - worst) / 2 / a.smooth));
                    const double surv = alive * f;
                    pay += div_z(alive - surv, num);
                    alive = surv;
                } else {
                    pay += div_z(alive, num);
                    pay -= div_z(div_z(alive * fmax(a.strike - worst, 0.0), a.strike), num);
                }
            } else if (PRD == CF_PRODUCT_BASKETS) {
                double b = 0.0;
#pragma unroll
                for (int k = 0; k < AMAX; ++k) if (k < A) b += __ldg(a.pweights + k) * F[k];
                for (int k = 0; k < a.n_strikes; ++k) emit(k, div_z(fmax(b - __ldg(a.strikes + k), 0.0), num));
            } else {                                       // MultiStats: levels now, differences after all levels
                for (int a1 = 0; a1 < A; ++a1) emit(payIdx++, F[a1]);
                for (int a1 = 0; a1 < A; ++a1)
                    for (int a2 = 0; a2 <= a1; ++a2) emit(payIdx++, F[a1] * F[a2]);
                if (e > 0) {
                    const int perDate = A + A * (A + 1) / 2;
                    int q = E * perDate + (e - 1) * perDate;
                    for (int a1 = 0; a1 < A; ++a1) emit(q++, F[a1] - Fprev[a1]);
                    for (int a1 = 0; a1 < A; ++a1)
                        for (int a2 = 0; a2 <= a1; ++a2) emit(q++, (F[a1] - Fprev[a1]) * (F[a2] - Fprev[a2]));
                }
#pragma unroll
                for (int k = 0; k < AMAX; ++k) Fprev[k] = F[k];
            }
        };

        int e = 0;
        if (a.today) { observe(e); ++e; }
        for (int i = 0; i < D; ++i) {
            double w[AMAX];
            for (int k0 = 0; k0 < A; k0 += kChunk) {
                const int cnt = min(kChunk, A - k0);
                gen.fill(i * A + k0, cnt);
#pragma unroll
                for (int k = 0; k < kChunk; ++k)
                    if (k < cnt && k0 + k < AMAX) w[k0 + k] = gen.get(k);
            }
            if (AAD) {
                double* h = a.hist + (size_t(i) * hRows) * nSlots + slot;
                h[size_t(2 * A) * nSlots] = alive;
#pragma unroll
                for (int k = 0; k < AMAX; ++k) if (k < A) h[size_t(A + k) * nSlots] = w[k];
            }
#pragma unroll
            for (int k = 0; k < AMAX; ++k) {
                if (k < A) {
                    double cw = 0.0;
#pragma unroll
                    for (int j = 0; j <= k; ++j) cw += a.cholv[k * (k + 1) / 2 + j] * w[j];
                    const double fwd = S[k] * __ldg(a.dynFwd + i * A + k);
                    const double sd = __ldg(a.stds + i * A + k), dr = __ldg(a.drifts + i * A + k);
                    const int dyn = __ldg(a.dyn + k);
                    const double al = __ldg(a.alphas + k);
                    // exp_core: the library's accuracy (< 1 ulp) without its range handling; the exponent is a few standard deviations
                    if (dyn == 0) S[k] = fwd * exp_core(dr + sd * cw);
                    else if (dyn == 1) S[k] = fwd + sd * cw;
                    else if (dyn == 2) S[k] = (fwd + al) * exp_core(dr + sd * cw) - al;
                    else S[k] = (fwd - al) * exp_core(dr + sd * cw) + al;
                }
            }
            if (AAD) {
                double* h = a.hist + (size_t(i) * hRows) * nSlots + slot;
#pragma unroll
                for (int k = 0; k < AMAX; ++k) if (k < A) h[size_t(k) * nSlots] = S[k];
            }
            observe(e); ++e;
        }
        if (PRD == CF_PRODUCT_AUTOCALL) emit(0, pay);
        if (AAD) {
            if (valid) { aggSum += agg; if (a.per_path_agg) a.per_path_agg[pidx] = agg; }
        }

        // ---- reverse sweep
        if (AAD) {
            double Sbar[AMAX];
#pragma unroll
            for (int k = 0; k < AMAX; ++k) Sbar[k] = 0.0;
            const double paybar = (PRD == CF_PRODUCT_AUTOCALL) ? a.w[0] : 0.0;
            double alivebar = 0.0;                          // adjoint of the notional alive AFTER the current event
            // adjoint of the sample of event ev, spots Sev, notional alive before the event
            auto reverseEvent = [&](int ev, const double* Sev, double aliveBefore) {
                const double num = a.num ? __ldg(a.num + ev) : 1.0;
                double F[AMAX], Fbar[AMAX];
#pragma unroll
                for (int k = 0; k < AMAX; ++k) { F[k] = (k < A) ? Sev[k] * __ldg(a.ff + ev * A + k) : 0.0; Fbar[k] = 0.0; }
                double numbar = 0.0;
                if (PRD == CF_PRODUCT_AUTOCALL) {
                    double worst = div_fast(F[0], __ldg(a.pweights));
                    int am = 0;
#pragma unroll
                    for (int k = 1; k < AMAX; ++k)
                        if (k < A) { const double pf = div_fast(F[k], __ldg(a.pweights + k)); if (pf < worst) { worst = pf; am = k; } }
                    double worstbar;
                    if (ev < E - 1) {
                        const double q = (a.ko + a.smooth - worst) / 2 / a.smooth;
                        const double f = fmin(1.0, fmax(0.0, q));
                        // surv = alive f; pay += alive cpn dt / num + alive (1 - f) / num
                        const double fbar = alivebar * aliveBefore - div_z(paybar * aliveBefore, num);
                        const double qbar = (q > 0.0 && q < 1.0) ? fbar : 0.0;      // max(0, .) then min(1, .), strict (AADExpr.h:571-598)
                        worstbar = div_z(-qbar, sm2);
                        numbar = div_z(-paybar * (aliveBefore * a.coupon * a.cpn_dt + (aliveBefore - aliveBefore * f)), num * num);
                        alivebar = alivebar * f + paybar * a.coupon * a.cpn_dt / num + div_z(paybar * (1.0 - f), num);
                    } else {
                        const double put = fmax(a.strike - worst, 0.0);
                        worstbar = (a.strike - worst > 0.0) ? div_z(div_z(paybar * aliveBefore, a.strike), num) : 0.0;
                        numbar = div_z(-paybar * (aliveBefore * a.coupon * a.cpn_dt + aliveBefore - div_z(aliveBefore * put, a.strike)), num * num);
                        alivebar = paybar * (a.coupon * a.cpn_dt / num + 1.0 / num - div_z(div_z(put, a.strike), num));
                    }
                    const double fb = div_z(worstbar, __ldg(a.pweights + am));      // one division: only the worst performer carries the adjoint
#pragma unroll
                    for (int k = 0; k < AMAX; ++k) if (k == am) Fbar[k] = fb;
                } else if (PRD == CF_PRODUCT_BASKETS) {
                    double b = 0.0;
#pragma unroll
                    for (int k = 0; k < AMAX; ++k) if (k < A) b += __ldg(a.pweights + k) * F[k];
                    double bbar = 0.0;
                    for (int k = 0; k < a.n_strikes; ++k) {
                        const double x = b - __ldg(a.strikes + k);
                        if (x > 0.0) { bbar += div_z(a.w[k], num); numbar -= div_z(a.w[k] * x, num * num); }
                    }
#pragma unroll
                    for (int k = 0; k < AMAX; ++k) if (k < A) Fbar[k] = bbar * __ldg(a.pweights + k);
                }
                // forwards[a][0] = spot * ff (fillScen, mcMdlMultiDisplaced.h:628-640)
                // rows 0 .. A - 1: forward-factor adjoints, row A: numeraire adjoint
#pragma unroll
                for (int k = 0; k < AMAX; ++k) {
                    if (k < A) {
                        scr[k * 32 + lane] = valid ? Fbar[k] * Sev[k] : 0.0;
                        Sbar[k] += Fbar[k] * __ldg(a.ff + ev * A + k);
                    }
                }
                scr[A * 32 + lane] = valid ? numbar : 0.0;
                __syncwarp();
                if (lane <= A) {
                    const double s = dlm_row_sum(scr + lane * 32, lane);
                    if (lane < A) myTab[oFf + ev * A + lane] += s;
                    else if (a.num) myTab[oNum + ev] += s;
                }
                __syncwarp();
            };

            int er = E - 1;
            for (int i = D - 1; i >= 0; --i) {
                const double* h = a.hist + (size_t(i) * hRows) * nSlots + slot;
                const double* hPrev = a.hist + (size_t(i > 0 ? i - 1 : 0) * hRows) * nSlots + slot;
                double Sn[AMAX], Sp[AMAX], w[AMAX], cwb[AMAX];
#pragma unroll
                for (int k = 0; k < AMAX; ++k) {
                    cwb[k] = 0.0;
                    Sn[k] = (k < A) ? h[size_t(k) * nSlots] : 0.0;
                    w[k] = (k < A) ? h[size_t(A + k) * nSlots] : 0.0;
                    Sp[k] = (k < A) ? (i > 0 ? hPrev[size_t(k) * nSlots] : __ldg(a.spots + k)) : 0.0;
                }
                const double aliveBefore = h[size_t(2 * A) * nSlots];
                reverseEvent(er, Sn, aliveBefore);
                --er;
#pragma unroll
                for (int k = 0; k < AMAX; ++k) {
                    if (k < A) {
                        double cw = 0.0;
#pragma unroll
                        for (int j = 0; j <= k; ++j) cw += a.cholv[k * (k + 1) / 2 + j] * w[j];
                        const double df = __ldg(a.dynFwd + i * A + k);
                        const double fwd = Sp[k] * df;
                        const double sd = __ldg(a.stds + i * A + k), dr = __ldg(a.drifts + i * A + k);
                        const int dyn = __ldg(a.dyn + k);
                        const double al = __ldg(a.alphas + k);
                        const double sb = valid ? Sbar[k] : 0.0;
                        double fwdbar, xbar = 0.0, cwbar, sdbar;
                        if (dyn == 1) {                               // S = fwd + std cw
                            fwdbar = sb; sdbar = sb * cw; cwbar = sb * sd;
                        } else {
                            const double ex = exp_core(dr + sd * cw);
                            fwdbar = sb * ex;
                            if (dyn == 0) xbar = sb * Sn[k];                                   // S = fwd e
                            else if (dyn == 2) { xbar = sb * (Sn[k] + al); alphaBar[k] += sb * (ex - 1.0); }   // S = (fwd + al) e - al
                            else { xbar = sb * (Sn[k] - al); alphaBar[k] += sb * (1.0 - ex); }                 // S = (fwd - al) e + al
                            sdbar = xbar * cw; cwbar = xbar * sd;
                        }
                        // rows [0, A): dynFwd adjoints, [A, 2 A): drift adjoints, [2 A, 3 A): std adjoints
                        scr[k * 32 + lane] = fwdbar * Sp[k];
                        scr[(A + k) * 32 + lane] = xbar;
                        scr[(2 * A + k) * 32 + lane] = sdbar;
                        cwb[k] = cwbar;
                        Sbar[k] = fwdbar * df;
                    }
                }
                __syncwarp();
                for (int r = lane; r < 3 * A; r += 32) {
                    const int t = r / A;
                    myTab[(t == 0 ? oFwd : t == 1 ? oDrift : oStd) + i * A + (r - t * A)] += dlm_row_sum(scr + r * 32, lane);
                }
                __syncwarp();
                // outer product cwBar (x) w over the warp's paths: rows [0, A) = cwBar, [A, 2 A) = w
#pragma unroll
                for (int k = 0; k < AMAX; ++k)
                    if (k < A) { scr[k * 32 + lane] = cwb[k]; scr[(A + k) * 32 + lane] = w[k]; }
                __syncwarp();
#pragma unroll
                for (int q = 0; q < kRounds; ++q) {
                    if (pairK[q] < A) {
                        const double* xr = scr + pairK[q] * 32;
                        const double* wr = scr + (A + pairJ[q]) * 32;
                        double s0 = 0.0, s1 = 0.0;
#pragma unroll
                        for (int c = 0; c < 32; c += 2) {
                            const int c0 = (lane + c) & 31, c1 = (lane + c + 1) & 31;
                            s0 = fma(xr[c0], wr[c0], s0);
                            s1 = fma(xr[c1], wr[c1], s1);
                        }
                        cholAcc[q] += s0 + s1;
                    }
                }
                __syncwarp();
            }
            if (a.today) {
                double S0[AMAX];
#pragma unroll
                for (int k = 0; k < AMAX; ++k) S0[k] = (k < A) ? __ldg(a.spots + k) : 0.0;
                reverseEvent(0, S0, 1.0);
            }
#pragma unroll
            for (int k = 0; k < AMAX; ++k) if (k < A && valid) spotBar[k] += Sbar[k];
        }
    }

    // ---- block results -> partial[blockIdx]: payoff sums | agg | table adjoints
    __syncthreads();
    double* out = a.partial + size_t(blockIdx.x) * a.partial_stride;
    for (int k = tid; k < nPay; k += kBlock) {
        double s = 0.0;
        for (int w = 0; w < kWarps; ++w) s += payRows[size_t(w) * nPay + k];
        out[k] = s;
    }
    if (AAD) {
        double s = block_sum(aggSum, red);
        if (tid == 0) out[nPay] = s;
        double* adj = out + nPay + 1;
#pragma unroll
        for (int k = 0; k < AMAX; ++k) {
            if (k < A) {
                s = block_sum(spotBar[k], red);
                if (tid == 0) adj[k] = s;
                s = block_sum(alphaBar[k], red);
                if (tid == 0) adj[A + k] = s;
            }
        }
        for (int k = tid; k < A * A; k += kBlock) adj[2 * A + k] = 0.0;
        // Cholesky adjoint: the warps' pair sums through the (now idle) scratch rows, combined in warp order
        const int nPairs = A * (A + 1) / 2, scrStride = 3 * A * 32;
#pragma unroll
        for (int q = 0; q < kRounds; ++q)
            if (pairK[q] < A) scr[lane + 32 * q] = cholAcc[q];
        __syncthreads();
        for (int pp = tid; pp < nPairs; pp += kBlock) {
            double t = 0.0;
            for (int w = 0; w < kWarps; ++w) t += scrAll[size_t(w) * scrStride + pp];
            int k = 0;
            while ((k + 1) * (k + 2) / 2 <= pp) ++k;
            adj[2 * A + k * A + (pp - k * (k + 1) / 2)] = t;
        }
        __syncthreads();
        for (int k = tid; k < nStepTab; k += kBlock) {
            double t = 0.0;
            for (int w = 0; w < kWarps; ++w) t += tabRows[size_t(w) * nStepTab + k];
            adj[2 * A + A * A + k] = t;
        }
    }
}

}  // namespace cf
