// cf_kernels.cuh -- single-asset path kernels (Black-Scholes, Dupire) x (European, UOC):
// one path per thread, RNG + generatePath + payoffs fused, and for AAD a hand-written reverse
// sweep that replays the path backwards from a per-thread history and accumulates the table
// adjoints deterministically per block.
//
// Replaces, on the device: mcBase.h:378-386 (value loop), 680-704 (AAD loop) with
// Sobol::nextG sobol.h:103-109 / mrg32k3a::nextG mrg32k3a.h:150-186, Dupire::generatePath
// mcMdlDupire.h:238-280, BlackScholes::generatePath mcMdlBS.h:321-350, European::payoffs
// mcPrd.h:113-125, UOC::payoffs mcPrd.h:235-288, and the tape (AADTape.h / AADExpr.h) on that path.
// Adjoint equations: SURVEY.md Appendix A.1 / A.2.
//
// Performance structure (B200, FP64-pipe bound):
//  * Gaussians are produced a chunk of kChunk steps at a time per warp; the rare, expensive tail
//    branch of the Moro inverse (log(-log u), 16 % of draws) is compacted across the chunk so the
//    warp pays ~2-3 dense passes per chunk instead of one divergent pass per step.
//  * The spot bucket of the local-vol row is found with a uniform-cell lookup table + one
//    correcting compare instead of a binary search (bit-identical bucket to std::upper_bound).
//  * Barrier tests are done in log space; exp() is evaluated only on the final date and on the
//    rare samples inside the smoothing zone, where the reference's own comparisons are replayed.
//  * The reverse sweep stores only log-spots: g_i - v_i is recovered from consecutive log-spots.
#pragma once

#include "cf_device.cuh"
#include "cf_comm.cuh"
#include "../../include/cf_b200.h"

namespace cf {

constexpr int kMaxPay = 2;    // payoffs held per thread (European: 1, UOC: 2)
constexpr int kChunk = 8;     // steps of Gaussians staged per warp

struct KArgs {
    // run
    uint64_t first_path;
    uint64_t n_paths;
    int      n_batches;
    // rng
    uint32_t seed1, seed2;
    int      dim;                  // = n_steps (single asset)
    const uint32_t* sobol_dir;     // [32][dim]
    const uint64_t* mrg_jump;      // [32][2][9]
    // model
    int      n_steps, n_events, n_knots;
    const uint8_t* is_event;       // [n_steps + 1]
    double   spot;
    const double* tabA;            // BS: drifts [n_steps]      Dupire: interp_vols [n_steps][n_knots]
    const double* tabB;            // BS: stds   [n_steps]      Dupire: log_spots [n_knots]
    const double* numeraires;      // [n_events] or null (BS only; Dupire leaves the Sample defaults)
    const double* fwd_factors;     // [n_events] or null
    const double* discounts;       // [n_events] or null
    const double* libors;          // [n_events] or null (first libor of each event date)
    // Dupire bucket lookup: cell = (L - lut_x0) * lut_scale, lut[cell] = #knots <= left edge of cell
    const uint8_t* lut;
    int      lut_n;                // 0 = no table, use binary search
    double   lut_x0, lut_scale;
    int      store_g;              // 1: keep g_i in the history (needed when some interp_vol ~ 0)
    int      big_tables;           // 1: table A and the block's table adjoints do not fit in shared memory: A is read from
                                   //    global memory (L1 / L2) and the adjoints accumulate in the block's row of `partial`
    // product
    int      n_payoffs, is_put;
    double   strike, barrier, smooth, coupon;
    const double* event_dt;        // ContingentBond: coverage of the period starting at event e
    double   w[kMaxPay];           // payoff weights of the aggregate (AAD)
    // Europeans (mcPrd.h:290-401): strikes of event e are strikes[strike_off[e] .. strike_off[e+1]), weights in memory
    const double*  strikes;
    const int32_t* strike_off;
    const double*  wlong;          // [n_payoffs] payoff weights of the aggregate (AAD), device memory
    // outputs
    double*  partial;              // [gridDim][partial_stride]: payoff sums, agg, table adjoints
    int      partial_stride;
    double*  per_path_payoffs;     // [n_paths][n_payoffs] or null
    double*  per_path_agg;         // [n_paths] or null
    // scratch
    double*  hist;                 // [1 or 2][n_steps][gridDim * kBlock]
};

// ---------------------------------------------------------------------------------------------
// Spot sources handed to the products.  Dupire carries the log-spot and evaluates exp lazily;
// Black-Scholes carries the forward itself.
// ---------------------------------------------------------------------------------------------
struct LogSpotSrc {      // Dupire: forwards[0][0] = spot = exp(L), numeraire = discount = 1
    static constexpr bool kHasLog = true;
    double L, S;
    bool have;
    __device__ explicit LogSpotSrc(double l) : L(l), S(0.0), have(false) {}
    __device__ double logFwd() const { return L; }
    __device__ double fwd() { if (!have) { S = exp(L); have = true; } return S; }
    __device__ double num() const { return 1.0; }
    __device__ double disc() const { return 1.0; }
    __device__ double lib() const { return 0.0; }
};
struct FwdSrc {          // Black-Scholes: forward = S * ff[e], numeraire / discount / libor from tables
    static constexpr bool kHasLog = false;
    double F, N, Dsc, Lib;
    __device__ FwdSrc(double f, double n, double d, double l = 0.0) : F(f), N(n), Dsc(d), Lib(l) {}
    __device__ double logFwd() const { return 0.0; }
    __device__ double fwd() { return F; }
    __device__ double num() const { return N; }
    __device__ double disc() const { return Dsc; }
    __device__ double lib() const { return Lib; }
};
struct SampleAdj { double fwd, num, disc, lib; };

// Sink for products with many payoffs (Europeans): payoff k of this path goes to the warp's row of payoff
// sums (fixed order: lanes by shuffle tree, warps combined at the end of the kernel), to the aggregate
// and, on request, to the per-path matrix.
struct PayCtx {
    double*       myPay;       // this warp's [n_payoffs] row in shared memory
    double*       fw;          // this warp's 32 doubles of shared scratch (call ladders)
    const double* w;           // aggregate weights or null
    double*       perPath;     // this path's [n_payoffs] row or null
    double        agg;
    bool          valid;
    int           lane;
    __device__ __forceinline__ void emit(int k, double v)
    {
        const double s = warp_sum(valid ? v : 0.0);
        if (lane == 0) myPay[k] += s;
        if (w) agg += __ldg(w + k) * v;
        if (perPath) perPath[k] = v;
    }
    // calls max(F - K_k, 0) / num of the strikes k0 .. k0 + n - 1 (device memory), all lanes together
    __device__ __forceinline__ void ladder(int k0, int n, const double* __restrict__ K, double F, double num)
    {
        warp_ladder_sums(fw, F, valid, K + k0, n, num, myPay + k0, lane);
        if (perPath) {
            for (int k = k0; k < k0 + n; ++k) {
                const double v = div_n(fmax(F - __ldg(K + k), 0.0), num);
                if (w) agg += __ldg(w + k) * v;
                perPath[k] = v;
            }
        } else if (w) {
            // this path's share of the aggregate: the weighted positive parts first, one division for the ladder
            double a0 = 0.0, a1 = 0.0;
            int k = k0;
            for (; k + 1 < k0 + n; k += 2) {
                a0 = fma(__ldg(w + k), pos_part(F - __ldg(K + k)), a0);
                a1 = fma(__ldg(w + k + 1), pos_part(F - __ldg(K + k + 1)), a1);
            }
            if (k < k0 + n) a0 = fma(__ldg(w + k), pos_part(F - __ldg(K + k)), a0);
            agg += div_n(a0 + a1, num);
        }
    }
};

// ---------------------------------------------------------------------------------------------
// Products (streaming form: one call per event date, in order)
// ---------------------------------------------------------------------------------------------
template <int PRD> struct Product;

// European call, mcPrd.h:113-125: payoff = max(F - K, 0) * disc / num at the single event date
template <> struct Product<CF_PRODUCT_EUROPEAN> {
    double strike, pay, wbar;
    __device__ void init(const KArgs& a) { strike = a.strike; pay = 0.0; }
    template <class Src> __device__ void observe(int e, int nEvents, Src& s, PayCtx&)
    {
        if (e == 0) pay = fmax(s.fwd() - strike, 0.0) * s.disc() / s.num();
    }
    __device__ void payoffs(double* out) const { out[0] = pay; }
    __device__ void begin_reverse(const double* w) { wbar = w[0]; }
    template <class Src> __device__ SampleAdj reverse(int e, int nEvents, Src& s, double /*prevFwd*/)
    {
        SampleAdj r = {0.0, 0.0, 0.0, 0.0};
        if (e == 0) {
            const double F = s.fwd();
            const double intrinsic = fmax(F - strike, 0.0);
            // max(x, 0) has derivative 1 iff x > 0 strictly (AADExpr.h:571-583)
            r.fwd = (F - strike > 0.0) ? wbar * s.disc() / s.num() : 0.0;
            r.disc = wbar * intrinsic / s.num();
            r.num = -wbar * intrinsic * s.disc() / s.num() / s.num();
        }
        return r;
    }
};

// Up-and-out call/put with smoothed barrier, mcPrd.h:235-288.
template <> struct Product<CF_PRODUCT_UOC> {
    double strike, barSmooth, minusSmooth, twoSmooth, logZone;
    double alive, euro;
    bool   killed, isPut;
    double abar, aliveCur, eurobar;     // reverse state

    __device__ void init(const KArgs& a)
    {
        strike = a.strike; isPut = a.is_put != 0;
        twoSmooth = 2 * a.smooth; barSmooth = a.barrier + a.smooth; minusSmooth = a.barrier - a.smooth;
        // log-space pre-filter: below this log-forward the sample is certainly outside the smoothing
        // zone (margin >> rounding of exp/log); at or above it the reference's comparisons are
        // replayed on the forward itself.
        logZone = minusSmooth > 0.0 ? log(minusSmooth) - 1.0e-9 : -1.0e300;
        alive = 1.0; euro = 0.0; killed = false;
    }
    template <class Src> __device__ void observe(int e, int nEvents, Src& s, PayCtx&)
    {
        if (!killed && (!Src::kHasLog || s.logFwd() > logZone)) {
            const double F = s.fwd();
            if (F > barSmooth) { killed = true; alive = 0.0; }
            else if (F > minusSmooth) alive *= (barSmooth - F) / twoSmooth;
        }
        if (e == nEvents - 1) {
            const double F = s.fwd();
            euro = (isPut ? fmax(strike - F, 0.0) : fmax(F - strike, 0.0)) / s.num();
        }
    }
    __device__ void payoffs(double* out) const { out[0] = alive * euro; out[1] = euro; }

    __device__ void begin_reverse(const double* w)
    {
        eurobar = w[0] * alive + w[1];          // alive is the constant 0 on killed paths
        abar = killed ? 0.0 : w[0] * euro;      // a killed `alive` is a fresh leaf: nothing flows
        aliveCur = alive;
    }
    template <class Src> __device__ SampleAdj reverse(int e, int nEvents, Src& s, double /*prevFwd*/)
    {
        SampleAdj r = {0.0, 0.0, 0.0, 0.0};
        if (e == nEvents - 1) {
            const double F = s.fwd();
            const double x = isPut ? strike - F : F - strike;
            if (x > 0.0) r.fwd = (isPut ? -eurobar : eurobar) / s.num();
            r.num = -eurobar * euro / s.num();
        }
        if (!killed && (!Src::kHasLog || s.logFwd() > logZone)) {
            const double F = s.fwd();
            if (F > minusSmooth) {              // fuzzy sample (F <= barSmooth on a live path)
                const double f = (barSmooth - F) / twoSmooth;
                // alive before this sample; f == 0 only if the spot sits exactly on the upper edge
                const double alivePrev = (f != 0.0) ? aliveCur / f : 0.0;
                r.fwd += abar * alivePrev * (-1.0 / twoSmooth);
                abar *= f;
                aliveCur = alivePrev;
            }
        }
        return r;
    }
};

// Portfolio of European calls, several strikes per maturity, mcPrd.h:374-399:
// payoff (event e, strike k) = max(F_e - k, 0) / num_e.  Payoffs are emitted as they are computed.
template <> struct Product<CF_PRODUCT_EUROPEANS> {
    const double*  K;
    const int32_t* off;
    const double*  w;
    __device__ void init(const KArgs& a) { K = a.strikes; off = a.strike_off; w = a.wlong; }
    template <class Src> __device__ void observe(int e, int nEvents, Src& s, PayCtx& c)
    {
        const double F = s.fwd(), num = s.num();
        const int k0 = __ldg(off + e), k1 = __ldg(off + e + 1);
        c.ladder(k0, k1 - k0, K, F, num);
    }
    __device__ void payoffs(double*) const {}
    __device__ void begin_reverse(const double*) {}
    template <class Src> __device__ SampleAdj reverse(int e, int nEvents, Src& s, double /*prevFwd*/)
    {
        SampleAdj r = {0.0, 0.0, 0.0, 0.0};
        const double F = s.fwd(), num = s.num();
        const int k1 = __ldg(off + e + 1);
        // max(x, 0) has derivative 1 iff x > 0 (AADExpr.h:571-583): the weights of the strikes in the money and their
        // moneyness, branch-free, then one division each for the ladder
        double sw = 0.0, swx = 0.0;
        for (int k = __ldg(off + e); k < k1; ++k) {
            const double x = F - __ldg(K + k);
            const double wk = x > 0.0 ? __ldg(w + k) : 0.0;
            sw += wk;
            swx = fma(wk, x, swx);
        }
        r.fwd = div_n(sw, num);
        r.num = -div_n(swx, num * num);
        return r;
    }
};

// Contingent floater, mcPrd.h:513-572: per period [T_e-1, T_e] the coupon (libor(T_e-1, T_e) + cpn) * coverage is paid
// at T_e if the asset went up over the period, with a smoothed digital of half-width `smooth`; redemption at maturity.
template <> struct Product<CF_PRODUCT_CONTINGENT> {
    double smooth, twoSmooth, cpn, pay, s0, libPrev, wbar, carryFwd, carryLib;
    const double* dt;
    const double* libors;
    __device__ void init(const KArgs& a)
    {
        smooth = a.smooth; twoSmooth = 2 * a.smooth; cpn = a.coupon; dt = a.event_dt; libors = a.libors;
        pay = 0.0; s0 = 0.0; libPrev = 0.0;
    }
    __device__ double digital(double d) const
    {
        if (d > smooth) return 1.0;
        if (d < -smooth) return 0.0;
        return (d + smooth) / twoSmooth;                      // "fuzzy" edge: interpolate (mcPrd.h:556-559)
    }
    template <class Src> __device__ void observe(int e, int nEvents, Src& s, PayCtx&)
    {
        const double s1 = s.fwd();
        if (e > 0) pay += digital(s1 - s0) * (libPrev + cpn) * __ldg(dt + e - 1) / s.num();
        if (e == nEvents - 1) pay += 1.0 / s.num();           // redemption at maturity
        s0 = s1; libPrev = s.lib();
    }
    __device__ void payoffs(double* out) const { out[0] = pay; }
    __device__ void begin_reverse(const double* w) { wbar = w[0]; carryFwd = 0.0; carryLib = 0.0; }
    template <class Src> __device__ SampleAdj reverse(int e, int nEvents, Src& s, double prevFwd)
    {
        // what the period starting here (paid at event e + 1) sent back to this date's spot and libor
        SampleAdj r = {carryFwd, 0.0, 0.0, carryLib};
        carryFwd = 0.0; carryLib = 0.0;
        const double num = s.num();
        if (e == nEvents - 1) r.num -= wbar / (num * num);
        if (e > 0) {
            const double d = s.fwd() - prevFwd, lib = libors ? __ldg(libors + e - 1) : 0.0, cov = __ldg(dt + e - 1);
            const double dig = digital(d);
            const double ddig = (d > smooth || d < -smooth) ? 0.0 : 1.0 / twoSmooth;
            r.num -= wbar * (dig * (lib + cpn) * cov / num) / num;
            const double g = wbar * ddig * (lib + cpn) * cov / num;
            r.fwd += g; carryFwd = -g;
            carryLib = wbar * dig * cov / num;
        }
        return r;
    }
};

// ---------------------------------------------------------------------------------------------
// Shared-memory carve-up
// ---------------------------------------------------------------------------------------------
struct Smem {
    double*   tabA;      // model table A
    double*   tabB;      // model table B
    double*   invdx;     // Dupire: 1 / (x[n+1] - x[n])
    double*   adj;       // block table adjoints (AAD)
    double2*  wrow;      // [2][kWarps][rowlen] warp rows (AAD)
    double*   red;       // [kWarps] reduction scratch
    double*   gq;        // [kWarps][kChunk][32] staged Gaussians
    uint16_t* tagq;      // [kWarps][kChunk*32] tail queue
    uint32_t* dirlow;    // [dim][8]
    uint32_t* base;      // [2][dim]
    uint8_t*  lut;       // [lut_n]
    uint8_t*  isev;      // [n_steps + 1]
    double*   pay;       // [kWarps][n_payoffs] payoff sums of many-payoff products
};

__host__ __device__ inline size_t align16(size_t x) { return (x + 15) & ~size_t(15); }

template <int MDL>
__host__ __device__ inline int table_a_size(int nSteps, int nKnots) { return MDL == CF_MODEL_DUPIRE ? nSteps * nKnots : nSteps; }
template <int MDL>
__host__ __device__ inline int table_b_size(int nSteps, int nKnots) { return MDL == CF_MODEL_DUPIRE ? nKnots : nSteps; }
template <int MDL>
__host__ __device__ inline int adj_table_size(int nSteps, int nKnots, int nEvents)
{
    return MDL == CF_MODEL_DUPIRE ? nSteps * nKnots : 2 * nSteps + 4 * nEvents;
}
template <int MDL>
__host__ __device__ inline int row_len(int nKnots) { return MDL == CF_MODEL_DUPIRE ? (nKnots > 1 ? nKnots - 1 : 1) : 3; }

struct SmemSizes { size_t tabA, tabB, invdx, adj, wrow, red, gq, tagq, dirlow, base, lut, isev, pay, total; };

template <int MDL, bool AAD>
__host__ __device__ inline SmemSizes smem_sizes(int nSteps, int nKnots, int nEvents, int dim, bool sobol, int lutN, int nPayRows = 0,
                                                bool bigTables = false)
{
    SmemSizes s{};
    s.tabA = bigTables ? 0 : align16(sizeof(double) * table_a_size<MDL>(nSteps, nKnots));
    s.tabB = align16(sizeof(double) * table_b_size<MDL>(nSteps, nKnots));
    s.invdx = align16(sizeof(double) * (nKnots > 0 ? nKnots : 1));
    s.adj = (AAD && !bigTables) ? align16(sizeof(double) * adj_table_size<MDL>(nSteps, nKnots, nEvents)) : 0;
    s.wrow = AAD ? align16(sizeof(double2) * 2 * kWarps * row_len<MDL>(nKnots)) : 0;
    s.red = align16(sizeof(double) * kWarps);
    s.gq = align16(sizeof(double) * kWarps * kChunk * 32);
    s.tagq = align16(sizeof(uint16_t) * kWarps * kChunk * 32);
    s.dirlow = sobol ? align16(sizeof(uint32_t) * dim * kLowBits) : 0;
    s.base = sobol ? align16(sizeof(uint32_t) * 2 * dim) : 0;
    s.lut = align16(size_t(lutN > 0 ? lutN : 1));
    s.isev = align16(size_t(nSteps) + 1);
    // wrow (reverse sweep) aliases gq + tagq (forward sweep): the two phases never overlap within a block
    if (s.gq + s.tagq < s.wrow) s.gq = s.wrow - s.tagq;
    s.isev = align16(s.isev);
    s.pay = align16(sizeof(double) * kWarps * (size_t(nPayRows) + (nPayRows > 0 ? 32 : 0)));      // + the ladders' scratch
    s.total = s.tabA + s.tabB + s.invdx + s.adj + s.red + s.gq + s.tagq + s.dirlow + s.base + s.lut + s.isev + s.pay;
    return s;
}

template <int MDL, bool AAD>
__device__ inline Smem carve(unsigned char* p, int nSteps, int nKnots, int nEvents, int dim, bool sobol, int lutN, int nPayRows,
                             bool bigTables)
{
    const SmemSizes z = smem_sizes<MDL, AAD>(nSteps, nKnots, nEvents, dim, sobol, lutN, nPayRows, bigTables);
    Smem s{};
    s.tabA = reinterpret_cast<double*>(p);    p += z.tabA;
    s.tabB = reinterpret_cast<double*>(p);    p += z.tabB;
    s.invdx = reinterpret_cast<double*>(p);   p += z.invdx;
    s.adj = reinterpret_cast<double*>(p);     p += z.adj;
    s.wrow = reinterpret_cast<double2*>(p + z.red);   // aliases gq / tagq
    s.red = reinterpret_cast<double*>(p);     p += z.red;
    s.gq = reinterpret_cast<double*>(p);      p += z.gq;
    s.tagq = reinterpret_cast<uint16_t*>(p);  p += z.tagq;
    s.dirlow = reinterpret_cast<uint32_t*>(p); p += z.dirlow;
    s.base = reinterpret_cast<uint32_t*>(p);  p += z.base;
    s.lut = reinterpret_cast<uint8_t*>(p);    p += z.lut;
    s.isev = reinterpret_cast<uint8_t*>(p);   p += z.isev;
    s.pay = reinterpret_cast<double*>(p);
    return s;
}

// ---------------------------------------------------------------------------------------------
// interp (interp.h:26-63) on the smem row y against knots x: upper_bound, flat extrapolation.
// ---------------------------------------------------------------------------------------------
struct Bucket { int n; int side; };   // left knot of the bucket; side: 0 inside, -1 / +1 flat extrapolation left / right

struct Locator {
    const double* x; const uint8_t* lut;
    int m, p2, lutN;
    double x0, scale;

    // ub = number of knots <= v  (std::upper_bound)
    __device__ __forceinline__ int upper_bound(double v) const
    {
        int ub;
        if (lutN > 0) {
            // uniform cells of width <= half the smallest knot spacing: the cell's entry is within
            // one of the answer, one compare each way settles it
            double c = (v - x0) * scale;
            c = fmin(fmax(c, 0.0), double(lutN - 1));
            ub = lut[__double2int_rz(c)];
            if (ub < m && x[ub] <= v) ++ub;
            else if (ub > 0 && x[ub - 1] > v) --ub;
        } else {
            ub = 0;
            for (int s = p2; s > 0; s >>= 1) {
                const int c = ub + s;
                if (c <= m && x[c - 1] <= v) ub = c;
            }
        }
        return ub;
    }
    __device__ __forceinline__ Bucket locate(double v) const
    {
        const int ub = upper_bound(v);
        Bucket b;
        if (ub == m) { b.n = m > 1 ? m - 2 : 0; b.side = 1; }
        else if (ub == 0) { b.n = 0; b.side = -1; }
        else { b.n = ub - 1; b.side = 0; }
        return b;
    }
};

// ---------------------------------------------------------------------------------------------
// Gaussians for a chunk of steps, per warp (all lanes must call together).
// Central branch of invNormalCdf (gaussians.h:73-78) inline; tail branch (80-86) compacted.
// ---------------------------------------------------------------------------------------------
template <int RNGK>
struct GaussGen {
    SobolThread sob;
    MrgThread   mrg;
    double      sign;        // mrg32k3a antithetic: -1 on odd paths
    double*     gq;          // this warp's [kChunk][32]
    uint16_t*   tagq;        // this warp's [kChunk*32]
    const uint32_t* dirlow; const uint32_t* base; int dim;

    __device__ __forceinline__ double uniform(int d)
    {
        if (RNGK == CF_RNG_SOBOL) return CF_ONEOVER2POW32 * double(sob.state(dirlow, base, dim, d));
        return mrg_uniform(mrg.next());
    }

    __device__ __forceinline__ void fill(int i0, int cnt)
    {
        const int lane = threadIdx.x & 31;
        const unsigned ltMask = (1u << lane) - 1u;
        int q = 0;
        __syncwarp();
        for (int k = 0; k < cnt; ++k) {
            const double p = uniform(i0 + k);
            const bool sup = p > 0.5;
            const double up = sup ? 1.0 - p : p;
            const double x = up - 0.5;
            const bool central = fabs(x) < 0.42;
            double r = x * x;
            const double num = ((-25.44106049637 * r + 41.39119773534) * r + -18.61500062529) * r + 2.50662823884;
            const double den = (((3.13082909833 * r + -21.06224101826) * r + 23.08336743743) * r + -8.47351093090) * r + 1.0;
            r = div_fast(x * num, den);                 // den in [0.11, 1]: the lean quotient equals the IEEE one (cf_device.cuh)
            gq[k * 32 + lane] = central ? (sup ? -r : r) : up;
            const unsigned ball = __ballot_sync(kFull, !central);
            if (!central) tagq[q + __popc(ball & ltMask)] = uint16_t((sup ? 0x8000u : 0u) | (unsigned(k) << 5) | unsigned(lane));
            q += __popc(ball);
        }
        __syncwarp();
        for (int b = 0; b < q; b += 32) {
            const int idx = b + lane;
            if (idx < q) {
                const unsigned t = tagq[idx];
                const int slot = int(t & 0x7fffu);            // k * 32 + lane
                double r = log(-log(gq[slot]));
                r = 0.3374754822726147 + r * (0.9761690190917186 + r * (0.1607979714918209 + r * (0.0276438810333863
                    + r * (0.0038405729373609 + r * (0.0003951896511919 + r * (0.0000321767881768
                    + r * (0.0000002888167364 + r * 0.0000003960315187)))))));
                gq[slot] = (t & 0x8000u) ? r : -r;
            }
        }
        __syncwarp();
    }
    __device__ __forceinline__ double get(int k) const
    {
        const double g = gq[k * 32 + (threadIdx.x & 31)];
        return RNGK == CF_RNG_SOBOL ? g : sign * g;
    }
};

// ---------------------------------------------------------------------------------------------
// The path kernel
// ---------------------------------------------------------------------------------------------
template <int MDL, int PRD, bool AAD, int RNGK>
__global__ void __launch_bounds__(kBlock, 2) path_kernel(const KArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int D = a.n_steps, m = a.n_knots, E = a.n_events;
    constexpr bool kSobol = (RNGK == CF_RNG_SOBOL);
    constexpr bool kDupire = (MDL == CF_MODEL_DUPIRE);
    constexpr bool kManyPay = (PRD == CF_PRODUCT_EUROPEANS);
    const int nPayRows = kManyPay ? a.n_payoffs : 0;
    const bool big = a.big_tables != 0;
    Smem sm = carve<MDL, AAD>(smem_raw, D, m, E, a.dim, kSobol, a.lut_n, nPayRows, big);
    const int rowLen = row_len<MDL>(m);
    const int nAdj = adj_table_size<MDL>(D, m, E);
    const bool storeG = !kDupire || a.store_g != 0;
    if (big) {
        // long schedules (n_steps * n_knots beyond shared memory): table A stays in global memory, the table adjoints
        // accumulate directly in this block's row of `partial` (one thread per entry: no race, program order)
        sm.tabA = const_cast<double*>(a.tabA);
        sm.adj = a.partial + size_t(blockIdx.x) * a.partial_stride + a.n_payoffs + 2;
    }

    // ---- stage tables
    if (!big) for (int i = tid; i < table_a_size<MDL>(D, m); i += kBlock) sm.tabA[i] = a.tabA[i];
    for (int i = tid; i < table_b_size<MDL>(D, m); i += kBlock) sm.tabB[i] = a.tabB[i];
    if (kDupire) {
        for (int i = tid; i + 1 < m; i += kBlock) sm.invdx[i] = 1.0 / (a.tabB[i + 1] - a.tabB[i]);
        for (int i = tid; i < a.lut_n; i += kBlock) sm.lut[i] = a.lut[i];
    }
    for (int i = tid; i <= D; i += kBlock) sm.isev[i] = a.is_event[i];
    if (AAD) for (int i = tid; i < nAdj; i += kBlock) sm.adj[i] = 0.0;
    for (int i = tid; i < kWarps * nPayRows; i += kBlock) sm.pay[i] = 0.0;
    if (kSobol) sobol_load_low(sm.dirlow, a.sobol_dir, a.dim);
    Locator loc;
    loc.x = sm.tabB; loc.lut = sm.lut; loc.m = m; loc.lutN = a.lut_n; loc.x0 = a.lut_x0; loc.scale = a.lut_scale;
    loc.p2 = 1;
    while (loc.p2 * 2 <= m) loc.p2 *= 2;
    __syncthreads();

    const size_t nSlots = size_t(gridDim.x) * kBlock;
    const size_t slot = size_t(blockIdx.x) * kBlock + tid;
    double* histL = a.hist;                         // Dupire: L_i            BS: S_{i+1}
    double* histG = a.hist + size_t(D) * nSlots;    // g_i (when stored)

    GaussGen<RNGK> gen;
    gen.gq = sm.gq + size_t(warp) * kChunk * 32;
    gen.tagq = sm.tagq + size_t(warp) * kChunk * 32;
    gen.dirlow = sm.dirlow; gen.base = sm.base; gen.dim = a.dim;

    double paySum[kMaxPay] = {0.0, 0.0};
    double aggSum = 0.0, spotBar = 0.0;
    const double logS0 = kDupire ? log(a.spot) : 0.0;

    // contiguous batches per block: with mrg32k3a a thread's next path is kBlock / 2 antithetic pairs down the stream,
    // one jump matrix instead of a full skip-ahead (a third of a short path's instructions)
    const int bBeg = int(int64_t(blockIdx.x) * a.n_batches / gridDim.x), bEnd = int(int64_t(blockIdx.x + 1) * a.n_batches / gridDim.x);
    MrgThread mrgStart;
    for (int batch = bBeg; batch < bEnd; ++batch) {
        const uint64_t p = uint64_t(batch) * kBlock + tid;     // path within this run
        const bool valid = p < a.n_paths;
        const uint64_t pabs = a.first_path + p;

        // ---- RNG positioning
        gen.sign = 1.0;
        if (kSobol) {
            const uint32_t n0 = uint32_t(a.first_path + uint64_t(batch) * kBlock + 1);
            const uint32_t H0 = n0 >> kLowBits;
            __syncthreads();                               // previous batch done with base[]
            sobol_block_base(sm.base, a.sobol_dir, a.dim, H0);
            __syncthreads();
            gen.sob.init(uint32_t(pabs + 1), H0);
        } else {
            if (batch == bBeg) mrgStart.init(a.seed1, a.seed2, pabs >> 1, a.mrg_jump);
            else mrgStart.advance(uint64_t(kBlock / 2), a.mrg_jump);
            gen.mrg = mrgStart;
            gen.sign = (pabs & 1ull) ? -1.0 : 1.0;
        }
        auto bsSample = [&](int e, double spotNow) -> FwdSrc {
            return FwdSrc(a.fwd_factors ? spotNow * __ldg(a.fwd_factors + e) : spotNow,
                          a.numeraires ? __ldg(a.numeraires + e) : 1.0,
                          a.discounts ? __ldg(a.discounts + e) : 1.0,
                          a.libors ? __ldg(a.libors + e) : 0.0);
        };

        // ---- forward: generatePath + payoffs
        Product<PRD> prd;
        prd.init(a);
        PayCtx ctx;
        ctx.myPay = sm.pay + size_t(warp) * nPayRows; ctx.w = AAD ? a.wlong : nullptr;
        ctx.fw = sm.pay + size_t(kWarps) * nPayRows + size_t(warp) * 32;
        ctx.perPath = (valid && a.per_path_payoffs) ? a.per_path_payoffs + p * a.n_payoffs : nullptr;
        ctx.agg = 0.0; ctx.valid = valid; ctx.lane = lane;
        int e = 0;
        double X = kDupire ? logS0 : a.spot;                   // Dupire: log spot, BS: spot
        if (sm.isev[0]) {
            if (kDupire) { LogSpotSrc s(X); prd.observe(e, E, s, ctx); }
            else { FwdSrc s = bsSample(e, X); prd.observe(e, E, s, ctx); }
            ++e;
        }
        for (int i0 = 0; i0 < D; i0 += kChunk) {
            const int cnt = min(kChunk, D - i0);
            gen.fill(i0, cnt);
            for (int k = 0; k < cnt; ++k) {
                const int i = i0 + k;
                const double g = gen.get(k);
                if (kDupire) {
                    if (AAD) {
                        histL[size_t(i) * nSlots + slot] = X;
                        if (storeG) histG[size_t(i) * nSlots + slot] = g;
                    }
                    const Bucket b = loc.locate(X);
                    const double* y = sm.tabA + i * m;
                    double v;
                    if (b.side != 0) v = y[b.side > 0 ? m - 1 : 0];
                    else { const double y1 = y[b.n]; v = y1 + (y[b.n + 1] - y1) * ((X - sm.tabB[b.n]) * sm.invdx[b.n]); }
                    X += v * (-0.5 * v + g);                   // mcMdlDupire.h:271
                    if (sm.isev[i + 1]) { LogSpotSrc s(X); prd.observe(e, E, s, ctx); ++e; }
                } else {
                    X = X * exp_core(sm.tabA[i] + sm.tabB[i] * g);  // mcMdlBS.h:343 (lean exp: the exponent is a few standard deviations)
                    if (AAD) { histL[size_t(i) * nSlots + slot] = X; histG[size_t(i) * nSlots + slot] = g; }
                    FwdSrc s = bsSample(e, X);
                    prd.observe(e, E, s, ctx); ++e;            // every BS step ends on an event date
                }
            }
        }
        double pay[kMaxPay] = {0.0, 0.0};
        prd.payoffs(pay);
        double agg = kManyPay ? ctx.agg : 0.0;
        if (!kManyPay) {
#pragma unroll
            for (int k = 0; k < kMaxPay; ++k) if (k < a.n_payoffs) agg += a.w[k] * pay[k];
        }
        if (valid) {
            if (!kManyPay) {
#pragma unroll
                for (int k = 0; k < kMaxPay; ++k) if (k < a.n_payoffs) paySum[k] += pay[k];
                if (a.per_path_payoffs)
                    for (int k = 0; k < a.n_payoffs; ++k) a.per_path_payoffs[p * a.n_payoffs + k] = pay[k];
            }
            aggSum += agg;
            if (a.per_path_agg) a.per_path_agg[p] = agg;
        }

        // ---- reverse sweep (block-synchronous: one barrier per step)
        if (AAD) {
            __syncthreads();            // warp rows alias the Gaussian staging area of the forward sweep
            prd.begin_reverse(a.w);
            double Xbar = 0.0;          // adjoint of X_{i+1} (Dupire: log spot; BS: spot)
            int er = E - 1;
            int buf = 0;
            for (int i = D - 1; i >= 0; --i, buf ^= 1) {
                double2* myRow = sm.wrow + size_t(buf * kWarps + warp) * rowLen;
                if (kDupire) {
                    if (sm.isev[i + 1]) {
                        LogSpotSrc s(X);
                        const SampleAdj sa = prd.reverse(er, E, s, 0.0);
                        if (sa.fwd != 0.0) Xbar += sa.fwd * s.fwd();   // dS/dL = S
                        --er;
                    }
                    const double L = histL[size_t(i) * nSlots + slot];
                    const Bucket b = loc.locate(L);
                    const double* y = sm.tabA + i * m;
                    double v, t, slope;
                    if (b.side != 0) {
                        const bool right = b.side > 0;
                        v = y[right ? m - 1 : 0]; t = (right && m > 1) ? 1.0 : 0.0; slope = 0.0;
                    } else {
                        const double y1 = y[b.n], dy = y[b.n + 1] - y1, idx = sm.invdx[b.n];
                        t = (L - sm.tabB[b.n]) * idx;
                        v = y1 + dy * t;
                        slope = dy * idx;
                    }
                    // g_i - v_i: stored, or recovered from L_{i+1} = L_i + v (g - v/2)
                    const double gmv = storeG ? histG[size_t(i) * nSlots + slot] - v : (X - L) / v - 0.5 * v;
                    const double vbar = valid ? Xbar * gmv : 0.0;
                    const double bb = vbar * t;
                    warp_keyed_accumulate(myRow, rowLen, b.n, vbar - bb, bb);
                    Xbar += vbar * slope;
                    X = L;
                    __syncthreads();
                    if (tid < m) {
                        double s = 0.0;
                        const double2* rows = sm.wrow + size_t(buf * kWarps) * rowLen;
                        for (int w = 0; w < kWarps; ++w) {
                            if (tid < rowLen) s += rows[w * rowLen + tid].x;
                            if (tid >= 1 && m > 1) s += rows[w * rowLen + tid - 1].y;
                        }
                        sm.adj[i * m + tid] += s;
                    }
                } else {
                    // BS: X currently holds S_{i+1}
                    const double g = histG[size_t(i) * nSlots + slot];
                    const double S1 = X;
                    const double S0 = (i > 0) ? histL[size_t(i - 1) * nSlots + slot] : a.spot;
                    FwdSrc smp = bsSample(er, S1);
                    // forward sampled on the previous event date (products that compare consecutive samples)
                    const double prevFwd = (er > 0 && a.fwd_factors) ? S0 * __ldg(a.fwd_factors + er - 1) : S0;
                    const SampleAdj sa = prd.reverse(er, E, smp, prevFwd);
                    const double ff = a.fwd_factors ? __ldg(a.fwd_factors + er) : 1.0;
                    // Xbar carries the adjoint of log S: S_{i+1} = S_i e_i makes Sbar_i S_i = Sbar_{i+1} S_{i+1}, so the sweep
                    // needs neither e_i nor a product per step; a sample adds sa.fwd ff S_{i+1}
                    if (sa.fwd != 0.0) Xbar += sa.fwd * ff * S1;
                    const double abar = valid ? Xbar : 0.0;                // adjoint of drift_i + std_i g_i
                    // dense per-step values: drift, std | numeraire, fwd factor, discount, libor of event er (mostly zero: one
                    // vote decides whether the warp needs their sums at all)
                    double v0 = warp_sum(abar), v1 = warp_sum(abar * g);
                    double v2 = 0.0, v3 = 0.0, v4 = 0.0, v5 = 0.0;
                    if (__any_sync(kFull, valid && (sa.num != 0.0 || sa.fwd != 0.0 || sa.disc != 0.0 || sa.lib != 0.0))) {
                        v2 = warp_sum(valid ? sa.num : 0.0); v3 = warp_sum(valid ? sa.fwd * S1 : 0.0);
                        v4 = warp_sum(valid ? sa.disc : 0.0); v5 = warp_sum(valid ? sa.lib : 0.0);
                    }
                    if (lane == 0) { myRow[0] = make_double2(v0, v1); myRow[1] = make_double2(v2, v3); myRow[2] = make_double2(v4, v5); }
                    X = S0;
                    __syncthreads();
                    if (tid < 6) {
                        double s = 0.0;
                        const double2* rows = sm.wrow + size_t(buf * kWarps) * rowLen;
                        for (int w = 0; w < kWarps; ++w) {
                            const double2 q = rows[w * rowLen + (tid >> 1)];
                            s += (tid & 1) ? q.y : q.x;
                        }
                        const int idx = tid == 0 ? i : tid == 1 ? D + i : 2 * D + (tid - 2) * E + er;
                        sm.adj[idx] += s;
                    }
                    --er;
                }
            }
            // today's sample (timeline point 0)
            if (sm.isev[0]) {
                if (kDupire) {
                    LogSpotSrc s(X);
                    const SampleAdj sa = prd.reverse(0, E, s, 0.0);
                    if (sa.fwd != 0.0) Xbar += sa.fwd * s.fwd();
                } else {
                    FwdSrc smp = bsSample(0, X);
                    const SampleAdj sa = prd.reverse(0, E, smp, 0.0);
                    const double ff = a.fwd_factors ? __ldg(a.fwd_factors) : 1.0;
                    Xbar += sa.fwd * ff * X;
                    double v2 = warp_sum(valid ? sa.num : 0.0), v3 = warp_sum(valid ? sa.fwd * X : 0.0);
                    double v4 = warp_sum(valid ? sa.disc : 0.0), v5 = warp_sum(valid ? sa.lib : 0.0);
                    __syncthreads();
                    if (lane == 0) { sm.wrow[warp * rowLen] = make_double2(v2, v3); sm.wrow[warp * rowLen + 1] = make_double2(v4, v5); }
                    __syncthreads();
                    if (tid < 4) {
                        double s = 0.0;
                        for (int w = 0; w < kWarps; ++w) {
                            const double2 q = sm.wrow[w * rowLen + (tid >> 1)];
                            s += (tid & 1) ? q.y : q.x;
                        }
                        sm.adj[2 * D + tid * E] += s;
                    }
                }
            }
            __syncthreads();
            // spot leaf: Dupire L0 = log(S0) -> 1/S0 (mcMdlDupire.h:245); BS: S_0 = spot
            if (valid) spotBar += Xbar / a.spot;          // Dupire: L0 = log(S0); Black-Scholes: the adjoint of log S_0, S_0 = spot
        }
    }

    // ---- block results -> partial[blockIdx]
    double* out = a.partial + size_t(blockIdx.x) * a.partial_stride;
    if (kManyPay) {
        __syncthreads();
        for (int k = tid; k < a.n_payoffs; k += kBlock) {
            double s = 0.0;
            for (int w = 0; w < kWarps; ++w) s += sm.pay[size_t(w) * nPayRows + k];
            out[k] = s;
        }
    } else {
        for (int k = 0; k < a.n_payoffs && k < kMaxPay; ++k) {
            const double s = block_sum(paySum[k], sm.red);
            if (tid == 0) out[k] = s;
        }
    }
    if (AAD) {
        double s = block_sum(aggSum, sm.red);
        if (tid == 0) out[a.n_payoffs] = s;
        s = block_sum(spotBar, sm.red);
        if (tid == 0) out[a.n_payoffs + 1] = s;
        __syncthreads();
        if (!big) for (int i = tid; i < nAdj; i += kBlock) out[a.n_payoffs + 2 + i] = sm.adj[i];
    }
}

// out[k] = sum over blocks (fixed order) of partial[b][k]
static __global__ void reduce_partials_kernel(const double* __restrict__ partial, int nBlocks, int stride, int n,
                                       double* __restrict__ out)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    double s = 0.0;
    for (int b = 0; b < nBlocks; ++b) s += partial[size_t(b) * stride + k];
    out[k] = s;
}


// in: [nHead] head values then ybar[D][m] (adjoints of interp_vols).  out: head then volsbar[m][nTimes]
// with volsbar[j][k] = sum_i (k1[i] == k ? c1[i] : 0) * ybar[i][j] + (k2[i] == k ? c2[i] : 0) * ybar[i][j]
static __global__ void collapse_time_kernel(const double* __restrict__ in, int nHead, int D, int m, int nTimes,
                                     const int32_t* __restrict__ k1, const int32_t* __restrict__ k2,
                                     const double* __restrict__ c1, const double* __restrict__ c2,
                                     double* __restrict__ out)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < nHead) { out[q] = in[q]; return; }
    const int r = q - nHead;
    if (r >= m * nTimes) return;
    const int j = r / nTimes, k = r % nTimes;
    double s = 0.0;
    for (int i = D - 1; i >= 0; --i) {
        const double y = in[nHead + i * m + j];
        if (k1[i] == k) s += c1[i] * y;
        if (k2[i] == k) s += c2[i] * y;
    }
    out[q] = s;
}

}  // namespace cf
