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
#pragma once

#include "cf_device.cuh"
#include "../../include/cf_b200.h"

namespace cf {

constexpr int kMaxPay = 2;   // payoffs held per thread (European: 1, UOC: 2)

// Kernel arguments: device pointers to the uploaded tables (see cf_model / cf_product).
struct KArgs {
    // run
    uint64_t first_path;
    uint64_t n_paths;
    int      n_batches;
    // rng
    int      rng_kind;
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
    const double* numeraires;      // [n_events] or null
    const double* fwd_factors;     // [n_events] or null
    const double* discounts;       // [n_events] or null
    // product
    int      n_payoffs, is_put;
    double   strike, barrier, smooth;
    double   w[kMaxPay];           // payoff weights of the aggregate (AAD)
    // outputs
    double*  partial;              // [gridDim][partial_stride]: payoff sums, agg, table adjoints
    int      partial_stride;
    double*  per_path_payoffs;     // [n_paths][n_payoffs] or null
    double*  per_path_agg;         // [n_paths] or null
    // scratch
    double*  hist;                 // [2][n_steps][gridDim * kBlock]
};

// ---------------------------------------------------------------------------------------------
// Products (streaming form: one call per event date, in order)
// ---------------------------------------------------------------------------------------------
struct Sample { double fwd, num, disc; };   // forwards[0][0], numeraire, discounts[0]
struct SampleAdj { double fwd, num, disc; };

template <int PRD> struct Product;

// European call, mcPrd.h:113-125: payoff = max(F - K, 0) * disc / num at the single event date
template <> struct Product<CF_PRODUCT_EUROPEAN> {
    double strike, pay;
    __device__ void init(const KArgs& a) { strike = a.strike; pay = 0.0; }
    __device__ void observe(int e, int nEvents, const Sample& s)
    {
        if (e == 0) pay = fmax(s.fwd - strike, 0.0) * s.disc / s.num;
    }
    __device__ void payoffs(double* out) const { out[0] = pay; }
    // reverse: agg = w[0] * pay
    __device__ void begin_reverse(const double* w) { wbar = w[0]; }
    __device__ SampleAdj reverse(int e, int nEvents, const Sample& s) const
    {
        SampleAdj r = {0.0, 0.0, 0.0};
        if (e == 0) {
            const double intrinsic = fmax(s.fwd - strike, 0.0);
            // max(x, 0) has derivative 1 iff x > 0 strictly (AADExpr.h:571-583)
            r.fwd = (s.fwd - strike > 0.0) ? wbar * s.disc / s.num : 0.0;
            r.disc = wbar * intrinsic / s.num;
            r.num = -wbar * intrinsic * s.disc / s.num / s.num;
        }
        return r;
    }
    double wbar;
};

// Up-and-out call/put with smoothed barrier, mcPrd.h:235-288.
template <> struct Product<CF_PRODUCT_UOC> {
    double strike, barSmooth, minusSmooth, twoSmooth;
    double alive, euro;
    bool   killed, isPut;
    // reverse state
    double abar, aliveCur, eurobar;

    __device__ void init(const KArgs& a)
    {
        strike = a.strike; isPut = a.is_put != 0;
        twoSmooth = 2 * a.smooth; barSmooth = a.barrier + a.smooth; minusSmooth = a.barrier - a.smooth;
        alive = 1.0; euro = 0.0; killed = false;
    }
    __device__ void observe(int e, int nEvents, const Sample& s)
    {
        if (!killed) {
            if (s.fwd > barSmooth) { killed = true; alive = 0.0; }
            else if (s.fwd > minusSmooth) alive *= (barSmooth - s.fwd) / twoSmooth;
        }
        if (e == nEvents - 1)
            euro = (isPut ? fmax(strike - s.fwd, 0.0) : fmax(s.fwd - strike, 0.0)) / s.num;
    }
    __device__ void payoffs(double* out) const { out[0] = alive * euro; out[1] = euro; }

    __device__ void begin_reverse(const double* w)
    {
        eurobar = w[0] * alive + w[1];          // alive is the constant 0 on killed paths
        abar = killed ? 0.0 : w[0] * euro;      // a killed `alive` is a fresh leaf: nothing flows
        aliveCur = alive;
    }
    __device__ SampleAdj reverse(int e, int nEvents, const Sample& s)
    {
        SampleAdj r = {0.0, 0.0, 0.0};
        if (e == nEvents - 1) {
            const double x = isPut ? strike - s.fwd : s.fwd - strike;
            if (x > 0.0) r.fwd = (isPut ? -eurobar : eurobar) / s.num;
            r.num = -eurobar * euro / s.num;
        }
        if (!killed && s.fwd > minusSmooth) {   // fuzzy sample (s.fwd <= barSmooth on a live path)
            const double f = (barSmooth - s.fwd) / twoSmooth;
            // alive before this sample; f == 0 only if the spot sits exactly on the upper edge
            const double alivePrev = (f != 0.0) ? aliveCur / f : 0.0;
            r.fwd += abar * alivePrev * (-1.0 / twoSmooth);
            abar *= f;
            aliveCur = alivePrev;
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
    uint32_t* dirlow;    // [dim][8]
    uint32_t* base;      // [2][dim]
};

__host__ __device__ inline size_t align16(size_t x) { return (x + 15) & ~size_t(15); }

template <int MDL>
__host__ __device__ inline int table_a_size(int nSteps, int nKnots) { return MDL == CF_MODEL_DUPIRE ? nSteps * nKnots : nSteps; }
template <int MDL>
__host__ __device__ inline int table_b_size(int nSteps, int nKnots) { return MDL == CF_MODEL_DUPIRE ? nKnots : nSteps; }
// adjoint slots accumulated through the per-step block table
template <int MDL>
__host__ __device__ inline int adj_table_size(int nSteps, int nKnots, int nEvents)
{
    return MDL == CF_MODEL_DUPIRE ? nSteps * nKnots : 2 * nSteps + 3 * nEvents;
}
template <int MDL>
__host__ __device__ inline int row_len(int nKnots) { return MDL == CF_MODEL_DUPIRE ? (nKnots > 1 ? nKnots - 1 : 1) : 3; }

template <int MDL, bool AAD>
__host__ __device__ inline size_t smem_bytes(int nSteps, int nKnots, int nEvents, int dim, bool sobol)
{
    size_t s = 0;
    s += align16(sizeof(double) * table_a_size<MDL>(nSteps, nKnots));
    s += align16(sizeof(double) * table_b_size<MDL>(nSteps, nKnots));
    s += align16(sizeof(double) * (nKnots > 0 ? nKnots : 1));
    if (AAD) {
        s += align16(sizeof(double) * adj_table_size<MDL>(nSteps, nKnots, nEvents));
        s += align16(sizeof(double2) * 2 * kWarps * row_len<MDL>(nKnots));
    }
    s += align16(sizeof(double) * kWarps);
    if (sobol) {
        s += align16(sizeof(uint32_t) * dim * kLowBits);
        s += align16(sizeof(uint32_t) * 2 * dim);
    }
    return s;
}

template <int MDL, bool AAD>
__device__ inline Smem carve(unsigned char* p, int nSteps, int nKnots, int nEvents, int dim, bool sobol)
{
    Smem s{};
    s.tabA = reinterpret_cast<double*>(p);  p += align16(sizeof(double) * table_a_size<MDL>(nSteps, nKnots));
    s.tabB = reinterpret_cast<double*>(p);  p += align16(sizeof(double) * table_b_size<MDL>(nSteps, nKnots));
    s.invdx = reinterpret_cast<double*>(p); p += align16(sizeof(double) * (nKnots > 0 ? nKnots : 1));
    if (AAD) {
        s.adj = reinterpret_cast<double*>(p);   p += align16(sizeof(double) * adj_table_size<MDL>(nSteps, nKnots, nEvents));
        s.wrow = reinterpret_cast<double2*>(p); p += align16(sizeof(double2) * 2 * kWarps * row_len<MDL>(nKnots));
    }
    s.red = reinterpret_cast<double*>(p);   p += align16(sizeof(double) * kWarps);
    if (sobol) {
        s.dirlow = reinterpret_cast<uint32_t*>(p); p += align16(sizeof(uint32_t) * dim * kLowBits);
        s.base = reinterpret_cast<uint32_t*>(p);
    }
    return s;
}

// ---------------------------------------------------------------------------------------------
// interp (interp.h:26-63) on the smem row y against knots x: upper_bound, flat extrapolation.
// Returns v; n = left knot of the bucket, t = weight of knot n+1, slope = dv/dx0 (0 when flat).
// ---------------------------------------------------------------------------------------------
struct Interp { double v, t, slope; int n; };

__device__ __forceinline__ Interp interp_row(const double* __restrict__ x, const double* __restrict__ invdx,
                                            const double* __restrict__ y, int m, int p2, double x0)
{
    // ub = number of knots <= x0  (std::upper_bound)
    int ub = 0;
    for (int s = p2; s > 0; s >>= 1) {
        const int c = ub + s;
        if (c <= m && x[c - 1] <= x0) ub = c;
    }
    Interp r;
    if (ub == m) { r.n = (m > 1 ? m - 2 : 0); r.t = (m > 1 ? 1.0 : 0.0); r.v = y[m - 1]; r.slope = 0.0; }
    else if (ub == 0) { r.n = 0; r.t = 0.0; r.v = y[0]; r.slope = 0.0; }
    else {
        const int n = ub - 1;
        const double y1 = y[n], y2 = y[n + 1];
        const double t = (x0 - x[n]) * invdx[n];
        r.n = n; r.t = t; r.slope = (y2 - y1) * invdx[n];
        r.v = y1 + (y2 - y1) * t;
    }
    return r;
}

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
    const Smem sm = carve<MDL, AAD>(smem_raw, D, m, E, a.dim, kSobol);
    const int rowLen = row_len<MDL>(m);
    const int nAdj = adj_table_size<MDL>(D, m, E);

    // ---- stage tables
    for (int i = tid; i < table_a_size<MDL>(D, m); i += kBlock) sm.tabA[i] = a.tabA[i];
    for (int i = tid; i < table_b_size<MDL>(D, m); i += kBlock) sm.tabB[i] = a.tabB[i];
    if (MDL == CF_MODEL_DUPIRE)
        for (int i = tid; i + 1 < m; i += kBlock) sm.invdx[i] = 1.0 / (a.tabB[i + 1] - a.tabB[i]);
    if (AAD) for (int i = tid; i < nAdj; i += kBlock) sm.adj[i] = 0.0;
    if (kSobol) sobol_load_low(sm.dirlow, a.sobol_dir, a.dim);
    int p2 = 1;
    while (p2 * 2 <= m) p2 *= 2;
    __syncthreads();

    const size_t nSlots = size_t(gridDim.x) * kBlock;
    const size_t slot = size_t(blockIdx.x) * kBlock + tid;
    double* histL = a.hist;                         // Dupire: L_i            BS: S_{i+1}
    double* histG = a.hist + size_t(D) * nSlots;    // g_i

    double paySum[kMaxPay] = {0.0, 0.0};
    double aggSum = 0.0, spotBar = 0.0;
    const double logS0 = (MDL == CF_MODEL_DUPIRE) ? log(a.spot) : 0.0;

    for (int batch = blockIdx.x; batch < a.n_batches; batch += gridDim.x) {
        const uint64_t p = uint64_t(batch) * kBlock + tid;     // path within this run
        const bool valid = p < a.n_paths;
        const uint64_t pabs = a.first_path + p;

        // ---- RNG positioning
        SobolThread sob;
        MrgThread mrg;
        double sign = 1.0;
        if (kSobol) {
            const uint32_t n0 = uint32_t(a.first_path + uint64_t(batch) * kBlock + 1);
            const uint32_t H0 = n0 >> kLowBits;
            __syncthreads();                               // previous batch done with base[]
            sobol_block_base(sm.base, a.sobol_dir, a.dim, H0);
            __syncthreads();
            sob.init(uint32_t(pabs + 1), H0);
        } else {
            mrg.init(a.seed1, a.seed2, pabs >> 1, a.mrg_jump);
            sign = (pabs & 1ull) ? -1.0 : 1.0;
        }
        auto gauss = [&](int d) -> double {
            if (kSobol) return inv_normal_cdf(CF_ONEOVER2POW32 * double(sob.state(sm.dirlow, sm.base, a.dim, d)));
            return sign * inv_normal_cdf(mrg_uniform(mrg.next()));
        };
        auto sampleAt = [&](int e, double spotNow) -> Sample {
            Sample s;
            s.fwd = a.fwd_factors ? spotNow * __ldg(a.fwd_factors + e) : spotNow;
            s.num = a.numeraires ? __ldg(a.numeraires + e) : 1.0;
            s.disc = a.discounts ? __ldg(a.discounts + e) : 1.0;
            return s;
        };

        // ---- forward: generatePath + payoffs
        Product<PRD> prd;
        prd.init(a);
        int e = 0;
        double X = (MDL == CF_MODEL_DUPIRE) ? logS0 : a.spot;   // Dupire: log spot, BS: spot
        if (a.is_event[0]) {
            prd.observe(e, E, sampleAt(e, (MDL == CF_MODEL_DUPIRE) ? exp(X) : X));
            ++e;
        }
        for (int i = 0; i < D; ++i) {
            const double g = gauss(i);
            if (MDL == CF_MODEL_DUPIRE) {
                if (AAD) { histL[size_t(i) * nSlots + slot] = X; histG[size_t(i) * nSlots + slot] = g; }
                const Interp it = interp_row(sm.tabB, sm.invdx, sm.tabA + i * m, m, p2, X);
                X += it.v * (-0.5 * it.v + g);                 // mcMdlDupire.h:271
                if (a.is_event[i + 1]) { prd.observe(e, E, sampleAt(e, exp(X))); ++e; }
            } else {
                X = X * exp(sm.tabA[i] + sm.tabB[i] * g);       // mcMdlBS.h:343
                if (AAD) { histL[size_t(i) * nSlots + slot] = X; histG[size_t(i) * nSlots + slot] = g; }
                prd.observe(e, E, sampleAt(e, X)); ++e;         // every BS step ends on an event date
            }
        }
        double pay[kMaxPay] = {0.0, 0.0};
        prd.payoffs(pay);
        double agg = 0.0;
#pragma unroll
        for (int k = 0; k < kMaxPay; ++k) if (k < a.n_payoffs) agg += a.w[k] * pay[k];
        if (valid) {
#pragma unroll
            for (int k = 0; k < kMaxPay; ++k) if (k < a.n_payoffs) paySum[k] += pay[k];
            aggSum += agg;
            if (a.per_path_payoffs)
                for (int k = 0; k < a.n_payoffs; ++k) a.per_path_payoffs[p * a.n_payoffs + k] = pay[k];
            if (a.per_path_agg) a.per_path_agg[p] = agg;
        }

        // ---- reverse sweep (block-synchronous: one barrier per step)
        if (AAD) {
            prd.begin_reverse(a.w);
            double Xbar = 0.0;          // adjoint of X_{i+1} (Dupire: log spot; BS: spot)
            int er = E - 1;
            int buf = 0;
            for (int i = D - 1; i >= 0; --i, buf ^= 1) {
                double2* myRow = sm.wrow + size_t(buf * kWarps + warp) * rowLen;
                const double g = histG[size_t(i) * nSlots + slot];
                if (MDL == CF_MODEL_DUPIRE) {
                    if (a.is_event[i + 1]) {
                        const double S = exp(X);
                        const SampleAdj sa = prd.reverse(er, E, sampleAt(er, S));
                        const double ff = a.fwd_factors ? __ldg(a.fwd_factors + er) : 1.0;
                        Xbar += sa.fwd * ff * S;            // S = exp(L)
                        --er;
                    }
                    const double L = histL[size_t(i) * nSlots + slot];
                    const Interp it = interp_row(sm.tabB, sm.invdx, sm.tabA + i * m, m, p2, L);
                    const double vbar = valid ? Xbar * (g - it.v) : 0.0;
                    const double bb = vbar * it.t;
                    warp_keyed_accumulate(myRow, rowLen, it.n, vbar - bb, bb);
                    Xbar += vbar * it.slope;
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
                    const double S1 = X;
                    const Sample smp = sampleAt(er, S1);
                    const SampleAdj sa = prd.reverse(er, E, smp);
                    const double ff = a.fwd_factors ? __ldg(a.fwd_factors + er) : 1.0;
                    Xbar += sa.fwd * ff;
                    const double abar = valid ? Xbar * S1 : 0.0;           // adjoint of drift_i + std_i g_i
                    const double ei = exp(sm.tabA[i] + sm.tabB[i] * g);
                    // dense per-step values: drift, std | numeraire, fwd factor, discount of event er
                    double v0 = warp_sum(abar), v1 = warp_sum(abar * g);
                    double v2 = warp_sum(valid ? sa.num : 0.0), v3 = warp_sum(valid ? sa.fwd * S1 : 0.0);
                    double v4 = warp_sum(valid ? sa.disc : 0.0);
                    if (lane == 0) { myRow[0] = make_double2(v0, v1); myRow[1] = make_double2(v2, v3); myRow[2] = make_double2(v4, 0.0); }
                    Xbar *= ei;
                    X = (i > 0) ? histL[size_t(i - 1) * nSlots + slot] : a.spot;
                    __syncthreads();
                    if (tid < 5) {
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
            if (a.is_event[0]) {
                const double S = (MDL == CF_MODEL_DUPIRE) ? exp(X) : X;
                const Sample smp = sampleAt(0, S);
                const SampleAdj sa = prd.reverse(0, E, smp);
                const double ff = a.fwd_factors ? __ldg(a.fwd_factors) : 1.0;
                Xbar += (MDL == CF_MODEL_DUPIRE) ? sa.fwd * ff * S : sa.fwd * ff;
                if (MDL == CF_MODEL_BS) {
                    double v2 = warp_sum(valid ? sa.num : 0.0), v3 = warp_sum(valid ? sa.fwd * S : 0.0);
                    double v4 = warp_sum(valid ? sa.disc : 0.0);
                    __syncthreads();
                    if (lane == 0) { sm.wrow[warp * rowLen] = make_double2(v2, v3); sm.wrow[warp * rowLen + 1] = make_double2(v4, 0.0); }
                    __syncthreads();
                    if (tid < 3) {
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
            if (valid) spotBar += (MDL == CF_MODEL_DUPIRE) ? Xbar / a.spot : Xbar;
        }
    }

    // ---- block results -> partial[blockIdx]
    double* out = a.partial + size_t(blockIdx.x) * a.partial_stride;
    for (int k = 0; k < a.n_payoffs && k < kMaxPay; ++k) {
        const double s = block_sum(paySum[k], sm.red);
        if (tid == 0) out[k] = s;
    }
    if (AAD) {
        double s = block_sum(aggSum, sm.red);
        if (tid == 0) out[a.n_payoffs] = s;
        s = block_sum(spotBar, sm.red);
        if (tid == 0) out[a.n_payoffs + 1] = s;
        __syncthreads();
        for (int i = tid; i < nAdj; i += kBlock) out[a.n_payoffs + 2 + i] = sm.adj[i];
    }
}

// out[k] = sum over blocks (fixed order) of partial[b][k]
__global__ void reduce_partials_kernel(const double* __restrict__ partial, int nBlocks, int stride, int n,
                                       double* __restrict__ out)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    double s = 0.0;
    for (int b = 0; b < nBlocks; ++b) s += partial[size_t(b) * stride + k];
    out[k] = s;
}

}  // namespace cf
