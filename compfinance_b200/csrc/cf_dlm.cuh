// cf_dlm.cuh -- multi-asset displaced-lognormal model x {Autocall, Baskets, MultiStats}: value and AAD.
//
// Replaces MultiDisplaced::generatePath (mcMdlMultiDisplaced.h:653-717) + Autocall / Baskets /
// MultiStats::payoffs (mcPrdMulti.h:388-439, 274-286, 141-178) under the loops of mcBase.h:378-386 /
// 680-704, and the per-path tape sweep on the AAD side.  Adjoint equations: SURVEY.md Appendix A.3.
//
// One path per thread, RNG + generatePath + payoffs fused; one block of NW warps per SM (12 with mrg32k3a: 168
// registers, 8 with Sobol whose window arithmetic wants 256-path batches).  Per step the warp draws its A Gaussians
// (step-major dimension i * A + k) into a per-warp queue, every thread correlates them with the lower Cholesky factor
// (coefficients in the constant bank) and advances its assets with ONE branch-free step for the four dynamics
//     S' = normal ? fwd + std cw : (fwd + sa) exp(drift + std cw) - sa        sa = 0, +alpha, -alpha
// so that the A chains of a step interleave.  The reverse sweep reads spots, Gaussians and the alive notional back from
// a per-thread history ([step][row][slot], coalesced) and needs sums over the 32 paths of a warp of
//   * the per-step table adjoints (dynFwd, drifts, stds [D][A]; forward factors [E][A], numeraires [E]), and
//   * the Cholesky adjoint  cholBar[k][j] = sum over paths and steps of cwBar_k w_j  (an outer product).
// Both go through a per-warp scratch block of rows [value][34]: every lane writes its values of the step once, then
// lane r adds row r (16 x LDS.128, conflict free) and sends the sum to the warp's own table in L2 with a
// fire-and-forget RED (same lane, same address: program order), and lane p adds the products of its (k, j) pairs,
// kept in registers over the whole run.  One __syncwarp pair per step.  Alpha adjoints accumulate in thread-private
// shared-memory columns, spot adjoints go through the rows once per path.  All orders are fixed by lane / warp / block
// index: results are bit-reproducible.
//
// mrg32k3a: antithetic partners (lanes 2 j, 2 j + 1 when the shard starts on an even path) share their Gaussians --
// both step the generator, each converts every other number -- and a thread reaches its next path by ONE or two jump
// matrices from the previous one (blocks own contiguous batch ranges) instead of a full skip-ahead.
#pragma once

#include "cf_kernels.cuh"

namespace cf {

struct LArgs {
    uint64_t first_path, n_paths;
    int      n_batches;            // batches of NW * 32 paths
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
    int      has_alpha;            // some asset is sur- or subnormal
    int      steps_in_smem;        // the [D][A] step tables fit in shared memory
    const double4* step_pack;      // [D][A]: (dynFwd, drift, std, -), the three step tables side by side
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
    double*  hist;                 // [D][2 A + 1][grid * NW * 32]
    double*  warp_tab;             // [grid * NW][dlm_step_tables + A], zeroed before the launch (AAD)
    double*  alpha_cols;           // [A][grid * NW * 32], zeroed before the launch (AAD with sur- / subnormal assets)
};

// Layout of the table-adjoint vector of the displaced model (after the aggregate):
//   spots [A] | alphas [A] | chol [A][A] | dynFwd [D][A] | drifts [D][A] | stds [D][A] | numeraires [E] | ff [E][A]
__host__ __device__ inline int dlm_adj_size(int A, int D, int E) { return 2 * A + A * A + 3 * D * A + E + E * A; }
__host__ __device__ inline int dlm_step_tables(int A, int D, int E) { return 3 * D * A + E + E * A; }
__host__ __device__ inline int dlm_warp_tab(int A, int D, int E) { return dlm_step_tables(A, D, E) + A; }

constexpr int kDlmRow = 34;         // doubles per scratch row: 32 lanes, rows 272 bytes apart (16-byte aligned, conflict-free LDS.128)
constexpr int kDlmStepSmem = 24576; // step tables kept in shared memory up to this many bytes

struct LSmemSizes { size_t pay, red, evc, inum, asset, steps, logt, gq, dirlow, base, alpha, work, total; };

// more than 12 assets: the rows of the outer product reuse those of the table adjoints (two phases per step)
__host__ __device__ inline int dlm_rows(int A, int AMAX) { return AMAX > 12 ? 4 * A + 1 : 6 * A + 1; }
__host__ __device__ inline bool dlm_steps_fit(int A, int D) { return size_t(D) * A * 32 <= size_t(kDlmStepSmem); }

__host__ __device__ inline LSmemSizes dlm_smem(int A, int AMAX, int D, int E, int nPay, int dim, bool sobol, bool aad, int nWarps, bool hasAlpha, bool stepsInSmem)
{
    LSmemSizes s{};
    s.pay = align16(sizeof(double) * nWarps * size_t(nPay));
    s.red = align16(sizeof(double) * nWarps);
    s.evc = align16(sizeof(double) * (size_t(E) * A + 4));
    s.inum = align16(sizeof(double) * size_t(E));
    s.asset = sizeof(double) * 4 * size_t(AMAX);
    s.steps = stepsInSmem ? sizeof(double) * 4 * (size_t(D) * A + 4) : 0;            // groups of four assets read past the last one
    s.logt = sizeof(double) * 2 * 128;
    s.dirlow = sobol ? align16(sizeof(uint32_t) * size_t(dim) * kLowBits) : 0;
    s.base = sobol ? align16(sizeof(uint32_t) * 2 * size_t(dim)) : 0;
    s.alpha = 0;
    s.gq = sizeof(double) * size_t(AMAX) * 32;                                            // forward: the warp's Gaussians of a step,
    const size_t fwd = s.gq + sizeof(uint16_t) * size_t(AMAX) * 32;                       //   then its tail tags
    const size_t scr = aad ? sizeof(double) * size_t(dlm_rows(A, AMAX)) * kDlmRow : 0;          // reverse: the warp's rows (same memory)
    s.work = align16(fwd > scr ? fwd : scr) * nWarps;
    s.total = s.pay + s.red + s.evc + s.inum + s.asset + s.steps + s.logt + s.dirlow + s.base + s.alpha + s.work;
    return s;
}

// Coefficients read as constant-bank operands of the FP64 instructions (a 64-bit immediate costs two moves each time)
static __constant__ double kDlmExp[13] = {
    1.60590438368216145994e-10, 2.08767569878680989792e-09, 2.50521083854417187751e-08, 2.75573192239858906526e-07,
    2.75573192239858906526e-06, 2.48015873015873015873e-05, 1.98412698412698412698e-04, 1.38888888888888888889e-03,
    8.33333333333333333333e-03, 4.16666666666666666667e-02, 1.66666666666666666667e-01, 1.4426950408889634e+00,
    6755399441055744.0};
static __constant__ double kDlmLn2[2] = {-6.93147180559945286227e-01, -2.31904681384629955842e-17};
static __constant__ double kDlmMoroA[4] = {-25.44106049637, 41.39119773534, -18.61500062529, 2.50662823884};
static __constant__ double kDlmMoroB[4] = {3.13082909833, -21.06224101826, 23.08336743743, -8.47351093090};
static __constant__ double kDlmMoroC[9] = {0.3374754822726147, 0.9761690190917186, 0.1607979714918209, 0.0276438810333863,
    0.0038405729373609, 0.0003951896511919, 0.0000321767881768, 0.0000002888167364, 0.0000003960315187};

// exp_core (cf_device.cuh) with its coefficients in the constant bank: same operations, same bits
__device__ __forceinline__ double dlm_exp(double x)
{
    const double kd = fma(x, kDlmExp[11], kDlmExp[12]);
    const int k = __double2loint(kd);
    const double kf = kd - kDlmExp[12];
    double r = fma(kf, kDlmLn2[0], x);
    r = fma(kf, kDlmLn2[1], r);
    double p = kDlmExp[0];
#pragma unroll
    for (int i = 1; i <= 10; ++i) p = fma(p, r, kDlmExp[i]);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
}

// log(x), x positive and normal: 128 entries (c, -log c), c = 11-bit reciprocal of the centre of the mantissa interval
// (one copy: only the Moro tail, one number in six, comes here); r = m c - 1 exact, log1p by its series to r^6; < 2 ulp
__device__ __forceinline__ double dlm_log(double x, const double2* __restrict__ tab)
{
    const int hx = __double2hiint(x);
    const double ed = double((hx >> 20) - 1023);
    const double mant = __hiloint2double((hx & 0x000fffff) | 0x3ff00000, __double2loint(x));
    const double2 t = tab[(uint32_t(hx) >> 13) & 127u];
    const double r = fma(mant, t.x, -1.0);
    double q = fma(r, -1.0 / 6.0, 0.2);
    q = fma(q, r, -0.25);
    q = fma(q, r, 1.0 / 3.0);
    q = fma(q, r, -0.5);
    q = fma(q, r, 1.0);
    return fma(ed, 6.93147180559945286227e-01, fma(r, q, t.y));
}

// Sum of one scratch row (32 doubles, one per lane of the warp that wrote them); lanes reading different rows at the
// same column fall in different 16-byte bank groups (rows are 17 groups apart).  Four partial sums, fixed order.
__device__ __forceinline__ double dlm_row_sum(const double* row)
{
    const double2* r2 = reinterpret_cast<const double2*>(row);
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
    for (int c = 0; c < 16; c += 2) {
        const double2 u = r2[c], v = r2[c + 1];
        s0 += u.x; s1 += u.y; s2 += v.x; s3 += v.y;
    }
    return (s0 + s1) + (s2 + s3);
}

// fire-and-forget addition in L2: no result, no scoreboard
__device__ __forceinline__ void dlm_red(double* addr, double v)
{
    asm volatile("red.global.add.f64 [%0], %1;" :: "l"(addr), "d"(v) : "memory");
}

// Rounds of 32 (k, j <= k) pairs of the Cholesky adjoint owned by a lane
template <int AMAX> struct DlmPairs { static constexpr int kRounds = (AMAX * (AMAX + 1) / 2 + 31) / 32; };

template <int AMAX, int PRD, bool AAD, int RNGK, int NW>
__global__ void __launch_bounds__(NW * 32, 1) dlm_kernel(const LArgs a)
{
    constexpr int kT = NW * 32;
    constexpr bool kSobol = (RNGK == CF_RNG_SOBOL);
    constexpr bool kAuto = (PRD == CF_PRODUCT_AUTOCALL);
    static_assert(!kSobol || NW == kWarps, "the Sobol window arithmetic wants 256-path batches");
    // asset k of the unrolled loops exists: buckets are 4 wide, so only the last three need the test
    #define CF_DLM_HAS(k) ((k) < AMAX - 3 || (k) < A)

    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int A = a.A, D = a.D, E = a.E, nPay = a.n_payoffs;
    const bool hasAlpha = a.has_alpha != 0;
    const bool stepsInSmem = a.steps_in_smem != 0;
    const LSmemSizes z = dlm_smem(A, AMAX, D, E, nPay, a.dim, kSobol, AAD, NW, hasAlpha, stepsInSmem);
    unsigned char* p = smem_raw;
    double* payRows = reinterpret_cast<double*>(p);     p += z.pay;       // [NW][n_payoffs]
    double* red = reinterpret_cast<double*>(p);         p += z.red;
    double* evC = reinterpret_cast<double*>(p);         p += z.evc;       // [E][A]: Autocall ff / reference, else ff
    double* invNum = reinterpret_cast<double*>(p);      p += z.inum;      // [E]: 1 / numeraire
    double4* assetC = reinterpret_cast<double4*>(p);    p += z.asset;     // [A]: (sa, dynamics, 1 / reference, alpha)
    double4* stepS = reinterpret_cast<double4*>(p);     p += z.steps;     // [D][A]: (dynFwd, drift, std, -)
    double2* logT = reinterpret_cast<double2*>(p);      p += z.logt;
    uint32_t* dirlow = reinterpret_cast<uint32_t*>(p);  p += z.dirlow;
    uint32_t* base = reinterpret_cast<uint32_t*>(p);    p += z.base;
    double* work = reinterpret_cast<double*>(p) + size_t(warp) * (z.work / NW / sizeof(double));
    double* gq = work;                                                    // forward: [AMAX][32]
    double* scr = work;                                                   // reverse: [6 A + 1][34]
    uint16_t* tagq = reinterpret_cast<uint16_t*>(reinterpret_cast<unsigned char*>(work) + z.gq);      // forward, after gq

    const int nStepTab = dlm_step_tables(A, D, E);
    const int oFwd = 0, oDrift = D * A, oStd = 2 * D * A, oNum = 3 * D * A, oFf = 3 * D * A + E, oSpot = nStepTab;
    for (int i = tid; i < NW * nPay; i += kT) payRows[i] = 0.0;
    for (int i = tid; i < E * A + 4; i += kT) {
        const double f = i < E * A ? __ldg(a.ff + i) : 0.0;
        evC[i] = (kAuto && i < E * A) ? f / __ldg(a.pweights + (i % A)) : f;
    }
    for (int i = tid; i < E; i += kT) invNum[i] = a.num ? 1.0 / __ldg(a.num + i) : 1.0;
    for (int i = tid; i < AMAX; i += kT) {
        const int dyn = i < A ? __ldg(a.dyn + i) : 0;
        const double al = i < A ? __ldg(a.alphas + i) : 0.0;
        assetC[i] = make_double4(dyn == 2 ? al : dyn == 3 ? -al : 0.0, double(dyn), (kAuto && i < A) ? 1.0 / __ldg(a.pweights + i) : 0.0, al);
    }
    if (stepsInSmem)
        for (int i = tid; i < D * A + 4; i += kT) stepS[i] = a.step_pack[i];          // the pack carries four spare entries
    for (int i = tid; i < 128; i += kT) {
        const double c0 = 1.0 / (1.0 + (double(i) + 0.5) * (1.0 / 128.0));
        const double c = __hiloint2double(__double2hiint(c0) & 0xfffffc00, 0);
        logT[i] = make_double2(c, -log(c));
    }
    if (kSobol) sobol_load_low(dirlow, a.sobol_dir, a.dim);
    __syncthreads();
    double* myPay = payRows + size_t(warp) * nPay;
    double* myTab = AAD ? a.warp_tab + (size_t(blockIdx.x) * NW + warp) * size_t(dlm_warp_tab(A, D, E)) : nullptr;

    const double4* stepP = stepsInSmem ? stepS : a.step_pack;      // shared or global, one code path
    auto stepConst = [&](int idx) -> double4 { return stepP[idx]; };

    SobolThread sob;
    MrgThread mrg, mrgStart;
    // antithetic partners: lane 2 j steps the first component of the pair's generator, lane 2 j + 1 the second
    uint32_t hs0 = 0, hs1 = 0, hs2 = 0;
    const unsigned ltMask = (1u << lane) - 1u;
    // antithetic partners in adjacent lanes share the conversion work
    const bool share = !kSobol && ((a.first_path & 1ull) == 0);
    const int parity = lane & 1;
    const uint32_t hC1 = parity ? 527612u : 1403580u, hC2 = parity ? 1370589u : 810728u;
    const uint32_t hM = parity ? uint32_t(kM2) : uint32_t(kM1), hFold = parity ? 22853u : 209u;
    // one number of the pair's stream (mrg32k3a.h:55-81): own component here, the partner's by shuffle
    auto nextShared = [&]() -> uint32_t {
        const uint64_t pr = uint64_t(hC1) * (parity ? hs0 : hs1) + uint64_t(hC2) * uint64_t(hM - hs2);     // < 2^54
        uint64_t t = uint64_t(uint32_t(pr)) + uint64_t(uint32_t(pr >> 32)) * hFold;                        // < 2^38
        t = uint64_t(uint32_t(t)) + uint64_t(uint32_t(t >> 32) * hFold);                                   // < 2^32 + 2^21
        const uint32_t v = uint32_t(t >= hM ? t - hM : t);
        hs2 = hs1; hs1 = hs0; hs0 = v;
        const uint32_t o = __shfl_xor_sync(kFull, v, 1);
        const uint32_t x = parity ? o : v, y = parity ? v : o;
        return x > y ? x - y : uint32_t(uint64_t(x) + kM1 - y);
    };

    // the Gaussian of one number of the stream, or its tail marker: central branch of invNormalCdf (gaussians.h:73-78)
    // inline, tail (80-86) parked.  Returns true when central.
    auto convert = [&](uint32_t zint, double& val, bool& sup) -> bool {
        const double pu = kSobol ? CF_ONEOVER2POW32 * double(zint) : mrg_uniform(zint);
        sup = pu > 0.5;
        const double up = sup ? 1.0 - pu : pu;
        const double x = up - 0.5;
        const bool central = fabs(x) < 0.42;
        double r = x * x;
        const double num = ((kDlmMoroA[0] * r + kDlmMoroA[1]) * r + kDlmMoroA[2]) * r + kDlmMoroA[3];
        const double den = (((kDlmMoroB[0] * r + kDlmMoroB[1]) * r + kDlmMoroB[2]) * r + kDlmMoroB[3]) * r + 1.0;
        r = div_fast(x * num, den);                 // den in [0.11, 1]: the lean quotient equals the IEEE one (cf_device.cuh)
        val = central ? (sup ? -r : r) : up;
        return central;
    };

    // the A Gaussians of step i of the warp's 32 paths -> gq[k][lane] (value of the even path of an antithetic pair)
    auto fillGauss = [&](int i) {
        int q = 0;
        __syncwarp();
        if (share) {
            for (int k0 = 0; k0 < A; k0 += 2) {
                const uint32_t z0 = nextShared();
                uint32_t z1 = z0;
                if (k0 + 1 < A) z1 = nextShared();
                const int k = k0 + parity;
                double val; bool sup;
                const bool central = convert(parity ? z1 : z0, val, sup) || k >= A;
                if (k < A) { gq[k * 32 + lane] = val; gq[k * 32 + (lane ^ 1)] = val; }
                const unsigned ball = __ballot_sync(kFull, !central);
                if (!central) tagq[q + __popc(ball & ltMask)] = uint16_t((sup ? 0x8000u : 0u) | (unsigned(k) << 5) | unsigned(lane));
                q += __popc(ball);
            }
        } else {
            for (int k = 0; k < A; ++k) {
                const uint32_t zi = kSobol ? sob.state(dirlow, base, a.dim, i * A + k) : mrg.next();
                double val; bool sup;
                const bool central = convert(zi, val, sup);
                gq[k * 32 + lane] = val;
                const unsigned ball = __ballot_sync(kFull, !central);
                if (!central) tagq[q + __popc(ball & ltMask)] = uint16_t((sup ? 0x8000u : 0u) | (unsigned(k) << 5) | unsigned(lane));
                q += __popc(ball);
            }
        }
        __syncwarp();
        for (int b = 0; b < q; b += 32) {
            const int idx = b + lane;
            if (idx < q) {
                const unsigned t = tagq[idx];
                const int slot = int(t & 0x7fffu);            // k * 32 + lane
                double r = dlm_log(-dlm_log(gq[slot], logT), logT);
                double g = kDlmMoroC[8];
#pragma unroll
                for (int c = 7; c >= 0; --c) g = fma(g, r, kDlmMoroC[c]);
                g = (t & 0x8000u) ? g : -g;
                gq[slot] = g;
                if (share) gq[slot ^ 1] = g;
            }
        }
        __syncwarp();
    };

    const uint32_t nSlots = uint32_t(gridDim.x) * kT;
    const uint32_t slot = uint32_t(blockIdx.x) * kT + tid;
    const uint32_t hRows = uint32_t(2 * A + 1);                // per step: spots after the step [A], Gaussians [A], alive before the event
    double aggSum = 0.0;
    // Cholesky adjoint: pair p = k (k + 1) / 2 + j of round q is owned by lane p - 32 q of every warp
    constexpr int kRounds = DlmPairs<AMAX>::kRounds;
    double cholAcc[kRounds];
    int pairX[kRounds], pairW[kRounds];                        // offsets of the pair's rows in the warp's scratch, -1: no pair
    constexpr bool kTwoPhase = AMAX > 12;
    const int rFf = 0, rNum = A, rFwd = A + 1, rDrift = 2 * A + 1, rStd = 3 * A + 1, nTabRows = 4 * A + 1;
    const int rCwb = kTwoPhase ? 0 : 4 * A + 1, rW = kTwoPhase ? A : 5 * A + 1;
    if (AAD) {
#pragma unroll
        for (int q = 0; q < kRounds; ++q) {
            const int pp = lane + 32 * q;
            int k = 0;
            while ((k + 1) * (k + 2) / 2 <= pp) ++k;
            const int j = pp - k * (k + 1) / 2;
            pairX[q] = k < A ? (rCwb + k) * kDlmRow : -1;
            pairW[q] = (rW + j) * kDlmRow;
            cholAcc[q] = 0.0;
        }
    }
    // table entry of row r = lane + 32 q of the step's rows: base + A * event (forward factors), + event (numeraire),
    // + A * step (step tables); -1: no such row / no numeraire table
    constexpr int kFlush = (4 * AMAX + 1 + 31) / 32;
    int flushBase[kFlush], flushMul[kFlush];
    if (AAD) {
#pragma unroll
        for (int q = 0; q < kFlush; ++q) {
            const int r = lane + 32 * q;
            if (r < A) { flushBase[q] = oFf + r; flushMul[q] = 0; }
            else if (r == A) { flushBase[q] = a.num ? oNum : -1; flushMul[q] = 1; }
            else if (r < nTabRows) { const int t = (r - rFwd) / A; flushBase[q] = (t == 0 ? oFwd : t == 1 ? oDrift : oStd) + (r - rFwd - t * A); flushMul[q] = 2; }
            else { flushBase[q] = -1; flushMul[q] = 2; }
        }
    }
    const double inv2s = 1.0 / (2.0 * a.smooth), invStrike = 1.0 / a.strike, cpn = a.coupon * a.cpn_dt;

    // contiguous batches per block: a thread's next path is kT / 2 antithetic pairs after its last one
    const int bBeg = int(int64_t(blockIdx.x) * a.n_batches / gridDim.x), bEnd = int(int64_t(blockIdx.x + 1) * a.n_batches / gridDim.x);
    for (int batch = bBeg; batch < bEnd; ++batch) {
        const uint64_t pidx = uint64_t(batch) * kT + tid;
        const bool valid = pidx < a.n_paths;
        const uint64_t pabs = a.first_path + pidx;
        double sign = 1.0;
        if (kSobol) {
            const uint32_t n0 = uint32_t(a.first_path + uint64_t(batch) * kT + 1);
            const uint32_t H0 = n0 >> kLowBits;
            __syncthreads();
            sobol_block_base(base, a.sobol_dir, a.dim, H0);
            __syncthreads();
            sob.init(uint32_t(pabs + 1), H0);
        } else {
            if (batch == bBeg) mrgStart.init(a.seed1, a.seed2, pabs >> 1, a.mrg_jump);
            else mrgStart.advance(uint64_t(kT / 2), a.mrg_jump);
            mrg = mrgStart;
            if (share) {
                hs0 = parity ? mrgStart.y0 : mrgStart.x0; hs1 = parity ? mrgStart.y1 : mrgStart.x1; hs2 = parity ? mrgStart.y2 : mrgStart.x2;
            }
            sign = (pabs & 1ull) ? -1.0 : 1.0;
        }

        // ---- payoff accumulation helpers
        double agg = 0.0;
        auto emit = [&](int k, double v) {                 // payoff k of this path
            const double s = warp_sum(valid ? v : 0.0);
            if (lane == 0) myPay[k] += s;
            if (AAD) agg += __ldg(a.w + k) * v;
            if (valid && a.per_path_payoffs) a.per_path_payoffs[pidx * nPay + k] = v;
        };

        // ---- forward
        double S[AMAX], Fprev[AMAX];
#pragma unroll
        for (int k = 0; k < AMAX; ++k) { S[k] = (k < A) ? __ldg(a.spots + k) : 0.0; Fprev[k] = 0.0; }
        double alive = 1.0, pay = 0.0;                     // Autocall state
        int payIdx = 0;                                    // MultiStats running payoff index

        auto observe = [&](int e) {                        // event date e with the current spots
            const double cn = invNum[e];
            const double* ec = evC + e * A;
            if (kAuto) {
                double worst = S[0] * ec[0];
#pragma unroll
                for (int k = 1; k < AMAX; ++k)
                    if (CF_DLM_HAS(k)) { const double pf = S[k] * ec[k]; worst = pf < worst ? pf : worst; }
                pay = fma(alive * cpn, cn, pay);
                if (e < E - 1) {
                    const double f = fmin(1.0, fmax(0.0, (a.ko + a.smooth - worst) * inv2s));
                    const double surv = alive * f;
                    pay = fma(alive - surv, cn, pay);
                    alive = surv;
                } else {
                    pay = fma(alive, cn, pay);
                    pay -= alive * fmax(a.strike - worst, 0.0) * invStrike * cn;
                }
            } else if (PRD == CF_PRODUCT_BASKETS) {
                double b = 0.0;
#pragma unroll
                for (int k = 0; k < AMAX; ++k) if (CF_DLM_HAS(k)) b += __ldg(a.pweights + k) * (S[k] * ec[k]);
                for (int k = 0; k < a.n_strikes; ++k) emit(k, fmax(b - __ldg(a.strikes + k), 0.0) * cn);
            } else {                                       // MultiStats: levels now, differences after all levels
                double F[AMAX];
#pragma unroll
                for (int k = 0; k < AMAX; ++k) F[k] = (k < A) ? S[k] * ec[k] : 0.0;
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
            fillGauss(i);
            double w[AMAX];
#pragma unroll
            for (int k = 0; k < AMAX; ++k) w[k] = CF_DLM_HAS(k) ? sign * gq[k * 32 + lane] : 0.0;
            double* h = AAD ? a.hist + size_t(uint32_t(i) * hRows) * nSlots + slot : nullptr;
            if (AAD) {
                h[(2u * uint32_t(A)) * nSlots] = alive;
#pragma unroll
                for (int k = 0; k < AMAX; ++k) if (CF_DLM_HAS(k)) h[(uint32_t(A) + k) * nSlots] = w[k];
            }
#pragma unroll
            for (int k = 0; k < AMAX; ++k) {
                if (CF_DLM_HAS(k)) {
                    double cw = 0.0;
#pragma unroll
                    for (int j = 0; j <= k; ++j) cw = fma(a.cholv[k * (k + 1) / 2 + j], w[j], cw);
                    const double4 sc = stepConst(i * A + k);
                    const double4 ac = assetC[k];
                    const double fwd = S[k] * sc.x;
                    const double ex = dlm_exp(fma(sc.z, cw, sc.y));
                    const double sLog = fma(fwd + ac.x, ex, -ac.x);
                    const double sNor = fma(sc.z, cw, fwd);
                    S[k] = ac.y == 1.0 ? sNor : sLog;
                }
            }
            if (AAD) {
#pragma unroll
                for (int k = 0; k < AMAX; ++k) if (CF_DLM_HAS(k)) h[uint32_t(k) * nSlots] = S[k];
            }
            observe(e); ++e;
        }
        if (kAuto) emit(0, pay);
        if (AAD) {
            if (valid) { aggSum += agg; if (a.per_path_agg) a.per_path_agg[pidx] = agg; }
        }

        // ---- reverse sweep
        if (AAD) {
            double Sbar[AMAX];
#pragma unroll
            for (int k = 0; k < AMAX; ++k) Sbar[k] = 0.0;
            // a lane past the end of the run sweeps with zero seeds: every adjoint it writes is an exact zero
            const double paybar = (kAuto && valid) ? __ldg(a.w) : 0.0;
            double alivebar = 0.0;                          // adjoint of the notional alive AFTER the current event
            // adjoint of the sample of event ev, spots Sev, notional alive before the event: rows rFf .. rNum
            auto reverseEvent = [&](int ev, const double* Sev, double aliveBefore) {
                const double cn = invNum[ev];
                const double* ec = evC + ev * A;
                double numbar = 0.0;
                if (kAuto) {
                    double worst = Sev[0] * ec[0];
                    int am = 0;
#pragma unroll
                    for (int k = 1; k < AMAX; ++k)
                        if (CF_DLM_HAS(k)) { const double pf = Sev[k] * ec[k]; if (pf < worst) { worst = pf; am = k; } }
                    double worstbar;
                    if (ev < E - 1) {
                        const double q = (a.ko + a.smooth - worst) * inv2s;
                        const double f = fmin(1.0, fmax(0.0, q));
                        // surv = alive f; pay += alive cpn dt / num + alive (1 - f) / num
                        const double fbar = alivebar * aliveBefore - paybar * aliveBefore * cn;
                        const double qbar = (q > 0.0 && q < 1.0) ? fbar : 0.0;      // max(0, .) then min(1, .), strict (AADExpr.h:571-598)
                        worstbar = -qbar * inv2s;
                        numbar = -paybar * (aliveBefore * cpn + (aliveBefore - aliveBefore * f)) * cn * cn;
                        alivebar = alivebar * f + paybar * cpn * cn + paybar * (1.0 - f) * cn;
                    } else {
                        const double put = fmax(a.strike - worst, 0.0);
                        worstbar = (a.strike - worst > 0.0) ? paybar * aliveBefore * invStrike * cn : 0.0;
                        numbar = -paybar * (aliveBefore * cpn + aliveBefore - aliveBefore * put * invStrike) * cn * cn;
                        alivebar = paybar * (cpn * cn + cn - put * invStrike * cn);
                    }
                    // worst = S ff / ref of the worst performer only
#pragma unroll
                    for (int k = 0; k < AMAX; ++k) {
                        if (CF_DLM_HAS(k)) {
                            const double wb = (k == am) ? worstbar : 0.0;
                            scr[(rFf + k) * kDlmRow + lane] = wb * Sev[k] * assetC[k].z;
                            Sbar[k] = fma(wb, ec[k], Sbar[k]);
                        }
                    }
                } else if (PRD == CF_PRODUCT_BASKETS) {
                    double b = 0.0;
#pragma unroll
                    for (int k = 0; k < AMAX; ++k) if (CF_DLM_HAS(k)) b += __ldg(a.pweights + k) * (Sev[k] * ec[k]);
                    double bbar = 0.0;
                    for (int k = 0; k < a.n_strikes; ++k) {
                        const double x = b - __ldg(a.strikes + k);
                        const double wk = valid ? __ldg(a.w + k) : 0.0;
                        if (x > 0.0) { bbar += wk * cn; numbar -= wk * x * cn * cn; }
                    }
#pragma unroll
                    for (int k = 0; k < AMAX; ++k) {
                        if (CF_DLM_HAS(k)) {
                            const double fb = bbar * __ldg(a.pweights + k);
                            scr[(rFf + k) * kDlmRow + lane] = fb * Sev[k];
                            Sbar[k] = fma(fb, ec[k], Sbar[k]);
                        }
                    }
                }
                scr[rNum * kDlmRow + lane] = numbar;
            };
            // sums of rows [0, nRows) over the warp's paths -> the warp's table; event rows first, then the step's
            auto flushRows = [&](int nRows, int ev, int i) {
#pragma unroll
                for (int q = 0; q < kFlush; ++q) {
                    const int r = lane + 32 * q;
                    if (r < nRows && flushBase[q] >= 0) {
                        const double s = dlm_row_sum(scr + r * kDlmRow);
                        dlm_red(myTab + flushBase[q] + (flushMul[q] == 0 ? ev * A : flushMul[q] == 1 ? ev : i * A), s);
                    }
                }
            };

            int er = E - 1;
            for (int i = D - 1; i >= 0; --i) {
                const double* h = a.hist + size_t(uint32_t(i) * hRows) * nSlots + slot;
                const double* hPrev = a.hist + size_t(uint32_t(i > 0 ? i - 1 : 0) * hRows) * nSlots + slot;
                double Sn[AMAX], w[AMAX], cwb[kTwoPhase ? AMAX : 1];
#pragma unroll
                for (int k = 0; k < AMAX; ++k) {
                    Sn[k] = CF_DLM_HAS(k) ? h[uint32_t(k) * nSlots] : 0.0;
                    w[k] = CF_DLM_HAS(k) ? h[(uint32_t(A) + k) * nSlots] : 0.0;
                }
                const double aliveBefore = h[(2u * uint32_t(A)) * nSlots];
                __syncwarp();                               // the rows of the previous step have been read
                reverseEvent(er, Sn, aliveBefore);
                --er;
#pragma unroll
                for (int k = 0; k < AMAX; ++k) {
                    if (CF_DLM_HAS(k)) {
                        double cw = 0.0;
#pragma unroll
                        for (int j = 0; j <= k; ++j) cw = fma(a.cholv[k * (k + 1) / 2 + j], w[j], cw);
                        const double4 sc = stepConst(i * A + k);
                        const double4 ac = assetC[k];
                        const double Sp = i > 0 ? hPrev[uint32_t(k) * nSlots] : __ldg(a.spots + k);
                        const double sb = Sbar[k];
                        const double ex = dlm_exp(fma(sc.z, cw, sc.y));
                        const bool normal = ac.y == 1.0;
                        const double fwdbar = normal ? sb : sb * ex;              // S = fwd + std cw | (fwd + sa) e - sa
                        const double xbar = normal ? 0.0 : sb * (Sn[k] + ac.x);
                        const double y = normal ? sb : xbar;
                        if (ac.y >= 2.0) {                                        // alpha: +-(e - 1) sb; thread-private column: program order
                            const double t = sb * (ex - 1.0);
                            dlm_red(a.alpha_cols + size_t(uint32_t(k) * nSlots) + slot, ac.y == 2.0 ? t : -t);
                        }
                        scr[(rFwd + k) * kDlmRow + lane] = fwdbar * Sp;
                        scr[(rDrift + k) * kDlmRow + lane] = xbar;
                        scr[(rStd + k) * kDlmRow + lane] = y * cw;
                        if (kTwoPhase) cwb[k] = y * sc.z;
                        else {
                            scr[(rCwb + k) * kDlmRow + lane] = y * sc.z;
                            scr[(rW + k) * kDlmRow + lane] = w[k];
                        }
                        Sbar[k] = fwdbar * sc.x;
                    }
                }
                __syncwarp();
                flushRows(nTabRows, er + 1, i);
                if (kTwoPhase) {
                    __syncwarp();
#pragma unroll
                    for (int k = 0; k < AMAX; ++k)
                        if (CF_DLM_HAS(k)) { scr[(rCwb + k) * kDlmRow + lane] = cwb[k]; scr[(rW + k) * kDlmRow + lane] = w[k]; }
                    __syncwarp();
                }
                // outer product cwBar (x) w over the warp's paths
#pragma unroll
                for (int q = 0; q < kRounds; ++q) {
                    if (pairX[q] >= 0) {
                        const double2* xr = reinterpret_cast<const double2*>(scr + pairX[q]);
                        const double2* wr = reinterpret_cast<const double2*>(scr + pairW[q]);
                        double s0 = 0.0, s1 = 0.0;
#pragma unroll
                        for (int c = 0; c < 16; ++c) {
                            const double2 xv = xr[c], wv = wr[c];
                            s0 = fma(xv.x, wv.x, s0);
                            s1 = fma(xv.y, wv.y, s1);
                        }
                        cholAcc[q] += s0 + s1;
                    }
                }
            }
            __syncwarp();
            if (a.today) {
                double S0[AMAX];
#pragma unroll
                for (int k = 0; k < AMAX; ++k) S0[k] = (k < A) ? __ldg(a.spots + k) : 0.0;
                reverseEvent(0, S0, 1.0);
                __syncwarp();
                flushRows(A + 1, 0, 0);
                __syncwarp();
            }
            // spot adjoints of the warp's paths through the rows, once per path
#pragma unroll
            for (int k = 0; k < AMAX; ++k) if (CF_DLM_HAS(k)) scr[k * kDlmRow + lane] = Sbar[k];
            __syncwarp();
            if (lane < A) dlm_red(myTab + oSpot + lane, dlm_row_sum(scr + lane * kDlmRow));
            __syncwarp();
        }
    }

    // ---- block results -> partial[blockIdx]: payoff sums | agg | table adjoints
    if (AAD) __threadfence();
    __syncthreads();
    double* out = a.partial + size_t(blockIdx.x) * a.partial_stride;
    for (int k = tid; k < nPay; k += kT) {
        double s = 0.0;
        for (int w = 0; w < NW; ++w) s += payRows[size_t(w) * nPay + k];
        out[k] = s;
    }
    if (AAD) {
        double s = block_sum(aggSum, red);
        if (tid == 0) out[nPay] = s;
        double* adj = out + nPay + 1;
        const size_t tabStride = size_t(dlm_warp_tab(A, D, E));
        const double* blockTab = a.warp_tab + size_t(blockIdx.x) * NW * tabStride;
        for (int k = tid; k < A; k += kT) {
            double t = 0.0;
            for (int w = 0; w < NW; ++w) t += __ldcg(blockTab + size_t(w) * tabStride + oSpot + k);
            adj[k] = t;
            double u0 = 0.0, u1 = 0.0, u2 = 0.0, u3 = 0.0;
            if (hasAlpha) {
                const double* col = a.alpha_cols + size_t(uint32_t(k) * nSlots) + size_t(blockIdx.x) * kT;
                for (int c = 0; c < kT; c += 4) {
                    u0 += __ldcg(col + c); u1 += __ldcg(col + c + 1);
                    u2 += __ldcg(col + c + 2); u3 += __ldcg(col + c + 3);
                }
            }
            adj[A + k] = (u0 + u1) + (u2 + u3);
        }
        for (int k = tid; k < A * A; k += kT) adj[2 * A + k] = 0.0;
        // Cholesky adjoint: the warps' pair sums through the (now idle) scratch rows, combined in warp order
        const int nPairs = A * (A + 1) / 2;
        const size_t workStride = z.work / NW / sizeof(double);
#pragma unroll
        for (int q = 0; q < kRounds; ++q) scr[lane + 32 * q] = cholAcc[q];
        __syncthreads();
        double* workAll = scr - size_t(warp) * workStride;
        for (int pp = tid; pp < nPairs; pp += kT) {
            double t = 0.0;
            for (int w = 0; w < NW; ++w) t += workAll[size_t(w) * workStride + pp];
            int k = 0;
            while ((k + 1) * (k + 2) / 2 <= pp) ++k;
            adj[2 * A + k * A + (pp - k * (k + 1) / 2)] = t;
        }
        for (int k = tid; k < nStepTab; k += kT) {
            double t = 0.0;
            for (int w = 0; w < NW; ++w) t += __ldcg(blockTab + size_t(w) * tabStride + k);
            adj[2 * A + A * A + k] = t;
        }
    }
    #undef CF_DLM_HAS
}

}  // namespace cf
