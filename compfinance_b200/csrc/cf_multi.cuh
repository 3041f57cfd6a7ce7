// cf_multi.cuh -- itemised AAD risk (one adjoint vector per payoff) for Dupire x Europeans.
//
// Replaces mcSimulAADMulti / mcParallelSimulAADMulti (mcBase.h:776, 859: every tape node carries
// nPay adjoints, Node::propagateAll AADNode.h:85-102) for the portfolio of European calls of
// mcPrd.h:290-401 under the Dupire model (BASELINE config 4: 12 maturities x 60 strikes = 720
// payoffs x 1081 parameters).  SURVEY.md Appendix A.4:
//
//   d payoff(e, k) / d theta = 1{S_e > K_k} dS_e / d theta        (numeraire 1 under Dupire)
//
// and for strikes sorted within a maturity the indicator is a prefix in k.  So a path only needs ONE
// reverse sweep per maturity (seed dS_e/dL = S_e), accumulated into the table of its strike class
// c = #{k : K_k < S_e}; the risk of payoff (e, k) is the sum of the tables of classes > rank(k),
// taken afterwards (suffix sums), together with the time interpolation of init().
//
// Accumulation: every thread adds into [event][class][1 + steps x knots] in L2.  To be bit-reproducible whatever
// the order the adds arrive in, an addend x is split into two FIXED-POINT integers -- hi = rint(x 2^10) and
// lo = rint((x - hi 2^-10) 2^b), b = 50 for up to 2^22 paths -- added with 64-bit integer atomics (associative:
// same bits on every run, every grid, every GPU).  What is dropped is below 2^-(b+1) per addend (1e-15);
// |x| N < 2^52 is required of hi (x < 1e9 for 2^22 paths) and checked by the host against the spot.
#pragma once

#include "cf_kernels.cuh"

namespace cf {

constexpr int kMultiMaxSteps = 64;     // per-thread history (local memory)

struct MArgs {
    uint64_t first_path, n_paths;
    int      n_batches;
    int      strided;              // batches dealt round-robin over the blocks (full skip-ahead per path) or contiguous ranges
    uint32_t seed1, seed2;
    int      dim;
    const uint32_t* sobol_dir;
    const uint64_t* mrg_jump;
    int      D, m, E;
    const uint8_t* is_event;       // [D + 1]
    double   spot;
    const double* interp_vols;     // [D][m]
    const double* log_spots;       // [m]
    const double* ksorted;         // [n_payoffs] strikes, ascending within each event
    const int32_t* koff;           // [E + 1]
    int      n_payoffs, cmax;      // cmax = 1 + max strikes per event
    double*  partial;              // [grid][n_payoffs] payoff sums (sorted order)
    long long* Thi;                // [E][cmax][1 + D * m]: spot adjoint, then interp_vols adjoints, fixed point 2^-10; zeroed by the host
    long long* Tlo;                // the remainders, fixed point 2^-lo_bits
    double   lo_scale;             // 2^lo_bits
    double*  per_path_payoffs;     // [n_paths][n_payoffs] (sorted order) or null
    // bucket lookup by uniform cells, as in the generic kernel (KArgs::lut): 0 entries = binary search
    const uint8_t* lut;
    int      lut_n;
    double   lut_x0, lut_scale;
};

struct MSmem { size_t y, x, invdx, ks, isev, pay, gq, tagq, dirlow, base, lut, total; };

__host__ __device__ inline MSmem multi_smem(int D, int m, int nPay, int dim, bool sobol, int lutN)
{
    MSmem s{};
    s.y = align16(sizeof(double) * size_t(D) * m);
    s.x = align16(sizeof(double) * m);
    s.invdx = align16(sizeof(double) * m);
    s.ks = align16(sizeof(double) * size_t(nPay));
    s.isev = align16(size_t(D) + 1);
    s.pay = align16(sizeof(double) * kWarps * (size_t(nPay) + 32));          // + 32 doubles per warp: the ladders' scratch
    s.gq = align16(sizeof(double) * kWarps * kChunk * 32);
    s.tagq = align16(sizeof(uint16_t) * kWarps * kChunk * 32);
    s.dirlow = sobol ? align16(sizeof(uint32_t) * size_t(dim) * kLowBits) : 0;
    s.base = sobol ? align16(sizeof(uint32_t) * 2 * size_t(dim)) : 0;
    s.lut = align16(size_t(lutN > 0 ? lutN : 1));
    s.total = s.y + s.x + s.invdx + s.ks + s.isev + s.pay + s.gq + s.tagq + s.dirlow + s.base + s.lut;
    return s;
}

template <int RNGK>
__global__ void __launch_bounds__(kBlock, 2) dupire_europeans_multi_kernel(const MArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int D = a.D, m = a.m, nPay = a.n_payoffs;
    constexpr bool kSobol = (RNGK == CF_RNG_SOBOL);
    const MSmem z = multi_smem(D, m, nPay, a.dim, kSobol, a.lut_n);
    unsigned char* p = smem_raw;
    double* ysm = reinterpret_cast<double*>(p);         p += z.y;
    double* xs = reinterpret_cast<double*>(p);          p += z.x;
    double* invdx = reinterpret_cast<double*>(p);       p += z.invdx;
    double* ks = reinterpret_cast<double*>(p);          p += z.ks;
    uint8_t* isev = reinterpret_cast<uint8_t*>(p);      p += z.isev;
    double* payRows = reinterpret_cast<double*>(p);     p += z.pay;
    double* gq = reinterpret_cast<double*>(p);          p += z.gq;
    uint16_t* tagq = reinterpret_cast<uint16_t*>(p);    p += z.tagq;
    uint32_t* dirlow = reinterpret_cast<uint32_t*>(p);  p += z.dirlow;
    uint32_t* base = reinterpret_cast<uint32_t*>(p);    p += z.base;
    uint8_t* lutS = reinterpret_cast<uint8_t*>(p);

    for (int i = tid; i < D * m; i += kBlock) ysm[i] = a.interp_vols[i];
    for (int i = tid; i < m; i += kBlock) xs[i] = a.log_spots[i];
    for (int i = tid; i + 1 < m; i += kBlock) invdx[i] = 1.0 / (a.log_spots[i + 1] - a.log_spots[i]);
    for (int i = tid; i < nPay; i += kBlock) ks[i] = a.ksorted[i];
    for (int i = tid; i <= D; i += kBlock) isev[i] = a.is_event[i];
    for (int i = tid; i < kWarps * nPay; i += kBlock) payRows[i] = 0.0;
    for (int i = tid; i < a.lut_n; i += kBlock) lutS[i] = a.lut[i];
    if (kSobol) sobol_load_low(dirlow, a.sobol_dir, a.dim);
    __syncthreads();
    double* myPay = payRows + size_t(warp) * nPay;
    double* fw = payRows + size_t(kWarps) * nPay + size_t(warp) * 32;

    Locator loc;
    loc.x = xs; loc.lut = lutS; loc.m = m; loc.lutN = a.lut_n; loc.x0 = a.lut_x0; loc.scale = a.lut_scale;
    loc.p2 = 1;
    while (loc.p2 * 2 <= m) loc.p2 *= 2;

    GaussGen<RNGK> gen;
    gen.gq = gq + size_t(warp) * kChunk * 32;
    gen.tagq = tagq + size_t(warp) * kChunk * 32;
    gen.dirlow = dirlow; gen.base = base; gen.dim = a.dim;

    const size_t tabLen = 1 + size_t(D) * m;
    const double logS0 = log(a.spot);

    // contiguous batches per block: with mrg32k3a a thread's next path is kBlock / 2 antithetic pairs down the stream,
    // one jump matrix instead of a full skip-ahead (a third of a short path's instructions)
    const int bBeg = int(int64_t(blockIdx.x) * a.n_batches / gridDim.x), bEnd = int(int64_t(blockIdx.x + 1) * a.n_batches / gridDim.x);
    MrgThread mrgStart;
    const int bFirst = a.strided ? int(blockIdx.x) : bBeg, bLast = a.strided ? a.n_batches : bEnd, bStep = a.strided ? int(gridDim.x) : 1;
    for (int batch = bFirst; batch < bLast; batch += bStep) {
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
            // round-robin batches take the full skip-ahead per path: advancing from the previous batch instead (three jump
            // matrices) measured SLOWER, 9.3 ms against 8.15 ms, like the contiguous ranges -- this kernel's time goes with
            // how its warps fall on the L2 atomics, not with its instruction count
            if (batch == bFirst || a.strided) mrgStart.init(a.seed1, a.seed2, pabs >> 1, a.mrg_jump);
            else mrgStart.advance(uint64_t(kBlock / 2), a.mrg_jump);
            gen.mrg = mrgStart;
            gen.sign = (pabs & 1ull) ? -1.0 : 1.0;
        }

        double Lh[kMultiMaxSteps + 1], gh[kMultiMaxSteps];     // log-spot at every timeline point, Gaussian of every step
        // sample of event e after nst steps at log-spot L: payoffs, then the sweep of its strike class
        auto sample = [&](int e, int nst, double L) {
            const double S = exp(L);
            const int k0 = a.koff[e], k1 = a.koff[e + 1];
            // payoff sums of the maturity's strike ladder; Dupire leaves the numeraire at 1 (mcBase.h:91-99)
            warp_ladder_sums(fw, S, valid, ks + k0, k1 - k0, 1.0, myPay + k0, lane);
            if (valid && a.per_path_payoffs) {
                // rarely taken: kept out of the instruction cache's way
#pragma unroll 1
                for (int k = k0; k < k1; ++k) a.per_path_payoffs[pidx * nPay + k] = fmax(S - ks[k], 0.0);
            }
            // class = #strikes strictly below S (max(x, 0) has derivative 1 iff x > 0, AADExpr.h:571-583)
            int lo = k0, hi = k1;
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (ks[mid] < S) lo = mid + 1; else hi = mid; }
            const int c = lo - k0;
            if (!valid || c == 0) return;
            const size_t tOff = (size_t(e) * a.cmax + c) * tabLen;
            unsigned long long* Thi = reinterpret_cast<unsigned long long*>(a.Thi) + tOff;
            unsigned long long* Tlo = reinterpret_cast<unsigned long long*>(a.Tlo) + tOff;
            const double loScale = a.lo_scale;
            auto addFixed = [&](size_t q, double x) {              // order-independent: two integer atomics
                const long long hi = __double2ll_rn(x * 1024.0);
                const long long lo = __double2ll_rn(fma(double(hi), -1.0 / 1024.0, x) * loScale);
                if (hi) atomicAdd(Thi + q, static_cast<unsigned long long>(hi));
                if (lo) atomicAdd(Tlo + q, static_cast<unsigned long long>(lo));
            };
            double Lbar = S;                                         // dS/dL
            for (int i = nst - 1; i >= 0; --i) {
                const double Li = Lh[i];
                const Bucket b = loc.locate(Li);
                const double* y = ysm + i * m;
                double v, t, slope;
                if (b.side != 0) {
                    const bool right = b.side > 0;
                    v = y[right ? m - 1 : 0]; t = (right && m > 1) ? 1.0 : 0.0; slope = 0.0;
                } else {
                    const double y1 = y[b.n], dy = y[b.n + 1] - y1, idx = invdx[b.n];
                    t = (Li - xs[b.n]) * idx;
                    v = y1 + dy * t;
                    slope = dy * idx;
                }
                const double vbar = Lbar * (gh[i] - v);
                const double bb = vbar * t;
                const size_t row = 1 + size_t(i) * m + b.n;
                if (vbar - bb != 0.0) addFixed(row, vbar - bb);
                if (bb != 0.0) addFixed(row + 1, bb);
                Lbar += vbar * slope;
            }
            addFixed(0, Lbar / a.spot);                              // L0 = log(S0), mcMdlDupire.h:245
        };

        double X = logS0;
        for (int i0 = 0; i0 < D; i0 += kChunk) {
            const int cnt = min(kChunk, D - i0);
            gen.fill(i0, cnt);
            for (int k = 0; k < cnt; ++k) {
                const int i = i0 + k;
                const double g = gen.get(k);
                Lh[i] = X; gh[i] = g;
                const Bucket b = loc.locate(X);
                const double* y = ysm + i * m;
                double v;
                if (b.side != 0) v = y[b.side > 0 ? m - 1 : 0];
                else { const double y1 = y[b.n]; v = y1 + (y[b.n + 1] - y1) * ((X - xs[b.n]) * invdx[b.n]); }
                X += v * (-0.5 * v + g);                              // mcMdlDupire.h:271
            }
        }
        Lh[D] = X;
        // the samples after the path, from ONE call site: inlined into the (unrolled) step loop the sweep was replicated
        // nine times, 190 KB of code, and the kernel waited on instruction fetch four cycles out of five
        int e = 0;
#pragma unroll 1
        for (int pt = 0; pt <= D; ++pt)
            if (isev[pt]) { sample(e, pt, Lh[pt]); ++e; }
    }

    __syncthreads();
    double* out = a.partial + size_t(blockIdx.x) * nPay;
    for (int k = tid; k < nPay; k += kBlock) {
        double s = 0.0;
        for (int w = 0; w < kWarps; ++w) s += payRows[size_t(w) * nPay + k];
        out[k] = s;
    }
}

// T[e][c] <- sum over classes c' >= c of the fixed-point tables (integer suffix sums: exact), converted to double;
// one thread per table entry
static __global__ void multi_suffix_kernel(const long long* __restrict__ Thi, const long long* __restrict__ Tlo, double loInv,
                                           double* __restrict__ T, int E, int cmax, size_t tabLen)
{
    const size_t idx = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= size_t(E) * tabLen) return;
    const size_t e = idx / tabLen, q = idx % tabLen;
    long long hi = 0, lo = 0;
    for (int c = cmax - 1; c >= 0; --c) {
        const size_t cell = (e * cmax + c) * tabLen + q;
        hi += Thi[cell]; lo += Tlo[cell];
        T[cell] = fma(double(lo), loInv, double(hi) * (1.0 / 1024.0));
    }
}

// risks[param][payoff] (sum over paths, not yet divided by N): param 0 = spot, then vols[j][kt] spot-major.
//   payoff p = (event e, sorted rank r): classes > r, i.e. the suffix table of class r + 1
//   vols[j][kt] <- sum_i (k1[i] == kt ? c1[i] : 0) + (k2[i] == kt ? c2[i] : 0)) * ybar[i][j]      (mcMdlDupire.h:202-216)
static __global__ void multi_collapse_kernel(const double* __restrict__ T, int E, int cmax, int D, int m, int nTimes,
                                      const int32_t* __restrict__ k1, const int32_t* __restrict__ k2,
                                      const double* __restrict__ c1, const double* __restrict__ c2,
                                      const int32_t* __restrict__ payEvent, const int32_t* __restrict__ payRank,
                                      const int32_t* __restrict__ payOrig, int nPay, double* __restrict__ out)   // out: [nParam][nPay]
{
    const size_t idx = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    const size_t nParam = 1 + size_t(m) * nTimes;
    if (idx >= nParam * nPay) return;
    const int ps = int(idx % nPay);                // payoff in sorted order
    const size_t param = idx / nPay;
    const int e = payEvent[ps], cls = payRank[ps] + 1;
    const size_t tabLen = 1 + size_t(D) * m;
    double s = 0.0;
    if (cls < cmax) {
        const double* U = T + (size_t(e) * cmax + cls) * tabLen;
        if (param == 0) s = U[0];
        else {
            const int j = int((param - 1) / nTimes), kt = int((param - 1) % nTimes);
            for (int i = D - 1; i >= 0; --i) {
                const double y = U[1 + size_t(i) * m + j];
                if (k1[i] == kt) s += c1[i] * y;
                if (k2[i] == kt) s += c2[i] * y;
            }
        }
    }
    out[param * nPay + payOrig[ps]] = s;
}

// payoff sums from sorted to original order
static __global__ void multi_unsort_kernel(const double* __restrict__ sorted, const int32_t* __restrict__ payOrig, int nPay,
                                           double* __restrict__ out)
{
    const int ps = blockIdx.x * blockDim.x + threadIdx.x;
    if (ps < nPay) out[payOrig[ps]] = sorted[ps];
}

}  // namespace cf
