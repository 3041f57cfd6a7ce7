// cf_bs.cuh -- Black-Scholes x {European, UOC}: the reverse sweep of the fast path, and the reduction of a fast-path run.
//
// Replaces, with cf::dupire_forward4_kernel<CF_MODEL_BS, ...> (cf_dupire.cuh), BlackScholes::generatePath
// (mcMdlBS.h:321-350) + European / UOC::payoffs (mcPrd.h:113-125, 235-288) under the loops of mcBase.h:378-386 /
// 680-704 and the per-path tape sweep.  Adjoint equations: SURVEY.md Appendix A.2, taken in LOG space:
//
//   X_{i+1} = X_i + drift_i + std_i g_i,  S = exp(X),  forward of a sample F = S ff
//   Lbar_i  = Lbar_{i+1} + b_i                  (b_i: adjoint of log F of the barrier sample at point i + 1; 0 outside
//                                                the smoothing zone; the payoff seeds Lbar at maturity)
//   driftbar_i += Lbar_i,  stdbar_i += Lbar_i g_i,  ffbar_e += b_e / ff_e,  spotbar += Lbar_0 / S_0
//
// The recursion has no multiplier, so a path is swept by ONE WARP AT ONCE like the span form of the Dupire kernel
// (lane l owns the S consecutive steps [S l, S l + S)): g_i is recovered from consecutive log-spots of the forward
// kernel's history, Lbar is the payoff seed plus a suffix sum of the (rare) barrier terms -- the warp scan runs only
// when a lane has one -- and, because a lane always works on the same steps, the per-step table adjoints accumulate in
// the lane's REGISTERS over all the paths of the warp: no shared-memory accumulators, no shuffles, no barriers in the
// loop.  Only live paths (non-zero payoff adjoints) are swept, as the reference's own sweep skips zero adjoints
// (AADNode.h:76).  Warp rows are combined per block in warp order, blocks by rows_reduce_kernel: bit-reproducible.
#pragma once

#include "cf_dupire.cuh"

namespace cf {

#ifndef CF_BS_REV_WARPS
#define CF_BS_REV_WARPS 24                   // measured on config 2: 16 warps (128 registers) 0.363 ms, 20: 0.345, 24 (80 registers): 0.336, 32 (64): 0.365
#endif
constexpr int kBsRevWarps = CF_BS_REV_WARPS;
constexpr int kBsRevBlock = kBsRevWarps * 32;
constexpr int kBsMaxWords = 1024;              // live-mask words (32 paths each) one block can own

// table-adjoint vector of Black-Scholes (include/cf_b200.h): spot | drifts [D] | stds [D] | numeraire, fwd factor, discount, libor [E] each
__host__ __device__ inline int bs_adj_size(int D, int E) { return 1 + 2 * D + 4 * E; }

struct BsSmemR { size_t ds, bits, live, rows, red, total; };
__host__ __device__ inline BsSmemR bs_smem_rev(int D, int E)
{
    BsSmemR s{};
    s.ds = align16(sizeof(double2) * size_t(D + 1));
    s.bits = align16(sizeof(uint32_t) * ((D + 31) / 32 + 1));
    s.live = align16(sizeof(uint32_t) * (2 * kBsMaxWords + 4 + kBsRevWarps));      // masks, prefix (+ 1), warp totals on their own 16 bytes
    s.rows = align16(sizeof(double) * size_t(bs_adj_size(D, E))) * kBsRevWarps;
    s.red = align16(sizeof(double) * kBsRevWarps);
    s.total = s.ds + s.bits + s.live + s.rows + s.red;
    return s;
}

// a.partial_rev: [grid][bs_adj_size] table adjoints of the block's live paths
template <int PRD, int S>
__global__ void __launch_bounds__(kBsRevBlock, 1) bs_reverse_span_kernel(const DArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t lane = uint32_t(tid & 31);
    const int D = a.n_steps, E = a.n_events;
    const int nAdj = bs_adj_size(D, E);

    const BsSmemR z = bs_smem_rev(D, E);
    unsigned char* p = smem_raw;
    double2* dsS = reinterpret_cast<double2*>(p);        p += z.ds;
    uint32_t* bitS = reinterpret_cast<uint32_t*>(p);     p += z.bits;
    uint32_t* maskS = reinterpret_cast<uint32_t*>(p);
    uint32_t* prefS = maskS + kBsMaxWords;
    uint32_t* wtotS = prefS + kBsMaxWords + 4;         p += z.live;      // 16-byte aligned: a widened load of the totals never touches the prefix array
    double* rowsS = reinterpret_cast<double*>(p);        p += z.rows;
    double* red = reinterpret_cast<double*>(p);
    const size_t rowStride = align16(sizeof(double) * size_t(nAdj)) / sizeof(double);
    double* myRow = rowsS + rowStride * size_t(warp);

    for (int i = tid; i < D; i += kBsRevBlock) dsS[i] = a.bs_ds[i];
    for (int i = tid; i < (D + 31) / 32; i += kBsRevBlock) bitS[i] = a.ev_bits[i];
    for (int i = int(lane); i < nAdj; i += 32) myRow[i] = 0.0;
    pdl_wait();                                        // the forward kernel's history, states and live mask are complete
    pdl_launch_dependents();

    // ---- live paths of this block: a contiguous range of mask words, compacted in path order (deterministic).  The
    // sweep of a path is a few dozen instructions whatever the path: equal ranges of paths are balanced enough.
    const uint32_t nWall = uint32_t(a.n_pad >> 5);
    const uint32_t wBeg = uint32_t(uint64_t(blockIdx.x) * nWall / gridDim.x), wEnd = uint32_t(uint64_t(blockIdx.x + 1) * nWall / gridDim.x);
    const uint32_t nW = wEnd - wBeg;                     // <= kBsMaxWords (host)
    const uint32_t wpt = (nW + kBsRevBlock - 1) / kBsRevBlock;
    {
        const uint32_t w0 = min(uint32_t(tid) * wpt, nW), w1 = min(w0 + wpt, nW);
        uint32_t mine = 0;
        for (uint32_t i = w0; i < w1; ++i) { const uint32_t mk = __ldcg(a.live + wBeg + i); maskS[i] = mk; mine += uint32_t(__popc(mk)); }
        uint32_t incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(kFull, incl, o); if (int(lane) >= o) incl += t; }
        if (lane == 31u) wtotS[warp] = incl;
        __syncthreads();
        uint32_t off = 0, total = 0;
        for (int w = 0; w < kBsRevWarps; ++w) { if (w < warp) off += wtotS[w]; total += wtotS[w]; }
        uint32_t run = off + incl - mine;
        for (uint32_t i = w0; i < w1; ++i) { prefS[i] = run; run += uint32_t(__popc(maskS[i])); }
        for (uint32_t i = nW + uint32_t(tid); i <= kBsMaxWords; i += kBsRevBlock) prefS[i] = total;
    }
    __syncthreads();
    const uint32_t nLive = prefS[kBsMaxWords];
    auto selectPath = [&](uint32_t q) -> uint32_t {
        uint32_t lo = 0u, hi = kBsMaxWords;
        while (hi - lo > 1u) {
            const uint32_t mid = (lo + hi) >> 1;
            if (prefS[mid] <= q) lo = mid; else hi = mid;
        }
        return (wBeg + lo) * 32u + __fns(maskS[lo], 0u, int(q - prefS[lo]) + 1);
    };

    const double strike = a.strike;
    const double twoSmooth = 2 * a.smooth, barSmooth = a.barrier + a.smooth, minusSmooth = a.barrier - a.smooth;
    const double logZone = (PRD == CF_PRODUCT_UOC) ? (minusSmooth > 0.0 ? log(minusSmooth) - 1.0e-9 : -DBL_MAX) : DBL_MAX;
    const bool isPut = a.is_put != 0;
    const double w0 = a.w[0], w1 = a.w[1];
    const double ff = a.fwd_factor, scale = a.pay_scale;
    constexpr uint32_t histStride = 1024;
    const size_t pathBlock = size_t(256) * size_t((D + 3) >> 2);

    // ---- this lane's steps i0 .. i0 + S - 1: drift, 1 / std, event flag of point i + 1, history element
    const int i0 = S * int(lane);
    double dr[S], isd[S];
    bool ev[S], in[S];
    uint32_t hoff[S];
#pragma unroll
    for (int j = 0; j < S; ++j) {
        const int i = i0 + j;
        in[j] = i < D;
        const uint32_t ii = uint32_t(in[j] ? i : D - 1);
        dr[j] = dsS[ii].x; isd[j] = 1.0 / dsS[ii].y;
        ev[j] = (PRD == CF_PRODUCT_UOC) && in[j] && ((bitS[ii >> 5] >> (ii & 31u)) & 1u);
        hoff[j] = (ii >> 2) * histStride + (ii & 3u);
    }
    // per-lane accumulators over all the paths of the warp
    double driftAcc[S], stdAcc[S], ffAcc[S];
#pragma unroll
    for (int j = 0; j < S; ++j) { driftAcc[j] = 0.0; stdAcc[j] = 0.0; ffAcc[j] = 0.0; }
    double spotAcc = 0.0, ffTodayAcc = 0.0, ffLastAcc = 0.0, numLastAcc = 0.0, discLastAcc = 0.0;
    const double X0 = log(a.spot);

    for (uint32_t qb = uint32_t(warp); qb < nLive; qb += 32u * kBsRevWarps) {
        // ---- set-up of the warp's next 32 paths, one per lane
        const uint32_t qMine = qb + kBsRevWarps * lane;
        const bool mine = qMine < nLive;
        const uint32_t pthMine = selectPath(mine ? qMine : nLive - 1u);
        const double XTm = __ldcg(a.state + pthMine);
        const double aenc = __ldcg(a.state + a.n_pad + pthMine);
        const bool killedM = aenc < 0.0;
        const double aliveM = killedM ? 0.0 : aenc;
        const double FTm = exp(XTm) * ff;
        const double euroM = (isPut ? fmax(strike - FTm, 0.0) : fmax(FTm - strike, 0.0)) * scale;
        const double eurobarM = !mine ? 0.0 : ((PRD == CF_PRODUCT_UOC) ? w0 * aliveM + w1 : w0);
        const double Km = (PRD == CF_PRODUCT_UOC && !killedM && mine) ? (w0 * euroM) * aliveM : 0.0;   // abar x alive: invariant
        const double xTm = isPut ? strike - FTm : FTm - strike;
        // adjoint of log F_T from the payoff, and from the barrier sample at maturity
        const double GpayM = (xTm > 0.0) ? (isPut ? -eurobarM : eurobarM) * scale * FTm : 0.0;
        double GbarM = 0.0;
        if (PRD == CF_PRODUCT_UOC && !killedM && XTm > logZone && FTm > minusSmooth) {
            const double f = div_fast(barSmooth - FTm, twoSmooth);
            GbarM = (f != 0.0) ? (Km / f) * (-1.0 / twoSmooth) * FTm : 0.0;
        }
        // table adjoints of the last event: numeraire, forward factor, discount (mcPrd.h:122-124, 276-287)
        numLastAcc -= eurobarM * euroM;                  // x 1 / numeraire at the end
        ffLastAcc += GpayM + GbarM;                      // x 1 / ff at the end
        if (PRD == CF_PRODUCT_EUROPEAN) discLastAcc += eurobarM * euroM;   // x 1 / discount at the end
        const double zoneM = killedM ? DBL_MAX : logZone;
        const uint32_t cnt = min(32u, (nLive - qb + kBsRevWarps - 1u) / kBsRevWarps);
        double hx[S];
        {
            const uint32_t pth = __shfl_sync(kFull, pthMine, 0);
            const double* hp = a.hist + 4 * (size_t(pth >> 8) * pathBlock + (pth & 255u));
#pragma unroll
            for (int j = 0; j < S; ++j) hx[j] = __ldcg(hp + hoff[j]);
        }
        for (uint32_t jj = 0; jj < cnt; ++jj) {
            const double XT = __shfl_sync(kFull, XTm, int(jj));
            const double GT = __shfl_sync(kFull, GpayM + GbarM, int(jj));
            const double K = __shfl_sync(kFull, Km, int(jj));
            const double zone = __shfl_sync(kFull, zoneM, int(jj));
            const double nextLane = __shfl_down_sync(kFull, hx[0], 1);
            double g[S], b[S];
            double bSum = 0.0;
#pragma unroll
            for (int j = 0; j < S; ++j) {
                const double L = hx[j];
                const double Lraw = (j + 1 < S) ? hx[(j + 1 < S) ? j + 1 : j] : nextLane;
                const double Ln = (i0 + j + 1 >= D) ? XT : Lraw;
                g[j] = (Ln - L - dr[j]) * isd[j];
                b[j] = 0.0;
                if (PRD == CF_PRODUCT_UOC && ev[j] && Ln > zone) {           // rare: a sample inside the smoothing zone
                    const double F = exp_core(Ln) * ff;
                    if (F > minusSmooth) {
                        const double f = div_fast(barSmooth - F, twoSmooth);
                        b[j] = (f != 0.0) ? (K / f) * (-1.0 / twoSmooth) * F : 0.0;
                    }
                }
                bSum += b[j];
            }
            if (jj + 1u < cnt) {
                const uint32_t pth = __shfl_sync(kFull, pthMine, int(jj + 1u));
                const double* hp = a.hist + 4 * (size_t(pth >> 8) * pathBlock + (pth & 255u));
#pragma unroll
                for (int j = 0; j < S; ++j) hx[j] = __ldcg(hp + hoff[j]);
            }
            // Lbar entering this lane's span = payoff seed + the barrier terms of the later lanes
            double later = 0.0;
            if (PRD == CF_PRODUCT_UOC && __any_sync(kFull, bSum != 0.0)) {
                double incl = bSum;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const double t = __shfl_down_sync(kFull, incl, o); if (int(lane) + o < 32) incl += t; }
                later = incl - bSum;
            }
            double Lbar = GT + later;
#pragma unroll
            for (int j = S - 1; j >= 0; --j) {
                Lbar += b[j];
                if (in[j]) { driftAcc[j] += Lbar; stdAcc[j] = fma(Lbar, g[j], stdAcc[j]); ffAcc[j] += b[j]; }
            }
            if (lane == 0u) {
                // today's sample (timeline point 0), then S_0 = spot
                double bToday = 0.0;
                if (PRD == CF_PRODUCT_UOC && a.ev0 && X0 > zone) {
                    const double F = exp_core(X0) * ff;
                    if (F > minusSmooth) {
                        const double f = div_fast(barSmooth - F, twoSmooth);
                        bToday = (f != 0.0) ? (K / f) * (-1.0 / twoSmooth) * F : 0.0;
                    }
                }
                ffTodayAcc += bToday;
                spotAcc += Lbar + bToday;
            }
        }
    }

    // ---- the warp's row in the layout of the adjoint vector, then the block's sum in warp order
    const int evShift = a.ev0 ? 1 : 0;                 // event index of timeline point i + 1 is i + evShift
#pragma unroll
    for (int j = 0; j < S; ++j) {
        const int i = i0 + j;
        if (i < D) {
            myRow[1 + i] = driftAcc[j];
            myRow[1 + D + i] = stdAcc[j];
            if (i + 1 < D) myRow[1 + 2 * D + E + (i + evShift)] = ffAcc[j] / ff;      // the last point is handled below
        }
    }
    __syncwarp();
    // the last event's adjoints were accumulated by the lane that set the path up: the warp's sums, fixed shuffle tree
    numLastAcc = warp_sum(numLastAcc); ffLastAcc = warp_sum(ffLastAcc); discLastAcc = warp_sum(discLastAcc);
    if (lane == 0u) {
        myRow[0] = spotAcc / a.spot;
        if (a.ev0) myRow[1 + 2 * D + E] = ffTodayAcc / ff;
        myRow[1 + 2 * D + (E - 1)] = numLastAcc / a.bs_num;
        myRow[1 + 2 * D + E + (E - 1)] = ffLastAcc / ff;
        if (PRD == CF_PRODUCT_EUROPEAN) myRow[1 + 2 * D + 2 * E + (E - 1)] = discLastAcc / a.bs_disc;
    }
    __syncthreads();
    double* out = a.partial_rev + size_t(blockIdx.x) * size_t(nAdj);
    for (int k = tid; k < nAdj; k += kBsRevBlock) {
        double t = 0.0;
        for (int w = 0; w < kBsRevWarps; ++w) t += rowsS[rowStride * size_t(w) + k];
        out[k] = (a.accumulate ? out[k] : 0.0) + t;
    }
    (void)red;
}

// out[k] = sum over blocks of headPartial[b][k] (k < nHead), then of tailPartial[b][k - nHead]; one warp per output,
// fixed order; with peers: the sum over the participants as well (cf_comm.cuh)
static __global__ void rows_reduce_kernel(const double* __restrict__ headPartial, int nBlocksH, int strideH, int nHead,
                                          const double* __restrict__ tailPartial, int nBlocksT, int strideT, int nTail,
                                          double* __restrict__ out, const DPeers peers)
{
    pdl_wait();
    const int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (k >= nHead + nTail) return;
    double s = 0.0;
    if (k < nHead) { for (int b = lane; b < nBlocksH; b += 32) s += headPartial[size_t(b) * strideH + k]; }
    else { for (int b = lane; b < nBlocksT; b += 32) s += tailPartial[size_t(b) * strideT + (k - nHead)]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(kFull, s, o);
    if (peers.world > 1) s = peer_warp_sum(peers, size_t(k), s, lane);
    if (lane == 0) out[k] = s;
}

}  // namespace cf
