// Instantiations of the Black-Scholes fast path (see cf_pick.h): the shared forward kernel with the log-normal step,
// the span reverse kernel of cf_bs.cuh.
#include "cf_bs.cuh"
#include "cf_pick.h"

namespace cf {
namespace {
template <int PRD, int P, int CH>
DKernel pickF(bool aad, int rng)
{
    if (aad) return rng == CF_RNG_SOBOL ? dupire_forward4_kernel<CF_MODEL_BS, PRD, true, CF_RNG_SOBOL, P, kFwdWarps, CH>
                                        : dupire_forward4_kernel<CF_MODEL_BS, PRD, true, CF_RNG_MRG32K3A, P, kFwdWarps, CH>;
    return rng == CF_RNG_SOBOL ? dupire_forward4_kernel<CF_MODEL_BS, PRD, false, CF_RNG_SOBOL, P, kFwdWarps, CH>
                               : dupire_forward4_kernel<CF_MODEL_BS, PRD, false, CF_RNG_MRG32K3A, P, kFwdWarps, CH>;
}
template <int S>
DKernel pickR(int prd) { return prd == CF_PRODUCT_UOC ? bs_reverse_span_kernel<CF_PRODUCT_UOC, S> : bs_reverse_span_kernel<CF_PRODUCT_EUROPEAN, S>; }
}  // namespace

DKernel pick_bs_forward(int prd, bool aad, int rng, int fwdP)
{
    const bool uoc = prd == CF_PRODUCT_UOC;
    if (fwdP == 1) return uoc ? pickF<CF_PRODUCT_UOC, 1, kFwdChunk1>(aad, rng) : pickF<CF_PRODUCT_EUROPEAN, 1, kFwdChunk1>(aad, rng);
    return uoc ? pickF<CF_PRODUCT_UOC, 2, kFwdChunk>(aad, rng) : pickF<CF_PRODUCT_EUROPEAN, 2, kFwdChunk>(aad, rng);
}

DKernel pick_bs_reverse(int prd, int S)
{
    switch (S) {
        case 1: return pickR<1>(prd);
        case 2: return pickR<2>(prd);
        case 3: return pickR<3>(prd);
        case 4: return pickR<4>(prd);
        case 5: return pickR<5>(prd);
        case 6: return pickR<6>(prd);
        case 8: return pickR<8>(prd);
        default: return nullptr;
    }
}
}  // namespace cf
