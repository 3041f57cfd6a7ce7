// Instantiations of the Dupire forward / reverse kernels (see cf_pick.h).
#include "cf_dupire.cuh"
#include "cf_pick.h"

namespace cf {
namespace {
template <int PRD, int P, int NW, int CH>
DKernel pickF(bool aad, int rng)
{
    if (aad) return rng == CF_RNG_SOBOL ? dupire_forward4_kernel<CF_MODEL_DUPIRE, PRD, true, CF_RNG_SOBOL, P, NW, CH> : dupire_forward4_kernel<CF_MODEL_DUPIRE, PRD, true, CF_RNG_MRG32K3A, P, NW, CH>;
    return rng == CF_RNG_SOBOL ? dupire_forward4_kernel<CF_MODEL_DUPIRE, PRD, false, CF_RNG_SOBOL, P, NW, CH> : dupire_forward4_kernel<CF_MODEL_DUPIRE, PRD, false, CF_RNG_MRG32K3A, P, NW, CH>;
}
}  // namespace

DKernel pick_dupire_forward(int prd, bool aad, int rng, int fwdP, int chunk)
{
    const bool uoc = prd == CF_PRODUCT_UOC;
    if (fwdP == 1 && chunk == kFwdChunk1) return uoc ? pickF<CF_PRODUCT_UOC, 1, kFwdWarps, kFwdChunk1>(aad, rng) : pickF<CF_PRODUCT_EUROPEAN, 1, kFwdWarps, kFwdChunk1>(aad, rng);
    if (fwdP == 1) return uoc ? pickF<CF_PRODUCT_UOC, 1, kFwdWarps, kFwdChunk>(aad, rng) : pickF<CF_PRODUCT_EUROPEAN, 1, kFwdWarps, kFwdChunk>(aad, rng);
    return uoc ? pickF<CF_PRODUCT_UOC, 2, kFwdWarps, kFwdChunk>(aad, rng) : pickF<CF_PRODUCT_EUROPEAN, 2, kFwdWarps, kFwdChunk>(aad, rng);
}

DKernel pick_dupire_reverse(int prd)
{
    return prd == CF_PRODUCT_UOC ? dupire_reverse_kernel<CF_PRODUCT_UOC> : dupire_reverse_kernel<CF_PRODUCT_EUROPEAN>;
}

namespace {
template <int S>
DKernel pickS(int prd) { return prd == CF_PRODUCT_UOC ? dupire_reverse_span_kernel<CF_PRODUCT_UOC, S> : dupire_reverse_span_kernel<CF_PRODUCT_EUROPEAN, S>; }
}  // namespace

// steps per lane for which the span kernel is built: the smallest one >= ceil(n_steps / 32), 0 if there is none
int dupire_span_steps(int n_steps)
{
    const int need = (n_steps + 31) / 32;
    for (int s : {1, 2, 3, 4, 5, 6, 8}) if (s >= need) return s;
    return 0;
}

DKernel pick_dupire_reverse_span(int prd, int S)
{
    switch (S) {
        case 1: return pickS<1>(prd);
        case 2: return pickS<2>(prd);
        case 3: return pickS<3>(prd);
        case 4: return pickS<4>(prd);
        case 5: return pickS<5>(prd);
        case 6: return pickS<6>(prd);
        case 8: return pickS<8>(prd);
        default: return nullptr;
    }
}

}  // namespace cf
