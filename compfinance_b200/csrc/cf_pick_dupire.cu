// Instantiations of the Dupire forward / reverse kernels (see cf_pick.h).
#include "cf_dupire.cuh"
#include "cf_pick.h"

namespace cf {
namespace {
template <int PRD, int P, int NW, int CH>
DKernel pickF(bool aad, int rng)
{
    if (aad) return rng == CF_RNG_SOBOL ? dupire_forward4_kernel<PRD, true, CF_RNG_SOBOL, P, NW, CH> : dupire_forward4_kernel<PRD, true, CF_RNG_MRG32K3A, P, NW, CH>;
    return rng == CF_RNG_SOBOL ? dupire_forward4_kernel<PRD, false, CF_RNG_SOBOL, P, NW, CH> : dupire_forward4_kernel<PRD, false, CF_RNG_MRG32K3A, P, NW, CH>;
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

DKernel pick_dupire_reverse_quad(int prd)
{
    return prd == CF_PRODUCT_UOC ? dupire_reverse_quad_kernel<CF_PRODUCT_UOC> : dupire_reverse_quad_kernel<CF_PRODUCT_EUROPEAN>;
}
}  // namespace cf
