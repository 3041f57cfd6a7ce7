// Instantiations of the Dupire forward / reverse kernels (see cf_pick.h).
#include "cf_dupire.cuh"
#include "cf_pick.h"

namespace cf {
namespace {
template <int PRD, int P, int NW>
DKernel pickF(bool aad, int rng)
{
    if (aad) return rng == CF_RNG_SOBOL ? dupire_forward4_kernel<PRD, true, CF_RNG_SOBOL, P, NW> : dupire_forward4_kernel<PRD, true, CF_RNG_MRG32K3A, P, NW>;
    return rng == CF_RNG_SOBOL ? dupire_forward4_kernel<PRD, false, CF_RNG_SOBOL, P, NW> : dupire_forward4_kernel<PRD, false, CF_RNG_MRG32K3A, P, NW>;
}
}  // namespace

DKernel pick_dupire_forward(int prd, bool aad, int rng, int fwdP)
{
    const bool uoc = prd == CF_PRODUCT_UOC;
    if (fwdP == 1) return uoc ? pickF<CF_PRODUCT_UOC, 1, kFwdWarps>(aad, rng) : pickF<CF_PRODUCT_EUROPEAN, 1, kFwdWarps>(aad, rng);
    return uoc ? pickF<CF_PRODUCT_UOC, 2, kFwdWarps>(aad, rng) : pickF<CF_PRODUCT_EUROPEAN, 2, kFwdWarps>(aad, rng);
}

DKernel pick_dupire_reverse(int prd)
{
    return prd == CF_PRODUCT_UOC ? dupire_reverse_kernel<CF_PRODUCT_UOC> : dupire_reverse_kernel<CF_PRODUCT_EUROPEAN>;
}
}  // namespace cf
