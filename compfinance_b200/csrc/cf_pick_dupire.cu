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
    if (fwdP == 4) return uoc ? pickF<CF_PRODUCT_UOC, 4, 16>(aad, rng) : pickF<CF_PRODUCT_EUROPEAN, 4, 16>(aad, rng);
    return uoc ? pickF<CF_PRODUCT_UOC, 2, 24>(aad, rng) : pickF<CF_PRODUCT_EUROPEAN, 2, 24>(aad, rng);
}

DKernel pick_dupire_reverse(int prd, int P)
{
    if (prd == CF_PRODUCT_UOC) return P == 4 ? dupire_reverse_kernel<CF_PRODUCT_UOC, 4> : dupire_reverse_kernel<CF_PRODUCT_UOC, 2>;
    return P == 4 ? dupire_reverse_kernel<CF_PRODUCT_EUROPEAN, 4> : dupire_reverse_kernel<CF_PRODUCT_EUROPEAN, 2>;
}
}  // namespace cf
