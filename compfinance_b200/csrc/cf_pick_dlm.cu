// Instantiations of the displaced multi-asset kernel (see cf_pick.h).
#include "cf_dlm.cuh"
#include "cf_pick.h"

namespace cf {
namespace {
template <int AMAX, int PRD>
LKernel pick2(bool aad, int rng)
{
    if (aad) {
        if constexpr (PRD == CF_PRODUCT_MULTISTATS) return nullptr;      // value-only test instrument
        else return rng == CF_RNG_SOBOL ? dlm_kernel<AMAX, PRD, true, CF_RNG_SOBOL> : dlm_kernel<AMAX, PRD, true, CF_RNG_MRG32K3A>;
    }
    return rng == CF_RNG_SOBOL ? dlm_kernel<AMAX, PRD, false, CF_RNG_SOBOL> : dlm_kernel<AMAX, PRD, false, CF_RNG_MRG32K3A>;
}
template <int AMAX>
LKernel pick1(int prd, bool aad, int rng)
{
    if (prd == CF_PRODUCT_AUTOCALL) return pick2<AMAX, CF_PRODUCT_AUTOCALL>(aad, rng);
    if (prd == CF_PRODUCT_BASKETS) return pick2<AMAX, CF_PRODUCT_BASKETS>(aad, rng);
    return pick2<AMAX, CF_PRODUCT_MULTISTATS>(aad, rng);
}
}  // namespace

LKernel pick_dlm_kernel(int amax, int prd, bool aad, int rng)
{
    return amax <= 4 ? pick1<4>(prd, aad, rng) : pick1<16>(prd, aad, rng);
}
}  // namespace cf
