// Instantiations of the generic path kernel and of the itemised-risk kernel (see cf_pick.h).
#include "cf_kernels.cuh"
#include "cf_multi.cuh"
#include "cf_pick.h"

namespace cf {
namespace {
template <int MDL, int PRD>
KernelFn pick2(bool aad, int rng)
{
    if (aad) return rng == CF_RNG_SOBOL ? path_kernel<MDL, PRD, true, CF_RNG_SOBOL> : path_kernel<MDL, PRD, true, CF_RNG_MRG32K3A>;
    return rng == CF_RNG_SOBOL ? path_kernel<MDL, PRD, false, CF_RNG_SOBOL> : path_kernel<MDL, PRD, false, CF_RNG_MRG32K3A>;
}
}  // namespace

KernelFn pick_path_kernel(int mdl, int prd, bool aad, int rng)
{
    if (mdl == CF_MODEL_BS && prd == CF_PRODUCT_EUROPEAN) return pick2<CF_MODEL_BS, CF_PRODUCT_EUROPEAN>(aad, rng);
    if (mdl == CF_MODEL_BS && prd == CF_PRODUCT_UOC) return pick2<CF_MODEL_BS, CF_PRODUCT_UOC>(aad, rng);
    if (mdl == CF_MODEL_DUPIRE && prd == CF_PRODUCT_EUROPEAN) return pick2<CF_MODEL_DUPIRE, CF_PRODUCT_EUROPEAN>(aad, rng);
    if (mdl == CF_MODEL_DUPIRE && prd == CF_PRODUCT_UOC) return pick2<CF_MODEL_DUPIRE, CF_PRODUCT_UOC>(aad, rng);
    if (mdl == CF_MODEL_BS && prd == CF_PRODUCT_EUROPEANS) return pick2<CF_MODEL_BS, CF_PRODUCT_EUROPEANS>(aad, rng);
    if (mdl == CF_MODEL_DUPIRE && prd == CF_PRODUCT_EUROPEANS) return pick2<CF_MODEL_DUPIRE, CF_PRODUCT_EUROPEANS>(aad, rng);
    if (mdl == CF_MODEL_BS && prd == CF_PRODUCT_CONTINGENT) return pick2<CF_MODEL_BS, CF_PRODUCT_CONTINGENT>(aad, rng);
    return nullptr;
}

MKernel pick_multi_kernel(int rng)
{
    return rng == CF_RNG_SOBOL ? dupire_europeans_multi_kernel<CF_RNG_SOBOL> : dupire_europeans_multi_kernel<CF_RNG_MRG32K3A>;
}
}  // namespace cf
