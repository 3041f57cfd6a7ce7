// Instantiations of the displaced multi-asset kernel (see cf_pick.h).  One translation unit per asset-count bucket
// (CF_DLM_AMAX = 4, 8, 12, 16, set by the build), compiled in parallel; the unit of bucket 4 also holds the dispatcher.
#include "cf_dlm.cuh"
#include "cf_pick.h"

#ifndef CF_DLM_AMAX
#error "CF_DLM_AMAX must be 4, 8, 12 or 16"
#endif

namespace cf {
namespace {
template <int AMAX, int PRD>
LKernel pick2(bool aad, int rng, int warps)
{
    constexpr int S = CF_RNG_SOBOL, M = CF_RNG_MRG32K3A;
    if (aad) {
        if constexpr (PRD == CF_PRODUCT_MULTISTATS) return nullptr;      // value-only test instrument
        else return rng == S ? dlm_kernel<AMAX, PRD, true, S, 8> : warps == 12 ? dlm_kernel<AMAX, PRD, true, M, 12> : dlm_kernel<AMAX, PRD, true, M, 8>;
    }
    return rng == S ? dlm_kernel<AMAX, PRD, false, S, 8> : warps == 12 ? dlm_kernel<AMAX, PRD, false, M, 12> : dlm_kernel<AMAX, PRD, false, M, 8>;
}
template <int AMAX>
LKernel pick1(int prd, bool aad, int rng, int warps)
{
    if (prd == CF_PRODUCT_AUTOCALL) return pick2<AMAX, CF_PRODUCT_AUTOCALL>(aad, rng, warps);
    if (prd == CF_PRODUCT_BASKETS) return pick2<AMAX, CF_PRODUCT_BASKETS>(aad, rng, warps);
    return pick2<AMAX, CF_PRODUCT_MULTISTATS>(aad, rng, warps);
}
}  // namespace

LKernel pick_dlm_kernel_4(int prd, bool aad, int rng, int warps);
LKernel pick_dlm_kernel_8(int prd, bool aad, int rng, int warps);
LKernel pick_dlm_kernel_12(int prd, bool aad, int rng, int warps);
LKernel pick_dlm_kernel_16(int prd, bool aad, int rng, int warps);

#if CF_DLM_AMAX == 4
LKernel pick_dlm_kernel_4(int prd, bool aad, int rng, int warps) { return pick1<4>(prd, aad, rng, warps); }
LKernel pick_dlm_kernel(int n_assets, int prd, bool aad, int rng, int warps)
{
    if (n_assets <= 4) return pick_dlm_kernel_4(prd, aad, rng, warps);
    if (n_assets <= 8) return pick_dlm_kernel_8(prd, aad, rng, warps);
    if (n_assets <= 12) return pick_dlm_kernel_12(prd, aad, rng, warps);
    return pick_dlm_kernel_16(prd, aad, rng, warps);
}
#elif CF_DLM_AMAX == 8
LKernel pick_dlm_kernel_8(int prd, bool aad, int rng, int warps) { return pick1<8>(prd, aad, rng, warps); }
#elif CF_DLM_AMAX == 12
LKernel pick_dlm_kernel_12(int prd, bool aad, int rng, int warps) { return pick1<12>(prd, aad, rng, warps); }
#else
LKernel pick_dlm_kernel_16(int prd, bool aad, int rng, int warps) { return pick1<16>(prd, aad, rng, warps); }
#endif
}  // namespace cf
