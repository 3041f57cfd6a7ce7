// cf_pick.h -- the template kernels are instantiated in their own translation units (cf_pick_*.cu), compiled in
// parallel; the API unit obtains them as host function pointers (each unit registers its own device code).
#pragma once

namespace cf {

struct KArgs; struct LArgs; struct DArgs; struct MArgs;
using KernelFn = void (*)(const KArgs);
using LKernel = void (*)(const LArgs);
using DKernel = void (*)(const DArgs);
using MKernel = void (*)(const MArgs);

// cf_pick_path.cu: generic path kernel (Black-Scholes / Dupire x European, UOC, Europeans); nullptr: not implemented
KernelFn pick_path_kernel(int mdl, int prd, bool aad, int rng);
// cf_pick_path.cu: itemised risk of Dupire x Europeans
MKernel pick_multi_kernel(int rng);
// cf_pick_dlm.cu: displaced multi-asset model, instantiated for up to 4 / 8 / 12 / 16 assets; nullptr: MultiStats with AAD
// warps: 8 (Sobol, or when the rows of 12 warps do not fit in shared memory) or 12 (mrg32k3a)
LKernel pick_dlm_kernel(int n_assets, int prd, bool aad, int rng, int warps);
inline int dlm_bucket(int n_assets) { return n_assets <= 4 ? 4 : n_assets <= 8 ? 8 : n_assets <= 12 ? 12 : 16; }
// cf_pick_dupire.cu: the north-star kernels; fwdP = paths per thread of the forward kernel, 1 or 2 (kFwdWarps warps per block),
// chunk = steps of Gaussians per fill (kFwdChunk; kFwdChunk1 is also built for fwdP = 1)
DKernel pick_dupire_forward(int prd, bool aad, int rng, int fwdP, int chunk);
DKernel pick_dupire_reverse(int prd);
int dupire_span_steps(int n_steps);             // steps per lane of the span reverse kernel for this timeline, 0: none
DKernel pick_dupire_reverse_span(int prd, int S);
// cf_pick_bs.cu: Black-Scholes x {European, UOC} fast path (fwdP = 1: 8-step chunks, 2: 4-step chunks)
DKernel pick_bs_forward(int prd, bool aad, int rng, int fwdP);
DKernel pick_bs_reverse(int prd, int S);   // one warp per live path, S steps per lane      // four lanes per live path (small shards)

}  // namespace cf
