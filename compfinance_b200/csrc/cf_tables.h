// cf_tables.h -- host-side integer tables of the RNGs: Sobol direction numbers regenerated from
// the Joe-Kuo initialisers, and mrg32k3a jump matrices.
#pragma once

#include <cstdint>
#include <vector>

namespace cf {

// jkDir[bit][dim] for bit 0..31, dim 0..maxDim-1 (sobol.cpp:16-3672 holds the same numbers as a
// literal table; here they are rebuilt with the published recurrence from data/joe_kuo_old_1111.txt).
const std::vector<uint32_t>& sobol_direction_table();   // [32][sobol_max_dim()]
int sobol_max_dim();

// Jump matrices of mrg32k3a for strides stride * 2^k (k = 0..31): out[k][c][9], c = 0 (mod m1), 1 (mod m2).
// The transition matrices are those of mrg32k3a.h:316-350 (skipNumbers), state vector (Xn, Xn1, Xn2).
std::vector<uint64_t> mrg_jump_matrices(uint64_t stride);

}  // namespace cf
