// cf_device.cuh -- device building blocks shared by the RNG kernels and the path kernels.
// fp64 throughout, no fast-math.  Integer RNG streams are bit-exact with the reference.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace cf {

constexpr int kBlock = 256;          // threads per block = paths per batch
constexpr int kWarps = kBlock / 32;
constexpr int kLowBits = 8;          // log2(kBlock): Sobol bits resolved per thread
constexpr unsigned kFull = 0xffffffffu;

// ---------------------------------------------------------------------------------------------
// invNormalCdf: Beasley-Springer-Moro, restated from gaussians.h:47-87 (same constants, same
// evaluation order).  The reference evaluates x * num / den with a true division.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double inv_normal_cdf(const double p)
{
    const bool sup = p > 0.5;
    const double up = sup ? 1.0 - p : p;
    const double x = up - 0.5;
    double r;
    if (fabs(x) < 0.42) {
        r = x * x;
        const double num = ((-25.44106049637 * r + 41.39119773534) * r + -18.61500062529) * r + 2.50662823884;
        const double den = (((3.13082909833 * r + -21.06224101826) * r + 23.08336743743) * r + -8.47351093090) * r + 1.0;
        r = x * num / den;
        return sup ? -r : r;
    }
    r = log(-log(up));
    r = 0.3374754822726147 + r * (0.9761690190917186 + r * (0.1607979714918209 + r * (0.0276438810333863
        + r * (0.0038405729373609 + r * (0.0003951896511919 + r * (0.0000321767881768
        + r * (0.0000002888167364 + r * 0.0000003960315187)))))));
    return sup ? r : -r;
}

// ---------------------------------------------------------------------------------------------
// Sobol (sobol.h:77-151).  Path p (0-based, absolute) is the sequence point of index n = p + 1
// (the first next() yields index 1); its integer state in dimension d is the XOR of the direction
// numbers jkDir[b][d] over the set bits b of Gray(n) = n ^ (n >> 1).  skipTo() (sobol.h:119) is
// the same closed form, so no sequential state is needed: every (path, dim) is random access.
//
// A block covers 256 consecutive indices.  With n = H * 256 + l:
//   Gray(n) >> 8          = Gray(H)                      -> block-uniform part, "base"
//   Gray(n) & 0xff        = (Gray(l) & 0xff) ^ ((H & 1) << 7)   -> 8 per-thread bits
// A 256-wide window that is not 256-aligned spans two consecutive H; Gray(H+1) differs from
// Gray(H) in one bit, so the second base is one more XOR.
// ---------------------------------------------------------------------------------------------
#define CF_ONEOVER2POW32 2.3283064365387E-10   /* sobol.h:26 -- NOT 2^-32 */

struct SobolThread {
    uint32_t mask[kLowBits];   // all-ones where the thread's low Gray bit k is set
    int      sel;              // 0: base of H0, 1: base of H0 + 1

    __device__ __forceinline__ void init(uint32_t nidx, uint32_t H0)
    {
        const uint32_t H = nidx >> kLowBits;
        const uint32_t l = nidx & (kBlock - 1);
        const uint32_t low = ((l ^ (l >> 1)) & (kBlock - 1)) ^ ((H & 1u) << (kLowBits - 1));
        sel = int((H - H0) & 1u);      // 0 or 1 for every real path; padding lanes past index 2^32 - 1 wrap around and stay in range
#pragma unroll
        for (int k = 0; k < kLowBits; ++k) mask[k] = 0u - ((low >> k) & 1u);
    }

    // dirlow: smem [dim][8]; base: smem [2][dim]
    __device__ __forceinline__ uint32_t state(const uint32_t* __restrict__ dirlow,
                                              const uint32_t* __restrict__ base, int dim, int d) const
    {
        const uint4 a = reinterpret_cast<const uint4*>(dirlow)[2 * d];
        const uint4 b = reinterpret_cast<const uint4*>(dirlow)[2 * d + 1];
        uint32_t x = base[sel * dim + d];
        x ^= (a.x & mask[0]) ^ (a.y & mask[1]);
        x ^= (a.z & mask[2]) ^ (a.w & mask[3]);
        x ^= (b.x & mask[4]) ^ (b.y & mask[5]);
        x ^= (b.z & mask[6]) ^ (b.w & mask[7]);
        return x;
    }
};

// Block-cooperative: fill base[2][dim] for the block whose first index is n0 (H0 = n0 >> 8).
// dir: global [32][dim] direction numbers.
__device__ __forceinline__ void sobol_block_base(uint32_t* base, const uint32_t* __restrict__ dir,
                                                 int dim, uint32_t H0)
{
    const uint32_t g0 = H0 ^ (H0 >> 1);
    const int flip = __ffs(~H0) - 1;          // Gray(H0+1) ^ Gray(H0) = 1 << ctz(~H0)
    for (int d = threadIdx.x; d < dim; d += blockDim.x) {
        uint32_t x = 0;
        uint32_t g = g0;
        while (g) {
            const int j = __ffs(g) - 1;
            g &= g - 1;
            x ^= __ldg(dir + (kLowBits + j) * dim + d);
        }
        base[d] = x;
        base[dim + d] = (flip >= 0 && flip + kLowBits < 32) ? (x ^ __ldg(dir + (kLowBits + flip) * dim + d)) : x;
    }
}

// Block-cooperative: dirlow[dim][8] <- dir[0..7][dim]
__device__ __forceinline__ void sobol_load_low(uint32_t* dirlow, const uint32_t* __restrict__ dir, int dim)
{
    for (int i = threadIdx.x; i < dim * kLowBits; i += blockDim.x) {
        const int d = i / kLowBits, k = i % kLowBits;
        dirlow[i] = __ldg(dir + k * dim + d);
    }
}

// ---------------------------------------------------------------------------------------------
// mrg32k3a (mrg32k3a.h:55-81).  The reference carries the state in doubles holding exact
// integers; the recurrences are exact modular arithmetic, reproduced here in uint64.
//   x_n = (1403580 x_{n-2} - 810728 x_{n-3}) mod m1      m1 = 2^32 - 209
//   y_n = (527612 y_{n-1} - 1370589 y_{n-3}) mod m2      m2 = 2^32 - 22853
//   u   = ((x - y) or (x - y + m1)) / (m1 + 1)           (true IEEE division)
// Antithetic pairing (mrg32k3a.h:107-186): path 2q draws D fresh numbers, path 2q+1 reuses them
// as 1-u / -g.  skipTo (mrg32k3a.h:192-238): path pair q starts at stream offset q*D, reached
// here with host-precomputed jump matrices A^(D 2^k) mod m (k = 0..31).
// ---------------------------------------------------------------------------------------------
constexpr uint64_t kM1 = 4294967087ull, kM2 = 4294944443ull;

__device__ __forceinline__ uint64_t mod_m1(uint64_t x)
{   // x < 2^64.  2^32 = 209 (mod m1)
    x = (x >> 32) * 209ull + (x & 0xffffffffull);      // < 2^41
    x = (x >> 32) * 209ull + (x & 0xffffffffull);      // < 2^32 + 209*2^9
    return x >= kM1 ? x - kM1 : x;
}
__device__ __forceinline__ uint64_t mod_m2(uint64_t x)
{   // 2^32 = 22853 (mod m2)
    x = (x >> 32) * 22853ull + (x & 0xffffffffull);    // < 2^47
    x = (x >> 32) * 22853ull + (x & 0xffffffffull);    // < 2^32 + 22853*2^15 < 2^33
    x = x >= kM2 ? x - kM2 : x;
    return x >= kM2 ? x - kM2 : x;
}

struct MrgThread {
    uint32_t x0, x1, x2, y0, y1, y2;    // (Xn, Xn1, Xn2), (Yn, Yn1, Yn2)

    // jump: global [32][2][9] uint64 (component 0: mod m1, 1: mod m2), row-major 3x3
    __device__ __forceinline__ void init(uint32_t seedA, uint32_t seedB, uint64_t pairIndex,
                                         const uint64_t* __restrict__ jump)
    {
        x0 = x1 = x2 = seedA;
        y0 = y1 = y2 = seedB;
        advance(pairIndex, jump);
    }

    // moves the state `count` path pairs (count * D numbers) down the stream: one jump matrix per set bit of count
    __device__ __forceinline__ void advance(uint64_t count, const uint64_t* __restrict__ jump)
    {
        uint64_t X[3] = {x0, x1, x2}, Y[3] = {y0, y1, y2};
        for (int k = 0; k < 64 && (count >> k); ++k) {
            if (!((count >> k) & 1ull)) continue;
            const uint64_t* M1 = jump + (k * 2 + 0) * 9;
            const uint64_t* M2 = jump + (k * 2 + 1) * 9;
            uint64_t nx[3], ny[3];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                nx[r] = mod_m1(mod_m1(__ldg(M1 + 3 * r) * X[0]) + mod_m1(__ldg(M1 + 3 * r + 1) * X[1])
                               + mod_m1(__ldg(M1 + 3 * r + 2) * X[2]));
                ny[r] = mod_m2(mod_m2(__ldg(M2 + 3 * r) * Y[0]) + mod_m2(__ldg(M2 + 3 * r + 1) * Y[1])
                               + mod_m2(__ldg(M2 + 3 * r + 2) * Y[2]));
            }
#pragma unroll
            for (int r = 0; r < 3; ++r) { X[r] = nx[r]; Y[r] = ny[r]; }
        }
        x0 = uint32_t(X[0]); x1 = uint32_t(X[1]); x2 = uint32_t(X[2]);
        y0 = uint32_t(Y[0]); y1 = uint32_t(Y[1]); y2 = uint32_t(Y[2]);
    }

    // next integer numerator z in [0, m1): u = z / (m1 + 1).  One reduction per component: the negative term is taken
    // as a13 (m - x), so p = a12 x1 + a13 (m1 - x2) < 2^54 is non-negative and congruent; 2^32 = 209 (mod m1) and
    // 22853 (mod m2) fold the high word twice, then one conditional subtraction.
    __device__ __forceinline__ uint32_t next()
    {
        const uint64_t px = 1403580ull * x1 + 810728ull * uint64_t(uint32_t(kM1) - x2);
        uint64_t tx = uint64_t(uint32_t(px)) + uint64_t(uint32_t(px >> 32) * 209u);            // high word < 2^22: < 2^33
        tx = uint64_t(uint32_t(tx)) + uint64_t(uint32_t(tx >> 32) * 209u);                    // < 2^32 + 209
        const uint32_t x = uint32_t(tx >= kM1 ? tx - kM1 : tx);
        x2 = x1; x1 = x0; x0 = x;
        const uint64_t py = 527612ull * y0 + 1370589ull * uint64_t(uint32_t(kM2) - y2);
        uint64_t ty = uint64_t(uint32_t(py)) + uint64_t(uint32_t(py >> 32)) * 22853ull;         // < 2^38
        ty = uint64_t(uint32_t(ty)) + uint64_t(uint32_t(ty >> 32) * 22853u);                  // < 2^32 + 2^21
        const uint32_t y = uint32_t(ty >= kM2 ? ty - kM2 : ty);
        y2 = y1; y1 = y0; y0 = y;
        return x > y ? x - y : uint32_t(uint64_t(x) + kM1 - y);
    }
};

__device__ __forceinline__ double div_fast(double a, double b);
// u = z / (m1 + 1) (mrg32k3a.h:75-80); the lean quotient is bit for bit the IEEE one for every numerator z < 2^32
// (checked exhaustively on the device: cf_selftest_mrg_uniform, tests/test_gpu_rng.py)
__device__ __forceinline__ double mrg_uniform(uint32_t z) { return div_fast(double(z), 4294967088.0); }
__device__ __forceinline__ double mrg_uniform_ieee(uint32_t z) { return double(z) / 4294967088.0; }

// a / b for b > 0 where a is often exactly zero (adjoints of inactive branches, dead notionals, options out of the
// money).  CUDA's double division leaves its inline path for a zero numerator and runs the out-of-line routine
// (~200 instructions); 0 / b is a itself, signed zero included, so the result is bit for bit the quotient.
__device__ __forceinline__ double div_z(double a, double b) { return a == 0.0 ? a : a / b; }
// the same for a numeraire-like divisor that is often exactly 1 (models without rates leave the Sample default): a / 1 = a
__device__ __forceinline__ double div_n(double a, double num) { return (a == 0.0 || num == 1.0) ? a : a / num; }

// ---------------------------------------------------------------------------------------------
// Lean arithmetic shared by the path kernels
// ---------------------------------------------------------------------------------------------
// a / b for normal operands well inside the exponent range: reciprocal seed (relative error < 2^-19.9, measured), one
// cubic Newton step (r (1 + e + e^2): error ~ e^3 = 2^-60), one residual correction of the quotient.  CUDA's own
// division adds a second Newton step and a range check; on 1.5e9 random operands of the ranges met here this
// sequence returned the IEEE quotient every time (tools/micro/divtest.cu).
__device__ __forceinline__ double div_fast(double a, double b)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
    double e = fma(-b, r, 1.0);
    e = fma(e, e, e);
    r = fma(r, e, r);
    const double q = a * r;
    return fma(fma(-b, q, a), r, q);
}

// exp(x) for |x| < 700 (no overflow / underflow / NaN handling): x = k ln 2 + r, |r| <= ln 2 / 2, Taylor to r^13 (2e-18),
// the exponent added to the high word.  Error < 1 ulp like the library exp; used on the few samples inside the barrier's
// log-space pre-filter, where the library routine's range handling is most of the cost.
__device__ __forceinline__ double exp_core(double x)
{
    const double kd = fma(x, 1.4426950408889634e+00, 6755399441055744.0);      // round to nearest: k in the low word
    const int k = __double2loint(kd);
    const double kf = kd - 6755399441055744.0;
    double r = fma(kf, -6.93147180559945286227e-01, x);
    r = fma(kf, -2.31904681384629955842e-17, r);
    double p = 1.60590438368216145994e-10;                                      // 1 / 13!
    p = fma(p, r, 2.08767569878680989792e-09);
    p = fma(p, r, 2.50521083854417187751e-08);
    p = fma(p, r, 2.75573192239858906526e-07);
    p = fma(p, r, 2.75573192239858906526e-06);
    p = fma(p, r, 2.48015873015873015873e-05);
    p = fma(p, r, 1.98412698412698412698e-04);
    p = fma(p, r, 1.38888888888888888889e-03);
    p = fma(p, r, 8.33333333333333333333e-03);
    p = fma(p, r, 4.16666666666666666667e-02);
    p = fma(p, r, 1.66666666666666666667e-01);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
}

// ---------------------------------------------------------------------------------------------
// Deterministic reductions
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

// Sum over the block in a fixed order; result valid in thread 0.  scratch: >= kWarps doubles.
__device__ __forceinline__ double block_sum(double v, double* scratch)
{
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
    if (threadIdx.x == 0) {
        for (int w = 0; w < int(blockDim.x >> 5); ++w) s += scratch[w];
    }
    return s;
}

// max(d, 0) for a finite d in three integer operations on the two halves (fmax carries NaN handling: nine)
__device__ __forceinline__ double pos_part(double d)
{
    const int hi = __double2hiint(d);
    const int keep = ~(hi >> 31);
    return __hiloint2double(hi & keep, __double2loint(d) & keep);
}

// Payoff sums of a call ladder over the 32 paths of a warp: sum over paths p of max(F_p - K_k, 0) / num for the strikes
// k < n.  The forwards go through 32 doubles of shared memory once and lane l adds the paths, in path order, for the
// strikes l, l + 32, ...: one pass instead of a five-level shuffle tree per strike.  All lanes must call together; a lane
// without a path contributes nothing.  myPay: the warp's row of payoff sums at the ladder's first strike.
__device__ __forceinline__ void warp_ladder_sums(double* fw, double F, bool valid, const double* __restrict__ K, int n,
                                                 double num, double* myPay, int lane)
{
    __syncwarp();
    fw[lane] = valid ? F : -1.0e300;                 // max(-1e300 - K, 0) = 0
    __syncwarp();
    const double2* f2 = reinterpret_cast<const double2*>(fw);
    // one strike per pass, not unrolled further: the body is 16 pairs already and code size matters (instruction cache)
#pragma unroll 1
    for (int k = lane; k < n; k += 32) {
        const double strike = K[k];
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int p = 0; p < 16; ++p) {
            const double2 f = f2[p];                 // every lane reads the same address: one broadcast
            s0 += pos_part(f.x - strike);
            s1 += pos_part(f.y - strike);
        }
        const double s = s0 + s1;
        myPay[k] += (num == 1.0) ? s : s / num;
    }
}

// Keyed warp accumulation (the scatter of knot adjoints): every lane adds (a, b) to row[key].
// Lanes sharing a key are serialised in lane order, so the result is bit-reproducible.
// row: this warp's private smem row of nbins double2, zeroed here.
__device__ __forceinline__ void warp_keyed_accumulate(double2* row, int nbins, int key, double a, double b)
{
    const int lane = threadIdx.x & 31;
    for (int j = lane; j < nbins; j += 32) row[j] = make_double2(0.0, 0.0);
    __syncwarp();
    const unsigned peers = __match_any_sync(kFull, key);
    const int rank = __popc(peers & ((1u << lane) - 1u));
    const int maxrank = __reduce_max_sync(kFull, rank);
    for (int r = 0; r <= maxrank; ++r) {
        if (rank == r) {
            double2 v = row[key];
            v.x += a; v.y += b;
            row[key] = v;
        }
        __syncwarp();
    }
}

}  // namespace cf
