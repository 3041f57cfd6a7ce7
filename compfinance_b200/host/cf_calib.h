// cf_calib.h -- the host stage of dupireSuperbucket (main.h:453-569): closed-form prices, an implied-volatility
// surface with a risk view on top, Dupire's formula and the calibration of the local-vol grid.  O(10^3) operations per
// maturity, once per run: it stays on the host and runs on the compact AD type of cf_aad.h when the risk view is
// differentiated.
//
// INTERFACE-MANDATED (main.h and the Excel wrappers are written against them): normalDens / normalCdf
// (gaussians.h:10-39), blackScholes / blackScholesIvol / merton (analytics.h:8-91), interp2D (interp.h:65-109),
// RiskView<T> (ivs.h:27-88), IVS with call() / localVol() and MertonIVS (ivs.h:90-179), dupireCalibMaturity / dupireCalib
// / DupireCalibResults (mcMdlDupire.h:289-388, main.h:413-447) -- and, because the calibrated local vols are compared
// BIT FOR BIT with the reference's, the floating-point order inside each formula (the polynomial of the normal
// distribution, d1 / d2, the bisection bracket and its final interpolation, the ten Poisson terms, the centred
// differences of Dupire's formula with their 0.5e04 / 1.0e08 factors).
// OWN STRUCTURE: the formulas live in namespace cfcal as small value types (a bisection bracket, a Poisson term
// generator, a five-point call stencil, a calibration window); the mandated names are thin fronts on them.
#pragma once

#include "cf_base.h"
#include "cf_util.h"

namespace cfcal {

constexpr double kSqrt2Pi = 2.506628274631;      // gaussians.h, as written there

// root of an increasing function by bisection, finished by one secant step between the last two brackets
struct Bisection
{
    double lo, hi, fLo, fHi;
    template <class F>
    double solve(const F& f, const double target, const double width)
    {
        while (hi - lo > width) {
            const double mid = 0.5 * (hi + lo);
            const double fMid = f(mid);
            if (fMid > target) { hi = mid; fHi = fMid; }
            else { lo = mid; fLo = fMid; }
        }
        return lo + (target - fLo) / (fHi - fLo) * (hi - lo);
    }
};

// the n-th term of a Poisson(intensity x maturity) mixture: its probability, produced one after the other
struct PoissonTerms
{
    double   expMinus, power = 1.0, rate;
    unsigned factorial = 1;
    explicit PoissonTerms(const double intensityTimesMaturity) : expMinus(std::exp(-intensityTimesMaturity)), rate(intensityTimesMaturity) {}
    double probability() const { return expMinus * power / factorial; }
    void next(const size_t n) { factorial *= unsigned(n + 1); power *= rate; }
};

// call prices around (strike, maturity): the point itself, +/- 1e-4 in maturity and +/- 1e-4 in strike
template <class T>
struct CallStencil
{
    T here, earlier, later, below, above;
    // Dupire: sigma_loc = sqrt(2 C_T / C_KK) / K
    T localVol(const double strike) const
    {
        const T dT = (later - earlier) * 0.5e04;
        const T dKK = (below + above - 2.0 * here) * 1.0e08;
        return sqrt(2.0 * dT / dKK) / strike;
    }
};

// indices [first, last] of the ascending spots within `halfWidth` of `centre` (last < first: none)
struct Window { int first, last; };
template <class IT>
inline Window windowAround(IT spots, const int n, const double centre, const double halfWidth)
{
    Window w{0, n - 1};
    while (w.first < n && spots[w.first] < centre - halfWidth) ++w.first;
    while (w.last >= 0 && spots[w.last] > centre + halfWidth) --w.last;
    return w;
}

}  // namespace cfcal

// ---- Gaussian functions (gaussians.h:10-39) with their AD overloads (AADExpr.h:420-446)
inline double normalDens(const double x)
{
    if (x < -10.0 || 10.0 < x) return 0.0;
    return std::exp(-0.5 * x * x) / cfcal::kSqrt2Pi;
}

// Zelen and Severo's approximation (gaussians.h:23-39)
inline double normalCdf(const double x)
{
    if (x < -10.0) return 0.0;
    if (x > 10.0) return 1.0;
    if (x < 0.0) return 1.0 - normalCdf(-x);
    static constexpr double p = 0.2316419, b[5] = {0.319381530, -0.356563782, 1.781477937, -1.821255978, 1.330274429};
    const double t = 1.0 / (1.0 + p * x);
    const double pol = t * (b[0] + t * (b[1] + t * (b[2] + t * (b[3] + t * b[4]))));
    return 1.0 - normalDens(x) * pol;
}
inline Number normalDens(const Number& x)
{
    const double d = normalDens(x.value());
    return Number::fromUnary(d, x, -x.value() * d);
}
inline Number normalCdf(const Number& x) { return Number::fromUnary(normalCdf(x.value()), x, normalDens(x.value())); }

// ---- Black-Scholes and Merton (analytics.h:8-91)
template <class T, class U, class V, class W>
inline T blackScholes(const U spot, const V strike, const T vol, const W mat)
{
    const auto stdev = vol * std::sqrt(mat);
    if (stdev <= EPS) return T(std::max(0.0, double(spot - strike)));
    const auto d2 = std::log(spot / strike) / stdev - 0.5 * stdev;
    const auto d1 = d2 + stdev;
    return spot * normalCdf(d1) - strike * normalCdf(d2);
}

// Implied vol: a bracket grown from [0.05, 0.5] by halving / doubling, bisected to 1e-12 (analytics.h:23-56)
inline double blackScholesIvol(const double spot, const double strike, const double prem, const double mat)
{
    if (prem <= std::max(0.0, spot - strike) + EPS) return 0.0;
    const auto price = [&](const double v) { return blackScholes(spot, strike, v, mat); };
    cfcal::Bisection b{0.05, 0.5, 0.0, 0.0};
    while (price(b.hi) < prem) b.hi *= 2;
    while (price(b.lo) > prem) b.lo /= 2;
    b.fHi = price(b.hi);
    b.fLo = price(b.lo);
    return b.solve(price, prem, 1.e-12);
}

// Merton's jump-diffusion call: ten terms of the Poisson mixture of Black-Scholes prices (analytics.h:59-91)
inline double merton(const double spot, const double strike, const double vol, const double mat, const double intens,
                     const double meanJmp, const double stdJmp)
{
    const double varJmp = stdJmp * stdJmp;
    const double mv2 = meanJmp + 0.5 * varJmp;
    const double comp = intens * (std::exp(mv2) - 1);
    const double var = vol * vol;
    cfcal::PoissonTerms jumps(intens * mat);
    double mixture = 0.0;
    for (size_t n = 0; n < 10; ++n) {
        const double fwdSpot = spot * std::exp(n * mv2 - comp * mat);
        const double condVol = std::sqrt(var + n * varJmp / mat);
        mixture += jumps.probability() * blackScholes(fwdSpot, strike, condVol, mat);
        jumps.next(n);
    }
    return mixture;
}

// ---- 2-D interpolation (interp.h:65-109): rows of z interpolated in y, then blended in x; flat outside
template <bool smoothStep = false, class T, class U, class V, class W, class X>
inline V interp2D(const std::vector<T>& x, const std::vector<U>& y, const matrix<V>& z, const W& x0, const X& y0)
{
    const auto alongRow = [&](const size_t row) { return interp<smoothStep>(y.begin(), y.end(), z[row], z[row] + y.size(), y0); };
    const cfnum::Bracket b = cfnum::bracket(x.begin(), x.end(), x0);
    if (b.side == cfnum::Bracket::Above) return alongRow(x.size() - 1);
    if (b.side == cfnum::Bracket::Below) return alongRow(0);
    auto zLeft = alongRow(b.lower);
    auto zRight = alongRow(b.lower + 1);
    auto t = (x0 - x[b.lower]) / (x[b.lower + 1] - x[b.lower]);
    return cfnum::blend<smoothStep>(zLeft, zRight, t);
}

// ---- Risk view: additive spreads to the implied vols on a (strike, maturity) grid (ivs.h:27-88)
template <class T>
class RiskView
{
    struct Grid { std::vector<double> strikes; std::vector<Time> mats; };
    Grid      grid;
    matrix<T> spreads;
    bool      none = true;

public:
    RiskView() = default;
    // every spread 0; for T = Number each is a leaf of the tape from here on, as in the reference
    RiskView(const std::vector<double>& strikes, const std::vector<Time>& mats)
        : grid{strikes, mats}, spreads(strikes.size(), mats.size()), none(false)
    {
        for (T& s : spreads) {
            s = T(0.0);
            if constexpr (std::is_same<T, Number>::value) s.putOnTape();
        }
    }
    T spread(const double strike, const Time mat) const
    {
        if (none) return T(0.0);
        return interp2D<true>(grid.strikes, grid.mats, spreads, strike, mat);
    }
    void bump(const size_t i, const size_t j, const double bumpBy) { spreads[i][j] += bumpBy; }

    bool   empty() const { return none; }
    size_t rows() const { return grid.strikes.size(); }
    size_t cols() const { return grid.mats.size(); }
    const std::vector<double>& strikes() const { return grid.strikes; }
    const std::vector<Time>&   mats() const { return grid.mats; }
    const matrix<T>&           risks() const { return spreads; }
    typename matrix<T>::iterator       begin() { return spreads.begin(); }
    typename matrix<T>::iterator       end() { return spreads.end(); }
    typename matrix<T>::const_iterator begin() const { return spreads.begin(); }
    typename matrix<T>::const_iterator end() const { return spreads.end(); }
};

// ---- Implied volatility surfaces (ivs.h:90-179)
class IVS
{
    double s0;

public:
    explicit IVS(const double spot) : s0(spot) {}
    virtual ~IVS() = default;
    double spot() const { return s0; }
    virtual double impliedVol(const double strike, const Time mat) const = 0;

    // call price off the surface, the risk view's spread added to the implied vol
    template <class T = double>
    T call(const double strike, const Time mat, const RiskView<T>* risk = nullptr) const
    {
        const T bump = risk ? risk->spread(strike, mat) : T(0.0);
        return blackScholes<T>(s0, strike, impliedVol(strike, mat) + bump, mat);
    }

    // Dupire's formula on centred differences of 1e-4 in maturity and strike (ivs.h:119-138)
    template <class T = double>
    T localVol(const double strike, const double mat, const RiskView<T>* risk = nullptr) const
    {
        cfcal::CallStencil<T> c;
        c.here = call(strike, mat, risk);
        c.earlier = call(strike, mat - 1.0e-04, risk);
        c.later = call(strike, mat + 1.0e-04, risk);
        c.below = call(strike - 1.0e-04, mat, risk);
        c.above = call(strike + 1.0e-04, mat, risk);
        return c.localVol(strike);
    }
};

class MertonIVS : public IVS
{
    struct Jumps { double intensity, mean, stdev; };
    double diffusionVol;
    Jumps  jumps;

public:
    MertonIVS(const double spot, const double vol, const double intens, const double aveJmp, const double stdJmp)
        : IVS(spot), diffusionVol(vol), jumps{intens, aveJmp, stdJmp} {}
    double impliedVol(const double strike, const Time mat) const override
    {
        const double premium = merton(spot(), strike, diffusionVol, mat, jumps.intensity, jumps.mean, jumps.stdev);
        return blackScholesIvol(spot(), strike, premium, mat);
    }
};

// ---- Calibration of the local-vol grid (mcMdlDupire.h:289-388)
// One maturity: Dupire's formula within 2.5 standard deviations of the spot (the at-the-money call x sqrt(2 pi) as the
// standard deviation), the edge values copied outwards -- for T = Number the copies share the edge's tape node.
template <class IT, class OT, class T = double>
inline void dupireCalibMaturity(const IVS& ivs, const Time maturity, IT spotsBegin, IT spotsEnd, OT lVolsBegin,
                                const RiskView<T>& riskView = RiskView<double>())
{
    const int n = int(std::distance(spotsBegin, spotsEnd));
    const double stdev = double(ivs.call(ivs.spot(), maturity)) * cfcal::kSqrt2Pi;
    const cfcal::Window w = cfcal::windowAround(spotsBegin, n, ivs.spot(), 2.5 * stdev);
    for (int i = w.first; i <= w.last; ++i) lVolsBegin[i] = ivs.localVol(spotsBegin[i], maturity, &riskView);
    for (int i = w.first - 1; i >= 0; --i) lVolsBegin[i] = lVolsBegin[w.first];
    for (int i = w.last + 1; i < n; ++i) lVolsBegin[i] = lVolsBegin[w.last];
}

template <class T>
struct DupireCalibResults
{
    std::vector<double> spots;
    std::vector<Time>   times;
    matrix<T>           lVols;      // spot major
};

// The grid: the spots to include filled to at most maxDs apart (1 cent tolerance), the times to include plus maxDt
// itself filled to at most maxDt apart (one hour tolerance); one calibration per time, stored spot major.
template <class T = double>
inline DupireCalibResults<T> dupireCalib(const IVS& ivs, const std::vector<double>& inclSpots, const double maxDs,
                                         const std::vector<Time>& inclTimes, const double maxDt,
                                         const RiskView<T>& riskView = RiskView<double>())
{
    constexpr double oneCent = 0.01, oneHour = 0.000114469;
    DupireCalibResults<T> grid;
    grid.spots = fillData(inclSpots, maxDs, oneCent);
    grid.times = fillData(inclTimes, maxDt, oneHour, &maxDt, &maxDt + 1);
    const size_t nS = grid.spots.size(), nT = grid.times.size();
    grid.lVols.resize(nS, nT);
    std::vector<T> column(nS);
    for (size_t j = 0; j < nT; ++j) {
        dupireCalibMaturity(ivs, grid.times[j], grid.spots.begin(), grid.spots.end(), column.begin(), riskView);
        for (size_t i = 0; i < nS; ++i) grid.lVols[i][j] = column[i];
    }
    return grid;
}

// main.h:413-447: the same on a Merton surface built here
inline DupireCalibResults<double> dupireCalib(const std::vector<double>& inclSpots, const double maxDs,
                                              const std::vector<Time>& inclTimes, const double maxDt, const double spot,
                                              const double vol, const double jmpIntens = 0.0, const double jmpAverage = 0.0,
                                              const double jmpStd = 0.0)
{
    return dupireCalib(MertonIVS(spot, vol, jmpIntens, jmpAverage, jmpStd), inclSpots, maxDs, inclTimes, maxDt);
}
