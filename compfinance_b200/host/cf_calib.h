// cf_calib.h -- the host stage of dupireSuperbucket (main.h:453-569): Black-Scholes / Merton
// analytics (analytics.h:8-91, gaussians.h:10-39), implied volatility surfaces and risk views
// (ivs.h:27-179), Dupire's formula and the calibration of the local-vol grid
// (mcMdlDupire.h:289-388).  O(10^3) operations per maturity, once per run: this stays on the host
// and runs on the compact AD type of cf_aad.h when the risk view is differentiated.
#pragma once

#include "cf_base.h"
#include "cf_util.h"

// ---- Gaussian functions (gaussians.h:10-39) with their AD overloads (AADExpr.h:420-446) ------------
inline double normalDens(const double x) { return x < -10.0 || 10.0 < x ? 0.0 : std::exp(-0.5 * x * x) / 2.506628274631; }

// Zelen and Severo's approximation (gaussians.h:23-39)
inline double normalCdf(const double x)
{
    if (x < -10.0) return 0.0;
    if (x > 10.0) return 1.0;
    if (x < 0.0) return 1.0 - normalCdf(-x);
    const double p = 0.2316419, b1 = 0.319381530, b2 = -0.356563782, b3 = 1.781477937, b4 = -1.821255978, b5 = 1.330274429;
    const double t = 1.0 / (1.0 + p * x);
    const double pol = t * (b1 + t * (b2 + t * (b3 + t * (b4 + t * b5))));
    return 1.0 - normalDens(x) * pol;
}
inline Number normalDens(const Number& x) { return Number::fromUnary(normalDens(x.value()), x, -x.value() * normalDens(x.value())); }
inline Number normalCdf(const Number& x) { return Number::fromUnary(normalCdf(x.value()), x, normalDens(x.value())); }

// ---- Black-Scholes and Merton (analytics.h:8-91) ------------------------------------------------------
template <class T, class U, class V, class W>
inline T blackScholes(const U spot, const V strike, const T vol, const W mat)
{
    const auto std_ = vol * std::sqrt(mat);
    if (std_ <= EPS) return T(std::max(0.0, double(spot - strike)));
    const auto d2 = std::log(spot / strike) / std_ - 0.5 * std_;
    const auto d1 = d2 + std_;
    return spot * normalCdf(d1) - strike * normalCdf(d2);
}

// Implied vol by bisection to 1e-12 then one linear interpolation (analytics.h:23-56)
inline double blackScholesIvol(const double spot, const double strike, const double prem, const double mat)
{
    if (prem <= std::max(0.0, spot - strike) + EPS) return 0.0;
    double p, pu, pl = 0.0;
    double u = 0.5;
    while (blackScholes(spot, strike, u, mat) < prem) u *= 2;
    double l = 0.05;
    while (blackScholes(spot, strike, l, mat) > prem) l /= 2;
    pu = blackScholes(spot, strike, u, mat);
    pl = blackScholes(spot, strike, l, mat);
    while (u - l > 1.e-12) {
        const double m = 0.5 * (u + l);
        p = blackScholes(spot, strike, m, mat);
        if (p > prem) { u = m; pu = p; }
        else { l = m; pl = p; }
    }
    return l + (prem - pl) / (pu - pl) * (u - l);
}

// Merton's jump-diffusion call as a Poisson mixture of Black-Scholes prices, 10 terms (analytics.h:59-91)
inline double merton(const double spot, const double strike, const double vol, const double mat, const double intens,
                     const double meanJmp, const double stdJmp)
{
    const double varJmp = stdJmp * stdJmp;
    const double mv2 = meanJmp + 0.5 * varJmp;
    const double comp = intens * (std::exp(mv2) - 1);
    const double var = vol * vol;
    const double intensT = intens * mat;
    unsigned fact = 1;
    double iT = 1.0;
    double result = 0.0;
    for (size_t n = 0; n < 10; ++n) {
        const double s = spot * std::exp(n * mv2 - comp * mat);
        const double v = std::sqrt(var + n * varJmp / mat);
        const double prob = std::exp(-intensT) * iT / fact;
        result += prob * blackScholes(s, strike, v, mat);
        fact *= unsigned(n + 1);
        iT *= intensT;
    }
    return result;
}

// ---- 2-D interpolation (interp.h:65-109): smooth-step or linear in both directions, flat outside ----------
template <bool smoothStep = false, class T, class U, class V, class W, class X>
inline V interp2D(const std::vector<T>& x, const std::vector<U>& y, const matrix<V>& z, const W& x0, const X& y0)
{
    const size_t n = x.size(), m = y.size();
    const size_t n2 = size_t(std::distance(x.begin(), std::upper_bound(x.begin(), x.end(), x0)));
    if (n2 == n) return interp<smoothStep>(y.begin(), y.end(), z[n2 - 1], z[n2 - 1] + m, y0);
    if (n2 == 0) return interp<smoothStep>(y.begin(), y.end(), z[0], z[0] + m, y0);
    const size_t n1 = n2 - 1;
    auto z1 = interp<smoothStep>(y.begin(), y.end(), z[n1], z[n1] + m, y0);
    auto z2 = interp<smoothStep>(y.begin(), y.end(), z[n2], z[n2] + m, y0);
    auto t = (x0 - x[n1]) / (x[n2] - x[n1]);
    if constexpr (smoothStep) return z1 + (z2 - z1) * t * t * (3.0 - 2 * t);
    else return z1 + (z2 - z1) * t;
}

// ---- Risk view: additive spreads to the implied vols on a (strike, maturity) grid (ivs.h:27-88) ----------
template <class T>
class RiskView
{
    bool                myEmpty;
    std::vector<double> myStrikes;
    std::vector<Time>   myMats;
    matrix<T>           mySpreads;

public:
    RiskView() : myEmpty(true) {}
    // All spreads 0; for T = Number they are put on tape here, as in the reference
    RiskView(const std::vector<double>& strikes, const std::vector<Time>& mats)
        : myEmpty(false), myStrikes(strikes), myMats(mats), mySpreads(strikes.size(), mats.size())
    {
        for (auto& spr : mySpreads) {
            spr = T(0.0);
            if constexpr (std::is_same<T, Number>::value) spr.putOnTape();
        }
    }
    T spread(const double strike, const Time mat) const
    {
        return myEmpty ? T(0.0) : interp2D<true>(myStrikes, myMats, mySpreads, strike, mat);
    }
    bool empty() const { return myEmpty; }
    size_t rows() const { return myStrikes.size(); }
    size_t cols() const { return myMats.size(); }
    const std::vector<double>& strikes() const { return myStrikes; }
    const std::vector<Time>& mats() const { return myMats; }
    const matrix<T>& risks() const { return mySpreads; }
    typename matrix<T>::iterator begin() { return mySpreads.begin(); }
    typename matrix<T>::iterator end() { return mySpreads.end(); }
    typename matrix<T>::const_iterator begin() const { return mySpreads.begin(); }
    typename matrix<T>::const_iterator end() const { return mySpreads.end(); }
    void bump(const size_t i, const size_t j, const double bumpBy) { mySpreads[i][j] += bumpBy; }
};

// ---- Implied volatility surfaces (ivs.h:90-179) ----------------------------------------------------------------
class IVS
{
    double mySpot;

public:
    IVS(const double spot) : mySpot(spot) {}
    double spot() const { return mySpot; }
    virtual double impliedVol(const double strike, const Time mat) const = 0;

    template <class T = double>
    T call(const double strike, const Time mat, const RiskView<T>* risk = nullptr) const
    {
        return blackScholes<T>(mySpot, strike, impliedVol(strike, mat) + (risk ? risk->spread(strike, mat) : T(0.0)), mat);
    }

    // Dupire's formula with centred differences of 1e-4 in time and strike (ivs.h:119-138)
    template <class T = double>
    T localVol(const double strike, const double mat, const RiskView<T>* risk = nullptr) const
    {
        const T c00 = call(strike, mat, risk);
        const T c01 = call(strike, mat - 1.0e-04, risk);
        const T c02 = call(strike, mat + 1.0e-04, risk);
        const T ct = (c02 - c01) * 0.5e04;
        const T c10 = call(strike - 1.0e-04, mat, risk);
        const T c20 = call(strike + 1.0e-04, mat, risk);
        const T ckk = (c10 + c20 - 2.0 * c00) * 1.0e08;
        return sqrt(2.0 * ct / ckk) / strike;
    }
    virtual ~IVS() {}
};

class MertonIVS : public IVS
{
    double myVol, myIntensity, myAverageJmp, myJmpStd;

public:
    MertonIVS(const double spot, const double vol, const double intens, const double aveJmp, const double stdJmp)
        : IVS(spot), myVol(vol), myIntensity(intens), myAverageJmp(aveJmp), myJmpStd(stdJmp) {}
    double impliedVol(const double strike, const Time mat) const override
    {
        return blackScholesIvol(spot(), strike, merton(spot(), strike, myVol, mat, myIntensity, myAverageJmp, myJmpStd), mat);
    }
};

// ---- Calibration of the local-vol grid (mcMdlDupire.h:289-388) ----------------------------------------------
// One maturity: Dupire's formula within 2.5 standard deviations of the spot, flat outside
template <class IT, class OT, class T = double>
inline void dupireCalibMaturity(const IVS& ivs, const Time maturity, IT spotsBegin, IT spotsEnd, OT lVolsBegin,
                                const RiskView<T>& riskView = RiskView<double>())
{
    IT spots = spotsBegin;
    const int nSpots = int(std::distance(spotsBegin, spotsEnd));
    const double atmCall = double(ivs.call(ivs.spot(), maturity));
    const double std_ = atmCall * 2.506628274631;
    int il = 0;
    while (il < nSpots && spots[il] < ivs.spot() - 2.5 * std_) ++il;
    int ih = nSpots - 1;
    while (ih >= 0 && spots[ih] > ivs.spot() + 2.5 * std_) --ih;
    for (int i = il; i <= ih; ++i) lVolsBegin[i] = ivs.localVol(spots[i], maturity, &riskView);
    for (int i = 0; i < il; ++i) lVolsBegin[i] = lVolsBegin[il];
    for (int i = ih + 1; i < nSpots; ++i) lVolsBegin[i] = lVolsBegin[ih];
}

template <class T>
struct DupireCalibResults
{
    std::vector<double> spots;
    std::vector<Time>   times;
    matrix<T>           lVols;      // spot major
};

template <class T = double>
inline DupireCalibResults<T> dupireCalib(const IVS& ivs, const std::vector<double>& inclSpots, const double maxDs,
                                         const std::vector<Time>& inclTimes, const double maxDt,
                                         const RiskView<T>& riskView = RiskView<double>())
{
    DupireCalibResults<T> results;
    results.spots = fillData(inclSpots, maxDs, 0.01);
    results.times = fillData(inclTimes, maxDt, 0.000114469 /* one hour */, &maxDt, &maxDt + 1);   // includes maxDt itself
    matrix<T> lVolsT(results.times.size(), results.spots.size());
    for (size_t j = 0; j < results.times.size(); ++j)
        dupireCalibMaturity(ivs, results.times[j], results.spots.begin(), results.spots.end(), lVolsT[j], riskView);
    results.lVols = transpose(lVolsT);
    return results;
}

// main.h:413-447
inline DupireCalibResults<double> dupireCalib(const std::vector<double>& inclSpots, const double maxDs,
                                              const std::vector<Time>& inclTimes, const double maxDt, const double spot,
                                              const double vol, const double jmpIntens = 0.0, const double jmpAverage = 0.0,
                                              const double jmpStd = 0.0)
{
    MertonIVS ivs(spot, vol, jmpIntens, jmpAverage, jmpStd);
    return dupireCalib(ivs, inclSpots, maxDs, inclTimes, maxDt);
}
