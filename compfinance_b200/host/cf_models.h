// cf_models.h -- host side of the single-asset models: Black-Scholes (mcMdlBS.h:25-350) and
// Dupire local volatility (mcMdlDupire.h:28-281).  Constructors, parameter order / labels,
// allocate() and init() follow the reference; generatePath() runs on the device, described by
// deviceImage().
#pragma once

#include "cf_base.h"
#include "cf_util.h"

#define HALF_DAY 0.00136986301369863      // mcMdlDupire.h:26

// ---------------------------------------------------------------------------------------------
// Black-Scholes: constant vol / rate / dividend yield, exact log-normal steps between event dates
// ---------------------------------------------------------------------------------------------
template <class T>
class BlackScholes : public Model<T>
{
    T mySpot, myRate, myDiv, myVol;
    const bool mySpotMeasure;

    std::vector<Time> myTimeline;          // today + event dates after today
    bool              myTodayOnTimeline = false;

    // init() tables (mcMdlBS.h:203-276)
    std::vector<T> myStds, myDrifts, myNumeraires;
    std::vector<std::vector<T>> myDiscounts, myForwardFactors, myLibors;

    std::vector<T*>          myParameters;
    std::vector<std::string> myParameterLabels;

    void setParamPointers()
    {
        myParameters[0] = &mySpot; myParameters[1] = &myVol; myParameters[2] = &myRate; myParameters[3] = &myDiv;
    }

public:
    template <class U>
    BlackScholes(const U spot, const U vol, const bool spotMeasure = false, const U rate = U(0.0), const U div = U(0.0))
        : mySpot(spot), myRate(rate), myDiv(div), myVol(vol), mySpotMeasure(spotMeasure), myParameters(4),
          myParameterLabels({"spot", "vol", "rate", "div"})
    {
        setParamPointers();
    }

    T spot() const { return mySpot; }
    const T vol() const { return myVol; }
    const T rate() const { return myRate; }
    const T div() const { return myDiv; }
    bool spotMeasure() const { return mySpotMeasure; }

    const std::vector<T*>& parameters() override { return myParameters; }
    const std::vector<std::string>& parameterLabels() const override { return myParameterLabels; }

    std::unique_ptr<Model<T>> clone() const override
    {
        auto c = std::make_unique<BlackScholes<T>>(*this);
        c->setParamPointers();
        return c;
    }

    void allocate(const std::vector<Time>& productTimeline, const std::vector<SampleDef>& defline) override
    {
        myTimeline.clear();
        myTimeline.push_back(systemTime);
        for (const auto& time : productTimeline)
            if (time > systemTime) myTimeline.push_back(time);
        myTodayOnTimeline = (productTimeline[0] == systemTime);

        myStds.resize(myTimeline.size() - 1);
        myDrifts.resize(myTimeline.size() - 1);
        const size_t n = productTimeline.size();
        myNumeraires.resize(n);
        myDiscounts.resize(n);
        myForwardFactors.resize(n);
        myLibors.resize(n);
        for (size_t j = 0; j < n; ++j) {
            myDiscounts[j].resize(defline[j].discountMats.size());
            myForwardFactors[j].resize(defline[j].forwardMats.front().size());
            myLibors[j].resize(defline[j].liborDefs.size());
        }
    }

    void init(const std::vector<Time>& productTimeline, const std::vector<SampleDef>& defline) override
    {
        const T mu = myRate - myDiv;
        const size_t n = myTimeline.size() - 1;
        for (size_t i = 0; i < n; ++i) {
            const double dt = myTimeline[i + 1] - myTimeline[i];
            myStds[i] = myVol * std::sqrt(dt);
            // spot measure: + vol^2 / 2, risk-neutral: - vol^2 / 2 (mcMdlBS.h:213-224)
            myDrifts[i] = mySpotMeasure ? (mu + 0.5 * myVol * myVol) * dt : (mu - 0.5 * myVol * myVol) * dt;
        }
        const size_t m = productTimeline.size();
        for (size_t i = 0; i < m; ++i) {
            if (defline[i].numeraire)
                myNumeraires[i] = mySpotMeasure ? exp(myDiv * productTimeline[i]) / mySpot : exp(myRate * productTimeline[i]);
            for (size_t j = 0; j < defline[i].discountMats.size(); ++j)
                myDiscounts[i][j] = exp(-myRate * (defline[i].discountMats[j] - productTimeline[i]));
            for (size_t j = 0; j < defline[i].forwardMats.front().size(); ++j)
                myForwardFactors[i][j] = exp(mu * (defline[i].forwardMats.front()[j] - productTimeline[i]));
            for (size_t j = 0; j < defline[i].liborDefs.size(); ++j) {
                const double dt = defline[i].liborDefs[j].end - defline[i].liborDefs[j].start;
                myLibors[i][j] = (exp(myRate * dt) - 1.0) / dt;
            }
        }
    }

    size_t simDim() const override { return myTimeline.size() - 1; }

    // Device image: drifts / stds per step; per event date the numeraire, first forward factor, first
    // discount and first libor (what the single-asset products read: forwards[0][0], discounts[0],
    // libors[0], numeraire).  Adjoint layout: [spot, drift[D], std[D], numeraire[E], fwd factor[E], discount[E], libor[E]].
    bool deviceImage(ModelImage& img, const std::vector<Time>& productTimeline, const std::vector<SampleDef>& defline) override
    {
        if (mySpotMeasure) return false;      // numeraire depends on the path under the spot measure
        const size_t D = myTimeline.size() - 1, E = productTimeline.size();
        if (E != D + (myTodayOnTimeline ? 1 : 0)) return false;   // an event date before today
        img = ModelImage();
        img.isEvent.assign(D + 1, 1);
        img.isEvent[0] = myTodayOnTimeline ? 1 : 0;
        img.tabA.resize(D); img.tabB.resize(D);
        for (size_t i = 0; i < D; ++i) { img.tabA[i] = cfValue(myDrifts[i]); img.tabB[i] = cfValue(myStds[i]); }
        img.numeraires.assign(E, 1.0); img.fwdFactors.assign(E, 1.0); img.discounts.assign(E, 1.0); img.libors.assign(E, 0.0);
        for (size_t e = 0; e < E; ++e) {
            if (defline[e].forwardMats.front().size() > 1 || defline[e].discountMats.size() > 1 || defline[e].liborDefs.size() > 1) return false;
            if (!myLibors[e].empty()) img.libors[e] = cfValue(myLibors[e][0]);
            if (defline[e].numeraire) img.numeraires[e] = cfValue(myNumeraires[e]);
            if (!myForwardFactors[e].empty()) img.fwdFactors[e] = cfValue(myForwardFactors[e][0]);
            if (!myDiscounts[e].empty()) img.discounts[e] = cfValue(myDiscounts[e][0]);
        }
        cf_model& p = img.pod;
        p.kind = CF_MODEL_BS; p.n_assets = 1; p.n_steps = int(D); p.n_events = int(E);
        p.is_event = img.isEvent.data(); p.spot = cfValue(mySpot);
        p.bs_drifts = img.tabA.data(); p.bs_stds = img.tabB.data();
        p.numeraires = img.numeraires.data(); p.fwd_factors = img.fwdFactors.data(); p.discounts = img.discounts.data();
        p.libors = img.libors.data();
        img.firstSampleIsToday = myTodayOnTimeline;
        img.firstSampleForward = cfValue(mySpot) * img.fwdFactors[0];
        if constexpr (std::is_same<T, Number>::value) {
            auto& t = img.adjointTargets;
            t.assign(1 + 2 * D + 4 * E, nullptr);
            t[0] = &mySpot;
            for (size_t i = 0; i < D; ++i) { t[1 + i] = &myDrifts[i]; t[1 + D + i] = &myStds[i]; }
            for (size_t e = 0; e < E; ++e) {
                if (defline[e].numeraire) t[1 + 2 * D + e] = &myNumeraires[e];
                if (!myForwardFactors[e].empty()) t[1 + 2 * D + E + e] = &myForwardFactors[e][0];
                if (!myDiscounts[e].empty()) t[1 + 2 * D + 2 * E + e] = &myDiscounts[e][0];
                if (!myLibors[e].empty()) t[1 + 2 * D + 3 * E + e] = &myLibors[e][0];
            }
            for (auto*& q : t) if (q && !q->onTape()) q = nullptr;     // constants carry no adjoint
        }
        return true;
    }
};

// ---------------------------------------------------------------------------------------------
// Dupire: local vol surface bilinear in (log spot, time), log-Euler steps of at most maxDt
// ---------------------------------------------------------------------------------------------
template <class T>
class Dupire : public Model<T>
{
    T                         mySpot;
    const std::vector<double> mySpots;
    std::vector<double>       myLogSpots;
    const std::vector<Time>   myTimes;
    matrix<T>                 myVols;            // spot major: vol(spot i, time j) = myVols[i][j]
    const Time                myMaxDt;

    std::vector<Time> myTimeline;
    std::vector<bool> myCommonSteps;             // timeline point is an event date
    matrix<T>         myInterpVols;              // time major, already multiplied by sqrt(dt)
    std::vector<int>    myCol1, myCol2;          // init(): interpVols[i][.] = w1[i] vols[.][col1[i]] + w2[i] vols[.][col2[i]]
    std::vector<double> myW1, myW2;

    std::vector<T*>          myParameters;
    std::shared_ptr<std::vector<std::string>> myParameterLabels;   // immutable after construction: shared by the clones

    void setParamPointers()
    {
        myParameters[0] = &mySpot;
        std::transform(myVols.begin(), myVols.end(), std::next(myParameters.begin()), [](auto& vol) { return &vol; });
    }

public:
    template <class U>
    Dupire(const U spot, const std::vector<double> spots, const std::vector<Time> times, const matrix<U> vols,
           const Time maxDt = 0.25)
        : mySpot(spot), mySpots(spots), myLogSpots(spots.size()), myTimes(times), myVols(vols), myMaxDt(maxDt),
          myParameters(vols.rows() * vols.cols() + 1),
          myParameterLabels(std::make_shared<std::vector<std::string>>(vols.rows() * vols.cols() + 1))
    {
        std::transform(mySpots.begin(), mySpots.end(), myLogSpots.begin(), [](const double s) { return std::log(s); });
        (*myParameterLabels)[0] = "spot";
        size_t p = 0;
        for (size_t i = 0; i < myVols.rows(); ++i)
            for (size_t j = 0; j < myVols.cols(); ++j) {
                std::ostringstream ost;
                ost << std::setprecision(2) << std::fixed << "lvol " << mySpots[i] << " " << myTimes[j];
                (*myParameterLabels)[++p] = ost.str();
            }
        setParamPointers();
    }

    T spot() const { return mySpot; }
    const std::vector<double>& spots() const { return mySpots; }
    const std::vector<Time>& times() const { return myTimes; }
    const matrix<T>& vols() const { return myVols; }
    Time maxDt() const { return myMaxDt; }
    const std::vector<Time>& simulationTimeline() const { return myTimeline; }

    const std::vector<T*>& parameters() override { return myParameters; }
    const std::vector<std::string>& parameterLabels() const override { return *myParameterLabels; }

    std::unique_ptr<Model<T>> clone() const override
    {
        auto c = std::make_unique<Dupire<T>>(*this);
        c->setParamPointers();
        return c;
    }

    void allocate(const std::vector<Time>& productTimeline, const std::vector<SampleDef>&) override
    {
        // product timeline + today, no step longer than maxDt (mcMdlDupire.h:173-177)
        myTimeline = fillData(productTimeline, myMaxDt, HALF_DAY, &systemTime, &systemTime + 1);
        myCommonSteps.resize(myTimeline.size());
        std::transform(myTimeline.begin(), myTimeline.end(), myCommonSteps.begin(), [&](const Time t) {
            return std::binary_search(productTimeline.begin(), productTimeline.end(), t);
        });
        myInterpVols.resize(myTimeline.size() - 1, mySpots.size());
    }

    void init(const std::vector<Time>&, const std::vector<SampleDef>&) override
    {
        // vols interpolated in time at the LEFT end of each step, times sqrt(dt) (mcMdlDupire.h:202-216).  The time
        // bracket of a step is the same for every spot, so interp's upper_bound (interp.h:36-46) is done once per step;
        // the arithmetic per entry is interp's, in its order.  On the host tape an entry is ONE node with its two
        // parents (the flattened expression sqrtdt * (y1 + (y2 - y1) * t)), and the step's map
        //   interpVols[i][j] = w1[i] vols[j][col1[i]] + w2[i] vols[j][col2[i]]
        // is kept for deviceImage().
        const size_t n = myTimeline.size() - 1, m = myLogSpots.size(), nT = myTimes.size();
        myCol1.assign(n, 0); myCol2.assign(n, 0); myW1.assign(n, 0.0); myW2.assign(n, 0.0);
        for (size_t i = 0; i < n; ++i) {
            const double sqrtdt = std::sqrt(myTimeline[i + 1] - myTimeline[i]);
            const Time x0 = myTimeline[i];
            const auto it = std::upper_bound(myTimes.begin(), myTimes.end(), x0);
            if (it == myTimes.end() || it == myTimes.begin()) {         // flat extrapolation: a copy of the edge column
                const size_t k = it == myTimes.end() ? nT - 1 : 0;
                for (size_t j = 0; j < m; ++j) myInterpVols[i][j] = sqrtdt * myVols[j][k];
                myCol1[i] = myCol2[i] = int(k); myW1[i] = sqrtdt; myW2[i] = 0.0;
                continue;
            }
            const size_t k = size_t(std::distance(myTimes.begin(), it)) - 1;
            const double t = (x0 - myTimes[k]) / (myTimes[k + 1] - myTimes[k]);
            // derivatives as the sweep over sqrtdt * (y1 + (y2 - y1) * t) accumulates them; a zero weight is not a parent
            const double w2 = sqrtdt * t, w1 = t != 0.0 ? sqrtdt + (-(sqrtdt * t)) : sqrtdt;
            for (size_t j = 0; j < m; ++j) {
                const T& y1 = myVols[j][k];
                const T& y2 = myVols[j][k + 1];
                if constexpr (std::is_same<T, Number>::value) {
                    const double v = sqrtdt * (y1.value() + (y2.value() - y1.value()) * t);
                    myInterpVols[i][j] = t != 0.0 ? Number::fromBinary(v, y1, w1, y2, w2) : Number::fromUnary(v, y1, w1);
                } else
                    myInterpVols[i][j] = sqrtdt * (y1 + (y2 - y1) * t);
            }
            myCol1[i] = int(k); myCol2[i] = t != 0.0 ? int(k + 1) : int(k); myW1[i] = w1; myW2[i] = t != 0.0 ? w2 : 0.0;
        }
    }

    size_t simDim() const override { return myTimeline.size() - 1; }

    // The same map READ OFF THE TAPE of init() (leaf gradients of every table entry): true when every entry is a
    // combination of at most two local vols of its own spot row, the same columns and weights along the row.
    // Independent of what init() recorded; deviceImage() checks one against the other under CF_CHECK_TIME_MAP.
    bool timeMapFromTape(ModelImage& img) const
    {
        if constexpr (!std::is_same<T, Number>::value) return false;
        else {
            const size_t D = myTimeline.size() - 1, m = myLogSpots.size(), nT = myTimes.size();
            const Tape& tape = *Number::tape;
            img.col1.assign(D, 0); img.col2.assign(D, 0); img.w1.assign(D, 0.0); img.w2.assign(D, 0.0);
            std::vector<std::pair<int, double>> g;
            for (size_t i = 0; i < D; ++i)
                for (size_t j = 0; j < m; ++j) {
                    g.clear();
                    if (!myInterpVols[i][j].onTape()) return false;
                    tape.leafGradient(myInterpVols[i][j].index(), 1.0, g);
                    if (g.empty() || g.size() > 2) return false;
                    int c[2] = {-1, -1}; double w[2] = {0.0, 0.0};
                    for (size_t q = 0; q < g.size(); ++q) {
                        // leaf must be one of vols[j][*]
                        const int off = g[q].first - myVols[j][0].index();
                        if (!myVols[j][0].onTape() || off < 0 || off >= int(nT) || myVols[j][off].index() != g[q].first) return false;
                        c[q] = off; w[q] = g[q].second;
                    }
                    if (g.size() == 1) { c[1] = c[0]; w[1] = 0.0; }
                    if (c[0] > c[1]) { std::swap(c[0], c[1]); std::swap(w[0], w[1]); }
                    if (j == 0) { img.col1[i] = c[0]; img.col2[i] = c[1]; img.w1[i] = w[0]; img.w2[i] = w[1]; }
                    else if (img.col1[i] != c[0] || img.col2[i] != c[1] || img.w1[i] != w[0] || img.w2[i] != w[1]) return false;
                }
            return true;
        }
    }

    // Device image.  Adjoint layout without the time map: [spot, interpVols[D][m]]; with it:
    // [spot, vols[m][nTimes]] (the parameter order).  For T = Number the time map
    //   interpVols[i][j] = w1[i] vols[j][col1[i]] + w2[i] vols[j][col2[i]]
    // is the one init() recorded, used when the spot and every local vol are on tape; otherwise the table
    // adjoints come back and the tape sweep mark -> start does the chain rule.
    bool deviceImage(ModelImage& img, const std::vector<Time>& productTimeline, const std::vector<SampleDef>& defline) override
    {
        const size_t D = myTimeline.size() - 1, m = myLogSpots.size(), E = productTimeline.size(), nT = myTimes.size();
        for (const auto& def : defline) if (!def.liborDefs.empty()) return false;
        img = ModelImage();
        img.isEvent.resize(D + 1);
        size_t nEv = 0;
        for (size_t i = 0; i <= D; ++i) { img.isEvent[i] = myCommonSteps[i] ? 1 : 0; nEv += img.isEvent[i]; }
        if (nEv != E) return false;
        img.tabA.resize(D * m);
        for (size_t i = 0; i < D; ++i)
            for (size_t j = 0; j < m; ++j) img.tabA[i * m + j] = cfValue(myInterpVols[i][j]);
        img.tabB = myLogSpots;
        cf_model& p = img.pod;
        p.kind = CF_MODEL_DUPIRE; p.n_assets = 1; p.n_steps = int(D); p.n_events = int(E);
        p.is_event = img.isEvent.data(); p.spot = cfValue(mySpot);
        p.n_knots = int(m); p.log_spots = img.tabB.data(); p.interp_vols = img.tabA.data();
        img.firstSampleIsToday = myCommonSteps[0];
        img.firstSampleForward = std::exp(std::log(cfValue(mySpot)));      // exp(logspot), mcMdlDupire.h:252
        if constexpr (std::is_same<T, Number>::value) {
            // the map recorded by init(); usable when the spot and every local vol are leaves of the tape
            bool structured = mySpot.onTape() && myCol1.size() == D
                              && std::all_of(myVols.begin(), myVols.end(), [](const Number& v) { return v.onTape(); });
            img.col1 = myCol1; img.col2 = myCol2; img.w1 = myW1; img.w2 = myW2;
            static const bool check = std::getenv("CF_CHECK_TIME_MAP") != nullptr;
            if (structured && check) {
                ModelImage ref;
                if (!timeMapFromTape(ref) || ref.col1 != img.col1 || ref.col2 != img.col2 || ref.w1 != img.w1 || ref.w2 != img.w2)
                    throw std::runtime_error("Dupire::deviceImage: init()'s time map differs from the one read off the tape");
            }
            auto& t = img.adjointTargets;
            if (structured) {
                p.n_times = int(nT);
                p.time_col1 = img.col1.data(); p.time_col2 = img.col2.data();
                p.time_w1 = img.w1.data(); p.time_w2 = img.w2.data();
                t.assign(1 + m * nT, nullptr);
                t[0] = &mySpot;
                for (size_t j = 0; j < m; ++j)
                    for (size_t k = 0; k < nT; ++k) t[1 + j * nT + k] = &myVols[j][k];
            } else {
                t.assign(1 + D * m, nullptr);
                t[0] = mySpot.onTape() ? &mySpot : nullptr;
                for (size_t i = 0; i < D; ++i)
                    for (size_t j = 0; j < m; ++j) t[1 + i * m + j] = myInterpVols[i][j].onTape() ? &myInterpVols[i][j] : nullptr;
            }
        }
        return true;
    }
};
