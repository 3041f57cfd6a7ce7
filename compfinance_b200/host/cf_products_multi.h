// cf_products_multi.h -- host side of the multi-asset products of mcPrdMulti.h: MultiStats
// (:11-179, moment payoffs, a test instrument), Baskets (:181-287, a strike ladder on a weighted
// basket) and Autocall (:289-439, worst-of autocallable with smoothed knock-out).  Constructors,
// timelines, deflines and labels follow the reference; the payoffs are evaluated inside the path
// kernel (cf_dlm.cuh), described by deviceImage().
#pragma once

#include "cf_base.h"

template <class T>
class MultiStats : public Product<T>
{
    std::vector<Time>        myFixDates, myFwdDates;
    size_t                   myNumAssets;
    std::vector<std::string> myAssetNames;
    std::vector<SampleDef>   myDefline;
    std::vector<std::string> myLabels;

public:
    MultiStats(const std::vector<std::string>& assets, const std::vector<Time>& fixDates, const std::vector<Time>& fwdDates)
        : myFixDates(fixDates), myFwdDates(fwdDates), myNumAssets(assets.size()), myAssetNames(assets)
    {
        const size_t nTimes = fixDates.size(), A = myNumAssets;
        myDefline.resize(nTimes);
        for (size_t i = 0; i < nTimes; ++i) {
            myDefline[i].numeraire = false;
            myDefline[i].forwardMats.assign(A, std::vector<Time>(1, myFwdDates[i]));
        }
        auto label = [&](const std::string& head, const std::string& tail) {
            myLabels.push_back(head + " " + tail);
        };
        auto dates = [&](const size_t t) {
            std::ostringstream ost;
            ost.precision(2);
            ost << std::fixed << myFixDates[t] << " " << myFwdDates[t];
            return ost.str();
        };
        // forwards and their products on every fixing date, then the same on increments (mcPrdMulti.h:60-110)
        for (size_t t = 0; t < nTimes; ++t) {
            for (size_t a1 = 0; a1 < A; ++a1) label(myAssetNames[a1], dates(t));
            for (size_t a1 = 0; a1 < A; ++a1)
                for (size_t a2 = 0; a2 <= a1; ++a2) label(myAssetNames[a1] + " " + myAssetNames[a2], dates(t));
        }
        for (size_t t2 = 1; t2 < nTimes; ++t2) {
            const std::string span = dates(t2 - 1) + " - " + dates(t2);
            for (size_t a1 = 0; a1 < A; ++a1) label(myAssetNames[a1], span);
            for (size_t a1 = 0; a1 < A; ++a1)
                for (size_t a2 = 0; a2 <= a1; ++a2) label(myAssetNames[a1] + " " + myAssetNames[a2], span);
        }
    }

    const size_t numAssets() const override { return myNumAssets; }
    const std::vector<std::string>& assetNames() const override { return myAssetNames; }
    const std::vector<Time>& fixDates() const { return myFixDates; }
    const std::vector<Time>& fwdDates() const { return myFwdDates; }
    std::unique_ptr<Product<T>> clone() const override { return std::make_unique<MultiStats<T>>(*this); }
    const std::vector<Time>& timeline() const override { return myFixDates; }
    const std::vector<SampleDef>& defline() const override { return myDefline; }
    const std::vector<std::string>& payoffLabels() const override { return myLabels; }

    bool deviceImage(ProductImage& img, const ModelImage& mdl) const override
    {
        if (size_t(mdl.pod.n_assets) != myNumAssets || mdl.pod.kind != CF_MODEL_DISPLACED) return false;
        img = ProductImage();
        img.pod.kind = CF_PRODUCT_MULTISTATS; img.pod.n_events = int(myFixDates.size()); img.pod.n_payoffs = int(myLabels.size());
        return true;
    }
};

template <class T>
class Baskets : public Product<T>
{
    size_t                   myNumAssets;
    std::vector<std::string> myAssetNames;
    std::vector<double>      myWeights;
    Time                     myMaturity;
    std::vector<double>      myStrikes;
    std::vector<Time>        myTimeline;
    std::vector<SampleDef>   myDefline;
    std::vector<std::string> myLabels;

public:
    Baskets(const std::vector<std::string>& assets, const std::vector<double> weights, const Time maturity,
            const std::vector<double>& strikes)
        : myNumAssets(assets.size()), myAssetNames(assets), myWeights(weights), myMaturity(maturity), myStrikes(strikes),
          myTimeline(1, maturity), myDefline(1)
    {
        myDefline[0].numeraire = true;
        myDefline[0].forwardMats = std::vector<std::vector<Time>>(myNumAssets, {maturity});
        for (const double strike : strikes) {
            std::ostringstream ost;
            ost.precision(2);
            ost << std::fixed << "basket strike " << strike;
            myLabels.push_back(ost.str());
        }
    }

    const size_t numAssets() const override { return myNumAssets; }
    const std::vector<std::string>& assetNames() const override { return myAssetNames; }
    const std::vector<double>& weights() const { return myWeights; }
    Time maturity() const { return myMaturity; }
    const std::vector<double>& strikes() const { return myStrikes; }
    std::unique_ptr<Product<T>> clone() const override { return std::make_unique<Baskets<T>>(*this); }
    const std::vector<Time>& timeline() const override { return myTimeline; }
    const std::vector<SampleDef>& defline() const override { return myDefline; }
    const std::vector<std::string>& payoffLabels() const override { return myLabels; }

    bool deviceImage(ProductImage& img, const ModelImage& mdl) const override
    {
        if (size_t(mdl.pod.n_assets) != myNumAssets || mdl.pod.kind != CF_MODEL_DISPLACED) return false;
        if (myWeights.size() != myNumAssets) return false;
        img = ProductImage();
        img.strikes = myStrikes; img.weights = myWeights;
        img.pod.kind = CF_PRODUCT_BASKETS; img.pod.n_events = 1; img.pod.n_payoffs = int(myStrikes.size());
        img.pod.strikes = img.strikes.data(); img.pod.weights = img.weights.data();
        return true;
    }
};

template <class T>
class Autocall : public Product<T>
{
    size_t                   myNumAssets;
    std::vector<std::string> myAssetNames;
    Time                     myMaturity;
    int                      myNumPeriods;
    std::vector<double>      myRefs;
    double                   myKO, myStrike, myCpn, mySmooth;
    std::vector<Time>        myTimeline;
    std::vector<SampleDef>   myDefline;
    std::vector<std::string> myLabels;

public:
    Autocall(const std::vector<std::string>& assets, const std::vector<double> refs, const Time maturity, const int periods,
             const double ko, const double strike, const double cpn, const double smooth)
        : myNumAssets(assets.size()), myAssetNames(assets), myMaturity(maturity), myNumPeriods(periods), myRefs(refs),
          myKO(ko), myStrike(strike), myCpn(cpn), mySmooth(std::max(smooth, EPS)), myTimeline(size_t(periods)),
          myDefline(size_t(periods)), myLabels(1)
    {
        // period ends by repeated addition of maturity / periods (mcPrdMulti.h:324-332)
        Time time = systemTime;
        const double dt = maturity / periods;
        for (int step = 0; step < periods; ++step) {
            time += dt;
            myTimeline[size_t(step)] = time;
            myDefline[size_t(step)].numeraire = true;
            myDefline[size_t(step)].forwardMats = std::vector<std::vector<Time>>(myNumAssets, {time});
        }
        myLabels[0] = "autocall strike " + std::to_string(int(100 * myStrike + EPS)) + " KO " + std::to_string(int(100 * myKO + EPS))
                      + " CPN " + std::to_string(int(100 * myCpn + EPS)) + " " + std::to_string(periods) + " periods of "
                      + std::to_string(int(12 * maturity / periods + EPS)) + "m";
    }

    const size_t numAssets() const override { return myNumAssets; }
    const std::vector<std::string>& assetNames() const override { return myAssetNames; }
    const std::vector<double>& refs() const { return myRefs; }
    Time maturity() const { return myMaturity; }
    int periods() const { return myNumPeriods; }
    double strike() const { return myStrike; }
    double ko() const { return myKO; }
    double cpn() const { return myCpn; }
    double smooth() const { return mySmooth; }
    std::unique_ptr<Product<T>> clone() const override { return std::make_unique<Autocall<T>>(*this); }
    const std::vector<Time>& timeline() const override { return myTimeline; }
    const std::vector<SampleDef>& defline() const override { return myDefline; }
    const std::vector<std::string>& payoffLabels() const override { return myLabels; }

    bool deviceImage(ProductImage& img, const ModelImage& mdl) const override
    {
        if (size_t(mdl.pod.n_assets) != myNumAssets || mdl.pod.kind != CF_MODEL_DISPLACED) return false;
        if (myRefs.size() != myNumAssets) return false;
        img = ProductImage();
        img.weights = myRefs;
        img.eventDt.assign(size_t(myNumPeriods), myMaturity / myNumPeriods);      // coupon accrual (mcPrdMulti.h:400)
        img.pod.kind = CF_PRODUCT_AUTOCALL; img.pod.n_events = myNumPeriods; img.pod.n_payoffs = 1;
        img.pod.strike = myStrike; img.pod.barrier = myKO; img.pod.smooth = mySmooth; img.pod.coupon = myCpn;
        img.pod.weights = img.weights.data(); img.pod.event_dt = img.eventDt.data();
        return true;
    }
};
