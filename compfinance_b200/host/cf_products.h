// cf_products.h -- host side of the single-asset products of mcPrd.h: European (:29-126),
// UOC up-and-out call/put with smoothed barrier (:128-288), Europeans portfolio (:290-401).
// Constructors, timelines, deflines and payoff labels follow the reference; the payoffs themselves
// are evaluated inside the path kernels (cf_kernels.cuh / cf_dupire.cuh), described by deviceImage().
#pragma once

#include "cf_base.h"

#define ONE_HOUR 0.000114469      // mcPrd.h:26
#define ONE_DAY 0.003773585       // mcPrd.h:27

template <class T>
class European : public Product<T>
{
    double myStrike;
    Time   myExerciseDate, mySettlementDate;
    std::vector<Time>        myTimeline;
    std::vector<SampleDef>   myDefline;
    std::vector<std::string> myLabels;

public:
    European(const double strike, const Time exerciseDate, const Time settlementDate)
        : myStrike(strike), myExerciseDate(exerciseDate), mySettlementDate(settlementDate), myLabels(1)
    {
        myTimeline.push_back(exerciseDate);
        myDefline.resize(1);
        SampleDef& def = myDefline.front();
        def.numeraire = true;
        def.forwardMats.push_back({settlementDate});
        def.discountMats.push_back(settlementDate);

        std::ostringstream ost;
        ost.precision(2);
        ost << std::fixed << "call " << myStrike << " " << exerciseDate;
        if (settlementDate != exerciseDate) ost << " " << settlementDate;
        myLabels[0] = ost.str();
    }
    European(const double strike, const Time exerciseDate) : European(strike, exerciseDate, exerciseDate) {}

    double strike() const { return myStrike; }
    Time exerciseDate() const { return myExerciseDate; }
    Time settlementDate() const { return mySettlementDate; }

    std::unique_ptr<Product<T>> clone() const override { return std::make_unique<European<T>>(*this); }
    const std::vector<Time>& timeline() const override { return myTimeline; }
    const std::vector<SampleDef>& defline() const override { return myDefline; }
    const std::vector<std::string>& payoffLabels() const override { return myLabels; }

    bool deviceImage(ProductImage& img, const ModelImage& mdl) const override
    {
        if (mdl.pod.n_assets != 1) return false;
        img = ProductImage();
        img.pod.kind = CF_PRODUCT_EUROPEAN; img.pod.n_events = 1; img.pod.n_payoffs = 1;
        img.pod.strike = myStrike;
        return true;
    }
};

template <class T>
class UOC : public Product<T>
{
    bool   myCallPut;          // false = call, true = put
    double myStrike, myBarrier;
    Time   myMaturity;
    double mySmooth;
    Time   myMonitorFreq;
    std::vector<Time>        myTimeline;
    std::vector<SampleDef>   myDefline;
    std::vector<std::string> myLabels;

public:
    UOC(const double strike, const double barrier, const Time maturity, const Time monitorFreq, const double smooth,
        const bool callPut = false)
        : myCallPut(callPut), myStrike(strike), myBarrier(barrier), myMaturity(maturity), mySmooth(smooth),
          myMonitorFreq(monitorFreq), myLabels(2)
    {
        // today, then every monitoring date by repeated addition, then maturity (mcPrd.h:165-176)
        myTimeline.push_back(systemTime);
        for (Time t = systemTime + monitorFreq; myMaturity - t > ONE_HOUR; t += monitorFreq) myTimeline.push_back(t);
        myTimeline.push_back(myMaturity);

        const size_t n = myTimeline.size();
        myDefline.resize(n);
        for (size_t i = 0; i < n; ++i) {
            myDefline[i].numeraire = (i + 1 == n);              // numeraire on the last date only
            myDefline[i].forwardMats.push_back({myTimeline[i]});  // spot(t) = forward(t, t)
        }

        std::ostringstream ost;
        ost.precision(2);
        ost << std::fixed << (myCallPut ? "put " : "call ") << myMaturity << " " << myStrike;
        myLabels[1] = ost.str();
        ost << " up and out " << myBarrier << " monitoring freq " << monitorFreq << " smooth " << mySmooth;
        myLabels[0] = ost.str();
    }

    double strike() const { return myStrike; }
    double barrier() const { return myBarrier; }
    Time maturity() const { return myMaturity; }
    Time monitorFreq() const { return myMonitorFreq; }
    double smooth() const { return mySmooth; }
    bool isPut() const { return myCallPut; }

    std::unique_ptr<Product<T>> clone() const override { return std::make_unique<UOC<T>>(*this); }
    const std::vector<Time>& timeline() const override { return myTimeline; }
    const std::vector<SampleDef>& defline() const override { return myDefline; }
    const std::vector<std::string>& payoffLabels() const override { return myLabels; }

    bool deviceImage(ProductImage& img, const ModelImage& mdl) const override
    {
        // The smoothing half-width is a plain double of the FIRST sample's forward (mcPrd.h:247); on the
        // device it must be path-independent, i.e. the first sample has to be today's.
        if (mdl.pod.n_assets != 1 || !mdl.firstSampleIsToday) return false;
        img = ProductImage();
        img.pod.kind = CF_PRODUCT_UOC; img.pod.n_events = int(myTimeline.size()); img.pod.n_payoffs = 2;
        img.pod.is_put = myCallPut ? 1 : 0;
        img.pod.strike = myStrike; img.pod.barrier = myBarrier;
        img.pod.smooth = double(mdl.firstSampleForward * mySmooth);
        return true;
    }
};

// Contingent floater (mcPrd.h:404-574): every period pays (libor + coupon) x coverage at its end if the asset went
// up over the period (smoothed digital), plus redemption at maturity.  One payoff.
template <class T>
class ContingentBond : public Product<T>
{
    Time   myMaturity;
    double myCpn, mySmooth;
    std::vector<Time>        myTimeline;
    std::vector<SampleDef>   myDefline;
    std::vector<std::string> myLabels;
    std::vector<double>      myDt;              // coverage of the period starting at timeline point i

public:
    ContingentBond(const Time maturity, const double cpn, const Time payFreq, const double smooth)
        : myMaturity(maturity), myCpn(cpn), mySmooth(smooth), myLabels(1)
    {
        // today, then every payment date by repeated addition, then maturity (mcPrd.h:437-452)
        myTimeline.push_back(systemTime);
        for (Time t = systemTime + payFreq; myMaturity - t > ONE_DAY; t += payFreq) {
            myDt.push_back(t - myTimeline.back());
            myTimeline.push_back(t);
        }
        myDt.push_back(myMaturity - myTimeline.back());
        myTimeline.push_back(myMaturity);

        const size_t n = myTimeline.size();
        myDefline.resize(n);
        for (size_t i = 0; i < n; ++i) {
            myDefline[i].forwardMats.push_back({myTimeline[i]});             // spot(T_i)
            if (i + 1 < n) myDefline[i].liborDefs.push_back(SampleDef::RateDef(myTimeline[i], myTimeline[i + 1], "libor"));
            myDefline[i].numeraire = i > 0;                                    // payments on every date but today
        }
        std::ostringstream ost;
        ost.precision(2);
        ost << std::fixed << "contingent bond " << myMaturity << " " << myCpn;
        myLabels[0] = ost.str();
    }

    Time maturity() const { return myMaturity; }
    double coupon() const { return myCpn; }
    double smooth() const { return mySmooth; }
    const std::vector<double>& coverages() const { return myDt; }

    std::unique_ptr<Product<T>> clone() const override { return std::make_unique<ContingentBond<T>>(*this); }
    const std::vector<Time>& timeline() const override { return myTimeline; }
    const std::vector<SampleDef>& defline() const override { return myDefline; }
    const std::vector<std::string>& payoffLabels() const override { return myLabels; }

    bool deviceImage(ProductImage& img, const ModelImage& mdl) const override
    {
        // the half-width of the digital is a plain double of the FIRST sample's forward (mcPrd.h:531): today's
        if (mdl.pod.n_assets != 1 || mdl.pod.kind != CF_MODEL_BS || !mdl.firstSampleIsToday) return false;
        img = ProductImage();
        img.pod.kind = CF_PRODUCT_CONTINGENT; img.pod.n_events = int(myTimeline.size()); img.pod.n_payoffs = 1;
        img.pod.coupon = myCpn;
        img.pod.smooth = double(mdl.firstSampleForward * mySmooth);
        img.eventDt = myDt;
        img.pod.event_dt = img.eventDt.data();
        return true;
    }
};

template <class T>
class Europeans : public Product<T>
{
    std::vector<Time>                myMaturities;
    std::vector<std::vector<double>> myStrikes;
    std::vector<SampleDef>           myDefline;
    std::vector<std::string>         myLabels;

public:
    Europeans(const std::map<Time, std::vector<double>>& options)
    {
        for (const auto& p : options) { myMaturities.push_back(p.first); myStrikes.push_back(p.second); }
        const size_t n = options.size();
        myDefline.resize(n);
        for (size_t i = 0; i < n; ++i) {
            myDefline[i].numeraire = true;
            myDefline[i].forwardMats.push_back({myMaturities[i]});
        }
        for (const auto& option : options)
            for (const auto& strike : option.second) {
                std::ostringstream ost;
                ost.precision(2);
                ost << std::fixed << "call " << option.first << " " << strike;
                myLabels.push_back(ost.str());
            }
    }

    const std::vector<Time>& maturities() const { return myMaturities; }
    const std::vector<std::vector<double>>& strikes() const { return myStrikes; }

    std::unique_ptr<Product<T>> clone() const override { return std::make_unique<Europeans<T>>(*this); }
    const std::vector<Time>& timeline() const override { return myMaturities; }
    const std::vector<SampleDef>& defline() const override { return myDefline; }
    const std::vector<std::string>& payoffLabels() const override { return myLabels; }

    bool deviceImage(ProductImage& img, const ModelImage& mdl) const override
    {
        if (mdl.pod.n_assets != 1) return false;
        img = ProductImage();
        img.strikeOffsets.push_back(0);
        for (const auto& ks : myStrikes) {
            img.strikes.insert(img.strikes.end(), ks.begin(), ks.end());
            img.strikeOffsets.push_back(int32_t(img.strikes.size()));
        }
        img.pod.kind = CF_PRODUCT_EUROPEANS; img.pod.n_events = int(myMaturities.size());
        img.pod.n_payoffs = int(img.strikes.size());
        img.pod.strike_offsets = img.strikeOffsets.data(); img.pod.strikes = img.strikes.data();
        return true;
    }
};
