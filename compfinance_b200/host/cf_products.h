// cf_products.h -- host side of the single-asset products of mcPrd.h: European (:29-126), UOC up-and-out call / put with
// a smoothed barrier (:128-288), Europeans portfolio (:290-401), ContingentBond (:404-574).
//
// INTERFACE-MANDATED: the four class names with their constructor argument lists (the order store.h:120-205 passes), the
// accessors the entry points use, the Product<T> virtuals, the payoff LABEL TEXTS (keys of the notionals maps and of the
// risk reports: "call 1.00 3.00", "call 3.00 120.00 up and out 150.00 monitoring freq 0.02 smooth 0.01", ...), the
// timelines -- schedules built by repeated addition of the period from systemTime, closed by the maturity when the last
// regular date falls short of it by more than ONE_HOUR / ONE_DAY (mcPrd.h:165-176, 437-452; they must match bit for bit)
// -- and what each sample asks of the model.
// OWN STRUCTURE: what the four share (timeline, defline, labels and the virtuals that return them) lives once in
// SingleAssetProduct<T>; the two schedules come from one helper; a product fills the base in its constructor and
// describes its constants to the engine in deviceImage().  The payoffs themselves are evaluated inside the path kernels
// (cf_kernels.cuh, cf_dupire.cuh, cf_bs.cuh).
#pragma once

#include "cf_base.h"

#define ONE_HOUR 0.000114469      // mcPrd.h:26
#define ONE_DAY 0.003773585       // mcPrd.h:27

namespace cfprd {

// "x.xx": how the reference's labels print dates, strikes, barriers (ostringstream, fixed, precision 2)
inline std::string fixed2(const double x)
{
    std::ostringstream text;
    text << std::fixed << std::setprecision(2) << x;
    return text.str();
}

// today, today + period, today + 2 periods (by repeated addition), ... while more than `slack` short of `end`; then `end`
inline std::vector<Time> schedule(const Time period, const Time end, const double slack)
{
    std::vector<Time> dates(1, systemTime);
    for (Time t = systemTime + period; end - t > slack; t += period) dates.push_back(t);
    dates.push_back(end);
    return dates;
}

}  // namespace cfprd

template <class T>
class SingleAssetProduct : public Product<T>
{
public:
    const std::vector<Time>& timeline() const override { return dates; }
    const std::vector<SampleDef>& defline() const override { return samples; }
    const std::vector<std::string>& payoffLabels() const override { return labels; }

protected:
    // one more event date observing the asset's forward to `forwardTo`
    SampleDef& addSample(const Time date, const Time forwardTo, const bool withNumeraire)
    {
        dates.push_back(date);
        samples.emplace_back();
        samples.back().numeraire = withNumeraire;
        samples.back().forwardMats.push_back(std::vector<Time>(1, forwardTo));
        return samples.back();
    }
    static bool oneAsset(const ModelImage& mdl) { return mdl.pod.n_assets == 1; }

    std::vector<Time>        dates;
    std::vector<SampleDef>   samples;
    std::vector<std::string> labels;
};

// ---- European call: max(F(exercise, settlement) - K, 0) discounted from settlement, over the numeraire (mcPrd.h:113-125)
template <class T>
class European : public SingleAssetProduct<T>
{
    using Base = SingleAssetProduct<T>;
    double k;
    Time   exercise, settlement;

public:
    European(const double strike, const Time exerciseDate, const Time settlementDate)
        : k(strike), exercise(exerciseDate), settlement(settlementDate)
    {
        Base::addSample(exerciseDate, settlementDate, true).discountMats.push_back(settlementDate);
        std::string text = "call " + cfprd::fixed2(strike) + " " + cfprd::fixed2(exerciseDate);
        if (settlementDate != exerciseDate) text += " " + cfprd::fixed2(settlementDate);
        Base::labels.push_back(text);
    }
    European(const double strike, const Time exerciseDate) : European(strike, exerciseDate, exerciseDate) {}

    double strike() const { return k; }
    Time exerciseDate() const { return exercise; }
    Time settlementDate() const { return settlement; }
    std::unique_ptr<Product<T>> clone() const override { return std::make_unique<European<T>>(*this); }

    bool deviceImage(ProductImage& img, const ModelImage& mdl) const override
    {
        if (!Base::oneAsset(mdl)) return false;
        img = ProductImage();
        img.pod.kind = CF_PRODUCT_EUROPEAN;
        img.pod.n_events = 1;
        img.pod.n_payoffs = 1;
        img.pod.strike = k;
        return true;
    }
};

// ---- UOC: payoff 0 the barrier option, payoff 1 its European; monitored on a schedule of `monitorFreq`, knocked out
// smoothly over [barrier - s, barrier + s] with s = smooth x the first sample's forward (mcPrd.h:235-288)
template <class T>
class UOC : public SingleAssetProduct<T>
{
    using Base = SingleAssetProduct<T>;
    struct Terms { double strike, barrier; Time maturity, monitorFreq; double smooth; bool put; };
    Terms terms;

public:
    UOC(const double strike, const double barrier, const Time maturity, const Time monitorFreq, const double smooth,
        const bool callPut = false)
        : terms{strike, barrier, maturity, monitorFreq, smooth, callPut}
    {
        const std::vector<Time> monitoring = cfprd::schedule(monitorFreq, maturity, ONE_HOUR);
        for (size_t i = 0; i < monitoring.size(); ++i)                   // spot(t) = forward(t, t); numeraire at maturity only
            Base::addSample(monitoring[i], monitoring[i], i + 1 == monitoring.size());
        const std::string european = std::string(callPut ? "put " : "call ") + cfprd::fixed2(maturity) + " " + cfprd::fixed2(strike);
        Base::labels.push_back(european + " up and out " + cfprd::fixed2(barrier) + " monitoring freq " + cfprd::fixed2(monitorFreq)
                               + " smooth " + cfprd::fixed2(smooth));
        Base::labels.push_back(european);
    }

    double strike() const { return terms.strike; }
    double barrier() const { return terms.barrier; }
    Time maturity() const { return terms.maturity; }
    Time monitorFreq() const { return terms.monitorFreq; }
    double smooth() const { return terms.smooth; }
    bool isPut() const { return terms.put; }
    std::unique_ptr<Product<T>> clone() const override { return std::make_unique<UOC<T>>(*this); }

    bool deviceImage(ProductImage& img, const ModelImage& mdl) const override
    {
        // the smoothing half-width is a plain double of the FIRST sample's forward (mcPrd.h:247); on the device it must
        // be path-independent, i.e. the first sample has to be today's
        if (!Base::oneAsset(mdl) || !mdl.firstSampleIsToday) return false;
        img = ProductImage();
        img.pod.kind = CF_PRODUCT_UOC;
        img.pod.n_events = int(Base::dates.size());
        img.pod.n_payoffs = 2;
        img.pod.is_put = terms.put ? 1 : 0;
        img.pod.strike = terms.strike;
        img.pod.barrier = terms.barrier;
        img.pod.smooth = double(mdl.firstSampleForward * terms.smooth);
        return true;
    }
};

// ---- ContingentBond (mcPrd.h:404-574): every period pays (libor + coupon) x coverage at its end if the asset went up
// over the period (smoothed digital), plus redemption at maturity.  One payoff.
template <class T>
class ContingentBond : public SingleAssetProduct<T>
{
    using Base = SingleAssetProduct<T>;
    Time                expiry;
    double              cpn, smoothing;
    std::vector<double> coverage;             // of the period starting at timeline point i

public:
    ContingentBond(const Time maturity, const double coupon, const Time payFreq, const double smooth)
        : expiry(maturity), cpn(coupon), smoothing(smooth)
    {
        const std::vector<Time> payments = cfprd::schedule(payFreq, maturity, ONE_DAY);
        for (size_t i = 0; i < payments.size(); ++i) {
            SampleDef& s = Base::addSample(payments[i], payments[i], i > 0);        // payments on every date but today
            if (i + 1 < payments.size()) {
                s.liborDefs.push_back(SampleDef::RateDef(payments[i], payments[i + 1], "libor"));
                coverage.push_back(payments[i + 1] - payments[i]);
            }
        }
        Base::labels.push_back("contingent bond " + cfprd::fixed2(maturity) + " " + cfprd::fixed2(coupon));
    }

    Time maturity() const { return expiry; }
    double coupon() const { return cpn; }
    double smooth() const { return smoothing; }
    const std::vector<double>& coverages() const { return coverage; }
    std::unique_ptr<Product<T>> clone() const override { return std::make_unique<ContingentBond<T>>(*this); }

    bool deviceImage(ProductImage& img, const ModelImage& mdl) const override
    {
        // the half-width of the digital is a plain double of the FIRST sample's forward (mcPrd.h:531): today's
        if (!Base::oneAsset(mdl) || mdl.pod.kind != CF_MODEL_BS || !mdl.firstSampleIsToday) return false;
        img = ProductImage();
        img.eventDt = coverage;
        img.pod.kind = CF_PRODUCT_CONTINGENT;
        img.pod.n_events = int(Base::dates.size());
        img.pod.n_payoffs = 1;
        img.pod.coupon = cpn;
        img.pod.smooth = double(mdl.firstSampleForward * smoothing);
        img.pod.event_dt = img.eventDt.data();
        return true;
    }
};

// ---- Europeans: calls of several strikes per maturity, in the order of the map (mcPrd.h:374-399)
template <class T>
class Europeans : public SingleAssetProduct<T>
{
    using Base = SingleAssetProduct<T>;
    std::vector<std::vector<double>> ladders;         // strikes per maturity

public:
    Europeans(const std::map<Time, std::vector<double>>& options)
    {
        for (const auto& [maturity, strikes] : options) {
            Base::addSample(maturity, maturity, true);
            ladders.push_back(strikes);
            for (const double strike : strikes) Base::labels.push_back("call " + cfprd::fixed2(maturity) + " " + cfprd::fixed2(strike));
        }
    }

    const std::vector<Time>& maturities() const { return Base::dates; }
    const std::vector<std::vector<double>>& strikes() const { return ladders; }
    std::unique_ptr<Product<T>> clone() const override { return std::make_unique<Europeans<T>>(*this); }

    bool deviceImage(ProductImage& img, const ModelImage& mdl) const override
    {
        if (!Base::oneAsset(mdl)) return false;
        img = ProductImage();
        img.strikeOffsets.push_back(0);
        for (const auto& ladder : ladders) {
            img.strikes.insert(img.strikes.end(), ladder.begin(), ladder.end());
            img.strikeOffsets.push_back(int32_t(img.strikes.size()));
        }
        img.pod.kind = CF_PRODUCT_EUROPEANS;
        img.pod.n_events = int(ladders.size());
        img.pod.n_payoffs = int(img.strikes.size());
        img.pod.strike_offsets = img.strikeOffsets.data();
        img.pod.strikes = img.strikes.data();
        return true;
    }
};
