// cf_products_multi.h -- host side of the multi-asset products of mcPrdMulti.h: MultiStats (:11-179, moment payoffs, a
// test instrument), Baskets (:181-287, a strike ladder on a weighted basket) and Autocall (:289-439, worst-of
// autocallable with a smoothed knock-out).
//
// INTERFACE-MANDATED: the three class names with their constructor argument lists (the order store.h:207-255 passes),
// the Product<T> virtuals, the payoff LABEL TEXTS (they are the keys of the notionals maps and of the risk reports:
// "<asset> <fix> <fwd>", "basket strike 100.00", "autocall strike 70 KO 100 CPN 10 12 periods of 3m"), the timelines --
// the Autocall's by repeated addition of maturity / periods from systemTime (mcPrdMulti.h:324-332) --
// and what each sample asks of the model (one forward per asset at the given maturity, numeraire or not).
// OWN STRUCTURE: what the three share (asset names, timeline, defline, labels and the virtuals that return them) lives
// once in MultiAssetProduct<T>; a product only fills it in its constructor and describes its payoff constants to the
// engine in deviceImage().  The payoffs themselves are evaluated inside the path kernel (cf_dlm.cuh).
#pragma once

#include "cf_base.h"
#include "cf_products.h"      // cfprd::fixed2

namespace cfprd {

// whole percent / whole months of the Autocall label: int(100 x + EPS)
inline std::string whole(const double x) { return std::to_string(int(x + EPS)); }

}  // namespace cfprd

template <class T>
class MultiAssetProduct : public Product<T>
{
public:
    const size_t numAssets() const override { return names.size(); }
    const std::vector<std::string>& assetNames() const override { return names; }
    const std::vector<Time>& timeline() const override { return dates; }
    const std::vector<SampleDef>& defline() const override { return samples; }
    const std::vector<std::string>& payoffLabels() const override { return labels; }

protected:
    explicit MultiAssetProduct(const std::vector<std::string>& assets) : names(assets) {}

    // one more event date: every asset's forward to `forwardTo`, with or without the numeraire
    void addSample(const Time date, const Time forwardTo, const bool withNumeraire)
    {
        dates.push_back(date);
        samples.emplace_back();
        samples.back().numeraire = withNumeraire;
        samples.back().forwardMats.assign(names.size(), std::vector<Time>(1, forwardTo));
    }
    // the engine runs these products under the displaced multi-asset model only, asset for asset
    bool fits(const ModelImage& mdl) const { return mdl.pod.kind == CF_MODEL_DISPLACED && size_t(mdl.pod.n_assets) == names.size(); }

    std::vector<std::string> names;
    std::vector<Time>        dates;
    std::vector<SampleDef>   samples;
    std::vector<std::string> labels;
};

// ---- MultiStats: forwards and their pairwise products on every fixing date, then the same on increments between
// consecutive fixings (mcPrdMulti.h:60-110): the moments of the model, read against closed forms in testDLM.xlsx
template <class T>
class MultiStats : public MultiAssetProduct<T>
{
    using Base = MultiAssetProduct<T>;
    std::vector<Time> fwdTo;

    // "<asset>" and "<asset> <asset'>" (asset' <= asset) with a common tail
    void addMoments(const std::string& tail)
    {
        const auto& n = Base::names;
        for (const auto& a : n) Base::labels.push_back(a + " " + tail);
        for (size_t i = 0; i < n.size(); ++i)
            for (size_t j = 0; j <= i; ++j) Base::labels.push_back(n[i] + " " + n[j] + " " + tail);
    }

public:
    MultiStats(const std::vector<std::string>& assets, const std::vector<Time>& fixDates, const std::vector<Time>& fwdDates)
        : Base(assets), fwdTo(fwdDates)
    {
        std::vector<std::string> stamp;          // "<fix> <fwd>" of every fixing
        for (size_t t = 0; t < fixDates.size(); ++t) {
            Base::addSample(fixDates[t], fwdDates[t], false);
            stamp.push_back(cfprd::fixed2(fixDates[t]) + " " + cfprd::fixed2(fwdDates[t]));
        }
        for (const auto& s : stamp) addMoments(s);
        for (size_t t = 1; t < stamp.size(); ++t) addMoments(stamp[t - 1] + " - " + stamp[t]);
    }

    const std::vector<Time>& fixDates() const { return Base::dates; }
    const std::vector<Time>& fwdDates() const { return fwdTo; }
    std::unique_ptr<Product<T>> clone() const override { return std::make_unique<MultiStats<T>>(*this); }

    bool deviceImage(ProductImage& img, const ModelImage& mdl) const override
    {
        if (!Base::fits(mdl)) return false;
        img = ProductImage();
        img.pod.kind = CF_PRODUCT_MULTISTATS;
        img.pod.n_events = int(Base::dates.size());
        img.pod.n_payoffs = int(Base::labels.size());
        return true;
    }
};

// ---- Baskets: calls of every strike on sum_a w_a F_a at one maturity (mcPrdMulti.h:181-287)
template <class T>
class Baskets : public MultiAssetProduct<T>
{
    using Base = MultiAssetProduct<T>;
    std::vector<double> w, ks;
    Time                expiry;

public:
    Baskets(const std::vector<std::string>& assets, const std::vector<double> weights, const Time maturity,
            const std::vector<double>& strikes)
        : Base(assets), w(weights), ks(strikes), expiry(maturity)
    {
        Base::addSample(maturity, maturity, true);
        for (const double k : ks) Base::labels.push_back("basket strike " + cfprd::fixed2(k));
    }

    const std::vector<double>& weights() const { return w; }
    const std::vector<double>& strikes() const { return ks; }
    Time maturity() const { return expiry; }
    std::unique_ptr<Product<T>> clone() const override { return std::make_unique<Baskets<T>>(*this); }

    bool deviceImage(ProductImage& img, const ModelImage& mdl) const override
    {
        if (!Base::fits(mdl) || w.size() != Base::names.size()) return false;
        img = ProductImage();
        img.strikes = ks;
        img.weights = w;
        img.pod.kind = CF_PRODUCT_BASKETS;
        img.pod.n_events = 1;
        img.pod.n_payoffs = int(ks.size());
        img.pod.strikes = img.strikes.data();
        img.pod.weights = img.weights.data();
        return true;
    }
};

// ---- Autocall: per period a coupon on the notional still alive, a knock-out on the worst performance smoothed over
// [KO - smooth, KO + smooth], at maturity the put on the worst performance struck at `strike` (mcPrdMulti.h:289-439)
template <class T>
class Autocall : public MultiAssetProduct<T>
{
    using Base = MultiAssetProduct<T>;
    struct Terms { Time maturity; int periods; double ko, strike, coupon, smooth; };
    std::vector<double> refLevels;
    Terms               terms;

public:
    Autocall(const std::vector<std::string>& assets, const std::vector<double> refs, const Time maturity, const int periods,
             const double ko, const double strike, const double cpn, const double smooth)
        : Base(assets), refLevels(refs), terms{maturity, periods, ko, strike, cpn, std::max(smooth, EPS)}
    {
        const double length = maturity / periods;
        Time end = systemTime;
        for (int p = 0; p < periods; ++p) {
            end += length;                       // accumulated, not p * length: the timeline must match bit for bit
            Base::addSample(end, end, true);
        }
        Base::labels.push_back("autocall strike " + cfprd::whole(100 * strike) + " KO " + cfprd::whole(100 * ko) + " CPN "
                               + cfprd::whole(100 * cpn) + " " + std::to_string(periods) + " periods of "
                               + cfprd::whole(12 * maturity / periods) + "m");
    }

    const std::vector<double>& refs() const { return refLevels; }
    Time   maturity() const { return terms.maturity; }
    int    periods() const { return terms.periods; }
    double strike() const { return terms.strike; }
    double ko() const { return terms.ko; }
    double cpn() const { return terms.coupon; }
    double smooth() const { return terms.smooth; }
    std::unique_ptr<Product<T>> clone() const override { return std::make_unique<Autocall<T>>(*this); }

    bool deviceImage(ProductImage& img, const ModelImage& mdl) const override
    {
        if (!Base::fits(mdl) || refLevels.size() != Base::names.size()) return false;
        img = ProductImage();
        img.weights = refLevels;
        img.eventDt.assign(size_t(terms.periods), terms.maturity / terms.periods);      // coupon accrual (mcPrdMulti.h:400)
        img.pod.kind = CF_PRODUCT_AUTOCALL;
        img.pod.n_events = terms.periods;
        img.pod.n_payoffs = 1;
        img.pod.strike = terms.strike;
        img.pod.barrier = terms.ko;
        img.pod.smooth = terms.smooth;
        img.pod.coupon = terms.coupon;
        img.pod.weights = img.weights.data();
        img.pod.event_dt = img.eventDt.data();
        return true;
    }
};
