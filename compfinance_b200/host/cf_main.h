// cf_main.h -- the entry points of the reference's main.h over the CUDA engine.
//
// INTERFACE-MANDATED (kept so that a caller of main.h compiles unchanged): the names, argument lists and result structs
// of value (main.h:43-96), AADriskOne (:99-173), AADriskAggregate (:176-254), AADriskMulti (:269-312), bumpRisk
// (:316-359), dupireAADRisk (:364-411), dupireSuperbucket (:453-569), dupireSuperbucketBump (:575-696), NumericalParam
// (:33-40), the error texts, and -- for the two finite-difference drivers -- the order of the floating-point operations
// of a difference quotient (bump sizes 1e-8 / 1e-5, "(bumped - base) * 1e+8"), which the parity tests compare with
// the reference's own drivers.
//
// OWN STRUCTURE: every entry point is a thin request on three helpers -- cfdrv::Objects (the stored pair and the key of
// its resident session, cf_base.h), cfdrv::Book (notionals -> weight vector -> book value) and cfdrv::differences (a
// finite-difference sweep over a list of bumps).  Means over paths come from deterministic device reductions instead
// of std::accumulate over a per-path matrix (main.h:69-74), and a run is sharded over the devices NumericalParam
// names (the reference: over the threads of its pool).
#pragma once

#include <functional>

#include "cf_calib.h"
#include "cf_rng.h"
#include "cf_store.h"

// main.h:33-40, plus the devices to run on.  `devices` empty: the engine's current context
// (cf_init, or the current CUDA device); otherwise the CUDA ordinals of a single-process multi-device context, opened
// when it differs from the current one.  `parallel` is kept for source compatibility: the serial and the parallel
// algorithms of the reference return the same numbers (exact skip-ahead), and so does the device.
struct NumericalParam
{
    bool parallel;
    bool useSobol;
    int  numPath;
    int  seed1 = 12345;
    int  seed2 = 1234;
    std::vector<int> devices;
};

struct ValueResults
{
    std::vector<std::string> identifiers;
    std::vector<double>      values;
};

struct AADRiskResults
{
    std::vector<std::string> payoffIds;
    std::vector<double>      payoffValues;
    double                   riskPayoffValue;
    std::vector<std::string> paramIds;
    std::vector<double>      risks;
};

// Values and a matrix of risks, payoffs in columns and parameters in rows (main.h:259-265)
struct RiskReports
{
    std::vector<std::string> payoffs;
    std::vector<std::string> params;
    std::vector<double>      values;
    matrix<double>           risks;
};

struct DupireRiskResults
{
    double         value;
    double         delta;
    matrix<double> vega;
};

struct SuperbucketResults
{
    double              value;
    double              delta;
    std::vector<double> strikes;
    std::vector<Time>   mats;
    matrix<double>      vega;
};

namespace cfdrv {

// Opens the device context NumericalParam asks for, if it is not the one already open.
inline void useDevices(const NumericalParam& num)
{
    static std::vector<int> current;
    if (num.devices.empty()) return;
    if (num.devices == current && cf_device_count() == int(current.size())) return;
    cfDropSessions();                                   // their plans live on the devices of the context being closed
    cfCheck(cf_init(int(num.devices.size()), num.devices.data()));
    current = num.devices;
}

inline std::unique_ptr<RNG> makeRng(const NumericalParam& num)
{
    if (num.useSobol) return std::make_unique<Sobol>();
    return std::make_unique<mrg32k3a>(num.seed1, num.seed2);
}

// The stored <T> twins of a (model, product) pair and what identifies their resident session.
template <class T>
struct Objects
{
    const Model<T>*      model = nullptr;
    const Product<T>*    product = nullptr;
    std::unique_ptr<RNG> rng;
    CfSessionKey         key;

    Objects(const std::string& modelId, const std::string& productId, const NumericalParam& num, const char* notFound)
    {
        useDevices(num);
        model = getModel<T>(modelId);
        product = getProduct<T>(productId);
        if (!model || !product) throw std::runtime_error(notFound);
        rng = makeRng(num);
        key = cfMakeKey(cfModelSerial(modelId), cfProductSerial(productId), *rng, std::is_same<T, Number>::value);
    }
};

inline std::vector<double> perPath(const std::vector<double>& sums, const int nPath)
{
    std::vector<double> means(sums.size());
    for (size_t i = 0; i < sums.size(); ++i) means[i] = sums[i] / nPath;
    return means;
}

// Position of a payoff label; `who` prefixes the reference's error text.
inline size_t payoffIndex(const std::vector<std::string>& labels, const std::string& label, const char* who)
{
    const auto it = std::find(labels.begin(), labels.end(), label);
    if (it == labels.end()) throw std::runtime_error(std::string(who) + " : payoff not found");
    return size_t(it - labels.begin());
}

// A book: notionals by payoff label (main.h:196-207) as a weight vector in payoff order.
struct Book
{
    std::vector<double> weights;
    Book(const std::vector<std::string>& labels, const std::map<std::string, double>& notionals, const char* who)
        : weights(labels.size(), 0.0)
    {
        for (const auto& entry : notionals) weights[payoffIndex(labels, entry.first, who)] = entry.second;
    }
    double value(const std::vector<double>& payoffValues) const
    {
        return std::inner_product(weights.begin(), weights.end(), payoffValues.begin(), 0.0);
    }
};

inline AADRiskResults report(const Objects<Number>& o, const std::vector<double>& weights, const NumericalParam& num)
{
    const AADSums sums = cfSimulAADSums(*o.product, *o.model, *o.rng, size_t(num.numPath), weights, nullptr, nullptr, &o.key);
    AADRiskResults r;
    r.payoffIds = o.product->payoffLabels();
    r.payoffValues = perPath(sums.payoffSums, num.numPath);
    r.riskPayoffValue = sums.aggSum / num.numPath;
    r.paramIds = o.model->parameterLabels();
    r.risks = sums.risks;
    return r;
}

// Finite differences: quotient(k) = (revalue(k) - base) * scale for every bump k, written to out(k).
inline void differences(const size_t nBumps, const double scale, const std::function<std::vector<double>(size_t)>& revalue,
                        const std::vector<double>& base, const std::function<void(size_t, size_t, double)>& out)
{
    for (size_t k = 0; k < nBumps; ++k) {
        const std::vector<double> bumped = revalue(k);
        for (size_t j = 0; j < base.size(); ++j) out(k, j, scale * (bumped[j] - base[j]));
    }
}

}  // namespace cfdrv

// ---- value (main.h:43-96)
inline ValueResults value(const Model<double>& model, const Product<double>& product, const NumericalParam& num)
{
    cfdrv::useDevices(num);
    const auto rng = cfdrv::makeRng(num);
    ValueResults r;
    r.identifiers = product.payoffLabels();
    r.values = cfdrv::perPath(cfSimulSums(product, model, *rng, size_t(num.numPath)), num.numPath);
    return r;
}

inline ValueResults value(const std::string& modelId, const std::string& productId, const NumericalParam& num)
{
    const cfdrv::Objects<double> o(modelId, productId, num, "value() : Could not retrieve model and product");
    ValueResults r;
    r.identifiers = o.product->payoffLabels();
    r.values = cfdrv::perPath(cfSimulSums(*o.product, *o.model, *o.rng, size_t(num.numPath), nullptr, &o.key), num.numPath);
    return r;
}

// ---- AAD risk of one payoff (main.h:99-173); an empty name means the first payoff
inline AADRiskResults AADriskOne(const std::string& modelId, const std::string& productId, const NumericalParam& num,
                                 const std::string& riskPayoff = "")
{
    const cfdrv::Objects<Number> o(modelId, productId, num, "AADrisk() : Could not retrieve model and product");
    const std::vector<std::string>& labels = o.product->payoffLabels();
    std::vector<double> pick(labels.size(), 0.0);
    pick[riskPayoff.empty() ? 0 : cfdrv::payoffIndex(labels, riskPayoff, "AADriskOne()")] = 1.0;
    return cfdrv::report(o, pick, num);
}

// ---- AAD risk of a book of payoffs (main.h:176-254)
inline AADRiskResults AADriskAggregate(const std::string& modelId, const std::string& productId,
                                       const std::map<std::string, double>& notionals, const NumericalParam& num)
{
    const cfdrv::Objects<Number> o(modelId, productId, num, "AADriskAggregate() : Could not retrieve model and product");
    return cfdrv::report(o, cfdrv::Book(o.product->payoffLabels(), notionals, "AADriskAggregate()").weights, num);
}

// ---- itemised AAD risk, one column per payoff (main.h:269-312)
inline RiskReports AADriskMulti(const std::string& modelId, const std::string& productId, const NumericalParam& num)
{
    const cfdrv::Objects<Number> o(modelId, productId, num, "AADrisk() : Could not retrieve model and product");
    AADMultiSums sums = cfSimulAADMultiSums(*o.product, *o.model, *o.rng, size_t(num.numPath), nullptr, &o.key);
    RiskReports r;
    r.params = o.model->parameterLabels();
    r.payoffs = o.product->payoffLabels();
    r.values = cfdrv::perPath(sums.payoffSums, num.numPath);
    r.risks = std::move(sums.risks);
    return r;
}

// ---- itemised risk by bumping every parameter by 1e-8 and re-valuing (main.h:316-359)
inline RiskReports bumpRisk(const std::string& modelId, const std::string& productId, const NumericalParam& num)
{
    const cfdrv::Objects<double> o(modelId, productId, num, "bumpRisk() : Could not retrieve model and product");
    const ValueResults base = value(*o.model, *o.product, num);
    RiskReports r;
    r.payoffs = base.identifiers;
    r.values = base.values;
    const auto scratch = o.model->clone();               // the stored model is never touched
    r.params = scratch->parameterLabels();
    const std::vector<double*>& knobs = scratch->parameters();
    r.risks.resize(knobs.size(), r.payoffs.size());
    cfdrv::differences(knobs.size(), 1.0e+08,
                       [&](const size_t i) {
                           *knobs[i] += 1.e-08;
                           std::vector<double> v = value(*scratch, *o.product, num).values;
                           *knobs[i] -= 1.e-08;
                           return v;
                       },
                       base.values, [&](const size_t i, const size_t j, const double q) { r.risks[i][j] = q; });
    return r;
}

// ---- Dupire: price, delta and the vega matrix to the local-vol surface (main.h:364-411) -- the north-star entry point
inline DupireRiskResults dupireAADRisk(const std::string& modelId, const std::string& productId,
                                       const std::map<std::string, double>& notionals, const NumericalParam& num)
{
    cfdrv::useDevices(num);
    const Model<Number>* stored = getModel<Number>(modelId);
    if (!stored) throw std::runtime_error("dupireAADRisk() : Model not found");
    const Dupire<Number>* dupire = dynamic_cast<const Dupire<Number>*>(stored);
    if (!dupire) throw std::runtime_error("dupireAADRisk() : Model not a Dupire");
    const cfdrv::Objects<Number> o(modelId, productId, num, "AADriskAggregate() : Could not retrieve model and product");
    // the aggregate risk without its 1081 parameter labels, which this entry point drops (main.h:399-408)
    const AADSums sums = cfSimulAADSums(*o.product, *o.model, *o.rng, size_t(num.numPath),
                                        cfdrv::Book(o.product->payoffLabels(), notionals, "AADriskAggregate()").weights,
                                        nullptr, nullptr, &o.key);
    DupireRiskResults r;
    r.value = sums.aggSum / num.numPath;
    r.delta = sums.risks[0];                             // parameter order: spot, then vols spot-major (mcMdlDupire.h:116-121)
    r.vega.resize(dupire->spots().size(), dupire->times().size());
    std::copy(sums.risks.begin() + 1, sums.risks.end(), r.vega.begin());
    return r;
}

// ---- Superbucket (main.h:453-569): calibrate, differentiate to the local vols on the GPU, chain to the risk view.
// The local-vol risks ("microbucket") seed the tape of a second calibration recorded with a RiskView<Number>.
inline SuperbucketResults dupireSuperbucket(const double spot, const double maxDt, const std::string& productId,
                                            const std::map<std::string, double>& notionals,
                                            const std::vector<double>& inclSpots, const double maxDs,
                                            const std::vector<Time>& inclTimes, const double maxDtVol,
                                            const std::vector<double>& strikes, const std::vector<Time>& mats, const double vol,
                                            const double jmpIntens, const double jmpAverage, const double jmpStd,
                                            const NumericalParam& num)
{
    SuperbucketResults r;
    r.strikes = strikes;
    r.mats = mats;

    // 1. the model, calibrated in plain doubles, and its risks to its own parameters
    const auto calibrated = dupireCalib(inclSpots, maxDs, inclTimes, maxDtVol, spot, vol, jmpIntens, jmpAverage, jmpStd);
    putDupire(spot, calibrated.spots, calibrated.times, calibrated.lVols, maxDt, "superbucket");
    const DupireRiskResults micro = dupireAADRisk("superbucket", productId, notionals, num);
    r.value = micro.value;
    r.delta = micro.delta;

    // 2. the same calibration on the tape, as a function of the spreads of the risk view
    Tape& tape = *Number::tape;
    tape.clear();
    const MertonIVS surface(spot, vol, jmpIntens, jmpAverage, jmpStd);
    RiskView<Number> view(strikes, mats);
    auto onTape = dupireCalib(surface, inclSpots, maxDs, inclTimes, maxDtVol, view);

    // 3. seed and sweep.  The seeds are ASSIGNED, as the reference does (main.h:541-547): a flat-extrapolated local vol
    // is a copy that shares its tape node with the edge of the calibrated range, so the last assignment wins there.
    matrix<Number>& lv = onTape.lVols;
    for (size_t i = 0; i < micro.vega.rows(); ++i)
        for (size_t j = 0; j < micro.vega.cols(); ++j)
            if (lv[i][j].onTape()) lv[i][j].adjoint() = micro.vega[i][j];
    if (tape.size() > 0) Number::propagateAdjoints(tape.end() - 1, tape.begin());
    r.vega.resize(view.rows(), view.cols());
    std::transform(view.begin(), view.end(), r.vega.begin(), [](const Number& spread) { return spread.adjoint(); });
    tape.clear();
    return r;
}

// ---- Superbucket by bumps (main.h:575-696): 1e-8 on the spot, 1e-5 on every spread of the risk view, recalibrating
inline SuperbucketResults dupireSuperbucketBump(const double spot, const double maxDt, const std::string& productId,
                                                const std::map<std::string, double>& notionals,
                                                const std::vector<double>& inclSpots, const double maxDs,
                                                const std::vector<Time>& inclTimes, const double maxDtVol,
                                                const std::vector<double>& strikes, const std::vector<Time>& mats,
                                                const double vol, const double jmpIntens, const double jmpAverage,
                                                const double jmpStd, const NumericalParam& num)
{
    cfdrv::useDevices(num);
    const Product<double>* product = getProduct<double>(productId);
    if (!product) throw std::runtime_error("dupireSuperbucketBump() : product not found");
    const cfdrv::Book book(product->payoffLabels(), notionals, "dupireSuperbucketBump()");
    const MertonIVS surface(spot, vol, jmpIntens, jmpAverage, jmpStd);
    RiskView<double> view(strikes, mats);
    // the book value under a model calibrated to the (possibly bumped) view, at a (possibly bumped) spot
    const auto calibrated = dupireCalib(inclSpots, maxDs, inclTimes, maxDtVol, spot, vol, jmpIntens, jmpAverage, jmpStd);
    auto bookValue = [&](const double s0, const std::vector<double>& gridSpots, const std::vector<Time>& gridTimes,
                         const matrix<double>& localVols) {
        const Dupire<double> model(s0, gridSpots, gridTimes, localVols, maxDt);
        return book.value(value(model, *product, num).values);
    };

    SuperbucketResults r;
    r.strikes = strikes;
    r.mats = mats;
    r.value = bookValue(spot, calibrated.spots, calibrated.times, calibrated.lVols);
    r.delta = (bookValue(spot + 1.0e-08, calibrated.spots, calibrated.times, calibrated.lVols) - r.value) * 1.0e+08;
    r.vega.resize(view.rows(), view.cols());
    for (size_t i = 0; i < view.rows(); ++i)
        for (size_t j = 0; j < view.cols(); ++j) {
            view.bump(i, j, 1.0e-05);
            const auto again = dupireCalib(surface, inclSpots, maxDs, inclTimes, maxDtVol, view);
            r.vega[i][j] = (bookValue(spot, again.spots, again.times, again.lVols) - r.value) * 1.0e+05;
            view.bump(i, j, -1.0e-05);
        }
    return r;
}
