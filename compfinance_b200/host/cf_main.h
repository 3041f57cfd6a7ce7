// cf_main.h -- the entry points of the reference's main.h, same names, arguments and result
// structs: value (:43-96), AADriskOne (:99-173), AADriskAggregate (:176-254), bumpRisk (:316-359),
// dupireAADRisk (:364-411).  The simulations behind them run on the CUDA engine; means over paths
// come from deterministic device reductions instead of std::accumulate over a per-path matrix.
#pragma once

#include "cf_calib.h"
#include "cf_rng.h"
#include "cf_store.h"

struct NumericalParam
{
    bool parallel;
    bool useSobol;
    int  numPath;
    int  seed1 = 12345;
    int  seed2 = 1234;
};

inline std::unique_ptr<RNG> cfMakeRng(const NumericalParam& num)
{
    if (num.useSobol) return std::make_unique<Sobol>();
    return std::make_unique<mrg32k3a>(num.seed1, num.seed2);
}

struct ValueResults
{
    std::vector<std::string> identifiers;
    std::vector<double>      values;
};

// Price product in model (main.h:43-77)
inline ValueResults value(const Model<double>& model, const Product<double>& product, const NumericalParam& num)
{
    auto rng = cfMakeRng(num);
    const auto sums = cfSimulSums(product, model, *rng, size_t(num.numPath));
    ValueResults results;
    results.identifiers = product.payoffLabels();
    results.values.resize(sums.size());
    for (size_t i = 0; i < sums.size(); ++i) results.values[i] = sums[i] / num.numPath;
    return results;
}

inline ValueResults value(const std::string& modelId, const std::string& productId, const NumericalParam& num)
{
    const Model<double>* model = getModel<double>(modelId);
    const Product<double>* product = getProduct<double>(productId);
    if (!model || !product) throw std::runtime_error("value() : Could not retrieve model and product");
    return value(*model, *product, num);
}

struct AADRiskResults
{
    std::vector<std::string> payoffIds;
    std::vector<double>      payoffValues;
    double                   riskPayoffValue;
    std::vector<std::string> paramIds;
    std::vector<double>      risks;
};

inline AADRiskResults cfAADrisk(const Model<Number>& model, const Product<Number>& product,
                                const std::vector<double>& weights, const NumericalParam& num)
{
    auto rng = cfMakeRng(num);
    const AADSums sums = cfSimulAADSums(product, model, *rng, size_t(num.numPath), weights);
    AADRiskResults results;
    results.payoffIds = product.payoffLabels();
    results.payoffValues.resize(sums.payoffSums.size());
    for (size_t i = 0; i < sums.payoffSums.size(); ++i) results.payoffValues[i] = sums.payoffSums[i] / num.numPath;
    results.riskPayoffValue = sums.aggSum / num.numPath;
    results.paramIds = model.parameterLabels();
    results.risks = sums.risks;
    return results;
}

// AAD risk, one payoff (main.h:99-173)
inline AADRiskResults AADriskOne(const std::string& modelId, const std::string& productId, const NumericalParam& num,
                                 const std::string& riskPayoff = "")
{
    const Model<Number>* model = getModel<Number>(modelId);
    const Product<Number>* product = getProduct<Number>(productId);
    if (!model || !product) throw std::runtime_error("AADrisk() : Could not retrieve model and product");
    const std::vector<std::string>& allPayoffs = product->payoffLabels();
    size_t riskPayoffIdx = 0;
    if (!riskPayoff.empty()) {
        auto it = std::find(allPayoffs.begin(), allPayoffs.end(), riskPayoff);
        if (it == allPayoffs.end()) throw std::runtime_error("AADriskOne() : payoff not found");
        riskPayoffIdx = size_t(std::distance(allPayoffs.begin(), it));
    }
    std::vector<double> weights(allPayoffs.size(), 0.0);
    weights[riskPayoffIdx] = 1.0;
    return cfAADrisk(*model, *product, weights, num);
}

// Notionals by payoff label -> weight vector in payoff order (main.h:196-207)
inline std::vector<double> cfNotionalWeights(const Product<Number>& product, const std::map<std::string, double>& notionals,
                                             const char* who)
{
    const std::vector<std::string>& allPayoffs = product.payoffLabels();
    std::vector<double> vnots(allPayoffs.size(), 0.0);
    for (const auto& notional : notionals) {
        auto it = std::find(allPayoffs.begin(), allPayoffs.end(), notional.first);
        if (it == allPayoffs.end()) throw std::runtime_error(std::string(who) + " : payoff not found");
        vnots[size_t(std::distance(allPayoffs.begin(), it))] = notional.second;
    }
    return vnots;
}

// AAD risk, aggregate portfolio (main.h:176-254)
inline AADRiskResults AADriskAggregate(const std::string& modelId, const std::string& productId,
                                       const std::map<std::string, double>& notionals, const NumericalParam& num)
{
    const Model<Number>* model = getModel<Number>(modelId);
    const Product<Number>* product = getProduct<Number>(productId);
    if (!model || !product) throw std::runtime_error("AADriskAggregate() : Could not retrieve model and product");
    return cfAADrisk(*model, *product, cfNotionalWeights(*product, notionals, "AADriskAggregate()"), num);
}

// Values and a matrix of risks, payoffs in columns and parameters in rows (main.h:259-265)
struct RiskReports
{
    std::vector<std::string> payoffs;
    std::vector<std::string> params;
    std::vector<double>      values;
    matrix<double>           risks;
};

// Itemized AAD risk, one per payoff (main.h:269-312)
inline RiskReports AADriskMulti(const std::string& modelId, const std::string& productId, const NumericalParam& num)
{
    const Model<Number>* model = getModel<Number>(modelId);
    const Product<Number>* product = getProduct<Number>(productId);
    if (!model || !product) throw std::runtime_error("AADrisk() : Could not retrieve model and product");
    RiskReports results;
    auto rng = cfMakeRng(num);
    AADMultiSums sums = cfSimulAADMultiSums(*product, *model, *rng, size_t(num.numPath));
    results.params = model->parameterLabels();
    results.payoffs = product->payoffLabels();
    results.risks = std::move(sums.risks);
    results.values.resize(sums.payoffSums.size());
    for (size_t i = 0; i < sums.payoffSums.size(); ++i) results.values[i] = sums.payoffSums[i] / num.numPath;
    return results;
}

// Bump risk, itemized (main.h:316-359): finite differences by re-running value()
inline RiskReports bumpRisk(const std::string& modelId, const std::string& productId, const NumericalParam& num)
{
    auto* orig = getModel<double>(modelId);
    const Product<double>* product = getProduct<double>(productId);
    if (!orig || !product) throw std::runtime_error("bumpRisk() : Could not retrieve model and product");
    RiskReports results;
    auto baseRes = value(*orig, *product, num);
    results.payoffs = baseRes.identifiers;
    results.values = baseRes.values;
    auto model = orig->clone();
    results.params = model->parameterLabels();
    const std::vector<double*> parameters = model->parameters();
    const size_t n = parameters.size(), m = results.payoffs.size();
    results.risks.resize(n, m);
    for (size_t i = 0; i < n; ++i) {
        *parameters[i] += 1.e-08;
        auto bumpRes = value(*model, *product, num);
        *parameters[i] -= 1.e-08;
        for (size_t j = 0; j < m; ++j) results.risks[i][j] = 1.0e+08 * (bumpRes.values[j] - baseRes.values[j]);
    }
    return results;
}

struct DupireRiskResults
{
    double         value;
    double         delta;
    matrix<double> vega;
};

// Dupire specific: price, delta and vega matrix to the local-vol surface (main.h:364-411)
inline DupireRiskResults dupireAADRisk(const std::string& modelId, const std::string& productId,
                                       const std::map<std::string, double>& notionals, const NumericalParam& num)
{
    const Model<Number>* model = getModel<Number>(modelId);
    if (!model) throw std::runtime_error("dupireAADRisk() : Model not found");
    const Dupire<Number>* dupire = dynamic_cast<const Dupire<Number>*>(model);
    if (!dupire) throw std::runtime_error("dupireAADRisk() : Model not a Dupire");
    const Product<Number>* product = getProduct<Number>(productId);
    if (!product) throw std::runtime_error("AADriskAggregate() : Could not retrieve model and product");
    // AADriskAggregate without its label vectors (1081 strings per call that this entry point drops, main.h:399-408)
    auto rng = cfMakeRng(num);
    const AADSums sums = cfSimulAADSums(*product, *model, *rng, size_t(num.numPath),
                                        cfNotionalWeights(*product, notionals, "AADriskAggregate()"));
    DupireRiskResults results;
    results.value = sums.aggSum / num.numPath;
    results.delta = sums.risks[0];
    results.vega.resize(dupire->spots().size(), dupire->times().size());
    std::copy(std::next(sums.risks.begin()), sums.risks.end(), results.vega.begin());
    return results;
}

// Superbucket (main.h:449-569): value, delta and vega to the implied vols of a risk view
struct SuperbucketResults
{
    double              value;
    double              delta;
    std::vector<double> strikes;
    std::vector<Time>   mats;
    matrix<double>      vega;
};

// Calibrate -> price and differentiate to the local vols on the GPU (dupireAADRisk) -> calibrate again on the
// host tape with a risk view -> seed the local vols with the microbucket -> sweep back to the implied-vol spreads.
inline SuperbucketResults dupireSuperbucket(const double spot, const double maxDt, const std::string& productId,
                                            const std::map<std::string, double>& notionals,
                                            const std::vector<double>& inclSpots, const double maxDs,
                                            const std::vector<Time>& inclTimes, const double maxDtVol,
                                            const std::vector<double>& strikes, const std::vector<Time>& mats, const double vol,
                                            const double jmpIntens, const double jmpAverage, const double jmpStd,
                                            const NumericalParam& num)
{
    SuperbucketResults results;
    Tape* tape = Number::tape;
    tape->rewind();
    auto params = dupireCalib(inclSpots, maxDs, inclTimes, maxDtVol, spot, vol, jmpIntens, jmpAverage, jmpStd);
    putDupire(spot, params.spots, params.times, params.lVols, maxDt, "superbucket");
    auto mdlDerivs = dupireAADRisk("superbucket", productId, notionals, num);
    results.value = mdlDerivs.value;
    results.delta = mdlDerivs.delta;
    const matrix<double>& microbucket = mdlDerivs.vega;

    tape->clear();
    MertonIVS ivs(spot, vol, jmpIntens, jmpAverage, jmpStd);
    RiskView<Number> riskView(strikes, mats);
    auto nParams = dupireCalib(ivs, inclSpots, maxDs, inclTimes, maxDtVol, riskView);
    matrix<Number>& nLvols = nParams.lVols;
    // Seeded by ASSIGNMENT, as the reference does (main.h:541-547): flat-extrapolated local vols are copies that
    // share their tape node with the edge of the calibrated range, so the last assignment wins on those nodes.
    for (size_t i = 0; i < microbucket.rows(); ++i)
        for (size_t j = 0; j < microbucket.cols(); ++j)
            if (nLvols[i][j].onTape()) nLvols[i][j].adjoint() = microbucket[i][j];
    if (tape->size() > 0) Number::propagateAdjoints(tape->end() - 1, tape->begin());
    results.strikes = strikes;
    results.mats = mats;
    results.vega.resize(riskView.rows(), riskView.cols());
    std::transform(riskView.begin(), riskView.end(), results.vega.begin(), [](const Number& n) { return n.adjoint(); });
    tape->clear();
    return results;
}

// Superbucket by bumps (main.h:575-696): 1e-8 on the spot, 1e-5 on every spread of the risk view, recalibrating each time
inline SuperbucketResults dupireSuperbucketBump(const double spot, const double maxDt, const std::string& productId,
                                                const std::map<std::string, double>& notionals,
                                                const std::vector<double>& inclSpots, const double maxDs,
                                                const std::vector<Time>& inclTimes, const double maxDtVol,
                                                const std::vector<double>& strikes, const std::vector<Time>& mats,
                                                const double vol, const double jmpIntens, const double jmpAverage,
                                                const double jmpStd, const NumericalParam& num)
{
    SuperbucketResults results;
    auto params = dupireCalib(inclSpots, maxDs, inclTimes, maxDtVol, spot, vol, jmpIntens, jmpAverage, jmpStd);
    Dupire<double> model(spot, params.spots, params.times, params.lVols, maxDt);
    const Product<double>* product = getProduct<double>(productId);
    if (!product) throw std::runtime_error("dupireSuperbucketBump() : product not found");
    auto baseVals = value(model, *product, num);
    const std::vector<std::string>& allPayoffs = baseVals.identifiers;
    std::vector<double> vnots(allPayoffs.size(), 0.0);
    for (const auto& notional : notionals) {
        auto it = std::find(allPayoffs.begin(), allPayoffs.end(), notional.first);
        if (it == allPayoffs.end()) throw std::runtime_error("dupireSuperbucketBump() : payoff not found");
        vnots[size_t(std::distance(allPayoffs.begin(), it))] = notional.second;
    }
    auto book = [&](const ValueResults& v) { return std::inner_product(vnots.begin(), vnots.end(), v.values.begin(), 0.0); };
    results.value = book(baseVals);
    MertonIVS ivs(spot, vol, jmpIntens, jmpAverage, jmpStd);
    RiskView<double> riskView(strikes, mats);
    Dupire<double> bumpedSpot(spot + 1.0e-08, params.spots, params.times, params.lVols, maxDt);
    results.delta = (book(value(bumpedSpot, *product, num)) - results.value) * 1.0e+08;
    const size_t n = riskView.rows(), m = riskView.cols();
    results.vega.resize(n, m);
    for (size_t i = 0; i < n; ++i)
        for (size_t j = 0; j < m; ++j) {
            riskView.bump(i, j, 1.0e-05);
            auto bumpedCalib = dupireCalib(ivs, inclSpots, maxDs, inclTimes, maxDtVol, riskView);
            Dupire<double> bumpedModel(spot, bumpedCalib.spots, bumpedCalib.times, bumpedCalib.lVols, maxDt);
            results.vega[i][j] = (book(value(bumpedModel, *product, num)) - results.value) * 1.0e+05;
            riskView.bump(i, j, -1.0e-05);
        }
    results.strikes = strikes;
    results.mats = mats;
    return results;
}
