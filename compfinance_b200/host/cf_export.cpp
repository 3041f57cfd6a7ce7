// cf_export.cpp -- flat C wrappers over the host API (cf_main.h), the counterpart of the
// reference's Excel wrappers (xlExport.cpp:83-1255: xPutBlackScholes, xPutDupire, xPutBarrier,
// xPutEuropean, xPutEuropeans, xValue, xAADrisk, xAADriskAggregate, xBumprisk ...).  Instead of
// FP12 / XLOPER12 they take plain arrays, so any FFI (ctypes, cgo, JNI ...) can bind them; errors
// follow the wrappers' convention (catch everything at the boundary, xlExport.cpp:585-588) and are
// reported as a non-zero return code + cfx_last_error().
// Built into compfinance_b200/lib/libcf_host.so, which links the CUDA engine libcf_b200.so.
#include "cf_main.h"

namespace {
thread_local std::string g_err;

template <class F>
int guarded(F&& f)
{
    try { f(); return 0; }
    catch (const std::exception& e) { g_err = e.what(); return 1; }
    catch (...) { g_err = "unknown exception"; return 1; }
}

// defaults of xl2num (xlExport.cpp:35-67)
NumericalParam mkNum(int parallel, int useSobol, int numPath, int seed1, int seed2)
{
    NumericalParam n;
    n.parallel = parallel != 0;
    n.useSobol = useSobol != 0;
    n.numPath = numPath;
    n.seed1 = seed1 > 0 ? seed1 : 1234;
    n.seed2 = seed2 > 0 ? seed2 : n.seed1 + 1;
    return n;
}

std::map<std::string, double> mkNotionals(const std::string& productId, const double* notionals)
{
    const auto* labels = getPayoffLabels(productId);
    if (!labels) throw std::runtime_error("product not found");
    std::map<std::string, double> m;
    for (size_t i = 0; i < labels->size(); ++i)
        if (notionals[i] != 0.0) m[(*labels)[i]] = notionals[i];
    return m;
}

int joinLabels(const std::vector<std::string>& v, char* out, int cap)
{
    std::string s;
    for (const auto& l : v) { s += l; s += '\n'; }
    if (int(s.size()) + 1 <= cap && out) std::memcpy(out, s.c_str(), s.size() + 1);
    return int(s.size()) + 1;
}
}  // namespace

extern "C" {

const char* cfx_last_error() { return g_err.c_str(); }

int cfx_init(int device)
{
    return guarded([&] { cfDropSessions(); cfCheck(cf_init(1, &device)); });
}

// A single-process multi-device context: every run through the entry points below is sharded over these devices.
int cfx_init_devices(int n, const int* devices)
{
    return guarded([&] { cfDropSessions(); cfCheck(cf_init(n, devices)); });
}

// Drops the resident sessions (clones, tapes, device plans) the entry points keep per (model, product, RNG).
void cfx_drop_sessions() { cfDropSessions(); }

void cfx_set_system_time(double t) { systemTime = t; }

int cfx_put_black_scholes(double spot, double vol, int spotMeasure, double rate, double div, const char* id)
{
    return guarded([&] { putBlackScholes(spot, vol, spotMeasure != 0, rate, div, id); });
}

int cfx_put_dupire(double spot, const double* spots, int nSpots, const double* times, int nTimes,
                   const double* vols /* [nSpots][nTimes] */, double maxDt, const char* id)
{
    return guarded([&] {
        matrix<double> v(nSpots, nTimes);
        std::copy(vols, vols + size_t(nSpots) * nTimes, v.begin());
        putDupire(spot, std::vector<double>(spots, spots + nSpots), std::vector<double>(times, times + nTimes), v, maxDt, id);
    });
}

int cfx_put_european(double strike, double exercise, double settlement, const char* id)
{
    return guarded([&] { putEuropean(strike, exercise, settlement <= 0 ? exercise : settlement, id); });
}

int cfx_put_barrier(double strike, double barrier, double maturity, double monitorFreq, double smooth, int callPut,
                    const char* id)
{
    return guarded([&] { putBarrier(strike, barrier, maturity, monitorFreq, smooth, callPut != 0, id); });
}

int cfx_put_contingent(double coupon, double maturity, double payFreq, double smooth, const char* id)
{
    return guarded([&] { putContingent(coupon, maturity, payFreq, smooth, id); });
}

int cfx_put_europeans(const double* maturities, const double* strikes, int n, const char* id)
{
    return guarded([&] {
        putEuropeans(std::vector<double>(maturities, maturities + n), std::vector<double>(strikes, strikes + n), id);
    });
}

// xPutDLM / xPutMultiStats / xPutBaskets / xPutAutocall (xlExport.cpp:186-543); assets are named a0, a1, ...
static std::vector<std::string> mkNames(int n)
{
    std::vector<std::string> v;
    for (int i = 0; i < n; ++i) v.push_back("a" + std::to_string(i));
    return v;
}
static matrix<double> mkMatrix(const double* data, int rows, int cols)
{
    matrix<double> m(rows, cols);
    if (rows * cols > 0) std::copy(data, data + size_t(rows) * cols, m.begin());
    return m;
}

int cfx_put_displaced(int nAssets, const double* spots, const double* atms, const double* skews, double discRate,
                      const double* repoSpreads, const double* divDates, int nDivs, const double* divs /* [nDivs][nAssets] */,
                      const double* correl /* [nAssets][nAssets] */, double lambda, const char* id)
{
    return guarded([&] {
        putDisplaced(mkNames(nAssets), std::vector<double>(spots, spots + nAssets), std::vector<double>(atms, atms + nAssets),
                     std::vector<double>(skews, skews + nAssets), discRate, std::vector<double>(repoSpreads, repoSpreads + nAssets),
                     std::vector<double>(divDates, divDates + nDivs), mkMatrix(divs, nDivs, nAssets),
                     mkMatrix(correl, nAssets, nAssets), lambda, id);
    });
}

int cfx_put_multistats(int nAssets, const double* fixDates, const double* fwdDates, int n, const char* id)
{
    return guarded([&] {
        putMultiStats(mkNames(nAssets), std::vector<double>(fixDates, fixDates + n), std::vector<double>(fwdDates, fwdDates + n), id);
    });
}

int cfx_put_baskets(int nAssets, const double* weights, double maturity, const double* strikes, int nStrikes, const char* id)
{
    return guarded([&] {
        putBaskets(mkNames(nAssets), std::vector<double>(weights, weights + nAssets), maturity,
                   std::vector<double>(strikes, strikes + nStrikes), id);
    });
}

int cfx_put_autocall(int nAssets, const double* refs, double maturity, int periods, double ko, double strike, double cpn,
                     double smooth, const char* id)
{
    return guarded([&] {
        putAutocall(mkNames(nAssets), std::vector<double>(refs, refs + nAssets), maturity, periods, ko, strike, cpn, smooth, id);
    });
}

int cfx_num_payoffs(const char* productId)
{
    const auto* l = getPayoffLabels(productId);
    return l ? int(l->size()) : -1;
}
int cfx_num_params(const char* modelId)
{
    const auto* m = getModel<double>(modelId);
    return m ? int(m->numParams()) : -1;
}
int cfx_payoff_labels(const char* productId, char* out, int cap)
{
    const auto* l = getPayoffLabels(productId);
    return l ? joinLabels(*l, out, cap) : -1;
}
int cfx_param_labels(const char* modelId, char* out, int cap)
{
    const auto* m = getModel<double>(modelId);
    return m ? joinLabels(m->parameterLabels(), out, cap) : -1;
}
int cfx_product_timeline(const char* productId, double* out, int cap)
{
    const auto* tl = getTimeline(productId);
    if (!tl) return -1;
    for (size_t i = 0; i < tl->size() && int(i) < cap; ++i) out[i] = (*tl)[i];
    return int(tl->size());
}

// xValue (xlExport.cpp:545-589) -> value() (main.h:80)
int cfx_value(const char* modelId, const char* productId, int useSobol, int seed1, int seed2, int numPath,
              int parallel, double* values)
{
    return guarded([&] {
        if (!numPath) throw std::runtime_error("numPath is zero");
        auto r = value(modelId, productId, mkNum(parallel, useSobol, numPath, seed1, seed2));
        std::copy(r.values.begin(), r.values.end(), values);
    });
}

// Per-path payoff matrix of mcSimul / mcParallelSimul (mcBase.h:267, 314): out[numPath][nPay]
int cfx_simul_paths(const char* modelId, const char* productId, int useSobol, int seed1, int seed2, int numPath,
                    int parallel, double* out)
{
    return guarded([&] {
        const Model<double>* mdl = getModel<double>(modelId);
        const Product<double>* prd = getProduct<double>(productId);
        if (!mdl || !prd) throw std::runtime_error("model / product not found");
        auto rng = cfdrv::makeRng(mkNum(parallel, useSobol, numPath, seed1, seed2));
        auto res = parallel ? mcParallelSimul(*prd, *mdl, *rng, numPath) : mcSimul(*prd, *mdl, *rng, numPath);
        const size_t nPay = prd->payoffLabels().size();
        for (size_t i = 0; i < res.size(); ++i) std::copy(res[i].begin(), res[i].end(), out + i * nPay);
    });
}

// xAADrisk (xlExport.cpp:644-697) -> AADriskOne (main.h:99); riskPayoffIdx < 0 = first payoff
int cfx_aad_risk_one(const char* modelId, const char* productId, int riskPayoffIdx, int useSobol, int seed1,
                     int seed2, int numPath, int parallel, double* payoffValues, double* riskPayoffValue, double* risks)
{
    return guarded([&] {
        const auto* labels = getPayoffLabels(productId);
        if (!labels) throw std::runtime_error("product not found");
        const std::string label = riskPayoffIdx >= 0 ? labels->at(size_t(riskPayoffIdx)) : std::string();
        auto r = AADriskOne(modelId, productId, mkNum(parallel, useSobol, numPath, seed1, seed2), label);
        std::copy(r.payoffValues.begin(), r.payoffValues.end(), payoffValues);
        *riskPayoffValue = r.riskPayoffValue;
        std::copy(r.risks.begin(), r.risks.end(), risks);
    });
}

// Per-path results of mcSimulAAD / mcParallelSimulAAD (mcBase.h:429, 566) with aggregator = payoff k
int cfx_simul_aad_paths(const char* modelId, const char* productId, int riskPayoffIdx, int useSobol, int seed1,
                        int seed2, int numPath, int parallel, double* payoffs, double* aggregated, double* risks)
{
    return guarded([&] {
        const Model<Number>* mdl = getModel<Number>(modelId);
        const Product<Number>* prd = getProduct<Number>(productId);
        if (!mdl || !prd) throw std::runtime_error("model / product not found");
        auto rng = cfdrv::makeRng(mkNum(parallel, useSobol, numPath, seed1, seed2));
        const size_t k = riskPayoffIdx < 0 ? 0 : size_t(riskPayoffIdx);
        auto agg = [k](const std::vector<Number>& v) { return v[k]; };
        auto res = parallel ? mcParallelSimulAAD(*prd, *mdl, *rng, numPath, agg) : mcSimulAAD(*prd, *mdl, *rng, numPath, agg);
        const size_t nPay = prd->payoffLabels().size();
        for (size_t i = 0; i < res.payoffs.size(); ++i) std::copy(res.payoffs[i].begin(), res.payoffs[i].end(), payoffs + i * nPay);
        std::copy(res.aggregated.begin(), res.aggregated.end(), aggregated);
        std::copy(res.risks.begin(), res.risks.end(), risks);
    });
}

// xAADriskAggregate (xlExport.cpp:699-772) -> AADriskAggregate (main.h:176); notionals[nPay], 0 = absent
int cfx_aad_risk_aggregate(const char* modelId, const char* productId, const double* notionals, int useSobol,
                           int seed1, int seed2, int numPath, int parallel, double* payoffValues,
                           double* riskPayoffValue, double* risks)
{
    return guarded([&] {
        auto r = AADriskAggregate(modelId, productId, mkNotionals(productId, notionals),
                                  mkNum(parallel, useSobol, numPath, seed1, seed2));
        std::copy(r.payoffValues.begin(), r.payoffValues.end(), payoffValues);
        *riskPayoffValue = r.riskPayoffValue;
        std::copy(r.risks.begin(), r.risks.end(), risks);
    });
}

// xAADriskMulti (xlExport.cpp:826-874) -> AADriskMulti (main.h:269): risks[nParam][nPay]
int cfx_aad_risk_multi(const char* modelId, const char* productId, int useSobol, int seed1, int seed2, int numPath,
                       int parallel, double* values, double* risks)
{
    return guarded([&] {
        auto r = AADriskMulti(modelId, productId, mkNum(parallel, useSobol, numPath, seed1, seed2));
        std::copy(r.values.begin(), r.values.end(), values);
        std::copy(r.risks.begin(), r.risks.end(), risks);
    });
}

// Per-path results of mcSimulAADMulti / mcParallelSimulAADMulti (mcBase.h:776, 859): payoffs [nPath][nPay] (the matrix the
// reference's result carries, mcBase.h:758-771), risks [nParam][nPay]
int cfx_simul_aad_multi_paths(const char* modelId, const char* productId, int useSobol, int seed1, int seed2, int numPath,
                              int parallel, double* payoffs, double* risks)
{
    return guarded([&] {
        const Model<Number>* mdl = getModel<Number>(modelId);
        const Product<Number>* prd = getProduct<Number>(productId);
        if (!mdl || !prd) throw std::runtime_error("model / product not found");
        auto rng = cfdrv::makeRng(mkNum(parallel, useSobol, numPath, seed1, seed2));
        auto res = parallel ? mcParallelSimulAADMulti(*prd, *mdl, *rng, numPath) : mcSimulAADMulti(*prd, *mdl, *rng, numPath);
        const size_t nPay = prd->payoffLabels().size();
        if (res.payoffs.size() != size_t(numPath)) throw std::runtime_error("mcSimulAADMulti: per-path payoffs not filled (too large)");
        for (size_t i = 0; i < res.payoffs.size(); ++i) std::copy(res.payoffs[i].begin(), res.payoffs[i].end(), payoffs + i * nPay);
        std::copy(res.risks.begin(), res.risks.end(), risks);
    });
}

// xBumprisk (xlExport.cpp:876-922) -> bumpRisk (main.h:316): risks[nParam][nPay]
int cfx_bump_risk(const char* modelId, const char* productId, int useSobol, int seed1, int seed2, int numPath,
                  int parallel, double* values, double* risks)
{
    return guarded([&] {
        auto r = bumpRisk(modelId, productId, mkNum(parallel, useSobol, numPath, seed1, seed2));
        std::copy(r.values.begin(), r.values.end(), values);
        std::copy(r.risks.begin(), r.risks.end(), risks);
    });
}

// dupireAADRisk (main.h:364): vega[nSpots][nTimes]
int cfx_dupire_aad_risk(const char* modelId, const char* productId, const double* notionals, int useSobol,
                        int seed1, int seed2, int numPath, int parallel, double* value_, double* delta, double* vega)
{
    return guarded([&] {
        auto r = dupireAADRisk(modelId, productId, mkNotionals(productId, notionals),
                               mkNum(parallel, useSobol, numPath, seed1, seed2));
        *value_ = r.value;
        *delta = r.delta;
        std::copy(r.vega.begin(), r.vega.end(), vega);
    });
}

// xDupireCalib (xlExport.cpp:924-963) -> dupireCalib (main.h:413): returns sizes, fills spots / times / lvols[nSpots][nTimes]
int cfx_dupire_calib(const double* inclSpots, int nInclSpots, double maxDs, const double* inclTimes, int nInclTimes,
                     double maxDt, double spot, double vol, double jmpIntens, double jmpAverage, double jmpStd, int* nSpots,
                     int* nTimes, double* spots, double* times, double* lvols, int cap)
{
    return guarded([&] {
        auto r = dupireCalib(std::vector<double>(inclSpots, inclSpots + nInclSpots), maxDs,
                             std::vector<double>(inclTimes, inclTimes + nInclTimes), maxDt, spot, vol, jmpIntens, jmpAverage, jmpStd);
        *nSpots = int(r.spots.size());
        *nTimes = int(r.times.size());
        if (int(r.spots.size() * r.times.size()) > cap) throw std::runtime_error("cfx_dupire_calib: cap too small");
        std::copy(r.spots.begin(), r.spots.end(), spots);
        std::copy(r.times.begin(), r.times.end(), times);
        std::copy(r.lVols.begin(), r.lVols.end(), lvols);
    });
}

// xDupireSuperbucket (xlExport.cpp:965-1104) -> dupireSuperbucket / dupireSuperbucketBump (main.h:453, 575): vega[nStrikes][nMats]
int cfx_dupire_superbucket(double spot, double maxDt, const char* productId, const double* notionals, const double* inclSpots,
                           int nInclSpots, double maxDs, const double* inclTimes, int nInclTimes, double maxDtVol,
                           const double* strikes, int nStrikes, const double* mats, int nMats, double vol, double jmpIntens,
                           double jmpAverage, double jmpStd, int useSobol, int seed1, int seed2, int numPath, int parallel,
                           int bump, double* value_, double* delta, double* vega)
{
    return guarded([&] {
        const auto nots = mkNotionals(productId, notionals);
        const std::vector<double> is(inclSpots, inclSpots + nInclSpots), it(inclTimes, inclTimes + nInclTimes);
        const std::vector<double> ks(strikes, strikes + nStrikes), ms(mats, mats + nMats);
        const auto num = mkNum(parallel, useSobol, numPath, seed1, seed2);
        auto r = bump ? dupireSuperbucketBump(spot, maxDt, productId, nots, is, maxDs, it, maxDtVol, ks, ms, vol, jmpIntens, jmpAverage, jmpStd, num)
                      : dupireSuperbucket(spot, maxDt, productId, nots, is, maxDs, it, maxDtVol, ks, ms, vol, jmpIntens, jmpAverage, jmpStd, num);
        *value_ = r.value;
        *delta = r.delta;
        std::copy(r.vega.begin(), r.vega.end(), vega);
    });
}

// Host-only inspection of the path-independent stage (no GPU work): the flattened device image of
// (model, product) after allocate + init.  Any output pointer may be null.  Returns sizes through dims:
// dims = {n_steps, n_events, n_knots, n_times (0 = no time map), adjoint size, is first sample today}
int cfx_describe(const char* modelId, const char* productId, int aad, int* dims, unsigned char* isEvent,
                 double* tabA, double* tabB, double* numeraires, double* fwdFactors, double* discounts,
                 int* col1, int* col2, double* w1, double* w2, double* productConsts /* strike, barrier, smooth */)
{
    return guarded([&] {
        CfDeviceSetup s;
        std::unique_ptr<Model<double>> md;
        std::unique_ptr<Model<Number>> mn;
        Sobol rng;
        if (aad) {
            const Model<Number>* mdl = getModel<Number>(modelId);
            const Product<Number>* prd = getProduct<Number>(productId);
            if (!mdl || !prd) throw std::runtime_error("model / product not found");
            mn = mdl->clone();
            mn->allocate(prd->timeline(), prd->defline());
            Number::tape->clear();
            mn->putParametersOnTape();
            mn->init(prd->timeline(), prd->defline());
            Number::tape->mark();
            cfBuildImages(*prd, *mn, rng, s);
        } else {
            const Model<double>* mdl = getModel<double>(modelId);
            const Product<double>* prd = getProduct<double>(productId);
            if (!mdl || !prd) throw std::runtime_error("model / product not found");
            md = mdl->clone();
            md->allocate(prd->timeline(), prd->defline());
            md->init(prd->timeline(), prd->defline());
            cfBuildImages(*prd, *md, rng, s);
        }
        const cf_model& p = s.mdl.pod;
        if (dims) {
            dims[0] = p.n_steps; dims[1] = p.n_events; dims[2] = p.n_knots; dims[3] = p.n_times;
            dims[4] = int(cf_table_adjoint_size(&s.mdl.pod, &s.prd.pod)); dims[5] = s.mdl.firstSampleIsToday ? 1 : 0;
        }
        auto cp = [](const auto& v, auto* out) { if (out) std::copy(v.begin(), v.end(), out); };
        cp(s.mdl.isEvent, isEvent); cp(s.mdl.tabA, tabA); cp(s.mdl.tabB, tabB);
        cp(s.mdl.numeraires, numeraires); cp(s.mdl.fwdFactors, fwdFactors); cp(s.mdl.discounts, discounts);
        if (p.n_times > 0) { cp(s.mdl.col1, col1); cp(s.mdl.col2, col2); cp(s.mdl.w1, w1); cp(s.mdl.w2, w2); }
        if (productConsts) { productConsts[0] = s.prd.pod.strike; productConsts[1] = s.prd.pod.barrier; productConsts[2] = s.prd.pod.smooth; }
        if (aad) Number::tape->clear();
    });
}

// A range of paths [firstPath, firstPath + nPaths) of (model, product) straight through the C ABI (cf_run_value /
// cf_run_aad): sums, not averages -- what one worker of the reference's parallel loops contributes (mcBase.h:346-395,
// 640-746).  For tests of arbitrary shard boundaries (odd antithetic starts) on every model.  tableAdjoints: dims[4] of
// cfx_describe doubles (AAD only); weights: the aggregate's payoff weights (AAD only).
int cfx_run_range(const char* modelId, const char* productId, int aad, int useSobol, int seed1, int seed2,
                  unsigned long long firstPath, unsigned long long nPaths, const double* weights, double* payoffSums,
                  double* aggSum, double* tableAdjoints)
{
    return guarded([&] {
        CfDeviceSetup s;
        auto rng = cfdrv::makeRng(mkNum(1, useSobol, int(nPaths), seed1, seed2));
        if (aad) {
            const Model<Number>* mdl = getModel<Number>(modelId);
            const Product<Number>* prd = getProduct<Number>(productId);
            if (!mdl || !prd) throw std::runtime_error("model / product not found");
            auto mn = mdl->clone();
            mn->allocate(prd->timeline(), prd->defline());
            Number::tape->clear();
            mn->putParametersOnTape();
            mn->init(prd->timeline(), prd->defline());
            Number::tape->mark();
            cfBuildImages(*prd, *mn, *rng, s);
            cfCheck(cf_run_aad(&s.mdl.pod, &s.prd.pod, &s.rng, firstPath, nPaths, weights, payoffSums, aggSum, tableAdjoints, nullptr, nullptr));
            Number::tape->clear();
        } else {
            const Model<double>* mdl = getModel<double>(modelId);
            const Product<double>* prd = getProduct<double>(productId);
            if (!mdl || !prd) throw std::runtime_error("model / product not found");
            auto md = mdl->clone();
            md->allocate(prd->timeline(), prd->defline());
            md->init(prd->timeline(), prd->defline());
            cfBuildImages(*prd, *md, *rng, s);
            cfCheck(cf_run_value(&s.mdl.pod, &s.prd.pod, &s.rng, firstPath, nPaths, payoffSums, nullptr));
        }
    });
}

// Sequential RNG interface (RNG::init / skipTo / nextU / nextG, mcBase.h:228-246) served by the device
int cfx_rng_sequence(int useSobol, int seed1, int seed2, int dim, unsigned skip, int n, int gaussian, double* out)
{
    return guarded([&] {
        auto rng = cfdrv::makeRng(mkNum(1, useSobol, n, seed1, seed2));
        rng->init(size_t(dim));
        if (skip) rng->skipTo(skip);
        std::vector<double> v(static_cast<size_t>(dim));
        for (int i = 0; i < n; ++i) {
            if (gaussian) rng->nextG(v); else rng->nextU(v);
            std::copy(v.begin(), v.end(), out + size_t(i) * dim);
        }
    });
}

}  // extern "C"

#include "cf_xl.h"
