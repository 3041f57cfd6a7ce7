// oracle/ref_driver.cpp -- C-ABI shim around the UNMODIFIED reference (asavine/CompFinance),
// compiled by oracle/build_ref.py into oracle/_ref/libcfref.so.
//
// TEST INFRASTRUCTURE ONLY: used by tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs as the checker / CPU baseline.  Never linked or loaded
// by the product (compfinance_b200/).
//
// Everything numerical below is executed by the reference's own code: main.h entry points
// (value main.h:80, AADriskOne main.h:99, AADriskAggregate main.h:176, AADriskMulti main.h:269,
// dupireAADRisk main.h:364, dupireCalib main.h:414, dupireSuperbucket main.h:453), the store
// (store.h:38-250), the RNG classes (sobol.h:31, mrg32k3a.h:23) and invNormalCdf
// (gaussians.h:47).  This file only moves flat arrays in and out.

#include "main.h"

#include <cstring>
#include <thread>

namespace {
thread_local std::string g_err;

template <class F>
int guarded(F&& f)
{
    try { f(); return 0; }
    catch (const std::exception& e) { g_err = e.what(); return 1; }
    catch (...) { g_err = "unknown exception"; return 1; }
}

NumericalParam mkNum(int parallel, int useSobol, int numPath, int seed1, int seed2)
{
    NumericalParam n;
    n.parallel = parallel != 0;
    n.useSobol = useSobol != 0;
    n.numPath = numPath;
    n.seed1 = seed1;
    n.seed2 = seed2;
    return n;
}

std::unique_ptr<RNG> mkRng(int useSobol, int seed1, int seed2)
{
    if (useSobol) return std::make_unique<Sobol>();
    return std::make_unique<mrg32k3a>(seed1, seed2);
}

std::map<std::string, double> mkNotionals(const Product<double>* prd, const double* notionals)
{
    std::map<std::string, double> m;
    const auto& labels = prd->payoffLabels();
    for (size_t i = 0; i < labels.size(); ++i)
        if (notionals[i] != 0.0) m[labels[i]] = notionals[i];
    return m;
}

matrix<double> mkMatrix(const double* data, int rows, int cols)
{
    matrix<double> m(rows, cols);
    std::copy(data, data + size_t(rows) * cols, m.begin());
    return m;
}

std::vector<std::string> mkNames(const char* prefix, int n)
{
    std::vector<std::string> v(n);
    for (int i = 0; i < n; ++i) v[i] = std::string(prefix) + std::to_string(i);
    return v;
}
}  // namespace

extern "C" {

const char* ref_last_error() { return g_err.c_str(); }

// Thread pool: reference default is hardware_concurrency()-1 workers (xlExport.cpp:1605).
int ref_start_pool(int nThread)
{
    return guarded([&] {
        ThreadPool* pool = ThreadPool::getInstance();
        pool->stop();
        if (nThread < 0) nThread = int(std::thread::hardware_concurrency()) - 1;
        if (nThread > 0) pool->start(nThread);
    });
}
int ref_stop_pool() { return guarded([] { ThreadPool::getInstance()->stop(); }); }
int ref_pool_threads() { return int(ThreadPool::getInstance()->numThreads()); }
void ref_set_system_time(double t) { systemTime = t; }

// ---- RNG streams ---------------------------------------------------------------------------
// Uniform / Gaussian vectors for paths [first, first + n); out is [n][dim] row-major.
int ref_rng_draw(int useSobol, int seed1, int seed2, int dim, unsigned first, int n,
                 int gaussian, double* out)
{
    return guarded([&] {
        auto rng = mkRng(useSobol, seed1, seed2);
        rng->init(dim);
        if (first) rng->skipTo(first);
        std::vector<double> v(dim);
        for (int i = 0; i < n; ++i) {
            if (gaussian) rng->nextG(v); else rng->nextU(v);
            std::memcpy(out + size_t(i) * dim, v.data(), sizeof(double) * dim);
        }
    });
}

void ref_inv_normal(const double* p, double* out, int n)
{
    for (int i = 0; i < n; ++i) out[i] = invNormalCdf(p[i]);
}

unsigned ref_sobol_dirnum(int bit, int dim) { return getjkDir()[bit][dim]; }

// ---- store ---------------------------------------------------------------------------------
int ref_put_bs(double spot, double vol, int spotMeasure, double rate, double div, const char* id)
{
    return guarded([&] { putBlackScholes(spot, vol, spotMeasure != 0, rate, div, id); });
}

int ref_put_dupire(double spot, const double* spots, int nSpots, const double* times, int nTimes,
                   const double* vols /*[nSpots][nTimes]*/, double maxDt, const char* id)
{
    return guarded([&] {
        putDupire(spot, std::vector<double>(spots, spots + nSpots),
                  std::vector<double>(times, times + nTimes), mkMatrix(vols, nSpots, nTimes), maxDt, id);
    });
}

int ref_put_displaced(int nAssets, const double* spots, const double* atms, const double* skews,
                      double discRate, const double* repoSpreads, const double* divDates, int nDivs,
                      const double* divs /*[nDivs][nAssets]*/, const double* correl /*[nA][nA]*/,
                      double lambda, const char* id)
{
    return guarded([&] {
        putDisplaced(mkNames("a", nAssets), std::vector<double>(spots, spots + nAssets),
                     std::vector<double>(atms, atms + nAssets), std::vector<double>(skews, skews + nAssets),
                     discRate, std::vector<double>(repoSpreads, repoSpreads + nAssets),
                     std::vector<double>(divDates, divDates + nDivs), mkMatrix(divs, nDivs, nAssets),
                     mkMatrix(correl, nAssets, nAssets), lambda, id);
    });
}

int ref_put_european(double strike, double exercise, double settlement, const char* id)
{
    return guarded([&] { putEuropean(strike, exercise, settlement, id); });
}

int ref_put_barrier(double strike, double barrier, double maturity, double monitorFreq, double smooth,
                    int callPut, const char* id)
{
    return guarded([&] { putBarrier(strike, barrier, maturity, monitorFreq, smooth, callPut != 0, id); });
}

int ref_put_contingent(double coupon, double maturity, double payFreq, double smooth, const char* id)
{
    return guarded([&] { putContingent(coupon, maturity, payFreq, smooth, id); });
}

int ref_put_europeans(const double* maturities, const double* strikes, int n, const char* id)
{
    return guarded([&] {
        putEuropeans(std::vector<double>(maturities, maturities + n),
                     std::vector<double>(strikes, strikes + n), id);
    });
}

int ref_put_multistats(int nAssets, const double* fixDates, const double* fwdDates, int n, const char* id)
{
    return guarded([&] {
        putMultiStats(mkNames("a", nAssets), std::vector<double>(fixDates, fixDates + n),
                      std::vector<double>(fwdDates, fwdDates + n), id);
    });
}

int ref_put_baskets(int nAssets, const double* weights, double maturity, const double* strikes,
                    int nStrikes, const char* id)
{
    return guarded([&] {
        putBaskets(mkNames("a", nAssets), std::vector<double>(weights, weights + nAssets), maturity,
                   std::vector<double>(strikes, strikes + nStrikes), id);
    });
}

int ref_put_autocall(int nAssets, const double* refs, double maturity, int periods, double ko,
                     double strike, double cpn, double smooth, const char* id)
{
    return guarded([&] {
        putAutocall(mkNames("a", nAssets), std::vector<double>(refs, refs + nAssets), maturity, periods,
                    ko, strike, cpn, smooth, id);
    });
}

int ref_num_payoffs(const char* productId)
{
    const auto* p = getProduct<double>(productId);
    return p ? int(p->payoffLabels().size()) : -1;
}
int ref_num_params(const char* modelId)
{
    const auto* m = getModel<double>(modelId);
    return m ? int(m->numParams()) : -1;
}
int ref_product_timeline(const char* productId, double* out, int cap)
{
    const auto* p = getProduct<double>(productId);
    if (!p) return -1;
    const auto& tl = p->timeline();
    for (size_t i = 0; i < tl.size() && int(i) < cap; ++i) out[i] = tl[i];
    return int(tl.size());
}
// labels joined by '\n'
int ref_labels(const char* id, int what /*0 payoffs, 1 params*/, char* out, int cap)
{
    const std::vector<std::string>* v = nullptr;
    if (what == 0) { const auto* p = getProduct<double>(id); if (p) v = &p->payoffLabels(); }
    else { const auto* m = getModel<double>(id); if (m) v = &m->parameterLabels(); }
    if (!v) return -1;
    std::string s;
    for (const auto& l : *v) { s += l; s += '\n'; }
    if (int(s.size()) + 1 > cap) return int(s.size()) + 1;
    std::memcpy(out, s.c_str(), s.size() + 1);
    return int(s.size()) + 1;
}

// ---- entry points --------------------------------------------------------------------------
// value(): main.h:80.  values[nPay]
int ref_value(const char* modelId, const char* productId, int parallel, int useSobol, int numPath,
              int seed1, int seed2, double* values)
{
    return guarded([&] {
        auto r = value(modelId, productId, mkNum(parallel, useSobol, numPath, seed1, seed2));
        std::copy(r.values.begin(), r.values.end(), values);
    });
}

// Per-path payoffs straight from mcSimul / mcParallelSimul (mcBase.h:267, 314). out[numPath][nPay]
int ref_simul_paths(const char* modelId, const char* productId, int parallel, int useSobol, int numPath,
                    int seed1, int seed2, double* out)
{
    return guarded([&] {
        const Model<double>* mdl = getModel<double>(modelId);
        const Product<double>* prd = getProduct<double>(productId);
        if (!mdl || !prd) throw std::runtime_error("ref_simul_paths: model/product not found");
        auto rng = mkRng(useSobol, seed1, seed2);
        auto res = parallel ? mcParallelSimul(*prd, *mdl, *rng, numPath) : mcSimul(*prd, *mdl, *rng, numPath);
        const size_t nPay = prd->payoffLabels().size();
        for (size_t i = 0; i < res.size(); ++i) std::copy(res[i].begin(), res[i].end(), out + i * nPay);
    });
}

// AADriskOne: main.h:99.  riskPayoffIdx < 0 = default (first payoff)
int ref_aad_risk_one(const char* modelId, const char* productId, int riskPayoffIdx, int parallel,
                     int useSobol, int numPath, int seed1, int seed2, double* payoffValues,
                     double* riskPayoffValue, double* risks)
{
    return guarded([&] {
        const Product<double>* prd = getProduct<double>(productId);
        if (!prd) throw std::runtime_error("ref_aad_risk_one: product not found");
        std::string label = riskPayoffIdx >= 0 ? prd->payoffLabels().at(riskPayoffIdx) : std::string();
        auto r = AADriskOne(modelId, productId, mkNum(parallel, useSobol, numPath, seed1, seed2), label);
        std::copy(r.payoffValues.begin(), r.payoffValues.end(), payoffValues);
        *riskPayoffValue = r.riskPayoffValue;
        std::copy(r.risks.begin(), r.risks.end(), risks);
    });
}

// AADriskAggregate: main.h:176.  notionals[nPay] (0 = absent from the map)
int ref_aad_risk_aggregate(const char* modelId, const char* productId, const double* notionals,
                           int parallel, int useSobol, int numPath, int seed1, int seed2,
                           double* payoffValues, double* riskPayoffValue, double* risks)
{
    return guarded([&] {
        const Product<double>* prd = getProduct<double>(productId);
        if (!prd) throw std::runtime_error("ref_aad_risk_aggregate: product not found");
        auto r = AADriskAggregate(modelId, productId, mkNotionals(prd, notionals),
                                  mkNum(parallel, useSobol, numPath, seed1, seed2));
        std::copy(r.payoffValues.begin(), r.payoffValues.end(), payoffValues);
        *riskPayoffValue = r.riskPayoffValue;
        std::copy(r.risks.begin(), r.risks.end(), risks);
    });
}

// AADriskMulti: main.h:269.  values[nPay], risks[nParam][nPay]
int ref_aad_risk_multi(const char* modelId, const char* productId, int parallel, int useSobol,
                       int numPath, int seed1, int seed2, double* values, double* risks)
{
    return guarded([&] {
        auto r = AADriskMulti(modelId, productId, mkNum(parallel, useSobol, numPath, seed1, seed2));
        std::copy(r.values.begin(), r.values.end(), values);
        std::copy(r.risks.begin(), r.risks.end(), risks);
    });
}

// bumpRisk: main.h:316
int ref_bump_risk(const char* modelId, const char* productId, int parallel, int useSobol, int numPath,
                  int seed1, int seed2, double* values, double* risks)
{
    return guarded([&] {
        auto r = bumpRisk(modelId, productId, mkNum(parallel, useSobol, numPath, seed1, seed2));
        std::copy(r.values.begin(), r.values.end(), values);
        std::copy(r.risks.begin(), r.risks.end(), risks);
    });
}

// dupireAADRisk: main.h:364.  vega[nSpots][nTimes]
int ref_dupire_aad_risk(const char* modelId, const char* productId, const double* notionals,
                        int parallel, int useSobol, int numPath, int seed1, int seed2, double* value_,
                        double* delta, double* vega)
{
    return guarded([&] {
        const Product<double>* prd = getProduct<double>(productId);
        if (!prd) throw std::runtime_error("ref_dupire_aad_risk: product not found");
        auto r = dupireAADRisk(modelId, productId, mkNotionals(prd, notionals),
                               mkNum(parallel, useSobol, numPath, seed1, seed2));
        *value_ = r.value;
        *delta = r.delta;
        std::copy(r.vega.begin(), r.vega.end(), vega);
    });
}

// dupireCalib: main.h:414.  Returns sizes through nSpots/nTimes; arrays sized by caps.
int ref_dupire_calib(const double* inclSpots, int nInclSpots, double maxDs, const double* inclTimes,
                     int nInclTimes, double maxDt, double spot, double vol, double jmpIntens,
                     double jmpAverage, double jmpStd, int* nSpots, int* nTimes, double* spots,
                     double* times, double* lvols, int cap)
{
    return guarded([&] {
        auto r = dupireCalib(std::vector<double>(inclSpots, inclSpots + nInclSpots), maxDs,
                             std::vector<double>(inclTimes, inclTimes + nInclTimes), maxDt, spot, vol,
                             jmpIntens, jmpAverage, jmpStd);
        *nSpots = int(r.spots.size());
        *nTimes = int(r.times.size());
        if (int(r.spots.size() * r.times.size()) > cap || int(r.spots.size()) > cap || int(r.times.size()) > cap)
            throw std::runtime_error("ref_dupire_calib: cap too small");
        std::copy(r.spots.begin(), r.spots.end(), spots);
        std::copy(r.times.begin(), r.times.end(), times);
        std::copy(r.lVols.begin(), r.lVols.end(), lvols);
    });
}

// dupireSuperbucket: main.h:453.  vega[nStrikes][nMats]
int ref_dupire_superbucket(double spot, double maxDt, const char* productId, const double* notionals,
                           const double* inclSpots, int nInclSpots, double maxDs, const double* inclTimes,
                           int nInclTimes, double maxDtVol, const double* strikes, int nStrikes,
                           const double* mats, int nMats, double vol, double jmpIntens, double jmpAverage,
                           double jmpStd, int parallel, int useSobol, int numPath, int seed1, int seed2,
                           double* value_, double* delta, double* vega)
{
    return guarded([&] {
        const Product<double>* prd = getProduct<double>(productId);
        if (!prd) throw std::runtime_error("ref_dupire_superbucket: product not found");
        auto r = dupireSuperbucket(spot, maxDt, productId, mkNotionals(prd, notionals),
                                   std::vector<double>(inclSpots, inclSpots + nInclSpots), maxDs,
                                   std::vector<double>(inclTimes, inclTimes + nInclTimes), maxDtVol,
                                   std::vector<double>(strikes, strikes + nStrikes),
                                   std::vector<double>(mats, mats + nMats), vol, jmpIntens, jmpAverage,
                                   jmpStd, mkNum(parallel, useSobol, numPath, seed1, seed2));
        *value_ = r.value;
        *delta = r.delta;
        std::copy(r.vega.begin(), r.vega.end(), vega);
    });
}

// dupireSuperbucketBump: main.h:575 (the reference's own finite-difference driver)
int ref_dupire_superbucket_bump(double spot, double maxDt, const char* productId, const double* notionals,
                                const double* inclSpots, int nInclSpots, double maxDs, const double* inclTimes,
                                int nInclTimes, double maxDtVol, const double* strikes, int nStrikes,
                                const double* mats, int nMats, double vol, double jmpIntens, double jmpAverage,
                                double jmpStd, int parallel, int useSobol, int numPath, int seed1, int seed2,
                                double* value_, double* delta, double* vega)
{
    return guarded([&] {
        const Product<double>* prd = getProduct<double>(productId);
        if (!prd) throw std::runtime_error("ref_dupire_superbucket_bump: product not found");
        auto r = dupireSuperbucketBump(spot, maxDt, productId, mkNotionals(prd, notionals),
                                       std::vector<double>(inclSpots, inclSpots + nInclSpots), maxDs,
                                       std::vector<double>(inclTimes, inclTimes + nInclTimes), maxDtVol,
                                       std::vector<double>(strikes, strikes + nStrikes),
                                       std::vector<double>(mats, mats + nMats), vol, jmpIntens, jmpAverage,
                                       jmpStd, mkNum(parallel, useSobol, numPath, seed1, seed2));
        *value_ = r.value;
        *delta = r.delta;
        std::copy(r.vega.begin(), r.vega.end(), vega);
    });
}

}  // extern "C"
