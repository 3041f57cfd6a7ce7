// cf_store.h -- the object store of the reference (store.h:30-282): two global maps holding a
// <double> and a <Number> twin of every model / product, addressed by name.
#pragma once

#include <unordered_map>

#include "cf_models.h"
#include "cf_models_multi.h"
#include "cf_products.h"
#include "cf_products_multi.h"

using ModelStore = std::unordered_map<std::string, std::pair<std::unique_ptr<Model<double>>, std::unique_ptr<Model<Number>>>>;
using ProductStore = std::unordered_map<std::string, std::pair<std::unique_ptr<Product<double>>, std::unique_ptr<Product<Number>>>>;

inline ModelStore modelStore;
inline ProductStore productStore;
// Every put gives the entry a new serial number: what the entry points key their resident device plans on
// (cf_base.h: CfSessionKey), so that overwriting a name invalidates the plans built from the old object.
inline std::unordered_map<std::string, uint64_t> modelSerial, productSerial;
inline uint64_t storeSerialCounter = 0;
inline uint64_t cfModelSerial(const std::string& store) { auto it = modelSerial.find(store); return it == modelSerial.end() ? 0 : it->second; }
inline uint64_t cfProductSerial(const std::string& store) { auto it = productSerial.find(store); return it == productSerial.end() ? 0 : it->second; }

template <template <class> class M, class... Args>
inline void cfPutModel(const std::string& store, const Args&... args)
{
    modelStore[store] = std::make_pair(std::unique_ptr<Model<double>>(new M<double>(args...)),
                                       std::unique_ptr<Model<Number>>(new M<Number>(args...)));
    modelSerial[store] = ++storeSerialCounter;
}
template <template <class> class P, class... Args>
inline void cfPutProduct(const std::string& store, const Args&... args)
{
    productStore[store] = std::make_pair(std::unique_ptr<Product<double>>(new P<double>(args...)),
                                         std::unique_ptr<Product<Number>>(new P<Number>(args...)));
    productSerial[store] = ++storeSerialCounter;
}

inline void putBlackScholes(const double spot, const double vol, const bool qSpot, const double rate, const double div,
                            const std::string& store)
{
    cfPutModel<BlackScholes>(store, spot, vol, qSpot, rate, div);
}

inline void putDupire(const double spot, const std::vector<double>& spots, const std::vector<Time>& times,
                      const matrix<double>& vols /* spot major */, const double maxDt, const std::string& store)
{
    cfPutModel<Dupire>(store, spot, spots, times, vols, maxDt);
}

// store.h:75-97
inline void putDisplaced(const std::vector<std::string>& assets, const std::vector<double>& spots, const std::vector<double>& atms,
                         const std::vector<double>& skews, const double& discRate, const std::vector<double>& repoSpreads,
                         const std::vector<Time>& divDates, const matrix<double>& divs, const matrix<double>& correl,
                         const double& lambda, const std::string& store)
{
    cfPutModel<MultiDisplaced>(store, assets, discRate, repoSpreads, spots, divDates, divs, atms, skews, correl, lambda);
}

template <class T> const Model<T>* getModel(const std::string& store);
template <> inline const Model<double>* getModel(const std::string& store)
{
    auto it = modelStore.find(store);
    return it == modelStore.end() ? nullptr : it->second.first.get();
}
template <> inline const Model<Number>* getModel(const std::string& store)
{
    auto it = modelStore.find(store);
    return it == modelStore.end() ? nullptr : it->second.second.get();
}

inline std::pair<const std::vector<std::string>*, const std::vector<double*>*> getModelParameters(const std::string& store)
{
    auto it = modelStore.find(store);
    if (it == modelStore.end()) return std::make_pair(nullptr, nullptr);
    auto* mdl = it->second.first.get();
    return std::make_pair(&mdl->parameterLabels(), &mdl->parameters());
}

inline void putEuropean(const double strike, const Time exerciseDate, const Time settlementDate, const std::string& store)
{
    cfPutProduct<European>(store, strike, exerciseDate, settlementDate);
}

inline void putBarrier(const double strike, const double barrier, const Time maturity, const double monitorFreq,
                       const double smooth, const bool callPut /* false: call, true: put */, const std::string& store)
{
    const double smoothFactor = smooth <= 0 ? EPS : smooth;      // store.h:161
    cfPutProduct<UOC>(store, strike, barrier, maturity, monitorFreq, smoothFactor, callPut);
}

// store.h:165-184 (note the argument order: coupon first)
inline void putContingent(const double coupon, const Time maturity, const double payFreq, const double smooth,
                          const std::string& store)
{
    const double smoothFactor = smooth <= 0 ? 0.0 : smooth;
    cfPutProduct<ContingentBond>(store, maturity, coupon, payFreq, smoothFactor);
}

inline void putEuropeans(const std::vector<Time>& maturities /* increasing */, const std::vector<double>& strikes,
                         const std::string& store)
{
    std::map<Time, std::vector<double>> options;
    for (size_t i = 0; i < maturities.size(); ++i) options[maturities[i]].push_back(strikes[i]);
    cfPutProduct<Europeans>(store, options);
}

// store.h:207-255
inline void putMultiStats(const std::vector<std::string>& assets, const std::vector<Time>& fixDates /* increasing */,
                          const std::vector<Time>& fwdDates /* on or after the fixings */, const std::string& store)
{
    cfPutProduct<MultiStats>(store, assets, fixDates, fwdDates);
}
inline void putBaskets(const std::vector<std::string>& assets, const std::vector<double>& weights, const Time maturity,
                       const std::vector<double> strikes, const std::string& store)
{
    cfPutProduct<Baskets>(store, assets, weights, maturity, strikes);
}
inline void putAutocall(const std::vector<std::string>& assets, const std::vector<double>& refs, const Time maturity,
                        const int periods, const double ko, const double strike, const double cpn, const double smooth,
                        const std::string& store)
{
    cfPutProduct<Autocall>(store, assets, refs, maturity, periods, ko, strike, cpn, smooth);
}

template <class T> const Product<T>* getProduct(const std::string& store);
template <> inline const Product<double>* getProduct(const std::string& store)
{
    auto it = productStore.find(store);
    return it == productStore.end() ? nullptr : it->second.first.get();
}
template <> inline const Product<Number>* getProduct(const std::string& store)
{
    auto it = productStore.find(store);
    return it == productStore.end() ? nullptr : it->second.second.get();
}

inline const std::vector<std::string>* getPayoffLabels(const std::string& store)
{
    auto it = productStore.find(store);
    return it == productStore.end() ? nullptr : &it->second.first->payoffLabels();
}
inline const std::vector<Time>* getTimeline(const std::string& store)
{
    auto it = productStore.find(store);
    return it == productStore.end() ? nullptr : &it->second.first->timeline();
}
