// cf_base.h -- host mirror of the reference's simulation abstractions (mcBase.h) with the six
// template algorithms re-implemented on top of the CUDA engine (include/cf_b200.h).
//
// Kept from the reference, same names and meaning (mcBase.h:45-246): Time / systemTime, SampleDef,
// Sample<T>, Scenario<T>, allocatePath, initializePath, Product<T>, Model<T>, RNG, and the free
// functions mcSimul (:267), mcParallelSimul (:314), mcSimulAAD (:429), mcParallelSimulAAD (:566)
// with their result structs.  What changes: the path loops run on the GPU.  A concrete model or
// product takes part by describing itself to the engine through deviceImage(); the host keeps
// everything path-independent (timelines, init() tables, the chain rule from tables to parameters
// on a small host tape = the reference's propagateMarkToStart, mcBase.h:518 / 721-733).
//
// There is no CPU path: generatePath()/payoffs() of the built-in classes are not evaluated on the
// host, and a model/product without a device image makes the algorithms throw runtime_error
// (the reference's error convention, mcBase.h:273).
#pragma once

#include <chrono>
#include <cstdio>
#include <cstdlib>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <iomanip>
#include <map>
#include <memory>
#include <numeric>
#include <sstream>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <unordered_map>
#include <vector>

#include "../../include/cf_b200.h"
#include "cf_aad.h"
#include "cf_matrix.h"

using Time = double;
inline Time systemTime = 0.0;     // mcBase.cpp:22

// ---------------------------------------------------------------------------------------------
// Scenarios (mcBase.h:45-117)
// ---------------------------------------------------------------------------------------------
struct SampleDef
{
    bool numeraire = true;
    struct RateDef
    {
        Time start, end;
        std::string curve;
        RateDef(const Time s, const Time e, const std::string& c) : start(s), end(e), curve(c) {}
    };
    std::vector<Time>               discountMats;
    std::vector<RateDef>            liborDefs;
    std::vector<std::vector<Time>>  forwardMats;   // forwardMats[a] = maturities for asset a
};

template <class T>
struct Sample
{
    T                           numeraire;
    std::vector<T>              discounts;
    std::vector<T>              libors;
    std::vector<std::vector<T>> forwards;

    void allocate(const SampleDef& data)
    {
        discounts.resize(data.discountMats.size());
        libors.resize(data.liborDefs.size());
        forwards.resize(data.forwardMats.size());
        for (size_t a = 0; a < forwards.size(); ++a) forwards[a].resize(data.forwardMats[a].size());
    }
    void initialize()
    {
        numeraire = T(1.0);
        std::fill(discounts.begin(), discounts.end(), T(1.0));
        std::fill(libors.begin(), libors.end(), T(0.0));
        for (auto& f : forwards) std::fill(f.begin(), f.end(), T(100.0));
    }
};

template <class T> using Scenario = std::vector<Sample<T>>;

template <class T>
inline void allocatePath(const std::vector<SampleDef>& defline, Scenario<T>& path)
{
    path.resize(defline.size());
    for (size_t i = 0; i < defline.size(); ++i) path[i].allocate(defline[i]);
}
template <class T>
inline void initializePath(Scenario<T>& path) { for (auto& s : path) s.initialize(); }

// ---------------------------------------------------------------------------------------------
// Device images: flat, self-owning descriptions handed to the C ABI
// ---------------------------------------------------------------------------------------------
struct ModelImage
{
    cf_model pod{};
    std::vector<uint8_t> isEvent;
    std::vector<double>  tabA, tabB, numeraires, fwdFactors, discounts, libors;
    std::vector<int32_t> col1, col2;
    std::vector<double>  w1, w2;
    // value of forwards[0][0] on the first sample when that sample is today (UOC smoothing, mcPrd.h:247)
    bool   firstSampleIsToday = false;
    double firstSampleForward = 0.0;
    // AAD: for each slot of the device adjoint vector, the host-tape Number it belongs to
    std::vector<Number*> adjointTargets;
};

struct ProductImage
{
    cf_product pod{};
    std::vector<int32_t> strikeOffsets;
    std::vector<double>  strikes, weights, eventDt;
};

// ---------------------------------------------------------------------------------------------
// Products, models, RNGs (mcBase.h:122-246)
// ---------------------------------------------------------------------------------------------
template <class T>
class Product
{
    inline static const std::vector<std::string> defaultAssetNames = {"spot"};

public:
    virtual const std::vector<Time>& timeline() const = 0;
    virtual const std::vector<SampleDef>& defline() const = 0;
    virtual const size_t numAssets() const { return 1; }
    virtual const std::vector<std::string>& assetNames() const { return defaultAssetNames; }
    virtual const std::vector<std::string>& payoffLabels() const = 0;

    // Payoffs of one path.  The built-in products are evaluated on the device inside the path
    // kernels; this host entry exists for interface compatibility and throws unless overridden.
    virtual void payoffs(const Scenario<T>& /*path*/, std::vector<T>& /*payoffs*/) const
    {
        throw std::runtime_error("Product::payoffs(): paths are evaluated on the GPU, no host evaluation");
    }

    virtual std::unique_ptr<Product<T>> clone() const = 0;
    virtual ~Product() {}

    // Describe the payoff to the CUDA engine.  Return false if the product cannot run on the device.
    virtual bool deviceImage(ProductImage& /*img*/, const ModelImage& /*mdl*/) const { return false; }
};

template <class T>
class Model
{
    inline static const std::vector<std::string> defaultAssetNames = {"spot"};

public:
    virtual const size_t numAssets() const { return 1; }
    virtual const std::vector<std::string>& assetNames() const { return defaultAssetNames; }

    virtual void allocate(const std::vector<Time>& prdTimeline, const std::vector<SampleDef>& prdDefline) = 0;
    virtual void init(const std::vector<Time>& prdTimeline, const std::vector<SampleDef>& prdDefline) = 0;
    virtual size_t simDim() const = 0;

    virtual void generatePath(const std::vector<double>& /*gaussVec*/, Scenario<T>& /*path*/) const
    {
        throw std::runtime_error("Model::generatePath(): paths are generated on the GPU, no host evaluation");
    }

    virtual std::unique_ptr<Model<T>> clone() const = 0;
    virtual ~Model() {}

    virtual const std::vector<T*>& parameters() = 0;
    virtual const std::vector<std::string>& parameterLabels() const = 0;
    size_t numParams() const { return const_cast<Model*>(this)->parameters().size(); }

    void putParametersOnTape()
    {
        if constexpr (std::is_same<T, Number>::value)
            for (Number* param : parameters()) param->putOnTape();
    }

    // Describe the initialised model (after allocate + init) to the CUDA engine.
    virtual bool deviceImage(ModelImage& /*img*/, const std::vector<Time>& /*prdTimeline*/,
                             const std::vector<SampleDef>& /*prdDefline*/) { return false; }
};

class RNG
{
public:
    virtual void init(const size_t simDim) = 0;
    virtual void nextU(std::vector<double>& uVec) = 0;
    virtual void nextG(std::vector<double>& gaussVec) = 0;
    virtual std::unique_ptr<RNG> clone() const = 0;
    virtual ~RNG() {}
    virtual void skipTo(const unsigned b) = 0;
    // Describe the generator to the CUDA engine.
    virtual bool deviceImage(cf_rng& /*img*/) const { return false; }
};

template <class T>
inline bool checkCompatiblity(const Product<T>& prd, const Model<T>& mdl) { return prd.assetNames() == mdl.assetNames(); }

// ---------------------------------------------------------------------------------------------
// Engine glue
// ---------------------------------------------------------------------------------------------
inline void cfCheck(int rc)
{
    if (rc != 0) throw std::runtime_error(cf_last_error());
}

struct CfDeviceSetup
{
    ModelImage   mdl;
    ProductImage prd;
    cf_rng       rng{};
};

template <class T>
inline void cfBuildImages(const Product<T>& prd, Model<T>& initialisedMdl, const RNG& rng, CfDeviceSetup& s)
{
    if (!initialisedMdl.deviceImage(s.mdl, prd.timeline(), prd.defline()))
        throw std::runtime_error("This model has no device image: it cannot run on the CUDA engine");
    if (!prd.deviceImage(s.prd, s.mdl))
        throw std::runtime_error("This product has no device image: it cannot run on the CUDA engine");
    if (!rng.deviceImage(s.rng))
        throw std::runtime_error("This RNG has no device image: it cannot run on the CUDA engine");
}

// ---------------------------------------------------------------------------------------------
// Resident sessions.  The reference's entry points look their model and product up in the store by name and run
// on a pool of threads that lives as long as the process (xlExport.cpp:1605); a risk report is asked for again and
// again on the same stored objects.  What does not depend on the call -- the clone of the model initialised on the
// product's timeline (for AAD: on its own tape, parameters and init() before the mark, mcBase.h:537-561), the flat
// device images and the engine's plan with its tables resident in HBM -- is therefore kept per (model, product, RNG)
// and reused until one of them is put again (cf_store.h serial numbers), the evaluation date moves or the device
// context changes.  CF_HOST_CACHE=0 in the environment switches the reuse off (every call rebuilds everything).
// ---------------------------------------------------------------------------------------------
struct CfSessionKey
{
    uint64_t modelSerial = 0, productSerial = 0;     // 0: not from the store, never cached
    int      rngKind = 0;
    uint32_t seed1 = 0, seed2 = 0;
    double   sysTime = 0.0;
    int      contextGen = 0;
    bool     aad = false;
    bool operator==(const CfSessionKey& o) const
    {
        return modelSerial == o.modelSerial && productSerial == o.productSerial && rngKind == o.rngKind && seed1 == o.seed1
               && seed2 == o.seed2 && sysTime == o.sysTime && contextGen == o.contextGen && aad == o.aad;
    }
};

struct CfSession
{
    CfSessionKey                   key;
    std::unique_ptr<Model<Number>> mdlN;      // AAD: initialised clone, recorded on `tape`
    std::unique_ptr<Model<double>> mdlD;      // value
    Tape                           tape;
    CfDeviceSetup                  setup;     // PODs point into its own vectors: a session is never moved
    cf_plan*                       plan = nullptr;
    size_t                         nAdj = 0, nPay = 0;
    // device adjoint q lands on parameter directParam[q] itself (-1: on nothing); directState 1: that holds for every
    // target, the host tape has nothing to propagate (Dupire with its time map: spot and local vols); -1: it does not
    std::vector<int>               directParam;
    int                            directState = 0;
    CfSession() = default;
    CfSession(const CfSession&) = delete;
    CfSession& operator=(const CfSession&) = delete;
    ~CfSession() { if (plan) cf_plan_destroy(plan); }
};

// Are the targets of the device adjoints the model's parameters themselves?  Then the sweep mark -> start of
// mcBase.h:518 only visits nodes with zero adjoints and the risks are the device sums, added in the same order.
inline bool cfDirectTargets(CfSession& s, const std::vector<Number*>& params)
{
    if (s.directState == 0) {
        const auto& targets = s.setup.mdl.adjointTargets;
        // the common case first: target q IS parameter q (Dupire with its time map lists spot and vols in parameter order)
        if (targets.size() == params.size() && std::equal(targets.begin(), targets.end(), params.begin())) {
            s.directParam.resize(targets.size());
            std::iota(s.directParam.begin(), s.directParam.end(), 0);
            s.directState = 1;
            return true;
        }
        std::unordered_map<const Number*, int> where;
        for (size_t j = 0; j < params.size(); ++j) where.emplace(params[j], int(j));
        s.directParam.assign(targets.size(), -1);
        s.directState = 1;
        for (size_t q = 0; q < targets.size() && s.directState == 1; ++q) {
            if (!targets[q]) continue;
            const auto it = where.find(targets[q]);
            if (it == where.end()) s.directState = -1; else s.directParam[q] = it->second;
        }
    }
    return s.directState == 1;
}

// Number::tape points at the session's tape while the session is worked on
struct CfTapeScope
{
    Tape* saved;
    explicit CfTapeScope(Tape* t) : saved(Number::tape) { Number::tape = t; }
    ~CfTapeScope() { Number::tape = saved; }
};

inline bool cfHostCacheEnabled()
{
    static const bool on = [] { const char* e = std::getenv("CF_HOST_CACHE"); return !e || std::atoi(e) != 0; }();
    return on;
}

inline std::vector<std::unique_ptr<CfSession>>& cfSessions() { static std::vector<std::unique_ptr<CfSession>> v; return v; }

inline CfSessionKey cfMakeKey(const uint64_t modelSerial, const uint64_t productSerial, const RNG& rng, const bool aad)
{
    CfSessionKey k;
    cf_rng r{};
    if (!rng.deviceImage(r)) throw std::runtime_error("This RNG has no device image: it cannot run on the CUDA engine");
    k.modelSerial = modelSerial; k.productSerial = productSerial; k.rngKind = r.kind; k.seed1 = r.seed1; k.seed2 = r.seed2;
    k.sysTime = systemTime; k.contextGen = cf_context_generation(); k.aad = aad;
    return k;
}

// Builds the session: clone, allocate, init (AAD: on the session's tape, then mark), device images, resident plan.
template <class T>
inline std::unique_ptr<CfSession> cfBuildSession(const Product<T>& prd, const Model<T>& mdl, const RNG& rng, const bool aad)
{
    if (!checkCompatiblity(prd, mdl)) throw std::runtime_error("Model and product are not compatible");
    auto s = std::make_unique<CfSession>();
    CfTapeScope scope(&s->tape);
    auto cMdl = mdl.clone();
    cMdl->allocate(prd.timeline(), prd.defline());
    if constexpr (std::is_same<T, Number>::value) {
        // AAD - 1 (mcBase.h:455-472): parameters and init() on tape, then mark
        s->tape.clear();
        cMdl->putParametersOnTape();
        cMdl->init(prd.timeline(), prd.defline());
        s->tape.mark();
    } else {
        cMdl->init(prd.timeline(), prd.defline());
    }
    cfBuildImages(prd, *cMdl, rng, s->setup);
    s->nPay = prd.payoffLabels().size();
    s->nAdj = cf_table_adjoint_size(&s->setup.mdl.pod, &s->setup.prd.pod);
    if (aad && s->nAdj != s->setup.mdl.adjointTargets.size())
        throw std::runtime_error("mcSimulAAD: device adjoint layout does not match the model's host tables");
    cfCheck(cf_plan_create(&s->setup.mdl.pod, &s->setup.prd.pod, &s->setup.rng, &s->plan));
    if constexpr (std::is_same<T, Number>::value) s->mdlN = std::move(cMdl); else s->mdlD = std::move(cMdl);
    return s;
}

// The session of `key` if it is resident, else a new one (kept when the key is cacheable; at most 16 are kept).
template <class T>
inline CfSession* cfAcquireSession(const Product<T>& prd, const Model<T>& mdl, const RNG& rng, const CfSessionKey* key,
                                   const bool aad, std::unique_ptr<CfSession>& owner)
{
    const bool cacheable = key && key->modelSerial && key->productSerial && cfHostCacheEnabled();
    auto& all = cfSessions();
    if (cacheable)
        for (auto& s : all) if (s->key == *key) return s.get();
    auto fresh = cfBuildSession(prd, mdl, rng, aad);
    if (!cacheable) { owner = std::move(fresh); return owner.get(); }
    fresh->key = *key;
    // sessions of objects that have been put again or of a closed context can never be hit: drop them
    all.erase(std::remove_if(all.begin(), all.end(), [&](const std::unique_ptr<CfSession>& s) {
                  return s->key.contextGen != key->contextGen;
              }), all.end());
    if (all.size() >= 16) all.erase(all.begin());
    // the sessions hold device plans: they are released at exit while the CUDA runtime is still up (exit handlers run in
    // reverse order of registration; the runtime registered its own before the plan above could be created)
    static const bool atExit = [] { std::atexit([] { cfSessions().clear(); }); return true; }();
    (void)atExit;
    all.push_back(std::move(fresh));
    return all.back().get();
}

inline void cfDropSessions() { cfSessions().clear(); }

// Sums only (what main.h actually consumes): payoff sums over paths.
inline std::vector<double> cfSimulSums(const Product<double>& prd, const Model<double>& mdl, const RNG& rng,
                                       const size_t nPath, std::vector<double>* perPath = nullptr,
                                       const CfSessionKey* key = nullptr)
{
    std::unique_ptr<CfSession> owner;
    CfSession* s = cfAcquireSession(prd, mdl, rng, key, false, owner);
    std::vector<double> sums(s->nPay);
    if (perPath) perPath->resize(nPath * s->nPay);
    cfCheck(cf_plan_run_value(s->plan, 0, nPath, sums.data(), perPath ? perPath->data() : nullptr));
    return sums;
}

// mcSimul / mcParallelSimul (mcBase.h:267-400): matrix (0..nPath-1, 0..nPay-1) of payoffs
inline std::vector<std::vector<double>> mcSimul(const Product<double>& prd, const Model<double>& mdl, const RNG& rng,
                                                const size_t nPath)
{
    std::vector<double> flat;
    cfSimulSums(prd, mdl, rng, nPath, &flat);
    const size_t nPay = prd.payoffLabels().size();
    std::vector<std::vector<double>> results(nPath, std::vector<double>(nPay));
    for (size_t i = 0; i < nPath; ++i) std::copy(flat.begin() + i * nPay, flat.begin() + (i + 1) * nPay, results[i].begin());
    return results;
}
inline std::vector<std::vector<double>> mcParallelSimul(const Product<double>& prd, const Model<double>& mdl,
                                                        const RNG& rng, const size_t nPath)
{
    return mcSimul(prd, mdl, rng, nPath);
}

// AAD results (mcBase.h:405-422)
struct AADSimulResults
{
    AADSimulResults(const size_t nPath, const size_t nPay, const size_t nParam)
        : payoffs(nPath, std::vector<double>(nPay)), aggregated(nPath), risks(nParam) {}
    std::vector<std::vector<double>> payoffs;
    std::vector<double>              aggregated;
    std::vector<double>              risks;
};

// Sums-only AAD results used by the main.h-level entry points
struct AADSums
{
    std::vector<double> payoffSums;
    double              aggSum = 0.0;
    std::vector<double> risks;      // already divided by nPath (mcBase.h:745)
};

const auto defaultAggregator = [](const std::vector<Number>& v) { return v[0]; };

// The aggregators the reference passes (main.h:135, 210-213) are linear in the payoffs; the device
// takes the weight vector.  Recover it by differentiating aggFun on the host tape at two points and
// refuse non-linear aggregators loudly.
template <class F>
inline std::vector<double> cfAggregatorWeights(const F& aggFun, const size_t nPay)
{
    auto gradAt = [&](const double base, const double step) {
        Tape localTape;
        Tape* saved = Number::tape;
        Number::tape = &localTape;
        std::vector<Number> pays(nPay);
        for (size_t k = 0; k < nPay; ++k) { pays[k] = Number(base + step * double(k)); pays[k].putOnTape(); }
        Number result = aggFun(pays);
        std::vector<double> g(nPay, 0.0);
        if (result.onTape()) {
            result.propagateToStart();
            for (size_t k = 0; k < nPay; ++k) g[k] = pays[k].adjoint();
        }
        Number::tape = saved;
        return g;
    };
    const auto g1 = gradAt(1.0, 0.25), g2 = gradAt(3.0, -0.125);
    for (size_t k = 0; k < nPay; ++k)
        if (std::fabs(g1[k] - g2[k]) > 1e-12 * (1.0 + std::fabs(g1[k])))
            throw std::runtime_error("mcSimulAAD: only linear aggregators of the payoffs are supported on the device");
    return g1;
}

// Core of mcSimulAAD / mcParallelSimulAAD: tape for the path-independent stage on the host,
// paths + adjoint sweep on the device, mark-to-start propagation on the host.
inline AADSums cfSimulAADSums(const Product<Number>& prd, const Model<Number>& mdl, const RNG& rng, const size_t nPath,
                              const std::vector<double>& weights, std::vector<double>* perPathPayoffs = nullptr,
                              std::vector<double>* perPathAgg = nullptr, const CfSessionKey* key = nullptr)
{
    static const bool timing = std::getenv("CF_TIMING") != nullptr;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto us = [](auto a, auto b) { return std::chrono::duration<double, std::micro>(b - a).count(); };
    const auto t0 = now();
    std::unique_ptr<CfSession> owner;
    CfSession* s = cfAcquireSession(prd, mdl, rng, key, true, owner);
    CfTapeScope scope(&s->tape);
    const std::vector<Number*>& params = s->mdlN->parameters();
    const size_t nParam = params.size(), nPay = s->nPay, nAdj = s->nAdj;
    if (weights.size() != nPay) throw std::runtime_error("mcSimulAAD: one weight per payoff");
    const auto t1 = now();

    AADSums out;
    out.payoffSums.resize(nPay);
    std::vector<double> adj(nAdj);
    if (perPathPayoffs) perPathPayoffs->resize(nPath * nPay);
    if (perPathAgg) perPathAgg->resize(nPath);
    cfCheck(cf_plan_run_aad(s->plan, weights.data(), 0, nPath, out.payoffSums.data(), &out.aggSum, adj.data(),
                            perPathPayoffs ? perPathPayoffs->data() : nullptr, perPathAgg ? perPathAgg->data() : nullptr));

    const auto t2 = now();
    // AAD - 4 (mcBase.h:512-527): adjoints accumulated over paths on the pre-mark nodes, one sweep mark -> start
    out.risks.assign(nParam, 0.0);
    if (cfDirectTargets(*s, params)) {
        for (size_t k = 0; k < nAdj; ++k)
            if (s->directParam[k] >= 0) out.risks[size_t(s->directParam[k])] += adj[k];
        for (size_t j = 0; j < nParam; ++j) out.risks[j] /= double(nPath);
    } else {
        s->tape.resetAdjoints();
        const auto& targets = s->setup.mdl.adjointTargets;
        for (size_t k = 0; k < nAdj; ++k)
            if (targets[k]) targets[k]->adjoint() += adj[k];
        Number::propagateMarkToStart();
        for (size_t j = 0; j < nParam; ++j) out.risks[j] = params[j]->adjoint() / double(nPath);
    }
    if (timing) std::fprintf(stderr, "cfSimulAADSums: session (clone + init on tape + images + plan, or reuse) %.0f us, device run %.0f us, chain rule %.0f us\n",
                             us(t0, t1), us(t1, t2), us(t2, now()));
    return out;
}

template <class F = decltype(defaultAggregator)>
inline AADSimulResults mcSimulAAD(const Product<Number>& prd, const Model<Number>& mdl, const RNG& rng,
                                  const size_t nPath, const F& aggFun = defaultAggregator)
{
    const size_t nPay = prd.payoffLabels().size();
    const auto weights = cfAggregatorWeights(aggFun, nPay);
    std::vector<double> pp, pa;
    AADSums sums = cfSimulAADSums(prd, mdl, rng, nPath, weights, &pp, &pa);
    AADSimulResults results(nPath, nPay, sums.risks.size());
    for (size_t i = 0; i < nPath; ++i) std::copy(pp.begin() + i * nPay, pp.begin() + (i + 1) * nPay, results.payoffs[i].begin());
    results.aggregated = std::move(pa);
    results.risks = std::move(sums.risks);
    return results;
}

template <class F = decltype(defaultAggregator)>
inline AADSimulResults mcParallelSimulAAD(const Product<Number>& prd, const Model<Number>& mdl, const RNG& rng,
                                          const size_t nPath, const F& aggFun = defaultAggregator)
{
    return mcSimulAAD(prd, mdl, rng, nPath, aggFun);
}

// Itemised AAD results (mcBase.h:758-771)
struct AADMultiSimulResults
{
    AADMultiSimulResults(const size_t nPath, const size_t nPay, const size_t nParam)
        : payoffs(nPath, std::vector<double>(nPay)), risks(nParam, nPay) {}
    std::vector<std::vector<double>> payoffs;     // filled only when requested (see mcSimulAADMulti)
    matrix<double>                   risks;       // (0..nParam-1, 0..nPay-1), averaged over paths
};

struct AADMultiSums
{
    std::vector<double> payoffSums;
    matrix<double>      risks;                     // already divided by nPath (mcBase.h:846-851)
};

// Core of mcSimulAADMulti / mcParallelSimulAADMulti (mcBase.h:776, 859): the device returns, per payoff, the
// adjoints of the init() tables summed over paths; the host sweeps its tape mark -> start once per payoff
// (the reference does the same sweep with nPay adjoints per node, AADNode.h:85-102).
inline AADMultiSums cfSimulAADMultiSums(const Product<Number>& prd, const Model<Number>& mdl, const RNG& rng, const size_t nPath,
                                        std::vector<double>* perPathPayoffs = nullptr, const CfSessionKey* key = nullptr)
{
    std::unique_ptr<CfSession> owner;
    CfSession* s = cfAcquireSession(prd, mdl, rng, key, true, owner);
    CfTapeScope scope(&s->tape);
    const std::vector<Number*>& params = s->mdlN->parameters();
    const size_t nParam = params.size(), nPay = s->nPay, nAdj = s->nAdj;

    AADMultiSums out;
    out.payoffSums.resize(nPay);
    std::vector<double> tables(nAdj * nPay);
    cfCheck(cf_plan_run_aad_multi(s->plan, 0, nPath, out.payoffSums.data(), tables.data()));
    if (perPathPayoffs) {
        // the per-path payoffs do not depend on the AAD mode: a value run of the same plan on the same paths
        std::vector<double> sums(nPay);
        perPathPayoffs->resize(nPath * nPay);
        cfCheck(cf_plan_run_value(s->plan, 0, nPath, sums.data(), perPathPayoffs->data()));
    }

    const auto& targets = s->setup.mdl.adjointTargets;
    out.risks.resize(nParam, nPay);
    if (cfDirectTargets(*s, params)) {
        for (size_t j = 0; j < nParam; ++j) std::fill(out.risks[j], out.risks[j] + nPay, 0.0);
        for (size_t q = 0; q < nAdj; ++q) {
            if (s->directParam[q] < 0) continue;
            double* row = out.risks[size_t(s->directParam[q])];
            const double* src = tables.data() + q * nPay;
            for (size_t k = 0; k < nPay; ++k) row[k] += src[k];
        }
        for (size_t j = 0; j < nParam; ++j)
            for (size_t k = 0; k < nPay; ++k) out.risks[j][k] /= double(nPath);
        return out;
    }
    for (size_t k = 0; k < nPay; ++k) {
        s->tape.resetAdjoints();
        for (size_t q = 0; q < nAdj; ++q)
            if (targets[q]) targets[q]->adjoint() += tables[q * nPay + k];
        Number::propagateMarkToStart();
        for (size_t j = 0; j < nParam; ++j) out.risks[j][k] = params[j]->adjoint() / double(nPath);
    }
    return out;
}

// mcSimulAADMulti (mcBase.h:776).  The reference result carries the per-path payoff matrix (mcBase.h:758-771): it is
// filled from a value run on the same paths as long as it stays below 2^27 doubles (1 GB; config 4 at its full 2^22
// paths x 720 payoffs would be 24 GB on the host -- the entry points of main.h only average it), else left empty.
inline AADMultiSimulResults mcSimulAADMulti(const Product<Number>& prd, const Model<Number>& mdl, const RNG& rng, const size_t nPath)
{
    const size_t nPay = prd.payoffLabels().size();
    const bool withPaths = double(nPath) * double(nPay) <= double(size_t(1) << 27);
    std::vector<double> flat;
    AADMultiSums sums = cfSimulAADMultiSums(prd, mdl, rng, nPath, withPaths ? &flat : nullptr);
    AADMultiSimulResults results(withPaths ? nPath : 0, nPay, sums.risks.rows());
    if (withPaths)
        for (size_t i = 0; i < nPath; ++i) std::copy(flat.begin() + i * nPay, flat.begin() + (i + 1) * nPay, results.payoffs[i].begin());
    results.risks = std::move(sums.risks);
    return results;
}
inline AADMultiSimulResults mcParallelSimulAADMulti(const Product<Number>& prd, const Model<Number>& mdl, const RNG& rng,
                                                    const size_t nPath)
{
    return mcSimulAADMulti(prd, mdl, rng, nPath);
}

// ThreadPool facade (threadPool.h:72-171): the path loops run on the GPU, the pool has nothing to
// do; kept so that client code starting / resizing the pool (xlExport.cpp:72-81, 1605) still links.
class ThreadPool
{
    size_t myThreads = 0;
public:
    static ThreadPool* getInstance() { static ThreadPool instance; return &instance; }
    size_t numThreads() const { return myThreads; }
    static size_t threadNum() { return 0; }
    void start(const size_t nThread = 0) { myThreads = nThread; }
    void stop() { myThreads = 0; }
};
