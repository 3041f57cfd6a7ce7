// cf_xl.h -- the Excel-facing wrappers of the reference (xlExport.cpp:72-1255), same names, argument lists,
// result shapes and error behaviour (#N/A on a bad argument or on any exception), over the GPU-backed
// entry points of cf_main.h.  Portable: the Excel types come from cf_xlcall.h, results live in a
// process arena that is released at the start of the next call (the reference's FreeAllTempMemory,
// xlMemoryPool.h:22).  Registration with Excel (xlAutoOpen, xlExport.cpp:1259-1608) is the add-in shell's
// job and is not part of the hot path; the toy-code wrappers (xToyDupireBarrierMc*) are out of scope.
//
// Included by cf_export.cpp (one translation unit: the stores are header-defined globals, store.h:35-36).
#pragma once

#include <cmath>
#include <ctime>
#include <deque>
#include <limits>
#include <memory>
#include <unordered_map>

#include "cf_xlcall.h"

namespace cfxl {

// ---- temporary memory: everything a wrapper returns is valid until the next wrapper call
inline std::deque<std::unique_ptr<unsigned char[]>>& arena()
{
    static std::deque<std::unique_ptr<unsigned char[]>> a;
    return a;
}
inline void freeAll() { arena().clear(); }
inline void* temp(const size_t bytes)
{
    arena().emplace_back(new unsigned char[bytes ? bytes : 1]());
    return arena().back().get();
}
inline LPXLOPER12 tempOper() { return static_cast<LPXLOPER12>(temp(sizeof(XLOPER12))); }

inline XLOPER12 makeStr(const std::string& s)
{
    XLOPER12 x{};
    x.xltype = xltypeStr;
    const size_t n = std::min<size_t>(s.size(), 32767);
    x.val.str = static_cast<XCHAR*>(temp((n + 1) * sizeof(XCHAR)));
    x.val.str[0] = XCHAR(n);
    for (size_t i = 0; i < n; ++i) x.val.str[i + 1] = XCHAR(static_cast<unsigned char>(s[i]));
    return x;
}
inline XLOPER12 makeNum(const double v) { XLOPER12 x{}; x.xltype = xltypeNum; x.val.num = v; return x; }

inline LPXLOPER12 TempStr12(const std::string& s) { LPXLOPER12 p = tempOper(); *p = makeStr(s); return p; }
inline LPXLOPER12 TempNum12(const double v) { LPXLOPER12 p = tempOper(); *p = makeNum(v); return p; }
inline LPXLOPER12 TempErr12(const int err) { LPXLOPER12 p = tempOper(); p->xltype = xltypeErr; p->val.err = err; return p; }

// ---- reading arguments (xlOper.h:63-128)
inline size_t getRows(const LPXLOPER12 o) { return !o ? 0 : o->xltype != xltypeMulti ? 1 : size_t(o->val.array.rows); }
inline size_t getCols(const LPXLOPER12 o) { return !o ? 0 : o->xltype != xltypeMulti ? 1 : size_t(o->val.array.columns); }

inline std::string getString(const LPXLOPER12 o, const size_t i = 0, const size_t j = 0)
{
    if (!o) return "";
    const XLOPER12* s = nullptr;
    if (o->xltype == xltypeStr && i == 0 && j == 0) s = o;
    else if (o->xltype == xltypeMulti) {
        s = o->val.array.lparray + i * getCols(o) + j;
        if (s->xltype != xltypeStr) return "";
    } else return "";
    std::string out(size_t(s->val.str[0]), ' ');
    for (size_t k = 0; k < out.size(); ++k) out[k] = char(s->val.str[k + 1]);
    return out;
}

inline std::vector<std::string> toStrVector(const LPXLOPER12 o)
{
    std::vector<std::string> v;
    if (!o) return v;
    if (o->xltype == xltypeStr) v.push_back(getString(o));
    else if (o->xltype == xltypeMulti)
        for (size_t i = 0; i < getRows(o); ++i) for (size_t j = 0; j < getCols(o); ++j) v.push_back(getString(o, i, j));
    return v;
}
inline size_t fpSize(const FP12* f) { return f ? size_t(f->rows) * size_t(f->columns) : 0; }
inline std::vector<double> toVector(const FP12* f) { return std::vector<double>(f->array, f->array + fpSize(f)); }
inline matrix<double> toMatrix(const FP12* f)
{
    matrix<double> m(size_t(f->rows), size_t(f->columns));
    std::copy(f->array, f->array + fpSize(f), m.begin());
    return m;
}

// ---- building results (xlOper.h:130-290)
inline LPXLOPER12 makeMulti(const size_t rows, const size_t cols)
{
    LPXLOPER12 o = tempOper();
    o->xltype = xltypeMulti;
    o->val.array.rows = RW(rows); o->val.array.columns = COL(cols);
    o->val.array.lparray = static_cast<LPXLOPER12>(temp(rows * cols * sizeof(XLOPER12)));
    const XLOPER12 blank = makeStr("");
    std::fill(o->val.array.lparray, o->val.array.lparray + rows * cols, blank);
    return o;
}
inline void setString(LPXLOPER12 o, const std::string& s, const size_t i, const size_t j) { o->val.array.lparray[i * getCols(o) + j] = makeStr(s); }
inline void setNum(LPXLOPER12 o, const double v, const size_t i, const size_t j) { o->val.array.lparray[i * getCols(o) + j] = makeNum(v); }

inline LPXLOPER12 fromStrVector(const std::vector<std::string>& v)
{
    LPXLOPER12 o = makeMulti(v.size(), 1);                     // one column
    for (size_t i = 0; i < v.size(); ++i) setString(o, v[i], i, 0);
    return o;
}
inline LPXLOPER12 fromLabelsAndNumbers(const std::vector<std::string>& labels, const std::vector<double>& numbers)
{
    const size_t n = labels.size();
    if (n == 0 || n != numbers.size()) return TempErr12(xlerrNA);
    LPXLOPER12 o = makeMulti(n, 2);
    for (size_t i = 0; i < n; ++i) { setString(o, labels[i], i, 0); setNum(o, numbers[i], i, 1); }
    return o;
}
inline LPXLOPER12 fromMatrix(const matrix<double>& m)
{
    if (m.rows() == 0 || m.cols() == 0) return TempErr12(xlerrNA);
    LPXLOPER12 o = makeMulti(m.rows(), m.cols());
    for (size_t i = 0; i < m.rows(); ++i) for (size_t j = 0; j < m.cols(); ++j) setNum(o, m[i][j], i, j);
    return o;
}
// labels down the first column and along the first row
inline LPXLOPER12 fromLabelledMatrix(const std::vector<double>& rowLabels, const std::vector<double>& colLabels, const matrix<double>& m)
{
    const size_t n = rowLabels.size(), k = colLabels.size();
    if (n == 0 || k == 0 || n != m.rows() || k != m.cols()) return TempErr12(xlerrNA);
    LPXLOPER12 o = makeMulti(n + 1, k + 1);
    for (size_t i = 0; i < n; ++i) setNum(o, rowLabels[i], i + 1, 0);
    for (size_t j = 0; j < k; ++j) setNum(o, colLabels[j], 0, j + 1);
    for (size_t i = 0; i < n; ++i) for (size_t j = 0; j < k; ++j) setNum(o, m[i][j], i + 1, j + 1);
    return o;
}
// string labels plus one labelled line (the values) between the header and the matrix
inline LPXLOPER12 fromLabelledMatrix(const std::vector<std::string>& rowLabels, const std::vector<std::string>& colLabels,
                                     const matrix<double>& m, const std::string& firstLineLabel, const std::vector<double>& firstLine)
{
    const size_t n = rowLabels.size(), k = colLabels.size();
    if (n == 0 || k == 0 || n != m.rows() || k != m.cols() || k != firstLine.size()) return TempErr12(xlerrNA);
    LPXLOPER12 o = makeMulti(n + 2, k + 1);
    setString(o, firstLineLabel, 1, 0);
    for (size_t j = 0; j < k; ++j) { setString(o, colLabels[j], 0, j + 1); setNum(o, firstLine[j], 1, j + 1); }
    for (size_t i = 0; i < n; ++i) {
        setString(o, rowLabels[i], i + 2, 0);
        for (size_t j = 0; j < k; ++j) setNum(o, m[i][j], i + 2, j + 1);
    }
    return o;
}

// numerical parameters as Excel passes them: doubles (xlExport.cpp:35-66)
inline NumericalParam xl2num(const double useSobol, const double seed1, const double seed2, const double numPath, const double parallel)
{
    NumericalParam num;
    num.numPath = static_cast<int>(numPath + EPS);
    num.parallel = parallel > EPS;
    num.seed1 = seed1 >= 1 ? static_cast<int>(seed1 + EPS) : 1234;
    num.seed2 = seed2 >= 1 ? static_cast<int>(seed2 + EPS) : num.seed1 + 1;
    num.useSobol = useSobol > EPS;
    return num;
}

// payoff labels and notionals from two ranges of the same shape, blanks and zero notionals dropped (xlExport.cpp:727-747)
inline bool readNotionals(const LPXLOPER12 xPayoffs, const FP12* xNotionals, std::map<std::string, double>& notionals)
{
    const size_t rows = getRows(xPayoffs), cols = getCols(xPayoffs);
    if (rows * cols == 0 || !xNotionals || fpSize(xNotionals) != rows * cols) return false;
    size_t idx = 0;
    for (size_t i = 0; i < rows; ++i) for (size_t j = 0; j < cols; ++j) {
        const std::string payoff = getString(xPayoffs, i, j);
        const double notional = xNotionals->array[idx++];
        if (!payoff.empty() && std::fabs(notional) > EPS) notionals[payoff] = notional;
    }
    return true;
}

inline LPXLOPER12 riskColumn(const AADRiskResults& r)
{
    const size_t n = r.risks.size();
    LPXLOPER12 o = makeMulti(n + 1, 2);
    setString(o, "value", 0, 0); setNum(o, r.riskPayoffValue, 0, 1);
    for (size_t i = 0; i < n; ++i) { setString(o, r.paramIds[i], i + 1, 0); setNum(o, r.risks[i], i + 1, 1); }
    return o;
}

inline std::unordered_map<std::string, RiskReports>& riskStore()
{
    static std::unordered_map<std::string, RiskReports> s;          // xlExport.cpp:774
    return s;
}

// every wrapper: release the previous results, map any exception to #N/A
template <class F>
inline LPXLOPER12 wrap(F&& body)
{
    freeAll();
    try { return body(); }
    catch (const std::exception&) { return TempErr12(xlerrNA); }
}

}  // namespace cfxl

extern "C" {

// xlExport.cpp:72-81
double xRestartThreadPool(double xNthread)
{
    const int numThread = int(xNthread + EPS);
    ThreadPool::getInstance()->stop();
    ThreadPool::getInstance()->start(numThread);
    return numThread;
}

// xlExport.cpp:83-115
LPXLOPER12 xPutDupire(double spot, FP12* spots, FP12* times, FP12* vols, double maxDt, LPXLOPER12 xid)
{
    using namespace cfxl;
    return wrap([&]() -> LPXLOPER12 {
        const std::string id = getString(xid);
        if (maxDt <= 0.0 || id.empty() || !spots || !times || !vols) return TempErr12(xlerrNA);
        if (fpSize(spots) * fpSize(times) != fpSize(vols)) return TempErr12(xlerrNA);
        putDupire(spot, toVector(spots), toVector(times), toMatrix(vols), maxDt, id);
        return TempStr12(id);
    });
}

// xlExport.cpp:117-137
LPXLOPER12 xPutBlackScholes(double spot, double vol, double qSpot, double rate, double div, LPXLOPER12 xid)
{
    using namespace cfxl;
    return wrap([&]() -> LPXLOPER12 {
        const std::string id = getString(xid);
        if (id.empty()) return TempErr12(xlerrNA);
        putBlackScholes(spot, vol, qSpot > 0, rate, div, id);
        return TempStr12(id);
    });
}

// xlExport.cpp:139-194
LPXLOPER12 xPutDLM(LPXLOPER12 assets, FP12* spots, FP12* atms, FP12* skews, double discRate, FP12* repoSpreads, FP12* divDates,
                   FP12* divs, FP12* correl, double lambda, LPXLOPER12 xid)
{
    using namespace cfxl;
    return wrap([&]() -> LPXLOPER12 {
        const std::string id = getString(xid);
        if (id.empty()) return TempErr12(xlerrNA);
        const std::vector<std::string> vassets = toStrVector(assets);
        if (vassets.empty()) return TempErr12(xlerrNA);
        for (const auto& a : vassets) if (a.empty()) return TempErr12(xlerrNA);
        const size_t n = vassets.size();
        if (!spots || !atms || !skews || !repoSpreads || !divDates || !divs || !correl) return TempErr12(xlerrNA);
        if (fpSize(spots) != n || fpSize(atms) != n || fpSize(skews) != n || fpSize(repoSpreads) != n
            || size_t(divs->columns) != n || size_t(divs->rows) != fpSize(divDates)
            || size_t(correl->rows) != n || size_t(correl->columns) != n) return TempErr12(xlerrNA);
        putDisplaced(vassets, toVector(spots), toVector(atms), toVector(skews), discRate, toVector(repoSpreads),
                     toVector(divDates), toMatrix(divs), toMatrix(correl), lambda, id);
        return TempStr12(id);
    });
}

// xlExport.cpp:270-290
LPXLOPER12 xPutEuropean(double strike, double exerciseDate, double settlementDate, LPXLOPER12 xid)
{
    using namespace cfxl;
    return wrap([&]() -> LPXLOPER12 {
        const std::string id = getString(xid);
        if (id.empty()) return TempErr12(xlerrNA);
        putEuropean(strike, exerciseDate, settlementDate <= 0 ? exerciseDate : settlementDate, id);
        return TempStr12(id);
    });
}

// xlExport.cpp:292-317: the call / put flag is a string starting with 'p' or 'P' for a put
LPXLOPER12 xPutBarrier(double strike, double barrier, double maturity, double monitorFreq, double smoothing,
                       LPXLOPER12 xcallput, LPXLOPER12 xid)
{
    using namespace cfxl;
    return wrap([&]() -> LPXLOPER12 {
        const std::string id = getString(xid);
        if (id.empty()) return TempErr12(xlerrNA);
        const std::string cp = getString(xcallput);
        const bool isPut = !cp.empty() && (cp[0] == 'p' || cp[0] == 'P');
        putBarrier(strike, barrier, maturity, monitorFreq, smoothing, isPut, id);
        return TempStr12(id);
    });
}

// xlExport.cpp:319-338
LPXLOPER12 xPutContingent(double coupon, double maturity, double payFreq, double smoothing, LPXLOPER12 xid)
{
    using namespace cfxl;
    return wrap([&]() -> LPXLOPER12 {
        const std::string id = getString(xid);
        if (id.empty()) return TempErr12(xlerrNA);
        putContingent(coupon, maturity, payFreq, smoothing, id);
        return TempStr12(id);
    });
}

// xlExport.cpp:340-385: pairs (maturity, strike) from two ranges of the same shape, blanks (<= EPS) dropped
LPXLOPER12 xPutEuropeans(FP12* maturities, FP12* strikes, LPXLOPER12 xid)
{
    using namespace cfxl;
    return wrap([&]() -> LPXLOPER12 {
        const std::string id = getString(xid);
        if (id.empty() || !maturities || !strikes || fpSize(strikes) == 0 || fpSize(strikes) != fpSize(maturities)) return TempErr12(xlerrNA);
        std::vector<double> vmats, vstrikes;
        for (size_t i = 0; i < fpSize(strikes); ++i)
            if (maturities->array[i] > EPS && strikes->array[i] > EPS) { vmats.push_back(maturities->array[i]); vstrikes.push_back(strikes->array[i]); }
        if (vmats.empty()) return TempErr12(xlerrNA);
        putEuropeans(vmats, vstrikes, id);
        return TempStr12(id);
    });
}

// xlExport.cpp:387-435
LPXLOPER12 xPutMultiStats(LPXLOPER12 assets, FP12* fix, FP12* fwd, LPXLOPER12 xid)
{
    using namespace cfxl;
    return wrap([&]() -> LPXLOPER12 {
        const std::string id = getString(xid);
        if (id.empty()) return TempErr12(xlerrNA);
        const std::vector<std::string> vassets = toStrVector(assets);
        if (vassets.empty() || !fix || !fwd || fpSize(fix) == 0 || fpSize(fix) != fpSize(fwd)) return TempErr12(xlerrNA);
        std::vector<double> vfix, vfwd;
        for (size_t i = 0; i < fpSize(fix); ++i) {
            const double a = fix->array[i], b = fwd->array[i];
            if (a > EPS && b > EPS) {
                if (!vfix.empty() && a <= vfix.back()) return TempErr12(xlerrNA);     // fixings must increase
                if (a > b) return TempErr12(xlerrNA);                                  // forward date on or after the fixing
                vfix.push_back(a); vfwd.push_back(b);
            }
        }
        if (vfix.empty()) return TempErr12(xlerrNA);
        putMultiStats(vassets, vfix, vfwd, id);
        return TempStr12(id);
    });
}

// xlExport.cpp:437-467
LPXLOPER12 xPutBaskets(LPXLOPER12 assets, FP12* weights, double maturity, FP12* strikes, LPXLOPER12 xid)
{
    using namespace cfxl;
    return wrap([&]() -> LPXLOPER12 {
        const std::string id = getString(xid);
        if (id.empty()) return TempErr12(xlerrNA);
        const std::vector<std::string> vassets = toStrVector(assets);
        if (vassets.empty() || maturity <= 0 || !weights || !strikes) return TempErr12(xlerrNA);
        const std::vector<double> vweights = toVector(weights), vstrikes = toVector(strikes);
        if (vweights.empty() || vstrikes.empty()) return TempErr12(xlerrNA);
        putBaskets(vassets, vweights, maturity, vstrikes, id);
        return TempStr12(id);
    });
}

// xlExport.cpp:469-504
LPXLOPER12 xPutAutocall(LPXLOPER12 assets, FP12* refs, double maturity, double periods, double ko, double strike, double cpn,
                        double smooth, LPXLOPER12 xid)
{
    using namespace cfxl;
    return wrap([&]() -> LPXLOPER12 {
        const std::string id = getString(xid);
        if (id.empty() || !refs) return TempErr12(xlerrNA);
        const std::vector<std::string> vassets = toStrVector(assets);
        const std::vector<double> vrefs = toVector(refs);
        if (vassets.empty() || vassets.size() != vrefs.size() || maturity <= 0 || int(periods + EPS) <= 0) return TempErr12(xlerrNA);
        putAutocall(vassets, vrefs, maturity, int(periods + EPS), ko, strike, cpn, smooth, id);
        return TempStr12(id);
    });
}

// xlExport.cpp:506-521
LPXLOPER12 xPayoffIds(LPXLOPER12 xid)
{
    using namespace cfxl;
    return wrap([&]() -> LPXLOPER12 {
        const std::string id = getString(xid);
        if (id.empty()) return TempErr12(xlerrNA);
        const auto* prd = getProduct<double>(id);
        if (!prd) return TempErr12(xlerrNA);
        return fromStrVector(prd->payoffLabels());
    });
}

// xlExport.cpp:523-543
LPXLOPER12 xParameters(LPXLOPER12 xid)
{
    using namespace cfxl;
    return wrap([&]() -> LPXLOPER12 {
        const std::string id = getString(xid);
        if (id.empty()) return TempErr12(xlerrNA);
        const auto params = getModelParameters(id);
        if (!params.first || !params.second) return TempErr12(xlerrNA);
        std::vector<double> values(params.second->size());
        std::transform(params.second->begin(), params.second->end(), values.begin(), [](const double* p) { return *p; });
        return fromLabelsAndNumbers(*params.first, values);
    });
}

// xlExport.cpp:545-589
LPXLOPER12 xValue(LPXLOPER12 modelid, LPXLOPER12 productid, double useSobol, double seed1, double seed2, double numPath, double parallel)
{
    using namespace cfxl;
    return wrap([&]() -> LPXLOPER12 {
        const std::string pid = getString(productid), mid = getString(modelid);
        if (pid.empty() || !getProduct<double>(pid) || mid.empty() || !getModel<double>(mid)) return TempErr12(xlerrNA);
        const auto num = xl2num(useSobol, seed1, seed2, numPath, parallel);
        if (!num.numPath) return TempErr12(xlerrNA);
        const auto results = value(mid, pid, num);
        return fromLabelsAndNumbers(results.identifiers, results.values);
    });
}

// xlExport.cpp:591-642: the values in one column, then the clock ticks the valuation took
LPXLOPER12 xValueTime(LPXLOPER12 modelid, LPXLOPER12 productid, double useSobol, double seed1, double seed2, double numPath, double parallel)
{
    using namespace cfxl;
    return wrap([&]() -> LPXLOPER12 {
        const std::string pid = getString(productid), mid = getString(modelid);
        if (pid.empty() || !getProduct<double>(pid) || mid.empty() || !getModel<double>(mid)) return TempErr12(xlerrNA);
        const auto num = xl2num(useSobol, seed1, seed2, numPath, parallel);
        if (!num.numPath) return TempErr12(xlerrNA);
        const std::clock_t t0 = std::clock();
        const auto results = value(mid, pid, num);
        const std::clock_t t1 = std::clock();
        LPXLOPER12 o = makeMulti(results.values.size() + 1, 1);
        for (size_t i = 0; i < results.values.size(); ++i) setNum(o, results.values[i], i, 0);
        setNum(o, double(t1 - t0), results.values.size(), 0);
        return o;
    });
}

// xlExport.cpp:644-697
LPXLOPER12 xAADrisk(LPXLOPER12 modelid, LPXLOPER12 productid, LPXLOPER12 xRiskPayoff, double useSobol, double seed1, double seed2,
                    double numPath, double parallel)
{
    using namespace cfxl;
    return wrap([&]() -> LPXLOPER12 {
        const std::string pid = getString(productid), mid = getString(modelid);
        if (pid.empty() || mid.empty()) return TempErr12(xlerrNA);
        const auto num = xl2num(useSobol, seed1, seed2, numPath, parallel);
        if (!num.numPath) return TempErr12(xlerrNA);
        return riskColumn(AADriskOne(mid, pid, num, getString(xRiskPayoff)));
    });
}

// xlExport.cpp:699-772
LPXLOPER12 xAADriskAggregate(LPXLOPER12 modelid, LPXLOPER12 productid, LPXLOPER12 xPayoffs, FP12* xNotionals, double useSobol,
                             double seed1, double seed2, double numPath, double parallel)
{
    using namespace cfxl;
    return wrap([&]() -> LPXLOPER12 {
        const std::string pid = getString(productid), mid = getString(modelid);
        if (pid.empty() || mid.empty()) return TempErr12(xlerrNA);
        const auto num = xl2num(useSobol, seed1, seed2, numPath, parallel);
        if (!num.numPath) return TempErr12(xlerrNA);
        std::map<std::string, double> notionals;
        if (!readNotionals(xPayoffs, xNotionals, notionals)) return TempErr12(xlerrNA);
        return riskColumn(AADriskAggregate(mid, pid, notionals, num));
    });
}

// xlExport.cpp:776-824 and 826-874: display the report now, or keep it under an id for xDisplayRisk
static LPXLOPER12 cfxlRiskReport(const bool bump, LPXLOPER12 modelid, LPXLOPER12 productid, double useSobol, double seed1, double seed2,
                                 double numPath, double parallel, double displayNow, LPXLOPER12 storeid)
{
    using namespace cfxl;
    return wrap([&]() -> LPXLOPER12 {
        const std::string pid = getString(productid), mid = getString(modelid);
        if (pid.empty() || mid.empty()) return TempErr12(xlerrNA);
        const auto num = xl2num(useSobol, seed1, seed2, numPath, parallel);
        if (!num.numPath) return TempErr12(xlerrNA);
        const RiskReports results = bump ? bumpRisk(mid, pid, num) : AADriskMulti(mid, pid, num);
        if (displayNow > 0.5) return fromLabelledMatrix(results.params, results.payoffs, results.risks, "value", results.values);
        const std::string riskId = getString(storeid);
        if (riskId.empty()) return TempErr12(xlerrNA);
        riskStore()[riskId] = results;
        return storeid;
    });
}
LPXLOPER12 xBumprisk(LPXLOPER12 modelid, LPXLOPER12 productid, double useSobol, double seed1, double seed2, double numPath,
                     double parallel, double displayNow, LPXLOPER12 storeid)
{
    return cfxlRiskReport(true, modelid, productid, useSobol, seed1, seed2, numPath, parallel, displayNow, storeid);
}
LPXLOPER12 xAADriskMulti(LPXLOPER12 modelid, LPXLOPER12 productid, double useSobol, double seed1, double seed2, double numPath,
                         double parallel, double displayNow, LPXLOPER12 storeid)
{
    return cfxlRiskReport(false, modelid, productid, useSobol, seed1, seed2, numPath, parallel, displayNow, storeid);
}

// xlExport.cpp:876-916: the columns of a stored report for the payoffs asked for
LPXLOPER12 xDisplayRisk(LPXLOPER12 riskid, LPXLOPER12 displayid)
{
    using namespace cfxl;
    return wrap([&]() -> LPXLOPER12 {
        const auto it = riskStore().find(getString(riskid));
        if (it == riskStore().end()) return TempErr12(xlerrNA);
        const RiskReports& rep = it->second;
        const std::vector<std::string> ids = toStrVector(displayid);
        if (ids.empty()) return TempErr12(xlerrNA);
        std::vector<size_t> cols;
        for (const auto& id : ids) {
            const auto f = std::find(rep.payoffs.begin(), rep.payoffs.end(), id);
            if (f == rep.payoffs.end()) return TempErr12(xlerrNA);
            cols.push_back(size_t(f - rep.payoffs.begin()));
        }
        const size_t nParam = rep.risks.rows();
        std::vector<double> vals(cols.size());
        matrix<double> risks(nParam, cols.size());
        for (size_t j = 0; j < cols.size(); ++j) {
            for (size_t i = 0; i < nParam; ++i) risks[i][j] = rep.risks[i][cols[j]];
            vals[j] = rep.values[cols[j]];
        }
        return fromLabelledMatrix(rep.params, ids, risks, "value", vals);
    });
}

// xlExport.cpp:918-963: local vols calibrated to a Merton implied-vol surface, spots down, times across
LPXLOPER12 xDupireCalib(const double spot, const double vol, const double jmpIntens, const double jmpAverage, const double jmpStd,
                        FP12* spots, const double maxDs, FP12* times, const double maxDt)
{
    using namespace cfxl;
    if (maxDs == 0 || maxDt == 0) return nullptr;
    return wrap([&]() -> LPXLOPER12 {
        if (!spots || !times) return TempErr12(xlerrNA);
        const auto results = dupireCalib(toVector(spots), maxDs, toVector(times), maxDt, spot, vol, jmpIntens, jmpAverage, jmpStd);
        return fromLabelledMatrix(results.spots, results.times, results.lVols);
    });
}

// xlExport.cpp:965-1104: value, delta, then the vega matrix with maturities across and strikes down
LPXLOPER12 xDupireSuperbucket(const double spot, const double vol, const double jmpIntens, const double jmpAverage, const double jmpStd,
                              FP12* strikes, FP12* mats, FP12* spots, double maxDs, FP12* times, double maxDtVol, double maxDtSimul,
                              LPXLOPER12 productid, LPXLOPER12 xPayoffs, FP12* xNotionals, double useSobol, double seed1, double seed2,
                              double numPath, double parallel, double bump)
{
    using namespace cfxl;
    return wrap([&]() -> LPXLOPER12 {
        if (numPath <= 0 || !strikes || !mats || !spots || !times) return TempErr12(xlerrNA);
        const std::string pid = getString(productid);
        if (pid.empty()) return TempErr12(xlerrNA);
        const auto num = xl2num(useSobol, seed1, seed2, numPath, parallel);
        if (!num.numPath) return TempErr12(xlerrNA);
        std::map<std::string, double> notionals;
        if (!readNotionals(xPayoffs, xNotionals, notionals)) return TempErr12(xlerrNA);
        const auto r = bump < EPS
            ? dupireSuperbucket(spot, maxDtSimul, pid, notionals, toVector(spots), maxDs, toVector(times), maxDtVol,
                                toVector(strikes), toVector(mats), vol, jmpIntens, jmpAverage, jmpStd, num)
            : dupireSuperbucketBump(spot, maxDtSimul, pid, notionals, toVector(spots), maxDs, toVector(times), maxDtVol,
                                    toVector(strikes), toVector(mats), vol, jmpIntens, jmpAverage, jmpStd, num);
        const size_t n = r.vega.rows(), m = r.vega.cols();
        LPXLOPER12 o = makeMulti(n + 4, m + 2);
        setString(o, "value", 0, 0); setNum(o, r.value, 0, 1);
        setString(o, "delta", 1, 0); setNum(o, r.delta, 1, 1);
        setString(o, "vega", 2, 0); setString(o, "mats", 2, 1);
        for (size_t j = 0; j < m; ++j) setNum(o, r.mats[j], 2, 2 + j);
        setString(o, "strikes", 3, 0);
        for (size_t i = 0; i < n; ++i) {
            setNum(o, r.strikes[i], 4 + i, 1);
            for (size_t j = 0; j < m; ++j) setNum(o, r.vega[i][j], 4 + i, 2 + j);
        }
        return o;
    });
}

// xlExport.cpp:1106-1110
double xMerton(double spot, double vol, double mat, double strike, double intens, double meanJmp, double stdJmp)
{
    return merton(spot, strike, vol, mat, intens, meanJmp, stdJmp);
}

// xlExport.cpp:1221-1255: Sobol points (uniforms), optionally followed by their antithetic 1 - u
LPXLOPER12 xSobolPoints(double numPoints, double dimension, double anti, double skip)
{
    using namespace cfxl;
    return wrap([&]() -> LPXLOPER12 {
        if (numPoints <= 0.0 || dimension <= 0.0) return TempErr12(xlerrNA);
        Sobol rng;
        rng.init(size_t(int(dimension)));
        rng.skipTo(unsigned(skip));
        const size_t dim = size_t(int(dimension)), nPts = size_t(int(numPoints));
        std::vector<double> pt(dim);
        matrix<double> pts(nPts, dim);
        int i = 0;
        while (i < int(numPoints)) {
            rng.nextU(pt);
            std::copy(pt.begin(), pt.end(), pts[size_t(i)]);
            ++i;
            if (i < int(numPoints) && anti > 0.5) {
                for (double& x : pt) x = 1 - x;
                std::copy(pt.begin(), pt.end(), pts[size_t(i)]);
                ++i;
            }
        }
        return fromMatrix(pts);
    });
}

}  // extern "C"
