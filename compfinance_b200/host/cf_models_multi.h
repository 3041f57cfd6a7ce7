// cf_models_multi.h -- host side of the multi-asset displaced-lognormal model of
// mcMdlMultiDisplaced.h: constructor and parameter order (:123-265), allocate (:359-417), init
// (:474-606: repo rates, displacement from ATM and skew, lambda-shifted correlation, Cholesky,
// per-step forward factors / drifts / stds, numeraires and forward factors on event dates).
// init() runs on the host tape for T = Number (the part of the reference's tape before
// tape.mark()); the paths themselves run in cf_dlm.cuh, described by deviceImage().
#pragma once

#include "cf_base.h"

// Cholesky decomposition with the reference's guards (choldc.h:7-50): a pivot below -1e-15 throws,
// below 1e-15 is set to 0, and a row under a zero pivot is 0.
template <class T>
inline void choldc(const matrix<T>& in, matrix<T>& out)
{
    const int n = int(in.rows());
    std::fill(out.begin(), out.end(), T(0.0));
    for (int i = 0; i < n; ++i) {
        for (int j = 0; j <= i; ++j) {
            T sum = in[i][j];
            for (int k = 0; k < j; ++k) sum -= out[i][k] * out[j][k];
            if (i == j) {
                if (sum < -1.0e-15) throw std::runtime_error("choldc : matrix not positive definite");
                if (sum < 1.0e-15) sum = 0.0;
                out[i][i] = sqrt(sum);
            } else {
                if (fabs(out[j][j]) < 1.0e-15) out[i][j] = 0.0;
                else out[i][j] = sum / out[j][j];
            }
        }
    }
}

template <class T>
class MultiDisplaced : public Model<T>
{
public:
    enum Dynamics { Lognormal = 0, Normal = 1, Surnormal = 2, Subnormal = 3 };

private:
    size_t                   myNumAssets;
    std::vector<std::string> myAssetNames;
    T                        myDiscRate;
    std::vector<T>           myRepoSpreads, myRepoRates, mySpots;
    std::vector<Time>        myDivDates;
    matrix<T>                myDivs;             // [div date][asset]
    std::vector<T>           myAtms, mySkews, myAlphas, myBetas;
    std::vector<Dynamics>    myDynamics;
    matrix<T>                myCorrel;           // parameters live in the lower triangle
    T                        myLambda;
    matrix<T>                myUsedCorrel, myChol;

    std::vector<Time> myTimeline;                // today + event dates after today
    bool              myTodayOnTimeline = false;

    matrix<T>                   myStds, myDrifts, myDynFwdFacts;    // [step][asset]
    std::vector<T>              myNumeraires;                       // [event]
    std::vector<std::vector<T>> myDiscounts, myLibors;              // [event][...]
    std::vector<std::vector<std::vector<T>>> myForwardFactors;      // [event][asset][maturity]

    std::vector<T*>          myParameters;
    std::vector<std::string> myParameterLabels;

    void setParamPointers()
    {
        size_t p = 0;
        myParameters[p++] = &myDiscRate;
        for (size_t i = 0; i < myNumAssets; ++i) myParameters[p++] = &mySpots[i];
        for (size_t i = 0; i < myNumAssets; ++i) myParameters[p++] = &myRepoSpreads[i];
        for (size_t i = 0; i < myDivDates.size(); ++i)
            for (size_t j = 0; j < myNumAssets; ++j) myParameters[p++] = &myDivs[i][j];
        for (size_t i = 0; i < myNumAssets; ++i) myParameters[p++] = &myAtms[i];
        for (size_t i = 0; i < myNumAssets; ++i) myParameters[p++] = &mySkews[i];
        for (size_t i = 1; i < myNumAssets; ++i)
            for (size_t j = 0; j < i; ++j) myParameters[p++] = &myCorrel[i][j];
        myParameters[p] = &myLambda;
    }

    // prod(1 - div) over dividend dates in [T1, T2) times exp(repo (T2 - T1)) (mcMdlMultiDisplaced.h:445-470).
    // The single-asset variant of the reference (:421-442) never advances its dividend index; it is only
    // ever reached with T2 == T1 by the shipped products, where both give exp(0) = 1 times an empty product.
    T forwardFactor(const Time T1, const Time T2, const size_t asset) const
    {
        T divProd = T(1.0);
        auto it = std::lower_bound(myDivDates.begin(), myDivDates.end(), T1);
        for (size_t d = size_t(std::distance(myDivDates.begin(), it)); d < myDivDates.size() && myDivDates[d] < T2; ++d)
            divProd *= 1.0 - myDivs[d][asset];
        return divProd * exp(myRepoRates[asset] * (T2 - T1));
    }

public:
    template <class U>
    MultiDisplaced(const std::vector<std::string>& assets, const U discRate, const std::vector<U>& repoSpreads,
                   const std::vector<U>& spots, const std::vector<Time>& divDates, const matrix<U>& divs,
                   const std::vector<U>& atms, const std::vector<U>& skews, const matrix<U>& correl, const U& lambda)
        : myNumAssets(assets.size()), myAssetNames(assets), myDiscRate(discRate), myDivDates(divDates), myDivs(divs),
          myLambda(lambda)
    {
        const size_t n = assets.size();
        auto copyTo = [n](const std::vector<U>& src, std::vector<T>& dst) { dst.resize(n); for (size_t i = 0; i < n; ++i) dst[i] = T(src[i]); };
        copyTo(spots, mySpots); copyTo(repoSpreads, myRepoSpreads); copyTo(atms, myAtms); copyTo(skews, mySkews);
        myCorrel = correl;
        myCorrel.resize(n, n);

        const size_t numParams = 1 + 2 * n + n * divDates.size() + 2 * n + n * (n - 1) / 2 + 1;
        myParameters.resize(numParams);
        myParameterLabels.resize(numParams);
        size_t p = 0;
        myParameterLabels[p++] = "disc rate";
        for (size_t i = 0; i < n; ++i) myParameterLabels[p++] = "spot " + myAssetNames[i];
        for (size_t i = 0; i < n; ++i) myParameterLabels[p++] = "repo spread " + myAssetNames[i];
        for (size_t i = 0; i < myDivDates.size(); ++i)
            for (size_t j = 0; j < n; ++j) {
                std::ostringstream ost;
                ost << std::setprecision(2) << std::fixed << "div " << myAssetNames[j] << " " << myDivDates[i];
                myParameterLabels[p++] = ost.str();
            }
        for (size_t i = 0; i < n; ++i) myParameterLabels[p++] = "ATM " + myAssetNames[i];
        for (size_t i = 0; i < n; ++i) myParameterLabels[p++] = "skew " + myAssetNames[i];
        for (size_t i = 1; i < n; ++i)
            for (size_t j = 0; j < i; ++j) myParameterLabels[p++] = "correl " + myAssetNames[i] + " " + myAssetNames[j];
        myParameterLabels[p] = "lambda";
        setParamPointers();
    }

    const size_t numAssets() const override { return myNumAssets; }
    const std::vector<std::string>& assetNames() const override { return myAssetNames; }
    const T rate() const { return myDiscRate; }
    const std::vector<T>& spots() const { return mySpots; }
    const std::vector<T>& repoSpreads() const { return myRepoSpreads; }
    const std::vector<Time>& divDates() const { return myDivDates; }
    const matrix<T>& divs() const { return myDivs; }
    const std::vector<T>& atms() const { return myAtms; }
    const std::vector<T>& skews() const { return mySkews; }
    const matrix<T>& correl() const { return myCorrel; }
    const T lambda() const { return myLambda; }
    const std::vector<Dynamics>& dynamics() const { return myDynamics; }
    const std::vector<T>& alphas() const { return myAlphas; }
    const std::vector<T>& betas() const { return myBetas; }
    const matrix<T>& chol() const { return myChol; }
    const std::vector<Time>& simulationTimeline() const { return myTimeline; }

    const std::vector<T*>& parameters() override { return myParameters; }
    const std::vector<std::string>& parameterLabels() const override { return myParameterLabels; }

    std::unique_ptr<Model<T>> clone() const override
    {
        auto c = std::make_unique<MultiDisplaced<T>>(*this);
        c->setParamPointers();
        return c;
    }

    void allocate(const std::vector<Time>& productTimeline, const std::vector<SampleDef>& defline) override
    {
        const size_t A = myNumAssets;
        myRepoRates.resize(A); myAlphas.resize(A); myBetas.resize(A); myDynamics.resize(A);
        myUsedCorrel.resize(A, A); myChol.resize(A, A);
        // simulation timeline = today + the product's dates after today (mcMdlMultiDisplaced.h:372-380)
        myTimeline.clear();
        myTimeline.push_back(systemTime);
        for (const auto& time : productTimeline) if (time > systemTime) myTimeline.push_back(time);
        myTodayOnTimeline = (productTimeline[0] == systemTime);
        const size_t D = myTimeline.size() - 1, n = productTimeline.size();
        myStds.resize(D, A); myDrifts.resize(D, A); myDynFwdFacts.resize(D, A);
        myNumeraires.resize(n); myDiscounts.resize(n); myLibors.resize(n); myForwardFactors.resize(n);
        for (size_t j = 0; j < n; ++j) {
            myDiscounts[j].resize(defline[j].discountMats.size());
            myLibors[j].resize(defline[j].liborDefs.size());
            myForwardFactors[j].resize(A);
            for (size_t k = 0; k < A; ++k) myForwardFactors[j][k].resize(defline[j].forwardMats[k].size());
        }
    }

    void init(const std::vector<Time>& productTimeline, const std::vector<SampleDef>& defline) override
    {
        const size_t A = myNumAssets;
        for (size_t a = 0; a < A; ++a) myRepoRates[a] = myDiscRate - myRepoSpreads[a];

        // displacement out of ATM and skew, four regimes (mcMdlMultiDisplaced.h:483-508)
        for (size_t a = 0; a < A; ++a) {
            myBetas[a] = myAtms[a] + 2 * mySkews[a];
            if (fabs(mySkews[a]) < 1.0e-05) { myDynamics[a] = Lognormal; myAlphas[a] = 0.0; }
            else if (fabs(myBetas[a]) < 1.0e-05) {
                myDynamics[a] = Normal;
                myAlphas[a] = -2 * mySpots[a] * mySkews[a];
                myBetas[a] = 0.0;
            } else if (myBetas[a] > 0) {
                myDynamics[a] = Surnormal;
                myAlphas[a] = -2 * mySpots[a] / myBetas[a] * mySkews[a];
            } else {
                myDynamics[a] = Subnormal;
                myBetas[a] *= -1.0;
                myAlphas[a] = -2 * mySpots[a] / myBetas[a] * mySkews[a];
            }
        }

        // full matrix from the lower triangle, unit diagonal; lambda shift; Cholesky (:510-528)
        for (size_t i = 0; i < A; ++i) {
            myCorrel[i][i] = 1.0;
            for (size_t j = i; j < A; ++j) myCorrel[i][j] = myCorrel[j][i];
        }
        {
            auto dst = myUsedCorrel.begin();
            for (auto src = myCorrel.begin(); src != myCorrel.end(); ++src, ++dst) *dst = myLambda * (1.0 - *src) + *src;
        }
        choldc(myUsedCorrel, myChol);

        // per step: forward factors of the dynamics, stds, drifts (:531-563)
        const size_t D = myTimeline.size() - 1;
        for (size_t i = 0; i < D; ++i) {
            const double dt = myTimeline[i + 1] - myTimeline[i];
            for (size_t a = 0; a < A; ++a) {
                myDynFwdFacts[i][a] = forwardFactor(myTimeline[i], myTimeline[i + 1], a);
                if (myDynamics[a] == Lognormal) {
                    myStds[i][a] = myBetas[a] * std::sqrt(dt);
                    myDrifts[i][a] = -0.5 * myStds[i][a] * myStds[i][a];
                } else if (myDynamics[a] == Normal) {
                    myStds[i][a] = myAlphas[a] * std::sqrt(dt);
                    myDrifts[i][a] = T(0.0);
                } else if (myDynamics[a] == Surnormal) {
                    myStds[i][a] = myBetas[a] * std::sqrt(dt);
                    myDrifts[i][a] = -0.5 * myBetas[a] * myBetas[a] * dt;
                } else {
                    myStds[i][a] = -myBetas[a] * std::sqrt(dt);
                    myDrifts[i][a] = -0.5 * myBetas[a] * myBetas[a] * dt;
                }
            }
        }

        // per event date: numeraire, discounts, libors, forward factors (:567-604)
        for (size_t i = 0; i < productTimeline.size(); ++i) {
            if (defline[i].numeraire) myNumeraires[i] = exp(myDiscRate * productTimeline[i]);
            for (size_t j = 0; j < defline[i].discountMats.size(); ++j)
                myDiscounts[i][j] = exp(-myDiscRate * (defline[i].discountMats[j] - productTimeline[i]));
            for (size_t j = 0; j < defline[i].liborDefs.size(); ++j) {
                const double dt = defline[i].liborDefs[j].end - defline[i].liborDefs[j].start;
                myLibors[i][j] = (exp(myDiscRate * dt) - 1.0) / dt;
            }
            for (size_t a = 0; a < A; ++a)
                for (size_t j = 0; j < defline[i].forwardMats[a].size(); ++j)
                    myForwardFactors[i][a][j] = forwardFactor(productTimeline[i], defline[i].forwardMats[a][j], a);
        }
    }

    size_t simDim() const override { return myNumAssets * (myTimeline.size() - 1); }

    // Device image.  Adjoint layout (cf_dlm.cuh):
    //   spots [A] | alphas [A] | chol [A][A] | dynFwd [D][A] | drifts [D][A] | stds [D][A] | numeraires [E] | ff [E][A]
    bool deviceImage(ModelImage& img, const std::vector<Time>& productTimeline, const std::vector<SampleDef>& defline) override
    {
        const size_t A = myNumAssets, D = myTimeline.size() - 1, E = productTimeline.size();
        if (E != D + (myTodayOnTimeline ? 1 : 0)) return false;
        bool anyNum = false;
        for (const auto& def : defline) {
            if (!def.liborDefs.empty() || !def.discountMats.empty()) return false;     // multi-asset products read forwards and the numeraire only
            if (def.forwardMats.size() != A) return false;
            for (const auto& mats : def.forwardMats) if (mats.size() != 1) return false;
            anyNum = anyNum || def.numeraire;
        }
        for (const auto& def : defline) if (def.numeraire != anyNum) return false;
        img = ModelImage();
        img.isEvent.assign(D + 1, 1);
        img.isEvent[0] = myTodayOnTimeline ? 1 : 0;
        auto flat = [](std::vector<double>& dst, const matrix<T>& src) {
            dst.clear();
            for (auto it = src.begin(); it != src.end(); ++it) dst.push_back(cfValue(*it));
        };
        dlmSpots.resize(A); dlmAlphas.resize(A); dlmDyn.resize(A);
        for (size_t a = 0; a < A; ++a) { dlmSpots[a] = cfValue(mySpots[a]); dlmAlphas[a] = cfValue(myAlphas[a]); dlmDyn[a] = int32_t(myDynamics[a]); }
        flat(dlmChol, myChol); flat(dlmDynFwd, myDynFwdFacts); flat(dlmDrifts, myDrifts); flat(dlmStds, myStds);
        dlmFf.resize(E * A);
        for (size_t e = 0; e < E; ++e) for (size_t a = 0; a < A; ++a) dlmFf[e * A + a] = cfValue(myForwardFactors[e][a][0]);
        img.numeraires.clear();
        if (anyNum) for (size_t e = 0; e < E; ++e) img.numeraires.push_back(cfValue(myNumeraires[e]));
        cf_model& p = img.pod;
        p.kind = CF_MODEL_DISPLACED; p.n_assets = int(A); p.n_steps = int(D); p.n_events = int(E);
        p.is_event = img.isEvent.data();
        p.numeraires = anyNum ? img.numeraires.data() : nullptr;
        p.dlm_spots = dlmSpots.data(); p.dlm_chol = dlmChol.data(); p.dlm_alphas = dlmAlphas.data(); p.dlm_dynamics = dlmDyn.data();
        p.dlm_dyn_fwd = dlmDynFwd.data(); p.dlm_drifts = dlmDrifts.data(); p.dlm_stds = dlmStds.data(); p.dlm_fwd_factors = dlmFf.data();
        if constexpr (std::is_same<T, Number>::value) {
            auto& t = img.adjointTargets;
            t.clear();
            auto push = [&t](Number& x) { t.push_back(x.onTape() ? &x : nullptr); };
            for (size_t a = 0; a < A; ++a) push(mySpots[a]);
            for (size_t a = 0; a < A; ++a) push(myAlphas[a]);
            for (size_t a = 0; a < A; ++a) for (size_t k = 0; k < A; ++k) push(myChol[a][k]);
            for (size_t i = 0; i < D; ++i) for (size_t a = 0; a < A; ++a) push(myDynFwdFacts[i][a]);
            for (size_t i = 0; i < D; ++i) for (size_t a = 0; a < A; ++a) push(myDrifts[i][a]);
            for (size_t i = 0; i < D; ++i) for (size_t a = 0; a < A; ++a) push(myStds[i][a]);
            for (size_t e = 0; e < E; ++e) { if (anyNum) push(myNumeraires[e]); else t.push_back(nullptr); }
            for (size_t e = 0; e < E; ++e) for (size_t a = 0; a < A; ++a) push(myForwardFactors[e][a][0]);
        }
        return true;
    }

private:
    // flat copies kept alive for the device image
    std::vector<double>  dlmSpots, dlmAlphas, dlmChol, dlmDynFwd, dlmDrifts, dlmStds, dlmFf;
    std::vector<int32_t> dlmDyn;
};
