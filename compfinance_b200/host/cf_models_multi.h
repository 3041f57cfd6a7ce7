// cf_models_multi.h -- host side of the multi-asset displaced-lognormal model (mcMdlMultiDisplaced.h).
//
// INTERFACE-MANDATED: the class name, the constructor's argument list (:123-214, the order store.h:75-97 passes), the
// Model<T> virtuals, the ORDER of the parameters = the order of the risk report (:216-265: disc rate, spots, repo
// spreads, divs [date][asset], ATMs, skews, lower-triangle correls, lambda) with their label texts, and the arithmetic of
// init() (:474-606) in its floating-point order, because its tables feed the device and its tape gives the risks that
// the shipped spreadsheets pin (AutocallPricer.xlsx, testDLM.xlsx) -- including the 1e-5 thresholds of the dynamics
// selection and the 1e-15 guards of the Cholesky decomposition (choldc.h:7-50).
// OWN STRUCTURE: one enumeration of the parameters (forEachParameter) serves labels and pointers alike; init() is a
// sequence of named stages; the device image and its adjoint targets are produced by the same section writer, so the
// flat layout of cf_dlm.cuh and the tape nodes it maps to cannot drift apart.  The paths run in cf_dlm.cuh.
#pragma once

#include <iterator>

#include "cf_base.h"

// Cholesky factor of a symmetric matrix, row by row (choldc.h:7-50): a pivot below -1e-15 throws, one below 1e-15
// counts as zero, and the column under a zero pivot is zero.
template <class T>
inline void choldc(const matrix<T>& in, matrix<T>& out)
{
    const int n = int(in.rows());
    for (T& cell : out) cell = T(0.0);
    // in[i][j] minus the inner product of the already known parts of rows i and j
    const auto residual = [&](const int i, const int j) {
        T r = in[i][j];
        for (int k = 0; k < j; ++k) r -= out[i][k] * out[j][k];
        return r;
    };
    for (int i = 0; i < n; ++i) {
        for (int j = 0; j < i; ++j) {
            T r = residual(i, j);
            if (fabs(out[j][j]) < 1.0e-15) out[i][j] = 0.0;
            else out[i][j] = r / out[j][j];
        }
        T pivot = residual(i, i);
        if (pivot < -1.0e-15) throw std::runtime_error("choldc : matrix not positive definite");
        if (pivot < 1.0e-15) pivot = 0.0;
        out[i][i] = sqrt(pivot);
    }
}

template <class T>
class MultiDisplaced : public Model<T>
{
public:
    enum Dynamics { Lognormal = 0, Normal = 1, Surnormal = 2, Subnormal = 3 };

    template <class U>
    MultiDisplaced(const std::vector<std::string>& assets, const U discRate, const std::vector<U>& repoSpreads,
                   const std::vector<U>& spots, const std::vector<Time>& divDates, const matrix<U>& divs,
                   const std::vector<U>& atms, const std::vector<U>& skews, const matrix<U>& correl, const U& lambda)
        : names(assets), discRate_(discRate), divDates_(divDates), divs_(divs), lambda_(lambda)
    {
        const size_t n = names.size();
        const auto convert = [n](const std::vector<U>& from) {
            std::vector<T> to(n);
            for (size_t i = 0; i < n; ++i) to[i] = T(from[i]);
            return to;
        };
        spots_ = convert(spots);
        repoSpreads_ = convert(repoSpreads);
        atms_ = convert(atms);
        skews_ = convert(skews);
        correl_ = correl;
        correl_.resize(n, n);
        forEachParameter([this](T&, const std::string& label) { labels.push_back(label); });
        bindParameters();
    }

    // ---- Model<T>
    const size_t numAssets() const override { return names.size(); }
    const std::vector<std::string>& assetNames() const override { return names; }
    const std::vector<T*>& parameters() override { return pointers; }
    const std::vector<std::string>& parameterLabels() const override { return labels; }
    size_t simDim() const override { return names.size() * (timeline.size() - 1); }

    std::unique_ptr<Model<T>> clone() const override
    {
        auto copy = std::make_unique<MultiDisplaced<T>>(*this);
        copy->bindParameters();                 // the copied pointers still address the original
        return copy;
    }

    // simulation timeline = today + the product's dates after today (mcMdlMultiDisplaced.h:359-417)
    void allocate(const std::vector<Time>& productTimeline, const std::vector<SampleDef>& defline) override
    {
        const size_t A = names.size(), E = productTimeline.size();
        timeline.assign(1, systemTime);
        std::copy_if(productTimeline.begin(), productTimeline.end(), std::back_inserter(timeline),
                     [](const Time t) { return t > systemTime; });
        todayIsEvent = productTimeline[0] == systemTime;
        const size_t D = timeline.size() - 1;

        repoRates.resize(A); alphas_.resize(A); betas_.resize(A); dynamics_.resize(A);
        usedCorrel.resize(A, A); chol_.resize(A, A);
        stds.resize(D, A); drifts.resize(D, A); stepFwd.resize(D, A);
        numeraires.resize(E); discounts.resize(E); libors.resize(E); eventFwd.resize(E);
        for (size_t e = 0; e < E; ++e) {
            discounts[e].resize(defline[e].discountMats.size());
            libors[e].resize(defline[e].liborDefs.size());
            eventFwd[e].resize(A);
            for (size_t a = 0; a < A; ++a) eventFwd[e][a].resize(defline[e].forwardMats[a].size());
        }
    }

    // mcMdlMultiDisplaced.h:474-606; on the host tape when T = Number (the part of the reference's tape before the mark)
    void init(const std::vector<Time>& productTimeline, const std::vector<SampleDef>& defline) override
    {
        stageRepoRates();
        stageDisplacements();
        stageCorrelation();
        stageSteps();
        stageEvents(productTimeline, defline);
    }

    // ---- read access
    const T rate() const { return discRate_; }
    const std::vector<T>& spots() const { return spots_; }
    const std::vector<T>& repoSpreads() const { return repoSpreads_; }
    const std::vector<Time>& divDates() const { return divDates_; }
    const matrix<T>& divs() const { return divs_; }
    const std::vector<T>& atms() const { return atms_; }
    const std::vector<T>& skews() const { return skews_; }
    const matrix<T>& correl() const { return correl_; }
    const T lambda() const { return lambda_; }
    const std::vector<Dynamics>& dynamics() const { return dynamics_; }
    const std::vector<T>& alphas() const { return alphas_; }
    const std::vector<T>& betas() const { return betas_; }
    const matrix<T>& chol() const { return chol_; }
    const std::vector<Time>& simulationTimeline() const { return timeline; }

    // Device image.  Flat layout of the tables = layout of the adjoint vector (cf_dlm.cuh):
    //   spots [A] | alphas [A] | chol [A][A] | dynFwd [D][A] | drifts [D][A] | stds [D][A] | numeraires [E] | ff [E][A]
    bool deviceImage(ModelImage& img, const std::vector<Time>& productTimeline, const std::vector<SampleDef>& defline) override
    {
        const size_t A = names.size(), D = timeline.size() - 1, E = productTimeline.size();
        if (E != D + (todayIsEvent ? 1 : 0)) return false;
        // the multi-asset products read one forward per asset and, all of them or none, the numeraire
        size_t withNumeraire = 0;
        for (const SampleDef& def : defline) {
            if (!def.liborDefs.empty() || !def.discountMats.empty() || def.forwardMats.size() != A) return false;
            for (const auto& mats : def.forwardMats) if (mats.size() != 1) return false;
            withNumeraire += def.numeraire ? 1 : 0;
        }
        if (withNumeraire != 0 && withNumeraire != defline.size()) return false;
        const bool hasNumeraire = withNumeraire != 0;

        img = ModelImage();
        img.isEvent.assign(D + 1, 1);
        img.isEvent[0] = todayIsEvent ? 1 : 0;
        // a section of the layout: values into `flat`, the tape node of each value (AAD) into the adjoint targets
        const auto section = [&img](std::vector<double>& flat, const size_t count, const auto& cell) {
            flat.resize(count);
            for (size_t k = 0; k < count; ++k) {
                T& x = cell(k);
                flat[k] = cfValue(x);
                if constexpr (std::is_same<T, Number>::value) img.adjointTargets.push_back(x.onTape() ? &x : nullptr);
            }
        };
        section(flatSpots, A, [&](const size_t k) -> T& { return spots_[k]; });
        section(flatAlphas, A, [&](const size_t k) -> T& { return alphas_[k]; });
        section(flatChol, A * A, [&](const size_t k) -> T& { return chol_[k / A][k % A]; });
        section(flatStepFwd, D * A, [&](const size_t k) -> T& { return stepFwd[k / A][k % A]; });
        section(flatDrifts, D * A, [&](const size_t k) -> T& { return drifts[k / A][k % A]; });
        section(flatStds, D * A, [&](const size_t k) -> T& { return stds[k / A][k % A]; });
        if (hasNumeraire) section(img.numeraires, E, [&](const size_t k) -> T& { return numeraires[k]; });
        else if constexpr (std::is_same<T, Number>::value) img.adjointTargets.insert(img.adjointTargets.end(), E, nullptr);
        section(flatEventFwd, E * A, [&](const size_t k) -> T& { return eventFwd[k / A][k % A][0]; });
        flatDynamics.resize(A);
        for (size_t a = 0; a < A; ++a) flatDynamics[a] = int32_t(dynamics_[a]);

        cf_model& pod = img.pod;
        pod.kind = CF_MODEL_DISPLACED;
        pod.n_assets = int(A); pod.n_steps = int(D); pod.n_events = int(E);
        pod.is_event = img.isEvent.data();
        pod.numeraires = hasNumeraire ? img.numeraires.data() : nullptr;
        pod.dlm_spots = flatSpots.data(); pod.dlm_alphas = flatAlphas.data(); pod.dlm_chol = flatChol.data();
        pod.dlm_dynamics = flatDynamics.data();
        pod.dlm_dyn_fwd = flatStepFwd.data(); pod.dlm_drifts = flatDrifts.data(); pod.dlm_stds = flatStds.data();
        pod.dlm_fwd_factors = flatEventFwd.data();
        return true;
    }

private:
    // ---- parameters
    std::vector<std::string> names;
    T                        discRate_;
    std::vector<T>           spots_, repoSpreads_;
    std::vector<Time>        divDates_;
    matrix<T>                divs_;              // [div date][asset]
    std::vector<T>           atms_, skews_;
    matrix<T>                correl_;            // the parameters are the lower triangle
    T                        lambda_;
    std::vector<T*>          pointers;
    std::vector<std::string> labels;

    // ---- what init() derives from them
    std::vector<T>        repoRates, alphas_, betas_;
    std::vector<Dynamics> dynamics_;
    matrix<T>             usedCorrel, chol_;
    std::vector<Time>     timeline;              // today + event dates after today
    bool                  todayIsEvent = false;
    matrix<T>             stds, drifts, stepFwd;            // [step][asset]
    std::vector<T>        numeraires;                       // [event]
    std::vector<std::vector<T>> discounts, libors;          // [event][...]
    std::vector<std::vector<std::vector<T>>> eventFwd;      // [event][asset][maturity]

    // ---- flat copies the device image points into
    std::vector<double>  flatSpots, flatAlphas, flatChol, flatStepFwd, flatDrifts, flatStds, flatEventFwd;
    std::vector<int32_t> flatDynamics;

    // Every parameter with its label, in the order of the risk report (mcMdlMultiDisplaced.h:216-265)
    template <class F>
    void forEachParameter(F&& visit)
    {
        const size_t n = names.size();
        visit(discRate_, "disc rate");
        for (size_t a = 0; a < n; ++a) visit(spots_[a], "spot " + names[a]);
        for (size_t a = 0; a < n; ++a) visit(repoSpreads_[a], "repo spread " + names[a]);
        for (size_t d = 0; d < divDates_.size(); ++d)
            for (size_t a = 0; a < n; ++a) {
                std::ostringstream text;
                text << std::setprecision(2) << std::fixed << "div " << names[a] << " " << divDates_[d];
                visit(divs_[d][a], text.str());
            }
        for (size_t a = 0; a < n; ++a) visit(atms_[a], "ATM " + names[a]);
        for (size_t a = 0; a < n; ++a) visit(skews_[a], "skew " + names[a]);
        for (size_t a = 1; a < n; ++a)
            for (size_t b = 0; b < a; ++b) visit(correl_[a][b], "correl " + names[a] + " " + names[b]);
        visit(lambda_, "lambda");
    }
    void bindParameters()
    {
        pointers.clear();
        forEachParameter([this](T& p, const std::string&) { pointers.push_back(&p); });
    }

    // prod(1 - div) over the dividend dates in [from, to) times exp(repo (to - from)) (mcMdlMultiDisplaced.h:445-470).
    // The reference's single-asset variant (:421-442) never advances its dividend index; the shipped products only
    // reach it with to == from, where both give exp(0) times an empty product.
    T forwardFactor(const Time from, const Time to, const size_t asset) const
    {
        T keep = T(1.0);
        for (auto d = std::lower_bound(divDates_.begin(), divDates_.end(), from); d != divDates_.end() && *d < to; ++d)
            keep *= 1.0 - divs_[size_t(d - divDates_.begin())][asset];
        return keep * exp(repoRates[asset] * (to - from));
    }

    void stageRepoRates()
    {
        for (size_t a = 0; a < names.size(); ++a) repoRates[a] = discRate_ - repoSpreads_[a];
    }

    // beta = ATM + 2 skew decides the dynamics; alpha is the displacement (mcMdlMultiDisplaced.h:483-508)
    void stageDisplacements()
    {
        for (size_t a = 0; a < names.size(); ++a) {
            T& alpha = alphas_[a];
            T& beta = betas_[a];
            beta = atms_[a] + 2 * skews_[a];
            if (fabs(skews_[a]) < 1.0e-05) {
                dynamics_[a] = Lognormal;
                alpha = 0.0;
            } else if (fabs(beta) < 1.0e-05) {
                dynamics_[a] = Normal;
                alpha = -2 * spots_[a] * skews_[a];
                beta = 0.0;
            } else {
                dynamics_[a] = beta > 0 ? Surnormal : Subnormal;
                if (dynamics_[a] == Subnormal) beta *= -1.0;
                alpha = -2 * spots_[a] / beta * skews_[a];
            }
        }
    }

    // the full matrix out of the lower triangle with a unit diagonal, shifted towards 1 by lambda, factorised (:510-528)
    void stageCorrelation()
    {
        const size_t A = names.size();
        for (size_t a = 0; a < A; ++a) {
            correl_[a][a] = 1.0;
            for (size_t b = a; b < A; ++b) correl_[a][b] = correl_[b][a];
        }
        std::transform(correl_.begin(), correl_.end(), usedCorrel.begin(),
                       [this](const T& rho) { return lambda_ * (1.0 - rho) + rho; });
        choldc(usedCorrel, chol_);
    }

    // per step and asset: forward factor of the dynamics, std, drift (:531-563)
    void stageSteps()
    {
        for (size_t i = 0; i + 1 < timeline.size(); ++i) {
            const double dt = timeline[i + 1] - timeline[i];
            for (size_t a = 0; a < names.size(); ++a) {
                stepFwd[i][a] = forwardFactor(timeline[i], timeline[i + 1], a);
                T& sd = stds[i][a];
                switch (dynamics_[a]) {
                    case Lognormal:
                        sd = betas_[a] * std::sqrt(dt);
                        drifts[i][a] = -0.5 * sd * sd;
                        break;
                    case Normal:
                        sd = alphas_[a] * std::sqrt(dt);
                        drifts[i][a] = T(0.0);
                        break;
                    case Surnormal:
                        sd = betas_[a] * std::sqrt(dt);
                        drifts[i][a] = -0.5 * betas_[a] * betas_[a] * dt;
                        break;
                    case Subnormal:
                        sd = -betas_[a] * std::sqrt(dt);
                        drifts[i][a] = -0.5 * betas_[a] * betas_[a] * dt;
                        break;
                }
            }
        }
    }

    // per event date: numeraire, discounts, libors, forward factors of every asset to its maturities (:567-604)
    void stageEvents(const std::vector<Time>& productTimeline, const std::vector<SampleDef>& defline)
    {
        for (size_t e = 0; e < productTimeline.size(); ++e) {
            const Time now = productTimeline[e];
            const SampleDef& def = defline[e];
            if (def.numeraire) numeraires[e] = exp(discRate_ * now);
            for (size_t j = 0; j < def.discountMats.size(); ++j) discounts[e][j] = exp(-discRate_ * (def.discountMats[j] - now));
            for (size_t j = 0; j < def.liborDefs.size(); ++j) {
                const double dt = def.liborDefs[j].end - def.liborDefs[j].start;
                libors[e][j] = (exp(discRate_ * dt) - 1.0) / dt;
            }
            for (size_t a = 0; a < names.size(); ++a)
                for (size_t j = 0; j < def.forwardMats[a].size(); ++j) eventFwd[e][a][j] = forwardFactor(now, def.forwardMats[a][j], a);
        }
    }
};
