// host_stage.cpp -- times the path-independent host stage of dupireAADRisk (config 3) without a GPU:
// clone, allocate, init() on the host tape, device images.  g++ -std=c++17 -O2 -I../../include -I../../compfinance_b200/host
// host_stage.cpp -L../../compfinance_b200/lib -lcf_b200 -Wl,-rpath,...   (tools/micro/run_host_stage.sh)
#include "cf_main.h"
#include <chrono>
#include <cstdio>

int main()
{
    std::vector<double> spots, times;
    for (int i = 0; i < 30; ++i) spots.push_back(55 + 5.0 * i);
    for (int j = 1; j <= 36; ++j) times.push_back(j / 12.0);
    matrix<double> vols(30, 36);
    for (int i = 0; i < 30; ++i) for (int j = 0; j < 36; ++j) { double l = std::log(spots[i] / 100); vols[i][j] = 0.15 + 0.1 * l * l + 0.02 * times[j]; }
    putDupire(100.0, spots, times, vols, 0.25, "m");
    putBarrier(120.0, 150.0, 3.0, 1.0 / 52, 0.01, false, "p");
    const Model<Number>* mdl = getModel<Number>("m");
    const Product<Number>* prd = getProduct<Number>("p");
    Sobol rng;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto us = [](auto a, auto b) { return std::chrono::duration<double, std::micro>(b - a).count(); };
    double best[5] = {1e9, 1e9, 1e9, 1e9, 1e9};
    for (int rep = 0; rep < 200; ++rep) {
        auto t0 = now();
        auto c = mdl->clone();
        auto t1 = now();
        c->allocate(prd->timeline(), prd->defline());
        auto t2 = now();
        Tape& tape = *Number::tape;
        tape.clear();
        c->putParametersOnTape();
        c->init(prd->timeline(), prd->defline());
        tape.mark();
        auto t3 = now();
        CfDeviceSetup s;
        cfBuildImages(*prd, *c, rng, s);
        auto t4 = now();
        c.reset();
        auto t5 = now();
        double v[5] = {us(t0, t1), us(t1, t2), us(t2, t3), us(t3, t4), us(t4, t5)};
        for (int k = 0; k < 5; ++k) best[k] = std::min(best[k], v[k]);
    }
    std::printf("clone %.1f us, allocate %.1f us, init on tape %.1f us, images %.1f us, destroy %.1f us\n", best[0], best[1], best[2], best[3], best[4]);
    return 0;
}
