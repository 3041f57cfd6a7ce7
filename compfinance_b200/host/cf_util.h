// cf_util.h -- host numeric utilities used by the path-independent stage:
// fillData (utility.h:11-79), interp (interp.h:26-63).  Templated so that they also run on the
// host AD type when init() is recorded.
#pragma once

#include <algorithm>
#include <iterator>
#include <vector>

#ifndef EPS
#define EPS 1.0e-08        // gaussians.h:8, utility.h:9
#endif

// Fill a sorted collection so that consecutive points are at most maxDx apart, after merging in
// the extra points [addBegin, addEnd) with equality tolerance minDx (utility.h:24-79).
template <class CONT, class T, class IT = T*>
inline CONT fillData(const CONT& original, const T& maxDx, const T& minDx = T(0.0), IT addBegin = nullptr,
                     IT addEnd = nullptr)
{
    CONT filled, added;
    const size_t addPoints = addBegin && addEnd ? std::distance(addBegin, addEnd) : 0;
    if (addPoints > 0)
        std::set_union(original.begin(), original.end(), addBegin, addEnd, std::back_inserter(added),
                       [minDx](const T x, const T y) { return x < y - minDx; });
    const CONT& sequence = addPoints > 0 ? added : original;

    auto it = sequence.begin();
    filled.push_back(*it);
    for (++it; it != sequence.end(); ++it) {
        const auto current = filled.back();
        const auto next = *it;
        if (next - current > maxDx) {
            const int nAdd = int((next - current) / maxDx - EPS) + 1;
            const auto spacing = (next - current) / nAdd;
            for (auto t = current + spacing; t < next - minDx; t += spacing) filled.push_back(t);
        }
        filled.push_back(next);
    }
    return filled;
}

// 1-D interpolation of ys against sorted knots xs at x0: upper_bound, flat extrapolation,
// linear (or smooth-step) inside (interp.h:26-63).
template <bool smoothStep = false, class ITX, class ITY, class T>
inline auto interp(ITX xBegin, ITX xEnd, ITY yBegin, ITY yEnd, const T& x0) -> std::remove_reference_t<decltype(*yBegin)>
{
    auto it = std::upper_bound(xBegin, xEnd, x0);
    if (it == xEnd) return *(yEnd - 1);
    if (it == xBegin) return *yBegin;
    const size_t n = std::distance(xBegin, it) - 1;
    auto x1 = xBegin[n];
    auto y1 = yBegin[n];
    auto x2 = xBegin[n + 1];
    auto y2 = yBegin[n + 1];
    auto t = (x0 - x1) / (x2 - x1);
    if constexpr (smoothStep) return y1 + (y2 - y1) * t * t * (3.0 - 2 * t);
    else return y1 + (y2 - y1) * t;
}
