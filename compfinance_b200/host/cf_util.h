// cf_util.h -- host numerics of the path-independent stage: grid filling and 1-D interpolation.
//
// INTERFACE-MANDATED: the names and call forms fillData(original, maxDx, minDx, addBegin, addEnd) (utility.h:11-79) and
// interp<smoothStep>(xBegin, xEnd, yBegin, yEnd, x0) (interp.h:26-63), which the models and the calibration call the way
// the reference's do, the tolerance EPS, and the FLOATING-POINT ORDER of a sub-division ("int(gap / maxDx - EPS) + 1"
// equal pieces, points accumulated by repeated addition) and of a blend ("y1 + (y2 - y1) * t"): timelines and init()
// tables must come out bit for bit as the reference's.
// OWN STRUCTURE: a tolerant merge of two sorted ranges, a gap sub-divider, a bracket locator and a blend, each on its
// own; fillData and interp are compositions of them (cf_calib.h's 2-D interpolation reuses the bracket and the blend).
#pragma once

#include <algorithm>
#include <iterator>
#include <type_traits>
#include <vector>

#ifndef EPS
#define EPS 1.0e-08        // gaussians.h:8, utility.h:9
#endif

namespace cfnum {

// Union of two ascending ranges where b is "the same point" as a when neither is more than tol below the other;
// the point of the first range is the one kept.
template <class Out, class ItA, class ItB, class T>
inline void mergeWithin(ItA a, const ItA aEnd, ItB b, const ItB bEnd, const T tol, Out& out)
{
    auto before = [tol](const T lhs, const T rhs) { return lhs < rhs - tol; };
    while (a != aEnd && b != bEnd) {
        if (before(*b, *a)) out.push_back(*b++);
        else {
            if (!before(*a, *b)) ++b;
            out.push_back(*a++);
        }
    }
    out.insert(out.end(), a, aEnd);
    out.insert(out.end(), b, bEnd);
}

// Points strictly inside (from, to) that cut it into equal pieces no longer than maxStep; none closer than tol to `to`.
template <class Out, class T>
inline void subdivide(const T from, const T to, const T maxStep, const T tol, Out& out)
{
    const T gap = to - from;
    if (!(gap > maxStep)) return;
    const int pieces = int(gap / maxStep - EPS) + 1;
    const T step = gap / pieces;
    for (T at = from + step; at < to - tol; at += step) out.push_back(at);
}

// Where x0 sits among ascending knots: below all of them, above all of them, or in [knot(lower), knot(lower + 1)).
struct Bracket { enum Side { Below, Inside, Above } side; size_t lower; };

template <class ItX, class T>
inline Bracket bracket(const ItX xBegin, const ItX xEnd, const T& x0)
{
    const ItX firstAbove = std::upper_bound(xBegin, xEnd, x0);
    if (firstAbove == xEnd) return {Bracket::Above, 0};
    if (firstAbove == xBegin) return {Bracket::Below, 0};
    return {Bracket::Inside, size_t(firstAbove - xBegin) - 1};
}

// y1 -> y2 as t goes 0 -> 1, linearly or along the smooth step 3 t^2 - 2 t^3
template <bool smoothStep, class Y, class W>
inline auto blend(const Y& y1, const Y& y2, const W& t)
{
    if constexpr (smoothStep) return y1 + (y2 - y1) * t * t * (3.0 - 2 * t);
    else return y1 + (y2 - y1) * t;
}

}  // namespace cfnum

// utility.h:24-79: `original` with the points [addBegin, addEnd) merged in (equal within minDx) and every gap longer
// than maxDx cut into equal pieces
template <class CONT, class T, class IT = T*>
inline CONT fillData(const CONT& original, const T& maxDx, const T& minDx = T(0.0), IT addBegin = nullptr,
                     IT addEnd = nullptr)
{
    CONT knots;
    if (addBegin && addEnd && addBegin != addEnd) cfnum::mergeWithin(original.begin(), original.end(), addBegin, addEnd, minDx, knots);
    else knots = original;

    CONT filled;
    for (const auto& knot : knots) {
        if (!filled.empty()) {
            const T from = filled.back();          // sub-division points are appended while `from` is held by value
            cfnum::subdivide(from, T(knot), maxDx, minDx, filled);
        }
        filled.push_back(knot);
    }
    return filled;
}

// interp.h:26-63: flat outside the knots, linear (or smooth-step) between two of them
template <bool smoothStep = false, class ITX, class ITY, class T>
inline auto interp(ITX xBegin, ITX xEnd, ITY yBegin, ITY yEnd, const T& x0) -> std::remove_reference_t<decltype(*yBegin)>
{
    const cfnum::Bracket b = cfnum::bracket(xBegin, xEnd, x0);
    if (b.side == cfnum::Bracket::Above) return *(yEnd - 1);
    if (b.side == cfnum::Bracket::Below) return *yBegin;
    const auto xLeft = xBegin[b.lower];
    const auto t = (x0 - xLeft) / (xBegin[b.lower + 1] - xLeft);
    return cfnum::blend<smoothStep>(yBegin[b.lower], yBegin[b.lower + 1], t);
}
