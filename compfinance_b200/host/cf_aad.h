// cf_aad.h -- a compact host reverse-AD scalar, used ONLY for the path-independent stage.
//
// In the reference every operation of a simulation is recorded on the tape (AAD.h, AADTape.h,
// AADNode.h, AADExpr.h).  Here the per-path part of the tape is replaced by hand-written adjoint
// kernels on the GPU; what remains on the host is the part the reference records BEFORE
// tape.mark() (mcBase.h:537-561): putting the parameters on tape and running Model<T>::init().
// The device returns the adjoints of the init() tables summed over paths; seeding them on this
// tape and sweeping it back to the parameters is the reference's Number::propagateMarkToStart()
// (AADExpr.h:1059-1062, called at mcBase.h:518 and 721-733).  The same type differentiates the
// calibration stage of dupireSuperbucket (main.h:521-562).
//
// Semantics kept from the reference (AADExpr.h): max(x, d) has derivative 1 iff x > d strictly
// (571-583), min 1 iff x < d (586-598), comparisons look at values only (742-836), double(x) cuts
// the dependency (960-961).  This is a plain operator-overloading tape, not expression templates:
// the init stage is O(10^3..10^4) operations, once per run.
#pragma once

#include <cmath>
#include <cstddef>
#include <stdexcept>
#include <utility>
#include <vector>

struct AADNode {
    double adjoint = 0.0;
    int    nArg = 0;
    int    arg[2] = {-1, -1};
    double der[2] = {0.0, 0.0};
};

class Tape {
    std::vector<AADNode> myNodes;
    size_t myMark = 0;

public:
    using iterator = size_t;   // position on the tape

    int record(int nArg, int a0, double d0, int a1 = -1, double d1 = 0.0)
    {
        AADNode& n = myNodes.emplace_back();
        n.nArg = nArg; n.arg[0] = a0; n.der[0] = d0; n.arg[1] = a1; n.der[1] = d1;
        return int(myNodes.size()) - 1;
    }
    AADNode& node(int i) { return myNodes[size_t(i)]; }
    const AADNode& node(int i) const { return myNodes[size_t(i)]; }
    size_t size() const { return myNodes.size(); }

    void clear() { myNodes.clear(); myMark = 0; }
    void rewind() { clear(); }
    void mark() { myMark = myNodes.size(); }
    void rewindToMark() { myNodes.resize(myMark); }
    void resetAdjoints() { for (auto& n : myNodes) n.adjoint = 0.0; }

    iterator begin() const { return 0; }
    iterator end() const { return myNodes.size(); }
    iterator markIt() const { return myMark; }

    // d node / d leaf for every leaf under `node`, by walking its (small) expression DAG.
    // Used to read a table entry of init() as a sparse linear map of the parameters.
    void leafGradient(int node, double seed, std::vector<std::pair<int, double>>& out) const
    {
        const AADNode& n = myNodes[size_t(node)];
        if (n.nArg == 0) {
            for (auto& e : out) if (e.first == node) { e.second += seed; return; }
            out.emplace_back(node, seed);
            return;
        }
        for (int k = 0; k < n.nArg; ++k)
            if (n.der[k] != 0.0) leafGradient(n.arg[k], seed * n.der[k], out);
    }

    // reverse sweep over nodes [to, from], from >= to
    void propagate(iterator from, iterator to)
    {
        if (myNodes.empty()) return;
        for (size_t i = from + 1; i-- > to;) {
            const AADNode& n = myNodes[i];
            if (n.adjoint == 0.0) continue;
            for (int k = 0; k < n.nArg; ++k) myNodes[size_t(n.arg[k])].adjoint += n.der[k] * n.adjoint;
        }
    }
};

class Number {
    double myValue = 0.0;
    int    myIdx = -1;     // node on *tape, -1 = constant

    Number(double v, int idx) : myValue(v), myIdx(idx) {}

    static Number unary(double v, const Number& a, double da)
    {
        if (a.myIdx < 0) return Number(v);
        return Number(v, tape->record(1, a.myIdx, da));
    }
    static Number binary(double v, const Number& a, double da, const Number& b, double db)
    {
        if (a.myIdx < 0 && b.myIdx < 0) return Number(v);
        if (b.myIdx < 0) return Number(v, tape->record(1, a.myIdx, da));
        if (a.myIdx < 0) return Number(v, tape->record(1, b.myIdx, db));
        return Number(v, tape->record(2, a.myIdx, da, b.myIdx, db));
    }

public:
    static thread_local Tape* tape;

    // result of a differentiable unary function evaluated outside this header (normalCdf, ...)
    static Number fromUnary(double v, const Number& a, double da) { return unary(v, a, da); }
    // result of a differentiable binary expression recorded as ONE node, the way the reference's expression
    // templates flatten an expression into a single multi-argument node (AADExpr.h:873-886)
    static Number fromBinary(double v, const Number& a, double da, const Number& b, double db) { return binary(v, a, da, b, db); }

    Number() = default;
    Number(const double v) : myValue(v) {}
    Number& operator=(const double v) { myValue = v; myIdx = -1; return *this; }

    void putOnTape() { myIdx = tape->record(0, -1, 0.0); }

    double value() const { return myValue; }
    double& value() { return myValue; }
    explicit operator double() const { return myValue; }
    bool onTape() const { return myIdx >= 0; }
    int  index() const { return myIdx; }

    double& adjoint()
    {
        if (myIdx < 0) throw std::runtime_error("Number::adjoint(): not on tape");
        return tape->node(myIdx).adjoint;
    }
    double adjoint() const
    {
        if (myIdx < 0) throw std::runtime_error("Number::adjoint(): not on tape");
        return tape->node(myIdx).adjoint;
    }

    // Seed 1 on this result and sweep
    void propagateToStart() { adjoint() = 1.0; tape->propagate(size_t(myIdx), tape->begin()); }
    void propagateToMark() { adjoint() = 1.0; if (size_t(myIdx) >= tape->markIt()) tape->propagate(size_t(myIdx), tape->markIt()); }
    static void propagateMarkToStart()
    {
        if (tape->markIt() > 0) tape->propagate(tape->markIt() - 1, tape->begin());
    }
    static void propagateAdjoints(Tape::iterator from, Tape::iterator to) { tape->propagate(from, to); }

    // arithmetic
    friend Number operator+(const Number& a, const Number& b) { return binary(a.myValue + b.myValue, a, 1.0, b, 1.0); }
    friend Number operator-(const Number& a, const Number& b) { return binary(a.myValue - b.myValue, a, 1.0, b, -1.0); }
    friend Number operator*(const Number& a, const Number& b) { return binary(a.myValue * b.myValue, a, b.myValue, b, a.myValue); }
    friend Number operator/(const Number& a, const Number& b)
    {
        return binary(a.myValue / b.myValue, a, 1.0 / b.myValue, b, -a.myValue / b.myValue / b.myValue);    // AADExpr.h:175-193
    }
    friend Number operator+(const Number& a, const double b) { return unary(a.myValue + b, a, 1.0); }
    friend Number operator+(const double a, const Number& b) { return unary(a + b.myValue, b, 1.0); }
    friend Number operator-(const Number& a, const double b) { return unary(a.myValue - b, a, 1.0); }
    friend Number operator-(const double a, const Number& b) { return unary(a - b.myValue, b, -1.0); }
    friend Number operator*(const Number& a, const double b) { return unary(a.myValue * b, a, b); }
    friend Number operator*(const double a, const Number& b) { return unary(a * b.myValue, b, a); }
    friend Number operator/(const Number& a, const double b) { return unary(a.myValue / b, a, 1.0 / b); }
    friend Number operator/(const double a, const Number& b) { return unary(a / b.myValue, b, -a / b.myValue / b.myValue); }
    Number operator-() const { return unary(-myValue, *this, -1.0); }
    Number operator+() const { return *this; }
    Number& operator+=(const Number& b) { return *this = *this + b; }
    Number& operator-=(const Number& b) { return *this = *this - b; }
    Number& operator*=(const Number& b) { return *this = *this * b; }
    Number& operator/=(const Number& b) { return *this = *this / b; }
    Number& operator+=(const double b) { return *this = *this + b; }
    Number& operator-=(const double b) { return *this = *this - b; }
    Number& operator*=(const double b) { return *this = *this * b; }
    Number& operator/=(const double b) { return *this = *this / b; }

    // functions
    friend Number exp(const Number& a) { const double e = std::exp(a.myValue); return unary(e, a, e); }
    friend Number log(const Number& a) { return unary(std::log(a.myValue), a, 1.0 / a.myValue); }
    friend Number sqrt(const Number& a) { const double s = std::sqrt(a.myValue); return unary(s, a, 0.5 / s); }
    friend Number fabs(const Number& a) { return unary(std::fabs(a.myValue), a, a.myValue > 0.0 ? 1.0 : -1.0); }
    friend Number pow(const Number& a, const double p) { return unary(std::pow(a.myValue, p), a, p * std::pow(a.myValue, p - 1.0)); }
    friend Number pow(const Number& a, const Number& b)
    {
        const double v = std::pow(a.myValue, b.myValue);
        return binary(v, a, b.myValue * v / a.myValue, b, std::log(a.myValue) * v);
    }
    friend Number max(const Number& a, const Number& b)
    {
        const bool left = a.myValue > b.myValue, right = b.myValue > a.myValue;   // ties give 0 to both (AADExpr.h:215-253)
        return binary(left ? a.myValue : b.myValue, a, left ? 1.0 : 0.0, b, right ? 1.0 : 0.0);
    }
    friend Number min(const Number& a, const Number& b)
    {
        const bool left = a.myValue < b.myValue, right = b.myValue < a.myValue;
        return binary(left ? a.myValue : b.myValue, a, left ? 1.0 : 0.0, b, right ? 1.0 : 0.0);
    }
    friend Number max(const Number& a, const double b) { return unary(a.myValue > b ? a.myValue : b, a, a.myValue > b ? 1.0 : 0.0); }
    friend Number max(const double a, const Number& b) { return max(b, a); }
    friend Number min(const Number& a, const double b) { return unary(a.myValue < b ? a.myValue : b, a, a.myValue < b ? 1.0 : 0.0); }
    friend Number min(const double a, const Number& b) { return min(b, a); }

    // comparisons: values only
    friend bool operator==(const Number& a, const Number& b) { return a.myValue == b.myValue; }
    friend bool operator!=(const Number& a, const Number& b) { return a.myValue != b.myValue; }
    friend bool operator<(const Number& a, const Number& b) { return a.myValue < b.myValue; }
    friend bool operator>(const Number& a, const Number& b) { return a.myValue > b.myValue; }
    friend bool operator<=(const Number& a, const Number& b) { return a.myValue <= b.myValue; }
    friend bool operator>=(const Number& a, const Number& b) { return a.myValue >= b.myValue; }
    friend bool operator==(const Number& a, const double b) { return a.myValue == b; }
    friend bool operator!=(const Number& a, const double b) { return a.myValue != b; }
    friend bool operator<(const Number& a, const double b) { return a.myValue < b; }
    friend bool operator>(const Number& a, const double b) { return a.myValue > b; }
    friend bool operator<=(const Number& a, const double b) { return a.myValue <= b; }
    friend bool operator>=(const Number& a, const double b) { return a.myValue >= b; }
    friend bool operator<(const double a, const Number& b) { return a < b.myValue; }
    friend bool operator>(const double a, const Number& b) { return a > b.myValue; }
    friend bool operator<=(const double a, const Number& b) { return a <= b.myValue; }
    friend bool operator>=(const double a, const Number& b) { return a >= b.myValue; }
};

inline Tape& cfGlobalTape() { static thread_local Tape t; return t; }
inline thread_local Tape* Number::tape = &cfGlobalTape();

// Value extraction usable in templated code for both T = double and T = Number
inline double cfValue(const double x) { return x; }
inline double cfValue(const Number& x) { return x.value(); }

// convertCollection (AAD.h:58-75): copy a collection of T into a collection of U
template <class It1, class It2>
inline void convertCollection(It1 srcBegin, It1 srcEnd, It2 destBegin)
{
    using destType = std::remove_reference_t<decltype(*destBegin)>;
    for (; srcBegin != srcEnd; ++srcBegin, ++destBegin) *destBegin = destType(cfValue(*srcBegin));
}
