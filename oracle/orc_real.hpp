// TEST INFRASTRUCTURE ONLY — CPU oracle for the batched C172 RK4 path.
// Nothing under oracle/ is linked, imported or executed by the product path
// (flight.jl_b200/); only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may use it, and only as the checker.
//
// Scalar abstraction: the oracle is compiled twice from the same sources,
//   -DORC_COUNT=0 : real == double                         (liborc.so)
//   -DORC_COUNT=1 : real == Counted, an op-counting double (liborc_count.so)
// The counting build fixes the "algorithmic FLOPs per step" figure of record
// under SURVEY.md §8(d)'s convention: add/sub/mul/compare-select = 1, every
// division, sqrt and transcendental call = 1 (an a*b+c pair therefore counts 2,
// which is the FMA = 2 rule).
#pragma once
#include <cmath>
#include <cstdint>

#ifndef ORC_COUNT
#define ORC_COUNT 0
#endif

#if ORC_COUNT
struct OpCounts {
    uint64_t plain = 0;   // add, sub, mul, neg-free compare/select, abs, min, max
    uint64_t div = 0;
    uint64_t sqrt = 0;
    uint64_t trans = 0;   // sin, cos, tan, atan, atan2, asin, acos, exp, log, pow
    uint64_t total() const { return plain + div + sqrt + trans; }
};
extern thread_local OpCounts g_ops;

struct Counted {
    double v;
    Counted() : v(0.0) {}
    Counted(double x) : v(x) {}
    Counted(int x) : v(x) {}
    explicit operator double() const { return v; }
};
inline Counted operator+(Counted a, Counted b) { g_ops.plain++; return Counted(a.v + b.v); }
inline Counted operator-(Counted a, Counted b) { g_ops.plain++; return Counted(a.v - b.v); }
inline Counted operator*(Counted a, Counted b) { g_ops.plain++; return Counted(a.v * b.v); }
inline Counted operator/(Counted a, Counted b) { g_ops.div++; return Counted(a.v / b.v); }
inline Counted operator-(Counted a) { return Counted(-a.v); }  // sign flip: free
inline Counted& operator+=(Counted& a, Counted b) { a = a + b; return a; }
inline Counted& operator-=(Counted& a, Counted b) { a = a - b; return a; }
inline Counted& operator*=(Counted& a, Counted b) { a = a * b; return a; }
inline Counted& operator/=(Counted& a, Counted b) { a = a / b; return a; }
#define ORC_CMP(op) inline bool operator op(Counted a, Counted b) { g_ops.plain++; return a.v op b.v; }
ORC_CMP(<) ORC_CMP(>) ORC_CMP(<=) ORC_CMP(>=) ORC_CMP(==) ORC_CMP(!=)
#undef ORC_CMP
#define ORC_MIX(op, R) \
    inline R operator op(Counted a, double b) { return a op Counted(b); } \
    inline R operator op(double a, Counted b) { return Counted(a) op b; } \
    inline R operator op(Counted a, int b) { return a op Counted(b); }    \
    inline R operator op(int a, Counted b) { return Counted(a) op b; }
ORC_MIX(+, Counted) ORC_MIX(-, Counted) ORC_MIX(*, Counted) ORC_MIX(/, Counted)
ORC_MIX(<, bool) ORC_MIX(>, bool) ORC_MIX(<=, bool) ORC_MIX(>=, bool) ORC_MIX(==, bool) ORC_MIX(!=, bool)
#undef ORC_MIX
#define ORC_T1(name) inline Counted name(Counted a) { g_ops.trans++; return Counted(std::name(a.v)); }
ORC_T1(sin) ORC_T1(cos) ORC_T1(tan) ORC_T1(atan) ORC_T1(asin) ORC_T1(acos) ORC_T1(exp) ORC_T1(log)
#undef ORC_T1
inline Counted sqrt(Counted a) { g_ops.sqrt++; return Counted(std::sqrt(a.v)); }
inline Counted atan2(Counted a, Counted b) { g_ops.trans++; return Counted(std::atan2(a.v, b.v)); }
inline Counted pow(Counted a, Counted b) { g_ops.trans++; return Counted(std::pow(a.v, b.v)); }
inline Counted pow(Counted a, double b) { g_ops.trans++; return Counted(std::pow(a.v, b)); }
inline Counted fabs(Counted a) { g_ops.plain++; return Counted(std::fabs(a.v)); }
inline Counted floor(Counted a) { g_ops.plain++; return Counted(std::floor(a.v)); }
inline Counted fmin(Counted a, Counted b) { g_ops.plain++; return Counted(a.v < b.v ? a.v : b.v); }
inline Counted fmax(Counted a, Counted b) { g_ops.plain++; return Counted(a.v > b.v ? a.v : b.v); }
typedef Counted real;
inline double val(real a) { return a.v; }
#else
typedef double real;
inline double val(real a) { return a; }
using std::sin; using std::cos; using std::tan; using std::atan; using std::atan2;
using std::asin; using std::acos; using std::exp; using std::log; using std::pow;
using std::sqrt; using std::fabs; using std::floor; using std::fmin; using std::fmax;
#endif

// Julia's min/max/clamp/sign/abs on Float64 (NaN handling is irrelevant on this path).
inline real rmin(real a, real b) { return fmin(a, b); }
inline real rmax(real a, real b) { return fmax(a, b); }
inline real rclamp(real x, real lo, real hi) { return rmin(rmax(x, lo), hi); }
inline real rabs(real a) { return fabs(a); }
inline int isign(real a) { return (a > 0.0) - (a < 0.0); }
inline real sq(real a) { return a * a; }
inline real cube(real a) { return a * a * a; }
