// TEST INFRASTRUCTURE ONLY (see orc_real.hpp).
// Linear interpolation with the semantics of Interpolations.jl 0.16 (a third-party
// dependency of the reference, compat "0.16" in lib/FlightPhysics/Project.toml:55;
// its source is not under /root/reference, so its published algorithm is restated):
//   * "gridded"  = linear_interpolation(knots::Vector, data)  -> Gridded(Linear())
//   * "scaled"   = scale(interpolate(A, BSpline(Linear())), range)
//   * extrapolate(itp, Flat())  : evaluate at the clamped coordinate
//   * extrapolate(itp, Line())  : val(clamped) + sum_d (x_d - clamped_d) * grad_d(clamped)
// Call sites pinned: FA/c172/c172.jl:162-195, FP/piston.jl:84-189,
// FP/propellers.jl:270-273, FP/control.jl:965, FP/geodesy.jl:195.
#pragma once
#include <vector>
#include <cassert>
#include "orc_real.hpp"

enum Extrap { FLAT = 0, LINE = 1 };

struct Axis {
    std::vector<double> knots;  // explicit knots (gridded) or materialised range (scaled)
    Extrap lo = FLAT, hi = FLAT;
    Axis() {}
    Axis(const std::vector<double>& k, Extrap l, Extrap h) : knots(k), lo(l), hi(h) {}
    static Axis range(double a, double b, int n, Extrap l, Extrap h) {
        std::vector<double> k(n);
        for (int i = 0; i < n; i++) k[i] = a + (b - a) * (double)i / (double)(n - 1);
        k[n - 1] = b;
        return Axis(k, l, h);
    }
    int n() const { return (int)knots.size(); }
    // cell index i (0-based, 0..n-2) and weight w in [0,1] for a coordinate already inside bounds
    void locate(real x, int& i, real& w) const {
        int nn = n();
        int lo_i = 0, hi_i = nn - 1;  // searchsortedlast
        while (hi_i - lo_i > 1) {
            int mid = (lo_i + hi_i) / 2;
            if (x >= knots[mid]) lo_i = mid; else hi_i = mid;
        }
        i = lo_i;
        if (i > nn - 2) i = nn - 2;
        w = (x - knots[i]) / (knots[i + 1] - knots[i]);
    }
};

struct Interp1 {
    Axis ax;
    std::vector<double> y;
    real operator()(real x) const {
        real xs = rclamp(x, ax.knots.front(), ax.knots.back());
        int i; real w;
        ax.locate(xs, i, w);
        real v = (1.0 - w) * y[i] + w * y[i + 1];
        bool below = x < ax.knots.front(), above = x > ax.knots.back();
        if ((below && ax.lo == LINE) || (above && ax.hi == LINE)) {
            real g = (y[i + 1] - y[i]) / (ax.knots[i + 1] - ax.knots[i]);
            v = v + (x - xs) * g;
        }
        return v;
    }
};

struct Interp2 {
    Axis a0, a1;
    std::vector<double> y;  // y[i0 * n1 + i1]
    double at(int i, int j) const { return y[(size_t)i * a1.n() + j]; }
    real operator()(real x0, real x1) const {
        real xs0 = rclamp(x0, a0.knots.front(), a0.knots.back());
        real xs1 = rclamp(x1, a1.knots.front(), a1.knots.back());
        int i, j; real u, w;
        a0.locate(xs0, i, u);
        a1.locate(xs1, j, w);
        real f00 = at(i, j), f01 = at(i, j + 1), f10 = at(i + 1, j), f11 = at(i + 1, j + 1);
        real v = (1.0 - u) * ((1.0 - w) * f00 + w * f01) + u * ((1.0 - w) * f10 + w * f11);
        bool b0 = x0 < a0.knots.front(), t0 = x0 > a0.knots.back();
        bool b1 = x1 < a1.knots.front(), t1 = x1 > a1.knots.back();
        if ((b0 && a0.lo == LINE) || (t0 && a0.hi == LINE)) {
            real g0 = ((1.0 - w) * (f10 - f00) + w * (f11 - f01)) / (a0.knots[i + 1] - a0.knots[i]);
            v = v + (x0 - xs0) * g0;
        }
        if ((b1 && a1.lo == LINE) || (t1 && a1.hi == LINE)) {
            real g1 = ((1.0 - u) * (f01 - f00) + u * (f11 - f10)) / (a1.knots[j + 1] - a1.knots[j]);
            v = v + (x1 - xs1) * g1;
        }
        return v;
    }
};

// three scaled-linear axes with Flat extrapolation: the propeller Lookup of a VariablePitch propeller
// (propellers.jl:253-273: extrapolate(scale(interpolate(coef, (BSpline(Linear()) x3)), J, Mt, dbeta), Flat()))
struct Interp3 {
    Axis a0, a1, a2;
    std::vector<double> y;  // y[(i0 * n1 + i1) * n2 + i2]
    double at(int i, int j, int k) const { return y[((size_t)i * a1.n() + j) * a2.n() + k]; }
    real operator()(real x0, real x1, real x2) const {
        real xs0 = rclamp(x0, a0.knots.front(), a0.knots.back());
        real xs1 = rclamp(x1, a1.knots.front(), a1.knots.back());
        real xs2 = rclamp(x2, a2.knots.front(), a2.knots.back());
        int i, j, k; real u, w, t;
        a0.locate(xs0, i, u);
        a1.locate(xs1, j, w);
        a2.locate(xs2, k, t);
        auto face = [&](int ii) {
            real f0 = (1.0 - t) * at(ii, j, k) + t * at(ii, j, k + 1);
            real f1 = (1.0 - t) * at(ii, j + 1, k) + t * at(ii, j + 1, k + 1);
            return (1.0 - w) * f0 + w * f1;
        };
        return (1.0 - u) * face(i) + u * face(i + 1);
    }
};
