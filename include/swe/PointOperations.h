// PointOperations.h — host math utilities of upstream include/PointOperations.h:8-51, src/PointOperations.cpp
// and include/CubicPolyMath.h, scalar (no Eigen). The same formulas, in the same operation order, run inside the
// device kernels (csrc/swe_device.cuh) and the CPU oracle; tests/cpp/host_api_test.cpp checks them against
// each other. A driver needs them for initial conditions (TriangAverage) and diagnostics.
#pragma once
#include <cmath>
#include <functional>

#include "Exceptions.h"
#include "TriangMesh.h"

using PointArray = Storage<3>;  // upstream include/PointOperations.h:6

inline double Len(const Point &a, const Point &b) noexcept {  // src/PointOperations.cpp:4-6
    return std::sqrt((a[0] - b[0]) * (a[0] - b[0]) + (a[1] - b[1]) * (a[1] - b[1]));
}
inline double Det(const Point &a, const Point &b) noexcept { return a[0] * b[1] - b[0] * a[1]; }  // :8-10
inline double TriangArea(const Point &a, const Point &b, const Point &c) noexcept {  // :12-14
    return 0.5 * std::fabs(Det(b - a, c - a));
}
inline Point Intersection(const Point &a1, const Point &a2, const Point &b1, const Point &b2) {  // :16-24
    const double den = Det(a2 - a1, b2 - b1);
    if (den < tol) throw SolverError("trying to find intersection of parallel vectors");
    return a1 + (Det(b1 - a1, b2 - b1) / den) * (a2 - a1);
}

// include/PointOperations.h:20-44: average of f over the triangle by n^2 congruent sub-triangles, f evaluated at
// every sub-centroid, in upstream's loop and accumulation order.
template <size_t k, size_t n>
Array<k> TriangAverage(const Point &p0, const Point &p1, const Point &p2, const std::function<Array<k>(const Point &)> &f) {
    constexpr double h = 1. / n;
    const Point di = h * (p1 - p0);
    const Point dj = h * (p2 - p0);
    const Point dt = 1. / 3. * (di + dj);
    Array<k> sum;
    Point pi = p0;
    for (unsigned i = 0; i < n; i++) {
        Point pt = pi + dt;
        for (unsigned j = 0; j < n - i - 1; j++) {
            sum += h * f(pt);
            sum += h * f(pt + dt);
            pt += dj;
        }
        sum += h * f(pt);
        pi += di;
    }
    return h * sum;
}

struct CubicPoly {  // include/CubicPolyMath.h:6-19: x^3 + b x^2 + c x + d, constructor order (d, c, b)
    CubicPoly(double d = 0, double c = 0, double b = 0) : m_b(b), m_c(c), m_d(d) {}
    double operator()(double x) const { return x * x * x + m_b * x * x + m_c * x + m_d; }

 private:
    double m_b, m_c, m_d;
};

// src/PointOperations.cpp:26-40: sign-bit bisection, accuracy + (int)log2(range) + 1 halvings
inline double Bisection(const std::function<double(double)> &f, double xmin = 0., double xmax = 1., const int accuracy = 50) {
    if (std::signbit(f(xmin)) == std::signbit(f(xmax))) return (std::fabs(f(xmin)) < std::fabs(f(xmax))) ? xmin : xmax;
    const int n = accuracy + static_cast<int>(std::log2(xmax - xmin));
    double x = xmin;
    for (int i = 0; i <= n; i++) {
        x = 0.5 * (xmin + xmax);
        ((std::signbit(f(xmin)) != std::signbit(f(x))) ? xmax : xmin) = x;
    }
    return x;
}

// src/PointOperations.cpp:42-48: slope (d/dx, d/dy) of the plane through three points (x, y, z); the 2x2 system is
// solved like Eigen's partialPivLu: pivot = row with the larger |a_i0| (first on ties), true divisions.
inline std::array<double, 2> Gradient(const Point &P0, const Point &P1, const Point &P2) {
    double a00 = P1[0] - P0[0], a01 = P1[1] - P0[1], r0 = P1[2] - P0[2];
    double a10 = P2[0] - P0[0], a11 = P2[1] - P0[1], r1 = P2[2] - P0[2];
    if (std::fabs(a10) > std::fabs(a00)) { std::swap(a00, a10); std::swap(a01, a11); std::swap(r0, r1); }
    const double l = a10 / a00;
    const double u11 = a11 - l * a01;
    const double c1 = r1 - l * r0;
    const double g1 = c1 / u11;
    return {(r0 - a01 * g1) / a00, g1};
}
