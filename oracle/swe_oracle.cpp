// swe_oracle.cpp — CPU ORACLE. TEST INFRASTRUCTURE ONLY.
//
// A scalar C++17 restatement of the reference's explicit finite-volume time step, used as the
// parity authority for the CUDA path and as the timed CPU baseline. Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it;
// the product (swe_fvm_b200/, include/) never does.
//
// PARITY STATUS: PINNED. upstream's own src/*.cpp are compiled (oracle/Makefile.ref, Eigen subset shim
// oracle/eigen_shim, named repairs in oracle/ref_patches/) into oracle/_ref/, and tests/test_ref_anchor.py
// demands BIT equality between that code and this file — function by function (Gradient, Bisection,
// ElemFlux, wavespeeds, assigners, Domain geometry, every reconstruction, fluxes, RHS, draining dt,
// TriangAverage) and over whole Solvers::Euler/SSPRK2/SSPRK3 steps — in the modes recon/pw2 = as written
// and repaired, libm=1, sequential=1. The DEFAULT mode differs from upstream only by the listed decisions
// S7/S8 (stage snapshot instead of order-dependent in-place loops; checked against upstream's own RHS
// applied out of place) and S9 (cbrt / (int)log2 restated with IEEE-only arithmetic, <= 1 ulp). The
// reference's committed outputs pin the mesh numbering and the initial condition exactly
// (notebooks/topology.dat, out0.dat) and the one-step result coarsely (out1.dat, print precision,
// written by an intermediate upstream revision). Citations below are paths relative to upstream; the
// semantic decisions S1-S11 of SURVEY.md App. A.10 are each selectable by an option.
//
// Arithmetic: build with `g++ -O2 -ffp-contract=off` on x86-64 (SSE2 doubles, no x87, no FMA),
// so every expression below is evaluated in the written order in IEEE binary64; the CUDA
// kernels are compiled with -fmad=false and written in the same order. cbrt and (int)log2 are
// restated with integer/IEEE-only algorithms (det_cbrt, ilog2_trunc) shared in spirit with the
// device code because libm and CUDA differ in the last ulp there (S9).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

using Idx = int64_t;
constexpr double tol = 1e-13;  // include/Includes.h:30
constexpr double kInf = std::numeric_limits<double>::infinity();

inline bool IsWet(double h) { return h > 1e-12; }  // include/Bathymetry.h:5-8

// --- S9 helpers -------------------------------------------------------------------------
// cbrt with IEEE + - * / only (<= 1 ulp from the exact value, like glibc's cbrt).
inline double det_cbrt(double x) {
    if (x == 0.0 || x != x || x == kInf || x == -kInf) return x;
    double a = std::fabs(x);
    double scale = 1.0;
    if (a < 1e-280) { a *= 0x1p+162; scale = 0x1p-54; }  // 2^162 = (2^54)^3, exact
    uint64_t bits;
    std::memcpy(&bits, &a, 8);
    bits = bits / 3 + 0x2A9F7893782DA1CEull;  // exponent/3 + bias: ~4% initial error
    double y;
    std::memcpy(&y, &bits, 8);
    for (int it = 0; it < 5; ++it) y = (2.0 * y + a / (y * y)) / 3.0;  // Newton
    y = y - (y * y * y - a) / (3.0 * (y * y));                         // final correction
    y *= scale;
    return x < 0 ? -y : y;
}
// (int)log2(x) for finite x > 0 (truncation toward zero), from the exponent bits.
inline int ilog2_trunc(double x) {
    uint64_t bits;
    std::memcpy(&bits, &x, 8);
    int e = int((bits >> 52) & 0x7ff);
    uint64_t frac = bits & 0xfffffffffffffull;
    if (e == 0) {  // subnormal: normalise
        if (frac == 0) return std::numeric_limits<int>::min();
        int sh = 0;
        while (!(frac & (1ull << 52))) { frac <<= 1; ++sh; }
        frac &= 0xfffffffffffffull;
        e = 1 - sh;
    }
    e -= 1023;
    if (e >= 0) return e;               // x >= 1: floor == trunc
    return frac == 0 ? e : e + 1;       // x < 1: trunc toward zero == ceil
}

// include/CubicPolyMath.h:6-19 — x^3 + b x^2 + c x + d, ctor order (d, c, b)
struct CubicPoly {
    double b, c, d;
    CubicPoly(double d_ = 0, double c_ = 0, double b_ = 0) : b(b_), c(c_), d(d_) {}
    double operator()(double x) const { return x * x * x + b * x * x + c * x + d; }
};

// src/PointOperations.cpp:26-40
template <class F>
double Bisection(const F &f, double xmin = 0., double xmax = 1., const int accuracy = 50, bool libm = false) {
    if (std::signbit(f(xmin)) == std::signbit(f(xmax)))
        return (std::fabs(f(xmin)) < std::fabs(f(xmax))) ? xmin : xmax;
    int n = accuracy + (libm ? static_cast<int>(std::log2(xmax - xmin)) : ilog2_trunc(xmax - xmin));
    double x = xmin;
    for (int i = 0; i <= n; i++) {
        x = 0.5 * (xmin + xmax);
        ((std::signbit(f(xmin)) != std::signbit(f(x))) ? xmax : xmin) = x;
    }
    return x;
}

// src/PointOperations.cpp:42-48 — slope of the plane through 3 points (columns of a 3x3),
// 2x2 system solved like Eigen's partialPivLu: pivot = row with the larger |a_i0| (first
// one on ties), true divisions.
inline void Gradient(const double P0[3], const double P1[3], const double P2[3], double g[2]) {
    double a00 = P1[0] - P0[0], a01 = P1[1] - P0[1], r0 = P1[2] - P0[2];
    double a10 = P2[0] - P0[0], a11 = P2[1] - P0[1], r1 = P2[2] - P0[2];
    if (std::fabs(a10) > std::fabs(a00)) {
        std::swap(a00, a10); std::swap(a01, a11); std::swap(r0, r1);
    }
    double l = a10 / a00;
    double u11 = a11 - l * a01;
    double c1 = r1 - l * r0;
    g[1] = c1 / u11;
    g[0] = (r0 - a01 * g[1]) / a00;
}

inline double Det(const double a[2], const double b[2]) { return a[0] * b[1] - b[0] * a[1]; }  // :8-10

// which branch the last reconstruction call of this thread took (debug tap for the tests that
// assert every branch of PartWet1 / PartWet2 / FullWet is exercised)
struct Trace { int pw1 = -1, pw2 = -1, fw = 0; };
thread_local Trace g_trace;
enum { FW_DRY_NB = 1, FW_PW_NB = 2, FW_VERTEX_ZERO = 4, FW_TVD_OFF = 8 };

struct MUSCL {  // include/MUSCLObject.h:15-55
    double o[3];
    double G[3][2];
    Idx i;
};

struct Oracle {
    Idx nn = 0, ne = 0, nt = 0;
    std::vector<double> geom;
    std::vector<Idx> ep, et, tp, te, tt;
    double cor = 0, tau = 0;
    // options
    int sequential = 0;  // S7/S8: 0 = Jacobi (snapshot), 1 = reference loop order, in place
    int recon = 0;       // S2: 0 = repaired (w,u,v plane gradients), 1 = as written, 2 = first order
    int roe_fix = 0;     // S5: 0 = cl*ur as written, 1 = cr*ur
    int cfl_abs = 0;     // S6: 0 = signed max as written, 1 = magnitudes
    int pw2 = 0;         // S3: 0 = PartWet2 points(r,c) read as (point, coordinate), 1 = as written
    int libm = 0;        // S9: 0 = det_cbrt / ilog2_trunc (shared with the device), 1 = libm cbrt, (int)log2
    int threads = 1;
    // geometry, computed once with the reference's formulas (the reference recomputes them on
    // every access: src/Bathymetry.cpp:20-31,69-90)
    std::vector<double> Tc, Ec, Len_, Area_, slope, norm0, norm1;
    std::vector<uint8_t> bnd_tri;
    // fields (include/SpaceDisc.h:39-44, include/MUSCLObject.h:67-68)
    std::vector<double> vol, edg, src, f, maxwp, volref;
    std::vector<uint8_t> cfl_mask;
    double min_len = 1.;
    // branch-hit counters of the last ComputeInterfaceValues (see Trace): [0..2] pass-1 PartWet1 of a
    // part-wet cell (submerged / cbrt / bisection), [3..6] full-wet cells (dry neighbour / part-wet
    // neighbour / vertex check zeroed the gradient / TVD switched a component off), [7..11] pass-2
    // PartWet2 (early PartWet1 / 1 wet / 3 wet / 2 wet / 2-wet fallback)
    int64_t branch[12] = {0};
    void count_pass1(int cls) {
        int64_t add[12] = {0};
        if (cls == 1) add[g_trace.pw1] = 1;
        else if (cls == 2) for (int b = 0; b < 4; ++b) add[3 + b] = (g_trace.fw >> b) & 1;
        for (int b = 0; b < 12; ++b) if (add[b]) {
#pragma omp atomic
            branch[b] += 1;
        }
    }
    void count_pass2() {
#pragma omp atomic
        branch[7 + g_trace.pw2] += 1;
    }

    // ---- Domain (src/Bathymetry.cpp) ----
    const double *P(Idx n) const { return &geom[3 * n]; }
    const double *T(Idx t) const { return &Tc[3 * t]; }
    const double *E(Idx e) const { return &Ec[3 * e]; }  // S1: edge midpoint incl. bed
    double L(Idx e) const { return Len_[e]; }
    double Area(Idx t) const { return Area_[t]; }
    const double *TriangSlope(Idx t) const { return &slope[2 * t]; }
    const double *Norm(Idx e, Idx t) const { return (et[2 * e] == t) ? &norm0[2 * e] : &norm1[2 * e]; }

    void tang(Idx e, Idx t, double tan[2]) const {  // src/Bathymetry.cpp:69-76
        const double *p0 = P(ep[2 * e]), *p1 = P(ep[2 * e + 1]);
        double len = L(e);
        tan[0] = (p1[0] - p0[0]) / len;
        tan[1] = (p1[1] - p0[1]) / len;
        double d[2] = {T(t)[0] - p0[0], T(t)[1] - p0[1]};
        if (Det(d, tan) > 0.) { tan[0] = -tan[0]; tan[1] = -tan[1]; }
    }

    void precompute() {
        Tc.resize(3 * nt); Ec.resize(3 * ne); Len_.resize(ne); Area_.resize(nt); slope.resize(2 * nt);
        norm0.assign(2 * ne, 0.); norm1.assign(2 * ne, 0.); bnd_tri.assign(nt, 0);
        const double third = 1. / 3.;
        for (Idx t = 0; t < nt; ++t) {
            const double *p0 = P(tp[3 * t]), *p1 = P(tp[3 * t + 1]), *p2 = P(tp[3 * t + 2]);
            for (int c = 0; c < 3; ++c) Tc[3 * t + c] = p0[c] * third + p1[c] * third + p2[c] * third;  // :24-27
            double a[2] = {p1[0] - p0[0], p1[1] - p0[1]}, b[2] = {p2[0] - p0[0], p2[1] - p0[1]};
            Area_[t] = 0.5 * std::fabs(Det(a, b));  // src/PointOperations.cpp:12-14
            Gradient(p0, p1, p2, &slope[2 * t]);    // src/Bathymetry.cpp:10-13
        }
        for (Idx e = 0; e < ne; ++e) {
            const double *p0 = P(ep[2 * e]), *p1 = P(ep[2 * e + 1]);
            for (int c = 0; c < 3; ++c) Ec[3 * e + c] = 0.5 * (p0[c] + p1[c]);
            Len_[e] = std::sqrt((p0[0] - p1[0]) * (p0[0] - p1[0]) + (p0[1] - p1[1]) * (p0[1] - p1[1]));  // :4-6
        }
        for (Idx e = 0; e < ne; ++e) {
            double tan[2];
            tang(e, et[2 * e], tan);
            // src/Bathymetry.cpp:78-80: [[0, 1], [-1, 0]] * Tang, multiplied out literally so that
            // even the signs of zero components are upstream's
            norm0[2 * e] = 0. * tan[0] + 1. * tan[1]; norm0[2 * e + 1] = -1. * tan[0] + 0. * tan[1];
            if (et[2 * e + 1] >= 0) {
                tang(e, et[2 * e + 1], tan);
                norm1[2 * e] = 0. * tan[0] + 1. * tan[1]; norm1[2 * e + 1] = -1. * tan[0] + 0. * tan[1];
            }
        }
        for (Idx t = 0; t < nt; ++t)  // src/TriangMesh.cpp:18-23
            for (int k = 0; k < 3; ++k)
                if (et[2 * te[3 * t + k] + 1] < 0) bnd_tri[t] = 1;
        vol.assign(3 * nt, 0.); volref.assign(3 * nt, 0.);
        edg.assign(6 * ne, 0.); src.assign(6 * ne, 0.); f.assign(3 * ne, 0.);
        maxwp.assign(nn, 0.);
    }

    // ---- VolumeField / EdgeField (include/ValueField.h:31-48,70-75; src/ValueField.cpp) ----
    double vb(Idx i) const { return Tc[3 * i + 2]; }
    static double vh(const std::vector<double> &v, const Oracle &o, Idx i) { return v[3 * i] - o.vb(i); }
    static Idx EId(Idx e, Idx from, Idx to) { return 2 * e + (Idx)(from < to); }
    double eb(Idx e) const { return Ec[3 * e + 2]; }

    // ---- classification (src/MUSCLObject.cpp:13-29) on a given state array ----
    bool IsDryCell(const std::vector<double> &v, Idx i) const { return !IsWet(v[3 * i] - vb(i)); }
    bool IsFullWetCell(Idx i) const {
        if (bnd_tri[i]) return false;
        double max_bp = std::max(std::max(P(tp[3 * i])[2], P(tp[3 * i + 1])[2]), P(tp[3 * i + 2])[2]);
        return max_bp < vol[3 * i];
    }
    bool IsPartWetCell(Idx i) const { return !IsDryCell(vol, i) && !IsFullWetCell(i); }

    // ---- MUSCL (include/MUSCLObject.h:26-48) ----
    void AtPoint(const MUSCL &m, const double p[3], double out[3]) const {
        const double dx = p[0] - T(m.i)[0], dy = p[1] - T(m.i)[1];
        for (int c = 0; c < 3; ++c) out[c] = m.o[c] + (m.G[c][0] * dx + m.G[c][1] * dy);
        if (!((out[0] - p[2]) >= 0)) { out[0] = p[2]; out[1] = 0.; out[2] = 0.; }
    }
    void GradientRow0(const MUSCL &m, const double p[3], double g[2]) const {
        double a[3];
        AtPoint(m, p, a);
        double h = a[0] - p[2];
        if (h >= 0) { g[0] = m.G[0][0]; g[1] = m.G[0][1]; }
        else { g[0] = TriangSlope(m.i)[0]; g[1] = TriangSlope(m.i)[1]; }  // dryGradient
    }

    MUSCL ReconstructDryCell(Idx i) const {  // src/MUSCLObject.cpp:31-36
        MUSCL m{};
        m.i = i;
        m.o[0] = vb(i); m.o[1] = 0.; m.o[2] = 0.;
        m.G[0][0] = TriangSlope(i)[0]; m.G[0][1] = TriangSlope(i)[1];
        return m;
    }

    MUSCL ReconstructPartWetCell1(Idx i) const {  // src/MUSCLObject.cpp:86-112
        const double z0 = P(tp[3 * i])[2], z1 = P(tp[3 * i + 1])[2], z2 = P(tp[3 * i + 2])[2];
        double b13 = std::max(std::max(z0, z1), z2);
        double b23 = std::min(std::min(z0, z1), z2);
        double b12 = 3. * vb(i) - b23 - b13;
        double b_delimiter = b12 + (1. / 3.) * (b13 - b12) * (b13 - b12) / (b13 - b23);
        const double wi = vol[3 * i], hi = vol[3 * i] - vb(i);
        double w_rec;
        if (wi >= b13) {
            w_rec = wi; g_trace.pw1 = 0;
        } else if (wi <= b_delimiter) {
            g_trace.pw1 = 1;
            w_rec = b23 + (libm ? std::cbrt(3. * hi * (b13 - b23) * (b12 - b23)) : det_cbrt(3. * hi * (b13 - b23) * (b12 - b23)));
        } else {
            double a = -3. * b13;
            double b = 3. * (b12 * b13 + b13 * b23 - b12 * b23);
            double c = (b13 - b23) * (3. * hi * (b13 - b12) - b12 * (b12 + b23)) - b23 * b23 * b13;
            g_trace.pw1 = 2;
            w_rec = Bisection(CubicPoly(c, b, a), b12, b13, 50, libm != 0);
        }
        MUSCL m{};
        m.i = i;
        m.o[0] = w_rec; m.o[1] = vol[3 * i + 1]; m.o[2] = vol[3 * i + 2];
        return m;
    }

    MUSCL ReconstructFullWetCell(Idx i) const {  // src/MUSCLObject.cpp:38-84 (S2)
        const Idx *ip = &tp[3 * i], *ie = &te[3 * i], *it = &tt[3 * i];
        const double *pi = &vol[3 * i];
        MUSCL m{};
        m.i = i;
        for (int c = 0; c < 3; ++c) m.o[c] = pi[c];
        g_trace.fw = 0;
        double X[3][3];  // grad_points: (x, y, bed) per support point
        double V[3][3];  // grad_values per support point
        for (int k = 0; k < 3; ++k) {
            if (IsFullWetCell(it[k])) {
                for (int c = 0; c < 3; ++c) { X[k][c] = T(it[k])[c]; V[k][c] = vol[3 * it[k] + c]; }
            } else if (IsDryCell(vol, it[k])) {
                g_trace.fw |= FW_DRY_NB;
                return m;  // zero gradient
            } else {
                g_trace.fw |= FW_PW_NB;
                MUSCL nb = ReconstructPartWetCell1(it[k]);
                const double *pt = E(ie[k]);
                double a[3];
                AtPoint(nb, pt, a);
                for (int c = 0; c < 3; ++c) { X[k][c] = pt[c]; V[k][c] = 0.5 * (pi[c] + a[c]); }
            }
        }
        if (recon == 2) return m;  // first-order option
        double df[3][2] = {{0, 0}, {0, 0}, {0, 0}};
        if (recon == 1) {
            Gradient(X[0], X[1], X[2], df[0]);  // as written: z-row = bed of the support points
        } else {
            for (int c = 0; c < 3; ++c) {
                double q0[3] = {X[0][0], X[0][1], V[0][c]}, q1[3] = {X[1][0], X[1][1], V[1][c]},
                       q2[3] = {X[2][0], X[2][1], V[2][c]};
                Gradient(q0, q1, q2, df[c]);
            }
        }
        // vertex positivity (:66-72): dx = P(ip) * (I - 1/3), literally
        const double md = 1. - 1. / 3., mo = 0. - 1. / 3.;
        bool all_wet = true;
        for (int k = 0; k < 3; ++k) {
            double dx = 0., dy = 0.;
            for (int j = 0; j < 3; ++j) {
                double mjk = (j == k) ? md : mo;
                if (j == 0) { dx = P(ip[j])[0] * mjk; dy = P(ip[j])[1] * mjk; }
                else { dx += P(ip[j])[0] * mjk; dy += P(ip[j])[1] * mjk; }
            }
            double wp = (df[0][0] * dx + df[0][1] * dy) + pi[0];
            double hp = wp - P(ip[k])[2];
            if (!IsWet(hp)) all_wet = false;
        }
        if (!all_wet) { g_trace.fw |= FW_VERTEX_ZERO; for (int c = 0; c < 3; ++c) df[c][0] = df[c][1] = 0.; }
        // on/off TVD limiter (:74-81)
        double TVD[3] = {1., 1., 1.};
        for (int k = 0; k < 3; ++k) {
            const double *pn = &vol[3 * it[k]];
            const double dx = E(ie[k])[0] - T(i)[0], dy = E(ie[k])[1] - T(i)[1];
            for (int c = 0; c < 3; ++c) {
                double vtmin = std::min(pi[c], pn[c]);
                double vtmax = std::max(pi[c], pn[c]);
                double vek = pi[c] + (df[c][0] * dx + df[c][1] * dy);
                if (!((vtmin <= vek) && (vek <= vtmax))) { TVD[c] = 0.; g_trace.fw |= FW_TVD_OFF; }
            }
        }
        for (int c = 0; c < 3; ++c) { m.G[c][0] = TVD[c] * df[c][0]; m.G[c][1] = TVD[c] * df[c][1]; }
        return m;
    }

    MUSCL ReconstructPartWetCell2(Idx i) const {  // src/MUSCLObject.cpp:114-191 (S3)
        Idx ip[3] = {tp[3 * i], tp[3 * i + 1], tp[3 * i + 2]};
        if (P(ip[0])[2] > P(ip[1])[2]) std::swap(ip[0], ip[1]);
        if (P(ip[1])[2] > P(ip[2])[2]) std::swap(ip[1], ip[2]);
        if (P(ip[0])[2] > P(ip[1])[2]) std::swap(ip[0], ip[1]);
        const double *Q0 = P(ip[0]), *Q1 = P(ip[1]), *Q2 = P(ip[2]);
        double b23 = Q0[2], b12 = Q1[2], b13 = Q2[2];
        if ((vol[3 * i] > b13) || (b13 - b23 < tol)) { g_trace.pw2 = 0; return ReconstructPartWetCell1(i); }

        double w23 = maxwp[ip[0]];
        double h23 = w23 - b23;
        double ratio_b = (b12 - b23) / (b13 - b23);
        double h_delimiter1 = 1. / 3. * h23 * ratio_b;
        double h_delimiter2 = 1. / 3. * h23 * (2. * b13 - b12 - b23) / (b13 - b23);
        double hi = vol[3 * i] - vb(i);
        double ratio_h = hi / h23;

        // as written (pw2 == 1) the scalar accesses points(0,2), points(1,2) hit x/y entries of the
        // third point that the following points.col(2) = ... overwrites: S0.z stays b23, S1.z stays b12
        double S0[3] = {Q0[0], Q0[1], pw2 ? Q0[2] : w23}, S1[3], S2[3];
        if (hi <= h_delimiter1) {  // 1 point wet, 2 dry
            g_trace.pw2 = 1;
            double k2 = std::sqrt(3. * ratio_h / ratio_b);
            for (int c = 0; c < 3; ++c) S1[c] = k2 * Q1[c] + (1. - k2) * Q0[c];
            double k3 = std::sqrt(3. * ratio_h * ratio_b);
            for (int c = 0; c < 3; ++c) S2[c] = k3 * Q2[c] + (1. - k3) * Q0[c];
        } else if (hi >= h_delimiter2) {  // 3 points wet
            g_trace.pw2 = 2;
            double delta_w = 1.5 * (hi - h_delimiter2);
            S1[0] = Q1[0]; S1[1] = Q1[1]; S1[2] = Q1[2];
            if (!pw2) {
                S1[2] += delta_w;
                S1[2] += (1. - ratio_b) * h23;
            }
            S2[0] = Q2[0]; S2[1] = Q2[1]; S2[2] = Q2[2];
            S2[2] += delta_w;
        } else {  // 2 points wet, 1 dry
            g_trace.pw2 = 3;
            double alpha = 3. * ratio_h;
            double beta = (b13 - b12) / (b13 - b23);
            double k1 = 1. - Bisection(CubicPoly{(1. + beta - alpha) / (beta * beta), (alpha - 3.) / beta}, 0., 1., 50, libm != 0);
            if (k1 < tol) { g_trace.pw2 = 4; return ReconstructPartWetCell1(i); }
            double k3 = 1. - beta * (1. - k1);
            double bp1 = k1 * b13 + (1. - k1) * b12;
            S1[0] = Q1[0]; S1[1] = Q1[1];
            S1[2] = pw2 ? Q1[2] : b12 + (k1 / k3) * beta * h23;
            S2[0] = k1 * Q2[0] + (1. - k1) * Q1[0];
            S2[1] = k1 * Q2[1] + (1. - k1) * Q1[1];
            S2[2] = bp1;
        }
        MUSCL m{};
        m.i = i;
        Gradient(S0, S1, S2, m.G[0]);
        double wt = w23 + (m.G[0][0] * (T(i)[0] - Q0[0]) + m.G[0][1] * (T(i)[1] - Q0[1]));
        m.o[0] = wt; m.o[1] = vol[3 * i + 1]; m.o[2] = vol[3 * i + 2];
        return m;
    }

    // ---- src/SpaceDisc.cpp:15-31 + PrimAssigner (src/Assigners.cpp:8-20) ----
    void UpdateInterfaceValues(const MUSCL &m, bool write_max) {
        const Idx i = m.i;
        const Idx *ip = &tp[3 * i], *ie = &te[3 * i], *it = &tt[3 * i];
        for (int k = 0; k < 3; ++k) {
            double a[3];
            if (write_max) {
                AtPoint(m, P(ip[k]), a);
                double *slot = &maxwp[ip[k]];
#ifdef _OPENMP
                if (threads > 1) {  // order-independent max (atomic CAS loop)
                    double cur = *slot;
                    while (cur < a[0] &&
                           !__atomic_compare_exchange(slot, &cur, &a[0], true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
                } else
#endif
                    *slot = std::max(*slot, a[0]);
            }
            const double *c = E(ie[k]);
            AtPoint(m, c, a);
            const Idx id = EId(ie[k], i, it[k]);
            double *e = &edg[3 * id];
            const double be = eb(ie[k]);
            double h = a[0] - be;
            if (!IsWet(h)) {
                e[0] = be; e[1] = 0.; e[2] = 0.;
            } else {
                e[0] = a[0]; e[1] = a[1]; e[2] = a[2];
                if (h < 1e-3) {
                    double fac = std::sqrt(2) * h / std::sqrt(h * h + 1e-6);
                    e[1] *= fac; e[2] *= fac;
                }
            }
            double g[2];
            GradientRow0(m, c, g);
            double u = e[1], v = e[2];
            src[3 * id + 1] = g[0] + cor * (-v);
            src[3 * id + 2] = g[1] + cor * u;
        }
    }

    void ComputeInterfaceValues() {  // src/SpaceDisc.cpp:33-52
        for (Idx n = 0; n < nn; ++n) maxwp[n] = geom[3 * n + 2];
        for (int b = 0; b < 12; ++b) branch[b] = 0;
        auto pass1 = [this](Idx i) {
            if (IsDryCell(vol, i)) { UpdateInterfaceValues(ReconstructDryCell(i), true); }
            else if (!IsFullWetCell(i)) { MUSCL m = ReconstructPartWetCell1(i); count_pass1(1); UpdateInterfaceValues(m, true); }
            else { MUSCL m = ReconstructFullWetCell(i); count_pass1(2); UpdateInterfaceValues(m, true); }
        };
        if (sequential) {
            for (Idx i = 0; i < nt; ++i) pass1(i);
            for (Idx i = 0; i < nt; ++i)
                if (IsPartWetCell(i)) { MUSCL m = ReconstructPartWetCell2(i); count_pass2(); UpdateInterfaceValues(m, true); }
            return;
        }
        // S8 (Jacobi): pass 2 reads the node maxima of pass 1 only and does not update them.
#pragma omp parallel for schedule(static) num_threads(threads) if (threads > 1)
        for (Idx i = 0; i < nt; ++i) pass1(i);
#pragma omp parallel for schedule(static) num_threads(threads) if (threads > 1)
        for (Idx i = 0; i < nt; ++i)
            if (IsPartWetCell(i)) { MUSCL m = ReconstructPartWetCell2(i); count_pass2(); UpdateInterfaceValues(m, false); }
    }

    // ---- include/SpaceDisc.h:4-18 ----
    static void ElemFlux(const double n[2], const double U[3], double res[3]) {
        res[0] = res[1] = res[2] = 0.;
        double h = U[0];
        if (IsWet(h)) {
            double hveln = U[1] * n[0] + U[2] * n[1];
            res[0] = hveln;
            res[1] = (hveln / h) * U[1] + (0.5 * h * h) * n[0];
            res[2] = (hveln / h) * U[2] + (0.5 * h * h) * n[1];
        }
    }

    // ---- src/Fluxes.cpp:5-26 ----
    void Wavespeeds(int ws, double ul, double hl, double ur, double hr, double a[2]) const {
        double cl = std::sqrt(hl), cr = std::sqrt(hr);
        if (ws == 0) {  // Rusanov
            double aplus = std::max(std::fabs(ul) + cl, std::fabs(ur) + cr);
            a[0] = -aplus; a[1] = aplus;
        } else if (ws == 1) {  // Davis
            a[0] = std::min(ul - cl, ur - cr); a[1] = std::max(ul + cl, ur + cr);
        } else {  // Einfeldt; S5: `cl * ur` as written (:22)
            double uRoe = (cl * ul + (roe_fix ? cr : cl) * ur) / (cl + cr);
            double cRoe = std::sqrt(0.5 * (hl + hr));
            a[0] = std::min(ul - cl, uRoe - cRoe); a[1] = std::max(ur + cr, uRoe + cRoe);
        }
    }

    void econs(Idx id, Idx e, double U[3]) const {  // ValueField::cons
        double h = edg[3 * id] - eb(e);
        U[0] = h; U[1] = h * edg[3 * id + 1]; U[2] = h * edg[3 * id + 2];
    }

    // include/Fluxes.h:14-54 (HLL) and :56-111 (HLLC)
    void Flux(int kind, int ws, Idx e, Idx from, Idx to, double *r, double F[3]) const {
        // n = Norm(e, from) = (t_y, -t_x)  (src/Bathymetry.cpp:78-80)  =>  t = (-n_y, n_x) exactly
        const double *n = Norm(e, from);
        const double t[2] = {-n[1], n[0]};
        const Idx il = EId(e, from, to), ir = EId(e, to, from);
        const double *el = &edg[3 * il], *er = &edg[3 * ir];
        double ul = el[1] * n[0] + el[2] * n[1];
        double vl = el[1] * t[0] + el[2] * t[1];
        double hl = el[0] - eb(e);
        double ur = er[1] * n[0] + er[2] * n[1];
        double vr = er[1] * t[0] + er[2] * t[1];
        double hr = er[0] - eb(e);
        F[0] = F[1] = F[2] = 0.;
        if (hl + hr <= 1e-10) return;
        double a[2];
        Wavespeeds(ws, ul, hl, ur, hr, a);
        double Ul[3], Ur[3];
        econs(il, e, Ul);
        econs(ir, e, Ur);
        if (kind == 0) {  // HLL
            double al = std::min(0., a[0]);
            double ar = std::max(0., a[1]);
            if (ar - al <= 1e-10) return;
            if (r) {
                double dl = 2. * Area(from) / L(e);
                double dr = 2. * Area(to) / L(e);
                double c = std::fabs(cor);
                double length_to_wavespeed = std::min(dl, dr) / (c + std::max(-al, ar));
                *r = std::min(*r, length_to_wavespeed);
            }
            double Fl[3], Fr[3];
            ElemFlux(n, Ul, Fl);
            ElemFlux(n, Ur, Fr);
            for (int c = 0; c < 3; ++c) F[c] = (ar * Fl[c] - al * Fr[c] + (al * ar) * (Ur[c] - Ul[c])) / (ar - al);
            return;
        }
        double al = a[0], ar = a[1];
        double ustar = (ar - ur) * hr * ur - (al - ul) * hl * ul + 0.5 * (hl * hl - hr * hr);
        ustar /= (hr * (ar - ur) - hl * (al - ul));
        if (r) {
            double dl = 2. * Area(from) / L(e);
            double dr = 2. * Area(to) / L(e);
            double c = std::fabs(cor);
            double amax = cfl_abs ? std::max(std::fabs(al), std::fabs(ar)) : std::max(al, ar);  // S6
            double length_to_wavespeed = std::min(dl, dr) / (c + std::max(tol, amax));
            *r = std::min(*r, length_to_wavespeed);
        }
        if (ustar <= 0) {
            double urstar = vr * t[0] + ustar * t[1];
            double vrstar = vr * n[0] + ustar * n[1];
            double hrstar = hr * (ar - ur) / (ar - ustar);
            double Us[3] = {hrstar, hrstar * urstar, hrstar * vrstar};
            double Fe[3];
            ElemFlux(n, Ur, Fe);
            double s = std::max(0., ar);
            for (int c = 0; c < 3; ++c) F[c] = Fe[c] + s * (Us[c] - Ur[c]);
        } else {
            double ulstar = vl * t[0] + ustar * t[1];
            double vlstar = vl * n[0] + ustar * n[1];
            double hlstar = hl * (al - ul) / (al - ustar);
            double Us[3] = {hlstar, hlstar * ulstar, hlstar * vlstar};
            double Fe[3];
            ElemFlux(n, Ul, Fe);
            double s = std::min(0., al);
            for (int c = 0; c < 3; ++c) F[c] = Fe[c] + s * (Us[c] - Ul[c]);
        }
    }

    void ComputeFluxes(int kind, int ws) {  // src/SpaceDisc.cpp:54-74
        double mn = 1.;
#pragma omp parallel for schedule(static) num_threads(threads) reduction(min : mn) if (threads > 1 && !sequential)
        for (Idx e = 0; e < ne; ++e) {
            Idx lf = et[2 * e], lt = et[2 * e + 1];
            if (lt == -1) {  // SOLID_WALL: cell-mean depth (S10c)
                double U[3] = {vol[3 * lf] - vb(lf), 0., 0.};
                ElemFlux(Norm(e, lf), U, &f[3 * e]);
            } else {
                double r = 1.;
                bool counted = cfl_mask.empty() || cfl_mask[e];
                Flux(kind, ws, e, lf, lt, counted ? &r : nullptr, &f[3 * e]);
                mn = std::min(mn, r);
            }
        }
        min_len = mn;
    }

    // ---- src/TimeDisc.cpp:43-66 on the state array `v` (live or stage snapshot, S7) ----
    double ComputeDrainingDt(const std::vector<double> &v, Idx i) const {
        if (i < 0) return kInf;
        if (IsDryCell(v, i)) return 0.;
        const Idx *ie = &te[3 * i];
        double sum = 0.;
        for (int k = 0; k < 3; k++) {
            Idx itk = et[2 * ie[k]];
            double f_ek = f[3 * ie[k]];
            sum += std::max(0., (i == itk ? f_ek : -f_ek));
        }
        return sum > tol ? Area(i) * (v[3 * i] - vb(i)) / sum : kInf;
    }

    void RHS(const std::vector<double> &v, Idx i, double dt, double res[3]) const {  // src/TimeDisc.cpp:3-41
        res[0] = res[1] = res[2] = 0.;
        const Idx *ie = &te[3 * i], *it = &tt[3 * i];
        double i_area = 1. / Area(i);
        double dti = ComputeDrainingDt(v, i);
        for (int k = 0; k < 3; k++) {
            int sgn = et[2 * ie[k]] == i ? 1 : -1;
            double dtik = ComputeDrainingDt(v, it[k]);
            double dtk = (sgn * f[3 * ie[k]]) > 0. ? std::min(dt, dti) : std::min(dt, dtik);
            double c_ek = i_area * L(ie[k]);
            const Idx id = EId(ie[k], i, it[k]);
            double h_ek = edg[3 * id] - eb(ie[k]);
            double s = dtk * sgn * c_ek;
            for (int c = 0; c < 3; ++c) res[c] -= s * f[3 * ie[k] + c];
            res[1] -= dt * (1. / 3.) * src[3 * id + 1] * h_ek;
            res[2] -= dt * (1. / 3.) * src[3 * id + 2] * h_ek;
            const double *n = Norm(ie[k], i);
            res[1] += dtk * (n[0] * c_ek * (0.5 * h_ek * h_ek));
            res[2] += dtk * (n[1] * c_ek * (0.5 * h_ek * h_ek));
        }
    }

    // ConsAssigner (src/Assigners.cpp:22-44)
    void cons_get(const std::vector<double> &v, Idx i, double U[3]) const {
        double h = v[3 * i] - vb(i);
        U[0] = h; U[1] = v[3 * i + 1] * h; U[2] = v[3 * i + 2] * h;
    }
    void cons_set(Idx i, const double U[3]) {
        double h = U[0];
        double *o = &vol[3 * i];
        if (!IsWet(h)) { o[0] = vb(i); o[1] = 0.; o[2] = 0.; return; }
        double ih;
        if (h < 1e-3) ih = std::sqrt(2) * h / std::sqrt(h * h * h * h + 1e-12);
        else ih = 1. / h;
        o[0] = h + vb(i); o[1] = U[1] * ih; o[2] = U[2] * ih;
    }

    // one stage: cons(i) = a0*U0.cons(i) + a1*cons(i) + RHS(i, dts); a0 == 0 means no U0 term.
    void StageUpdate(const std::vector<double> *U0, double a0, double a1, double dts, bool plain_sum) {
        if (!sequential) volref = vol;  // S7: drain()/dry() read the stage snapshot
        const std::vector<double> &ref = sequential ? vol : volref;
#pragma omp parallel for schedule(static) num_threads(threads) if (threads > 1 && !sequential)
        for (Idx i = 0; i < nt; ++i) {
            double r[3], U[3], Uc[3];
            RHS(ref, i, dts, r);
            cons_get(vol, i, Uc);
            if (plain_sum) {  // Euler `+=` and the first RK stage: U + RHS
                if (U0) cons_get(*U0, i, Uc);
                for (int c = 0; c < 3; ++c) U[c] = Uc[c] + r[c];
            } else {
                double Ua[3];
                cons_get(*U0, i, Ua);
                for (int c = 0; c < 3; ++c) U[c] = a0 * Ua[c] + a1 * Uc[c] + r[c];
            }
            cons_set(i, U);
        }
    }

    void Step(int scheme, int kind, int ws, double dt) {  // src/Solvers.cpp
        ComputeInterfaceValues();
        ComputeFluxes(kind, ws);
        if (scheme == 0) {  // Euler :5-14
            StageUpdate(nullptr, 0., 1., dt, true);
            return;
        }
        const std::vector<double> U0 = vol;  // deep copy, :22 / :41
        StageUpdate(&U0, 1., 0., dt, true);
        ComputeInterfaceValues();
        ComputeFluxes(kind, ws);
        if (scheme == 1) {  // SSPRK2 :16-33
            StageUpdate(&U0, 0.5, 0.5, 0.5 * dt, false);
            return;
        }
        StageUpdate(&U0, 0.75, 0.25, 0.25 * dt, false);  // SSPRK3 :35-59
        ComputeInterfaceValues();
        ComputeFluxes(kind, ws);
        StageUpdate(&U0, (1. / 3.), (2. / 3.), (2. / 3.) * dt, false);
    }
};

}  // namespace

// -------------------------------------------------------------------------------------------
// C interface for ctypes (tests/, bench.py cpu_baseline)
// -------------------------------------------------------------------------------------------
extern "C" {

void *oracle_create(int64_t nn, int64_t ne, int64_t nt, const double *geom, const int64_t *ep, const int64_t *et,
                    const int64_t *tp, const int64_t *te, const int64_t *tt, double cor, double tau) {
    Oracle *o = new Oracle();
    o->nn = nn; o->ne = ne; o->nt = nt; o->cor = cor; o->tau = tau;
    o->geom.assign(geom, geom + 3 * nn);
    o->ep.assign(ep, ep + 2 * ne); o->et.assign(et, et + 2 * ne);
    o->tp.assign(tp, tp + 3 * nt); o->te.assign(te, te + 3 * nt); o->tt.assign(tt, tt + 3 * nt);
    o->precompute();
    return o;
}
void oracle_destroy(void *p) { delete static_cast<Oracle *>(p); }

int oracle_set_option(void *p, const char *key, int value) {
    Oracle *o = static_cast<Oracle *>(p);
    if (!std::strcmp(key, "sequential")) o->sequential = value;
    else if (!std::strcmp(key, "recon")) o->recon = value;
    else if (!std::strcmp(key, "roe_fix")) o->roe_fix = value;
    else if (!std::strcmp(key, "cfl_abs")) o->cfl_abs = value;
    else if (!std::strcmp(key, "pw2")) o->pw2 = value;
    else if (!std::strcmp(key, "libm")) o->libm = value;
    else if (!std::strcmp(key, "threads")) {
#ifdef _OPENMP
        o->threads = value > 0 ? value : omp_get_max_threads();
#else
        o->threads = 1;
#endif
    } else return -1;
    return 0;
}
int oracle_get_threads(void *p) { return static_cast<Oracle *>(p)->threads; }

void oracle_set_cfl_edge_mask(void *p, const uint8_t *mask) {
    Oracle *o = static_cast<Oracle *>(p);
    if (mask) o->cfl_mask.assign(mask, mask + o->ne); else o->cfl_mask.clear();
}
void oracle_set_state(void *p, const double *prim) {
    Oracle *o = static_cast<Oracle *>(p);
    std::copy(prim, prim + 3 * o->nt, o->vol.begin());
}
void oracle_get_state(void *p, double *prim) {
    Oracle *o = static_cast<Oracle *>(p);
    std::copy(o->vol.begin(), o->vol.end(), prim);
}
void oracle_step(void *p, int scheme, int flux, int ws, double dt) { static_cast<Oracle *>(p)->Step(scheme, flux, ws, dt); }
void oracle_run(void *p, int scheme, int flux, int ws, int64_t nsteps, double dt, double dt0) {
    Oracle *o = static_cast<Oracle *>(p);
    double d = dt > 0 ? dt : dt0;
    for (int64_t s = 0; s < nsteps; ++s) {
        o->Step(scheme, flux, ws, d);
        if (!(dt > 0)) d = 0.15 * o->min_len;
    }
}
void oracle_compute_interface_values(void *p) { static_cast<Oracle *>(p)->ComputeInterfaceValues(); }
void oracle_compute_fluxes(void *p, int flux, int ws) { static_cast<Oracle *>(p)->ComputeFluxes(flux, ws); }
void oracle_stage_update(void *p, const double *U0prim, double a0, double a1, double dts, int plain_sum) {
    Oracle *o = static_cast<Oracle *>(p);
    if (U0prim) {
        std::vector<double> U0(U0prim, U0prim + 3 * o->nt);
        o->StageUpdate(&U0, a0, a1, dts, plain_sum != 0);
    } else {
        o->StageUpdate(nullptr, a0, a1, dts, true);
    }
}
double oracle_min_len_to_wavespeed(void *p) { return static_cast<Oracle *>(p)->min_len; }
double oracle_cfl_dt(void *p) { return 0.15 * static_cast<Oracle *>(p)->min_len; }  // include/TimeDisc.h:13,22

void oracle_get_edge_states(void *p, double *out) { Oracle *o = static_cast<Oracle *>(p); std::copy(o->edg.begin(), o->edg.end(), out); }
void oracle_get_sources(void *p, double *out) { Oracle *o = static_cast<Oracle *>(p); std::copy(o->src.begin(), o->src.end(), out); }
void oracle_get_fluxes(void *p, double *out) { Oracle *o = static_cast<Oracle *>(p); std::copy(o->f.begin(), o->f.end(), out); }
void oracle_get_node_max_w(void *p, double *out) { Oracle *o = static_cast<Oracle *>(p); std::copy(o->maxwp.begin(), o->maxwp.end(), out); }
// draining dt of every cell for the snapshot of the last stage (Jacobi) / the live state
void oracle_get_draining_dt(void *p, double *out) {
    Oracle *o = static_cast<Oracle *>(p);
    const std::vector<double> &ref = o->sequential ? o->vol : o->volref;
    for (Idx i = 0; i < o->nt; ++i) out[i] = o->ComputeDrainingDt(ref, i);
}
void oracle_get_branch_counts(void *p, int64_t *out12) {
    Oracle *o = static_cast<Oracle *>(p);
    for (int b = 0; b < 12; ++b) out12[b] = o->branch[b];
}
void oracle_get_cell_class(void *p, int8_t *out) {
    Oracle *o = static_cast<Oracle *>(p);
    for (Idx i = 0; i < o->nt; ++i) out[i] = o->IsDryCell(o->vol, i) ? 0 : (o->IsFullWetCell(i) ? 2 : 1);
}
// geometry taps (for the mesh/geometry parity tests)
void oracle_get_geometry(void *p, double *T3, double *E3, double *L, double *A, double *n0, double *slope) {
    Oracle *o = static_cast<Oracle *>(p);
    if (T3) std::copy(o->Tc.begin(), o->Tc.end(), T3);
    if (E3) std::copy(o->Ec.begin(), o->Ec.end(), E3);
    if (L) std::copy(o->Len_.begin(), o->Len_.end(), L);
    if (A) std::copy(o->Area_.begin(), o->Area_.end(), A);
    if (n0) std::copy(o->norm0.begin(), o->norm0.end(), n0);
    if (slope) std::copy(o->slope.begin(), o->slope.end(), slope);
}
// diagnostics: mass, kinetic, potential (commented ComputeIntegrals, src/SpaceDisc.cpp:77-104, cell-mean form)
void oracle_diagnostics(void *p, double out[6]) {
    Oracle *o = static_cast<Oracle *>(p);
    double mass = 0, kin = 0, pot = 0, vmax = 0, hmin = kInf, wet = 0;
    for (Idx i = 0; i < o->nt; ++i) {
        double h = o->vol[3 * i] - o->vb(i), u = o->vol[3 * i + 1], v = o->vol[3 * i + 2], b = o->vb(i);
        mass += o->Area(i) * h;
        kin += o->Area(i) * (0.5 * h * (u * u + v * v));
        pot += o->Area(i) * (0.5 * h * h + h * b);
        vmax = std::max(vmax, std::max(std::fabs(u), std::fabs(v)));
        hmin = std::min(hmin, h);
        wet += IsWet(h) ? 1. : 0.;
    }
    out[0] = mass; out[1] = kin; out[2] = pot; out[3] = vmax; out[4] = hmin; out[5] = wet;
}

// unit-level helpers
double oracle_cbrt(double x) { return det_cbrt(x); }
int oracle_ilog2_trunc(double x) { return ilog2_trunc(x); }
double oracle_bisection_cubic(double d, double c, double b, double lo, double hi) { return Bisection(CubicPoly(d, c, b), lo, hi); }
double oracle_bisection_cubic2(double d, double c, double b, double lo, double hi, int libm) {
    return Bisection(CubicPoly(d, c, b), lo, hi, 50, libm != 0);
}
// TimeDisc::RHS(i, dt) / one edge flux / wavespeeds on the current fields (live state)
void oracle_rhs(void *p, int64_t i, double dt, double *out3) {
    Oracle *o = static_cast<Oracle *>(p);
    o->RHS(o->vol, i, dt, out3);
}
void oracle_edge_flux(void *p, int flux, int ws, int64_t e, double *F3, double *r) {
    Oracle *o = static_cast<Oracle *>(p);
    o->Flux(flux, ws, e, o->et[2 * e], o->et[2 * e + 1], r, F3);
}
void oracle_wavespeeds(void *p, int ws, double ul, double hl, double ur, double hr, double *a2) {
    static_cast<Oracle *>(p)->Wavespeeds(ws, ul, hl, ur, hr, a2);
}
void oracle_get_draining_dt_live(void *p, double *out) {
    Oracle *o = static_cast<Oracle *>(p);
    for (Idx i = 0; i < o->nt; ++i) out[i] = o->ComputeDrainingDt(o->vol, i);
}
// ConsAssigner::operator= / Get and PrimAssigner on cell i
void oracle_assign_cons(void *p, int64_t i, const double *U3) { static_cast<Oracle *>(p)->cons_set(i, U3); }
void oracle_get_cons(void *p, int64_t i, double *U3) { Oracle *o = static_cast<Oracle *>(p); o->cons_get(o->vol, i, U3); }
void oracle_gradient(const double *P9, double *g2) { Gradient(P9, P9 + 3, P9 + 6, g2); }
void oracle_elem_flux(const double *n2, const double *U3, double *F3) { Oracle::ElemFlux(n2, U3, F3); }
// reconstruction of one cell: kind 0 dry, 1 partwet1, 2 fullwet, 3 partwet2 -> origin(3), G(6)
void oracle_reconstruct(void *p, int kind, int64_t i, double *o3, double *G6) {
    Oracle *o = static_cast<Oracle *>(p);
    MUSCL m = kind == 0 ? o->ReconstructDryCell(i) : kind == 1 ? o->ReconstructPartWetCell1(i)
              : kind == 2 ? o->ReconstructFullWetCell(i) : o->ReconstructPartWetCell2(i);
    for (int c = 0; c < 3; ++c) { o3[c] = m.o[c]; G6[2 * c] = m.G[c][0]; G6[2 * c + 1] = m.G[c][1]; }
}
void oracle_muscl_at_point(void *p, int64_t i, const double *o3, const double *G6, const double *pt3, double *out3) {
    Oracle *o = static_cast<Oracle *>(p);
    MUSCL m{};
    m.i = i;
    for (int c = 0; c < 3; ++c) { m.o[c] = o3[c]; m.G[c][0] = G6[2 * c]; m.G[c][1] = G6[2 * c + 1]; }
    o->AtPoint(m, pt3, out3);
}
void oracle_set_node_max_w(void *p, const double *in) { Oracle *o = static_cast<Oracle *>(p); std::copy(in, in + o->nn, o->maxwp.begin()); }

}  // extern "C"

// -------------------------------------------------------------------------------------------
// Mesh generator + analytic cases + cell-average initial conditions, restated for the oracle so that
// the CPU arm needs nothing from the product library (bench.py --impl reference) and so that the
// product's host / device versions are checked against an independent statement.
// -------------------------------------------------------------------------------------------
namespace {

// StructTriangMesh(ni, nj, h) (include/StructTriangMesh.h:4-15; body missing upstream): conventions of
// SURVEY App. B recovered from notebooks/topology.dat — (ni+1)(nj+1) corner nodes row by row, then ni*nj
// centre nodes; per square the triangles Bottom, Right, Top, Left, CCW, last node = centre; edges numbered
// in order of first visit (triangles in order, k = 0,1,2, edge k joins nodes k and k+1); EdgePoints sorted;
// EdgeTriangs = (later visitor, earlier visitor), boundary (owner, -1). Everything in closed form.
struct StructIdx {
    Idx ni, nj;
    Idx row_base(Idx j) const { return j == 0 ? 0 : (7 * ni + 1) + (j - 1) * (6 * ni + 1); }
    Idx sq_base(Idx j, Idx i) const { return row_base(j) + i * (j == 0 ? 7 : 6) + (i > 0 ? 1 : 0); }
    Idx off(Idx j) const { return j == 0 ? 1 : 0; }
    Idx e_b1(Idx j, Idx i) const { return sq_base(j, i) + off(j); }
    Idx e_b2(Idx j, Idx i) const { return e_b1(j, i) + 1; }
    Idx e_right(Idx j, Idx i) const { return e_b1(j, i) + 2; }
    Idx e_r1(Idx j, Idx i) const { return e_b1(j, i) + 3; }
    Idx e_top(Idx j, Idx i) const { return e_b1(j, i) + 4; }
    Idx e_t1(Idx j, Idx i) const { return e_b1(j, i) + 5; }
    Idx e_bot(Idx j, Idx i) const { return j == 0 ? sq_base(j, i) : e_top(j - 1, i); }
    Idx e_left(Idx j, Idx i) const { return i == 0 ? e_b1(j, i) + 6 : e_right(j, i - 1); }
};

}  // namespace

extern "C" {

void oracle_struct_mesh_sizes(int64_t ni, int64_t nj, int64_t *nn, int64_t *ne, int64_t *nt) {
    *nn = (ni + 1) * (nj + 1) + ni * nj; *ne = 6 * ni * nj + ni + nj; *nt = 4 * ni * nj;
}
void oracle_struct_mesh(int64_t ni, int64_t nj, double h, int64_t i0, int64_t j0, double *geom, int64_t *ep, int64_t *et,
                        int64_t *tp, int64_t *te, int64_t *tt) {
    const Idx nv = (ni + 1) * (nj + 1);
    const StructIdx S{ni, nj};
#pragma omp parallel for schedule(static)
    for (Idx j = 0; j <= nj; ++j)
        for (Idx i = 0; i <= ni; ++i) {
            double *g = &geom[3 * (j * (ni + 1) + i)];
            g[0] = double(i0 + i) * h; g[1] = double(j0 + j) * h; g[2] = 0.;
        }
#pragma omp parallel for schedule(static)
    for (Idx j = 0; j < nj; ++j)
        for (Idx i = 0; i < ni; ++i) {
            const Idx s = j * ni + i;
            double *g = &geom[3 * (nv + s)];
            g[0] = (double(i0 + i) + 0.5) * h; g[1] = (double(j0 + j) + 0.5) * h; g[2] = 0.;
            const Idx v00 = j * (ni + 1) + i, v10 = v00 + 1, v01 = v00 + ni + 1, v11 = v01 + 1, c = nv + s;
            const Idx B = 4 * s, R = B + 1, T = B + 2, L = B + 3;
            const Idx P[4][3] = {{v00, v10, c}, {v10, v11, c}, {v11, v01, c}, {v01, v00, c}};
            const Idx bot = S.e_bot(j, i), b1 = S.e_b1(j, i), b2 = S.e_b2(j, i), rgt = S.e_right(j, i), r1 = S.e_r1(j, i),
                      top = S.e_top(j, i), t1 = S.e_t1(j, i), lft = S.e_left(j, i);
            const Idx E[4][3] = {{bot, b1, b2}, {rgt, r1, b1}, {top, t1, r1}, {lft, b2, t1}};
            const Idx below = j > 0 ? 4 * (s - ni) + 2 : -1, east = i < ni - 1 ? 4 * (s + 1) + 3 : -1,
                      above = j < nj - 1 ? 4 * (s + ni) : -1, west = i > 0 ? 4 * (s - 1) + 1 : -1;
            const Idx N[4][3] = {{below, R, L}, {east, T, B}, {above, L, R}, {west, B, T}};
            for (int q = 0; q < 4; ++q)
                for (int k = 0; k < 3; ++k) { tp[3 * (B + q) + k] = P[q][k]; te[3 * (B + q) + k] = E[q][k]; tt[3 * (B + q) + k] = N[q][k]; }
            auto edge = [&](Idx e, Idx a, Idx b, Idx later, Idx earlier) {
                ep[2 * e] = std::min(a, b); ep[2 * e + 1] = std::max(a, b);
                et[2 * e] = later; et[2 * e + 1] = earlier;
            };
            // every edge is written by exactly one square: the one that visits it first
            if (j == 0) edge(bot, v00, v10, B, -1);
            edge(b1, v10, c, R, B);
            edge(b2, c, v00, L, B);
            edge(rgt, v10, v11, i < ni - 1 ? east : R, i < ni - 1 ? R : -1);
            edge(r1, v11, c, T, R);
            edge(top, v11, v01, j < nj - 1 ? above : T, j < nj - 1 ? T : -1);
            edge(t1, v01, c, L, T);
            if (i == 0) edge(lft, v01, v00, L, -1);
        }
}

// examples/Tests.h: kind 0 LakeAtRestTest (:32-43), 1 ClassicThackerTest (:46-57,135-162,237-280), 2 the Gaussian
// hump of testGaussWave (examples/Main.cpp:183-186, flat bed), 3 the synthetic fully-wet workload of SURVEY §8d,
// 4 BowlTest bed + still lake + hump. par = (mid_x, mid_y, length, cor, tau, delta, H0, p0, q0, level, amp).
// out = (b, h, u, v) at (x, y, t).
void oracle_case_eval(int kind, const double *par, double x, double y, double t, double *out) {
    const double mx = par[0], my = par[1], len = par[2], cor = par[3], delta = par[5], H0 = par[6], p0 = par[7], q0 = par[8],
                 level = par[9], amp = par[10];
    double b = 0., h = 0., u = 0., v = 0.;
    if (kind == 0) {
        b = (1. < x) && (x < 3.) && (1. < y) && (y < 3.) ? -0.2 : -1.;
        h = std::max(0., -b);
    } else if (kind == 1) {
        b = delta * ((x - mx) * (x - mx) + (y - my) * (y - my) - 1.0);
        const double w = std::sqrt(cor * cor + 8. * delta);
        const double qz = (q0 - 0.5 * cor) * (q0 - 0.5 * cor);
        const double rz = qz + 2. * H0 * H0 + p0 * p0 - 0.25 * w * w;
        const double a = std::sqrt(rz * rz + w * w * p0 * p0) / (rz + 0.5 * w * w);
        const double bb = std::atan(w * p0 / rz);
        const double p = 0.5 * w * a * std::sin(w * t + bb) / (1. - a * std::cos(w * t + bb));
        const double q = (q0 - 0.5 * cor) * (1. - a * std::cos(bb)) / (1. - a * std::cos(w * t + bb)) + 0.5 * cor;
        u = p * (x - mx) + q * (y - my);
        v = q * (mx - x) + p * (y - my);
        const double Hc = H0 * (1. - a * std::cos(bb)) / (1. - a * std::cos(w * t + bb));
        const double az0 = (1. - a * std::cos(bb)) * (1. - a * std::cos(bb));
        const double azt = (1. - a * std::cos(w * t + bb)) * (1. - a * std::cos(w * t + bb));
        const double Hxx = (0.25 * w * w * (a * a - 1.) + qz * az0) / azt, Hyy = Hxx, Hxy = 0.;
        const double res = Hc + 0.5 * Hxx * (x - mx) * (x - mx) + Hxy * (x - mx) * (y - my) + 0.5 * Hyy * (y - my) * (y - my);
        h = std::max(0., res);
    } else if (kind == 2) {
        h = 1. + std::exp(-5. * ((x - mx) * (x - mx) + (y - my) * (y - my)));
    } else if (kind == 3) {
        const double two_pi = 6.283185307179586476925286766559;
        b = 0.1 * std::sin(two_pi * x / len) * std::sin(two_pi * y / len) - 1.;
        h = amp * std::exp(-5. * ((x - mx) * (x - mx) + (y - my) * (y - my))) - b;
    } else {
        b = delta * ((x - mx) * (x - mx) + (y - my) * (y - my) - 1.0);
        h = std::max(0., level + amp * std::exp(-5. * ((x - mx) * (x - mx) + (y - my) * (y - my))) - b);
    }
    out[0] = b; out[1] = h; out[2] = u; out[3] = v;
}

}  // extern "C"

namespace {
// TriangAverage<3,n> (include/PointOperations.h:20-44), n at run time: n^2 congruent sub-triangles, the integrand at
// every sub-centroid, in upstream's loop and accumulation order (sum += h*f; result h*sum).
template <class F>
void TriangAverage3(int n, const double *p0, const double *p1, const double *p2, const F &f, double out[3]) {
    const double h = 1. / n;
    double di[3], dj[3], dt[3], pi[3], sum[3] = {0., 0., 0.};
    for (int c = 0; c < 3; ++c) {
        di[c] = h * (p1[c] - p0[c]); dj[c] = h * (p2[c] - p0[c]);
        dt[c] = 1. / 3. * (di[c] + dj[c]); pi[c] = p0[c];
    }
    double v[3], q[3];
    for (int i = 0; i < n; i++) {
        double pt[3] = {pi[0] + dt[0], pi[1] + dt[1], pi[2] + dt[2]};
        for (int j = 0; j < n - i - 1; j++) {
            f(pt, v);
            for (int c = 0; c < 3; ++c) sum[c] += h * v[c];
            for (int c = 0; c < 3; ++c) q[c] = pt[c] + dt[c];
            f(q, v);
            for (int c = 0; c < 3; ++c) sum[c] += h * v[c];
            for (int c = 0; c < 3; ++c) pt[c] += dj[c];
        }
        f(pt, v);
        for (int c = 0; c < 3; ++c) sum[c] += h * v[c];
        for (int c = 0; c < 3; ++c) pi[c] += di[c];
    }
    for (int c = 0; c < 3; ++c) out[c] = h * sum[c];
}
}  // namespace

extern "C" {

// polynomial test integrand (coefficients c[0..5]) for the bit comparison with upstream's TriangAverage
void oracle_triang_average_poly(int n, const double *p0, const double *p1, const double *p2, const double *c, double *out3) {
    TriangAverage3(n, p0, p1, p2, [c](const double *q, double *v) {
        v[0] = c[0] + c[1] * q[0] + c[2] * q[1] * q[1]; v[1] = c[3] * q[0] * q[1]; v[2] = c[4] + c[5] * q[2];
    }, out3);
}

// nodal bathymetry geometry row 2 <- b(x, y) (examples/Main.cpp:204-207, 319-322)
void oracle_case_set_bathymetry(int kind, const double *par, int64_t nn, double *geom) {
#pragma omp parallel for schedule(static)
    for (Idx n = 0; n < nn; ++n) {
        double o[4];
        oracle_case_eval(kind, par, geom[3 * n], geom[3 * n + 1], 0., o);
        geom[3 * n + 2] = o[0];
    }
}

// cell initial state like examples/Main.cpp:211-223: (h, u, v) averaged by TriangAverage<3, quad_n> at time t,
// then w = h_avg + b_i, through PrimAssigner (src/Assigners.cpp:8-20). LakeAtRest: the average of max(0, bed)
// over the cell (examples/Main.cpp:333-336); the Gaussian wave samples w at the centroid (:183-186).
void oracle_case_initial_state(int kind, const double *par, int64_t nt, const double *geom, const int64_t *tp, int quad_n,
                               double t, double *prim) {
    const double third = 1. / 3.;
#pragma omp parallel for schedule(static)
    for (Idx i = 0; i < nt; ++i) {
        const double *p0 = &geom[3 * tp[3 * i]], *p1 = &geom[3 * tp[3 * i + 1]], *p2 = &geom[3 * tp[3 * i + 2]];
        const double cx = p0[0] * third + p1[0] * third + p2[0] * third;
        const double cy = p0[1] * third + p1[1] * third + p2[1] * third;
        const double bi = p0[2] * third + p1[2] * third + p2[2] * third;
        double x[3] = {0., 0., 0.};
        if (kind == 2) {
            double o[4];
            oracle_case_eval(kind, par, cx, cy, t, o);
            x[0] = o[1];
        } else if (kind == 0) {
            if (!(p0[2] <= 0. && p1[2] <= 0. && p2[2] <= 0.)) {
                const double det = (p1[0] - p0[0]) * (p2[1] - p0[1]) - (p2[0] - p0[0]) * (p1[1] - p0[1]);
                TriangAverage3(quad_n, p0, p1, p2, [&](const double *q, double *v) {
                    const double l1 = ((q[0] - p0[0]) * (p2[1] - p0[1]) - (p2[0] - p0[0]) * (q[1] - p0[1])) / det;
                    const double l2 = ((p1[0] - p0[0]) * (q[1] - p0[1]) - (q[0] - p0[0]) * (p1[1] - p0[1])) / det;
                    v[0] = std::max(0., p0[2] + l1 * (p1[2] - p0[2]) + l2 * (p2[2] - p0[2])); v[1] = 0.; v[2] = 0.;
                }, x);
            }
            x[1] = 0.; x[2] = 0.;
        } else {
            TriangAverage3(quad_n, p0, p1, p2, [&](const double *q, double *v) {
                double o[4];
                oracle_case_eval(kind, par, q[0], q[1], t, o);
                v[0] = o[1]; v[1] = o[2]; v[2] = o[3];
            }, x);
            x[0] += bi;
        }
        const double h = x[0] - bi;
        double *o = &prim[3 * i];
        if (!IsWet(h)) { o[0] = bi; o[1] = 0.; o[2] = 0.; continue; }
        o[0] = x[0]; o[1] = x[1]; o[2] = x[2];
        if (h < 1e-3) {
            const double f = std::sqrt(2) * h / std::sqrt(h * h + 1e-6);
            o[1] *= f; o[2] *= f;
        }
    }
}

}  // extern "C"
