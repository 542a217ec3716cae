// Tests.h — the analytic test cases of upstream examples/Tests.h with the same class hierarchy:
//   Test (abstract, :9-30) -> LakeAtRestTest (:32-43)
//                          -> BowlTest (abstract, :46-57) -> ThackerTest (abstract, :135-162) -> ClassicThackerTest (:237-280)
// b/u/v/h are virtual exactly as upstream (:20-24), so a USER-DEFINED case plugs in by deriving from Test (or
// BowlTest / ThackerTest) — SetBathymetry and InitialState then evaluate it on the host (TriangAverage,
// include/PointOperations.h:20-44). The cases named by the benchmark configs additionally carry a device
// description (Builtin()), which lets swe_case_* build bathymetry and initial state on the host in C or on the GPU
// (swe_case_initial_state_device) without a per-point virtual call.
// Upstream reads cor/tau/delta/H0/p0/q0 from a config.ini that is not in its tree; both constructor forms exist:
// plain arguments with the defaults of SURVEY.md App. C, and upstream's (Parser, DimensionManager, mid_x, mid_y).
#pragma once
#include <algorithm>
#include <cmath>

#include "Bathymetry.h"
#include "DimensionManager.h"
#include "PointOperations.h"
#include "ValueField.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

class Test {
 protected:
    double m_mid_x, m_mid_y;  // location
    double m_cor, m_tau;      // sources
    Test(double mid_x, double mid_y, double cor = 0., double tau = 0.) : m_mid_x(mid_x), m_mid_y(mid_y), m_cor(cor), m_tau(tau) {}
    Test(const Parser &par, const DimensionManager &dim, double mid_x, double mid_y)  // examples/Tests.h:13-17
        : m_mid_x(mid_x), m_mid_y(mid_y),
          m_cor(dim.Unscale<Scales::source>(par.Get("Common", "cor"))),
          m_tau(dim.Unscale<Scales::source>(par.Get("Common", "tau"))) {}

 public:
    virtual ~Test() = default;
    virtual double b(double x, double y) const = 0;
    virtual double u(double x, double y, double t) const = 0;
    virtual double v(double x, double y, double t) const = 0;
    virtual double h(double x, double y, double t) const = 0;
    double w(double x, double y, double t) const { return h(x, y, t) + b(x, y); }
    double hu(double x, double y, double t) const { return h(x, y, t) * u(x, y, t); }
    double hv(double x, double y, double t) const { return h(x, y, t) * v(x, y, t); }
    bool IsWet(double x, double y, double t) const { return h(x, y, t) >= tol; }
    double Cor() const { return m_cor; }
    double Tau() const { return m_tau; }

    // device / C description of the case, or nullptr for a user-defined one
    virtual const swe_case *Builtin() const { return nullptr; }
    // the averaging rule of the lake-at-rest driver (examples/Main.cpp:333-336) instead of examples/Main.cpp:211-223
    virtual bool AveragesClippedBed() const { return false; }

    // nodal bathymetry b(x, y) on every node (examples/Main.cpp:202-205)
    void SetBathymetry(Domain &d) const {
        if (const swe_case *c = Builtin()) { swe_detail::check(swe_case_set_bathymetry(c, d.Mesh().Handle())); return; }
        for (size_t n = 0; n < d.Size(); ++n) { const Point p = d.P((NodeTag)n); d.AtNode((NodeTag)n) = b(p[0], p[1]); }
    }
    // cell averages of (h, u, v) by TriangAverage<3, n>, w = h_avg + b_i, through PrimAssigner (examples/Main.cpp:211-223)
    VolumeField InitialState(const Domain &d, int quad_n = 4, double t = 0.) const {
        VolumeField v0(d, (size_t)d.Mesh().NumTriangles());
        if (const swe_case *c = Builtin()) {
            swe_detail::check(swe_case_initial_state(c, d.Mesh().Handle(), quad_n, t, v0.Raw().data.data()));
            return v0;
        }
        const auto f = [&](const Point &p) { return Array<3>{h(p[0], p[1], t), u(p[0], p[1], t), v(p[0], p[1], t)}; };
        for (Idx i = 0; i < d.Mesh().NumTriangles(); ++i) {
            const TriangTag tp = d.Mesh().TriangPoints(i);
            Array<3> a = Average(d.P(tp[0]), d.P(tp[1]), d.P(tp[2]), quad_n, f);
            a[0] += v0.b(i);
            v0.prim(i) = a;
        }
        return v0;
    }
    // TriangAverage<3, n> with n chosen at run time (same loop as include/PointOperations.h:20-44)
    static Array<3> Average(const Point &p0, const Point &p1, const Point &p2, int n, const std::function<Array<3>(const Point &)> &f) {
        const double hq = 1. / n;
        const Point di = hq * (p1 - p0), dj = hq * (p2 - p0), dt = 1. / 3. * (di + dj);
        Array<3> sum;
        Point pi = p0;
        for (int i = 0; i < n; i++) {
            Point pt = pi + dt;
            for (int j = 0; j < n - i - 1; j++) {
                sum += hq * f(pt);
                sum += hq * f(pt + dt);
                pt += dj;
            }
            sum += hq * f(pt);
            pi += di;
        }
        return hq * sum;
    }
};

class LakeAtRestTest : public Test {  // examples/Tests.h:32-43
 public:
    LakeAtRestTest(double mid_x, double mid_y) : Test(mid_x, mid_y) { init(); }
    LakeAtRestTest(const Parser &par, const DimensionManager &dim, double mid_x, double mid_y) : Test(par, dim, mid_x, mid_y) { init(); }
    double b(double x, double y) const override { return (1. < x) && (x < 3.) && (1. < y) && (y < 3.) ? -0.2 : -1.; }
    double u(double, double, double) const override { return 0.; }
    double v(double, double, double) const override { return 0.; }
    double h(double x, double y, double) const override { return std::max(0., -b(x, y)); }
    const swe_case *Builtin() const override { return &m_c; }
    bool AveragesClippedBed() const override { return true; }

 private:
    void init() { swe_case_defaults(&m_c, SWE_CASE_LAKE_AT_REST, m_mid_x, m_mid_y, 2 * m_mid_x); m_c.cor = m_cor; m_c.tau = m_tau; }
    swe_case m_c{};
};

class BowlTest : public Test {  // abstract, examples/Tests.h:46-57: paraboloid bed delta (r^2 - 1)
 protected:
    double m_delta;
    BowlTest(double mid_x, double mid_y, double cor, double tau, double delta) : Test(mid_x, mid_y, cor, tau), m_delta(delta) {}
    BowlTest(const Parser &par, const DimensionManager &dim, double mid_x, double mid_y)
        : Test(par, dim, mid_x, mid_y), m_delta(par.Get("Common", "delta")) {}

 public:
    double b(double x, double y) const override {
        return m_delta * ((x - m_mid_x) * (x - m_mid_x) + (y - m_mid_y) * (y - m_mid_y) - 1.0);
    }
    double Delta() const { return m_delta; }
};

class ThackerTest : public BowlTest {  // abstract, examples/Tests.h:135-162
 protected:
    double H0, p0, q0;
    ThackerTest(double mid_x, double mid_y, double cor, double tau, double delta, double H0_, double p0_, double q0_)
        : BowlTest(mid_x, mid_y, cor, tau, delta), H0(H0_), p0(p0_), q0(q0_) {}
    ThackerTest(const Parser &par, const DimensionManager &dim, double mid_x, double mid_y)
        : BowlTest(par, dim, mid_x, mid_y),
          H0(dim.Unscale<Scales::height>(par.Get("Thacker", "H0"))),
          p0(dim.Unscale<Scales::source>(par.Get("Thacker", "p0"))),
          q0(dim.Unscale<Scales::source>(par.Get("Thacker", "q0"))) {}

 public:
    virtual double p(double t) const = 0;
    virtual double q(double t) const = 0;
    double u(double x, double y, double t) const override { return p(t) * (x - m_mid_x) + q(t) * (y - m_mid_y); }
    virtual double Hc(double t) const = 0;
    virtual double Hxx(double t) const = 0;
    virtual double Hxy(double t) const = 0;
    virtual double Hyy(double t) const = 0;
    double h(double x, double y, double t) const override {
        const double res = Hc(t) + 0.5 * Hxx(t) * (x - m_mid_x) * (x - m_mid_x) + Hxy(t) * (x - m_mid_x) * (y - m_mid_y) +
                           0.5 * Hyy(t) * (y - m_mid_y) * (y - m_mid_y);
        return std::max(0., res);
    }
};

class ClassicThackerTest : public ThackerTest {  // examples/Tests.h:237-280
 public:
    ClassicThackerTest(double mid_x, double mid_y, double cor = 0., double tau = 0., double delta = 1., double H0_ = 0.5,
                       double p0_ = 0., double q0_ = 0.)
        : ThackerTest(mid_x, mid_y, cor, tau, delta, H0_, p0_, q0_) { init(); }
    ClassicThackerTest(const Parser &par, const DimensionManager &dim, double mid_x, double mid_y)
        : ThackerTest(par, dim, mid_x, mid_y) { init(); }
    double p(double t) const override { return 0.5 * m_w * m_a * std::sin(m_w * t + m_phi) / (1. - m_a * std::cos(m_w * t + m_phi)); }
    double q(double t) const override {
        return (q0 - 0.5 * m_cor) * (1. - m_a * std::cos(m_phi)) / (1. - m_a * std::cos(m_w * t + m_phi)) + 0.5 * m_cor;
    }
    double v(double x, double y, double t) const override { return q(t) * (m_mid_x - x) + p(t) * (y - m_mid_y); }
    double Hc(double t) const override { return H0 * (1. - m_a * std::cos(m_phi)) / (1. - m_a * std::cos(m_w * t + m_phi)); }
    double Hxx(double t) const override {
        const double qz = (q0 - 0.5 * m_cor) * (q0 - 0.5 * m_cor);
        const double az0 = (1. - m_a * std::cos(m_phi)) * (1. - m_a * std::cos(m_phi));
        const double azt = (1. - m_a * std::cos(m_w * t + m_phi)) * (1. - m_a * std::cos(m_w * t + m_phi));
        return (0.25 * m_w * m_w * (m_a * m_a - 1.) + qz * az0) / azt;
    }
    double Hxy(double) const override { return 0; }
    double Hyy(double t) const override { return Hxx(t); }
    double Omega() const { return m_w; }  // angular frequency: the driver's end time is pi / omega (examples/Main.cpp:268)
    const swe_case *Builtin() const override { return &m_c; }

 private:
    void init() {
        m_w = std::sqrt(m_cor * m_cor + 8. * m_delta);
        const double qz = (q0 - 0.5 * m_cor) * (q0 - 0.5 * m_cor);
        const double rz = qz + 2. * H0 * H0 + p0 * p0 - 0.25 * m_w * m_w;
        m_a = std::sqrt(rz * rz + m_w * m_w * p0 * p0) / (rz + 0.5 * m_w * m_w);
        m_phi = std::atan(m_w * p0 / rz);
        swe_case_defaults(&m_c, SWE_CASE_CLASSIC_THACKER, m_mid_x, m_mid_y, 2 * m_mid_x);
        m_c.cor = m_cor; m_c.tau = m_tau; m_c.delta = m_delta; m_c.H0 = H0; m_c.p0 = p0; m_c.q0 = q0;
    }
    double m_w = 0, m_a = 0, m_phi = 0;
    swe_case m_c{};
};

class GaussWaveTest : public Test {  // the IC of testGaussWave (examples/Main.cpp:183-186): flat bed, hump sampled at the centroid
 public:
    GaussWaveTest(double mid_x, double mid_y) : Test(mid_x, mid_y) { swe_case_defaults(&m_c, SWE_CASE_GAUSS_WAVE, mid_x, mid_y, 2 * mid_x); }
    double b(double, double) const override { return 0.; }
    double u(double, double, double) const override { return 0.; }
    double v(double, double, double) const override { return 0.; }
    double h(double x, double y, double) const override {
        return 1. + std::exp(-5. * ((x - m_mid_x) * (x - m_mid_x) + (y - m_mid_y) * (y - m_mid_y)));
    }
    const swe_case *Builtin() const override { return &m_c; }

 private:
    swe_case m_c{};
};

class BowlHumpTest : public BowlTest {  // BowlTest bed + still lake at `level` + Gaussian hump (config 2)
 public:
    BowlHumpTest(double mid_x, double mid_y, double delta = 1., double level = 3., double amp = 0.5)
        : BowlTest(mid_x, mid_y, 0., 0., delta), m_level(level), m_amp(amp) {
        swe_case_defaults(&m_c, SWE_CASE_BOWL_HUMP, mid_x, mid_y, 2 * mid_x);
        m_c.delta = delta; m_c.level = level; m_c.amp = amp;
    }
    double u(double, double, double) const override { return 0.; }
    double v(double, double, double) const override { return 0.; }
    double h(double x, double y, double) const override {
        const double r2 = (x - m_mid_x) * (x - m_mid_x) + (y - m_mid_y) * (y - m_mid_y);
        return std::max(0., m_level + m_amp * std::exp(-5. * r2) - b(x, y));
    }
    const swe_case *Builtin() const override { return &m_c; }

 private:
    double m_level, m_amp;
    swe_case m_c{};
};
