// Tests.h — the analytic test cases of upstream examples/Tests.h (Test base :9-30,
// LakeAtRestTest :32-43, BowlTest :46-57, ThackerTest :135-162, ClassicThackerTest :237-280).
// Upstream reads cor/tau/delta/H0/p0/q0 from a config.ini that is not in its tree; here they are
// plain constructor arguments with the defaults suggested in SURVEY.md App. C.
#pragma once
#include "Bathymetry.h"
#include "DimensionManager.h"
#include "ValueField.h"

class Test {
 protected:
    swe_case m_c{};
    Test(int kind, double mid_x, double mid_y, double length) { swe_case_defaults(&m_c, kind, mid_x, mid_y, length); }
    // the reference's constructor (examples/Tests.h:13-17): cor, tau from [Common], unscaled as sources
    Test(int kind, const Parser &par, const DimensionManager &dim, double mid_x, double mid_y)
        : Test(kind, mid_x, mid_y, 2 * mid_x) {
        m_c.cor = dim.Unscale<Scales::source>(par.Get("Common", "cor"));
        m_c.tau = dim.Unscale<Scales::source>(par.Get("Common", "tau"));
    }

 public:
    virtual ~Test() = default;
    double b(double x, double y) const { return eval(x, y, 0.)[0]; }
    double h(double x, double y, double t) const { return eval(x, y, t)[1]; }
    double u(double x, double y, double t) const { return eval(x, y, t)[2]; }
    double v(double x, double y, double t) const { return eval(x, y, t)[3]; }
    double w(double x, double y, double t) const { const auto e = eval(x, y, t); return e[1] + e[0]; }
    double hu(double x, double y, double t) const { const auto e = eval(x, y, t); return e[1] * e[2]; }
    double hv(double x, double y, double t) const { const auto e = eval(x, y, t); return e[1] * e[3]; }
    bool IsWet(double x, double y, double t) const { return h(x, y, t) >= tol; }
    // nodal bathymetry b(x, y) on every node (examples/Main.cpp:202-205)
    void SetBathymetry(Domain &d) const { swe_detail::check(swe_case_set_bathymetry(&m_c, d.Mesh().Handle())); }
    // cell averages by TriangAverage<3, n> + dry clamp (examples/Main.cpp:211-223)
    VolumeField InitialState(const Domain &d, int quad_n = 4, double t = 0.) const {
        VolumeField v0(d, (size_t)d.Mesh().NumTriangles());
        swe_detail::check(swe_case_initial_state(&m_c, d.Mesh().Handle(), quad_n, t, v0.Raw().data.data()));
        return v0;
    }
    const swe_case &Params() const { return m_c; }

 private:
    std::array<double, 4> eval(double x, double y, double t) const {
        std::array<double, 4> o{};
        swe_detail::check(swe_case_eval(&m_c, x, y, t, o.data()));
        return o;
    }
};

class LakeAtRestTest : public Test {
 public:
    LakeAtRestTest(double mid_x, double mid_y) : Test(SWE_CASE_LAKE_AT_REST, mid_x, mid_y, 2 * mid_x) {}
    LakeAtRestTest(const Parser &par, const DimensionManager &dim, double mid_x, double mid_y)  // examples/Tests.h:34
        : Test(SWE_CASE_LAKE_AT_REST, par, dim, mid_x, mid_y) {}
};

class ClassicThackerTest : public Test {
 public:
    ClassicThackerTest(double mid_x, double mid_y, double cor = 0., double tau = 0., double delta = 1., double H0 = 0.5,
                       double p0 = 0., double q0 = 0.)
        : Test(SWE_CASE_CLASSIC_THACKER, mid_x, mid_y, 2 * mid_x) {
        m_c.cor = cor; m_c.tau = tau; m_c.delta = delta; m_c.H0 = H0; m_c.p0 = p0; m_c.q0 = q0;
    }
    // examples/Tests.h:242 with BowlTest (:49-51, [Common] delta) and ThackerTest (:138-142, [Thacker] H0 p0 q0)
    ClassicThackerTest(const Parser &par, const DimensionManager &dim, double mid_x, double mid_y)
        : Test(SWE_CASE_CLASSIC_THACKER, par, dim, mid_x, mid_y) {
        m_c.delta = par.Get("Common", "delta");
        m_c.H0 = dim.Unscale<Scales::height>(par.Get("Thacker", "H0"));
        m_c.p0 = dim.Unscale<Scales::source>(par.Get("Thacker", "p0"));
        m_c.q0 = dim.Unscale<Scales::source>(par.Get("Thacker", "q0"));
    }
};

class GaussWaveTest : public Test {  // the IC of testGaussWave (examples/Main.cpp:183-186)
 public:
    GaussWaveTest(double mid_x, double mid_y) : Test(SWE_CASE_GAUSS_WAVE, mid_x, mid_y, 2 * mid_x) {}
};

class BowlHumpTest : public Test {  // BowlTest bed + still lake + Gaussian hump (config 2)
 public:
    BowlHumpTest(double mid_x, double mid_y, double delta = 1., double level = 3., double amp = 0.5)
        : Test(SWE_CASE_BOWL_HUMP, mid_x, mid_y, 2 * mid_x) { m_c.delta = delta; m_c.level = level; m_c.amp = amp; }
};
