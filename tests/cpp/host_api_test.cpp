// Host-only checks of the C++17 API mirror (include/swe/*.h): mesh classes, Domain geometry, the
// reference's proxy assigners (the commented TestValueFields of upstream examples/Main.cpp:102-123)
// and the analytic Test cases. No GPU needed; returns non-zero on the first failed check.
#include <cmath>
#include <cstdio>
#include <string>
#include <utility>

#include "swe/Tests.h"
#include "swe/Fluxes.h"

static int fails = 0;
#define CHECK(cond)                                                         \
    do {                                                                    \
        if (!(cond)) { std::printf("FAILED %s:%d  %s\n", __FILE__, __LINE__, #cond); ++fails; } \
    } while (0)

int main(int argc, char **argv) {
    const std::string bowl = argc > 1 ? argv[1] : "tests/golden/bowl.msh";
    // StructTriangMesh(ni, nj, h): sizes and conventions
    StructTriangMesh sm(3, 2, 0.5);
    CHECK(sm.NumTriangles() == 24 && sm.NumEdges() == 41 && sm.NumNodes() == 18);
    CHECK(sm.Ni() == 3 && sm.Nj() == 2);
    const Topology topo = sm.GetTopology();
    Idx walls = 0;
    for (Idx e = 0; e < topo.NumEdges(); ++e) {
        const EdgeTag ep = topo.EdgePoints(e), et = topo.EdgeTriangs(e);
        CHECK(ep[0] < ep[1]);
        if (topo.IsEdgeBoundary(e)) { ++walls; CHECK(et[1] == (Idx)Boundaries::SOLID_WALL); }
        else CHECK(et[0] > et[1]);
    }
    CHECK(walls == 10);
    for (Idx t = 0; t < topo.NumTriangles(); ++t) {
        const TriangTag te = topo.TriangEdges(t), tt = topo.TriangTriangs(t);
        for (int k = 0; k < 3; ++k) {
            const EdgeTag et = topo.EdgeTriangs(te[k]);
            CHECK(et[0] == t || et[1] == t);
            CHECK(tt[k] == (et[0] == t ? et[1] : et[0]));
        }
    }
    // Domain geometry: areas sum to the domain, outward normals are unit and opposite on both sides
    Domain d(&sm);
    double area = 0;
    for (Idx t = 0; t < sm.NumTriangles(); ++t) area += d.Area(t);
    CHECK(std::fabs(area - 1.5) < 1e-14);
    for (Idx e = 0; e < sm.NumEdges(); ++e) {
        const EdgeTag et = sm.EdgeTriangs(e);
        const auto n0 = d.Norm(e, et[0]);
        CHECK(std::fabs(n0[0] * n0[0] + n0[1] * n0[1] - 1.) < 1e-14);
        const Point c = d.T(et[0]), m = d.E(e);
        CHECK((m[0] - c[0]) * n0[0] + (m[1] - c[1]) * n0[1] > 0);  // points out of et[0]
        if (et[1] >= 0) { const auto n1 = d.Norm(e, et[1]); CHECK(n1[0] == -n0[0] && n1[1] == -n0[1]); }
    }
    // proxy assigners (src/Assigners.cpp): dry clamp, velocity damping, cons <-> prim
    for (size_t i = 0; i < d.Size(); ++i) d.AtNode(i) = -1.0;
    VolumeField v(d, (size_t)sm.NumTriangles());
    v.prim(0) = Array<3>{0.5, 0.3, -0.2};
    CHECK(v.w(0) == 0.5 && v.u(0) == 0.3 && v.v(0) == -0.2 && std::fabs(v.h(0) - 1.5) < 1e-15);
    const Array<3> c0 = std::as_const(v).cons(0);
    CHECK(std::fabs(c0[1] - 1.5 * 0.3) < 1e-15 && std::fabs(c0[2] + 1.5 * 0.2) < 1e-15);
    v.prim(1) = Array<3>{-1.0 + 5e-13, 1.0, 1.0};  // h <= 1e-12: dry state (b, 0, 0)
    CHECK(v.w(1) == v.b(1) && v.u(1) == 0. && v.v(1) == 0.);
    v.prim(2) = Array<3>{-1.0 + 5e-4, 1.0, 0.};    // 1e-12 < h < 1e-3: velocities damped
    const double h2 = v.h(2);
    CHECK(std::fabs(v.u(2) - std::sqrt(2) * h2 / std::sqrt(h2 * h2 + 1e-6)) < 1e-15);
    v.cons(3) = Array<3>{2.0, 1.0, -4.0};           // ConsAssigner: (h, hu, hv) -> (w, u, v)
    CHECK(v.w(3) == 1.0 && v.u(3) == 0.5 && v.v(3) == -2.0);
    v.cons(3) += Array<3>{0.5, 0.25, 0.};            // += is Get() + rhs, then operator=
    CHECK(std::fabs(v.h(3) - 2.5) < 1e-15 && std::fabs(v.hu(3) - 1.25) < 1e-15);
    v.cons(4) = Array<3>{1e-13, 5.0, 5.0};
    CHECK(v.w(4) == v.b(4) && v.u(4) == 0.);
    // Gmsh reader + refinement through the C++ classes
    TriangMesh g(bowl);
    CHECK(g.NumNodes() == 7555 && g.NumEdges() == 22342 && g.NumTriangles() == 14788);
    TriangMesh r = g.Refine();
    CHECK(r.NumTriangles() == 4 * g.NumTriangles() && r.NumNodes() == g.NumNodes() + g.NumEdges());
    bool threw = false;
    try { TriangMesh bad("/nonexistent.msh"); } catch (const MeshError &) { threw = true; }
    CHECK(threw);
    // analytic cases
    ClassicThackerTest th(2., 2.);
    CHECK(th.b(2., 2.) == -1. && std::fabs(th.h(2., 2., 0.) - 0.5) < 1e-15 && th.u(2.3, 2.1, 0.) == 0.);
    CHECK(std::fabs(th.h(2., 2., M_PI / std::sqrt(8.)) - 0.125) < 1e-14);
    LakeAtRestTest lake(2., 2.);
    CHECK(lake.b(2., 2.) == -0.2 && lake.b(0.5, 0.5) == -1. && lake.w(2., 2., 0.) == 0.);
    Domain ld{StructTriangMesh{16, 16, 0.25}};
    lake.SetBathymetry(ld);
    const VolumeField l0 = lake.InitialState(ld);
    for (Idx t = 0; t < ld.Mesh().NumTriangles(); ++t) CHECK(l0.w(t) == 0. && l0.u(t) == 0.);
    // .ini Parser + DimensionManager + the reference's Test constructor signatures
    {
        const std::string ini = argc > 2 ? argv[2] : "examples/config.ini";
        Parser parser(ini);
        DimensionManager dimer(parser);
        CHECK(parser.Get("Common", "delta") == 1.0 && parser.Get("Thacker", "H0") == 0.5);
        bool perr = false;
        try { parser.Get("Common", "nope"); } catch (const ParserError &) { perr = true; }
        CHECK(perr);
        perr = false;
        try { Parser missing("/nonexistent.ini"); } catch (const ParserError &) { perr = true; }
        CHECK(perr);
        const DimensionManager d2(2.0, 10.0, 4.0);
        CHECK(d2.Scale<Scales::height>(1.5) == 3.0 && d2.Unscale<Scales::length>(5.0) == 0.5);
        CHECK(d2.Scale<Scales::source>(1.0) == 0.4 && d2.Unscale<Scales::time>(5.0) == 2.0);
        ClassicThackerTest a(parser, dimer, 2., 2.), b(2., 2.);
        for (double x : {1.7, 2.0, 2.2}) CHECK(a.h(x, 2.1, 0.3) == b.h(x, 2.1, 0.3) && a.u(x, 2.1, 0.3) == b.u(x, 2.1, 0.3));
        LakeAtRestTest l2(parser, dimer, 2., 2.);
        CHECK(l2.b(2., 2.) == -0.2);
    }
    // the virtual Test hierarchy: built-in cases agree with their C / device description, and a USER-DEFINED case
    // (deriving from BowlTest, like upstream's BallTest would) plugs into SetBathymetry / InitialState
    {
        ClassicThackerTest ct(2., 2., 0.3, 0., 1., 0.5, 0.1, 0.4);
        const swe_case *cc = ct.Builtin();
        CHECK(cc != nullptr);
        for (double x : {1.6, 2.0, 2.35})
            for (double t : {0., 0.2, 0.7}) {
                double o[4];
                swe_case_eval(cc, x, 2.2, t, o);
                CHECK(std::fabs(ct.b(x, 2.2) - o[0]) < 1e-15 && std::fabs(ct.h(x, 2.2, t) - o[1]) < 1e-15);
                CHECK(std::fabs(ct.u(x, 2.2, t) - o[2]) < 1e-15 && std::fabs(ct.v(x, 2.2, t) - o[3]) < 1e-15);
            }
        const Test &base = ct;  // through the abstract interface, as upstream's drivers use it
        CHECK(base.w(2., 2., 0.) == base.h(2., 2., 0.) + base.b(2., 2.) && base.IsWet(2., 2., 0.) && !base.IsWet(0.1, 0.1, 0.));
        struct TiltedPool : BowlTest {  // user case: paraboloid bed, tilted free surface, solid-body rotation
            TiltedPool() : BowlTest(2., 2., 0., 0., 1.) {}
            double u(double, double y, double) const override { return -0.1 * (y - m_mid_y); }
            double v(double x, double, double) const override { return 0.1 * (x - m_mid_x); }
            double h(double x, double y, double) const override { return std::max(0., -0.5 + 0.05 * (x - m_mid_x) - b(x, y)); }
        } pool;
        CHECK(pool.Builtin() == nullptr);
        Domain pd{StructTriangMesh{12, 12, 4. / 12}};
        pool.SetBathymetry(pd);
        CHECK(pd.AtNode(0) == pool.b(0., 0.));
        const VolumeField p0 = pool.InitialState(pd, 3);
        Idx wet = 0, dry = 0;
        for (Idx t = 0; t < pd.Mesh().NumTriangles(); ++t) {
            const Point c = pd.T(t);
            if (p0.h(t) > 0.05) { ++wet; CHECK(std::fabs(p0.w(t) - (-0.5 + 0.05 * (c[0] - 2.))) < 0.05); }
            else if (p0.h(t) == 0.) { ++dry; CHECK(p0.u(t) == 0. && p0.v(t) == 0.); }
        }
        CHECK(wet > 20 && dry > 100);
        // the built-in path and the generic virtual path give the same initial state for a built-in case
        struct ThackerViaVirtuals : ClassicThackerTest {
            using ClassicThackerTest::ClassicThackerTest;
            const swe_case *Builtin() const override { return nullptr; }
        } tv(2., 2.);
        ClassicThackerTest tb(2., 2.);
        Domain td1{StructTriangMesh{10, 10, 0.4}};
        tb.SetBathymetry(td1);
        const VolumeField a0 = tb.InitialState(td1, 4, 0.1), a1 = tv.InitialState(td1, 4, 0.1);
        for (Idx t = 0; t < td1.Mesh().NumTriangles(); ++t)
            CHECK(std::fabs(a0.w(t) - a1.w(t)) < 1e-14 && std::fabs(a0.u(t) - a1.u(t)) < 1e-14 && std::fabs(a0.v(t) - a1.v(t)) < 1e-14);
    }
    // PointOperations.h mirror (upstream include/PointOperations.h:8-51, include/CubicPolyMath.h)
    {
        const Point a{0., 0., 1.}, b{3., 4., 2.}, c{0., 2., 5.};
        CHECK(Len(a, b) == 5. && Det(b, c) == 6. && TriangArea(a, b, c) == 3.);
        const Point x = Intersection(Point{0., 0., 0.}, Point{2., 2., 0.}, Point{2., 0., 0.}, Point{0., 2., 0.});  // den > 0 (S10d)
        CHECK(std::fabs(x[0] - 1.) < 1e-15 && std::fabs(x[1] - 1.) < 1e-15);
        bool par = false;
        try { Intersection(a, b, a, b); } catch (const SolverError &) { par = true; }
        CHECK(par);
        const auto g = Gradient(Point{0., 0., 1.}, Point{1., 0., 3.}, Point{0., 1., 0.5});  // z = 1 + 2x - 0.5y
        CHECK(std::fabs(g[0] - 2.) < 1e-15 && std::fabs(g[1] + 0.5) < 1e-15);
        const auto gp = Gradient(Point{0., 0., 1.}, Point{1e-9, 1., 0.5 + 2e-9}, Point{1., 0., 3.});  // pivoting path
        CHECK(std::fabs(gp[0] - 2.) < 1e-12 && std::fabs(gp[1] + 0.5) < 1e-12);
        CHECK(std::fabs(Bisection(CubicPoly(-0.5), 0., 1.) - std::cbrt(0.5)) < 1e-15);
        CHECK(Bisection(CubicPoly(1.), 0., 1.) == 0.);  // same sign at both ends: the end with the smaller |f|
        const std::function<Array<2>(const Point &)> lin = [](const Point &q) { return Array<2>{1. + 2. * q[0] - q[1], q[2]}; };
        const Array<2> avg = TriangAverage<2, 7>(a, b, c, lin);  // exact for linear integrands: value at the centroid
        CHECK(std::fabs(avg[0] - (1. + 2. * 1. - 2.)) < 1e-13 && std::fabs(avg[1] - 8. / 3.) < 1e-13);
    }
    // flux tags keep the reference's spelling
    constexpr Fluxer f = Fluxes::HLLC<Wavespeeds::Einfeldt>;
    static_assert(f.flux == SWE_HLLC && f.wavespeed == SWE_EINFELDT, "tag mapping");
    constexpr Fluxer f2 = Fluxes::HLL<Wavespeeds::Rusanov>;
    static_assert(f2.flux == SWE_HLL && f2.wavespeed == SWE_RUSANOV, "tag mapping");
    CHECK(f.RegistryId() == 5 && f2.RegistryId() == 0);
    CHECK(swe_fluxer_count() >= 7 && swe_fluxer_find("HLLC<Einfeldt>") == 5);
    const Fluxer llf = Fluxes::Registered("LocalLaxFriedrichs");  // a flux added through csrc/user_fluxes.cuh
    CHECK(llf.id == 6 && llf.RegistryId() == 6);
    bool nof = false;
    try { Fluxes::Registered("NoSuchFlux"); } catch (const DomainError &) { nof = true; }
    CHECK(nof);
    std::printf("%s (%d failed checks)\n", fails ? "FAILED" : "ok", fails);
    return fails ? 1 : 0;
}
