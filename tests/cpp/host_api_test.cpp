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
    // flux tags keep the reference's spelling
    constexpr Fluxer f = Fluxes::HLLC<Wavespeeds::Einfeldt>;
    static_assert(f.flux == SWE_HLLC && f.wavespeed == SWE_EINFELDT, "tag mapping");
    constexpr Fluxer f2 = Fluxes::HLL<Wavespeeds::Rusanov>;
    static_assert(f2.flux == SWE_HLL && f2.wavespeed == SWE_RUSANOV, "tag mapping");
    std::printf("%s (%d failed checks)\n", fails ? "FAILED" : "ok", fails);
    return fails ? 1 : 0;
}
