// Main.cpp — the reference's drivers (upstream examples/Main.cpp) re-created on the B200 path:
//   testGaussWave  (:172-195)  bowl.msh, flat bed, HLL<Einfeldt>, one Euler step, dumps in the
//                              reference's text format (out0.dat / out1.dat, topology.dat)
//   TestLakeAtRest (:308-373)  StructTriangMesh, HLLC<Einfeldt>, SSPRK3, prints the final error
//   TestThacker                ClassicThackerTest on StructTriangMesh(n), SSPRK2, CFL-driven dt
// Build: g++ -std=c++17 -Iinclude examples/Main.cpp -Lswe_fvm_b200 -lswe_b200 -Wl,-rpath,... -o swe_main
#include <cmath>
#include <cstdio>
#include <fstream>
#include <iostream>
#include <string>

#include "swe/Solvers.h"
#include "swe/Tests.h"

static void dumpFields(const SpaceDisc &sd, const std::string &filename) {  // upstream :65-73
    std::ofstream fout(filename);
    const auto &vol = sd.GetVolField();
    for (Idx i = 0; i < sd.GetDomain().Mesh().NumTriangles(); i++)
        fout << vol.w(i) << '\t' << vol.hu(i) << '\t' << vol.hv(i) << '\n';
}

static void dumpTopology(const Domain &b, const std::string &topologyFile) {  // upstream :146-165
    std::ofstream fout(topologyFile);
    const auto &m = b.Mesh();
    fout << m.NumEdges() << '\n';
    for (Idx i = 0; i < m.NumEdges(); i++) {
        const auto ep = m.EdgePoints(i), et = m.EdgeTriangs(i);
        fout << ep[0] << '\t' << ep[1] << '\t' << et[0] << '\t' << et[1] << '\n';
    }
    fout << m.NumTriangles() << '\n';
    for (Idx i = 0; i < m.NumTriangles(); i++) {
        const auto tp = m.TriangPoints(i), te = m.TriangEdges(i), tt = m.TriangTriangs(i);
        fout << tp[0] << '\t' << tp[1] << '\t' << tp[2] << '\t' << te[0] << '\t' << te[1] << '\t' << te[2] << '\t'
             << tt[0] << '\t' << tt[1] << '\t' << tt[2] << '\n';
    }
}

static void testGaussWave(const std::string &mesh_file) {
    TriangMesh m(mesh_file);
    Domain b(&m);
    for (size_t i = 0; i < b.Size(); i++) b.AtNode(i) = 0.;
    dumpTopology(b, "topology.dat");
    VolumeField v0{b, (size_t)m.NumTriangles()};
    for (Idx i = 0; i < m.NumTriangles(); ++i) {
        const Point t = m.T(i);
        const double r = (t[0] - 4.) * (t[0] - 4.) + (t[1] - 4.) * (t[1] - 4.);
        v0.prim(i) = Array<3>{1. + std::exp(-5. * r), 0., 0.};
    }
    SpaceDisc sd{Fluxes::HLL<Wavespeeds::Einfeldt>, b, v0};
    TimeDisc td{&sd};
    const double dt = 0.001;
    dumpFields(sd, "out0.dat");
    Solvers::Euler(&td, dt);
    dumpFields(sd, "out1.dat");
    std::printf("testGaussWave: %lld cells, one Euler step, CFLdt = %.6e\n", (long long)m.NumTriangles(), td.CFLdt());
}

static void TestLakeAtRest() {
    const size_t n = 32;
    const double l = 4;
    LakeAtRestTest test(0.5 * l, 0.5 * l);
    Domain bathymetry{StructTriangMesh{n, n, l / n}};
    test.SetBathymetry(bathymetry);
    VolumeField v0 = test.InitialState(bathymetry);
    SpaceDisc sd{Fluxes::HLLC<Wavespeeds::Einfeldt>, bathymetry, v0};
    TimeDisc td{&sd};
    const double t_end = 0.5, dt = 1e-3;
    int cnt = 0;
    for (double t = 0.; t < t_end; t += dt, ++cnt) Solvers::SSPRK3(&td, dt);
    const auto &vol = sd.GetVolField();
    double maxv = 0, maxw = 0;
    for (Idx i = 0; i < bathymetry.Mesh().NumTriangles(); ++i) {
        maxv = std::max(maxv, std::max(std::fabs(vol.u(i)), std::fabs(vol.v(i))));
        maxw = std::max(maxw, std::fabs(vol.w(i)));
    }
    std::printf("TestLakeAtRest: %d SSPRK3 steps, max|u|,|v| = %.3e, max|w| = %.3e\n", cnt, maxv, maxw);
}

static void TestThacker(size_t n) {
    const double l = 4;
    ClassicThackerTest test(0.5 * l, 0.5 * l);
    Domain bathymetry{StructTriangMesh{n, n, l / n}};
    test.SetBathymetry(bathymetry);
    VolumeField v0 = test.InitialState(bathymetry, 8);
    SpaceDisc sd{Fluxes::HLLC<Wavespeeds::Einfeldt>, bathymetry, v0};
    TimeDisc td{&sd};
    const double t_end = 0.25 * M_PI / std::sqrt(8.);
    Solvers::SSPRK2(&td, 1e-4);
    double t = 1e-4;
    int cnt = 1;
    while (t < t_end) {
        const double dt = std::min(td.CFLdt(), t_end - t);
        Solvers::SSPRK2(&td, dt);
        t += dt; ++cnt;
    }
    const auto &vol = sd.GetVolField();
    double err = 0, mass = 0;
    for (Idx i = 0; i < bathymetry.Mesh().NumTriangles(); ++i) {
        const Point c = bathymetry.T(i);
        const double A = bathymetry.Area(i), d = vol.h(i) - test.h(c[0], c[1], t);
        err += A * d * d; mass += A * vol.h(i);
    }
    std::printf("TestThacker n=%zu: %d steps to t=%.4f, L2 error of h = %.4e, mass = %.12f\n", n, cnt, t, std::sqrt(err), mass);
}

int main(int argc, char **argv) {
    try {
        const std::string mesh = argc > 1 ? argv[1] : "tests/golden/bowl.msh";
        testGaussWave(mesh);
        TestLakeAtRest();
        TestThacker(64);
    } catch (const MeshError &e) { std::cerr << "Mesh error: " << e.what() << std::endl; return 2;
    } catch (const DomainError &e) { std::cerr << "Domain error: " << e.what() << std::endl; return 3;
    } catch (const SolverError &e) { std::cerr << "Solver error: " << e.what() << std::endl; return 4;
    } catch (const std::exception &e) { std::cerr << "Error: " << e.what() << std::endl; return 1; }
    return 0;
}
