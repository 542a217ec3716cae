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

#include <chrono>
#include <cstdlib>
#include <cstring>

#include "swe/MultiGpu.h"
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

// A time loop written in upstream's own style (src/Solvers.cpp:5-14) on top of the per-cell accessors: every
// RHS(i, dt) is a tap of the device kernels. One such Euler step equals Solvers::Euler on the device.
static void TestReferenceStyleLoop() {
    const size_t n = 48;
    ClassicThackerTest test(2., 2.);
    Domain b{StructTriangMesh{n, n, 4. / n}};
    test.SetBathymetry(b);
    const VolumeField v0 = test.InitialState(b, 4);
    SpaceDisc sd{Fluxes::HLLC<Wavespeeds::Einfeldt>, b, v0}, sd2{Fluxes::HLLC<Wavespeeds::Einfeldt>, b, v0};
    TimeDisc td{&sd}, td2{&sd2};
    const double dt = 2e-3;
    Idx dry = 0, part = 0, full = 0;
    for (int step = 0; step < 5; ++step) {
        sd.ComputeInterfaceValues();
        sd.ComputeFluxes();
        const auto &m = sd.GetDomain().Mesh();
        for (Idx i = 0; i < m.NumTriangles(); ++i) sd.GetVolFieldForWrite().cons(i) += td.RHS(i, dt);
        sd.Upload();
        Solvers::Euler(&td2, dt);
    }
    double diff = 0;
    const auto &a = sd.GetVolField(), &c = sd2.GetVolField();
    for (Idx i = 0; i < b.Mesh().NumTriangles(); ++i) {
        diff = std::max(diff, std::max(std::fabs(a.w(i) - c.w(i)), std::max(std::fabs(a.u(i) - c.u(i)), std::fabs(a.v(i) - c.v(i)))));
        dry += sd.IsDryCell(i); part += sd.IsPartWetCell(i); full += sd.IsFullWetCell(i);
    }
    std::printf("TestReferenceStyleLoop: 5 Euler steps via cons(i) += RHS(i, dt), max diff to Solvers::Euler = %.3e; cells dry/part/full = %lld/%lld/%lld\n",
                diff, (long long)dry, (long long)part, (long long)full);
}

// The same SpaceDisc / Solvers calls on several GPUs of this process (any mesh, RCB partition): the result must
// equal the single-GPU run bit for bit.
static void TestThackerMultiGpu(size_t n, int ngpus) {
    ClassicThackerTest test(2., 2.);
    Domain b{StructTriangMesh{n, n, 4. / n}};
    test.SetBathymetry(b);
    const VolumeField v0 = test.InitialState(b, 4);
    std::vector<int> devices;  // more ranks than GPUs: several ranks share a device (exercises the full transport on one GPU)
    for (int g = 0; g < ngpus; ++g) devices.push_back(g % std::max(1, (int)swe_device_count()));
    SpaceDisc one{Fluxes::HLLC<Wavespeeds::Einfeldt>, b, v0};
    SpaceDisc many{Fluxes::HLLC<Wavespeeds::Einfeldt>, b, v0, 0., 0., devices};
    TimeDisc td1{&one}, tdn{&many};
    for (int s = 0; s < 20; ++s) { Solvers::SSPRK2(&td1, 2e-3); Solvers::SSPRK2(&tdn, 2e-3); }
    const auto &a = one.GetVolField(), &c = many.GetVolField();
    Idx differ = 0;
    for (Idx i = 0; i < b.Mesh().NumTriangles(); ++i) differ += !(a.w(i) == c.w(i) && a.u(i) == c.u(i) && a.v(i) == c.v(i));
    std::printf("TestThackerMultiGpu n=%zu on %d GPU(s): %lld cells differ from the single-GPU run, CFLdt %.17g vs %.17g\n", n, ngpus,
                (long long)differ, td1.CFLdt(), tdn.CFLdt());
}

// configs[4]: synthetic Thacker domain on StructTriangMesh(n, n) split into strips over the GPUs of this process
// (n = 8192: 268M cells); strips are generated rank by rank, the initial state on the device.
static void TestConfig4(size_t n, int ngpus, int nsteps) {
    ClassicThackerTest test(2., 2.);
    std::vector<int> devices;
    for (int g = 0; g < ngpus; ++g) devices.push_back(g % std::max(1, (int)swe_device_count()));
    const auto t0 = std::chrono::steady_clock::now();
    StripsSolver s{Fluxes::HLLC<Wavespeeds::Einfeldt>, n, n, 4. / n, test, devices};
    const auto t1 = std::chrono::steady_clock::now();
    s.Run(SWE_SSPRK2, 3, 0., 1e-6);  // primes the CFL dt
    const auto t2 = std::chrono::steady_clock::now();
    s.Run(SWE_SSPRK2, nsteps, 0., s.CFLdt());
    const auto t3 = std::chrono::steady_clock::now();
    const double sec = std::chrono::duration<double>(t3 - t2).count();
    std::printf("TestConfig4 n=%zu (%lld cells) on %d GPU(s): set-up %.1f s, %d adaptive SSPRK2 steps in %.3f s = %.3e cell-updates/s, "
                "CFLdt = %.6e, state hash = %016llx\n", n, (long long)s.NumTriangles(), ngpus, std::chrono::duration<double>(t1 - t0).count(),
                nsteps, sec, (double)s.NumTriangles() * nsteps / sec, s.CFLdt(), (unsigned long long)s.StateHash());
}

int main(int argc, char **argv) {
    try {
        // swe_main [mesh.msh] [--gpus N] [--config4 n [steps]]
        std::string mesh = "tests/golden/bowl.msh";
        int gpus = 1, c4steps = 10;
        size_t c4n = 0;
        for (int k = 1; k < argc; ++k) {
            if (!std::strcmp(argv[k], "--gpus") && k + 1 < argc) gpus = std::atoi(argv[++k]);
            else if (!std::strcmp(argv[k], "--config4") && k + 1 < argc) {
                c4n = (size_t)std::atoll(argv[++k]);
                if (k + 1 < argc && argv[k + 1][0] != '-') c4steps = std::atoi(argv[++k]);
            } else mesh = argv[k];
        }
        if (c4n) { TestConfig4(c4n, gpus, c4steps); return 0; }
        testGaussWave(mesh);
        TestLakeAtRest();
        TestThacker(64);
        TestReferenceStyleLoop();
        TestThackerMultiGpu(64, gpus);
    } catch (const MeshError &e) { std::cerr << "Mesh error: " << e.what() << std::endl; return 2;
    } catch (const DomainError &e) { std::cerr << "Domain error: " << e.what() << std::endl; return 3;
    } catch (const SolverError &e) { std::cerr << "Solver error: " << e.what() << std::endl; return 4;
    } catch (const std::exception &e) { std::cerr << "Error: " << e.what() << std::endl; return 1; }
    return 0;
}
