// MultiGpu.h — configs[4]-size structured runs on several GPUs of one process.
// A StructTriangMesh(ni, nj, h) with hundreds of millions of cells is never built as ONE host mesh: every rank's
// strip (its rows of squares + 3 halo rows per open side) is generated directly (swe_dist_plan_struct), bathymetry
// and initial state come from a built-in Test evaluated on the device (TriangAverage<3, quad_n> as a kernel).
// For meshes that fit on the host, SpaceDisc itself takes a device list (any mesh, RCB partition).
#pragma once
#include <string>
#include <vector>

#include "Solvers.h"
#include "Tests.h"

class StripsSolver {
 public:
    StripsSolver(const Fluxer &fluxer, size_t ni, size_t nj, double h, const Test &test, const std::vector<int> &devices,
                 int quad_n = 1, bool reorder = true)
        : m_fluxer(fluxer), m_nt(4 * (Idx)ni * (Idx)nj) {
        const swe_case *c = test.Builtin();
        if (!c) throw DomainError("StripsSolver needs a built-in Test (device-side initial state)");
        const int32_t world = (int32_t)devices.size();
        std::vector<swe_dist_plan *> plans((size_t)world, nullptr);
        for (int32_t r = 0; r < world; ++r) swe_detail::check(swe_dist_plan_struct(&plans[r], r, world, (Idx)ni, (Idx)nj, h));
        swe_dist_config cfg{};
        cfg.reorder = reorder ? 1 : 0; cfg.overlap = 1; cfg.cor = test.Cor(); cfg.tau = test.Tau();
        std::vector<int32_t> devs(devices.begin(), devices.end());
        m_ranks.assign((size_t)world, nullptr);
        const int rc = swe_dist_group_create(m_ranks.data(), plans.data(), devs.data(), world, &cfg);
        if (rc != SWE_OK) { for (auto *p : plans) swe_dist_plan_free(p); m_ranks.clear(); SpaceDisc::dist_check(rc, nullptr); }
        for (auto *d : m_ranks) {
            swe_ctx *x = swe_dist_ctx(d);
            if (fluxer.id >= 0) swe_detail::check(swe_set_fluxer(x, fluxer.id), x);
            swe_detail::check(swe_case_set_bathymetry_device(x, c), x);
            swe_detail::check(swe_case_initial_state_device(x, c, quad_n, 0.), x);
        }
        for (auto *d : m_ranks) SpaceDisc::dist_check(swe_dist_exchange(d), d);
    }
    ~StripsSolver() { for (auto *d : m_ranks) swe_dist_destroy(d); }
    StripsSolver(const StripsSolver &) = delete;
    StripsSolver &operator=(const StripsSolver &) = delete;

    Idx NumTriangles() const { return m_nt; }
    int NumGpus() const { return (int)m_ranks.size(); }
    // nsteps steps on all GPUs; dt <= 0: dt = CFLdt() of the previous step (global minimum), first step dt0
    void Run(swe_scheme scheme, Idx nsteps, double dt, double dt0 = 0.) {
        int rc = swe_dist_group_run(m_ranks.data(), (int32_t)m_ranks.size(), scheme, m_fluxer.flux, m_fluxer.wavespeed, nsteps, dt, dt0);
        for (size_t k = 0; k < m_ranks.size() && rc != SWE_OK; ++k) SpaceDisc::dist_check(rc, m_ranks[k]);
        for (auto *d : m_ranks) SpaceDisc::dist_check(swe_dist_synchronize(d), d);
    }
    double CFLdt() const { double dt; SpaceDisc::dist_check(swe_dist_cfl_dt(m_ranks[0], &dt), m_ranks[0]); return dt; }
    uint64_t StateHash() const {  // order-independent hash of all owned cells: equal for any number of GPUs
        uint64_t tot = 0;
        for (auto *d : m_ranks) { uint64_t h = 0; SpaceDisc::dist_check(swe_dist_state_hash(d, &h), d); tot += h; }
        return tot;
    }
    double Mass() const {  // sum over the ranks' OWNED cells is not available from swe_diagnostics (it covers local cells);
        double m = 0;      // report rank-local masses instead: each one is conserved up to the exchange through the cuts
        for (auto *d : m_ranks) { double o[6]; swe_detail::check(swe_diagnostics(swe_dist_ctx(d), o), swe_dist_ctx(d)); m += o[0]; }
        return m;
    }

 private:
    Fluxer m_fluxer;
    Idx m_nt;
    std::vector<swe_dist *> m_ranks;
};
