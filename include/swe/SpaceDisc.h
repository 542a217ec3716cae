// SpaceDisc.h — upstream include/SpaceDisc.h:20-48 (and its base MUSCLObject, include/MUSCLObject.h:4-69) over the
// device context. The constructor uploads mesh + bathymetry + initial state once; ComputeInterfaceValues /
// ComputeFluxes launch the reconstruction and flux kernels; the Get* accessors download (parity taps, not part of
// the time loop). With a list of devices the same object runs on N GPUs of this process (swe_dist_group_*: RCB
// partition, halo cells, peer-memory exchange fused into the stage) — Solvers::X(&td, dt) is unchanged.
#pragma once
#include <memory>
#include <string>

#include "Fluxes.h"
#include "ValueField.h"

struct SpaceDisc {
    SpaceDisc(const Fluxer &fluxer, const Domain &b, const VolumeField &v0, double cor = 0, double tau = 0,
              int device = 0, bool reorder = true)
        : m_fluxer(fluxer), m_b(b), m_cor(cor), m_tau(tau), m_vol(v0) {
        swe_mesh mv = b.Mesh().View();
        mv.cor = cor; mv.tau = tau;
        swe_detail::check(swe_create(&m_ctx, &mv, device, reorder ? 1 : 0));
        if (m_fluxer.id >= 0) swe_detail::check(swe_set_fluxer(m_ctx, m_fluxer.id), m_ctx);
        Upload();
    }
    // N GPUs driven by this process: cells partitioned by recursive coordinate bisection, one rank per device
    SpaceDisc(const Fluxer &fluxer, const Domain &b, const VolumeField &v0, double cor, double tau,
              const std::vector<int> &devices, bool reorder = true)
        : m_fluxer(fluxer), m_b(b), m_cor(cor), m_tau(tau), m_vol(v0) {
        const int32_t world = (int32_t)devices.size();
        if (world < 1) throw DomainError("SpaceDisc: empty device list");
        std::vector<int32_t> part((size_t)b.Mesh().NumTriangles());
        swe_detail::check(swe_partition_rcb(b.Mesh().Handle(), world, part.data()));
        std::vector<swe_dist_plan *> plans((size_t)world, nullptr);
        for (int32_t r = 0; r < world; ++r) swe_detail::check(swe_dist_plan_mesh(&plans[r], r, world, b.Mesh().Handle(), part.data()));
        swe_dist_config cfg{};
        cfg.reorder = reorder ? 1 : 0; cfg.overlap = 1; cfg.cor = cor; cfg.tau = tau;
        std::vector<int32_t> devs(devices.begin(), devices.end());
        m_ranks.assign((size_t)world, nullptr);
        const int rc = swe_dist_group_create(m_ranks.data(), plans.data(), devs.data(), world, &cfg);
        if (rc != SWE_OK) {
            for (auto *p : plans) swe_dist_plan_free(p);
            m_ranks.clear();
            dist_check(rc, nullptr);
        }
        if (m_fluxer.id >= 0)
            for (auto *d : m_ranks) swe_detail::check(swe_set_fluxer(swe_dist_ctx(d), m_fluxer.id), swe_dist_ctx(d));
        Upload();
    }
    ~SpaceDisc() {
        for (auto *d : m_ranks) swe_dist_destroy(d);
        if (m_ctx) swe_destroy(m_ctx);
    }
    SpaceDisc(const SpaceDisc &) = delete;
    SpaceDisc &operator=(const SpaceDisc &) = delete;

    bool IsDistributed() const noexcept { return !m_ranks.empty(); }
    const std::vector<swe_dist *> &Ranks() const noexcept { return m_ranks; }
    const Domain &GetDomain() const noexcept { return m_b; }

    // host copy of the state, refreshed from the device when it is stale
    const VolumeField &GetVolField() const {
        if (!m_host_valid) {
            if (IsDistributed()) for (auto *d : m_ranks) dist_check(swe_dist_get_owned_state(d, m_vol.Raw().data.data()), d);
            else swe_detail::check(swe_get_state(m_ctx, m_vol.Raw().data.data()), m_ctx);
            m_host_valid = true;
        }
        return m_vol;
    }
    VolumeField &GetVolFieldForWrite() { GetVolField(); return m_vol; }  // edit (e.g. cons(i) += td->RHS(i, dt)), then Upload()
    void Upload() {
        if (IsDistributed()) for (auto *d : m_ranks) dist_check(swe_dist_set_state_global(d, m_vol.Raw().data.data()), d);
        else swe_detail::check(swe_set_state(m_ctx, m_vol.Raw().data.data()), m_ctx);
        m_host_valid = true;
        ++m_epoch; m_cls_epoch = 0;
    }
    // the device state changed behind the host copy (called by Solvers::*)
    void Touch() { m_host_valid = false; ++m_epoch; }
    uint64_t Epoch() const noexcept { return m_epoch; }

    // MUSCLObject::IsDryCell / IsFullWetCell / IsPartWetCell (include/MUSCLObject.h:11-13) of the current state
    bool IsDryCell(Idx i) const { return CellClass(i) == 0; }
    bool IsPartWetCell(Idx i) const { return CellClass(i) == 1; }
    bool IsFullWetCell(Idx i) const { return CellClass(i) == 2; }

    void ComputeInterfaceValues() { single("ComputeInterfaceValues"); swe_detail::check(swe_compute_interface_values(m_ctx), m_ctx); ++m_epoch; }
    void ComputeFluxes() { single("ComputeFluxes"); swe_detail::check(swe_compute_fluxes(m_ctx, m_fluxer.flux, m_fluxer.wavespeed), m_ctx); ++m_epoch; }

    Storage<3> GetFluxes() const { single("GetFluxes"); Storage<3> f((size_t)m_b.Mesh().NumEdges()); swe_detail::check(swe_get_fluxes(m_ctx, f.data.data()), m_ctx); return f; }
    Storage<3> GetEdgField() const { single("GetEdgField"); Storage<3> f(2 * (size_t)m_b.Mesh().NumEdges()); swe_detail::check(swe_get_edge_states(m_ctx, f.data.data()), m_ctx); return f; }
    Storage<3> GetSrcField() const { single("GetSrcField"); Storage<3> f(2 * (size_t)m_b.Mesh().NumEdges()); swe_detail::check(swe_get_sources(m_ctx, f.data.data()), m_ctx); return f; }
    void EnableTaps(bool on = true) { single("EnableTaps"); swe_detail::check(swe_enable_taps(m_ctx, on ? 1 : 0), m_ctx); }
    double GetTau() const noexcept { return m_tau; }
    double GetCor() const noexcept { return m_cor; }
    double GetMinLenToWavespeed() const {
        double v;
        swe_ctx *c = IsDistributed() ? swe_dist_ctx(m_ranks[0]) : m_ctx;
        swe_detail::check(swe_get_min_len_to_wavespeed(c, &v), c);
        return v;
    }
    const Fluxer &GetFluxer() const noexcept { return m_fluxer; }
    swe_ctx *Context() const noexcept { return m_ctx; }

    // the switches for the places where upstream HEAD is unfinished: "recon", "pw2", "roe_fix", "cfl_abs"
    void SetOption(const char *key, int value) {
        if (IsDistributed()) for (auto *d : m_ranks) swe_detail::check(swe_set_option(swe_dist_ctx(d), key, value), swe_dist_ctx(d));
        else swe_detail::check(swe_set_option(m_ctx, key, value), m_ctx);
    }
    // binary checkpoint / restart: state, time, dt, min_len, settings (upstream only has text dumps)
    void SaveCheckpoint(const std::string &path) const { single("SaveCheckpoint"); swe_detail::check(swe_checkpoint_save(m_ctx, path.c_str()), m_ctx); }
    void LoadCheckpoint(const std::string &path) { single("LoadCheckpoint"); swe_detail::check(swe_checkpoint_load(m_ctx, path.c_str()), m_ctx); Touch(); m_cls_epoch = 0; }
    double Time() const { single("Time"); double t; swe_detail::check(swe_get_time(m_ctx, &t), m_ctx); return t; }

    static void dist_check(int rc, const swe_dist *d) {
        if (rc == SWE_OK) return;
        const char *m = swe_dist_last_error(d);
        const std::string msg = (m && *m) ? m : "multi-GPU error";
        switch (rc) {
            case SWE_ERR_INVALID: throw DomainError(msg);
            case SWE_ERR_NUMERIC: throw SolverError(msg);
            default: throw DeviceError(msg);
        }
    }

 private:
    void single(const char *what) const {
        if (IsDistributed()) throw DomainError(std::string(what) + " is a single-GPU tap; not available on a multi-GPU SpaceDisc");
    }
    int CellClass(Idx i) const {
        if (m_cls_epoch != m_epoch) {
            m_cls.resize((size_t)m_b.Mesh().NumTriangles());
            if (IsDistributed()) {  // same predicates on the host copy (src/MUSCLObject.cpp:13-29)
                const VolumeField &v = GetVolField();
                const Topology t = m_b.GetTopology();
                for (Idx c = 0; c < (Idx)m_cls.size(); ++c) {
                    if (!IsWet(v.h(c))) { m_cls[(size_t)c] = 0; continue; }
                    const TriangTag tp = t.TriangPoints(c);
                    const double bmax = std::max(std::max(m_b.P(tp[0])[2], m_b.P(tp[1])[2]), m_b.P(tp[2])[2]);
                    m_cls[(size_t)c] = (!t.IsTriangleBoundary(c) && bmax < v.w(c)) ? 2 : 1;
                }
            } else {
                swe_detail::check(swe_classify(m_ctx, m_cls.data()), m_ctx);
            }
            m_cls_epoch = m_epoch;
        }
        return m_cls[(size_t)i];
    }

    Fluxer m_fluxer;
    const Domain &m_b;
    double m_cor, m_tau;
    mutable VolumeField m_vol;
    mutable bool m_host_valid = false;
    mutable std::vector<int8_t> m_cls;
    mutable uint64_t m_cls_epoch = 0;
    uint64_t m_epoch = 1;
    swe_ctx *m_ctx = nullptr;
    std::vector<swe_dist *> m_ranks;
};
