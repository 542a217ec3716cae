// SpaceDisc.h — upstream include/SpaceDisc.h:20-48 over the device context. The constructor
// uploads mesh + bathymetry + initial state once; ComputeInterfaceValues / ComputeFluxes launch
// the reconstruction and flux kernels; the Get* accessors download (they are parity taps, not
// part of the time loop).
#pragma once
#include "Fluxes.h"
#include "ValueField.h"

struct SpaceDisc {
    SpaceDisc(const Fluxer &fluxer, const Domain &b, const VolumeField &v0, double cor = 0, double tau = 0,
              int device = 0, bool reorder = true)
        : m_fluxer(fluxer), m_b(b), m_cor(cor), m_tau(tau), m_vol(v0) {
        swe_mesh mv = b.Mesh().View();
        mv.cor = cor; mv.tau = tau;
        swe_detail::check(swe_create(&m_ctx, &mv, device, reorder ? 1 : 0));
        Upload();
    }
    ~SpaceDisc() { swe_destroy(m_ctx); }
    SpaceDisc(const SpaceDisc &) = delete;
    SpaceDisc &operator=(const SpaceDisc &) = delete;

    const Domain &GetDomain() const noexcept { return m_b; }
    // host copy of the state, refreshed from the device on every call
    const VolumeField &GetVolField() const { swe_detail::check(swe_get_state(m_ctx, m_vol.Raw().data.data()), m_ctx); return m_vol; }
    VolumeField &GetVolFieldForWrite() { return m_vol; }  // edit, then Upload()
    void Upload() { swe_detail::check(swe_set_state(m_ctx, m_vol.Raw().data.data()), m_ctx); }

    void ComputeInterfaceValues() { swe_detail::check(swe_compute_interface_values(m_ctx), m_ctx); }
    void ComputeFluxes() { swe_detail::check(swe_compute_fluxes(m_ctx, m_fluxer.flux, m_fluxer.wavespeed), m_ctx); }

    Storage<3> GetFluxes() const { Storage<3> f((size_t)m_b.Mesh().NumEdges()); swe_detail::check(swe_get_fluxes(m_ctx, f.data.data()), m_ctx); return f; }
    Storage<3> GetEdgField() const { Storage<3> f(2 * (size_t)m_b.Mesh().NumEdges()); swe_detail::check(swe_get_edge_states(m_ctx, f.data.data()), m_ctx); return f; }
    Storage<3> GetSrcField() const { Storage<3> f(2 * (size_t)m_b.Mesh().NumEdges()); swe_detail::check(swe_get_sources(m_ctx, f.data.data()), m_ctx); return f; }
    void EnableTaps(bool on = true) { swe_detail::check(swe_enable_taps(m_ctx, on ? 1 : 0), m_ctx); }
    double GetTau() const noexcept { return m_tau; }
    double GetCor() const noexcept { return m_cor; }
    double GetMinLenToWavespeed() const { double v; swe_detail::check(swe_get_min_len_to_wavespeed(m_ctx, &v), m_ctx); return v; }
    const Fluxer &GetFluxer() const noexcept { return m_fluxer; }
    swe_ctx *Context() const noexcept { return m_ctx; }

 private:
    Fluxer m_fluxer;
    const Domain &m_b;
    double m_cor, m_tau;
    mutable VolumeField m_vol;
    swe_ctx *m_ctx = nullptr;
};
