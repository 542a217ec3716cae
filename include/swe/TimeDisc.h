// TimeDisc.h — upstream include/TimeDisc.h:4-23. RHS / ComputeDrainingDt run inside the device stage-update
// kernel during Solvers::X; the per-cell accessors below are whole-array taps of the same kernels (cached until
// the state or the fluxes change), so a time loop written in upstream's own style,
//     sd->ComputeInterfaceValues(); sd->ComputeFluxes();
//     for (i...) sd->GetVolFieldForWrite().cons(i) += td->RHS(i, dt);   sd->Upload();
// works unchanged (with snapshot semantics: every RHS(i) sees the same state, SURVEY decision S7).
#pragma once
#include <limits>

#include "SpaceDisc.h"

struct TimeDisc {
    explicit TimeDisc(SpaceDisc *sd = nullptr) : m_sd(sd) {}
    SpaceDisc *GetSpaceDisc() { return m_sd; }
    const SpaceDisc *GetSpaceDisc() const { return m_sd; }
    void SetSpaceDisc(SpaceDisc *sd) { m_sd = sd; m_rhs_epoch = 0; }
    double CFLdt() const {  // 0.15 * min_len_to_wavespeed (include/TimeDisc.h:13,22), global over all GPUs
        double dt;
        if (m_sd->IsDistributed()) SpaceDisc::dist_check(swe_dist_cfl_dt(m_sd->Ranks()[0], &dt), m_sd->Ranks()[0]);
        else swe_detail::check(swe_cfl_dt(m_sd->Context(), &dt), m_sd->Context());
        return dt;
    }
    // TimeDisc::RHS(i, dt) (include/TimeDisc.h:15, src/TimeDisc.cpp:3-41)
    Array<3> RHS(Idx i, double dt) const {
        refresh(dt);
        return {m_rhs[3 * (size_t)i], m_rhs[3 * (size_t)i + 1], m_rhs[3 * (size_t)i + 2]};
    }
    // TimeDisc::ComputeDrainingDt(i) (src/TimeDisc.cpp:43-66): +inf for ghost ids
    double ComputeDrainingDt(Idx i) const {
        if (i < 0) return std::numeric_limits<double>::infinity();
        refresh(m_rhs_dt);
        return m_drain[(size_t)i];
    }
    std::vector<double> DrainingDt() const {  // per cell, of the last stage
        std::vector<double> d((size_t)m_sd->GetDomain().Mesh().NumTriangles());
        swe_detail::check(swe_get_draining_dt(m_sd->Context(), d.data()), m_sd->Context());
        return d;
    }

 protected:
    void refresh(double dt) const {
        if (m_sd->IsDistributed()) throw DomainError("TimeDisc::RHS(i, dt) is a single-GPU tap; not available on a multi-GPU SpaceDisc");
        if (m_rhs_epoch == m_sd->Epoch() && dt == m_rhs_dt) return;
        const size_t nt = (size_t)m_sd->GetDomain().Mesh().NumTriangles();
        m_rhs.resize(3 * nt); m_drain.resize(nt);
        swe_detail::check(swe_compute_rhs(m_sd->Context(), dt, m_rhs.data()), m_sd->Context());
        swe_detail::check(swe_get_draining_dt(m_sd->Context(), m_drain.data()), m_sd->Context());
        m_rhs_epoch = m_sd->Epoch(); m_rhs_dt = dt;
    }
    SpaceDisc *m_sd;
    mutable std::vector<double> m_rhs, m_drain;
    mutable uint64_t m_rhs_epoch = 0;
    mutable double m_rhs_dt = 0.;
};
