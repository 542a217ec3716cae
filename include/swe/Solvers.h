// Solvers.h — upstream include/Solvers.h:6-8, src/Solvers.cpp: one explicit time step of size dt.
// Same signatures; the bodies forward to the device (swe_step), or to every GPU of a multi-GPU SpaceDisc
// (swe_dist_group_run). Solvers::Run keeps nsteps steps on the device(s) without host synchronisation
// (dt <= 0: every step uses CFLdt() of the previous one, the global minimum over all GPUs).
#pragma once
#include "TimeDisc.h"

namespace Solvers {
inline void Run(TimeDisc *const td, swe_scheme scheme, Idx nsteps, double dt, double dt0 = 0.) {
    SpaceDisc *sd = td->GetSpaceDisc();
    const Fluxer &f = sd->GetFluxer();
    if (sd->IsDistributed()) {
        auto &r = sd->Ranks();
        int rc = swe_dist_group_run(const_cast<swe_dist **>(r.data()), (int32_t)r.size(), scheme, f.flux, f.wavespeed, nsteps, dt, dt0);
        for (size_t k = 0; k < r.size() && rc != SWE_OK; ++k) SpaceDisc::dist_check(rc, r[k]);
        for (auto *d : r) SpaceDisc::dist_check(swe_dist_synchronize(d), d);
    } else {
        swe_detail::check(swe_run(sd->Context(), scheme, f.flux, f.wavespeed, nsteps, dt, dt0), sd->Context());
        swe_detail::check(swe_synchronize(sd->Context()), sd->Context());
    }
    sd->Touch();
}
inline void Step(TimeDisc *const td, swe_scheme scheme, double dt) {
    SpaceDisc *sd = td->GetSpaceDisc();
    if (!(dt > 0.)) throw DomainError("Solvers: dt must be positive");
    if (sd->IsDistributed()) { Run(td, scheme, 1, dt); return; }
    swe_detail::check(swe_step(sd->Context(), scheme, sd->GetFluxer().flux, sd->GetFluxer().wavespeed, dt), sd->Context());
    sd->Touch();
}
inline void Euler(TimeDisc *const td, double dt) { Step(td, SWE_EULER, dt); }
inline void SSPRK2(TimeDisc *const td, double dt) { Step(td, SWE_SSPRK2, dt); }
inline void SSPRK3(TimeDisc *const td, double dt) { Step(td, SWE_SSPRK3, dt); }
}  // namespace Solvers
