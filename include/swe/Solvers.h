// Solvers.h — upstream include/Solvers.h:6-8, src/Solvers.cpp: one explicit time step of size dt.
// Same signatures; the bodies forward to the device (swe_step). Solvers::Run keeps nsteps steps on
// the device without host synchronisation (dt <= 0: every step uses CFLdt() of the previous one).
#pragma once
#include "TimeDisc.h"

namespace Solvers {
inline void Step(TimeDisc *const td, swe_scheme scheme, double dt) {
    SpaceDisc *sd = td->GetSpaceDisc();
    swe_detail::check(swe_step(sd->Context(), scheme, sd->GetFluxer().flux, sd->GetFluxer().wavespeed, dt), sd->Context());
}
inline void Euler(TimeDisc *const td, double dt) { Step(td, SWE_EULER, dt); }
inline void SSPRK2(TimeDisc *const td, double dt) { Step(td, SWE_SSPRK2, dt); }
inline void SSPRK3(TimeDisc *const td, double dt) { Step(td, SWE_SSPRK3, dt); }
inline void Run(TimeDisc *const td, swe_scheme scheme, Idx nsteps, double dt, double dt0 = 0.) {
    SpaceDisc *sd = td->GetSpaceDisc();
    swe_detail::check(swe_run(sd->Context(), scheme, sd->GetFluxer().flux, sd->GetFluxer().wavespeed, nsteps, dt, dt0), sd->Context());
    swe_detail::check(swe_synchronize(sd->Context()), sd->Context());
}
}  // namespace Solvers
