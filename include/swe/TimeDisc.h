// TimeDisc.h — upstream include/TimeDisc.h:4-23. RHS / ComputeDrainingDt run inside the device
// stage-update kernel; CFLdt reads the reduced min_len_to_wavespeed back.
#pragma once
#include "SpaceDisc.h"

struct TimeDisc {
    explicit TimeDisc(SpaceDisc *sd = nullptr) : m_sd(sd) {}
    SpaceDisc *GetSpaceDisc() { return m_sd; }
    const SpaceDisc *GetSpaceDisc() const { return m_sd; }
    void SetSpaceDisc(SpaceDisc *sd) { m_sd = sd; }
    double CFLdt() const { double dt; swe_detail::check(swe_cfl_dt(m_sd->Context(), &dt), m_sd->Context()); return dt; }
    std::vector<double> DrainingDt() const {  // per cell, of the last stage (ComputeDrainingDt, src/TimeDisc.cpp:43-66)
        std::vector<double> d((size_t)m_sd->GetDomain().Mesh().NumTriangles());
        swe_detail::check(swe_get_draining_dt(m_sd->Context(), d.data()), m_sd->Context());
        return d;
    }

 protected:
    SpaceDisc *m_sd;
};
