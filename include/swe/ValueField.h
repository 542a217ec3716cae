// ValueField.h — host-side VolumeField with the reference's proxy assigners
// (upstream include/ValueField.h:17-62, include/Assigners.h, src/Assigners.cpp). It is the
// container a driver fills with the initial state before handing it to SpaceDisc; on the device
// the field lives as structure-of-arrays inside the swe_ctx.
#pragma once
#include <cmath>

#include "Bathymetry.h"

struct PrimAssigner {  // upstream src/Assigners.cpp:4-20
    Storage<3> *f; double b; Idx col;
    Array<3> Get() const { return f->col(col); }
    void operator=(const Array<3> &rhs) {
        const double h = rhs[0] - b;
        if (!IsWet(h)) { f->set_col(col, DryState(b)); return; }
        Array<3> v = rhs;
        if (h < SWE_DAMP_DEPTH) { const double s = std::sqrt(2) * h / std::sqrt(h * h + SWE_DAMP_EPS_PRIM); v[1] *= s; v[2] *= s; }
        f->set_col(col, v);
    }
};
struct ConsAssigner {  // upstream src/Assigners.cpp:22-44
    Storage<3> *f; double b; Idx col;
    Array<3> Get() const { Array<3> r = f->col(col); r[0] -= b; r[1] *= r[0]; r[2] *= r[0]; return r; }
    void operator=(const Array<3> &rhs) {
        const double h = rhs[0];
        if (!IsWet(h)) { f->set_col(col, DryState(b)); return; }
        const double ih = (h < SWE_DAMP_DEPTH) ? std::sqrt(2) * h / std::sqrt(h * h * h * h + SWE_DAMP_EPS_CONS) : 1. / h;
        f->set_col(col, {h + b, rhs[1] * ih, rhs[2] * ih});
    }
    void operator+=(const Array<3> &rhs) { *this = Get() + rhs; }
};

struct VolumeField {
    VolumeField(const Domain &b, size_t size) : m_b(&b), m_str(size) {}
    size_t Size() const { return m_str.cols(); }
    double b(Idx t) const { return m_b->T(t)[2]; }  // VolumeDomainWrapper::At (upstream src/ValueField.cpp:8-10)
    Array<3> prim(Idx t) const { return m_str.col(t); }
    double w(Idx t) const { return m_str(0, t); }
    double u(Idx t) const { return m_str(1, t); }
    double v(Idx t) const { return m_str(2, t); }
    double h(Idx t) const { return m_str(0, t) - b(t); }
    double hu(Idx t) const { return h(t) * u(t); }
    double hv(Idx t) const { return h(t) * v(t); }
    Array<3> cons(Idx t) const { return {h(t), hu(t), hv(t)}; }
    PrimAssigner prim(Idx t) { return PrimAssigner{&m_str, b(t), t}; }
    ConsAssigner cons(Idx t) { return ConsAssigner{&m_str, b(t), t}; }
    Storage<3> &Raw() { return m_str; }
    const Storage<3> &Raw() const { return m_str; }

 private:
    const Domain *m_b;
    Storage<3> m_str;
};
