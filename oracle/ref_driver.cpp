// ref_driver.cpp — C interface over the UPSTREAM solver sources. TEST INFRASTRUCTURE ONLY.
//
// This file is compiled together with the reference's own, unmodified src/*.cpp (taken where
// they lie under /root/reference, plus the named repairs in oracle/ref_patches/) against the
// Eigen subset shim oracle/eigen_shim, into oracle/_ref/libswe_ref_*.so (oracle/Makefile.ref).
// Everything numerical below is executed by upstream code: Topology, Domain, VolumeField,
// SpaceDisc, TimeDisc, Solvers::*, Fluxes::HLL/HLLC<Wavespeeds::*>, Gradient, Bisection,
// TriangAverage. This file only marshals plain arrays in and out (same entry-point shapes as
// oracle/swe_oracle.cpp, prefix ref_ instead of oracle_) and reaches protected members through
// derived classes. It is used by tests/test_ref_anchor.py to pin oracle/swe_oracle.cpp.
#include <CubicPolyMath.h>
#include <Fluxes.h>
#include <Solvers.h>

#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

namespace {

// protected members of upstream classes, reached the legal way (through a derived class)
struct FieldAccess : VolumeField {
    static Storage<3> &str(BaseValueField &f) { return f.*(&FieldAccess::m_str); }
    static const Storage<3> &str(const BaseValueField &f) { return f.*(&FieldAccess::m_str); }
};
struct SD : SpaceDisc {
    using SpaceDisc::SpaceDisc;
    Storage<1> &MaxWp() { return m_max_wp; }
    void Update(const MUSCL &m) { UpdateInterfaceValues(m); }
    double &MinLen() { return m_min_length_to_wavespeed; }
};
struct TD : TimeDisc {
    using TimeDisc::TimeDisc;
    double Drain(Idx i) const { return ComputeDrainingDt(i); }
};

SpaceDisc::Fluxer pick_flux(int kind, int ws) {
    if (kind == 0) {
        if (ws == 0) return Fluxes::HLL<Wavespeeds::Rusanov>;
        if (ws == 1) return Fluxes::HLL<Wavespeeds::Davis>;
        return Fluxes::HLL<Wavespeeds::Einfeldt>;
    }
    if (ws == 0) return Fluxes::HLLC<Wavespeeds::Rusanov>;
    if (ws == 1) return Fluxes::HLLC<Wavespeeds::Davis>;
    return Fluxes::HLLC<Wavespeeds::Einfeldt>;
}

struct Ctx {
    Idx nn, ne, nt;
    EdgeTagArray ep, et;
    TriangTagArray tp, te, tt;
    PointArray geom;
    std::unique_ptr<Eigen::Ref<PointArray>> gref;
    std::unique_ptr<Topology> topo;
    std::unique_ptr<Domain> dom;
    std::unique_ptr<VolumeField> v0;
    std::unique_ptr<SD> sd;
    std::unique_ptr<TD> td;
    double cor = 0, tau = 0;
    int flux = -1, ws = -1;
    std::string err;

    // SpaceDisc takes its Fluxer at construction (include/SpaceDisc.h:24): rebuild it, keeping
    // the state and the edge fields, when another flux is asked for.
    void ensure(int kind, int w) {
        if (sd && kind == flux && w == ws) return;
        std::unique_ptr<SD> neu(new SD(pick_flux(kind, w), *dom, sd ? sd->GetVolField() : *v0, cor, tau));
        if (sd) {
            FieldAccess::str(const_cast<EdgeField &>(neu->GetEdgField())) = FieldAccess::str(sd->GetEdgField());
            FieldAccess::str(const_cast<EdgeField &>(neu->GetSrcField())) = FieldAccess::str(sd->GetSrcField());
            const_cast<Storage<3> &>(neu->GetFluxes()) = sd->GetFluxes();
            neu->MaxWp() = sd->MaxWp();
            neu->MinLen() = sd->MinLen();
        } else {
            // fresh fields: define every coefficient (upstream leaves them uninitialised)
            FieldAccess::str(const_cast<EdgeField &>(neu->GetEdgField())).setZero();
            FieldAccess::str(const_cast<EdgeField &>(neu->GetSrcField())).setZero();
            const_cast<Storage<3> &>(neu->GetFluxes()).setZero();
            neu->MaxWp().resize(Eigen::NoChange, nn);
            neu->MaxWp().setZero();
            neu->MinLen() = 1.;
        }
        sd = std::move(neu);
        td.reset(new TD(sd.get()));
        flux = kind; ws = w;
    }
};

template <class A> void fill_rows(A &a, const int64_t *src, Idx rows, Idx cols) {
    a.resize(rows, cols);
    for (Idx r = 0; r < rows; ++r)
        for (Idx c = 0; c < cols; ++c) a(r, c) = src[r * cols + c];
}

}  // namespace

extern "C" {

void *ref_create(int64_t nn, int64_t ne, int64_t nt, const double *geom, const int64_t *ep, const int64_t *et,
                 const int64_t *tp, const int64_t *te, const int64_t *tt, double cor, double tau) {
    Ctx *c = new Ctx();
    c->nn = nn; c->ne = ne; c->nt = nt; c->cor = cor; c->tau = tau;
    fill_rows(c->ep, ep, ne, 2); fill_rows(c->et, et, ne, 2);
    fill_rows(c->tp, tp, nt, 3); fill_rows(c->te, te, nt, 3); fill_rows(c->tt, tt, nt, 3);
    c->geom.resize(Eigen::NoChange, nn);
    for (Idx n = 0; n < nn; ++n)
        for (int k = 0; k < 3; ++k) c->geom(k, n) = geom[3 * n + k];
    c->gref.reset(new Eigen::Ref<PointArray>(c->geom));
    c->topo.reset(new Topology(nn, c->ep, c->et, c->tp, c->te, c->tt));
    c->dom.reset(new Domain(*c->gref, *c->topo));
    c->v0.reset(new VolumeField(*c->dom, size_t(nt)));
    FieldAccess::str(*c->v0).setZero();
    c->ensure(1, 2);
    return c;
}
void ref_destroy(void *p) { delete static_cast<Ctx *>(p); }
const char *ref_last_error(void *p) { return static_cast<Ctx *>(p)->err.c_str(); }

// raw state I/O (no assigner in between, like oracle_set_state)
void ref_set_state(void *p, const double *prim) {
    Ctx *c = static_cast<Ctx *>(p);
    Storage<3> &s = FieldAccess::str(c->sd->GetVolField());
    for (Idx i = 0; i < c->nt; ++i)
        for (int k = 0; k < 3; ++k) s(k, i) = prim[3 * i + k];
}
void ref_get_state(void *p, double *prim) {
    Ctx *c = static_cast<Ctx *>(p);
    const Storage<3> &s = FieldAccess::str(c->sd->GetVolField());
    for (Idx i = 0; i < c->nt; ++i)
        for (int k = 0; k < 3; ++k) prim[3 * i + k] = s(k, i);
}
// state through upstream's PrimAssigner (examples/Main.cpp: v0.prim(i) = ...)
void ref_assign_prim(void *p, int64_t i, const double *prim3) {
    Ctx *c = static_cast<Ctx *>(p);
    c->sd->GetVolField().prim(i) = Array<3>{prim3[0], prim3[1], prim3[2]};
}
void ref_assign_cons(void *p, int64_t i, const double *cons3) {
    Ctx *c = static_cast<Ctx *>(p);
    c->sd->GetVolField().cons(i) = Array<3>{cons3[0], cons3[1], cons3[2]};
}
void ref_get_cons(void *p, int64_t i, double *cons3) {
    Ctx *c = static_cast<Ctx *>(p);
    Array<3> u = std::as_const(c->sd->GetVolField()).cons(i);
    for (int k = 0; k < 3; ++k) cons3[k] = u[k];
}

// Solvers::X, exactly upstream's loops (sequential, in place: S7/S8 "sequential" semantics)
int ref_step(void *p, int scheme, int flux, int ws, double dt) {
    Ctx *c = static_cast<Ctx *>(p);
    try {
        c->ensure(flux, ws);
        if (scheme == 0) Solvers::Euler(c->td.get(), dt);
        else if (scheme == 1) Solvers::SSPRK2(c->td.get(), dt);
        else Solvers::SSPRK3(c->td.get(), dt);
    } catch (const std::exception &e) { c->err = e.what(); return -1; }
    return 0;
}
int ref_compute_interface_values(void *p) {
    Ctx *c = static_cast<Ctx *>(p);
    try { c->sd->ComputeInterfaceValues(); } catch (const std::exception &e) { c->err = e.what(); return -1; }
    return 0;
}
int ref_compute_fluxes(void *p, int flux, int ws) {
    Ctx *c = static_cast<Ctx *>(p);
    try { c->ensure(flux, ws); c->sd->ComputeFluxes(); } catch (const std::exception &e) { c->err = e.what(); return -1; }
    return 0;
}
// One stage with snapshot semantics built from upstream's own pieces: every TimeDisc::RHS(i, dts)
// is evaluated on the pre-stage state first, then cons(i) = a0*U0.cons(i) + a1*cons(i) + RHS.
// (The loop ORDER is this driver's; each RHS, drain and assigner call is upstream code.)
int ref_stage_update_snapshot(void *p, const double *U0prim, double a0, double a1, double dts, int plain_sum) {
    Ctx *c = static_cast<Ctx *>(p);
    try {
        std::vector<Array<3>> rhs(size_t(c->nt));
        for (Idx i = 0; i < c->nt; ++i) rhs[size_t(i)] = c->td->RHS(i, dts);
        VolumeField U0(*c->dom, c->sd->GetVolField());
        if (U0prim) {
            Storage<3> &s = FieldAccess::str(U0);
            for (Idx i = 0; i < c->nt; ++i)
                for (int k = 0; k < 3; ++k) s(k, i) = U0prim[3 * i + k];
        }
        VolumeField &V = c->sd->GetVolField();
        for (Idx i = 0; i < c->nt; ++i) {
            if (plain_sum) V.cons(i) = std::as_const(U0).cons(i) + rhs[size_t(i)];
            else V.cons(i) = a0 * std::as_const(U0).cons(i) + a1 * std::as_const(V).cons(i) + rhs[size_t(i)];
        }
    } catch (const std::exception &e) { c->err = e.what(); return -1; }
    return 0;
}
double ref_min_len_to_wavespeed(void *p) { return static_cast<Ctx *>(p)->sd->GetMinLenToWavespeed(); }
double ref_cfl_dt(void *p) { return static_cast<Ctx *>(p)->td->CFLdt(); }

void ref_get_edge_states(void *p, double *out) {
    Ctx *c = static_cast<Ctx *>(p);
    const Storage<3> &s = FieldAccess::str(c->sd->GetEdgField());
    for (Idx j = 0; j < 2 * c->ne; ++j) for (int k = 0; k < 3; ++k) out[3 * j + k] = s(k, j);
}
void ref_get_sources(void *p, double *out) {
    Ctx *c = static_cast<Ctx *>(p);
    const Storage<3> &s = FieldAccess::str(c->sd->GetSrcField());
    for (Idx j = 0; j < 2 * c->ne; ++j) for (int k = 0; k < 3; ++k) out[3 * j + k] = s(k, j);
}
void ref_get_fluxes(void *p, double *out) {
    Ctx *c = static_cast<Ctx *>(p);
    const Storage<3> &s = c->sd->GetFluxes();
    for (Idx j = 0; j < c->ne; ++j) for (int k = 0; k < 3; ++k) out[3 * j + k] = s(k, j);
}
void ref_get_node_max_w(void *p, double *out) {
    Ctx *c = static_cast<Ctx *>(p);
    for (Idx n = 0; n < c->nn; ++n) out[n] = c->sd->MaxWp()[n];
}
void ref_set_node_max_w(void *p, const double *in) {
    Ctx *c = static_cast<Ctx *>(p);
    c->sd->MaxWp().resize(Eigen::NoChange, c->nn);
    for (Idx n = 0; n < c->nn; ++n) c->sd->MaxWp()[n] = in[n];
}
void ref_get_draining_dt(void *p, double *out) {
    Ctx *c = static_cast<Ctx *>(p);
    for (Idx i = 0; i < c->nt; ++i) out[i] = c->td->Drain(i);
}
void ref_rhs(void *p, int64_t i, double dt, double *out3) {
    Ctx *c = static_cast<Ctx *>(p);
    Array<3> r = c->td->RHS(i, dt);
    for (int k = 0; k < 3; ++k) out3[k] = r[k];
}
void ref_get_cell_class(void *p, int8_t *out) {
    Ctx *c = static_cast<Ctx *>(p);
    for (Idx i = 0; i < c->nt; ++i) out[i] = c->sd->IsDryCell(i) ? 0 : (c->sd->IsFullWetCell(i) ? 2 : 1);
}
int ref_is_part_wet(void *p, int64_t i) { return static_cast<Ctx *>(p)->sd->IsPartWetCell(i) ? 1 : 0; }

// Domain taps: T(t), E(e), L(e), Area(t), Norm(e, EdgeTriangs(e)[0]), TriangSlope(t)
void ref_get_geometry(void *p, double *T3, double *E3, double *L, double *A, double *n0, double *slope) {
    Ctx *c = static_cast<Ctx *>(p);
    const Domain &d = *c->dom;
    for (Idx t = 0; t < c->nt; ++t) {
        if (T3) { Point q = d.T(t); for (int k = 0; k < 3; ++k) T3[3 * t + k] = q[k]; }
        if (A) A[t] = d.Area(t);
        if (slope) { Eigen::Vector2d s = d.TriangSlope(t); slope[2 * t] = s[0]; slope[2 * t + 1] = s[1]; }
    }
    for (Idx e = 0; e < c->ne; ++e) {
        if (E3) { Point q = d.E(e); for (int k = 0; k < 3; ++k) E3[3 * e + k] = q[k]; }
        if (L) L[e] = d.L(e);
        if (n0) { Eigen::Vector2d n = d.Norm(e, c->topo->EdgeTriangs(e)[0]); n0[2 * e] = n[0]; n0[2 * e + 1] = n[1]; }
    }
}
void ref_norm(void *p, int64_t e, int64_t t, double *n2) {
    Eigen::Vector2d n = static_cast<Ctx *>(p)->dom->Norm(e, t);
    n2[0] = n[0]; n2[1] = n[1];
}
void ref_tang(void *p, int64_t e, int64_t t, double *t2) {
    Eigen::Vector2d n = static_cast<Ctx *>(p)->dom->Tang(e, t);
    t2[0] = n[0]; t2[1] = n[1];
}

// unit-level entry points
double ref_bisection_cubic(double d, double cc, double b, double lo, double hi) { return Bisection(CubicPoly(d, cc, b), lo, hi); }
void ref_gradient(const double *P9, double *g2) {  // three points (x,y,z) = columns of the 3x3
    Eigen::Matrix3d m;
    for (int k = 0; k < 3; ++k) for (int r = 0; r < 3; ++r) m(r, k) = P9[3 * k + r];
    Eigen::Vector2d g = Gradient(m);
    g2[0] = g[0]; g2[1] = g[1];
}
void ref_elem_flux(const double *n2, const double *U3, double *F3) {
    Array<3> f = ElemFlux(Eigen::Vector2d(n2[0], n2[1]), Array<3>{U3[0], U3[1], U3[2]});
    for (int k = 0; k < 3; ++k) F3[k] = f[k];
}
void ref_wavespeeds(int ws, double ul, double hl, double ur, double hr, double *a2) {
    Array<2> a = ws == 0 ? Wavespeeds::Rusanov(ul, hl, ur, hr) : ws == 1 ? Wavespeeds::Davis(ul, hl, ur, hr)
                                                                         : Wavespeeds::Einfeldt(ul, hl, ur, hr);
    a2[0] = a[0]; a2[1] = a[1];
}
double ref_len(const double *a3, const double *b3) { return Len(Point{a3[0], a3[1], a3[2]}, Point{b3[0], b3[1], b3[2]}); }
double ref_det(const double *a3, const double *b3) { return Det(Point{a3[0], a3[1], a3[2]}, Point{b3[0], b3[1], b3[2]}); }
double ref_triang_area(const double *a3, const double *b3, const double *c3) {
    return TriangArea(Point{a3[0], a3[1], a3[2]}, Point{b3[0], b3[1], b3[2]}, Point{c3[0], c3[1], c3[2]});
}
// reconstruction of one cell: kind 0 dry, 1 partwet1, 2 fullwet, 3 partwet2 -> origin(3), G(3x2)
int ref_reconstruct(void *p, int kind, int64_t i, double *o3, double *G6) {
    Ctx *c = static_cast<Ctx *>(p);
    try {
        const SD &sd = *c->sd;
        MUSCLObject::MUSCL m = kind == 0 ? sd.ReconstructDryCell(i) : kind == 1 ? sd.ReconstructPartWetCell1(i)
                               : kind == 2 ? sd.ReconstructFullWetCell(i) : sd.ReconstructPartWetCell2(i);
        Array<3> o = m.AtOrigin();
        // m_grad is private: MUSCL::Gradient(p) returns it unchanged wherever the depth at p is
        // non-negative (include/MUSCLObject.h:41-47), so ask at a probe far below any surface.
        Point T = c->dom->T(i);
        Point probe{T[0], T[1], -1e300};
        Eigen::Matrix32d G = m.Gradient(probe);
        for (int k = 0; k < 3; ++k) { o3[k] = o[k]; G6[2 * k] = G(k, 0); G6[2 * k + 1] = G(k, 1); }
    } catch (const std::exception &e) { c->err = e.what(); return -1; }
    return 0;
}
// MUSCL{o,G,i}.AtPoint(pt) and .Gradient(pt).row(0)
void ref_muscl_at_point(void *p, int64_t i, const double *o3, const double *G6, const double *pt3, double *out3, double *g2) {
    Ctx *c = static_cast<Ctx *>(p);
    Eigen::Matrix32d G;
    for (int k = 0; k < 3; ++k) { G(k, 0) = G6[2 * k]; G(k, 1) = G6[2 * k + 1]; }
    MUSCLObject::MUSCL m(*c->dom, Array<3>{o3[0], o3[1], o3[2]}, G, i);
    Point pt{pt3[0], pt3[1], pt3[2]};
    Array<3> a = m.AtPoint(pt);
    for (int k = 0; k < 3; ++k) out3[k] = a[k];
    if (g2) { Eigen::Matrix32d g = m.Gradient(pt); g2[0] = g(0, 0); g2[1] = g(0, 1); }
}
// flux of one edge through the Fluxer (reads the edge fields as they are)
int ref_edge_flux(void *p, int flux, int ws, int64_t e, double *F3, double *r) {
    Ctx *c = static_cast<Ctx *>(p);
    try {
        c->ensure(flux, ws);
        auto et = c->topo->EdgeTriangs(e);
        Array<3> f = pick_flux(flux, ws)(c->sd.get(), e, et[0], et[1], r);
        for (int k = 0; k < 3; ++k) F3[k] = f[k];
    } catch (const std::exception &e2) { c->err = e2.what(); return -1; }
    return 0;
}

// TriangAverage<3,n> (include/PointOperations.h:20-44) of a caller-supplied integrand
typedef void (*ref_fn3)(const double *pt3, double *out3, void *user);
int ref_triang_average3(int n, const double *p0, const double *p1, const double *p2, ref_fn3 fn, void *user, double *out3) {
    auto f = [&](const Point &q) {
        double in[3] = {q[0], q[1], q[2]}, o[3];
        fn(in, o, user);
        return Array<3>{o[0], o[1], o[2]};
    };
    Point a{p0[0], p0[1], p0[2]}, b{p1[0], p1[1], p1[2]}, cc{p2[0], p2[1], p2[2]};
    Array<3> r;
    switch (n) {
        case 1: r = TriangAverage<3, 1>(a, b, cc, f); break;
        case 2: r = TriangAverage<3, 2>(a, b, cc, f); break;
        case 3: r = TriangAverage<3, 3>(a, b, cc, f); break;
        case 4: r = TriangAverage<3, 4>(a, b, cc, f); break;
        case 8: r = TriangAverage<3, 8>(a, b, cc, f); break;
        case 10: r = TriangAverage<3, 10>(a, b, cc, f); break;
        case 100: r = TriangAverage<3, 100>(a, b, cc, f); break;
        default: return -1;
    }
    for (int k = 0; k < 3; ++k) out3[k] = r[k];
    return 0;
}

// which repairs this build contains (set by the Makefile)
#ifndef SWE_REF_PATCHES
#define SWE_REF_PATCHES "?"
#endif
const char *ref_patches(void) { return SWE_REF_PATCHES; }

}  // extern "C"
