// Bathymetry.h — Domain = mesh + nodal bathymetry (upstream include/Bathymetry.h:12-49,
// src/Bathymetry.cpp). Host-side geometry accessors use the same formulas and operation order
// as the device set-up kernels (csrc/swe_kernels.cuh), so both agree bit for bit.
#pragma once
#include <cmath>

#include "TriangMesh.h"

constexpr inline bool IsWet(double h) noexcept { return h > SWE_WET_DEPTH; }  // upstream include/Bathymetry.h:5-8
inline Point DryState(double b) noexcept { return {b, 0., 0.}; }

struct Domain {
    explicit Domain(TriangMesh *mesh) : m_mesh(mesh) {}
    explicit Domain(TriangMesh mesh) : m_owned(std::make_shared<TriangMesh>(std::move(mesh))), m_mesh(m_owned.get()) {}
    const TriangMesh &Mesh() const { return *m_mesh; }
    Topology GetTopology() const { return m_mesh->GetTopology(); }
    size_t Size() const { return (size_t)m_mesh->NumNodes(); }
    double &AtNode(NodeTag i) { return m_mesh->Geometry()[3 * i + 2]; }  // nodal bed elevation
    double AtNode(NodeTag i) const { return m_mesh->P(i)[2]; }
    Point P(NodeTag p) const { return m_mesh->P(p); }
    Point T(NodeTag t) const { return m_mesh->T(t); }
    Point E(NodeTag e) const {  // edge midpoint in x, y and b (decision S1: upstream's body is a stub)
        const EdgeTag ep = m_mesh->EdgePoints(e);
        Point r;
        for (int c = 0; c < 3; ++c) r[c] = 0.5 * (P(ep[0])[c] + P(ep[1])[c]);
        return r;
    }
    double L(NodeTag e) const {
        const EdgeTag ep = m_mesh->EdgePoints(e);
        const Point a = P(ep[0]), b = P(ep[1]);
        return std::sqrt((a[0] - b[0]) * (a[0] - b[0]) + (a[1] - b[1]) * (a[1] - b[1]));
    }
    double Area(NodeTag t) const {
        const TriangTag tp = m_mesh->TriangPoints(t);
        const Point a = P(tp[0]), b = P(tp[1]), c = P(tp[2]);
        return 0.5 * std::fabs((b[0] - a[0]) * (c[1] - a[1]) - (c[0] - a[0]) * (b[1] - a[1]));
    }
    std::array<double, 2> Tang(NodeTag e, NodeTag t) const {  // upstream src/Bathymetry.cpp:69-76
        const EdgeTag ep = m_mesh->EdgePoints(e);
        const Point a = P(ep[0]), b = P(ep[1]), c = T(t);
        const double len = L(e);
        double tx = (b[0] - a[0]) / len, ty = (b[1] - a[1]) / len;
        if ((c[0] - a[0]) * ty - tx * (c[1] - a[1]) > 0.) { tx = -tx; ty = -ty; }
        return {tx, ty};
    }
    std::array<double, 2> Norm(NodeTag e, NodeTag t) const { const auto tg = Tang(e, t); return {tg[1], -tg[0]}; }

 private:
    std::shared_ptr<TriangMesh> m_owned;
    TriangMesh *m_mesh;
};
