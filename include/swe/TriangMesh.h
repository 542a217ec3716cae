// TriangMesh.h — Topology / TriangMesh / StructTriangMesh of the reference
// (upstream include/TriangMesh.h:14-65, include/StructTriangMesh.h:4-15, examples/Main.cpp:174).
// Unlike upstream's Topology (const& members, dangling in its pybind binding) the mesh OWNS its
// arrays (inside the swe_hostmesh handle); accessors return small value types.
#pragma once
#include <memory>
#include <string>

#include "Exceptions.h"
#include "Includes.h"

using NodeTag = Idx;
using EdgeTag = std::array<Idx, 2>;
using TriangTag = std::array<Idx, 3>;
using Point = Array<3>;  // (x, y, b): upstream include/PointOperations.h:5

struct Topology {
    explicit Topology(const swe_mesh &v) : m_v(v) {}
    NodeTag NumNodes() const { return m_v.nn; }
    NodeTag NumEdges() const { return m_v.ne; }
    NodeTag NumTriangles() const { return m_v.nt; }
    EdgeTag EdgePoints(NodeTag i) const { return {m_v.edge_nodes[2 * i], m_v.edge_nodes[2 * i + 1]}; }
    EdgeTag EdgeTriangs(NodeTag i) const { return {m_v.edge_elements[2 * i], m_v.edge_elements[2 * i + 1]}; }
    bool IsEdgeBoundary(NodeTag i) const { return m_v.edge_elements[2 * i + 1] < 0; }
    TriangTag TriangPoints(NodeTag i) const { return {m_v.element_nodes[3 * i], m_v.element_nodes[3 * i + 1], m_v.element_nodes[3 * i + 2]}; }
    TriangTag TriangEdges(NodeTag i) const { return {m_v.element_edges[3 * i], m_v.element_edges[3 * i + 1], m_v.element_edges[3 * i + 2]}; }
    TriangTag TriangTriangs(NodeTag i) const { return {m_v.element_neighbours[3 * i], m_v.element_neighbours[3 * i + 1], m_v.element_neighbours[3 * i + 2]}; }
    bool IsTriangleBoundary(NodeTag i) const {
        for (Idx e : TriangEdges(i)) if (IsEdgeBoundary(e)) return true;
        return false;
    }
    const swe_mesh &View() const { return m_v; }

 private:
    swe_mesh m_v;
};

class TriangMesh {
 public:
    // Gmsh >= 4.1 ASCII file, numbering as in upstream's notebooks/topology.dat
    explicit TriangMesh(const std::string &filename) {
        swe_hostmesh *h = nullptr;
        swe_detail::check(swe_hostmesh_gmsh(&h, filename.c_str()));
        reset(h);
    }
    NodeTag NumNodes() const { return m_view.nn; }
    NodeTag NumEdges() const { return m_view.ne; }
    NodeTag NumTriangles() const { return m_view.nt; }
    Topology GetTopology() const { return Topology(m_view); }
    EdgeTag EdgePoints(NodeTag i) const { return GetTopology().EdgePoints(i); }
    EdgeTag EdgeTriangs(NodeTag i) const { return GetTopology().EdgeTriangs(i); }
    TriangTag TriangPoints(NodeTag i) const { return GetTopology().TriangPoints(i); }
    TriangTag TriangEdges(NodeTag i) const { return GetTopology().TriangEdges(i); }
    TriangTag TriangTriangs(NodeTag i) const { return GetTopology().TriangTriangs(i); }
    Point P(NodeTag i) const { const double *g = m_view.geometry + 3 * i; return {g[0], g[1], g[2]}; }
    Point T(NodeTag t) const {  // centroid in all three coordinates (upstream src/Bathymetry.cpp:24-27)
        const TriangTag tp = TriangPoints(t);
        const double third = 1. / 3.;
        Point r;
        for (int c = 0; c < 3; ++c) r[c] = P(tp[0])[c] * third + P(tp[1])[c] * third + P(tp[2])[c] * third;
        return r;
    }
    TriangMesh Refine() const {  // uniform 1 -> 4 split
        swe_hostmesh *h = nullptr;
        swe_detail::check(swe_hostmesh_refine(&h, m_h.get()));
        return TriangMesh(h);
    }
    swe_hostmesh *Handle() const { return m_h.get(); }
    double *Geometry() { return swe_hostmesh_geometry(m_h.get()); }
    const swe_mesh &View() const { return m_view; }

 protected:
    TriangMesh() = default;
    explicit TriangMesh(swe_hostmesh *h) { reset(h); }
    void reset(swe_hostmesh *h) {
        m_h = std::shared_ptr<swe_hostmesh>(h, swe_hostmesh_free);
        swe_detail::check(swe_hostmesh_view(h, &m_view));
    }
    std::shared_ptr<swe_hostmesh> m_h;
    swe_mesh m_view{};
};

// StructTriangMesh(ni, nj, h): [0, ni h] x [0, nj h], every square split into Bottom/Right/Top/Left
// triangles around its centre node (upstream include/StructTriangMesh.h:4-15).
struct StructTriangMesh : public TriangMesh {
    StructTriangMesh(size_t ni, size_t nj, double h, Idx i0 = 0, Idx j0 = 0) : m_ni(ni), m_nj(nj) {
        swe_hostmesh *hm = nullptr;
        swe_detail::check(swe_hostmesh_struct(&hm, (Idx)ni, (Idx)nj, h, i0, j0));
        reset(hm);
    }
    size_t Ni() const { return m_ni; }
    size_t Nj() const { return m_nj; }

 private:
    size_t m_ni, m_nj;
};
