// swe_fvm_module.cpp — pybind11 module `SWE_FVM` over the C++17 host API (include/swe/*.h), i.e. the
// working version of upstream's pybind/Topology.cpp:8-31 (which exposes only Topology and stores
// references to the converted temporaries, so its counters return garbage). Here Topology OWNS
// copies of its arrays, and the mesh classes and the time step (SpaceDisc / TimeDisc / Solvers)
// are exposed as well. The module links against libswe_b200.so (the C-ABI); no kernels live here.
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <memory>
#include <stdexcept>
#include <vector>

#include "swe/Solvers.h"
#include "swe/Tests.h"

namespace py = pybind11;
using namespace pybind11::literals;
using IdxArray = py::array_t<Idx, py::array::c_style | py::array::forcecast>;
using DblArray = py::array_t<double, py::array::c_style | py::array::forcecast>;

namespace {

struct PyTopology {  // owning counterpart of upstream's Topology (include/TriangMesh.h:14-65)
    Idx nn;
    std::vector<Idx> ep, et, tp, te, tt;
    PyTopology(Idx numNodes, const IdxArray &edgeNodes, const IdxArray &edgeElements, const IdxArray &elementNodes,
               const IdxArray &elementEdges, const IdxArray &elementNeighbours)
        : nn(numNodes) {
        auto take = [](const IdxArray &a, py::ssize_t cols, const char *name) {
            if (a.ndim() != 2 || a.shape(1) != cols) throw std::invalid_argument(std::string(name) + ": expected an (N, " + std::to_string(cols) + ") integer array");
            return std::vector<Idx>(a.data(), a.data() + a.size());
        };
        ep = take(edgeNodes, 2, "edgeNodes"); et = take(edgeElements, 2, "edgeElements");
        tp = take(elementNodes, 3, "elementNodes"); te = take(elementEdges, 3, "elementEdges");
        tt = take(elementNeighbours, 3, "elementNeighbours");
        if (ep.size() != et.size() || tp.size() != te.size() || tp.size() != tt.size())
            throw std::invalid_argument("Topology: inconsistent array lengths");
    }
    Idx NumNodes() const { return nn; }
    Idx NumEdges() const { return (Idx)ep.size() / 2; }
    Idx NumTriangles() const { return (Idx)tp.size() / 3; }
    bool IsEdgeBoundary(Idx e) const { return et.at(2 * e + 1) < 0; }
};

template <class T>
py::array_t<T> view2d(const T *p, Idx rows, Idx cols, py::handle owner, bool writable) {
    py::array_t<T> a({(py::ssize_t)rows, (py::ssize_t)cols}, {(py::ssize_t)(cols * sizeof(T)), (py::ssize_t)sizeof(T)}, p, owner);
    if (!writable) py::detail::array_proxy(a.ptr())->flags &= ~py::detail::npy_api::NPY_ARRAY_WRITEABLE_;
    return a;
}

Fluxer make_fluxer(const std::string &flux, const std::string &ws) {
    Fluxer f{};
    if (flux == "hll" || flux == "HLL") f.flux = SWE_HLL;
    else if (flux == "hllc" || flux == "HLLC") f.flux = SWE_HLLC;
    else throw std::invalid_argument("flux must be 'hll' or 'hllc'");
    if (ws == "rusanov" || ws == "Rusanov") f.wavespeed = SWE_RUSANOV;
    else if (ws == "davis" || ws == "Davis") f.wavespeed = SWE_DAVIS;
    else if (ws == "einfeldt" || ws == "Einfeldt") f.wavespeed = SWE_EINFELDT;
    else throw std::invalid_argument("wavespeed must be 'rusanov', 'davis' or 'einfeldt'");
    return f;
}

struct PySpaceDisc {  // keeps the Domain alive next to the SpaceDisc that refers to it
    std::shared_ptr<TriangMesh> mesh;
    std::unique_ptr<Domain> domain;
    std::unique_ptr<SpaceDisc> sd;
    PySpaceDisc(const std::string &flux, const std::string &ws, std::shared_ptr<TriangMesh> m, const DblArray &v0, double cor,
                double tau, int device, bool reorder)
        : mesh(std::move(m)) {
        if (v0.ndim() != 2 || v0.shape(0) != mesh->NumTriangles() || v0.shape(1) != 3)
            throw std::invalid_argument("v0: expected an (num_elements, 3) array of (w, u, v)");
        domain = std::make_unique<Domain>(mesh.get());
        VolumeField vf(*domain, (size_t)mesh->NumTriangles());
        std::copy(v0.data(), v0.data() + v0.size(), vf.Raw().data.begin());
        sd = std::make_unique<SpaceDisc>(make_fluxer(flux, ws), *domain, vf, cor, tau, device, reorder);
    }
    DblArray state() const {
        const VolumeField &v = sd->GetVolField();
        DblArray out({(py::ssize_t)mesh->NumTriangles(), (py::ssize_t)3});
        std::copy(v.Raw().data.begin(), v.Raw().data.end(), out.mutable_data());
        return out;
    }
    void set_state(const DblArray &v) {
        if (v.ndim() != 2 || v.shape(0) != mesh->NumTriangles() || v.shape(1) != 3) throw std::invalid_argument("state: expected (num_elements, 3)");
        std::copy(v.data(), v.data() + v.size(), sd->GetVolFieldForWrite().Raw().data.begin());
        sd->Upload();
    }
    DblArray fluxes() const {
        Storage<3> f = sd->GetFluxes();
        DblArray out({(py::ssize_t)mesh->NumEdges(), (py::ssize_t)3});
        std::copy(f.data.begin(), f.data.end(), out.mutable_data());
        return out;
    }
};

struct PyTimeDisc {
    std::shared_ptr<PySpaceDisc> sd;
    TimeDisc td;
    explicit PyTimeDisc(std::shared_ptr<PySpaceDisc> s) : sd(std::move(s)), td(sd->sd.get()) {}
};

}  // namespace

PYBIND11_MODULE(SWE_FVM, m) {
    m.doc() = "SWE_FVM explicit finite-volume time step on B200 (pybind11 over the C++ host API / C-ABI)";

    py::class_<PyTopology>(m, "Topology")
        .def(py::init<Idx, const IdxArray &, const IdxArray &, const IdxArray &, const IdxArray &, const IdxArray &>(),
             "numNodes"_a, "edgeNodes"_a, "edgeElements"_a, "elementNodes"_a, "elementEdges"_a, "elementNeighbours"_a)
        .def_static("create", [](Idx n, const IdxArray &a, const IdxArray &b, const IdxArray &c, const IdxArray &d, const IdxArray &e) {
                return PyTopology(n, a, b, c, d, e); },
            "numNodes"_a, "edgeNodes"_a, "edgeElements"_a, "elementNodes"_a, "elementEdges"_a, "elementNeighbours"_a)
        .def("num_nodes", &PyTopology::NumNodes)
        .def("num_edges", &PyTopology::NumEdges)
        .def("num_elements", &PyTopology::NumTriangles)
        .def("is_edge_boundary", &PyTopology::IsEdgeBoundary);

    py::class_<TriangMesh, std::shared_ptr<TriangMesh>>(m, "TriangMesh")
        .def(py::init<const std::string &>(), "filename"_a, "Gmsh >= 4.1 ASCII mesh, numbered like the reference's reader")
        .def("num_nodes", &TriangMesh::NumNodes)
        .def("num_edges", &TriangMesh::NumEdges)
        .def("num_elements", &TriangMesh::NumTriangles)
        .def("refine", [](const TriangMesh &t) { return std::make_shared<TriangMesh>(t.Refine()); })
        .def("topology", [](const TriangMesh &t) {
            const swe_mesh &v = t.View();
            auto mk = [](const int64_t *p, Idx r, Idx c) { IdxArray a({(py::ssize_t)r, (py::ssize_t)c}); std::copy(p, p + r * c, a.mutable_data()); return a; };
            return PyTopology(v.nn, mk(v.edge_nodes, v.ne, 2), mk(v.edge_elements, v.ne, 2), mk(v.element_nodes, v.nt, 3),
                              mk(v.element_edges, v.nt, 3), mk(v.element_neighbours, v.nt, 3)); })
        .def_property_readonly("geometry", [](py::object self) {
            TriangMesh &t = self.cast<TriangMesh &>();
            return view2d<double>(t.Geometry(), t.NumNodes(), 3, self, true); }, "(num_nodes, 3) array of (x, y, b); b is writable")
        .def_property_readonly("edge_nodes", [](py::object self) { auto &t = self.cast<TriangMesh &>(); return view2d<int64_t>(t.View().edge_nodes, t.NumEdges(), 2, self, false); })
        .def_property_readonly("edge_elements", [](py::object self) { auto &t = self.cast<TriangMesh &>(); return view2d<int64_t>(t.View().edge_elements, t.NumEdges(), 2, self, false); })
        .def_property_readonly("element_nodes", [](py::object self) { auto &t = self.cast<TriangMesh &>(); return view2d<int64_t>(t.View().element_nodes, t.NumTriangles(), 3, self, false); })
        .def_property_readonly("element_edges", [](py::object self) { auto &t = self.cast<TriangMesh &>(); return view2d<int64_t>(t.View().element_edges, t.NumTriangles(), 3, self, false); })
        .def_property_readonly("element_neighbours", [](py::object self) { auto &t = self.cast<TriangMesh &>(); return view2d<int64_t>(t.View().element_neighbours, t.NumTriangles(), 3, self, false); });

    py::class_<StructTriangMesh, TriangMesh, std::shared_ptr<StructTriangMesh>>(m, "StructTriangMesh")
        .def(py::init<size_t, size_t, double, Idx, Idx>(), "ni"_a, "nj"_a, "h"_a, "i0"_a = 0, "j0"_a = 0)
        .def("ni", &StructTriangMesh::Ni)
        .def("nj", &StructTriangMesh::Nj);

    py::class_<PySpaceDisc, std::shared_ptr<PySpaceDisc>>(m, "SpaceDisc")
        .def(py::init<const std::string &, const std::string &, std::shared_ptr<TriangMesh>, const DblArray &, double, double, int, bool>(),
             "flux"_a, "wavespeed"_a, "mesh"_a, "v0"_a, "cor"_a = 0., "tau"_a = 0., "device"_a = 0, "reorder"_a = true)
        .def("get_vol_field", &PySpaceDisc::state, "(num_elements, 3) primitive state (w, u, v)")
        .def("set_vol_field", &PySpaceDisc::set_state)
        .def("compute_interface_values", [](PySpaceDisc &s) { s.sd->ComputeInterfaceValues(); })
        .def("compute_fluxes", [](PySpaceDisc &s) { s.sd->ComputeFluxes(); })
        .def("get_fluxes", &PySpaceDisc::fluxes)
        .def("get_min_len_to_wavespeed", [](PySpaceDisc &s) { return s.sd->GetMinLenToWavespeed(); })
        .def("get_cor", [](PySpaceDisc &s) { return s.sd->GetCor(); })
        .def("get_tau", [](PySpaceDisc &s) { return s.sd->GetTau(); });

    py::class_<PyTimeDisc>(m, "TimeDisc")
        .def(py::init<std::shared_ptr<PySpaceDisc>>(), "space_disc"_a)
        .def("cfl_dt", [](PyTimeDisc &t) { return t.td.CFLdt(); })
        .def("draining_dt", [](PyTimeDisc &t) { return t.td.DrainingDt(); });

    py::module_ solvers = m.def_submodule("Solvers", "Solvers::Euler / SSPRK2 / SSPRK3 (src/Solvers.cpp)");
    solvers.def("euler", [](PyTimeDisc &t, double dt) { Solvers::Euler(&t.td, dt); }, "td"_a, "dt"_a);
    solvers.def("ssprk2", [](PyTimeDisc &t, double dt) { Solvers::SSPRK2(&t.td, dt); }, "td"_a, "dt"_a);
    solvers.def("ssprk3", [](PyTimeDisc &t, double dt) { Solvers::SSPRK3(&t.td, dt); }, "td"_a, "dt"_a);
    solvers.def("run", [](PyTimeDisc &t, const std::string &scheme, Idx nsteps, double dt, double dt0) {
        const swe_scheme sc = scheme == "euler" ? SWE_EULER : scheme == "ssprk2" ? SWE_SSPRK2 : scheme == "ssprk3" ? SWE_SSPRK3
                              : throw std::invalid_argument("scheme must be 'euler', 'ssprk2' or 'ssprk3'");
        Solvers::Run(&t.td, sc, nsteps, dt, dt0); }, "td"_a, "scheme"_a, "nsteps"_a, "dt"_a = 0., "dt0"_a = 0.);
}
