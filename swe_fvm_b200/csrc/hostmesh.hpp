// hostmesh.hpp — owning host mesh behind the opaque swe_hostmesh handle of swe_b200.h.
// Unlike the reference's Topology (non-owning const& members, include/TriangMesh.h:56-64,
// the cause of the dangling pybind binding) this object owns its arrays.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/swe_b200.h"

struct swe_hostmesh {
    int64_t nn = 0, ne = 0, nt = 0;
    std::vector<double> geom;        // 3 x nn column-major (x, y, b)
    std::vector<int64_t> ep, et;     // ne x 2
    std::vector<int64_t> tp, te, tt; // nt x 3
    // filled by swe_hostmesh_extract only
    std::vector<int64_t> global_cells, global_edges;
    std::vector<int32_t> owner;
};

namespace swe {
int host_threads();  // threads for the host-side set-up loops (SWE_HOST_THREADS, else cores / LOCAL_WORLD_SIZE, <= 32)
void set_host_error(const std::string &s);
const char *host_error();
void build_topology(swe_hostmesh &m, const std::vector<int64_t> &bnd_pairs);
void build_struct(swe_hostmesh &m, int64_t ni, int64_t nj, double h, int64_t i0, int64_t j0);
int read_gmsh(swe_hostmesh &m, const char *path);
void refine(swe_hostmesh &out, const swe_hostmesh &in);
void partition_rcb(const swe_hostmesh &m, int32_t nparts, int32_t *part);
void extract(swe_hostmesh &out, const swe_hostmesh &g, const int32_t *part, int32_t rank, int32_t layers);
void case_eval(const swe_case &c, double x, double y, double t, double out[4]);
void case_initial_state(const swe_case &c, const swe_hostmesh &m, int quad_n, double t, double *prim);
}  // namespace swe
