// hostmesh.hpp — owning host mesh behind the opaque swe_hostmesh handle of swe_b200.h.
// Unlike the reference's Topology (non-owning const& members, include/TriangMesh.h:56-64,
// the cause of the dangling pybind binding) this object owns its arrays.
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "../../include/swe_b200.h"

namespace swe {
// std::vector whose resize() leaves new elements uninitialised: the big mesh arrays are filled by OpenMP loops right
// after the resize, and a value-initialising resize would first zero ~10 GB on ONE thread (2 s at 64M cells) and
// place every page on that thread's NUMA node.
template <class T>
struct default_init_allocator : std::allocator<T> {
    template <class U> struct rebind { using other = default_init_allocator<U>; };
    using std::allocator<T>::allocator;
    template <class U> void construct(U *p) noexcept(std::is_nothrow_default_constructible<U>::value) { ::new (static_cast<void *>(p)) U; }
    template <class U, class... A> void construct(U *p, A &&...a) { ::new (static_cast<void *>(p)) U(std::forward<A>(a)...); }
};
using dvec = std::vector<double, default_init_allocator<double>>;
using ivec = std::vector<int64_t, default_init_allocator<int64_t>>;
}  // namespace swe

struct swe_hostmesh {
    int64_t nn = 0, ne = 0, nt = 0;
    swe::dvec geom;        // 3 x nn column-major (x, y, b)
    swe::ivec ep, et;      // ne x 2
    swe::ivec tp, te, tt;  // nt x 3
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
