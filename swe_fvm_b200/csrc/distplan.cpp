// distplan.cpp — host-side domain decomposition for the multi-GPU time step (no GPU needed).
//
// One rank = one GPU. A rank owns a set of cells plus a halo deep enough that ONE exchange of cell
// states per RK stage suffices: the new state of an owned cell depends on the reconstructions of its
// ring-2 cells (the fluxes of the neighbours' edges enter their draining dt, src/TimeDisc.cpp:25,47-65),
// which depend on ring-3 states and, through the node maxima of the part-wet pass
// (src/SpaceDisc.cpp:23,47-51), on every cell sharing a node with those. Halo cells are
// recomputed redundantly with the SAME kernels and the GLOBAL edge orientation, reductions are min/max
// only, so owned cells come out bit-identical to the single-GPU run for any number of ranks.
//
// The reference has no counterpart (it is single-threaded); this replaces what an
// MPI-style driver around Solvers::X (src/Solvers.cpp) would have to do.
#include "distplan.hpp"

#include <algorithm>
#include <cstring>
#include <numeric>

namespace swe {

static void finish_plan(swe_dist_plan &p) {
    const swe_hostmesh &m = *p.mesh;
    p.n_owned = 0;
    for (uint8_t o : p.owned) p.n_owned += o;
    // ordering classes (device ranges, in this order):
    //   0 interior: neither sent nor dependent on halo data
    //   1 sent to a peer, reconstruction stencil free of halo cells
    //   2 owned, but the stencil (the cell + its three edge neighbours) contains a halo cell
    //   3 halo cells (received every stage)
    // => K1 of classes 0-1 overlaps the exchange in flight, classes 2-3 wait for it; K4 of classes 1-2
    //    runs first and is followed at once by the pack-and-signal kernels.
    std::vector<uint8_t> sent((size_t)m.nt, 0);
    for (auto &pe : p.peers)
        for (int64_t c : pe.send) sent[(size_t)c] = 1;
    p.cls.assign((size_t)m.nt, 0);
#pragma omp parallel for schedule(static) num_threads(host_threads())
    for (int64_t t = 0; t < m.nt; ++t) {
        if (!p.owned[(size_t)t]) { p.cls[(size_t)t] = 3; continue; }
        bool dep = false;
        for (int k = 0; k < 3; ++k) {
            const int64_t j = m.tt[3 * t + k];
            if (j >= 0 && !p.owned[(size_t)j]) dep = true;
        }
        p.cls[(size_t)t] = dep ? 2 : (sent[(size_t)t] ? 1 : 0);
    }
    // CFL candidates are valid (and needed) only on edges that touch an owned cell
    p.cfl_mask.assign((size_t)m.ne, 0);
#pragma omp parallel for schedule(static) num_threads(host_threads())
    for (int64_t e = 0; e < m.ne; ++e) {
        const int64_t a = m.et[2 * e], b = m.et[2 * e + 1];
        p.cfl_mask[(size_t)e] = (p.owned[(size_t)a] || (b >= 0 && p.owned[(size_t)b])) ? 1 : 0;
    }
}

void strip_rows(int64_t nj, int32_t world, int32_t rank, int64_t &j0, int64_t &j1) {
    const int64_t base = nj / world, rem = nj % world;
    j0 = rank * base + std::min<int64_t>(rank, rem);
    j1 = j0 + base + (rank < rem ? 1 : 0);
}

// Structured strips: rank r owns rows [j0, j1) of squares of the global StructTriangMesh(ni, nj, h) and
// builds its block directly (no global mesh in memory), halo_rows rows of squares on each open side.
// Local cell order = global order restricted; node coordinates are bitwise those of the global mesh.
int plan_struct(swe_dist_plan &p, int32_t rank, int32_t world, int64_t ni, int64_t nj, double h, int32_t halo_rows) {
    if (world < 1 || rank < 0 || rank >= world || ni < 1 || nj < world) { set_host_error("swe_dist_plan_struct: bad arguments"); return SWE_ERR_INVALID; }
    int64_t j0, j1;
    strip_rows(nj, world, rank, j0, j1);
    if (world > 1 && j1 - j0 < halo_rows) { set_host_error("swe_dist_plan_struct: strip thinner than the halo"); return SWE_ERR_INVALID; }
    const int64_t lo = std::max<int64_t>(0, j0 - halo_rows), hi = std::min<int64_t>(nj, j1 + halo_rows);
    p.rank = rank; p.world = world;
    p.mesh = new swe_hostmesh();
    build_struct(*p.mesh, ni, hi - lo, h, 0, lo);
    const int64_t cpr = 4 * ni;  // cells per row of squares
    p.owned.assign((size_t)p.mesh->nt, 0);
    std::fill(p.owned.begin() + (j0 - lo) * cpr, p.owned.begin() + (j1 - lo) * cpr, 1);
    p.gcell.resize((size_t)p.mesh->nt);
    std::iota(p.gcell.begin(), p.gcell.end(), lo * cpr);
    auto rows = [&](int64_t ja, int64_t jb) {
        std::vector<int64_t> v((size_t)std::max<int64_t>(0, (jb - ja) * cpr));
        std::iota(v.begin(), v.end(), (ja - lo) * cpr);
        return v;
    };
    p.peers.clear();
    if (rank > 0) {  // the lower neighbour needs my first halo_rows rows, I need its last halo_rows rows
        int64_t pj0, pj1;
        strip_rows(nj, world, rank - 1, pj0, pj1);
        swe_dist_plan::Peer pe;
        pe.rank = rank - 1;
        pe.send = rows(j0, std::min(j1, j0 + halo_rows));
        pe.recv = rows(std::max(pj0, j0 - halo_rows), j0);
        p.peers.push_back(std::move(pe));
    }
    if (rank < world - 1) {
        int64_t pj0, pj1;
        strip_rows(nj, world, rank + 1, pj0, pj1);
        swe_dist_plan::Peer pe;
        pe.rank = rank + 1;
        pe.send = rows(std::max(j0, j1 - halo_rows), j1);
        pe.recv = rows(j1, std::min(pj1, j1 + halo_rows));
        p.peers.push_back(std::move(pe));
    }
    finish_plan(p);
    return SWE_OK;
}

// members of `rank`'s sub-mesh (its cells + `layers` vertex-adjacent rings), increasing global id
static void submesh_members(const swe_hostmesh &g, const std::vector<int64_t> &nstart, const std::vector<int64_t> &ncell,
                            const int32_t *part, int32_t rank, int32_t layers, std::vector<int64_t> &members) {
    std::vector<int8_t> in((size_t)g.nt, 0);
    std::vector<int64_t> frontier;
    for (int64_t t = 0; t < g.nt; ++t)
        if (part[t] == rank) { in[(size_t)t] = 1; frontier.push_back(t); }
    for (int32_t l = 0; l < layers; ++l) {
        std::vector<int64_t> next;
        for (int64_t t : frontier)
            for (int k = 0; k < 3; ++k) {
                const int64_t q0 = g.tp[3 * t + k];
                for (int64_t q = nstart[(size_t)q0]; q < nstart[(size_t)q0 + 1]; ++q) {
                    const int64_t c = ncell[(size_t)q];
                    if (!in[(size_t)c]) { in[(size_t)c] = 1; next.push_back(c); }
                }
            }
        frontier.swap(next);
    }
    members.clear();
    for (int64_t t = 0; t < g.nt; ++t) if (in[(size_t)t]) members.push_back(t);
}

// Any mesh, any partition vector (every rank passes the same global mesh and vector). Receive lists are
// the halo cells grouped by owner; send lists are what the peers' sub-meshes contain of my cells — each
// rank derives them itself by repeating the peers' (deterministic) ring growth, so no list exchange is needed.
int plan_mesh(swe_dist_plan &p, int32_t rank, int32_t world, const swe_hostmesh &g, const int32_t *part, int32_t layers) {
    if (world < 1 || rank < 0 || rank >= world || !part) { set_host_error("swe_dist_plan_mesh: bad arguments"); return SWE_ERR_INVALID; }
    for (int64_t t = 0; t < g.nt; ++t)
        if (part[t] < 0 || part[t] >= world) { set_host_error("swe_dist_plan_mesh: partition id out of range"); return SWE_ERR_INVALID; }
    p.rank = rank; p.world = world;
    p.mesh = new swe_hostmesh();
    extract(*p.mesh, g, part, rank, layers);
    const swe_hostmesh &m = *p.mesh;
    p.gcell = m.global_cells;
    p.owned.resize((size_t)m.nt);
    for (int64_t l = 0; l < m.nt; ++l) p.owned[(size_t)l] = m.owner[(size_t)l] == rank;
    p.peers.clear();
    if (world > 1) {
        std::vector<int64_t> nstart((size_t)g.nn + 1, 0);
        for (int64_t k = 0; k < 3 * g.nt; ++k) nstart[(size_t)g.tp[k] + 1]++;
        for (int64_t q = 0; q < g.nn; ++q) nstart[(size_t)q + 1] += nstart[(size_t)q];
        std::vector<int64_t> ncell((size_t)3 * g.nt), fill(nstart.begin(), nstart.end() - 1);
        for (int64_t t = 0; t < g.nt; ++t)
            for (int k = 0; k < 3; ++k) ncell[(size_t)fill[(size_t)g.tp[3 * t + k]]++] = t;
        std::vector<int64_t> members;
        for (int32_t peer = 0; peer < world; ++peer) {
            if (peer == rank) continue;
            swe_dist_plan::Peer pe;
            pe.rank = peer;
            for (int64_t l = 0; l < m.nt; ++l)
                if (m.owner[(size_t)l] == peer) pe.recv.push_back(l);
            submesh_members(g, nstart, ncell, part, peer, layers, members);
            for (int64_t gc : members)
                if (part[gc] == rank) {
                    const auto it = std::lower_bound(p.gcell.begin(), p.gcell.end(), gc);
                    pe.send.push_back((int64_t)(it - p.gcell.begin()));
                }
            if (!pe.send.empty() || !pe.recv.empty()) p.peers.push_back(std::move(pe));
        }
    }
    finish_plan(p);
    return SWE_OK;
}

}  // namespace swe

extern "C" {

SWE_API int swe_dist_plan_struct(swe_dist_plan **out, int32_t rank, int32_t world, int64_t ni, int64_t nj, double h) {
    if (!out) return SWE_ERR_INVALID;
    *out = nullptr;
    swe_dist_plan *p = new swe_dist_plan();
    const int rc = swe::plan_struct(*p, rank, world, ni, nj, h, 3);
    if (rc) { swe_dist_plan_free(p); return rc; }
    *out = p;
    return SWE_OK;
}
SWE_API int swe_dist_plan_mesh(swe_dist_plan **out, int32_t rank, int32_t world, const swe_hostmesh *global, const int32_t *part_nt) {
    if (!out || !global) return SWE_ERR_INVALID;
    *out = nullptr;
    swe_dist_plan *p = new swe_dist_plan();
    const int rc = swe::plan_mesh(*p, rank, world, *global, part_nt, 4);
    if (rc) { swe_dist_plan_free(p); return rc; }
    *out = p;
    return SWE_OK;
}
SWE_API void swe_dist_plan_free(swe_dist_plan *p) {
    if (!p) return;
    delete p->mesh;
    delete p;
}
SWE_API const swe_hostmesh *swe_dist_plan_local_mesh(const swe_dist_plan *p) { return p ? p->mesh : nullptr; }
SWE_API int64_t swe_dist_plan_owned_count(const swe_dist_plan *p) { return p ? p->n_owned : 0; }
SWE_API const uint8_t *swe_dist_plan_owned(const swe_dist_plan *p) { return p ? p->owned.data() : nullptr; }
SWE_API const int64_t *swe_dist_plan_global_cells(const swe_dist_plan *p) { return p ? p->gcell.data() : nullptr; }
SWE_API const uint8_t *swe_dist_plan_classes(const swe_dist_plan *p) { return p ? p->cls.data() : nullptr; }
SWE_API const uint8_t *swe_dist_plan_cfl_mask(const swe_dist_plan *p) { return p ? p->cfl_mask.data() : nullptr; }
SWE_API int32_t swe_dist_plan_npeers(const swe_dist_plan *p) { return p ? (int32_t)p->peers.size() : 0; }
SWE_API int swe_dist_plan_peer(const swe_dist_plan *p, int32_t k, int32_t *peer_rank, int64_t *nsend, const int64_t **send,
                               int64_t *nrecv, const int64_t **recv) {
    if (!p || k < 0 || k >= (int32_t)p->peers.size()) return SWE_ERR_INVALID;
    const auto &pe = p->peers[(size_t)k];
    if (peer_rank) *peer_rank = pe.rank;
    if (nsend) *nsend = (int64_t)pe.send.size();
    if (send) *send = pe.send.data();
    if (nrecv) *nrecv = (int64_t)pe.recv.size();
    if (recv) *recv = pe.recv.data();
    return SWE_OK;
}

}  // extern "C"
