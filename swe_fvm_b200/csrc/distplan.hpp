// distplan.hpp — host-side decomposition plan behind the opaque swe_dist_plan handle of swe_b200.h.
#pragma once
#include <cstdint>
#include <vector>

#include "hostmesh.hpp"

struct swe_dist_plan {
    int32_t rank = 0, world = 1;
    swe_hostmesh *mesh = nullptr;            // local sub-mesh: owned + halo cells (owned by the plan)
    std::vector<uint8_t> owned, cls, cfl_mask;  // per local cell / cell / edge
    std::vector<int64_t> gcell;              // local cell -> global cell id (increasing)
    struct Peer { int32_t rank; std::vector<int64_t> send, recv; };  // local cell ids, both sides agree on the order
    std::vector<Peer> peers;                 // increasing rank
    int64_t n_owned = 0;
};

namespace swe {
int plan_struct(swe_dist_plan &p, int32_t rank, int32_t world, int64_t ni, int64_t nj, double h, int32_t halo_rows);
int plan_mesh(swe_dist_plan &p, int32_t rank, int32_t world, const swe_hostmesh &g, const int32_t *part, int32_t layers);
}  // namespace swe
