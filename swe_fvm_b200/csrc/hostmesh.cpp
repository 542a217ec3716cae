// hostmesh.cpp — host-side mesh construction behind the C-ABI of include/swe_b200.h.
//
// Rebuilds the parts of the reference whose bodies are missing upstream:
//   * StructTriangMesh(ni,nj,h)        (include/StructTriangMesh.h:4-15, declaration only)
//   * TriangMesh(filename) Gmsh reader (examples/Main.cpp:174, docs/TriangMesh_8h_source.html)
// from the numbering conventions pinned by notebooks/topology.dat:
//   1. node id = gmsh tag - 1;  2. triangles in file order, nodes in file order (CCW);
//   3. edges: boundary line elements first (file order), then first-visit order over
//      (triangle, k) with edge k = {ip[k], ip[(k+1)%3]}; EdgePoints sorted ascending;
//   4. TriangEdges[k] joins ip[k], ip[k+1]; TriangTriangs[k] is the cell across it (-1 wall);
//   5. EdgeTriangs = (later-visiting triangle, earlier-visiting triangle), walls (owner,-1).
// Plus what the multi-GPU path needs: 1->4 refinement, RCB partition, sub-mesh extraction.
#include "hostmesh.hpp"
#include "../../include/swe_constants.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <numeric>
#include <sstream>
#include <cstdlib>
#include <string>
#include <thread>
#include <unordered_map>

namespace swe {

static thread_local std::string g_host_error;

// Launchers such as torchrun export OMP_NUM_THREADS=1, which would make the one-off set-up of a 64M-cell rank take
// half a minute; SWE_HOST_THREADS overrides, otherwise the cores are shared among the ranks of this node.
int host_threads() {
    static int n = 0;
    if (n) return n;
    if (const char *e = std::getenv("SWE_HOST_THREADS")) { n = std::max(1, std::atoi(e)); return n; }
    int hw = (int)std::thread::hardware_concurrency();
    if (hw <= 0) hw = 8;
    int lw = 1;
    if (const char *e = std::getenv("LOCAL_WORLD_SIZE")) lw = std::max(1, std::atoi(e));
    n = std::max(1, std::min(32, hw / lw));
    return n;
}
void set_host_error(const std::string &s) { g_host_error = s; }
const char *host_error() { return g_host_error.c_str(); }

// ---------------------------------------------------------------------------------------
// Generic topology builder (rules 3-5). tp is already filled; bnd = boundary node pairs.
// ---------------------------------------------------------------------------------------
namespace {

struct EdgeTable {  // open addressing, key = (lo << 32) | hi
    std::vector<uint64_t> keys;
    std::vector<int64_t> vals;
    uint64_t mask;
    explicit EdgeTable(size_t expected) {
        size_t cap = 16;
        while (cap < expected * 2 + 8) cap <<= 1;
        keys.assign(cap, ~0ull);
        vals.assign(cap, -1);
        mask = cap - 1;
    }
    static uint64_t hash(uint64_t k) {
        k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
        return k;
    }
    // returns reference to the value slot for key (inserted with -1 if new)
    int64_t &slot(uint64_t key) {
        uint64_t h = hash(key) & mask;
        while (keys[h] != ~0ull && keys[h] != key) h = (h + 1) & mask;
        keys[h] = key;
        return vals[h];
    }
};

inline uint64_t edge_key(int64_t a, int64_t b) {
    if (a > b) std::swap(a, b);
    return (uint64_t(a) << 32) | uint64_t(b);
}

}  // namespace

void build_topology(swe_hostmesh &m, const std::vector<int64_t> &bnd) {
    const int64_t nt = m.nt;
    const int64_t nb = (int64_t)bnd.size() / 2;
    EdgeTable tab((size_t)(nt * 3 / 2 + nb + 16));
    m.ep.clear(); m.et.clear();
    m.ep.reserve((size_t)(nt * 3 + 2 * nb));
    m.et.reserve((size_t)(nt * 3 + 2 * nb));
    m.te.assign((size_t)nt * 3, -1);
    m.tt.assign((size_t)nt * 3, -1);
    int64_t ne = 0;
    for (int64_t k = 0; k < nb; ++k) {
        int64_t a = bnd[2 * k], b = bnd[2 * k + 1];
        int64_t &s = tab.slot(edge_key(a, b));
        if (s >= 0) continue;  // duplicated line element
        s = ne++;
        m.ep.push_back(std::min(a, b)); m.ep.push_back(std::max(a, b));
        m.et.push_back(-1); m.et.push_back(-1);
    }
    for (int64_t t = 0; t < nt; ++t) {
        for (int k = 0; k < 3; ++k) {
            int64_t a = m.tp[3 * t + k], b = m.tp[3 * t + (k + 1) % 3];
            int64_t &s = tab.slot(edge_key(a, b));
            if (s < 0) {
                s = ne++;
                m.ep.push_back(std::min(a, b)); m.ep.push_back(std::max(a, b));
                m.et.push_back(t); m.et.push_back(-1);
            } else if (m.et[2 * s] < 0) {
                m.et[2 * s] = t;  // pre-registered boundary line, first (only) owner
            } else {
                m.et[2 * s + 1] = m.et[2 * s];  // (later, earlier)
                m.et[2 * s] = t;
            }
            m.te[3 * t + k] = s;
        }
    }
    m.ne = ne;
    for (int64_t t = 0; t < nt; ++t)
        for (int k = 0; k < 3; ++k) {
            int64_t e = m.te[3 * t + k];
            m.tt[3 * t + k] = (m.et[2 * e] == t) ? m.et[2 * e + 1] : m.et[2 * e];
        }
}

// ---------------------------------------------------------------------------------------
// StructTriangMesh: closed-form single pass with a running edge counter (m_curr_e upstream)
// ---------------------------------------------------------------------------------------
void build_struct(swe_hostmesh &m, int64_t ni, int64_t nj, double h, int64_t i0, int64_t j0) {
    const int64_t nv = (ni + 1) * (nj + 1);
    m.nn = nv + ni * nj;
    m.nt = 4 * ni * nj;
    m.ne = 6 * ni * nj + ni + nj;
    m.geom.resize((size_t)m.nn * 3);
    m.tp.resize((size_t)m.nt * 3); m.te.resize((size_t)m.nt * 3); m.tt.resize((size_t)m.nt * 3);
    m.ep.resize((size_t)m.ne * 2); m.et.resize((size_t)m.ne * 2);
    // Edge ids in closed form (the running counter m_curr_e of upstream's generator, resolved): a square visits its
    // edges in the order [bottom if j == 0], B-R diagonal, B-L diagonal, right, R-T diagonal, top, T-L diagonal,
    // [left if i == 0]; row 0 creates 7 edges per square (+1 for i == 0), the other rows 6 (+1).
    auto row_base = [&](int64_t j) { return j == 0 ? (int64_t)0 : (7 * ni + 1) + (j - 1) * (6 * ni + 1); };
    auto first = [&](int64_t j, int64_t i) { return row_base(j) + i * (j == 0 ? 7 : 6) + (i > 0 ? 1 : 0) + (j == 0 ? 1 : 0); };  // id of the B-R diagonal
    auto e_top = [&](int64_t j, int64_t i) { return first(j, i) + 4; };
    auto e_right = [&](int64_t j, int64_t i) { return first(j, i) + 2; };
#pragma omp parallel for schedule(static) num_threads(host_threads())
    for (int64_t j = 0; j <= nj; ++j)
        for (int64_t i = 0; i <= ni; ++i) {
            double *g = &m.geom[3 * (j * (ni + 1) + i)];
            g[0] = double(i0 + i) * h; g[1] = double(j0 + j) * h; g[2] = 0.;
        }
#pragma omp parallel for schedule(static) num_threads(host_threads())
    for (int64_t j = 0; j < nj; ++j)
        for (int64_t i = 0; i < ni; ++i) {
            const int64_t s = j * ni + i;
            const int64_t v00 = j * (ni + 1) + i, v10 = v00 + 1, v01 = v00 + (ni + 1), v11 = v01 + 1;
            const int64_t c = nv + s;
            double *g = &m.geom[3 * c];
            g[0] = (double(i0 + i) + 0.5) * h; g[1] = (double(j0 + j) + 0.5) * h; g[2] = 0.;
            const int64_t B = 4 * s, R = B + 1, T = B + 2, L = B + 3;
            int64_t *tp = &m.tp[3 * B];
            tp[0] = v00; tp[1] = v10; tp[2] = c;    // Bottom
            tp[3] = v10; tp[4] = v11; tp[5] = c;    // Right
            tp[6] = v11; tp[7] = v01; tp[8] = c;    // Top
            tp[9] = v01; tp[10] = v00; tp[11] = c;  // Left
            const int64_t f = first(j, i);
            const int64_t e_b1 = f, e_b2 = f + 1, e_rt = f + 2, e_r1 = f + 3, e_tp = f + 4, e_t1 = f + 5;
            const int64_t e_bot = j == 0 ? f - 1 : e_top(j - 1, i);
            const int64_t e_left = i == 0 ? f + 6 : e_right(j, i - 1);
            int64_t *te = &m.te[3 * B];
            te[0] = e_bot;  te[1] = e_b1; te[2] = e_b2;
            te[3] = e_rt;   te[4] = e_r1; te[5] = e_b1;
            te[6] = e_tp;   te[7] = e_t1; te[8] = e_r1;
            te[9] = e_left; te[10] = e_b2; te[11] = e_t1;
            const int64_t below = (j > 0) ? 4 * (s - ni) + 2 : -1, east = (i < ni - 1) ? 4 * (s + 1) + 3 : -1;
            const int64_t above = (j < nj - 1) ? 4 * (s + ni) : -1, west = (i > 0) ? 4 * (s - 1) + 1 : -1;
            int64_t *tt = &m.tt[3 * B];
            tt[0] = below; tt[1] = R; tt[2] = L;
            tt[3] = east;  tt[4] = T; tt[5] = B;
            tt[6] = above; tt[7] = L; tt[8] = R;
            tt[9] = west;  tt[10] = B; tt[11] = T;
            // every edge is written by the square that visits it first: sorted end points, (later, earlier) cells
            auto edge = [&](int64_t e, int64_t a, int64_t b, int64_t later, int64_t earlier) {
                m.ep[2 * e] = std::min(a, b); m.ep[2 * e + 1] = std::max(a, b);
                m.et[2 * e] = later; m.et[2 * e + 1] = earlier;
            };
            if (j == 0) edge(e_bot, v00, v10, B, -1);
            edge(e_b1, v10, c, R, B);
            edge(e_b2, c, v00, L, B);
            edge(e_rt, v10, v11, east >= 0 ? east : R, east >= 0 ? R : -1);
            edge(e_r1, v11, c, T, R);
            edge(e_tp, v11, v01, above >= 0 ? above : T, above >= 0 ? T : -1);
            edge(e_t1, v01, c, L, T);
            if (i == 0) edge(e_left, v01, v00, L, -1);
        }
}

// ---------------------------------------------------------------------------------------
// Gmsh ASCII 4.1 / 4.2 reader (examples/bowl.msh, notebooks/basic.msh). Unknown sections
// ($Entities, $PhysicalNames, $Projection ...) are skipped.
// ---------------------------------------------------------------------------------------
static int nodes_per_element_type(int type) {
    switch (type) {
        case 1: return 2;   // line
        case 2: return 3;   // triangle
        case 3: return 4;   // quad
        case 4: return 4;   // tet
        case 15: return 1;  // point
        case 8: return 3;   // 2nd-order line
        case 9: return 6;   // 2nd-order triangle
        default: return -1;
    }
}

int read_gmsh(swe_hostmesh &m, const char *path) {
    std::ifstream in(path);
    if (!in) { set_host_error(std::string("cannot open mesh file ") + path); return SWE_ERR_IO; }
    std::string line;
    double version = 0;
    std::vector<double> xyz;
    std::unordered_map<int64_t, int64_t> tag2id;
    bool dense_tags = true;
    std::vector<int64_t> tris, bnd;
    bool have_nodes = false, have_elems = false;
    while (std::getline(in, line)) {
        while (!line.empty() && (line.back() == '\r' || line.back() == ' ')) line.pop_back();
        if (line == "$MeshFormat") {
            int ftype = 0, dsize = 0;
            in >> version >> ftype >> dsize;
            if (ftype != 0) { set_host_error("binary Gmsh files are not supported"); return SWE_ERR_IO; }
            if (version < 4.0) { set_host_error("Gmsh format >= 4.1 required (README.md:21)"); return SWE_ERR_IO; }
        } else if (line == "$Nodes") {
            int64_t nblocks, nnodes, mintag, maxtag;
            in >> nblocks >> nnodes >> mintag >> maxtag;
            xyz.assign((size_t)nnodes * 3, 0.0);
            dense_tags = (mintag == 1 && maxtag == nnodes);
            int64_t next = 0;
            std::vector<int64_t> tags;
            for (int64_t b = 0; b < nblocks; ++b) {
                int64_t edim, etag, parametric, nb;
                in >> edim >> etag >> parametric >> nb;
                tags.resize((size_t)nb);
                for (int64_t k = 0; k < nb; ++k) in >> tags[k];
                for (int64_t k = 0; k < nb; ++k) {
                    double x, y, z;
                    in >> x >> y >> z;
                    for (int64_t q = 0; q < parametric * edim; ++q) { double dummy; in >> dummy; }
                    int64_t id = dense_tags ? tags[k] - 1 : next;
                    if (id < 0 || id >= nnodes) { set_host_error("node tag out of range"); return SWE_ERR_IO; }
                    if (!dense_tags) tag2id[tags[k]] = id;
                    ++next;
                    xyz[3 * id] = x; xyz[3 * id + 1] = y; xyz[3 * id + 2] = z;
                }
            }
            if (!in) { set_host_error("truncated $Nodes section"); return SWE_ERR_IO; }
            have_nodes = true;
        } else if (line == "$Elements") {
            int64_t nblocks, nelems, mintag, maxtag;
            in >> nblocks >> nelems >> mintag >> maxtag;
            for (int64_t b = 0; b < nblocks; ++b) {
                int64_t edim, etag, nb; int type;
                in >> edim >> etag >> type >> nb;
                int npe = nodes_per_element_type(type);
                if (npe < 0) { set_host_error("unsupported Gmsh element type " + std::to_string(type)); return SWE_ERR_IO; }
                for (int64_t k = 0; k < nb; ++k) {
                    int64_t tag, n[8];
                    in >> tag;
                    for (int q = 0; q < npe; ++q) {
                        in >> n[q];
                        n[q] = dense_tags ? n[q] - 1 : tag2id.at(n[q]);
                    }
                    if (type == 1) { bnd.push_back(n[0]); bnd.push_back(n[1]); }
                    else if (type == 2) { tris.push_back(n[0]); tris.push_back(n[1]); tris.push_back(n[2]); }
                }
            }
            if (!in) { set_host_error("truncated $Elements section"); return SWE_ERR_IO; }
            have_elems = true;
        }
    }
    if (!have_nodes || !have_elems || tris.empty()) {
        set_host_error("no $Nodes/$Elements with triangles found in mesh file");
        return SWE_ERR_IO;
    }
    m.nn = (int64_t)xyz.size() / 3;
    m.geom.assign((size_t)m.nn * 3, 0.0);
    for (int64_t p = 0; p < m.nn; ++p) {
        m.geom[3 * p] = xyz[3 * p];
        m.geom[3 * p + 1] = xyz[3 * p + 1];
        m.geom[3 * p + 2] = 0.0;  // bathymetry set by the caller (Domain::AtNode upstream)
    }
    m.nt = (int64_t)tris.size() / 3;
    m.tp.assign(tris.begin(), tris.end());
    // enforce CCW (all CCW in gmsh output; flip defensively, keeping the first node)
    for (int64_t t = 0; t < m.nt; ++t) {
        const double *a = &m.geom[3 * m.tp[3 * t]], *b = &m.geom[3 * m.tp[3 * t + 1]], *c = &m.geom[3 * m.tp[3 * t + 2]];
        double det = (b[0] - a[0]) * (c[1] - a[1]) - (c[0] - a[0]) * (b[1] - a[1]);
        if (det < 0) std::swap(m.tp[3 * t + 1], m.tp[3 * t + 2]);
    }
    build_topology(m, bnd);
    return SWE_OK;
}

// ---------------------------------------------------------------------------------------
// Uniform 1 -> 4 refinement: new node Nn + e at every edge midpoint; children
// (p0,m0,m2) (m0,p1,m1) (m2,m1,p2) (m0,m1,m2), all CCW; boundary lines split in two.
// ---------------------------------------------------------------------------------------
void refine(swe_hostmesh &out, const swe_hostmesh &in) {
    out.nn = in.nn + in.ne;
    out.geom.assign((size_t)out.nn * 3, 0.0);
    std::copy(in.geom.begin(), in.geom.end(), out.geom.begin());
    for (int64_t e = 0; e < in.ne; ++e) {
        const double *a = &in.geom[3 * in.ep[2 * e]], *b = &in.geom[3 * in.ep[2 * e + 1]];
        double *p = &out.geom[3 * (in.nn + e)];
        for (int c = 0; c < 3; ++c) p[c] = 0.5 * (a[c] + b[c]);
    }
    out.nt = 4 * in.nt;
    out.tp.resize((size_t)out.nt * 3);
    for (int64_t t = 0; t < in.nt; ++t) {
        const int64_t p0 = in.tp[3 * t], p1 = in.tp[3 * t + 1], p2 = in.tp[3 * t + 2];
        const int64_t m0 = in.nn + in.te[3 * t], m1 = in.nn + in.te[3 * t + 1], m2 = in.nn + in.te[3 * t + 2];
        int64_t *q = &out.tp[12 * t];
        q[0] = p0; q[1] = m0; q[2] = m2;
        q[3] = m0; q[4] = p1; q[5] = m1;
        q[6] = m2; q[7] = m1; q[8] = p2;
        q[9] = m0; q[10] = m1; q[11] = m2;
    }
    std::vector<int64_t> bnd;
    for (int64_t e = 0; e < in.ne; ++e)
        if (in.et[2 * e + 1] < 0) {
            bnd.push_back(in.ep[2 * e]); bnd.push_back(in.nn + e);
            bnd.push_back(in.nn + e); bnd.push_back(in.ep[2 * e + 1]);
        }
    build_topology(out, bnd);
}

// ---------------------------------------------------------------------------------------
// Partitioning: recursive coordinate bisection of centroids; sub-mesh extraction with
// `layers` rings of vertex-adjacent halo cells (local order = increasing global id so the
// (later, earlier) edge orientation and EdgeIndexer slots are those of the global mesh).
// ---------------------------------------------------------------------------------------
static void rcb_rec(const std::vector<double> &cx, const std::vector<double> &cy, std::vector<int64_t> &ids,
                    int64_t lo, int64_t hi, int32_t p0, int32_t np, int32_t *part) {
    if (np == 1) { for (int64_t k = lo; k < hi; ++k) part[ids[k]] = p0; return; }
    double xmin = 1e300, xmax = -1e300, ymin = 1e300, ymax = -1e300;
    for (int64_t k = lo; k < hi; ++k) {
        xmin = std::min(xmin, cx[ids[k]]); xmax = std::max(xmax, cx[ids[k]]);
        ymin = std::min(ymin, cy[ids[k]]); ymax = std::max(ymax, cy[ids[k]]);
    }
    const std::vector<double> &key = (xmax - xmin > ymax - ymin) ? cx : cy;
    int32_t npl = np / 2;
    int64_t mid = lo + (hi - lo) * npl / np;
    std::nth_element(ids.begin() + lo, ids.begin() + mid, ids.begin() + hi, [&](int64_t a, int64_t b) {
        return key[a] < key[b] || (key[a] == key[b] && a < b);
    });
    rcb_rec(cx, cy, ids, lo, mid, p0, npl, part);
    rcb_rec(cx, cy, ids, mid, hi, p0 + npl, np - npl, part);
}

void partition_rcb(const swe_hostmesh &m, int32_t nparts, int32_t *part) {
    std::vector<double> cx((size_t)m.nt), cy((size_t)m.nt);
    for (int64_t t = 0; t < m.nt; ++t) {
        const double *a = &m.geom[3 * m.tp[3 * t]], *b = &m.geom[3 * m.tp[3 * t + 1]], *c = &m.geom[3 * m.tp[3 * t + 2]];
        cx[t] = (a[0] + b[0] + c[0]) / 3.0; cy[t] = (a[1] + b[1] + c[1]) / 3.0;
    }
    std::vector<int64_t> ids((size_t)m.nt);
    std::iota(ids.begin(), ids.end(), 0);
    rcb_rec(cx, cy, ids, 0, m.nt, 0, nparts, part);
}

void extract(swe_hostmesh &out, const swe_hostmesh &g, const int32_t *part, int32_t rank, int32_t layers) {
    // node -> cells CSR
    std::vector<int64_t> nstart((size_t)g.nn + 1, 0);
    for (int64_t k = 0; k < 3 * g.nt; ++k) nstart[g.tp[k] + 1]++;
    for (int64_t p = 0; p < g.nn; ++p) nstart[p + 1] += nstart[p];
    std::vector<int64_t> ncell((size_t)3 * g.nt), fill(nstart.begin(), nstart.end() - 1);
    for (int64_t t = 0; t < g.nt; ++t)
        for (int k = 0; k < 3; ++k) ncell[fill[g.tp[3 * t + k]]++] = t;
    std::vector<int8_t> in((size_t)g.nt, 0);
    std::vector<int64_t> frontier;
    for (int64_t t = 0; t < g.nt; ++t)
        if (part[t] == rank) { in[t] = 1; frontier.push_back(t); }
    for (int32_t l = 0; l < layers; ++l) {
        std::vector<int64_t> next;
        for (int64_t t : frontier)
            for (int k = 0; k < 3; ++k) {
                int64_t p = g.tp[3 * t + k];
                for (int64_t q = nstart[p]; q < nstart[p + 1]; ++q) {
                    int64_t c = ncell[q];
                    if (!in[c]) { in[c] = 1; next.push_back(c); }
                }
            }
        frontier.swap(next);
    }
    out.global_cells.clear();
    for (int64_t t = 0; t < g.nt; ++t) if (in[t]) out.global_cells.push_back(t);
    out.nt = (int64_t)out.global_cells.size();
    out.owner.resize((size_t)out.nt);
    std::vector<int64_t> cell_l((size_t)g.nt, -1), node_l((size_t)g.nn, -1);
    for (int64_t l = 0; l < out.nt; ++l) { cell_l[out.global_cells[l]] = l; out.owner[l] = part[out.global_cells[l]]; }
    // nodes in increasing global id
    for (int64_t l = 0; l < out.nt; ++l)
        for (int k = 0; k < 3; ++k) node_l[g.tp[3 * out.global_cells[l] + k]] = 0;
    int64_t nn = 0;
    for (int64_t p = 0; p < g.nn; ++p) if (node_l[p] == 0) node_l[p] = nn++;
    out.nn = nn;
    out.geom.assign((size_t)nn * 3, 0.0);
    for (int64_t p = 0; p < g.nn; ++p)
        if (node_l[p] >= 0) std::copy(&g.geom[3 * p], &g.geom[3 * p] + 3, &out.geom[3 * node_l[p]]);
    out.tp.resize((size_t)out.nt * 3);
    for (int64_t l = 0; l < out.nt; ++l)
        for (int k = 0; k < 3; ++k) out.tp[3 * l + k] = node_l[g.tp[3 * out.global_cells[l] + k]];
    // Edges: keep the GLOBAL edge's orientation (ep order, et order) verbatim, so fluxes are
    // evaluated with the same (from, to) as in the undecomposed mesh.
    std::vector<int64_t> edge_l((size_t)g.ne, -1);
    int64_t ne = 0;
    for (int64_t l = 0; l < out.nt; ++l)
        for (int k = 0; k < 3; ++k) {
            int64_t e = g.te[3 * out.global_cells[l] + k];
            if (edge_l[e] < 0) edge_l[e] = 0;
        }
    for (int64_t e = 0; e < g.ne; ++e) if (edge_l[e] == 0) edge_l[e] = ne++;
    out.ne = ne;
    out.ep.assign((size_t)ne * 2, -1); out.et.assign((size_t)ne * 2, -1);
    out.global_edges.assign((size_t)ne, -1);
    for (int64_t e = 0; e < g.ne; ++e) {
        int64_t le = edge_l[e];
        if (le < 0) continue;
        out.global_edges[le] = e;
        out.ep[2 * le] = node_l[g.ep[2 * e]]; out.ep[2 * le + 1] = node_l[g.ep[2 * e + 1]];
        int64_t a = g.et[2 * e], b = g.et[2 * e + 1];
        int64_t la = cell_l[a], lb = (b >= 0) ? cell_l[b] : b;
        if (b >= 0 && lb < 0) lb = -1;             // neighbour outside the sub-mesh: cut = wall
        if (la < 0) { la = lb; lb = -1; }          // owner slot must hold the cell we do have
        out.et[2 * le] = la; out.et[2 * le + 1] = lb;
    }
    out.te.resize((size_t)out.nt * 3); out.tt.resize((size_t)out.nt * 3);
    for (int64_t l = 0; l < out.nt; ++l)
        for (int k = 0; k < 3; ++k) {
            int64_t gt = out.global_cells[l];
            out.te[3 * l + k] = edge_l[g.te[3 * gt + k]];
            int64_t nb = g.tt[3 * gt + k];
            out.tt[3 * l + k] = (nb >= 0) ? (cell_l[nb] >= 0 ? cell_l[nb] : -1) : nb;
        }
}

// ---------------------------------------------------------------------------------------
// Analytic cases (examples/Tests.h)
// ---------------------------------------------------------------------------------------
struct ThackerCoefs { double w, a, b; };

static ThackerCoefs classic_thacker(const swe_case &c) {  // examples/Tests.h:242-250
    ThackerCoefs k;
    k.w = std::sqrt(c.cor * c.cor + 8. * c.delta);
    double qz = (c.q0 - 0.5 * c.cor) * (c.q0 - 0.5 * c.cor);
    double rz = qz + 2. * c.H0 * c.H0 + c.p0 * c.p0 - 0.25 * k.w * k.w;
    k.a = std::sqrt(rz * rz + k.w * k.w * c.p0 * c.p0) / (rz + 0.5 * k.w * k.w);
    k.b = std::atan(k.w * c.p0 / rz);
    return k;
}

void case_eval(const swe_case &c, double x, double y, double t, double out[4]) {
    const double dx = x - c.mid_x, dy = y - c.mid_y;
    double b = 0, h = 0, u = 0, v = 0;
    switch (c.kind) {
        case SWE_CASE_LAKE_AT_REST:  // examples/Tests.h:37-42
            b = ((1. < x) && (x < 3.) && (1. < y) && (y < 3.)) ? -0.2 : -1.;
            h = std::max(0., -b);
            break;
        case SWE_CASE_CLASSIC_THACKER: {  // examples/Tests.h:54-56,147-161,256-279
            b = c.delta * (dx * dx + dy * dy - 1.0);
            ThackerCoefs k = classic_thacker(c);
            double ph = k.w * t + k.b;
            double den = 1. - k.a * std::cos(ph);
            double p = 0.5 * k.w * k.a * std::sin(ph) / den;
            double q = (c.q0 - 0.5 * c.cor) * (1. - k.a * std::cos(k.b)) / den + 0.5 * c.cor;
            u = p * dx + q * dy;
            v = q * (c.mid_x - x) + p * dy;
            double Hc = c.H0 * (1. - k.a * std::cos(k.b)) / den;
            double qz = (c.q0 - 0.5 * c.cor) * (c.q0 - 0.5 * c.cor);
            double az0 = (1. - k.a * std::cos(k.b)) * (1. - k.a * std::cos(k.b));
            double Hxx = (0.25 * k.w * k.w * (k.a * k.a - 1.) + qz * az0) / (den * den);
            double res = Hc + 0.5 * Hxx * dx * dx + 0.5 * Hxx * dy * dy;
            h = std::max(0., res);
            break;
        }
        case SWE_CASE_GAUSS_WAVE:  // examples/Main.cpp:183-186 (flat bed b = 0)
            b = 0.;
            h = 1. + std::exp(-5. * (dx * dx + dy * dy));
            break;
        case SWE_CASE_FULLY_WET: {  // SURVEY.md §8d fully-wet synthetic variant
            const double two_pi = 6.283185307179586476925286766559;
            b = 0.1 * std::sin(two_pi * x / c.length) * std::sin(two_pi * y / c.length) - 1.;
            h = c.amp * std::exp(-5. * (dx * dx + dy * dy)) - b;
            break;
        }
        case SWE_CASE_BOWL_HUMP:  // BowlTest bed + still lake at `level` + Gaussian hump
            b = c.delta * (dx * dx + dy * dy - 1.0);
            h = std::max(0., c.level + c.amp * std::exp(-5. * (dx * dx + dy * dy)) - b);
            break;
        default: break;
    }
    out[0] = b; out[1] = h; out[2] = u; out[3] = v;
}

// TriangAverage<3,n> (include/PointOperations.h:20-44) with run-time n; same loop order.
template <class F>
static void triang_average(const double *p0, const double *p1, const double *p2, int n, F &&f, double res[3]) {
    const double hq = 1. / n;
    double di[2] = {hq * (p1[0] - p0[0]), hq * (p1[1] - p0[1])};
    double dj[2] = {hq * (p2[0] - p0[0]), hq * (p2[1] - p0[1])};
    double dt[2] = {1. / 3. * (di[0] + dj[0]), 1. / 3. * (di[1] + dj[1])};
    double sum[3] = {0, 0, 0};
    double pi[2] = {p0[0], p0[1]};
    double o[3];
    auto add = [&](double x, double y) {
        f(x, y, o);
        sum[0] += hq * o[0]; sum[1] += hq * o[1]; sum[2] += hq * o[2];
    };
    for (int i = 0; i < n; i++) {
        double pt[2] = {pi[0] + dt[0], pi[1] + dt[1]};
        for (int j = 0; j < n - i - 1; j++) {
            add(pt[0], pt[1]);
            add(pt[0] + dt[0], pt[1] + dt[1]);
            pt[0] += dj[0]; pt[1] += dj[1];
        }
        add(pt[0], pt[1]);
        pi[0] += di[0]; pi[1] += di[1];
    }
    for (int k = 0; k < 3; ++k) res[k] = hq * sum[k];
}

void case_initial_state(const swe_case &c, const swe_hostmesh &m, int quad_n, double t, double *prim) {
    const double third = 1. / 3.;
    // the reference's IC loop carries `#pragma omp parallel for` too (examples/Main.cpp:210)
#pragma omp parallel for schedule(static) num_threads(host_threads())
    for (int64_t i = 0; i < m.nt; ++i) {
        const double *p0 = &m.geom[3 * m.tp[3 * i]], *p1 = &m.geom[3 * m.tp[3 * i + 1]], *p2 = &m.geom[3 * m.tp[3 * i + 2]];
        // VolumeDomainWrapper::At = Domain::T(i)[2] (src/ValueField.cpp:8-10, src/Bathymetry.cpp:24-27)
        const double cx = p0[0] * third + p1[0] * third + p2[0] * third;
        const double cy = p0[1] * third + p1[1] * third + p2[1] * third;
        const double bi = p0[2] * third + p1[2] * third + p2[2] * third;
        double x[3];
        if (c.kind == SWE_CASE_GAUSS_WAVE) {
            double o[4];
            case_eval(c, cx, cy, t, o);
            x[0] = o[1]; x[1] = 0.; x[2] = 0.;  // w sampled at the centroid, b = 0
        } else if (c.kind == SWE_CASE_LAKE_AT_REST) {
            // examples/Main.cpp:333-336: w = average of max(0, linear bed over the cell), u = v = 0
            if (p0[2] <= 0. && p1[2] <= 0. && p2[2] <= 0.) {
                x[0] = 0.;
            } else {
                const double det = (p1[0] - p0[0]) * (p2[1] - p0[1]) - (p2[0] - p0[0]) * (p1[1] - p0[1]);
                triang_average(p0, p1, p2, quad_n, [&](double px, double py, double *o) {
                    double l1 = ((px - p0[0]) * (p2[1] - p0[1]) - (p2[0] - p0[0]) * (py - p0[1])) / det;
                    double l2 = ((p1[0] - p0[0]) * (py - p0[1]) - (px - p0[0]) * (p1[1] - p0[1])) / det;
                    o[0] = std::max(0., p0[2] + l1 * (p1[2] - p0[2]) + l2 * (p2[2] - p0[2]));
                    o[1] = 0.; o[2] = 0.;
                }, x);
            }
            x[1] = 0.; x[2] = 0.;
        } else {
            triang_average(p0, p1, p2, quad_n, [&](double px, double py, double *o) {
                double r[4];
                case_eval(c, px, py, t, r);
                o[0] = r[1]; o[1] = r[2]; o[2] = r[3];
            }, x);
            x[0] += bi;  // examples/Main.cpp:221
        }
        // PrimAssigner::operator= (src/Assigners.cpp:8-20)
        double h = x[0] - bi;
        double *o = &prim[3 * i];
        if (!(h > SWE_WET_DEPTH)) { o[0] = bi; o[1] = 0.; o[2] = 0.; continue; }
        o[0] = x[0]; o[1] = x[1]; o[2] = x[2];
        if (h < SWE_DAMP_DEPTH) {
            double f = std::sqrt(2) * h / std::sqrt(h * h + SWE_DAMP_EPS_PRIM);
            o[1] *= f; o[2] *= f;
        }
    }
}

}  // namespace swe

// ---------------------------------------------------------------------------------------
// C-ABI
// ---------------------------------------------------------------------------------------
using swe::set_host_error;

extern "C" {

SWE_API int swe_hostmesh_struct(swe_hostmesh **out, int64_t ni, int64_t nj, double h, int64_t i0, int64_t j0) {
    if (!out || ni <= 0 || nj <= 0 || !(h > 0)) { set_host_error("swe_hostmesh_struct: ni, nj, h must be positive"); return SWE_ERR_INVALID; }
    if (4 * ni * nj > (int64_t)2000000000 / 3) { set_host_error("swe_hostmesh_struct: mesh too large for int32 device ids"); return SWE_ERR_INVALID; }
    auto *m = new (std::nothrow) swe_hostmesh();
    if (!m) return SWE_ERR_NOMEM;
    try { swe::build_struct(*m, ni, nj, h, i0, j0); }
    catch (const std::bad_alloc &) { delete m; set_host_error("out of host memory"); return SWE_ERR_NOMEM; }
    *out = m;
    return SWE_OK;
}

SWE_API int swe_hostmesh_gmsh(swe_hostmesh **out, const char *path) {
    if (!out || !path) { set_host_error("swe_hostmesh_gmsh: null argument"); return SWE_ERR_INVALID; }
    auto *m = new (std::nothrow) swe_hostmesh();
    if (!m) return SWE_ERR_NOMEM;
    int rc;
    try { rc = swe::read_gmsh(*m, path); }
    catch (const std::exception &e) { rc = SWE_ERR_IO; set_host_error(std::string("gmsh reader: ") + e.what()); }
    if (rc != SWE_OK) { delete m; return rc; }
    *out = m;
    return SWE_OK;
}

SWE_API int swe_hostmesh_from_triangles(swe_hostmesh **out, int64_t nn, const double *xy, int64_t nt,
                                        const int64_t *tri, int64_t nb, const int64_t *bnd) {
    if (!out || !xy || !tri || nn <= 0 || nt <= 0) { set_host_error("swe_hostmesh_from_triangles: bad argument"); return SWE_ERR_INVALID; }
    for (int64_t k = 0; k < 3 * nt; ++k)
        if (tri[k] < 0 || tri[k] >= nn) { set_host_error("triangle node id out of range"); return SWE_ERR_INVALID; }
    auto *m = new (std::nothrow) swe_hostmesh();
    if (!m) return SWE_ERR_NOMEM;
    m->nn = nn; m->nt = nt;
    m->geom.assign((size_t)nn * 3, 0.0);
    for (int64_t p = 0; p < nn; ++p) { m->geom[3 * p] = xy[2 * p]; m->geom[3 * p + 1] = xy[2 * p + 1]; }
    m->tp.assign(tri, tri + 3 * nt);
    std::vector<int64_t> b;
    if (bnd && nb > 0) b.assign(bnd, bnd + 2 * nb);
    swe::build_topology(*m, b);
    *out = m;
    return SWE_OK;
}

SWE_API int swe_hostmesh_refine(swe_hostmesh **out, const swe_hostmesh *in) {
    if (!out || !in) { set_host_error("swe_hostmesh_refine: null argument"); return SWE_ERR_INVALID; }
    auto *m = new (std::nothrow) swe_hostmesh();
    if (!m) return SWE_ERR_NOMEM;
    swe::refine(*m, *in);
    *out = m;
    return SWE_OK;
}

SWE_API void swe_hostmesh_free(swe_hostmesh *m) { delete m; }

SWE_API int swe_hostmesh_view(const swe_hostmesh *m, swe_mesh *v) {
    if (!m || !v) { set_host_error("swe_hostmesh_view: null argument"); return SWE_ERR_INVALID; }
    v->nn = m->nn; v->ne = m->ne; v->nt = m->nt;
    v->geometry = m->geom.data();
    v->edge_nodes = m->ep.data(); v->edge_elements = m->et.data();
    v->element_nodes = m->tp.data(); v->element_edges = m->te.data(); v->element_neighbours = m->tt.data();
    v->cor = 0.; v->tau = 0.;
    return SWE_OK;
}

SWE_API double *swe_hostmesh_geometry(swe_hostmesh *m) { return m ? m->geom.data() : nullptr; }

SWE_API int swe_hostmesh_extract(swe_hostmesh **out, const swe_hostmesh *g, const int32_t *part, int32_t rank, int32_t layers) {
    if (!out || !g || !part || layers < 0) { set_host_error("swe_hostmesh_extract: bad argument"); return SWE_ERR_INVALID; }
    auto *m = new (std::nothrow) swe_hostmesh();
    if (!m) return SWE_ERR_NOMEM;
    swe::extract(*m, *g, part, rank, layers);
    if (m->nt == 0) { delete m; set_host_error("swe_hostmesh_extract: rank owns no cells"); return SWE_ERR_INVALID; }
    *out = m;
    return SWE_OK;
}

SWE_API const int64_t *swe_hostmesh_global_cells(const swe_hostmesh *m) { return (m && !m->global_cells.empty()) ? m->global_cells.data() : nullptr; }
SWE_API const int32_t *swe_hostmesh_cell_owner(const swe_hostmesh *m) { return (m && !m->owner.empty()) ? m->owner.data() : nullptr; }

SWE_API int swe_partition_rcb(const swe_hostmesh *m, int32_t nparts, int32_t *part) {
    if (!m || !part || nparts < 1) { set_host_error("swe_partition_rcb: bad argument"); return SWE_ERR_INVALID; }
    swe::partition_rcb(*m, nparts, part);
    return SWE_OK;
}

SWE_API void swe_case_defaults(swe_case *c, int32_t kind, double mid_x, double mid_y, double length) {
    if (!c) return;
    std::memset(c, 0, sizeof(*c));
    c->kind = kind; c->mid_x = mid_x; c->mid_y = mid_y; c->length = length;
    c->cor = 0.; c->tau = 0.; c->delta = 1.; c->H0 = 0.5; c->p0 = 0.; c->q0 = 0.;
    c->level = 3.0;
    c->amp = (kind == SWE_CASE_FULLY_WET) ? 0.05 : 0.5;
}

SWE_API int swe_case_eval(const swe_case *c, double x, double y, double t, double out[4]) {
    if (!c || !out || c->kind < 0 || c->kind > SWE_CASE_BOWL_HUMP) { set_host_error("swe_case_eval: bad case"); return SWE_ERR_INVALID; }
    swe::case_eval(*c, x, y, t, out);
    return SWE_OK;
}

SWE_API int swe_case_set_bathymetry(const swe_case *c, swe_hostmesh *m) {
    if (!c || !m || c->kind < 0 || c->kind > SWE_CASE_BOWL_HUMP) { set_host_error("swe_case_set_bathymetry: bad argument"); return SWE_ERR_INVALID; }
#pragma omp parallel for schedule(static) num_threads(swe::host_threads())
    for (int64_t p = 0; p < m->nn; ++p) {
        double o[4];
        swe::case_eval(*c, m->geom[3 * p], m->geom[3 * p + 1], 0., o);
        m->geom[3 * p + 2] = o[0];
    }
    return SWE_OK;
}

SWE_API int swe_case_initial_state(const swe_case *c, const swe_hostmesh *m, int32_t quad_n, double t, double *prim) {
    if (!c || !m || !prim || quad_n < 1 || c->kind < 0 || c->kind > SWE_CASE_BOWL_HUMP) { set_host_error("swe_case_initial_state: bad argument"); return SWE_ERR_INVALID; }
    swe::case_initial_state(*c, *m, quad_n, t, prim);
    return SWE_OK;
}

}  // extern "C"
