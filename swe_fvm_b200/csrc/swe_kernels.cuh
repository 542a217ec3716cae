// swe_kernels.cuh — the hot path: one RK stage = reconstruct (fast kernel + two list-driven passes
// for the wet/dry front) -> edge flux + CFL min -> draining dt -> stage update. Hand-written fp64
// CUDA for sm_100a; no tensor cores (nothing here is a dense contraction).
//
// Data layout in HBM (structure of arrays, int32 ids, device numbering):
//   cells : tt/te/tp[k*nt+i] (k-major so a warp reads 32 consecutive ids), cgeo[i] = 32-byte
//           packet (cx, cy, cb, bfull) gathered for neighbours with one sector, area[i]
//   nodes : node[p] = 32-byte packet (x, y, b, -)
//   edges : slotL/slotR[e] (where the two cells keep their side of edge e), en[e] = (nx, ny)
//           outward normal of EdgeTriangs[0], elen[e], dmin[e] = min(2A_l/L, 2A_r/L)
//   state : w/u/v[i] SoA, ping-pong buffers (no U0 copy kernels)
//   node maxima of the reconstructed surface (m_max_wp upstream) are NOT stored: the part-wet
//   pass gathers them over the node's incident cells (deterministic, no atomics on doubles)
//   edge-side values (m_edg, m_src upstream) are stored CELL-major, c*[k*nt+i] = value of cell
//   i's side of its k-th edge: written coalesced by the reconstruction, read coalesced by the
//   stage update, gathered once by the flux kernel through slotL/slotR.
//   te is sign-encoded: e >= 0 -> cell is EdgeTriangs(e)[0] (sgn = +1), ~e -> cell is [1] (-1).
#pragma once
#include <cuda_runtime.h>

#include "swe_device.cuh"
#include "swe_flux_registry.cuh"

namespace swe {

struct DevMesh {
    int nt, ne, nn;
    const int *tt, *te, *tp;
    const double4 *cgeo;
    const double *area, *cb;
    const double4 *node;
    const int *n2c_start, *n2c_cells;  // node -> incident cells (CSR), pass 2 and taps only
    const int *slotL, *slotR;
    const double2 *en;
    const double *elen, *dmin;  // dmin = +inf on edges excluded from the CFL min (multi-GPU halo edges)
};

struct DevFields {
    double *w, *u, *v;               // current state (stage input)
    double *ceh, *ceu, *cev;         // [3*nt] edge-side depth h_e and (damped) velocities
    double *cgx, *cgy;               // [nt] w-gradient of the cell's reconstruction (row 0 of m_grad);
                                     // m_src = (cgx, cgy) + cor * (-v_e, u_e) is formed in the update
    double *cew;                     // [3*nt] edge-side w, taps only (nullable)
    double *f0, *f1, *f2;            // [ne] fluxes
    double *dti;                     // [nt] draining dt
    double *pwl;                     // [nt] pass-1 free-surface level of the part-wet cells (PartWet1), read by pass 2
    // Dry-region skipping. tile_dry[t] (written by K1, preset to 1) stays 1 iff EVERY cell of the 128-cell tile t is
    // "deep dry" in the stage-begin state: dry, stored canonically as (cb, +0, +0), and all its edge neighbours dry. Then
    // every flux through its edges is exactly +0, its draining dt is 0 and the stage update returns (cb, +0, +0) again,
    // so K2 writes zeros without reading the edge states, K3 writes 0 and K4 leaves the tile alone — bit-identical to
    // doing the work. td / td0: the flags K2-K4 may use for this stage / for the state saved by swe_save_state (an
    // all-zero array when they do not describe the current data).
    // tdp: the flags of the PREVIOUS complete reconstruction pass (all zero when unknown). Invariant kept by K1: while a
    // tile stays flagged from one pass to the next, the edge-side values of its cells are (0, 0, 0), their gradient is
    // the bed slope and their class is 0 — written when the tile first became deep dry — so a deep-dry cell of a tile
    // that was already flagged has nothing to compute or store.
    unsigned char *tile_dry;
    const unsigned char *td, *td0, *tdp;
    // tdf / tdd: the previous pass's flags as far as the FLUX / DRAINING-DT arrays are concerned: a tile flagged then and
    // now already holds +0 fluxes on all its edges / dt = 0 in all its cells (written or computed after that pass), so
    // even the zero stores are skipped. All zero unless a flux / draining pass ran after the previous reconstruction.
    const unsigned char *tdf, *tdd;
    // dry-region form of the stage update: compacted ids of the tiles it has to process ([0] = count, then the ids), built
    // by k_tile_compact right before it; a block beyond the count exits after one cached load instead of waiting for its
    // own flag byte (half a million blocks of one flag load each cost 0.56 ms at 64M cells, measured)
    int *tile_list;
    signed char *cls;                // [nt] 0 dry, 1 part-wet, 2 full-wet
    int *pw_list;                    // [nt] compacted ids of part-wet cells (pass 2 work list)
    int *rs_list;                    // [nt] cells left to the generic reconstruction kernel (K1s)
    double *scal;                    // [0] min_len_to_wavespeed, [1] dt, [2] time, [3] running min
    int *flags;                      // [0] non-finite state seen, [1] part-wet list count, [3] flux block ticket,
                                     // [4] generic-reconstruction list count, [5] peer-memory wait timed out
    unsigned long long *dbg;         // [12] branch-hit counters (TAPS builds only; see swe_get_branch_counts)
    int recon;                       // S2: 0 repaired (w,u,v plane gradients), 1 as written upstream, 2 first order
    int pw2;                         // S3: 0 PartWet2 points(r,c) read as (point, coordinate), 1 as written
};
// branch-hit counter slots (the CPU checker in the test tree keeps the same twelve counters)
enum { BR_PW1_SUBMERGED = 0, BR_PW1_CBRT, BR_PW1_BISECTION, BR_FW_DRY_NB, BR_FW_PW_NB, BR_FW_VERTEX_ZERO, BR_FW_TVD_OFF,
       BR_PW2_TO_PW1, BR_PW2_ONE_WET, BR_PW2_THREE_WET, BR_PW2_TWO_WET, BR_PW2_FALLBACK, BR_COUNT };
template <bool TAPS> __device__ __forceinline__ void count_branch(const DevFields &s, int b) {
    if (TAPS) atomicAdd(&s.dbg[b], 1ull);
}

#ifndef SWE_K1_SPLIT
#define SWE_K1_SPLIT 1
#endif
constexpr int kBlock = 128;
constexpr int kUpdTileShift = 7;  // tiles of kBlock = 128 consecutive cells (device numbering): stage update blocks, dry-region flags
static_assert((1 << kUpdTileShift) == kBlock, "tile = one thread block of the stage update");

// stores of write-once intermediates (edge-side values, gradients, fluxes, dti) may carry the
// evict-first hint so that they do not displace the gathered neighbour data in L2 (SWE_STCS=1)
#ifndef SWE_STCS
#define SWE_STCS 0
#endif
__device__ __forceinline__ void st_once(double *p, double v) {
#if SWE_STCS
    __stcs(p, v);
#else
    *p = v;
#endif
}

// MUSCL::AtPoint (include/MUSCLObject.h:26-33) for one component set
struct Muscl {
    double o0, o1, o2;
    double g00, g01, g10, g11, g20, g21;
};

// UpdateInterfaceValues (src/SpaceDisc.cpp:15-31) + PrimAssigner (src/Assigners.cpp:8-20) for
// local edge k whose midpoint is (mx, my, mb); writes the cell-major slots.
template <bool TAPS>
__device__ __forceinline__ void emit_edge(const DevFields &s, int slot, const Muscl &M, double cx, double cy,
                                          double mx, double my, double mb) {
    const double dx = mx - cx, dy = my - cy;
    const double a0 = M.o0 + (M.g00 * dx + M.g01 * dy);
    double eu = M.o1 + (M.g10 * dx + M.g11 * dy);
    double ev = M.o2 + (M.g20 * dx + M.g21 * dy);
    // AtPoint's dry clamp (a0 - mb < 0 -> (mb,0,0)) followed by PrimAssigner's (h <= 1e-12 ->
    // (mb,0,0)) collapse into one test: a clamped point has h = 0, which is not wet either.
    const double h = a0 - mb;
    double ew = a0, eh = h;
    if (!is_wet(h)) {
        ew = mb; eh = 0.; eu = 0.; ev = 0.;
    } else if (h < SWE_DAMP_DEPTH) {
        const double fac = sqrt(2.0) * h / sqrt(h * h + SWE_DAMP_EPS_PRIM);
        eu *= fac; ev *= fac;
    }
    st_once(s.ceh + slot, eh); st_once(s.ceu + slot, eu); st_once(s.cev + slot, ev);
    if (TAPS) s.cew[slot] = ew;
    // m_src (src/SpaceDisc.cpp:26-29) = Gradient(edge).row(0) + cor * (-v_e, u_e). MUSCL::Gradient(p)
    // (include/MUSCLObject.h:41-48) always evaluates to m_grad: when the reconstructed depth is
    // negative AtPoint returns the dry state, whose depth is exactly 0. So only the cell's
    // (g00, g01) is stored (once per cell) and the sum is formed where it is consumed.
}

// K1 dispatch: persistent grid-stride, exactly one resident wave (4 CTAs/SM at 128 registers for the
// fast path; the generic routine needs ~160).
// Measured alternatives that were slower at 64M cells (profiles/README.md): one CTA per 128-cell
// tile (3.78 vs 3.30 ms), in-order atomic tile dispatch (3.66), register caps that spill (5.2-5.9),
// shared-memory parking of the neighbour states for 16-20 warps/SM (3.55-3.69), a cp.async (LDGSTS)
// double-buffered smem staging of all 36 inputs (5.1: no L1 left beside 221 KB smem and
// persistent CTAs drift apart), register-free prefetch.global of the next cell's gathers (3.68).
#ifndef SWE_K1_BLOCK
#define SWE_K1_BLOCK 128
#endif
constexpr int kK1Block = SWE_K1_BLOCK;
#ifndef SWE_K1_MIN_BLOCKS
#define SWE_K1_MIN_BLOCKS 4
#endif
#ifndef SWE_K1_GRID_PER_SM
#define SWE_K1_GRID_PER_SM 4
#endif
// 32-byte read-only gather of one geometry packet. sm_100 has 256-bit global loads (ld.global.nc.v4.f64 ->
// LDG.E.ENL2.256.CONSTANT): one instruction and one L1 request per packet instead of two 128-bit ones. The packet
// arrays come from cudaMalloc and hold 32-byte elements, so every element is 32-byte aligned.
#ifndef SWE_LD256
#define SWE_LD256 1
#endif
__device__ __forceinline__ double4 ldg4(const double4 *p) {
#if SWE_LD256
    double4 r;
    asm("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
    return r;
#else
    const double2 a = __ldg(reinterpret_cast<const double2 *>(p));
    const double2 b = __ldg(reinterpret_cast<const double2 *>(p) + 1);
    return make_double4(a.x, a.y, b.x, b.y);
#endif
}
template <bool TAPS>
__device__ __forceinline__ void reconstruct_cell(const DevMesh &m, const DevFields &s, const int i, const int ip0,
                                                 const int ip1, const int ip2, const int it0, const int it1,
                                                 const int it2) {
    const int nt = m.nt;
    // all gathers are issued up front (two dependent round trips in total: ids -> data); the
    // neighbour ids of boundary triangles are clamped, their values are never used
    const int jt[3] = {max(it0, 0), max(it1, 0), max(it2, 0)};
    const double4 P0 = ldg4(m.node + ip0), P1 = ldg4(m.node + ip1), P2 = ldg4(m.node + ip2);
    const double4 Gi = ldg4(m.cgeo + i);
    const double w = __ldg(s.w + i), u = __ldg(s.u + i), v = __ldg(s.v + i);
    double N[3][3];
    double4 Gn[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        N[k][0] = __ldg(s.w + jt[k]); N[k][1] = __ldg(s.u + jt[k]); N[k][2] = __ldg(s.v + jt[k]);
        Gn[k] = ldg4(m.cgeo + jt[k]);
    }
    const double cx = Gi.x, cy = Gi.y, cb = Gi.z;

    // edge midpoints E(ie[k]) = 0.5 (P(ep0) + P(ep1)), edge k joins ip[k], ip[k+1] (S1)
    const double mxk[3] = {0.5 * (P0.x + P1.x), 0.5 * (P1.x + P2.x), 0.5 * (P2.x + P0.x)};
    const double myk[3] = {0.5 * (P0.y + P1.y), 0.5 * (P1.y + P2.y), 0.5 * (P2.y + P0.y)};
    const double mbk[3] = {0.5 * (P0.z + P1.z), 0.5 * (P1.z + P2.z), 0.5 * (P2.z + P0.z)};

    const bool bnd = (it0 | it1 | it2) < 0;
    const double bmax = smax(smax(P0.z, P1.z), P2.z);
    const bool dry = !is_wet(w - cb);
    const bool full = !bnd && (bmax < w);
    s.cls[i] = dry ? 0 : (full ? 2 : 1);
#if !SWE_K1_SPLIT
    s.tile_dry[i >> kUpdTileShift] = 0;  // the generic routine does not classify tiles: nothing is skipped
#endif

    Muscl M;
    M.g00 = M.g01 = M.g10 = M.g11 = M.g20 = M.g21 = 0.;
    if (dry) {  // ReconstructDryCell (:31-36): origin (b_i,0,0), w-gradient = bed slope
        M.o0 = cb; M.o1 = 0.; M.o2 = 0.;
        gradient3(P0.x, P0.y, P0.z, P1.x, P1.y, P1.z, P2.x, P2.y, P2.z, M.g00, M.g01);
    } else if (!full) {  // ReconstructPartWetCell1 (:86-112)
        const double bmin = smin(smin(P0.z, P1.z), P2.z);
        int br = 0;
        M.o0 = partwet1_level(w, cb, bmax, bmin, &br); M.o1 = u; M.o2 = v;
        s.pwl[i] = M.o0;  // pass 2 evaluates this cell's pass-1 surface at its nodes: keep the level (up to 50 bisection steps)
        count_branch<TAPS>(s, BR_PW1_SUBMERGED + br);
        s.pw_list[atomicAdd(&s.flags[1], 1)] = i;
    } else {  // ReconstructFullWetCell (:38-84), S2: plane gradients of w, u, v
        M.o0 = w; M.o1 = u; M.o2 = v;
        double X[3][2], V[3][3], Z[3];  // Z: bed of the support point (the as-written w-gradient uses it)
        bool zero_grad = false, pw_nb = false;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int j = jt[k];
            const double wj = N[k][0], uj = N[k][1], vj = N[k][2];
            const double4 Gj = Gn[k];
            if (Gj.w < wj) {  // IsFullWetCell(j): bfull = +inf on boundary triangles
                X[k][0] = Gj.x; X[k][1] = Gj.y; Z[k] = Gj.z;
                V[k][0] = wj; V[k][1] = uj; V[k][2] = vj;
            } else if (!is_wet(wj - Gj.z)) {  // IsDryCell(j) -> zero gradient
                zero_grad = true;
                X[k][0] = X[k][1] = 0.; Z[k] = 0.; V[k][0] = V[k][1] = V[k][2] = 0.;
            } else {  // part-wet neighbour (rare): its PartWet1 value at the shared edge midpoint
                if (!zero_grad) pw_nb = true;  // upstream returns at the first dry neighbour (:52-53)
                const double z0 = m.node[m.tp[j]].z, z1 = m.node[m.tp[nt + j]].z, z2 = m.node[m.tp[2 * nt + j]].z;
                const double b13 = smax(smax(z0, z1), z2), b23 = smin(smin(z0, z1), z2);
                double a0 = partwet1_level(wj, Gj.z, b13, b23, nullptr), a1 = uj, a2 = vj;
                // AtPoint with a zero gradient: o + (0*dx + 0*dy)
                const double dx = mxk[k] - Gj.x, dy = myk[k] - Gj.y;
                a0 = a0 + (0. * dx + 0. * dy); a1 = a1 + (0. * dx + 0. * dy); a2 = a2 + (0. * dx + 0. * dy);
                if (!((a0 - mbk[k]) >= 0)) { a0 = mbk[k]; a1 = 0.; a2 = 0.; }
                X[k][0] = mxk[k]; X[k][1] = myk[k]; Z[k] = mbk[k];
                V[k][0] = 0.5 * (w + a0); V[k][1] = 0.5 * (u + a1); V[k][2] = 0.5 * (v + a2);
            }
        }
        if (zero_grad) count_branch<TAPS>(s, BR_FW_DRY_NB);
        if (pw_nb) count_branch<TAPS>(s, BR_FW_PW_NB);
        if (!zero_grad && s.recon != 2) {  // recon == 2: first order, the gradient stays zero
            const Lu2 lu = lu2_factor(X[0][0], X[0][1], X[1][0], X[1][1], X[2][0], X[2][1]);
            double df[3][2];
            const bool asw = s.recon == 1;  // as written (:63-64): df.row(0) = Gradient(grad_points), u/v rows zero
            if (asw) { V[0][0] = Z[0]; V[1][0] = Z[1]; V[2][0] = Z[2]; }
            lu2_solve(lu, V[0][0], V[1][0], V[2][0], df[0][0], df[0][1]);
            // vertex positivity (:66-72): dx = P(ip) * (I - 1/3), evaluated literally
            const double md = 1. - 1. / 3., mo = 0. - 1. / 3.;
            const double dx0 = (P0.x * md + P1.x * mo) + P2.x * mo, dy0 = (P0.y * md + P1.y * mo) + P2.y * mo;
            const double dx1 = (P0.x * mo + P1.x * md) + P2.x * mo, dy1 = (P0.y * mo + P1.y * md) + P2.y * mo;
            const double dx2 = (P0.x * mo + P1.x * mo) + P2.x * md, dy2 = (P0.y * mo + P1.y * mo) + P2.y * md;
            const double hp0 = ((df[0][0] * dx0 + df[0][1] * dy0) + w) - P0.z;
            const double hp1 = ((df[0][0] * dx1 + df[0][1] * dy1) + w) - P1.z;
            const double hp2 = ((df[0][0] * dx2 + df[0][1] * dy2) + w) - P2.z;
            const bool positive = is_wet(hp0) && is_wet(hp1) && is_wet(hp2);
            lu2_solve(lu, V[0][1], V[1][1], V[2][1], df[1][0], df[1][1]);
            lu2_solve(lu, V[0][2], V[1][2], V[2][2], df[2][0], df[2][1]);
            if (asw) df[1][0] = df[1][1] = df[2][0] = df[2][1] = 0.;
            if (!positive) {
                count_branch<TAPS>(s, BR_FW_VERTEX_ZERO);
#pragma unroll
                for (int c = 0; c < 3; ++c) df[c][0] = df[c][1] = 0.;
            }
            // on/off TVD limiter (:74-81) against the neighbours' cell means
            double tvd[3] = {1., 1., 1.};
            const double own[3] = {w, u, v};
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const double dx = mxk[k] - cx, dy = myk[k] - cy;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const double lo = smin(own[c], N[k][c]), hi = smax(own[c], N[k][c]);
                    const double vek = own[c] + (df[c][0] * dx + df[c][1] * dy);
                    if (!((lo <= vek) && (vek <= hi))) tvd[c] = 0.;
                }
            }
            if (tvd[0] == 0. || tvd[1] == 0. || tvd[2] == 0.) count_branch<TAPS>(s, BR_FW_TVD_OFF);
            M.g00 = tvd[0] * df[0][0]; M.g01 = tvd[0] * df[0][1];
            M.g10 = tvd[1] * df[1][0]; M.g11 = tvd[1] * df[1][1];
            M.g20 = tvd[2] * df[2][0]; M.g21 = tvd[2] * df[2][1];
        }
    }
    // UpdateInterfaceValues (src/SpaceDisc.cpp:15-31); the node maxima (:23) are gathered in pass 2
    st_once(s.cgx + i, M.g00); st_once(s.cgy + i, M.g01);
    emit_edge<TAPS>(s, i, M, cx, cy, mxk[0], myk[0], mbk[0]);
    emit_edge<TAPS>(s, nt + i, M, cx, cy, mxk[1], myk[1], mbk[1]);
    emit_edge<TAPS>(s, 2 * nt + i, M, cx, cy, mxk[2], myk[2], mbk[2]);
}

// K1 fast path: dry cells and full-wet cells whose three neighbours are all full-wet (the bulk of
// any mesh). Without the part-wet patching logic the support points are the neighbour centroids and
// the support values the neighbour means, which needs fewer registers and selects. Every other
// cell (part-wet cells, full-wet cells next to a dry / part-wet one) is appended to rs_list and
// reconstructed by the generic routine in k_reconstruct_slow — same formulas, same bits.
// arithmetic of the fast path on already loaded data (shared by the gather kernel and the tiled kernel)
template <bool TAPS, int RECON, bool DRY = false>
__device__ __forceinline__ bool fast_compute(const DevFields &s, const int nt, const int i, const int wall, const double4 P0,
                                             const double4 P1, const double4 P2, const double4 Gi, const double w,
                                             const double u, const double v, const double N00, const double N01,
                                             const double N02, const double N10, const double N11, const double N12,
                                             const double N20, const double N21, const double N22, const double4 G0,
                                             const double4 G1, const double4 G2);
// DRY: bit k set iff side k of the cell is a wall (no neighbour); otherwise only "any side is a wall" matters
template <bool DRY>
__device__ __forceinline__ int wall_mask(int it0, int it1, int it2) {
    if (DRY) return (it0 < 0 ? 1 : 0) | (it1 < 0 ? 2 : 0) | (it2 < 0 ? 4 : 0);
    return ((it0 | it1 | it2) < 0) ? 1 : 0;
}

template <bool TAPS, int RECON, bool DRY>
__device__ __forceinline__ bool reconstruct_cell_fast(const DevMesh &m, const DevFields &s, const int i, const int ip0,
                                                      const int ip1, const int ip2, const int it0, const int it1,
                                                      const int it2) {
    const int j0 = max(it0, 0), j1 = max(it1, 0), j2 = max(it2, 0);
    const double4 P0 = ldg4(m.node + ip0), P1 = ldg4(m.node + ip1), P2 = ldg4(m.node + ip2);
    const double4 Gi = ldg4(m.cgeo + i);
    const double w = __ldg(s.w + i), u = __ldg(s.u + i), v = __ldg(s.v + i);
    const double N00 = __ldg(s.w + j0), N01 = __ldg(s.u + j0), N02 = __ldg(s.v + j0);
    const double N10 = __ldg(s.w + j1), N11 = __ldg(s.u + j1), N12 = __ldg(s.v + j1);
    const double N20 = __ldg(s.w + j2), N21 = __ldg(s.u + j2), N22 = __ldg(s.v + j2);
    const double4 G0 = ldg4(m.cgeo + j0), G1 = ldg4(m.cgeo + j1), G2 = ldg4(m.cgeo + j2);
    return fast_compute<TAPS, RECON, DRY>(s, m.nt, i, wall_mask<DRY>(it0, it1, it2), P0, P1, P2, Gi, w, u, v, N00, N01, N02, N10, N11,
                                          N12, N20, N21, N22, G0, G1, G2);
}

template <bool TAPS, int RECON, bool DRY>
__device__ __forceinline__ bool fast_compute(const DevFields &s, const int nt, const int i, const int wall, const double4 P0,
                                             const double4 P1, const double4 P2, const double4 Gi, const double w,
                                             const double u, const double v, const double N00, const double N01,
                                             const double N02, const double N10, const double N11, const double N12,
                                             const double N20, const double N21, const double N22, const double4 G0,
                                             const double4 G1, const double4 G2) {
    const double cx = Gi.x, cy = Gi.y, cb = Gi.z;
    const bool bnd = wall != 0;
    const bool dry = !is_wet(w - cb);
    if (DRY) {  // deep-dry test for the tile flag (see DevFields::tile_dry): one cell that fails clears its tile
        const bool canon = __double_as_longlong(w) == __double_as_longlong(cb) && __double_as_longlong(u) == 0ll &&
                           __double_as_longlong(v) == 0ll;
        const bool nb_dry = ((wall & 1) || !is_wet(N00 - G0.z)) && ((wall & 2) || !is_wet(N10 - G1.z)) &&
                            ((wall & 4) || !is_wet(N20 - G2.z));
        if (!(canon && nb_dry)) s.tile_dry[i >> kUpdTileShift] = 0;
        else if (__ldg(s.tdp + (i >> kUpdTileShift))) return true;  // outputs already in place (see DevFields::tdp)
    }
    const bool full = !bnd && (smax(smax(P0.z, P1.z), P2.z) < w);
    const bool nbfull = (G0.w < N00) && (G1.w < N10) && (G2.w < N20);
    if (!dry && !(full && nbfull)) return false;
    s.cls[i] = dry ? 0 : 2;

    const double mx0 = 0.5 * (P0.x + P1.x), my0 = 0.5 * (P0.y + P1.y), mb0 = 0.5 * (P0.z + P1.z);
    const double mx1 = 0.5 * (P1.x + P2.x), my1 = 0.5 * (P1.y + P2.y), mb1 = 0.5 * (P1.z + P2.z);
    const double mx2 = 0.5 * (P2.x + P0.x), my2 = 0.5 * (P2.y + P0.y), mb2 = 0.5 * (P2.z + P0.z);
    Muscl M;
    M.g00 = M.g01 = M.g10 = M.g11 = M.g20 = M.g21 = 0.;
    if (dry) {  // ReconstructDryCell (src/MUSCLObject.cpp:31-36)
        M.o0 = cb; M.o1 = 0.; M.o2 = 0.;
        gradient3(P0.x, P0.y, P0.z, P1.x, P1.y, P1.z, P2.x, P2.y, P2.z, M.g00, M.g01);
    } else if (RECON == 2) {  // first-order option: MUSCL{prim(i), 0}
        M.o0 = w; M.o1 = u; M.o2 = v;
    } else {  // ReconstructFullWetCell (:38-84) with three full-wet neighbours
        M.o0 = w; M.o1 = u; M.o2 = v;
        const Lu2 lu = lu2_factor(G0.x, G0.y, G1.x, G1.y, G2.x, G2.y);
        double d00, d01, d10, d11, d20, d21;
        // RECON 1 = as written upstream (:63-64): the w-gradient is the slope of the BED values of the
        // support points (Gradient(grad_points)), the u and v rows stay zero
        if (RECON == 1) lu2_solve(lu, G0.z, G1.z, G2.z, d00, d01);
        else lu2_solve(lu, N00, N10, N20, d00, d01);
        const double md = 1. - 1. / 3., mo = 0. - 1. / 3.;
        const double dx0 = (P0.x * md + P1.x * mo) + P2.x * mo, dy0 = (P0.y * md + P1.y * mo) + P2.y * mo;
        const double dx1 = (P0.x * mo + P1.x * md) + P2.x * mo, dy1 = (P0.y * mo + P1.y * md) + P2.y * mo;
        const double dx2 = (P0.x * mo + P1.x * mo) + P2.x * md, dy2 = (P0.y * mo + P1.y * mo) + P2.y * md;
        const double hp0 = ((d00 * dx0 + d01 * dy0) + w) - P0.z;
        const double hp1 = ((d00 * dx1 + d01 * dy1) + w) - P1.z;
        const double hp2 = ((d00 * dx2 + d01 * dy2) + w) - P2.z;
        const bool positive = is_wet(hp0) && is_wet(hp1) && is_wet(hp2);
        if (RECON == 1) { d10 = d11 = d20 = d21 = 0.; }
        else {
            lu2_solve(lu, N01, N11, N21, d10, d11);
            lu2_solve(lu, N02, N12, N22, d20, d21);
        }
        if (!positive) { count_branch<TAPS>(s, BR_FW_VERTEX_ZERO); d00 = d01 = d10 = d11 = d20 = d21 = 0.; }
        // on/off TVD limiter (:74-81): lo <= v_e <= hi with lo/hi = min/max of the two cell means
        const double ex0 = mx0 - cx, ey0 = my0 - cy, ex1 = mx1 - cx, ey1 = my1 - cy, ex2 = mx2 - cx, ey2 = my2 - cy;
        double t0 = 1., t1 = 1., t2 = 1.;
#define SWE_TVD(T, OWN, NB, DA, DB, EX, EY)                                             \
        {                                                                               \
            const double vek = (OWN) + ((DA) * (EX) + (DB) * (EY));                     \
            /* min(own,nb) <= vek <= max(own,nb) without forming min / max */           \
            const bool inb = ((OWN) <= (NB)) ? (((OWN) <= vek) && (vek <= (NB)))        \
                                             : (((NB) <= vek) && (vek <= (OWN)));       \
            if (!inb) T = 0.;                                                           \
        }
        SWE_TVD(t0, w, N00, d00, d01, ex0, ey0) SWE_TVD(t1, u, N01, d10, d11, ex0, ey0) SWE_TVD(t2, v, N02, d20, d21, ex0, ey0)
        SWE_TVD(t0, w, N10, d00, d01, ex1, ey1) SWE_TVD(t1, u, N11, d10, d11, ex1, ey1) SWE_TVD(t2, v, N12, d20, d21, ex1, ey1)
        SWE_TVD(t0, w, N20, d00, d01, ex2, ey2) SWE_TVD(t1, u, N21, d10, d11, ex2, ey2) SWE_TVD(t2, v, N22, d20, d21, ex2, ey2)
#undef SWE_TVD
        if (TAPS && (t0 == 0. || t1 == 0. || t2 == 0.)) count_branch<TAPS>(s, BR_FW_TVD_OFF);
        M.g00 = t0 * d00; M.g01 = t0 * d01;
        M.g10 = t1 * d10; M.g11 = t1 * d11;
        M.g20 = t2 * d20; M.g21 = t2 * d21;
    }
    st_once(s.cgx + i, M.g00); st_once(s.cgy + i, M.g01);
    emit_edge<TAPS>(s, i, M, cx, cy, mx0, my0, mb0);
    emit_edge<TAPS>(s, nt + i, M, cx, cy, mx1, my1, mb1);
    emit_edge<TAPS>(s, 2 * nt + i, M, cx, cy, mx2, my2, mb2);
    return true;
}

// generic reconstruction of the cells the fast path skipped (compacted list, O(front length))
template <bool TAPS>
__global__ void __launch_bounds__(kBlock) k_reconstruct_slow(DevMesh m, DevFields s) {
    const int nt = m.nt;
    const int n = min(s.flags[4], nt);
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += gridDim.x * blockDim.x) {
        const int i = s.rs_list[q];
        reconstruct_cell<TAPS>(m, s, i, m.tp[i], m.tp[nt + i], m.tp[2 * nt + i], m.tt[i], m.tt[nt + i], m.tt[2 * nt + i]);
    }
}

// persistent grid-stride kernel over the cell range [first, last); the ids of the thread's next
// cell are fetched before the current cell is processed (hides the first memory round trip)
// DRY: the instantiation that also maintains the dry-tile flags (DevFields::tile_dry); used while a sizeable part of the
// domain is dry (the flag logic costs ~3 % here and 5-8 % in K2 / K4 on a fully wet mesh, measured)
template <bool TAPS, int RECON, bool DRY = false>
__global__ void __launch_bounds__(kK1Block, SWE_K1_MIN_BLOCKS) k_reconstruct(DevMesh m, DevFields s, int first, int last) {
    const int nt = m.nt;
    const int stride = gridDim.x * blockDim.x;
    int i = first + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= last) return;
    int ip0 = __ldg(m.tp + i), ip1 = __ldg(m.tp + nt + i), ip2 = __ldg(m.tp + 2 * nt + i);
    int it0 = __ldg(m.tt + i), it1 = __ldg(m.tt + nt + i), it2 = __ldg(m.tt + 2 * nt + i);
    for (;;) {
        const int nx = i + stride;
        int np0 = 0, np1 = 0, np2 = 0, nt0 = 0, nt1 = 0, nt2 = 0;
        if (nx < last) {
            np0 = __ldg(m.tp + nx); np1 = __ldg(m.tp + nt + nx); np2 = __ldg(m.tp + 2 * nt + nx);
            nt0 = __ldg(m.tt + nx); nt1 = __ldg(m.tt + nt + nx); nt2 = __ldg(m.tt + 2 * nt + nx);
        }
#if SWE_K1_SPLIT
        bool done = false;
        if (DRY && __ldg(s.tdp + (i >> kUpdTileShift))) {
            // The tile was deep dry in the previous pass, so this cell most likely still is: decide that from the cell's
            // own state and its neighbours' depths alone (56 instead of 96 bytes per cell, no node / geometry packets).
            // Then nothing has to be computed or stored (DevFields::tdp) and the tile flag keeps its preset.
            const double cb = __ldg(m.cb + i);
            const double w = __ldg(s.w + i), u = __ldg(s.u + i), v = __ldg(s.v + i);
            const int j0 = max(it0, 0), j1 = max(it1, 0), j2 = max(it2, 0);
            const double h0 = __ldg(s.w + j0) - __ldg(m.cb + j0), h1 = __ldg(s.w + j1) - __ldg(m.cb + j1),
                         h2 = __ldg(s.w + j2) - __ldg(m.cb + j2);
            done = __double_as_longlong(w) == __double_as_longlong(cb) && __double_as_longlong(u) == 0ll &&
                   __double_as_longlong(v) == 0ll && (it0 < 0 || !is_wet(h0)) && (it1 < 0 || !is_wet(h1)) && (it2 < 0 || !is_wet(h2));
        }
        if (!done && !reconstruct_cell_fast<TAPS, RECON, DRY>(m, s, i, ip0, ip1, ip2, it0, it1, it2)) s.rs_list[atomicAdd(&s.flags[4], 1)] = i;
#else
        reconstruct_cell<TAPS>(m, s, i, ip0, ip1, ip2, it0, it1, it2);
#endif
        if (nx >= last) break;
        i = nx; ip0 = np0; ip1 = np1; ip2 = np2; it0 = nt0; it1 = nt1; it2 = nt2;
    }
}

// ---------------------------------------------------------------------------------------
// K1, tiled form (option k1_tiled = 1; compiled in with SWE_K1_TILED, NOT the default: measured slower than the gather
// kernel, profiles/r2_k1_forms_ncu.csv): the block's cell patch is staged in shared memory.
// Cells are numbered along a Hilbert curve, so a tile of kTile consecutive cells is a compact patch
// and ~94-96 % of its neighbour references fall inside the tile itself (measured on the structured
// and the Gmsh meshes). Per tile one thread issues 1-D TMA bulk copies (cp.async.bulk -> mbarrier
// complete_tx) of the tile's CONTIGUOUS ranges of tp/tt ids, w/u/v and the cgeo packets into a
// double-buffered stage; the copies of tile n+1 fly while tile n is computed. Neighbour data is then
// read from shared memory (LDS) when the neighbour lies in the tile and gathered from global memory
// otherwise; only the node packets are always gathered (they are shared by ~6 cells and hit L1).
// Partial tiles at the ends of a cell range use plain guarded loads into the same stage.
// ---------------------------------------------------------------------------------------
#ifndef SWE_K1_TILED
#define SWE_K1_TILED 1
#endif
#ifndef SWE_K1_TILE
#define SWE_K1_TILE 128
#endif
constexpr int kTile = SWE_K1_TILE;
struct __align__(128) K1Stage {
    double4 cgeo[kTile];
    double w[kTile], u[kTile], v[kTile];
    int tp[3][kTile], tt[3][kTile];
};
constexpr unsigned kK1StageBytes = kTile * (32 + 24 + 24);

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned phase) {
    unsigned ok;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok)
                     : "r"(smem_u32(bar)), "r"(phase)
                     : "memory");
    } while (!ok);
}

// issue the copies of tile [base, base + kTile) (must be a full, aligned tile) into stage st
__device__ __forceinline__ void k1_stage_issue(const DevMesh &m, const DevFields &s, K1Stage &st, unsigned long long *bar, int base) {
    const int nt = m.nt;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // earlier generic reads of this stage are done
    mbar_expect_tx(bar, kK1StageBytes);
    tma_load_1d(st.cgeo, m.cgeo + base, kTile * 32, bar);
    tma_load_1d(st.w, s.w + base, kTile * 8, bar);
    tma_load_1d(st.u, s.u + base, kTile * 8, bar);
    tma_load_1d(st.v, s.v + base, kTile * 8, bar);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        tma_load_1d(st.tp[k], m.tp + (size_t)k * nt + base, kTile * 4, bar);
        tma_load_1d(st.tt[k], m.tt + (size_t)k * nt + base, kTile * 4, bar);
    }
}

template <bool TAPS, int RECON>
__global__ void __launch_bounds__(kTile, SWE_K1_MIN_BLOCKS) k_reconstruct_tiled(DevMesh m, DevFields s, int first, int last) {
    __shared__ K1Stage stage[2];
    __shared__ unsigned long long bar[2];
    const int nt = m.nt;
    const int tid = threadIdx.x;
    const int tile0 = first / kTile, tile1 = (last + kTile - 1) / kTile;  // tiles on absolute kTile boundaries
    if (tid == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    // a tile can be fetched by TMA when it lies entirely inside [first, last) and nt keeps the k-major id rows 16-byte aligned
    const bool rows_aligned = (nt & 3) == 0;
    auto full = [&](int t) { return rows_aligned && t * kTile >= first && (t + 1) * kTile <= last; };
    unsigned phase[2] = {0u, 0u};
    int t = tile0 + blockIdx.x;
    if (t < tile1 && full(t) && tid == 0) k1_stage_issue(m, s, stage[0], &bar[0], t * kTile);
    int b = 0;
    for (; t < tile1; t += gridDim.x, b ^= 1) {
        const int base = t * kTile;
        const int tn = t + gridDim.x;
        if (tn < tile1 && full(tn) && tid == 0) k1_stage_issue(m, s, stage[b ^ 1], &bar[b ^ 1], tn * kTile);
        K1Stage &st = stage[b];
        const int i = base + tid;
        const bool active = i >= first && i < last;
        if (full(t)) {
            mbar_wait(&bar[b], phase[b]);
            phase[b] ^= 1u;
        } else {  // partial tile: guarded loads
            if (active) {
                st.cgeo[tid] = ldg4(m.cgeo + i);
                st.w[tid] = s.w[i]; st.u[tid] = s.u[i]; st.v[tid] = s.v[i];
#pragma unroll
                for (int k = 0; k < 3; ++k) { st.tp[k][tid] = m.tp[(size_t)k * nt + i]; st.tt[k][tid] = m.tt[(size_t)k * nt + i]; }
            }
            __syncthreads();
        }
        if (active) {
            const int lo = max(base, first), hi = min(base + kTile, last);  // staged range of this tile
            const int ip0 = st.tp[0][tid], ip1 = st.tp[1][tid], ip2 = st.tp[2][tid];
            const int it0 = st.tt[0][tid], it1 = st.tt[1][tid], it2 = st.tt[2][tid];
            const double4 P0 = ldg4(m.node + ip0), P1 = ldg4(m.node + ip1), P2 = ldg4(m.node + ip2);
            // boundary sides point at the cell itself (values unused); in-tile neighbours come from the stage
            const int j0 = it0 < 0 ? i : it0, j1 = it1 < 0 ? i : it1, j2 = it2 < 0 ? i : it2;
            double N00, N01, N02, N10, N11, N12, N20, N21, N22;
            double4 G0, G1, G2;
#define SWE_NB(J, A, B, C, G)                                                           \
            if (J >= lo && J < hi) { const int q = J - base; A = st.w[q]; B = st.u[q]; C = st.v[q]; G = st.cgeo[q]; } \
            else { A = __ldg(s.w + J); B = __ldg(s.u + J); C = __ldg(s.v + J); G = ldg4(m.cgeo + J); }
            SWE_NB(j0, N00, N01, N02, G0) SWE_NB(j1, N10, N11, N12, G1) SWE_NB(j2, N20, N21, N22, G2)
#undef SWE_NB
            if (!fast_compute<TAPS, RECON>(s, nt, i, wall_mask<false>(it0, it1, it2), P0, P1, P2, st.cgeo[tid], st.w[tid], st.u[tid], st.v[tid],
                                           N00, N01, N02, N10, N11, N12, N20, N21, N22, G0, G1, G2))
                s.rs_list[atomicAdd(&s.flags[4], 1)] = i;
        }
        __syncthreads();  // every thread is done with stage b before it is refilled two tiles later
    }
}

// ---------------------------------------------------------------------------------------
// K1, software-pipelined form (k1_tiled = 2): every thread copies the inputs of its NEXT cell — own packet and
// state, the three node packets, the three neighbour packets and states: 26 cp.async (LDGSTS, through L1) of 8 or 16
// bytes — into its private column of a shared-memory stage while it computes the current cell, so the ~1 us memory
// round trip that the gather kernel exposes once per cell (ncu: long-scoreboard = half of all warp cycles, fp64
// pipe 43 %, DRAM 62 %) is hidden behind ~550 instructions of arithmetic. The stage is single-buffered: a cell's
// values are moved to registers first, then the copies of the next cell are issued into the same slots (same thread,
// program order). 296 bytes per thread = 37 KB per 128-thread block, structure-of-arrays so that both the 8- and
// the 16-byte accesses of a warp are bank-conflict free. No block barrier: a thread only ever touches its own slots.
// ---------------------------------------------------------------------------------------
struct K1Pf {
    double2 pxy[3][kK1Block];   // node packets (x, y)
    double2 ga[4][kK1Block];    // cgeo (cx, cy): neighbours 0-2, own cell 3
    double2 gb[4][kK1Block];    // cgeo (cb, bfull)
    double pz[3][kK1Block];     // node bed
    double st[12][kK1Block];    // w,u,v of neighbours 0-2 (3k + c), own cell 9-11
};
__device__ __forceinline__ void cp_async16(void *dst, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(void *dst, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void k1pf_issue(const DevMesh &m, const DevFields &s, K1Pf &st, const int tid, const int i, const int ip0,
                                           const int ip1, const int ip2, const int it0, const int it1, const int it2) {
    const int ip[3] = {ip0, ip1, ip2};
    const int jt[4] = {max(it0, 0), max(it1, 0), max(it2, 0), i};  // boundary sides: clamped, values never used
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const double *np = reinterpret_cast<const double *>(m.node + ip[k]);
        cp_async16(&st.pxy[k][tid], np);
        cp_async8(&st.pz[k][tid], np + 2);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const double *gp = reinterpret_cast<const double *>(m.cgeo + jt[k]);
        cp_async16(&st.ga[k][tid], gp);
        cp_async16(&st.gb[k][tid], gp + 2);
        cp_async8(&st.st[3 * k][tid], s.w + jt[k]);
        cp_async8(&st.st[3 * k + 1][tid], s.u + jt[k]);
        cp_async8(&st.st[3 * k + 2][tid], s.v + jt[k]);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}
template <bool TAPS, int RECON>
__global__ void __launch_bounds__(kK1Block, SWE_K1_MIN_BLOCKS) k_reconstruct_pf(DevMesh m, DevFields s, int first, int last) {
    __shared__ K1Pf st;
    const int nt = m.nt;
    const int tid = threadIdx.x;
    const int stride = gridDim.x * blockDim.x;
    int i = first + blockIdx.x * blockDim.x + tid;
    if (i >= last) return;
    int it0, it1, it2;
    {
        const int ip0 = __ldg(m.tp + i), ip1 = __ldg(m.tp + nt + i), ip2 = __ldg(m.tp + 2 * nt + i);
        it0 = __ldg(m.tt + i); it1 = __ldg(m.tt + nt + i); it2 = __ldg(m.tt + 2 * nt + i);
        k1pf_issue(m, s, st, tid, i, ip0, ip1, ip2, it0, it1, it2);
    }
    // ids of the next cell (consumed when its copies are issued, one iteration from now)
    int nx = i + stride;
    int np0 = 0, np1 = 0, np2 = 0, nt0 = 0, nt1 = 0, nt2 = 0;
    if (nx < last) {
        np0 = __ldg(m.tp + nx); np1 = __ldg(m.tp + nt + nx); np2 = __ldg(m.tp + 2 * nt + nx);
        nt0 = __ldg(m.tt + nx); nt1 = __ldg(m.tt + nt + nx); nt2 = __ldg(m.tt + 2 * nt + nx);
    }
    for (;;) {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        double2 a, b;
        a = st.pxy[0][tid]; const double4 P0 = make_double4(a.x, a.y, st.pz[0][tid], 0.);
        a = st.pxy[1][tid]; const double4 P1 = make_double4(a.x, a.y, st.pz[1][tid], 0.);
        a = st.pxy[2][tid]; const double4 P2 = make_double4(a.x, a.y, st.pz[2][tid], 0.);
        a = st.ga[0][tid]; b = st.gb[0][tid]; const double4 G0 = make_double4(a.x, a.y, b.x, b.y);
        a = st.ga[1][tid]; b = st.gb[1][tid]; const double4 G1 = make_double4(a.x, a.y, b.x, b.y);
        a = st.ga[2][tid]; b = st.gb[2][tid]; const double4 G2 = make_double4(a.x, a.y, b.x, b.y);
        a = st.ga[3][tid]; b = st.gb[3][tid]; const double4 Gi = make_double4(a.x, a.y, b.x, b.y);
        const double N00 = st.st[0][tid], N01 = st.st[1][tid], N02 = st.st[2][tid];
        const double N10 = st.st[3][tid], N11 = st.st[4][tid], N12 = st.st[5][tid];
        const double N20 = st.st[6][tid], N21 = st.st[7][tid], N22 = st.st[8][tid];
        const double w = st.st[9][tid], u = st.st[10][tid], v = st.st[11][tid];
        const int bnd = wall_mask<false>(it0, it1, it2);
        const bool more = nx < last;
        if (more) {  // the slots are free again: start the copies of the next cell, then fetch the ids of the one after
            k1pf_issue(m, s, st, tid, nx, np0, np1, np2, nt0, nt1, nt2);
            it0 = nt0; it1 = nt1; it2 = nt2;
            const int n2 = nx + stride;
            if (n2 < last) {
                np0 = __ldg(m.tp + n2); np1 = __ldg(m.tp + nt + n2); np2 = __ldg(m.tp + 2 * nt + n2);
                nt0 = __ldg(m.tt + n2); nt1 = __ldg(m.tt + nt + n2); nt2 = __ldg(m.tt + 2 * nt + n2);
            }
        }
        if (!fast_compute<TAPS, RECON>(s, nt, i, bnd, P0, P1, P2, Gi, w, u, v, N00, N01, N02, N10, N11, N12, N20, N21, N22, G0, G1, G2))
            s.rs_list[atomicAdd(&s.flags[4], 1)] = i;
        if (!more) break;
        i = nx; nx = i + stride;
    }
}

// Value at node Q = (qx, qy, qz) of cell t's PASS-1 reconstruction, i.e. one term of the max in
// m_max_wp[node] = max(m_max_wp[node], muscl.AtPoint(P(node))[0]) (src/SpaceDisc.cpp:23).
__device__ __noinline__ double pass1_w_at_node(const DevMesh &m, const DevFields &s, int t, double qx, double qy, double qz) {
    const double4 G = m.cgeo[t];
    const int c = s.cls[t];
    double o0, g0, g1;
    if (c == 1) {  // PartWet1: flat level stored by pass 1, zero gradient (its cgx/cgy may already hold pass-2 values)
        o0 = s.pwl[t];
        g0 = 0.; g1 = 0.;
    } else {
        o0 = (c == 0) ? G.z : s.w[t];
        g0 = s.cgx[t]; g1 = s.cgy[t];
    }
    double a0 = o0 + (g0 * (qx - G.x) + g1 * (qy - G.y));
    if (!((a0 - qz) >= 0)) a0 = qz;
    return a0;
}
__device__ __forceinline__ double node_max_w(const DevMesh &m, const DevFields &s, int p, double qx, double qy, double qz) {
    double mx = qz;  // m_max_wp starts from the node bathymetry (src/SpaceDisc.cpp:34)
    for (int q = m.n2c_start[p]; q < m.n2c_start[p + 1]; ++q) mx = smax(mx, pass1_w_at_node(m, s, m.n2c_cells[q], qx, qy, qz));
    return mx;
}

// ---------------------------------------------------------------------------------------
// K1b: pass 2, ReconstructPartWetCell2 (src/MUSCLObject.cpp:114-191, S3) over the compacted list
// of part-wet cells. The node maximum it needs is gathered from the PASS-1 reconstructions of
// the cells around the lowest vertex (S8: pass 2 never sees pass-2 values) — a deterministic
// gather instead of the reference's scatter-max, no atomics on doubles.
// ---------------------------------------------------------------------------------------
constexpr int kPw2Blocks = 296;
template <bool TAPS>
__global__ void __launch_bounds__(kBlock) k_partwet2(DevMesh m, DevFields s) {
    const int nt = m.nt;
    const int npw = min(s.flags[1], nt);
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < npw; q += gridDim.x * blockDim.x) {
        const int i = s.pw_list[q];
        const int ip0 = m.tp[i], ip1 = m.tp[nt + i], ip2 = m.tp[2 * nt + i];
        const double4 P0 = m.node[ip0], P1 = m.node[ip1], P2 = m.node[ip2];
        const double4 Gi = m.cgeo[i];
        const double cx = Gi.x, cy = Gi.y, cb = Gi.z;
        const double w = s.w[i], u = s.u[i], v = s.v[i];
        // three conditional swaps (:119-121)
        double4 Q0 = P0, Q1 = P1, Q2 = P2;
        int q0 = ip0, q1 = ip1, q2 = ip2;
        if (Q0.z > Q1.z) { double4 t = Q0; Q0 = Q1; Q1 = t; int ti = q0; q0 = q1; q1 = ti; }
        if (Q1.z > Q2.z) { double4 t = Q1; Q1 = Q2; Q2 = t; int ti = q1; q1 = q2; q2 = ti; }
        if (Q0.z > Q1.z) { double4 t = Q0; Q0 = Q1; Q1 = t; int ti = q0; q0 = q1; q1 = ti; }
        const double b23 = Q0.z, b12 = Q1.z, b13 = Q2.z;
        if ((w > b13) || (b13 - b23 < kTol)) { count_branch<TAPS>(s, BR_PW2_TO_PW1); continue; }  // PartWet1 = what pass 1 wrote

        const double w23 = node_max_w(m, s, q0, Q0.x, Q0.y, Q0.z);
        const double h23 = w23 - b23;
        const double ratio_b = (b12 - b23) / (b13 - b23);
        const double h_delimiter1 = 1. / 3. * h23 * ratio_b;
        const double h_delimiter2 = 1. / 3. * h23 * (2. * b13 - b12 - b23) / (b13 - b23);
        const double hi = w - cb;
        const double ratio_h = hi / h23;
        // as written (s.pw2 == 1, :143,158-159,180) the scalar accesses points(0,2) / points(1,2) land on entries
        // that points.col(2) = ... overwrites afterwards: S0.z stays b23 and S1.z stays b12
        const bool asw = s.pw2 != 0;
        const double S0x = Q0.x, S0y = Q0.y, S0z = asw ? Q0.z : w23;
        double S1x, S1y, S1z, S2x, S2y, S2z;
        if (hi <= h_delimiter1) {  // one vertex wet
            count_branch<TAPS>(s, BR_PW2_ONE_WET);
            const double k2 = sqrt(3. * ratio_h / ratio_b);
            S1x = k2 * Q1.x + (1. - k2) * Q0.x; S1y = k2 * Q1.y + (1. - k2) * Q0.y; S1z = k2 * Q1.z + (1. - k2) * Q0.z;
            const double k3 = sqrt(3. * ratio_h * ratio_b);
            S2x = k3 * Q2.x + (1. - k3) * Q0.x; S2y = k3 * Q2.y + (1. - k3) * Q0.y; S2z = k3 * Q2.z + (1. - k3) * Q0.z;
        } else if (hi >= h_delimiter2) {  // three vertices wet
            count_branch<TAPS>(s, BR_PW2_THREE_WET);
            const double delta_w = 1.5 * (hi - h_delimiter2);
            S1x = Q1.x; S1y = Q1.y; S1z = Q1.z;
            if (!asw) {
                S1z += delta_w;
                S1z += (1. - ratio_b) * h23;
            }
            S2x = Q2.x; S2y = Q2.y; S2z = Q2.z;
            S2z += delta_w;
        } else {  // two vertices wet
            const double alpha = 3. * ratio_h;
            const double beta = (b13 - b12) / (b13 - b23);
            CubicPoly p;
            p.d = (1. + beta - alpha) / (beta * beta); p.c = (alpha - 3.) / beta; p.b = 0.;
            const double k1 = 1. - bisection(p, 0., 1.);
            if (k1 < kTol) { count_branch<TAPS>(s, BR_PW2_FALLBACK); continue; }
            count_branch<TAPS>(s, BR_PW2_TWO_WET);
            const double k3 = 1. - beta * (1. - k1);
            const double bp1 = k1 * b13 + (1. - k1) * b12;
            S1x = Q1.x; S1y = Q1.y;
            S1z = asw ? Q1.z : b12 + (k1 / k3) * beta * h23;
            S2x = k1 * Q2.x + (1. - k1) * Q1.x;
            S2y = k1 * Q2.y + (1. - k1) * Q1.y;
            S2z = bp1;
        }
        Muscl M;
        M.g10 = M.g11 = M.g20 = M.g21 = 0.;
        gradient3(S0x, S0y, S0z, S1x, S1y, S1z, S2x, S2y, S2z, M.g00, M.g01);
        M.o0 = w23 + (M.g00 * (cx - Q0.x) + M.g01 * (cy - Q0.y));
        M.o1 = u; M.o2 = v;
        s.cgx[i] = M.g00; s.cgy[i] = M.g01;
        emit_edge<TAPS>(s, i, M, cx, cy, 0.5 * (P0.x + P1.x), 0.5 * (P0.y + P1.y), 0.5 * (P0.z + P1.z));
        emit_edge<TAPS>(s, nt + i, M, cx, cy, 0.5 * (P1.x + P2.x), 0.5 * (P1.y + P2.y), 0.5 * (P1.z + P2.z));
        emit_edge<TAPS>(s, 2 * nt + i, M, cx, cy, 0.5 * (P2.x + P0.x), 0.5 * (P2.y + P0.y), 0.5 * (P2.z + P0.z));
    }
}

// tap: m_max_wp of every node after pass 1 (src/SpaceDisc.cpp:23,34)
__global__ void k_node_maxw_out(DevMesh m, DevFields s, const int *old, double *dst) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= m.nn) return;
    const double4 Q = m.node[p];
    dst[old ? old[p] : p] = node_max_w(m, s, p, Q.x, Q.y, Q.z);
}

// ---------------------------------------------------------------------------------------
// K2: edge fluxes (src/SpaceDisc.cpp:54-74) + CFL min as warp-shuffle + block reduction
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double t = __shfl_xor_sync(0xffffffffu, v, o);
        v = (t < v) ? t : v;
    }
    return v;
}

// cell of a cell-major slot (slot = k * nt + i, k in 0..2) and the dry-tile test of an edge (see DevFields::tile_dry)
__device__ __forceinline__ int slot_cell(int slot, int nt) {
    int i = slot;
    if (i >= nt) i -= nt;
    if (i >= nt) i -= nt;
    return i;
}
// 0: compute the edge; 1: a cell of a deep-dry tile on either side, the flux is +0; 2: ... and that tile was already
// flagged in the previous pass, the flux arrays hold +0 for this edge
__device__ __forceinline__ int edge_in_dry_tile(const unsigned char *td, const unsigned char *tdf, int nt, int sl, int sr) {
    const int tl = slot_cell(sl, nt) >> kUpdTileShift;
    unsigned f = __ldg(td + tl);
    unsigned g = f & __ldg(tdf + tl);
    if (sr >= 0) {
        const int tr = slot_cell(sr, nt) >> kUpdTileShift;
        const unsigned fr = __ldg(td + tr);
        f |= fr;
        g |= fr & __ldg(tdf + tr);
    }
    return g ? 2 : (f ? 1 : 0);
}
#ifndef SWE_K2_GRID_PER_SM
#define SWE_K2_GRID_PER_SM 64  // A/B at 64M cells: 16 -> 2.114 ms, 64 -> 2.062 ms, one block per 128 edges -> 2.889 ms
#endif
#ifndef SWE_K2_MIN_BLOCKS
#define SWE_K2_MIN_BLOCKS 8
#endif
// CFL = false: the instantiation for the non-final stages of a multi-stage step. m_min_length_to_wavespeed is reset
// and rebuilt by EVERY ComputeFluxes (src/SpaceDisc.cpp:56) and only read after the step (CFLdt, include/TimeDisc.h:13),
// so the minima of the earlier stages are dead values: no dmin load, no division, no reduction, no publish.
template <class FLUXER, bool OPT, bool CFL = true, bool DRY = false>
__global__ void __launch_bounds__(kBlock, SWE_K2_MIN_BLOCKS) k_flux(DevMesh m, DevFields s, double abscor, int roe_fix, int cfl_abs) {
    const int ne = m.ne;
    const int stride = gridDim.x * blockDim.x;
    double l2w = 1.0;  // reset value of m_min_length_to_wavespeed (:56)
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (!DRY) {
        if (e < ne) {
            int sl = __ldg(m.slotL + e), sr = __ldg(m.slotR + e);
            for (;;) {
                const int nx = e + stride;
                int nsl = 0, nsr = 0;
                if (nx < ne) { nsl = __ldg(m.slotL + nx); nsr = __ldg(m.slotR + nx); }
                const double2 n = __ldg(m.en + e);
                double f0, f1, f2;
                if (sr < 0) {  // SOLID_WALL: ElemFlux(n, {h(lf), 0, 0}) with the cell-mean depth (:67-69)
                    const int lf = sl % m.nt;
                    const double h = s.w[lf] - m.cb[lf];
                    elem_flux(n.x, n.y, h, 0., 0., f0, f1, f2);
                } else {
                    double cand = 1.0;
                    FLUXER::template eval<OPT>(n.x, n.y, __ldg(s.ceh + sl), __ldg(s.ceu + sl), __ldg(s.cev + sl), __ldg(s.ceh + sr),
                                               __ldg(s.ceu + sr), __ldg(s.cev + sr), CFL ? __ldg(m.dmin + e) : 1.0, abscor, f0, f1, f2,
                                               cand, roe_fix, cfl_abs);
                    if (CFL) l2w = (cand < l2w) ? cand : l2w;  // edges excluded from the CFL min carry dmin = +inf
                }
                st_once(s.f0 + e, f0); st_once(s.f1 + e, f1); st_once(s.f2 + e, f2);
                if (nx >= ne) break;
                e = nx; sl = nsl; sr = nsr;
            }
        }
    } else if (e < ne) {
        // dry-region form: slot ids two edges ahead, dry-tile flags one edge ahead, so neither is on the critical path
        int sl = __ldg(m.slotL + e), sr = __ldg(m.slotR + e);
        int nx = e + stride;
        int nsl = 0, nsr = 0;
        if (nx < ne) { nsl = __ldg(m.slotL + nx); nsr = __ldg(m.slotR + nx); }
        int skip = edge_in_dry_tile(s.td, s.tdf, m.nt, sl, sr);
        for (;;) {
            const int n2 = nx + stride;
            int n2sl = 0, n2sr = 0;
            if (n2 < ne) { n2sl = __ldg(m.slotL + n2); n2sr = __ldg(m.slotR + n2); }
            const int nskip = (nx < ne) ? edge_in_dry_tile(s.td, s.tdf, m.nt, nsl, nsr) : 0;
            double f0, f1, f2;
            if (skip) {  // a cell of a deep-dry tile on either side: both cells are dry, the flux is exactly +0
                f0 = 0.; f1 = 0.; f2 = 0.;
            } else {
                const double2 n = __ldg(m.en + e);
                if (sr < 0) {
                    const int lf = sl % m.nt;
                    const double h = s.w[lf] - m.cb[lf];
                    elem_flux(n.x, n.y, h, 0., 0., f0, f1, f2);
                } else {
                    double cand = 1.0;
                    FLUXER::template eval<OPT>(n.x, n.y, __ldg(s.ceh + sl), __ldg(s.ceu + sl), __ldg(s.cev + sl), __ldg(s.ceh + sr),
                                               __ldg(s.ceu + sr), __ldg(s.cev + sr), CFL ? __ldg(m.dmin + e) : 1.0, abscor, f0, f1, f2,
                                               cand, roe_fix, cfl_abs);
                    if (CFL) l2w = (cand < l2w) ? cand : l2w;
                }
            }
            if (skip != 2) { st_once(s.f0 + e, f0); st_once(s.f1 + e, f1); st_once(s.f2 + e, f2); }
            if (nx >= ne) break;
            e = nx; nx = n2; sl = nsl; sr = nsr; nsl = n2sl; nsr = n2sr; skip = nskip;
        }
    }
    if (!CFL) return;
    __shared__ double red[kBlock / 32];
    l2w = warp_min(l2w);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = l2w;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = (threadIdx.x < kBlock / 32) ? red[threadIdx.x] : 1.0;
        v = warp_min(v);
        if (threadIdx.x == 0) {
            // running min in scal[3]; the last block to finish publishes it as
            // m_min_length_to_wavespeed (scal[0]) and re-arms the accumulator with the reset
            // value 1.0 (src/SpaceDisc.cpp:56), so no separate reset / finalise kernels exist.
            if (v < 1.0) atomic_min_pos_double(&s.scal[3], v);
            __threadfence();
            const int ticket = atomicAdd(&s.flags[3], 1);
            if (ticket == (int)gridDim.x - 1) {
                const double mn = __longlong_as_double(atomicAdd((unsigned long long *)&s.scal[3], 0ull));
                s.scal[0] = mn;
                *(volatile double *)&s.scal[3] = 1.0;
                s.flags[3] = 0;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// K3: draining time step (src/TimeDisc.cpp:43-66) of every cell from the stage-begin state
// ---------------------------------------------------------------------------------------
#ifndef SWE_K3_PERSISTENT
#define SWE_K3_PERSISTENT 0  // measured: 0.905 ms persistent vs 0.716 ms one cell per short-lived thread at 64M cells
#endif
#ifndef SWE_K3_ILP
#define SWE_K3_ILP 2  // A/B at 64M cells: one cell per thread 0.727 ms, two 0.615 ms
#endif
#ifndef SWE_K3_GRID_PER_SM
#define SWE_K3_GRID_PER_SM 16
#endif
__device__ __forceinline__ double drain_dt_cell(double h, double area, double fe0, double fe1, double fe2);
template <bool DRY = false>
__global__ void __launch_bounds__(kBlock) k_drain(DevMesh m, DevFields s) {
    const int nt = m.nt;
#if SWE_K3_PERSISTENT
    // persistent grid-stride form: the edge ids of the thread's next cell are fetched while the current cell waits
    // for its flux gathers, so only one memory round trip per cell is exposed instead of two per short-lived thread
    const int stride = gridDim.x * blockDim.x;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nt) return;
    int t0 = __ldg(m.te + i), t1 = __ldg(m.te + nt + i), t2 = __ldg(m.te + 2 * nt + i);
    for (;;) {
        const int nx = i + stride;
        int n0 = 0, n1 = 0, n2 = 0;
        if (nx < nt) { n0 = __ldg(m.te + nx); n1 = __ldg(m.te + nt + nx); n2 = __ldg(m.te + 2 * nt + nx); }
        const double f0 = s.f0[t0 >= 0 ? t0 : ~t0], f1 = s.f0[t1 >= 0 ? t1 : ~t1], f2 = s.f0[t2 >= 0 ? t2 : ~t2];
        const double h = s.w[i] - m.cb[i];
        st_once(s.dti + i, drain_dt_cell(h, m.area[i], t0 >= 0 ? f0 : -f0, t1 >= 0 ? f1 : -f1, t2 >= 0 ? f2 : -f2));
        if (nx >= nt) break;
        i = nx; t0 = n0; t1 = n1; t2 = n2;
    }
#elif SWE_K3_ILP == 2
    // two cells per thread (tiles 2b and 2b + 1 of the block): both cells' edge ids, then both cells' gathers are in
    // flight together, which halves the number of exposed memory round trips per cell
    const int i0 = (2 * blockIdx.x) * blockDim.x + threadIdx.x, i1 = i0 + blockDim.x;
    if (i0 >= nt) return;
    const bool two = i1 < nt;
    const int j1 = two ? i1 : i0;
    if (DRY) {
        // dry-region form: the tile flags are tested before anything else is loaded (like the stage update: a wet tile pays
        // one extra memory round trip, a deep-dry tile costs two bytes per thread)
        const int ta = i0 >> kUpdTileShift, tb = j1 >> kUpdTileShift;
        if (__ldg(s.td + ta) && __ldg(s.td + tb)) {  // dry cells: draining dt = 0 (already stored if the tiles were flagged before)
            if (!(__ldg(s.tdd + ta) && __ldg(s.tdd + tb))) {
                st_once(s.dti + i0, 0.);
                if (two) st_once(s.dti + i1, 0.);
            }
            return;
        }
    }
    const int a0 = __ldg(m.te + i0), a1 = __ldg(m.te + nt + i0), a2 = __ldg(m.te + 2 * nt + i0);
    const int b0 = __ldg(m.te + j1), b1 = __ldg(m.te + nt + j1), b2 = __ldg(m.te + 2 * nt + j1);
    const double ha = s.w[i0] - m.cb[i0], hb = s.w[j1] - m.cb[j1];
    const double aa = m.area[i0], ab = m.area[j1];
    const double fa0 = s.f0[a0 >= 0 ? a0 : ~a0], fa1 = s.f0[a1 >= 0 ? a1 : ~a1], fa2 = s.f0[a2 >= 0 ? a2 : ~a2];
    const double fb0 = s.f0[b0 >= 0 ? b0 : ~b0], fb1 = s.f0[b1 >= 0 ? b1 : ~b1], fb2 = s.f0[b2 >= 0 ? b2 : ~b2];
    st_once(s.dti + i0, drain_dt_cell(ha, aa, a0 >= 0 ? fa0 : -fa0, a1 >= 0 ? fa1 : -fa1, a2 >= 0 ? fa2 : -fa2));
    if (two) st_once(s.dti + i1, drain_dt_cell(hb, ab, b0 >= 0 ? fb0 : -fb0, b1 >= 0 ? fb1 : -fb1, b2 >= 0 ? fb2 : -fb2));
#else
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nt) return;
    const double h = s.w[i] - m.cb[i];
    double r;
    if (!is_wet(h)) {
        r = 0.;
    } else {
        double sum = 0.;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int te = m.te[k * nt + i];
            const double fe = (te >= 0) ? s.f0[te] : -s.f0[~te];
            sum += smax(0., fe);
        }
        r = (sum > kTol) ? m.area[i] * h / sum : __longlong_as_double(0x7ff0000000000000ll);
    }
    st_once(s.dti + i, r);
#endif
}

// K3', the list-driven form used with the fused stage update: only the cells whose draining dt is read from
// global memory by some neighbour — cells with a neighbour in another 128-cell update tile or in another
// ordering class (a static list built at set-up, a few per cent of the mesh). Same arithmetic as k_drain.
__device__ __forceinline__ double drain_dt_cell(double h, double area, double fe0, double fe1, double fe2) {
    if (!is_wet(h)) return 0.;
    double sum = 0.;
    sum += smax(0., fe0); sum += smax(0., fe1); sum += smax(0., fe2);
    return (sum > kTol) ? area * h / sum : __longlong_as_double(0x7ff0000000000000ll);
}
__global__ void __launch_bounds__(kBlock) k_drain_list(DevMesh m, DevFields s, const int *__restrict__ list, int n) {
    const int nt = m.nt;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += gridDim.x * blockDim.x) {
        const int i = __ldg(list + q);
        const int t0 = __ldg(m.te + i), t1 = __ldg(m.te + nt + i), t2 = __ldg(m.te + 2 * nt + i);
        const double fe0 = (t0 >= 0) ? s.f0[t0] : -s.f0[~t0];
        const double fe1 = (t1 >= 0) ? s.f0[t1] : -s.f0[~t1];
        const double fe2 = (t2 >= 0) ? s.f0[t2] : -s.f0[~t2];
        s.dti[i] = drain_dt_cell(s.w[i] - m.cb[i], m.area[i], fe0, fe1, fe2);
    }
}
// set-up: flag the cells of the K3' list. cf[q] = first device id of ordering class q (classes are contiguous).
struct ClassFirst { int f[6]; };
__device__ __forceinline__ int class_of(const ClassFirst &cf, int i) {
    return (i >= cf.f[1]) + (i >= cf.f[2]) + (i >= cf.f[3]) + (i >= cf.f[4]);
}
__global__ void k_mark_drain_boundary(int nt, const int *tt, ClassFirst cf, unsigned char *flag) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nt) return;
    const int ci = class_of(cf, i), ti = i >> kUpdTileShift;
    bool b = false;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int j = tt[k * nt + i];
        if (j >= 0 && ((j >> kUpdTileShift) != ti || class_of(cf, j) != ci)) b = true;
    }
    flag[i] = b ? 1 : 0;
}

// dry-region helpers of the stage update: the compacted list of tiles to process, and the fill of the skipped tiles
// when the stage writes the other state buffer (their cells are (cb, +0, +0) by the definition of the flag)
__global__ void k_tile_compact(int ntiles, const unsigned char *td, const unsigned char *td0, int use_td0, int *list) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const bool todo = t < ntiles && !(td[t] && (!use_td0 || td0[t]));
    const unsigned bal = __ballot_sync(0xffffffffu, todo);
    if (!bal) return;
    const int lane = threadIdx.x & 31;
    int pos = 0;
    if (lane == 0) pos = atomicAdd(list, __popc(bal));
    pos = __shfl_sync(0xffffffffu, pos, 0);
    if (todo) list[1 + pos + __popc(bal & ((1u << lane) - 1u))] = t;
}
__global__ void k_fill_dry_tiles(int first, int last, const unsigned char *td, const unsigned char *td0, int use_td0,
                                 const double *cb, double *wout, double *uout, double *vout) {
    const int i = first + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= last) return;
    const int t = i >> kUpdTileShift;
    if (td[t] && (!use_td0 || td0[t])) { wout[i] = cb[i]; uout[i] = 0.; vout[i] = 0.; }
}

// ---------------------------------------------------------------------------------------
// K4: RHS gather (src/TimeDisc.cpp:3-41) + RK combination (src/Solvers.cpp) + ConsAssigner
// (src/Assigners.cpp:22-44). Deterministic: fixed k order, no float atomics.
//   PLAIN: U = cons(W) + RHS                 (Euler `+=`, first RK stage)
//   else : U = a0*cons(U0) + a1*cons(W) + RHS
// Reads W (stage input) and writes Wout (may alias W: only the own cell is read).
// ---------------------------------------------------------------------------------------
#ifndef SWE_K4_MIN_BLOCKS
#define SWE_K4_MIN_BLOCKS 1
#endif
// RHS_ONLY (tap for TimeDisc::RHS(i, dt)): store the increment (r0, r1, r2) instead of applying it.
// FUSED: the draining dt (K3, src/TimeDisc.cpp:43-66) of the block's own cells is computed here from the already
// loaded mass fluxes and shared through shared memory; blocks work on ABSOLUTE 128-cell tiles, so a neighbour in the
// same tile (and inside [first, last)) is served from shared memory and only the few neighbours outside read the
// global dti array, which k_drain_list filled for exactly those cells. Saves the k_drain pass over all cells and
// three 8-byte gathers per cell.
template <bool PLAIN, bool COR, bool RHS_ONLY = false, bool FUSED = false, bool DRY = false>
__global__ void __launch_bounds__(kBlock, SWE_K4_MIN_BLOCKS) k_update(DevMesh m, DevFields s, const double *__restrict__ w0,
                                                   const double *__restrict__ u0, const double *__restrict__ v0,
                                                   double *wout, double *uout, double *vout, double a0, double a1,
                                                   double dt_host, double dt_coef, double cor, int first, int last) {
    // cell range [first, last) in device numbering; FUSED: tiles on absolute kBlock boundaries
    // dry-region form (DRY): block b processes the b-th tile of the compacted list of tiles that are NOT deep dry (now and,
    // when U0 enters the combination, at swe_save_state); the other tiles keep (cb, +0, +0) (k_fill_dry_tiles writes
    // that into the other buffer when the stage is out of place). Tiles are absolute here, lanes outside [first, last) idle.
    // On small meshes (tile_list == nullptr; the extra launches would cost more than they save) every block tests its own flag.
    const bool listed = DRY && s.tile_list != nullptr;
    if (listed && (int)blockIdx.x >= __ldg(s.tile_list)) return;
    const int base = listed ? __ldg(s.tile_list + 1 + blockIdx.x) << kUpdTileShift
                            : (FUSED ? ((first >> kUpdTileShift) + (int)blockIdx.x) << kUpdTileShift : first + blockIdx.x * blockDim.x);
    int i = base + threadIdx.x;
    const int nt = m.nt;
    __shared__ double sdt[FUSED ? kBlock : 1];
    const bool active = FUSED ? (i >= first && i < last) : true;
    if (listed) { if (i < first || i >= last) return; }
    else if (!FUSED) { if (i >= last) return; }
    else if (!active) i = first;  // idle lanes of a partial tile shadow a valid cell (loads stay in bounds), never store
    if (DRY && !listed) {
        const int t = i >> kUpdTileShift;
        unsigned skip_t = __ldg(s.td + t);
        if (!PLAIN) skip_t &= __ldg(s.td0 + t);
        if (skip_t) {
            if (wout != s.w) { wout[i] = __ldg(m.cb + i); uout[i] = 0.; vout[i] = 0.; }  // out of place (first stage after swe_save_state)
            return;
        }
    }
    // All loads are issued before any arithmetic (ids -> gathers: two dependent round trips, every
    // gather of the cell in flight at once). Written out explicitly because the compiler's own
    // schedule flipped between a batched (2.2 ms) and an interleaved (2.6 ms at 64M cells) form
    // when an unrelated kernel parameter changed.
    int te[3], tn[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) { te[k] = __ldg(m.te + k * nt + i); tn[k] = __ldg(m.tt + k * nt + i); }
    // dt_coef != 0: stage dt = dt_coef * (device-resident dt), else the host value
    const double dt = (dt_coef != 0.) ? dt_coef * s.scal[1] : dt_host;
    const double cb = __ldg(m.cb + i);
    const double area_i = __ldg(m.area + i);
    double dti = FUSED ? 0. : s.dti[i];
    const double gx = s.cgx[i], gy = s.cgy[i];
    const double wc = s.w[i], uc = s.u[i], vc = s.v[i];
    double wa = 0., ua = 0., va = 0.;
    if (!PLAIN) { wa = w0[i]; ua = u0[i]; va = v0[i]; }
    double F0[3], F1[3], F2[3], dtn[3], len[3], hek[3], cu[3], cv[3];
    double2 nrm[3];
    const int lo_t = max(base, first), hi_t = min(base + kBlock, last);  // FUSED: cells whose dti this block computes
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int e = te[k] >= 0 ? te[k] : ~te[k];
        F0[k] = s.f0[e]; F1[k] = s.f1[e]; F2[k] = s.f2[e];
        if (FUSED) {  // in-tile neighbours come from shared memory after the barrier below
            const bool in_tile = tn[k] >= lo_t && tn[k] < hi_t;
            dtn[k] = (tn[k] < 0) ? __longlong_as_double(0x7ff0000000000000ll) : (in_tile ? 0. : s.dti[tn[k]]);
        } else {
            dtn[k] = (tn[k] < 0) ? __longlong_as_double(0x7ff0000000000000ll) : s.dti[tn[k]];
        }
        len[k] = __ldg(m.elen + e);
        nrm[k] = __ldg(m.en + e);
        hek[k] = s.ceh[k * nt + i];
        if (COR) { cu[k] = s.ceu[k * nt + i]; cv[k] = s.cev[k * nt + i]; }
    }
    if (FUSED) {
        dti = drain_dt_cell(wc - cb, area_i, te[0] >= 0 ? F0[0] : -F0[0], te[1] >= 0 ? F0[1] : -F0[1], te[2] >= 0 ? F0[2] : -F0[2]);
        sdt[threadIdx.x] = dti;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 3; ++k)
            if (tn[k] >= lo_t && tn[k] < hi_t) dtn[k] = sdt[tn[k] - base];
        if (!active) return;
    }
    const double i_area = 1. / area_i;
    double r0 = 0., r1 = 0., r2 = 0.;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const double sgn = te[k] >= 0 ? 1. : -1.;
        const double dtk = (sgn * F0[k] > 0.) ? smin(dt, dti) : smin(dt, dtn[k]);
        const double c_ek = i_area * len[k];
        const double h_ek = hek[k];
        const double sc = dtk * sgn * c_ek;
        r0 -= sc * F0[k]; r1 -= sc * F1[k]; r2 -= sc * F2[k];
        // m_src = grad w + cor * (-v_e, u_e) (src/SpaceDisc.cpp:26-29); with cor == 0 the second
        // term is a signed zero and the sum equals the gradient
        double sx = gx, sy = gy;
        if (COR) { sx = gx + cor * (-cv[k]); sy = gy + cor * cu[k]; }
        r1 -= dt * (1. / 3.) * sx * h_ek;
        r2 -= dt * (1. / 3.) * sy * h_ek;
        const double nx = sgn * nrm[k].x, ny = sgn * nrm[k].y;  // Norm(e, i) = -Norm(e, other) exactly
        r1 += dtk * (nx * c_ek * (0.5 * h_ek * h_ek));
        r2 += dtk * (ny * c_ek * (0.5 * h_ek * h_ek));
    }
    if (RHS_ONLY) { wout[i] = r0; uout[i] = r1; vout[i] = r2; return; }
    const double hc = wc - cb;
    double U0, U1, U2;
    if (PLAIN) {
        U0 = hc + r0; U1 = uc * hc + r1; U2 = vc * hc + r2;
    } else {
        const double ha = wa - cb;
        const double A1 = ua * ha, A2 = va * ha;
        U0 = (a0 * ha + a1 * hc) + r0;
        U1 = (a0 * A1 + a1 * (uc * hc)) + r1;
        U2 = (a0 * A2 + a1 * (vc * hc)) + r2;
    }
    double ow, ou, ov;
    if (!is_wet(U0)) {
        ow = cb; ou = 0.; ov = 0.;
    } else {
        double ih;
        if (U0 < SWE_DAMP_DEPTH) ih = sqrt(2.0) * U0 / sqrt(U0 * U0 * U0 * U0 + SWE_DAMP_EPS_CONS);
        else ih = 1. / U0;
        ow = U0 + cb; ou = U1 * ih; ov = U2 * ih;
    }
    if (!(isfinite(ow) && isfinite(ou) && isfinite(ov))) s.flags[0] = 1;
    wout[i] = ow; uout[i] = ou; vout[i] = ov;
}

// number of dry cells of the current state (decides whether the dry-region instantiations are worth their overhead)
__global__ void k_count_dry(int nt, const double *w, const double *cb, unsigned long long *out) {
    unsigned long long n = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nt; i += gridDim.x * blockDim.x) n += is_wet(w[i] - cb[i]) ? 0 : 1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
    if ((threadIdx.x & 31) == 0 && n) atomicAdd(out, n);
}

// IsDryCell / IsFullWetCell / IsPartWetCell (src/MUSCLObject.cpp:13-29) of the CURRENT state, without reconstructing
__global__ void k_classify(DevMesh m, const double *w, signed char *cls) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m.nt) return;
    const double4 G = m.cgeo[i];  // (cx, cy, cb, bfull): bfull = max node bed, +inf on boundary triangles
    const double wi = w[i];
    cls[i] = !is_wet(wi - G.z) ? 0 : ((G.w < wi) ? 2 : 1);
}

// after a step: time += dt_used; in adaptive mode dt = 0.15 * min_len (include/TimeDisc.h:13,22)
// min_slot: where min_len_to_wavespeed of the finished step lives (0 on one GPU; 4 = the global minimum pulled from
// the peer-memory table on several GPUs)
__global__ void k_post_step(DevFields s, double dt_host, int adaptive, int min_slot) {
    const double used = adaptive ? s.scal[1] : dt_host;
    s.scal[2] += used;
    if (adaptive) s.scal[1] = SWE_CFL * s.scal[min_slot];
}
__global__ void k_set_scalar(double *p, double v) { *p = v; }

// ---------------------------------------------------------------------------------------
// setup: geometry from node coordinates with the reference's formulas (src/Bathymetry.cpp)
// ---------------------------------------------------------------------------------------
__global__ void k_setup_cells(int nt, const int *tp, const int *tt, const double4 *node, double4 *cgeo, double *area, double *cb) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nt) return;
    const double4 P0 = node[tp[i]], P1 = node[tp[nt + i]], P2 = node[tp[2 * nt + i]];
    const double third = 1. / 3.;
    double4 g;
    g.x = P0.x * third + P1.x * third + P2.x * third;  // Domain::T (:24-27)
    g.y = P0.y * third + P1.y * third + P2.y * third;
    g.z = P0.z * third + P1.z * third + P2.z * third;
    const bool bnd = (tt[i] | tt[nt + i] | tt[2 * nt + i]) < 0;
    g.w = bnd ? __longlong_as_double(0x7ff0000000000000ll) : smax(smax(P0.z, P1.z), P2.z);
    cgeo[i] = g;
    cb[i] = g.z;
    const double ax = P1.x - P0.x, ay = P1.y - P0.y, bx = P2.x - P0.x, by = P2.y - P0.y;
    area[i] = 0.5 * fabs(ax * by - bx * ay);  // Domain::Area (:87-90)
}

// ---------------------------------------------------------------------------------------
// set-up on the device: the caller's int64 incidence arrays and 3 x nn geometry are uploaded as they are; the
// locality-preserving numbering (Hilbert keys + CUB sort), the conversion to int32 structure-of-arrays in device
// numbering and the consistency checks of the local edge order run as kernels (at 64M cells the host loops that did
// this before cost 6.5 of the 10 s of set-up).
// ---------------------------------------------------------------------------------------
// Hilbert-curve index of a point on the 2^21 x 2^21 grid (no quadrant jumps, unlike the Z-order / Morton curve)
__host__ __device__ inline unsigned long long hilbert21(unsigned long long x, unsigned long long y) {
    const unsigned long long n = 1ull << 21;
    unsigned long long d = 0;
    for (unsigned long long s = n >> 1; s > 0; s >>= 1) {
        const unsigned long long rx = (x & s) ? 1 : 0, ry = (y & s) ? 1 : 0;
        d += s * s * ((3 * rx) ^ ry);
        if (ry == 0) {
            if (rx == 1) { x = n - 1 - x; y = n - 1 - y; }
            const unsigned long long t = x; x = y; y = t;
        }
    }
    return d;
}
struct KeyBox { double x0, y0, scale; };
__device__ __forceinline__ unsigned long long curve_key(double x, double y, const KeyBox &b) {
    const unsigned long long qx = (unsigned long long)fmin(2097151.0, fmax(0.0, (x - b.x0) * b.scale));
    const unsigned long long qy = (unsigned long long)fmin(2097151.0, fmax(0.0, (y - b.y0) * b.scale));
    return hilbert21(qx, qy);
}
// which: 0 cells (centroid; ids = element_nodes, 3 per row), 1 edges (midpoint; ids = edge_nodes, 2 per row), 2 nodes.
// curve == 0 keeps the caller's order (inside every class); cls (cells only): class id in the bits above the curve key.
__global__ void k_order_keys(int which, long long n, const long long *ids, const double *geom, KeyBox box, const unsigned char *cls,
                             int curve, unsigned long long *keys, int *vals) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long long key = (unsigned long long)i;
    if (curve) {
        double x, y;
        if (which == 0) {
            const long long a = ids[3 * i], b = ids[3 * i + 1], c = ids[3 * i + 2];
            x = (geom[3 * a] + geom[3 * b] + geom[3 * c]) / 3.;
            y = (geom[3 * a + 1] + geom[3 * b + 1] + geom[3 * c + 1]) / 3.;
        } else if (which == 1) {
            const long long a = ids[2 * i], b = ids[2 * i + 1];
            x = 0.5 * (geom[3 * a] + geom[3 * b]);
            y = 0.5 * (geom[3 * a + 1] + geom[3 * b + 1]);
        } else {
            x = geom[3 * i]; y = geom[3 * i + 1];
        }
        key = curve_key(x, y, box);
    }
    if (cls) key |= (unsigned long long)cls[i] << 42;
    keys[i] = key;
    vals[i] = (int)i;
}
__global__ void k_invert_perm(int n, const int *order, int *newid) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) newid[order[k]] = k;
}
// maps: caller id -> device id, nullptr = identity
__device__ __forceinline__ int map_id(const int *map, long long i) { return map ? map[i] : (int)i; }
// cell incidence -> k-major int32 SoA in device numbering; bad[0] = max code of a violated local convention
// (1 element_edges / edge_elements inconsistent, 2 edge k does not join nodes k, k+1, 3 neighbour k is not across edge k)
__global__ void k_convert_cells(int nt, const long long *tp64, const long long *te64, const long long *tt64, const long long *ep64,
                                const long long *et64, const int *cell_new, const int *edge_new, const int *node_new, int *tp,
                                int *tt, int *te, int *bad) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nt) return;
    const int d = map_id(cell_new, t);
    int worst = 0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        tp[(size_t)k * nt + d] = map_id(node_new, tp64[3 * (size_t)t + k]);
        const long long nb = tt64[3 * (size_t)t + k];
        tt[(size_t)k * nt + d] = nb >= 0 ? map_id(cell_new, nb) : (int)nb;
        const long long e = te64[3 * (size_t)t + k];
        const bool first = et64[2 * e] == t;
        if (!first && et64[2 * e + 1] != t) worst = max(worst, 1);
        const long long a = ep64[2 * e], b = ep64[2 * e + 1];
        const long long p = tp64[3 * (size_t)t + k], q = tp64[3 * (size_t)t + (k + 1) % 3];
        if (!((a == p && b == q) || (a == q && b == p))) worst = max(worst, 2);
        const long long across = first ? et64[2 * e + 1] : et64[2 * e];
        if (across != nb) worst = max(worst, 3);
        const int en = map_id(edge_new, e);
        te[(size_t)k * nt + d] = first ? en : ~en;
    }
    if (worst) atomicMax(bad, worst);
}
__global__ void k_convert_edges(int ne, const long long *ep64, const long long *et64, const int *cell_new, const int *edge_new,
                                const int *node_new, int *ep0, int *ep1, int *et0, int *et1) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= ne) return;
    const int d = map_id(edge_new, e);
    ep0[d] = map_id(node_new, ep64[2 * (size_t)e]); ep1[d] = map_id(node_new, ep64[2 * (size_t)e + 1]);
    et0[d] = map_id(cell_new, et64[2 * (size_t)e]);
    const long long b = et64[2 * (size_t)e + 1];
    et1[d] = b >= 0 ? map_id(cell_new, b) : (int)b;
}
__global__ void k_convert_nodes(int nn, const double *geom, const int *node_new, double4 *node) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nn) return;
    node[map_id(node_new, p)] = make_double4(geom[3 * (size_t)p], geom[3 * (size_t)p + 1], geom[3 * (size_t)p + 2], 0.);
}

// node -> cells CSR set-up: count the incident cells of every node and emit the cell id of each (corner, cell) pair
__global__ void k_n2c_count(int n3, int nt, const int *tp, int *count, int *cell_of_pair) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n3) return;
    atomicAdd(&count[tp[q]], 1);
    cell_of_pair[q] = q % nt;
}

__global__ void k_setup_slots(int nt, const int *te, int *slotL, int *slotR) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nt) return;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int t = te[k * nt + i];
        if (t >= 0) slotL[t] = k * nt + i; else slotR[~t] = k * nt + i;
    }
}

// ep = EdgePoints in the CALLER's order (sorted by caller node id), et0/et1 = device cell ids
__global__ void k_setup_edges(int ne, int nt, const int *ep0, const int *ep1, const int *et0, const int *et1,
                              const double4 *node, const double4 *cgeo, const double *area, double2 *en,
                              double *elen, double *dmin) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= ne) return;
    const double4 A = node[ep0[e]], B = node[ep1[e]];
    const double len = sqrt((A.x - B.x) * (A.x - B.x) + (A.y - B.y) * (A.y - B.y));  // Len (:4-6)
    double tx = (B.x - A.x) / len, ty = (B.y - A.y) / len;                           // Tang (:69-76)
    const int lf = et0[e], lt = et1[e];
    const double4 T = cgeo[lf];
    const double dx = T.x - A.x, dy = T.y - A.y;
    if (dx * ty - tx * dy > 0.) { tx = -tx; ty = -ty; }
    en[e] = make_double2(ty, -tx);  // Norm (:78-80)
    elen[e] = len;
    const double dl = 2. * area[lf] / len;
    double d = dl;
    if (lt >= 0) { const double dr = 2. * area[lt] / len; d = smin(dl, dr); }
    dmin[e] = d;
}

// ---------------------------------------------------------------------------------------
// layout conversion between the caller's Storage<3> (3 x n column-major, caller numbering)
// and the device SoA. old[] = caller id of device id (nullptr: identity).
// ---------------------------------------------------------------------------------------
__global__ void k_state_in(int nt, const int *old, const double *__restrict__ aos, double *w, double *u, double *v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nt) return;
    const size_t o = old ? (size_t)old[i] : (size_t)i;
    w[i] = aos[3 * o]; u[i] = aos[3 * o + 1]; v[i] = aos[3 * o + 2];
}
__global__ void k_state_out(int nt, const int *old, const double *w, const double *u, const double *v, double *aos) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nt) return;
    const size_t o = old ? (size_t)old[i] : (size_t)i;
    aos[3 * o] = w[i]; aos[3 * o + 1] = u[i]; aos[3 * o + 2] = v[i];
}
__global__ void k_scalar_out(int n, const int *old, const double *src, double *dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    dst[old ? old[i] : i] = src[i];
}
__global__ void k_cls_out(int n, const int *old, const signed char *src, signed char *dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    dst[old ? old[i] : i] = src[i];
}
__global__ void k_flux_out(int ne, const int *old, const double *f0, const double *f1, const double *f2, double *aos) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= ne) return;
    const size_t o = old ? (size_t)old[e] : (size_t)e;
    aos[3 * o] = f0[e]; aos[3 * o + 1] = f1[e]; aos[3 * o + 2] = f2[e];
}
// edge-side taps into the reference's EdgeField layout: column 2e + (from < to), caller ids
// (include/ValueField.h:70-75). which = 0: (w,u,v) of m_edg, 1: (0, sx, sy) of m_src.
__global__ void k_edge_out(DevMesh m, DevFields s, const int *cell_old, const int *edge_old, int which, double cor, double *aos) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int nt = m.nt;
    if (i >= nt) return;
    const long long from = cell_old ? cell_old[i] : i;
    for (int k = 0; k < 3; ++k) {
        const int te = m.te[k * nt + i];
        const int e = te >= 0 ? te : ~te;
        const int tn = m.tt[k * nt + i];
        const long long to = (tn < 0) ? (long long)tn : (cell_old ? cell_old[tn] : tn);
        const size_t eo = edge_old ? (size_t)edge_old[e] : (size_t)e;
        const size_t col = 2 * eo + (from < to ? 1 : 0);
        const int slot = k * nt + i;
        if (which == 0) {
            aos[3 * col] = s.cew[slot]; aos[3 * col + 1] = s.ceu[slot]; aos[3 * col + 2] = s.cev[slot];
        } else {
            aos[3 * col] = 0.;
            aos[3 * col + 1] = s.cgx[i] + cor * (-s.cev[slot]);
            aos[3 * col + 2] = s.cgy[i] + cor * s.ceu[slot];
        }
    }
}

// ---------------------------------------------------------------------------------------
// halo pack / unpack (multi-GPU): buffers hold (w,u,v) per listed cell, buf[3k + c], so the
// segment of each peer is contiguous
// ---------------------------------------------------------------------------------------
__global__ void k_halo_pack(int n, const int *cells, const double *w, const double *u, const double *v, double *buf) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int c = cells[k];
    buf[3 * (size_t)k] = w[c]; buf[3 * (size_t)k + 1] = u[c]; buf[3 * (size_t)k + 2] = v[c];
}
__global__ void k_halo_unpack(int n, const int *cells, const double *buf, double *w, double *u, double *v) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int c = cells[k];
    w[c] = buf[3 * (size_t)k]; u[c] = buf[3 * (size_t)k + 1]; v[c] = buf[3 * (size_t)k + 2];
}

// peer-memory halo transport: after the pack kernel stored this rank's boundary states straight into
// the neighbour GPU's receive buffer (NVLink stores through a CUDA-IPC mapping), one thread publishes
// the exchange sequence number in the neighbour's flag slot; the receiver spins on its own flags.
// CFL edge mask (multi-GPU: only edges touching owned cells count): excluded edges get dmin = +inf,
// so their length/wavespeed candidate is +inf and never wins the min — no mask test in the flux kernel
__global__ void k_apply_cfl_mask(int ne, const unsigned char *mask, const double *dmin0, double *dmin) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= ne) return;
    dmin[e] = mask[e] ? dmin0[e] : __longlong_as_double(0x7ff0000000000000ll);
}
__global__ void k_halo_signal(volatile int *peer_flag, int seq) {
    __threadfence_system();
    *peer_flag = seq;
    __threadfence_system();
}
// waits until every peer has published `seq` (or ~timeout_cycles passed: sets err[0] = 1 instead of hanging)
__global__ void k_halo_wait(volatile int *flags, int npeers, int seq, long long timeout_cycles, int *err) {
    const int p = threadIdx.x;
    if (p >= npeers) return;
    const long long t0 = clock64();
    while (flags[p] < seq) {
        if (clock64() - t0 > timeout_cycles) { err[0] = 1; break; }
        __nanosleep(200);
    }
    __threadfence_system();
}

// Fused send side of the peer-memory transport: pack this rank's boundary states straight into the
// neighbour GPU's receive buffer (NVLink stores) and, from the last block to finish, publish the
// exchange number in the neighbour's flag slot. One launch per peer, right after the boundary cells were
// updated, so the stores fly while the interior is still being updated.
__global__ void k_halo_pack_signal(int n, const int *cells, const double *w, const double *u, const double *v, double *peer_buf,
                                   volatile int *peer_flag, int seq, int *ticket) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) {
        const int c = cells[k];
        peer_buf[3 * (size_t)k] = w[c]; peer_buf[3 * (size_t)k + 1] = u[c]; peer_buf[3 * (size_t)k + 2] = v[c];
    }
    __threadfence_system();  // this thread's remote stores are visible system-wide before the ticket
    __syncthreads();
    if (threadIdx.x == 0) {
        const int t = atomicAdd(ticket, 1);
        if (t == (int)gridDim.x - 1) {
            *ticket = 0;
            __threadfence_system();
            *peer_flag = seq;
            __threadfence_system();
        }
    }
}
// Fused receive side: every block waits until all peers have published `seq` (time-out: err[0] = 1 instead
// of hanging), then the grid unpacks the receive buffer into the halo cells. Buffer reads bypass L1 (the
// data was written by another GPU).
__global__ void k_halo_wait_unpack(volatile int *flags, int npeers, int seq, long long timeout_cycles, int *err, int n,
                                   const int *cells, const double *buf, double *w, double *u, double *v) {
    if ((int)threadIdx.x < npeers) {
        const long long t0 = clock64();
        while (flags[threadIdx.x] < seq) {
            if (clock64() - t0 > timeout_cycles) { err[0] = 1; break; }
            __nanosleep(100);
        }
        __threadfence_system();
    }
    __syncthreads();
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const int c = cells[k];
        w[c] = __ldcg(buf + 3 * (size_t)k); u[c] = __ldcg(buf + 3 * (size_t)k + 1); v[c] = __ldcg(buf + 3 * (size_t)k + 2);
    }
}
// Global CFL minimum over peer memory (one double per rank, no NCCL on the path): push stores this rank's
// min_len_to_wavespeed into slot [seq & 1][rank] of EVERY rank's table and then raises that rank's flag;
// pull waits for all world flags and takes the minimum (exact, order independent => the same dt on every
// GPU count). Double buffering by the parity of seq suffices: a rank cannot push step n+2 before every rank
// has pulled step n (each pull needs every rank's push of the same step).
constexpr int kMaxRanks = 16;
struct MinPeers { double *buf[kMaxRanks]; int *flag[kMaxRanks]; };
__global__ void k_min_push(const double *scal, MinPeers t, int world, int rank, int seq) {
    const int p = threadIdx.x;
    if (p >= world) return;
    t.buf[p][(seq & 1) * kMaxRanks + rank] = scal[0];
    __threadfence_system();
    *(volatile int *)(t.flag[p] + rank) = seq;
    __threadfence_system();
}
// The global minimum goes to scal[4]; scal[0] keeps belonging to the flux kernel (the pull is deferred into the
// NEXT step, after its first flux evaluation has already overwritten scal[0]). also_scal0: a host query forces
// the pull at the end of a step and wants the classic slot updated as well.
__global__ void k_min_pull(double *scal, const double *buf, volatile int *flag, int world, int seq, long long timeout_cycles, int *err,
                           int also_scal0) {
    const int p = threadIdx.x;
    if (p < world) {
        const long long t0 = clock64();
        while (flag[p] < seq) {
            if (clock64() - t0 > timeout_cycles) { err[0] = 1; break; }
            __nanosleep(100);
        }
        __threadfence_system();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double mn = __ldcg(buf + (seq & 1) * kMaxRanks);
        for (int q = 1; q < world; ++q) { const double x = __ldcg(buf + (seq & 1) * kMaxRanks + q); mn = (x < mn) ? x : mn; }
        scal[4] = mn;
        if (also_scal0) scal[0] = mn;
    }
}

// Order-independent 64-bit hash of the cell states: sum over the selected cells of mix(global id, component,
// bit pattern). Equal on any partition of the same global mesh iff every owned cell state is bit-identical.
__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__global__ void k_state_hash(int nt, const int *old, const long long *gid, const unsigned char *mask, const double *w,
                             const double *u, const double *v, unsigned long long *out) {
    unsigned long long acc = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nt; i += gridDim.x * blockDim.x) {
        const int o = old ? old[i] : i;  // caller (local) id
        if (mask && !mask[o]) continue;
        const unsigned long long g = gid ? (unsigned long long)gid[o] : (unsigned long long)o;
        acc += mix64(mix64(3ull * g) ^ (unsigned long long)__double_as_longlong(w[i]));
        acc += mix64(mix64(3ull * g + 1ull) ^ (unsigned long long)__double_as_longlong(u[i]));
        acc += mix64(mix64(3ull * g + 2ull) ^ (unsigned long long)__double_as_longlong(v[i]));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, acc);
}

// ---------------------------------------------------------------------------------------
// diagnostics: fixed-shape two-level tree => deterministic for a given mesh size
// ---------------------------------------------------------------------------------------
constexpr int kDiagBlocks = 592;  // 4 per SM
constexpr int kDiagThreads = 256;
__global__ void __launch_bounds__(kDiagThreads) k_diag_partial(DevMesh m, DevFields s, double *partial) {
    double mass = 0., kin = 0., pot = 0., vmax = 0., hmin = __longlong_as_double(0x7ff0000000000000ll), wet = 0.;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m.nt; i += gridDim.x * blockDim.x) {
        const double b = m.cgeo[i].z, A = m.area[i];
        const double h = s.w[i] - b, u = s.u[i], v = s.v[i];
        mass += A * h;
        kin += A * (0.5 * h * (u * u + v * v));
        pot += A * (0.5 * h * h + h * b);
        vmax = fmax(vmax, fmax(fabs(u), fabs(v)));
        hmin = fmin(hmin, h);
        wet += is_wet(h) ? 1. : 0.;
    }
    __shared__ double sh[6][kDiagThreads];
    sh[0][threadIdx.x] = mass; sh[1][threadIdx.x] = kin; sh[2][threadIdx.x] = pot;
    sh[3][threadIdx.x] = vmax; sh[4][threadIdx.x] = hmin; sh[5][threadIdx.x] = wet;
    __syncthreads();
    for (int st = kDiagThreads / 2; st > 0; st >>= 1) {
        if (threadIdx.x < st) {
            sh[0][threadIdx.x] += sh[0][threadIdx.x + st];
            sh[1][threadIdx.x] += sh[1][threadIdx.x + st];
            sh[2][threadIdx.x] += sh[2][threadIdx.x + st];
            sh[3][threadIdx.x] = fmax(sh[3][threadIdx.x], sh[3][threadIdx.x + st]);
            sh[4][threadIdx.x] = fmin(sh[4][threadIdx.x], sh[4][threadIdx.x + st]);
            sh[5][threadIdx.x] += sh[5][threadIdx.x + st];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0)
        for (int q = 0; q < 6; ++q) partial[q * kDiagBlocks + blockIdx.x] = sh[q][0];
}
__global__ void k_diag_final(const double *partial, double *out) {
    const int q = threadIdx.x;
    if (q >= 6) return;
    double acc = partial[q * kDiagBlocks];
    for (int b = 1; b < kDiagBlocks; ++b) {
        const double x = partial[q * kDiagBlocks + b];
        if (q == 3) acc = fmax(acc, x); else if (q == 4) acc = fmin(acc, x); else acc += x;
    }
    out[q] = acc;
}

}  // namespace swe
