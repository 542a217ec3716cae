// swe_kernels.cuh — the hot path: one RK stage = reconstruct -> (part-wet pass 2) -> edge flux
// + CFL min -> draining dt -> stage update. Hand-written fp64 CUDA for sm_100a.
//
// Data layout in HBM (structure of arrays, int32 ids, device numbering):
//   cells : tt/te/tp[k*nt+i] (k-major so a warp reads 32 consecutive ids), cgeo[i] = 32-byte
//           packet (cx, cy, cb, bfull) gathered for neighbours with one sector, area[i]
//   nodes : node[p] = 32-byte packet (x, y, b, -)
//   edges : slotL/slotR[e] (where the two cells keep their side of edge e), en[e] = (nx, ny)
//           outward normal of EdgeTriangs[0], elen[e], dmin[e] = min(2A_l/L, 2A_r/L)
//   state : w/u/v[i] SoA, ping-pong buffers (no U0 copy kernels)
//   edge-side values (m_edg, m_src upstream) are stored CELL-major, c*[k*nt+i] = value of cell
//   i's side of its k-th edge: written coalesced by the reconstruction, read coalesced by the
//   stage update, gathered once by the flux kernel through slotL/slotR.
//   te is sign-encoded: e >= 0 -> cell is EdgeTriangs(e)[0] (sgn = +1), ~e -> cell is [1] (-1).
#pragma once
#include <cuda_runtime.h>

#include "swe_device.cuh"

namespace swe {

struct DevMesh {
    int nt, ne, nn;
    const int *tt, *te, *tp;
    const double4 *cgeo;
    const double *area;
    const double4 *node;
    const int *slotL, *slotR;
    const double2 *en;
    const double *elen, *dmin;
    const unsigned char *cfl_mask;  // nullable
};

struct DevFields {
    double *w, *u, *v;               // current state (stage input)
    double *ceh, *ceu, *cev;         // [3*nt] edge-side depth h_e and (damped) velocities
    double *csx, *csy;               // [3*nt] edge-side source (grad w + cor * (-v, u))
    double *cew;                     // [3*nt] edge-side w, taps only (nullable)
    double *f0, *f1, *f2;            // [ne] fluxes
    double *maxw;                    // [nn] node maxima of reconstructed w (pass 1)
    double *dti;                     // [nt] draining dt
    signed char *cls;                // [nt] 0 dry, 1 part-wet, 2 full-wet
    double *scal;                    // [0] min_len_to_wavespeed, [1] dt, [2] time
    int *flags;                      // [0] non-finite state seen
};

constexpr int kBlock = 128;

// ---------------------------------------------------------------------------------------
// stage begin: m_max_wp = node bathymetry (src/SpaceDisc.cpp:34); min_len = 1 (:56)
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_stage_begin(DevMesh m, DevFields s) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p == 0) s.scal[0] = 1.0;
    if (p < m.nn) s.maxw[p] = m.node[p].z;
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
                                          double mx, double my, double mb, double cor) {
    const double dx = mx - cx, dy = my - cy;
    double a0 = M.o0 + (M.g00 * dx + M.g01 * dy);
    double a1 = M.o1 + (M.g10 * dx + M.g11 * dy);
    double a2 = M.o2 + (M.g20 * dx + M.g21 * dy);
    if (!((a0 - mb) >= 0)) { a0 = mb; a1 = 0.; a2 = 0.; }
    double h = a0 - mb;
    double ew, eh, eu, ev;
    if (!is_wet(h)) {
        ew = mb; eh = 0.; eu = 0.; ev = 0.;
    } else {
        ew = a0; eh = h; eu = a1; ev = a2;
        if (h < 1e-3) {
            double fac = sqrt(2.0) * h / sqrt(h * h + 1e-6);
            eu *= fac; ev *= fac;
        }
    }
    s.ceh[slot] = eh; s.ceu[slot] = eu; s.cev[slot] = ev;
    if (TAPS) s.cew[slot] = ew;
    // MUSCL::Gradient(p) (include/MUSCLObject.h:41-48) always evaluates to m_grad: when the
    // reconstructed depth is negative AtPoint returns the dry state, whose depth is exactly 0.
    s.csx[slot] = M.g00 + cor * (-ev);
    s.csy[slot] = M.g01 + cor * eu;
}

__device__ __forceinline__ double muscl_w_at(const Muscl &M, double cx, double cy, double px, double py, double pz) {
    double a0 = M.o0 + (M.g00 * (px - cx) + M.g01 * (py - cy));
    if (!((a0 - pz) >= 0)) a0 = pz;
    return a0;
}

// ---------------------------------------------------------------------------------------
// K1: classification + pass-1 reconstruction of every cell (src/SpaceDisc.cpp:37-45,
// src/MUSCLObject.cpp:13-112), node maxima by order-independent integer atomics.
// ---------------------------------------------------------------------------------------
template <bool TAPS>
__global__ void __launch_bounds__(kBlock) k_reconstruct(DevMesh m, DevFields s, double cor) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int nt = m.nt;
    if (i >= nt) return;
    const int ip0 = m.tp[i], ip1 = m.tp[nt + i], ip2 = m.tp[2 * nt + i];
    const int it0 = m.tt[i], it1 = m.tt[nt + i], it2 = m.tt[2 * nt + i];
    const double4 P0 = m.node[ip0], P1 = m.node[ip1], P2 = m.node[ip2];
    const double4 Gi = m.cgeo[i];
    const double cx = Gi.x, cy = Gi.y, cb = Gi.z;
    const double w = s.w[i], u = s.u[i], v = s.v[i];

    // edge midpoints E(ie[k]) = 0.5 (P(ep0) + P(ep1)), edge k joins ip[k], ip[k+1] (S1)
    const double mx0 = 0.5 * (P0.x + P1.x), my0 = 0.5 * (P0.y + P1.y), mb0 = 0.5 * (P0.z + P1.z);
    const double mx1 = 0.5 * (P1.x + P2.x), my1 = 0.5 * (P1.y + P2.y), mb1 = 0.5 * (P1.z + P2.z);
    const double mx2 = 0.5 * (P2.x + P0.x), my2 = 0.5 * (P2.y + P0.y), mb2 = 0.5 * (P2.z + P0.z);

    const bool bnd = (it0 | it1 | it2) < 0;
    const double bmax = smax(smax(P0.z, P1.z), P2.z);
    const bool dry = !is_wet(w - cb);
    const bool full = !bnd && (bmax < w);
    s.cls[i] = dry ? 0 : (full ? 2 : 1);

    Muscl M;
    M.g00 = M.g01 = M.g10 = M.g11 = M.g20 = M.g21 = 0.;
    if (dry) {  // ReconstructDryCell (:31-36): origin (b_i,0,0), w-gradient = bed slope
        M.o0 = cb; M.o1 = 0.; M.o2 = 0.;
        gradient3(P0.x, P0.y, P0.z, P1.x, P1.y, P1.z, P2.x, P2.y, P2.z, M.g00, M.g01);
    } else if (!full) {  // ReconstructPartWetCell1 (:86-112)
        const double bmin = smin(smin(P0.z, P1.z), P2.z);
        M.o0 = partwet1_level(w, cb, bmax, bmin); M.o1 = u; M.o2 = v;
    } else {  // ReconstructFullWetCell (:38-84), S2: plane gradients of w, u, v
        M.o0 = w; M.o1 = u; M.o2 = v;
        double X[3][2], V[3][3], N[3][3];
        bool zero_grad = false;
        const int itk[3] = {it0, it1, it2};
        const double mxk[3] = {mx0, mx1, mx2}, myk[3] = {my0, my1, my2}, mbk[3] = {mb0, mb1, mb2};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int j = itk[k];
            const double wj = s.w[j], uj = s.u[j], vj = s.v[j];
            const double4 Gj = m.cgeo[j];
            N[k][0] = wj; N[k][1] = uj; N[k][2] = vj;
            if (Gj.w < wj) {  // IsFullWetCell(j): bfull = +inf on boundary triangles
                X[k][0] = Gj.x; X[k][1] = Gj.y;
                V[k][0] = wj; V[k][1] = uj; V[k][2] = vj;
            } else if (!is_wet(wj - Gj.z)) {  // IsDryCell(j) -> zero gradient
                zero_grad = true;
                X[k][0] = X[k][1] = 0.; V[k][0] = V[k][1] = V[k][2] = 0.;
            } else {  // part-wet neighbour: its PartWet1 value at the shared edge midpoint
                const double z0 = m.node[m.tp[j]].z, z1 = m.node[m.tp[nt + j]].z, z2 = m.node[m.tp[2 * nt + j]].z;
                const double b13 = smax(smax(z0, z1), z2), b23 = smin(smin(z0, z1), z2);
                double a0 = partwet1_level(wj, Gj.z, b13, b23), a1 = uj, a2 = vj;
                // AtPoint with a zero gradient: o + (0*dx + 0*dy)
                const double dx = mxk[k] - Gj.x, dy = myk[k] - Gj.y;
                a0 = a0 + (0. * dx + 0. * dy); a1 = a1 + (0. * dx + 0. * dy); a2 = a2 + (0. * dx + 0. * dy);
                if (!((a0 - mbk[k]) >= 0)) { a0 = mbk[k]; a1 = 0.; a2 = 0.; }
                X[k][0] = mxk[k]; X[k][1] = myk[k];
                V[k][0] = 0.5 * (w + a0); V[k][1] = 0.5 * (u + a1); V[k][2] = 0.5 * (v + a2);
            }
        }
        if (!zero_grad) {
            double df[3][2];
            const Lu2 lu = lu2_factor(X[0][0], X[0][1], X[1][0], X[1][1], X[2][0], X[2][1]);
#pragma unroll
            for (int c = 0; c < 3; ++c) lu2_solve(lu, V[0][c], V[1][c], V[2][c], df[c][0], df[c][1]);
            // vertex positivity (:66-72): dx = P(ip) * (I - 1/3), evaluated literally
            const double md = 1. - 1. / 3., mo = 0. - 1. / 3.;
            const double dx0 = (P0.x * md + P1.x * mo) + P2.x * mo, dy0 = (P0.y * md + P1.y * mo) + P2.y * mo;
            const double dx1 = (P0.x * mo + P1.x * md) + P2.x * mo, dy1 = (P0.y * mo + P1.y * md) + P2.y * mo;
            const double dx2 = (P0.x * mo + P1.x * mo) + P2.x * md, dy2 = (P0.y * mo + P1.y * mo) + P2.y * md;
            const double hp0 = ((df[0][0] * dx0 + df[0][1] * dy0) + w) - P0.z;
            const double hp1 = ((df[0][0] * dx1 + df[0][1] * dy1) + w) - P1.z;
            const double hp2 = ((df[0][0] * dx2 + df[0][1] * dy2) + w) - P2.z;
            if (!(is_wet(hp0) && is_wet(hp1) && is_wet(hp2))) {
#pragma unroll
                for (int c = 0; c < 3; ++c) df[c][0] = df[c][1] = 0.;
            }
            // on/off TVD limiter (:74-81)
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
            M.g00 = tvd[0] * df[0][0]; M.g01 = tvd[0] * df[0][1];
            M.g10 = tvd[1] * df[1][0]; M.g11 = tvd[1] * df[1][1];
            M.g20 = tvd[2] * df[2][0]; M.g21 = tvd[2] * df[2][1];
        }
    }
    // UpdateInterfaceValues (src/SpaceDisc.cpp:15-31)
    atomic_max_double(&s.maxw[ip0], muscl_w_at(M, cx, cy, P0.x, P0.y, P0.z));
    atomic_max_double(&s.maxw[ip1], muscl_w_at(M, cx, cy, P1.x, P1.y, P1.z));
    atomic_max_double(&s.maxw[ip2], muscl_w_at(M, cx, cy, P2.x, P2.y, P2.z));
    emit_edge<TAPS>(s, i, M, cx, cy, mx0, my0, mb0, cor);
    emit_edge<TAPS>(s, nt + i, M, cx, cy, mx1, my1, mb1, cor);
    emit_edge<TAPS>(s, 2 * nt + i, M, cx, cy, mx2, my2, mb2, cor);
}

// ---------------------------------------------------------------------------------------
// K1b: pass 2, ReconstructPartWetCell2 (src/MUSCLObject.cpp:114-191, S3) for part-wet cells;
// reads the node maxima of pass 1 only and does not update them (S8).
// ---------------------------------------------------------------------------------------
template <bool TAPS>
__global__ void __launch_bounds__(kBlock) k_partwet2(DevMesh m, DevFields s, double cor) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int nt = m.nt;
    if (i >= nt) return;
    if (s.cls[i] != 1) return;
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
    if ((w > b13) || (b13 - b23 < kTol)) return;  // falls back to PartWet1 = what pass 1 wrote

    const double w23 = s.maxw[q0];
    const double h23 = w23 - b23;
    const double ratio_b = (b12 - b23) / (b13 - b23);
    const double h_delimiter1 = 1. / 3. * h23 * ratio_b;
    const double h_delimiter2 = 1. / 3. * h23 * (2. * b13 - b12 - b23) / (b13 - b23);
    const double hi = w - cb;
    const double ratio_h = hi / h23;
    const double S0x = Q0.x, S0y = Q0.y, S0z = w23;
    double S1x, S1y, S1z, S2x, S2y, S2z;
    if (hi <= h_delimiter1) {  // one vertex wet
        const double k2 = sqrt(3. * ratio_h / ratio_b);
        S1x = k2 * Q1.x + (1. - k2) * Q0.x; S1y = k2 * Q1.y + (1. - k2) * Q0.y; S1z = k2 * Q1.z + (1. - k2) * Q0.z;
        const double k3 = sqrt(3. * ratio_h * ratio_b);
        S2x = k3 * Q2.x + (1. - k3) * Q0.x; S2y = k3 * Q2.y + (1. - k3) * Q0.y; S2z = k3 * Q2.z + (1. - k3) * Q0.z;
    } else if (hi >= h_delimiter2) {  // three vertices wet
        const double delta_w = 1.5 * (hi - h_delimiter2);
        S1x = Q1.x; S1y = Q1.y; S1z = Q1.z;
        S1z += delta_w;
        S1z += (1. - ratio_b) * h23;
        S2x = Q2.x; S2y = Q2.y; S2z = Q2.z;
        S2z += delta_w;
    } else {  // two vertices wet
        const double alpha = 3. * ratio_h;
        const double beta = (b13 - b12) / (b13 - b23);
        CubicPoly p;
        p.d = (1. + beta - alpha) / (beta * beta); p.c = (alpha - 3.) / beta; p.b = 0.;
        const double k1 = 1. - bisection(p, 0., 1.);
        if (k1 < kTol) return;
        const double k3 = 1. - beta * (1. - k1);
        const double bp1 = k1 * b13 + (1. - k1) * b12;
        S1x = Q1.x; S1y = Q1.y;
        S1z = b12 + (k1 / k3) * beta * h23;
        S2x = k1 * Q2.x + (1. - k1) * Q1.x;
        S2y = k1 * Q2.y + (1. - k1) * Q1.y;
        S2z = bp1;
    }
    Muscl M;
    M.g10 = M.g11 = M.g20 = M.g21 = 0.;
    gradient3(S0x, S0y, S0z, S1x, S1y, S1z, S2x, S2y, S2z, M.g00, M.g01);
    M.o0 = w23 + (M.g00 * (cx - Q0.x) + M.g01 * (cy - Q0.y));
    M.o1 = u; M.o2 = v;
    const double mx0 = 0.5 * (P0.x + P1.x), my0 = 0.5 * (P0.y + P1.y), mb0 = 0.5 * (P0.z + P1.z);
    const double mx1 = 0.5 * (P1.x + P2.x), my1 = 0.5 * (P1.y + P2.y), mb1 = 0.5 * (P1.z + P2.z);
    const double mx2 = 0.5 * (P2.x + P0.x), my2 = 0.5 * (P2.y + P0.y), mb2 = 0.5 * (P2.z + P0.z);
    emit_edge<TAPS>(s, i, M, cx, cy, mx0, my0, mb0, cor);
    emit_edge<TAPS>(s, nt + i, M, cx, cy, mx1, my1, mb1, cor);
    emit_edge<TAPS>(s, 2 * nt + i, M, cx, cy, mx2, my2, mb2, cor);
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

template <int FLUX, int WS>
__global__ void __launch_bounds__(kBlock) k_flux(DevMesh m, DevFields s, double abscor) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    double l2w = 1.0;  // reset value of m_min_length_to_wavespeed (:56)
    if (e < m.ne) {
        const int sl = m.slotL[e], sr = m.slotR[e];
        const double2 n = m.en[e];
        double f0, f1, f2;
        if (sr < 0) {  // SOLID_WALL: ElemFlux(n, {h(lf), 0, 0}) with the cell-mean depth (:67-69)
            const int lf = sl % m.nt;
            const double h = s.w[lf] - m.cgeo[lf].z;
            elem_flux(n.x, n.y, h, 0., 0., f0, f1, f2);
        } else {
            double cand = 1.0;
            riemann_flux<FLUX, WS>(n.x, n.y, s.ceh[sl], s.ceu[sl], s.cev[sl], s.ceh[sr], s.ceu[sr], s.cev[sr],
                                   m.dmin[e], abscor, f0, f1, f2, cand);
            if (m.cfl_mask == nullptr || m.cfl_mask[e]) l2w = cand;
        }
        s.f0[e] = f0; s.f1[e] = f1; s.f2[e] = f2;
    }
    __shared__ double red[kBlock / 32];
    l2w = warp_min(l2w);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = l2w;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = (threadIdx.x < kBlock / 32) ? red[threadIdx.x] : 1.0;
        v = warp_min(v);
        if (threadIdx.x == 0 && v < 1.0) atomic_min_pos_double(&s.scal[0], v);
    }
}

// ---------------------------------------------------------------------------------------
// K3: draining time step (src/TimeDisc.cpp:43-66) of every cell from the stage-begin state
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) k_drain(DevMesh m, DevFields s) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int nt = m.nt;
    if (i >= nt) return;
    const double h = s.w[i] - m.cgeo[i].z;
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
    s.dti[i] = r;
}

// ---------------------------------------------------------------------------------------
// K4: RHS gather (src/TimeDisc.cpp:3-41) + RK combination (src/Solvers.cpp) + ConsAssigner
// (src/Assigners.cpp:22-44). Deterministic: fixed k order, no float atomics.
//   PLAIN: U = cons(W) + RHS                 (Euler `+=`, first RK stage)
//   else : U = a0*cons(U0) + a1*cons(W) + RHS
// Reads W (stage input) and writes Wout (may alias W: only the own cell is read).
// ---------------------------------------------------------------------------------------
template <bool PLAIN>
__global__ void __launch_bounds__(kBlock) k_update(DevMesh m, DevFields s, const double *__restrict__ w0,
                                                   const double *__restrict__ u0, const double *__restrict__ v0,
                                                   double *wout, double *uout, double *vout, double a0, double a1,
                                                   double dt_host, double dt_coef) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int nt = m.nt;
    if (i >= nt) return;
    // dt_coef != 0: stage dt = dt_coef * (device-resident dt), else the host value
    const double dt = (dt_coef != 0.) ? dt_coef * s.scal[1] : dt_host;
    const double cb = m.cgeo[i].z;
    const double i_area = 1. / m.area[i];
    const double dti = s.dti[i];
    double r0 = 0., r1 = 0., r2 = 0.;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int te = m.te[k * nt + i];
        const int tn = m.tt[k * nt + i];
        const bool first = te >= 0;
        const int e = first ? te : ~te;
        const double sgn = first ? 1. : -1.;
        const double F0 = s.f0[e], F1 = s.f1[e], F2 = s.f2[e];
        const double dtik = (tn < 0) ? __longlong_as_double(0x7ff0000000000000ll) : s.dti[tn];
        const double dtk = (sgn * F0 > 0.) ? smin(dt, dti) : smin(dt, dtik);
        const double c_ek = i_area * m.elen[e];
        const double h_ek = s.ceh[k * nt + i];
        const double sc = dtk * sgn * c_ek;
        r0 -= sc * F0; r1 -= sc * F1; r2 -= sc * F2;
        r1 -= dt * (1. / 3.) * s.csx[k * nt + i] * h_ek;
        r2 -= dt * (1. / 3.) * s.csy[k * nt + i] * h_ek;
        const double2 n = m.en[e];
        const double nx = sgn * n.x, ny = sgn * n.y;  // Norm(e, i) = -Norm(e, other) exactly
        r1 += dtk * (nx * c_ek * (0.5 * h_ek * h_ek));
        r2 += dtk * (ny * c_ek * (0.5 * h_ek * h_ek));
    }
    const double wc = s.w[i], uc = s.u[i], vc = s.v[i];
    const double hc = wc - cb;
    double U0, U1, U2;
    if (PLAIN) {
        U0 = hc + r0; U1 = uc * hc + r1; U2 = vc * hc + r2;
    } else {
        const double ha = w0[i] - cb;
        const double A1 = u0[i] * ha, A2 = v0[i] * ha;
        U0 = (a0 * ha + a1 * hc) + r0;
        U1 = (a0 * A1 + a1 * (uc * hc)) + r1;
        U2 = (a0 * A2 + a1 * (vc * hc)) + r2;
    }
    double ow, ou, ov;
    if (!is_wet(U0)) {
        ow = cb; ou = 0.; ov = 0.;
    } else {
        double ih;
        if (U0 < 1e-3) ih = sqrt(2.0) * U0 / sqrt(U0 * U0 * U0 * U0 + 1e-12);
        else ih = 1. / U0;
        ow = U0 + cb; ou = U1 * ih; ov = U2 * ih;
    }
    if (!(isfinite(ow) && isfinite(ou) && isfinite(ov))) s.flags[0] = 1;
    wout[i] = ow; uout[i] = ou; vout[i] = ov;
}

// after a step: time += dt_used; in adaptive mode dt = 0.15 * min_len (include/TimeDisc.h:13,22)
__global__ void k_post_step(DevFields s, double dt_host, int adaptive) {
    const double used = adaptive ? s.scal[1] : dt_host;
    s.scal[2] += used;
    if (adaptive) s.scal[1] = 0.15 * s.scal[0];
}
__global__ void k_set_scalar(double *p, double v) { *p = v; }

// ---------------------------------------------------------------------------------------
// setup: geometry from node coordinates with the reference's formulas (src/Bathymetry.cpp)
// ---------------------------------------------------------------------------------------
__global__ void k_setup_cells(int nt, const int *tp, const int *tt, const double4 *node, double4 *cgeo, double *area) {
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
    const double ax = P1.x - P0.x, ay = P1.y - P0.y, bx = P2.x - P0.x, by = P2.y - P0.y;
    area[i] = 0.5 * fabs(ax * by - bx * ay);  // Domain::Area (:87-90)
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
__global__ void k_edge_out(DevMesh m, DevFields s, const int *cell_old, const int *edge_old, int which, double *aos) {
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
            aos[3 * col] = 0.; aos[3 * col + 1] = s.csx[slot]; aos[3 * col + 2] = s.csy[slot];
        }
    }
}

// ---------------------------------------------------------------------------------------
// halo pack / unpack (multi-GPU): component-major buffers [3][n]
// ---------------------------------------------------------------------------------------
__global__ void k_halo_pack(int n, const int *cells, const double *w, const double *u, const double *v, double *buf) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int c = cells[k];
    buf[k] = w[c]; buf[n + k] = u[c]; buf[2 * (size_t)n + k] = v[c];
}
__global__ void k_halo_unpack(int n, const int *cells, const double *buf, double *w, double *u, double *v) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int c = cells[k];
    w[c] = buf[k]; u[c] = buf[n + k]; v[c] = buf[2 * (size_t)n + k];
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
