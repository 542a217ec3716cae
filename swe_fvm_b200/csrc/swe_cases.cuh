// swe_cases.cuh — the analytic test cases (upstream examples/Tests.h) on the device, so that
// initial conditions and error norms at 10^7-10^8 cells need no host loop.
// Same formulas as swe::case_eval in hostmesh.cpp; device libm differs from glibc in the last
// ulp of exp/sin/cos/atan, so a device-built initial state is an INPUT in its own right (the
// parity tests feed GPU and oracle the same host-built state).
#pragma once
#include "../../include/swe_b200.h"
#include "swe_kernels.cuh"

namespace swe {

__device__ __forceinline__ void dev_case_eval(const swe_case &c, double x, double y, double t, double &b, double &h,
                                              double &u, double &v) {
    const double dx = x - c.mid_x, dy = y - c.mid_y;
    b = 0.; h = 0.; u = 0.; v = 0.;
    switch (c.kind) {
        case SWE_CASE_LAKE_AT_REST:
            b = ((1. < x) && (x < 3.) && (1. < y) && (y < 3.)) ? -0.2 : -1.;
            h = fmax(0., -b);
            break;
        case SWE_CASE_CLASSIC_THACKER: {
            b = c.delta * (dx * dx + dy * dy - 1.0);
            const double w = sqrt(c.cor * c.cor + 8. * c.delta);
            const double qz = (c.q0 - 0.5 * c.cor) * (c.q0 - 0.5 * c.cor);
            const double rz = qz + 2. * c.H0 * c.H0 + c.p0 * c.p0 - 0.25 * w * w;
            const double a = sqrt(rz * rz + w * w * c.p0 * c.p0) / (rz + 0.5 * w * w);
            const double ph0 = atan(w * c.p0 / rz);
            const double ph = w * t + ph0;
            const double den = 1. - a * cos(ph);
            const double p = 0.5 * w * a * sin(ph) / den;
            const double q = (c.q0 - 0.5 * c.cor) * (1. - a * cos(ph0)) / den + 0.5 * c.cor;
            u = p * dx + q * dy;
            v = q * (c.mid_x - x) + p * dy;
            const double Hc = c.H0 * (1. - a * cos(ph0)) / den;
            const double az0 = (1. - a * cos(ph0)) * (1. - a * cos(ph0));
            const double Hxx = (0.25 * w * w * (a * a - 1.) + qz * az0) / (den * den);
            h = fmax(0., Hc + 0.5 * Hxx * dx * dx + 0.5 * Hxx * dy * dy);
            break;
        }
        case SWE_CASE_GAUSS_WAVE:
            h = 1. + exp(-5. * (dx * dx + dy * dy));
            break;
        case SWE_CASE_FULLY_WET: {
            const double two_pi = 6.283185307179586476925286766559;
            b = 0.1 * sin(two_pi * x / c.length) * sin(two_pi * y / c.length) - 1.;
            h = c.amp * exp(-5. * (dx * dx + dy * dy)) - b;
            break;
        }
        case SWE_CASE_BOWL_HUMP:
            b = c.delta * (dx * dx + dy * dy - 1.0);
            h = fmax(0., c.level + c.amp * exp(-5. * (dx * dx + dy * dy)) - b);
            break;
        default: break;
    }
}

// nodal bathymetry b(x, y) -> node[p].z (examples/Main.cpp:202-205)
__global__ void k_case_bathymetry(int nn, double4 *node, swe_case c) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nn) return;
    double4 P = node[p];
    double b, h, u, v;
    dev_case_eval(c, P.x, P.y, 0., b, h, u, v);
    P.z = b;
    node[p] = P;
}

// cell averages by TriangAverage<3, n> (include/PointOperations.h:20-44, same loop order), then
// w = h_avg + b_i and the PrimAssigner clamp (examples/Main.cpp:211-223, src/Assigners.cpp:8-20)
__global__ void k_case_init(DevMesh m, swe_case c, int n, double t, double *w, double *u, double *v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int nt = m.nt;
    if (i >= nt) return;
    const double4 P0 = m.node[m.tp[i]], P1 = m.node[m.tp[nt + i]], P2 = m.node[m.tp[2 * nt + i]];
    const double third = 1. / 3.;
    const double cx = P0.x * third + P1.x * third + P2.x * third;
    const double cy = P0.y * third + P1.y * third + P2.y * third;
    const double bi = P0.z * third + P1.z * third + P2.z * third;
    double x0 = 0., x1 = 0., x2 = 0.;
    if (c.kind == SWE_CASE_GAUSS_WAVE) {
        double b, h, uu, vv;
        dev_case_eval(c, cx, cy, t, b, h, uu, vv);
        x0 = h;
    } else if (c.kind == SWE_CASE_LAKE_AT_REST && P0.z <= 0. && P1.z <= 0. && P2.z <= 0.) {
        x0 = 0.;
    } else {
        const bool lake = c.kind == SWE_CASE_LAKE_AT_REST;
        const double hq = 1. / n;
        const double dix = hq * (P1.x - P0.x), diy = hq * (P1.y - P0.y);
        const double djx = hq * (P2.x - P0.x), djy = hq * (P2.y - P0.y);
        const double dtx = 1. / 3. * (dix + djx), dty = 1. / 3. * (diy + djy);
        const double det = (P1.x - P0.x) * (P2.y - P0.y) - (P2.x - P0.x) * (P1.y - P0.y);
        double s0 = 0., s1 = 0., s2 = 0.;
        double pix = P0.x, piy = P0.y;
        auto add = [&](double px, double py) {
            double b, h, uu, vv;
            if (lake) {  // examples/Main.cpp:333-336: max(0, linear bed over the cell)
                const double l1 = ((px - P0.x) * (P2.y - P0.y) - (P2.x - P0.x) * (py - P0.y)) / det;
                const double l2 = ((P1.x - P0.x) * (py - P0.y) - (px - P0.x) * (P1.y - P0.y)) / det;
                h = fmax(0., P0.z + l1 * (P1.z - P0.z) + l2 * (P2.z - P0.z)); uu = 0.; vv = 0.;
            } else {
                dev_case_eval(c, px, py, t, b, h, uu, vv);
            }
            s0 += hq * h; s1 += hq * uu; s2 += hq * vv;
        };
        for (int a = 0; a < n; a++) {
            double ptx = pix + dtx, pty = piy + dty;
            for (int j = 0; j < n - a - 1; j++) {
                add(ptx, pty);
                add(ptx + dtx, pty + dty);
                ptx += djx; pty += djy;
            }
            add(ptx, pty);
            pix += dix; piy += diy;
        }
        x0 = hq * s0; x1 = hq * s1; x2 = hq * s2;
        if (!lake) x0 += bi;
    }
    const double h = x0 - bi;
    if (!is_wet(h)) { w[i] = bi; u[i] = 0.; v[i] = 0.; return; }
    if (h < SWE_DAMP_DEPTH) {
        const double f = sqrt(2.0) * h / sqrt(h * h + SWE_DAMP_EPS_PRIM);
        x1 *= f; x2 *= f;
    }
    w[i] = x0; u[i] = x1; v[i] = x2;
}

// L2 error of (h, hu, hv) against the exact solution at time t, sampled at the cell centroids
// (the commented CompareWith / TriangAverage<3,1> of upstream src/SpaceDisc.cpp:106-138):
// partial[q*blocks + b] = sum over the block's cells of A_i * diff_q^2
__global__ void __launch_bounds__(kDiagThreads) k_case_error_partial(DevMesh m, DevFields s, swe_case c, double t, double *partial) {
    double e0 = 0., e1 = 0., e2 = 0.;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m.nt; i += gridDim.x * blockDim.x) {
        const double4 G = m.cgeo[i];
        double b, h, u, v;
        dev_case_eval(c, G.x, G.y, t, b, h, u, v);
        const double hn = s.w[i] - G.z;
        const double d0 = hn - h, d1 = hn * s.u[i] - h * u, d2 = hn * s.v[i] - h * v;
        const double A = m.area[i];
        e0 += A * d0 * d0; e1 += A * d1 * d1; e2 += A * d2 * d2;
    }
    __shared__ double sh[3][kDiagThreads];
    sh[0][threadIdx.x] = e0; sh[1][threadIdx.x] = e1; sh[2][threadIdx.x] = e2;
    __syncthreads();
    for (int st = kDiagThreads / 2; st > 0; st >>= 1) {
        if (threadIdx.x < st)
            for (int q = 0; q < 3; ++q) sh[q][threadIdx.x] += sh[q][threadIdx.x + st];
        __syncthreads();
    }
    if (threadIdx.x == 0)
        for (int q = 0; q < 3; ++q) partial[q * kDiagBlocks + blockIdx.x] = sh[q][0];
}
__global__ void k_case_error_final(const double *partial, double *out) {
    const int q = threadIdx.x;
    if (q >= 3) return;
    double acc = 0.;
    for (int b = 0; b < kDiagBlocks; ++b) acc += partial[q * kDiagBlocks + b];
    out[q] = sqrt(acc);
}

}  // namespace swe
