// swe_device.cuh — fp64 device numerics of the SWE_FVM time step (sm_100a, no tensor cores:
// nothing here is a dense contraction). Compiled with -fmad=false so that every expression is
// evaluated in the written order in IEEE binary64, exactly like the reference compiled for
// x86-64 SSE2; fp64 div/sqrt are IEEE-rounded on device. Citations: upstream SWE_FVM tree.
#pragma once
#include <cstdint>

#include "../../include/swe_constants.h"

namespace swe {

constexpr double kTol = SWE_TOL;  // include/Includes.h:30

__device__ __forceinline__ bool is_wet(double h) { return h > SWE_WET_DEPTH; }  // include/Bathymetry.h:5-8
// std::min / std::max semantics (argument order matters for NaN and signed zeros)
__device__ __forceinline__ double smin(double a, double b) { return (b < a) ? b : a; }
__device__ __forceinline__ double smax(double a, double b) { return (a < b) ? b : a; }

// cbrt with IEEE + - * / only (S9: libm and CUDA cbrt differ in the last ulp).
__device__ __noinline__ double det_cbrt(double x) {
    if (x == 0.0 || x != x || isinf(x)) return x;
    double a = fabs(x);
    double scale = 1.0;
    if (a < 1e-280) { a *= 0x1p+162; scale = 0x1p-54; }
    unsigned long long bits = (unsigned long long)__double_as_longlong(a);
    bits = bits / 3 + 0x2A9F7893782DA1CEull;
    double y = __longlong_as_double((long long)bits);
    for (int it = 0; it < 5; ++it) y = (2.0 * y + a / (y * y)) / 3.0;
    y = y - (y * y * y - a) / (3.0 * (y * y));
    y *= scale;
    return x < 0 ? -y : y;
}

// (int)log2(x), x > 0 finite (src/PointOperations.cpp:33), from the exponent bits.
__device__ __forceinline__ int ilog2_trunc(double x) {
    unsigned long long bits = (unsigned long long)__double_as_longlong(x);
    int e = int((bits >> 52) & 0x7ff);
    unsigned long long frac = bits & 0xfffffffffffffull;
    if (e == 0) {
        if (frac == 0) return INT_MIN;
        int sh = 0;
        while (!(frac & (1ull << 52))) { frac <<= 1; ++sh; }
        frac &= 0xfffffffffffffull;
        e = 1 - sh;
    }
    e -= 1023;
    if (e >= 0) return e;
    return frac == 0 ? e : e + 1;
}

struct CubicPoly {  // include/CubicPolyMath.h:6-19
    double b, c, d;
    __device__ __forceinline__ double operator()(double x) const { return x * x * x + b * x * x + c * x + d; }
};

__device__ __forceinline__ bool sbit(double v) { return __double_as_longlong(v) < 0; }

// src/PointOperations.cpp:26-40
__device__ __noinline__ double bisection(CubicPoly f, double xmin, double xmax) {
    if (sbit(f(xmin)) == sbit(f(xmax))) return (fabs(f(xmin)) < fabs(f(xmax))) ? xmin : xmax;
    int n = 50 + ilog2_trunc(xmax - xmin);
    double x = xmin;
    bool smin_ = sbit(f(xmin));  // f(xmin) only changes when xmin does: evaluate the cubic once per iteration, same bits
    for (int i = 0; i <= n; i++) {
        x = 0.5 * (xmin + xmax);
        const bool sx = sbit(f(x));
        if (smin_ != sx) xmax = x; else { xmin = x; smin_ = sx; }
    }
    return x;
}

// src/PointOperations.cpp:42-48: plane slope through 3 points, 2x2 partial-pivot LU.
struct Lu2 {  // factorisation shared by the three components of one stencil
    double a00, a01, l, u11;
    bool swapped;
};
__device__ __forceinline__ Lu2 lu2_factor(double x0, double y0, double x1, double y1, double x2, double y2) {
    Lu2 f;
    double a00 = x1 - x0, a01 = y1 - y0, a10 = x2 - x0, a11 = y2 - y0;
    f.swapped = fabs(a10) > fabs(a00);
    if (f.swapped) { double t = a00; a00 = a10; a10 = t; t = a01; a01 = a11; a11 = t; }
    f.a00 = a00; f.a01 = a01;
    f.l = a10 / a00;
    f.u11 = a11 - f.l * a01;
    return f;
}
__device__ __forceinline__ void lu2_solve(const Lu2 &f, double z0, double z1, double z2, double &g0, double &g1) {
    double r0 = z1 - z0, r1 = z2 - z0;
    if (f.swapped) { double t = r0; r0 = r1; r1 = t; }
    double c1 = r1 - f.l * r0;
    g1 = c1 / f.u11;
    g0 = (r0 - f.a01 * g1) / f.a00;
}
__device__ __forceinline__ void gradient3(double x0, double y0, double z0, double x1, double y1, double z1,
                                          double x2, double y2, double z2, double &g0, double &g1) {
    Lu2 f = lu2_factor(x0, y0, x1, y1, x2, y2);
    lu2_solve(f, z0, z1, z2, g0, g1);
}

// ReconstructPartWetCell1 (src/MUSCLObject.cpp:86-112): flat free surface holding the cell volume.
// *branch (optional): 0 submerged, 1 closed form (cbrt), 2 cubic by bisection.
__device__ __noinline__ double partwet1_level(double w, double cb, double b13, double b23, int *branch) {
    double b12 = 3. * cb - b23 - b13;
    double b_delimiter = b12 + (1. / 3.) * (b13 - b12) * (b13 - b12) / (b13 - b23);
    double hi = w - cb;
    if (w >= b13) { if (branch) *branch = 0; return w; }
    if (w <= b_delimiter) { if (branch) *branch = 1; return b23 + det_cbrt(3. * hi * (b13 - b23) * (b12 - b23)); }
    if (branch) *branch = 2;
    CubicPoly p;
    p.b = -3. * b13;
    p.c = 3. * (b12 * b13 + b13 * b23 - b12 * b23);
    p.d = (b13 - b23) * (3. * hi * (b13 - b12) - b12 * (b12 + b23)) - b23 * b23 * b13;
    return bisection(p, b12, b13);
}

// ElemFlux (include/SpaceDisc.h:4-18)
__device__ __forceinline__ void elem_flux(double nx, double ny, double h, double hu, double hv,
                                          double &f0, double &f1, double &f2) {
    f0 = 0.; f1 = 0.; f2 = 0.;
    if (is_wet(h)) {
        double hveln = hu * nx + hv * ny;
        f0 = hveln;
        f1 = (hveln / h) * hu + (0.5 * h * h) * nx;
        f2 = (hveln / h) * hv + (0.5 * h * h) * ny;
    }
}

enum { WS_RUSANOV = 0, WS_DAVIS = 1, WS_EINFELDT = 2 };
enum { FLUX_HLL = 0, FLUX_HLLC = 1 };

// Wavespeeds (src/Fluxes.cpp:5-26). Einfeldt keeps `cl * ur` as written upstream (S5).
// roe_fix (S5 alternative, only read when OPT): cr * ur instead of the as-written cl * ur.
template <int WS, bool OPT = false>
__device__ __forceinline__ void wavespeeds(double ul, double hl, double ur, double hr, double &a0, double &a1, int roe_fix = 0) {
    double cl = sqrt(hl), cr = sqrt(hr);
    if (WS == WS_RUSANOV) {
        double aplus = smax(fabs(ul) + cl, fabs(ur) + cr);
        a0 = -aplus; a1 = aplus;
    } else if (WS == WS_DAVIS) {
        a0 = smin(ul - cl, ur - cr); a1 = smax(ul + cl, ur + cr);
    } else {
        double uRoe = (cl * ul + ((OPT && roe_fix) ? cr : cl) * ur) / (cl + cr);
        double cRoe = sqrt(0.5 * (hl + hr));
        a0 = smin(ul - cl, uRoe - cRoe); a1 = smax(ur + cr, uRoe + cRoe);
    }
}

// Fluxes::HLL<W> (include/Fluxes.h:14-54) / Fluxes::HLLC<W> (:56-111) on one interior edge.
// (nx, ny) = outward normal of the `from` cell; t = (-ny, nx). l2w = length/wavespeed candidate
// for the CFL min (left untouched on the early-outs, like the reference).
// OPT: the S5 / S6 alternatives (roe_fix, cfl_abs) are honoured; the default instantiation is as written upstream.
template <int FLUX, int WS, bool OPT = false>
__device__ __forceinline__ void riemann_flux(double nx, double ny, double hl, double uxl, double uyl, double hr,
                                             double uxr, double uyr, double dmin, double abscor, double &f0,
                                             double &f1, double &f2, double &l2w, int roe_fix = 0, int cfl_abs = 0) {
    const double tx = -ny, ty = nx;
    double ul = uxl * nx + uyl * ny;
    double ur = uxr * nx + uyr * ny;
    f0 = 0.; f1 = 0.; f2 = 0.;
    if (hl + hr <= SWE_FLUX_DRY_SUM) return;
    double al, ar;
    wavespeeds<WS, OPT>(ul, hl, ur, hr, al, ar, roe_fix);
    const double Ul0 = hl, Ul1 = hl * uxl, Ul2 = hl * uyl;
    const double Ur0 = hr, Ur1 = hr * uxr, Ur2 = hr * uyr;
    if (FLUX == FLUX_HLL) {
        al = smin(0., al);
        ar = smax(0., ar);
        if (ar - al <= SWE_FLUX_DRY_SUM) return;
        l2w = dmin / (abscor + smax(-al, ar));
        double l0, l1, l2, r0, r1, r2;
        elem_flux(nx, ny, Ul0, Ul1, Ul2, l0, l1, l2);
        elem_flux(nx, ny, Ur0, Ur1, Ur2, r0, r1, r2);
        f0 = (ar * l0 - al * r0 + (al * ar) * (Ur0 - Ul0)) / (ar - al);
        f1 = (ar * l1 - al * r1 + (al * ar) * (Ur1 - Ul1)) / (ar - al);
        f2 = (ar * l2 - al * r2 + (al * ar) * (Ur2 - Ul2)) / (ar - al);
        return;
    }
    double vl = uxl * tx + uyl * ty;
    double vr = uxr * tx + uyr * ty;
    double ustar = (ar - ur) * hr * ur - (al - ul) * hl * ul + 0.5 * (hl * hl - hr * hr);
    ustar /= (hr * (ar - ur) - hl * (al - ul));
    {   // S6: signed max as written; cfl_abs: magnitudes
        const double amax = (OPT && cfl_abs) ? smax(fabs(al), fabs(ar)) : smax(al, ar);
        l2w = dmin / (abscor + smax(kTol, amax));
    }
    if (ustar <= 0) {
        double urstar = vr * tx + ustar * ty;
        double vrstar = vr * nx + ustar * ny;
        double hrstar = hr * (ar - ur) / (ar - ustar);
        double e0, e1, e2;
        elem_flux(nx, ny, Ur0, Ur1, Ur2, e0, e1, e2);
        double s = smax(0., ar);
        f0 = e0 + s * (hrstar - Ur0);
        f1 = e1 + s * (hrstar * urstar - Ur1);
        f2 = e2 + s * (hrstar * vrstar - Ur2);
    } else {
        double ulstar = vl * tx + ustar * ty;
        double vlstar = vl * nx + ustar * ny;
        double hlstar = hl * (al - ul) / (al - ustar);
        double e0, e1, e2;
        elem_flux(nx, ny, Ul0, Ul1, Ul2, e0, e1, e2);
        double s = smin(0., al);
        f0 = e0 + s * (hlstar - Ul0);
        f1 = e1 + s * (hlstar * ulstar - Ul1);
        f2 = e2 + s * (hlstar * vlstar - Ul2);
    }
}

// Order-independent max on doubles (any signs) with integer atomics: deterministic because
// max is associative/commutative. v must not be NaN; -0.0 is canonicalised to +0.0.
__device__ __forceinline__ void atomic_max_double(double *addr, double v) {
    v = v + 0.0;
    if (v >= 0.0) atomicMax((long long *)addr, __double_as_longlong(v));
    else atomicMin((unsigned long long *)addr, (unsigned long long)__double_as_longlong(v));
}
// min of positive doubles
__device__ __forceinline__ void atomic_min_pos_double(double *addr, double v) {
    atomicMin((unsigned long long *)addr, (unsigned long long)__double_as_longlong(v));
}

}  // namespace swe
