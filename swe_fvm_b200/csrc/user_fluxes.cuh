// user_fluxes.cuh — fluxes registered in addition to the reference's HLL / HLLC families.
// Add a functor (interface: swe_flux_registry.cuh) and one X(id, "name", Type) line; ids continue at 6.
#pragma once

namespace swe {

// Local Lax-Friedrichs (Rusanov-type central flux): F = 1/2 (F(Ul) + F(Ur)) . n - 1/2 a (Ur - Ul),
// a = max(|ul| + sqrt(hl), |ur| + sqrt(hr)). Not part of upstream: it is here to show (and test) that a new flux is
// one functor + one registry line; same early-outs and CFL candidate shape as Fluxes::HLL (include/Fluxes.h:28-48).
struct LocalLaxFriedrichs {
    template <bool OPT>
    __device__ __forceinline__ static void eval(double nx, double ny, double hl, double uxl, double uyl, double hr, double uxr,
                                                double uyr, double dmin, double abscor, double &f0, double &f1, double &f2,
                                                double &l2w, int, int) {
        f0 = 0.; f1 = 0.; f2 = 0.;
        if (hl + hr <= 1e-10) return;
        const double ul = uxl * nx + uyl * ny, ur = uxr * nx + uyr * ny;
        const double a = smax(fabs(ul) + sqrt(hl), fabs(ur) + sqrt(hr));
        if (a <= 1e-10) return;
        l2w = dmin / (abscor + a);
        double l0, l1, l2, r0, r1, r2;
        elem_flux(nx, ny, hl, hl * uxl, hl * uyl, l0, l1, l2);
        elem_flux(nx, ny, hr, hr * uxr, hr * uyr, r0, r1, r2);
        f0 = 0.5 * (l0 + r0) - 0.5 * a * (hr - hl);
        f1 = 0.5 * (l1 + r1) - 0.5 * a * (hr * uxr - hl * uxl);
        f2 = 0.5 * (l2 + r2) - 0.5 * a * (hr * uyr - hl * uyl);
    }
};

}  // namespace swe

#define SWE_USER_FLUXES(X) X(6, "LocalLaxFriedrichs", swe::LocalLaxFriedrichs)
