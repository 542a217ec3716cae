// swe_flux_registry.cuh — the pluggable Riemann fluxes of the device path.
//
// Upstream plugs the flux into SpaceDisc as a std::function<Array<3>(SpaceDisc*, Idx e, Idx from, Idx to,
// double* r)> (include/SpaceDisc.h:22) instantiated from Fluxes::HLL<W> / Fluxes::HLLC<W> (include/Fluxes.h:14,56).
// A std::function cannot run on the device, so a flux here is a DEVICE FUNCTOR registered at compile time:
//
//   struct MyFlux {
//       template <bool OPT>   // OPT: honour the S5/S6 alternatives (roe_fix, cfl_abs); ignore it if not applicable
//       __device__ static void eval(double nx, double ny,                    // outward unit normal of the `from` cell
//                                   double hl, double uxl, double uyl,        // edge-side depth and velocity, from side
//                                   double hr, double uxr, double uyr,        // ... to side
//                                   double dmin, double abscor,               // min(2A_l/L, 2A_r/L), |cor|
//                                   double &f0, double &f1, double &f2,       // flux through the edge (mass, x-, y-momentum)
//                                   double &l2w,                              // length / wavespeed candidate of the CFL min
//                                   int roe_fix, int cfl_abs);                //   (leave untouched to contribute nothing)
//   };
//
// and one line in SWE_USER_FLUXES (csrc/user_fluxes.cuh): X(6, "MyFlux", MyFlux). The id / name pair is then
// visible through swe_fluxer_count / swe_fluxer_name / swe_fluxer_find and selectable with swe_set_fluxer (C),
// Fluxes::Registered("MyFlux") (C++) or SpaceDisc(flux="MyFlux") (Python). k_flux<F, OPT> is instantiated for
// every entry, edge loop / wall branch / CFL reduction are shared.
#pragma once
#include "swe_device.cuh"

namespace swe {

template <int FLUX, int WS>
struct BuiltinFlux {  // Fluxes::HLL<W> / Fluxes::HLLC<W>, W in Wavespeeds::{Rusanov, Davis, Einfeldt}
    template <bool OPT>
    __device__ __forceinline__ static void eval(double nx, double ny, double hl, double uxl, double uyl, double hr, double uxr,
                                                double uyr, double dmin, double abscor, double &f0, double &f1, double &f2,
                                                double &l2w, int roe_fix, int cfl_abs) {
        riemann_flux<FLUX, WS, OPT>(nx, ny, hl, uxl, uyl, hr, uxr, uyr, dmin, abscor, f0, f1, f2, l2w, roe_fix, cfl_abs);
    }
};

// comma-free names for the registry macro
using FluxHllRusanov = BuiltinFlux<FLUX_HLL, WS_RUSANOV>;
using FluxHllDavis = BuiltinFlux<FLUX_HLL, WS_DAVIS>;
using FluxHllEinfeldt = BuiltinFlux<FLUX_HLL, WS_EINFELDT>;
using FluxHllcRusanov = BuiltinFlux<FLUX_HLLC, WS_RUSANOV>;
using FluxHllcDavis = BuiltinFlux<FLUX_HLLC, WS_DAVIS>;
using FluxHllcEinfeldt = BuiltinFlux<FLUX_HLLC, WS_EINFELDT>;

}  // namespace swe

#include "user_fluxes.cuh"
#ifndef SWE_USER_FLUXES
#define SWE_USER_FLUXES(X)
#endif

// id = 3 * swe_flux + swe_wavespeed for the built-in ones (so the enum pair of swe_step keeps working)
#define SWE_FLUX_LIST(X)                                              \
    X(0, "HLL<Rusanov>", swe::FluxHllRusanov)     \
    X(1, "HLL<Davis>", swe::FluxHllDavis)         \
    X(2, "HLL<Einfeldt>", swe::FluxHllEinfeldt)   \
    X(3, "HLLC<Rusanov>", swe::FluxHllcRusanov)   \
    X(4, "HLLC<Davis>", swe::FluxHllcDavis)       \
    X(5, "HLLC<Einfeldt>", swe::FluxHllcEinfeldt) \
    SWE_USER_FLUXES(X)
