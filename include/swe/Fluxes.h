// Fluxes.h — the pluggable Riemann fluxes. Upstream passes `Fluxes::HLLC<Wavespeeds::Einfeldt>`
// (a function template instance) into SpaceDisc as a std::function (include/SpaceDisc.h:22,
// include/Fluxes.h:14,56, src/Fluxes.cpp:5-26). A std::function cannot run on the device, so the
// same spelling here names a compile-time tag that selects the matching device kernel
// instantiation (k_flux<FLUX, WS> in csrc/swe_kernels.cuh). New fluxes are added by registering
// another instantiation there and another tag here.
#pragma once
#include "../swe_b200.h"

struct Fluxer {
    swe_flux flux;
    swe_wavespeed wavespeed;
};

namespace Wavespeeds {
struct Rusanov { static constexpr swe_wavespeed id = SWE_RUSANOV; };
struct Davis { static constexpr swe_wavespeed id = SWE_DAVIS; };
struct Einfeldt { static constexpr swe_wavespeed id = SWE_EINFELDT; };
}  // namespace Wavespeeds

namespace Fluxes {
template <class W> inline constexpr Fluxer HLL{SWE_HLL, W::id};
template <class W> inline constexpr Fluxer HLLC{SWE_HLLC, W::id};
}  // namespace Fluxes
