// Fluxes.h — the pluggable Riemann fluxes. Upstream passes `Fluxes::HLLC<Wavespeeds::Einfeldt>`
// (a function template instance) into SpaceDisc as a std::function (include/SpaceDisc.h:22,
// include/Fluxes.h:14,56, src/Fluxes.cpp:5-26). A std::function cannot run on the device, so the
// same spelling here names a compile-time tag that selects the matching device kernel
// instantiation (k_flux<FLUX, WS> in csrc/swe_kernels.cuh). New fluxes are added by registering
// another instantiation there and another tag here.
#pragma once
#include <string>

#include "../swe_b200.h"
#include "Exceptions.h"

struct Fluxer {
    swe_flux flux;
    swe_wavespeed wavespeed;
    int32_t id = -1;  // registry id of a flux that is not one of the six built-in ones (-1: use flux / wavespeed)
    int32_t RegistryId() const { return id >= 0 ? id : 3 * (int32_t)flux + (int32_t)wavespeed; }
};

namespace Wavespeeds {
struct Rusanov { static constexpr swe_wavespeed id = SWE_RUSANOV; };
struct Davis { static constexpr swe_wavespeed id = SWE_DAVIS; };
struct Einfeldt { static constexpr swe_wavespeed id = SWE_EINFELDT; };
}  // namespace Wavespeeds

namespace Fluxes {
template <class W> inline constexpr Fluxer HLL{SWE_HLL, W::id, -1};
template <class W> inline constexpr Fluxer HLLC{SWE_HLLC, W::id, -1};
// any flux of the device registry (csrc/swe_flux_registry.cuh, csrc/user_fluxes.cuh) by name,
// e.g. Fluxes::Registered("LocalLaxFriedrichs"); throws DomainError if there is no such flux
inline Fluxer Registered(const char *name) {
    const int32_t id = swe_fluxer_find(name);
    if (id < 0) throw DomainError(std::string("flux not registered: ") + name);
    return Fluxer{SWE_HLLC, SWE_EINFELDT, id};
}
}  // namespace Fluxes
