/* swe_constants.h — the numeric constants of the SWE_FVM time step, stated once for the product side
 * (device kernels, host mesh/case builders, C++ mirror). Each is a literal of the upstream tree; the CPU
 * checker in the test tree restates them independently from the same citations and is pinned to upstream's
 * compiled sources, so a wrong value here fails the parity tests. C and C++. */
#ifndef SWE_CONSTANTS_H
#define SWE_CONSTANTS_H

#define SWE_TOL 1e-13            /* tol, include/Includes.h:30 */
#define SWE_WET_DEPTH 1e-12      /* IsWet(h): h > 1e-12, include/Bathymetry.h:5-8 */
#define SWE_DAMP_DEPTH 1e-3      /* velocities are damped below this depth, src/Assigners.cpp:13-17,33-37 */
#define SWE_DAMP_EPS_PRIM 1e-6   /* sqrt(2) h / sqrt(h^2 + 1e-6), PrimAssigner, src/Assigners.cpp:14 */
#define SWE_DAMP_EPS_CONS 1e-12  /* sqrt(2) h / sqrt(h^4 + 1e-12), ConsAssigner, src/Assigners.cpp:35 */
#define SWE_FLUX_DRY_SUM 1e-10   /* hl + hr <= 1e-10 and ar - al <= 1e-10 early-outs, include/Fluxes.h:27,37,69 */
#define SWE_CFL 0.15             /* CFLdt() = 0.15 * min_length_to_wavespeed, include/TimeDisc.h:13,22 */

#endif
