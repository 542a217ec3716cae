/*
 * swe_b200.h — C-ABI of the B200-native SWE_FVM time step (the drop-in boundary).
 *
 * Plain C: pointers, sizes, enums. No C++/torch types cross this boundary. All functions
 * return 0 on success or a negative swe_status; none throws. The C++17 host API in
 * include/swe/ (SpaceDisc / TimeDisc / Solvers::*) forwards to these entry points and
 * re-raises errors as the reference's exception types.
 *
 * Each entry point names the reference interface it replaces (paths relative to the
 * upstream SWE_FVM tree):
 *
 *   swe_mesh                <- Topology ctor arguments + Domain geometry
 *                              (include/TriangMesh.h:27-33, include/Bathymetry.h:14,
 *                               pybind/Topology.cpp:12-30)
 *   swe_create              <- SpaceDisc::SpaceDisc (include/SpaceDisc.h:24, src/SpaceDisc.cpp:4-13)
 *   swe_set_state/get_state <- MUSCLObject::GetVolField (include/MUSCLObject.h:6-7); layout of
 *                              Storage<3> (include/Includes.h:26-27): 3 x Nt column-major (w,u,v)
 *   swe_step                <- Solvers::Euler/SSPRK2/SSPRK3 (include/Solvers.h:6-8, src/Solvers.cpp)
 *                              with the Fluxer plug-in (include/SpaceDisc.h:22) selected by enum:
 *                              Fluxes::HLL<W>/HLLC<W> (include/Fluxes.h:14,56),
 *                              W in Wavespeeds::{Rusanov,Davis,Einfeldt} (src/Fluxes.cpp:5-26)
 *   swe_cfl_dt              <- TimeDisc::CFLdt (include/TimeDisc.h:13)
 *   swe_compute_interface_values <- SpaceDisc::ComputeInterfaceValues (src/SpaceDisc.cpp:33-52)
 *   swe_compute_fluxes      <- SpaceDisc::ComputeFluxes (src/SpaceDisc.cpp:54-74)
 *   swe_get_edge_states     <- SpaceDisc::GetEdgField (include/SpaceDisc.h:26), EdgeIndexer order
 *                              (include/ValueField.h:70-75)
 *   swe_get_sources         <- SpaceDisc::GetSrcField (include/SpaceDisc.h:27)
 *   swe_get_fluxes          <- SpaceDisc::GetFluxes (include/SpaceDisc.h:28)
 *   swe_get_min_len_to_wavespeed <- SpaceDisc::GetMinLenToWavespeed (include/SpaceDisc.h:31)
 *   swe_get_node_max_w      <- MUSCLObject::m_max_wp (include/MUSCLObject.h:68)
 *   swe_get_draining_dt     <- TimeDisc::ComputeDrainingDt (src/TimeDisc.cpp:43-66)
 *   swe_hostmesh_struct     <- StructTriangMesh(ni,nj,h) (include/StructTriangMesh.h:4-15; body
 *                              missing upstream, conventions of notebooks/topology.dat)
 *   swe_hostmesh_gmsh       <- TriangMesh(filename) Gmsh reader (examples/Main.cpp:174; body missing)
 *   swe_case_*              <- Test::{b,u,v,h} (examples/Tests.h:20-27) + the IC loops of
 *                              examples/Main.cpp:202-223,317-337 (TriangAverage,
 *                              include/PointOperations.h:20-44)
 */
#ifndef SWE_B200_H
#define SWE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define SWE_API
#else
#define SWE_API __attribute__((visibility("default")))
#endif

typedef enum swe_status {
    SWE_OK = 0,
    SWE_ERR_INVALID = -1,   /* bad argument / unsupported mesh (DomainError upstream)   */
    SWE_ERR_CUDA = -2,      /* CUDA runtime failure, incl. "no device"                  */
    SWE_ERR_IO = -3,        /* mesh file could not be read (MeshError upstream)         */
    SWE_ERR_NUMERIC = -4,   /* non-finite state detected on device (SolverError)        */
    SWE_ERR_NOMEM = -5
} swe_status;

typedef enum swe_scheme { SWE_EULER = 0, SWE_SSPRK2 = 1, SWE_SSPRK3 = 2 } swe_scheme;
typedef enum swe_flux { SWE_HLL = 0, SWE_HLLC = 1 } swe_flux;
typedef enum swe_wavespeed { SWE_RUSANOV = 0, SWE_DAVIS = 1, SWE_EINFELDT = 2 } swe_wavespeed;

/* Boundaries enum of include/Includes.h:21. Only SOLID_WALL is implemented upstream
 * (src/SpaceDisc.cpp:66-72); swe_create rejects the others. */
enum { SWE_SOLID_WALL = -1, SWE_FREE_FLOW = -2, SWE_PERIODIC = -3, SWE_CUSTOM = -4 };

/* Mesh + bathymetry exactly as the reference's Topology/Domain take them. All arrays are
 * HOST pointers, copied during swe_create (the caller may free them right after).
 * Local ordering (the convention of upstream's mesh builders, notebooks/topology.dat): for every cell t and
 * k = 0,1,2, element_edges[t][k] joins element_nodes[t][k] and element_nodes[t][(k+1)%3], and
 * element_neighbours[t][k] is the cell across that edge (or a negative boundary tag). swe_create checks this
 * and returns SWE_ERR_INVALID otherwise (upstream itself would accept any local order because it looks the
 * edge points up per edge; the kernels here take the edge midpoints from the cell's own nodes). */
typedef struct swe_mesh {
    int64_t nn, ne, nt;
    const double *geometry;             /* 3 x nn column-major: (x, y, b) per node        */
    const int64_t *edge_nodes;          /* ne x 2 row-major, EdgePoints                    */
    const int64_t *edge_elements;       /* ne x 2 row-major, EdgeTriangs; [1] < 0 = wall   */
    const int64_t *element_nodes;       /* nt x 3 row-major, TriangPoints (CCW)            */
    const int64_t *element_edges;       /* nt x 3 row-major, TriangEdges                   */
    const int64_t *element_neighbours;  /* nt x 3 row-major, TriangTriangs; < 0 = boundary */
    double cor;                         /* Coriolis parameter (SpaceDisc ctor)             */
    double tau;                         /* friction parameter: accepted, unused upstream   */
} swe_mesh;

/* ------------------------------------------------------------------------------------ */
/* Device context = SpaceDisc + TimeDisc state on one GPU                                */
/* ------------------------------------------------------------------------------------ */
typedef struct swe_ctx swe_ctx;

/* reorder: 0 = keep caller numbering on device, 1 = locality-preserving (Hilbert curve) renumbering
 * of cells/edges/nodes on device. Results and every get/set use the CALLER's numbering. */
SWE_API int swe_create(swe_ctx **out, const swe_mesh *mesh, int device, int reorder);
/* Same, with an ordering class (0..3) per cell: cells of one class get one contiguous range of
 * device ids (inside it: Hilbert order if reorder != 0, caller order otherwise), so that
 * swe_compute_interface_values_class can reconstruct class by class. A multi-GPU driver puts the
 * cells whose stencil touches halo cells in their own class and overlaps the halo exchange with
 * the reconstruction of all the others. */
SWE_API int swe_create_classes(swe_ctx **out, const swe_mesh *mesh, int device, int reorder,
                               const uint8_t *cell_class_nt);
SWE_API void swe_destroy(swe_ctx *ctx);
/* message of the last failure on ctx (ctx may be NULL: last failure of swe_create/hostmesh). */
SWE_API const char *swe_last_error(const swe_ctx *ctx);

/* run all work of ctx on this cudaStream_t (default: the legacy default stream 0). */
SWE_API int swe_set_stream(swe_ctx *ctx, void *cuda_stream);
SWE_API int swe_synchronize(swe_ctx *ctx);

/* Primitive state (w,u,v), 3 x nt column-major, caller numbering. HOST buffers.
 * swe_set_state stores the values verbatim (use swe_case_initial_state or the C++
 * PrimAssigner mirror to apply the dry clamp first). */
SWE_API int swe_set_state(swe_ctx *ctx, const double *prim_3xnt);
SWE_API int swe_get_state(swe_ctx *ctx, double *prim_3xnt);
/* same, asynchronous on the ctx stream (buffers should be pinned). */
SWE_API int swe_set_state_async(swe_ctx *ctx, const double *prim_3xnt);
SWE_API int swe_get_state_async(swe_ctx *ctx, double *prim_3xnt);

/* Host-buffer pipeline: a stream of independent states (ensemble members, batches), one time step each. Each call
 * enqueues upload of host_in (3 x nt, should be pinned) -> one step of size dt -> download into host_out and returns
 * at once; the upload of the next batch and the download of the previous one overlap the step of the current one
 * (three streams, double-buffered staging: steady-state cost per batch = max(H2D, step, D2H)). swe_wait_host blocks
 * until every submitted batch has landed in its host_out (and reports SWE_ERR_NUMERIC like swe_synchronize). */
SWE_API int swe_submit_step_host(swe_ctx *ctx, const double *host_in_3xnt, double *host_out_3xnt, swe_scheme scheme,
                                 swe_flux flux, swe_wavespeed ws, double dt);
SWE_API int swe_wait_host(swe_ctx *ctx);

/* One time step of the chosen scheme with a fixed dt (like every reference driver). */
SWE_API int swe_step(swe_ctx *ctx, swe_scheme scheme, swe_flux flux, swe_wavespeed ws, double dt);
/* nsteps steps without host synchronisation. dt > 0: fixed dt. dt <= 0: adaptive, every
 * step uses dt = CFLdt() of the previous step's last stage (device-resident, no read-back);
 * the first adaptive step uses dt0, or with dt0 <= 0 the dt already stored on the device by the
 * previous swe_run / swe_checkpoint_load (restart). */
SWE_API int swe_run(swe_ctx *ctx, swe_scheme scheme, swe_flux flux, swe_wavespeed ws,
                    int64_t nsteps, double dt, double dt0);
/* 0.15 * min(1, min over edges of length/wavespeed) of the last flux evaluation. */
SWE_API int swe_cfl_dt(swe_ctx *ctx, double *dt);
SWE_API int swe_get_min_len_to_wavespeed(swe_ctx *ctx, double *v);
/* simulated time accumulated by swe_step/swe_run since the last swe_set_state. */
SWE_API int swe_get_time(swe_ctx *ctx, double *t);
/* number of kernels launched by this ctx so far (bench.py's gpu_launches). */
SWE_API int64_t swe_launch_count(const swe_ctx *ctx);

/* Per-kernel timing with CUDA events on the ctx stream (bench.py's roofline numbers).
 * swe_kernel_timing(ctx, 1) starts (and clears) the recording, swe_kernel_times returns the
 * number of kernel kinds K (<= max_kinds) and fills total milliseconds, launch counts and
 * static name strings per kind. */
SWE_API int swe_kernel_timing(swe_ctx *ctx, int enable);
SWE_API int swe_kernel_times(swe_ctx *ctx, int32_t max_kinds, double *ms_total, int64_t *counts,
                             const char **names);

/* The two halves of a stage, callable on their own (parity taps, custom TimeDisc loops). */
SWE_API int swe_compute_interface_values(swe_ctx *ctx);
SWE_API int swe_compute_fluxes(swe_ctx *ctx, swe_flux flux, swe_wavespeed ws);
/* ComputeInterfaceValues split by cell range [first_cell, last_cell) so that a multi-GPU driver
 * can reconstruct the interior while the halo is still in flight and the boundary rows after it
 * arrived. begin != 0 on the first range of a stage, finish != 0 on the last one (runs the
 * part-wet second pass, which needs every first-pass result). Caller numbering; requires a
 * context created with reorder = 0. */
SWE_API int swe_compute_interface_values_range(swe_ctx *ctx, int64_t first_cell, int64_t last_cell,
                                               int begin, int finish);
/* the same for all cells of one ordering class (see swe_create_classes) */
SWE_API int swe_compute_interface_values_class(swe_ctx *ctx, int32_t cell_class, int begin, int finish);
/* stage update: cons(i) = a0*U0.cons(i) + a1*cons(i) + RHS(i, dt_stage), U0 = state saved by
 * swe_save_state(). Euler: a0=0,a1=1. Uses the fluxes of the last swe_compute_fluxes. */
SWE_API int swe_save_state(swe_ctx *ctx);
SWE_API int swe_stage_update(swe_ctx *ctx, double a0, double a1, double dt_stage);
/* same with dt_stage = coef * (device-resident dt): lets a multi-GPU driver step with the
 * globally reduced CFL dt without reading it back. swe_set_dt stores dt on the device;
 * swe_advance_dt does time += dt and, if adaptive, dt = 0.15 * min_len_to_wavespeed. */
SWE_API int swe_stage_update_dev(swe_ctx *ctx, double a0, double a1, double coef);
SWE_API int swe_set_dt(swe_ctx *ctx, double dt);
SWE_API int swe_advance_dt(swe_ctx *ctx, int adaptive, double dt_fixed);
/* keep the reconstructed edge-side w as well (needed only by swe_get_edge_states) and count branch hits. */
SWE_API int swe_enable_taps(swe_ctx *ctx, int on);

/* Flux registry: upstream's plug-in point (i), SpaceDisc::Fluxer (include/SpaceDisc.h:22). Fluxes are device
 * functors registered at compile time (csrc/swe_flux_registry.cuh, csrc/user_fluxes.cuh: one functor + one line).
 * Entries 0..5 are Fluxes::HLL/HLLC<Wavespeeds::Rusanov/Davis/Einfeldt> with id = 3 * swe_flux + swe_wavespeed;
 * swe_set_fluxer(ctx, id) makes swe_step / swe_run / swe_compute_fluxes / swe_dist_* use that flux regardless
 * of their (flux, wavespeed) arguments; id < 0 switches back. */
SWE_API int32_t swe_fluxer_count(void);
SWE_API const char *swe_fluxer_name(int32_t k);   /* k-th registry entry */
SWE_API int32_t swe_fluxer_id(int32_t k);
SWE_API int32_t swe_fluxer_find(const char *name); /* id, or -1 */
SWE_API int swe_set_fluxer(swe_ctx *ctx, int32_t id);

/* Switches for the places where upstream HEAD is unfinished. Defaults = the
 * repaired scheme; the alternatives reproduce upstream exactly as written and are bit-checked against
 * upstream's own sources compiled by the test tree (tests/test_ref_anchor.py, tests/test_gpu_parity.py):
 *   "recon"   0 plane gradients of w,u,v from grad_values (S2, default) | 1 as written:
 *             src/MUSCLObject.cpp:63-64 df.row(0) = Gradient(grad_points), u/v first order | 2 first order
 *   "pw2"     0 ReconstructPartWetCell2 reads points(r,c) as (point, coordinate) (S3, default) |
 *             1 as written (src/MUSCLObject.cpp:141-183)
 *   "roe_fix" 0 Einfeldt Roe velocity cl*ur as written (src/Fluxes.cpp:22, default) | 1 cr*ur
 *   "cfl_abs" 0 HLLC CFL candidate max(tol, max(al, ar)) as written (include/Fluxes.h:89, default) | 1 magnitudes
 * Tuning switches (identical bits either way; A/B evidence in profiles/README.md):
 *   "graph"       -1 replay whole steps as a CUDA graph on meshes below 4M cells (default) | 0 off | 1 on
 *   "k1_tiled"    0 register-prefetched gather reconstruction (default) | 1 TMA-staged shared-memory tiles
 *   "fused_drain" 0 separate draining-dt pass (default) | 1 draining dt computed inside the stage update
 *                 (swe_get_draining_dt then needs swe_enable_taps or swe_compute_rhs)
 *   "skip_cfl"    1 only the last stage of a step rebuilds the CFL minimum (default; the earlier ones are dead values
 *                 upstream too) | 0 every stage
 *   "dry_skip"    -1 auto (default): while >= 20 % of the cells are dry, tiles of 128 cells that are all dry, stored
 *                 as (b, +0, +0) and surrounded by dry cells are skipped by the flux / draining-dt / update kernels and
 *                 by the stores of the reconstruction | 0 off | 1 on
 *   "dry_list"    -1 auto: on meshes of >= 4M cells the dry-region stage update runs over a compacted list of the tiles
 *                 it has to process | 0 every block tests its own tile flag | 1 always the list */
SWE_API int swe_set_option(swe_ctx *ctx, const char *key, int32_t value);
SWE_API int swe_get_option(swe_ctx *ctx, const char *key, int32_t *value);
/* Debug tap (taps enabled): how many cells took each branch in the last swe_compute_interface_values.
 * [0..2] pass-1 ReconstructPartWetCell1 of a part-wet cell: submerged / closed form (cbrt) / bisection
 * (src/MUSCLObject.cpp:100-108); [3..6] full-wet cells: dry neighbour (:52-53) / part-wet neighbour
 * (:54-59) / gradient zeroed by the vertex check (:70-72) / a TVD component switched off (:80);
 * [7..11] pass-2 ReconstructPartWetCell2: early PartWet1 (:127) / 1 vertex wet (:145) / 3 wet (:153) /
 * 2 wet (:164) / 2-wet fall-back k1 < tol (:170). */
SWE_API int swe_get_branch_counts(swe_ctx *ctx, int64_t out12[12]);

/* Parity taps. HOST output buffers, caller numbering, reference layouts. */
SWE_API int swe_get_edge_states(swe_ctx *ctx, double *edg_3x2ne); /* col = 2e + (from<to)  */
SWE_API int swe_get_sources(swe_ctx *ctx, double *src_3x2ne);     /* row 0 is zero (unused) */
SWE_API int swe_get_fluxes(swe_ctx *ctx, double *f_3xne);
SWE_API int swe_get_node_max_w(swe_ctx *ctx, double *maxw_nn);
SWE_API int swe_get_draining_dt(swe_ctx *ctx, double *dti_nt);    /* of the last stage      */
SWE_API int swe_get_cell_class(swe_ctx *ctx, int8_t *cls_nt);     /* 0 dry 1 part 2 full    */

/* Per-cell accessors of the reference API, as whole-array taps (HOST buffers, caller numbering):
 * swe_classify    <- MUSCLObject::IsDryCell / IsFullWetCell / IsPartWetCell (include/MUSCLObject.h:11-13) of the
 *                    CURRENT state: 0 dry, 1 part-wet, 2 full-wet
 * swe_compute_rhs <- TimeDisc::RHS(i, dt) for every i (include/TimeDisc.h:15, src/TimeDisc.cpp:3-41), 3 x nt,
 *                    for the edge values / fluxes of the last swe_compute_interface_values + swe_compute_fluxes;
 *                    also refreshes swe_get_draining_dt (TimeDisc::ComputeDrainingDt). Snapshot semantics (S7):
 *                    every RHS sees the same state. The state itself is not changed. */
SWE_API int swe_classify(swe_ctx *ctx, int8_t *cls_nt);
SWE_API int swe_compute_rhs(swe_ctx *ctx, double dt, double *rhs_3xnt);
SWE_API int swe_set_time(swe_ctx *ctx, double t);
SWE_API int swe_get_dt(swe_ctx *ctx, double *dt); /* the device-resident dt of adaptive runs */

/* Binary checkpoint / restart: state (caller numbering), simulated time, the device-resident dt and
 * min_len_to_wavespeed (so an adaptive run continues bit-identically) and the solver settings, which must match
 * on load (SWE_ERR_INVALID otherwise). A checkpoint can be loaded into a context with another device numbering. */
SWE_API int swe_checkpoint_save(swe_ctx *ctx, const char *path);
SWE_API int swe_checkpoint_load(swe_ctx *ctx, const char *path);

/* Device reductions (diagnostics; the commented ComputeIntegrals of src/SpaceDisc.cpp:77-104):
 * out[0] = sum A_i h_i (mass), out[1] = sum A_i 0.5 h (u^2+v^2), out[2] = sum A_i (0.5 h^2 + h b),
 * out[3] = max |u|,|v|, out[4] = min h, out[5] = number of wet cells. Deterministic
 * (fixed-shape tree). */
SWE_API int swe_diagnostics(swe_ctx *ctx, double out[6]);

/* ------------------------------------------------------------------------------------ */
/* Multi-GPU support: a rank owns a sub-mesh with halo cells   */
/* ------------------------------------------------------------------------------------ */
/* CFL min only over edges with cfl_mask[e] != 0 (edges touching an owned cell). NULL = all. */
SWE_API int swe_set_cfl_edge_mask(swe_ctx *ctx, const uint8_t *mask_ne);
/* Register gather/scatter lists (caller numbering of local cells). */
SWE_API int swe_halo_set_lists(swe_ctx *ctx, int64_t nsend, const int64_t *send_cells,
                               int64_t nrecv, const int64_t *recv_cells);
/* pack: sendbuf[3k + c] = prim[c] of send_cells[k] (Storage<3> layout, so each peer's
 * segment is contiguous); unpack: the inverse into recv_cells. Buffers are DEVICE pointers
 * (NCCL or peer-mapped buffers). */
SWE_API int swe_halo_pack(swe_ctx *ctx, double *dev_sendbuf);
SWE_API int swe_halo_unpack(swe_ctx *ctx, const double *dev_recvbuf);
/* Peer-memory halo transport: the pack kernel stores this rank's boundary states straight into the
 * neighbour GPU's receive buffer over NVLink (CUDA IPC mapping) and publishes the exchange number
 * in the neighbour's flag slot; the receiver waits on its flags (with a time-out) and unpacks. No
 * NCCL call on the data path. Set-up: swe_halo_set_lists, then swe_halo_p2p_alloc (returns three
 * 64-byte IPC handles: receive buffer 0, receive buffer 1, flags), exchange of the handles and of
 * the segment offsets by the host layer, one swe_halo_p2p_connect per peer. Per exchange:
 * swe_halo_p2p_push then swe_halo_p2p_pull (both asynchronous on the ctx stream). The flag slot
 * of peer k is k (its position in this rank's peer list). */
SWE_API int swe_halo_p2p_alloc(swe_ctx *ctx, int32_t npeers, unsigned char *handles_3x64);
SWE_API int swe_halo_p2p_connect(swe_ctx *ctx, int64_t send_start, int64_t send_count,
                                 const unsigned char *peer_handles_3x64, int64_t dst_offset_cells,
                                 int32_t my_slot_in_peer_flags);
SWE_API int swe_halo_p2p_push(swe_ctx *ctx);
SWE_API int swe_halo_p2p_pull(swe_ctx *ctx);
SWE_API int swe_halo_p2p_error(swe_ctx *ctx); /* 1 if a wait timed out */
/* replace the running min_len_to_wavespeed (after the global min all-reduce). */
SWE_API int swe_set_min_len_to_wavespeed(swe_ctx *ctx, double v);
/* device address of the fp64 scalar holding min_len_to_wavespeed (for in-place NCCL all-reduce). */
SWE_API int swe_min_len_device_ptr(swe_ctx *ctx, void **dev_ptr);

/* ------------------------------------------------------------------------------------ */
/* Host-side mesh construction (no GPU needed)                                            */
/* ------------------------------------------------------------------------------------ */
typedef struct swe_hostmesh swe_hostmesh;

/* StructTriangMesh(ni, nj, h): [0,ni*h] x [0,nj*h], every square split by both diagonals
 * into Bottom/Right/Top/Left triangles around a centre node; 4*ni*nj cells. i0/j0 shift
 * the block inside a larger global grid (node coordinates are (i0+i)*h, bitwise equal to
 * the global mesh), which is how a rank builds its strip of a larger grid directly. */
SWE_API int swe_hostmesh_struct(swe_hostmesh **out, int64_t ni, int64_t nj, double h,
                                int64_t i0, int64_t j0);
/* Gmsh ASCII 4.1/4.2 reader reproducing the reference's numbering (notebooks/topology.dat). */
SWE_API int swe_hostmesh_gmsh(swe_hostmesh **out, const char *path);
/* build from raw triangles (+ optional boundary line list), same numbering rules. */
SWE_API int swe_hostmesh_from_triangles(swe_hostmesh **out, int64_t nn, const double *xy_2xnn,
                                        int64_t nt, const int64_t *tri_ntx3,
                                        int64_t nb, const int64_t *bnd_nbx2);
/* uniform 1 -> 4 refinement (edge midpoints); bathymetry of new nodes = mean of the ends. */
SWE_API int swe_hostmesh_refine(swe_hostmesh **out, const swe_hostmesh *in);
SWE_API void swe_hostmesh_free(swe_hostmesh *m);
/* borrow views into m (valid until free); cor/tau are left 0. geometry is writable through
 * swe_hostmesh_geometry so the caller can set the nodal bathymetry (row 2). */
SWE_API int swe_hostmesh_view(const swe_hostmesh *m, swe_mesh *view);
SWE_API double *swe_hostmesh_geometry(swe_hostmesh *m);

/* Sub-mesh extraction for domain decomposition: cells with part[i] == rank plus `layers`
 * rings of vertex-adjacent halo cells, in increasing global id (orientation preserving).
 * Outputs (malloc'ed inside the returned hostmesh, borrowed): global ids of local cells,
 * owner rank of each local cell. */
SWE_API int swe_hostmesh_extract(swe_hostmesh **out, const swe_hostmesh *global,
                                 const int32_t *part_nt, int32_t rank, int32_t layers);
SWE_API const int64_t *swe_hostmesh_global_cells(const swe_hostmesh *m); /* nt, or NULL */
SWE_API const int32_t *swe_hostmesh_cell_owner(const swe_hostmesh *m);   /* nt, or NULL */
/* recursive coordinate bisection of cell centroids into nparts (any nparts >= 1). */
SWE_API int swe_partition_rcb(const swe_hostmesh *m, int32_t nparts, int32_t *part_nt);

/* ------------------------------------------------------------------------------------ */
/* Multi-GPU time step behind the C-ABI: decomposition plan (host) + distributed context    */
/* ------------------------------------------------------------------------------------ */
/* The reference steps one SpaceDisc with Solvers::X(TimeDisc*, dt) (include/Solvers.h:6-8); swe_dist_step /
 * swe_dist_run are the same call on N GPUs. Everything multi-GPU lives below this boundary: the
 * decomposition (strips of a StructTriangMesh, or any mesh + partition vector), the per-rank device
 * contexts with halo cells, the peer-memory halo exchange fused into the stage (pack-and-signal right
 * after the boundary cells are updated, wait-and-unpack before the halo-dependent reconstruction), and
 * the global CFL minimum over peer memory. A launcher only provides ONE bootstrap primitive, an
 * all-gather of a small fixed-size blob (swe_allgather_fn), used during creation to exchange CUDA-IPC
 * handles; no launcher code runs on the data path. One process per GPU (swe_dist_create_*) or one
 * process driving several GPUs (swe_dist_group_create_*, no bootstrap needed). Results on owned cells
 * are bit-identical to the single-GPU run for any GPU count. */
typedef struct swe_dist_plan swe_dist_plan; /* host only: who owns what, halo lists (no GPU needed)      */
typedef struct swe_dist swe_dist;           /* one rank: plan + device context + peer mappings           */

/* rank r owns an even share of the nj rows of squares of StructTriangMesh(ni, nj, h) plus 3 halo rows per
 * open side; the local block is generated directly (no global mesh in memory). */
SWE_API int swe_dist_plan_struct(swe_dist_plan **out, int32_t rank, int32_t world, int64_t ni, int64_t nj, double h);
/* any mesh + partition vector (e.g. swe_partition_rcb); every rank passes the same inputs. 4 vertex rings. */
SWE_API int swe_dist_plan_mesh(swe_dist_plan **out, int32_t rank, int32_t world, const swe_hostmesh *global,
                               const int32_t *part_nt);
SWE_API void swe_dist_plan_free(swe_dist_plan *plan);
SWE_API const swe_hostmesh *swe_dist_plan_local_mesh(const swe_dist_plan *plan);   /* owned + halo cells   */
SWE_API int64_t swe_dist_plan_owned_count(const swe_dist_plan *plan);
SWE_API const uint8_t *swe_dist_plan_owned(const swe_dist_plan *plan);             /* per local cell       */
SWE_API const int64_t *swe_dist_plan_global_cells(const swe_dist_plan *plan);      /* local -> global id   */
SWE_API const uint8_t *swe_dist_plan_classes(const swe_dist_plan *plan);           /* ordering class 0..3  */
SWE_API const uint8_t *swe_dist_plan_cfl_mask(const swe_dist_plan *plan);          /* per local edge       */
SWE_API int32_t swe_dist_plan_npeers(const swe_dist_plan *plan);
SWE_API int swe_dist_plan_peer(const swe_dist_plan *plan, int32_t k, int32_t *peer_rank, int64_t *nsend,
                               const int64_t **send_local, int64_t *nrecv, const int64_t **recv_local);

/* bootstrap: gather `bytes` bytes from every rank into recv (world * bytes, rank order). 0 = success. */
typedef int (*swe_allgather_fn)(void *user, const void *send, void *recv, int64_t bytes);
typedef struct swe_dist_config {
    int32_t device;        /* CUDA device ordinal of this rank                                          */
    int32_t reorder;       /* 1: Hilbert numbering inside every ordering class (default in the drivers)  */
    int32_t overlap;       /* 1: the exchange overlaps the interior reconstruction / update (default)    */
    int32_t reserved;
    double cor, tau;       /* SpaceDisc ctor arguments                                                   */
    double wait_timeout_s; /* peer waits give up after this long and raise SWE_ERR_CUDA (0: 60 s)        */
    swe_allgather_fn allgather;
    void *user;
} swe_dist_config;

/* one process per GPU: every rank calls this collectively with its own plan (takes ownership of it) */
SWE_API int swe_dist_create(swe_dist **out, swe_dist_plan *plan, const swe_dist_config *cfg);
/* one process, `world` GPUs: plans[r] / devices[r] per rank; out[r] receives the rank objects */
SWE_API int swe_dist_group_create(swe_dist **out_world, swe_dist_plan **plans_world, const int32_t *devices_world,
                                  int32_t world, const swe_dist_config *cfg);
SWE_API void swe_dist_destroy(swe_dist *d);
SWE_API const char *swe_dist_last_error(const swe_dist *d);
SWE_API swe_ctx *swe_dist_ctx(swe_dist *d);                 /* the rank's device context (local numbering)   */
SWE_API const swe_dist_plan *swe_dist_get_plan(const swe_dist *d);
/* fill the halo cells from their owners (after swe_set_state / device-side initial conditions) */
SWE_API int swe_dist_exchange(swe_dist *d);
/* Solvers::Euler/SSPRK2/SSPRK3 on all ranks; asynchronous on each rank's stream. dt <= 0 in swe_dist_run:
 * every step uses 0.15 * GLOBAL min_len_to_wavespeed of the previous step (device resident); first step dt0. */
SWE_API int swe_dist_step(swe_dist *d, swe_scheme scheme, swe_flux flux, swe_wavespeed ws, double dt);
SWE_API int swe_dist_run(swe_dist *d, swe_scheme scheme, swe_flux flux, swe_wavespeed ws, int64_t nsteps, double dt,
                         double dt0);
/* group form: issues every step for all ranks of one process in turn */
SWE_API int swe_dist_group_run(swe_dist **ranks, int32_t world, swe_scheme scheme, swe_flux flux, swe_wavespeed ws,
                               int64_t nsteps, double dt, double dt0);
/* the host-buffer pipeline of swe_submit_step_host on a rank: LOCAL state (owned + halo cells) in and out */
SWE_API int swe_dist_submit_step_host(swe_dist *d, const double *host_in_3xntlocal, double *host_out_3xntlocal,
                                      swe_scheme scheme, swe_flux flux, swe_wavespeed ws, double dt);
SWE_API int swe_dist_wait_host(swe_dist *d);
/* waits for the rank's stream; SWE_ERR_NUMERIC on a non-finite state, SWE_ERR_CUDA if a peer wait timed out */
SWE_API int swe_dist_synchronize(swe_dist *d);
SWE_API int swe_dist_cfl_dt(swe_dist *d, double *dt);     /* 0.15 * global min (after a step)               */
/* order-independent 64-bit hash of this rank's OWNED cell states keyed by global cell id: the sum over the
 * ranks (mod 2^64) equals swe_state_hash of the undecomposed run iff every cell is bit-identical */
SWE_API int swe_dist_state_hash(swe_dist *d, uint64_t *partial);
SWE_API int swe_state_hash(swe_ctx *ctx, uint64_t *hash);
/* owned cell states into a GLOBAL 3 x nt_global array (only this rank's owned columns are written) */
SWE_API int swe_dist_get_owned_state(swe_dist *d, double *prim_3xnt_global);
/* set the local state (owned + halo) from a GLOBAL 3 x nt_global array */
SWE_API int swe_dist_set_state_global(swe_dist *d, const double *prim_3xnt_global);

/* ------------------------------------------------------------------------------------ */
/* Analytic test cases (examples/Tests.h) — bathymetry and initial conditions             */
/* ------------------------------------------------------------------------------------ */
typedef enum swe_case_kind {
    SWE_CASE_LAKE_AT_REST = 0,    /* LakeAtRestTest, examples/Tests.h:32-43                 */
    SWE_CASE_CLASSIC_THACKER = 1, /* ClassicThackerTest, examples/Tests.h:237-280           */
    SWE_CASE_GAUSS_WAVE = 2,      /* testGaussWave, examples/Main.cpp:172-195 (flat bed)    */
    SWE_CASE_FULLY_WET = 3,       /* synthetic fully-wet variant               */
    SWE_CASE_BOWL_HUMP = 4        /* BowlTest bed (examples/Tests.h:46-57) + Gaussian hump  */
} swe_case_kind;

typedef struct swe_case {
    int32_t kind;
    double mid_x, mid_y; /* centre of the domain feature                                   */
    double length;       /* domain side l (FULLY_WET bed wavelength)                       */
    double cor, tau;     /* [Common] cor, tau                                              */
    double delta;        /* [Common] delta                                                 */
    double H0, p0, q0;   /* [Thacker]                                                      */
    double level, amp;   /* BOWL_HUMP still-water level; Gaussian hump amplitude           */
} swe_case;

SWE_API void swe_case_defaults(swe_case *c, int32_t kind, double mid_x, double mid_y, double length);
/* exact solution at a point: out = (b, h, u, v). */
SWE_API int swe_case_eval(const swe_case *c, double x, double y, double t, double out[4]);
/* nodal bathymetry: geometry row 2 <- case.b(x,y). */
SWE_API int swe_case_set_bathymetry(const swe_case *c, swe_hostmesh *m);
/* cell initial state like examples/Main.cpp:211-223: (h,u,v) averaged by TriangAverage<3,n>
 * at time t, then w = h_avg + b_i, then the PrimAssigner dry clamp. GAUSS_WAVE instead
 * samples w at the centroid like examples/Main.cpp:183-186. out: 3 x nt column-major. */
SWE_API int swe_case_initial_state(const swe_case *c, const swe_hostmesh *m, int32_t quad_n,
                                   double t, double *prim_3xnt);

/* The same on the device (no host loop, for 10^7-10^8 cells): nodal bathymetry b(x,y) written
 * into the context's node array (and the bed-dependent cell geometry refreshed), the cell state
 * built by TriangAverage<3, quad_n> on the device, and the L2 error of (h, hu, hv) against the
 * exact solution at time t sampled at the centroids (upstream's commented CompareWith,
 * src/SpaceDisc.cpp:106-138). Device libm differs from glibc in the last ulp, so device-built
 * states are inputs in their own right. */
SWE_API int swe_case_set_bathymetry_device(swe_ctx *ctx, const swe_case *c);
SWE_API int swe_case_initial_state_device(swe_ctx *ctx, const swe_case *c, int32_t quad_n, double t);
SWE_API int swe_case_l2_error(swe_ctx *ctx, const swe_case *c, double t, double out[3]);

SWE_API const char *swe_version(void);
SWE_API int32_t swe_device_count(void); /* CUDA devices visible to this process (0 without a GPU) */

#ifdef __cplusplus
}
#endif
#endif /* SWE_B200_H */
