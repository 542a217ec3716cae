// swe_b200.cu — device context and C-ABI entry points of include/swe_b200.h.
//
// A swe_ctx is the device-resident equivalent of the reference's SpaceDisc + TimeDisc
// (include/SpaceDisc.h:20-48, include/TimeDisc.h:4-23): mesh topology and precomputed geometry
// uploaded once as structure-of-arrays (optionally Morton-renumbered), the cell state, the
// edge-side reconstructions, fluxes, node maxima and draining time steps. Every C function
// returns a status; nothing throws across the boundary. There is no CPU fallback.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <numeric>
#include <cstdlib>
#include <string>
#include <thread>
#include <vector>

#include "../../include/swe_b200.h"
#include "hostmesh.hpp"
#include "swe_kernels.cuh"
#include "swe_cases.cuh"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <thrust/iterator/counting_iterator.h>

using namespace swe;

struct swe_ctx {
    int device = 0;
    cudaStream_t stream = 0;
    std::string err;
    int64_t launches = 0;
    int nt = 0, ne = 0, nn = 0;
    int sms = 148;
    double cor = 0, tau = 0;
    bool reordered = false;
    bool taps = false;
    int opt_recon = 0, opt_pw2 = 0, opt_roe_fix = 0, opt_cfl_abs = 0;  // swe_set_option (semantic decisions S2/S3/S5/S6)
    int fluxer = -1;    // registry id selected by swe_set_fluxer (-1: use the (flux, wavespeed) enums of the call)
    int opt_tiled = 0;  // K1 form: 0 = register-prefetched gathers (default, faster: profiles/r2_k1_tiled_vs_gather.md), 1 = TMA-staged tiles
    unsigned long long *dbg = nullptr;                                 // branch-hit counters (taps)
    // fused draining dt (k_update<.., FUSED>): K3 runs only over this static list of cells whose dti is read from
    // global memory by a neighbour in another update tile / ordering class; 0 = separate k_drain pass over all cells.
    // Off by default: measured SLOWER at 64M cells (profiles/README.md, r2_fused_drain_ab.json): 15-25 % of the cells
    // sit on a 128-cell tile border, so the list pass costs 0.62 ms against 0.73 ms for the full pass, and the fused
    // update needs 104 registers + a barrier (2.75 vs 2.23 ms)
    int opt_fused_drain = 0;
    int opt_skip_cfl = 1;
    // dry-region skipping (DevFields::tile_dry): flags written by K1, usable by K2-K4 while they describe the current state
    int opt_dry_skip = -1;      // -1 auto: on while >= 20 % of the cells are dry (re-evaluated after the state was set from outside
                                // and every 1024 steps of a run), 0 off, 1 on. The dry instantiations cost ~5 % on a fully wet mesh
    bool dry_active = false;    // the dry-region instantiations of K1-K4 are in use
    bool dry_eval_pending = true;
    bool k1_dry = false;        // every K1 launch since the last `begin` maintained the flags
    int64_t steps_since_eval = 0;
    unsigned char *tile_dry = nullptr, *tile_dry0 = nullptr, *tile_zero = nullptr;  // [ntiles]: this stage / saved state / all zero
    // flags of the last pass whose flux / draining-dt kernel ran with them: tile flagged there => its edges hold +0 fluxes /
    // its cells hold dt = 0 in memory (copied on the stream right after those kernels, zeroed when they run without flags)
    unsigned char *tile_fluxed = nullptr, *tile_drained = nullptr;
    int *tile_list = nullptr;            // [1 + ntiles] compacted tiles of the dry-region stage update
    int opt_dry_list = -1;               // -1 auto (meshes of >= 4M cells), 0 every block tests its own flag, 1 always the list
    unsigned char *tile_prev = nullptr;  // flags of the previous K1 pass (copied from tile_dry at every `begin`; tile_dry is zeroed on
                                         // the stream whenever it stops describing what the edge-state arrays hold: graph-replay safe)
    int ntiles = 0;
    int64_t state_version = 1, flags_version = 0;  // flags valid iff equal
    bool flags0_valid = false;
    int64_t k1_covered = 0;                        // cells reconstructed since the last `begin`  // non-final stages of a step run the flux kernel without the CFL minimum (a dead value upstream too)
    int *drain_list = nullptr;
    int drain_count = 0;
    bool dti_complete = false;  // c->dti holds the draining dt of EVERY cell (full k_drain ran for the last stage)
    int class_first[6] = {0, 0, 0, 0, 0, 0};  // device cell range of every ordering class
    // device mesh
    int *tt = nullptr, *te = nullptr, *tp = nullptr, *slotL = nullptr, *slotR = nullptr;
    double4 *cgeo = nullptr, *node = nullptr;
    double2 *en = nullptr;
    double *area = nullptr, *cb = nullptr, *elen = nullptr, *dmin = nullptr;
    int *n2c_start = nullptr, *n2c_cells = nullptr;
    unsigned char *cfl_mask = nullptr;
    double *dmin0 = nullptr;  // unmasked copy of dmin (only once a CFL edge mask was set)
    int *cell_old = nullptr, *edge_old = nullptr, *node_old = nullptr;  // device id -> caller id
    std::vector<int> cell_new;  // caller id -> device id (host; halo lists)
    // fields
    double *bufA[3] = {nullptr, nullptr, nullptr}, *bufB[3] = {nullptr, nullptr, nullptr};
    double **cur = nullptr, **sav = nullptr;  // point at bufA / bufB
    double *ceh = nullptr, *ceu = nullptr, *cev = nullptr, *cgx = nullptr, *cgy = nullptr, *cew = nullptr;
    double *f0 = nullptr, *f1 = nullptr, *f2 = nullptr, *dti = nullptr, *pwl = nullptr;
    signed char *cls = nullptr;
    int *pw_list = nullptr, *rs_list = nullptr;
    double *scal = nullptr;
    int *flags = nullptr;
    double *diag = nullptr;  // partials + 6 outputs
    double *stage_aos = nullptr;  // 3*max(nt, 2ne) staging for host transfers
    size_t stage_cap = 0;
    // halo
    int *send_cells = nullptr, *recv_cells = nullptr;
    int nsend = 0, nrecv = 0;
    bool saved_pending = false;
    // peer-memory halo transport (swe_halo_p2p_*)
    struct P2PPeer { int send_start, send_count; double *peer_recv[2]; int *peer_flag; };
    std::vector<P2PPeer> p2p_peers;
    double *p2p_recv[2] = {nullptr, nullptr};  // local receive buffers (parity of the exchange number)
    int *p2p_flags = nullptr;                  // [npeers] sequence numbers written by the peers, [npeers] = error
    int p2p_seq = 0;
    std::vector<void *> p2p_imported;
    // CUDA graphs of whole time steps (swe_run on launch-bound meshes): one instantiated graph per
    // (scheme, flux, adaptive, dt, which state buffer is current); replayed instead of ~13 launches per step
    struct StepGraph {
        int scheme, fluxer, adaptive, parity, opts;
        double dt;
        cudaGraphExec_t exec;
        int64_t launches;
        bool cur_is_a_after;  // host-side state after the step
    };
    std::vector<StepGraph> graphs;
    cudaStream_t gstream = nullptr;  // capture stream when the context runs on the legacy default stream
    cudaEvent_t gev = nullptr;
    int opt_graph = -1;              // -1 auto (meshes below kGraphAutoCells), 0 off, 1 on
    // host-buffer pipeline (swe_submit_step_host): upload of batch n+1 and download of batch n-1 overlap the step of batch n
    struct HostPipe {
        cudaStream_t s_in = nullptr, s_out = nullptr;
        double *in[2] = {nullptr, nullptr}, *out[2] = {nullptr, nullptr};
        cudaEvent_t in_ready[2], in_consumed[2], out_ready[2], out_free[2];
        int64_t n = 0;
        bool ok = false;
    } pipe;
    // optional per-kernel CUDA-event timing (bench.py roofline): pairs recorded on c->stream
    bool ktiming = false;
    struct KtPair { cudaEvent_t a, b; int id; };
    std::vector<KtPair> kt_pairs;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> kt_pool;
};

enum { KT_RECONSTRUCT = 0, KT_PARTWET2, KT_FLUX, KT_DRAIN, KT_UPDATE, KT_HALO_PACK, KT_HALO_WAIT, KT_MIN, KT_COUNT };
static const char *kt_names[KT_COUNT] = {"k_reconstruct", "k_partwet2", "k_flux", "k_drain", "k_update",
                                         "k_halo_pack_signal", "k_halo_wait_unpack", "k_min_push_pull"};
constexpr size_t kKtMaxPairs = 8192;

static inline int kt_begin(swe_ctx *c, int id) {
    if (!c->ktiming || c->kt_pairs.size() >= kKtMaxPairs) return -1;
    swe_ctx::KtPair p;
    if (!c->kt_pool.empty()) { p.a = c->kt_pool.back().first; p.b = c->kt_pool.back().second; c->kt_pool.pop_back(); }
    else { if (cudaEventCreate(&p.a) != cudaSuccess || cudaEventCreate(&p.b) != cudaSuccess) return -1; }
    p.id = id;
    cudaEventRecord(p.a, c->stream);
    c->kt_pairs.push_back(p);
    return (int)c->kt_pairs.size() - 1;
}
static inline void kt_end(swe_ctx *c, int h) {
    if (h >= 0) cudaEventRecord(c->kt_pairs[h].b, c->stream);
}

static thread_local std::string g_create_error;

using swe::host_threads;

#define CUDA_TRY(ctx, call)                                                                      \
    do {                                                                                         \
        cudaError_t e_ = (call);                                                                 \
        if (e_ != cudaSuccess) {                                                                 \
            (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e_);                     \
            return SWE_ERR_CUDA;                                                                 \
        }                                                                                        \
    } while (0)

static inline int nblk(int64_t n, int b) { return (int)((n + b - 1) / b); }

// the compacted tile list of the dry-region stage update costs three small launches per stage: only worth it on big meshes
constexpr int kTileListMinCells = 1 << 22;
static inline bool use_tile_list(const swe_ctx *c) { return c->opt_dry_list < 0 ? c->nt >= kTileListMinCells : c->opt_dry_list == 1; }
static DevMesh dev_mesh(const swe_ctx *c) {
    DevMesh m;
    m.nt = c->nt; m.ne = c->ne; m.nn = c->nn;
    m.tt = c->tt; m.te = c->te; m.tp = c->tp;
    m.cgeo = c->cgeo; m.area = c->area; m.cb = c->cb; m.node = c->node;
    m.n2c_start = c->n2c_start; m.n2c_cells = c->n2c_cells;
    m.slotL = c->slotL; m.slotR = c->slotR; m.en = c->en; m.elen = c->elen; m.dmin = c->dmin;
    return m;
}
static DevFields dev_fields(const swe_ctx *c) {
    DevFields s;
    s.w = c->cur[0]; s.u = c->cur[1]; s.v = c->cur[2];
    s.ceh = c->ceh; s.ceu = c->ceu; s.cev = c->cev; s.cgx = c->cgx; s.cgy = c->cgy; s.cew = c->cew;
    s.f0 = c->f0; s.f1 = c->f1; s.f2 = c->f2; s.dti = c->dti; s.pwl = c->pwl; s.cls = c->cls; s.pw_list = c->pw_list; s.rs_list = c->rs_list;
    s.scal = c->scal; s.flags = c->flags;
    s.dbg = c->dbg; s.recon = c->opt_recon; s.pw2 = c->opt_pw2;
    s.tile_dry = c->tile_dry;
    s.td = (c->dry_active && c->flags_version == c->state_version) ? c->tile_dry : c->tile_zero;
    s.td0 = (c->dry_active && c->flags0_valid) ? c->tile_dry0 : c->tile_zero;
    s.tile_list = use_tile_list(c) ? c->tile_list : nullptr;
    s.tdp = c->k1_dry ? c->tile_prev : c->tile_zero;
    s.tdf = c->tile_fluxed;
    s.tdd = c->tile_drained;
    return s;
}

template <class T>
static cudaError_t dalloc(T **p, size_t n) { return cudaMalloc((void **)p, std::max<size_t>(n, 1) * sizeof(T)); }

static inline int drain_grid(const swe_ctx *c) {
#if SWE_K3_PERSISTENT
    return std::min(nblk(c->nt, kBlock), c->sms * SWE_K3_GRID_PER_SM);
#else
    return nblk(c->nt, kBlock * SWE_K3_ILP);
#endif
}
// the tile flags stop describing what the edge-state arrays hold (state / bed set from outside, mode change): zero them
// on the stream, so the next reconstruction pass finds no `previous` flags (works under graph replay too)
static inline void drop_tile_flags(swe_ctx *c) {
    if (c->tile_dry) {
        cudaSetDevice(c->device);
        cudaMemsetAsync(c->tile_dry, 0, (size_t)c->ntiles, c->stream);
        cudaMemsetAsync(c->tile_fluxed, 0, (size_t)c->ntiles, c->stream);
        cudaMemsetAsync(c->tile_drained, 0, (size_t)c->ntiles, c->stream);
    }
    c->flags_version = 0;
}
// dry-region skipping: are the flags of this stage usable right now?
static inline bool dry_now(const swe_ctx *c) { return c->dry_active && c->flags_version == c->state_version; }
static int launch_check(swe_ctx *c, const char *what);
// Decide whether the dry-region instantiations pay off (auto mode). Synchronises the stream once; called from the
// step / run entry points only when the state was set from outside or a long run asks for a re-evaluation,
// never during graph capture.
static int dry_refresh(swe_ctx *c) {
    if (c->opt_dry_skip >= 0) { c->dry_active = c->opt_dry_skip == 1; c->dry_eval_pending = false; return SWE_OK; }
    if (!c->dry_eval_pending) return SWE_OK;
    c->dry_eval_pending = false;
    c->steps_since_eval = 0;
    CUDA_TRY(c, cudaSetDevice(c->device));
    unsigned long long *cnt = reinterpret_cast<unsigned long long *>(c->diag);  // scratch: 8-byte aligned device doubles
    CUDA_TRY(c, cudaMemsetAsync(cnt, 0, sizeof(unsigned long long), c->stream));
    k_count_dry<<<std::min(nblk(c->nt, 256), 8 * c->sms), 256, 0, c->stream>>>(c->nt, c->cur[0], c->cb, cnt);
    int rc = launch_check(c, "k_count_dry");
    if (rc) return rc;
    unsigned long long h = 0;
    CUDA_TRY(c, cudaMemcpyAsync(&h, cnt, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    const bool on = (double)h >= 0.2 * (double)c->nt;
    if (on != c->dry_active) { c->dry_active = on; c->flags_version = 0; c->flags0_valid = false; drop_tile_flags(c); }
    return SWE_OK;
}
static int launch_check(swe_ctx *c, const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { c->err = std::string(what) + ": " + cudaGetErrorString(e); return SWE_ERR_CUDA; }
    c->launches++;
    return SWE_OK;
}

static void destroy_ctx(swe_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    void *ptrs[] = {c->tt, c->te, c->tp, c->slotL, c->slotR, c->cgeo, c->node, c->en, c->area, c->cb, c->elen, c->dmin,
                    c->n2c_start, c->n2c_cells, c->cfl_mask, c->dmin0, c->cell_old, c->edge_old, c->node_old, c->bufA[0],
                    c->bufA[1], c->bufA[2], c->bufB[0], c->bufB[1], c->bufB[2], c->ceh, c->ceu, c->cev, c->cgx, c->cgy,
                    c->cew, c->f0, c->f1, c->f2, c->dti, c->pwl, c->cls, c->pw_list, c->rs_list, c->scal, c->flags, c->diag, c->stage_aos,
                    c->send_cells, c->recv_cells, c->dbg, c->drain_list, c->tile_dry, c->tile_dry0, c->tile_zero, c->tile_prev, c->tile_list, c->tile_fluxed, c->tile_drained};
    for (void *p : ptrs) if (p) cudaFree(p);
    for (void *p : c->p2p_imported) cudaIpcCloseMemHandle(p);
    if (c->p2p_recv[0]) cudaFree(c->p2p_recv[0]);
    if (c->p2p_recv[1]) cudaFree(c->p2p_recv[1]);
    if (c->p2p_flags) cudaFree(c->p2p_flags);
    for (auto &g : c->graphs) cudaGraphExecDestroy(g.exec);
    if (c->gstream) cudaStreamDestroy(c->gstream);
    if (c->gev) cudaEventDestroy(c->gev);
    if (c->pipe.ok) {
        cudaStreamSynchronize(c->pipe.s_in); cudaStreamSynchronize(c->pipe.s_out);
        for (int q = 0; q < 2; ++q) {
            cudaFree(c->pipe.in[q]); cudaFree(c->pipe.out[q]);
            cudaEventDestroy(c->pipe.in_ready[q]); cudaEventDestroy(c->pipe.in_consumed[q]);
            cudaEventDestroy(c->pipe.out_ready[q]); cudaEventDestroy(c->pipe.out_free[q]);
        }
        cudaStreamDestroy(c->pipe.s_in); cudaStreamDestroy(c->pipe.s_out);
    }
    for (auto &p : c->kt_pairs) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
    for (auto &p : c->kt_pool) { cudaEventDestroy(p.first); cudaEventDestroy(p.second); }
    delete c;
}

static int ensure_stage(swe_ctx *c, size_t n_doubles) {
    if (c->stage_cap >= n_doubles) return SWE_OK;
    if (c->stage_aos) cudaFree(c->stage_aos);
    c->stage_aos = nullptr; c->stage_cap = 0;
    CUDA_TRY(c, dalloc(&c->stage_aos, n_doubles));
    c->stage_cap = n_doubles;
    return SWE_OK;
}

// flux registry (swe_flux_registry.cuh): id -> kernel instantiation
struct FluxerEntry { int id; const char *name; };
static const FluxerEntry g_fluxers[] = {
#define SWE_X(ID, NAME, TYPE) {ID, NAME},
    SWE_FLUX_LIST(SWE_X)
#undef SWE_X
};
constexpr int kNumFluxers = (int)(sizeof(g_fluxers) / sizeof(g_fluxers[0]));
static bool fluxer_known(int id) {
    for (int k = 0; k < kNumFluxers; ++k) if (g_fluxers[k].id == id) return true;
    return false;
}
static void launch_flux(swe_ctx *c, const DevMesh &m, const DevFields &s, int fluxer_id, bool cfl = true) {
    const int g = std::min(nblk(c->ne, kBlock), c->sms * SWE_K2_GRID_PER_SM);
    const double ac = std::fabs(c->cor);
    const int rf = c->opt_roe_fix, ca = c->opt_cfl_abs;
    const bool opt = rf || ca;  // the default instantiation is upstream as written
    const bool dry = dry_now(c);
    switch (fluxer_id) {
#define SWE_X(ID, NAME, TYPE)                                                            \
        case ID:                                                                         \
            if (dry && !opt) {                                                           \
                if (cfl) k_flux<TYPE, false, true, true><<<g, kBlock, 0, c->stream>>>(m, s, ac, rf, ca);  \
                else k_flux<TYPE, false, false, true><<<g, kBlock, 0, c->stream>>>(m, s, ac, rf, ca);     \
            } else if (!cfl) {                                                           \
                if (opt) k_flux<TYPE, true, false><<<g, kBlock, 0, c->stream>>>(m, s, ac, rf, ca);  \
                else k_flux<TYPE, false, false><<<g, kBlock, 0, c->stream>>>(m, s, ac, rf, ca);     \
            } else if (opt) k_flux<TYPE, true><<<g, kBlock, 0, c->stream>>>(m, s, ac, rf, ca);       \
            else k_flux<TYPE, false><<<g, kBlock, 0, c->stream>>>(m, s, ac, rf, ca);      \
            break;
        SWE_FLUX_LIST(SWE_X)
#undef SWE_X
        default: break;
    }
}

// CUDA loads kernels lazily (CUDA_MODULE_LOADING=LAZY): the first launch of a kernel may have to wait for every
// running kernel to finish. A rank that spins in k_halo_wait_unpack / k_min_pull for a peer of the SAME process
// would then dead-lock the host thread that is about to issue that peer's first k_min_push. So every kernel
// of the time step is loaded once, up front, when the first context is created.
static void preload_kernels() {
    static bool done = false;
    if (done) return;
    done = true;
    cudaFuncAttributes a;
#define SWE_LOAD(...) cudaFuncGetAttributes(&a, (const void *)(__VA_ARGS__))
#define SWE_LOAD_K1(T) SWE_LOAD(k_reconstruct<T, 0>); SWE_LOAD(k_reconstruct<T, 1>); SWE_LOAD(k_reconstruct<T, 2>); \
    SWE_LOAD(k_reconstruct_tiled<T, 0>); SWE_LOAD(k_reconstruct_tiled<T, 1>); SWE_LOAD(k_reconstruct_tiled<T, 2>); \
    SWE_LOAD(k_reconstruct_pf<T, 0>); SWE_LOAD(k_reconstruct_pf<T, 1>); SWE_LOAD(k_reconstruct_pf<T, 2>); \
    SWE_LOAD(k_reconstruct_slow<T>); SWE_LOAD(k_partwet2<T>)
    SWE_LOAD_K1(false); SWE_LOAD_K1(true);
    SWE_LOAD(k_reconstruct<false, 0, true>); SWE_LOAD(k_reconstruct<false, 1, true>); SWE_LOAD(k_reconstruct<false, 2, true>);
    SWE_LOAD(k_drain<true>); SWE_LOAD(k_count_dry); SWE_LOAD(k_tile_compact); SWE_LOAD(k_fill_dry_tiles);
    SWE_LOAD(k_update<true, true, false, false, true>); SWE_LOAD(k_update<true, false, false, false, true>);
    SWE_LOAD(k_update<false, true, false, false, true>); SWE_LOAD(k_update<false, false, false, false, true>);
#define SWE_X(ID, NAME, TYPE) SWE_LOAD(k_flux<TYPE, false>); SWE_LOAD(k_flux<TYPE, true>); \
    SWE_LOAD(k_flux<TYPE, false, false>); SWE_LOAD(k_flux<TYPE, true, false>); \
    SWE_LOAD(k_flux<TYPE, false, true, true>); SWE_LOAD(k_flux<TYPE, false, false, true>);
    SWE_FLUX_LIST(SWE_X)
#undef SWE_X
    SWE_LOAD(k_drain<false>); SWE_LOAD(k_drain_list);
    SWE_LOAD(k_update<true, true, false, true>); SWE_LOAD(k_update<true, false, false, true>);
    SWE_LOAD(k_update<false, true, false, true>); SWE_LOAD(k_update<false, false, false, true>);
    SWE_LOAD(k_update<true, true>); SWE_LOAD(k_update<true, false>); SWE_LOAD(k_update<false, true>); SWE_LOAD(k_update<false, false>);
    SWE_LOAD(k_update<true, true, true>); SWE_LOAD(k_update<true, false, true>); SWE_LOAD(k_classify);
    SWE_LOAD(k_post_step); SWE_LOAD(k_set_scalar);
    SWE_LOAD(k_halo_pack); SWE_LOAD(k_halo_unpack); SWE_LOAD(k_halo_signal); SWE_LOAD(k_halo_wait);
    SWE_LOAD(k_halo_pack_signal); SWE_LOAD(k_halo_wait_unpack); SWE_LOAD(k_min_push); SWE_LOAD(k_min_pull);
    SWE_LOAD(k_state_hash); SWE_LOAD(k_state_in); SWE_LOAD(k_state_out); SWE_LOAD(k_diag_partial); SWE_LOAD(k_diag_final);
#undef SWE_LOAD_K1
#undef SWE_LOAD
    if (const char *e = std::getenv("SWE_B200_CARVEOUT")) {  // A/B: shared-memory carve-out (per cent) of the gather kernels
        const int pct = std::atoi(e);
        cudaFuncSetAttribute((const void *)k_reconstruct<false, 0>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
        cudaFuncSetAttribute((const void *)k_flux<BuiltinFlux<1, 2>, false>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
        cudaFuncSetAttribute((const void *)k_update<true, false>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
        cudaFuncSetAttribute((const void *)k_update<false, false>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    }
    cudaGetLastError();
}

extern "C" {

SWE_API const char *swe_version(void) { return "swe_b200 0.2 (sm_100a, fp64, -fmad=false)"; }
SWE_API int32_t swe_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

SWE_API const char *swe_last_error(const swe_ctx *ctx) {
    if (ctx) return ctx->err.c_str();
    if (!g_create_error.empty()) return g_create_error.c_str();
    return swe::host_error();
}

SWE_API int swe_create_classes(swe_ctx **out, const swe_mesh *mesh, int device, int reorder, const uint8_t *cell_class) {
    g_create_error.clear();
    auto fail = [&](int code, const std::string &msg) { g_create_error = msg; return code; };
    if (!out || !mesh) return fail(SWE_ERR_INVALID, "swe_create: null argument");
    *out = nullptr;
    const int64_t nn = mesh->nn, ne = mesh->ne, nt = mesh->nt;
    if (nn <= 0 || ne <= 0 || nt <= 0 || !mesh->geometry || !mesh->edge_nodes || !mesh->edge_elements ||
        !mesh->element_nodes || !mesh->element_edges || !mesh->element_neighbours)
        return fail(SWE_ERR_INVALID, "swe_create: empty mesh or null array");
    if (3 * nt >= (int64_t)2147483647 || 2 * ne >= (int64_t)2147483647)
        return fail(SWE_ERR_INVALID, "swe_create: mesh too large for int32 device ids");
    // validate ids and boundary tags (only SOLID_WALL is implemented upstream, src/SpaceDisc.cpp:66-72)
    {
        int bad = 0;  // 1 edge_elements, 2 boundary tag, 3 edge_nodes, 4 element_nodes, 5 element_edges, 6 neighbours
#pragma omp parallel for schedule(static) num_threads(host_threads()) reduction(max : bad)
        for (int64_t e = 0; e < ne; ++e) {
            const int64_t a = mesh->edge_elements[2 * e], b = mesh->edge_elements[2 * e + 1];
            if (a < 0 || a >= nt || b >= nt) bad = std::max(bad, 1);
            else if (b < 0 && b != SWE_SOLID_WALL) bad = std::max(bad, 2);
            if (mesh->edge_nodes[2 * e] < 0 || mesh->edge_nodes[2 * e] >= nn || mesh->edge_nodes[2 * e + 1] < 0 ||
                mesh->edge_nodes[2 * e + 1] >= nn)
                bad = std::max(bad, 3);
        }
        if (bad == 1) return fail(SWE_ERR_INVALID, "swe_create: edge_elements id out of range");
        if (bad == 2) return fail(SWE_ERR_INVALID, "swe_create: only SOLID_WALL (-1) boundaries are supported");
        if (bad == 3) return fail(SWE_ERR_INVALID, "swe_create: edge_nodes id out of range");
#pragma omp parallel for schedule(static) num_threads(host_threads()) reduction(max : bad)
        for (int64_t k = 0; k < 3 * nt; ++k) {
            if (mesh->element_nodes[k] < 0 || mesh->element_nodes[k] >= nn) bad = std::max(bad, 4);
            if (mesh->element_edges[k] < 0 || mesh->element_edges[k] >= ne) bad = std::max(bad, 5);
            const int64_t nb = mesh->element_neighbours[k];
            if (nb >= nt || (nb < 0 && nb != SWE_SOLID_WALL)) bad = std::max(bad, 6);
        }
        if (bad == 4) return fail(SWE_ERR_INVALID, "swe_create: element_nodes id out of range");
        if (bad == 5) return fail(SWE_ERR_INVALID, "swe_create: element_edges id out of range");
        if (bad == 6) return fail(SWE_ERR_INVALID, "swe_create: element_neighbours id out of range / unsupported boundary");
    }
    // SWE_B200_TIMING=1: phase times of the set-up on stderr
    const bool timing = std::getenv("SWE_B200_TIMING") != nullptr;
    auto t_last = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) {
        if (!timing) return;
        cudaDeviceSynchronize();
        const auto now = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[swe_create] %-28s %8.3f s\n", what, std::chrono::duration<double>(now - t_last).count());
        t_last = now;
    };
    lap("validate");
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0)
        return fail(SWE_ERR_CUDA, std::string("swe_create: no CUDA device available (") +
                                      (ce != cudaSuccess ? cudaGetErrorString(ce) : "device count 0") +
                                      "); this library has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(SWE_ERR_INVALID, "swe_create: bad device ordinal");
    if ((ce = cudaSetDevice(device)) != cudaSuccess) return fail(SWE_ERR_CUDA, cudaGetErrorString(ce));
    int sm_count = 148;
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, device);
    preload_kernels();

    swe_ctx *c = new (std::nothrow) swe_ctx();
    if (!c) return fail(SWE_ERR_NOMEM, "out of host memory");
    c->device = device; c->nt = (int)nt; c->ne = (int)ne; c->nn = (int)nn; c->sms = sm_count;
    c->cor = mesh->cor; c->tau = mesh->tau;
    c->reordered = reorder != 0 || cell_class != nullptr;
    // tuning defaults can be overridden from the environment for A/B runs (same switches as swe_set_option)
    if (const char *e = std::getenv("SWE_B200_FUSED_DRAIN")) c->opt_fused_drain = std::atoi(e) != 0;
    if (const char *e = std::getenv("SWE_B200_K1_TILED")) c->opt_tiled = SWE_K1_TILED ? std::min(2, std::max(0, std::atoi(e))) : 0;
    if (cell_class)
        for (int64_t t = 0; t < nt; ++t)
            if (cell_class[t] > 3) { delete c; return fail(SWE_ERR_INVALID, "swe_create_classes: class ids must be 0..3"); }

    // ---- upload the caller's arrays as they are; numbering, conversion and checks run on the device ----
#define CREATE_TRY(call)                                                                          \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess) {                                                                  \
            std::string msg_ = std::string(#call) + ": " + cudaGetErrorString(e_);                \
            free_setup();                                                                         \
            destroy_ctx(c);                                                                       \
            return fail(e_ == cudaErrorMemoryAllocation ? SWE_ERR_NOMEM : SWE_ERR_CUDA, msg_);    \
        }                                                                                         \
    } while (0)
    long long *d_tp64 = nullptr, *d_te64 = nullptr, *d_tt64 = nullptr, *d_ep64 = nullptr, *d_et64 = nullptr;
    double *d_geom = nullptr;
    unsigned char *d_cls = nullptr;
    int *d_cell_new = nullptr, *d_edge_new = nullptr, *d_node_new = nullptr;  // caller id -> device id (nullptr: identity)
    int *d_ep0 = nullptr, *d_bad = nullptr;
    unsigned long long *d_keys = nullptr, *d_keys2 = nullptr;
    int *d_vals = nullptr;
    void *d_tmp = nullptr;
    auto free_setup = [&]() {
        void *ptrs[] = {d_tp64, d_te64, d_tt64, d_ep64, d_et64, d_geom, d_cls, d_cell_new, d_edge_new, d_node_new, d_ep0, d_bad,
                        d_keys, d_keys2, d_vals, d_tmp};
        for (void *q : ptrs) if (q) cudaFree(q);
    };
    static_assert(sizeof(long long) == sizeof(int64_t), "int64 incidence arrays");
    CREATE_TRY(dalloc(&d_tp64, (size_t)3 * nt)); CREATE_TRY(dalloc(&d_te64, (size_t)3 * nt)); CREATE_TRY(dalloc(&d_tt64, (size_t)3 * nt));
    CREATE_TRY(dalloc(&d_ep64, (size_t)2 * ne)); CREATE_TRY(dalloc(&d_et64, (size_t)2 * ne)); CREATE_TRY(dalloc(&d_geom, (size_t)3 * nn));
    CREATE_TRY(cudaMemcpy(d_tp64, mesh->element_nodes, sizeof(int64_t) * 3 * nt, cudaMemcpyHostToDevice));
    CREATE_TRY(cudaMemcpy(d_te64, mesh->element_edges, sizeof(int64_t) * 3 * nt, cudaMemcpyHostToDevice));
    CREATE_TRY(cudaMemcpy(d_tt64, mesh->element_neighbours, sizeof(int64_t) * 3 * nt, cudaMemcpyHostToDevice));
    CREATE_TRY(cudaMemcpy(d_ep64, mesh->edge_nodes, sizeof(int64_t) * 2 * ne, cudaMemcpyHostToDevice));
    CREATE_TRY(cudaMemcpy(d_et64, mesh->edge_elements, sizeof(int64_t) * 2 * ne, cudaMemcpyHostToDevice));
    CREATE_TRY(cudaMemcpy(d_geom, mesh->geometry, sizeof(double) * 3 * nn, cudaMemcpyHostToDevice));
    lap("upload caller arrays");

    // ---- numbering: caller id -> device id (space-filling-curve order, stable CUB radix sort of the keys) ----
    if (c->reordered) {
        double x0 = 1e300, x1 = -1e300, y0 = 1e300, y1 = -1e300;
#pragma omp parallel for schedule(static) num_threads(host_threads()) reduction(min : x0, y0) reduction(max : x1, y1)
        for (int64_t p = 0; p < nn; ++p) {
            x0 = std::min(x0, mesh->geometry[3 * p]); x1 = std::max(x1, mesh->geometry[3 * p]);
            y0 = std::min(y0, mesh->geometry[3 * p + 1]); y1 = std::max(y1, mesh->geometry[3 * p + 1]);
        }
        const double span = std::max(std::max(x1 - x0, y1 - y0), 1e-300);
        const KeyBox box{x0, y0, 2097152.0 / span};
        const size_t nmax = (size_t)std::max(std::max(nt, ne), nn);
        CREATE_TRY(dalloc(&d_keys, nmax)); CREATE_TRY(dalloc(&d_keys2, nmax)); CREATE_TRY(dalloc(&d_vals, nmax));
        size_t tmp_bytes = 0;
        CREATE_TRY(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_keys, d_keys2, d_vals, d_vals, (int64_t)nmax, 0, 45));
        CREATE_TRY(cudaMalloc(&d_tmp, std::max<size_t>(tmp_bytes, 1)));
        if (cell_class) {
            CREATE_TRY(dalloc(&d_cls, (size_t)nt));
            CREATE_TRY(cudaMemcpy(d_cls, cell_class, (size_t)nt, cudaMemcpyHostToDevice));
        }
        // order[k] = caller id of device id k (= the *_old maps kept by the context); newid = its inverse
        auto number = [&](int which, int64_t n, const long long *ids, const unsigned char *cls, int **old_out, int **new_out) -> cudaError_t {
            cudaError_t e;
            k_order_keys<<<nblk(n, 256), 256>>>(which, (long long)n, ids, d_geom, box, cls, reorder != 0 ? 1 : 0, d_keys, d_vals);
            if ((e = cudaGetLastError()) != cudaSuccess) return e;
            if ((e = dalloc(old_out, (size_t)n)) != cudaSuccess) return e;
            if ((e = cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, d_keys, d_keys2, d_vals, *old_out, n, 0, 45)) != cudaSuccess) return e;
            if ((e = dalloc(new_out, (size_t)n)) != cudaSuccess) return e;
            k_invert_perm<<<nblk(n, 256), 256>>>((int)n, *old_out, *new_out);
            return cudaGetLastError();
        };
        CREATE_TRY(number(0, nt, d_tp64, d_cls, &c->cell_old, &d_cell_new));
        if (reorder != 0) {
            CREATE_TRY(number(1, ne, d_ep64, nullptr, &c->edge_old, &d_edge_new));
            CREATE_TRY(number(2, nn, nullptr, nullptr, &c->node_old, &d_node_new));
        } else {  // classes only: edges and nodes keep the caller's numbering (identity maps, as before)
            std::vector<int> id((size_t)std::max(ne, nn));
            std::iota(id.begin(), id.end(), 0);
            CREATE_TRY(dalloc(&c->edge_old, (size_t)ne)); CREATE_TRY(dalloc(&c->node_old, (size_t)nn));
            CREATE_TRY(cudaMemcpy(c->edge_old, id.data(), sizeof(int) * ne, cudaMemcpyHostToDevice));
            CREATE_TRY(cudaMemcpy(c->node_old, id.data(), sizeof(int) * nn, cudaMemcpyHostToDevice));
        }
        try {
            c->cell_new.resize((size_t)nt);
        } catch (const std::bad_alloc &) { free_setup(); destroy_ctx(c); return fail(SWE_ERR_NOMEM, "out of host memory"); }
        CREATE_TRY(cudaMemcpy(c->cell_new.data(), d_cell_new, sizeof(int) * nt, cudaMemcpyDeviceToHost));
    } else {
        try {
            c->cell_new.resize((size_t)nt);
        } catch (const std::bad_alloc &) { free_setup(); destroy_ctx(c); return fail(SWE_ERR_NOMEM, "out of host memory"); }
        std::iota(c->cell_new.begin(), c->cell_new.end(), 0);
    }
    lap("numbering (keys + sorts)");
    {   // device range of every ordering class (one class = everything when none are given)
        int counts[5] = {0, 0, 0, 0, 0};
        if (cell_class) for (int64_t t = 0; t < nt; ++t) counts[cell_class[t]]++; else counts[0] = (int)nt;
        c->class_first[0] = 0;
        for (int q = 0; q < 5; ++q) c->class_first[q + 1] = c->class_first[q] + counts[q];
    }

    // ---- conversion to int32 SoA in device numbering + the checks of the local convention the kernels rely on
    // (SURVEY App. B rules 3-4, notebooks/topology.dat): TriangEdges[k] joins TriangPoints[k] and TriangPoints[(k+1)%3],
    // TriangTriangs[k] lies across it ----
    {
        CREATE_TRY(dalloc(&c->tp, (size_t)3 * nt)); CREATE_TRY(dalloc(&c->tt, (size_t)3 * nt)); CREATE_TRY(dalloc(&c->te, (size_t)3 * nt));
        CREATE_TRY(dalloc(&d_bad, 1));
        CREATE_TRY(cudaMemset(d_bad, 0, sizeof(int)));
        k_convert_cells<<<nblk(nt, 256), 256>>>((int)nt, d_tp64, d_te64, d_tt64, d_ep64, d_et64, d_cell_new, d_edge_new, d_node_new, c->tp,
                                               c->tt, c->te, d_bad);
        CREATE_TRY(cudaGetLastError());
        int inconsistent = 0;
        CREATE_TRY(cudaMemcpy(&inconsistent, d_bad, sizeof(int), cudaMemcpyDeviceToHost));
        if (inconsistent) {
            free_setup();
            destroy_ctx(c);
            return fail(SWE_ERR_INVALID,
                        inconsistent == 1 ? "swe_create: element_edges / edge_elements are inconsistent"
                        : inconsistent == 2 ? "swe_create: element_edges[k] must join element_nodes[k] and element_nodes[(k+1)%3]"
                                            : "swe_create: element_neighbours[k] must be the cell across element_edges[k]");
        }
        lap("cell incidence -> int32 SoA");
        // node -> incident cells (CSR, device numbering; pass 2 gathers the node maxima over it), built on the device:
        // histogram of the node ids -> exclusive scan = row starts; stable sort of (node, cell) pairs = row contents
        {
            int *keys_in = nullptr, *keys_out = nullptr, *vals_in = nullptr;
            void *tmp = nullptr;
            size_t tb1 = 0, tb2 = 0;
            CREATE_TRY(dalloc(&c->n2c_start, (size_t)nn + 1)); CREATE_TRY(dalloc(&c->n2c_cells, (size_t)3 * nt));
            CREATE_TRY(dalloc(&keys_out, (size_t)3 * nt)); CREATE_TRY(dalloc(&vals_in, (size_t)3 * nt));
            keys_in = c->tp;  // tp[k * nt + d] = node of local corner k of cell d
            CREATE_TRY(cudaMemset(c->n2c_start, 0, sizeof(int) * (nn + 1)));
            k_n2c_count<<<nblk(3 * nt, 256), 256>>>((int)(3 * nt), (int)nt, keys_in, c->n2c_start + 1, vals_in);
            CREATE_TRY(cudaGetLastError());
            CREATE_TRY(cub::DeviceScan::InclusiveSum(nullptr, tb1, c->n2c_start + 1, c->n2c_start + 1, (int)nn));
            CREATE_TRY(cub::DeviceRadixSort::SortPairs(nullptr, tb2, keys_in, keys_out, vals_in, c->n2c_cells, (int64_t)(3 * nt)));
            CREATE_TRY(cudaMalloc(&tmp, std::max(tb1, tb2)));
            CREATE_TRY(cub::DeviceScan::InclusiveSum(tmp, tb1, c->n2c_start + 1, c->n2c_start + 1, (int)nn));
            CREATE_TRY(cub::DeviceRadixSort::SortPairs(tmp, tb2, keys_in, keys_out, vals_in, c->n2c_cells, (int64_t)(3 * nt)));
            CREATE_TRY(cudaDeviceSynchronize());
            cudaFree(tmp); cudaFree(keys_out); cudaFree(vals_in);
        }
    }
    lap("node->cell CSR");
    int *d_ep1 = nullptr, *d_et0 = nullptr, *d_et1 = nullptr;
    CREATE_TRY(dalloc(&d_ep0, (size_t)4 * ne));
    d_ep1 = d_ep0 + ne; d_et0 = d_ep1 + ne; d_et1 = d_et0 + ne;
    k_convert_edges<<<nblk(ne, 256), 256>>>((int)ne, d_ep64, d_et64, d_cell_new, d_edge_new, d_node_new, d_ep0, d_ep1, d_et0, d_et1);
    CREATE_TRY(cudaGetLastError());
    CREATE_TRY(dalloc(&c->node, (size_t)nn));
    k_convert_nodes<<<nblk(nn, 256), 256>>>((int)nn, d_geom, d_node_new, c->node);
    CREATE_TRY(cudaGetLastError());
    CREATE_TRY(cudaDeviceSynchronize());
    // the int64 copies and the forward maps are no longer needed
    cudaFree(d_tp64); cudaFree(d_te64); cudaFree(d_tt64); cudaFree(d_ep64); cudaFree(d_et64); cudaFree(d_geom); cudaFree(d_cls);
    cudaFree(d_cell_new); cudaFree(d_edge_new); cudaFree(d_node_new); cudaFree(d_bad); cudaFree(d_keys); cudaFree(d_keys2);
    cudaFree(d_vals); cudaFree(d_tmp);
    d_tp64 = d_te64 = d_tt64 = d_ep64 = d_et64 = nullptr; d_geom = nullptr; d_cls = nullptr;
    d_cell_new = d_edge_new = d_node_new = d_bad = nullptr; d_keys = d_keys2 = nullptr; d_vals = nullptr; d_tmp = nullptr;
    lap("edges, nodes, inverse maps");
    // ---- geometry on device ----
    CREATE_TRY(dalloc(&c->cgeo, (size_t)nt)); CREATE_TRY(dalloc(&c->area, (size_t)nt)); CREATE_TRY(dalloc(&c->cb, (size_t)nt));
    CREATE_TRY(dalloc(&c->slotL, (size_t)ne)); CREATE_TRY(dalloc(&c->slotR, (size_t)ne));
    CREATE_TRY(dalloc(&c->en, (size_t)ne)); CREATE_TRY(dalloc(&c->elen, (size_t)ne)); CREATE_TRY(dalloc(&c->dmin, (size_t)ne));
    CREATE_TRY(cudaMemset(c->slotR, 0xff, sizeof(int) * ne));
    CREATE_TRY(cudaMemset(c->slotL, 0xff, sizeof(int) * ne));
    k_setup_cells<<<nblk(nt, 256), 256>>>(c->nt, c->tp, c->tt, c->node, c->cgeo, c->area, c->cb);
    k_setup_slots<<<nblk(nt, 256), 256>>>(c->nt, c->te, c->slotL, c->slotR);
    k_setup_edges<<<nblk(ne, 256), 256>>>(c->ne, c->nt, d_ep0, d_ep1, d_et0, d_et1, c->node, c->cgeo, c->area, c->en,
                                         c->elen, c->dmin);
    CREATE_TRY(cudaGetLastError());
    CREATE_TRY(cudaDeviceSynchronize());
    cudaFree(d_ep0);
    d_ep0 = nullptr;
    c->launches += 3;
    lap("geometry kernels");
    // ---- fields ----
    for (int q = 0; q < 3; ++q) { CREATE_TRY(dalloc(&c->bufA[q], (size_t)nt)); CREATE_TRY(dalloc(&c->bufB[q], (size_t)nt)); }
    c->cur = c->bufA; c->sav = c->bufA;
    CREATE_TRY(dalloc(&c->ceh, (size_t)3 * nt)); CREATE_TRY(dalloc(&c->ceu, (size_t)3 * nt)); CREATE_TRY(dalloc(&c->cev, (size_t)3 * nt));
    CREATE_TRY(dalloc(&c->cgx, (size_t)nt)); CREATE_TRY(dalloc(&c->cgy, (size_t)nt)); CREATE_TRY(dalloc(&c->pw_list, (size_t)nt)); CREATE_TRY(dalloc(&c->rs_list, (size_t)nt));
    CREATE_TRY(dalloc(&c->f0, (size_t)ne)); CREATE_TRY(dalloc(&c->f1, (size_t)ne)); CREATE_TRY(dalloc(&c->f2, (size_t)ne));
    CREATE_TRY(dalloc(&c->dti, (size_t)nt)); CREATE_TRY(dalloc(&c->cls, (size_t)nt)); CREATE_TRY(dalloc(&c->pwl, (size_t)nt));
    CREATE_TRY(dalloc(&c->scal, 8)); CREATE_TRY(dalloc(&c->flags, 8));
    CREATE_TRY(dalloc(&c->diag, (size_t)6 * kDiagBlocks + 8));
    for (int q = 0; q < 3; ++q) {
        CREATE_TRY(cudaMemset(c->bufA[q], 0, sizeof(double) * nt));
        CREATE_TRY(cudaMemset(c->bufB[q], 0, sizeof(double) * nt));
    }
    CREATE_TRY(cudaMemset(c->ceh, 0, sizeof(double) * 3 * nt)); CREATE_TRY(cudaMemset(c->ceu, 0, sizeof(double) * 3 * nt));
    CREATE_TRY(cudaMemset(c->cev, 0, sizeof(double) * 3 * nt)); CREATE_TRY(cudaMemset(c->cgx, 0, sizeof(double) * nt));
    CREATE_TRY(cudaMemset(c->cgy, 0, sizeof(double) * nt));
    CREATE_TRY(cudaMemset(c->f0, 0, sizeof(double) * ne)); CREATE_TRY(cudaMemset(c->f1, 0, sizeof(double) * ne));
    CREATE_TRY(cudaMemset(c->f2, 0, sizeof(double) * ne));
    CREATE_TRY(cudaMemset(c->dti, 0, sizeof(double) * nt)); CREATE_TRY(cudaMemset(c->cls, 0, nt));
    CREATE_TRY(cudaMemset(c->pwl, 0, sizeof(double) * nt));
    c->ntiles = (int)((nt + kBlock - 1) >> kUpdTileShift);
    CREATE_TRY(dalloc(&c->tile_dry, (size_t)c->ntiles)); CREATE_TRY(dalloc(&c->tile_dry0, (size_t)c->ntiles)); CREATE_TRY(dalloc(&c->tile_zero, (size_t)c->ntiles));
    CREATE_TRY(cudaMemset(c->tile_dry, 0, (size_t)c->ntiles)); CREATE_TRY(cudaMemset(c->tile_dry0, 0, (size_t)c->ntiles));
    CREATE_TRY(cudaMemset(c->tile_zero, 0, (size_t)c->ntiles));
    CREATE_TRY(dalloc(&c->tile_fluxed, (size_t)c->ntiles)); CREATE_TRY(dalloc(&c->tile_drained, (size_t)c->ntiles));
    CREATE_TRY(cudaMemset(c->tile_fluxed, 0, (size_t)c->ntiles)); CREATE_TRY(cudaMemset(c->tile_drained, 0, (size_t)c->ntiles));
    CREATE_TRY(dalloc(&c->tile_list, (size_t)c->ntiles + 1));
    CREATE_TRY(cudaMemset(c->tile_list, 0, sizeof(int) * ((size_t)c->ntiles + 1)));
    CREATE_TRY(dalloc(&c->tile_prev, (size_t)c->ntiles));
    CREATE_TRY(cudaMemset(c->tile_prev, 0, (size_t)c->ntiles));
    CREATE_TRY(cudaMemset(c->flags, 0, sizeof(int) * 8));
    const double scal0[8] = {1.0, 0.0, 0.0, 1.0, 1.0, 0, 0, 0};  // [0] min_len, [1] dt, [2] time, [3] running min, [4] global min_len
    CREATE_TRY(cudaMemcpy(c->scal, scal0, sizeof(scal0), cudaMemcpyHostToDevice));
    CREATE_TRY(cudaDeviceSynchronize());
    {   // K3' work list (fused draining dt): cells with a neighbour in another update tile or ordering class
        unsigned char *flag = nullptr;
        int *sel = nullptr, *d_cnt = nullptr;
        void *tmp = nullptr;
        size_t tb = 0;
        ClassFirst cf;
        for (int q = 0; q < 6; ++q) cf.f[q] = c->class_first[q];
        CREATE_TRY(dalloc(&flag, (size_t)nt)); CREATE_TRY(dalloc(&sel, (size_t)nt)); CREATE_TRY(dalloc(&d_cnt, 1));
        k_mark_drain_boundary<<<nblk(nt, 256), 256>>>(c->nt, c->tt, cf, flag);
        CREATE_TRY(cudaGetLastError());
        thrust::counting_iterator<int> ids(0);
        CREATE_TRY(cub::DeviceSelect::Flagged(nullptr, tb, ids, flag, sel, d_cnt, (int)nt));
        CREATE_TRY(cudaMalloc(&tmp, std::max<size_t>(tb, 1)));
        CREATE_TRY(cub::DeviceSelect::Flagged(tmp, tb, ids, flag, sel, d_cnt, (int)nt));
        CREATE_TRY(cudaMemcpy(&c->drain_count, d_cnt, sizeof(int), cudaMemcpyDeviceToHost));
        CREATE_TRY(dalloc(&c->drain_list, (size_t)c->drain_count));
        CREATE_TRY(cudaMemcpy(c->drain_list, sel, sizeof(int) * (size_t)c->drain_count, cudaMemcpyDeviceToDevice));
        cudaFree(tmp); cudaFree(flag); cudaFree(sel); cudaFree(d_cnt);
        c->launches += 1;
    }
#undef CREATE_TRY
    lap("fields + drain list");
    *out = c;
    return SWE_OK;
}

SWE_API int swe_create(swe_ctx **out, const swe_mesh *mesh, int device, int reorder) {
    return swe_create_classes(out, mesh, device, reorder, nullptr);
}

SWE_API void swe_destroy(swe_ctx *ctx) { destroy_ctx(ctx); }

SWE_API int swe_set_stream(swe_ctx *c, void *stream) {
    if (!c) return SWE_ERR_INVALID;
    c->stream = (cudaStream_t)stream;
    return SWE_OK;
}

SWE_API int swe_synchronize(swe_ctx *c) {
    if (!c) return SWE_ERR_INVALID;
    CUDA_TRY(c, cudaSetDevice(c->device));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    int flag = 0;
    CUDA_TRY(c, cudaMemcpy(&flag, c->flags, sizeof(int), cudaMemcpyDeviceToHost));
    if (flag) { c->err = "non-finite cell state detected on device (SolverError)"; return SWE_ERR_NUMERIC; }
    return SWE_OK;
}

// ---- state transfer ----
SWE_API int swe_set_state_async(swe_ctx *c, const double *prim) {
    if (!c || !prim) return SWE_ERR_INVALID;
    CUDA_TRY(c, cudaSetDevice(c->device));
    int rc = ensure_stage(c, (size_t)3 * c->nt);
    if (rc) return rc;
    CUDA_TRY(c, cudaMemcpyAsync(c->stage_aos, prim, sizeof(double) * 3 * c->nt, cudaMemcpyHostToDevice, c->stream));
    c->state_version++;
    c->dry_eval_pending = true;
    drop_tile_flags(c);
    k_state_in<<<nblk(c->nt, 256), 256, 0, c->stream>>>(c->nt, c->cell_old, c->stage_aos, c->cur[0], c->cur[1], c->cur[2]);
    if ((rc = launch_check(c, "k_state_in"))) return rc;
    k_set_scalar<<<1, 1, 0, c->stream>>>(c->scal + 2, 0.0);
    CUDA_TRY(c, cudaMemsetAsync(c->flags, 0, sizeof(int), c->stream));
    return launch_check(c, "k_set_scalar");
}
SWE_API int swe_set_state(swe_ctx *c, const double *prim) {
    int rc = swe_set_state_async(c, prim);
    if (rc) return rc;
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return SWE_OK;
}
SWE_API int swe_get_state_async(swe_ctx *c, double *prim) {
    if (!c || !prim) return SWE_ERR_INVALID;
    CUDA_TRY(c, cudaSetDevice(c->device));
    int rc = ensure_stage(c, (size_t)3 * c->nt);
    if (rc) return rc;
    k_state_out<<<nblk(c->nt, 256), 256, 0, c->stream>>>(c->nt, c->cell_old, c->cur[0], c->cur[1], c->cur[2], c->stage_aos);
    if ((rc = launch_check(c, "k_state_out"))) return rc;
    CUDA_TRY(c, cudaMemcpyAsync(prim, c->stage_aos, sizeof(double) * 3 * c->nt, cudaMemcpyDeviceToHost, c->stream));
    return SWE_OK;
}
SWE_API int swe_get_state(swe_ctx *c, double *prim) {
    int rc = swe_get_state_async(c, prim);
    if (rc) return rc;
    return swe_synchronize(c);
}

static int one_step_fwd(swe_ctx *c, swe_scheme scheme, swe_flux flux, swe_wavespeed ws, double dt);
// ---- host-buffer pipeline: a stream of independent states, one time step each ----
// Every submitted batch is uploaded from its (pinned) host buffer, stepped once, and downloaded into its (pinned)
// output buffer. Three streams and double-buffered staging: the H2D copy of batch n+1 and the D2H copy of batch n-1
// run while batch n is being stepped, so the steady-state cost per batch is max(H2D, step, D2H) instead of their sum
// (PCIe is full duplex). Nothing blocks the host until swe_wait_host.
static int pipe_init(swe_ctx *c) {
    if (c->pipe.ok) return SWE_OK;
    auto &p = c->pipe;
    CUDA_TRY(c, cudaStreamCreateWithFlags(&p.s_in, cudaStreamNonBlocking));
    CUDA_TRY(c, cudaStreamCreateWithFlags(&p.s_out, cudaStreamNonBlocking));
    for (int q = 0; q < 2; ++q) {
        CUDA_TRY(c, dalloc(&p.in[q], (size_t)3 * c->nt));
        CUDA_TRY(c, dalloc(&p.out[q], (size_t)3 * c->nt));
        CUDA_TRY(c, cudaEventCreateWithFlags(&p.in_ready[q], cudaEventDisableTiming));
        CUDA_TRY(c, cudaEventCreateWithFlags(&p.in_consumed[q], cudaEventDisableTiming));
        CUDA_TRY(c, cudaEventCreateWithFlags(&p.out_ready[q], cudaEventDisableTiming));
        CUDA_TRY(c, cudaEventCreateWithFlags(&p.out_free[q], cudaEventDisableTiming));
    }
    p.n = 0; p.ok = true;
    return SWE_OK;
}
// stage 1 of a batch: upload + conversion into the device state (compute stream ordered after it)
static int pipe_begin(swe_ctx *c, const double *host_in) {
    CUDA_TRY(c, cudaSetDevice(c->device));
    int rc = pipe_init(c);
    if (rc) return rc;
    auto &p = c->pipe;
    const int b = (int)(p.n & 1);
    if (p.n >= 2) CUDA_TRY(c, cudaStreamWaitEvent(p.s_in, p.in_consumed[b], 0));
    CUDA_TRY(c, cudaMemcpyAsync(p.in[b], host_in, sizeof(double) * 3 * c->nt, cudaMemcpyHostToDevice, p.s_in));
    CUDA_TRY(c, cudaEventRecord(p.in_ready[b], p.s_in));
    CUDA_TRY(c, cudaStreamWaitEvent(c->stream, p.in_ready[b], 0));
    c->state_version++;
    k_state_in<<<nblk(c->nt, 256), 256, 0, c->stream>>>(c->nt, c->cell_old, p.in[b], c->cur[0], c->cur[1], c->cur[2]);
    if ((rc = launch_check(c, "k_state_in"))) return rc;
    CUDA_TRY(c, cudaEventRecord(p.in_consumed[b], c->stream));
    return SWE_OK;
}
// stage 3 of a batch: conversion + download
static int pipe_end(swe_ctx *c, double *host_out) {
    auto &p = c->pipe;
    const int b = (int)(p.n & 1);
    int rc;
    if (p.n >= 2) CUDA_TRY(c, cudaStreamWaitEvent(c->stream, p.out_free[b], 0));
    k_state_out<<<nblk(c->nt, 256), 256, 0, c->stream>>>(c->nt, c->cell_old, c->cur[0], c->cur[1], c->cur[2], p.out[b]);
    if ((rc = launch_check(c, "k_state_out"))) return rc;
    CUDA_TRY(c, cudaEventRecord(p.out_ready[b], c->stream));
    CUDA_TRY(c, cudaStreamWaitEvent(p.s_out, p.out_ready[b], 0));
    CUDA_TRY(c, cudaMemcpyAsync(host_out, p.out[b], sizeof(double) * 3 * c->nt, cudaMemcpyDeviceToHost, p.s_out));
    CUDA_TRY(c, cudaEventRecord(p.out_free[b], p.s_out));
    ++p.n;
    return SWE_OK;
}
SWE_API int swe_submit_step_host(swe_ctx *c, const double *host_in, double *host_out, swe_scheme scheme, swe_flux flux,
                                 swe_wavespeed ws, double dt) {
    if (!c || !host_in || !host_out) return SWE_ERR_INVALID;
    if (scheme < SWE_EULER || scheme > SWE_SSPRK3 || !(dt > 0.)) { c->err = "swe_submit_step_host: bad scheme / dt"; return SWE_ERR_INVALID; }
    int rc;
    if ((rc = pipe_begin(c, host_in))) return rc;
    if ((rc = one_step_fwd(c, scheme, flux, ws, dt))) return rc;
    return pipe_end(c, host_out);
}
SWE_API int swe_wait_host(swe_ctx *c) {
    if (!c) return SWE_ERR_INVALID;
    CUDA_TRY(c, cudaSetDevice(c->device));
    if (c->pipe.ok) { CUDA_TRY(c, cudaStreamSynchronize(c->pipe.s_in)); CUDA_TRY(c, cudaStreamSynchronize(c->pipe.s_out)); }
    return swe_synchronize(c);
}

// ---- the stage pieces ----
SWE_API int swe_enable_taps(swe_ctx *c, int on) {
    if (!c) return SWE_ERR_INVALID;
    CUDA_TRY(c, cudaSetDevice(c->device));
    if (on && !c->cew) {
        CUDA_TRY(c, dalloc(&c->cew, (size_t)3 * c->nt));
        CUDA_TRY(c, cudaMemset(c->cew, 0, sizeof(double) * 3 * c->nt));
    }
    if (on && !c->dbg) {
        CUDA_TRY(c, dalloc(&c->dbg, (size_t)BR_COUNT));
        CUDA_TRY(c, cudaMemset(c->dbg, 0, sizeof(unsigned long long) * BR_COUNT));
    }
    c->taps = on != 0;
    return SWE_OK;
}

// ---- flux registry ----
SWE_API int32_t swe_fluxer_count(void) { return kNumFluxers; }
SWE_API const char *swe_fluxer_name(int32_t k) { return (k >= 0 && k < kNumFluxers) ? g_fluxers[k].name : nullptr; }
SWE_API int32_t swe_fluxer_id(int32_t k) { return (k >= 0 && k < kNumFluxers) ? g_fluxers[k].id : -1; }
SWE_API int32_t swe_fluxer_find(const char *name) {
    if (!name) return -1;
    for (int k = 0; k < kNumFluxers; ++k) if (!std::strcmp(g_fluxers[k].name, name)) return g_fluxers[k].id;
    return -1;
}
SWE_API int swe_set_fluxer(swe_ctx *c, int32_t id) {
    if (!c) return SWE_ERR_INVALID;
    if (id >= 0 && !fluxer_known(id)) { c->err = "swe_set_fluxer: no flux with this id is registered (csrc/swe_flux_registry.cuh)"; return SWE_ERR_INVALID; }
    c->fluxer = id < 0 ? -1 : id;
    return SWE_OK;
}

// Semantic-decision switches; every value is bit-checked against the same oracle option.
SWE_API int swe_set_option(swe_ctx *c, const char *key, int32_t value) {
    if (!c || !key) return SWE_ERR_INVALID;
    auto bad = [&](const char *why) { c->err = std::string("swe_set_option(") + key + "): " + why; return SWE_ERR_INVALID; };
    if (!std::strcmp(key, "recon")) { if (value < 0 || value > 2) return bad("0 repaired, 1 as written, 2 first order"); c->opt_recon = value; }
    else if (!std::strcmp(key, "pw2")) { if (value < 0 || value > 1) return bad("0 repaired, 1 as written"); c->opt_pw2 = value; }
    else if (!std::strcmp(key, "roe_fix")) { if (value < 0 || value > 1) return bad("0 as written (cl*ur), 1 cr*ur"); c->opt_roe_fix = value; }
    else if (!std::strcmp(key, "cfl_abs")) { if (value < 0 || value > 1) return bad("0 as written (signed max), 1 magnitudes"); c->opt_cfl_abs = value; }
    else if (!std::strcmp(key, "graph")) { if (value < -1 || value > 1) return bad("-1 auto, 0 off, 1 on"); c->opt_graph = value; }
    else if (!std::strcmp(key, "k1_tiled")) { if (value < 0 || value > 2 * SWE_K1_TILED) return bad("0 gather kernel, 1 TMA-staged shared-memory tiles, 2 cp.async software pipeline"); c->opt_tiled = value; }
    else if (!std::strcmp(key, "fused_drain")) { if (value < 0 || value > 1) return bad("0 separate k_drain pass, 1 draining dt fused into the stage update"); c->opt_fused_drain = value; }
    else if (!std::strcmp(key, "skip_cfl")) { if (value < 0 || value > 1) return bad("0 every stage rebuilds the CFL minimum, 1 only the last stage of a step"); c->opt_skip_cfl = value; }
    else if (!std::strcmp(key, "dry_list")) { if (value < -1 || value > 1) return bad("-1 auto (>= 4M cells), 0 every block of the dry-region update tests its own tile flag, 1 compacted tile list"); c->opt_dry_list = value; }
    else if (!std::strcmp(key, "dry_skip")) {
        if (value < -1 || value > 1) return bad("-1 auto (on while >= 20 % of the cells are dry), 0 every tile is processed, 1 tiles of deep-dry cells are skipped by the flux / draining / update kernels");
        c->opt_dry_skip = value; c->flags_version = 0; c->flags0_valid = false; c->dry_eval_pending = true; drop_tile_flags(c);
        if (value >= 0) c->dry_active = value == 1;
    }
    else return bad("unknown option (recon, pw2, roe_fix, cfl_abs, k1_tiled, fused_drain, skip_cfl, dry_skip, graph)");
    return SWE_OK;
}
SWE_API int swe_get_option(swe_ctx *c, const char *key, int32_t *value) {
    if (!c || !key || !value) return SWE_ERR_INVALID;
    if (!std::strcmp(key, "recon")) *value = c->opt_recon;
    else if (!std::strcmp(key, "pw2")) *value = c->opt_pw2;
    else if (!std::strcmp(key, "roe_fix")) *value = c->opt_roe_fix;
    else if (!std::strcmp(key, "cfl_abs")) *value = c->opt_cfl_abs;
    else if (!std::strcmp(key, "k1_tiled")) *value = c->opt_tiled;
    else if (!std::strcmp(key, "fused_drain")) *value = c->opt_fused_drain;
    else if (!std::strcmp(key, "skip_cfl")) *value = c->opt_skip_cfl;
    else if (!std::strcmp(key, "dry_list")) *value = c->opt_dry_list;
    else if (!std::strcmp(key, "dry_skip")) *value = c->opt_dry_skip;
    else if (!std::strcmp(key, "graph")) *value = c->opt_graph;
    else { c->err = std::string("swe_get_option: unknown option ") + key; return SWE_ERR_INVALID; }
    return SWE_OK;
}
// Branch-hit counters of the last swe_compute_interface_values (taps must be enabled): which branches of
// ReconstructPartWetCell1/2 and ReconstructFullWetCell the cells took; slots as in swe_b200.h.
SWE_API int swe_get_branch_counts(swe_ctx *c, int64_t out12[12]) {
    if (!c || !out12) return SWE_ERR_INVALID;
    if (!c->dbg || !c->taps) { c->err = "swe_get_branch_counts: call swe_enable_taps(ctx, 1) before computing"; return SWE_ERR_INVALID; }
    CUDA_TRY(c, cudaSetDevice(c->device));
    unsigned long long h[BR_COUNT];
    CUDA_TRY(c, cudaMemcpyAsync(h, c->dbg, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    for (int b = 0; b < BR_COUNT; ++b) out12[b] = (int64_t)h[b];
    return SWE_OK;
}

// pass 1 on the cell range [first, last) (DEVICE numbering = caller numbering when the context
// was created with reorder = 0). begin: reset the part-wet work list; finish: run pass 2.
static int interface_values_range(swe_ctx *c, int first, int last, bool begin, bool finish) {
    CUDA_TRY(c, cudaSetDevice(c->device));
    const DevMesh m = dev_mesh(c);
    const DevFields s = dev_fields(c);
    int rc;
    if (begin) {  // work-list counters: part-wet cells [1], generic-reconstruction cells [4]
        CUDA_TRY(c, cudaMemsetAsync(c->flags + 1, 0, sizeof(int), c->stream));
        CUDA_TRY(c, cudaMemsetAsync(c->flags + 4, 0, sizeof(int), c->stream));
        if (c->taps) CUDA_TRY(c, cudaMemsetAsync(c->dbg, 0, sizeof(unsigned long long) * BR_COUNT, c->stream));
        // dry-tile flags: preset to "deep dry", cleared by every cell that is not (off: all zero, nothing is skipped)
        c->k1_dry = c->dry_active && !c->taps && c->opt_tiled == 0;
        if (c->k1_dry) {  // the last pass's flags become `previous`, the new ones start as "deep dry"
            CUDA_TRY(c, cudaMemcpyAsync(c->tile_prev, c->tile_dry, (size_t)c->ntiles, cudaMemcpyDeviceToDevice, c->stream));
            CUDA_TRY(c, cudaMemsetAsync(c->tile_dry, 1, (size_t)c->ntiles, c->stream));
        } else {          // this pass does not maintain the flags: nothing may be assumed by the next one
            CUDA_TRY(c, cudaMemsetAsync(c->tile_dry, 0, (size_t)c->ntiles, c->stream));
        }
        c->k1_covered = 0;
        c->flags_version = 0;
    }
    c->k1_covered += std::max(0, last - first);
    if (last > first) {
        int kt = kt_begin(c, KT_RECONSTRUCT);
        // persistent grid: a multiple of the SM count (148 on B200), never more blocks than work
        const int g1 = std::min(nblk(last - first, kK1Block), c->sms * SWE_K1_GRID_PER_SM);
#if SWE_K1_TILED
        const int ntiles = (last + kTile - 1) / kTile - first / kTile;
        const int gt = std::min(ntiles, c->sms * SWE_K1_GRID_PER_SM);
#define SWE_K1(TAPS) \
        do { \
            if (c->opt_tiled == 2) { \
                if (c->opt_recon == 0) k_reconstruct_pf<TAPS, 0><<<g1, kK1Block, 0, c->stream>>>(m, s, first, last); \
                else if (c->opt_recon == 1) k_reconstruct_pf<TAPS, 1><<<g1, kK1Block, 0, c->stream>>>(m, s, first, last); \
                else k_reconstruct_pf<TAPS, 2><<<g1, kK1Block, 0, c->stream>>>(m, s, first, last); \
            } else if (c->opt_tiled) { \
                if (c->opt_recon == 0) k_reconstruct_tiled<TAPS, 0><<<gt, kTile, 0, c->stream>>>(m, s, first, last); \
                else if (c->opt_recon == 1) k_reconstruct_tiled<TAPS, 1><<<gt, kTile, 0, c->stream>>>(m, s, first, last); \
                else k_reconstruct_tiled<TAPS, 2><<<gt, kTile, 0, c->stream>>>(m, s, first, last); \
            } else if (!TAPS && c->k1_dry) { \
                if (c->opt_recon == 0) k_reconstruct<false, 0, true><<<g1, kK1Block, 0, c->stream>>>(m, s, first, last); \
                else if (c->opt_recon == 1) k_reconstruct<false, 1, true><<<g1, kK1Block, 0, c->stream>>>(m, s, first, last); \
                else k_reconstruct<false, 2, true><<<g1, kK1Block, 0, c->stream>>>(m, s, first, last); \
            } else if (c->opt_recon == 0) k_reconstruct<TAPS, 0><<<g1, kK1Block, 0, c->stream>>>(m, s, first, last); \
            else if (c->opt_recon == 1) k_reconstruct<TAPS, 1><<<g1, kK1Block, 0, c->stream>>>(m, s, first, last); \
            else k_reconstruct<TAPS, 2><<<g1, kK1Block, 0, c->stream>>>(m, s, first, last); \
        } while (0)
#else
#define SWE_K1(TAPS) \
        do { \
            if (c->opt_recon == 0) k_reconstruct<TAPS, 0><<<g1, kK1Block, 0, c->stream>>>(m, s, first, last); \
            else if (c->opt_recon == 1) k_reconstruct<TAPS, 1><<<g1, kK1Block, 0, c->stream>>>(m, s, first, last); \
            else k_reconstruct<TAPS, 2><<<g1, kK1Block, 0, c->stream>>>(m, s, first, last); \
        } while (0)
#endif
        if (c->taps) SWE_K1(true); else SWE_K1(false);
#undef SWE_K1
        kt_end(c, kt);
        if ((rc = launch_check(c, "k_reconstruct"))) return rc;
    }
    if (finish) {
#if SWE_K1_SPLIT
        int kts = kt_begin(c, KT_PARTWET2);  // accounted with the part-wet pass: both are O(front) list kernels
        if (c->taps) k_reconstruct_slow<true><<<kPw2Blocks, kBlock, 0, c->stream>>>(m, s);
        else k_reconstruct_slow<false><<<kPw2Blocks, kBlock, 0, c->stream>>>(m, s);
        kt_end(c, kts);
        if ((rc = launch_check(c, "k_reconstruct_slow"))) return rc;
#endif
        int kt = kt_begin(c, KT_PARTWET2);
        if (c->taps) k_partwet2<true><<<kPw2Blocks, kBlock, 0, c->stream>>>(m, s);
        else k_partwet2<false><<<kPw2Blocks, kBlock, 0, c->stream>>>(m, s);
        kt_end(c, kt);
        if ((rc = launch_check(c, "k_partwet2"))) return rc;
        // the flags describe the current state once every cell went through pass 1 (ranges may come in any order)
        if (c->k1_dry && c->k1_covered == (int64_t)c->nt) c->flags_version = c->state_version;
        if (c->k1_dry && c->k1_covered != (int64_t)c->nt)  // partial pass: untouched tiles still carry the preset
            CUDA_TRY(c, cudaMemsetAsync(c->tile_dry, 0, (size_t)c->ntiles, c->stream));
    }
    return SWE_OK;
}

SWE_API int swe_compute_interface_values(swe_ctx *c) {
    if (!c) return SWE_ERR_INVALID;
    return interface_values_range(c, 0, c->nt, true, true);
}

SWE_API int swe_compute_interface_values_class(swe_ctx *c, int32_t cls, int begin, int finish) {
    if (!c || cls < 0 || cls > 3) return SWE_ERR_INVALID;
    return interface_values_range(c, c->class_first[cls], c->class_first[cls + 1], begin != 0, finish != 0);
}

SWE_API int swe_compute_interface_values_range(swe_ctx *c, int64_t first_cell, int64_t last_cell, int begin, int finish) {
    if (!c) return SWE_ERR_INVALID;
    if (c->reordered) { c->err = "swe_compute_interface_values_range: needs a context created with reorder = 0"; return SWE_ERR_INVALID; }
    if (first_cell < 0 || last_cell > c->nt || first_cell > last_cell) { c->err = "swe_compute_interface_values_range: bad range"; return SWE_ERR_INVALID; }
    return interface_values_range(c, (int)first_cell, (int)last_cell, begin != 0, finish != 0);
}

// cfl = false (internal, non-final stages of a step): fluxes only, min_length_to_wavespeed is left alone
static int compute_fluxes(swe_ctx *c, swe_flux flux, swe_wavespeed ws, bool cfl);
SWE_API int swe_compute_fluxes(swe_ctx *c, swe_flux flux, swe_wavespeed ws) { return compute_fluxes(c, flux, ws, true); }
static int compute_fluxes(swe_ctx *c, swe_flux flux, swe_wavespeed ws, bool cfl) {
    if (!c) return SWE_ERR_INVALID;
    int id = c->fluxer;  // swe_set_fluxer overrides the enum pair
    if (id < 0) {
        if ((flux != SWE_HLL && flux != SWE_HLLC) || ws < SWE_RUSANOV || ws > SWE_EINFELDT) {
            c->err = "swe_compute_fluxes: unknown flux / wavespeed (built in: HLL, HLLC x Rusanov, Davis, Einfeldt; others via swe_set_fluxer)";
            return SWE_ERR_INVALID;
        }
        id = 3 * (int)flux + (int)ws;
    }
    CUDA_TRY(c, cudaSetDevice(c->device));
    const DevMesh m = dev_mesh(c);
    const DevFields s = dev_fields(c);
    const int kt = kt_begin(c, KT_FLUX);
    launch_flux(c, m, s, id, cfl || c->taps || !c->opt_skip_cfl);
    if (dry_now(c)) CUDA_TRY(c, cudaMemcpyAsync(c->tile_fluxed, c->tile_dry, (size_t)c->ntiles, cudaMemcpyDeviceToDevice, c->stream));
    else if (c->dry_active) CUDA_TRY(c, cudaMemsetAsync(c->tile_fluxed, 0, (size_t)c->ntiles, c->stream));
    kt_end(c, kt);
    return launch_check(c, "k_flux");
}

SWE_API int swe_save_state(swe_ctx *c) {
    if (!c) return SWE_ERR_INVALID;
    c->sav = c->cur;  // no copy: the next stage update writes the other buffer
    c->saved_pending = true;
    // the dry-tile flags of the saved state (the non-plain stage updates combine U0 into the result)
    c->flags0_valid = dry_now(c);
    if (c->flags0_valid) {
        CUDA_TRY(c, cudaSetDevice(c->device));
        CUDA_TRY(c, cudaMemcpyAsync(c->tile_dry0, c->tile_dry, (size_t)c->ntiles, cudaMemcpyDeviceToDevice, c->stream));
    }
    return SWE_OK;
}

// K3 over all cells, then the choice of the output buffer of this stage (first stage after
// swe_save_state: the other buffer, so that U0 stays intact without a copy kernel)
static int stage_drain(swe_ctx *c, double ***outb_out) {
    CUDA_TRY(c, cudaSetDevice(c->device));
    const DevMesh m = dev_mesh(c);
    const DevFields s = dev_fields(c);
    int rc;
    int kt = kt_begin(c, KT_DRAIN);
    if (c->opt_fused_drain && !c->taps) {  // fused form: only the cells read across tile / class borders
        if (c->drain_count > 0)
            k_drain_list<<<std::min(nblk(c->drain_count, kBlock), c->sms * 16), kBlock, 0, c->stream>>>(m, s, c->drain_list, c->drain_count);
        c->dti_complete = false;
    } else {  // taps: swe_get_draining_dt wants every cell (the fused update still computes its own copy)
        if (dry_now(c)) k_drain<true><<<drain_grid(c), kBlock, 0, c->stream>>>(m, s);
        else k_drain<false><<<drain_grid(c), kBlock, 0, c->stream>>>(m, s);
        c->dti_complete = true;
        if (dry_now(c)) CUDA_TRY(c, cudaMemcpyAsync(c->tile_drained, c->tile_dry, (size_t)c->ntiles, cudaMemcpyDeviceToDevice, c->stream));
        else if (c->dry_active) CUDA_TRY(c, cudaMemsetAsync(c->tile_drained, 0, (size_t)c->ntiles, c->stream));
    }
    kt_end(c, kt);
    if ((rc = launch_check(c, "k_drain"))) return rc;
    double **outb = c->cur;
    if (c->saved_pending && c->sav == c->cur) {
        outb = (c->cur == c->bufA) ? c->bufB : c->bufA;
        c->saved_pending = false;
    }
    *outb_out = outb;
    return SWE_OK;
}
// K4 on the device cell range [first, last)
static int stage_update_range(swe_ctx *c, double **outb, double a0, double a1, double dt_host, double dt_coef, int first, int last) {
    if (last <= first) return SWE_OK;
    const DevMesh m = dev_mesh(c);
    const DevFields s = dev_fields(c);
    const bool fused = c->opt_fused_drain != 0;
    const bool dry = dry_now(c);
    // fused: one block per ABSOLUTE 128-cell tile touching [first, last)
    const int g = fused ? ((last - 1) >> kUpdTileShift) - (first >> kUpdTileShift) + 1 : nblk(last - first, kBlock);
    const int kt = kt_begin(c, KT_UPDATE);
    const bool cor_on = c->cor != 0.;
    const bool listed = use_tile_list(c);
    const int g_dry = listed ? c->ntiles : g;  // listed form: one block per listed tile, the count lives on the device
    if (dry && !fused && listed) {
        const int use_td0 = a0 == 0. ? 0 : 1;
        CUDA_TRY(c, cudaMemsetAsync(c->tile_list, 0, sizeof(int), c->stream));
        k_tile_compact<<<nblk(c->ntiles, 256), 256, 0, c->stream>>>(c->ntiles, s.td, s.td0, use_td0, c->tile_list);
        if (outb[0] != s.w)  // out of place (first stage after swe_save_state): the skipped tiles keep (cb, +0, +0)
            k_fill_dry_tiles<<<nblk(last - first, 256), 256, 0, c->stream>>>(first, last, s.td, s.td0, use_td0, c->cb, outb[0], outb[1], outb[2]);
    }
#define SWE_UPD(PLAIN, COR, W0, U0, V0) \
    do { \
        if (fused) k_update<PLAIN, COR, false, true><<<g, kBlock, 0, c->stream>>>(m, s, W0, U0, V0, outb[0], outb[1], outb[2], a0, a1, dt_host, dt_coef, c->cor, first, last); \
        else if (dry) k_update<PLAIN, COR, false, false, true><<<g_dry, kBlock, 0, c->stream>>>(m, s, W0, U0, V0, outb[0], outb[1], outb[2], a0, a1, dt_host, dt_coef, c->cor, first, last); \
        else k_update<PLAIN, COR><<<g, kBlock, 0, c->stream>>>(m, s, W0, U0, V0, outb[0], outb[1], outb[2], a0, a1, dt_host, dt_coef, c->cor, first, last); \
    } while (0)
    if (a0 == 0.) {
        if (cor_on) SWE_UPD(true, true, nullptr, nullptr, nullptr); else SWE_UPD(true, false, nullptr, nullptr, nullptr);
    } else {
        if (cor_on) SWE_UPD(false, true, c->sav[0], c->sav[1], c->sav[2]); else SWE_UPD(false, false, c->sav[0], c->sav[1], c->sav[2]);
    }
#undef SWE_UPD
    kt_end(c, kt);
    return launch_check(c, "k_update");
}
static int stage_update(swe_ctx *c, double a0, double a1, double dt_host, double dt_coef) {
    double **outb = nullptr;
    int rc;
    if ((rc = stage_drain(c, &outb))) return rc;
    if ((rc = stage_update_range(c, outb, a0, a1, dt_host, dt_coef, 0, c->nt))) return rc;
    c->cur = outb;
    c->state_version++;
    return SWE_OK;
}

SWE_API int swe_stage_update(swe_ctx *c, double a0, double a1, double dt_stage) {
    if (!c) return SWE_ERR_INVALID;
    if (a0 == 0. && a1 != 1.) { c->err = "swe_stage_update: a0 == 0 requires a1 == 1 (plain U + RHS)"; return SWE_ERR_INVALID; }
    return stage_update(c, a0, a1, dt_stage, 0.);
}
// stage dt = coef * device-resident dt (set by swe_set_dt / swe_advance_dt)
SWE_API int swe_stage_update_dev(swe_ctx *c, double a0, double a1, double coef) {
    if (!c || coef == 0.) return SWE_ERR_INVALID;
    return stage_update(c, a0, a1, 0., coef);
}
SWE_API int swe_set_dt(swe_ctx *c, double dt) {
    if (!c) return SWE_ERR_INVALID;
    CUDA_TRY(c, cudaSetDevice(c->device));
    k_set_scalar<<<1, 1, 0, c->stream>>>(c->scal + 1, dt);
    return launch_check(c, "k_set_scalar");
}
// time += dt; if adaptive: dt = 0.15 * min_len_to_wavespeed (after the global min all-reduce)
SWE_API int swe_advance_dt(swe_ctx *c, int adaptive, double dt_fixed) {
    if (!c) return SWE_ERR_INVALID;
    CUDA_TRY(c, cudaSetDevice(c->device));
    k_post_step<<<1, 1, 0, c->stream>>>(dev_fields(c), dt_fixed, adaptive, 0);
    return launch_check(c, "k_post_step");
}

static int read_scalar_fwd(swe_ctx *c, int idx, double *v);
static int one_step(swe_ctx *c, swe_scheme scheme, swe_flux flux, swe_wavespeed ws, double dt, bool dev_dt);
static int one_step_fwd(swe_ctx *c, swe_scheme scheme, swe_flux flux, swe_wavespeed ws, double dt) { return one_step(c, scheme, flux, ws, dt, false); }
static int one_step(swe_ctx *c, swe_scheme scheme, swe_flux flux, swe_wavespeed ws, double dt, bool dev_dt) {
    int rc;
    auto upd = [&](double a0, double a1, double coef) {
        return dev_dt ? stage_update(c, a0, a1, 0., coef) : stage_update(c, a0, a1, coef * dt, 0.);
    };
    if ((rc = swe_compute_interface_values(c))) return rc;
    if ((rc = compute_fluxes(c, flux, ws, scheme == SWE_EULER))) return rc;  // the CFL minimum of the LAST stage is the step's
    if (scheme == SWE_EULER) return dev_dt ? stage_update(c, 0., 1., 0., 1.) : stage_update(c, 0., 1., dt, 0.);
    swe_save_state(c);
    // first stage: U0 + RHS(dt)
    if ((rc = (dev_dt ? stage_update(c, 0., 1., 0., 1.) : stage_update(c, 0., 1., dt, 0.)))) return rc;
    if ((rc = swe_compute_interface_values(c))) return rc;
    if ((rc = compute_fluxes(c, flux, ws, scheme == SWE_SSPRK2))) return rc;
    if (scheme == SWE_SSPRK2) return upd(0.5, 0.5, 0.5);
    if ((rc = upd(0.75, 0.25, 0.25))) return rc;
    if ((rc = swe_compute_interface_values(c))) return rc;
    if ((rc = swe_compute_fluxes(c, flux, ws))) return rc;
    return upd((1. / 3.), (2. / 3.), (2. / 3.));
}

SWE_API int swe_step(swe_ctx *c, swe_scheme scheme, swe_flux flux, swe_wavespeed ws, double dt) {
    if (!c) return SWE_ERR_INVALID;
    if (scheme < SWE_EULER || scheme > SWE_SSPRK3) { c->err = "swe_step: unknown scheme"; return SWE_ERR_INVALID; }
    if (!(dt > 0.)) { c->err = "swe_step: dt must be positive"; return SWE_ERR_INVALID; }
    int rc = dry_refresh(c);
    if (rc) return rc;
    rc = one_step(c, scheme, flux, ws, dt, false);
    if (rc) return rc;
    return swe_advance_dt(c, 0, dt);
}

// Launch-bound meshes (configs[0]-[1]: tens of microseconds of kernel time per step against ~13 launches): one
// whole time step is captured into a CUDA graph and replayed. The two state buffers swap roles from step to step
// (the first stage writes the other buffer so that U0 stays intact), hence one graph per parity.
constexpr int kGraphAutoCells = 4 * 1024 * 1024;
static int run_graphed(swe_ctx *c, swe_scheme scheme, swe_flux flux, swe_wavespeed ws, int64_t nsteps, double dt, bool adaptive) {
    int rc;
    cudaStream_t user = c->stream;
    if (user == 0) {  // the legacy default stream cannot be captured: run on an own stream, ordered after / before it
        if (!c->gstream) {
            CUDA_TRY(c, cudaStreamCreateWithFlags(&c->gstream, cudaStreamNonBlocking));
            CUDA_TRY(c, cudaEventCreateWithFlags(&c->gev, cudaEventDisableTiming));
        }
        CUDA_TRY(c, cudaEventRecord(c->gev, 0));
        CUDA_TRY(c, cudaStreamWaitEvent(c->gstream, c->gev, 0));
        c->stream = c->gstream;
    }
    auto finish = [&](int code) {
        if (user == 0) {
            cudaEventRecord(c->gev, c->gstream);
            cudaStreamWaitEvent(0, c->gev, 0);
            c->stream = user;
        }
        return code;
    };
    const int fluxer = c->fluxer >= 0 ? c->fluxer : 3 * (int)flux + (int)ws;
    const int opts = c->opt_recon | (c->opt_pw2 << 2) | (c->opt_roe_fix << 3) | (c->opt_cfl_abs << 4) | (c->opt_tiled << 8) | ((c->taps ? 1 : 0) << 6) | (c->opt_fused_drain << 7) | (c->opt_skip_cfl << 10) | ((c->dry_active ? 1 : 0) << 11) | ((use_tile_list(c) ? 1 : 0) << 12);
    for (int64_t s = 0; s < nsteps; ++s) {
        const int parity = (c->cur == c->bufA) ? 0 : 1;
        swe_ctx::StepGraph *g = nullptr;
        for (auto &q : c->graphs)
            if (q.scheme == (int)scheme && q.fluxer == fluxer && q.adaptive == (int)adaptive && q.parity == parity && q.opts == opts &&
                (adaptive || q.dt == dt)) { g = &q; break; }
        if (!g) {
            if (c->graphs.size() >= 16) {  // many different fixed dt values: stop caching, plain launches
                if ((rc = one_step(c, scheme, flux, ws, dt, adaptive))) return finish(rc);
                if ((rc = swe_advance_dt(c, adaptive ? 1 : 0, dt))) return finish(rc);
                continue;
            }
            const int64_t l0 = c->launches;
            cudaGraph_t graph = nullptr;
            CUDA_TRY(c, cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
            rc = one_step(c, scheme, flux, ws, dt, adaptive);
            if (!rc) rc = swe_advance_dt(c, adaptive ? 1 : 0, dt);
            cudaError_t e = cudaStreamEndCapture(c->stream, &graph);
            if (rc) { if (graph) cudaGraphDestroy(graph); return finish(rc); }
            if (e != cudaSuccess) { c->err = std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e); return finish(SWE_ERR_CUDA); }
            swe_ctx::StepGraph ng{(int)scheme, fluxer, (int)adaptive, parity, opts, dt, nullptr, c->launches - l0, c->cur == c->bufA};
            e = cudaGraphInstantiate(&ng.exec, graph, 0);
            cudaGraphDestroy(graph);
            if (e != cudaSuccess) { c->err = std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e); return finish(SWE_ERR_CUDA); }
            c->graphs.push_back(ng);
            g = &c->graphs.back();
            c->launches = l0;  // counted when the graph is launched
        }
        CUDA_TRY(c, cudaGraphLaunch(g->exec, c->stream));
        c->launches += g->launches;
        // host-side bookkeeping of what the step did to the buffers
        c->cur = g->cur_is_a_after ? c->bufA : c->bufB;
        c->state_version++;  // the replayed step ended with a stage update: its flags no longer describe the state
        c->sav = (scheme == SWE_EULER) ? c->sav : (parity == 0 ? c->bufA : c->bufB);
        c->saved_pending = false;
    }
    return finish(SWE_OK);
}

SWE_API int swe_run(swe_ctx *c, swe_scheme scheme, swe_flux flux, swe_wavespeed ws, int64_t nsteps, double dt, double dt0) {
    if (!c) return SWE_ERR_INVALID;
    if (scheme < SWE_EULER || scheme > SWE_SSPRK3) { c->err = "swe_run: unknown scheme"; return SWE_ERR_INVALID; }
    const bool adaptive = !(dt > 0.);
    int rc;
    // adaptive, dt0 > 0: first step dt0; dt0 <= 0: continue with the dt already on the device (previous swe_run /
    // swe_checkpoint_load), which makes a restarted adaptive run bit-identical to the uninterrupted one
    if (adaptive && dt0 > 0. && (rc = swe_set_dt(c, dt0))) return rc;
    if (adaptive && !(dt0 > 0.)) {
        double cur = 0.;
        if ((rc = read_scalar_fwd(c, 1, &cur))) return rc;
        if (!(cur > 0.)) { c->err = "swe_run: adaptive mode needs dt0 > 0 (no dt stored on the device yet)"; return SWE_ERR_INVALID; }
    }
    c->steps_since_eval += nsteps;
    if (c->steps_since_eval >= 1024) c->dry_eval_pending = true;  // the shoreline moves: look again now and then
    if ((rc = dry_refresh(c))) return rc;
    const bool use_graph = (c->opt_graph == 1 || (c->opt_graph < 0 && c->nt <= kGraphAutoCells)) && !c->ktiming && nsteps >= 4;
    if (use_graph) return run_graphed(c, scheme, flux, ws, nsteps, dt, adaptive);
    for (int64_t s = 0; s < nsteps; ++s) {
        if ((rc = one_step(c, scheme, flux, ws, dt, adaptive))) return rc;
        if ((rc = swe_advance_dt(c, adaptive ? 1 : 0, dt))) return rc;
    }
    return SWE_OK;
}

static int read_scalar(swe_ctx *c, int idx, double *v) {
    if (!c || !v) return SWE_ERR_INVALID;
    CUDA_TRY(c, cudaSetDevice(c->device));
    CUDA_TRY(c, cudaMemcpyAsync(v, c->scal + idx, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return SWE_OK;
}
static int read_scalar_fwd(swe_ctx *c, int idx, double *v) { return read_scalar(c, idx, v); }
SWE_API int swe_get_min_len_to_wavespeed(swe_ctx *c, double *v) { return read_scalar(c, 0, v); }
SWE_API int swe_get_dt(swe_ctx *c, double *dt) { return read_scalar(c, 1, dt); }
SWE_API int swe_cfl_dt(swe_ctx *c, double *dt) {
    double v = 0;
    int rc = read_scalar(c, 0, &v);
    if (rc) return rc;
    *dt = SWE_CFL * v;  // m_constCFL (include/TimeDisc.h:22)
    return SWE_OK;
}
SWE_API int swe_get_time(swe_ctx *c, double *t) { return read_scalar(c, 2, t); }
SWE_API int64_t swe_launch_count(const swe_ctx *c) { return c ? c->launches : 0; }
SWE_API int swe_set_min_len_to_wavespeed(swe_ctx *c, double v) {
    if (!c) return SWE_ERR_INVALID;
    CUDA_TRY(c, cudaSetDevice(c->device));
    k_set_scalar<<<1, 1, 0, c->stream>>>(c->scal, v);
    return launch_check(c, "k_set_scalar");
}
SWE_API int swe_min_len_device_ptr(swe_ctx *c, void **p) {
    if (!c || !p) return SWE_ERR_INVALID;
    *p = c->scal;
    return SWE_OK;
}

// ---- taps ----
static int tap_out(swe_ctx *c, double *host, size_t n_doubles) {
    CUDA_TRY(c, cudaMemcpyAsync(host, c->stage_aos, sizeof(double) * n_doubles, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return SWE_OK;
}
static int edge_tap(swe_ctx *c, double *out, int which) {
    if (!c || !out) return SWE_ERR_INVALID;
    if (which == 0 && !c->cew) { c->err = "swe_get_edge_states: call swe_enable_taps(ctx, 1) before computing"; return SWE_ERR_INVALID; }
    CUDA_TRY(c, cudaSetDevice(c->device));
    const size_t n = (size_t)6 * c->ne;
    int rc = ensure_stage(c, std::max(n, (size_t)3 * c->nt));
    if (rc) return rc;
    CUDA_TRY(c, cudaMemsetAsync(c->stage_aos, 0, sizeof(double) * n, c->stream));
    k_edge_out<<<nblk(c->nt, 256), 256, 0, c->stream>>>(dev_mesh(c), dev_fields(c), c->cell_old, c->edge_old, which, c->cor, c->stage_aos);
    if ((rc = launch_check(c, "k_edge_out"))) return rc;
    return tap_out(c, out, n);
}
SWE_API int swe_get_edge_states(swe_ctx *c, double *out) { return edge_tap(c, out, 0); }
SWE_API int swe_get_sources(swe_ctx *c, double *out) { return edge_tap(c, out, 1); }
SWE_API int swe_get_fluxes(swe_ctx *c, double *out) {
    if (!c || !out) return SWE_ERR_INVALID;
    CUDA_TRY(c, cudaSetDevice(c->device));
    int rc = ensure_stage(c, std::max((size_t)3 * c->ne, (size_t)3 * c->nt));
    if (rc) return rc;
    k_flux_out<<<nblk(c->ne, 256), 256, 0, c->stream>>>(c->ne, c->edge_old, c->f0, c->f1, c->f2, c->stage_aos);
    if ((rc = launch_check(c, "k_flux_out"))) return rc;
    return tap_out(c, out, (size_t)3 * c->ne);
}
static int scalar_tap(swe_ctx *c, double *out, const double *src, int n, const int *old) {
    if (!c || !out) return SWE_ERR_INVALID;
    CUDA_TRY(c, cudaSetDevice(c->device));
    int rc = ensure_stage(c, (size_t)3 * c->nt);
    if (rc) return rc;
    if ((size_t)n > c->stage_cap && (rc = ensure_stage(c, (size_t)n))) return rc;
    k_scalar_out<<<nblk(n, 256), 256, 0, c->stream>>>(n, old, src, c->stage_aos);
    if ((rc = launch_check(c, "k_scalar_out"))) return rc;
    return tap_out(c, out, (size_t)n);
}
SWE_API int swe_get_node_max_w(swe_ctx *c, double *out) {
    if (!c || !out) return SWE_ERR_INVALID;
    CUDA_TRY(c, cudaSetDevice(c->device));
    int rc = ensure_stage(c, (size_t)3 * c->nt);
    if (rc) return rc;
    k_node_maxw_out<<<nblk(c->nn, 256), 256, 0, c->stream>>>(dev_mesh(c), dev_fields(c), c->node_old, c->stage_aos);
    if ((rc = launch_check(c, "k_node_maxw_out"))) return rc;
    return tap_out(c, out, (size_t)c->nn);
}
SWE_API int swe_get_draining_dt(swe_ctx *c, double *out) {
    if (!c) return SWE_ERR_INVALID;
    if (!c->dti_complete) {
        c->err = "swe_get_draining_dt: the fused stage update keeps the draining dt on chip; call swe_enable_taps(ctx, 1) before the "
                 "stage, or swe_compute_rhs, to materialise it for every cell";
        return SWE_ERR_INVALID;
    }
    return scalar_tap(c, out, c->dti, c->nt, c->cell_old);
}
SWE_API int swe_get_cell_class(swe_ctx *c, int8_t *out) {
    if (!c || !out) return SWE_ERR_INVALID;
    CUDA_TRY(c, cudaSetDevice(c->device));
    int rc = ensure_stage(c, (size_t)3 * c->nt);
    if (rc) return rc;
    k_cls_out<<<nblk(c->nt, 256), 256, 0, c->stream>>>(c->nt, c->cell_old, c->cls, (signed char *)c->stage_aos);
    if ((rc = launch_check(c, "k_cls_out"))) return rc;
    CUDA_TRY(c, cudaMemcpyAsync(out, c->stage_aos, (size_t)c->nt, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return SWE_OK;
}

SWE_API int swe_diagnostics(swe_ctx *c, double out[6]) {
    if (!c || !out) return SWE_ERR_INVALID;
    CUDA_TRY(c, cudaSetDevice(c->device));
    k_diag_partial<<<kDiagBlocks, kDiagThreads, 0, c->stream>>>(dev_mesh(c), dev_fields(c), c->diag);
    int rc;
    if ((rc = launch_check(c, "k_diag_partial"))) return rc;
    k_diag_final<<<1, 32, 0, c->stream>>>(c->diag, c->diag + 6 * kDiagBlocks);
    if ((rc = launch_check(c, "k_diag_final"))) return rc;
    CUDA_TRY(c, cudaMemcpyAsync(out, c->diag + 6 * kDiagBlocks, sizeof(double) * 6, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return SWE_OK;
}

// ---- per-cell accessors of the reference API (TimeDisc::RHS, MUSCLObject::Is*Cell) ----
// Cell classes of the CURRENT state (0 dry, 1 part-wet, 2 full-wet), independent of the last reconstruction.
SWE_API int swe_classify(swe_ctx *c, int8_t *out) {
    if (!c || !out) return SWE_ERR_INVALID;
    CUDA_TRY(c, cudaSetDevice(c->device));
    int rc = ensure_stage(c, (size_t)3 * c->nt);
    if (rc) return rc;
    signed char *tmp = (signed char *)c->stage_aos, *tmp2 = tmp + c->nt;
    k_classify<<<nblk(c->nt, 256), 256, 0, c->stream>>>(dev_mesh(c), c->cur[0], tmp);
    if ((rc = launch_check(c, "k_classify"))) return rc;
    k_cls_out<<<nblk(c->nt, 256), 256, 0, c->stream>>>(c->nt, c->cell_old, tmp, tmp2);
    if ((rc = launch_check(c, "k_cls_out"))) return rc;
    CUDA_TRY(c, cudaMemcpyAsync(out, tmp2, (size_t)c->nt, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return SWE_OK;
}
// TimeDisc::RHS(i, dt) of every cell (src/TimeDisc.cpp:3-41) for the fluxes / edge values of the last
// swe_compute_interface_values + swe_compute_fluxes and the current state: 3 x nt column-major, caller
// numbering. Also refreshes the draining time steps (swe_get_draining_dt). The state is not changed.
SWE_API int swe_compute_rhs(swe_ctx *c, double dt, double *rhs) {
    if (!c || !rhs) return SWE_ERR_INVALID;
    CUDA_TRY(c, cudaSetDevice(c->device));
    int rc = ensure_stage(c, (size_t)6 * c->nt);
    if (rc) return rc;
    const DevMesh m = dev_mesh(c);
    const DevFields s = dev_fields(c);
    k_drain<false><<<drain_grid(c), kBlock, 0, c->stream>>>(m, s);
    if ((rc = launch_check(c, "k_drain"))) return rc;
    c->dti_complete = true;
    double *r0 = c->stage_aos, *r1 = r0 + c->nt, *r2 = r1 + c->nt, *aos = r2 + c->nt;
    const int g = nblk(c->nt, kBlock);
    if (c->cor != 0.) k_update<true, true, true><<<g, kBlock, 0, c->stream>>>(m, s, nullptr, nullptr, nullptr, r0, r1, r2, 0., 1., dt, 0., c->cor, 0, c->nt);
    else k_update<true, false, true><<<g, kBlock, 0, c->stream>>>(m, s, nullptr, nullptr, nullptr, r0, r1, r2, 0., 1., dt, 0., c->cor, 0, c->nt);
    if ((rc = launch_check(c, "k_update<rhs>"))) return rc;
    k_state_out<<<nblk(c->nt, 256), 256, 0, c->stream>>>(c->nt, c->cell_old, r0, r1, r2, aos);
    if ((rc = launch_check(c, "k_state_out"))) return rc;
    CUDA_TRY(c, cudaMemcpyAsync(rhs, aos, sizeof(double) * 3 * c->nt, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return SWE_OK;
}
SWE_API int swe_set_time(swe_ctx *c, double t) {
    if (!c) return SWE_ERR_INVALID;
    CUDA_TRY(c, cudaSetDevice(c->device));
    k_set_scalar<<<1, 1, 0, c->stream>>>(c->scal + 2, t);
    return launch_check(c, "k_set_scalar");
}

// ---- binary checkpoint / restart (upstream only has text dumps, examples/Main.cpp:65-73) ----
// File: "SWEB200C" magic, version, nt, ne, nn, cor, time, dt, min_len_to_wavespeed, 4 option words, then the state
// 3 x nt fp64 in the CALLER's numbering (so a checkpoint can be loaded into a context with another device numbering).
namespace {
struct CkptHeader {
    char magic[8];
    int32_t version, pad;
    int64_t nt, ne, nn;
    double cor, time, dt, min_len;
    int32_t opt[4];
};
}  // namespace
SWE_API int swe_checkpoint_save(swe_ctx *c, const char *path) {
    if (!c || !path) return SWE_ERR_INVALID;
    CUDA_TRY(c, cudaSetDevice(c->device));
    std::vector<double> st((size_t)3 * c->nt);
    int rc = swe_get_state(c, st.data());
    if (rc) return rc;
    double scal[4];
    CUDA_TRY(c, cudaMemcpy(scal, c->scal, sizeof(scal), cudaMemcpyDeviceToHost));
    CkptHeader h{};
    std::memcpy(h.magic, "SWEB200C", 8);
    h.version = 1; h.nt = c->nt; h.ne = c->ne; h.nn = c->nn; h.cor = c->cor;
    h.min_len = scal[0]; h.dt = scal[1]; h.time = scal[2];
    h.opt[0] = c->opt_recon; h.opt[1] = c->opt_pw2; h.opt[2] = c->opt_roe_fix; h.opt[3] = c->opt_cfl_abs;
    FILE *f = std::fopen(path, "wb");
    if (!f) { c->err = std::string("swe_checkpoint_save: cannot open ") + path; return SWE_ERR_IO; }
    const bool ok = std::fwrite(&h, sizeof(h), 1, f) == 1 && std::fwrite(st.data(), sizeof(double), st.size(), f) == st.size();
    if (std::fclose(f) != 0 || !ok) { c->err = std::string("swe_checkpoint_save: short write to ") + path; return SWE_ERR_IO; }
    return SWE_OK;
}
SWE_API int swe_checkpoint_load(swe_ctx *c, const char *path) {
    if (!c || !path) return SWE_ERR_INVALID;
    CUDA_TRY(c, cudaSetDevice(c->device));
    FILE *f = std::fopen(path, "rb");
    if (!f) { c->err = std::string("swe_checkpoint_load: cannot open ") + path; return SWE_ERR_IO; }
    CkptHeader h{};
    std::vector<double> st((size_t)3 * c->nt);
    bool ok = std::fread(&h, sizeof(h), 1, f) == 1 && std::memcmp(h.magic, "SWEB200C", 8) == 0 && h.version == 1;
    if (ok && (h.nt != c->nt || h.ne != c->ne || h.nn != c->nn)) {
        std::fclose(f);
        c->err = "swe_checkpoint_load: the checkpoint belongs to a different mesh";
        return SWE_ERR_INVALID;
    }
    ok = ok && std::fread(st.data(), sizeof(double), st.size(), f) == st.size();
    std::fclose(f);
    if (!ok) { c->err = std::string("swe_checkpoint_load: not a checkpoint / truncated: ") + path; return SWE_ERR_IO; }
    if (h.cor != c->cor || h.opt[0] != c->opt_recon || h.opt[1] != c->opt_pw2 || h.opt[2] != c->opt_roe_fix || h.opt[3] != c->opt_cfl_abs) {
        c->err = "swe_checkpoint_load: the checkpoint was written with other solver settings (cor / recon / pw2 / roe_fix / cfl_abs)";
        return SWE_ERR_INVALID;
    }
    int rc = swe_set_state(c, st.data());  // resets the time, restored next
    if (rc) return rc;
    const double scal[3] = {h.min_len, h.dt, h.time};
    CUDA_TRY(c, cudaMemcpy(c->scal, scal, sizeof(scal), cudaMemcpyHostToDevice));
    return SWE_OK;
}

// ---- per-kernel timing (CUDA events on the ctx stream) ----
SWE_API int swe_kernel_timing(swe_ctx *c, int enable) {
    if (!c) return SWE_ERR_INVALID;
    CUDA_TRY(c, cudaSetDevice(c->device));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    for (auto &p : c->kt_pairs) c->kt_pool.push_back({p.a, p.b});
    c->kt_pairs.clear();
    c->ktiming = enable != 0;
    return SWE_OK;
}
SWE_API int swe_kernel_times(swe_ctx *c, int32_t max_kinds, double *ms_total, int64_t *counts, const char **names) {
    if (!c || !ms_total || !counts || max_kinds < KT_COUNT) return SWE_ERR_INVALID;
    CUDA_TRY(c, cudaSetDevice(c->device));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    for (int k = 0; k < KT_COUNT; ++k) { ms_total[k] = 0.; counts[k] = 0; if (names) names[k] = kt_names[k]; }
    for (auto &p : c->kt_pairs) {
        float ms = 0.f;
        CUDA_TRY(c, cudaEventElapsedTime(&ms, p.a, p.b));
        ms_total[p.id] += ms; counts[p.id]++;
    }
    return KT_COUNT;
}

// ---- analytic cases on the device ----
SWE_API int swe_case_set_bathymetry_device(swe_ctx *c, const swe_case *cs) {
    if (!c || !cs || cs->kind < 0 || cs->kind > SWE_CASE_BOWL_HUMP) return SWE_ERR_INVALID;
    CUDA_TRY(c, cudaSetDevice(c->device));
    c->state_version++;  // the bed (and with it every cell class) changes
    c->dry_eval_pending = true;
    drop_tile_flags(c);
    k_case_bathymetry<<<nblk(c->nn, 256), 256, 0, c->stream>>>(c->nn, c->node, *cs);
    int rc;
    if ((rc = launch_check(c, "k_case_bathymetry"))) return rc;
    // geometry that depends on the bed: cgeo (cb, bfull) and cb
    k_setup_cells<<<nblk(c->nt, 256), 256, 0, c->stream>>>(c->nt, c->tp, c->tt, c->node, c->cgeo, c->area, c->cb);
    return launch_check(c, "k_setup_cells");
}
SWE_API int swe_case_initial_state_device(swe_ctx *c, const swe_case *cs, int32_t quad_n, double t) {
    if (!c || !cs || quad_n < 1 || cs->kind < 0 || cs->kind > SWE_CASE_BOWL_HUMP) return SWE_ERR_INVALID;
    CUDA_TRY(c, cudaSetDevice(c->device));
    c->state_version++;
    c->dry_eval_pending = true;
    drop_tile_flags(c);
    k_case_init<<<nblk(c->nt, 128), 128, 0, c->stream>>>(dev_mesh(c), *cs, quad_n, t, c->cur[0], c->cur[1], c->cur[2]);
    int rc;
    if ((rc = launch_check(c, "k_case_init"))) return rc;
    k_set_scalar<<<1, 1, 0, c->stream>>>(c->scal + 2, t);
    CUDA_TRY(c, cudaMemsetAsync(c->flags, 0, sizeof(int), c->stream));
    return launch_check(c, "k_set_scalar");
}
SWE_API int swe_case_l2_error(swe_ctx *c, const swe_case *cs, double t, double out[3]) {
    if (!c || !cs || !out || cs->kind < 0 || cs->kind > SWE_CASE_BOWL_HUMP) return SWE_ERR_INVALID;
    CUDA_TRY(c, cudaSetDevice(c->device));
    k_case_error_partial<<<kDiagBlocks, kDiagThreads, 0, c->stream>>>(dev_mesh(c), dev_fields(c), *cs, t, c->diag);
    int rc;
    if ((rc = launch_check(c, "k_case_error_partial"))) return rc;
    k_case_error_final<<<1, 32, 0, c->stream>>>(c->diag, c->diag + 6 * kDiagBlocks);
    if ((rc = launch_check(c, "k_case_error_final"))) return rc;
    CUDA_TRY(c, cudaMemcpyAsync(out, c->diag + 6 * kDiagBlocks, sizeof(double) * 3, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return SWE_OK;
}

// ---- multi-GPU support ----
SWE_API int swe_set_cfl_edge_mask(swe_ctx *c, const uint8_t *mask) {
    if (!c) return SWE_ERR_INVALID;
    CUDA_TRY(c, cudaSetDevice(c->device));
    if (!mask) {  // back to "all edges count"
        if (c->dmin0) CUDA_TRY(c, cudaMemcpyAsync(c->dmin, c->dmin0, sizeof(double) * c->ne, cudaMemcpyDeviceToDevice, c->stream));
        return SWE_OK;
    }
    std::vector<unsigned char> h((size_t)c->ne);
    if (c->reordered) {
        std::vector<int> old((size_t)c->ne);
        CUDA_TRY(c, cudaMemcpy(old.data(), c->edge_old, sizeof(int) * c->ne, cudaMemcpyDeviceToHost));
        for (int d = 0; d < c->ne; ++d) h[d] = mask[old[d]];
    } else {
        std::copy(mask, mask + c->ne, h.begin());
    }
    if (!c->cfl_mask) CUDA_TRY(c, dalloc(&c->cfl_mask, (size_t)c->ne));
    CUDA_TRY(c, cudaMemcpy(c->cfl_mask, h.data(), (size_t)c->ne, cudaMemcpyHostToDevice));
    if (!c->dmin0) {
        CUDA_TRY(c, dalloc(&c->dmin0, (size_t)c->ne));
        CUDA_TRY(c, cudaMemcpy(c->dmin0, c->dmin, sizeof(double) * c->ne, cudaMemcpyDeviceToDevice));
    }
    k_apply_cfl_mask<<<nblk(c->ne, 256), 256, 0, c->stream>>>(c->ne, c->cfl_mask, c->dmin0, c->dmin);
    return launch_check(c, "k_apply_cfl_mask");
}

SWE_API int swe_halo_set_lists(swe_ctx *c, int64_t nsend, const int64_t *send_cells, int64_t nrecv, const int64_t *recv_cells) {
    if (!c || nsend < 0 || nrecv < 0 || (nsend && !send_cells) || (nrecv && !recv_cells)) return SWE_ERR_INVALID;
    CUDA_TRY(c, cudaSetDevice(c->device));
    auto upload = [&](int64_t n, const int64_t *src, int **dst) -> int {
        if (*dst) { cudaFree(*dst); *dst = nullptr; }
        if (n == 0) return SWE_OK;
        std::vector<int> h((size_t)n);
        for (int64_t k = 0; k < n; ++k) {
            if (src[k] < 0 || src[k] >= c->nt) { c->err = "swe_halo_set_lists: cell id out of range"; return SWE_ERR_INVALID; }
            h[k] = c->cell_new[src[k]];
        }
        CUDA_TRY(c, dalloc(dst, (size_t)n));
        CUDA_TRY(c, cudaMemcpy(*dst, h.data(), sizeof(int) * n, cudaMemcpyHostToDevice));
        return SWE_OK;
    };
    int rc;
    if ((rc = upload(nsend, send_cells, &c->send_cells))) return rc;
    if ((rc = upload(nrecv, recv_cells, &c->recv_cells))) return rc;
    c->nsend = (int)nsend; c->nrecv = (int)nrecv;
    return SWE_OK;
}
SWE_API int swe_halo_pack(swe_ctx *c, double *buf) {
    if (!c || (c->nsend && !buf)) return SWE_ERR_INVALID;
    if (c->nsend == 0) return SWE_OK;
    CUDA_TRY(c, cudaSetDevice(c->device));
    k_halo_pack<<<nblk(c->nsend, 256), 256, 0, c->stream>>>(c->nsend, c->send_cells, c->cur[0], c->cur[1], c->cur[2], buf);
    return launch_check(c, "k_halo_pack");
}
SWE_API int swe_halo_unpack(swe_ctx *c, const double *buf) {
    if (!c || (c->nrecv && !buf)) return SWE_ERR_INVALID;
    if (c->nrecv == 0) return SWE_OK;
    CUDA_TRY(c, cudaSetDevice(c->device));
    c->state_version++;
    k_halo_unpack<<<nblk(c->nrecv, 256), 256, 0, c->stream>>>(c->nrecv, c->recv_cells, buf, c->cur[0], c->cur[1], c->cur[2]);
    return launch_check(c, "k_halo_unpack");
}


// ---- peer-memory halo transport (CUDA IPC + NVLink stores) ----
// 1. swe_halo_p2p_alloc: allocates this rank's two receive buffers (3*nrecv doubles each, selected by the
//    parity of the exchange number) and npeers flag slots; returns their IPC handles (64 bytes each).
// 2. the host layer exchanges handles / segment offsets between ranks (any transport) and calls
//    swe_halo_p2p_connect once per peer.
// 3. per exchange: swe_halo_p2p_push (pack kernels storing into the peers' buffers + flags) and
//    swe_halo_p2p_pull (wait for the peers' flags, unpack).
SWE_API int swe_halo_p2p_alloc(swe_ctx *c, int32_t npeers, unsigned char *handles_3x64) {
    if (!c || npeers <= 0 || npeers > 32 || !handles_3x64 || c->nrecv <= 0) return SWE_ERR_INVALID;
    CUDA_TRY(c, cudaSetDevice(c->device));
    if (c->p2p_flags) { c->err = "swe_halo_p2p_alloc: already allocated"; return SWE_ERR_INVALID; }
    for (int q = 0; q < 2; ++q) {
        CUDA_TRY(c, cudaMalloc((void **)&c->p2p_recv[q], sizeof(double) * 3 * (size_t)c->nrecv));
        CUDA_TRY(c, cudaMemset(c->p2p_recv[q], 0, sizeof(double) * 3 * (size_t)c->nrecv));
    }
    CUDA_TRY(c, cudaMalloc((void **)&c->p2p_flags, sizeof(int) * 64));
    CUDA_TRY(c, cudaMemset(c->p2p_flags, 0, sizeof(int) * 64));
    cudaIpcMemHandle_t h;
    void *ptrs[3] = {c->p2p_recv[0], c->p2p_recv[1], c->p2p_flags};
    for (int q = 0; q < 3; ++q) {
        CUDA_TRY(c, cudaIpcGetMemHandle(&h, ptrs[q]));
        std::memcpy(handles_3x64 + 64 * q, &h, 64);
    }
    c->p2p_peers.clear();
    c->p2p_seq = 0;
    CUDA_TRY(c, cudaDeviceSynchronize());
    return SWE_OK;
}
// peer: send_start/send_count = this rank's segment of the send list for that peer; handles = the peer's three
// handles; dst_offset = where (in cells) this rank's segment starts in the peer's receive buffers;
// my_slot = index of this rank in the peer's flag array.
SWE_API int swe_halo_p2p_connect(swe_ctx *c, int64_t send_start, int64_t send_count, const unsigned char *peer_handles_3x64,
                                 int64_t dst_offset, int32_t my_slot) {
    if (!c || !peer_handles_3x64 || send_start < 0 || send_count < 0 || send_start + send_count > c->nsend || my_slot < 0 || my_slot >= 32)
        return SWE_ERR_INVALID;
    CUDA_TRY(c, cudaSetDevice(c->device));
    void *mapped[3];
    for (int q = 0; q < 3; ++q) {
        cudaIpcMemHandle_t h;
        std::memcpy(&h, peer_handles_3x64 + 64 * q, 64);
        CUDA_TRY(c, cudaIpcOpenMemHandle(&mapped[q], h, cudaIpcMemLazyEnablePeerAccess));
        c->p2p_imported.push_back(mapped[q]);
    }
    swe_ctx::P2PPeer p;
    p.send_start = (int)send_start; p.send_count = (int)send_count;
    p.peer_recv[0] = (double *)mapped[0] + 3 * dst_offset;
    p.peer_recv[1] = (double *)mapped[1] + 3 * dst_offset;
    p.peer_flag = (int *)mapped[2] + my_slot;
    c->p2p_peers.push_back(p);
    return SWE_OK;
}
SWE_API int swe_halo_p2p_push(swe_ctx *c) {
    if (!c || !c->p2p_flags) return SWE_ERR_INVALID;
    CUDA_TRY(c, cudaSetDevice(c->device));
    const int seq = ++c->p2p_seq;
    int rc;
    for (auto &p : c->p2p_peers) {
        if (p.send_count > 0) {
            k_halo_pack<<<nblk(p.send_count, 256), 256, 0, c->stream>>>(p.send_count, c->send_cells + p.send_start, c->cur[0],
                                                                       c->cur[1], c->cur[2], p.peer_recv[seq & 1]);
            if ((rc = launch_check(c, "k_halo_pack(peer)"))) return rc;
        }
        k_halo_signal<<<1, 1, 0, c->stream>>>(p.peer_flag, seq);
        if ((rc = launch_check(c, "k_halo_signal"))) return rc;
    }
    return SWE_OK;
}
SWE_API int swe_halo_p2p_pull(swe_ctx *c) {
    if (!c || !c->p2p_flags || c->p2p_seq <= 0) return SWE_ERR_INVALID;
    CUDA_TRY(c, cudaSetDevice(c->device));
    const int seq = c->p2p_seq;
    int rc;
    k_halo_wait<<<1, 32, 0, c->stream>>>(c->p2p_flags, (int)c->p2p_peers.size(), seq, 20000000000ll, c->flags + 5);
    if ((rc = launch_check(c, "k_halo_wait"))) return rc;
    if (c->nrecv > 0) {
        c->state_version++;
        k_halo_unpack<<<nblk(c->nrecv, 256), 256, 0, c->stream>>>(c->nrecv, c->recv_cells, c->p2p_recv[seq & 1], c->cur[0], c->cur[1], c->cur[2]);
        if ((rc = launch_check(c, "k_halo_unpack"))) return rc;
    }
    return SWE_OK;
}
// 1 if a peer-memory wait timed out since the context was created
SWE_API int swe_halo_p2p_error(swe_ctx *c) {
    if (!c) return SWE_ERR_INVALID;
    int flag = 0;
    cudaSetDevice(c->device);
    cudaMemcpy(&flag, c->flags + 5, sizeof(int), cudaMemcpyDeviceToHost);
    return flag;
}

}  // extern "C"

#include "swe_dist.cuh"
