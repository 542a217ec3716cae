// swe_dist.cuh — the multi-GPU time step behind the C-ABI (included at the end of swe_b200.cu).
//
// Upstream steps one SpaceDisc with Solvers::X(TimeDisc*, dt) (include/Solvers.h:6-8, src/Solvers.cpp);
// swe_dist_step is that call on N GPUs. One rank = one GPU = one swe_ctx over the rank's sub-mesh
// (distplan.cpp). Per RK stage, all on the rank's ONE stream (no side stream, nothing for a persistent
// kernel to starve):
//
//   K1 classes 0-1 (stencil free of halo cells)      <- overlaps the exchange still in flight
//   k_halo_wait_unpack                               <- peers' flags, then receive buffer -> halo cells
//   K1 classes 2-3, list passes, K2 (+ k_min_push after the last stage's K2)
//   K3 all cells, K4 classes 1-2 (the cells the peers need)
//   k_halo_pack_signal per peer                      <- NVLink stores into the peer's buffer + flag
//   K4 class 0 (interior)                            <- the stores fly meanwhile
//
// and once per step k_min_pull (global CFL minimum from the peer-memory table) before the dt update.
// A rank waits only if a neighbour is more than (interior update + interior reconstruction) behind.
// Peer buffers are reached through CUDA IPC (one process per GPU) or directly (one process, many
// GPUs); the launcher supplies one all-gather of a fixed-size blob during creation and nothing else.
#pragma once
#include <unistd.h>

#include "distplan.hpp"

struct DistHello {  // what every rank tells every other rank during creation
    int64_t pid;
    int32_t rank, device, npeers, pad;
    cudaIpcMemHandle_t h_recv0, h_recv1, h_ctrl;
    uint64_t raw_recv0, raw_recv1, raw_ctrl;
    int32_t peer_rank[kMaxRanks];
    int64_t recv_off[kMaxRanks], recv_cnt[kMaxRanks];
};

struct swe_dist {
    swe_dist_config cfg{};
    swe_dist_plan *plan = nullptr;
    swe_ctx *ctx = nullptr;
    std::string err;
    struct Peer { int rank; int send_start, send_count, recv_start, recv_count; double *peer_recv[2]; int *peer_flag; };
    std::vector<Peer> peers;
    int nsend = 0, nrecv = 0;
    // device memory of this rank
    double *recvbuf[2] = {nullptr, nullptr};
    unsigned char *ctrl = nullptr;   // [0,128) halo flags (int x 32) | [128,256) min flags | [512,768) min table (double x 2 x kMaxRanks)
    int *tickets = nullptr;          // [kMaxRanks] block tickets of the pack kernels
    long long *gid_dev = nullptr;
    unsigned char *owned_dev = nullptr;
    unsigned long long *hash_dev = nullptr;
    MinPeers minpeers{};
    std::vector<void *> imported;
    int seq = 0, minseq = 0;
    bool pending = false;            // an exchange was pushed and not yet pulled
    // the global CFL minimum of the last step was pushed but not yet pulled: it is first NEEDED by the stage update of
    // the next step, so the pull (a wait for the slowest rank) is deferred behind that step's first K1 + K2
    bool min_pending = false;
    int min_adaptive = 0;
    double min_dt_fixed = 0.;
    cudaStream_t own_stream = nullptr;  // group mode: every rank of the process gets its own non-blocking stream
    long long timeout_cycles = 0;
    DistHello hello{};
    int *halo_flags() const { return (int *)ctrl; }
    int *min_flags() const { return (int *)(ctrl + 128); }
    double *min_table() const { return (double *)(ctrl + 512); }
};

#define DIST_TRY(d, call)                                                                       \
    do {                                                                                        \
        cudaError_t e_ = (call);                                                                \
        if (e_ != cudaSuccess) { (d)->err = std::string(#call) + ": " + cudaGetErrorString(e_); return SWE_ERR_CUDA; } \
    } while (0)
#define DIST_CTX(d, call)                                                                       \
    do {                                                                                        \
        int rc_ = (call);                                                                       \
        if (rc_ != SWE_OK) { (d)->err = (d)->ctx ? (d)->ctx->err : std::string("context call failed"); return rc_; } \
    } while (0)

static thread_local std::string g_dist_error;

// phase A: context over the local sub-mesh, halo lists, receive buffers, the hello blob
static int dist_phase_a(swe_dist *d) {
    const swe_dist_plan &p = *d->plan;
    if (p.world > kMaxRanks) { d->err = "swe_dist: at most 16 ranks (one node)"; return SWE_ERR_INVALID; }
    swe_mesh mv;
    swe_hostmesh_view(p.mesh, &mv);
    mv.cor = d->cfg.cor; mv.tau = d->cfg.tau;
    int rc = swe_create_classes(&d->ctx, &mv, d->cfg.device, d->cfg.reorder, p.world > 1 ? p.cls.data() : nullptr);
    if (rc) { d->err = swe_last_error(nullptr); return rc; }
    swe_ctx *c = d->ctx;
    DIST_TRY(d, cudaSetDevice(c->device));
    if (p.world > 1) DIST_CTX(d, swe_set_cfl_edge_mask(c, p.cfl_mask.data()));
    // concatenated halo lists (peer order), translated to device numbering by swe_halo_set_lists
    std::vector<int64_t> send, recv;
    d->peers.clear();
    for (const auto &pe : p.peers) {
        swe_dist::Peer q{};
        q.rank = pe.rank;
        q.send_start = (int)send.size(); q.send_count = (int)pe.send.size();
        q.recv_start = (int)recv.size(); q.recv_count = (int)pe.recv.size();
        send.insert(send.end(), pe.send.begin(), pe.send.end());
        recv.insert(recv.end(), pe.recv.begin(), pe.recv.end());
        d->peers.push_back(q);
    }
    d->nsend = (int)send.size(); d->nrecv = (int)recv.size();
    DIST_CTX(d, swe_halo_set_lists(c, d->nsend, send.data(), d->nrecv, recv.data()));
    for (int q = 0; q < 2; ++q) {
        DIST_TRY(d, cudaMalloc((void **)&d->recvbuf[q], sizeof(double) * 3 * (size_t)std::max(d->nrecv, 1)));
        DIST_TRY(d, cudaMemset(d->recvbuf[q], 0, sizeof(double) * 3 * (size_t)std::max(d->nrecv, 1)));
    }
    DIST_TRY(d, cudaMalloc((void **)&d->ctrl, 1024));
    DIST_TRY(d, cudaMemset(d->ctrl, 0, 1024));
    DIST_TRY(d, dalloc(&d->tickets, (size_t)kMaxRanks));
    DIST_TRY(d, cudaMemset(d->tickets, 0, sizeof(int) * kMaxRanks));
    // global ids / ownership in LOCAL (caller) numbering for the state hash
    DIST_TRY(d, dalloc(&d->gid_dev, (size_t)c->nt));
    DIST_TRY(d, dalloc(&d->owned_dev, (size_t)c->nt));
    DIST_TRY(d, dalloc(&d->hash_dev, (size_t)1));
    DIST_TRY(d, cudaMemcpy(d->gid_dev, p.gcell.data(), sizeof(long long) * c->nt, cudaMemcpyHostToDevice));
    DIST_TRY(d, cudaMemcpy(d->owned_dev, p.owned.data(), (size_t)c->nt, cudaMemcpyHostToDevice));
    int khz = 1965000;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, c->device);
    const double tmo = d->cfg.wait_timeout_s > 0 ? d->cfg.wait_timeout_s : 60.0;
    d->timeout_cycles = (long long)(tmo * 1e3 * (double)khz);
    DistHello &h = d->hello;
    std::memset(&h, 0, sizeof(h));
    h.pid = (int64_t)getpid(); h.rank = p.rank; h.device = c->device; h.npeers = (int32_t)d->peers.size();
    if (p.world > 1) {
        DIST_TRY(d, cudaIpcGetMemHandle(&h.h_recv0, d->recvbuf[0]));
        DIST_TRY(d, cudaIpcGetMemHandle(&h.h_recv1, d->recvbuf[1]));
        DIST_TRY(d, cudaIpcGetMemHandle(&h.h_ctrl, d->ctrl));
    }
    h.raw_recv0 = (uint64_t)d->recvbuf[0]; h.raw_recv1 = (uint64_t)d->recvbuf[1]; h.raw_ctrl = (uint64_t)d->ctrl;
    for (size_t k = 0; k < d->peers.size(); ++k) {
        h.peer_rank[k] = d->peers[k].rank; h.recv_off[k] = d->peers[k].recv_start; h.recv_cnt[k] = d->peers[k].recv_count;
    }
    DIST_TRY(d, cudaDeviceSynchronize());
    return SWE_OK;
}

// phase B: map every rank's control block (global min) and the neighbours' receive buffers
static int dist_phase_b(swe_dist *d, const DistHello *all) {
    const swe_dist_plan &p = *d->plan;
    swe_ctx *c = d->ctx;
    DIST_TRY(d, cudaSetDevice(c->device));
    auto map = [&](const DistHello &h, int which, void **out) -> int {  // 0 recv0, 1 recv1, 2 ctrl
        const uint64_t raw = which == 0 ? h.raw_recv0 : which == 1 ? h.raw_recv1 : h.raw_ctrl;
        if (h.pid == (int64_t)getpid()) {  // same process: direct peer access
            if (h.device != c->device) {
                int can = 0;
                DIST_TRY(d, cudaDeviceCanAccessPeer(&can, c->device, h.device));
                if (!can) { d->err = "swe_dist: no peer access between the GPUs of this process"; return SWE_ERR_CUDA; }
                cudaError_t e = cudaDeviceEnablePeerAccess(h.device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { d->err = cudaGetErrorString(e); return SWE_ERR_CUDA; }
                cudaGetLastError();
            }
            *out = (void *)raw;
            return SWE_OK;
        }
        const cudaIpcMemHandle_t &hd = which == 0 ? h.h_recv0 : which == 1 ? h.h_recv1 : h.h_ctrl;
        DIST_TRY(d, cudaIpcOpenMemHandle(out, hd, cudaIpcMemLazyEnablePeerAccess));
        d->imported.push_back(*out);
        return SWE_OK;
    };
    int rc;
    for (int r = 0; r < p.world; ++r) {
        if (all[r].rank != r) { d->err = "swe_dist: the bootstrap all-gather did not return the blobs in rank order"; return SWE_ERR_INVALID; }
        void *ctrl = d->ctrl;
        if (r != p.rank && (rc = map(all[r], 2, &ctrl))) return rc;
        d->minpeers.buf[r] = (double *)((unsigned char *)ctrl + 512);
        d->minpeers.flag[r] = (int *)((unsigned char *)ctrl + 128);
    }
    for (auto &q : d->peers) {
        const DistHello &h = all[q.rank];
        int slot = -1;
        for (int k = 0; k < h.npeers; ++k) if (h.peer_rank[k] == p.rank) slot = k;
        if (slot < 0 || h.recv_cnt[slot] != q.send_count) { d->err = "swe_dist: halo lists of neighbouring ranks disagree"; return SWE_ERR_INVALID; }
        void *r0 = nullptr, *r1 = nullptr;
        if ((rc = map(h, 0, &r0)) || (rc = map(h, 1, &r1))) return rc;
        q.peer_recv[0] = (double *)r0 + 3 * h.recv_off[slot];
        q.peer_recv[1] = (double *)r1 + 3 * h.recv_off[slot];
        // control block of that rank was mapped above: flag slot = my position in ITS peer list
        q.peer_flag = (int *)((unsigned char *)d->minpeers.flag[q.rank] - 128) + slot;
    }
    DIST_TRY(d, cudaDeviceSynchronize());
    return SWE_OK;
}

static void dist_free(swe_dist *d) {
    if (!d) return;
    if (d->ctx) { cudaSetDevice(d->ctx->device); cudaStreamSynchronize(d->ctx->stream); }
    if (d->own_stream) cudaStreamDestroy(d->own_stream);
    for (void *p : d->imported) cudaIpcCloseMemHandle(p);
    void *ptrs[] = {d->recvbuf[0], d->recvbuf[1], d->ctrl, d->tickets, d->gid_dev, d->owned_dev, d->hash_dev};
    for (void *p : ptrs) if (p) cudaFree(p);
    if (d->ctx) destroy_ctx(d->ctx);
    swe_dist_plan_free(d->plan);
    delete d;
}

// ---- the exchange, in stream order -----------------------------------------------------------------
static int dist_push(swe_dist *d, double **buf) {  // pack-and-signal the state in buf[0..2] to every peer
    swe_ctx *c = d->ctx;
    if (d->peers.empty()) return SWE_OK;
    const int seq = ++d->seq;
    int rc, k = 0;
    const int kt = kt_begin(c, KT_HALO_PACK);
    for (auto &q : d->peers) {
        const int g = std::max(1, nblk(q.send_count, 256));
        k_halo_pack_signal<<<g, 256, 0, c->stream>>>(q.send_count, c->send_cells + q.send_start, buf[0], buf[1], buf[2],
                                                     q.peer_recv[seq & 1], q.peer_flag, seq, d->tickets + k);
        if ((rc = launch_check(c, "k_halo_pack_signal"))) { d->err = c->err; return rc; }
        ++k;
    }
    kt_end(c, kt);
    d->pending = true;
    return SWE_OK;
}
static int dist_pull(swe_dist *d) {  // wait for every peer's flag, unpack into the halo cells
    swe_ctx *c = d->ctx;
    if (!d->pending) return SWE_OK;
    d->pending = false;
    const int g = std::max(1, std::min(nblk(d->nrecv, 256), 4 * c->sms));
    const int kt = kt_begin(c, KT_HALO_WAIT);  // includes the time spent waiting for a slower neighbour
    c->state_version++;  // halo cells change
    k_halo_wait_unpack<<<g, 256, 0, c->stream>>>(d->halo_flags(), (int)d->peers.size(), d->seq, d->timeout_cycles, c->flags + 5,
                                                 d->nrecv, c->recv_cells, d->recvbuf[d->seq & 1], c->cur[0], c->cur[1], c->cur[2]);
    kt_end(c, kt);
    int rc = launch_check(c, "k_halo_wait_unpack");
    if (rc) d->err = c->err;
    return rc;
}

// pull the global minimum of the previous step and advance time / dt with it
static int dist_finish_min(swe_dist *d, bool also_scal0) {
    swe_ctx *c = d->ctx;
    if (!d->min_pending) return SWE_OK;
    d->min_pending = false;
    int rc;
    if (d->plan->world > 1) {
        const int kt = kt_begin(c, KT_MIN);
        k_min_pull<<<1, 32, 0, c->stream>>>(c->scal, d->min_table(), d->min_flags(), d->plan->world, d->minseq, d->timeout_cycles,
                                            c->flags + 5, also_scal0 ? 1 : 0);
        kt_end(c, kt);
        if ((rc = launch_check(c, "k_min_pull"))) { d->err = c->err; return rc; }
    }
    k_post_step<<<1, 1, 0, c->stream>>>(dev_fields(c), d->min_dt_fixed, d->min_adaptive, d->plan->world > 1 ? 4 : 0);
    if ((rc = launch_check(c, "k_post_step"))) { d->err = c->err; return rc; }
    return SWE_OK;
}

static int dist_interface_values(swe_dist *d) {
    swe_ctx *c = d->ctx;
    if (!d->pending) DIST_CTX(d, interface_values_range(c, 0, c->nt, true, true));
    else {
        DIST_CTX(d, interface_values_range(c, c->class_first[0], c->class_first[2], true, false));
        int rc = dist_pull(d);
        if (rc) return rc;
        DIST_CTX(d, interface_values_range(c, c->class_first[2], c->class_first[4], false, true));
    }
    return SWE_OK;
}

static int dist_stage(swe_dist *d, swe_flux flux, swe_wavespeed ws, double a0, double a1, double dt_host, double dt_coef,
                      bool save, bool last) {
    swe_ctx *c = d->ctx;
    const int world = d->plan->world;
    int rc;
    if ((rc = dist_interface_values(d))) return rc;
    DIST_CTX(d, compute_fluxes(c, flux, ws, last));  // only the last stage's CFL minimum is ever read
    // dt of THIS step from the previous step's global minimum: first needed by the stage update below, so the wait
    // for the slowest rank hides behind the reconstruction + flux kernels above (must precede this step's push)
    if ((rc = dist_finish_min(d, false))) return rc;
    if (last && world > 1) {  // this rank's CFL minimum to every rank's table (needed only by the next step's dt)
        const int seq = ++d->minseq;
        const int kt = kt_begin(c, KT_MIN);
        k_min_push<<<1, 32, 0, c->stream>>>(c->scal, d->minpeers, world, d->plan->rank, seq);
        kt_end(c, kt);
        if ((rc = launch_check(c, "k_min_push"))) { d->err = c->err; return rc; }
    }
    if (save) swe_save_state(c);
    double **outb = nullptr;
    DIST_CTX(d, stage_drain(c, &outb));
    if (world > 1 && d->cfg.overlap) {
        // the cells the peers need first, then their NVLink stores fly while the interior is updated
        DIST_CTX(d, stage_update_range(c, outb, a0, a1, dt_host, dt_coef, c->class_first[1], c->class_first[3]));
        if ((rc = dist_push(d, outb))) return rc;
        DIST_CTX(d, stage_update_range(c, outb, a0, a1, dt_host, dt_coef, c->class_first[0], c->class_first[1]));
        c->cur = outb;  // only now: the stage input (c->cur) had to stay in place for the second range
        c->state_version++;
        // class 3 (halo cells) is not updated: every halo cell is overwritten by the exchange
    } else {
        DIST_CTX(d, stage_update_range(c, outb, a0, a1, dt_host, dt_coef, 0, c->class_first[3]));
        c->cur = outb;
        c->state_version++;
        if (world > 1) {
            if ((rc = dist_push(d, c->cur))) return rc;
            if ((rc = dist_pull(d))) return rc;
        }
    }
    return SWE_OK;
}

static int dist_one_step(swe_dist *d, swe_scheme scheme, swe_flux flux, swe_wavespeed ws, double dt, bool dev_dt) {
    struct St { double a0, a1, coef; };
    static const St tab[3][3] = {{{0., 1., 1.}, {0, 0, 0}, {0, 0, 0}},
                                 {{0., 1., 1.}, {0.5, 0.5, 0.5}, {0, 0, 0}},
                                 {{0., 1., 1.}, {0.75, 0.25, 0.25}, {1. / 3., 2. / 3., 2. / 3.}}};
    const int ns = scheme == SWE_EULER ? 1 : scheme == SWE_SSPRK2 ? 2 : 3;
    int rc;
    for (int k = 0; k < ns; ++k) {
        const St &s = tab[scheme][k];
        if ((rc = dist_stage(d, flux, ws, s.a0, s.a1, dev_dt ? 0. : s.coef * dt, dev_dt ? s.coef : 0., k == 0 && ns > 1, k == ns - 1)))
            return rc;
    }
    return SWE_OK;
}

extern "C" {

SWE_API int swe_dist_create(swe_dist **out, swe_dist_plan *plan, const swe_dist_config *cfg) {
    if (!out || !plan || !cfg) return SWE_ERR_INVALID;
    *out = nullptr;
    if (plan->world > 1 && !cfg->allgather) { g_dist_error = "swe_dist_create: world > 1 needs the bootstrap all-gather"; return SWE_ERR_INVALID; }
    swe_dist *d = new swe_dist();
    d->cfg = *cfg; d->plan = plan;
    int rc = dist_phase_a(d);
    if (!rc && plan->world > 1) {
        std::vector<DistHello> all((size_t)plan->world);
        if (cfg->allgather(cfg->user, &d->hello, all.data(), (int64_t)sizeof(DistHello)) != 0) { d->err = "swe_dist_create: bootstrap all-gather failed"; rc = SWE_ERR_INVALID; }
        if (!rc) rc = dist_phase_b(d, all.data());
        if (!rc) {  // nobody may start pushing before everybody has mapped: one more (empty) round
            char tok = 0;
            std::vector<char> toks((size_t)plan->world);
            if (cfg->allgather(cfg->user, &tok, toks.data(), 1) != 0) { d->err = "swe_dist_create: bootstrap barrier failed"; rc = SWE_ERR_INVALID; }
        }
    }
    if (rc) { g_dist_error = d->err; d->plan = nullptr; dist_free(d); return rc; }
    *out = d;
    return SWE_OK;
}

SWE_API int swe_dist_group_create(swe_dist **out, swe_dist_plan **plans, const int32_t *devices, int32_t world,
                                  const swe_dist_config *cfg) {
    if (!out || !plans || !devices || !cfg || world < 1 || world > kMaxRanks) return SWE_ERR_INVALID;
    std::vector<swe_dist *> ds((size_t)world, nullptr);
    std::vector<DistHello> all((size_t)world);
    int rc = SWE_OK;
    for (int r = 0; r < world && !rc; ++r) {
        ds[r] = new swe_dist();
        ds[r]->cfg = *cfg; ds[r]->cfg.device = devices[r]; ds[r]->plan = plans[r];
        rc = dist_phase_a(ds[r]);
        if (!rc) {  // the ranks of one process must not serialise on the legacy default stream: a wait kernel of one
                    // rank spins until the pack kernel of another rank has run
            if (cudaStreamCreateWithFlags(&ds[r]->own_stream, cudaStreamNonBlocking) != cudaSuccess) { ds[r]->err = "cudaStreamCreate failed"; rc = SWE_ERR_CUDA; }
            else ds[r]->ctx->stream = ds[r]->own_stream;
        }
        if (rc) g_dist_error = ds[r]->err; else all[r] = ds[r]->hello;
    }
    for (int r = 0; r < world && !rc && world > 1; ++r) {
        rc = dist_phase_b(ds[r], all.data());
        if (rc) g_dist_error = ds[r]->err;
    }
    if (rc) {
        for (auto *d : ds) if (d) { d->plan = nullptr; dist_free(d); }
        return rc;
    }
    for (int r = 0; r < world; ++r) out[r] = ds[r];
    return SWE_OK;
}

SWE_API void swe_dist_destroy(swe_dist *d) { dist_free(d); }
SWE_API const char *swe_dist_last_error(const swe_dist *d) { return d ? d->err.c_str() : g_dist_error.c_str(); }
SWE_API swe_ctx *swe_dist_ctx(swe_dist *d) { return d ? d->ctx : nullptr; }
SWE_API const swe_dist_plan *swe_dist_get_plan(const swe_dist *d) { return d ? d->plan : nullptr; }

SWE_API int swe_dist_exchange(swe_dist *d) {
    if (!d) return SWE_ERR_INVALID;
    DIST_TRY(d, cudaSetDevice(d->ctx->device));
    int rc;
    if (d->pending && (rc = dist_pull(d))) return rc;
    if ((rc = dist_push(d, d->ctx->cur))) return rc;
    return dist_pull(d);
}

SWE_API int swe_dist_step(swe_dist *d, swe_scheme scheme, swe_flux flux, swe_wavespeed ws, double dt) {
    if (!d) return SWE_ERR_INVALID;
    if (scheme < SWE_EULER || scheme > SWE_SSPRK3) { d->err = "swe_dist_step: unknown scheme"; return SWE_ERR_INVALID; }
    if (!(dt > 0.)) { d->err = "swe_dist_step: dt must be positive"; return SWE_ERR_INVALID; }
    DIST_TRY(d, cudaSetDevice(d->ctx->device));
    DIST_CTX(d, dry_refresh(d->ctx));
    int rc = dist_one_step(d, scheme, flux, ws, dt, false);
    if (rc) return rc;
    d->min_pending = true; d->min_adaptive = 0; d->min_dt_fixed = dt;
    if (d->plan->world == 1) return dist_finish_min(d, false);  // nothing to wait for on one GPU
    return SWE_OK;
}

static int dist_run_begin(swe_dist *d, swe_scheme scheme, double dt, double dt0) {
    if (scheme < SWE_EULER || scheme > SWE_SSPRK3) { d->err = "swe_dist_run: unknown scheme"; return SWE_ERR_INVALID; }
    if (!(dt > 0.) && !(dt0 > 0.)) { d->err = "swe_dist_run: adaptive mode needs dt0 > 0"; return SWE_ERR_INVALID; }
    DIST_TRY(d, cudaSetDevice(d->ctx->device));
    int rc = dist_finish_min(d, true);
    if (rc) return rc;
    // dry-region instantiations or not: decided here, where everything issued so far has its matching peer work
    // issued too (the evaluation synchronises this rank's stream once after the state was set from outside)
    DIST_CTX(d, dry_refresh(d->ctx));
    if (!(dt > 0.)) DIST_CTX(d, swe_set_dt(d->ctx, dt0));
    return SWE_OK;
}
static int dist_run_one(swe_dist *d, swe_scheme scheme, swe_flux flux, swe_wavespeed ws, double dt) {
    const bool adaptive = !(dt > 0.);
    DIST_TRY(d, cudaSetDevice(d->ctx->device));
    int rc = dist_one_step(d, scheme, flux, ws, dt, adaptive);
    if (rc) return rc;
    d->min_pending = true; d->min_adaptive = adaptive ? 1 : 0; d->min_dt_fixed = dt;
    if (d->plan->world == 1) return dist_finish_min(d, false);
    return SWE_OK;
}

SWE_API int swe_dist_run(swe_dist *d, swe_scheme scheme, swe_flux flux, swe_wavespeed ws, int64_t nsteps, double dt, double dt0) {
    if (!d) return SWE_ERR_INVALID;
    int rc = dist_run_begin(d, scheme, dt, dt0);
    for (int64_t s = 0; s < nsteps && !rc; ++s) rc = dist_run_one(d, scheme, flux, ws, dt);
    return rc;
}

SWE_API int swe_dist_group_run(swe_dist **ranks, int32_t world, swe_scheme scheme, swe_flux flux, swe_wavespeed ws, int64_t nsteps,
                               double dt, double dt0) {
    if (!ranks || world < 1) return SWE_ERR_INVALID;
    int rc = SWE_OK;
    for (int r = 0; r < world && !rc; ++r) rc = dist_run_begin(ranks[r], scheme, dt, dt0);
    // step by step over the ranks: every launch is asynchronous, the device-side waits pair up across GPUs
    for (int64_t s = 0; s < nsteps && !rc; ++s)
        for (int r = 0; r < world && !rc; ++r) rc = dist_run_one(ranks[r], scheme, flux, ws, dt);
    return rc;
}

// host-buffer pipeline on a rank (see swe_submit_step_host): local state (owned + halo cells) in, one step, local state out
SWE_API int swe_dist_submit_step_host(swe_dist *d, const double *host_in, double *host_out, swe_scheme scheme, swe_flux flux,
                                      swe_wavespeed ws, double dt) {
    if (!d || !host_in || !host_out) return SWE_ERR_INVALID;
    if (scheme < SWE_EULER || scheme > SWE_SSPRK3 || !(dt > 0.)) { d->err = "swe_dist_submit_step_host: bad scheme / dt"; return SWE_ERR_INVALID; }
    { int rc0 = dist_finish_min(d, false); if (rc0) return rc0; }
    DIST_CTX(d, pipe_begin(d->ctx, host_in));
    d->pending = false;  // the uploaded local state carries its own halo cells
    int rc = dist_one_step(d, scheme, flux, ws, dt, false);
    if (rc) return rc;
    d->min_pending = true; d->min_adaptive = 0; d->min_dt_fixed = dt;
    if (d->plan->world == 1 && (rc = dist_finish_min(d, false))) return rc;
    if ((rc = dist_pull(d))) return rc;  // halo cells of the result before it is downloaded
    DIST_CTX(d, pipe_end(d->ctx, host_out));
    return SWE_OK;
}
SWE_API int swe_dist_wait_host(swe_dist *d) {
    if (!d) return SWE_ERR_INVALID;
    swe_ctx *c = d->ctx;
    DIST_TRY(d, cudaSetDevice(c->device));
    if (c->pipe.ok) { DIST_TRY(d, cudaStreamSynchronize(c->pipe.s_in)); DIST_TRY(d, cudaStreamSynchronize(c->pipe.s_out)); }
    return swe_dist_synchronize(d);
}

SWE_API int swe_dist_synchronize(swe_dist *d) {
    if (!d) return SWE_ERR_INVALID;
    swe_ctx *c = d->ctx;
    DIST_TRY(d, cudaSetDevice(c->device));
    int rc;
    if (d->pending && (rc = dist_pull(d))) return rc;  // order the stream after the exchange in flight
    if ((rc = dist_finish_min(d, true))) return rc;    // ... and after the deferred global minimum
    DIST_TRY(d, cudaStreamSynchronize(c->stream));
    int flags[8];
    DIST_TRY(d, cudaMemcpy(flags, c->flags, sizeof(flags), cudaMemcpyDeviceToHost));
    if (flags[5]) { d->err = "swe_dist: a peer-memory wait timed out (a neighbouring rank stopped?)"; return SWE_ERR_CUDA; }
    if (flags[0]) { d->err = c->err = "non-finite cell state detected on device (SolverError)"; return SWE_ERR_NUMERIC; }
    return SWE_OK;
}

SWE_API int swe_dist_cfl_dt(swe_dist *d, double *dt) {
    if (!d || !dt) return SWE_ERR_INVALID;
    DIST_TRY(d, cudaSetDevice(d->ctx->device));
    int rc = dist_finish_min(d, true);
    if (rc) return rc;
    double v = 0.;  // scal[4] = global minimum on several GPUs, scal[0] on one
    DIST_TRY(d, cudaMemcpyAsync(&v, d->ctx->scal + (d->plan->world > 1 ? 4 : 0), sizeof(double), cudaMemcpyDeviceToHost, d->ctx->stream));
    DIST_TRY(d, cudaStreamSynchronize(d->ctx->stream));
    *dt = SWE_CFL * v;
    return SWE_OK;
}

static int state_hash(swe_ctx *c, const long long *gid, const unsigned char *mask, unsigned long long *dev, uint64_t *out) {
    CUDA_TRY(c, cudaSetDevice(c->device));
    CUDA_TRY(c, cudaMemsetAsync(dev, 0, sizeof(unsigned long long), c->stream));
    k_state_hash<<<std::min(nblk(c->nt, 256), 8 * c->sms), 256, 0, c->stream>>>(c->nt, c->cell_old, gid, mask, c->cur[0], c->cur[1], c->cur[2], dev);
    int rc = launch_check(c, "k_state_hash");
    if (rc) return rc;
    unsigned long long h = 0;
    CUDA_TRY(c, cudaMemcpyAsync(&h, dev, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    *out = (uint64_t)h;
    return SWE_OK;
}
SWE_API int swe_dist_state_hash(swe_dist *d, uint64_t *partial) {
    if (!d || !partial) return SWE_ERR_INVALID;
    int rc;
    if (d->pending && (rc = dist_pull(d))) return rc;
    DIST_CTX(d, state_hash(d->ctx, d->gid_dev, d->owned_dev, d->hash_dev, partial));
    return SWE_OK;
}
SWE_API int swe_state_hash(swe_ctx *c, uint64_t *hash) {
    if (!c || !hash) return SWE_ERR_INVALID;
    CUDA_TRY(c, cudaSetDevice(c->device));
    unsigned long long *dev = nullptr;
    CUDA_TRY(c, cudaMalloc((void **)&dev, sizeof(unsigned long long)));
    const int rc = state_hash(c, nullptr, nullptr, dev, hash);
    cudaFree(dev);
    return rc;
}

SWE_API int swe_dist_get_owned_state(swe_dist *d, double *prim_global) {
    if (!d || !prim_global) return SWE_ERR_INVALID;
    swe_ctx *c = d->ctx;
    int rc;
    if (d->pending && (rc = dist_pull(d))) return rc;
    std::vector<double> loc((size_t)3 * c->nt);
    DIST_CTX(d, swe_get_state(c, loc.data()));
    const swe_dist_plan &p = *d->plan;
    for (int64_t l = 0; l < c->nt; ++l)
        if (p.owned[(size_t)l]) std::copy(&loc[3 * l], &loc[3 * l] + 3, &prim_global[3 * p.gcell[(size_t)l]]);
    return SWE_OK;
}
SWE_API int swe_dist_set_state_global(swe_dist *d, const double *prim_global) {
    if (!d || !prim_global) return SWE_ERR_INVALID;
    swe_ctx *c = d->ctx;
    const swe_dist_plan &p = *d->plan;
    std::vector<double> loc((size_t)3 * c->nt);
    for (int64_t l = 0; l < c->nt; ++l) std::copy(&prim_global[3 * p.gcell[(size_t)l]], &prim_global[3 * p.gcell[(size_t)l]] + 3, &loc[3 * l]);
    DIST_CTX(d, swe_set_state(c, loc.data()));
    d->pending = false;
    return SWE_OK;
}

}  // extern "C"
