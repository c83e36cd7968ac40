// C-ABI implementation (include/sph_b200.h): persistent device state, the step pipeline and the
// host<->device boundary. One translation unit: the kernels live in the .cuh files next to it.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo --fmad=false -shared ...
// (--fmad=false is part of the parity contract, see sph_device.cuh).
#include "../../include/sph_b200.h"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <utility>
#include <vector>

#include "sph_device.cuh"
#include "sph_grid.cuh"
#include "sph_physics.cuh"
#include "sph_io.cuh"
#include "sph_slab.cuh"
#include "sph_scene.cuh"

using namespace sphb;

namespace {

thread_local char g_create_error[512] = "";

constexpr float REF_PI = 3.14159265f;  // src/SPHSystem.cpp:6

struct PassEvents {
    cudaEvent_t e[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
};

}  // namespace

struct sph_handle {
    int device = 0;
    int num_sms = 148;
    cudaStream_t stream = nullptr;
    uint64_t cap = 0, n = 0, steps = 0;
    sph_settings settings{};
    sph_derived derived{};
    Params P{};

    float4 *pos[2] = {nullptr, nullptr}, *vel[2] = {nullptr, nullptr};
    int cur = 0;
    float4 *force = nullptr;
    uint32_t *hash16 = nullptr;  // start-of-step hash16 of every sorted row (density lives in vel.w)
    uint32_t *nlist = nullptr, *ncount = nullptr;  // neighbour lists written by the density pass
    uint2 *cell_rank = nullptr, *slot = nullptr;
    uint32_t *order = nullptr, *map = nullptr;
    uint32_t *tile_claim = nullptr;  // per tile of CLUMP_ROWS rows: the build epoch in which a warp claimed it; behind it, the claimed tiles
    uint32_t *inverse = nullptr;          // sorted row of each pre-sort row (slab mode)
    uint32_t *halo_rows[2] = {nullptr, nullptr};  // pre-sort rows packed into each halo message
    uint64_t halo_n[2] = {0, 0};
    uint64_t ghost_first[2] = {0, 0}, ghost_n[2] = {0, 0};  // appended ghost batches (pre-sort rows)
    uint64_t n_ghost = 0;                 // ghost rows among the n rows
    bool slab_mode = false;
    bool slab_fast = false;       // sync-free slab path: exact row count is read back one step late
    bool rows_pending = false;    // pinned_rows holds a newer row count than h->n
    uint32_t *pinned_rows = nullptr;  // [0] rows surviving the last build, [1] slab violation bits
    cudaEvent_t ev_rows = nullptr;
    uint64_t fast_halo_cap = 0;
    bool have_bbox_from_integration = false;
    int bbox_expand = 0;          // cells added around the box by the next grid plan
    // peer-memory path
    char *mailbox = nullptr;           // local mailbox (IPC-exported)
    char *peer_mailbox[2] = {nullptr, nullptr};  // mapped mailboxes of the left / right neighbour
    P2PLayout p2p{};
    uint64_t p2p_H = 0, p2p_M = 0;
    uint32_t p2p_epoch = 0;
    uint64_t p2p_M_step = 0;            // migrant message capacity of the current step (<= p2p_M)
    uint32_t *mig_rows[2] = {nullptr, nullptr};  // per migrant slot: the local row kept as a ghost, or P2P_NO_ROW
    uint64_t arr_first[2] = {0, 0};    // pre-sort row of the first arrival of each side
    int p2p_lo = 0, p2p_hi = 0;    // this rank's cuts in the current peer step
    int slab_lo = 0, slab_hi = 0;  // this rank's x-cell range in the current sync-free step (0, 0: unknown -> no interior range)
    cudaStream_t stream2 = nullptr;  // peer steps: the density exchange runs here, next to the interior rows' force pass
    cudaEvent_t ev_dens = nullptr, ev_rho = nullptr;
    bool p2p_overlap = false;      // SPH_B200_P2P_OVERLAP=1: density exchange on the second stream, force pass split (measured: no gain)
    bool rho_pending = false;      // ev_rho has been recorded for this step's force pass to wait on
    bool p2p_clean = false;   // the peer step's device cursors / done-counters are zero (it re-zeroes them itself)
    int forces_cfg = 0, density_cfg = 0;
    int scan_blocks = 8;  // blocks per SM of the scan's persistent grid (SPH_B200_SCAN_BLOCKS: A/B)
    int grid_rows[3] = {4, 2, 1};  // rows per thread of k_cell_hist, k_place, k_order_gather
    int heavy_blocks = HEAVY_BLOCKS_PER_SM;  // blocks per SM of the heavy kernels' persistent grids (SPH_B200_HEAVY_BLOCKS: A/B)
    // Sync-free slab steps scan only the edge x-layers (sph_slab.cuh, "edge scans"): valid while the rows
    // are in the cell order of the last build, i.e. from a slab force step until anything else touches them.
    bool edge_scan_enabled = true;   // SPH_B200_EDGE_SCAN=0 scans every row (see DESIGN.md §5 for the measurements)
    bool edge_ok = false;        // rows [0, n) are in the cell order h->cells describes
    bool edge_all = true;        // this step's scans look at every row
    uint64_t edge_sorted = 0;    // rows in cell order when this step began (arrivals are appended behind them)
    bool tile_armed = false;  // k_forces_tile's dynamic shared memory limit has been raised
    bool tma_armed = false;   // ... and k_density_tma's
    unsigned long long *slab_counts = nullptr;  // SLAB_MAX_RANKS counters + cursors
    uint32_t *cells = nullptr;
    uint32_t max_cells = 0;
    uint32_t *h16_cells = nullptr;  // 65536 + 2 counters for the hash16 ordering
    uint32_t *const_65536 = nullptr;
    unsigned long long *tile_state = nullptr;
    GridDesc *gd = nullptr;
    StepCounters *ctr = nullptr;
    StatsAccum *stats_acc = nullptr;
    int parity = 0;
    bool have_state = false;  // particles uploaded
    bool have_step = false;   // density / hash16 rows are valid and aligned with pos / vel
    bool have_force = false;  // the force column holds the last step's forces
    bool write_force = false; // the next step writes the force column (drop-in call)

    void *scratch = nullptr;
    size_t scratch_bytes = 0;

    // Reset point: a device copy of the rows at the time of sph_set_reset_point (SPHSystem::reset, the GUI's R key)
    float4 *reset_pos = nullptr, *reset_vel = nullptr;
    uint64_t reset_n = 0;
    bool have_reset = false;

    // Per-pass CUDA-event timing (the Timer blocks of src/sph.cpp:235,249,262): one event set per
    // timed step, drawn from a pool that grows on demand and is read back in sph_pass_times.
    bool timing = false;
    std::vector<PassEvents> ev_pool;
    size_t ev_used = 0;
    static constexpr size_t kMaxTimedSteps = 16384;
    uint64_t launches = 0;  // kernels launched by sph_step / sph_update_particles_aos so far

    // Captured steps. A resident step takes no decision on the host (the grid plan, the row counts and
    // the scan epoch all live on the device), so sph_step replays CUDA graphs of 1 and kGraphLong
    // steps, one per bounding-box parity at entry, captured on first use for the current key.
    static constexpr int kGraphLong = 16;
    struct GraphKey {
        uint64_t n;
        float dt;
        int cur, write_force, bbox_expand, forces_cfg, density_cfg;
        uint32_t max_cells;
        Params P;
    };
    GraphKey graph_key;
    cudaGraphExec_t graphs[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};  // [1 | kGraphLong steps][parity]
    uint64_t graph_launches[2] = {0, 0};  // kernel nodes of a captured 1-step / kGraphLong-step graph (counted at capture)
    bool graph_enabled = true;  // SPH_B200_GRAPH=0 launches every kernel from the host
    bool pdl = false;           // SPH_B200_PDL=1: step kernels launched with programmatic dependent launch (measured: no gain)

    char err[512] = "";
};

namespace {

int fail(sph_handle *h, int code, const char *fmt, ...)
{
    char *dst = h ? h->err : g_create_error;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(dst, 512, fmt, ap);
    va_end(ap);
    return code;
}

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(h, SPH_ERR_CUDA, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__,      \
                        cudaGetErrorString(e_));                                                   \
    } while (0)
#define CK_LAUNCH() CK(cudaGetLastError())
// after a kernel launch of the step pipeline: checked and counted (sph_launch_count)
#define CK_STEP_LAUNCH() do { CK(cudaGetLastError()); ++h->launches; } while (0)

inline unsigned blocks_for(uint64_t n, int threads) { return (unsigned)((n + threads - 1) / threads); }

// Blocks of a persistent grid: per_sm blocks on every SM, but no more than the work can use (a 3 375-particle cube
// does not need 1 184 blocks that look at a counter and leave: every one of them is launch latency of a 30 us step).
inline unsigned persistent_blocks(const sph_handle *h, int per_sm, uint64_t useful)
{
    return (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)h->num_sms * per_sm, useful));
}

// Launch of a step kernel with programmatic stream serialization (pdl_enter in sph_device.cuh): the kernel's
// blocks may become resident while the kernel in front of it drains and wait on the device for its completion,
// which would hide launch latency between the nine kernels of a step. A captured step keeps these as
// programmatic edges of the graph. Opt-in (SPH_B200_PDL=1): measured on B200 (tools/ab_env.py, bit-identical
// either way) the captured step gains nothing — 0.2641 vs 0.2650 ms at 1 M, 0.9685 vs 0.9719 at 8 M — and the
// 3 375-particle cube loses 2 us per step (29.2 vs 27.4): kernel nodes of a graph already follow each other
// without a host-side gap, what is left between them is each kernel's own fill and drain.
template <class... P, class... A>
inline void launch_step(const sph_handle *h, void (*kern)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, A &&...args)
{
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at{};
    at.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at.val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = &at;
    cfg.numAttrs = h->pdl ? 1 : 0;
    (void)cudaLaunchKernelEx(&cfg, kern, std::forward<A>(args)...);  // a failure is picked up by CK_STEP_LAUNCH
}

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// SPHSettings::SPHSettings, src/SPHSystem.cpp:8-26. `pow(h, 9)` with a float and an int is the
// double pow; the float*float products are formed first and the quotient is rounded once when
// stored to the float member.
void derive(const sph_settings &s, sph_derived &d)
{
    const float h = s.h;
    d.poly6 = (float)((double)315.0f / ((double)(64.0f * REF_PI) * std::pow((double)h, 9.0)));
    d.spiky_grad = (float)((double)-45.0f / ((double)REF_PI * std::pow((double)h, 6.0)));
    d.spiky_lap = (float)((double)45.0f / ((double)REF_PI * std::pow((double)h, 6.0)));
    d.h2 = h * h;
    d.self_dens = (float)((double)(s.mass * d.poly6) * std::pow((double)h, 6.0));
    d.mass_poly6 = s.mass * d.poly6;
    d.sphere_scale = h / 2.f;
}

void make_params(const sph_settings &s, const sph_derived &d, Params &P)
{
    P.h = s.h;
    P.h2 = d.h2;
    P.mass = s.mass;
    P.mass_poly6 = s.mass * d.poly6;  // recomputed as src/sph.cpp:33 does
    P.self_dens = d.self_dens;
    P.gas_constant = s.gas_constant;
    P.rest_density = s.rest_density;
    P.visc_mass = s.viscosity * s.mass;
    P.spiky_grad = d.spiky_grad;
    P.spiky_lap = d.spiky_lap;
    P.g = s.g;
    P.h_minus_box = s.h - s.box_half_width;
    P.box_minus_h = -s.h + s.box_half_width;
    P.two_h = 2 * s.h;
    P.two_hmb = 2 * (s.h - s.box_half_width);
    P.two_nhmb = 2 * -(s.h - s.box_half_width);
    P.wall_offset = s.wall_offset;
    P.elasticity = s.elasticity;
    P.sphere_scale = d.sphere_scale;
    P.clump_cell = CLUMP_CELL;
    if (const char *e = std::getenv("SPH_B200_CLUMP_CELL")) {  // A/B and tests: 0 = every deferred row gets a warp
        const long v = std::atol(e);
        P.clump_cell = v <= 0 ? 0xFFFFFFFFu : (uint32_t)v;
    }
}

int validate_settings(sph_handle *h, const sph_settings *s)
{
    if (!s) return fail(h, SPH_ERR_INVALID, "settings is NULL");
    if (!(s->h > 0.f) || !std::isfinite(s->h)) return fail(h, SPH_ERR_INVALID, "settings.h must be a positive finite number");
    return SPH_OK;
}

int ensure_scratch(sph_handle *h, size_t bytes)
{
    if (bytes <= h->scratch_bytes) return SPH_OK;
    if (h->scratch) CK(cudaFree(h->scratch));
    h->scratch = nullptr;
    h->scratch_bytes = 0;
    CK(cudaMalloc(&h->scratch, bytes));
    h->scratch_bytes = bytes;
    return SPH_OK;
}

int enter(sph_handle *h)
{
    if (!h) return fail(nullptr, SPH_ERR_INVALID, "handle is NULL");
    CK(cudaSetDevice(h->device));
    return SPH_OK;
}

// Sync-free slab path: pick up the exact row count (and violation bits) the last build published.
int resolve_rows(sph_handle *h)
{
    if (!h->rows_pending) return SPH_OK;
    CK(cudaEventSynchronize(h->ev_rows));
    h->rows_pending = false;
    h->n = h->pinned_rows[0];
    const uint32_t err = h->pinned_rows[1];
    if (err)
        return fail(h, SPH_ERR_CAPACITY,
                    "slab exchange violation in an earlier step:%s%s%s%s (raise the message capacities or rebalance with the "
                    "general path)",
                    (err & SLAB_ERR_MIGRANT_OVERFLOW) ? " migrant message overflow" : "",
                    (err & SLAB_ERR_HALO_OVERFLOW) ? " halo message overflow" : "",
                    (err & SLAB_ERR_NOT_ADJACENT) ? " a particle left for a non-adjacent slab" : "",
                    (err & SLAB_ERR_P2P_TIMEOUT) ? " a neighbour never raised its mailbox flag (time-out)" : "");
    return SPH_OK;
}

// For calls that need the exact row count on the host (downloads, diagnostics, the general slab path).
int enter_exact(sph_handle *h)
{
    int rc = enter(h);
    if (rc) return rc;
    return resolve_rows(h);
}

// bbox of the current positions into bbox[parity] (after an upload; in steady state the
// integration kernel has already produced it).
int compute_bbox(sph_handle *h)
{
    k_reset_bbox<<<1, 32, 0, h->stream>>>(h->ctr, h->parity);
    CK_LAUNCH();
    if (h->n) {
        k_bbox<<<blocks_for(h->n, GRID_THREADS), GRID_THREADS, 0, h->stream>>>(h->pos[h->cur], (uint32_t)h->n, h->P.h,
                                                                            h->ctr, h->parity);
        CK_LAUNCH();
    }
    return SPH_OK;
}

// Neighbour-search build for the current positions: plan -> zero -> histogram -> scan -> place
// -> canonical order -> gather. Flips pos/vel buffers; afterwards rows are in cell order and
// h->cells holds the cell start offsets. Dropped rows (slab mode) do not survive the build, so
// h->n shrinks to the surviving row count.
int build_grid(sph_handle *h)
{
    h->edge_ok = false;  // rows move; a slab force step declares them ordered again
    const uint32_t n = (uint32_t)h->n;
    cudaStream_t s = h->stream;
    launch_step(h, k_plan_zero, persistent_blocks(h, 8, (uint64_t)h->max_cells / (4 * GRID_THREADS) + 1), GRID_THREADS, 0, s, h->ctr, h->gd, h->parity, h->max_cells, h->bbox_expand, h->cells);
    CK_STEP_LAUNCH();
    // Rows per thread of the three latency-bound build kernels (SPH_B200_GRID_CFG="hist,place,gather" for A/B: same bits)
#define LAUNCH_HIST(R)                                                                                                     \
    launch_step(h, k_cell_hist<R>, blocks_for(n, R * GRID_THREADS), GRID_THREADS, 0, s, h->pos[h->cur], n, h->P.h, h->gd,  \
                h->cells, h->cell_rank, h->ctr)
    switch (h->grid_rows[0]) {
    case 1: LAUNCH_HIST(1); break;
    case 2: LAUNCH_HIST(2); break;
    case 8: LAUNCH_HIST(8); break;
    default: LAUNCH_HIST(4); break;
    }
#undef LAUNCH_HIST
    CK_STEP_LAUNCH();
    launch_step(h, k_scan_exclusive, persistent_blocks(h, h->scan_blocks, (uint64_t)h->max_cells / SCAN_TILE + 2), SCAN_THREADS, 0, s, h->cells, &h->gd->ncells, h->tile_state,
                &h->ctr->ticket, &h->ctr->epoch);
    CK_STEP_LAUNCH();
#define LAUNCH_PLACE(R)                                                                                                    \
    launch_step(h, k_place<R>, blocks_for(n, R * GRID_THREADS), GRID_THREADS, 0, s, h->cell_rank, h->pos[h->cur], n, h->cells, \
                h->slot, h->gd, h->slab_mode ? h->ctr : nullptr, h->slab_lo, h->slab_hi)
    switch (h->grid_rows[1]) {
    case 1: LAUNCH_PLACE(1); break;
    case 4: LAUNCH_PLACE(4); break;
    default: LAUNCH_PLACE(2); break;
    }
#undef LAUNCH_PLACE
    CK_STEP_LAUNCH();
    uint32_t n_sorted = n;
    const uint32_t *n_dev = nullptr;
    if (h->slab_mode) {
        if (h->slab_fast) {
            // No host sync: later kernels run over the bound n and skip the dropped tail; the exact
            // count (and the violation bits next to it) reach the host before the next step.
            CK(cudaMemcpyAsync(h->pinned_rows, &h->ctr->aux[2], 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
            CK(cudaEventRecord(h->ev_rows, s));
            h->rows_pending = true;
            n_dev = &h->ctr->aux[2];
        } else {
            CK(cudaMemcpyAsync(&n_sorted, &h->ctr->aux[2], sizeof n_sorted, cudaMemcpyDeviceToHost, s));
            CK(cudaStreamSynchronize(s));
        }
    }
    if (n_sorted) {
#define LAUNCH_GATHER(F, R)                                                                                                \
    launch_step(h, k_order_gather<F, R>, blocks_for(n_sorted, R * GRID_THREADS), GRID_THREADS, 0, s, h->slot, h->cell_rank,   \
                n_sorted, n_dev, h->cells, h->P.h, h->pos[h->cur], h->vel[h->cur], h->pos[h->cur ^ 1], h->vel[h->cur ^ 1],    \
                h->hash16, h->slab_mode ? h->inverse : nullptr, h->gd)
        switch (h->grid_rows[2]) {
        case 0: LAUNCH_GATHER(false, 1); break;  // cell looked up in cell_rank (round-1 form)
        case 2: LAUNCH_GATHER(true, 2); break;
        case 4: LAUNCH_GATHER(true, 4); break;
        default: LAUNCH_GATHER(true, 1); break;
        }
#undef LAUNCH_GATHER
        CK_STEP_LAUNCH();
    }
    h->cur ^= 1;
    h->n = n_sorted;
    return SPH_OK;
}

// Density pass. Default: the staged row walk (sph_physics.cuh); SPH_B200_DENSITY_CFG >= 10 selects the
// round-1 kernel and its launch shapes, 1..3 other launch shapes of the staged walk (A/B measurements:
// all variants produce bit-identical lists and densities). The staged walk keeps 32-bit list offsets;
// capacities beyond that take the round-1 kernel.
int launch_density(sph_handle *h, uint32_t n)
{
    cudaStream_t s = h->stream;
#define LAUNCH_D(S, B, U)                                                                                  \
    launch_step(h, k_density<S, B, U>, blocks_for(n, PHYS_THREADS), PHYS_THREADS, 0, s,                       \
        h->pos[h->cur], n, h->gd, h->cells, h->P, h->vel[h->cur], h->nlist, h->ncount, (uint32_t)h->cap,      \
        h->order, h->ctr)
#define LAUNCH_S(B, U)                                                                                     \
    launch_step(h, k_density_staged<B, U>, blocks_for(n, PHYS_THREADS), PHYS_THREADS, 0, s,                   \
        h->pos[h->cur], n, h->gd, h->cells, h->P, h->vel[h->cur], h->nlist, h->ncount, (uint32_t)h->cap,      \
        h->order, h->ctr)
    int cfg = h->density_cfg;
    // Default shape by size: once the rows no longer fit L2 (3 M rows x 32 B) the walk waits on memory and the
    // 32-register shape's extra warps pay (0.398 vs 0.422 ms at 8 M rows); below, the 40-register one wins.
    // Same bits either way.
    if (cfg == 0 && n >= 3000000u) cfg = 3;
    if ((cfg < 10 || cfg >= 50) && (uint64_t)(NLIST_ROWS + 1) * h->cap >= (1ull << 32)) cfg = 10;
    if (cfg >= 50) {
        // TMA-staged neighbourhoods (measured alternative): persistent blocks, 3 per SM at 58 KB each
        auto kern = cfg == 51 ? k_density_tma<4> : k_density_tma<2>;
        if (!h->tma_armed) {
            CK(cudaFuncSetAttribute(k_density_tma<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TmaTile)));
            CK(cudaFuncSetAttribute(k_density_tma<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TmaTile)));
            h->tma_armed = true;
        }
        const unsigned tiles = blocks_for(n, PHYS_THREADS);
        launch_step(h, kern, std::min(tiles, (unsigned)h->num_sms * 3u), PHYS_THREADS, sizeof(TmaTile), s,
            h->pos[h->cur], n, h->gd, h->cells, h->P, h->vel[h->cur], h->nlist, h->ncount, (uint32_t)h->cap, h->order, h->ctr);
    } else
    switch (cfg) {
    case 1: LAUNCH_S(12, 4); break;
    case 2: LAUNCH_S(10, 2); break;
    case 3: LAUNCH_S(16, 2); break;
    case 10: LAUNCH_D(24, 12, 2); break;  // round-1 default
    case 11: LAUNCH_D(24, 1, 4); break;
    case 12: LAUNCH_D(24, 1, 2); break;
    default: LAUNCH_S(12, 2); break;  // 40 registers, 12.8 KB of shared memory: best at 1 M (dense); cfg 3 wins by 5 % at 8 M (sparse)
    }
#undef LAUNCH_D
#undef LAUNCH_S
    CK_STEP_LAUNCH();
    // the heavy tail (clumps, hash-collision cells), one warp per deferred particle; exits at once when empty
    launch_step(h, k_density_heavy, persistent_blocks(h, h->heavy_blocks, n / 64 + 8), HEAVY_THREADS, 0, s, h->pos[h->cur], n, h->gd, h->cells, h->P, h->vel[h->cur],
                h->nlist, h->ncount, (uint32_t)h->cap, h->order, h->tile_claim, h->tile_claim + h->cap / CLUMP_ROWS + 1, h->ctr);
    CK_STEP_LAUNCH();
    return SPH_OK;
}

// Launch of the fused forces+integration kernel. mode: FI_STEP (resident simulation: no force
// column), FI_STEP_WRITE_FORCE (the stateless drop-in call fills Particle::force every step) or
// FI_FORCE_ONLY (recompute the force column of the last step on demand; start-of-step rows are in
// the other buffers). SPH_B200_FORCES_CFG selects launch shapes for experiments.
int launch_forces_integrate(sph_handle *h, uint32_t n, float dt, int mode, int part = 0)
{
    cudaStream_t s = h->stream;
    // FI_FORCE_ONLY reads the start-of-step rows, which after the step live in the non-current buffers
    const int in = mode == FI_FORCE_ONLY ? (h->cur ^ 1) : h->cur;
#define LAUNCH_FI(T, B, M) LAUNCH_FIP(T, B, M, true, 1)
#define LAUNCH_FIP(T, B, M, PK, PIPE)                                                                            \
    launch_step(h, k_forces_integrate<T, B, M, PK, PIPE>, blocks_for(n, T), T, 0, s,                             \
        h->pos[in], h->vel[in], n, h->gd, h->cells, h->P, h->nlist, h->ncount, (uint32_t)h->cap, dt,              \
        h->pos[in ^ 1], h->vel[in ^ 1], h->force, h->ctr, h->parity ^ 1, h->map, part)
#define LAUNCH_FH(M)                                                                                              \
    launch_step(h, k_forces_heavy<M>, persistent_blocks(h, h->heavy_blocks, n / 64 + 8), HEAVY_THREADS, 0, s, h->pos[in], h->vel[in], n, h->gd, h->cells, \
                h->P, h->ncount, dt, h->pos[in ^ 1], h->vel[in ^ 1], h->force, h->ctr, h->parity ^ 1, h->map,             \
                h->tile_claim + h->cap / CLUMP_ROWS + 1)
    if (mode == FI_FORCE_ONLY) {
        CK(cudaMemsetAsync(&h->ctr->heavy[1], 0, sizeof(uint32_t), s));  // the step's deferral list is rebuilt
        CK(cudaMemsetAsync(&h->ctr->clump_rows[1], 0, sizeof(uint32_t), s));
        CK(cudaMemsetAsync(&h->ctr->clump_ticket[1], 0, 2 * sizeof(uint32_t), s));  // force list and tile tickets
        LAUNCH_FI(128, 12, FI_FORCE_ONLY);
    } else if (mode == FI_STEP_WRITE_FORCE) {
        LAUNCH_FI(128, 12, FI_STEP_WRITE_FORCE);
    } else {
#define LAUNCH_FT(T, B, CAPR)                                                                                     \
    do {                                                                                                         \
        auto kern = k_forces_tile<T, B, FI_STEP, CAPR>;                                                          \
        if (!h->tile_armed) { /* per handle: the attribute belongs to the handle's device */                    \
            CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * CAPR));              \
            h->tile_armed = true;                                                                                \
        }                                                                                                        \
        launch_step(h, kern, blocks_for(n, T), T, 32 * CAPR, s, h->pos[in], h->vel[in], n, h->gd, h->cells, h->P,  \
                    h->nlist, h->ncount, (uint32_t)h->cap, dt, h->pos[in ^ 1], h->vel[in ^ 1], h->force, h->ctr,    \
                    h->parity ^ 1, h->map);                                                                      \
    } while (0)
        switch ((part != 0 && h->forces_cfg == 6) ? 2 : h->forces_cfg) {  // the tile-staged variant has no split form
        case 6: LAUNCH_FT(128, 5, 1408); break;  // shared-memory staged neighbourhoods: measured 2-4x slower
        case 7: LAUNCH_FIP(128, 10, FI_STEP, false, 1); break;  // the scalar form of the force terms (A/B against the packed default: same bits)
        case 8: LAUNCH_FIP(128, 10, FI_STEP, true, 0); break;   // round-2-start form: no software pipeline, 48 registers
        case 9: LAUNCH_FIP(128, 10, FI_STEP, true, 1); break;
        case 10: LAUNCH_FIP(128, 10, FI_STEP, true, 2); break;  // next neighbour's rows in flight too
        case 11: LAUNCH_FIP(128, 12, FI_STEP, true, 2); break;
        case 12: LAUNCH_FIP(128, 8, FI_STEP, true, 2); break;
        case 1: LAUNCH_FI(128, 8, FI_STEP); break;
        case 3: LAUNCH_FI(64, 16, FI_STEP); break;
        case 4: LAUNCH_FI(256, 4, FI_STEP); break;
        case 5: LAUNCH_FI(128, 16, FI_STEP); break;
        default: LAUNCH_FI(128, 12, FI_STEP); break;  // list entry one neighbour ahead, 40 registers
        }
    }
#undef LAUNCH_FI
#undef LAUNCH_FIP
    CK_STEP_LAUNCH();
    if (part == 1) return SPH_OK;  // the rows outside the interior range follow in a second launch, which completes the pass
    if (mode == FI_FORCE_ONLY) LAUNCH_FH(FI_FORCE_ONLY);
    else if (mode == FI_STEP_WRITE_FORCE) LAUNCH_FH(FI_STEP_WRITE_FORCE);
    else LAUNCH_FH(FI_STEP);
#undef LAUNCH_FH
    CK_STEP_LAUNCH();
    if (mode != FI_FORCE_ONLY) {
        h->cur ^= 1;
        h->have_force = mode == FI_STEP_WRITE_FORCE;
    } else {
        h->have_force = true;
    }
    return SPH_OK;
}

// Make sure the force column of the last step exists (it is not written by a plain step).
int ensure_force(sph_handle *h)
{
    if (h->have_force) return SPH_OK;
    if (!h->have_step) return fail(h, SPH_ERR_STATE, "forces are only defined after a step");
    return launch_forces_integrate(h, (uint32_t)h->n, 0.f, FI_FORCE_ONLY);
}

int step_once(sph_handle *h, float dt)
{
    const uint32_t n = (uint32_t)h->n;
    cudaStream_t s = h->stream;
    bool timed = h->timing && h->ev_used < sph_handle::kMaxTimedSteps;
    cudaEvent_t *ev = nullptr;
    if (timed) {
        if (h->ev_used == h->ev_pool.size()) {
            PassEvents pe;
            for (auto &e : pe.e) CK(cudaEventCreate(&e));
            h->ev_pool.push_back(pe);
        }
        ev = h->ev_pool[h->ev_used++].e;
    }
    if (timed) CK(cudaEventRecord(ev[0], s));
    int rc = build_grid(h);
    if (rc) return rc;
    if (timed) CK(cudaEventRecord(ev[1], s));
    rc = launch_density(h, n);
    if (rc) return rc;
    if (timed) CK(cudaEventRecord(ev[2], s));
    rc = launch_forces_integrate(h, n, dt, h->write_force ? FI_STEP_WRITE_FORCE : FI_STEP);
    if (rc) return rc;
    if (timed) CK(cudaEventRecord(ev[3], s));
    if (timed) CK(cudaEventRecord(ev[4], s));
    h->parity ^= 1;
    ++h->steps;
    h->have_step = true;
    return SPH_OK;
}

void drop_graphs(sph_handle *h)
{
    for (auto &row : h->graphs)
        for (auto &g : row)
            if (g) { cudaGraphExecDestroy(g); g = nullptr; }
}

// Captures `len` steps from the current host state into an executable graph. Nothing runs, so the
// bookkeeping step_once did on the host is rolled back.
int capture_steps(sph_handle *h, float dt, int len, cudaGraphExec_t *out)
{
    const int cur = h->cur, parity = h->parity;
    const uint64_t steps = h->steps, launches = h->launches;
    const bool have_step = h->have_step, have_force = h->have_force;
    CK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
    int rc = SPH_OK;
    for (int k = 0; k < len && !rc; ++k) rc = step_once(h, dt);
    cudaGraph_t g = nullptr;
    const cudaError_t e = cudaStreamEndCapture(h->stream, &g);
    h->graph_launches[len == 1 ? 0 : 1] = h->launches - launches;
    h->cur = cur; h->parity = parity; h->steps = steps; h->launches = launches;
    h->have_step = have_step; h->have_force = have_force;
    if (rc) {
        if (g) cudaGraphDestroy(g);
        return rc;
    }
    CK(e);
    const cudaError_t ei = cudaGraphInstantiate(out, g, 0);
    cudaGraphDestroy(g);
    CK(ei);
    return SPH_OK;
}

// nsteps resident steps through the captured graphs.
int step_graphs(sph_handle *h, float dt, int nsteps)
{
    sph_handle::GraphKey key;
    std::memset(&key, 0, sizeof key);
    key.n = h->n; key.dt = dt; key.cur = h->cur; key.write_force = h->write_force; key.bbox_expand = h->bbox_expand;
    key.forces_cfg = h->forces_cfg; key.density_cfg = h->density_cfg; key.max_cells = h->max_cells; key.P = h->P;
    if (std::memcmp(&key, &h->graph_key, sizeof key)) {
        drop_graphs(h);
        h->graph_key = key;
    }
    for (int k = 0; k < nsteps;) {
        const int li = nsteps - k >= sph_handle::kGraphLong ? 1 : 0;
        const int len = li ? sph_handle::kGraphLong : 1;
        cudaGraphExec_t &ex = h->graphs[li][h->parity];
        if (!ex) {
            int rc = capture_steps(h, dt, len, &ex);
            if (rc) return rc;
        }
        CK(cudaGraphLaunch(ex, h->stream));
        h->steps += len;
        h->launches += h->graph_launches[li];
        if (len & 1) h->parity ^= 1;
        h->have_step = true;
        h->have_force = h->write_force;
        k += len;
    }
    return SPH_OK;
}

// map[d] = device row that lands at row d when rows are stably sorted by start-of-step hash16.
// Leaves the bucket start offsets (65537 entries) in h->h16_cells.
int build_hash16_order(sph_handle *h, bool need_map)
{
    const uint32_t n = (uint32_t)h->n;
    cudaStream_t s = h->stream;
    CK(cudaMemsetAsync(h->h16_cells, 0, sizeof(uint32_t) * 65540, s));
    k_scan_arm<<<1, 32, 0, s>>>(&h->ctr->aux[0], &h->ctr->epoch);
    CK_LAUNCH();
    if (n) {
        k_hash16_hist<<<blocks_for(n, IO_THREADS), IO_THREADS, 0, s>>>(h->hash16, n, h->h16_cells,
                                                                     need_map ? h->cell_rank : nullptr);
        CK_LAUNCH();
    }
    k_scan_exclusive<<<32, SCAN_THREADS, 0, s>>>(h->h16_cells, h->const_65536, h->tile_state, &h->ctr->aux[0],
                                                &h->ctr->epoch);
    CK_LAUNCH();
    if (need_map && n) {
        k_place<1><<<blocks_for(n, GRID_THREADS), GRID_THREADS, 0, s>>>(h->cell_rank, h->pos[h->cur], n, h->h16_cells,
                                                                   h->slot, nullptr, nullptr, 0, 0);
        CK_LAUNCH();
        k_stable_order<<<blocks_for(n, GRID_THREADS), GRID_THREADS, 0, s>>>(h->slot, h->cell_rank, n, h->h16_cells,
                                                                          h->map, nullptr);
        CK_LAUNCH();
    }
    return SPH_OK;
}

int after_upload(sph_handle *h, uint64_t n)
{
    h->edge_ok = false;
    h->n = n;
    h->steps = 0;
    h->have_state = true;
    h->have_step = false;
    h->n_ghost = 0;
    h->ghost_n[0] = h->ghost_n[1] = h->halo_n[0] = h->halo_n[1] = 0;
    h->have_bbox_from_integration = false;
    CK(cudaMemsetAsync(&h->ctr->aux[3], 0, sizeof(uint32_t), h->stream));  // slab violation word
    return compute_bbox(h);
}

}  // namespace

// =================================================================================================
extern "C" {

int sph_settings_default(sph_settings *out)
{
    if (!out) return SPH_ERR_INVALID;
    out->mass = 0.02f;          // src/Tester.cpp:90
    out->rest_density = 1000.f;
    out->gas_constant = 1.f;
    out->viscosity = 1.04f;
    out->h = 0.15f;
    out->g = -9.8f;
    out->tension = 0.2f;
    out->dt = 0.003f;           // src/SPHSystem.cpp:113
    out->box_half_width = 8.f;  // src/sph.cpp:139
    out->elasticity = 0.5f;     // src/sph.cpp:140
    out->wall_offset = 0.0001f; // src/sph.cpp:154
    return SPH_OK;
}

int sph_settings_derive(const sph_settings *s, sph_derived *out)
{
    if (!s || !out) return SPH_ERR_INVALID;
    derive(*s, *out);
    return SPH_OK;
}

const char *sph_last_error(const sph_handle *h) { return h ? h->err : g_create_error; }

int sph_create(const sph_settings *s, uint64_t capacity, int device, sph_handle **out)
{
    sph_handle *h = nullptr;  // errors before the handle exists go to the thread-local buffer
    if (!out) return fail(h, SPH_ERR_INVALID, "out is NULL");
    *out = nullptr;
    int rc = validate_settings(h, s);
    if (rc) return rc;
    if (capacity == 0 || capacity > 0xfffffff0ull)
        return fail(h, SPH_ERR_INVALID, "capacity must be in [1, 2^32-16], got %llu", (unsigned long long)capacity);
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev)
        return fail(h, SPH_ERR_INVALID, "device %d out of range (%d CUDA devices visible)", device, ndev);
    CK(cudaSetDevice(device));

    sph_handle *nh = new (std::nothrow) sph_handle();
    if (!nh) return fail(h, SPH_ERR_INVALID, "out of host memory");
    nh->device = device;
    nh->cap = capacity;
    nh->settings = *s;
    derive(*s, nh->derived);
    make_params(*s, nh->derived, nh->P);

    // From here on errors are recorded in the new handle and copied out on failure.
    h = nh;
    auto bail = [&](int code) {
        std::snprintf(g_create_error, sizeof g_create_error, "%s", nh->err);
        sph_destroy(nh);
        return code;
    };
#define CKC(call)                                                                                 \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess) {                                                                  \
            fail(h, SPH_ERR_CUDA, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__,            \
                 cudaGetErrorString(e_));                                                         \
            return bail(SPH_ERR_CUDA);                                                            \
        }                                                                                         \
    } while (0)

    cudaDeviceProp prop;
    CKC(cudaGetDeviceProperties(&prop, device));
    nh->num_sms = prop.multiProcessorCount;
    CKC(cudaStreamCreateWithFlags(&nh->stream, cudaStreamNonBlocking));

    // Dense-grid allocation: 32 cells per particle of capacity, within [2^22, 2^30]
    // (overridable for experiments with SPH_B200_MAX_CELLS).
    uint64_t mc = capacity * 32ull;
    if (mc < (1ull << 22)) mc = 1ull << 22;
    if (mc > (1ull << 30)) mc = 1ull << 30;
    if (const char *e = std::getenv("SPH_B200_MAX_CELLS")) {
        const unsigned long long v = std::strtoull(e, nullptr, 10);
        if (v >= 27 && v <= (1ull << 31)) mc = v;
    }
    nh->max_cells = (uint32_t)mc;
    nh->forces_cfg = 2;  // 128 threads, <= 48 registers: the pass is latency-bound, occupancy wins
    if (const char *e = std::getenv("SPH_B200_FORCES_CFG")) nh->forces_cfg = std::atoi(e);
    if (const char *e = std::getenv("SPH_B200_DENSITY_CFG")) nh->density_cfg = std::atoi(e);
    if (const char *e = std::getenv("SPH_B200_SCAN_BLOCKS")) nh->scan_blocks = std::min(std::max(std::atoi(e), 1), 8);
    if (const char *e = std::getenv("SPH_B200_GRID_CFG")) std::sscanf(e, "%d,%d,%d", &nh->grid_rows[0], &nh->grid_rows[1], &nh->grid_rows[2]);
    if (const char *e = std::getenv("SPH_B200_HEAVY_BLOCKS")) nh->heavy_blocks = std::min(std::max(std::atoi(e), 1), HEAVY_BLOCKS_PER_SM);
    if (const char *e = std::getenv("SPH_B200_GRAPH")) nh->graph_enabled = std::atoi(e) != 0;
    if (const char *e = std::getenv("SPH_B200_PDL")) nh->pdl = std::atoi(e) != 0;
    if (const char *e = std::getenv("SPH_B200_EDGE_SCAN")) nh->edge_scan_enabled = std::atoi(e) != 0;
    std::memset(&nh->graph_key, 0, sizeof nh->graph_key);

    const size_t cap = (size_t)capacity;
    for (int b = 0; b < 2; ++b) {
        CKC(cudaMalloc(&nh->pos[b], sizeof(float4) * cap));
        CKC(cudaMalloc(&nh->vel[b], sizeof(float4) * cap));
    }
    CKC(cudaMalloc(&nh->force, sizeof(float4) * cap));
    CKC(cudaMalloc(&nh->hash16, sizeof(uint32_t) * cap));
    CKC(cudaMalloc(&nh->nlist, sizeof(uint32_t) * cap * NLIST_ROWS));
    CKC(cudaMalloc(&nh->ncount, sizeof(uint32_t) * cap));
    CKC(cudaMalloc(&nh->cell_rank, sizeof(uint2) * cap));
    CKC(cudaMalloc(&nh->slot, sizeof(uint2) * cap));
    CKC(cudaMalloc(&nh->inverse, sizeof(uint32_t) * cap));
    CKC(cudaMalloc(&nh->halo_rows[0], sizeof(uint32_t) * cap));
    CKC(cudaMalloc(&nh->halo_rows[1], sizeof(uint32_t) * cap));
    CKC(cudaMalloc(&nh->slab_counts, sizeof(unsigned long long) * (2 * SLAB_MAX_RANKS + 8)));
    CKC(cudaMalloc(&nh->order, sizeof(uint32_t) * cap));
    CKC(cudaMalloc(&nh->map, sizeof(uint32_t) * cap));
    CKC(cudaMalloc(&nh->tile_claim, sizeof(uint32_t) * 2 * (cap / CLUMP_ROWS + 1)));
    CKC(cudaMemsetAsync(nh->tile_claim, 0, sizeof(uint32_t) * 2 * (cap / CLUMP_ROWS + 1), nh->stream));
    CKC(cudaMalloc(&nh->cells, sizeof(uint32_t) * ((size_t)nh->max_cells + 8)));
    CKC(cudaMalloc(&nh->h16_cells, sizeof(uint32_t) * 65540));
    CKC(cudaMalloc(&nh->const_65536, sizeof(uint32_t)));
    const size_t ntile = ((size_t)nh->max_cells + 1 + SCAN_TILE - 1) / SCAN_TILE + 1;
    CKC(cudaMalloc(&nh->tile_state, sizeof(unsigned long long) * ntile));
    CKC(cudaMemsetAsync(nh->tile_state, 0, sizeof(unsigned long long) * ntile, nh->stream));
    CKC(cudaMalloc(&nh->gd, sizeof(GridDesc)));
    CKC(cudaMalloc(&nh->ctr, sizeof(StepCounters)));
    CKC(cudaMemsetAsync(nh->ctr, 0, sizeof(StepCounters), nh->stream));
    CKC(cudaMemsetAsync(nh->gd, 0, sizeof(GridDesc), nh->stream));
    CKC(cudaMalloc(&nh->stats_acc, sizeof(StatsAccum)));
    CKC(cudaHostAlloc(&nh->pinned_rows, 2 * sizeof(uint32_t), cudaHostAllocDefault));
    nh->pinned_rows[0] = nh->pinned_rows[1] = 0;
    CKC(cudaEventCreateWithFlags(&nh->ev_rows, cudaEventDisableTiming));
    {
        // highest priority: its few small exchange kernels must get SM slots while the main stream's force pass has
        // tens of thousands of blocks queued (at equal priority they only start when that grid has been issued)
        int prio_lo = 0, prio_hi = 0;
        CKC(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        CKC(cudaStreamCreateWithPriority(&nh->stream2, cudaStreamNonBlocking, prio_hi));
    }
    CKC(cudaEventCreateWithFlags(&nh->ev_dens, cudaEventDisableTiming));
    CKC(cudaEventCreateWithFlags(&nh->ev_rho, cudaEventDisableTiming));
    if (const char *e = std::getenv("SPH_B200_P2P_OVERLAP")) nh->p2p_overlap = std::atoi(e) != 0;
    const uint32_t c65536 = 65536u;
    CKC(cudaMemcpyAsync(nh->const_65536, &c65536, sizeof c65536, cudaMemcpyHostToDevice, nh->stream));
    CKC(cudaStreamSynchronize(nh->stream));
#undef CKC
    *out = nh;
    return SPH_OK;
}

int sph_destroy(sph_handle *h)
{
    if (!h) return SPH_ERR_INVALID;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    for (int b = 0; b < 2; ++b) { cudaFree(h->pos[b]); cudaFree(h->vel[b]); }
    cudaFree(h->force); cudaFree(h->hash16); cudaFree(h->nlist); cudaFree(h->ncount); cudaFree(h->cell_rank); cudaFree(h->slot); cudaFree(h->inverse);
    cudaFree(h->halo_rows[0]); cudaFree(h->halo_rows[1]); cudaFree(h->slab_counts);
    cudaFree(h->order); cudaFree(h->map); cudaFree(h->tile_claim); cudaFree(h->cells); cudaFree(h->h16_cells);
    cudaFree(h->const_65536); cudaFree(h->tile_state); cudaFree(h->gd); cudaFree(h->ctr);
    cudaFree(h->stats_acc); cudaFree(h->scratch); cudaFree(h->reset_pos); cudaFree(h->reset_vel);
    for (int k = 0; k < 2; ++k)
        if (h->peer_mailbox[k]) cudaIpcCloseMemHandle(h->peer_mailbox[k]);
    cudaFree(h->mailbox); cudaFree(h->mig_rows[0]); cudaFree(h->mig_rows[1]);
    drop_graphs(h);
    if (h->pinned_rows) cudaFreeHost(h->pinned_rows);
    if (h->ev_rows) cudaEventDestroy(h->ev_rows);
    if (h->ev_dens) cudaEventDestroy(h->ev_dens);
    if (h->ev_rho) cudaEventDestroy(h->ev_rho);
    if (h->stream2) { cudaStreamSynchronize(h->stream2); cudaStreamDestroy(h->stream2); }
    for (auto &pe : h->ev_pool)
        for (auto &e : pe.e) if (e) cudaEventDestroy(e);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return SPH_OK;
}

int sph_set_settings(sph_handle *h, const sph_settings *s)
{
    int rc = enter_exact(h);
    if (rc) return rc;
    rc = validate_settings(h, s);
    if (rc) return rc;
    CK(cudaStreamSynchronize(h->stream));
    h->settings = *s;
    derive(*s, h->derived);
    make_params(*s, h->derived, h->P);
    h->have_step = false;
    if (h->have_state) return compute_bbox(h);  // cells depend on h
    return SPH_OK;
}

uint64_t sph_count(const sph_handle *h) { return h ? h->n : 0; }
uint64_t sph_capacity(const sph_handle *h) { return h ? h->cap : 0; }
void *sph_stream(sph_handle *h) { return h ? (void *)h->stream : nullptr; }

// ---- state in / out ---------------------------------------------------------------------------

int sph_upload(sph_handle *h, uint64_t n, const float *host_pos, const float *host_vel, const uint32_t *host_id)
{
    int rc = enter_exact(h);
    if (rc) return rc;
    if (n > h->cap) return fail(h, SPH_ERR_CAPACITY, "upload of %llu particles exceeds capacity %llu",
                                (unsigned long long)n, (unsigned long long)h->cap);
    if (n && (!host_pos || !host_vel)) return fail(h, SPH_ERR_INVALID, "pos/vel is NULL");
    if (host_id)
        for (uint64_t k = 0; k < n; ++k)
            if (host_id[k] & W_GHOST)
                return fail(h, SPH_ERR_INVALID, "id %u of row %llu has bit 31 set: ids must be unique and below 2^31", host_id[k],
                            (unsigned long long)k);
    const size_t b3 = align_up(sizeof(float) * 3 * n, 256), b1 = align_up(sizeof(uint32_t) * n, 256);
    rc = ensure_scratch(h, 2 * b3 + b1 + 256);
    if (rc) return rc;
    char *sc = (char *)h->scratch;
    float *dpos = (float *)sc, *dvel = (float *)(sc + b3);
    uint32_t *did = (uint32_t *)(sc + 2 * b3);
    if (n) {
        CK(cudaMemcpyAsync(dpos, host_pos, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, h->stream));
        CK(cudaMemcpyAsync(dvel, host_vel, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, h->stream));
        if (host_id) CK(cudaMemcpyAsync(did, host_id, sizeof(uint32_t) * n, cudaMemcpyHostToDevice, h->stream));
        k_import_xyz<<<blocks_for(n, IO_THREADS), IO_THREADS, 0, h->stream>>>(dpos, dvel, host_id ? did : nullptr,
                                                                            (uint32_t)n, h->pos[h->cur], h->vel[h->cur]);
        CK_LAUNCH();
    }
    rc = after_upload(h, n);
    if (rc) return rc;
    CK(cudaStreamSynchronize(h->stream));  // the caller may reuse its buffers
    return SPH_OK;
}

int sph_upload_device(sph_handle *h, uint64_t n, const void *dev_pos, const void *dev_vel)
{
    int rc = enter_exact(h);
    if (rc) return rc;
    if (n > h->cap) return fail(h, SPH_ERR_CAPACITY, "upload of %llu particles exceeds capacity %llu",
                                (unsigned long long)n, (unsigned long long)h->cap);
    if (n && (!dev_pos || !dev_vel)) return fail(h, SPH_ERR_INVALID, "pos/vel is NULL");
    if (n) {
        k_import_xyzw<<<blocks_for(n, IO_THREADS), IO_THREADS, 0, h->stream>>>((const float4 *)dev_pos, (const float4 *)dev_vel,
                                                                             (uint32_t)n, h->pos[h->cur], h->vel[h->cur]);
        CK_LAUNCH();
    }
    rc = after_upload(h, n);
    if (rc) return rc;
    CK(cudaStreamSynchronize(h->stream));
    return SPH_OK;
}

int sph_download(sph_handle *h, int order, float *host_pos, float *host_vel, float *host_force, float *host_density,
                 float *host_pressure, uint16_t *host_hash16, uint32_t *host_id)
{
    int rc = enter_exact(h);
    if (rc) return rc;
    if (!h->have_state) return fail(h, SPH_ERR_STATE, "no particles uploaded");
    if ((host_force || host_density || host_pressure || host_hash16 || order == SPH_ORDER_HASH16) && !h->have_step)
        return fail(h, SPH_ERR_STATE, "force/density/pressure/hash16 are only defined after a step");
    if (order != SPH_ORDER_DEVICE && order != SPH_ORDER_ID && order != SPH_ORDER_HASH16)
        return fail(h, SPH_ERR_INVALID, "unknown order %d", order);
    const uint64_t n = h->n;
    if (n == 0) return SPH_OK;
    if (host_force) {
        rc = ensure_force(h);
        if (rc) return rc;
    }
    const size_t b3 = align_up(sizeof(float) * 3 * n, 256), b1 = align_up(sizeof(float) * n, 256);
    rc = ensure_scratch(h, 3 * b3 + 4 * b1 + 256);
    if (rc) return rc;
    char *sc = (char *)h->scratch;
    ExportPtrs out{};
    out.pos3 = host_pos ? (float *)sc : nullptr;
    out.vel3 = host_vel ? (float *)(sc + b3) : nullptr;
    out.force3 = host_force ? (float *)(sc + 2 * b3) : nullptr;
    out.density = host_density ? (float *)(sc + 3 * b3) : nullptr;
    out.pressure = host_pressure ? (float *)(sc + 3 * b3 + b1) : nullptr;
    out.id = host_id ? (uint32_t *)(sc + 3 * b3 + 2 * b1) : nullptr;
    out.hash16 = host_hash16 ? (uint16_t *)(sc + 3 * b3 + 3 * b1) : nullptr;

    const uint32_t *map = nullptr;
    if (order == SPH_ORDER_ID) {
        CK(cudaMemsetAsync(&h->ctr->aux[1], 0, sizeof(uint32_t), h->stream));
        k_rows_by_id<<<blocks_for(n, IO_THREADS), IO_THREADS, 0, h->stream>>>(h->pos[h->cur], (uint32_t)n, h->map,
                                                                            &h->ctr->aux[1]);
        CK_LAUNCH();
        uint32_t bad = 0;
        CK(cudaMemcpyAsync(&bad, &h->ctr->aux[1], sizeof bad, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        if (bad) return fail(h, SPH_ERR_STATE, "SPH_ORDER_ID needs ids in [0, n): %u rows are out of range", bad);
        map = h->map;
    } else if (order == SPH_ORDER_HASH16) {
        rc = build_hash16_order(h, true);
        if (rc) return rc;
        map = h->map;
    }
    k_export<<<blocks_for(n, IO_THREADS), IO_THREADS, 0, h->stream>>>(h->pos[h->cur], h->vel[h->cur], h->force, h->hash16,
                                                                    (uint32_t)n, map, h->P, out);
    CK_LAUNCH();
    if (host_pos) CK(cudaMemcpyAsync(host_pos, out.pos3, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost, h->stream));
    if (host_vel) CK(cudaMemcpyAsync(host_vel, out.vel3, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost, h->stream));
    if (host_force) CK(cudaMemcpyAsync(host_force, out.force3, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost, h->stream));
    if (host_density) CK(cudaMemcpyAsync(host_density, out.density, sizeof(float) * n, cudaMemcpyDeviceToHost, h->stream));
    if (host_pressure) CK(cudaMemcpyAsync(host_pressure, out.pressure, sizeof(float) * n, cudaMemcpyDeviceToHost, h->stream));
    if (host_id) CK(cudaMemcpyAsync(host_id, out.id, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, h->stream));
    if (host_hash16) CK(cudaMemcpyAsync(host_hash16, out.hash16, sizeof(uint16_t) * n, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return SPH_OK;
}

int sph_read_positions_device(sph_handle *h, void *dev_xyzw)
{
    int rc = enter_exact(h);
    if (rc) return rc;
    if (!h->have_state) return fail(h, SPH_ERR_STATE, "no particles uploaded");
    if (!dev_xyzw) return fail(h, SPH_ERR_INVALID, "destination is NULL");
    if (h->n) {
        k_positions_xyz1<<<blocks_for(h->n, IO_THREADS), IO_THREADS, 0, h->stream>>>(h->pos[h->cur], (uint32_t)h->n,
                                                                                   (float4 *)dev_xyzw);
        CK_LAUNCH();
    }
    return SPH_OK;
}

int sph_read_positions(sph_handle *h, float *host_xyzw)
{
    int rc = enter_exact(h);
    if (rc) return rc;
    if (!host_xyzw) return fail(h, SPH_ERR_INVALID, "destination is NULL");
    rc = ensure_scratch(h, sizeof(float4) * h->n + 256);
    if (rc) return rc;
    rc = sph_read_positions_device(h, h->scratch);
    if (rc) return rc;
    if (h->n) CK(cudaMemcpyAsync(host_xyzw, h->scratch, sizeof(float4) * h->n, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return SPH_OK;
}

int sph_write_transforms_device(sph_handle *h, void *dev_mat4)
{
    int rc = enter_exact(h);
    if (rc) return rc;
    if (!h->have_state) return fail(h, SPH_ERR_STATE, "no particles uploaded");
    if (!dev_mat4) return fail(h, SPH_ERR_INVALID, "destination is NULL");
    if (h->n) {
        k_transforms<<<blocks_for(h->n, IO_THREADS), IO_THREADS, 0, h->stream>>>(h->pos[h->cur], (uint32_t)h->n, nullptr,
                                                                               h->P.sphere_scale, (float4 *)dev_mat4);
        CK_LAUNCH();
    }
    return SPH_OK;
}

int sph_write_transforms(sph_handle *h, float *host_mat4)
{
    int rc = enter_exact(h);
    if (rc) return rc;
    if (!host_mat4) return fail(h, SPH_ERR_INVALID, "destination is NULL");
    rc = ensure_scratch(h, 64 * h->n + 256);
    if (rc) return rc;
    rc = sph_write_transforms_device(h, h->scratch);
    if (rc) return rc;
    if (h->n) CK(cudaMemcpyAsync(host_mat4, h->scratch, 64 * h->n, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return SPH_OK;
}

// ---- reset point and device-side scenes -----------------------------------------------------------

int sph_set_reset_point(sph_handle *h)
{
    int rc = enter_exact(h);
    if (rc) return rc;
    if (!h->have_state) return fail(h, SPH_ERR_STATE, "no particles uploaded");
    if (!h->reset_pos) {
        CK(cudaMalloc(&h->reset_pos, sizeof(float4) * h->cap));
        CK(cudaMalloc(&h->reset_vel, sizeof(float4) * h->cap));
    }
    if (h->n) {
        CK(cudaMemcpyAsync(h->reset_pos, h->pos[h->cur], sizeof(float4) * h->n, cudaMemcpyDeviceToDevice, h->stream));
        CK(cudaMemcpyAsync(h->reset_vel, h->vel[h->cur], sizeof(float4) * h->n, cudaMemcpyDeviceToDevice, h->stream));
    }
    h->reset_n = h->n;
    h->have_reset = true;
    return SPH_OK;
}

int sph_reset(sph_handle *h)
{
    int rc = enter_exact(h);
    if (rc) return rc;
    if (!h->have_reset) return fail(h, SPH_ERR_STATE, "no reset point: call sph_set_reset_point after loading the initial state");
    if (h->reset_n) {
        CK(cudaMemcpyAsync(h->pos[h->cur], h->reset_pos, sizeof(float4) * h->reset_n, cudaMemcpyDeviceToDevice, h->stream));
        CK(cudaMemcpyAsync(h->vel[h->cur], h->reset_vel, sizeof(float4) * h->reset_n, cudaMemcpyDeviceToDevice, h->stream));
    }
    return after_upload(h, h->reset_n);  // rows, step count, bounding box: as after an upload (no host traffic, no sync)
}

namespace {

// glibc's rand() after srand(seed) as a linear recurrence (sph_scene.cuh): r[i] = r[i-31] + r[i-3] mod 2^32,
// output k = r[k + 344] >> 1. seed_state fills r[0..30] the way __srandom_r does.
void glibc_seed_state(unsigned seed, uint32_t r[31])
{
    int32_t word = seed ? (int32_t)seed : 1;
    r[0] = (uint32_t)word;
    for (int i = 1; i < 31; ++i) {
        const long hi = word / 127773, lo = word % 127773;
        long w = 16807 * lo - 2836 * hi;
        if (w < 0) w += 2147483647;
        word = (int32_t)w;
        r[i] = (uint32_t)word;
    }
}

// c = a * b mod (x^31 - x^28 - 1), coefficients mod 2^32.
void poly_mulmod(const uint32_t a[31], const uint32_t b[31], uint32_t c[31])
{
    uint32_t t[61] = {0};
    for (int i = 0; i < 31; ++i)
        for (int j = 0; j < 31; ++j) t[i + j] += a[i] * b[j];
    for (int k = 60; k >= 31; --k) {  // x^k = x^(k-3) + x^(k-31)
        t[k - 3] += t[k];
        t[k - 31] += t[k];
    }
    std::memcpy(c, t, sizeof(uint32_t) * 31);
}

// For every chunk q (chunk_draws outputs each, nchunks of them): the 31 words r[m .. m+30] with
// m = q * chunk_draws + 344 - 31, i.e. the window from which output q * chunk_draws is the next one produced.
void glibc_rand_jump(unsigned seed, uint64_t chunk_draws, uint64_t nchunks, std::vector<uint32_t> &out)
{
    out.resize(nchunks * 31);
    uint32_t r[31 + 344];
    glibc_seed_state(seed, r);
    for (int i = 31; i < 34; ++i) r[i] = r[i - 31];  // glibc copies the first three words, the recurrence starts at 34
    for (int i = 34; i < 31 + 344; ++i) r[i] = r[i - 31] + r[i - 3];
    // x^chunk_draws mod p by square and multiply
    uint32_t A[31] = {0}, base[31] = {0}, tmp[31];
    A[0] = 1;
    base[1] = 1;
    for (uint64_t e = chunk_draws; e; e >>= 1) {
        if (e & 1) { poly_mulmod(A, base, tmp); std::memcpy(A, tmp, sizeof tmp); }
        poly_mulmod(base, base, tmp);
        std::memcpy(base, tmp, sizeof tmp);
    }
    uint32_t win[61];
    std::memcpy(win, r + 344 - 31, sizeof(uint32_t) * 31);  // chunk 0
    for (uint64_t q = 0; q < nchunks; ++q) {
        std::memcpy(&out[q * 31], win, sizeof(uint32_t) * 31);
        if (q + 1 == nchunks) break;
        for (int i = 31; i < 61; ++i) win[i] = win[i - 31] + win[i - 3];
        uint32_t next[31];
        for (int j = 0; j < 31; ++j) {  // r[m + C + j] = sum_k A_k r[m + k + j]
            uint32_t acc = 0;
            for (int k = 0; k < 31; ++k) acc += A[k] * win[k + j];
            next[j] = acc;
        }
        std::memcpy(win, next, sizeof next);
    }
}

int scene_device(sph_handle *h, const SceneDesc &d, unsigned seed)
{
    int rc = enter_exact(h);
    if (rc) return rc;
    if (d.nx < 0 || d.ny < 0 || d.nz < 0 || d.i0 < 0 || d.i1 < d.i0 || d.i1 > d.nx) return fail(h, SPH_ERR_INVALID, "bad lattice / x-range");
    const uint64_t n_total = (uint64_t)d.nx * d.ny * d.nz, n = (uint64_t)(d.i1 - d.i0) * d.ny * d.nz;
    if (n_total >= (1ull << 31)) return fail(h, SPH_ERR_INVALID, "lattice of %llu particles: ids must stay below 2^31", (unsigned long long)n_total);
    if (n > h->cap) return fail(h, SPH_ERR_CAPACITY, "%llu particles exceed capacity %llu", (unsigned long long)n, (unsigned long long)h->cap);
    if (n_total) {
        // only the chunks up to the last produced row are needed (draws are consumed x-outer)
        const uint64_t last = (uint64_t)d.i1 * d.ny * d.nz;
        const uint64_t nchunks = (last + SCENE_CHUNK - 1) / SCENE_CHUNK;
        std::vector<uint32_t> st;
        glibc_rand_jump(seed, 3ull * SCENE_CHUNK, nchunks, st);
        rc = ensure_scratch(h, st.size() * sizeof(uint32_t) + 256);
        if (rc) return rc;
        CK(cudaMemcpyAsync(h->scratch, st.data(), st.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
        k_scene_block<<<blocks_for(nchunks, SCENE_THREADS), SCENE_THREADS, 0, h->stream>>>((const uint32_t *)h->scratch, last, d,
                                                                                          h->pos[h->cur], h->vel[h->cur]);
        CK_LAUNCH();
        CK(cudaStreamSynchronize(h->stream));  // st goes out of scope
    }
    return after_upload(h, n);
}

}  // namespace

// Host-only self-test of the jump-ahead (no CUDA call): the first `ndraws` outputs of glibc's rand() after
// srand(seed) against the same stream rebuilt chunk by chunk from the jumped states.
int sph_selftest_glibc_rand(unsigned seed, uint64_t ndraws, uint64_t chunk_draws, uint64_t *mismatches_out)
{
    if (!mismatches_out || chunk_draws == 0 || ndraws == 0 || ndraws > (1ull << 30)) return SPH_ERR_INVALID;
    const uint64_t nchunks = (ndraws + chunk_draws - 1) / chunk_draws;
    std::vector<uint32_t> st;
    glibc_rand_jump(seed, chunk_draws, nchunks, st);
    std::srand(seed);
    uint64_t bad = 0;
    for (uint64_t q = 0; q < nchunks; ++q) {
        uint32_t w[31];
        std::memcpy(w, &st[q * 31], sizeof w);
        int p = 0;
        const uint64_t m = std::min(chunk_draws, ndraws - q * chunk_draws);
        for (uint64_t t = 0; t < m; ++t) {
            const uint32_t r = w[p] + w[(p + 28) % 31];
            w[p] = r;
            p = (p + 1) % 31;
            bad += (uint32_t)std::rand() != (r >> 1);
        }
    }
    *mismatches_out = bad;
    return SPH_OK;
}

int sph_scene_cube_device(sph_handle *h, int width)
{
    if (!h) return SPH_ERR_INVALID;
    SceneDesc d{};
    d.nx = d.ny = d.nz = width;
    d.i0 = 0;
    d.i1 = width;
    d.h = h->settings.h;
    d.sep = d.h + 0.01f;  // src/SPHSystem.cpp:80
    d.x0 = d.z0 = -1.5f;
    d.y0 = d.h;           // "+ h + 0.1f" (src/SPHSystem.cpp:94): two additions, in this order
    d.y1 = 0.1f;
    d.cube = 1;
    return scene_device(h, d, 1024u);
}

int sph_scene_block_device(sph_handle *h, int nx, int ny, int nz, float sep, float x0, float y0, float z0, unsigned seed,
                           int i0, int i1)
{
    if (!h) return SPH_ERR_INVALID;
    SceneDesc d{};
    d.nx = nx; d.ny = ny; d.nz = nz;
    d.i0 = i0; d.i1 = i1;
    d.h = h->settings.h;
    d.sep = sep; d.x0 = x0; d.y0 = y0; d.z0 = z0;
    return scene_device(h, d, seed);
}

// ---- the step ---------------------------------------------------------------------------------

int sph_step(sph_handle *h, float dt, int nsteps)
{
    int rc = enter_exact(h);
    if (rc) return rc;
    if (!h->have_state) return fail(h, SPH_ERR_STATE, "sph_step before sph_upload");
    if (nsteps < 0) return fail(h, SPH_ERR_INVALID, "nsteps must be >= 0");
    if (!(dt > 0.f)) dt = h->settings.dt;  // SPHSystem::update's fixed step (src/SPHSystem.cpp:113)
    if (h->n == 0) return SPH_OK;
    if (h->graph_enabled && !h->timing && !h->slab_mode) return step_graphs(h, dt, nsteps);
    for (int k = 0; k < nsteps; ++k) {
        rc = step_once(h, dt);
        if (rc) return rc;
    }
    return SPH_OK;
}

// Neighbour search only (cell hash, counting sort by cell, cell start offsets) for the current
// positions: BASELINE.json config 4. Rows end up cell-sorted; no physics is run.
int sph_neighbor_search(sph_handle *h, int repeats)
{
    int rc = enter_exact(h);
    if (rc) return rc;
    if (!h->have_state) return fail(h, SPH_ERR_STATE, "no particles uploaded");
    if (h->n == 0) return SPH_OK;
    for (int k = 0; k < repeats; ++k) {
        // the box of the current positions is in bbox[parity]; the plan re-arms the other slot only
        rc = build_grid(h);
        if (rc) return rc;
    }
    h->have_step = false;
    return SPH_OK;
}

int sph_sync(sph_handle *h)
{
    int rc = enter_exact(h);
    if (rc) return rc;
    CK(cudaStreamSynchronize(h->stream));
    return SPH_OK;
}

int sph_update_particles_aos(sph_handle *h, void *host_particles, float *host_mat4, uint64_t n, float dt)
{
    int rc = enter_exact(h);
    if (rc) return rc;
    if (n > h->cap) return fail(h, SPH_ERR_CAPACITY, "%llu particles exceed capacity %llu", (unsigned long long)n,
                                (unsigned long long)h->cap);
    if (n && !host_particles) return fail(h, SPH_ERR_INVALID, "particles is NULL");
    if (!(dt > 0.f)) dt = h->settings.dt;
    if (n == 0) { h->n = 0; h->have_state = true; h->have_step = false; return SPH_OK; }
    const size_t ba = align_up((size_t)60 * n, 256), bm = align_up((size_t)64 * n, 256);
    rc = ensure_scratch(h, 2 * ba + bm + 256);
    if (rc) return rc;
    char *sc = (char *)h->scratch;
    uint32_t *aos_in = (uint32_t *)sc, *aos_out = (uint32_t *)(sc + ba);
    float4 *mats = (float4 *)(sc + 2 * ba);
    cudaStream_t s = h->stream;
    CK(cudaMemcpyAsync(aos_in, host_particles, (size_t)60 * n, cudaMemcpyHostToDevice, s));
    k_import_aos<<<blocks_for(n, IO_THREADS), IO_THREADS, 0, s>>>(aos_in, (uint32_t)n, h->pos[h->cur], h->vel[h->cur]);
    CK_LAUNCH();
    rc = after_upload(h, n);
    if (rc) return rc;
    h->write_force = true;  // Particle::force is part of this call's output
    rc = step_once(h, dt);
    h->write_force = false;
    if (rc) return rc;
    rc = build_hash16_order(h, true);
    if (rc) return rc;
    k_export_aos<<<blocks_for(n, IO_THREADS), IO_THREADS, 0, s>>>(h->pos[h->cur], h->vel[h->cur], h->force, h->hash16,
                                                                (uint32_t)n, h->map, h->P, aos_in, aos_out);
    CK_LAUNCH();
    CK(cudaMemcpyAsync(host_particles, aos_out, (size_t)60 * n, cudaMemcpyDeviceToHost, s));
    if (host_mat4) {
        k_transforms<<<blocks_for(n, IO_THREADS), IO_THREADS, 0, s>>>(h->pos[h->cur], (uint32_t)n, h->map,
                                                                    h->P.sphere_scale, mats);
        CK_LAUNCH();
        CK(cudaMemcpyAsync(host_mat4, mats, (size_t)64 * n, cudaMemcpyDeviceToHost, s));
    }
    CK(cudaStreamSynchronize(s));
    return SPH_OK;
}

// ---- parity / diagnostics ---------------------------------------------------------------------

int sph_hash_table(sph_handle *h, uint32_t *host_table)
{
    int rc = enter_exact(h);
    if (rc) return rc;
    if (!h->have_step) return fail(h, SPH_ERR_STATE, "the hash table is only defined after a step");
    if (!host_table) return fail(h, SPH_ERR_INVALID, "destination is NULL");
    rc = build_hash16_order(h, false);
    if (rc) return rc;
    rc = ensure_scratch(h, sizeof(uint32_t) * REF_TABLE_SIZE);
    if (rc) return rc;
    k_ref_table<<<blocks_for(REF_TABLE_SIZE, IO_THREADS), IO_THREADS, 0, h->stream>>>(h->h16_cells, (uint32_t *)h->scratch);
    CK_LAUNCH();
    CK(cudaMemcpyAsync(host_table, h->scratch, sizeof(uint32_t) * REF_TABLE_SIZE, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return SPH_OK;
}

int sph_neighbor_lists(sph_handle *h, uint32_t *host_counts, uint64_t *host_offsets, uint32_t *host_list,
                       uint64_t list_capacity, uint32_t *host_ids_out)
{
    int rc = enter_exact(h);
    if (rc) return rc;
    if (!h->have_state) return fail(h, SPH_ERR_STATE, "no particles uploaded");
    if (!host_counts) return fail(h, SPH_ERR_INVALID, "counts is NULL");
    if (host_list && !host_offsets) return fail(h, SPH_ERR_INVALID, "offsets is required with list");
    const uint64_t n = h->n;
    if (n == 0) { if (host_offsets) host_offsets[0] = 0; return SPH_OK; }
    // Rebuild the grid for the current positions. Rows get re-sorted, so the previous step's
    // force/density rows no longer line up with pos/vel.
    h->have_step = false;
    rc = build_grid(h);
    if (rc) return rc;
    // The density pass writes the lists the force pass consumes; report exactly those.
    rc = launch_density(h, (uint32_t)n);
    if (rc) return rc;
    uint32_t *dcounts = reinterpret_cast<uint32_t *>(h->slot);  // free after build_grid
    k_neighbor_lists<<<blocks_for(n, PHYS_THREADS), PHYS_THREADS, 0, h->stream>>>(h->pos[h->cur], h->vel[h->cur], (uint32_t)n,
                                                                                h->gd, h->cells, h->P, h->nlist, h->ncount,
                                                                                (uint32_t)h->cap, dcounts, nullptr, nullptr);
    CK_LAUNCH();
    CK(cudaMemcpyAsync(host_counts, dcounts, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, h->stream));
    if (host_ids_out) {
        rc = ensure_scratch(h, sizeof(uint32_t) * n + 256);
        if (rc) return rc;
        ExportPtrs out{};
        out.id = (uint32_t *)h->scratch;
        k_export<<<blocks_for(n, IO_THREADS), IO_THREADS, 0, h->stream>>>(h->pos[h->cur], h->vel[h->cur], h->force, h->hash16,
                                                                        (uint32_t)n, nullptr, h->P, out);
        CK_LAUNCH();
        CK(cudaMemcpyAsync(host_ids_out, out.id, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, h->stream));
    }
    CK(cudaStreamSynchronize(h->stream));
    if (!host_offsets) return SPH_OK;
    uint64_t total = 0;
    for (uint64_t i = 0; i < n; ++i) { host_offsets[i] = total; total += host_counts[i]; }
    host_offsets[n] = total;
    if (!host_list) return SPH_OK;
    if (total > list_capacity)
        return fail(h, SPH_ERR_CAPACITY, "neighbour list needs %llu entries, capacity is %llu",
                    (unsigned long long)total, (unsigned long long)list_capacity);
    if (total == 0) return SPH_OK;
    const size_t bo = align_up(sizeof(uint64_t) * (n + 1), 256);
    rc = ensure_scratch(h, bo + sizeof(uint32_t) * total + 256);
    if (rc) return rc;
    unsigned long long *doff = (unsigned long long *)h->scratch;
    uint32_t *dlist = (uint32_t *)((char *)h->scratch + bo);
    CK(cudaMemcpyAsync(doff, host_offsets, sizeof(uint64_t) * (n + 1), cudaMemcpyHostToDevice, h->stream));
    k_neighbor_lists<<<blocks_for(n, PHYS_THREADS), PHYS_THREADS, 0, h->stream>>>(h->pos[h->cur], h->vel[h->cur], (uint32_t)n,
                                                                                h->gd, h->cells, h->P, h->nlist, h->ncount,
                                                                                (uint32_t)h->cap, dcounts, doff, dlist);
    CK_LAUNCH();
    CK(cudaMemcpyAsync(host_list, dlist, sizeof(uint32_t) * total, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return SPH_OK;
}

int sph_get_stats(sph_handle *h, sph_stats *out)
{
    int rc = enter_exact(h);
    if (rc) return rc;
    if (!out) return fail(h, SPH_ERR_INVALID, "out is NULL");
    std::memset(out, 0, sizeof *out);
    out->count = h->n - h->n_ghost;
    out->steps = h->steps;
    if (!h->have_state || h->n == 0) return SPH_OK;
    GridDesc g{};
    StepCounters c{};
    StatsAccum a{};
    CK(cudaMemsetAsync(h->stats_acc, 0, sizeof(StatsAccum), h->stream));
    if (h->have_step) {
        k_stats<<<h->num_sms * 4, IO_THREADS, 0, h->stream>>>(h->pos[h->cur], h->vel[h->cur], (uint32_t)h->n, h->stats_acc);
        CK_LAUNCH();
    }
    CK(cudaMemcpyAsync(&a, h->stats_acc, sizeof a, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(&g, h->gd, sizeof g, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(&c, h->ctr, sizeof c, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    out->grid_origin[0] = g.ox; out->grid_origin[1] = g.oy; out->grid_origin[2] = g.oz;
    out->grid_dim[0] = g.nx; out->grid_dim[1] = g.ny; out->grid_dim[2] = g.nz;
    out->grid_cells = g.ncells;
    out->clamped = c.clamped;
    out->deferred_density = c.heavy[0];                    // every deferred row, clump rows (served a tile at a time) included
    out->deferred_forces = c.heavy[1] + c.clump_rows[1];   // rows the force pass deferred + the clump rows it served by tile
    out->nlist_rows = NLIST_ROWS;
    out->nan_count = a.nan_count;
    if (h->have_step) out->count = a.owned;
    out->mean_density = a.sum_rho / (double)(a.owned ? a.owned : 1);
    float mx;
    std::memcpy(&mx, &a.max_rho_bits, 4);
    out->max_density = mx;
    out->kinetic_energy = 0.5 * (double)h->settings.mass * a.sum_v2;
    return SPH_OK;
}

int sph_candidate_count(sph_handle *h, uint64_t *candidates_out, uint64_t *rows_out)
{
    int rc = enter_exact(h);
    if (rc) return rc;
    if (!candidates_out || !rows_out) return fail(h, SPH_ERR_INVALID, "NULL argument");
    if (!h->have_step) return fail(h, SPH_ERR_STATE, "candidates are those of the last step: take a step first");
    rc = ensure_scratch(h, 256);
    if (rc) return rc;
    unsigned long long *d = (unsigned long long *)h->scratch;
    CK(cudaMemsetAsync(d, 0, 2 * sizeof(unsigned long long), h->stream));
    // the start-of-step rows (the cell order h->cells describes) are in the non-current buffers after a step
    k_candidate_count<<<blocks_for(h->n, PHYS_THREADS), PHYS_THREADS, 0, h->stream>>>(h->pos[h->cur ^ 1], (uint32_t)h->n, h->gd,
                                                                                     h->cells, h->P.h, d);
    CK_LAUNCH();
    unsigned long long out[2] = {0, 0};
    CK(cudaMemcpyAsync(out, d, sizeof out, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    *candidates_out = out[0];
    *rows_out = out[1];
    return SPH_OK;
}

int sph_enable_pass_timing(sph_handle *h, int enable)
{
    if (!h) return SPH_ERR_INVALID;
    h->timing = enable != 0;
    h->ev_used = 0;
    return SPH_OK;
}

int sph_pass_times(sph_handle *h, float *ms4, uint64_t *steps_out)
{
    int rc = enter(h);
    if (rc) return rc;
    if (!ms4) return fail(h, SPH_ERR_INVALID, "destination is NULL");
    if (h->ev_used == 0)
        return fail(h, SPH_ERR_STATE, "no timed step: call sph_enable_pass_timing(h, 1) and sph_step first");
    CK(cudaStreamSynchronize(h->stream));
    double acc[4] = {0, 0, 0, 0};
    for (size_t i = 0; i < h->ev_used; ++i)
        for (int k = 0; k < 4; ++k) {
            float ms = 0.f;
            CK(cudaEventElapsedTime(&ms, h->ev_pool[i].e[k], h->ev_pool[i].e[k + 1]));
            acc[k] += ms;
        }
    for (int k = 0; k < 4; ++k) ms4[k] = (float)(acc[k] / (double)h->ev_used);
    if (steps_out) *steps_out = h->ev_used;
    h->ev_used = 0;
    return SPH_OK;
}

uint64_t sph_launch_count(const sph_handle *h) { return h ? h->launches : 0; }

int sph_selftest_division(sph_handle *h, uint64_t n, uint64_t seed, uint64_t *mismatches_out)
{
    int rc = enter(h);
    if (rc) return rc;
    if (!mismatches_out || n == 0 || n > (1ull << 28)) return fail(h, SPH_ERR_INVALID, "bad arguments");
    // Log-uniform magnitudes over the ranges the force pass sees (and well beyond), both signs.
    std::vector<float> a(n), d(n);
    uint64_t x = seed * 6364136223846793005ull + 1442695040888963407ull;
    auto next = [&]() { x = x * 6364136223846793005ull + 1442695040888963407ull; return (double)(x >> 11) / 9007199254740992.0; };
    for (uint64_t i = 0; i < n; ++i) {
        const double ea = -20.0 + 40.0 * next(), ed = -12.0 + 24.0 * next();
        a[i] = (float)((next() < 0.5 ? -1.0 : 1.0) * std::pow(10.0, ea));
        d[i] = (float)((next() < 0.5 ? -1.0 : 1.0) * std::pow(10.0, ed));
    }
    const size_t bytes = align_up(sizeof(float) * n, 256);
    rc = ensure_scratch(h, 2 * bytes + 256);
    if (rc) return rc;
    float *da = (float *)h->scratch, *dd = (float *)((char *)h->scratch + bytes);
    uint32_t *dout = (uint32_t *)((char *)h->scratch + 2 * bytes);
    CK(cudaMemcpyAsync(da, a.data(), sizeof(float) * n, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(dd, d.data(), sizeof(float) * n, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemsetAsync(dout, 0, sizeof(uint32_t), h->stream));
    k_selftest_div<<<blocks_for(n, 256), 256, 0, h->stream>>>(da, dd, (uint32_t)n, dout);
    CK_LAUNCH();
    uint32_t bad = 0;
    CK(cudaMemcpyAsync(&bad, dout, sizeof bad, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    *mismatches_out = bad;
    return SPH_OK;
}

int sph_selftest_packed_dist2(sph_handle *h, uint64_t n, uint64_t seed, uint64_t *mismatches_out)
{
    int rc = enter(h);
    if (rc) return rc;
    if (!mismatches_out || n == 0 || n > (1ull << 26)) return fail(h, SPH_ERR_INVALID, "bad arguments");
    // Positions as the candidate loops see them: two rows and a candidate a fraction of a cell apart,
    // anywhere in the box (magnitudes up to 16), plus exact coincidences and far-apart triples.
    std::vector<float> v(12 * n);
    uint64_t x = seed * 6364136223846793005ull + 1442695040888963407ull;
    auto next = [&]() { x = x * 6364136223846793005ull + 1442695040888963407ull; return (double)(x >> 11) / 9007199254740992.0; };
    for (uint64_t i = 0; i < n; ++i) {
        const double scale = std::pow(10.0, -3.0 + 3.0 * next());  // separation scale 1e-3 .. 1
        float *a = &v[4 * i], *b = &v[4 * (n + i)], *c = &v[4 * (2 * n + i)];
        for (int k = 0; k < 3; ++k) {
            const double base = -16.0 + 32.0 * next();
            a[k] = (float)base;
            b[k] = (float)(base + scale * (next() - 0.5));
            c[k] = (i % 97 == 0) ? a[k] : (float)(base + scale * (next() - 0.5));
        }
        a[3] = b[3] = c[3] = 0.f;
    }
    const size_t bytes = align_up(sizeof(float) * 4 * n, 256);
    rc = ensure_scratch(h, 3 * bytes + 256);
    if (rc) return rc;
    char *sc = (char *)h->scratch;
    uint32_t *dout = (uint32_t *)(sc + 3 * bytes);
    for (int k = 0; k < 3; ++k)
        CK(cudaMemcpyAsync(sc + k * bytes, v.data() + 4 * n * k, sizeof(float) * 4 * n, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemsetAsync(dout, 0, sizeof(uint32_t), h->stream));
    k_selftest_packed_dist2<<<blocks_for(n, 256), 256, 0, h->stream>>>((const float4 *)sc, (const float4 *)(sc + bytes),
                                                                    (const float4 *)(sc + 2 * bytes), (uint32_t)n, dout);
    CK_LAUNCH();
    uint32_t bad = 0;
    CK(cudaMemcpyAsync(&bad, dout, sizeof bad, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    *mismatches_out = bad;
    return SPH_OK;
}

// ---- slab decomposition ---------------------------------------------------------------------------

namespace {
int make_cuts(sph_handle *h, const int32_t *cuts, int world, SlabCuts &c)
{
    if (!cuts || world < 1 || world > SLAB_MAX_RANKS)
        return fail(h, SPH_ERR_INVALID, "world must be in [1, %d] and cuts non-NULL", SLAB_MAX_RANKS);
    c.world = world;
    for (int k = 0; k <= world; ++k) c.lo[k] = cuts[k];
    for (int k = 1; k < world; ++k)
        if (cuts[k] > cuts[k + 1] && k + 1 < world)
            return fail(h, SPH_ERR_INVALID, "cuts must be non-decreasing");
    return SPH_OK;
}
}  // namespace

int sph_slab_enable(sph_handle *h, int enable)
{
    int rc = enter(h);
    if (rc) return rc;
    h->slab_mode = enable != 0;
    return SPH_OK;
}

uint64_t sph_slab_owned(const sph_handle *h) { return h ? h->n - h->n_ghost : 0; }

int sph_slab_count(sph_handle *h, const int32_t *cuts, int world, uint64_t *host_counts)
{
    int rc = enter_exact(h);
    if (rc) return rc;
    if (!h->have_state) return fail(h, SPH_ERR_STATE, "no particles uploaded");
    if (!host_counts) return fail(h, SPH_ERR_INVALID, "counts is NULL");
    SlabCuts c;
    rc = make_cuts(h, cuts, world, c);
    if (rc) return rc;
    CK(cudaMemsetAsync(h->slab_counts, 0, sizeof(unsigned long long) * SLAB_MAX_RANKS, h->stream));
    if (h->n) {
        k_slab_count<<<blocks_for(h->n, SLAB_THREADS), SLAB_THREADS, 0, h->stream>>>(h->pos[h->cur], (uint32_t)h->n, h->P.h, c,
                                                                                   h->slab_counts);
        CK_LAUNCH();
    }
    unsigned long long tmp[SLAB_MAX_RANKS];
    CK(cudaMemcpyAsync(tmp, h->slab_counts, sizeof(unsigned long long) * world, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    for (int k = 0; k < world; ++k) host_counts[k] = tmp[k];
    return SPH_OK;
}

int sph_slab_pack(sph_handle *h, const int32_t *cuts, int world, int self, void *dev_buf, const uint64_t *row_offsets)
{
    int rc = enter_exact(h);
    if (rc) return rc;
    if (!h->have_state) return fail(h, SPH_ERR_STATE, "no particles uploaded");
    if (self < 0 || self >= world || !row_offsets) return fail(h, SPH_ERR_INVALID, "bad self / offsets");
    SlabCuts c;
    rc = make_cuts(h, cuts, world, c);
    if (rc) return rc;
    SlabOffsets off{};
    for (int k = 0; k < world; ++k) off.row[k] = row_offsets[k];
    unsigned long long *cursors = h->slab_counts + SLAB_MAX_RANKS;
    CK(cudaMemsetAsync(cursors, 0, sizeof(unsigned long long) * SLAB_MAX_RANKS, h->stream));
    if (h->n) {
        k_slab_pack<<<blocks_for(h->n, SLAB_THREADS), SLAB_THREADS, 0, h->stream>>>(
            h->pos[h->cur], h->vel[h->cur], (uint32_t)h->n, h->P.h, c, self, off, cursors, (float4 *)dev_buf);
        CK_LAUNCH();
    }
    // a general step starts from a clean violation word (the sync-free steps only ever set bits in it)
    CK(cudaMemsetAsync(&h->ctr->aux[3], 0, sizeof(uint32_t), h->stream));
    h->n_ghost = 0;  // last step's ghosts are dropped rows now
    h->ghost_n[0] = h->ghost_n[1] = 0;
    h->have_step = false;
    h->slab_fast = false;
    h->slab_lo = h->slab_hi = 0;
    h->edge_ok = false;
    return SPH_OK;
}

int sph_slab_append(sph_handle *h, const void *dev_rows, uint64_t nrows, int kind)
{
    int rc = enter(h);
    if (rc) return rc;
    if (kind < 0 || kind > 2) return fail(h, SPH_ERR_INVALID, "kind must be 0 (owned), 1 or 2 (ghost batch)");
    if (h->n + nrows > h->cap)
        return fail(h, SPH_ERR_CAPACITY, "append of %llu rows to %llu exceeds capacity %llu", (unsigned long long)nrows,
                    (unsigned long long)h->n, (unsigned long long)h->cap);
    if (nrows) {
        if (!dev_rows) return fail(h, SPH_ERR_INVALID, "rows is NULL");
        k_slab_append<<<blocks_for(nrows, SLAB_THREADS), SLAB_THREADS, 0, h->stream>>>(
            (const float4 *)dev_rows, (uint32_t)nrows, (uint32_t)h->n, kind != 0, h->pos[h->cur], h->vel[h->cur]);
        CK_LAUNCH();
    }
    if (kind) {
        h->ghost_first[kind - 1] = h->n;
        h->ghost_n[kind - 1] = nrows;
        h->n_ghost += nrows;
    }
    h->n += nrows;
    h->have_state = true;
    h->have_step = false;
    h->edge_ok = false;
    return SPH_OK;
}

int sph_slab_pack_halo(sph_handle *h, int32_t cell_x, int side, void *dev_buf, uint64_t capacity_rows, uint64_t *nrows_out)
{
    int rc = enter_exact(h);
    if (rc) return rc;
    if (side < 0 || side > 1 || !nrows_out) return fail(h, SPH_ERR_INVALID, "bad side / nrows_out");
    unsigned long long *cursor = h->slab_counts + 2 * SLAB_MAX_RANKS + side;
    h->p2p_clean = false;
    CK(cudaMemsetAsync(cursor, 0, sizeof(unsigned long long), h->stream));
    if (h->n) {
        k_slab_pack_halo<<<blocks_for(h->n, SLAB_THREADS), SLAB_THREADS, 0, h->stream>>>(
            h->pos[h->cur], h->vel[h->cur], (uint32_t)h->n, h->P.h, cell_x, cursor,
            (uint32_t)(capacity_rows > h->cap ? h->cap : capacity_rows), (float4 *)dev_buf, h->halo_rows[side]);
        CK_LAUNCH();
    }
    unsigned long long cnt = 0;
    CK(cudaMemcpyAsync(&cnt, cursor, sizeof cnt, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (cnt > capacity_rows || cnt > h->cap)
        return fail(h, SPH_ERR_CAPACITY, "halo layer has %llu rows, buffer holds %llu", cnt, (unsigned long long)capacity_rows);
    h->halo_n[side] = cnt;
    *nrows_out = cnt;
    return SPH_OK;
}

int sph_slab_step_density(sph_handle *h)
{
    int rc = enter(h);
    if (rc) return rc;
    if (!h->have_state) return fail(h, SPH_ERR_STATE, "no particles uploaded");
    if (!h->slab_mode) return fail(h, SPH_ERR_STATE, "call sph_slab_enable(h, 1) first");
    if (h->n == 0) return SPH_OK;
    if (h->slab_fast && h->have_bbox_from_integration) {
        // Sync-free paths: keep the box the last integration produced for the owned rows and widen it
        // by one cell for the ghost layers and arrivals (anything further out is clamped, which is safe).
        h->bbox_expand = 1;
    } else {
        rc = compute_bbox(h);  // arrivals and new ghosts are not covered by the integration's box
        if (rc) return rc;
    }
    rc = build_grid(h);
    h->bbox_expand = 0;
    if (rc) return rc;
    const uint32_t n = (uint32_t)h->n;
    if (n) {
        rc = launch_density(h, n);
        if (rc) return rc;
    }
    return SPH_OK;
}

int sph_slab_pack_halo_density(sph_handle *h, int side, void *dev_buf)
{
    int rc = enter(h);
    if (rc) return rc;
    if (side < 0 || side > 1) return fail(h, SPH_ERR_INVALID, "bad side");
    const uint64_t n = h->halo_n[side];
    if (n) {
        if (!dev_buf) return fail(h, SPH_ERR_INVALID, "buffer is NULL");
        k_slab_pack_density<<<blocks_for(n, SLAB_THREADS), SLAB_THREADS, 0, h->stream>>>(h->vel[h->cur], h->inverse, h->halo_rows[side],
                                                                                       (uint32_t)n, (float *)dev_buf);
        CK_LAUNCH();
    }
    return SPH_OK;
}

int sph_slab_set_ghost_density(sph_handle *h, int side, const void *dev_buf, uint64_t nrows)
{
    int rc = enter(h);
    if (rc) return rc;
    if (side < 0 || side > 1) return fail(h, SPH_ERR_INVALID, "bad side");
    if (nrows != h->ghost_n[side])
        return fail(h, SPH_ERR_INVALID, "ghost batch %d has %llu rows, got %llu densities", side,
                    (unsigned long long)h->ghost_n[side], (unsigned long long)nrows);
    if (nrows) {
        k_slab_set_ghost_density<<<blocks_for(nrows, SLAB_THREADS), SLAB_THREADS, 0, h->stream>>>(
            h->vel[h->cur], h->inverse, (uint32_t)h->ghost_first[side], (uint32_t)nrows, (const float *)dev_buf);
        CK_LAUNCH();
    }
    return SPH_OK;
}

int sph_slab_step_forces(sph_handle *h, float dt)
{
    int rc = enter(h);
    if (rc) return rc;
    if (!h->have_state) return fail(h, SPH_ERR_STATE, "no particles uploaded");
    if (!(dt > 0.f)) dt = h->settings.dt;
    const uint32_t n = (uint32_t)h->n;
    if (n) {
        if (h->rho_pending) {
            // the halo densities are still on their way (second stream): rows without ghost neighbours first
            rc = launch_forces_integrate(h, n, dt, FI_STEP, 1);
            if (rc) return rc;
            CK(cudaStreamWaitEvent(h->stream, h->ev_rho, 0));
            h->rho_pending = false;
            rc = launch_forces_integrate(h, n, dt, FI_STEP, 2);
        } else {
            rc = launch_forces_integrate(h, n, dt, FI_STEP);
        }
        if (rc) return rc;
        h->parity ^= 1;  // the integration accumulated the next step's box into the other slot
        h->have_bbox_from_integration = true;
    }
    ++h->steps;
    h->have_step = true;
    h->edge_ok = true;
    return SPH_OK;
}

int sph_slab_download_owned(sph_handle *h, float *host_pos_xyz, float *host_vel_xyz, uint32_t *host_id,
                            uint64_t capacity_rows, uint64_t *count_out)
{
    int rc = enter_exact(h);
    if (rc) return rc;
    if (!h->have_state) return fail(h, SPH_ERR_STATE, "no particles uploaded");
    if (!host_pos_xyz || !host_vel_xyz || !host_id || !count_out) return fail(h, SPH_ERR_INVALID, "NULL argument");
    const uint64_t n = h->n;
    const uint64_t cap = capacity_rows < n ? capacity_rows : n;
    const size_t b3 = align_up(sizeof(float) * 3 * (cap ? cap : 1), 256), b1 = align_up(sizeof(uint32_t) * (cap ? cap : 1), 256);
    rc = ensure_scratch(h, 2 * b3 + b1 + 256);
    if (rc) return rc;
    char *sc = (char *)h->scratch;
    unsigned long long *cursor = h->slab_counts + 2 * SLAB_MAX_RANKS + 6;
    CK(cudaMemsetAsync(cursor, 0, sizeof(unsigned long long), h->stream));
    if (n) {
        k_slab_export_owned<<<blocks_for(n, SLAB_THREADS), SLAB_THREADS, 0, h->stream>>>(
            h->pos[h->cur], h->vel[h->cur], (uint32_t)n, (uint32_t)cap, cursor, (float *)sc, (float *)(sc + b3),
            (uint32_t *)(sc + 2 * b3));
        CK_LAUNCH();
    }
    unsigned long long cnt = 0;
    CK(cudaMemcpyAsync(&cnt, cursor, sizeof cnt, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (cnt > capacity_rows)
        return fail(h, SPH_ERR_CAPACITY, "%llu owned rows, buffers hold %llu", cnt, (unsigned long long)capacity_rows);
    if (cnt) {
        CK(cudaMemcpyAsync(host_pos_xyz, sc, sizeof(float) * 3 * cnt, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaMemcpyAsync(host_vel_xyz, sc + b3, sizeof(float) * 3 * cnt, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaMemcpyAsync(host_id, sc + 2 * b3, sizeof(uint32_t) * cnt, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
    }
    *count_out = cnt;
    return SPH_OK;
}

int sph_slab_xcell_histogram(sph_handle *h, int32_t x_cell_lo, uint32_t nbins, uint64_t *host_hist)
{
    int rc = enter_exact(h);
    if (rc) return rc;
    if (!host_hist || nbins == 0 || nbins > (1u << 24)) return fail(h, SPH_ERR_INVALID, "bad histogram arguments");
    rc = ensure_scratch(h, sizeof(unsigned long long) * nbins + 256);
    if (rc) return rc;
    unsigned long long *d = (unsigned long long *)h->scratch;
    CK(cudaMemsetAsync(d, 0, sizeof(unsigned long long) * nbins, h->stream));
    if (h->n) {
        k_slab_xhist<<<blocks_for(h->n, SLAB_THREADS), SLAB_THREADS, 0, h->stream>>>(h->pos[h->cur], (uint32_t)h->n, h->P.h,
                                                                                   x_cell_lo, nbins, d);
        CK_LAUNCH();
    }
    CK(cudaMemcpyAsync(host_hist, d, sizeof(unsigned long long) * nbins, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return SPH_OK;
}

// ---- sync-free slab path: fixed-size messages to the adjacent ranks, no host synchronisation ----

// Start of a sync-free slab step: decide whether its two scans may be confined to the edge layers.
static void begin_edge_scans(sph_handle *h)
{
    h->edge_all = !h->edge_ok || !h->edge_scan_enabled;
    h->edge_sorted = h->n;
    h->edge_ok = false;  // consumed: rows get dropped and appended from here on
}

// Grid of the edge scans (grid-stride kernels): enough blocks for every row when the scan is not confined.
static unsigned edge_blocks(const sph_handle *h)
{
    const unsigned full = blocks_for(h->n, SLAB_THREADS);
    return h->edge_all ? full : std::min(full, (unsigned)h->num_sms * 8u);
}

int sph_slab_fast_begin(sph_handle *h, int32_t lo, int32_t hi, int32_t lo_prev, int32_t hi_next, uint64_t cap_rows,
                        void *dev_send_left, void *dev_send_right)
{
    int rc = enter_exact(h);  // the only wait of the step: the row count of the PREVIOUS step's build
    if (rc) return rc;
    if (!h->have_state) return fail(h, SPH_ERR_STATE, "no particles uploaded");
    if (!h->slab_mode) return fail(h, SPH_ERR_STATE, "call sph_slab_enable(h, 1) first");
    if (cap_rows == 0 || cap_rows > h->cap) return fail(h, SPH_ERR_INVALID, "bad message capacity");
    cudaStream_t s = h->stream;
    unsigned long long *cur = h->slab_counts + 2 * SLAB_MAX_RANKS;  // [0,1] migrants, [2,3] halos
    h->p2p_clean = false;
    CK(cudaMemsetAsync(cur, 0, 4 * sizeof(unsigned long long), s));
    CK(cudaMemsetAsync(&h->ctr->aux[3], 0, sizeof(uint32_t), s));
    if (dev_send_left) CK(cudaMemsetAsync(dev_send_left, 0xFF, cap_rows * 32, s));
    if (dev_send_right) CK(cudaMemsetAsync(dev_send_right, 0xFF, cap_rows * 32, s));
    begin_edge_scans(h);
    if (h->n) {
        k_slab_fast_begin<<<edge_blocks(h), SLAB_THREADS, 0, s>>>(
            h->pos[h->cur], h->vel[h->cur], (uint32_t)h->n, h->P.h, dev_send_left ? lo : -0x7fffffff - 1,
            dev_send_right ? hi : 0x7fffffff, lo_prev, hi_next, (uint32_t)cap_rows, (float4 *)dev_send_left,
            (float4 *)dev_send_right, cur, &h->ctr->aux[3], h->gd, h->cells, h->ctr, h->edge_all);
        CK_LAUNCH();
    }
    h->slab_fast = true;
    h->n_ghost = 0;
    h->ghost_n[0] = h->ghost_n[1] = 0;
    h->have_step = false;
    h->slab_lo = h->slab_hi = 0;
    ++h->launches;
    return SPH_OK;
}

// Append two fixed-size row messages (either may be NULL) after the current rows.
static int fast_append(sph_handle *h, const void *left, const void *right, uint64_t cap_rows, bool ghost)
{
    const void *src[2] = {left, right};
    for (int side = 0; side < 2; ++side) {
        if (ghost) { h->ghost_first[side] = h->n; h->ghost_n[side] = src[side] ? cap_rows : 0; }
        if (!src[side]) continue;
        if (h->n + cap_rows > h->cap)
            return fail(h, SPH_ERR_CAPACITY, "appending a %llu-row message to %llu rows exceeds capacity %llu",
                        (unsigned long long)cap_rows, (unsigned long long)h->n, (unsigned long long)h->cap);
        k_slab_append<<<blocks_for(cap_rows, SLAB_THREADS), SLAB_THREADS, 0, h->stream>>>(
            (const float4 *)src[side], (uint32_t)cap_rows, (uint32_t)h->n, ghost, h->pos[h->cur], h->vel[h->cur]);
        CK_LAUNCH();
        h->n += cap_rows;
        ++h->launches;
    }
    return SPH_OK;
}

int sph_slab_fast_arrivals(sph_handle *h, const void *dev_recv_left, const void *dev_recv_right, uint64_t cap_rows)
{
    int rc = enter(h);
    if (rc) return rc;
    return fast_append(h, dev_recv_left, dev_recv_right, cap_rows, false);
}

int sph_slab_fast_halo(sph_handle *h, int32_t lo, int32_t hi, uint64_t cap_rows, void *dev_send_left, void *dev_send_right)
{
    int rc = enter(h);
    if (rc) return rc;
    if (cap_rows == 0 || cap_rows > h->cap) return fail(h, SPH_ERR_INVALID, "bad message capacity");
    cudaStream_t s = h->stream;
    if (dev_send_left) CK(cudaMemsetAsync(dev_send_left, 0xFF, cap_rows * 32, s));
    if (dev_send_right) CK(cudaMemsetAsync(dev_send_right, 0xFF, cap_rows * 32, s));
    if (h->n) {
        k_slab_fast_halo<<<edge_blocks(h), SLAB_THREADS, 0, s>>>(
            h->pos[h->cur], h->vel[h->cur], (uint32_t)h->n, h->P.h, lo, hi, dev_send_left != nullptr,
            dev_send_right != nullptr, (uint32_t)cap_rows, (float4 *)dev_send_left, (float4 *)dev_send_right,
            h->halo_rows[0], h->halo_rows[1], h->slab_counts + 2 * SLAB_MAX_RANKS, &h->ctr->aux[3], h->gd, h->cells,
            h->ctr, (uint32_t)h->edge_sorted, h->edge_all);
        CK_LAUNCH();
    }
    h->fast_halo_cap = cap_rows;
    ++h->launches;
    return SPH_OK;
}

int sph_slab_fast_ghosts(sph_handle *h, const void *dev_recv_left, const void *dev_recv_right, uint64_t cap_rows)
{
    int rc = enter(h);
    if (rc) return rc;
    return fast_append(h, dev_recv_left, dev_recv_right, cap_rows, true);
}

int sph_slab_fast_pack_density(sph_handle *h, uint64_t cap_rows, void *dev_send_left, void *dev_send_right)
{
    int rc = enter(h);
    if (rc) return rc;
    if (cap_rows != h->fast_halo_cap) return fail(h, SPH_ERR_INVALID, "capacity differs from the halo messages'");
    void *dst[2] = {dev_send_left, dev_send_right};
    for (int side = 0; side < 2; ++side) {
        if (!dst[side]) continue;
        CK(cudaMemsetAsync(dst[side], 0xFF, cap_rows * sizeof(float), h->stream));
        k_slab_fast_pack_density<<<blocks_for(cap_rows, SLAB_THREADS), SLAB_THREADS, 0, h->stream>>>(
            h->vel[h->cur], h->inverse, h->halo_rows[side], h->slab_counts + 2 * SLAB_MAX_RANKS + 2 + side,
            (uint32_t)cap_rows, (float *)dst[side]);
        CK_LAUNCH();
        ++h->launches;
    }
    return SPH_OK;
}

int sph_slab_fast_set_ghost_density(sph_handle *h, const void *dev_recv_left, const void *dev_recv_right, uint64_t cap_rows)
{
    int rc = enter(h);
    if (rc) return rc;
    const void *src[2] = {dev_recv_left, dev_recv_right};
    for (int side = 0; side < 2; ++side) {
        if (!src[side] || h->ghost_n[side] == 0) continue;
        if (cap_rows != h->ghost_n[side]) return fail(h, SPH_ERR_INVALID, "capacity differs from the ghost batch's");
        k_slab_fast_set_ghost_density<<<blocks_for(cap_rows, SLAB_THREADS), SLAB_THREADS, 0, h->stream>>>(
            h->vel[h->cur], h->inverse, (uint32_t)h->ghost_first[side], (uint32_t)cap_rows, (const float *)src[side]);
        CK_LAUNCH();
        ++h->launches;
    }
    return SPH_OK;
}

// ---- peer-memory slab path: pack kernels store into the neighbours' mailboxes over NVLink ----------

int sph_slab_p2p_create(sph_handle *h, uint64_t halo_rows, uint64_t migrant_rows, void *ipc_handle_out64)
{
    int rc = enter_exact(h);
    if (rc) return rc;
    if (!ipc_handle_out64 || halo_rows == 0 || migrant_rows == 0) return fail(h, SPH_ERR_INVALID, "bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle is exchanged as 64 bytes");
    if (h->mailbox) return fail(h, SPH_ERR_STATE, "mailbox already created");
    P2PLayout L{};
    unsigned long long off = 0;
    auto take = [&](unsigned long long bytes) { const unsigned long long o = off; off += (bytes + 255) / 256 * 256; return o; };
    for (int s2 = 0; s2 < 2; ++s2)
        for (int b = 0; b < 2; ++b) {
            L.mig[s2][b] = take(migrant_rows * 32);
            L.halo[s2][b] = take(halo_rows * 32);
            L.rho[s2][b] = take((halo_rows + migrant_rows) * 4);  // halo densities, then those of the arrivals
        }
    for (int s2 = 0; s2 < 2; ++s2) {
        for (int t = 0; t < 3; ++t)
            for (int b = 0; b < 2; ++b) L.count[s2][t][b] = take(4);
        for (int f = 0; f < 2; ++f) L.flag[s2][f] = take(4);
    }
    L.bytes = off;
    CK(cudaMalloc(&h->mailbox, L.bytes));
    CK(cudaMemset(h->mailbox, 0, L.bytes));
    cudaIpcMemHandle_t ih;
    CK(cudaIpcGetMemHandle(&ih, h->mailbox));
    std::memcpy(ipc_handle_out64, &ih, 64);
    h->p2p = L;
    h->p2p_H = halo_rows;
    h->p2p_M = migrant_rows;
    h->p2p_epoch = 0;
    if (!h->mig_rows[0]) {
        CK(cudaMalloc(&h->mig_rows[0], sizeof(uint32_t) * migrant_rows));
        CK(cudaMalloc(&h->mig_rows[1], sizeof(uint32_t) * migrant_rows));
    }
    return SPH_OK;
}

int sph_slab_p2p_connect(sph_handle *h, int side, const void *peer_ipc_handle64)
{
    int rc = enter(h);
    if (rc) return rc;
    if (side < 0 || side > 1 || !peer_ipc_handle64) return fail(h, SPH_ERR_INVALID, "bad arguments");
    if (!h->mailbox) return fail(h, SPH_ERR_STATE, "create the local mailbox first");
    cudaIpcMemHandle_t ih;
    std::memcpy(&ih, peer_ipc_handle64, 64);
    void *ptr = nullptr;
    CK(cudaIpcOpenMemHandle(&ptr, ih, cudaIpcMemLazyEnablePeerAccess));
    h->peer_mailbox[side] = (char *)ptr;
    return SPH_OK;
}

namespace {
// Mailbox side indices: messages that arrive FROM my left neighbour land in my side 0; so when I send
// to my LEFT neighbour I write into ITS side 1 (I am its right neighbour), and vice versa.
inline char *peer_slot(sph_handle *h, int to_side, unsigned long long off) { return h->peer_mailbox[to_side] + off; }
}  // namespace

int sph_slab_p2p_begin(sph_handle *h, int32_t lo, int32_t hi, int32_t lo_prev, int32_t hi_next, uint64_t migrant_rows)
{
    int rc = enter_exact(h);
    if (rc) return rc;
    if (!h->have_state || !h->slab_mode || !h->mailbox) return fail(h, SPH_ERR_STATE, "slab mode with a mailbox required");
    if (h->n == 0) return fail(h, SPH_ERR_STATE, "a peer step needs at least one row on every rank");
    const bool has_l = h->peer_mailbox[0] != nullptr, has_r = h->peer_mailbox[1] != nullptr;
    if ((has_l || has_r) && (long long)hi - (long long)lo < 3 && (has_l ? 1 : 0) + (has_r ? 1 : 0) + ((long long)hi - lo) < 4)
        return fail(h, SPH_ERR_INVALID, "a slab of %lld x-cells is too narrow for the peer step (3 needed): use the NCCL transport",
                    (long long)hi - (long long)lo);
    if (migrant_rows == 0 || migrant_rows > h->p2p_M) migrant_rows = h->p2p_M;
    cudaStream_t s = h->stream;
    ++h->p2p_epoch;
    const int b = h->p2p_epoch & 1;
    const P2PLayout &L = h->p2p;
    unsigned long long *cur = h->slab_counts + 2 * SLAB_MAX_RANKS;
    unsigned int *done = reinterpret_cast<unsigned int *>(cur + P2P_CUR_WORDS);
    uint32_t *saved = done + P2P_DONE_WORDS;  // [type][side]
    if (!h->p2p_clean) {  // first peer step, or the general path used these words since: the step itself re-zeroes them
        CK(cudaMemsetAsync(cur, 0, (P2P_CUR_WORDS + P2P_DONE_WORDS / 2) * sizeof(unsigned long long), s));
        h->p2p_clean = true;
    }
    // The violation word (aux[3]) is NOT cleared here: a time-out raised by the density leg of the previous
    // step — after that step's copy to the host — is picked up by this step's copy and reported by the next.
    P2PPublish pub{};
    float4 *mig[2] = {nullptr, nullptr}, *halo[2] = {nullptr, nullptr};
    for (int side = 0; side < 2; ++side) {
        if (!h->peer_mailbox[side]) continue;
        mig[side] = (float4 *)peer_slot(h, side, L.mig[side ^ 1][b]);
        halo[side] = (float4 *)peer_slot(h, side, L.halo[side ^ 1][b]);
        pub.peer_count[0][side] = (uint32_t *)peer_slot(h, side, L.count[side ^ 1][P2P_MIG][b]);
        pub.peer_count[1][side] = (uint32_t *)peer_slot(h, side, L.count[side ^ 1][P2P_HALO][b]);
        pub.peer_flag[side] = (uint32_t *)peer_slot(h, side, L.flag[side ^ 1][0]);
        pub.cursor[0][side] = cur + side;
        pub.cursor[1][side] = cur + 2 + side;
        pub.saved[0][side] = saved + side;
        pub.saved[1][side] = saved + 2 + side;
    }
    pub.cap[0] = (uint32_t)migrant_rows;
    pub.cap[1] = (uint32_t)h->p2p_H;
    pub.done = done;
    pub.epoch = h->p2p_epoch;
    pub.ntypes = 2;
    begin_edge_scans(h);
    // migrants and halo rows stored straight into the neighbours' mailboxes by one pass over the edge rows; the
    // last block publishes the four counts and raises the two flags
    k_p2p_pack<<<edge_blocks(h), SLAB_THREADS, 0, s>>>(
        h->pos[h->cur], h->vel[h->cur], (uint32_t)h->n, h->P.h, lo, hi, lo_prev, hi_next, has_l, has_r, (uint32_t)migrant_rows,
        (uint32_t)h->p2p_H, mig[0], mig[1], halo[0], halo[1], h->mig_rows[0], h->mig_rows[1], h->halo_rows[0], h->halo_rows[1], cur,
        &h->ctr->aux[3], h->gd, h->cells, h->ctr, h->edge_all, pub);
    CK_STEP_LAUNCH();
    h->p2p_M_step = migrant_rows;
    h->fast_halo_cap = h->p2p_H;
    h->slab_fast = true;
    h->n_ghost = 0;
    h->ghost_n[0] = h->ghost_n[1] = 0;
    h->have_step = false;
    // the x-layers of this slab that touch a neighbour's ghosts: lo (if there is a left neighbour), hi - 1 (right)
    h->slab_lo = has_l ? lo : -0x3fffffff;
    h->slab_hi = has_r ? hi : 0x3fffffff;
    h->p2p_lo = lo;
    h->p2p_hi = hi;
    return SPH_OK;
}

// Both neighbours' migrant and halo messages appended behind the current rows by one kernel that waits for
// their flags: regions [arrivals L][arrivals R][ghosts L][ghosts R] of fixed size.
int sph_slab_p2p_arrivals(sph_handle *h)
{
    int rc = enter(h);
    if (rc) return rc;
    const P2PLayout &L = h->p2p;
    const int b = h->p2p_epoch & 1;
    const uint64_t cm = h->p2p_M_step, ch = h->p2p_H;
    P2PIncoming in{};
    int nsides = 0;
    for (int side = 0; side < 2; ++side) nsides += h->peer_mailbox[side] != nullptr;
    if (h->n + (uint64_t)nsides * (cm + ch) > h->cap)
        return fail(h, SPH_ERR_CAPACITY, "appending %d x (%llu + %llu) message rows to %llu rows exceeds capacity %llu", nsides,
                    (unsigned long long)cm, (unsigned long long)ch, (unsigned long long)h->n, (unsigned long long)h->cap);
    for (int side = 0; side < 2; ++side) {
        h->arr_first[side] = h->n;
        if (!h->peer_mailbox[side]) continue;
        in.mig[side] = (const float4 *)(h->mailbox + L.mig[side][b]);
        in.mig_count[side] = (const uint32_t *)(h->mailbox + L.count[side][P2P_MIG][b]);
        in.flag[side] = (const uint32_t *)(h->mailbox + L.flag[side][0]);
        in.mig_first[side] = (uint32_t)h->n;
        h->n += cm;
    }
    for (int side = 0; side < 2; ++side) {
        h->ghost_first[side] = h->n;
        h->ghost_n[side] = h->peer_mailbox[side] ? ch : 0;
        if (!h->peer_mailbox[side]) continue;
        in.halo[side] = (const float4 *)(h->mailbox + L.halo[side][b]);
        in.halo_count[side] = (const uint32_t *)(h->mailbox + L.count[side][P2P_HALO][b]);
        in.halo_first[side] = (uint32_t)h->n;
        h->n += ch;
        h->n_ghost += ch;
    }
    if (nsides) {
        // few blocks (they all poll the flag word first, see p2p_wait_flag): one per SM over the two sides
        k_p2p_append<<<dim3(std::min(blocks_for(cm + ch, SLAB_THREADS), (unsigned)(h->num_sms + 1) / 2), 2), SLAB_THREADS, 0, h->stream>>>(
            in, (uint32_t)cm, (uint32_t)ch, h->p2p_epoch, h->P.h, h->p2p_lo, h->p2p_hi, h->peer_mailbox[0] != nullptr,
            h->peer_mailbox[1] != nullptr, h->pos[h->cur], h->vel[h->cur], &h->ctr->aux[3]);
        CK_STEP_LAUNCH();
    }
    return SPH_OK;
}

// Densities of my boundary rows and of this step's arrivals into the neighbours' mailboxes, then the
// neighbours' into my ghosts (appended ones and the migrants I kept as ghosts).
int sph_slab_p2p_density(sph_handle *h)
{
    int rc = enter(h);
    if (rc) return rc;
    cudaStream_t s = h->stream;
    const P2PLayout &L = h->p2p;
    const int b = h->p2p_epoch & 1;
    const uint32_t ch = (uint32_t)h->p2p_H, cm = (uint32_t)h->p2p_M_step;
    unsigned long long *cur = h->slab_counts + 2 * SLAB_MAX_RANKS;
    unsigned int *done = reinterpret_cast<unsigned int *>(cur + P2P_CUR_WORDS);
    uint32_t *saved = done + P2P_DONE_WORDS;
    if (!h->peer_mailbox[0] && !h->peer_mailbox[1]) return SPH_OK;
    P2PRhoOut o{};
    P2PRhoIn in{};
    P2PPublish pub{};
    for (int side = 0; side < 2; ++side) {
        if (!h->peer_mailbox[side]) continue;
        o.out[side] = (float *)peer_slot(h, side, L.rho[side ^ 1][b]);
        o.halo_rows[side] = h->halo_rows[side];
        o.halo_sent[side] = saved + 2 + side;
        o.arrived[side] = (const uint32_t *)(h->mailbox + L.count[side][P2P_MIG][b]);
        o.arr_first[side] = (uint32_t)h->arr_first[side];
        pub.peer_flag[side] = (uint32_t *)peer_slot(h, side, L.flag[side ^ 1][1]);
        in.rho[side] = (const float *)(h->mailbox + L.rho[side][b]);
        in.flag[side] = (const uint32_t *)(h->mailbox + L.flag[side][1]);
        in.halo_count[side] = (const uint32_t *)(h->mailbox + L.count[side][P2P_HALO][b]);
        in.mig_sent[side] = saved + side;
        in.mig_rows[side] = h->mig_rows[side];
        in.ghost_first[side] = (uint32_t)h->ghost_first[side];
    }
    pub.done = done + 1;
    pub.epoch = h->p2p_epoch;
    pub.ntypes = 0;
    // With the overlap on, the exchange runs on the second stream while the main stream integrates the rows
    // that have no ghost among their neighbours (sph_slab_step_forces); the boundary rows' launch waits for it.
    cudaStream_t xs = s;
    if (h->p2p_overlap) {
        CK(cudaEventRecord(h->ev_dens, s));
        CK(cudaStreamWaitEvent(h->stream2, h->ev_dens, 0));
        xs = h->stream2;
    }
    const unsigned few = std::min(blocks_for(ch + cm, SLAB_THREADS), (unsigned)(h->num_sms + 1) / 2);
    k_p2p_rho_pack<<<dim3(few, 2), SLAB_THREADS, 0, xs>>>(h->vel[h->cur], h->inverse, o, ch, cm, pub);
    CK_STEP_LAUNCH();
    k_p2p_rho_apply<<<dim3(few, 2), SLAB_THREADS, 0, xs>>>(in, ch, cm, h->p2p_epoch, h->vel[h->cur], h->inverse, &h->ctr->aux[3], cur);
    CK_STEP_LAUNCH();
    if (h->p2p_overlap) {
        CK(cudaEventRecord(h->ev_rho, xs));
        h->rho_pending = true;
    }
    return SPH_OK;
}

}  // extern "C"
