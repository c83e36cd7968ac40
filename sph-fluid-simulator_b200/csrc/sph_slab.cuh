// Slab decomposition primitives (one process per GPU; the exchange itself is done by the caller
// with NCCL, see sph-fluid-simulator_b200/slab.py). The domain is cut along x at cell
// boundaries: rank r owns the particles whose cell.x (reference getCell, src/neighborTable.cpp:
// 14-17) lies in [cuts[r], cuts[r+1]). The reference has no multi-GPU path; this is new work
//  and is validated against the single-GPU path and the oracle by particle id.
//
// Wire format of a particle row: two float4 — (x, y, z, id bits) and (vx, vy, vz, 0) — 32 bytes.
#pragma once

#include "sph_device.cuh"

namespace sphb {

constexpr int SLAB_THREADS = 256;
constexpr int SLAB_MAX_RANKS = 64;

struct SlabCuts {
    int world;
    int lo[SLAB_MAX_RANKS + 1];  // lo[0] is treated as -inf, lo[world] as +inf
};

__device__ __forceinline__ int slab_owner(const SlabCuts &c, int cx)
{
    int r = 0;
    for (int k = 1; k < c.world; ++k) r += (cx >= c.lo[k]);
    return r;
}

// counts[r] += live owned rows (not dropped, not ghost) whose owner is rank r.
__global__ void __launch_bounds__(SLAB_THREADS)
k_slab_count(const float4 *__restrict__ pos, uint32_t n, float h, const SlabCuts cuts,
             unsigned long long *__restrict__ counts)
{
    __shared__ unsigned int s_cnt[SLAB_MAX_RANKS];
    for (int k = threadIdx.x; k < SLAB_MAX_RANKS; k += blockDim.x) s_cnt[k] = 0;
    __syncthreads();
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const float4 p = pos[i];
        const uint32_t w = __float_as_uint(p.w);
        if (w != W_DROP && !(w & W_GHOST)) atomicAdd(&s_cnt[slab_owner(cuts, cell_of(p.x, h))], 1u);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < cuts.world; k += blockDim.x)
        if (s_cnt[k]) atomicAdd(&counts[k], (unsigned long long)s_cnt[k]);
}

// Rows owned by another rank are copied to that rank's segment of the send buffer and dropped
// here; last step's ghosts are dropped. cursors[r] must start at 0; offsets[r] is the first row
// of rank r's segment. The order inside a segment is arbitrary — nothing downstream depends on
// row order (cells are ordered by particle id).
struct SlabOffsets {
    unsigned long long row[SLAB_MAX_RANKS];
};

__global__ void __launch_bounds__(SLAB_THREADS)
k_slab_pack(float4 *__restrict__ pos, const float4 *__restrict__ vel, uint32_t n, float h, const SlabCuts cuts,
            int self, const SlabOffsets offsets, unsigned long long *__restrict__ cursors,
            float4 *__restrict__ sendbuf)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 p = pos[i];
    const uint32_t w = __float_as_uint(p.w);
    if (w == W_DROP) return;
    bool drop = (w & W_GHOST) != 0u;
    if (!drop) {
        const int dest = slab_owner(cuts, cell_of(p.x, h));
        if (dest != self) {
            const unsigned long long k = offsets.row[dest] + atomicAdd(&cursors[dest], 1ull);
            float4 v = vel[i];
            v.w = 0.f;
            sendbuf[2 * k] = p;
            sendbuf[2 * k + 1] = v;
            drop = true;
        }
    }
    if (drop) {
        p.w = __uint_as_float(W_DROP);
        pos[i] = p;
    }
}

// Append received rows after the current rows, as owned particles or as ghosts. Rows whose id
// lane is the dropped-row pattern (the unused tail of a fixed-size message, pre-filled with 0xFF)
// stay dropped rows.
__global__ void __launch_bounds__(SLAB_THREADS)
k_slab_append(const float4 *__restrict__ rows, uint32_t nrows, uint32_t first, bool ghost,
              float4 *__restrict__ pos, float4 *__restrict__ vel)
{
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nrows) return;
    float4 p = rows[2 * k];
    const float4 v = rows[2 * k + 1];
    uint32_t w = __float_as_uint(p.w);
    if (w != W_DROP) {
        w &= W_ID_MASK;
        if (ghost) w |= W_GHOST;
    }
    p.w = __uint_as_float(w);
    pos[first + k] = p;
    vel[first + k] = v;
}

// ---- sync-free path: fixed-size messages to the two adjacent ranks, counts stay on the device ----

// Violation bits accumulated in StepCounters::aux[3] and read back one step late.
constexpr uint32_t SLAB_ERR_MIGRANT_OVERFLOW = 1u, SLAB_ERR_HALO_OVERFLOW = 2u, SLAB_ERR_NOT_ADJACENT = 4u;
constexpr uint32_t SLAB_ERR_P2P_TIMEOUT = 8u;

// Peer-memory transport (described at the end of this file): what the LAST block of a pack kernel
// publishes to the adjacent ranks once every block has stored its rows — the row counts of up to two
// message types, then, after a system-scope fence, ONE epoch flag per neighbour that the receiver spins on.
struct P2PPublish {
    uint32_t *peer_count[2][2];          // [type][side] count slot in the neighbour's mailbox (nullptr: none)
    uint32_t *peer_flag[2];              // [side] epoch flag in the neighbour's mailbox (nullptr: no neighbour)
    const unsigned long long *cursor[2][2];  // [type][side] local cursors the counts come from
    uint32_t *saved[2][2];               // [type][side] local copy of the published count (read by later kernels)
    uint32_t cap[2];                     // [type] message capacity
    unsigned int *done;                  // blocks of this kernel that have finished (zero at launch)
    uint32_t epoch;
    int ntypes;
};

// Called by every thread at the end of a pack kernel (no thread may have returned early).
// wrote: this thread stored rows into a peer mailbox.
__device__ __forceinline__ void p2p_publish_when_last(const P2PPublish &pub, bool wrote)
{
    if (wrote) __threadfence_system();  // my rows are visible to the peer before anything ordered after this
    __syncthreads();
    if (threadIdx.x != 0) return;
    __threadfence();
    if (atomicAdd(pub.done, 1u) != gridDim.x * gridDim.y - 1u) return;
    __threadfence();  // every other block's cursor updates and fences are behind us
    for (int t = 0; t < pub.ntypes; ++t)
#pragma unroll
        for (int side = 0; side < 2; ++side) {
            if (!pub.cursor[t][side]) continue;
            const unsigned long long c = __ldcg(pub.cursor[t][side]);
            const uint32_t cnt = (uint32_t)(c < pub.cap[t] ? c : pub.cap[t]);
            if (pub.saved[t][side]) *pub.saved[t][side] = cnt;
            if (pub.peer_count[t][side]) *pub.peer_count[t][side] = cnt;
        }
    __threadfence_system();
#pragma unroll
    for (int side = 0; side < 2; ++side)
        if (pub.peer_flag[side]) *reinterpret_cast<volatile uint32_t *>(pub.peer_flag[side]) = pub.epoch;
}

// First thing a consuming block does: thread 0 spins until the local flag reaches the epoch (gives up
// after ~4 s and records it), then the block proceeds. The consuming kernels run FEW blocks (grid-stride
// over the message): every block polls the same word, and a thousand pollers on one L2 line starve the
// neighbour's incoming store of that very word (measured: 7.8 ms per step with one block per 256 rows).
__device__ __forceinline__ void p2p_wait_flag(const uint32_t *flag, uint32_t epoch, uint32_t *err)
{
    if (threadIdx.x == 0) {
        const volatile uint32_t *f = flag;
        const long long t0 = clock64();
        while ((int)(*f - epoch) < 0) {
            __nanosleep(500);
            if (clock64() - t0 > 8000000000ll) {
                atomicOr(err, SLAB_ERR_P2P_TIMEOUT);
                break;
            }
        }
        __threadfence_system();
    }
    __syncthreads();
}

// ---- edge scans ---------------------------------------------------------------------------------
//
// Between two grid builds the rows stay in the start-of-step cell order (x slowest), so the rows of
// the slab's outer x-layers are two contiguous ranges delimited by the cell-start table, and rows
// appended since (arrivals) are a third one at the end. As long as no particle crosses more than
// one cell per step (StepCounters::fast_x, set by the integration), migrants, stale ghosts and the
// next halo can only be found there: the sync-free step scans ~2 % of the rows instead of all of
// them, twice. `all` (rows not in cell order: after an upload or a general-path step) or a fast
// particle fall back to every row.
struct EdgeScan {
    uint32_t first[3], len[3];
};

// Rows whose start-of-step cell.x <= lo + depth - 1 or >= hi - depth, plus the rows [sorted, n).
__device__ __forceinline__ void edge_scan_plan(EdgeScan &e, const GridDesc &g, const uint32_t *__restrict__ starts,
                                               uint32_t n, uint32_t sorted, int lo, int hi, int depth, bool all)
{
    sorted = min(sorted, n);
    uint32_t l_end = sorted, r_begin = sorted;
    if (!all) {
        const long long gl = min(max((long long)lo + depth - 1 - g.ox, 1LL), (long long)g.nx - 2);  // last left layer
        const long long gr = min(max((long long)hi - depth - g.ox, 1LL), (long long)g.nx - 2);      // first right layer
        if (gr > gl) {
            l_end = min(starts[(uint32_t)(gl + 1) * g.sx], sorted);
            r_begin = max(min(starts[(uint32_t)gr * g.sx], sorted), l_end);
        }
    }
    e.first[0] = 0;       e.len[0] = l_end;
    e.first[1] = r_begin; e.len[1] = sorted - r_begin;
    e.first[2] = sorted;  e.len[2] = n - sorted;
}

__device__ __forceinline__ uint32_t edge_scan_row(const EdgeScan &e, uint32_t t)
{
    if (t < e.len[0]) return t;
    t -= e.len[0];
    if (t < e.len[1]) return e.first[1] + t;
    return e.first[2] + (t - e.len[1]);
}

// One pass over the edge rows: last step's ghosts are dropped; owned rows whose cell.x left [lo, hi)
// are copied to the left / right migrant message (pre-filled with dropped rows by the caller)
// and dropped here. cur[0], cur[1] = message cursors (start at 0). A row that would have to travel
// further than the adjacent rank, or does not fit, raises a violation bit and stays put.
__global__ void __launch_bounds__(SLAB_THREADS)
k_slab_fast_begin(float4 *__restrict__ pos, const float4 *__restrict__ vel, uint32_t n, float h, int lo, int hi,
                  int lo_prev, int hi_next, uint32_t cap, float4 *__restrict__ send_l, float4 *__restrict__ send_r,
                  unsigned long long *__restrict__ cur, uint32_t *__restrict__ err, const GridDesc *__restrict__ gd,
                  const uint32_t *__restrict__ starts, const StepCounters *__restrict__ ctr, bool all)
{
    __shared__ EdgeScan s_e;
    if (threadIdx.x == 0) edge_scan_plan(s_e, *gd, starts, n, n, lo, hi, 1, all || ctr->fast_x != 0u);
    __syncthreads();
    const EdgeScan e = s_e;
    const uint32_t total = e.len[0] + e.len[1] + e.len[2];
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
        const uint32_t i = edge_scan_row(e, t);
        float4 p = pos[i];
        const uint32_t w = __float_as_uint(p.w);
        if (w == W_DROP) continue;
        bool drop = (w & W_GHOST) != 0u;
        if (!drop) {
            const int cx = cell_of(p.x, h);
            const int side = cx < lo ? 0 : (cx >= hi ? 1 : -1);
            if (side >= 0) {
                if ((side == 0 && cx < lo_prev) || (side == 1 && cx >= hi_next)) {
                    atomicOr(err, SLAB_ERR_NOT_ADJACENT);
                } else {
                    const unsigned long long k = atomicAdd(&cur[side], 1ull);
                    if (k >= cap) {
                        atomicOr(err, SLAB_ERR_MIGRANT_OVERFLOW);
                    } else {
                        float4 *dst = side == 0 ? send_l : send_r;
                        float4 v = vel[i];
                        v.w = 0.f;
                        dst[2 * k] = p;
                        dst[2 * k + 1] = v;
                        drop = true;
                    }
                }
            }
        }
        if (drop) {
            p.w = __uint_as_float(W_DROP);
            pos[i] = p;
        }
    }
}

// Both halo messages in one pass over the edge rows (two layers deep: a row of layer lo + 1 may have
// moved into layer lo) and the arrivals appended at [sorted, n): owned rows of x-cell lo go left, of
// x-cell hi-1 go right (a one-cell slab sends its rows both ways). cur[2], cur[3] = cursors;
// rows_l / rows_r remember the source rows so the densities can follow in the same order.
__global__ void __launch_bounds__(SLAB_THREADS)
k_slab_fast_halo(const float4 *__restrict__ pos, const float4 *__restrict__ vel, uint32_t n, float h, int lo, int hi,
                 bool has_left, bool has_right, uint32_t cap, float4 *__restrict__ send_l, float4 *__restrict__ send_r,
                 uint32_t *__restrict__ rows_l, uint32_t *__restrict__ rows_r, unsigned long long *__restrict__ cur,
                 uint32_t *__restrict__ err, const GridDesc *__restrict__ gd, const uint32_t *__restrict__ starts,
                 const StepCounters *__restrict__ ctr, uint32_t sorted, bool all)
{
    __shared__ EdgeScan s_e;
    if (threadIdx.x == 0) edge_scan_plan(s_e, *gd, starts, n, sorted, lo, hi, 2, all || ctr->fast_x != 0u);
    __syncthreads();
    const EdgeScan e = s_e;
    const uint32_t total = e.len[0] + e.len[1] + e.len[2];
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
        const uint32_t i = edge_scan_row(e, t);
        const float4 p = pos[i];
        const uint32_t w = __float_as_uint(p.w);
        if (w == W_DROP || (w & W_GHOST)) continue;
        const int cx = cell_of(p.x, h);
        const bool to_l = has_left && cx == lo, to_r = has_right && cx == hi - 1;
        if (!to_l && !to_r) continue;
        float4 v = vel[i];
        v.w = 0.f;
        if (to_l) {
            const unsigned long long k = atomicAdd(&cur[2], 1ull);
            if (k < cap) { send_l[2 * k] = p; send_l[2 * k + 1] = v; rows_l[k] = i; }
            else atomicOr(err, SLAB_ERR_HALO_OVERFLOW);
        }
        if (to_r) {
            const unsigned long long k = atomicAdd(&cur[3], 1ull);
            if (k < cap) { send_r[2 * k] = p; send_r[2 * k + 1] = v; rows_r[k] = i; }
            else atomicOr(err, SLAB_ERR_HALO_OVERFLOW);
        }
    }
}

// Densities of the rows of one halo message; the message count is on the device (cursor). The
// output was pre-filled with 0xFF words, which the receiver treats as "no row".
__global__ void __launch_bounds__(SLAB_THREADS)
k_slab_fast_pack_density(const float4 *__restrict__ vel, const uint32_t *__restrict__ inverse,
                         const uint32_t *__restrict__ rows, const unsigned long long *__restrict__ count, uint32_t cap,
                         float *__restrict__ out)
{
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long c = *count;
    if (k < cap && k < c) out[k] = vel[inverse[rows[k]]].w;
}

__global__ void __launch_bounds__(SLAB_THREADS)
k_slab_fast_set_ghost_density(float4 *__restrict__ vel, const uint32_t *__restrict__ inverse, uint32_t first,
                              uint32_t cap, const float *__restrict__ in)
{
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= cap) return;
    const float r = in[k];
    if (__float_as_uint(r) != 0xFFFFFFFFu) vel[inverse[first + k]].w = r;
}

// Copy the live owned rows of x-cell `cell_x` (a boundary layer of the slab) into a halo
// message and remember which rows they were, so their densities can follow in the same order.
__global__ void __launch_bounds__(SLAB_THREADS)
k_slab_pack_halo(const float4 *__restrict__ pos, const float4 *__restrict__ vel, uint32_t n, float h, int cell_x,
                 unsigned long long *__restrict__ cursor, uint32_t capacity, float4 *__restrict__ buf,
                 uint32_t *__restrict__ rows_out)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = pos[i];
    const uint32_t w = __float_as_uint(p.w);
    if (w == W_DROP || (w & W_GHOST) || cell_of(p.x, h) != cell_x) return;
    const unsigned long long k = atomicAdd(cursor, 1ull);
    if (k >= capacity) return;  // caller checks the cursor against the capacity
    float4 v = vel[i];
    v.w = 0.f;
    buf[2 * k] = p;
    buf[2 * k + 1] = v;
    rows_out[k] = i;
}

// Densities of the rows a halo message was packed from (pre-sort row -> sorted row via inverse).
__global__ void __launch_bounds__(SLAB_THREADS)
k_slab_pack_density(const float4 *__restrict__ vel, const uint32_t *__restrict__ inverse,
                    const uint32_t *__restrict__ rows, uint32_t nrows, float *__restrict__ out)
{
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < nrows) out[k] = vel[inverse[rows[k]]].w;  // density rides in vel.w
}

// Densities for a batch of ghost rows that was appended at pre-sort rows [first, first + nrows).
__global__ void __launch_bounds__(SLAB_THREADS)
k_slab_set_ghost_density(float4 *__restrict__ vel, const uint32_t *__restrict__ inverse, uint32_t first,
                         uint32_t nrows, const float *__restrict__ in)
{
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < nrows) vel[inverse[first + k]].w = in[k];
}

// Compact the live owned rows into xyz / xyz / id columns (order arbitrary), for host read-back.
__global__ void __launch_bounds__(SLAB_THREADS)
k_slab_export_owned(const float4 *__restrict__ pos, const float4 *__restrict__ vel, uint32_t n, uint32_t cap,
                    unsigned long long *__restrict__ cursor, float *__restrict__ pos3, float *__restrict__ vel3,
                    uint32_t *__restrict__ ids)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = pos[i];
    const uint32_t w = __float_as_uint(p.w);
    if (w & W_GHOST) return;  // ghost or dropped row
    const unsigned long long k = atomicAdd(cursor, 1ull);
    if (k >= cap) return;
    const float4 v = vel[i];
    pos3[3 * k] = p.x; pos3[3 * k + 1] = p.y; pos3[3 * k + 2] = p.z;
    vel3[3 * k] = v.x; vel3[3 * k + 1] = v.y; vel3[3 * k + 2] = v.z;
    ids[k] = w;
}

// Histogram of cell.x over live owned rows, bins [x_lo, x_lo + nbins) with clamping at both ends
// (used to choose balanced cuts).
__global__ void __launch_bounds__(SLAB_THREADS)
k_slab_xhist(const float4 *__restrict__ pos, uint32_t n, float h, int x_lo, uint32_t nbins,
             unsigned long long *__restrict__ hist)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = pos[i];
    const uint32_t w = __float_as_uint(p.w);
    if (w == W_DROP || (w & W_GHOST)) return;
    long long b = (long long)cell_of(p.x, h) - x_lo;
    b = min(max(b, 0LL), (long long)nbins - 1);
    atomicAdd(&hist[b], 1ull);
}


// ---- peer-memory path: the pack kernels store straight into the neighbour's mailbox over NVLink ----
//
// Every rank owns a mailbox (one device allocation, exported with cudaIpcGetMemHandle and mapped
// by both neighbours). For each side it holds, double-buffered by step parity, a migrant message,
// a halo message and a density message, plus per-message row counts and one epoch flag per
// message type. The sender's pack kernel writes rows directly into the peer mailbox (no send
// buffer, no NCCL kernel), a one-thread publish kernel then stores the row count, fences at system
// scope and raises the peer's flag to the step epoch; the receiver's stream spins on its own flag
// before it appends. Two buffers are enough: a sender reaches epoch e+2 only after it has waited
// for the neighbour's epoch e+1 message, which the neighbour sent after consuming epoch e.
struct P2PLayout {
    unsigned long long mig[2][2], halo[2][2], rho[2][2];  // byte offsets [side][parity]; rho = H + M floats
    unsigned long long count[2][3][2];                     // uint32 [side][type][parity]
    unsigned long long flag[2][2];                         // uint32 [side][0 = rows (migrants + halo), 1 = densities]
    unsigned long long bytes;
};
enum { P2P_MIG = 0, P2P_HALO = 1, P2P_RHO = 2 };
constexpr uint32_t P2P_NO_ROW = 0xFFFFFFFFu;

// Device-side scratch of the peer step, in sph_handle::slab_counts behind the general path's counters:
// cur[0..1] migrant cursors, cur[2..3] halo cursors, then two "blocks done" counters (rows kernel, density
// kernel) and the four published counts saved for the density leg. All of it is zero when a peer step
// begins: k_p2p_rho_apply, the last exchange kernel of a step, re-zeroes cursors and done-counters.
constexpr int P2P_CUR_WORDS = 4;    // unsigned long long cursors
constexpr int P2P_DONE_WORDS = 4;   // unsigned int: [0] rows kernel, [1] density kernel, [2..3] spare
constexpr int P2P_SAVED_WORDS = 4;  // unsigned int after them: my published counts [type][side]

// ONE pass over the slab's edge rows (two x-layers deep on either side; every row when a fast particle was
// seen) that does what used to take two exchange rounds:
//   * last step's ghosts are dropped;
//   * an owned row whose cell.x left [lo, hi) is stored into the left / right neighbour's migrant message.
//     If it lands in the neighbour's BOUNDARY layer (cell.x == lo - 1 or == hi) the neighbour would send it
//     straight back as a ghost — so it simply stays here as a ghost (same position, same velocity), its slot
//     remembered in mig_rows so that the density the new owner computes can find it; otherwise it is dropped;
//   * an owned row that stays and lies in my first / last x-layer is stored into the left / right halo message.
// An arrival never has to go into a halo message of the same step (the sender kept it, see above), so the
// halo messages do not wait for the migrants any more: one kernel, one flag, one wait on the other side —
// provided a slab is at least three cells wide (an arrival cannot reach the far boundary layer; checked on
// arrival). The last block publishes all four counts and the flags.
__global__ void __launch_bounds__(SLAB_THREADS)
k_p2p_pack(float4 *__restrict__ pos, const float4 *__restrict__ vel, uint32_t n, float h, int lo, int hi, int lo_prev,
           int hi_next, bool has_left, bool has_right, uint32_t cap_m, uint32_t cap_h, float4 *__restrict__ mig_l,
           float4 *__restrict__ mig_r, float4 *__restrict__ halo_l, float4 *__restrict__ halo_r,
           uint32_t *__restrict__ mig_rows_l, uint32_t *__restrict__ mig_rows_r, uint32_t *__restrict__ halo_rows_l,
           uint32_t *__restrict__ halo_rows_r, unsigned long long *__restrict__ cur, uint32_t *__restrict__ err,
           const GridDesc *__restrict__ gd, const uint32_t *__restrict__ starts, const StepCounters *__restrict__ ctr,
           bool all, const P2PPublish pub)
{
    __shared__ EdgeScan s_e;
    if (threadIdx.x == 0) edge_scan_plan(s_e, *gd, starts, n, n, lo, hi, 2, all || ctr->fast_x != 0u);
    __syncthreads();
    const EdgeScan e = s_e;
    const uint32_t total = e.len[0] + e.len[1] + e.len[2];
    bool wrote = false;
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
        const uint32_t i = edge_scan_row(e, t);
        float4 p = pos[i];
        const uint32_t w = __float_as_uint(p.w);
        if (w == W_DROP) continue;
        if (w & W_GHOST) {  // last step's halo copy
            p.w = __uint_as_float(W_DROP);
            pos[i] = p;
            continue;
        }
        const int cx = cell_of(p.x, h);
        const int side = (has_left && cx < lo) ? 0 : ((has_right && cx >= hi) ? 1 : -1);
        if (side >= 0) {
            if ((side == 0 && cx < lo_prev) || (side == 1 && cx >= hi_next)) {
                atomicOr(err, SLAB_ERR_NOT_ADJACENT);
                continue;
            }
            const unsigned long long k = atomicAdd(&cur[side], 1ull);
            if (k >= cap_m) {
                atomicOr(err, SLAB_ERR_MIGRANT_OVERFLOW);
                continue;
            }
            float4 *dst = side == 0 ? mig_l : mig_r;
            float4 v = vel[i];
            v.w = 0.f;
            dst[2 * k] = p;
            dst[2 * k + 1] = v;
            wrote = true;
            const bool keep = side == 0 ? cx == lo - 1 : cx == hi;
            (side == 0 ? mig_rows_l : mig_rows_r)[k] = keep ? i : P2P_NO_ROW;
            p.w = __uint_as_float(keep ? (w | W_GHOST) : W_DROP);
            pos[i] = p;
            continue;
        }
        const bool to_l = has_left && cx == lo, to_r = has_right && cx == hi - 1;
        if (!to_l && !to_r) continue;
        float4 v = vel[i];
        v.w = 0.f;
        if (to_l) {
            const unsigned long long k = atomicAdd(&cur[2], 1ull);
            if (k < cap_h) { halo_l[2 * k] = p; halo_l[2 * k + 1] = v; halo_rows_l[k] = i; wrote = true; }
            else atomicOr(err, SLAB_ERR_HALO_OVERFLOW);
        }
        if (to_r) {
            const unsigned long long k = atomicAdd(&cur[3], 1ull);
            if (k < cap_h) { halo_r[2 * k] = p; halo_r[2 * k + 1] = v; halo_rows_r[k] = i; wrote = true; }
            else atomicOr(err, SLAB_ERR_HALO_OVERFLOW);
        }
    }
    p2p_publish_when_last(pub, wrote);
}

// Both neighbours' messages appended behind the current rows by one kernel (blockIdx.y = side): the arrivals
// into a region of cap_m rows, the ghosts into a region of cap_h rows, rows past a message's count become
// dropped rows. The block first waits for that neighbour's flag. An arrival that lands in my FAR boundary
// layer would have had to be in a halo message that is already on its way: a violation (slab too narrow).
struct P2PIncoming {
    const float4 *mig[2], *halo[2];   // [side] messages in the local mailbox (nullptr: no neighbour)
    const uint32_t *mig_count[2], *halo_count[2];
    const uint32_t *flag[2];
    uint32_t mig_first[2], halo_first[2];  // first destination rows
};

__global__ void __launch_bounds__(SLAB_THREADS)
k_p2p_append(const P2PIncoming in, uint32_t cap_m, uint32_t cap_h, uint32_t epoch, float h, int lo, int hi, bool has_left,
             bool has_right, float4 *__restrict__ pos, float4 *__restrict__ vel, uint32_t *err)
{
    const int side = blockIdx.y;
    if (!in.mig[side]) return;
    p2p_wait_flag(in.flag[side], epoch, err);
    const uint32_t nm = *in.mig_count[side], nh = *in.halo_count[side];
    const uint32_t total = cap_m + cap_h;
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
        const bool ghost = t >= cap_m;
        const uint32_t k = ghost ? t - cap_m : t;
        const uint32_t row = (ghost ? in.halo_first[side] : in.mig_first[side]) + k;
        if (k >= (ghost ? nh : nm)) {
            pos[row] = make_float4(0.f, 0.f, 0.f, __uint_as_float(W_DROP));
            continue;
        }
        const float4 *src = ghost ? in.halo[side] : in.mig[side];
        float4 p = src[2 * k];
        const float4 v = src[2 * k + 1];
        uint32_t w = __float_as_uint(p.w) & W_ID_MASK;
        if (ghost) {
            w |= W_GHOST;
        } else {
            const int cx = cell_of(p.x, h);
            if ((side == 0 && has_right && cx >= hi - 1) || (side == 1 && has_left && cx <= lo)) atomicOr(err, SLAB_ERR_NOT_ADJACENT);
        }
        p.w = __uint_as_float(w);
        pos[row] = p;
        vel[row] = v;
    }
}

// Densities into the neighbours' mailboxes (blockIdx.y = side): first those of the rows of my halo message to
// that side (same order), then — at offset cap_h — those of the rows that ARRIVED from that side this step, in
// the order of its migrant message (the sender kept some of them as ghosts and needs their densities back).
struct P2PRhoOut {
    float *out[2];                    // [side] density message in the neighbour's mailbox (nullptr: no neighbour)
    const uint32_t *halo_rows[2];     // pre-sort rows of my halo messages
    const uint32_t *halo_sent[2];     // saved counts of my halo messages
    const uint32_t *arrived[2];       // counts of the migrant messages I received
    uint32_t arr_first[2];            // pre-sort row of the first arrival of each side
};

__global__ void __launch_bounds__(SLAB_THREADS)
k_p2p_rho_pack(const float4 *__restrict__ vel, const uint32_t *__restrict__ inverse, const P2PRhoOut o, uint32_t cap_h,
               uint32_t cap_m, const P2PPublish pub)
{
    const int side = blockIdx.y;
    bool wrote = false;
    if (o.out[side]) {
        const uint32_t nh = *o.halo_sent[side], na = *o.arrived[side];
        for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < cap_h + cap_m; t += gridDim.x * blockDim.x) {
            if (t < cap_h) {
                if (t < nh) { o.out[side][t] = vel[inverse[o.halo_rows[side][t]]].w; wrote = true; }
            } else if (t - cap_h < na) {
                o.out[side][t] = vel[inverse[o.arr_first[side] + (t - cap_h)]].w;
                wrote = true;
            }
        }
    }
    p2p_publish_when_last(pub, wrote);
}

// The neighbours' densities into my ghost rows (blockIdx.y = side): the appended ghost batch (pre-sort rows
// ghost_first[side] + k) and the migrants I kept as ghosts (pre-sort row mig_rows[side][k] of slot k). Then the
// step's cursors and done-counters are zeroed for the next step.
struct P2PRhoIn {
    const float *rho[2];
    const uint32_t *flag[2];
    const uint32_t *halo_count[2];    // count of the halo message I received from that side
    const uint32_t *mig_sent[2];      // saved count of the migrant message I sent to that side
    const uint32_t *mig_rows[2];
    uint32_t ghost_first[2];
};

__global__ void __launch_bounds__(SLAB_THREADS)
k_p2p_rho_apply(const P2PRhoIn in, uint32_t cap_h, uint32_t cap_m, uint32_t epoch, float4 *__restrict__ vel,
                const uint32_t *__restrict__ inverse, uint32_t *err, unsigned long long *cur)
{
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x < P2P_CUR_WORDS + P2P_DONE_WORDS / 2) cur[threadIdx.x] = 0ull;
    const int side = blockIdx.y;
    if (!in.rho[side]) return;
    p2p_wait_flag(in.flag[side], epoch, err);
    const uint32_t nh = min(cap_h, *in.halo_count[side]), nm = min(cap_m, *in.mig_sent[side]);
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < cap_h + cap_m; t += gridDim.x * blockDim.x) {
        if (t < cap_h) {
            if (t < nh) vel[inverse[in.ghost_first[side] + t]].w = in.rho[side][t];
        } else if (t - cap_h < nm) {
            const uint32_t r = in.mig_rows[side][t - cap_h];
            if (r != P2P_NO_ROW) vel[inverse[r]].w = in.rho[side][t];
        }
    }
}

}  // namespace sphb
