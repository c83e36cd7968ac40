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
// publishes to the adjacent ranks once every block has stored its rows — the row counts, then, after a
// system-scope fence, the epoch flags the receivers spin on. done == nullptr: not a peer step.
struct P2PPublish {
    uint32_t *peer_count[2];  // [side] row count slot in the neighbour's mailbox (nullptr: no neighbour)
    uint32_t *peer_flag[2];   // [side] epoch flag in the neighbour's mailbox
    unsigned int *done;       // blocks of this kernel that have finished (zero at launch)
    uint32_t epoch, cap;
};

// Called by every thread at the end of a pack kernel (no thread may have returned early).
// wrote: this thread stored rows into a peer mailbox.
__device__ __forceinline__ void p2p_publish_when_last(const P2PPublish &pub, const unsigned long long *cursor_l,
                                                      const unsigned long long *cursor_r, bool wrote)
{
    if (!pub.done) return;
    if (wrote) __threadfence_system();  // my rows are visible to the peer before anything ordered after this
    __syncthreads();
    if (threadIdx.x != 0) return;
    __threadfence();
    if (atomicAdd(pub.done, 1u) != gridDim.x * gridDim.y - 1u) return;
    __threadfence();  // every other block's cursor updates and fences are behind us
    const unsigned long long *cur[2] = {cursor_l, cursor_r};
#pragma unroll
    for (int side = 0; side < 2; ++side)
        if (pub.peer_count[side]) {
            const unsigned long long c = __ldcg(cur[side]);
            *pub.peer_count[side] = (uint32_t)(c < pub.cap ? c : pub.cap);
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
                  const uint32_t *__restrict__ starts, const StepCounters *__restrict__ ctr, bool all,
                  const P2PPublish pub)
{
    __shared__ EdgeScan s_e;
    if (threadIdx.x == 0) edge_scan_plan(s_e, *gd, starts, n, n, lo, hi, 1, all || ctr->fast_x != 0u);
    __syncthreads();
    const EdgeScan e = s_e;
    const uint32_t total = e.len[0] + e.len[1] + e.len[2];
    bool wrote = false;
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
                        wrote = true;
                    }
                }
            }
        }
        if (drop) {
            p.w = __uint_as_float(W_DROP);
            pos[i] = p;
        }
    }
    p2p_publish_when_last(pub, &cur[0], &cur[1], wrote);
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
                 const StepCounters *__restrict__ ctr, uint32_t sorted, bool all, const P2PPublish pub)
{
    __shared__ EdgeScan s_e;
    if (threadIdx.x == 0) edge_scan_plan(s_e, *gd, starts, n, sorted, lo, hi, 2, all || ctr->fast_x != 0u);
    __syncthreads();
    const EdgeScan e = s_e;
    const uint32_t total = e.len[0] + e.len[1] + e.len[2];
    bool wrote = false;
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
            if (k < cap) { send_l[2 * k] = p; send_l[2 * k + 1] = v; rows_l[k] = i; wrote = true; }
            else atomicOr(err, SLAB_ERR_HALO_OVERFLOW);
        }
        if (to_r) {
            const unsigned long long k = atomicAdd(&cur[3], 1ull);
            if (k < cap) { send_r[2 * k] = p; send_r[2 * k + 1] = v; rows_r[k] = i; wrote = true; }
            else atomicOr(err, SLAB_ERR_HALO_OVERFLOW);
        }
    }
    p2p_publish_when_last(pub, &cur[2], &cur[3], wrote);
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
    unsigned long long mig[2][2], halo[2][2], rho[2][2];  // byte offsets [side][parity]
    unsigned long long count[2][3][2];                     // uint32 [side][type][parity]
    unsigned long long flag[2][3];                         // uint32 [side][type]
    unsigned long long bytes;
};
enum { P2P_MIG = 0, P2P_HALO = 1, P2P_RHO = 2 };

// Device-side scratch of the peer step, in sph_handle::slab_counts behind the general path's counters:
// cur[0..1] migrant cursors, cur[2..3] halo cursors, then one "blocks done" counter per publishing kernel.
// All of it is zero when a peer step begins: k_p2p_rho_apply, the last exchange kernel of a step, re-zeroes it.
constexpr int P2P_CUR_WORDS = 4;  // unsigned long long cursors
constexpr int P2P_DONE_WORDS = 4; // unsigned int counters after them (migrants, halo, rho, spare)

// Both incoming messages of one type (side = blockIdx.y) appended behind the current rows, each into a
// fixed region of `cap` rows: rows past the message's count become dropped rows. The block first waits for
// the neighbour's flag (the flag-wait used to be a kernel of its own).
struct P2PIncoming {
    const float4 *rows[2];    // [side] message in the local mailbox (nullptr: no neighbour)
    const uint32_t *count[2];
    const uint32_t *flag[2];
    uint32_t first[2];        // first destination row
};

__global__ void __launch_bounds__(SLAB_THREADS)
k_p2p_append(const P2PIncoming in, uint32_t cap, uint32_t epoch, bool ghost, float4 *__restrict__ pos,
             float4 *__restrict__ vel, uint32_t *err)
{
    const int side = blockIdx.y;
    if (!in.rows[side]) return;
    p2p_wait_flag(in.flag[side], epoch, err);
    const uint32_t first = in.first[side], count = *in.count[side];
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < cap; k += gridDim.x * blockDim.x) {
        if (k >= count) {
            pos[first + k] = make_float4(0.f, 0.f, 0.f, __uint_as_float(W_DROP));
            continue;
        }
        float4 p = in.rows[side][2 * k];
        const float4 v = in.rows[side][2 * k + 1];
        uint32_t w = __float_as_uint(p.w) & W_ID_MASK;
        if (ghost) w |= W_GHOST;
        p.w = __uint_as_float(w);
        pos[first + k] = p;
        vel[first + k] = v;
    }
}

// Densities of my boundary rows (the rows of the two halo messages, same order) stored into the
// neighbours' mailboxes; the last block publishes counts and flags.
__global__ void __launch_bounds__(SLAB_THREADS)
k_p2p_rho_pack(const float4 *__restrict__ vel, const uint32_t *__restrict__ inverse, const uint32_t *__restrict__ rows_l,
               const uint32_t *__restrict__ rows_r, const unsigned long long *__restrict__ cur, uint32_t cap,
               float *__restrict__ out_l, float *__restrict__ out_r, const P2PPublish pub)
{
    const int side = blockIdx.y;
    float *out = side ? out_r : out_l;
    bool wrote = false;
    if (out) {
        const uint32_t *rows = side ? rows_r : rows_l;
        const unsigned long long c = cur[2 + side];
        const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
        if (k < cap && k < c) {
            out[k] = vel[inverse[rows[k]]].w;
            wrote = true;
        }
    }
    p2p_publish_when_last(pub, &cur[2], &cur[3], wrote);
}

// The neighbours' densities into my ghost rows (ghost batch `side` was appended at pre-sort rows
// [first[side], first[side] + cap)); then the step's cursors and done-counters are zeroed for the next step.
struct P2PRhoIncoming {
    const float *rho[2];
    const uint32_t *count[2];
    const uint32_t *flag[2];
    uint32_t first[2];
};

__global__ void __launch_bounds__(SLAB_THREADS)
k_p2p_rho_apply(const P2PRhoIncoming in, uint32_t cap, uint32_t epoch, float4 *__restrict__ vel,
                const uint32_t *__restrict__ inverse, uint32_t *err, unsigned long long *cur)
{
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x < P2P_CUR_WORDS + P2P_DONE_WORDS / 2) cur[threadIdx.x] = 0ull;
    const int side = blockIdx.y;
    if (!in.rho[side]) return;
    p2p_wait_flag(in.flag[side], epoch, err);
    const uint32_t count = min(cap, *in.count[side]);
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < count; k += gridDim.x * blockDim.x)
        vel[inverse[in.first[side] + k]].w = in.rho[side][k];
}

}  // namespace sphb
