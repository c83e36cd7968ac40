// Neighbour-search build: cell bounding box -> dense grid plan -> per-cell histogram ->
// exclusive scan (cell start offsets) -> placement -> stable order -> gather of the particle
// arrays into cell-sorted SoA float4 rows.
//
// This replaces, in the reference: parallelCalculateHashes (src/sph.cpp:17-24), sortParticles
// (src/sph.cpp:184-192) and createNeighborTable (src/neighborTable.cpp:19-37). The reference
// sorts by the truncated 16-bit hash and walks whole hash buckets; here particles are counting-
// sorted by their TRUE cell on a dense grid, so a particle's 27 neighbour cells are 9 contiguous
// runs of the sorted array and bucket-mates from unrelated cells are never touched. The hash16
// the reference would have stored is still computed bit-exactly and kept in pos.w, both as a
// parity artefact and because it decides the reference's double-count rule (sph_device.cuh).
#pragma once

#include "sph_device.cuh"

namespace sphb {

constexpr int GRID_THREADS = 256;

// ---- bounding box of true cells -----------------------------------------------------------

__device__ __forceinline__ void bbox_accumulate(int *bb, int cx, int cy, int cz, bool valid)
{
    const int big = 0x7fffffff;
    int mnx = valid ? cx : big, mny = valid ? cy : big, mnz = valid ? cz : big;
    int mxx = valid ? cx : -big - 1, mxy = valid ? cy : -big - 1, mxz = valid ? cz : -big - 1;
    mnx = __reduce_min_sync(0xffffffffu, mnx);
    mny = __reduce_min_sync(0xffffffffu, mny);
    mnz = __reduce_min_sync(0xffffffffu, mnz);
    mxx = __reduce_max_sync(0xffffffffu, mxx);
    mxy = __reduce_max_sync(0xffffffffu, mxy);
    mxz = __reduce_max_sync(0xffffffffu, mxz);
    if ((threadIdx.x & 31) == 0) {
        // Read first (ld.cg: past the L1, which could hold a stale line for the whole kernel): the box
        // is monotone, so after the first few warps almost no atomic is issued.
        if (mnx < __ldcg(&bb[0])) atomicMin(&bb[0], mnx);
        if (mny < __ldcg(&bb[1])) atomicMin(&bb[1], mny);
        if (mnz < __ldcg(&bb[2])) atomicMin(&bb[2], mnz);
        if (mxx > __ldcg(&bb[3])) atomicMax(&bb[3], mxx);
        if (mxy > __ldcg(&bb[4])) atomicMax(&bb[4], mxy);
        if (mxz > __ldcg(&bb[5])) atomicMax(&bb[5], mxz);
    }
}

// Block-level variant for kernels that already run one thread per particle: warp reductions,
// then one set of (at most six) atomics per block. The box is read with ld.cg so a stale L1 line
// cannot make every block issue redundant atomics.
__device__ __forceinline__ void bbox_accumulate_block(int *bb, int cx, int cy, int cz, bool valid)
{
    __shared__ int s_box[6][32];
    const int big = 0x7fffffff;
    int v[6] = {valid ? cx : big, valid ? cy : big, valid ? cz : big,
                valid ? cx : -big - 1, valid ? cy : -big - 1, valid ? cz : -big - 1};
#pragma unroll
    for (int k = 0; k < 3; ++k) v[k] = __reduce_min_sync(0xffffffffu, v[k]);
#pragma unroll
    for (int k = 3; k < 6; ++k) v[k] = __reduce_max_sync(0xffffffffu, v[k]);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 6; ++k) s_box[k][warp] = v[k];
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < 6; ++k) v[k] = lane < nwarps ? s_box[k][lane] : (k < 3 ? big : -big - 1);
#pragma unroll
        for (int k = 0; k < 3; ++k) v[k] = __reduce_min_sync(0xffffffffu, v[k]);
#pragma unroll
        for (int k = 3; k < 6; ++k) v[k] = __reduce_max_sync(0xffffffffu, v[k]);
        if (lane < 3) {
            if (v[lane] < __ldcg(&bb[lane])) atomicMin(&bb[lane], v[lane]);
        } else if (lane < 6) {
            if (v[lane] > __ldcg(&bb[lane])) atomicMax(&bb[lane], v[lane]);
        }
    }
}

// Barrier-free block-level variant for kernels whose warps finish at very different times: every
// warp folds its reduction into shared memory with atomics and retires; the last warp of the block
// to arrive (shared counter) carries the block's box to global memory. The caller initialises
// s_box / s_done at kernel entry (BboxShared::init, one early barrier while all warps are in step).
struct BboxShared {
    int box[6];
    unsigned int done;
    __device__ __forceinline__ void init()
    {
        if (threadIdx.x < 3) box[threadIdx.x] = 0x7fffffff;
        else if (threadIdx.x < 6) box[threadIdx.x] = -0x7fffffff - 1;
        if (threadIdx.x == 6) done = 0;
        __syncthreads();
    }
};

__device__ __forceinline__ void bbox_accumulate_late(int *bb, BboxShared &sh, int cx, int cy, int cz, bool valid)
{
    const int big = 0x7fffffff;
    int v[6] = {valid ? cx : big, valid ? cy : big, valid ? cz : big,
                valid ? cx : -big - 1, valid ? cy : -big - 1, valid ? cz : -big - 1};
#pragma unroll
    for (int k = 0; k < 3; ++k) v[k] = __reduce_min_sync(0xffffffffu, v[k]);
#pragma unroll
    for (int k = 3; k < 6; ++k) v[k] = __reduce_max_sync(0xffffffffu, v[k]);
    const int lane = threadIdx.x & 31, nwarps = (blockDim.x + 31) >> 5;
    if (lane < 3) atomicMin(&sh.box[lane], v[lane]);
    else if (lane < 6) atomicMax(&sh.box[lane], v[lane]);
    __syncwarp();
    unsigned int last = 0;
    if (lane == 0) {
        __threadfence_block();
        last = atomicAdd(&sh.done, 1u) == (unsigned)(nwarps - 1);
    }
    last = __shfl_sync(0xffffffffu, last, 0);
    if (last) {
        __threadfence_block();
        if (lane < 3) {
            const int b = sh.box[lane];
            if (b < __ldcg(&bb[lane])) atomicMin(&bb[lane], b);
        } else if (lane < 6) {
            const int b = sh.box[lane];
            if (b > __ldcg(&bb[lane])) atomicMax(&bb[lane], b);
        }
    }
}

__global__ void __launch_bounds__(GRID_THREADS)
k_bbox(const float4 *__restrict__ pos, uint32_t n, float h, StepCounters *ctr, int parity)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    bool valid = i < n;
    int cx = 0, cy = 0, cz = 0;
    if (valid) {
        const float4 p = pos[i];
        valid = __float_as_uint(p.w) != W_DROP;
        cx = cell_of(p.x, h); cy = cell_of(p.y, h); cz = cell_of(p.z, h);
    }
    bbox_accumulate_block(ctr->bbox[parity], cx, cy, cz, valid);
}

__global__ void k_reset_bbox(StepCounters *ctr, int parity)
{
    if (threadIdx.x < 3) ctr->bbox[parity][threadIdx.x] = 0x7fffffff;
    else if (threadIdx.x < 6) ctr->bbox[parity][threadIdx.x] = -0x7fffffff - 1;
}

// ---- plan: bounding box -> GridDesc -------------------------------------------------------

// Grid description of a step from the bounding box of occupied cells: one padding layer on every
// side; if the box exceeds the dense-grid allocation the high side of the longest axis is clipped
// (y first: it is the only unbounded axis) and grid_index() clamps the outliers into the last
// interior layer.
__device__ __forceinline__ GridDesc plan_grid(const int *b, uint32_t max_cells, int expand)
{
    // expand: extra cells on every side, for rows the box was not computed from (slab mode: the
    // box comes from the owned rows' integration; ghosts and arrivals sit at most one cell outside).
    long long lo[3] = {(long long)b[0] - expand, (long long)b[1] - expand, (long long)b[2] - expand};
    long long dim[3];
    for (int a = 0; a < 3; ++a) {
        long long ext = (long long)b[3 + a] + expand - lo[a] + 1;  // occupied cells on this axis
        if (ext < 1) ext = 1;                              // empty box (n == 0)
        // A box saturated on two axes (positions at +/-inf or beyond 2^31 h after a blow-up: cell_of saturates)
        // would overflow the 64-bit product below; 2^20 cells per axis is already far beyond any allocation,
        // the halving loop takes it from there.
        if (ext > (1LL << 20)) ext = 1LL << 20;
        dim[a] = ext + 2;
    }
    while (dim[0] * dim[1] * dim[2] > (long long)max_cells) {
        int a = 1;
        if (dim[0] > dim[a]) a = 0;
        if (dim[2] > dim[a]) a = 2;
        dim[a] = max(3LL, (dim[a] + 1) / 2);
        if (dim[0] == 3 && dim[1] == 3 && dim[2] == 3) break;
    }
    GridDesc g;
    g.ox = (int)max(lo[0] - 1, (long long)(-0x7fffffff - 1));
    g.oy = (int)max(lo[1] - 1, (long long)(-0x7fffffff - 1));
    g.oz = (int)max(lo[2] - 1, (long long)(-0x7fffffff - 1));
    g.nx = (int)dim[0]; g.ny = (int)dim[1]; g.nz = (int)dim[2];
    g.ncells = (uint32_t)(dim[0] * dim[1] * dim[2]);
    g.sz = (uint32_t)dim[1];
    g.sx = (uint32_t)(dim[1] * dim[2]);
    g.pad_ = 0;
    return g;
}

// Plan + zero in one launch. Every block derives the same plan from bbox[parity] (filled by the
// previous step's integration, or by k_bbox after an upload) and zeroes its share of
// counts[0 .. ncells] (the extra entry becomes the end sentinel after the scan); block 0 publishes
// the plan, re-arms bbox[parity^1] for this step's integration and resets the step counters.
__global__ void __launch_bounds__(GRID_THREADS)
k_plan_zero(StepCounters *ctr, GridDesc *gd, int parity, uint32_t max_cells, int expand, uint32_t *__restrict__ counts)
{
    pdl_enter();
    __shared__ GridDesc s_g;
    if (threadIdx.x == 0) s_g = plan_grid(ctr->bbox[parity], max_cells, expand);
    __syncthreads();
    const uint32_t n = s_g.ncells + 1;
    const uint32_t n4 = n >> 2;
    uint4 *c4 = reinterpret_cast<uint4 *>(counts);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x)
        c4[i] = make_uint4(0, 0, 0, 0);
    if (blockIdx.x == 0) {
        if (threadIdx.x < (n & 3u)) counts[(n4 << 2) + threadIdx.x] = 0;
        if (threadIdx.x == 0) {
            *gd = s_g;
            ctr->ticket = 0;
            ++ctr->epoch;
            ctr->clamped = 0;
            ctr->heavy[0] = ctr->heavy[1] = 0;
            ctr->clump_tiles[0] = ctr->clump_tiles[1] = 0;
            ctr->clump_rows[0] = ctr->clump_rows[1] = 0;
            ctr->clump_ticket[0] = ctr->clump_ticket[1] = ctr->clump_ticket[2] = 0;
            ctr->fast_x = 0;  // consumed by this step's slab edge scans, set again by its integration
            int *nb = ctr->bbox[parity ^ 1];
            nb[0] = nb[1] = nb[2] = 0x7fffffff;
            nb[3] = nb[4] = nb[5] = -0x7fffffff - 1;
        }
    }
}

// ---- histogram ------------------------------------------------------------------------------

// One thread per particle (array is in last step's cell order, so neighbouring threads hit the
// same or adjacent counters): cell -> grid index -> rank within the cell by atomicAdd.
// ROWS rows per thread (row i, i + blockDim, ...: loads stay coalesced): the kernel waits on the returning atomics, and
// a thread that has issued ROWS of them before it needs the first result keeps ROWS times as many in flight.
template <int ROWS>
__global__ void __launch_bounds__(GRID_THREADS)
k_cell_hist(const float4 *__restrict__ pos, uint32_t n, float h, const GridDesc *__restrict__ gd,
            uint32_t *__restrict__ counts, uint2 *__restrict__ cell_rank, StepCounters *ctr)
{
    pdl_enter();
    const uint32_t i0 = blockIdx.x * (blockDim.x * ROWS) + threadIdx.x;
    if (i0 >= n) return;
    const GridDesc g = *gd;
    float4 p[ROWS];
#pragma unroll
    for (int k = 0; k < ROWS; ++k) {
        const uint32_t i = i0 + k * blockDim.x;
        p[k] = i < n ? pos[i] : make_float4(0.f, 0.f, 0.f, __uint_as_float(W_DROP));
    }
    uint32_t c[ROWS], r[ROWS];
    uint32_t nclamped = 0;
#pragma unroll
    for (int k = 0; k < ROWS; ++k) {
        c[k] = CELL_NONE;  // migrated away / stale ghost: leaves the arrays here
        r[k] = 0u;
        if (__float_as_uint(p[k].w) != W_DROP) {
            bool clamped;
            c[k] = grid_index(g, cell_of(p[k].x, h), cell_of(p[k].y, h), cell_of(p[k].z, h), clamped);
            r[k] = atomicAdd(&counts[c[k]], 1u);
            nclamped += clamped;
        }
    }
#pragma unroll
    for (int k = 0; k < ROWS; ++k) {
        const uint32_t i = i0 + k * blockDim.x;
        if (i < n) cell_rank[i] = make_uint2(c[k], r[k]);
    }
    if (nclamped) atomicAdd(&ctr->clamped, nclamped);
}

// ---- single-pass exclusive scan (decoupled look-back) ---------------------------------------

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 16;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;
constexpr uint32_t SCAN_AGG = 1, SCAN_INCL = 2;

__device__ __forceinline__ unsigned long long scan_pack(uint32_t epoch, uint32_t status, uint32_t v)
{
    return ((unsigned long long)(epoch & 0x3fffffffu) << 34) | ((unsigned long long)status << 32) | v;
}

// In-place exclusive scan of data[0 .. n) where n = *n_ptr + 1. Tiles are handed out by an atomic
// ticket so a tile only ever waits on tiles that are already running; tile_state words carry an
// epoch (device-resident, bumped by whatever arms the scan) so they never need clearing between steps.
__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_exclusive(uint32_t *__restrict__ data, const uint32_t *__restrict__ n_ptr,
                 unsigned long long *tile_state, uint32_t *ticket, const uint32_t *__restrict__ epoch_ptr)
{
    pdl_enter();
    const uint32_t epoch = *epoch_ptr;
    __shared__ uint32_t s_tile, s_excl;
    __shared__ uint32_t s_wsum[SCAN_THREADS / 32];
    const uint32_t n = *n_ptr + 1;
    const uint32_t ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    for (;;) {
        if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
        __syncthreads();
        const uint32_t tile = s_tile;
        if (tile >= ntiles) return;

        const uint32_t base = tile * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
        uint32_t v[SCAN_ITEMS];
        if (base + SCAN_ITEMS <= n) {
#pragma unroll
            for (int q = 0; q < SCAN_ITEMS / 4; ++q) {
                const uint4 a = *reinterpret_cast<const uint4 *>(data + base + 4 * q);
                v[4 * q] = a.x; v[4 * q + 1] = a.y; v[4 * q + 2] = a.z; v[4 * q + 3] = a.w;
            }
        } else {
#pragma unroll
            for (int k = 0; k < SCAN_ITEMS; ++k) v[k] = (base + k < n) ? data[base + k] : 0u;
        }
        uint32_t tsum = 0;
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; ++k) tsum += v[k];
        const uint32_t inc = warp_incl_scan(tsum, lane);
        if (lane == 31) s_wsum[warp] = inc;
        __syncthreads();

        if (warp == 0) {
            const uint32_t w = lane < SCAN_THREADS / 32 ? s_wsum[lane] : 0u;
            const uint32_t winc = warp_incl_scan(w, lane);
            const uint32_t total = __shfl_sync(0xffffffffu, winc, SCAN_THREADS / 32 - 1);
            if (lane < SCAN_THREADS / 32) s_wsum[lane] = winc - w;
            uint32_t excl = 0;
            volatile unsigned long long *st = tile_state;
            if (tile == 0) {
                if (lane == 0) st[0] = scan_pack(epoch, SCAN_INCL, total);
            } else {
                if (lane == 0) st[tile] = scan_pack(epoch, SCAN_AGG, total);
                int look = (int)tile - 1;
                for (;;) {
                    const int idx = look - lane;
                    unsigned long long s = idx >= 0 ? st[idx] : scan_pack(epoch, SCAN_INCL, 0u);
                    const uint32_t status = (uint32_t)(s >> 32) & 3u;
                    const bool ready = ((uint32_t)(s >> 34) == (epoch & 0x3fffffffu)) && status != 0u;
                    if (!__all_sync(0xffffffffu, ready)) continue;
                    const uint32_t incl_mask = __ballot_sync(0xffffffffu, status == SCAN_INCL);
                    const int first = incl_mask ? __ffs(incl_mask) - 1 : 32;
                    uint32_t contrib = lane <= first ? (uint32_t)s : 0u;
                    contrib = __reduce_add_sync(0xffffffffu, contrib);
                    excl += contrib;
                    if (incl_mask) break;
                    look -= 32;
                }
                if (lane == 0) st[tile] = scan_pack(epoch, SCAN_INCL, excl + total);
            }
            if (lane == 0) s_excl = excl;
        }
        __syncthreads();

        uint32_t run = s_excl + s_wsum[warp] + (inc - tsum);
        uint32_t o[SCAN_ITEMS];
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; ++k) { o[k] = run; run += v[k]; }
        if (base + SCAN_ITEMS <= n) {
#pragma unroll
            for (int q = 0; q < SCAN_ITEMS / 4; ++q)
                *reinterpret_cast<uint4 *>(data + base + 4 * q) = make_uint4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
        } else {
#pragma unroll
            for (int k = 0; k < SCAN_ITEMS; ++k)
                if (base + k < n) data[base + k] = o[k];
        }
        __syncthreads();
    }
}

// Arms a scan outside the step (hash16 ordering): fresh ticket, fresh epoch.
__global__ void k_scan_arm(uint32_t *ticket, uint32_t *epoch)
{
    if (threadIdx.x == 0) { *ticket = 0; ++*epoch; }
}

// ---- placement and stable order -------------------------------------------------------------

// slot[start[cell] + rank] = (source row, particle id). Within a cell the rank came from
// atomicAdd, so the order inside a cell segment is arbitrary at this point.
template <int ROWS>
__global__ void __launch_bounds__(GRID_THREADS)
k_place(const uint2 *__restrict__ cell_rank, const float4 *__restrict__ pos, uint32_t n,
        const uint32_t *__restrict__ starts, uint2 *__restrict__ slot, const GridDesc *__restrict__ gd,
        StepCounters *publish_rows, int slab_lo, int slab_hi)
{
    pdl_enter();
    const uint32_t i = blockIdx.x * (blockDim.x * ROWS) + threadIdx.x;
    // slab mode: the rows that survive the build (= the scan's end sentinel) are published for the gather
    // kernel and for the host (this used to be a one-thread kernel of its own), and so is the range of
    // sorted rows that cannot have a ghost among their neighbours: the owned rows of the x-layers
    // slab_lo + 1 .. slab_hi - 2 (x is the slowest index: one contiguous range). The force pass of those
    // rows does not have to wait for the ghosts' densities. Empty when something was clamped into the grid
    // (ghost and owned rows may then share a layer).
    if (publish_rows && i == 0) {
        const GridDesc g = *gd;
        const uint32_t rows = starts[g.ncells];
        publish_rows->aux[2] = rows;
        uint32_t r0 = 0, r1 = 0;
        if (publish_rows->clamped == 0u && (long long)slab_hi - (long long)slab_lo > 2) {
            const long long a = min(max((long long)slab_lo + 1 - g.ox, 0LL), (long long)g.nx);
            const long long b = min(max((long long)slab_hi - 1 - g.ox, 0LL), (long long)g.nx);
            r0 = starts[(uint32_t)a * g.sx];
            r1 = b > a ? starts[(uint32_t)b * g.sx] : r0;
        }
        publish_rows->interior[0] = r0;
        publish_rows->interior[1] = r1;
    }
    uint2 cr[ROWS];
    uint32_t w[ROWS], at[ROWS];
#pragma unroll
    for (int k = 0; k < ROWS; ++k) {
        const uint32_t r = i + k * blockDim.x;
        cr[k] = r < n ? cell_rank[r] : make_uint2(CELL_NONE, 0u);
        w[k] = r < n ? __float_as_uint(pos[r].w) : 0u;
    }
#pragma unroll
    for (int k = 0; k < ROWS; ++k) at[k] = cr[k].x != CELL_NONE ? starts[cr[k].x] + cr[k].y : 0u;
#pragma unroll
    for (int k = 0; k < ROWS; ++k)
        if (cr[k].x != CELL_NONE) slot[at[k]] = make_uint2(i + k * blockDim.x, w[k] & W_ID_MASK);
}

// Canonical order inside every cell segment: ascending particle id (ids are unique). Each slot
// counts the smaller ids of its own segment. Because the order depends on nothing but the ids,
// sums over a cell do not depend on how rows were laid out before the sort — run to run, and
// between a single-GPU run and a slab-decomposed one. inverse (optional) = sorted row of each
// source row.
__global__ void __launch_bounds__(GRID_THREADS)
k_stable_order(const uint2 *__restrict__ slot, const uint2 *__restrict__ cell_rank, uint32_t n_sorted,
               const uint32_t *__restrict__ starts, uint32_t *__restrict__ order, uint32_t *__restrict__ inverse)
{
    const uint32_t d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= n_sorted) return;
    const uint2 me = slot[d];
    const uint32_t c = cell_rank[me.x].x;
    const uint32_t s = starts[c], e = starts[c + 1];
    uint32_t k = 0;
    for (uint32_t t = s; t < e; ++t) k += (slot[t].y < me.y);
    order[s + k] = me.x;
    if (inverse) inverse[me.x] = s + k;
}

// Canonical order + gather in one pass: every slot finds its rank by particle id inside its cell
// segment and moves its row straight to that place (a scatter confined to the cell segment).
// pos.w (identity) rides along; the hash16 of the start-of-step cell goes to its own column.
// FROM_POS: the row's cell is recomputed from its position (the row is gathered anyway, and both gathers are issued
// as soon as the source row is known) instead of being looked up in cell_rank: one dependent load and 8 bytes per
// row less in a kernel that is a chain of dependent loads.
template <bool FROM_POS, int ROWS>
__global__ void __launch_bounds__(GRID_THREADS)
k_order_gather(const uint2 *__restrict__ slot, const uint2 *__restrict__ cell_rank, uint32_t n_bound,
               const uint32_t *__restrict__ n_sorted_dev, const uint32_t *__restrict__ starts, float h,
               const float4 *__restrict__ pos_in, const float4 *__restrict__ vel_in, float4 *__restrict__ pos_out,
               float4 *__restrict__ vel_out, uint32_t *__restrict__ hash_out, uint32_t *__restrict__ inverse,
               const GridDesc *__restrict__ gd)
{
    pdl_enter();
    const uint32_t d0 = blockIdx.x * (blockDim.x * ROWS) + threadIdx.x;
    if (d0 >= n_bound) return;
    // Sync-free slab mode: the host only knows an upper bound of the surviving rows; the exact
    // count is on the device. Rows past it become dropped rows, which every later kernel skips.
    const uint32_t n_sorted = n_sorted_dev ? min(*n_sorted_dev, n_bound) : n_bound;
    bool live[ROWS];
    uint2 me[ROWS];
    float4 p[ROWS], v[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
        const uint32_t d = d0 + r * blockDim.x;
        live[r] = d < n_sorted;
        if (!live[r] && d < n_bound) pos_out[d] = make_float4(0.f, 0.f, 0.f, __uint_as_float(W_DROP));
        me[r] = live[r] ? slot[d] : make_uint2(0u, 0u);
    }
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {  // both gathers are issued as soon as the source row is known
        p[r] = live[r] ? pos_in[me[r].x] : make_float4(0.f, 0.f, 0.f, 0.f);
        v[r] = live[r] ? vel_in[me[r].x] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    uint32_t s[ROWS], e[ROWS], hsh[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
        const int cx = cell_of(p[r].x, h), cy = cell_of(p[r].y, h), cz = cell_of(p[r].z, h);
        hsh[r] = hash16_of(cx, cy, cz);
        uint32_t c = 0;
        if (live[r]) {
            if (FROM_POS) {
                bool clamped;
                c = grid_index(*gd, cx, cy, cz, clamped);  // what k_cell_hist computed for this row
            } else {
                c = cell_rank[me[r].x].x;
            }
        }
        s[r] = live[r] ? starts[c] : 0u;
        e[r] = live[r] ? starts[c + 1] : 0u;
    }
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
        if (!live[r]) continue;
        uint32_t k = s[r];
        for (uint32_t t = s[r]; t < e[r]; ++t) k += (slot[t].y < me[r].y);
        pos_out[k] = p[r];
        vel_out[k] = v[r];
        hash_out[k] = hsh[r];
        if (inverse) inverse[me[r].x] = k;
    }
}

}  // namespace sphb
