// Shared device-side types and exact-arithmetic helpers for the SPH step kernels.
//
// Parity rule for this file: every floating-point operation that decides a neighbour or feeds a
// reference-visible field is written with a round-to-nearest intrinsic (__fmul_rn, __fadd_rn,
// __fdiv_rn, __fsqrt_rn, __dmul_rn, __dadd_rn). Those are never contracted into FMAs, so the
// results match the reference's x86-64 (no-FMA) arithmetic operation for operation. The
// library is additionally compiled with --fmad=false.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace sphb {

// ---- pos.w / vel.w bit lanes --------------------------------------------------------------
// pos.w carries, as raw bits, the particle's identity: [30:0] id, bit 31 "ghost" (a slab halo copy
// owned by another rank: takes part in neighbour sums as j only, is never integrated). The value
// 0xFFFFFFFF marks a dropped row (a migrated particle or last step's ghost) that the next grid
// build skips, which is how rows leave the arrays.
// vel.w carries the hash16 of the particle's start-of-step cell (what Particle::hash holds after
// the reference step), for read-out and for the hash-table parity artefact.
constexpr uint32_t W_ID_MASK = 0x7FFFFFFFu;
constexpr uint32_t W_GHOST = 0x80000000u;
constexpr uint32_t W_DROP = 0xFFFFFFFFu;
constexpr uint32_t W_HASH_MASK = 0xFFFFu;
constexpr uint32_t CELL_NONE = 0xFFFFFFFFu;  // cell_rank.x of a dropped row

// Multipliers of getHash (reference src/neighborTable.cpp:8-10).
constexpr uint32_t HASH_MX = 73856093u, HASH_MY = 19349663u, HASH_MZ = 83492791u;
constexpr uint32_t REF_TABLE_SIZE = 262144u;  // src/neighborTable.h:9
constexpr uint32_t REF_NO_PARTICLE = 0xFFFFFFFFu;

// Dense cell grid of one step. Index = (ix * nz + iz) * ny + iy: y fastest, x slowest, so the
// three y-neighbours of a cell are adjacent in the sorted particle array and an x-slab is one
// contiguous range. Layers 0 and n-1 of every axis are empty padding; particles whose true cell
// lies outside are clamped into layers 1 / n-2 (candidates become a superset, results do not
// change because acceptance is by distance).
struct GridDesc {
    int ox, oy, oz;   // true cell coordinate of grid index 0 on each axis
    int nx, ny, nz;   // cells per axis including padding
    uint32_t ncells;  // nx * ny * nz
    uint32_t sx, sz;  // index strides of x and z (sy == 1)
    uint32_t pad_;
};

// Device-resident mutable counters shared by the kernels of a step.
struct StepCounters {
    int bbox[2][6];       // [parity][minx,miny,minz,maxx,maxy,maxz] of true cells
    uint32_t ticket;      // tile ticket of the scan
    uint32_t clamped;     // particles clamped into the grid this step
    uint32_t aux[4];
    uint32_t heavy[2];    // particles deferred to the warp-cooperative density / force kernels
    uint32_t epoch;       // tag of the current scan's tile states; bumped on the device so a captured step replays
    uint32_t fast_x;      // some particle moved half a cell or more along x in the last integration (slab edge scans)
    uint32_t interior[2]; // slab mode: sorted rows [interior[0], interior[1]) have no ghost among their neighbours
    // clump rows (deferred rows of crowded cells, sph_physics.cuh): [0] 32-row tiles the density pass claimed (the
    // force pass serves the same list), rows served by the tiled phase of each pass, and the tickets the heavy kernels
    // draw work with: [0] density deferral list, [1] force deferral list, [2] tiles in the force pass
    uint32_t clump_tiles[2], clump_rows[2], clump_ticket[3];
};

// Settings + derived constants, passed to kernels by value.
struct Params {
    float h, h2, mass, mass_poly6, self_dens, gas_constant, rest_density;
    float visc_mass;    // settings.viscosity * settings.mass  (src/sph.cpp:120, left-assoc)
    float spiky_grad, spiky_lap, g;
    float h_minus_box;  // settings.h - boxWidth               (src/sph.cpp:158,168)
    float box_minus_h;  // -settings.h + boxWidth              (src/sph.cpp:163,173)
    float two_h;        // 2 * settings.h                      (src/sph.cpp:154)
    float two_hmb;      // 2 * (settings.h - boxWidth)         (src/sph.cpp:159,169)
    float two_nhmb;     // 2 * -(settings.h - boxWidth)        (src/sph.cpp:164,174)
    float wall_offset, elasticity, sphere_scale;
    uint32_t clump_cell;  // rows of a cell with at least this many rows are deferred as clump rows (sph_physics.cuh)
};

// ---- programmatic dependent launch -----------------------------------------------------------
// The kernels of a step are launched with cudaLaunchAttributeProgrammaticStreamSerialization (launch_step
// in sph_api.cu): the next kernel's blocks may become resident while this one drains, and they block here
// until every block of the previous kernel has exited and its writes are visible. First statement of every
// step kernel, executed by every thread before anything else: a block that left without waiting would let
// its grid complete — and release the grid after it — before the grid in front of it has finished.
// Without the launch attribute both instructions do nothing.
__device__ __forceinline__ void pdl_enter()
{
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

// ---- cell and hash (bit-exact targets) ----------------------------------------------------

// getCell, src/neighborTable.cpp:14-17: IEEE fp32 divide, then truncation toward zero.
__device__ __forceinline__ int cell_of(float x, float h) { return __float2int_rz(__fdiv_rn(x, h)); }

// getHash narrowed to the uint16_t every consumer stores it in (src/neighborTable.cpp:5-12,
// src/Particle.h:9): "% 262144" followed by the 16-bit store is "& 0xFFFF".
__device__ __forceinline__ uint32_t hash16_of(int cx, int cy, int cz)
{
    return (((uint32_t)cx * HASH_MX) ^ ((uint32_t)cy * HASH_MY) ^ ((uint32_t)cz * HASH_MZ)) & W_HASH_MASK;
}

// True when two of the 27 offsets around (cx,cy,cz) share a hash16. hash(o) = a[ox]^b[oy]^c[oz]
// with a,b,c the per-axis products, so two offsets collide iff the xor of one per-axis
// difference from each axis is zero; per axis the differences are {0, a0^a1, a1^a2, a0^a2}.
__device__ __forceinline__ bool nbhd_has_duplicate_hash(int cx, int cy, int cz)
{
    uint32_t a[3], b[3], c[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        a[k] = ((uint32_t)(cx + k - 1) * HASH_MX) & W_HASH_MASK;
        b[k] = ((uint32_t)(cy + k - 1) * HASH_MY) & W_HASH_MASK;
        c[k] = ((uint32_t)(cz + k - 1) * HASH_MZ) & W_HASH_MASK;
    }
    const uint32_t da[4] = {0u, a[0] ^ a[1], a[1] ^ a[2], a[0] ^ a[2]};
    const uint32_t db[4] = {0u, b[0] ^ b[1], b[1] ^ b[2], b[0] ^ b[2]};
    const uint32_t dc[4] = {0u, c[0] ^ c[1], c[1] ^ c[2], c[0] ^ c[2]};
    bool dup = false;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (i == 0 && j == 0) continue;  // same x and y offset: z offsets always differ in hash
            const uint32_t t = da[i] ^ db[j];
#pragma unroll
            for (int k = 0; k < 4; ++k) dup |= (t == dc[k]);
        }
    return dup;
}

// How many of the 27 buckets the reference walks for a particle in cell (cx,cy,cz) are the
// bucket `hj` — i.e. how many times a neighbour whose hash16 is hj gets accepted
// (src/sph.cpp:40-65).
__device__ __noinline__ uint32_t bucket_multiplicity(int cx, int cy, int cz, uint32_t hj)
{
    uint32_t m = 0;
    for (int x = -1; x <= 1; ++x)
        for (int y = -1; y <= 1; ++y)
            for (int z = -1; z <= 1; ++z) m += (hash16_of(cx + x, cy + y, cz + z) == hj);
    return m;
}

// Grid index of a true cell, clamped into the non-padding layers.
__device__ __forceinline__ uint32_t grid_index(const GridDesc &g, int cx, int cy, int cz, bool &clamped)
{
    long long ix = (long long)cx - g.ox, iy = (long long)cy - g.oy, iz = (long long)cz - g.oz;
    const long long jx = min(max(ix, 1LL), (long long)g.nx - 2);
    const long long jy = min(max(iy, 1LL), (long long)g.ny - 2);
    const long long jz = min(max(iz, 1LL), (long long)g.nz - 2);
    clamped = (jx != ix) | (jy != iy) | (jz != iz);
    return (uint32_t)((jx * g.nz + jz) * g.ny + jy);
}

// glm::length2(pj - pi): (dx*dx + dy*dy) + dz*dz, every operation rounded (src/sph.cpp:57,107).
__device__ __forceinline__ float dist2_rn(float dx, float dy, float dz)
{
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// ---- packed fp32x2 arithmetic (sm_100a FADD2 / FMUL2 / FFMA2) ------------------------------
// Two IEEE round-to-nearest fp32 operations per instruction on a 64-bit register pair. Each half
// is rounded exactly like the scalar __fadd_rn / __fmul_rn / __fmaf_rn, so results are
// bit-identical to the scalar forms; the point is half the issue slots in the neighbour loops.
typedef unsigned long long f32x2;

__device__ __forceinline__ f32x2 pk2(float lo, float hi)
{
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk2(f32x2 v, float &lo, float &hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b)
{
    f32x2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b)
{
    f32x2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b)
{
    f32x2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c)
{
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

// ---- small warp helpers -------------------------------------------------------------------
__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane)
{
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

}  // namespace sphb
