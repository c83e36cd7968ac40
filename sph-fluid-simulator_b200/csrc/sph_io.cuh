// Layout conversion at the boundary: host-facing xyz / reference-AoS rows <-> device SoA float4,
// renderer read-out, diagnostics. None of this is on the per-step hot path of a resident
// simulation; it is on the path of the stateless drop-in call (sph_update_particles_aos).
#pragma once

#include "sph_device.cuh"

namespace sphb {

constexpr int IO_THREADS = 256;

// xyz triples (+ optional ids) -> float4 rows. pos.w = id bits, vel.w = 0 (density lane).
__global__ void __launch_bounds__(IO_THREADS)
k_import_xyz(const float *__restrict__ pos3, const float *__restrict__ vel3, const uint32_t *__restrict__ ids,
             uint32_t n, float4 *__restrict__ pos, float4 *__restrict__ vel)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    pos[i] = make_float4(pos3[3 * i], pos3[3 * i + 1], pos3[3 * i + 2], __uint_as_float((ids ? ids[i] : i) & W_ID_MASK));
    vel[i] = make_float4(vel3[3 * i], vel3[3 * i + 1], vel3[3 * i + 2], 0.f);
}

__global__ void __launch_bounds__(IO_THREADS)
k_import_xyzw(const float4 *__restrict__ pos_in, const float4 *__restrict__ vel_in, uint32_t n,
              float4 *__restrict__ pos, float4 *__restrict__ vel)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 p = pos_in[i], v = vel_in[i];
    p.w = __uint_as_float(i);
    v.w = 0.f;
    pos[i] = p;
    vel[i] = v;
}

// row_of_dest[id(row)] = row, for SPH_ORDER_ID exports.
__global__ void __launch_bounds__(IO_THREADS)
k_rows_by_id(const float4 *__restrict__ pos, uint32_t n, uint32_t *__restrict__ row_of_dest, uint32_t *bad)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t id = __float_as_uint(pos[i].w);  // a ghost bit makes the id out of range on purpose
    if (id < n) row_of_dest[id] = i;
    else atomicAdd(bad, 1u);
}

// Export the requested fields (null pointers are skipped) with destination row d taken from
// device row map[d] (or d when map is null).
struct ExportPtrs {
    float *pos3, *vel3, *force3, *density, *pressure;
    uint16_t *hash16;
    uint32_t *id;
};

__global__ void __launch_bounds__(IO_THREADS)
k_export(const float4 *__restrict__ pos, const float4 *__restrict__ vel, const float4 *__restrict__ force,
         const uint32_t *__restrict__ hash16, uint32_t n, const uint32_t *__restrict__ map, const Params P,
         ExportPtrs out)
{
    const uint32_t d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= n) return;
    const uint32_t r = map ? map[d] : d;
    const float4 p = pos[r], v = vel[r];
    if (out.pos3) { out.pos3[3 * d] = p.x; out.pos3[3 * d + 1] = p.y; out.pos3[3 * d + 2] = p.z; }
    if (out.vel3) { out.vel3[3 * d] = v.x; out.vel3[3 * d + 1] = v.y; out.vel3[3 * d + 2] = v.z; }
    if (out.force3) {
        const float4 f = force[r];
        out.force3[3 * d] = f.x; out.force3[3 * d + 1] = f.y; out.force3[3 * d + 2] = f.z;
    }
    if (out.density) out.density[d] = v.w;
    if (out.pressure) out.pressure[d] = __fmul_rn(P.gas_constant, __fsub_rn(v.w, P.rest_density));
    if (out.hash16) out.hash16[d] = (uint16_t)(hash16[r] & W_HASH_MASK);
    if (out.id) out.id[d] = __float_as_uint(p.w);  // bit 31 set = ghost row (slab mode)
}

// Renderer read-out: float4 (x, y, z, 1).
__global__ void __launch_bounds__(IO_THREADS)
k_positions_xyz1(const float4 *__restrict__ pos, uint32_t n, float4 *__restrict__ out)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = pos[i];
    out[i] = make_float4(p.x, p.y, p.z, 1.f);
}

// translate(position) * scale(h/2), column-major (src/sph.cpp:178-179). Four float4 stores per
// particle; map selects the device row of each output row (null = identity).
__global__ void __launch_bounds__(IO_THREADS)
k_transforms(const float4 *__restrict__ pos, uint32_t n, const uint32_t *__restrict__ map, float s,
             float4 *__restrict__ out)
{
    const uint32_t d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= n) return;
    const float4 p = pos[map ? map[d] : d];
    float4 *m = out + 4ull * d;
    m[0] = make_float4(s, 0.f, 0.f, 0.f);
    m[1] = make_float4(0.f, s, 0.f, 0.f);
    m[2] = make_float4(0.f, 0.f, s, 0.f);
    m[3] = make_float4(p.x + 0.f, p.y + 0.f, p.z + 0.f, 1.f);  // +0: glm's column sum turns -0 into +0
}

// ---- reference AoS rows (src/Particle.h:4-10), 15 x 32-bit words = 60 bytes -------------------
// words 0-2 position, 3-5 velocity, 6-8 acceleration (dead, carried through), 9-11 force,
// 12 density, 13 pressure, 14 = uint16 hash + 2 bytes padding.
constexpr int AOS_WORDS = 15;

__global__ void __launch_bounds__(IO_THREADS)
k_import_aos(const uint32_t *__restrict__ aos, uint32_t n, float4 *__restrict__ pos, float4 *__restrict__ vel)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t *r = aos + (size_t)AOS_WORDS * i;
    pos[i] = make_float4(__uint_as_float(r[0]), __uint_as_float(r[1]), __uint_as_float(r[2]), __uint_as_float(i));
    vel[i] = make_float4(__uint_as_float(r[3]), __uint_as_float(r[4]), __uint_as_float(r[5]), 0.f);
}

// Output row d = device row map[d]; the dead acceleration words come from the caller's input row
// (id = input row index).
__global__ void __launch_bounds__(IO_THREADS)
k_export_aos(const float4 *__restrict__ pos, const float4 *__restrict__ vel, const float4 *__restrict__ force,
             const uint32_t *__restrict__ hash16, uint32_t n, const uint32_t *__restrict__ map, const Params P,
             const uint32_t *__restrict__ aos_in, uint32_t *__restrict__ aos_out)
{
    const uint32_t d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= n) return;
    const uint32_t r = map ? map[d] : d;
    const float4 p = pos[r], v = vel[r], f = force[r];
    const float rh = v.w;
    const uint32_t id = __float_as_uint(p.w) & W_ID_MASK;
    const uint32_t *in = aos_in + (size_t)AOS_WORDS * id;
    uint32_t *o = aos_out + (size_t)AOS_WORDS * d;
    o[0] = __float_as_uint(p.x); o[1] = __float_as_uint(p.y); o[2] = __float_as_uint(p.z);
    o[3] = __float_as_uint(v.x); o[4] = __float_as_uint(v.y); o[5] = __float_as_uint(v.z);
    o[6] = in[6]; o[7] = in[7]; o[8] = in[8];
    o[9] = __float_as_uint(f.x); o[10] = __float_as_uint(f.y); o[11] = __float_as_uint(f.z);
    o[12] = __float_as_uint(rh);
    o[13] = __float_as_uint(__fmul_rn(P.gas_constant, __fsub_rn(rh, P.rest_density)));
    o[14] = hash16[r] & W_HASH_MASK;
}

// ---- hash16 ordering (order class of the reference's std::sort) and table ----------------------

__global__ void __launch_bounds__(IO_THREADS)
k_hash16_hist(const uint32_t *__restrict__ hash16, uint32_t n, uint32_t *__restrict__ counts,
              uint2 *__restrict__ key_rank)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t k = hash16[i] & W_HASH_MASK;
    const uint32_t r = atomicAdd(&counts[k], 1u);
    if (key_rank) key_rank[i] = make_uint2(k, r);
}

// createNeighborTable (src/neighborTable.cpp:19-37) from the bucket starts: first index of each
// non-empty bucket, NO_PARTICLE elsewhere (including the 196608 slots a uint16 can never reach).
__global__ void __launch_bounds__(IO_THREADS)
k_ref_table(const uint32_t *__restrict__ starts65537, uint32_t *__restrict__ table)
{
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= REF_TABLE_SIZE) return;
    uint32_t v = REF_NO_PARTICLE;
    if (k < 65536u && starts65537[k + 1] > starts65537[k]) v = starts65537[k];
    table[k] = v;
}

// ---- diagnostics ------------------------------------------------------------------------------

struct StatsAccum {
    double sum_rho, sum_v2;
    unsigned long long nan_count, owned;
    uint32_t max_rho_bits;  // densities are positive: uint order == float order
    uint32_t pad_;
};

__global__ void __launch_bounds__(IO_THREADS)
k_stats(const float4 *__restrict__ pos, const float4 *__restrict__ vel, uint32_t n, StatsAccum *acc)
{
    double sr = 0.0, sv = 0.0;
    unsigned long long nn = 0, no = 0;
    float mx = 0.f;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 p = pos[i], v = vel[i];
        if (__float_as_uint(p.w) & W_GHOST) continue;  // ghost or dropped row
        ++no;
        const float r = v.w;
        if (!(isfinite(p.x) && isfinite(p.y) && isfinite(p.z))) ++nn;
        sr += (double)r;
        sv += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z;
        mx = fmaxf(mx, r);
    }
    for (int o = 16; o > 0; o >>= 1) {
        sr += __shfl_xor_sync(0xffffffffu, sr, o);
        sv += __shfl_xor_sync(0xffffffffu, sv, o);
        nn += __shfl_xor_sync(0xffffffffu, nn, o);
        no += __shfl_xor_sync(0xffffffffu, no, o);
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&acc->sum_rho, sr);
        atomicAdd(&acc->sum_v2, sv);
        atomicAdd(&acc->nan_count, nn);
        atomicAdd(&acc->owned, no);
        atomicMax(&acc->max_rho_bits, __float_as_uint(mx));
    }
}

}  // namespace sphb
