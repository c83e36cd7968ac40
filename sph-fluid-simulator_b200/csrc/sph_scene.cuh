// Device-side scene generation: the lattice + jitter of SPHSystem::initParticles (reference
// src/SPHSystem.cpp:76-108) and of its dam-break generalisation, produced on the GPU
// bit-identically to the serial host loop.
//
// The jitter comes from glibc's rand() after srand(seed): the TYPE_3 additive-feedback generator,
//     r[i] = r[i-31] + r[i-3]  (mod 2^32) for i >= 34,   output k = r[k + 344] >> 1,
// with r[0..30] seeded by the Park-Miller steps of __srandom_r and r[31..33] = r[0..2]. The recurrence is linear over Z/2^32, so
// the state at any position follows from the seed state by a polynomial jump (x^n mod x^31 - x^28 - 1):
// the host computes the 31-word state at the start of every chunk of the output stream (glibc_rand_jump in
// sph_api.cu, a few milliseconds for 64 M particles), one thread per chunk then runs the recurrence and
// lays its particles out. rand() is consumed in the reference's loop order (x outer, y, z inner, three
// draws per particle), particle index = i + (j + ny * k) * nx.
#pragma once

#include "sph_device.cuh"

namespace sphb {

constexpr int SCENE_THREADS = 128;
constexpr uint32_t SCENE_CHUNK = 1024;  // particles per thread (3 draws each)

struct SceneDesc {
    int nx, ny, nz;      // lattice
    int i0, i1;          // rows with lattice x-index in [i0, i1) are produced (one x-range per rank)
    float sep, x0, y0, z0, h;
    float y1;            // added after y0 when cube != 0: initParticles' "+ h + 0.1f" is two additions
    int cube;
};

// (float(rand()) / float(RAND_MAX) * 0.5f - 1) * h / 10   (src/SPHSystem.cpp:83-91); float(RAND_MAX) = 2^31
__device__ __forceinline__ float scene_jitter(uint32_t r, float h)
{
    const float u = __fdiv_rn(__int2float_rn((int)r), 2147483648.0f);
    return __fdiv_rn(__fmul_rn(__fsub_rn(__fmul_rn(u, 0.5f), 1.0f), h), 10.0f);
}

__global__ void __launch_bounds__(SCENE_THREADS)
k_scene_block(const uint32_t *__restrict__ chunk_state, unsigned long long n_total, const SceneDesc d,
              float4 *__restrict__ pos, float4 *__restrict__ vel)
{
    __shared__ uint32_t s_r[31][SCENE_THREADS];  // this thread's 31-word window of the recurrence, circular
    const unsigned long long q = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long u0 = q * SCENE_CHUNK;
    if (u0 >= n_total) return;
    for (int k = 0; k < 31; ++k) s_r[k][threadIdx.x] = chunk_state[q * 31 + k];
    int p = 0;  // slot of r[i - 31]
    auto draw = [&]() {
        const int p3 = p + 28 >= 31 ? p + 28 - 31 : p + 28;  // slot of r[i - 3]
        const uint32_t r = s_r[p][threadIdx.x] + s_r[p3][threadIdx.x];
        s_r[p][threadIdx.x] = r;
        p = p + 1 == 31 ? 0 : p + 1;
        return r >> 1;
    };
    const unsigned long long u1 = min(u0 + SCENE_CHUNK, n_total);
    const unsigned long long plane = (unsigned long long)d.ny * d.nz;
    int i = (int)(u0 / plane);
    int j = (int)((u0 - (unsigned long long)i * plane) / d.nz);
    int k = (int)(u0 - (unsigned long long)i * plane - (unsigned long long)j * d.nz);
    const uint32_t w = (uint32_t)(d.i1 - d.i0);
    for (unsigned long long u = u0; u < u1; ++u) {
        const float rx = scene_jitter(draw(), d.h), ry = scene_jitter(draw(), d.h), rz = scene_jitter(draw(), d.h);
        if (i >= d.i0 && i < d.i1) {
            float4 o;
            o.x = __fadd_rn(__fadd_rn(__fmul_rn(__int2float_rn(i), d.sep), rx), d.x0);
            o.y = __fadd_rn(__fadd_rn(__fmul_rn(__int2float_rn(j), d.sep), ry), d.y0);
            if (d.cube) o.y = __fadd_rn(o.y, d.y1);
            o.z = __fadd_rn(__fadd_rn(__fmul_rn(__int2float_rn(k), d.sep), rz), d.z0);
            const uint32_t id = (uint32_t)i + ((uint32_t)j + (uint32_t)d.ny * (uint32_t)k) * (uint32_t)d.nx;
            o.w = __uint_as_float(id);
            const size_t row = (size_t)(i - d.i0) + ((size_t)j + (size_t)d.ny * k) * w;
            pos[row] = o;
            vel[row] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (++k == d.nz) { k = 0; if (++j == d.ny) { j = 0; ++i; } }
    }
}

}  // namespace sphb
