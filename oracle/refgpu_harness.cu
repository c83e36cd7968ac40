// TEST INFRASTRUCTURE — not product code, never on the product path.
//
// Timing harness around the reference's OWN CUDA step (src/kernels/sphGPU.cu, compiled unmodified
// by path for sm_100a): the "reference GPU kernel on the same box" of SURVEY.md §8(f). It is a
// secondary baseline only — that file is not result-compatible with the CPU oracle (box 10,
// elasticity 1, 32-neighbour cap without an overflow guard, off-by-one bounds), so it is only ever
// run on sparse states, in a subprocess, and its output is never compared with anything.
#include <chrono>
#include <cstdint>
#include <cstdlib>
#include <cstring>

#include <Particle.h>
#include <SPHSystem.h>
#include <kernels/sphGPU.h>

extern "C" double refgpu_time_steps(uint64_t n, const float *s7, float dt, int warmup, int steps, const float *pos,
                                    const float *vel)
{
    SPHSettings s(s7[0], s7[1], s7[2], s7[3], s7[4], s7[5], s7[6]);
    Particle *p = (Particle *)std::calloc(n + 1, sizeof(Particle));
    glm::mat4 *t = (glm::mat4 *)std::calloc(n + 1, sizeof(glm::mat4));
    for (uint64_t i = 0; i < n; ++i) {
        p[i].position = glm::vec3(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]);
        p[i].velocity = glm::vec3(vel[3 * i], vel[3 * i + 1], vel[3 * i + 2]);
    }
    for (int k = 0; k < warmup; ++k) updateParticlesGPU(p, t, n, s, dt);
    auto t0 = std::chrono::steady_clock::now();
    for (int k = 0; k < steps; ++k) updateParticlesGPU(p, t, n, s, dt);
    auto t1 = std::chrono::steady_clock::now();
    std::free(p);
    std::free(t);
    return std::chrono::duration<double>(t1 - t0).count();
}
