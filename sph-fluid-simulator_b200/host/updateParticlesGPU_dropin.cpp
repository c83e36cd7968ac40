// Drop-in replacement for the reference's GPU leg, with the reference's own signature
// (src/kernels/sphGPU.h:8-11). Compile this file INSIDE the reference tree (it includes the
// reference's Particle.h / SPHSystem.h) instead of src/kernels/sphGPU.cu, and link libsph_b200.so:
// the dispatcher updateParticles(..., onGPU = true) (src/sph.cpp:277-290) then runs the B200 step.
// INTEGRATION.md shows the CMake lines. It is not part of libsph_b200.so itself, because it
// needs the reference's headers.
//
// Behaviour: same data contract as the reference call — `particles` is updated in place and comes
// back sorted by start-of-step hash, every Particle field filled, `particleTransforms[i]` matching
// `particles[i]` — but with the CPU path's constants (box 8, elasticity 0.5), i.e. the semantics
// of updateParticlesCPU, not of the divergent reference kernel file.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include <Particle.h>
#include <SPHSystem.h>
#include <kernels/sphGPU.h>

#include "sph_b200.h"

static_assert(sizeof(Particle) == 60, "sph_update_particles_aos expects the 60-byte reference Particle");
static_assert(sizeof(glm::mat4) == 64, "column-major 4x4 float matrix expected");

void updateParticlesGPU(Particle *particles, glm::mat4 *particleTransforms, const size_t particleCount,
                        const SPHSettings &settings, float deltaTime)
{
    static sph_handle *handle = nullptr;
    static size_t capacity = 0;
    static sph_settings current;

    sph_settings s;
    sph_settings_default(&s);  // dt, box 8, elasticity 0.5, wall offset 1e-4 (src/sph.cpp:139-154)
    s.mass = settings.mass;
    s.rest_density = settings.restDensity;
    s.gas_constant = settings.gasConstant;
    s.viscosity = settings.viscosity;
    s.h = settings.h;
    s.g = settings.g;
    s.tension = settings.tension;

    int rc = SPH_OK;
    if (!handle || particleCount > capacity) {
        if (handle) sph_destroy(handle);
        handle = nullptr;
        capacity = particleCount ? particleCount : 1;
        rc = sph_create(&s, capacity, 0, &handle);
        current = s;
    } else if (std::memcmp(&current, &s, sizeof s) != 0) {
        rc = sph_set_settings(handle, &s);
        current = s;
    }
    if (rc == SPH_OK)
        rc = sph_update_particles_aos(handle, particles, reinterpret_cast<float *>(particleTransforms), particleCount,
                                      deltaTime);
    if (rc != SPH_OK) {
        // The reference call returns void and has no error channel; there is no CPU fallback.
        std::fprintf(stderr, "updateParticlesGPU (sph_b200): %s\n", sph_last_error(handle));
        std::abort();
    }
}
