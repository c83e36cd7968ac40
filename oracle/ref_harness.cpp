// TEST INFRASTRUCTURE — not product code, never on the product path.
//
// Headless C-ABI harness around the UNMODIFIED reference sources. The Makefile compiles
// /root/reference/src/{sph,neighborTable,timer,SPHSystem,Geometry}.cpp by path, against the
// stand-in glm / GL headers in oracle/shim/, and links them with this file into
// oracle/_ref/libsph_ref.so. Nothing of the reference is copied into the repository.
//
// What it is for:
//   * pinning oracle/sph_oracle.c (the plain-C restatement) to the real reference code;
//   * minting the golden fixtures under tests/golden/ (tests/golden/make_golden.py);
//   * the "reference" CPU baseline in bench.py (updateParticlesCPU, all host threads).
//
// Particle identity: the reference Particle has no id and std::sort permutes the array
// (src/sph.cpp:184-192). The `acceleration` member is never read or written by the step
// (src/Particle.h:6), so the harness stores the id's bit pattern in acceleration.x and it
// rides through the sort.
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <sstream>

#include <Particle.h>
#include <SPHSystem.h>
#include <neighborTable.h>
#include <sph.h>
#include <kernels/sphGPU.h>

// ---------------------------------------------------------------------------------------
// The reference dispatcher (src/sph.cpp:277-290) needs an updateParticlesGPU symbol. The
// harness forwards it to a hook so that a drop-in replacement (the product's
// updateParticlesGPU shim) can be installed at run time by the GPU parity tests.
// ---------------------------------------------------------------------------------------
typedef void (*gpu_hook_t)(void *particles, void *transforms, size_t n, const void *settings,
                           float dt);
static gpu_hook_t g_gpu_hook = nullptr;

// With -DREF_HARNESS_EXTERNAL_GPU the symbol comes from the product's drop-in
// (sph-fluid-simulator_b200/host/updateParticlesGPU_dropin.cpp) linked into the same library.
#ifndef REF_HARNESS_EXTERNAL_GPU
void updateParticlesGPU(Particle *particles, glm::mat4 *particleTransforms,
                        const size_t particleCount, const SPHSettings &settings,
                        float deltaTime)
{
    if (!g_gpu_hook) {
        std::fprintf(stderr, "ref_harness: updateParticlesGPU called with no hook installed\n");
        std::abort();
    }
    g_gpu_hook(particles, particleTransforms, particleCount, &settings, deltaTime);
}
#endif

namespace {

struct QuietStreams {
    // Timer prints three lines per step to std::cerr (src/timer.cpp:23) and Geometry prints
    // one line to std::cout when the OBJ is missing (src/Geometry.cpp:87).
    std::ios_base::iostate e, o;
    QuietStreams() : e(std::cerr.rdstate()), o(std::cout.rdstate()) {
        std::cerr.setstate(std::ios_base::failbit);
        std::cout.setstate(std::ios_base::failbit);
    }
    ~QuietStreams() { std::cerr.clear(e); std::cout.clear(o); }
};

SPHSettings make(const float *s7) { return SPHSettings(s7[0], s7[1], s7[2], s7[3], s7[4], s7[5], s7[6]); }

Particle *pack(uint64_t n, const float *pos, const float *vel, const uint32_t *id)
{
    Particle *p = (Particle *)std::calloc(n, sizeof(Particle));
    for (uint64_t i = 0; i < n; ++i) {
        p[i].position = glm::vec3(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]);
        p[i].velocity = glm::vec3(vel[3 * i], vel[3 * i + 1], vel[3 * i + 2]);
        uint32_t v = id ? id[i] : (uint32_t)i;
        std::memcpy(&p[i].acceleration.x, &v, 4);
    }
    return p;
}

void unpack(uint64_t n, const Particle *p, float *pos, float *vel, uint32_t *id, float *force,
            float *density, float *pressure, uint16_t *hash)
{
    for (uint64_t i = 0; i < n; ++i) {
        if (pos) { pos[3 * i] = p[i].position.x; pos[3 * i + 1] = p[i].position.y; pos[3 * i + 2] = p[i].position.z; }
        if (vel) { vel[3 * i] = p[i].velocity.x; vel[3 * i + 1] = p[i].velocity.y; vel[3 * i + 2] = p[i].velocity.z; }
        if (force) { force[3 * i] = p[i].force.x; force[3 * i + 1] = p[i].force.y; force[3 * i + 2] = p[i].force.z; }
        if (id) std::memcpy(&id[i], &p[i].acceleration.x, 4);
        if (density) density[i] = p[i].density;
        if (pressure) pressure[i] = p[i].pressure;
        if (hash) hash[i] = p[i].hash;
    }
}

} // namespace

extern "C" {

int ref_sizeof_particle(void) { return (int)sizeof(Particle); }
int ref_sizeof_settings(void) { return (int)sizeof(SPHSettings); }
int ref_hardware_concurrency(void) { return (int)std::thread::hardware_concurrency(); }
uint32_t ref_table_size(void) { return TABLE_SIZE; }

void ref_set_gpu_hook(gpu_hook_t hook) { g_gpu_hook = hook; }

// out13 = poly6, spikyGrad, spikyLap, gasConstant, mass, h2, selfDens, restDensity, viscosity,
//         h, g, tension, massPoly6Product    (src/SPHSystem.h:21-22 member order)
void ref_make_settings(const float *s7, float *out13, float *sphere_scale16)
{
    SPHSettings s = make(s7);
    const float v[13] = {s.poly6, s.spikyGrad, s.spikyLap, s.gasConstant, s.mass, s.h2, s.selfDens,
                         s.restDensity, s.viscosity, s.h, s.g, s.tension, s.massPoly6Product};
    std::memcpy(out13, v, sizeof v);
    if (sphere_scale16) std::memcpy(sphere_scale16, &s.sphereScale, 64);
}

uint32_t ref_get_hash(int cx, int cy, int cz) { return getHash(glm::ivec3(cx, cy, cz)); }

void ref_get_cell(float x, float y, float z, float h, int *out3)
{
    Particle p;
    p.position = glm::vec3(x, y, z);
    glm::ivec3 c = getCell(&p, h);
    out3[0] = c.x; out3[1] = c.y; out3[2] = c.z;
}

// createNeighborTable over an already-sorted hash column (src/neighborTable.cpp:19-37).
void ref_neighbor_table(uint64_t n, const uint16_t *sorted_hash, uint32_t *table_out)
{
    Particle *p = (Particle *)std::calloc(n ? n : 1, sizeof(Particle));
    for (uint64_t i = 0; i < n; ++i) p[i].hash = sorted_hash[i];
    size_t cnt = n;
    uint32_t *t = createNeighborTable(p, cnt);
    std::memcpy(table_out, t, sizeof(uint32_t) * TABLE_SIZE);
    std::free(t);
    std::free(p);
}

// SPHSystem(W, settings, false) -> initParticles() (src/SPHSystem.cpp:28-69, 76-108).
// The object is leaked on purpose: the reference destructor mismatches free/delete[].
void ref_init_cube(int width, const float *s7, float *pos, float *vel, float *transforms16)
{
    QuietStreams q;
    SPHSystem *sys = new SPHSystem((size_t)width, make(s7), false);
    uint64_t n = sys->particleCount;
    unpack(n, sys->particles, pos, vel, nullptr, nullptr, nullptr, nullptr, nullptr);
    (void)transforms16;
}

// SPHSystem::update / reset / startSimulation driven exactly as Tester does
// (src/Tester.cpp:110-125, 210-221): returns positions after `nsteps` update() calls.
// started_first = 0 checks that update() is a no-op until startSimulation().
void ref_class_run(int width, const float *s7, int nsteps, int start, int reset_after, float *pos,
                   float *vel, uint32_t *id)
{
    QuietStreams q;
    SPHSystem *sys = new SPHSystem((size_t)width, make(s7), false);
    for (uint32_t i = 0; i < sys->particleCount; ++i)  // ids ride in the dead acceleration field
        std::memcpy(&sys->particles[i].acceleration.x, &i, 4);
    if (start) sys->startSimulation();
    for (int i = 0; i < nsteps; ++i) sys->update(0.016f);
    if (reset_after) sys->reset();
    unpack(sys->particleCount, sys->particles, pos, vel, id, nullptr, nullptr, nullptr, nullptr);
}

// nsteps calls of updateParticles(..., onGPU) (src/sph.cpp:277-290) from the given state.
// All arrays are in/out in the reference's own post-step order (sorted by start-of-step hash).
void ref_step(uint64_t n, const float *s7, float dt, int nsteps, int on_gpu, float *pos, float *vel,
              uint32_t *id, float *force, float *density, float *pressure, uint16_t *hash,
              float *transforms16)
{
    QuietStreams q;
    SPHSettings s = make(s7);
    Particle *p = pack(n, pos, vel, id);
    glm::mat4 *t = new glm::mat4[n ? n : 1];
    for (int k = 0; k < nsteps; ++k) updateParticles(p, t, n, s, dt, on_gpu != 0);
    unpack(n, p, pos, vel, id, force, density, pressure, hash);
    if (transforms16) std::memcpy(transforms16, t, n * 64);
    delete[] t;
    std::free(p);
}

// Wall-clock seconds for `steps` timed calls of updateParticles(onGPU=false) after `warmup`
// untimed ones; state is advanced in place. This is the reference CPU baseline.
double ref_time_steps(uint64_t n, const float *s7, float dt, int warmup, int steps, float *pos,
                      float *vel)
{
    QuietStreams q;
    SPHSettings s = make(s7);
    Particle *p = pack(n, pos, vel, nullptr);
    glm::mat4 *t = new glm::mat4[n ? n : 1];
    for (int k = 0; k < warmup; ++k) updateParticles(p, t, n, s, dt, false);
    auto t0 = std::chrono::steady_clock::now();
    for (int k = 0; k < steps; ++k) updateParticles(p, t, n, s, dt, false);
    auto t1 = std::chrono::steady_clock::now();
    unpack(n, p, pos, vel, nullptr, nullptr, nullptr, nullptr, nullptr);
    delete[] t;
    std::free(p);
    return std::chrono::duration<double>(t1 - t0).count();
}

} // extern "C"
