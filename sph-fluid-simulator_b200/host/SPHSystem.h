// Host-side mirror of the reference's simulation object (reference src/SPHSystem.h:13-57), without
// any OpenGL: same constructor arguments, update(dt) / reset() / startSimulation(), public
// particleCount, and a position / model-matrix read-out for a renderer. The particle state lives
// on the GPU behind the C-ABI (include/sph_b200.h); there is no CPU path.
#pragma once

#include <cstddef>
#include <cstdint>
#include <vector>

#include "../../include/sph_b200.h"

namespace sphb200 {

/// Settings which can alter the SPH simulation (reference src/SPHSystem.h:13-22).
/// Same constructor signature and member names; the derived members are evaluated exactly as
/// the reference constructor evaluates them (src/SPHSystem.cpp:8-26).
struct SPHSettings {
    SPHSettings(float mass, float restDensity, float gasConst, float viscosity, float h, float g,
                float tension);

    float sphereScale;  // diagonal of the reference's glm::scale(vec3(h / 2.f)) matrix
    float poly6, spikyGrad, spikyLap, gasConstant, mass, h2, selfDens, restDensity, viscosity, h, g,
        tension, massPoly6Product;

    /// The plain-C settings block of the C-ABI, with the reference's buried constants.
    sph_settings abi() const;
};

class SPHSystem {
public:
    /// particleCubeWidth^3 particles laid out by initParticles (src/SPHSystem.cpp:76-108).
    /// runOnGPU must be true: this implementation has no CPU step and says so loudly.
    SPHSystem(size_t particleCubeWidth, const SPHSettings &settings, const bool &runOnGPU = true,
              int device = 0);
    ~SPHSystem();
    SPHSystem(const SPHSystem &) = delete;
    SPHSystem &operator=(const SPHSystem &) = delete;

    size_t particleCount;

    /// No-op until startSimulation(); the step is fixed at 0.003 whatever the caller passes
    /// (src/SPHSystem.cpp:110-117).
    void update(float deltaTime);
    /// Re-seeds the cube and stops the simulation (src/SPHSystem.cpp:136-139).
    void reset();
    void startSimulation();  // src/SPHSystem.cpp:141-143

    /// Renderer read-out, replacing the public `particles` array and the sphereModelMtxs the
    /// reference's draw() uploads (src/SPHSystem.cpp:119-134). Rows are in device order; the
    /// two arrays of one call sequence without an update() in between line up.
    const std::vector<float> &positions();      // particleCount * 4 floats: x, y, z, 1
    const std::vector<float> &modelMatrices();  // particleCount * 16 floats, column-major
    /// Full state by particle id (the index initParticles gave the particle).
    void download(std::vector<float> &pos_xyz, std::vector<float> &vel_xyz);

    sph_handle *handle() { return handle_; }
    bool started() const { return started_; }

private:
    void initParticles();

    SPHSettings settings;
    size_t particleCubeWidth;
    bool started_;
    sph_handle *handle_;
    std::vector<float> positions_, matrices_;
};

/// initParticles (src/SPHSystem.cpp:76-108): glibc srand(1024)/rand() lattice with jitter.
void sceneCube(size_t width, float h, float *pos_xyz, float *vel_xyz);
/// The same generator for an nx*ny*nz block with spacing `sep` and origin (x0,y0,z0)
/// (SURVEY.md §8(d) dam-break recipe): loop order x outer, y, z inner; index i+(j+ny*k)*nx.
void sceneBlock(size_t nx, size_t ny, size_t nz, float sep, float x0, float y0, float z0, float h,
                unsigned seed, float *pos_xyz, float *vel_xyz);

/// Rows of sceneBlock with lattice index i in [i0, i1) only, plus their global ids
/// (id = i + (j + ny*k)*nx); used to seed one rank of a slab-decomposed run.
void sceneBlockSlice(size_t nx, size_t ny, size_t nz, float sep, float x0, float y0, float z0, float h,
                     unsigned seed, size_t i0, size_t i1, float *pos_xyz, float *vel_xyz, uint32_t *ids);

}  // namespace sphb200
