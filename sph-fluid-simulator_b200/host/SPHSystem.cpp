// Host-side mirror of the reference's SPHSystem class over the C-ABI. See SPHSystem.h.
#include "SPHSystem.h"

#include <algorithm>
#include <cstdlib>
#include <stdexcept>
#include <string>

namespace sphb200 {

namespace {
void check(int rc, sph_handle *h, const char *what)
{
    if (rc != SPH_OK)
        throw std::runtime_error(std::string(what) + ": " + sph_last_error(h));
}
}  // namespace

SPHSettings::SPHSettings(float mass, float restDensity, float gasConst, float viscosity, float h,
                         float g, float tension)
    : gasConstant(gasConst), mass(mass), restDensity(restDensity), viscosity(viscosity), h(h), g(g),
      tension(tension)
{
    sph_settings s = abi();
    sph_derived d;
    sph_settings_derive(&s, &d);
    poly6 = d.poly6;
    spikyGrad = d.spiky_grad;
    spikyLap = d.spiky_lap;
    h2 = d.h2;
    selfDens = d.self_dens;
    massPoly6Product = d.mass_poly6;
    sphereScale = d.sphere_scale;
}

sph_settings SPHSettings::abi() const
{
    sph_settings s;
    sph_settings_default(&s);  // dt, box, elasticity, wall offset: the reference's buried constants
    s.mass = mass;
    s.rest_density = restDensity;
    s.gas_constant = gasConstant;
    s.viscosity = viscosity;
    s.h = h;
    s.g = g;
    s.tension = tension;
    return s;
}

// ---- scenes -------------------------------------------------------------------------------------

// Jitter of src/SPHSystem.cpp:83-91: (float(rand()) / float(RAND_MAX) * 0.5f - 1) * h / 10.
static inline float jitter(float h)
{
    return (float(std::rand()) / float(RAND_MAX) * 0.5f - 1) * h / 10;
}

void sceneCube(size_t w, float h, float *pos, float *vel)
{
    std::srand(1024);
    const float sep = h + 0.01f;
    for (size_t i = 0; i < w; i++)
        for (size_t j = 0; j < w; j++)
            for (size_t k = 0; k < w; k++) {
                const float rx = jitter(h), ry = jitter(h), rz = jitter(h);
                const size_t idx = i + (j + w * k) * w;
                pos[3 * idx + 0] = int(i) * sep + rx - 1.5f;
                pos[3 * idx + 1] = int(j) * sep + ry + h + 0.1f;
                pos[3 * idx + 2] = int(k) * sep + rz - 1.5f;
                vel[3 * idx + 0] = vel[3 * idx + 1] = vel[3 * idx + 2] = 0.f;
            }
}

void sceneBlock(size_t nx, size_t ny, size_t nz, float sep, float x0, float y0, float z0, float h,
                unsigned seed, float *pos, float *vel)
{
    std::srand(seed);
    for (size_t i = 0; i < nx; i++)
        for (size_t j = 0; j < ny; j++)
            for (size_t k = 0; k < nz; k++) {
                const float rx = jitter(h), ry = jitter(h), rz = jitter(h);
                const size_t idx = i + (j + ny * k) * nx;
                pos[3 * idx + 0] = int(i) * sep + rx + x0;
                pos[3 * idx + 1] = int(j) * sep + ry + y0;
                pos[3 * idx + 2] = int(k) * sep + rz + z0;
                vel[3 * idx + 0] = vel[3 * idx + 1] = vel[3 * idx + 2] = 0.f;
            }
}

void sceneBlockSlice(size_t nx, size_t ny, size_t nz, float sep, float x0, float y0, float z0, float h,
                     unsigned seed, size_t i0, size_t i1, float *pos, float *vel, uint32_t *ids)
{
    // The same lattice, jitter stream and ids as sceneBlock, but only the rows with i in [i0, i1)
    // are stored (one x-range of the lattice per rank). rand() is consumed for the others too so
    // that every rank draws the same jitter for the same particle.
    std::srand(seed);
    const size_t w = i1 - i0;
    for (size_t i = 0; i < nx; i++)
        for (size_t j = 0; j < ny; j++)
            for (size_t k = 0; k < nz; k++) {
                const float rx = jitter(h), ry = jitter(h), rz = jitter(h);
                if (i < i0 || i >= i1) continue;
                const size_t row = (i - i0) + (j + ny * k) * w;
                pos[3 * row + 0] = int(i) * sep + rx + x0;
                pos[3 * row + 1] = int(j) * sep + ry + y0;
                pos[3 * row + 2] = int(k) * sep + rz + z0;
                vel[3 * row + 0] = vel[3 * row + 1] = vel[3 * row + 2] = 0.f;
                ids[row] = (uint32_t)(i + (j + ny * k) * nx);
            }
}

// ---- SPHSystem ----------------------------------------------------------------------------------

SPHSystem::SPHSystem(size_t particleCubeWidth, const SPHSettings &settings, const bool &runOnGPU,
                     int device)
    : settings(settings), particleCubeWidth(particleCubeWidth), started_(false), handle_(nullptr)
{
    if (!runOnGPU)
        throw std::runtime_error("sphb200::SPHSystem has no CPU step: construct with runOnGPU = true");
    particleCount = particleCubeWidth * particleCubeWidth * particleCubeWidth;
    sph_settings s = settings.abi();
    const int rc = sph_create(&s, particleCount ? particleCount : 1, device, &handle_);
    if (rc != SPH_OK) throw std::runtime_error(std::string("sph_create: ") + sph_last_error(nullptr));
    initParticles();
    started_ = false;
}

SPHSystem::~SPHSystem()
{
    if (handle_) sph_destroy(handle_);
}

void SPHSystem::initParticles()
{
    // The cube of src/SPHSystem.cpp:76-108, generated on the device (bit-identical to sceneCube: the tests
    // compare the two), and kept as the reset point so that reset() is a device-to-device copy.
    check(sph_scene_cube_device(handle_, (int)particleCubeWidth), handle_, "sph_scene_cube_device");
    check(sph_set_reset_point(handle_), handle_, "sph_set_reset_point");
}

void SPHSystem::update(float deltaTime)
{
    if (!started_) return;
    // To increase system stability, a fixed deltaTime is set (src/SPHSystem.cpp:112-113).
    deltaTime = 0.003f;
    check(sph_step(handle_, deltaTime, 1), handle_, "sph_step");
}

void SPHSystem::reset()
{
    check(sph_reset(handle_), handle_, "sph_reset");  // src/SPHSystem.cpp:136-139 without the host loop and the upload
    started_ = false;
}

void SPHSystem::startSimulation() { started_ = true; }

const std::vector<float> &SPHSystem::positions()
{
    positions_.resize(4 * particleCount);
    check(sph_read_positions(handle_, positions_.data()), handle_, "sph_read_positions");
    return positions_;
}

const std::vector<float> &SPHSystem::modelMatrices()
{
    matrices_.resize(16 * particleCount);
    check(sph_write_transforms(handle_, matrices_.data()), handle_, "sph_write_transforms");
    return matrices_;
}

void SPHSystem::download(std::vector<float> &pos, std::vector<float> &vel)
{
    pos.resize(3 * particleCount);
    vel.resize(3 * particleCount);
    check(sph_download(handle_, SPH_ORDER_ID, pos.data(), vel.data(), nullptr, nullptr, nullptr, nullptr, nullptr),
          handle_, "sph_download");
}

}  // namespace sphb200

// ---- the class surface through the C-ABI (for non-C++ hosts and the tests) -----------------------

struct sph_system {
    sphb200::SPHSystem *sys;
    std::string err;
};

static thread_local std::string g_system_error;

extern "C" {

int sph_scene_cube(int width, float h, float *host_pos_xyz, float *host_vel_xyz)
{
    if (width < 0 || !host_pos_xyz || !host_vel_xyz) return SPH_ERR_INVALID;
    sphb200::sceneCube((size_t)width, h, host_pos_xyz, host_vel_xyz);
    return SPH_OK;
}

int sph_scene_block(int nx, int ny, int nz, float sep, float x0, float y0, float z0, float h, unsigned seed,
                    float *host_pos_xyz, float *host_vel_xyz)
{
    if (nx < 0 || ny < 0 || nz < 0 || !host_pos_xyz || !host_vel_xyz) return SPH_ERR_INVALID;
    sphb200::sceneBlock((size_t)nx, (size_t)ny, (size_t)nz, sep, x0, y0, z0, h, seed, host_pos_xyz, host_vel_xyz);
    return SPH_OK;
}

int sph_scene_block_slice(int nx, int ny, int nz, float sep, float x0, float y0, float z0, float h, unsigned seed,
                          int i0, int i1, float *host_pos_xyz, float *host_vel_xyz, uint32_t *host_ids)
{
    if (nx < 0 || ny < 0 || nz < 0 || i0 < 0 || i1 < i0 || i1 > nx || !host_pos_xyz || !host_vel_xyz || !host_ids)
        return SPH_ERR_INVALID;
    sphb200::sceneBlockSlice((size_t)nx, (size_t)ny, (size_t)nz, sep, x0, y0, z0, h, seed, (size_t)i0, (size_t)i1,
                             host_pos_xyz, host_vel_xyz, host_ids);
    return SPH_OK;
}

int sph_system_create(int cube_width, const sph_settings *s, int run_on_gpu, int device, sph_system **out)
{
    if (!out || !s || cube_width < 0) return SPH_ERR_INVALID;
    *out = nullptr;
    try {
        sphb200::SPHSettings st(s->mass, s->rest_density, s->gas_constant, s->viscosity, s->h, s->g, s->tension);
        sphb200::SPHSystem *sys = new sphb200::SPHSystem((size_t)cube_width, st, run_on_gpu != 0, device);
        *out = new sph_system{sys, {}};
        return SPH_OK;
    } catch (const std::exception &e) {
        g_system_error = e.what();
        return SPH_ERR_STATE;
    }
}

const char *sph_system_last_error(const sph_system *w) { return w ? w->err.c_str() : g_system_error.c_str(); }

int sph_system_destroy(sph_system *w)
{
    if (!w) return SPH_ERR_INVALID;
    delete w->sys;
    delete w;
    return SPH_OK;
}

#define SYS_TRY(...)                                    \
    if (!w) return SPH_ERR_INVALID;                     \
    try { __VA_ARGS__; return SPH_OK; }                 \
    catch (const std::exception &e) { w->err = e.what(); return SPH_ERR_CUDA; }

int sph_system_start(sph_system *w) { SYS_TRY(w->sys->startSimulation()) }
int sph_system_update(sph_system *w, float dt) { SYS_TRY(w->sys->update(dt)) }
int sph_system_reset(sph_system *w) { SYS_TRY(w->sys->reset()) }
uint64_t sph_system_count(const sph_system *w) { return w ? w->sys->particleCount : 0; }
sph_handle *sph_system_handle(sph_system *w) { return w ? w->sys->handle() : nullptr; }

int sph_system_positions(sph_system *w, float *host_xyzw)
{
    SYS_TRY({
        const std::vector<float> &p = w->sys->positions();
        std::copy(p.begin(), p.end(), host_xyzw);
    })
}

int sph_system_model_matrices(sph_system *w, float *host_mat4)
{
    SYS_TRY({
        const std::vector<float> &m = w->sys->modelMatrices();
        std::copy(m.begin(), m.end(), host_mat4);
    })
}

int sph_system_download(sph_system *w, float *host_pos_xyz, float *host_vel_xyz)
{
    SYS_TRY({
        std::vector<float> p, v;
        w->sys->download(p, v);
        std::copy(p.begin(), p.end(), host_pos_xyz);
        std::copy(v.begin(), v.end(), host_vel_xyz);
    })
}

}  // extern "C"
