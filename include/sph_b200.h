/*
 * sph_b200.h — C-ABI of the B200-native SPH simulation step.
 *
 * Drop-in scope: the SPHSystem::update() hot path of lijenicol/SPH-Fluid-Simulator, i.e. what sits
 * behind
 *     void updateParticles(Particle*, glm::mat4*, size_t, const SPHSettings&, float, bool onGPU)
 *         (reference src/sph.h:19-22, src/sph.cpp:277-290)
 * and its GPU leg
 *     void updateParticlesGPU(Particle*, glm::mat4*, size_t, const SPHSettings&, float)
 *         (reference src/kernels/sphGPU.h:8-11).
 * Semantics are those of the reference's CPU step updateParticlesCPU (src/sph.cpp:195-275):
 * same cell/hash function, same neighbour multisets (including the 16-bit hash-collision
 * double count), same arithmetic order; see DESIGN.md.
 *
 * Conventions
 *   - plain C types only; every function returns an int status (SPH_OK == 0) and never throws;
 *   - a handle is single-owner: calls on one handle must be serialised by the caller;
 *   - sph_step is asynchronous on the handle's internal stream; every call that returns data to
 *     the host synchronises that stream first;
 *   - "host" pointers are ordinary (pageable or pinned) host memory, "dev" pointers are device
 *     memory on the handle's device;
 *   - positions / velocities / forces are xyz-interleaved float triples unless stated otherwise.
 *
 * There is no CPU fallback: if no CUDA device is usable sph_create fails with SPH_ERR_CUDA.
 */
#ifndef SPH_B200_H
#define SPH_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPH_B200_ABI_VERSION 2

enum sph_status {
    SPH_OK = 0,
    SPH_ERR_INVALID = 1,  /* bad argument (null handle, n > capacity, ...) */
    SPH_ERR_CUDA = 2,     /* a CUDA runtime call failed; see sph_last_error */
    SPH_ERR_STATE = 3,    /* call not valid in the current state (e.g. step before upload) */
    SPH_ERR_CAPACITY = 4  /* a caller-provided buffer or the handle's capacity is too small */
};

/* Replaces `struct SPHSettings`'s constructor inputs (reference src/SPHSystem.h:13-22) plus the
 * constants the reference buries in code. */
typedef struct sph_settings {
    float mass;           /* SPHSettings::mass                         default 0.02  (src/Tester.cpp:90) */
    float rest_density;   /* SPHSettings::restDensity                  default 1000 */
    float gas_constant;   /* SPHSettings::gasConstant                  default 1 */
    float viscosity;      /* SPHSettings::viscosity                    default 1.04 */
    float h;              /* SPHSettings::h  (kernel radius = cell size) default 0.15 */
    float g;              /* SPHSettings::g                            default -9.8 */
    float tension;        /* SPHSettings::tension (unused by the step, as in the reference) 0.2 */
    float dt;             /* fixed step of SPHSystem::update (src/SPHSystem.cpp:113)   0.003 */
    float box_half_width; /* boxWidth  (src/sph.cpp:139)                               8 */
    float elasticity;     /* elasticity (src/sph.cpp:140)                              0.5 */
    float wall_offset;    /* the 0.0001f in the wall reflections (src/sph.cpp:154)     1e-4 */
} sph_settings;

/* Replaces the derived members SPHSettings' constructor computes (src/SPHSystem.cpp:19-25),
 * evaluated on the host with the same float/double expression order. */
typedef struct sph_derived {
    float poly6, spiky_grad, spiky_lap, h2, self_dens, mass_poly6, sphere_scale;
} sph_derived;

/* Diagnostics of the most recent step (replaces nothing in the reference). */
typedef struct sph_stats {
    uint64_t count;           /* particles resident */
    uint64_t steps;           /* steps taken since the last upload */
    int32_t grid_origin[3];   /* cell coordinate of grid index (0,0,0) */
    int32_t grid_dim[3];      /* cells per axis of the dense grid used by the last step */
    uint64_t grid_cells;      /* product of grid_dim */
    uint64_t clamped;         /* particles whose true cell fell outside the grid and were clamped */
    uint64_t nan_count;       /* particles with a non-finite position component */
    double mean_density;      /* mean of the last step's density */
    double max_density;
    double kinetic_energy;    /* 0.5 * mass * sum |v|^2 */
    uint64_t deferred_density; /* particles the last step handed to the heavy density kernel (one warp per row, or - rows of
                                  crowded cells - an 8-row tile at a time) */
    uint64_t deferred_forces;  /* ... and to the heavy force kernel */
    uint64_t nlist_rows;       /* rows of the per-particle neighbour list; particles with more neighbours are deferred */
} sph_stats;

typedef struct sph_handle sph_handle;

/* Order of the rows in a download. */
enum sph_order {
    SPH_ORDER_DEVICE = 0, /* the device's cell-sorted order of the last step */
    SPH_ORDER_ID = 1,     /* row k holds the particle whose id is k (ids must be 0..n-1) */
    SPH_ORDER_HASH16 = 2  /* stably sorted by start-of-step hash16: the order class the reference's
                             std::sort leaves the array in (src/sph.cpp:184-192) */
};

/* ---- settings -------------------------------------------------------------------------- */

/* The shipped defaults: SPHSettings(0.02,1000,1,1.04,0.15,-9.8,0.2) (src/Tester.cpp:90),
 * dt 0.003 (src/SPHSystem.cpp:113), box 8 / elasticity 0.5 / offset 1e-4 (src/sph.cpp:139-154). */
int sph_settings_default(sph_settings *out);

/* SPHSettings::SPHSettings (src/SPHSystem.cpp:8-26). */
int sph_settings_derive(const sph_settings *s, sph_derived *out);

/* ---- lifetime -------------------------------------------------------------------------- */

/* Replaces the per-call allocations of updateParticlesGPU (src/kernels/sphGPU.cu:253-299) with
 * persistent device state for up to `capacity` particles on CUDA device `device`. */
int sph_create(const sph_settings *s, uint64_t capacity, int device, sph_handle **out);
int sph_destroy(sph_handle *h);

/* Replace the settings of a live handle (GUI RESET semantics, src/Tester.cpp:169-173). */
int sph_set_settings(sph_handle *h, const sph_settings *s);

/* Human-readable description of the last failure on this handle (or of the last failed
 * sph_create when h is NULL). Never NULL. */
const char *sph_last_error(const sph_handle *h);

/* ---- state in / out -------------------------------------------------------------------- */

/* Load n particles (n <= capacity). id may be NULL (ids become 0..n-1); given ids must be UNIQUE and below
 * 2^31 (bit 31 marks a slab ghost row inside the library; an id with it set is refused with SPH_ERR_INVALID;
 * uniqueness is the caller's contract: rows inside a cell are ordered by id, two equal ids in one cell would
 * lose a row). Replaces the H2D copy of src/kernels/sphGPU.cu:253-255. */
int sph_upload(sph_handle *h, uint64_t n, const float *host_pos_xyz, const float *host_vel_xyz,
               const uint32_t *host_id);
/* Same, from device memory holding float4 rows (xyz + ignored w); ids 0..n-1. */
int sph_upload_device(sph_handle *h, uint64_t n, const void *dev_pos_xyzw, const void *dev_vel_xyzw);

/* Copy out whatever is non-NULL. pos/vel are the current state; force, density, pressure and
 * hash16 are the values the last step computed (start-of-step hash, as Particle::hash holds
 * after updateParticlesCPU). Replaces the D2H copies of src/kernels/sphGPU.cu:304-307. */
int sph_download(sph_handle *h, int order, float *host_pos_xyz, float *host_vel_xyz,
                 float *host_force_xyz, float *host_density, float *host_pressure,
                 uint16_t *host_hash16, uint32_t *host_id);

/* Renderer read-out (what SPHSystem::draw consumes, src/SPHSystem.cpp:119-134), device order.
 * sph_read_positions: n rows of float4 (x, y, z, 1).
 * sph_write_transforms: n column-major 4x4 matrices translate(pos) * scale(h/2)
 *                       (src/sph.cpp:178-179). Destination is HOST memory. */
int sph_read_positions(sph_handle *h, float *host_xyzw);
int sph_write_transforms(sph_handle *h, float *host_mat4);
/* Device-to-device variants for callers that own a mapped/interop buffer on the same device. */
int sph_read_positions_device(sph_handle *h, void *dev_xyzw);
int sph_write_transforms_device(sph_handle *h, void *dev_mat4);

uint64_t sph_count(const sph_handle *h);
uint64_t sph_capacity(const sph_handle *h);

/* ---- the step -------------------------------------------------------------------------- */

/* nsteps steps of updateParticlesCPU semantics with time step dt (src/sph.cpp:195-275). dt <= 0
 * means "use settings.dt", which is what SPHSystem::update does (src/SPHSystem.cpp:113). */
int sph_step(sph_handle *h, float dt, int nsteps);
int sph_sync(sph_handle *h);

/* Neighbour search alone, `repeats` times: hash cells, counting-sort the rows by cell, cell start
 * offsets (replaces parallelCalculateHashes + sortParticles + createNeighborTable,
 * src/sph.cpp:211-231). For the neighbour-search-only microbenchmark of BASELINE.json. */
int sph_neighbor_search(sph_handle *h, int repeats);

/* Stateless form with the exact data contract of updateParticlesGPU
 * (src/kernels/sphGPU.h:8-11): `host_particles` is an array of n 60-byte reference Particle
 * records (src/Particle.h:4-10: position, velocity, acceleration, force, density, pressure,
 * uint16 hash, 2 bytes padding), updated in place and returned sorted by start-of-step hash16;
 * `host_mat4` (may be NULL) receives n column-major transforms in the same order. The
 * `acceleration` field is carried through untouched. */
int sph_update_particles_aos(sph_handle *h, void *host_particles, float *host_mat4, uint64_t n, float dt);

/* ---- parity / diagnostics -------------------------------------------------------------- */

/* The hash16 -> first-index table of createNeighborTable (src/neighborTable.cpp:19-37) for the
 * last step's start-of-step hashes: 262144 entries, 0xFFFFFFFF where empty. */
int sph_hash_table(sph_handle *h, uint32_t *host_table);

/* Neighbour multisets of the CURRENT positions, enumerated by the same traversal code the
 * density / force kernels use. counts[k] (device order) = accepted entries of row k with
 * multiplicity; if list != NULL it receives, for row k, the ids of its neighbours at
 * [offsets[k], offsets[k+1]) where offsets is the exclusive prefix sum of counts (n+1 entries,
 * filled by this call); list_capacity is in entries. ids_out (optional) = id of each row. */
int sph_neighbor_lists(sph_handle *h, uint32_t *host_counts, uint64_t *host_offsets,
                       uint32_t *host_list, uint64_t list_capacity, uint32_t *host_ids_out);

int sph_get_stats(sph_handle *h, sph_stats *out);

/* Work statistics of the last step for the bench's FP32 model: the number of candidate
 * rows the neighbour passes looked at (sum over owned rows of the rows in their 27 cells, the row itself
 * included) and the number of owned rows. Accepted neighbours come from sph_neighbor_lists. */
int sph_candidate_count(sph_handle *h, uint64_t *candidates_out, uint64_t *rows_out);

/* Per-pass device times (replaces the Timer blocks of src/sph.cpp:235,249,262). While enabled,
 * every step records CUDA events between its passes; sph_pass_times synchronises, returns the
 * MEAN milliseconds per step over the steps recorded since the last call (or since enabling) in
 * the order: grid build (hash + sort + cell ranges), density, forces, integrate — and how many
 * steps that was — then starts a new window. At most 16384 steps are kept per window. */
int sph_enable_pass_timing(sph_handle *h, int enable);
int sph_pass_times(sph_handle *h, float *ms4, uint64_t *steps_out);

/* Kernels launched so far by sph_step / sph_update_particles_aos on this handle. */
uint64_t sph_launch_count(const sph_handle *h);

/* Device self-test: the force pass divides with a reciprocal shared between numerators of equal
 * denominator (csrc/sph_physics.cuh, Recip); this checks it bit-for-bit against IEEE division on
 * n pseudo-random pairs and returns the number of mismatches (expected 0). */
int sph_selftest_division(sph_handle *h, uint64_t n, uint64_t seed, uint64_t *mismatches_out);

/* Device self-test: the density pass forms dx, dy and their squares with packed fp32x2 arithmetic
 * (csrc/sph_physics.cuh, row_dist2); this checks it bit-for-bit against the scalar
 * (dx*dx + dy*dy) + dz*dz of glm::length2 (src/sph.cpp:57) on 2n pseudo-random pairs (expected 0). */
int sph_selftest_packed_dist2(sph_handle *h, uint64_t n, uint64_t seed, uint64_t *mismatches_out);

/* Raw CUDA stream of the handle (cudaStream_t as void*), so a caller can order its own work or
 * record events on it. */
void *sph_stream(sph_handle *h);


/* ---- scenes and the SPHSystem class surface ------------------------------------------- */

/* initParticles (src/SPHSystem.cpp:76-108): width^3 lattice, spacing h + 0.01, origin
 * (-1.5, h + 0.1, -1.5), glibc srand(1024)/rand() jitter, zero velocity; row index
 * i + (j + width*k)*width. Host arrays of 3*width^3 floats. */
int sph_scene_cube(int width, float h, float *host_pos_xyz, float *host_vel_xyz);
/* The same generator for an nx*ny*nz block (dam-break scenes). */
int sph_scene_block(int nx, int ny, int nz, float sep, float x0, float y0, float z0, float h, unsigned seed,
                    float *host_pos_xyz, float *host_vel_xyz);

/* Rows i in [i0, i1) of that block and their global ids (one lattice x-range per rank). */
int sph_scene_block_slice(int nx, int ny, int nz, float sep, float x0, float y0, float z0, float h, unsigned seed,
                          int i0, int i1, float *host_pos_xyz, float *host_vel_xyz, uint32_t *host_ids);

/* The same scenes generated ON THE DEVICE straight into the handle's state, bit-identical to the host
 * generators above (the glibc rand() stream is reproduced by a polynomial jump of its linear recurrence, one
 * chunk of the stream per thread): a 64 M-particle lattice takes milliseconds instead of seconds of serial
 * rand() calls, and nothing crosses PCIe. Uses the handle's settings.h. Rows get the lattice ids
 * i + (j + ny*k)*nx; sph_scene_block_device produces the rows with lattice x-index in [i0, i1) (one x-range
 * per rank). Equivalent to generating on the host and calling sph_upload. */
int sph_scene_cube_device(sph_handle *h, int width);
int sph_scene_block_device(sph_handle *h, int nx, int ny, int nz, float sep, float x0, float y0, float z0, unsigned seed,
                           int i0, int i1);

/* Host-only self-test of the rand() jump-ahead behind the device generators (no CUDA call): the first
 * ndraws outputs of glibc's rand() after srand(seed) against the stream rebuilt in chunks of chunk_draws from
 * the jumped states; returns the number of mismatches (expected 0). */
int sph_selftest_glibc_rand(unsigned seed, uint64_t ndraws, uint64_t chunk_draws, uint64_t *mismatches_out);

/* Reset point: sph_set_reset_point keeps a device copy of the current rows; sph_reset restores it (device
 * to device, no host traffic, no synchronisation) and rewinds the step count — what SPHSystem::reset
 * (src/SPHSystem.cpp:136-139) and the GUI's R key (src/Tester.cpp:169-173) need, without regenerating and
 * re-uploading the scene. */
int sph_set_reset_point(sph_handle *h);
int sph_reset(sph_handle *h);

/* class SPHSystem (src/SPHSystem.h:24-57) for hosts that cannot include the C++ header
 * (sph-fluid-simulator_b200/host/SPHSystem.h): constructor, update, reset, startSimulation,
 * particleCount, and the renderer read-out that replaces `particles` / `sphereModelMtxs`.
 * run_on_gpu = 0 fails: there is no CPU step. */
typedef struct sph_system sph_system;
int sph_system_create(int cube_width, const sph_settings *s, int run_on_gpu, int device, sph_system **out);
int sph_system_destroy(sph_system *sys);
const char *sph_system_last_error(const sph_system *sys);
int sph_system_start(sph_system *sys);             /* SPHSystem::startSimulation */
int sph_system_update(sph_system *sys, float dt);  /* SPHSystem::update (no-op until started; dt := 0.003) */
int sph_system_reset(sph_system *sys);             /* SPHSystem::reset */
uint64_t sph_system_count(const sph_system *sys);  /* SPHSystem::particleCount */
sph_handle *sph_system_handle(sph_system *sys);
int sph_system_positions(sph_system *sys, float *host_xyzw);        /* count rows of (x,y,z,1) */
int sph_system_model_matrices(sph_system *sys, float *host_mat4);   /* count column-major mat4 */
int sph_system_download(sph_system *sys, float *host_pos_xyz, float *host_vel_xyz); /* by particle id */


/* ---- slab decomposition (one process per GPU; DESIGN.md "Multi-GPU") -------------------- */
/*
 * The reference has no multi-GPU path; this is new work. The domain is cut along
 * x at cell boundaries: rank r owns the particles whose cell.x (reference getCell) lies in
 * [cuts[r], cuts[r+1]); cuts has world+1 entries, cuts[0] / cuts[world] are treated as -inf / +inf.
 * The library provides the device-side primitives; the exchange itself is the caller's (NCCL
 * through torch.distributed in sph-fluid-simulator_b200/slab.py). Buffers are DEVICE memory.
 * A particle row on the wire is two float4: (x, y, z, id bits) and (vx, vy, vz, 0) = 32 bytes.
 * Per step: count + pack (migrants leave, last step's ghosts are dropped) -> append arrivals ->
 * pack_halo both sides -> append ghosts -> step_density -> pack_halo_density -> exchange ->
 * set_ghost_density -> step_forces.
 */
int sph_slab_enable(sph_handle *h, int enable);
uint64_t sph_slab_owned(const sph_handle *h); /* rows that are not ghosts */
/* counts[r] = live owned rows whose owner under `cuts` is rank r (host arrays). */
int sph_slab_count(sph_handle *h, const int32_t *cuts, int world, uint64_t *host_counts);
/* Copy the rows owned by rank r != self to dev_buf at row row_offsets[r] + k and drop them here. */
int sph_slab_pack(sph_handle *h, const int32_t *cuts, int world, int self, void *dev_buf, const uint64_t *row_offsets);
/* Append rows: kind 0 = owned arrivals, 1 / 2 = ghost batch of side 0 / 1 (one batch per side). */
int sph_slab_append(sph_handle *h, const void *dev_rows, uint64_t nrows, int kind);
/* Halo message of the owned rows in x-cell cell_x; remembers the rows for pack_halo_density. */
int sph_slab_pack_halo(sph_handle *h, int32_t cell_x, int side, void *dev_buf, uint64_t capacity_rows, uint64_t *nrows_out);
/* Grid build over owned + ghost rows, then density (and neighbour lists) of the owned rows. */
int sph_slab_step_density(sph_handle *h);
/* Densities (float) of the rows of halo message `side`, same order; nrows = that message's. */
int sph_slab_pack_halo_density(sph_handle *h, int side, void *dev_buf);
/* Densities for ghost batch `side`, in the order its rows were appended. */
int sph_slab_set_ghost_density(sph_handle *h, int side, const void *dev_buf, uint64_t nrows);
/* Forces + integration of the owned rows. */
int sph_slab_step_forces(sph_handle *h, float dt);
/* Read back the owned rows only (ghost and dropped rows skipped), compacted, in arbitrary order,
 * with their ids; host buffers hold capacity_rows rows. */
int sph_slab_download_owned(sph_handle *h, float *host_pos_xyz, float *host_vel_xyz, uint32_t *host_id,
                            uint64_t capacity_rows, uint64_t *count_out);
/* Histogram of cell.x over owned rows, bins [x_cell_lo, x_cell_lo + nbins), ends clamped. */
int sph_slab_xcell_histogram(sph_handle *h, int32_t x_cell_lo, uint32_t nbins, uint64_t *host_hist);


/* Sync-free variant of the same step for steady state: messages have a FIXED capacity and go
 * only to the two adjacent ranks (NULL pointer = no neighbour on that side), unused message rows
 * are dropped-row patterns (0xFF), and every count stays on the device, so the host enqueues a
 * whole step without waiting for the GPU. The exact row count and any violation (message
 * overflow, a particle that would have to travel further than the adjacent slab) are read back
 * one step late by the next sph_slab_fast_begin, which then fails with SPH_ERR_CAPACITY: use the
 * general path above for the step after cuts change.
 * Order: fast_begin -> exchange migrants -> fast_arrivals -> fast_halo -> exchange -> fast_ghosts
 * -> sph_slab_step_density -> fast_pack_density -> exchange -> fast_set_ghost_density ->
 * sph_slab_step_forces. Row messages hold cap_rows x 32 bytes, density messages cap_rows floats. */
int sph_slab_fast_begin(sph_handle *h, int32_t lo, int32_t hi, int32_t lo_prev, int32_t hi_next, uint64_t cap_rows,
                        void *dev_send_left, void *dev_send_right);
int sph_slab_fast_arrivals(sph_handle *h, const void *dev_recv_left, const void *dev_recv_right, uint64_t cap_rows);
int sph_slab_fast_halo(sph_handle *h, int32_t lo, int32_t hi, uint64_t cap_rows, void *dev_send_left, void *dev_send_right);
int sph_slab_fast_ghosts(sph_handle *h, const void *dev_recv_left, const void *dev_recv_right, uint64_t cap_rows);
int sph_slab_fast_pack_density(sph_handle *h, uint64_t cap_rows, void *dev_send_left, void *dev_send_right);
int sph_slab_fast_set_ghost_density(sph_handle *h, const void *dev_recv_left, const void *dev_recv_right, uint64_t cap_rows);


/* Peer-memory variant of the sync-free step (one box, NVLink/NVSwitch): instead of packing into a
 * send buffer and handing it to NCCL, ONE pack kernel stores migrants and halo rows straight into the
 * adjacent ranks' mailboxes (device allocations exported with CUDA IPC); its last block publishes the row
 * counts and raises an epoch flag in the peers' memory, and the receiving kernel (one launch appends both
 * neighbours' arrivals and ghosts) spins on its own flag first. A migrant that lands in the neighbour's
 * boundary layer stays with the sender as a ghost (the neighbour would send it straight back), so the halo
 * messages do not wait for the migrants: one exchange round for rows, one for densities. No collective
 * library is involved in the step at all. Slabs must be at least three x-cells wide.
 * Setup: every rank calls sph_slab_p2p_create (64-byte IPC handle out), the handles are exchanged by
 * the caller, then sph_slab_p2p_connect(side 0 = left neighbour, 1 = right) for each neighbour.
 * Step: p2p_begin -> p2p_arrivals -> sph_slab_step_density -> p2p_density -> sph_slab_step_forces; all
 * ranks must step in lockstep. migrant_rows = capacity of this step's migrant messages (0 or more than
 * the mailbox holds: the mailbox's): a few thousand in steady state, a whole layer right after cuts moved. */
int sph_slab_p2p_create(sph_handle *h, uint64_t halo_rows, uint64_t migrant_rows, void *ipc_handle_out64);
int sph_slab_p2p_connect(sph_handle *h, int side, const void *peer_ipc_handle64);
int sph_slab_p2p_begin(sph_handle *h, int32_t lo, int32_t hi, int32_t lo_prev, int32_t hi_next, uint64_t migrant_rows);
int sph_slab_p2p_arrivals(sph_handle *h);
int sph_slab_p2p_density(sph_handle *h);

#ifdef __cplusplus
}
#endif
#endif /* SPH_B200_H */
