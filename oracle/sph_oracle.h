/* TEST INFRASTRUCTURE — not product code, never on the product path.
 *
 * Plain-C restatement of the reference's CPU simulation step (SPHSystem::update ->
 * updateParticles -> updateParticlesCPU), function by function, with the reference file:line each
 * one follows. It is the parity oracle for the CUDA path: only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * Parity status: PINNED. tests/test_oracle_vs_reference.py checks this restatement against the
 * real reference code (oracle/_ref/libsph_ref.so, the unmodified sources compiled headless):
 * cell/hash/table bit-exact, and — when fed the reference's own post-sort particle order —
 * density, pressure, force, position and velocity BIT-EXACT. The reference ships no golden
 * vectors of its own, so the fixtures in tests/golden/ were minted from that
 * reference build by tests/golden/make_golden.py.
 *
 * Arrays are structure-of-arrays with xyz interleaved: pos[3*i+0..2], vel[3*i+0..2].
 */
#ifndef SPH_ORACLE_H
#define SPH_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORACLE_TABLE_SIZE 262144u      /* src/neighborTable.h:9  */
#define ORACLE_NO_PARTICLE 0xFFFFFFFFu /* src/neighborTable.h:10 */

typedef struct oracle_settings {
    /* constructor inputs, src/SPHSystem.h:15-17 */
    float mass, restDensity, gasConstant, viscosity, h, g, tension;
    /* derived, src/SPHSystem.cpp:19-25 */
    float poly6, spikyGrad, spikyLap, h2, selfDens, massPoly6Product;
    float sphereScale; /* diagonal of glm::scale(vec3(h / 2.f)) */
} oracle_settings;

/* src/SPHSystem.cpp:8-26 */
void oracle_make_settings(float mass, float restDensity, float gasConst, float viscosity, float h,
                          float g, float tension, oracle_settings *out);

/* src/neighborTable.cpp:14-17 and :5-12 */
void oracle_get_cell(const float *pos3, float h, int *cell3);
uint32_t oracle_get_hash(const int *cell3);

/* src/sph.cpp:17-24 (the uint16_t narrowing is src/Particle.h:9) */
void oracle_hashes(uint64_t n, const float *pos, float h, uint16_t *hash);

/* src/sph.cpp:184-192. std::sort is unstable, so the reference's intra-bucket order is
 * implementation-defined; the restatement defines it: stable counting sort on hash16.
 * order[k] = source index of the particle that lands at sorted slot k. */
void oracle_sort_order(uint64_t n, const uint16_t *hash, uint32_t *order);

/* src/neighborTable.cpp:19-37 — table has ORACLE_TABLE_SIZE entries. */
void oracle_neighbor_table(uint64_t n, const uint16_t *sorted_hash, uint32_t *table);

/* src/sph.cpp:28-76, :80-129, :133-181 — all take hash-sorted arrays. */
void oracle_density_pressure(uint64_t n, const float *pos, const uint16_t *hash, const uint32_t *table,
                             const oracle_settings *s, float *density, float *pressure);
void oracle_forces(uint64_t n, const float *pos, const float *vel, const float *density,
                   const float *pressure, const uint16_t *hash, const uint32_t *table,
                   const oracle_settings *s, float *force);
void oracle_integrate(uint64_t n, float *pos, float *vel, const float *force, const float *density,
                      const oracle_settings *s, float dt, float *transforms16);

/* One updateParticlesCPU step, src/sph.cpp:195-275. In/out: pos, vel, id. Out (any may be NULL):
 * force, density, pressure, hash, transforms16. Everything comes back in the post-sort order.
 * If `order` is non-NULL it is used as the sort result instead of the stable counting sort
 * (it must arrange the particles in non-decreasing hash16 order; returns -1 if it does not);
 * this is how the tests replay the reference's own std::sort outcome for bit-exact comparison. */
int oracle_step(uint64_t n, const oracle_settings *s, float dt, float *pos, float *vel, uint32_t *id,
                const uint32_t *order, float *force, float *density, float *pressure, uint16_t *hash,
                float *transforms16);

/* The neighbour MULTISET of every particle exactly as the 27-bucket walk of
 * src/sph.cpp:40-65 / :90-126 accepts it (dist2 < h2, self skipped by index), in walk order,
 * expressed as sorted-array indices. counts[i] = accepted entries of particle i (a neighbour
 * reached through k colliding buckets appears k times). cand[i] (optional) = candidates
 * scanned. list may be NULL to only count; otherwise offsets[i] (exclusive prefix of counts,
 * n+1 entries) must be given and list must hold offsets[n] entries. Returns the total. */
uint64_t oracle_neighbor_lists(uint64_t n, const float *pos, const uint16_t *hash, const uint32_t *table,
                               const oracle_settings *s, uint32_t *counts, uint32_t *cand,
                               const uint64_t *offsets, uint32_t *list);

/* src/SPHSystem.cpp:76-108 (glibc srand(1024)/rand()), cube of width^3 particles. */
void oracle_init_cube(int width, const oracle_settings *s, float *pos, float *vel);

/* Generalisation of initParticles to an nx*ny*nz block with lattice spacing `sep` and origin
 * (x0,y0,z0) (SURVEY.md §8(d) scaling recipe): same loop order i(x) outer, j(y), k(z) inner,
 * same jitter formula, index i + (j + ny*k)*nx, srand(seed). */
void oracle_init_block(int nx, int ny, int nz, float sep, float x0, float y0, float z0, float h,
                       unsigned seed, float *pos, float *vel);

/* Wall-clock seconds for `steps` oracle steps after `warmup` untimed ones ("port" CPU baseline). */
double oracle_time_steps(uint64_t n, const oracle_settings *s, float dt, int warmup, int steps,
                         float *pos, float *vel);

int oracle_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
