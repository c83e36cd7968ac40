/* TEST INFRASTRUCTURE — not product code, never on the product path. See sph_oracle.h.
 *
 * Plain-C restatement of the reference CPU step. Every function names the reference lines it
 * follows; the arithmetic is written in the reference's evaluation order (glm 0.9.9 scalar
 * semantics) and must be compiled with -ffp-contract=off and without
 * -ffast-math. The per-particle loops are split over OpenMP threads the way the reference
 * splits them over std::threads (src/sph.cpp:200-209); no result depends on the split.
 */
#include "sph_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define PI_F 3.14159265f /* src/SPHSystem.cpp:6 */

/* ------------------------------------------------------------------ settings */

/* src/SPHSystem.cpp:8-26. pow(float,int) is the C++11 promoted overload, i.e. pow(double,double);
 * the float*float products are formed first, then everything continues in double and is
 * rounded once when stored to the float member. */
void oracle_make_settings(float mass, float restDensity, float gasConst, float viscosity, float h,
                          float g, float tension, oracle_settings *o)
{
    o->mass = mass;
    o->restDensity = restDensity;
    o->gasConstant = gasConst;
    o->viscosity = viscosity;
    o->h = h;
    o->g = g;
    o->tension = tension;
    o->poly6 = (float)((double)315.0f / ((double)(64.0f * PI_F) * pow((double)h, 9.0)));   /* :19 */
    o->spikyGrad = (float)((double)-45.0f / ((double)PI_F * pow((double)h, 6.0)));         /* :20 */
    o->spikyLap = (float)((double)45.0f / ((double)PI_F * pow((double)h, 6.0)));           /* :21 */
    o->h2 = h * h;                                                                         /* :22 */
    o->selfDens = (float)((double)(mass * o->poly6) * pow((double)h, 6.0));                /* :23 */
    o->massPoly6Product = mass * o->poly6;                                                 /* :24 */
    o->sphereScale = h / 2.f;                                                              /* :25 */
}

/* ------------------------------------------------------------------ cell + hash */

/* src/neighborTable.cpp:14-17: ivec3{x/h, y/h, z/h} — fp32 divide, then float->int truncation. */
void oracle_get_cell(const float *p, float h, int *c)
{
    c[0] = (int)(p[0] / h);
    c[1] = (int)(p[1] / h);
    c[2] = (int)(p[2] / h);
}

/* src/neighborTable.cpp:5-12. The reference multiplies signed ints and relies on wrap-around;
 * unsigned multiplication gives the same bits without undefined behaviour. */
uint32_t oracle_get_hash(const int *c)
{
    return (((uint32_t)c[0] * 73856093u) ^ ((uint32_t)c[1] * 19349663u) ^ ((uint32_t)c[2] * 83492791u))
           % ORACLE_TABLE_SIZE;
}

static inline uint16_t hash16_of(int cx, int cy, int cz)
{
    int c[3] = {cx, cy, cz};
    return (uint16_t)oracle_get_hash(c); /* narrowing: src/Particle.h:9, src/sph.cpp:43,93 */
}

/* src/sph.cpp:17-24 */
void oracle_hashes(uint64_t n, const float *pos, float h, uint16_t *hash)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n; i++) {
        int c[3];
        oracle_get_cell(pos + 3 * i, h, c);
        hash[i] = (uint16_t)oracle_get_hash(c);
    }
}

/* src/sph.cpp:184-192, made deterministic (stable counting sort by hash16). */
void oracle_sort_order(uint64_t n, const uint16_t *hash, uint32_t *order)
{
    uint64_t *start = (uint64_t *)calloc(65537, sizeof(uint64_t));
    for (uint64_t i = 0; i < n; i++) start[(uint32_t)hash[i] + 1]++;
    for (uint32_t b = 0; b < 65536; b++) start[b + 1] += start[b];
    for (uint64_t i = 0; i < n; i++) order[start[hash[i]]++] = (uint32_t)i;
    free(start);
}

/* src/neighborTable.cpp:19-37 */
void oracle_neighbor_table(uint64_t n, const uint16_t *sorted_hash, uint32_t *table)
{
    for (uint32_t i = 0; i < ORACLE_TABLE_SIZE; ++i) table[i] = ORACLE_NO_PARTICLE;
    uint32_t prev = ORACLE_NO_PARTICLE;
    for (uint64_t i = 0; i < n; ++i) {
        uint32_t cur = sorted_hash[i];
        if (cur != prev) {
            table[cur] = (uint32_t)i;
            prev = cur;
        }
    }
}

/* glm::length2(pj - pi): (dx*dx + dy*dy) + dz*dz, each operation rounded to fp32. */
static inline float dist2_of(const float *pj, const float *pi)
{
    float dx = pj[0] - pi[0], dy = pj[1] - pi[1], dz = pj[2] - pi[2];
    float xx = dx * dx, yy = dy * dy, zz = dz * dz;
    return (xx + yy) + zz;
}

/* ------------------------------------------------------------------ density + pressure */

/* src/sph.cpp:28-76 */
void oracle_density_pressure(uint64_t n, const float *pos, const uint16_t *hash, const uint32_t *table,
                             const oracle_settings *s, float *density, float *pressure)
{
    const float mp = s->mass * s->poly6; /* :33 */
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t i = 0; i < (int64_t)n; i++) {
        float dens = 0;
        const float *pi = pos + 3 * i;
        int cell[3];
        oracle_get_cell(pi, s->h, cell); /* :38 */
        for (int x = -1; x <= 1; x++)
            for (int y = -1; y <= 1; y++)
                for (int z = -1; z <= 1; z++) {
                    uint16_t ch = hash16_of(cell[0] + x, cell[1] + y, cell[2] + z); /* :43 */
                    uint32_t j = table[ch];
                    if (j == ORACLE_NO_PARTICLE) continue;
                    while (j < n) {
                        if (j == (uint64_t)i) { j++; continue; } /* :49-52 self skipped by index */
                        if (hash[j] != ch) break;                 /* :54-56 */
                        float d2 = dist2_of(pos + 3 * (uint64_t)j, pi);
                        if (d2 < s->h2) {
                            /* :59-60  float += float * std::pow(float,int): evaluated in double,
                             * rounded to float by the compound assignment. */
                            dens = (float)((double)dens + (double)mp * pow((double)(s->h2 - d2), 3.0));
                        }
                        j++;
                    }
                }
        density[i] = dens + s->selfDens;                                  /* :69 */
        pressure[i] = s->gasConstant * (density[i] - s->restDensity);     /* :72-74 */
    }
}

/* ------------------------------------------------------------------ forces */

/* src/sph.cpp:80-129 */
void oracle_forces(uint64_t n, const float *pos, const float *vel, const float *density,
                   const float *pressure, const uint16_t *hash, const uint32_t *table,
                   const oracle_settings *s, float *force)
{
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t i = 0; i < (int64_t)n; i++) {
        float fx = 0, fy = 0, fz = 0; /* :87 */
        const float *pi = pos + 3 * i;
        const float *vi = vel + 3 * i;
        int cell[3];
        oracle_get_cell(pi, s->h, cell);
        for (int x = -1; x <= 1; x++)
            for (int y = -1; y <= 1; y++)
                for (int z = -1; z <= 1; z++) {
                    uint16_t ch = hash16_of(cell[0] + x, cell[1] + y, cell[2] + z);
                    uint32_t j = table[ch];
                    if (j == ORACLE_NO_PARTICLE) continue;
                    while (j < n) {
                        if (j == (uint64_t)i) { j++; continue; }
                        if (hash[j] != ch) break;
                        const float *pj = pos + 3 * (uint64_t)j;
                        const float *vj = vel + 3 * (uint64_t)j;
                        float d2 = dist2_of(pj, pi);
                        if (d2 < s->h2) {
                            /* :110  ::sqrt(double) on a float, stored to float (== sqrtf) */
                            float dist = (float)sqrt((double)d2);
                            /* :111  normalize(v) = v * (1.0f / sqrtf(dot(v,v))) */
                            float dx = pj[0] - pi[0], dy = pj[1] - pi[1], dz = pj[2] - pi[2];
                            float inv = 1.0f / sqrtf(d2);
                            float nx = dx * inv, ny = dy * inv, nz = dz * inv;
                            /* :114  ((((-dir) * mass) * (p_i + p_j)) / (2 * rho_j)) * spikyGrad */
                            float psum = pressure[i] + pressure[j];
                            float den = 2 * density[j];
                            float px = ((((-nx) * s->mass) * psum) / den) * s->spikyGrad;
                            float py = ((((-ny) * s->mass) * psum) / den) * s->spikyGrad;
                            float pz = ((((-nz) * s->mass) * psum) / den) * s->spikyGrad;
                            /* :115  vec3 *= std::pow(float,int): double pow, cast to float first */
                            float w2 = (float)pow((double)(s->h - dist), 2.0);
                            px *= w2; py *= w2; pz *= w2;
                            fx += px; fy += py; fz += pz; /* :116 */
                            /* :119-120  (((visc*mass) * ((v_j - v_i) / rho_j)) * spikyLap) * (h - dist) */
                            float vm = s->viscosity * s->mass;
                            float hd = s->h - dist;
                            float ux = vj[0] - vi[0], uy = vj[1] - vi[1], uz = vj[2] - vi[2];
                            float qx = ((vm * (ux / density[j])) * s->spikyLap) * hd;
                            float qy = ((vm * (uy / density[j])) * s->spikyLap) * hd;
                            float qz = ((vm * (uz / density[j])) * s->spikyLap) * hd;
                            fx += qx; fy += qy; fz += qz; /* :121 */
                        }
                        j++;
                    }
                }
        force[3 * i] = fx; force[3 * i + 1] = fy; force[3 * i + 2] = fz;
    }
}

/* ------------------------------------------------------------------ integration + walls */

/* src/sph.cpp:133-181 */
void oracle_integrate(uint64_t n, float *pos, float *vel, const float *force, const float *density,
                      const oracle_settings *s, float dt, float *transforms16)
{
    const float boxWidth = 8.f;    /* :139 */
    const float elasticity = 0.5f; /* :140 */
    const float h = s->h;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n; i++) {
        float *p = pos + 3 * i, *v = vel + 3 * i;
        const float *f = force + 3 * i;
        /* :146  force / density + vec3(0, g, 0) */
        float ax = f[0] / density[i] + 0.f;
        float ay = f[1] / density[i] + s->g;
        float az = f[2] / density[i] + 0.f;
        v[0] += ax * dt; v[1] += ay * dt; v[2] += az * dt; /* :147 */
        p[0] += v[0] * dt; p[1] += v[1] * dt; p[2] += v[2] * dt; /* :150 */

        if (p[1] < h) { /* :153-156 */
            p[1] = -p[1] + 2 * h + 0.0001f;
            v[1] = -v[1] * elasticity;
        }
        if (p[0] < h - boxWidth) { /* :158-161 */
            p[0] = -p[0] + 2 * (h - boxWidth) + 0.0001f;
            v[0] = -v[0] * elasticity;
        }
        if (p[0] > -h + boxWidth) { /* :163-166 */
            p[0] = -p[0] + 2 * -(h - boxWidth) - 0.0001f;
            v[0] = -v[0] * elasticity;
        }
        if (p[2] < h - boxWidth) { /* :168-171 */
            p[2] = -p[2] + 2 * (h - boxWidth) + 0.0001f;
            v[2] = -v[2] * elasticity;
        }
        if (p[2] > -h + boxWidth) { /* :173-176 */
            p[2] = -p[2] + 2 * -(h - boxWidth) - 0.0001f;
            v[2] = -v[2] * elasticity;
        }
        if (transforms16) { /* :178-179  translate(position) * scale(h/2), column-major */
            float *m = transforms16 + 16 * i;
            const float sc = s->sphereScale;
            m[0] = sc; m[1] = 0; m[2] = 0; m[3] = 0;
            m[4] = 0; m[5] = sc; m[6] = 0; m[7] = 0;
            m[8] = 0; m[9] = 0; m[10] = sc; m[11] = 0;
            /* last column = col0*0 + col1*0 + col2*0 + (x,y,z,1)*1, summed left to right */
            m[12] = ((0.f + 0.f) + 0.f) + p[0] * 1.f;
            m[13] = ((0.f + 0.f) + 0.f) + p[1] * 1.f;
            m[14] = ((0.f + 0.f) + 0.f) + p[2] * 1.f;
            m[15] = 1.f;
        }
    }
}

/* ------------------------------------------------------------------ one step */

static void gather3(uint64_t n, const uint32_t *order, float *a)
{
    float *t = (float *)malloc(sizeof(float) * 3 * (n ? n : 1));
    for (uint64_t k = 0; k < n; k++) {
        t[3 * k] = a[3 * (uint64_t)order[k]];
        t[3 * k + 1] = a[3 * (uint64_t)order[k] + 1];
        t[3 * k + 2] = a[3 * (uint64_t)order[k] + 2];
    }
    memcpy(a, t, sizeof(float) * 3 * n);
    free(t);
}

/* src/sph.cpp:195-275 */
int oracle_step(uint64_t n, const oracle_settings *s, float dt, float *pos, float *vel, uint32_t *id,
                const uint32_t *order_in, float *force, float *density, float *pressure, uint16_t *hash,
                float *transforms16)
{
    uint64_t m = n ? n : 1;
    uint16_t *h0 = (uint16_t *)malloc(sizeof(uint16_t) * m);
    uint16_t *hs = hash ? hash : (uint16_t *)malloc(sizeof(uint16_t) * m);
    uint32_t *order = (uint32_t *)malloc(sizeof(uint32_t) * m);
    uint32_t *table = (uint32_t *)malloc(sizeof(uint32_t) * ORACLE_TABLE_SIZE);
    float *f = force ? force : (float *)malloc(sizeof(float) * 3 * m);
    float *d = density ? density : (float *)malloc(sizeof(float) * m);
    float *pr = pressure ? pressure : (float *)malloc(sizeof(float) * m);
    int rc = 0;

    oracle_hashes(n, pos, s->h, h0); /* :211-222 */
    if (order_in) memcpy(order, order_in, sizeof(uint32_t) * n);
    else oracle_sort_order(n, h0, order); /* :224-228 */
    for (uint64_t k = 0; k < n; k++) hs[k] = h0[order[k]];
    for (uint64_t k = 1; k < n; k++)
        if (hs[k] < hs[k - 1]) rc = -1;
    if (rc == 0) {
        gather3(n, order, pos);
        gather3(n, order, vel);
        if (id) {
            uint32_t *t = (uint32_t *)malloc(sizeof(uint32_t) * m);
            for (uint64_t k = 0; k < n; k++) t[k] = id[order[k]];
            memcpy(id, t, sizeof(uint32_t) * n);
            free(t);
        }
        oracle_neighbor_table(n, hs, table);                              /* :230-231 */
        oracle_density_pressure(n, pos, hs, table, s, d, pr);             /* :233-245 */
        oracle_forces(n, pos, vel, d, pr, hs, table, s, f);               /* :247-258 */
        oracle_integrate(n, pos, vel, f, d, s, dt, transforms16);         /* :260-272 */
    }
    free(h0);
    free(order);
    free(table);
    if (!hash) free(hs);
    if (!force) free(f);
    if (!density) free(d);
    if (!pressure) free(pr);
    return rc;
}

/* ------------------------------------------------------------------ neighbour multisets */

uint64_t oracle_neighbor_lists(uint64_t n, const float *pos, const uint16_t *hash, const uint32_t *table,
                               const oracle_settings *s, uint32_t *counts, uint32_t *cand,
                               const uint64_t *offsets, uint32_t *list)
{
    uint64_t total = 0;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : total)
    for (int64_t i = 0; i < (int64_t)n; i++) {
        const float *pi = pos + 3 * i;
        int cell[3];
        uint32_t cnt = 0, scanned = 0;
        oracle_get_cell(pi, s->h, cell);
        for (int x = -1; x <= 1; x++)
            for (int y = -1; y <= 1; y++)
                for (int z = -1; z <= 1; z++) {
                    uint16_t ch = hash16_of(cell[0] + x, cell[1] + y, cell[2] + z);
                    uint32_t j = table[ch];
                    if (j == ORACLE_NO_PARTICLE) continue;
                    while (j < n) {
                        if (j == (uint64_t)i) { j++; continue; }
                        if (hash[j] != ch) break;
                        scanned++;
                        if (dist2_of(pos + 3 * (uint64_t)j, pi) < s->h2) {
                            if (list) list[offsets[i] + cnt] = j;
                            cnt++;
                        }
                        j++;
                    }
                }
        if (counts) counts[i] = cnt;
        if (cand) cand[i] = scanned;
        total += cnt;
    }
    return total;
}

/* ------------------------------------------------------------------ initial conditions */

/* jitter of src/SPHSystem.cpp:83-91: (float(rand())/float(RAND_MAX) * 0.5f - 1) * h / 10 */
static inline float jitter(float h)
{
    return (((float)rand() / (float)RAND_MAX) * 0.5f - 1) * h / 10;
}

/* src/SPHSystem.cpp:76-108 */
void oracle_init_cube(int w, const oracle_settings *s, float *pos, float *vel)
{
    srand(1024);
    float sep = s->h + 0.01f;
    for (int i = 0; i < w; i++)
        for (int j = 0; j < w; j++)
            for (int k = 0; k < w; k++) {
                float rx = jitter(s->h), ry = jitter(s->h), rz = jitter(s->h);
                uint64_t idx = (uint64_t)i + ((uint64_t)j + (uint64_t)w * k) * w;
                pos[3 * idx] = i * sep + rx - 1.5f;
                pos[3 * idx + 1] = j * sep + ry + s->h + 0.1f;
                pos[3 * idx + 2] = k * sep + rz - 1.5f;
                vel[3 * idx] = vel[3 * idx + 1] = vel[3 * idx + 2] = 0.f;
            }
}

void oracle_init_block(int nx, int ny, int nz, float sep, float x0, float y0, float z0, float h,
                       unsigned seed, float *pos, float *vel)
{
    srand(seed);
    for (int i = 0; i < nx; i++)
        for (int j = 0; j < ny; j++)
            for (int k = 0; k < nz; k++) {
                float rx = jitter(h), ry = jitter(h), rz = jitter(h);
                uint64_t idx = (uint64_t)i + ((uint64_t)j + (uint64_t)ny * k) * nx;
                pos[3 * idx] = i * sep + rx + x0;
                pos[3 * idx + 1] = j * sep + ry + y0;
                pos[3 * idx + 2] = k * sep + rz + z0;
                vel[3 * idx] = vel[3 * idx + 1] = vel[3 * idx + 2] = 0.f;
            }
}

/* ------------------------------------------------------------------ timing */

double oracle_time_steps(uint64_t n, const oracle_settings *s, float dt, int warmup, int steps,
                         float *pos, float *vel)
{
    struct timespec a, b;
    for (int k = 0; k < warmup; k++) oracle_step(n, s, dt, pos, vel, NULL, NULL, NULL, NULL, NULL, NULL, NULL);
    clock_gettime(CLOCK_MONOTONIC, &a);
    for (int k = 0; k < steps; k++) oracle_step(n, s, dt, pos, vel, NULL, NULL, NULL, NULL, NULL, NULL, NULL);
    clock_gettime(CLOCK_MONOTONIC, &b);
    return (double)(b.tv_sec - a.tv_sec) + 1e-9 * (double)(b.tv_nsec - a.tv_nsec);
}

int oracle_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
