/* Minimal C host for the C-ABI (include/sph_b200.h): the reference's default scene — a 15^3 cube with the
 * settings of src/Tester.cpp:90-91 — stepped on the GPU, with the renderer's read-out every 100 steps.
 *
 *   make -C sph-fluid-simulator_b200
 *   gcc -std=c99 -O2 -I include examples/step_cube.c -L sph-fluid-simulator_b200/lib -lsph_b200 \
 *       -Wl,-rpath,$PWD/sph-fluid-simulator_b200/lib -o step_cube && ./step_cube
 */
#include <stdio.h>
#include <stdlib.h>

#include "sph_b200.h"

#define CHECK(call)                                                              \
    do {                                                                         \
        int rc_ = (call);                                                        \
        if (rc_ != SPH_OK) {                                                     \
            fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, sph_last_error(h)); \
            return 1;                                                            \
        }                                                                        \
    } while (0)

int main(void)
{
    const int width = 15;
    const uint64_t n = (uint64_t)width * width * width;
    sph_settings s;
    sph_handle *h = NULL;
    sph_settings_default(&s);

    float *pos = malloc(sizeof(float) * 3 * n), *vel = malloc(sizeof(float) * 3 * n);
    float *xyzw = malloc(sizeof(float) * 4 * n);
    if (!pos || !vel || !xyzw) return 1;
    sph_scene_cube(width, s.h, pos, vel); /* SPHSystem::initParticles, src/SPHSystem.cpp:76-108 */

    CHECK(sph_create(&s, n, 0, &h));
    CHECK(sph_upload(h, n, pos, vel, NULL));
    for (int frame = 0; frame < 10; ++frame) {
        sph_stats st;
        CHECK(sph_step(h, 0.f, 100));          /* dt <= 0: the fixed 0.003 of SPHSystem::update */
        CHECK(sph_read_positions(h, xyzw));    /* what the renderer draws */
        CHECK(sph_get_stats(h, &st));
        printf("step %4llu  mean density %.3f  kinetic energy %.4f  first particle (%.3f, %.3f, %.3f)\n",
               (unsigned long long)st.steps, st.mean_density, st.kinetic_energy, xyzw[0], xyzw[1], xyzw[2]);
    }
    sph_destroy(h);
    free(pos); free(vel); free(xyzw);
    return 0;
}
