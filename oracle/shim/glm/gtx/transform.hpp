// TEST INFRASTRUCTURE — stand-in for glm/gtx/transform.hpp (see ../glm.hpp).
// translate(v) = identity with last column (v, 1); scale(s) = diag(s.x, s.y, s.z, 1).
// Both are built the way glm builds them (column sums of the identity), which is exact.
#pragma once
#include "../glm.hpp"

namespace glm {
GLM_SHIM_HD inline mat4 translate(const vec3 &v) {
    mat4 m(1.0f);
    m[3] = m[0] * v.x + m[1] * v.y + m[2] * v.z + m[3];
    return m;
}
GLM_SHIM_HD inline mat4 scale(const vec3 &s) {
    mat4 m(1.0f), r;
    r[0] = m[0] * s.x; r[1] = m[1] * s.y; r[2] = m[2] * s.z; r[3] = m[3];
    return r;
}
} // namespace glm
