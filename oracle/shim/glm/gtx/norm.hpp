// TEST INFRASTRUCTURE — stand-in for glm/gtx/norm.hpp (see ../glm.hpp).
#pragma once
#include "../glm.hpp"

namespace glm {
GLM_SHIM_HD inline float length2(const vec3 &v) { return dot(v, v); }
} // namespace glm
