// TEST INFRASTRUCTURE ONLY: the reference includes glm/gtx/quaternion.hpp (cuda/includes/utils.cu:4)
// but the mesh-generation path only needs the `quat` type, which glm.hpp of this shim provides.
#pragma once
#include "../glm.hpp"
