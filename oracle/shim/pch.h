/* ORACLE BUILD SHIM (test infrastructure, not product code).
 *
 * Stand-in for the reference's precompiled header (reference: src/pch.h:1-45) so that the
 * reference's own src/scene/WSTessendorf.cpp can be compiled VERBATIM from /root/reference
 * without GLFW / Vulkan / spdlog.  Only what WSTessendorf.{h,cpp} needs is provided:
 * the std containers, glm, and empty logging / assert macros (reference: core/Log.h:50-52,
 * core/Assert.h:15-40).  VKP_PROFILE stays undefined, so VKP_PROFILE_SCOPE() expands to
 * nothing (reference: core/Profile.h:15-32).
 */
#ifndef WSO_ORACLE_SHIM_PCH_H_
#define WSO_ORACLE_SHIM_PCH_H_

#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstdlib>
#include <limits>
#include <memory>
#include <vector>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

#define GLM_FORCE_AVX
#define GLM_FORCE_INLINE
#define GLM_FORCE_RADIANS
#define GLM_FORCE_DEPTH_ZERO_TO_ONE
#include <glm/glm.hpp>
#include <glm/gtc/random.hpp>

#define VKP_REGISTER_FUNCTION()
#define VKP_LOG_INFO(...)
#define VKP_LOG_WARN(...)
#define VKP_LOG_ERR(...)
#define VKP_ASSERT(...)
#define VKP_ASSERT_MSG(...)

#endif
