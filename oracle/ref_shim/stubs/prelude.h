// Force-included before every reference translation unit: portability shims for MSVC-isms.
#pragma once
#include <cmath>
#include <cstdio>
#include "glm/glm.hpp"
namespace glm {
// P/utils.cpp:248,255 multiply a vec3 by the double result of an unqualified pow()
inline vec3 operator*(const vec3& v, double s) { return v * static_cast<float>(s); }
inline vec3 operator*(double s, const vec3& v) { return v * static_cast<float>(s); }
}
typedef unsigned char BYTE;  // FreeImage / Win32 typedef used by P/Texture.h
struct FIBITMAP;
