#pragma once
#define _USE_MATH_DEFINES
#include <cmath>
// MSVC's <corecrt_math.h> declares the float overloads of the C math functions in the GLOBAL namespace, so the
// unqualified acos(dir.z) of P/SphericalMap.cpp:12 is acos(float) there. libstdc++ puts them into the global namespace
// only through its <math.h> wrapper (using std::acos; ...); without it the call would silently become acos(double).
#include <math.h>
