// Stand-in for the reference's precompiled header (P/stdafx.h is UTF-16 and pulls Win32 / D3D11 / ImGui).
#pragma once
#define _CRT_SECURE_NO_WARNINGS
#include <stdio.h>
#include <cstdlib>
#include <cstring>
#include <cfloat>
#include <cmath>
#include <cassert>
#include <cstdint>
#include <string>
#include <chrono>
#include <mutex>
#include <thread>
#include <atomic>
#include <vector>
#include <map>
#include <memory>
#include <random>
#include <functional>
#include <algorithm>
#include <iostream>
#include <embree3/rtcore.h>
#include <imgui.h>
#define _fseeki64 fseeko
#define _ftelli64 ftello
#define sprintf_s snprintf
#define strcpy_s(d, n, s) strncpy(d, s, n)
