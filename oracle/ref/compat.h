/*
 * TEST INFRASTRUCTURE (oracle) -- not part of the product.
 *
 * Force-included portability shim (g++ -include) that lets the UNMODIFIED
 * reference sources under /root/reference/src compile with g++/libstdc++.
 * The reference is MSVC-flavoured; these are the five MSVC-isms it relies on
 * (SURVEY.md section 8(c)).  No reference source is copied or edited.
 */
#ifndef GOBLIN_B200_ORACLE_REF_COMPAT_H
#define GOBLIN_B200_ORACLE_REF_COMPAT_H
#ifdef __cplusplus
#include <cmath>
#include <cstdio>
#include <cstdint>
#include <condition_variable> /* GoblinThreadPool.h:33 uses it without the include */
#include <random>
#include <algorithm>

using std::isinf; /* GoblinGeometry.cpp:58 calls unqualified isinf */

typedef int errno_t; /* GoblinImageIO.cpp:103 */
static inline errno_t fopen_s(FILE** f, const char* name, const char* mode) {
    *f = std::fopen(name, mode);
    return *f ? 0 : 1;
}

namespace std {
/* GoblinTexture.cpp:374 calls std::max(double, float) */
inline double max(double a, float b) { return a < (double)b ? (double)b : a; }
/* GoblinUtils.cpp:14-15,21-22 use the pre-standard distribution names */
template <typename T> using uniform_real = uniform_real_distribution<T>;
template <typename T> using uniform_int = uniform_int_distribution<T>;
}
#endif
#endif
