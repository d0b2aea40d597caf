// TEST INFRASTRUCTURE ONLY (oracle/): stand-in for libmorton v0.2.12 (pinned by the
// reference at cmake/morton.cmake:9; call sites include/chad/detail/morton.hpp:27,31).
// Its BMI2 path is pure integer: x -> bits 0,3,6,..., y -> bits 1,4,..., z -> bits 2,5,...
#pragma once
#include <cstdint>
#include <immintrin.h>

namespace libmorton {
static constexpr uint64_t kMaskX = 0x9249249249249249ull;
static constexpr uint64_t kMaskY = 0x2492492492492492ull;
static constexpr uint64_t kMaskZ = 0x4924924924924924ull;
inline uint_fast64_t morton3D_64_encode(uint_fast32_t x, uint_fast32_t y, uint_fast32_t z) {
    return _pdep_u64(uint64_t(x), kMaskX) | _pdep_u64(uint64_t(y), kMaskY) | _pdep_u64(uint64_t(z), kMaskZ);
}
inline void morton3D_64_decode(uint_fast64_t m, uint_fast32_t& x, uint_fast32_t& y, uint_fast32_t& z) {
    x = uint_fast32_t(_pext_u64(m, kMaskX));
    y = uint_fast32_t(_pext_u64(m, kMaskY));
    z = uint_fast32_t(_pext_u64(m, kMaskZ));
}
}  // namespace libmorton
