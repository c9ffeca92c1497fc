// prv_keys.hpp -- OctoMap 1.9.6 key arithmetic (OcTreeBaseImpl::coordToKeyChecked / keyToCoord, tree
// depth 16) and the begin_leafs() ordering, shared by host code and device code.
#pragma once
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define PRV_HD __host__ __device__ __forceinline__
#else
#define PRV_HD inline
#endif

namespace prv {

constexpr int kTreeMaxVal = 32768;

// key = (int)floor(resolution_factor * coord) + 32768, valid when in [0, 65536)
PRV_HD bool coord_to_key_checked(double coord, double resolution_factor, uint16_t& key) {
    // the range test is made on the double, so NaN and coordinates beyond the int range are rejected on every platform
    // ((int) of such a double is INT_MIN on x86 -- rejected by the original test too -- but 0 / saturated in CUDA)
    const double f = floor(resolution_factor * coord);
    if (!(f >= -(double)kTreeMaxVal && f < (double)kTreeMaxVal)) return false;
    key = (uint16_t)((int)f + kTreeMaxVal);
    return true;
}

// centre coordinate of a key: (double(key - 32768) + 0.5) * resolution
PRV_HD double key_to_coord(int key, double resolution) { return ((double)(key - kTreeMaxVal) + 0.5) * resolution; }

// spread the 16 bits of v so that bit b lands at bit 3b
PRV_HD uint64_t spread3(uint64_t v) {
    v &= 0xFFFFull;
    v = (v | (v << 16)) & 0x0000FF0000FFull;
    v = (v | (v << 8)) & 0x00F00F00F00Full;
    v = (v | (v << 4)) & 0x0C30C30C30C3ull;
    v = (v | (v << 2)) & 0x249249249249ull;
    return v;
}
PRV_HD uint32_t compact3(uint64_t v) {
    v &= 0x249249249249ull;
    v = (v | (v >> 2)) & 0x0C30C30C30C3ull;
    v = (v | (v >> 4)) & 0x00F00F00F00Full;
    v = (v | (v >> 8)) & 0x0000FF0000FFull;
    v = (v | (v >> 16)) & 0xFFFFull;
    return (uint32_t)v;
}

// begin_leafs() visits children in ascending child index with x = bit 0, y = bit 1, z = bit 2 at every
// level, i.e. ascending 48-bit Morton code with z the most significant bit of each triple.
PRV_HD uint64_t morton_code(uint16_t kx, uint16_t ky, uint16_t kz) {
    return spread3(kx) | (spread3(ky) << 1) | (spread3(kz) << 2);
}
PRV_HD void morton_decode(uint64_t m, uint16_t k[3]) {
    k[0] = (uint16_t)compact3(m);
    k[1] = (uint16_t)compact3(m >> 1);
    k[2] = (uint16_t)compact3(m >> 2);
}

}  // namespace prv
