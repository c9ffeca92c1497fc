// (float)sqrt((double)x) == sqrtf(x) for every float x >= 0: the double rounding is innocuous because 53 >= 2 * 24 + 2, so
// ray_init may take octomath::Vector3::norm() -- sqrt in double of the float norm_sq, converted to float -- with ONE
// correctly rounded float square root (__fsqrt_rn).  Exhaustive: all 2^31 non-negative bit patterns up to +inf (2 s on 16 cores).
// Build: g++ -O2 -fopenmp -ffp-contract=off.  Test infrastructure.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
int main() {
    long bad = 0;
#pragma omp parallel for reduction(+ : bad) schedule(static)
    for (int64_t i = 0; i <= 0x7F800000ll; i++) {  // every non-negative float up to +inf
        uint32_t b = (uint32_t)i;
        float x;
        memcpy(&x, &b, 4);
        const float a = (float)std::sqrt((double)x);
        const float c = sqrtf(x);
        uint32_t ba, bc;
        memcpy(&ba, &a, 4);
        memcpy(&bc, &c, 4);
        if (ba != bc) bad++;
    }
    printf("mismatches: %ld of %lld\n", bad, 0x7F800001ll);
    return bad != 0;
}
