// ddiv_pair (prv_kernels.cuh: castRay's two divisions of an axis with ONE refined reciprocal) against __ddiv_rn, bit for bit, on
// the GPU.  TEST INFRASTRUCTURE.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -I nerf-prv_b200/csrc -I include
// Usage: test_ddiv_pair [log2 of the triples per class, default 28]   -> "mismatches: 0" and exit code 0
#include <cstdio>
#include <cstdlib>

#include "prv.h"
#include "prv_kernels.cuh"

__device__ __forceinline__ unsigned long long mix(unsigned long long x) {  // splitmix64
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__device__ __forceinline__ double unit(unsigned long long h) { return (double)(h >> 11) * (1.0 / 9007199254740992.0); }  // [0, 1)

// cls 0: the march's own ranges (b a float direction component, a ~ +-resolution / 2, c a resolution)
// cls 1: raw 64-bit patterns (NaNs, infinities, denormals, zeros, every exponent)
// cls 2: normal mantissas with exponents drawn near the ends of the range (the library's slow-path territory)
__global__ void check(unsigned long long n, int cls, unsigned long long seed, unsigned long long* bad, double* first) {
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long h0 = mix(seed + 3 * i), h1 = mix(seed + 3 * i + 1), h2 = mix(seed + 3 * i + 2);
        double a, b, c;
        if (cls == 0) {
            float dir = (float)(unit(h0) * 2.0 - 1.0);
            if ((h1 & 15) == 0) dir = (float)ldexp(unit(h0) + 0.5, -(int)(h1 >> 60) * 3 - 1) * ((h0 & 1) ? -1.0f : 1.0f);  // small components too
            b = (double)dir;
            const double res = (h1 & 32) ? 0.001 : ((h1 & 64) ? 0.002 : 1.0e-4 + unit(h1) * 0.1);
            a = ((h2 & 1) ? 0.5 : -0.5) * res + (unit(h2) - 0.5) * 1.0e-7 * res;  // voxelBorder - origin: +-res / 2 up to the float rounding of the centre
            c = res;
        } else if (cls == 1) {
            a = __longlong_as_double((long long)h0);
            b = __longlong_as_double((long long)h1);
            c = __longlong_as_double((long long)h2);
        } else {
            const int ra = (int)((h0 >> 52) % 122), rb = (int)((h1 >> 52) % 122);  // exponents within 60 of either end
            a = ldexp(1.0 + unit(h0), ra < 61 ? -1022 + ra : 1023 - (ra - 61)) * ((h0 & 1) ? -1.0 : 1.0);
            b = ldexp(1.0 + unit(h1), rb < 61 ? -1022 + rb : 1023 - (rb - 61)) * ((h1 & 1) ? -1.0 : 1.0);
            c = ldexp(1.0 + unit(h2), (int)((h2 >> 52) % 2040) - 1020);
        }
        double q1, q2;
        prvk::ddiv_pair(a, c, b, q1, q2);
        const double r1 = __ddiv_rn(a, b), r2 = __ddiv_rn(c, fabs(b));
        if (__double_as_longlong(q1) != __double_as_longlong(r1) || __double_as_longlong(q2) != __double_as_longlong(r2)) {
            if (atomicAdd(bad, 1ull) == 0ull) {
                first[0] = a; first[1] = b; first[2] = c; first[3] = q1; first[4] = r1; first[5] = q2; first[6] = r2;
            }
        }
    }
}

int main(int argc, char** argv) {
    const int lg = argc > 1 ? atoi(argv[1]) : 28;
    const unsigned long long n = 1ull << lg;
    unsigned long long* bad;
    double* first;
    cudaMallocManaged(&bad, 8);
    cudaMallocManaged(&first, 7 * 8);
    unsigned long long total_bad = 0;
    for (int cls = 0; cls < 3; cls++) {
        *bad = 0;
        check<<<148 * 8, 256>>>(cls == 0 ? 4 * n : n, cls, 0x1234567ull * (cls + 1), bad, first);
        if (cudaDeviceSynchronize() != cudaSuccess) {
            printf("CUDA error: %s\n", cudaGetErrorString(cudaGetLastError()));
            return 2;
        }
        printf("class %d: %llu triples, %llu differ\n", cls, cls == 0 ? 4 * n : n, *bad);
        if (*bad) printf("  first: a=%a b=%a c=%a  a/b: got %a want %a   c/|b|: got %a want %a\n", first[0], first[1], first[2], first[3], first[4], first[5], first[6]);
        total_bad += *bad;
    }
    printf("mismatches: %llu\n", total_bad);
    return total_bad != 0;
}
