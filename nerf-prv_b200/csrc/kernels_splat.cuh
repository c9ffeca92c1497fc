// kernels_splat.cuh -- splat z-buffer render
// Part of the single translation unit prv_device.cu (included there, in order); see DESIGN.md section 4.
#pragma once

// splat z-buffer -------------------------------------------------------------------------------------
// Stage 1: one 64-bit atomicMin per point on the CORNER cell of its footprint; stage 2 takes the min over the
// point_size x point_size corner cells that cover a pixel.  min is associative, so this equals point_size^2 atomics
// per point on the pixels themselves.
__global__ void __launch_bounds__(256) splat_points_kernel(const float* xyz, uint64_t P, DevCam cam, const ViewConst* views,
                                                           uint32_t view_base, float focal, int point_size, unsigned long long* corner,
                                                           int Wc, int Hc) {
    const uint32_t view_local = blockIdx.y;
    const ViewConst& vc = views[view_local + view_base];
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const double x = (double)xyz[3 * i + 0], y = (double)xyz[3 * i + 1], z = (double)xyz[3 * i + 2];
    const float xc = (float)row_apply(vc.inv + 0, x, y, z);
    const float yc = (float)row_apply(vc.inv + 4, x, y, z);
    const float zc = (float)row_apply(vc.inv + 8, x, y, z);
    if (!(zc > 0.01f && zc < 1000.01f)) return;
    const float u = fadd(fmul(fdiv(xc, zc), focal), fmul((float)cam.W, 0.5f));
    const float v = fadd(fmul(fdiv(yc, zc), focal), fmul((float)cam.H, 0.5f));
    if (!(u > -64.0f && u < (float)cam.W + 64.0f && v > -64.0f && v < (float)cam.H + 64.0f)) return;
    const float off = fsub(0.5f, fmul(0.5f, (float)point_size));
    const int lx = (int)floorf(fadd(u, off)), ly = (int)floorf(fadd(v, off));
    const int cx = lx + point_size - 1, cy = ly + point_size - 1;
    if (cx < 0 || cx >= Wc || cy < 0 || cy >= Hc) return;
    const unsigned long long packed = ((unsigned long long)__float_as_uint(zc) << 32) | (unsigned long long)(uint32_t)i;
    atomicMin(corner + ((size_t)view_local * Hc + cy) * Wc + cx, packed);
}

// The window minimum is separable: a pass over the rows of the shared-memory tile (min of point_size horizontal neighbours,
// (8 + s - 1) x 32 values per block) and then, per pixel, the min of point_size of those -- 2 s shared-memory reads per pixel
// instead of s^2 (25 for the reference's point_size 5; the s^2 version ran at 0.2 of the HBM roofline, bound by LDS.64 +
// 64-bit compares).  Dynamic shared memory: tile (32 + s - 1) x (8 + s - 1) u64, then the row minima 32 x (8 + s - 1) u64.
__global__ void __launch_bounds__(256) splat_resolve_kernel(const unsigned long long* corner, int Wc, int Hc, int W, int H, int point_size,
                                                            const uint8_t* rgb, uint8_t* rgba, float* depth, uint32_t view_base) {
    extern __shared__ unsigned long long s_tile[];
    const uint32_t view_local = blockIdx.z;
    const int tw = 32 + point_size - 1, th = 8 + point_size - 1;
    unsigned long long* s_hmin = s_tile + tw * th;
    const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 8;
    const unsigned long long* src = corner + (size_t)view_local * Hc * Wc;
    // flat tile index t -> (row, column) carried incrementally: one division per thread instead of one per element
    {
        const int dr = (int)blockDim.x / tw, dc = (int)blockDim.x % tw;
        int r = (int)threadIdx.x / tw, c = (int)threadIdx.x % tw;
        for (int t = threadIdx.x; t < tw * th; t += blockDim.x) {
            const int cx = x0 + c, cy = y0 + r;
            s_tile[t] = (cx < Wc && cy < Hc) ? src[(size_t)cy * Wc + cx] : ~0ull;
            c += dc;
            r += dr;
            if (c >= tw) {
                c -= tw;
                r++;
            }
        }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < 32 * th; t += blockDim.x) {
        const int c = t & 31, r = t >> 5;
        unsigned long long m = ~0ull;
        for (int dx = 0; dx < point_size; dx++) m = min(m, s_tile[r * tw + c + dx]);
        s_hmin[t] = m;
    }
    __syncthreads();
    const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
    const int x = x0 + lx, y = y0 + ly;
    if (x >= W || y >= H) return;
    unsigned long long best = ~0ull;
    for (int dy = 0; dy < point_size; dy++) best = min(best, s_hmin[(ly + dy) * 32 + lx]);
    const size_t pix = ((size_t)(view_local + view_base) * H + y) * W + x;
    uchar4 o;
    float d = 0.0f;
    if (best == ~0ull) {
        o = make_uchar4(255, 255, 255, 0);
    } else {
        const uint32_t idx = (uint32_t)(best & 0xFFFFFFFFull);
        o.x = rgb[3 * (size_t)idx + 0];
        o.y = rgb[3 * (size_t)idx + 1];
        o.z = rgb[3 * (size_t)idx + 2];
        o.w = (o.x == 255 && o.y == 255 && o.z == 255) ? 0 : 255;  // convertToAlpha, Share_Data.hpp:771-784
        d = __uint_as_float((uint32_t)(best >> 32));
    }
    reinterpret_cast<uchar4*>(rgba)[pix] = o;
    if (depth) depth[pix] = d;
}

// size-augmentation probe (main.cpp:913-931): pixels of a rendered view that are not white, counted where the image lies
__global__ void __launch_bounds__(256) count_nonwhite_kernel(const uchar4* __restrict__ rgba, uint32_t npix, uint32_t* __restrict__ counts) {
    __shared__ uint32_t s_red[8];
    const uint32_t view = blockIdx.y;
    const uchar4* img = rgba + (size_t)view * npix;
    uint32_t c = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += gridDim.x * blockDim.x) {
        const uchar4 v = img[i];
        c += (v.x != 255 || v.y != 255 || v.z != 255) ? 1u : 0u;
    }
    const uint32_t t = block_reduce_sum(c, s_red);
    if (threadIdx.x == 0 && t) atomicAdd(counts + view, t);
}
