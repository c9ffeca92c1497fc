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

__global__ void __launch_bounds__(256) splat_resolve_kernel(const unsigned long long* corner, int Wc, int Hc, int W, int H, int point_size,
                                                            const uint8_t* rgb, uint8_t* rgba, float* depth, uint32_t view_base) {
    extern __shared__ unsigned long long s_tile[];
    const uint32_t view_local = blockIdx.z;
    const int tw = 32 + point_size - 1, th = 8 + point_size - 1;
    const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 8;
    const unsigned long long* src = corner + (size_t)view_local * Hc * Wc;
    for (int t = threadIdx.x; t < tw * th; t += blockDim.x) {
        const int cx = x0 + t % tw, cy = y0 + t / tw;
        s_tile[t] = (cx < Wc && cy < Hc) ? src[(size_t)cy * Wc + cx] : ~0ull;
    }
    __syncthreads();
    const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
    const int x = x0 + lx, y = y0 + ly;
    if (x >= W || y >= H) return;
    unsigned long long best = ~0ull;
    for (int dy = 0; dy < point_size; dy++)
        for (int dx = 0; dx < point_size; dx++) best = min(best, s_tile[(ly + dy) * tw + lx + dx]);
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
