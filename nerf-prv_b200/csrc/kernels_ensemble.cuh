// kernels_ensemble.cuh -- ensemble-uncertainty view scoring (nbv_loop cases 2/3)
// Part of the single translation unit prv_device.cu (included there, in order); see DESIGN.md section 4.
#pragma once

// ensemble-uncertainty scoring (nbv_loop cases 2/3, main.cpp:2039-2161) ------------------------------------------
// Stage 1: one thread per (view, pixel) computes the pixel's contribution(s) in the reference's double arithmetic.
// Stage 2: one thread per view adds them in the reference's order (row-major pixels, channel order), so the score is
// the same sequence of double additions.  NaN marks "no term" (method 2 skips variances <= 1e-10).
__global__ void __launch_bounds__(256) ensemble_terms_kernel(const uint8_t* __restrict__ images, uint32_t E, uint32_t npix, int method,
                                                             const double* __restrict__ log_lut, double* __restrict__ terms) {
    const uint32_t view = blockIdx.y;
    const uint32_t pix = blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= npix) return;
    const uchar4* base = reinterpret_cast<const uchar4*>(images) + (size_t)view * E * npix + pix;
    double mean[3] = {0.0, 0.0, 0.0};
    double mean_density = 0.0;
    for (uint32_t e = 0; e < E; e++) {
        const uchar4 px = base[(size_t)e * npix];
        mean[0] = prvk::dadd(mean[0], (double)px.x);
        mean[1] = prvk::dadd(mean[1], (double)px.y);
        mean[2] = prvk::dadd(mean[2], (double)px.z);
        mean_density = prvk::dadd(mean_density, prvk::ddiv((double)px.w, 255.0));
    }
    for (int c = 0; c < 3; c++) mean[c] = prvk::ddiv(mean[c], (double)E);
    mean_density = prvk::ddiv(mean_density, (double)E);
    double variance[3] = {0.0, 0.0, 0.0};
    for (uint32_t e = 0; e < E; e++) {
        const uchar4 px = base[(size_t)e * npix];
        const double v[3] = {(double)px.x, (double)px.y, (double)px.z};
        for (int c = 0; c < 3; c++) {
            const double d = prvk::dsub(v[c], mean[c]);
            variance[c] = prvk::dadd(variance[c], prvk::dmul(d, d));
        }
    }
    for (int c = 0; c < 3; c++) variance[c] = prvk::ddiv(variance[c], (double)E);
    double* out = terms + ((size_t)view * npix + pix) * 3;
    const double nan = __longlong_as_double(0x7FF8000000000000ll);
    if (method == 2) {
        for (int c = 0; c < 3; c++) {
            double t = nan;
            if (variance[c] > 1e-10) {
                if (log_lut) {  // E == 2: variance = (|a-b|/2)^2 exactly; host libm values
                    const uchar4 p0 = base[0], p1 = base[npix];
                    const int a = c == 0 ? p0.x : (c == 1 ? p0.y : p0.z), b = c == 0 ? p1.x : (c == 1 ? p1.y : p1.z);
                    t = log_lut[a > b ? a - b : b - a];
                } else {
                    t = log(variance[c]);
                }
            }
            out[c] = t;
        }
    } else {
        out[0] = prvk::ddiv(prvk::dadd(prvk::dadd(variance[0], variance[1]), variance[2]), 3.0);
        const double q = prvk::dsub(1.0, mean_density);
        out[1] = prvk::dmul(q, q);
        out[2] = nan;
    }
}

__global__ void __launch_bounds__(32) ensemble_sum_kernel(const double* __restrict__ terms, uint32_t V, uint32_t npix, double* __restrict__ scores) {
    const uint32_t view = blockIdx.x * blockDim.x + threadIdx.x;
    if (view >= V) return;
    const double* t = terms + (size_t)view * npix * 3;
    double acc = 0.0;
    for (size_t i = 0; i < (size_t)npix * 3; i++) {
        const double v = t[i];
        if (v == v) acc = prvk::dadd(acc, v);
    }
    scores[view] = acc;
}
