// march_model.cpp -- PLANNING TOOL (not a test, not a product path): warp-level instruction model of coarse_kernel and
// march_kernel, driven by the kernels' own per-ray code compiled for the CPU (tests/cpp/kernel_on_host.cpp; its stage
// counters equal the B200's).  Rays are grouped into warps the way the pipeline groups them
// (cull_kernel region order -> 256-ray coarse chunks -> 32-ray march chunks, atomics assumed to land in block order) and
// every loop is charged max-over-lanes trip count x SASS instructions per trip (counted in the sm_100a SASS of round 1).
// The march kernel is issue-bound (ncu: 73 % issue-active), so warp-instructions are the quantity that moves its time.
// Used to rank candidate changes before spending GPU minutes on them: finer brick cull, starting the exact march at the
// first set fine cell instead of the AABB face, sorting hit-like and miss-like rays into different warps.
//
// Build / run: python tools/march_model.py [C1|C2|C3|C5] [view stride]
#include "../tests/cpp/kernel_on_host.cpp"

#include <cstdio>

namespace {

struct RayRec {
    uint32_t pid;
    uint8_t hit, fine_keep[3];  // survives the fine cull with K = 4, 2, 1
    uint16_t walk8, walkf[3];   // float cell steps of the 8^3 walk alone / nested walk with K = 4, 2, 1
    uint16_t p1[3];             // phase-1 additions per axis
    uint16_t m2[3], e2[3];      // phase-2 counted additions / compare-loop iterations per axis
    uint16_t nprobe;            // in-AABB probes including the terminating set bit
    uint16_t kentry[3];         // in-AABB steps before the ray is inside the box of the first set fine cell grown by one voxel (K = 4, 2, 1)
};

// march_axis with its trip counts exposed; must give the same (rank, steps, probes) as the header's march_axis
bool trace_ray(const DevMap& m, const ViewConst& vc, RayState r, RayRec& rec, CastResult& out, const int (*box)[6]) {
    out.rank = kNone;
    out.steps = out.probes = 0;
    for (int a = 0; a < 3; a++) rec.p1[a] = rec.m2[a] = rec.e2[a] = 0;
    rec.nprobe = 0;
    for (int k = 0; k < 3; k++) rec.kentry[k] = 0;
    int q0 = vc.okey[0] - m.lo[0], q1 = vc.okey[1] - m.lo[1], q2 = vc.okey[2] - m.lo[2];
    int a0, a1, a2, b0, b1, b2;
    if (!axis_window(q0, m.n[0], r.s0, a0, b0) || !axis_window(q1, m.n[1], r.s1, a1, b1) || !axis_window(q2, m.n[2], r.s2, a2, b2)) return true;
    if (slab_miss(r, a0, a1, a2, b0, b1, b2)) return true;
    uint32_t nsteps = 0;
    bool probe_first = false;
    if ((a0 | a1 | a2) != 0) {
        probe_first = true;
        int n0 = a0 > 0 ? a0 - 1 : 0, n1 = a1 > 0 ? a1 - 1 : 0, n2 = a2 > 0 ? a2 - 1 : 0;
        rec.p1[0] = (uint16_t)n0; rec.p1[1] = (uint16_t)n1; rec.p1[2] = (uint16_t)n2;
        for (int k = 0; k < n0; k++) r.t0 = dadd(r.t0, r.d0);
        for (int k = 0; k < n1; k++) r.t1 = dadd(r.t1, r.d1);
        for (int k = 0; k < n2; k++) r.t2 = dadd(r.t2, r.d2);
        double tstar = -1.0;
        int j = -1;
        if (a2 > 0) { tstar = r.t2; j = 2; }
        if (a1 > 0 && r.t1 >= tstar) { tstar = r.t1; j = 1; }
        if (a0 > 0 && r.t0 >= tstar) { tstar = r.t0; j = 0; }
        const double tup = next_up_pos(tstar);
        auto adv = [&](double& t, double d, double thr, int& n, int axis) {
            const float est = __fmul_rn((float)(thr - t), __frcp_rn((float)d));
            const int mm = est > 2.0f ? (int)est - 2 : 0;
            const int before = n;
            advance_below(t, d, thr, n);
            rec.m2[axis] = (uint16_t)(mm > 0 ? mm : 0);
            rec.e2[axis] = (uint16_t)(n - before - (mm > 0 ? mm : 0));
        };
        if (j != 0) adv(r.t0, r.d0, tstar, n0, 0);
        if (j != 1) adv(r.t1, r.d1, j < 1 ? tup : tstar, n1, 1);
        if (j != 2) adv(r.t2, r.d2, j < 2 ? tup : tstar, n2, 2);
        if (j == 0) { r.t0 = dadd(r.t0, r.d0); n0 = a0; }
        else if (j == 1) { r.t1 = dadd(r.t1, r.d1); n1 = a1; }
        else { r.t2 = dadd(r.t2, r.d2); n2 = a2; }
        nsteps = (uint32_t)(n0 + n1 + n2);
        if (n0 > b0 || n1 > b1 || n2 > b2) {
            out.steps = nsteps;
            return true;
        }
        q0 += r.s0 * n0; q1 += r.s1 * n1; q2 += r.s2 * n2;
    }
    // in-AABB: one probe at a time (same cells as the 4-deep loop of the header)
    uint32_t nprobe = 0;
    bool found = false;
    bool entered[3] = {false, false, false};
    auto note_entry = [&]() {
        for (int k = 0; k < 3; k++)
            if (!entered[k]) {
                if (box[k][0] > box[k][3]) { entered[k] = true; continue; }  // no box: ray was culled at this level
                if (q0 >= box[k][0] && q0 <= box[k][3] && q1 >= box[k][1] && q1 <= box[k][4] && q2 >= box[k][2] && q2 <= box[k][5]) entered[k] = true;
                else rec.kentry[k]++;
            }
    };
    auto bit_at = [&](int c0, int c1, int c2) -> bool {
        if ((unsigned)c0 >= (unsigned)m.n[0] || (unsigned)c1 >= (unsigned)m.n[1] || (unsigned)c2 >= (unsigned)m.n[2]) return true;  // shell
        const uint32_t w = m.bitmap[(size_t)(c2 * m.n[1] + c1) * m.wx + (c0 >> 5)];
        return (w >> (c0 & 31)) & 1u;
    };
    if (probe_first) {
        nprobe = 1;
        note_entry();
        found = bit_at(q0, q1, q2);
    }
    while (!found) {
        const uint32_t inc = dda_step(r.t0, r.t1, r.t2, r.d0, r.d1, r.d2, 0u, 1u, 2u);
        if (inc == 0) q0 += r.s0; else if (inc == 1) q1 += r.s1; else q2 += r.s2;
        nprobe++;
        note_entry();
        found = bit_at(q0, q1, q2);
    }
    rec.nprobe = (uint16_t)nprobe;
    out.steps = nsteps + nprobe - (probe_first ? 1u : 0u);
    if ((unsigned)q0 >= (unsigned)m.n[0] || (unsigned)q1 >= (unsigned)m.n[1] || (unsigned)q2 >= (unsigned)m.n[2]) {
        out.probes = nprobe - 1;
        return true;
    }
    out.probes = nprobe;
    out.rank = probe(m, q0, q1, q2);
    return true;
}

// nested brick walk that also reports the first set fine cell (AABB-relative voxel box grown by one voxel) and its cost
struct FineGrid {
    int K, nf[3];
    std::vector<uint32_t> bits;
};
void build_fine(const HostMap& hm, int K, FineGrid& g) {
    g.K = K;
    for (int a = 0; a < 3; a++) g.nf[a] = (hm.m.n[a] + K - 1) / K;
    g.bits.assign(((size_t)g.nf[0] * g.nf[1] * g.nf[2] + 31) / 32 + 1, 0u);
    for (uint32_t i = 0; i < hm.m.n_occ; i++) {
        const int q[3] = {hm.keys[3 * i] - hm.m.lo[0], hm.keys[3 * i + 1] - hm.m.lo[1], hm.keys[3 * i + 2] - hm.m.lo[2]};
        for (int d2 = -1; d2 <= 1; d2++)
            for (int d1 = -1; d1 <= 1; d1++)
                for (int d0 = -1; d0 <= 1; d0++) {
                    const int p0 = q[0] + d0, p1 = q[1] + d1, p2 = q[2] + d2;
                    if (p0 < 0 || p1 < 0 || p2 < 0 || p0 >= hm.m.n[0] || p1 >= hm.m.n[1] || p2 >= hm.m.n[2]) continue;
                    const uint32_t c = (uint32_t)(((p2 / K) * g.nf[1] + p1 / K) * g.nf[0] + p0 / K);
                    g.bits[c >> 5] |= 1u << (c & 31);
                }
    }
}
// returns keep; iters = float cell steps (coarse + fine); box = voxel box of the first set fine cell grown by one voxel
bool nested_walk(const DevMap& m, const FineGrid& f, const ViewConst& vc, float dx, float dy, float dz, int& iters, int box[6]) {
    const float o[3] = {(float)(vc.okey[0] - m.lo[0]) + 0.5f, (float)(vc.okey[1] - m.lo[1]) + 0.5f, (float)(vc.okey[2] - m.lo[2]) + 0.5f};
    const float d[3] = {dx, dy, dz};
    float inv[3], t0 = 0.0f, t1 = 3.0e38f;
    iters = 0;
    box[0] = 1; box[3] = 0;
    for (int a = 0; a < 3; a++) {
        const float hi = (float)m.n[a] + 1.0f;
        if (fabsf(d[a]) > 1.0e-12f) {
            inv[a] = 1.0f / d[a];
            const float ta = (-1.0f - o[a]) * inv[a], tb = (hi - o[a]) * inv[a];
            t0 = fmaxf(t0, fminf(ta, tb));
            t1 = fminf(t1, fmaxf(ta, tb));
        } else {
            inv[a] = 0.0f;
            if (o[a] < -1.0f || o[a] > hi) return false;
        }
    }
    if (!(t0 <= t1)) return t0 <= t1 * 1.0001f + 1.0e-3f;  // grazing: kept, no box (march from the AABB face)
    int c[3], st[3];
    float tm[3], td[3];
    for (int a = 0; a < 3; a++) {
        const float pa = o[a] + t0 * d[a];
        c[a] = max(0, min((int)floorf(pa / (float)kCoarse), m.nc[a] - 1));
        if (inv[a] != 0.0f) {
            st[a] = d[a] > 0.0f ? 1 : -1;
            tm[a] = ((float)((c[a] + (st[a] > 0 ? 1 : 0)) * kCoarse) - o[a]) * inv[a];
            td[a] = (float)kCoarse * fabsf(inv[a]);
        } else { st[a] = 0; tm[a] = 3.0e38f; td[a] = 0.0f; }
    }
    const int limit = m.nc[0] + m.nc[1] + m.nc[2] + 3;
    float tcur = t0;
    for (int it = 0; it < limit; it++) {
        iters++;
        const uint32_t bit = (uint32_t)((c[2] * m.nc[1] + c[1]) * m.nc[0] + c[0]);
        const float texit = fminf(tm[0], fminf(tm[1], tm[2]));
        if ((m.coarse[bit >> 5] >> (bit & 31)) & 1u) {
            // fine sub-walk (same logic as fine_cells_hit, plus the cell it stops in)
            const int K = f.K, R = kCoarse / K;
            int fc[3], fs[3], lo[3], hi[3];
            float fm[3], fd[3];
            for (int a = 0; a < 3; a++) {
                lo[a] = c[a] * R;
                hi[a] = min(f.nf[a] - 1, lo[a] + R - 1);
                const float pa = o[a] + tcur * d[a];
                fc[a] = max(lo[a], min((int)floorf(pa / (float)K), hi[a]));
                if (inv[a] != 0.0f) {
                    fs[a] = d[a] > 0.0f ? 1 : -1;
                    fm[a] = ((float)((fc[a] + (fs[a] > 0 ? 1 : 0)) * K) - o[a]) * inv[a];
                    fd[a] = (float)K * fabsf(inv[a]);
                } else { fs[a] = 0; fm[a] = 3.0e38f; fd[a] = 0.0f; }
            }
            const float tend = fmaf(texit, 1.0001f, 1.0e-3f);
            for (int s = 0; s < 3 * R + 3; s++) {
                iters++;
                const uint32_t fb = (uint32_t)((fc[2] * f.nf[1] + fc[1]) * f.nf[0] + fc[0]);
                if ((f.bits[fb >> 5] >> (fb & 31)) & 1u) {
                    for (int a = 0; a < 3; a++) {
                        box[a] = max(0, fc[a] * K - 1);
                        box[3 + a] = min(m.n[a] - 1, fc[a] * K + K);
                    }
                    return true;
                }
                const int a = (fm[0] <= fm[1] && fm[0] <= fm[2]) ? 0 : (fm[1] <= fm[2] ? 1 : 2);
                if (fm[a] > tend) break;
                fc[a] += fs[a];
                fm[a] += fd[a];
                if (fc[a] < lo[a] || fc[a] > hi[a]) break;
            }
        }
        const int a = (tm[0] <= tm[1] && tm[0] <= tm[2]) ? 0 : (tm[1] <= tm[2] ? 1 : 2);
        c[a] += st[a];
        tm[a] += td[a];
        if ((unsigned)c[a] >= (unsigned)m.nc[a]) return false;
        tcur = texit;
    }
    return true;
}
int walk8_iters(const DevMap& m, const ViewConst& vc, float dx, float dy, float dz) {  // coarse_miss, counting cell steps
    const float o[3] = {(float)(vc.okey[0] - m.lo[0]) + 0.5f, (float)(vc.okey[1] - m.lo[1]) + 0.5f, (float)(vc.okey[2] - m.lo[2]) + 0.5f};
    const float d[3] = {dx, dy, dz};
    float inv[3], t0 = 0.0f, t1 = 3.0e38f;
    for (int a = 0; a < 3; a++) {
        const float hi = (float)m.n[a] + 1.0f;
        if (fabsf(d[a]) > 1.0e-12f) {
            inv[a] = 1.0f / d[a];
            const float ta = (-1.0f - o[a]) * inv[a], tb = (hi - o[a]) * inv[a];
            t0 = fmaxf(t0, fminf(ta, tb));
            t1 = fminf(t1, fmaxf(ta, tb));
        } else { inv[a] = 0.0f; if (o[a] < -1.0f || o[a] > hi) return 0; }
    }
    if (!(t0 <= t1)) return 0;
    int c[3], st[3];
    float tm[3], td[3];
    for (int a = 0; a < 3; a++) {
        const float pa = o[a] + t0 * d[a];
        c[a] = max(0, min((int)floorf(pa / (float)kCoarse), m.nc[a] - 1));
        if (inv[a] != 0.0f) { st[a] = d[a] > 0.0f ? 1 : -1; tm[a] = ((float)((c[a] + (st[a] > 0 ? 1 : 0)) * kCoarse) - o[a]) * inv[a]; td[a] = (float)kCoarse * fabsf(inv[a]); }
        else { st[a] = 0; tm[a] = 3.0e38f; td[a] = 0.0f; }
    }
    int iters = 0;
    for (int it = 0; it < m.nc[0] + m.nc[1] + m.nc[2] + 3; it++) {
        iters++;
        const uint32_t bit = (uint32_t)((c[2] * m.nc[1] + c[1]) * m.nc[0] + c[0]);
        if ((m.coarse[bit >> 5] >> (bit & 31)) & 1u) return iters;
        const int a = (tm[0] <= tm[1] && tm[0] <= tm[2]) ? 0 : (tm[1] <= tm[2] ? 1 : 2);
        c[a] += st[a];
        tm[a] += td[a];
        if ((unsigned)c[a] >= (unsigned)m.nc[a]) return iters;
    }
    return iters;
}

}  // namespace

extern "C" {

// One view: fills recs (capacity cap) with the slab-surviving rays in pipeline (queue 1) order; returns their number, or
// -1 on a mismatch between trace_ray and the header's march_axis.  flags_out: bit0 = ray also survives the 8^3 brick cull.
int model_view(const uint16_t* keys, uint32_t N, double resolution, const prv_intrinsics* intr, const double* pose_world, const double* init_pos,
               void* recs_out, uint8_t* brick_keep_out, int cap) {
    static HostMap hm;
    static FineGrid fg[3];
    static const uint16_t* cached = nullptr;
    if (cached != keys) {
        hm = HostMap();
        build_map(hm, keys, nullptr, N, resolution, 1.0, 0);
        const int Ks[3] = {4, 2, 1};
        for (int k = 0; k < 3; k++) build_fine(hm, Ks[k], fg[k]);
        cached = keys;
    }
    const DevCam cam = make_cam(*intr, 1.0, -1);
    ViewConst vc;
    std::memset(&vc, 0, sizeof(vc));
    make_view_const(hm.setup, pose_world, init_pos, 0, vc);
    if (!(vc.flags & kViewFastOk) || !cam.region_cull_ok) return -2;
    RayRec* recs = (RayRec*)recs_out;
    int n = 0;
    const int W = cam.W, H = cam.H, regions_x = (W + 31) >> 5, regions_y = (H + 31) >> 5;
    for (int ry = 0; ry < regions_y; ry++)
        for (int rx = 0; rx < regions_x; rx++) {
            uint32_t bal = 0;
            for (int lane = 0; lane < 32; lane++)
                if (region_corner_outside(hm.m, cam, vc, rx, ry, lane)) bal |= 1u << lane;
            if (region_skip_from_ballot(bal)) continue;
            for (int t = 0; t < 4; t++)          // cull_kernel's order inside a region: row-tile, warp, lane
                for (int warp = 0; warp < 8; warp++)
                    for (int lane = 0; lane < 32; lane++) {
                        const int px = (rx << 5) + ((warp & 3) << 3) + (lane & 7);
                        const int py = (ry << 5) + (t << 3) + ((warp >> 2) << 2) + (lane >> 3);
                        if (px >= W || py >= H) continue;
                        float dx, dy, dz;
                        ray_direction_approx(cam, vc, (float)px, (float)py, dx, dy, dz);
                        if (loose_miss(hm.m, vc, dx, dy, dz)) continue;
                        if (n >= cap) return -3;
                        RayRec& rec = recs[n];
                        std::memset(&rec, 0, sizeof(rec));
                        rec.pid = (uint32_t)py * W + px;
                        const bool keep8 = !coarse_miss(hm.m, vc, dx, dy, dz);
                        brick_keep_out[n] = keep8 ? 1 : 0;
                        rec.walk8 = (uint16_t)walk8_iters(hm.m, vc, dx, dy, dz);
                        int box[3][6];
                        for (int k = 0; k < 3; k++) {
                            int it;
                            rec.fine_keep[k] = nested_walk(hm.m, fg[k], vc, dx, dy, dz, it, box[k]) ? 1 : 0;
                            rec.walkf[k] = (uint16_t)it;
                        }
                        RayState r;
                        float ex, ey, ez;
                        ray_direction(cam, vc, px, py, ex, ey, ez);
                        CastResult ref, got;
                        ref.rank = kNone; ref.steps = ref.probes = 0;
                        if (ray_init(vc, hm.m.resolution, ex, ey, ez, r)) {
                            march_axis(hm.m, vc, r, ref);
                            trace_ray(hm.m, vc, r, rec, got, box);
                            if (got.rank != ref.rank || got.steps != ref.steps || got.probes != ref.probes) return -1;
                        }
                        rec.hit = ref.rank != kNone;
                        if (rec.hit && !keep8) return -4;  // a cull removed a hit
                        for (int k = 0; k < 3; k++)
                            if (rec.hit && !rec.fine_keep[k]) return -5;
                        n++;
                    }
        }
    return n;
}

int model_rec_size() { return (int)sizeof(RayRec); }

}  // extern "C"
