// kernels_map.cuh -- GPU map build (bitmaps, coarse grid, prefix, rank table) and GPU ingest of the cloud
// Part of the single translation unit prv_device.cu (included there, in order); see DESIGN.md section 4.
#pragma once

// ground-truth map build on the GPU (replaces the CPU octree insertion as far as the cast needs it) ------------------
struct MapBuild {
    int lo[3], n[3], wx, row_log2, nc[3];
    uint32_t n_occ;
    unsigned long long slack_bits, nwords;
    uint32_t* bitmap;
    uint32_t* pad;
    uint32_t* coarse;
    uint32_t* prefix;
    uint32_t* leaf_of_raster;
    const uint16_t* keys;
    int cs;  // brick edge in voxels
};

// Leaf-order check and AABB of a key table, on the device (main.cpp:116-121: the table must be in begin_leafs() order, strictly
// ascending Morton code).  out[0] = index of the first key whose code is not above its predecessor's (left at N when the
// table is in order), out[1..3] = smallest key per axis, out[4..6] = largest.  The host initialises out to {N, 65535 x3, 0 x3}.
// (Round 1 ran both loops on the host at the start of every prv_set_map while the GPU waited: 0.1 ms for 45 k keys.)
__global__ void __launch_bounds__(256) map_check_kernel(const uint16_t* __restrict__ keys, uint32_t N, uint32_t* out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const uint16_t k[3] = {keys[3 * (size_t)i], keys[3 * (size_t)i + 1], keys[3 * (size_t)i + 2]};
    if (i > 0 && prv::morton_code(k[0], k[1], k[2]) <= prv::morton_code(keys[3 * (size_t)i - 3], keys[3 * (size_t)i - 2], keys[3 * (size_t)i - 1]))
        atomicMin(out, i);
    // thousands of keys fold into six words: test first (a stale read only costs a redundant atomic), as map_scatter_kernel does
#pragma unroll
    for (int a = 0; a < 3; a++) {
        if ((uint32_t)k[a] < __ldcg(out + 1 + a)) atomicMin(out + 1 + a, (uint32_t)k[a]);
        if ((uint32_t)k[a] > __ldcg(out + 4 + a)) atomicMax(out + 4 + a, (uint32_t)k[a]);
    }
}

// per occupied voxel: occupancy bit, padded-bitmap bit, and the (<= 8) coarse cells within one voxel of it
__global__ void __launch_bounds__(256) map_scatter_kernel(MapBuild b) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= b.n_occ) return;
    const int q0 = b.keys[3 * i] - b.lo[0], q1 = b.keys[3 * i + 1] - b.lo[1], q2 = b.keys[3 * i + 2] - b.lo[2];
    atomicOr(b.bitmap + ((size_t)(q2 * b.n[1] + q1) * b.wx + (q0 >> 5)), 1u << (q0 & 31));
    const unsigned long long L = b.slack_bits + ((unsigned long long)((q2 + 1) * (b.n[1] + 2) + (q1 + 1)) << b.row_log2) + (unsigned long long)(q0 + 1);
    atomicOr(b.pad + (L >> 5), 1u << (L & 31));
    const int q[3] = {q0, q1, q2};
    int cl[3], ch[3];
    for (int a = 0; a < 3; a++) {
        cl[a] = max(0, (q[a] - 1) / b.cs);
        ch[a] = min(b.nc[a] - 1, (q[a] + 1) / b.cs);
    }
    for (int K = cl[2]; K <= ch[2]; K++)
        for (int J = cl[1]; J <= ch[1]; J++)
            for (int I = cl[0]; I <= ch[0]; I++) {
                // thousands of voxels mark the same few coarse words: test first (a stale read only costs a redundant
                // atomic), otherwise the same-address atomics serialise in L2 and dominate the kernel
                const uint32_t c = (uint32_t)((K * b.nc[1] + J) * b.nc[0] + I);
                const uint32_t bit = 1u << (c & 31);
                if (!(__ldcg(b.coarse + (c >> 5)) & bit)) atomicOr(b.coarse + (c >> 5), bit);
            }
}

// the fully-set one-voxel shell of the padded bitmap: one thread per padded row
__global__ void __launch_bounds__(256) map_shell_kernel(MapBuild b) {
    const uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n1p = (uint32_t)b.n[1] + 2, n2p = (uint32_t)b.n[2] + 2;
    if (row >= n1p * n2p) return;
    const uint32_t c1 = row % n1p, c2 = row / n1p;
    const unsigned long long L0 = b.slack_bits + ((unsigned long long)row << b.row_log2);  // multiple of 32
    uint32_t* w = b.pad + (L0 >> 5);
    const uint32_t last = (uint32_t)b.n[0] + 1;  // padded x of the far shell cell
    if (c1 == 0 || c1 == n1p - 1 || c2 == 0 || c2 == n2p - 1) {
        for (uint32_t x = 0; x <= last; x += 32) {
            const uint32_t cnt = min(32u, last + 1 - x);
            atomicOr(w + (x >> 5), cnt == 32 ? 0xFFFFFFFFu : ((1u << cnt) - 1u));
        }
    } else {
        atomicOr(w, 1u);
        atomicOr(w + (last >> 5), 1u << (last & 31));
    }
}

// exclusive popcount prefix over the occupancy words, two coalesced passes over kPrefixTile-word tiles:
// map_tilesum_kernel: popcount of every tile; map_prefix_kernel: tile base = sum of the tiles before it (at most a few
// hundred values, re-added by every block) + a block scan inside the tile.  (A single 1024-thread block with one
// contiguous run per thread took 48 us on C2's 280k words.)
constexpr int kPrefixThreads = 256, kPrefixPerThread = 4, kPrefixTile = kPrefixThreads * kPrefixPerThread;
__global__ void __launch_bounds__(kPrefixThreads) map_tilesum_kernel(MapBuild b, uint32_t* tile_sums) {
    __shared__ uint32_t s_red[8];
    const unsigned long long w0 = (unsigned long long)blockIdx.x * kPrefixTile;
    uint32_t c = 0;
#pragma unroll
    for (int j = 0; j < kPrefixPerThread; j++) {
        const unsigned long long w = w0 + (unsigned long long)j * kPrefixThreads + threadIdx.x;
        if (w < b.nwords) c += __popc(b.bitmap[w]);
    }
    const uint32_t t = block_reduce_sum(c, s_red);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = t;
}

__global__ void __launch_bounds__(kPrefixThreads) map_prefix_kernel(MapBuild b, const uint32_t* tile_sums) {
    __shared__ uint32_t s_red[8];
    __shared__ uint32_t s_warp[8];
    __shared__ uint32_t s_base;
    uint32_t c = 0;
    for (uint32_t i = threadIdx.x; i < blockIdx.x; i += blockDim.x) c += tile_sums[i];
    const uint32_t base = block_reduce_sum(c, s_red);
    if (threadIdx.x == 0) s_base = base;
    // each thread owns kPrefixPerThread consecutive words of the tile
    const unsigned long long w0 = (unsigned long long)blockIdx.x * kPrefixTile + (unsigned long long)threadIdx.x * kPrefixPerThread;
    uint32_t pc[kPrefixPerThread];
    uint32_t sum = 0;
#pragma unroll
    for (int j = 0; j < kPrefixPerThread; j++) {
        pc[j] = w0 + j < b.nwords ? __popc(b.bitmap[w0 + j]) : 0u;
        sum += pc[j];
    }
    uint32_t incl = sum;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t run = s_base + incl - sum;
    for (int w = 0; w < warp; w++) run += s_warp[w];
#pragma unroll
    for (int j = 0; j < kPrefixPerThread; j++) {
        if (w0 + j < b.nwords) b.prefix[w0 + j] = run;
        run += pc[j];
    }
}

// raster rank -> leaf rank
__global__ void __launch_bounds__(256) map_rank_kernel(MapBuild b) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= b.n_occ) return;
    const int q0 = b.keys[3 * i] - b.lo[0], q1 = b.keys[3 * i + 1] - b.lo[1], q2 = b.keys[3 * i + 2] - b.lo[2];
    const size_t w = (size_t)(q2 * b.n[1] + q1) * b.wx + (q0 >> 5);
    b.leaf_of_raster[b.prefix[w] + __popc(b.bitmap[w] & ((1u << (q0 & 31)) - 1u))] = i;
}

// GPU ingest: cloud points -> leaf-ordered unique keys + first-point colours ------------------------
// key = (int)floor(resolution_factor * (double)coord) + 32768 (coordToKeyChecked, main.cpp:1015), invalid points sort last
__global__ void __launch_bounds__(256) ingest_keys_kernel(const float* __restrict__ xyz, uint32_t P, double resolution_factor,
                                                          unsigned long long* __restrict__ codes, uint32_t* __restrict__ index) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    uint16_t k[3];
    bool ok = true;
    for (int a = 0; a < 3; a++) {
        // the range test is made on the double: (int) of a NaN / out-of-range double is 0 / saturated on the device but INT_MIN
        // on x86 (where the reference, the oracle and prv_host_build_map then reject the point)
        const double f = floor(prvk::dmul(resolution_factor, (double)xyz[3 * (size_t)i + a]));
        const bool in_range = f >= -(double)prv::kTreeMaxVal && f < (double)prv::kTreeMaxVal;
        ok = ok && in_range;
        k[a] = in_range ? (uint16_t)((int)f + prv::kTreeMaxVal) : (uint16_t)0;
    }
    codes[i] = ok ? prv::morton_code(k[0], k[1], k[2]) : (1ull << 48);
    index[i] = i;
}

// head[i] = 1 when sorted entry i starts a new voxel (head[P] = 0 pads the scan so pos[P] = number of voxels)
__global__ void __launch_bounds__(256) ingest_heads_kernel(const unsigned long long* __restrict__ codes, uint32_t P, uint32_t* __restrict__ head) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > P) return;
    uint32_t h = 0;
    if (i < P) {
        const unsigned long long c = codes[i];
        h = (c < (1ull << 48)) && (i == 0 || codes[i - 1] != c) ? 1u : 0u;
    }
    head[i] = h;
}

__global__ void __launch_bounds__(256) ingest_compact_kernel(const unsigned long long* __restrict__ codes, const uint32_t* __restrict__ index,
                                                             const uint32_t* __restrict__ head, const uint32_t* __restrict__ pos, uint32_t P,
                                                             const uint8_t* __restrict__ rgb_in, uint16_t* __restrict__ keys_out, uint8_t* __restrict__ rgb_out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P || !head[i]) return;
    const uint32_t j = pos[i];
    uint16_t k[3];
    prv::morton_decode(codes[i], k);
    const uint32_t src = index[i];
    for (int a = 0; a < 3; a++) {
        keys_out[3 * (size_t)j + a] = k[a];
        rgb_out[3 * (size_t)j + a] = rgb_in[3 * (size_t)src + a];
    }
}
