// kernels_greedy.cuh -- coverage counts and the greedy set-cover kernels
// Part of the single translation unit prv_device.cu (included there, in order); see DESIGN.md section 4.
#pragma once

// coverage_count[v] = popcount(vis[v])
__global__ void __launch_bounds__(256) popcount_rows_kernel(const uint64_t* rows, uint32_t words64, uint32_t* counts) {
    __shared__ uint32_t s_red[8];
    const ulonglong2* row = reinterpret_cast<const ulonglong2*>(rows + (size_t)blockIdx.x * words64);
    uint32_t c = 0;
    for (uint32_t w = threadIdx.x; w < words64 / 2; w += blockDim.x) {
        const ulonglong2 v = row[w];
        c += __popcll(v.x) + __popcll(v.y);
    }
    const uint32_t t = block_reduce_sum(c, s_red);
    if (threadIdx.x == 0) counts[blockIdx.x] = t;
}

// greedy set cover ---------------------------------------------------------------------------------
// best[k] = max over views of (gain << 32) | (0xFFFFFFFF - view_id): largest gain, then LOWEST view id.
__global__ void __launch_bounds__(256) greedy_init_kernel(const uint64_t* rows, uint32_t words64, uint32_t first_row, uint32_t first_id,
                                                          unsigned long long* best) {
    __shared__ uint32_t s_red[8];
    const uint64_t* row = rows + (size_t)first_row * words64;
    uint32_t c = 0;
    for (uint32_t w = threadIdx.x; w < words64; w += blockDim.x) c += __popcll(row[w]);
    const uint32_t t = block_reduce_sum(c, s_red);
    if (threadIdx.x == 0) best[0] = ((unsigned long long)t << 32) | (unsigned long long)(0xFFFFFFFFu - first_id);
}

// iteration k >= 1: covered_k = covered_{k-1} | row[best_{k-1}];  score every view against covered_k.
// Block 0 also materialises covered_k for the next launch.  score_only_cover: last launch, no scoring.
__global__ void __launch_bounds__(256) greedy_iter_kernel(const uint64_t* rows, uint32_t words64, const uint32_t* view_ids,
                                                          const uint32_t* row_of_id, uint32_t k, unsigned long long* best,
                                                          const uint64_t* cov_prev, uint64_t* cov_next, int cover_only) {
    __shared__ uint32_t s_red[8];
    const unsigned long long prev = best[k - 1];
    if (k > 1 && (prev >> 32) == 0ull) return;  // previous argmax had zero gain: selection is over
    const uint32_t prev_id = 0xFFFFFFFFu - (uint32_t)(prev & 0xFFFFFFFFull);
    const ulonglong2* rb = reinterpret_cast<const ulonglong2*>(rows + (size_t)row_of_id[prev_id] * words64);
    const ulonglong2* cp = reinterpret_cast<const ulonglong2*>(cov_prev);
    const ulonglong2* rv = reinterpret_cast<const ulonglong2*>(rows + (size_t)blockIdx.x * words64);
    ulonglong2* cn = reinterpret_cast<ulonglong2*>(cov_next);
    uint32_t c = 0;
    for (uint32_t w = threadIdx.x; w < words64 / 2; w += blockDim.x) {
        ulonglong2 cov = rb[w];
        if (k > 1) {
            const ulonglong2 o = cp[w];
            cov.x |= o.x;
            cov.y |= o.y;
        }
        if (blockIdx.x == 0) cn[w] = cov;
        if (!cover_only) {
            const ulonglong2 v = rv[w];
            c += __popcll(v.x & ~cov.x) + __popcll(v.y & ~cov.y);
        }
    }
    if (cover_only) return;
    const uint32_t t = block_reduce_sum(c, s_red);
    if (threadIdx.x == 0) atomicMax(best + k, ((unsigned long long)t << 32) | (unsigned long long)(0xFFFFFFFFu - view_ids[blockIdx.x]));
}

// Whole greedy loop in ONE persistent kernel (cooperative launch: every block is resident).  Each block keeps the
// covered mask in shared memory and scores its rows (row r -> block r mod gridDim) against it; the per-iteration argmax
// is one 64-bit atomicMax per block followed by a grid barrier (cooperative_groups grid sync); every block
// then ORs the winner's row into its own copy of the mask.  Same selection rule and results as greedy_iter_kernel.
__global__ void __launch_bounds__(256) greedy_persistent_kernel(const uint64_t* __restrict__ rows, uint32_t words64, uint32_t nrows,
                                                                const uint32_t* __restrict__ view_ids, const uint32_t* __restrict__ row_of_id,
                                                                uint32_t first_row, uint32_t first_id, uint32_t max_iter, unsigned long long* best,
                                                                uint64_t* cov_out, unsigned int* arrive, int rows_in_smem) {
    extern __shared__ uint64_t s_cov[];
    __shared__ uint32_t s_red[8];
    __shared__ unsigned long long s_best;
    const uint32_t half = words64 / 2;
    ulonglong2* cov2 = reinterpret_cast<ulonglong2*>(s_cov);
    {
        const ulonglong2* r0 = reinterpret_cast<const ulonglong2*>(rows + (size_t)first_row * words64);
        uint32_t c = 0;
        for (uint32_t w = threadIdx.x; w < half; w += blockDim.x) {
            const ulonglong2 v = r0[w];
            cov2[w] = v;
            c += __popcll(v.x) + __popcll(v.y);
        }
        const uint32_t t = block_reduce_sum(c, s_red);
        if (blockIdx.x == 0 && threadIdx.x == 0) best[0] = ((unsigned long long)t << 32) | (unsigned long long)(0xFFFFFFFFu - first_id);
        __syncthreads();
    }
    // the block's own rows live in shared memory after the mask (the grid barrier's __threadfence invalidates L1 every
    // iteration, so rows left in global memory would be re-fetched from L2 each time)
    ulonglong2* srows = cov2 + half;
    if (rows_in_smem) {
        uint32_t slot = 0;
        for (uint32_t r = blockIdx.x; r < nrows; r += gridDim.x, slot++) {
            const ulonglong2* rv = reinterpret_cast<const ulonglong2*>(rows + (size_t)r * words64);
            for (uint32_t w = threadIdx.x; w < half; w += blockDim.x) srows[(size_t)slot * half + w] = rv[w];
        }
        __syncthreads();
    }
    for (uint32_t k = 1; k <= max_iter; k++) {
        unsigned long long local = 0ull;
        uint32_t slot = 0;
        for (uint32_t r = blockIdx.x; r < nrows; r += gridDim.x, slot++) {
            const ulonglong2* rv = rows_in_smem ? srows + (size_t)slot * half : reinterpret_cast<const ulonglong2*>(rows + (size_t)r * words64);
            uint32_t c = 0;
            for (uint32_t w = threadIdx.x; w < half; w += blockDim.x) {
                const ulonglong2 v = rv[w];
                const ulonglong2 cv = cov2[w];
                c += __popcll(v.x & ~cv.x) + __popcll(v.y & ~cv.y);
            }
            const uint32_t t = block_reduce_sum(c, s_red);
            if (threadIdx.x == 0) {
                const unsigned long long packed = ((unsigned long long)t << 32) | (unsigned long long)(0xFFFFFFFFu - view_ids[r]);
                local = packed > local ? packed : local;
            }
            __syncthreads();  // s_red reuse
        }
        // argmax across blocks + grid barrier
        // argmax across blocks: one 64-bit atomicMax per block, then the cooperative-groups grid barrier (measured 3-7 %
        // faster than a hand-written arrival counter with __threadfence + polling)
        if (threadIdx.x == 0) atomicMax(best + k, local);
        cooperative_groups::this_grid().sync();
        if (threadIdx.x == 0) s_best = *reinterpret_cast<volatile unsigned long long*>(best + k);
        __syncthreads();
        const unsigned long long b = s_best;
        if ((b >> 32) == 0ull) break;  // nothing left to gain: selection is over
        const uint32_t rb = row_of_id[0xFFFFFFFFu - (uint32_t)(b & 0xFFFFFFFFull)];
        const ulonglong2* rw = reinterpret_cast<const ulonglong2*>(rows + (size_t)rb * words64);
        for (uint32_t w = threadIdx.x; w < half; w += blockDim.x) {
            const ulonglong2 v = rw[w];
            ulonglong2 cv = cov2[w];
            cv.x |= v.x;
            cv.y |= v.y;
            cov2[w] = cv;
        }
        __syncthreads();
    }
    if (blockIdx.x == 0) {
        __syncthreads();
        for (uint32_t w = threadIdx.x; w < words64; w += blockDim.x) cov_out[w] = s_cov[w];
    }
}

// Greedy loop inside ONE thread-block cluster: the whole coverage table lives in the distributed shared memory of the
// cluster's CTAs (row r -> CTA r mod C, slot r div C), every CTA keeps its own copy of the covered mask.  One iteration =
// score own rows from shared memory, post the CTA's best to CTA 0 through DSMEM, ONE hardware cluster barrier, every CTA
// reduces the C candidates itself and ORs the winner's row (read through DSMEM from its owner) into its mask.  No global
// atomics, no polling; candidates are double-buffered so one barrier per iteration suffices.
__global__ void __launch_bounds__(512) greedy_cluster_kernel(const uint64_t* __restrict__ rows, uint32_t words64, uint32_t nrows,
                                                             const uint32_t* __restrict__ view_ids, const uint32_t* __restrict__ row_of_id,
                                                             uint32_t first_row, uint32_t first_id, uint32_t max_iter, uint32_t rows_per_cta,
                                                             unsigned long long* best, uint64_t* cov_out) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const uint32_t rank = cluster.block_rank(), C = cluster.num_blocks();
    extern __shared__ uint64_t s_mem[];
    __shared__ unsigned long long s_cand[2][16];
    __shared__ uint32_t s_red[16];
    __shared__ unsigned long long s_local;
    const uint32_t half = words64 / 2;
    ulonglong2* cov2 = reinterpret_cast<ulonglong2*>(s_mem);
    ulonglong2* srows = cov2 + half;
    // own rows -> shared memory; covered = row[first_row]
    uint32_t nown = 0;
    for (uint32_t r = rank; r < nrows; r += C, nown++) {
        const ulonglong2* rv = reinterpret_cast<const ulonglong2*>(rows + (size_t)r * words64);
        for (uint32_t w = threadIdx.x; w < half; w += blockDim.x) srows[(size_t)nown * half + w] = rv[w];
    }
    {
        const ulonglong2* r0 = reinterpret_cast<const ulonglong2*>(rows + (size_t)first_row * words64);
        uint32_t c = 0;
        for (uint32_t w = threadIdx.x; w < half; w += blockDim.x) {
            const ulonglong2 v = r0[w];
            cov2[w] = v;
            c += __popcll(v.x) + __popcll(v.y);
        }
        for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xFFFFFFFFu, c, o);
        if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = c;
        __syncthreads();
        if (rank == 0 && threadIdx.x == 0) {
            uint32_t t = 0;
            for (uint32_t w = 0; w < (blockDim.x >> 5); w++) t += s_red[w];
            best[0] = ((unsigned long long)t << 32) | (unsigned long long)(0xFFFFFFFFu - first_id);
        }
    }
    cluster.sync();  // every CTA's rows are resident before anybody reads them remotely
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    for (uint32_t k = 1; k <= max_iter; k++) {
        // score own rows: one warp per row
        if (threadIdx.x == 0) s_local = 0ull;
        __syncthreads();
        for (uint32_t slot = warp; slot < nown; slot += nwarp) {
            const ulonglong2* rv = srows + (size_t)slot * half;
            uint32_t c = 0;
            for (uint32_t w = lane; w < half; w += 32) {
                const ulonglong2 v = rv[w];
                const ulonglong2 cv = cov2[w];
                c += __popcll(v.x & ~cv.x) + __popcll(v.y & ~cv.y);
            }
            for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xFFFFFFFFu, c, o);
            if (lane == 0) atomicMax(&s_local, ((unsigned long long)c << 32) | (unsigned long long)(0xFFFFFFFFu - view_ids[rank + slot * C]));
        }
        __syncthreads();
        if (threadIdx.x == 0) *cluster.map_shared_rank(&s_cand[k & 1][rank], 0) = s_local;
        cluster.sync();
        // every CTA reduces the candidates posted at CTA 0
        unsigned long long b = 0ull;
        {
            const unsigned long long* cand0 = cluster.map_shared_rank(&s_cand[k & 1][0], 0);
            for (uint32_t c = 0; c < C; c++) {
                const unsigned long long v = cand0[c];
                b = v > b ? v : b;
            }
        }
        if (rank == 0 && threadIdx.x == 0) best[k] = b;
        if ((b >> 32) == 0ull) break;  // uniform across the cluster
        const uint32_t rb = row_of_id[0xFFFFFFFFu - (uint32_t)(b & 0xFFFFFFFFull)];
        const ulonglong2* rw = cluster.map_shared_rank(srows + (size_t)(rb / C) * half, rb % C);
        for (uint32_t w = threadIdx.x; w < half; w += blockDim.x) {
            const ulonglong2 v = rw[w];
            ulonglong2 cv = cov2[w];
            cv.x |= v.x;
            cv.y |= v.y;
            cov2[w] = cv;
        }
        __syncthreads();
    }
    cluster.sync();  // nobody exits while its shared memory may still be read remotely
    if (rank == 0)
        for (uint32_t w = threadIdx.x; w < words64; w += blockDim.x) cov_out[w] = s_mem[w];
}
