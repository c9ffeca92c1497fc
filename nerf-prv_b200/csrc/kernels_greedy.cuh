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

// Fused coverage count + all-gather send side (multi-GPU, views sharded): the pass that reads every local coverage row once
// to count it also STORES the row into the gathered table of every rank -- peer memory mapped over NVLink (cudaIpc), 128-bit
// stores -- together with the view id, and the last block to finish publishes the step number in every rank's flag word.  No
// NCCL call, no extra read of the rows, no host involvement.  PeerArena describes one rank's arena (same layout everywhere):
//   flags[kMaxPeers] u32 (flag[r] = last step rank r has published here) | error u32 | 4 x { rows [G*V][words] u64, ids [G*V] u32 }
// A step uses buffer step % 4: a rank can run at most three published casts ahead of the slowest reader of its stores (each
// cast waits for the rank's own selection two steps back, which waited for everybody's publication of that step), so four
// buffers are never overwritten while read.
#ifndef PRVK_HOST_CHECK  // (system-scope PTX: device only)
constexpr int kMaxPeers = 8;
constexpr uint32_t kPeerBuffers = 4;
struct PeerArena {
    uint32_t* flags[kMaxPeers];
    uint64_t* rows[kMaxPeers];  // buffer (step % 4) of each rank's arena
    uint32_t* ids[kMaxPeers];
    uint32_t nranks, rank;
};
__global__ void __launch_bounds__(256) count_publish_kernel(const uint64_t* __restrict__ rows, uint32_t words64, uint32_t V, uint32_t* __restrict__ counts,
                                                            const uint32_t* __restrict__ view_ids, PeerArena pa, uint32_t step, uint32_t* done) {
    __shared__ uint32_t s_red[8];
    __shared__ uint32_t s_last;
    const uint32_t v = blockIdx.x;
    const ulonglong2* row = reinterpret_cast<const ulonglong2*>(rows + (size_t)v * words64);
    const size_t slot = (size_t)pa.rank * V + v;
    uint32_t c = 0;
    for (uint32_t w = threadIdx.x; w < words64 / 2; w += blockDim.x) {
        const ulonglong2 x = row[w];
        c += __popcll(x.x) + __popcll(x.y);
        for (uint32_t r = 0; r < pa.nranks; r++) reinterpret_cast<ulonglong2*>(pa.rows[r] + slot * words64)[w] = x;  // local + 7 NVLink stores
    }
    if (threadIdx.x < pa.nranks) pa.ids[threadIdx.x][slot] = view_ids[v];
    const uint32_t t = block_reduce_sum(c, s_red);
    __threadfence_system();  // this block's peer stores are visible system-wide before it is counted as done
    __syncthreads();
    if (threadIdx.x == 0) {
        counts[v] = t;
        s_last = atomicAdd(done, 1u) == V - 1 ? 1u : 0u;
    }
    __syncthreads();
    if (s_last) {  // every block's stores have been fenced: publish the step (release) in every rank's arena
        if (threadIdx.x == 0) *done = 0;
        __threadfence_system();
        if (threadIdx.x < pa.nranks) {
            uint32_t* f = pa.flags[threadIdx.x] + pa.rank;
            asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f), "r"(step) : "memory");
        }
    }
}

// all-gather receive side: wait until every rank has published `step` here (one lane per rank, acquire loads), then the
// selection may read the table.  A rank that never shows up is reported after ~20 s instead of hanging the GPU.
__global__ void __launch_bounds__(32) wait_published_kernel(const uint32_t* flags, uint32_t nranks, uint32_t step, uint32_t* error) {
    const uint32_t r = threadIdx.x;
    if (r >= nranks) return;
    const long long t0 = clock64();
    for (;;) {
        uint32_t f;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(f) : "l"(flags + r) : "memory");
        if ((int32_t)(f - step) >= 0) break;
        if (clock64() - t0 > 40000000000ll) {  // ~20 s
            atomicExch(error, 1u + r);
            break;
        }
        __nanosleep(200);
    }
}
#endif

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

// (The two kernels below need cooperative grid sync / thread-block clusters with PTX: not part of the CPU emulation of
// tests/cpp/pipeline_on_host.cpp, which defines PRVK_HOST_CHECK and runs the kernels above.)
#ifndef PRVK_HOST_CHECK
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

// ---- cluster greedy --------------------------------------------------------------------------------------------------
// PTX helpers: one-sided stores into a peer CTA's shared memory that complete a transaction count on the PEER's mbarrier
// (st.async ... mbarrier::complete_tx::bytes), so the receiver just sleeps on its own barrier: no cluster-wide barrier,
// no fences, no polling of memory.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t map_to_cta(uint32_t smem_addr, uint32_t cta) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(cta));
    return r;
}
__device__ __forceinline__ void st_async_u32(uint32_t remote_addr, uint32_t v, uint32_t remote_mbar) {
    asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.u32 [%0], %1, [%2];" ::"r"(remote_addr), "r"(v), "r"(remote_mbar) : "memory");
}
__device__ __forceinline__ void st_async_u64(uint32_t remote_addr, unsigned long long v, uint32_t remote_mbar) {
    asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.u64 [%0], %1, [%2];" ::"r"(remote_addr), "l"(v), "r"(remote_mbar) : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t mbar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(mbar), "r"(parity) : "memory");
}

__device__ __forceinline__ void named_barrier_sync(uint32_t id, uint32_t nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// warp arg-max of (gain, nid) with the row riding along: largest gain, then largest nid (= lowest view id); two hardware
// warp reductions (redux.sync) instead of five rounds of three shuffles
__device__ __forceinline__ void warp_argmax(uint32_t& gain, uint32_t& nid, uint32_t& row) {
    const uint32_t m = __reduce_max_sync(0xFFFFFFFFu, gain);
    const uint32_t n = __reduce_max_sync(0xFFFFFFFFu, gain == m ? nid : 0u);
    const uint32_t bal = __ballot_sync(0xFFFFFFFFu, gain == m && nid == n);
    row = __shfl_sync(0xFFFFFFFFu, row, __ffs(bal) - 1);
    gain = m;
    nid = n;
}

// Greedy loop inside ONE thread-block cluster, coverage table sliced by COLUMNS over the cluster's CTAs: CTA c keeps words
// [c*slice, (c+1)*slice) (128-bit units) of EVERY row and of the covered mask in its shared memory, transposed to
// tab[w][v] so that "one thread per view" reads are conflict-free.  One iteration:
//   1. every CTA scores all V views over its slice (shared memory only) and sends view v's partial gain to the CTA that
//      owns v for the reduction (v / vper): st.async into the owner's shared memory, completing bytes on the owner's mbarrier;
//   2. each owner sleeps on that mbarrier, adds the partials of its views, takes its arg-max and sends it to every CTA
//      the same way;
//   3. every CTA sleeps on its candidate mbarrier, reduces the C candidates and ORs ITS slice of the winner's row (already
//      in its shared memory) into its slice of the mask.
// No global memory traffic, no atomics, no cluster-wide barrier inside the loop: two DSMEM store latencies per iteration.
// Measured history: a row-partitioned cluster variant that pulled the winner's row from its owner through DSMEM
// (~20 B/clk) cost 3.8 us per iteration; this column-sliced layout with two cluster.sync() per iteration 2.8 us
// (barrier.cluster with release/acquire is ~1 us with remote stores in flight); the grid-barrier kernel 2.9 us.
constexpr int kGreedyClusterThreads = 512;
constexpr int kGreedyClusterMax = 16;
// parts = threads that share one view's slice (each lane of a warp a different view, so the transposed table is read
// without bank conflicts); vper = views owned per CTA for the reduction (<= blockDim).
__global__ void __launch_bounds__(kGreedyClusterThreads) greedy_cluster_kernel(const uint64_t* __restrict__ rows, uint32_t words64, uint32_t nrows,
                                                                               const uint32_t* __restrict__ view_ids, uint32_t first_row, uint32_t first_id,
                                                                               uint32_t max_iter, uint32_t slice, uint32_t vper, uint32_t parts,
                                                                               unsigned long long* best, uint64_t* cov_out) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const uint32_t rank = cluster.block_rank(), C = cluster.num_blocks();
    extern __shared__ ulonglong2 s_tab[];  // [slice][nrows], cov[slice], s_new[slice] (u128); then u32: s_in[C*parts][vper], s_pg[parts][nrows], s_nz[slice]
    __shared__ __align__(8) unsigned long long s_mbar[2];        // [0]: partial gains have arrived, [1]: candidates have arrived
    __shared__ unsigned long long s_cand[kGreedyClusterMax];     // candidates of all owners (written remotely)
    __shared__ uint32_t s_candrow[kGreedyClusterMax];
    __shared__ uint32_t s_wgain[kGreedyClusterThreads / 32], s_wnid[kGreedyClusterThreads / 32], s_wrow[kGreedyClusterThreads / 32];
    __shared__ uint32_t s_red[kGreedyClusterThreads / 32];
    __shared__ uint32_t s_nnz[2];
    ulonglong2* cov = s_tab + (size_t)slice * nrows;
    ulonglong2* s_new = cov + slice;                                  // newly covered bits of the non-zero words, compacted
    uint32_t* s_in = reinterpret_cast<uint32_t*>(s_new + slice);
    uint32_t* s_pg = s_in + (size_t)C * parts * vper;                 // running partial gain of (part, view) over this slice
    uint32_t* s_nz = s_pg + (size_t)parts * nrows;                    // word index of each entry of s_new
    const uint32_t half = words64 / 2;
    const uint32_t w_lo = rank * slice;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    const uint32_t mbar_in = smem_u32(&s_mbar[0]), mbar_cand = smem_u32(&s_mbar[1]);
    if (threadIdx.x == 0) {
        mbar_init(mbar_in, 1);
        mbar_init(mbar_cand, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // load this CTA's column slice of every row (zero beyond the row's end)
    for (uint32_t i = threadIdx.x; i < nrows * slice; i += blockDim.x) {
        const uint32_t v = i / slice, w = i - v * slice;
        ulonglong2 x = make_ulonglong2(0ull, 0ull);
        if (w_lo + w < half) x = reinterpret_cast<const ulonglong2*>(rows + (size_t)v * words64)[w_lo + w];
        s_tab[(size_t)w * nrows + v] = x;
    }
    __syncthreads();
    {   // covered = row[first_row]; best[0] = its popcount (summed over the cluster with one global atomic per CTA, once)
        uint32_t c = 0;
        for (uint32_t w = threadIdx.x; w < slice; w += blockDim.x) {
            const ulonglong2 x = s_tab[(size_t)w * nrows + first_row];
            cov[w] = x;
            c += __popcll(x.x) + __popcll(x.y);
        }
        const uint32_t t = block_reduce_sum(c, s_red);
        if (threadIdx.x == 0) {
            unsigned long long add = (unsigned long long)t << 32;
            if (rank == 0) add |= (unsigned long long)(0xFFFFFFFFu - first_id);
            atomicAdd(best, add);  // best[0] was zeroed by the host
        }
        __syncthreads();
    }
    // the view this thread owns for the reduction
    const uint32_t own_v = rank * vper + threadIdx.x;
    const bool owner_thread = threadIdx.x < vper && own_v < nrows;
    const uint32_t own_nid = owner_thread ? 0xFFFFFFFFu - view_ids[own_v] : 0u;
    const uint32_t nown = rank * vper < nrows ? min(vper, nrows - rank * vper) : 0u;
    const uint32_t bytes_in = 4u * C * parts * nown, bytes_cand = 12u * C;
    const uint32_t own_warps = (vper + 31u) >> 5;  // warps that hold owner threads
    // scoring assignment: part = which share of the slice; parts > 1: one view per thread, else views v0, v0 + blockDim, ...
    const uint32_t vp32 = (nrows + 31u) & ~31u;
    const uint32_t part = parts > 1 ? threadIdx.x / vp32 : 0u;
    const uint32_t v0 = parts > 1 ? threadIdx.x - part * vp32 : threadIdx.x;
    const bool scorer = part < parts && v0 < nrows;
    const uint32_t vper_magic = vper > 1 ? (uint32_t)((0x100000000ull + vper - 1) / vper) : 0u;  // v / vper == umulhi(v, magic) for v * vper < 2^32
    const uint32_t s_in_addr = smem_u32(s_in), s_cand_addr = smem_u32(s_cand), s_candrow_addr = smem_u32(s_candrow);
    // destination of the first view's partial gain (the only one when parts > 1)
    uint32_t dst0 = 0, dst0_bar = 0;
    if (scorer) {
        const uint32_t o = vper > 1 ? __umulhi(v0, vper_magic) : v0;
        dst0 = map_to_cta(s_in_addr + 4u * ((rank * parts + part) * vper + (v0 - o * vper)), o);
        dst0_bar = map_to_cta(mbar_in, o);
    }
    if (threadIdx.x == 0) {
        mbar_expect_tx(mbar_in, bytes_in);
        mbar_expect_tx(mbar_cand, bytes_cand);
        s_nnz[0] = s_nnz[1] = 0u;
    }
    // initial partial gains against covered = row[first_row]: the only full pass over the slice.  Afterwards
    // gain(v) -= popcount(row(v) & newly covered), and words without newly covered bits are skipped for every view.
    if (scorer) {
        for (uint32_t v = v0; v < nrows; v += blockDim.x) {
            uint32_t c = 0;
#pragma unroll 4
            for (uint32_t w = part; w < slice; w += parts) {
                const ulonglong2 x = s_tab[(size_t)w * nrows + v];
                const ulonglong2 cv = cov[w];
                c += __popcll(x.x & ~cv.x) + __popcll(x.y & ~cv.y);
            }
            s_pg[(size_t)part * nrows + v] = c;
            if (parts > 1) break;
        }
    }
    cluster.sync();  // every CTA's mbarriers and shared memory are initialised before anybody writes into them remotely
    for (uint32_t k = 1; k <= max_iter; k++) {
        const uint32_t ph = (k - 1u) & 1u;
        // 1. partial gains of all views over this slice -> their owners
        if (scorer) {
            uint32_t v = v0, dst = dst0, dst_bar = dst0_bar;
            const uint32_t nnz = s_nnz[(k - 1u) & 1u];  // words that gained covered bits in the previous iteration
            for (;;) {
                uint32_t c = s_pg[(size_t)part * nrows + v];
                for (uint32_t i = part; i < nnz; i += parts) {
                    const ulonglong2 x = s_tab[(size_t)s_nz[i] * nrows + v];
                    const ulonglong2 nw = s_new[i];
                    c -= __popcll(x.x & nw.x) + __popcll(x.y & nw.y);  // partial sums may wrap; their total does not
                }
                s_pg[(size_t)part * nrows + v] = c;
                st_async_u32(dst, c, dst_bar);
                v += blockDim.x;
                if (parts > 1 || v >= nrows) break;
                const uint32_t o = vper > 1 ? __umulhi(v, vper_magic) : v;
                dst = map_to_cta(s_in_addr + 4u * (rank * vper + (v - o * vper)), o);
                dst_bar = map_to_cta(mbar_in, o);
            }
        }
        // Block barrier: this iteration's reads of s_nnz / s_nz / s_new are finished before step 3 rewrites them.  (The
        // message chain partials -> candidates already orders them, but only through other CTAs; the barrier makes the
        // order local and visible to compute-sanitizer racecheck.  It is off the critical path: the owners below have to
        // wait for every CTA's partials anyway.)
        __syncthreads();
        // 2. owner: total gain of each owned view, arg-max (largest gain, then lowest view id; the row index rides along),
        //    sent to every CTA.  Only the warps that hold owner threads take part.
        if (warp < own_warps) {
            mbar_wait(mbar_in, ph);
            uint32_t g = 0, nid = 0, row = 0;
            if (owner_thread) {
#pragma unroll 8
                for (uint32_t src = 0; src < C * parts; src++) g += s_in[(size_t)src * vper + threadIdx.x];
                nid = own_nid;
                row = own_v;
            }
            warp_argmax(g, nid, row);
            if (own_warps > 1) {  // uniform across the block's owner warps
                if (lane == 0) { s_wgain[warp] = g; s_wnid[warp] = nid; s_wrow[warp] = row; }
                named_barrier_sync(1, own_warps * 32u);
                if (warp == 0) {
                    g = lane < own_warps ? s_wgain[lane] : 0u;
                    nid = lane < own_warps ? s_wnid[lane] : 0u;
                    row = lane < own_warps ? s_wrow[lane] : 0u;
                    warp_argmax(g, nid, row);
                }
            }
            if (warp == 0) {
                // every view's partials have been consumed: re-arm the barrier for the next iteration (peers cannot send
                // the next partials before they have this CTA's candidate, which is sent below)
                if (lane == 0) mbar_expect_tx(mbar_in, bytes_in);
                if (lane < C) {
                    const uint32_t rb = map_to_cta(mbar_cand, lane);
                    st_async_u64(map_to_cta(s_cand_addr + 8u * rank, lane), ((unsigned long long)g << 32) | nid, rb);
                    st_async_u32(map_to_cta(s_candrow_addr + 4u * rank, lane), row, rb);
                }
            }
        }
        // 3. every CTA: winner of the C candidates, OR its slice of the winner's row into its slice of the mask
        mbar_wait(mbar_cand, ph);
        uint32_t bg = 0, bnid = 0, win_row = 0;
        if (lane < C) {
            const unsigned long long cnd = s_cand[lane];
            bg = (uint32_t)(cnd >> 32);
            bnid = (uint32_t)cnd;
            win_row = s_candrow[lane];
        }
        warp_argmax(bg, bnid, win_row);
        if (rank == 0 && threadIdx.x == 0) best[k] = ((unsigned long long)bg << 32) | bnid;
        if (bg == 0u) break;  // uniform across the cluster: nothing left to gain
        if (threadIdx.x == 0) s_nnz[(k + 1u) & 1u] = 0u;  // the list read by this iteration's scoring; refilled next iteration
        for (uint32_t w = threadIdx.x; w < slice; w += blockDim.x) {
            const ulonglong2 x = s_tab[(size_t)w * nrows + win_row];
            const ulonglong2 cv = cov[w];
            const ulonglong2 nw = make_ulonglong2(x.x & ~cv.x, x.y & ~cv.y);
            if (nw.x | nw.y) {
                cov[w] = make_ulonglong2(cv.x | x.x, cv.y | x.y);
                const uint32_t i = atomicAdd(&s_nnz[k & 1u], 1u);
                s_nz[i] = w;
                s_new[i] = nw;
            }
        }
        // every thread has read the candidates; only now may this CTA's next partials let the owners overwrite them
        __syncthreads();
        if (threadIdx.x == 0) mbar_expect_tx(mbar_cand, bytes_cand);
    }
    cluster.sync();  // nobody exits while a peer may still write into its shared memory
    for (uint32_t w = threadIdx.x; w < slice; w += blockDim.x)
        if (w_lo + w < half) reinterpret_cast<ulonglong2*>(cov_out)[w_lo + w] = cov[w];
}
#endif  // PRVK_HOST_CHECK
