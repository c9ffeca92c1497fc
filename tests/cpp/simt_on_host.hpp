// simt_on_host.hpp -- TEST INFRASTRUCTURE: just enough of the CUDA execution model to run this repo's kernels, as written, on
// the CPU.  One OS thread per CUDA thread of a block (blocks run one after the other), so the kernels' own barriers, warp
// collectives and atomics do the synchronising:
//   threadIdx / blockIdx / blockDim / gridDim   thread-local variables
//   __shared__                                  function-local static storage (one block is resident at a time)
//   __syncthreads()                             std::barrier over the block's live threads
//   __ballot_sync / __shfl*_sync / __reduce_*   exchange through a 32-slot buffer per warp between two warp barriers
//   atomicAdd / atomicOr / atomicMax / atomicMin  GCC __atomic builtins
// A thread that returns from the kernel drops out of its barriers, as an exited CUDA thread does.  Full-mask collectives only
// (all this repo uses).  Not emulated: clusters, mbarrier / st.async, cooperative grid sync -- the cluster and grid-barrier
// greedy kernels stay with the GPU tests.  Slow (a barrier costs microseconds): for small images and grids of a few blocks.
#pragma once
#include <barrier>
#include <cstdint>
#include <memory>
#include <thread>
#include <vector>

#include <cuda_runtime.h>  // dim3, uint3, make_uint4 ... (host side of the headers)

namespace simt {
struct Warp {
    std::barrier<> bar{32};
    unsigned long long slot[32];
    explicit Warp(int lanes) : bar(lanes) {}
};
struct Block {
    std::barrier<> bar;
    std::vector<std::unique_ptr<Warp>> warps;
    explicit Block(int threads) : bar(threads) {
        for (int t = 0; t < threads; t += 32) warps.emplace_back(new Warp(std::min(32, threads - t)));
    }
};
inline thread_local Block* tl_block = nullptr;
inline thread_local Warp* tl_warp = nullptr;
inline thread_local int tl_lane = 0;
}  // namespace simt

static thread_local uint3 threadIdx, blockIdx;
static thread_local dim3 blockDim, gridDim;

#define __shared__ static
#undef __launch_bounds__
#define __launch_bounds__(...)

static inline void __syncthreads() { simt::tl_block->bar.arrive_and_wait(); }
static inline void __syncwarp(unsigned = 0xFFFFFFFFu) { simt::tl_warp->bar.arrive_and_wait(); }

// every lane deposits `mine`, all lanes read what they need, and a second barrier frees the slots for the next collective
template <class Read>
static inline auto simt_exchange(unsigned long long mine, Read read) {
    simt::Warp& w = *simt::tl_warp;
    w.slot[simt::tl_lane] = mine;
    w.bar.arrive_and_wait();
    auto r = read(w.slot);
    w.bar.arrive_and_wait();
    return r;
}
static inline uint32_t __ballot_sync(unsigned, bool pred) {
    return simt_exchange(pred ? 1ull : 0ull, [](const unsigned long long* s) {
        uint32_t m = 0;
        for (int l = 0; l < 32; l++) m |= (uint32_t)(s[l] & 1ull) << l;
        return m;
    });
}
template <typename T>
static inline T __shfl_sync(unsigned, T v, int src) {
    return (T)simt_exchange((unsigned long long)v, [src](const unsigned long long* s) { return s[src & 31]; });
}
template <typename T>
static inline T __shfl_down_sync(unsigned, T v, int delta) {
    const int lane = simt::tl_lane;
    return (T)simt_exchange((unsigned long long)v, [=](const unsigned long long* s) { return lane + delta < 32 ? s[lane + delta] : (unsigned long long)v; });
}
template <typename T>
static inline T __shfl_up_sync(unsigned, T v, int delta) {
    const int lane = simt::tl_lane;
    return (T)simt_exchange((unsigned long long)v, [=](const unsigned long long* s) { return lane >= delta ? s[lane - delta] : (unsigned long long)v; });
}
static inline uint32_t __reduce_add_sync(unsigned, uint32_t v) {
    return simt_exchange(v, [](const unsigned long long* s) {
        uint32_t t = 0;
        for (int l = 0; l < 32; l++) t += (uint32_t)s[l];
        return t;
    });
}
static inline uint32_t __reduce_max_sync(unsigned, uint32_t v) {
    return simt_exchange(v, [](const unsigned long long* s) {
        uint32_t t = 0;
        for (int l = 0; l < 32; l++) t = std::max(t, (uint32_t)s[l]);
        return t;
    });
}

static inline uint32_t atomicAdd(uint32_t* p, uint32_t v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline uint32_t atomicOr(uint32_t* p, uint32_t v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
static inline unsigned long long atomicMax(unsigned long long* p, unsigned long long v) {
    unsigned long long old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (old < v && !__atomic_compare_exchange_n(p, &old, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {
    }
    return old;
}
static inline unsigned long long atomicMin(unsigned long long* p, unsigned long long v) {
    unsigned long long old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (old > v && !__atomic_compare_exchange_n(p, &old, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {
    }
    return old;
}
template <typename T>
static inline T __ldcg(const T* p) { return *p; }
static inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }

// <<<grid, block>>>: one OS thread per CUDA thread of a block, reused for every block of the launch; blocks run one after the
// other (a launch-wide barrier separates them, because __shared__ storage is reused), each with its own barrier objects
template <class Kernel>
static void simt_launch(dim3 grid, dim3 block, Kernel kernel) {
    const int threads = (int)(block.x * block.y * block.z);
    const size_t nblocks = (size_t)grid.x * grid.y * grid.z;
    std::barrier<> between_blocks(threads);
    std::vector<std::unique_ptr<simt::Block>> blocks(2);  // double-buffered: block b+1 is built while stragglers leave block b
    blocks[0].reset(new simt::Block(threads));
    std::vector<std::thread> pool;
    pool.reserve(threads);
    for (int t = 0; t < threads; t++)
        pool.emplace_back([&, t] {
            blockDim = block;
            gridDim = grid;
            threadIdx = {(unsigned)t % block.x, ((unsigned)t / block.x) % block.y, (unsigned)t / (block.x * block.y)};
            simt::tl_lane = t & 31;
            for (size_t b = 0; b < nblocks; b++) {
                simt::Block& blk = *blocks[b & 1];
                blockIdx = {(unsigned)(b % grid.x), (unsigned)((b / grid.x) % grid.y), (unsigned)(b / ((size_t)grid.x * grid.y))};
                simt::tl_block = &blk;
                simt::tl_warp = blk.warps[t >> 5].get();
                if (t == 0 && b + 1 < nblocks) blocks[(b + 1) & 1].reset(new simt::Block(threads));
                kernel();
                // an exited thread no longer takes part in the block's barriers
                simt::tl_warp->bar.arrive_and_drop();
                blk.bar.arrive_and_drop();
                between_blocks.arrive_and_wait();
            }
        });
    for (auto& th : pool) th.join();
}
