// tests/emu/tm_cpu_emu.h -- TEST TOOL.  Lets the CPU tests run the *product's own kernel source*
// (trophic_maelstorm_b200/csrc/tm_kernels.cuh) on host threads: one std::thread per CUDA thread,
// std::barrier for __syncthreads, a memcpy-backed stand-in for the TMA bulk copy + mbarrier.
// It is not a fallback: nothing under trophic_maelstorm_b200/ can reach it (the kernels header
// only includes it when compiled by a non-CUDA compiler, which only tests/emu/Makefile does).
#pragma once
#include <atomic>
#include <barrier>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

struct emu_dim3 { unsigned x = 1, y = 1, z = 1; };
inline thread_local emu_dim3 threadIdx, blockIdx, blockDim, gridDim;
inline thread_local std::barrier<> *emu_block_barrier = nullptr;
inline thread_local unsigned char *emu_dyn_smem_ptr = nullptr;

inline void __syncthreads() { emu_block_barrier->arrive_and_wait(); }
inline unsigned char *tm_emu_dyn_smem() { return emu_dyn_smem_ptr; }
#define TM_DYN_SMEM(name) unsigned char *name = tm_emu_dyn_smem()

inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v)
{
    return __atomic_fetch_add(p, v, __ATOMIC_RELAXED);
}

inline int atomicMax(int *p, int v)
{
    int old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (old < v && !__atomic_compare_exchange_n(p, &old, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}

namespace tml {
// mbarrier stand-in.  The 64-bit word holds {completed phases : 24 | expected arrivals : 8 | pending arrivals : 8 |
// pending transaction bytes : 24}; a phase completes when both pending counts reach zero (arrivals are then re-armed),
// as the hardware object does.  Bulk copies are synchronous memcpys that complete their bytes at once; every update is
// a compare-and-swap, so several issuing threads may share one barrier.
inline uint64_t emu_mbar_pack(uint64_t phase, uint64_t count, uint64_t arr, uint64_t tx)
{
    return (phase << 40) | (count << 32) | (arr << 24) | tx;
}
inline void mbar_init(uint64_t *b, int count) { __atomic_store_n(b, emu_mbar_pack(0, (uint64_t)count, (uint64_t)count, 0), __ATOMIC_RELEASE); }
inline void mbar_init_fence() {}
inline void emu_mbar_update(uint64_t *b, int d_arr, long long d_tx)
{
    uint64_t old = __atomic_load_n(b, __ATOMIC_ACQUIRE), neu;
    do {
        uint64_t phase = old >> 40, count = (old >> 32) & 0xFF, arr = (old >> 24) & 0xFF, tx = old & 0xFFFFFF;
        arr -= (uint64_t)d_arr; tx = (uint64_t)((long long)tx + d_tx);
        if (arr == 0 && tx == 0) { phase = (phase + 1) & 0xFFFFFF; arr = count; }
        neu = emu_mbar_pack(phase, count, arr, tx);
    } while (!__atomic_compare_exchange_n(b, &old, neu, true, __ATOMIC_ACQ_REL, __ATOMIC_ACQUIRE));
}
inline void mbar_expect_tx(uint64_t *b, uint32_t bytes) { emu_mbar_update(b, 1, (long long)bytes); } // arrive.expect_tx
inline void mbar_arrive_expect_tx(uint64_t *b, uint32_t bytes) { emu_mbar_update(b, 1, (long long)bytes); }
inline void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *b)
{
    std::memcpy(dst, src, bytes);
    emu_mbar_update(b, 0, -(long long)bytes);
}
inline bool mbar_try_wait(uint64_t *b, uint32_t parity)
{
    const uint32_t phase = (uint32_t)(__atomic_load_n(b, __ATOMIC_ACQUIRE) >> 40);
    if ((phase & 1u) != parity) return true;
    std::this_thread::yield();
    return false;
}
}

// run `kernel(args)` over grid x block on host threads, blocks one after another.  One set of `block` threads walks all blocks
// (a thread per (block, thread) pair made the reduction kernels' 296 x 256 launches the bulk of the CPU suite's time); between two
// blocks everybody meets at a second barrier, whose completion step re-poisons the shared memory.
template <class K, class A>
inline void emu_launch(K kernel, const A &args, unsigned grid, unsigned block, size_t smem_bytes)
{
    std::vector<unsigned char> smem(smem_bytes + 256);
    unsigned char *sbase = (unsigned char *)(((uintptr_t)smem.data() + 127) & ~(uintptr_t)127);
    std::memset(smem.data(), 0xA5, smem.size()); // shared memory starts as garbage
    std::barrier<> bar((std::ptrdiff_t)block);
    auto repoison = [&]() noexcept { std::memset(smem.data(), 0xA5, smem.size()); };
    std::barrier<decltype(repoison)> next_block((std::ptrdiff_t)block, repoison);
    std::vector<std::thread> th;
    th.reserve(block);
    for (unsigned t = 0; t < block; t++)
        th.emplace_back([&, t] {
            threadIdx.x = t; blockDim.x = block; gridDim.x = grid;
            emu_block_barrier = &bar; emu_dyn_smem_ptr = sbase;
            for (unsigned b = 0; b < grid; b++) {
                blockIdx.x = b;
                kernel(args);
                next_block.arrive_and_wait();
            }
        });
    for (auto &x : th) x.join();
}
