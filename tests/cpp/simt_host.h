// simt_host.h — TEST INFRASTRUCTURE.  Runs the library's CUDA kernels on the CPU, one OS thread per CUDA thread,
// so that their control flow (warp votes, shuffles, work sharing between lanes, hand-back lists) can be checked
// against the oracle in the container that has no GPU, and so that a kernel whose lanes disagree about a
// warp-wide primitive dead-locks HERE (a test time-out) and not on the GPU box.
//
// Not part of the product: nothing under physkit_b200/ or include/ includes this file, and the product has no
// CPU path (tests/test_abi_cpu.py).  The arithmetic is the kernels' own source compiled by g++ with
// -ffp-contract=off for SSE2 (no FMA), which is what -fmad=false gives on the device; pk_div_by_rcp's explicit
// fma calls go to libm's correctly rounded fma.
//
// Only what the kernels under test use is provided: full-mask votes / shuffles / match.any / __syncwarp (a std::barrier
// over the 32 threads of a warp), __syncthreads (a barrier over the block), atomicAdd, the bit and rounding intrinsics.  `__shared__` becomes a
// function-local static, so ONE block of a given kernel instance runs at a time (blocks are run one after the
// other; the kernels under test are persistent and take their work from an atomic cursor, so a single block does
// all of it).
#pragma once

#include <cuda_runtime.h> // vector types and make_*; under g++ the __device__ / __global__ annotations vanish

#include <atomic>
#include <barrier>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

#include <unistd.h>

#undef __shared__
#define __shared__ static
#ifndef __launch_bounds__
#define __launch_bounds__(...)
#endif
#ifndef __noinline__
#define __noinline__ __attribute__((noinline))
#endif
#ifndef __maxnreg__
#define __maxnreg__(...)
#endif

namespace simt
{
struct Idx
{
    unsigned x = 0, y = 0, z = 0;
};
struct Warp
{
    std::barrier<> bar{32};
    uint64_t buf[32];
    // sub-groups of a warp (the peers of a match.any): a rendezvous per group, keyed by its lowest lane
    uint64_t sub_buf[32];
    std::atomic<unsigned> sub_arrived[32];
    std::atomic<unsigned> sub_left[32];
    Warp()
    {
        for (int i = 0; i < 32; ++i)
        {
            sub_arrived[i].store(0);
            sub_left[i].store(0);
        }
    }
};
inline thread_local Idx tl_threadIdx, tl_blockIdx, tl_blockDim, tl_gridDim;
inline thread_local Warp *tl_warp = nullptr;
inline thread_local int tl_lane = 0;
inline thread_local std::barrier<> *tl_block_bar = nullptr; // __syncthreads

// The OS threads that stand in for the CUDA threads of a block are kept between launches: creating them anew for every
// block of every launch (a radix pass is a grid of 256 scan blocks of 256 threads) cost more than the kernels.
class Pool
{
public:
    static Pool &get()
    {
        static Pool p;
        return p;
    }
    // job(t) on workers t = 0 … block−1, returns when all of them are done
    void run(unsigned block, const std::function<void(unsigned)> &job)
    {
        ensure(block);
        job_ = &job;
        active_ = block;
        start_->arrive_and_wait();
        done_->arrive_and_wait();
    }
    ~Pool() { stop(); }

private:
    void ensure(unsigned n)
    {
        if (n <= size_) return;
        stop();
        size_ = n < 256u ? 256u : n;
        owner_ = getpid();
        start_ = std::make_unique<std::barrier<>>(static_cast<std::ptrdiff_t>(size_) + 1);
        done_ = std::make_unique<std::barrier<>>(static_cast<std::ptrdiff_t>(size_) + 1);
        quit_ = false;
        for (unsigned t = 0; t < size_; ++t)
            th_.emplace_back(
                [this, t]()
                {
                    for (;;)
                    {
                        start_->arrive_and_wait();
                        if (quit_) return;
                        if (t < active_) (*job_)(t);
                        done_->arrive_and_wait();
                    }
                });
    }
    void stop()
    {
        if (!size_) return;
        if (getpid() != owner_)
        {
            // a fork()ed child (multiprocessing in the test session) inherits this object but none of its threads:
            // nothing to wake, nothing to join
            for (auto &x : th_) x.detach();
            th_.clear();
            size_ = 0;
            return;
        }
        quit_ = true;
        start_->arrive_and_wait();
        for (auto &x : th_) x.join();
        th_.clear();
        size_ = 0;
    }
    std::vector<std::thread> th_;
    std::unique_ptr<std::barrier<>> start_, done_;
    const std::function<void(unsigned)> *job_ = nullptr;
    unsigned size_ = 0, active_ = 0;
    bool quit_ = false;
    pid_t owner_ = 0;
};

// Blocks run one after the other (their __shared__ arrays are function-local statics), each thread of the pool playing
// the same threadIdx in every block; the block barrier also separates consecutive blocks.
template <class K> void launch(unsigned grid, unsigned block, K &&kernel)
{
    assert(block % 32 == 0);
    const unsigned nwarps = (block + 31) / 32;
    std::vector<Warp> warps(nwarps);
    std::barrier<> block_bar(static_cast<std::ptrdiff_t>(block));
    const std::function<void(unsigned)> job = [&](unsigned t)
    {
        tl_threadIdx = Idx{t, 0, 0};
        tl_blockDim = Idx{block, 1, 1};
        tl_gridDim = Idx{grid, 1, 1};
        tl_warp = &warps[t / 32];
        tl_lane = static_cast<int>(t % 32);
        tl_block_bar = &block_bar;
        for (unsigned b = 0; b < grid; ++b)
        {
            tl_blockIdx = Idx{b, 0, 0};
            kernel();
            block_bar.arrive_and_wait();
        }
    };
    Pool::get().run(block, job);
}
} // namespace simt

#define threadIdx simt::tl_threadIdx
#define blockIdx simt::tl_blockIdx
#define blockDim simt::tl_blockDim
#define gridDim simt::tl_gridDim

constexpr unsigned SIMT_FULL = 0xFFFFFFFFu;

// Opportunistic groups (__activemask) are groups of ONE here: every thread is its own coalesced group, which the
// kernels must (and do) handle like any other group size.  Votes and shuffles over that one-lane mask are local.
inline unsigned __activemask() { return 1u << simt::tl_lane; }

inline unsigned __ballot_sync(unsigned mask, int pred)
{
    if (mask == (1u << simt::tl_lane)) return pred ? mask : 0u;
    assert(mask == SIMT_FULL);
    simt::Warp &w = *simt::tl_warp;
    w.buf[simt::tl_lane] = pred ? 1u : 0u;
    w.bar.arrive_and_wait();
    unsigned r = 0;
    for (int i = 0; i < 32; ++i) r |= static_cast<unsigned>(w.buf[i]) << i;
    w.bar.arrive_and_wait();
    return r;
}
// Exchange inside a proper sub-group of the warp (every lane of `mask`, and only those, calls with the same mask; groups
// that are in flight at the same time are disjoint, as the peers of a match.any are): all write, the last to arrive
// opens the gate, all read, the last to leave resets the rendezvous for the group's next exchange.
inline uint64_t simt_subgroup_exchange(unsigned mask, uint64_t bits, int src)
{
    simt::Warp &w = *simt::tl_warp;
    const int key = __builtin_ctz(mask);
    const unsigned k = static_cast<unsigned>(__builtin_popcount(mask));
    while (w.sub_left[key].load(std::memory_order_acquire) != 0) std::this_thread::yield(); // the previous exchange is still draining
    w.sub_buf[simt::tl_lane] = bits;
    w.sub_arrived[key].fetch_add(1, std::memory_order_acq_rel);
    while (w.sub_arrived[key].load(std::memory_order_acquire) < k) std::this_thread::yield();
    const uint64_t got = (src >= 0 && src < 32 && ((mask >> src) & 1u)) ? w.sub_buf[src] : bits;
    if (w.sub_left[key].fetch_add(1, std::memory_order_acq_rel) + 1 == k)
    {
        w.sub_arrived[key].store(0, std::memory_order_release);
        w.sub_left[key].store(0, std::memory_order_release);
    }
    return got;
}
template <class T> inline T simt_exchange(unsigned mask, T v, int src)
{
    if (mask == (1u << simt::tl_lane)) return v;
    static_assert(sizeof(T) <= 8, "shuffle of at most 64 bits");
    if (mask != SIMT_FULL)
    {
        assert((mask >> simt::tl_lane) & 1u);
        uint64_t b = 0;
        std::memcpy(&b, &v, sizeof(T));
        const uint64_t got = simt_subgroup_exchange(mask, b, src);
        T out;
        std::memcpy(&out, &got, sizeof(T));
        return out;
    }
    simt::Warp &w = *simt::tl_warp;
    uint64_t bits = 0;
    std::memcpy(&bits, &v, sizeof(T));
    w.buf[simt::tl_lane] = bits;
    w.bar.arrive_and_wait();
    uint64_t got = (src >= 0 && src < 32) ? w.buf[src] : bits;
    w.bar.arrive_and_wait();
    T out;
    std::memcpy(&out, &got, sizeof(T));
    return out;
}
template <class T> inline T __shfl_sync(unsigned mask, T v, int src) { return simt_exchange(mask, v, src & 31); }
template <class T> inline T __shfl_up_sync(unsigned mask, T v, unsigned delta)
{
    const int src = simt::tl_lane - static_cast<int>(delta);
    return simt_exchange(mask, v, src < 0 ? simt::tl_lane : src);
}
template <class T> inline T __shfl_xor_sync(unsigned mask, T v, int lane_mask) { return simt_exchange(mask, v, (simt::tl_lane ^ lane_mask) & 31); }
inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0u; }
inline void __syncwarp(unsigned mask = SIMT_FULL)
{
    assert(mask == SIMT_FULL);
    simt::tl_warp->bar.arrive_and_wait();
}
// Block barrier: every thread of the block must reach it (the sort kernels; blocks must be a whole number of warps'
// worth of threads that all stay alive until their last barrier, which the kernels under test satisfy).
inline void __syncthreads() { simt::tl_block_bar->arrive_and_wait(); }
// lanes of the (full) warp that hold the same value
template <class T> inline unsigned __match_any_sync(unsigned mask, T v)
{
    if (mask == (1u << simt::tl_lane)) return mask;
    assert(mask == SIMT_FULL);
    static_assert(sizeof(T) <= 8, "match of at most 64 bits");
    simt::Warp &w = *simt::tl_warp;
    uint64_t bits = 0;
    std::memcpy(&bits, &v, sizeof(T));
    w.buf[simt::tl_lane] = bits;
    w.bar.arrive_and_wait();
    unsigned r = 0;
    for (int i = 0; i < 32; ++i)
        if (w.buf[i] == bits) r |= 1u << i;
    w.bar.arrive_and_wait();
    return r;
}

template <class T> inline T __ldg(const T *p) { return *p; }
inline float rsqrtf(float x) { return 1.0f / std::sqrt(x); }
inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
inline int __ffs(int x) { return __builtin_ffs(x); }
inline int __ffsll(long long x) { return __builtin_ffsll(x); }
inline float __int_as_float(int x)
{
    float f;
    std::memcpy(&f, &x, 4);
    return f;
}
inline float __double2float_rd(double d)
{
    float f = static_cast<float>(d);
    if (static_cast<double>(f) > d) f = std::nextafterf(f, -INFINITY);
    return f;
}
inline float __double2float_ru(double d)
{
    float f = static_cast<float>(d);
    if (static_cast<double>(f) < d) f = std::nextafterf(f, INFINITY);
    return f;
}
inline double __longlong_as_double(long long x)
{
    double d;
    std::memcpy(&d, &x, 8);
    return d;
}
inline double __drcp_rn(double x) { return 1.0 / x; }
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline unsigned atomicAdd(unsigned *p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline unsigned atomicMin(unsigned *p, unsigned v)
{
    unsigned old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (v < old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}
inline unsigned atomicMax(unsigned *p, unsigned v)
{
    unsigned old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (v > old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}
inline int atomicExch(int *p, int v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
// loads with cache hints are plain loads here; the hierarchy kernel's hand-over reads go through an atomic load so
// that the host's memory model gives what __threadfence + ld.cg give on the device
template <class T> inline T __ldcg(const T *p)
{
    __atomic_thread_fence(__ATOMIC_SEQ_CST);
    return *p;
}
template <class T> inline T __ldcs(const T *p) { return *p; }
inline unsigned __float_as_uint(float f)
{
    unsigned u;
    std::memcpy(&u, &f, 4);
    return u;
}
inline float __uint_as_float(unsigned u)
{
    float f;
    std::memcpy(&f, &u, 4);
    return f;
}
// a + b rounded towards −inf / +inf: round to nearest, then step by the sign of the exact error (TwoSum)
inline float simt_fadd_dir(float a, float b, bool up)
{
    const float s = a + b;
    if (!(std::fabs(s) < INFINITY) || s != s) return s;
    const float bb = s - a;
    const float err = (a - (s - bb)) + (b - bb);
    if (up && err > 0.f) return std::nextafterf(s, INFINITY);
    if (!up && err < 0.f) return std::nextafterf(s, -INFINITY);
    return s;
}
inline float __fadd_rd(float a, float b) { return simt_fadd_dir(a, b, false); }
inline float __fadd_ru(float a, float b) { return simt_fadd_dir(a, b, true); }
inline unsigned atomicOr(unsigned *p, unsigned v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline unsigned long long atomicCAS(unsigned long long *p, unsigned long long expected, unsigned long long desired)
{
    __atomic_compare_exchange_n(p, &expected, desired, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED);
    return expected; // the old value, as the device intrinsic returns it
}
