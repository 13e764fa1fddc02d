// TEST INFRASTRUCTURE -- a minimal host-side execution model for the
// generated CUDA sources, so that the *actual kernel text* the product
// generates can be executed (functionally, not fast) in the GPU-less test
// suite.  One OS thread per CUDA thread, CTAs one after another,
// __syncthreads as a barrier, shared memory as one static buffer, TMA bulk
// copies as synchronous memcpy tracked by emulated mbarriers.  Never used
// by the product: tests/ inject it through monkeypatching only.
#pragma once

#include <atomic>
#include <cmath>
#include <condition_variable>
#include <cstring>
#include <map>
#include <mutex>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __align__(n) alignas(n)
#define __constant__ static const
#define __shared__ static      // one CTA runs at a time

struct emu_dim3 { unsigned x, y, z; };
// Vector types carry CUDA's alignment requirements; the kernels are built
// with -fsanitize=alignment (trapping), so a 16-byte access at an address
// the device would fault on ends the process here too
struct alignas(16) int4 { int x, y, z, w; };
struct alignas(16) double2 { double x, y; };
struct alignas(8) float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };

static thread_local emu_dim3 threadIdx, blockIdx;
static emu_dim3 blockDim, gridDim;

alignas(128) unsigned char smem_raw[256*1024];

using std::fma;
using std::sqrt;
using std::fabs;
using std::pow;
using std::fmin;
using std::fmax;

template <class T> static inline T __ldg(const T *p) { return *p; }
// (a warp broadcast of a value every lane already holds)
template <class T> static inline T __shfl_sync(unsigned, T v, int) { return v; }

// ---- atomics / bit casts -------------------------------------------------------
static std::mutex emu_atomic_mutex;

static inline void __threadfence()
{
    std::atomic_thread_fence(std::memory_order_seq_cst);
}

template <class T> static inline T atomicAdd(T *p, T v)
{
    std::lock_guard<std::mutex> lk(emu_atomic_mutex);
    T old = *p;
    *p = old + v;
    return old;
}

template <class T> static inline T atomicCAS(T *p, T cmp, T val)
{
    std::lock_guard<std::mutex> lk(emu_atomic_mutex);
    T old = *p;
    if (old == cmp)
        *p = val;
    return old;
}

static inline double __longlong_as_double(long long v)
{ double d; std::memcpy(&d, &v, 8); return d; }
static inline long long __double_as_longlong(double d)
{ long long v; std::memcpy(&v, &d, 8); return v; }
static inline float __uint_as_float(unsigned v)
{ float f; std::memcpy(&f, &v, 4); return f; }
static inline unsigned __float_as_uint(float f)
{ unsigned v; std::memcpy(&v, &f, 4); return v; }

// ---- CTA-wide barrier ---------------------------------------------------------
struct EmuBarrier
{
    std::mutex m;
    std::condition_variable cv;
    unsigned n = 0, waiting = 0, gen = 0;

    void wait()
    {
        std::unique_lock<std::mutex> lk(m);
        const unsigned g = gen;
        if (++waiting == n)
        {
            waiting = 0;
            gen++;
            cv.notify_all();
        }
        else
            cv.wait(lk, [&] { return gen != g; });
    }
};

static EmuBarrier emu_cta_barrier;
static inline void __syncthreads() { emu_cta_barrier.wait(); }

// ---- named barriers (bar.sync id, n) and one-shot flags -----------------------
static EmuBarrier emu_named_barriers[16];

static inline void bar_sync_named(int id, int n)
{
    EmuBarrier &b = emu_named_barriers[id];
    {
        std::lock_guard<std::mutex> lk(b.m);
        b.n = (unsigned) n;
    }
    b.wait();
}

static inline void flag_set(int *f)
{
    reinterpret_cast<std::atomic<int> *>(f)->store(1, std::memory_order_release);
}

static inline void flag_wait(int *f)
{
    while (!reinterpret_cast<std::atomic<int> *>(f)->load(std::memory_order_acquire))
        std::this_thread::yield();
}

// ---- warp-level matrix product (mma.sync m8n8k4, FP64) --------------------------
// Every thread of the CTA must take part the same number of times (the
// kernels that use it have uniform control flow): fragments are parked in
// a CTA-wide scratch array between two barriers.
static double emu_mma_a[1024], emu_mma_b[1024];

template <class T>
static inline void mma_m8n8k4(T &c0, T &c1, T a, T b)
{
    const unsigned tid = threadIdx.x, w0 = tid & ~31u, lane = tid & 31u;
    emu_mma_a[tid] = a;
    emu_mma_b[tid] = b;
    emu_cta_barrier.wait();
    const unsigned g = lane >> 2, t = lane & 3u;
    for (unsigned k = 0; k < 4; k++)
    {
        // A[g][k] sits in lane 4g + k, B[k][n] in lane 4n + k
        c0 += (T) (emu_mma_a[w0 + 4*g + k]*emu_mma_b[w0 + 4*(2*t) + k]);
        c1 += (T) (emu_mma_a[w0 + 4*g + k]*emu_mma_b[w0 + 4*(2*t + 1) + k]);
    }
    emu_cta_barrier.wait();
}

// ---- mbarrier + TMA bulk copy ---------------------------------------------------
// The 8-byte shared-memory slot holds {completed phases, pending bytes}
struct EmuMbar { std::atomic<unsigned> completed; std::atomic<unsigned> pending; };
static_assert(sizeof(EmuMbar) == 8, "mbarrier slot");

static inline unsigned smem_u32(const void *p) { return 0; }

// Arrivals a phase expects beyond the one that carries the byte count
// (``mbar_init(bar, 1 + k)``: k threads signal the completion of their
// cp.async copies); modelled as k further bytes
static std::mutex emu_mbar_mutex;
static std::map<const void *, unsigned> emu_mbar_extra;

static inline void mbar_init(unsigned long long *bar, unsigned n)
{
    EmuMbar *b = reinterpret_cast<EmuMbar *>(bar);
    b->completed.store(0);
    b->pending.store(0);
    std::lock_guard<std::mutex> lk(emu_mbar_mutex);
    emu_mbar_extra[bar] = n - 1;
}

static inline void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    // (copies issued by other threads may have completed already: the
    // byte count runs negative until the expectation arrives)
    EmuMbar *b = reinterpret_cast<EmuMbar *>(bar);
    {
        std::lock_guard<std::mutex> lk(emu_mbar_mutex);
        bytes += emu_mbar_extra[bar];
    }
    if (b->pending.fetch_add(bytes) + bytes == 0)
        b->completed.fetch_add(1, std::memory_order_release);
}

static inline void mbar_wait(unsigned long long *bar, unsigned parity)
{
    EmuMbar *b = reinterpret_cast<EmuMbar *>(bar);
    while ((b->completed.load(std::memory_order_acquire) & 1) == parity)
        std::this_thread::yield();
}

static inline void tma_load_1d(void *dst, const void *src, unsigned bytes,
                               unsigned long long *bar)
{
    EmuMbar *b = reinterpret_cast<EmuMbar *>(bar);
    std::memcpy(dst, src, bytes);
    if (b->pending.fetch_sub(bytes) == bytes)
        b->completed.fetch_add(1, std::memory_order_release);
}

// ---- per-thread asynchronous copies (executed at once) ---------------------------
template <int N>
static inline void cp_async(void *dst, const void *src)
{
    std::memcpy(dst, src, N);
}

static inline void cp_async16(void *dst, const void *src)
{
    std::memcpy(dst, src, 16);
}

static inline void cp_async_wait_all() {}

// All of this thread's copies so far have landed (they are executed at
// once here): one arrival on the barrier
static inline void cp_async_mbar_arrive(unsigned long long *bar);


static inline void cp_async_mbar_arrive(unsigned long long *bar)
{
    EmuMbar *b = reinterpret_cast<EmuMbar *>(bar);
    if (b->pending.fetch_sub(1) == 1)
        b->completed.fetch_add(1, std::memory_order_release);
}

// ---- launcher -----------------------------------------------------------------
template <class F>
static void emu_launch(F &&body, unsigned gx, unsigned gy, unsigned gz,
                       unsigned bx, unsigned by, unsigned bz, bool threaded)
{
    gridDim = {gx, gy, gz};
    blockDim = {bx, by, bz};
    const unsigned nth = bx*by*bz;

    for (unsigned cz = 0; cz < gz; cz++)
    for (unsigned cy = 0; cy < gy; cy++)
    for (unsigned cx = 0; cx < gx; cx++)
    {
        auto run = [&](unsigned t)
        {
            blockIdx = {cx, cy, cz};
            threadIdx = {t % bx, (t / bx) % by, t / (bx*by)};
            body();
        };

        if (!threaded)
        {
            for (unsigned t = 0; t < nth; t++)
                run(t);
        }
        else
        {
            emu_cta_barrier.n = nth;
            emu_cta_barrier.waiting = 0;
            std::vector<std::thread> pool;
            pool.reserve(nth);
            for (unsigned t = 0; t < nth; t++)
                pool.emplace_back(run, t);
            for (auto &th : pool)
                th.join();
        }
    }
}
