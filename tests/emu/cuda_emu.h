/*
 * cuda_emu.h -- TEST-ONLY single-process emulator for the subset of CUDA the kernels in
 * lordfast_b200/csrc/lf_kernels.cuh use.  This container has no GPU; the emulator lets the very
 * same kernel source be compiled with g++ and stepped on the CPU so that logic errors are found
 * before a gpurun call is spent.  It is NEVER linked into liblfgpu.so and is not a fallback: the
 * product library is nvcc-only and fails loudly without a device.
 *
 * Model: a block's threads are ucontext fibers scheduled round-robin on one OS thread; a fiber
 * yields inside __syncthreads() and inside every *_sync warp primitive, which are implemented as
 * (block / warp) barriers around an exchange buffer.  Warp primitives require every live lane of
 * the warp to take part (the kernels only use full masks).  Blocks run one after another.
 */
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <functional>
#include <vector>

#define LF_EMU 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __restrict__ __restrict
#define __launch_bounds__(...)
#define __shared__ static
#define __constant__ static

struct emu_dim3 { unsigned x, y, z; emu_dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
typedef emu_dim3 dim3;
struct uint2 { uint32_t x, y; };
struct uint4 { uint32_t x, y, z, w; };
struct int2 { int x, y; };
static inline uint2 make_uint2(uint32_t a, uint32_t b) { uint2 r = { a, b }; return r; }
static inline uint4 make_uint4(uint32_t a, uint32_t b, uint32_t c, uint32_t d) { uint4 r = { a, b, c, d }; return r; }

namespace emu {
/* minimal x86-64 context switch (callee-saved registers + stack pointer); ucontext's swapcontext makes a
 * sigprocmask system call per switch, which made shuffle-heavy kernels take minutes to emulate */
struct Ctx { void *sp; };
extern "C" void emu_switch(Ctx *from, Ctx *to);

struct Fiber {
    Ctx ctx;
    char *stack = nullptr;
    bool done = false;
    emu_dim3 tid;
    uint64_t xchg = 0;       /* value offered to a warp exchange */
    unsigned warp_gen = 0;   /* warp barrier generation this fiber has reached */
    unsigned block_gen = 0;
};
struct Block {
    std::vector<Fiber> fibers;
    Ctx sched;
    int cur = -1;
    emu_dim3 bid, bdim, gdim;
    std::vector<unsigned> warp_arrived, warp_gen; /* per warp */
    unsigned block_arrived = 0, block_gen = 0, live = 0;
    std::vector<unsigned> warp_live;
    std::function<void()> body;
    char *dyn_smem = nullptr;
};
extern Block *g_blk;
extern size_t g_stack_bytes;
extern unsigned long g_progress;

inline Fiber &self() { return g_blk->fibers[(size_t)g_blk->cur]; }
inline void yield() { emu_switch(&self().ctx, &g_blk->sched); }
inline int lane_id() { return g_blk->cur & 31; }
inline int warp_id() { return g_blk->cur >> 5; }

inline void warp_barrier()
{
    Block *b = g_blk; int w = warp_id(); Fiber &f = self();
    unsigned gen = b->warp_gen[(size_t)w];
    if (++b->warp_arrived[(size_t)w] == b->warp_live[(size_t)w]) { b->warp_arrived[(size_t)w] = 0; b->warp_gen[(size_t)w]++; g_progress++; }
    (void)f;
    while (b->warp_gen[(size_t)w] == gen) yield();
}
inline void block_barrier()
{
    Block *b = g_blk; unsigned gen = b->block_gen;
    if (++b->block_arrived == b->live) { b->block_arrived = 0; b->block_gen++; g_progress++; }
    while (b->block_gen == gen) yield();
}
/* every live lane publishes v, then reads lane src's value */
inline uint64_t warp_exchange(uint64_t v, int src)
{
    self().xchg = v;
    warp_barrier();
    Block *b = g_blk; size_t idx = (size_t)(warp_id() * 32 + src);
    uint64_t r = (src >= 0 && src < 32 && idx < b->fibers.size() && !b->fibers[idx].done) ? b->fibers[idx].xchg : v;
    warp_barrier();
    return r;
}
inline unsigned warp_ballot(int pred)
{
    self().xchg = pred ? 1 : 0;
    warp_barrier();
    Block *b = g_blk; unsigned m = 0;
    for (int l = 0; l < 32; l++) {
        size_t idx = (size_t)(warp_id() * 32 + l);
        if (idx < b->fibers.size() && !b->fibers[idx].done && b->fibers[idx].xchg) m |= 1u << l;
    }
    warp_barrier();
    return m;
}
void fiber_entry();
void launch(emu_dim3 grid, emu_dim3 block, size_t smem, const std::function<void()> &body);
} // namespace emu

#define threadIdx (emu::self().tid)
#define blockIdx (emu::g_blk->bid)
#define blockDim (emu::g_blk->bdim)
#define gridDim (emu::g_blk->gdim)
#define warpSize 32

static inline void __syncthreads() { emu::block_barrier(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu::warp_barrier(); }
static inline void __threadfence() {}
static inline void __threadfence_block() {}

/* width w (a power of two): the exchange stays inside the lane's segment of w lanes */
template <typename T> static inline T __shfl_sync(unsigned, T v, int src, int w = 32)
{ int l = emu::lane_id(); int s = (l & ~(w - 1)) | (src & (w - 1)); uint64_t x = 0; memcpy(&x, &v, sizeof(T)); x = emu::warp_exchange(x, s); T r; memcpy(&r, &x, sizeof(T)); return r; }
template <typename T> static inline T __shfl_up_sync(unsigned, T v, unsigned d, int w = 32)
{ int l = emu::lane_id(); int lw = l & (w - 1); int src = lw >= (int)d ? l - (int)d : l; uint64_t x = 0; memcpy(&x, &v, sizeof(T)); x = emu::warp_exchange(x, src); T r; memcpy(&r, &x, sizeof(T)); return r; }
template <typename T> static inline T __shfl_down_sync(unsigned, T v, unsigned d, int w = 32)
{ int l = emu::lane_id(); int lw = l & (w - 1); int src = lw + (int)d < w ? l + (int)d : l; uint64_t x = 0; memcpy(&x, &v, sizeof(T)); x = emu::warp_exchange(x, src); T r; memcpy(&r, &x, sizeof(T)); return r; }
template <typename T> static inline T __shfl_xor_sync(unsigned, T v, int m, int w = 32)
{ int l = emu::lane_id(); int src = l ^ m; if ((src & ~(w - 1)) != (l & ~(w - 1))) src = l; uint64_t x = 0; memcpy(&x, &v, sizeof(T)); x = emu::warp_exchange(x, src); T r; memcpy(&r, &x, sizeof(T)); return r; }
static inline unsigned __ballot_sync(unsigned, int p) { return emu::warp_ballot(p); }
static inline int __any_sync(unsigned, int p) { return emu::warp_ballot(p) != 0; }
static inline int __all_sync(unsigned, int p) { return emu::warp_ballot(!p) == 0; }
static inline unsigned __activemask() { return 0xffffffffu; }

static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline unsigned __brev(unsigned x)
{ x = (x >> 16) | (x << 16); x = ((x & 0xff00ff00u) >> 8) | ((x & 0x00ff00ffu) << 8); x = ((x & 0xf0f0f0f0u) >> 4) | ((x & 0x0f0f0f0fu) << 4);
  x = ((x & 0xccccccccu) >> 2) | ((x & 0x33333333u) << 2); x = ((x & 0xaaaaaaaau) >> 1) | ((x & 0x55555555u) << 1); return x; }
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned s) { s &= 31; return s ? (lo >> s) | (hi << (32 - s)) : lo; }
static inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned s) { s &= 31; return s ? (hi << s) | (lo >> (32 - s)) : hi; }
template <typename T> static inline T __ldg(const T *p) { return *p; }
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned sel)
{ unsigned long long src = ((unsigned long long)b << 32) | a; unsigned r = 0; for (int k = 0; k < 4; k++) { unsigned n = (sel >> (4 * k)) & 7u; r |= (unsigned)((src >> (8 * n)) & 0xffu) << (8 * k); } return r; }

template <typename T> static inline T atomicAdd(T *p, T v) { T o = *p; *p = (T)(o + v); return o; }
template <typename T> static inline T atomicMax(T *p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <typename T> static inline T atomicMin(T *p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <typename T> static inline T atomicOr(T *p, T v) { T o = *p; *p = (T)(o | v); return o; }
template <typename T> static inline T atomicExch(T *p, T v) { T o = *p; *p = v; return o; }

using std::max;
using std::min;

/* kernel launch: EMU_LAUNCH(kernel, grid, block, smem_bytes, args...) */
#define EMU_DYN_SMEM(type, name) type *name = (type *)emu::g_blk->dyn_smem
