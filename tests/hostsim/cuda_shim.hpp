// cuda_shim.hpp — a small SIMT emulator: runs the repository's __global__ functions on the CPU.
//
// TEST INFRASTRUCTURE ONLY.  Nothing in the product (libg2p.so, gaf2paf, gaf2unstable) includes
// this file.  It exists because the build container has no GPU: the kernels are warp-cooperative
// (shuffles, ballots, shared memory), so a scalar "host instantiation" cannot exercise them.  Here
// every CUDA thread of one CTA is a fiber; warp collectives and __syncthreads are rendezvous points
// resolved by a scheduler, CTAs run one after the other.  Only the subset of CUDA the kernels use
// is provided.
#pragma once
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#define G2P_HOSTSIM 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define G2P_NOINLINE __attribute__((noinline))
#define __restrict__
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __shared__ static   /* one CTA runs at a time on one OS thread */

struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct __attribute__((aligned(16))) uint4 { unsigned x, y, z, w; };
struct __attribute__((aligned(8))) uint2 { unsigned x, y; };
struct __attribute__((aligned(16))) ulonglong2 { unsigned long long x, y; };
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }

namespace hs {

enum Kind { K_NONE = 0, K_SHFL_IDX, K_SHFL_UP, K_SHFL_DOWN, K_SHFL_XOR, K_BALLOT, K_ANY, K_ALL, K_MATCH_ANY, K_SYNCWARP, K_SYNCTHREADS };

struct Thread {
    void* sp = nullptr;
    char* stack = nullptr;
    uint3 tid{0, 0, 0};
    bool done = false, waiting = false;
    int kind = K_NONE;
    unsigned mask = 0;
    uint64_t val = 0, result = 0;
    int arg = 0, width = 32;
};

extern Thread* cur;
extern uint3 g_block;
extern dim3 g_bdim, g_gdim;
extern unsigned char* g_dyn_smem;
uint64_t collective(int kind, unsigned mask, uint64_t val, int arg, int width);
void launch(dim3 grid, dim3 block, size_t dyn_smem, const std::function<void()>& body);

}  // namespace hs

#define threadIdx (hs::cur->tid)
#define blockIdx (hs::g_block)
#define blockDim (hs::g_bdim)
#define gridDim (hs::g_gdim)

// ---- collectives
template <class T> static inline uint64_t hs_pack(T v) { uint64_t u = 0; static_assert(sizeof(T) <= 8, ""); std::memcpy(&u, &v, sizeof(T)); return u; }
template <class T> static inline T hs_unpack(uint64_t u) { T v; std::memcpy(&v, &u, sizeof(T)); return v; }
template <class T> static inline T __shfl_sync(unsigned m, T v, int src, int w = 32) { return hs_unpack<T>(hs::collective(hs::K_SHFL_IDX, m, hs_pack(v), src, w)); }
template <class T> static inline T __shfl_up_sync(unsigned m, T v, unsigned d, int w = 32) { return hs_unpack<T>(hs::collective(hs::K_SHFL_UP, m, hs_pack(v), (int)d, w)); }
template <class T> static inline T __shfl_down_sync(unsigned m, T v, unsigned d, int w = 32) { return hs_unpack<T>(hs::collective(hs::K_SHFL_DOWN, m, hs_pack(v), (int)d, w)); }
template <class T> static inline T __shfl_xor_sync(unsigned m, T v, int x, int w = 32) { return hs_unpack<T>(hs::collective(hs::K_SHFL_XOR, m, hs_pack(v), x, w)); }
static inline unsigned __ballot_sync(unsigned m, int p) { return (unsigned)hs::collective(hs::K_BALLOT, m, p != 0, 0, 32); }
static inline int __any_sync(unsigned m, int p) { return (int)hs::collective(hs::K_ANY, m, p != 0, 0, 32); }
static inline int __all_sync(unsigned m, int p) { return (int)hs::collective(hs::K_ALL, m, p != 0, 0, 32); }
template <class T> static inline unsigned __match_any_sync(unsigned m, T v) { return (unsigned)hs::collective(hs::K_MATCH_ANY, m, hs_pack(v), 0, 32); }
static inline void __syncwarp(unsigned m = 0xffffffffu) { hs::collective(hs::K_SYNCWARP, m, 0, 0, 32); }
static inline void __syncthreads() { hs::collective(hs::K_SYNCTHREADS, 0, 0, 0, 32); }
static inline void __threadfence() {}
static inline void __threadfence_block() {}
static inline void __nanosleep(unsigned) {}
static inline void __trap() { std::fprintf(stderr, "hostsim: __trap()\n"); std::abort(); }

// ---- integer intrinsics
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __ffsll(long long x) { return __builtin_ffsll(x); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline int __clzll(long long x) { return x ? __builtin_clzll((unsigned long long)x) : 64; }
static inline unsigned __brev(unsigned x) { unsigned r = 0; for (int i = 0; i < 32; ++i) r |= ((x >> i) & 1u) << (31 - i); return r; }
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((uint64_t)a * b) >> 32); }
static inline unsigned long long __umul64hi(unsigned long long a, unsigned long long b) { return (unsigned long long)(((unsigned __int128)a * b) >> 64); }
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned s) { s &= 31; return s ? (lo >> s) | (hi << (32 - s)) : lo; }
static inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned s) { s &= 31; return s ? (hi << s) | (lo >> (32 - s)) : hi; }
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned sel) {
    uint64_t ab = ((uint64_t)b << 32) | a;
    unsigned r = 0;
    for (int i = 0; i < 4; ++i) {
        unsigned s = (sel >> (4 * i)) & 0xf;
        unsigned byte = (unsigned)((ab >> (8 * (s & 7))) & 0xff);
        if (s & 8) byte = (byte & 0x80) ? 0xff : 0;
        r |= byte << (8 * i);
    }
    return r;
}
template <class T> static inline T __ldg(const T* p) { return *p; }
template <class T> static inline T __ldcs(const T* p) { return *p; }
template <class T> static inline void __stcs(T* p, T v) { *p = v; }
static inline double __ddiv_rn(double a, double b) { volatile double r = a / b; return r; }
static inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
static inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }

static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned long long min(unsigned long long a, unsigned long long b) { return a < b ? a : b; }
static inline unsigned long long max(unsigned long long a, unsigned long long b) { return a > b ? a : b; }
static inline unsigned long min(unsigned long a, unsigned long b) { return a < b ? a : b; }
static inline unsigned long max(unsigned long a, unsigned long b) { return a > b ? a : b; }
static inline long long min(long long a, long long b) { return a < b ? a : b; }
static inline long long max(long long a, long long b) { return a > b ? a : b; }

// ---- atomics (sequential: one fiber runs at a time)
template <class T> static inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
template <class T> static inline T atomicMin(T* p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <class T> static inline T atomicMax(T* p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <class T> static inline T atomicOr(T* p, T v) { T o = *p; *p = o | v; return o; }
template <class T> static inline T atomicExch(T* p, T v) { T o = *p; *p = v; return o; }
template <class T> static inline T atomicCAS(T* p, T c, T v) { T o = *p; if (o == c) *p = v; return o; }

// =====================================================================================
#ifdef HS_IMPLEMENTATION
namespace hs {

Thread* cur = nullptr;
uint3 g_block{0, 0, 0};
dim3 g_bdim, g_gdim;
unsigned char* g_dyn_smem = nullptr;
static void* g_sched_sp = nullptr;
static const std::function<void()>* g_body = nullptr;

extern "C" void hs_switch(void** save_sp, void* new_sp);
asm(R"(
.text
.globl hs_switch
.type hs_switch,@function
hs_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size hs_switch,.-hs_switch
)");

static void trampoline() {
    (*g_body)();
    cur->done = true;
    hs_switch(&cur->sp, g_sched_sp);
    std::abort();
}

uint64_t collective(int kind, unsigned mask, uint64_t val, int arg, int width) {
    Thread* t = cur;
    t->kind = kind; t->mask = mask; t->val = val; t->arg = arg; t->width = width; t->waiting = true;
    hs_switch(&t->sp, g_sched_sp);
    return t->result;
}

static const size_t kStack = 128 * 1024;

static void prepare(Thread& t) {
    uintptr_t top = ((uintptr_t)t.stack + kStack) & ~(uintptr_t)15;
    uint64_t* sp = (uint64_t*)(top - 8);
    *--sp = (uint64_t)(uintptr_t)&trampoline;
    for (int i = 0; i < 6; ++i) *--sp = 0;
    t.sp = sp;
    t.done = false; t.waiting = false; t.kind = K_NONE;
}

// Resolve collectives of one warp; returns true if any lane was released.
static bool resolve_warp(Thread* w, int nl) {
    bool released = false;
    for (int l = 0; l < nl; ++l) {
        Thread& a = w[l];
        if (!a.waiting || a.kind == K_SYNCTHREADS) continue;
        const unsigned m = a.mask;
        if (!((m >> l) & 1u)) { std::fprintf(stderr, "hostsim: lane %d not in its own mask %08x\n", l, m); std::abort(); }
        bool ready = true;
        for (int k = 0; k < 32 && ready; ++k) {
            if (!((m >> k) & 1u)) continue;
            if (k >= nl) { ready = false; std::fprintf(stderr, "hostsim: mask %08x names lane %d beyond the CTA\n", m, k); std::abort(); }
            Thread& b = w[k];
            if (b.done) { std::fprintf(stderr, "hostsim: lane %d exited but is named in mask %08x (kind %d)\n", k, m, a.kind); std::abort(); }
            if (!b.waiting || b.kind == K_SYNCTHREADS) { ready = false; break; }
            if (b.mask != m) { ready = false; break; }   // waits on a different collective: not yet
            if (b.kind != a.kind) { std::fprintf(stderr, "hostsim: lanes %d/%d meet in different collectives (%d vs %d)\n", l, k, a.kind, b.kind); std::abort(); }
        }
        if (!ready) continue;
        // compute results
        unsigned ballot = 0;
        for (int k = 0; k < 32; ++k) if (((m >> k) & 1u) && w[k].val) ballot |= 1u << k;
        uint64_t res[32];
        for (int k = 0; k < 32; ++k) {
            if (!((m >> k) & 1u)) continue;
            Thread& b = w[k];
            const int wd = b.width, seg = k & ~(wd - 1), rel = k & (wd - 1);
            int src = k;
            switch (b.kind) {
                case K_SHFL_IDX: src = seg | (b.arg & (wd - 1)); break;
                case K_SHFL_UP: src = rel >= b.arg ? k - b.arg : k; break;
                case K_SHFL_DOWN: src = rel + b.arg < wd ? k + b.arg : k; break;
                case K_SHFL_XOR: src = ((rel ^ b.arg) < wd) ? (seg | (rel ^ b.arg)) : k; break;
                default: break;
            }
            switch (b.kind) {
                case K_SHFL_IDX: case K_SHFL_UP: case K_SHFL_DOWN: case K_SHFL_XOR:
                    // reading a lane outside the mask is undefined in CUDA: poison it
                    res[k] = ((m >> src) & 1u) ? w[src].val : 0xDEADBEEFDEADBEEFULL;
                    break;
                case K_BALLOT: res[k] = ballot; break;
                case K_ANY: res[k] = ballot != 0; break;
                case K_ALL: res[k] = ballot == m; break;
                case K_MATCH_ANY: {
                    unsigned mm = 0;
                    for (int j = 0; j < 32; ++j) if (((m >> j) & 1u) && w[j].val == b.val) mm |= 1u << j;
                    res[k] = mm;
                    break;
                }
                default: res[k] = 0;
            }
        }
        for (int k = 0; k < 32; ++k) {
            if (!((m >> k) & 1u)) continue;
            w[k].result = res[k]; w[k].waiting = false; w[k].kind = K_NONE;
        }
        released = true;
    }
    return released;
}

void launch(dim3 grid, dim3 block, size_t dyn_smem, const std::function<void()>& body) {
    const int nt = (int)(block.x * block.y * block.z);
    std::vector<Thread> th(nt);
    char* stacks = static_cast<char*>(std::malloc((size_t)nt * kStack + 64));   // untouched pages cost nothing
    std::vector<unsigned char> smem(dyn_smem + 64);
    for (int i = 0; i < nt; ++i) th[i].stack = stacks + (size_t)i * kStack;
    g_bdim = block; g_gdim = grid; g_body = &body;
    g_dyn_smem = (unsigned char*)(((uintptr_t)smem.data() + 15) & ~(uintptr_t)15);
    // G2P_SIMT_REVERSE=1 runs the CTAs of every launch in reverse order: kernels must not depend on
    // the order their CTAs happen to execute in (tickets / look-back excepted, which do not use blockIdx)
    static const bool reverse = std::getenv("G2P_SIMT_REVERSE") != nullptr;
    for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
    for (unsigned bi = 0; bi < grid.x; ++bi) {
        const unsigned bx = reverse ? grid.x - 1 - bi : bi;
        g_block = uint3{bx, by, bz};
        std::memset(g_dyn_smem, 0xCD, dyn_smem);
        for (int i = 0; i < nt; ++i) {
            th[i].tid = uint3{(unsigned)i % block.x, ((unsigned)i / block.x) % block.y, (unsigned)i / (block.x * block.y)};
            prepare(th[i]);
        }
        for (;;) {
            bool progress = false;
            int alive = 0;
            for (int i = 0; i < nt; ++i) {
                Thread& t = th[i];
                if (t.done) continue;
                ++alive;
                if (t.waiting) continue;
                cur = &t;
                hs_switch(&g_sched_sp, t.sp);
                cur = nullptr;
                progress = true;
            }
            if (!alive) break;
            for (int w0 = 0; w0 < nt; w0 += 32)
                while (resolve_warp(&th[w0], nt - w0 < 32 ? nt - w0 : 32)) progress = true;
            // __syncthreads: all live threads must have arrived
            int at_bar = 0, live = 0;
            for (int i = 0; i < nt; ++i) { if (th[i].done) continue; ++live; if (th[i].waiting && th[i].kind == K_SYNCTHREADS) ++at_bar; }
            if (live && at_bar == live) {
                for (int i = 0; i < nt; ++i) if (!th[i].done) { th[i].waiting = false; th[i].kind = K_NONE; th[i].result = 0; }
                progress = true;
            }
            if (!progress) {
                std::fprintf(stderr, "hostsim: deadlock in block (%u,%u,%u)\n", bx, by, bz);
                for (int i = 0; i < nt; ++i) if (!th[i].done) std::fprintf(stderr, "  thread %d kind %d mask %08x\n", i, th[i].kind, th[i].mask);
                std::abort();
            }
        }
    }
    g_body = nullptr;
    std::free(stacks);
}

}  // namespace hs
#endif  // HS_IMPLEMENTATION
