// TEST INFRASTRUCTURE ONLY.  A minimal single-OS-thread CUDA execution emulator (one ucontext fiber per CUDA thread,
// one block at a time) so that the kernels in openvo_b200/csrc/*.cu can be compiled with g++ (-DOVO_EMU) and their
// logic — indexing, warp collectives, barriers, packed-integer arithmetic — checked against the oracle in the CPU-only
// container.  It is never part of the product: openvo_b200/_native.py loads only the nvcc-built library and raises
// if that is missing.  Performance is irrelevant here; semantics of the subset of CUDA the kernels use are not.
#pragma once
#include <ucontext.h>
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

// ---- vector / runtime types ------------------------------------------------------------------------------
struct uint2 { unsigned x, y; };
struct uint3 { unsigned x, y, z; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct int2 { int x, y; };
struct alignas(16) int4 { int x, y, z, w; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
static inline uint2 make_uint2(unsigned a, unsigned b) { return {a, b}; }
static inline uint4 make_uint4(unsigned a, unsigned b, unsigned c, unsigned d) { return {a, b, c, d}; }
static inline float2 make_float2(float a, float b) { return {a, b}; }
static inline float4 make_float4(float a, float b, float c, float d) { return {a, b, c, d}; }
typedef void* cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
static inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }
static inline cudaError_t cudaGetLastError() { return 0; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return 0; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return 0; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { memset(d, v, n); return 0; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
static inline cudaError_t cudaDeviceSynchronize() { return 0; }
static inline cudaError_t cudaMallocHost(void** p, size_t n) { *p = malloc(n); return 0; }
static inline cudaError_t cudaFreeHost(void* p) { free(p); return 0; }
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = malloc(n); return 0; }
static inline cudaError_t cudaFree(void* p) { free(p); return 0; }
static inline cudaError_t cudaSetDevice(int) { return 0; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return 0; }
template <typename T> static inline cudaError_t cudaFuncSetAttribute(T, int, int) { return 0; }
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __shared__ static
#define __restrict__
#define __launch_bounds__(...)
#define __constant__ static
#define __align__(n) __attribute__((aligned(n)))
using std::isinf;
using std::isnan;
using std::isfinite;

// ---- execution engine ---------------------------------------------------------------------------------------
namespace ovo_emu {
struct Warp {
    uint64_t buf[2][32];
    int arrived = 0, alive = 0;
    unsigned gen = 0;
};
struct NamedBar {  // bar.arrive / bar.sync with an explicit thread count (producer / consumer hand-over)
    int arrived = 0;
    unsigned gen = 0;
    std::vector<char> seen;
};
struct Engine {
    NamedBar nbar[16];
    ucontext_t sched;
    std::vector<ucontext_t> ctx;
    std::vector<char> done;
    std::vector<Warp> warps;
    std::vector<char*> stacks;
    std::function<void()> body;
    int nthreads = 0, cur = 0, live = 0;
    int bar_arrived = 0;
    unsigned bar_gen = 0;
    std::vector<uint64_t> dyn_smem;
};
inline Engine& eng() { static Engine e; return e; }
inline uint3 g_threadIdx, g_blockIdx;
inline dim3 g_blockDim, g_gridDim;
inline void yield() { Engine& e = eng(); swapcontext(&e.ctx[e.cur], &e.sched); }
inline void trampoline() {
    Engine& e = eng();
    e.body();
    e.done[e.cur] = 1;
    e.live--;
    e.warps[e.cur >> 5].alive--;
    // a thread that exits while its warp / block waits must not dead-lock the others
    swapcontext(&e.ctx[e.cur], &e.sched);
}
inline void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& fn) {
    Engine& e = eng();
    const int nt = block.x * block.y * block.z;
    const size_t STK = 256 * 1024;
    while ((int)e.stacks.size() < nt) e.stacks.push_back((char*)malloc(STK));
    e.ctx.resize(nt);
    e.done.assign(nt, 0);
    e.dyn_smem.assign(smem / 8 + 2, 0);
    e.body = fn;
    g_blockDim = block;
    g_gridDim = grid;
    for (unsigned bz = 0; bz < grid.z; bz++)
        for (unsigned by = 0; by < grid.y; by++)
            for (unsigned bx = 0; bx < grid.x; bx++) {
                g_blockIdx = {bx, by, bz};
                e.nthreads = e.live = nt;
                e.bar_arrived = 0;
                for (auto& nb : e.nbar) { nb.arrived = 0; nb.gen = 0; nb.seen.assign(nt, 0); }
                e.warps.assign((nt + 31) / 32, Warp());
                // OVO_EMU_POISON=<byte>: fill the thread stacks and the dynamic shared memory of every CTA with that byte, so that a
                // read of an uninitialised local or shared-memory word gives the same (wrong) value every run instead of leftovers
                static const char* poison = getenv("OVO_EMU_POISON");
                if (poison) {
                    const int pb = (int)strtol(poison, nullptr, 0);
                    for (int t = 0; t < nt; t++) memset(e.stacks[t], pb, STK);
                    memset(e.dyn_smem.data(), pb, e.dyn_smem.size() * 8);
                }
                for (int t = 0; t < nt; t++) {
                    e.done[t] = 0;
                    e.warps[t >> 5].alive++;
                    getcontext(&e.ctx[t]);
                    e.ctx[t].uc_stack.ss_sp = e.stacks[t];
                    e.ctx[t].uc_stack.ss_size = STK;
                    e.ctx[t].uc_link = &e.sched;
                    makecontext(&e.ctx[t], (void (*)())trampoline, 0);
                }
                // OVO_EMU_SCHED: 0 / unset = round robin in thread order, 1 = reverse order, n >= 2 = pseudo-random order (seed n)
                // reshuffled every sweep: other interleavings of the warps for kernels that synchronise by named barriers / flags
                static const long sched_mode = getenv("OVO_EMU_SCHED") ? atol(getenv("OVO_EMU_SCHED")) : 0;
                std::vector<int> order(nt);
                for (int t = 0; t < nt; t++) order[t] = sched_mode == 1 ? nt - 1 - t : t;
                uint64_t rng = 0x9E3779B97F4A7C15ull * (uint64_t)(sched_mode + 1);
                while (e.live > 0)
                    for (int oi = 0; oi < nt; oi++) {
                        if (sched_mode >= 2 && oi == 0)
                            for (int i = nt - 1; i > 0; i--) {
                                rng = rng * 6364136223846793005ull + 1442695040888963407ull;
                                std::swap(order[i], order[(rng >> 33) % (uint64_t)(i + 1)]);
                            }
                        const int t = order[oi];
                        if (e.done[t]) continue;
                        e.cur = t;
                        g_threadIdx = {(unsigned)(t % block.x), (unsigned)((t / block.x) % block.y), (unsigned)(t / (block.x * block.y))};
                        swapcontext(&e.sched, &e.ctx[t]);
                    }
            }
}
inline void restore_tid() {  // after a yield other fibers changed the globals
    Engine& e = eng();
    const int t = e.cur;
    g_threadIdx = {(unsigned)(t % g_blockDim.x), (unsigned)((t / g_blockDim.x) % g_blockDim.y), (unsigned)(t / (g_blockDim.x * g_blockDim.y))};
}
inline void block_barrier() {
    Engine& e = eng();
    const unsigned g = e.bar_gen;
    if (++e.bar_arrived >= e.live) { e.bar_arrived = 0; e.bar_gen++; return; }
    while (e.bar_gen == g) {
        yield();
        if (e.bar_gen == g && e.bar_arrived >= e.live) { e.bar_arrived = 0; e.bar_gen++; }
    }
}
// named barrier: completes when `count` threads have arrived (arriving or waiting); a thread may not arrive twice in one phase
// (on hardware that corrupts the phase), and a wait that never completes is a dead-lock: both abort the test
inline void named_barrier(int id, int count, bool wait) {
    Engine& e = eng();
    NamedBar& b = e.nbar[id];
    if (b.seen[e.cur]) { fprintf(stderr, "emu: thread %d arrives twice at named barrier %d\n", e.cur, id); abort(); }
    b.seen[e.cur] = 1;
    const unsigned g = b.gen;
    if (++b.arrived >= count) {
        b.arrived = 0;
        b.gen++;
        std::fill(b.seen.begin(), b.seen.end(), 0);
        return;
    }
    if (!wait) return;
    for (long spins = 0; b.gen == g; spins++) {
        if (spins > 20000000) { fprintf(stderr, "emu: dead-lock at named barrier %d (thread %d)\n", id, e.cur); abort(); }
        yield();
    }
}
// wait (yielding) until a flag written by another thread has the expected value
inline void spin_until(const volatile int* flag, int value) {
    for (long spins = 0; *flag != value; spins++) {
        if (spins > 20000000) { fprintf(stderr, "emu: dead-lock waiting for a flag (thread %d)\n", eng().cur); abort(); }
        yield();
    }
}
// all live lanes of the warp deposit v; returns pointer to the 32 deposited values (valid until the next-but-one collective)
inline const uint64_t* warp_gather(uint64_t v) {
    Engine& e = eng();
    Warp& w = e.warps[e.cur >> 5];
    const unsigned g = w.gen;
    uint64_t* b = w.buf[g & 1];
    b[e.cur & 31] = v;
    if (++w.arrived >= w.alive) { w.arrived = 0; w.gen++; return b; }
    while (w.gen == g) {
        yield();
        if (w.gen == g && w.arrived >= w.alive) { w.arrived = 0; w.gen++; }
    }
    return b;
}
inline int lane_id() { return eng().cur & 31; }
inline int warp_lanes() {
    Engine& e = eng();
    const int w = e.cur >> 5;
    return std::min(32, e.nthreads - 32 * w);
}
}  // namespace ovo_emu

#define threadIdx ovo_emu::g_threadIdx
#define blockIdx ovo_emu::g_blockIdx
#define blockDim ovo_emu::g_blockDim
#define gridDim ovo_emu::g_gridDim
#define OVO_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(ovo_emu::eng().dyn_smem.data())
#define OVO_LAUNCH(kern, grid, block, smem, st, ...) ovo_emu::launch(grid, block, smem, [&] { kern(__VA_ARGS__); })

static inline void __syncthreads() { ovo_emu::block_barrier(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { ovo_emu::warp_gather(0); }
static inline void __threadfence() {}
static inline void __threadfence_block() {}

// ---- warp collectives (full-mask use only) -----------------------------------------------------------------
template <typename T> static inline uint64_t emu_bits(T v) { uint64_t b = 0; memcpy(&b, &v, sizeof(T)); return b; }
template <typename T> static inline T emu_from(uint64_t b) { T v; memcpy(&v, &b, sizeof(T)); return v; }
template <typename T> static inline T __shfl_sync(unsigned, T v, int src, int width = 32) {
    const int l = ovo_emu::lane_id();
    const uint64_t* b = ovo_emu::warp_gather(emu_bits(v));
    const int s = (l / width) * width + (src % width);
    return emu_from<T>(b[s]);
}
template <typename T> static inline T __shfl_up_sync(unsigned, T v, unsigned d, int width = 32) {
    const int l = ovo_emu::lane_id();
    const uint64_t* b = ovo_emu::warp_gather(emu_bits(v));
    const int s = l - (int)d;
    return (s < (l / width) * width) ? v : emu_from<T>(b[s]);
}
template <typename T> static inline T __shfl_down_sync(unsigned, T v, unsigned d, int width = 32) {
    const int l = ovo_emu::lane_id();
    const uint64_t* b = ovo_emu::warp_gather(emu_bits(v));
    const int s = l + (int)d;
    return (s >= (l / width + 1) * width || s >= ovo_emu::warp_lanes()) ? v : emu_from<T>(b[s]);
}
template <typename T> static inline T __shfl_xor_sync(unsigned, T v, int m, int width = 32) {
    const int l = ovo_emu::lane_id();
    const uint64_t* b = ovo_emu::warp_gather(emu_bits(v));
    const int s = l ^ m;
    return s >= ovo_emu::warp_lanes() ? v : emu_from<T>(b[s]);
}
static inline unsigned __ballot_sync(unsigned, int p) {
    const uint64_t* b = ovo_emu::warp_gather(p ? 1 : 0);
    unsigned r = 0;
    for (int i = 0; i < ovo_emu::warp_lanes(); i++) r |= (b[i] ? 1u : 0u) << i;
    return r;
}
static inline unsigned __match_any_sync(unsigned, int v) {
    const uint64_t* b = ovo_emu::warp_gather((uint64_t)(int64_t)v);
    unsigned r = 0;
    for (int i = 0; i < ovo_emu::warp_lanes(); i++) r |= ((int)(int64_t)b[i] == v ? 1u : 0u) << i;
    return r;
}
static inline int __any_sync(unsigned m, int p) { return __ballot_sync(m, p) != 0; }
static inline int __all_sync(unsigned m, int p) { const int n = ovo_emu::warp_lanes(); return __ballot_sync(m, p) == (n == 32 ? 0xffffffffu : ((1u << n) - 1)); }
static inline unsigned __reduce_min_sync(unsigned mask, unsigned v) {  // sub-warp masks: every lane names its own group
    const uint64_t* b = ovo_emu::warp_gather(v);
    unsigned r = 0xffffffffu;
    for (int i = 0; i < ovo_emu::warp_lanes(); i++)
        if (mask & (1u << i)) r = std::min(r, (unsigned)b[i]);
    return r;
}
static inline unsigned __reduce_max_sync(unsigned, unsigned v) {
    const uint64_t* b = ovo_emu::warp_gather(v);
    unsigned r = 0;
    for (int i = 0; i < ovo_emu::warp_lanes(); i++) r = std::max(r, (unsigned)b[i]);
    return r;
}
static inline int __reduce_add_sync(unsigned, int v) {
    const uint64_t* b = ovo_emu::warp_gather((uint64_t)(int64_t)v);
    int r = 0;
    for (int i = 0; i < ovo_emu::warp_lanes(); i++) r += (int)(int64_t)b[i];
    return r;
}
static inline unsigned __reduce_add_sync(unsigned m, unsigned v) { return (unsigned)__reduce_add_sync(m, (int)v); }

// ---- scalar intrinsics -------------------------------------------------------------------------------------------
template <typename T> static inline T __ldg(const T* p) { return *p; }
static inline unsigned __byte_perm(unsigned x, unsigned y, unsigned s) {
    const uint64_t v = ((uint64_t)y << 32) | x;
    unsigned r = 0;
    for (int i = 0; i < 4; i++) {
        const unsigned sel = (s >> (4 * i)) & 0xF;
        unsigned byte = (unsigned)((v >> (8 * (sel & 7))) & 0xFF);
        if (sel & 8) byte = (byte & 0x80) ? 0xFF : 0;
        r |= byte << (8 * i);
    }
    return r;
}
static inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned s) { s &= 31; return s ? (hi << s) | (lo >> (32 - s)) : hi; }
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned s) { s &= 31; return s ? (lo >> s) | (hi << (32 - s)) : lo; }
static inline unsigned emu_h2(unsigned a, unsigned b, unsigned (*op)(unsigned, unsigned)) {
    return (op(a & 0xFFFF, b & 0xFFFF) & 0xFFFF) | (op(a >> 16, b >> 16) << 16);
}
static inline unsigned __vminu2(unsigned a, unsigned b) { return emu_h2(a, b, [](unsigned x, unsigned y) { return std::min(x, y); }); }
static inline unsigned __vmaxu2(unsigned a, unsigned b) { return emu_h2(a, b, [](unsigned x, unsigned y) { return std::max(x, y); }); }
static inline unsigned __vmins2(unsigned a, unsigned b) {
    const short al = (short)(a & 0xFFFF), ah = (short)(a >> 16), bl = (short)(b & 0xFFFF), bh = (short)(b >> 16);
    return (unsigned)(unsigned short)std::min(al, bl) | ((unsigned)(unsigned short)std::min(ah, bh) << 16);
}
static inline unsigned __vmaxs2(unsigned a, unsigned b) {
    const short al = (short)(a & 0xFFFF), ah = (short)(a >> 16), bl = (short)(b & 0xFFFF), bh = (short)(b >> 16);
    return (unsigned)(unsigned short)std::max(al, bl) | ((unsigned)(unsigned short)std::max(ah, bh) << 16);
}
static inline unsigned __vimin3_u16x2(unsigned a, unsigned b, unsigned c) { return __vminu2(__vminu2(a, b), c); }
static inline unsigned __vimax3_u16x2(unsigned a, unsigned b, unsigned c) { return __vmaxu2(__vmaxu2(a, b), c); }
static inline unsigned __viaddmin_u16x2(unsigned a, unsigned b, unsigned c) {
    const unsigned lo = std::min(((a & 0xFFFF) + (b & 0xFFFF)) & 0xFFFF, c & 0xFFFF);
    const unsigned hi = std::min(((a >> 16) + (b >> 16)) & 0xFFFF, c >> 16);
    return lo | (hi << 16);
}
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
static inline float __fdiv_rn(float a, float b) { volatile float r = a / b; return r; }
static inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
static inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
static inline double __dsub_rn(double a, double b) { volatile double r = a - b; return r; }
static inline double __ddiv_rn(double a, double b) { volatile double r = a / b; return r; }
static inline int __float2int_rn(float v) { return (int)lrintf(v); }
static inline float __int2float_rn(int v) { return (float)v; }
static inline float __double2float_rn(double v) { return (float)v; }
static inline float __uint2float_rn(unsigned v) { return (float)v; }
template <typename T> static inline T atomicMin(T* p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <typename T> static inline T atomicMax(T* p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <typename T> static inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
template <typename T> static inline T atomicOr(T* p, T v) { T o = *p; *p = o | v; return o; }
template <typename T> static inline T atomicCAS(T* p, T c, T v) { T o = *p; if (o == c) *p = v; return o; }
template <typename T> static inline T atomicExch(T* p, T v) { T o = *p; *p = v; return o; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
static inline float fminf_(float a, float b) { return a < b ? a : b; }
