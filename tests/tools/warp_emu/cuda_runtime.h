// Stand-in for <cuda_runtime.h> when a kernel header is compiled for the HOST by the harnesses of this directory: one
// thread block runs as fibers on one OS thread (warp_emu.h), switching only at synchronisation points, so the device
// code executes unchanged -- shuffles, __syncwarp, __syncthreads, shared memory, atomics, the fp64 intrinsics.
// Compile with -ffp-contract=off (the library is built with --fmad=false).  Test infrastructure only; nothing in
// nanopore_b200/ includes it.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <algorithm>

#define __device__
#define __global__
#define __host__
#define __forceinline__ inline
#define __constant__ static
#define __grid_constant__
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))
#define PHMM_DYN_SHARED(name) unsigned char *const name = warp_emu::dyn_smem

namespace warp_emu {
int tid();
int lane();
int nthreads();
void barrier_block();
void barrier_warp();
extern unsigned long long slot[];
extern unsigned char *dyn_smem;
}

struct EmuIdx { int x; };
#define threadIdx (EmuIdx{warp_emu::tid()})
#define blockIdx (EmuIdx{0})
#define blockDim (EmuIdx{warp_emu::nthreads()})
#define gridDim (EmuIdx{1})

using std::max;
using std::min;

inline void __syncwarp(unsigned = 0xffffffffu) { warp_emu::barrier_warp(); }
inline void __syncthreads() { warp_emu::barrier_block(); }

template <typename T>
inline T emu_exchange(T v, int src_lane) {
    static_assert(sizeof(T) <= 8, "shuffle payload");
    unsigned long long raw = 0;
    memcpy(&raw, &v, sizeof(T));
    warp_emu::slot[warp_emu::tid()] = raw;
    warp_emu::barrier_warp();
    raw = warp_emu::slot[(warp_emu::tid() & ~31) + src_lane];
    warp_emu::barrier_warp();
    T r;
    memcpy(&r, &raw, sizeof(T));
    return r;
}
template <typename T> inline T __shfl_sync(unsigned, T v, int src) { return emu_exchange(v, src & 31); }
template <typename T> inline T __shfl_up_sync(unsigned, T v, int o) { const int l = warp_emu::lane(); return emu_exchange(v, l - o >= 0 ? l - o : l); }
template <typename T> inline T __shfl_down_sync(unsigned, T v, int o) { const int l = warp_emu::lane(); return emu_exchange(v, l + o < 32 ? l + o : l); }
inline int __reduce_max_sync(unsigned, int v) {
    for (int o = 16; o >= 1; o >>= 1) v = std::max(v, emu_exchange(v, warp_emu::lane() ^ o));
    return v;
}

// fibers switch only at synchronisation points, so plain read-modify-write is atomic here
template <typename T> inline T atomicAdd(T *p, T v) { const T o = *p; *p = o + v; return o; }
inline int atomicMin(int *p, int v) { const int o = *p; *p = std::min(o, v); return o; }
inline int atomicMax(int *p, int v) { const int o = *p; *p = std::max(o, v); return o; }

inline long long __double2ll_rz(double v) { return (long long)v; }
inline long long __double2ll_rd(double v) { return (long long)floor(v); }
inline double __longlong_as_double(long long v) { double d; memcpy(&d, &v, 8); return d; }
inline int __double2hiint(double d) { long long v; memcpy(&v, &d, 8); return (int)(v >> 32); }
inline int __double2loint(double d) { long long v; memcpy(&v, &d, 8); return (int)v; }
inline double __hiloint2double(int hi, int lo) { const long long v = (long long)(((unsigned long long)(unsigned)hi << 32) | (unsigned)lo); double d; memcpy(&d, &v, 8); return d; }
struct double2 { double x, y; };
inline double2 make_double2(double x, double y) { return double2{x, y}; }
struct int4 { int x, y, z, w; };
