// Stand-in for <cuda_runtime.h> when a kernel header is compiled for the HOST by tests/tools/warp_emu/decode_w_emu.cpp:
// one warp of 32 lanes runs as 32 fibers (ucontext) on one OS thread, switching only at warp-level synchronisation
// points, so the device code executes unchanged -- shuffles, __syncwarp, shared memory, integer atomics.  Test
// infrastructure only; nothing in nanopore_b200/ includes it.
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

namespace warp_emu {
int lane();
void barrier();
extern unsigned long long slot[32];
}

struct EmuIdx { int x; };
#define threadIdx (EmuIdx{warp_emu::lane()})
#define blockIdx (EmuIdx{0})

using std::max;
using std::min;

inline void __syncwarp(unsigned = 0xffffffffu) { warp_emu::barrier(); }

template <typename T>
inline T emu_exchange(T v, int src) {
    static_assert(sizeof(T) <= 8, "shuffle payload");
    unsigned long long raw = 0;
    memcpy(&raw, &v, sizeof(T));
    warp_emu::slot[warp_emu::lane()] = raw;
    warp_emu::barrier();
    raw = warp_emu::slot[src];
    warp_emu::barrier();
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

// fibers switch only at barriers, so plain read-modify-write is atomic here
inline int atomicAdd(int *p, int v) { const int o = *p; *p = o + v; return o; }
inline int atomicMin(int *p, int v) { const int o = *p; *p = std::min(o, v); return o; }
inline int atomicMax(int *p, int v) { const int o = *p; *p = std::max(o, v); return o; }

inline long long __double2ll_rz(double v) { return (long long)v; }
inline double __longlong_as_double(long long v) { double d; memcpy(&d, &v, 8); return d; }
inline int __double2hiint(double d) { long long v; memcpy(&v, &d, 8); return (int)(v >> 32); }
inline int __double2loint(double d) { long long v; memcpy(&v, &d, 8); return (int)v; }
inline double __hiloint2double(int hi, int lo) { const long long v = ((long long)hi << 32) | (unsigned)lo; double d; memcpy(&d, &v, 8); return d; }
