// warp_emu.h -- runs ONE thread block of a CUDA kernel on the host: every thread is a fiber (ucontext) on one OS
// thread; fibers switch only at synchronisation points (__syncthreads, __syncwarp, shuffles), in an order that is
// re-shuffled every scheduling round when a seed is given, so that a missing synchronisation shows as a wrong
// result for some seed.  Include exactly once per harness, after cuda_runtime.h of this directory.  Test
// infrastructure only; nothing in nanopore_b200/ includes it.
#pragma once
#include <ucontext.h>
#include <functional>
#include <vector>

namespace warp_emu {
constexpr int MAX_THREADS = 256;
unsigned long long slot[MAX_THREADS];
unsigned char *dyn_smem = nullptr;
static ucontext_t main_ctx, ctx[MAX_THREADS];
static int cur = 0, n_threads = 32, done[MAX_THREADS];
static long arrived_block[MAX_THREADS], arrived_warp[MAX_THREADS];
static std::function<void()> body;

int tid() { return cur; }
int lane() { return cur & 31; }
int nthreads() { return n_threads; }
static void yield() { swapcontext(&ctx[cur], &main_ctx); }
static void wait_for(long *arr, int lo, int hi) {
    const long g = ++arr[cur];
    for (;;) {
        bool all = true;
        for (int t = lo; t < hi; t++) if (!done[t] && arr[t] < g) { all = false; break; }
        if (all) return;
        yield();
    }
}
void barrier_block() { wait_for(arrived_block, 0, n_threads); }
void barrier_warp() { const int lo = cur & ~31; wait_for(arrived_warp, lo, lo + 32 < n_threads ? lo + 32 : n_threads); }
static void entry() {
    body();
    done[cur] = 1;
    yield();
}
// one block of `threads` threads running `kernel_call` to completion
static void run_block(int threads, unsigned seed, std::function<void()> kernel_call) {
    n_threads = threads;
    body = kernel_call;
    std::vector<std::vector<char>> stacks(threads, std::vector<char>(1 << 18));
    for (int t = 0; t < threads; t++) {
        done[t] = 0; arrived_block[t] = 0; arrived_warp[t] = 0;
        getcontext(&ctx[t]);
        ctx[t].uc_stack.ss_sp = stacks[t].data();
        ctx[t].uc_stack.ss_size = stacks[t].size();
        ctx[t].uc_link = &main_ctx;
        makecontext(&ctx[t], entry, 0);
    }
    std::vector<int> perm(threads);
    for (int t = 0; t < threads; t++) perm[t] = t;
    unsigned s = seed;
    for (;;) {
        bool any = false;
        if (seed) for (int t = threads - 1; t > 0; t--) { s = s * 1664525u + 1013904223u; std::swap(perm[t], perm[(s >> 8) % (t + 1)]); }
        for (int q = 0; q < threads; q++) {
            const int t = perm[q];
            if (done[t]) continue;
            any = true;
            cur = t;
            swapcontext(&main_ctx, &ctx[t]);
        }
        if (!any) break;
    }
}
}  // namespace warp_emu
