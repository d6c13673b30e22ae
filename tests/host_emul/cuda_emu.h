// cuda_emu.h -- a small CPU emulation of the CUDA execution model, TEST INFRASTRUCTURE ONLY.
//
// Lets the kernels of archive_b200/csrc/bzip2_enc_kernels.cu be compiled with g++ (-DB200Z_EMU) and executed on the
// build container, which has no GPU, so that their logic is covered by the `-m "not gpu"` tier. Every CUDA thread is a
// fibre (hand-rolled x86-64 context switch); the threads of one CTA run round-robin and only switch at
// __syncthreads() and at warp collectives, CTAs run one after another in blockIdx order. `__shared__` becomes `static`
// (valid because only one CTA is alive at a time). Nothing here is used by the product library.
//
// Several translation units of one library may include it: each gets its own scheduler state (launches complete before they
// return, so nothing is shared) and the context switch is a weak symbol.
#pragma once
#if !defined(__x86_64__)
#error "cuda_emu.h needs x86-64"
#endif
#include <mutex>
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <tuple>
#include <vector>

struct uint3 {
  unsigned x, y, z;
};
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct uint2 {
  unsigned x, y;
};
struct uint4 {
  unsigned x, y, z, w;
};
static inline uint4 make_uint4(unsigned a, unsigned b, unsigned c, unsigned d) { return uint4{a, b, c, d}; }
static inline uint2 make_uint2(unsigned a, unsigned b) { return uint2{a, b}; }

typedef int cudaError_t;
typedef void *cudaStream_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorLaunchFailure = 719 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyHostToHost };
static inline cudaError_t cudaMemsetAsync(void *p, int v, size_t n, cudaStream_t = nullptr) {
  memset(p, v, n);
  return cudaSuccess;
}
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) {
  memmove(d, s, n);
  return cudaSuccess;
}
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline const char *cudaGetErrorString(cudaError_t) { return "emulated"; }

// ---- the rest of the runtime API the product's host layer uses: "device" memory is host memory, streams and events are
// tokens (every launch and copy completes before it returns) ----
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
template <typename K>
static inline cudaError_t cudaFuncSetAttribute(K, cudaFuncAttribute, int) {
  return cudaSuccess;
}
enum { cudaHostAllocDefault = 0, cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2 };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16 };
typedef void *cudaEvent_t;
static inline cudaError_t cudaMalloc(void **p, size_t n) {
  *p = malloc(n ? n : 1);
  if (!*p) return 2;
  memset(*p, 0xCD, n);  // device memory is not zeroed: poison it so that reliance on zeros shows
  return cudaSuccess;
}
template <typename T>
static inline cudaError_t cudaMalloc(T **p, size_t n) {
  return cudaMalloc((void **)p, n);
}
static inline cudaError_t cudaFree(void *p) {
  free(p);
  return cudaSuccess;
}
static inline cudaError_t cudaHostAlloc(void **p, size_t n, unsigned) {
  *p = malloc(n ? n : 1);
  return *p ? cudaSuccess : 2;
}
template <typename T>
static inline cudaError_t cudaHostAlloc(T **p, size_t n, unsigned f) {
  return cudaHostAlloc((void **)p, n, f);
}
static inline cudaError_t cudaFreeHost(void *p) {
  free(p);
  return cudaSuccess;
}
static inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) {
  memmove(d, s, n);
  return cudaSuccess;
}
static inline cudaError_t cudaMemGetInfo(size_t *free_b, size_t *total_b) {
  *free_b = (size_t)6 << 30;
  *total_b = (size_t)8 << 30;
  return cudaSuccess;
}
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int *d) {
  *d = 0;
  return cudaSuccess;
}
static inline cudaError_t cudaGetDeviceCount(int *n) {
  *n = 1;
  return cudaSuccess;
}
static inline cudaError_t cudaDeviceGetAttribute(int *v, cudaDeviceAttr, int) {
  *v = 148;
  return cudaSuccess;
}
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) {
  *s = malloc(1);
  return cudaSuccess;
}
static inline cudaError_t cudaStreamDestroy(cudaStream_t s) {
  free(s);
  return cudaSuccess;
}
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t *e) {
  *e = malloc(1);
  return cudaSuccess;
}
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { return cudaEventCreate(e); }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) {
  free(e);
  return cudaSuccess;
}
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = nullptr) { return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t) {
  *ms = 0.f;
  return cudaSuccess;
}

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __shared__ static
#define __launch_bounds__(...)
#define __constant__ static

extern "C" void cuemu_switch(void **save_sp, void *new_sp);
asm(R"(
.text
.weak cuemu_switch
.type cuemu_switch,@function
cuemu_switch:
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
.size cuemu_switch, .-cuemu_switch
)");

namespace cuemu {
struct Fiber {
  void *sp = nullptr;
  char *stack = nullptr;
  bool done = false, at_barrier = false;
  uint3 tid{0, 0, 0};
  unsigned lane = 0, warp = 0;
};
struct WarpState {
  unsigned arrived = 0, released = 0, exists = 0;
  unsigned long long val[32], res[32], aux[32];
};
static Fiber *cur = nullptr;
static uint3 g_bid{0, 0, 0};
static dim3 g_bdim, g_gdim;
static void *sched_sp = nullptr;
static std::vector<WarpState> warps;
static const std::function<void()> *g_body = nullptr;
static unsigned long long events = 0;
static size_t stack_bytes = 256 * 1024;

static inline void yield() { cuemu_switch(&cur->sp, sched_sp); }
static void fiber_entry() {
  (*g_body)();
  cur->done = true;
  events++;
  yield();
  abort();
}
static inline void init_fiber(Fiber &f) {
  if (!f.stack) f.stack = (char *)malloc(stack_bytes);
  uintptr_t top = ((uintptr_t)f.stack + stack_bytes) & ~(uintptr_t)15;
  void **sp = (void **)(top - 64);
  for (int i = 0; i < 6; ++i) sp[i] = nullptr;
  sp[6] = (void *)&fiber_entry;
  f.sp = sp;
  f.done = f.at_barrier = false;
}

static std::vector<Fiber> pool;
// One launch at a time in the whole process (the scheduler state above is per translation unit and not re-entrant): host
// code that drives the device from several threads (b200z_deflate_batch's lanes) runs here with its kernels serialised.
inline std::mutex &launch_mutex() {
  static std::mutex m;
  return m;
}
static inline void launch(dim3 grid, dim3 block, const std::function<void()> &body) {
  std::lock_guard<std::mutex> one_launch(launch_mutex());
  size_t nthr = (size_t)block.x * block.y * block.z;
  if (pool.size() < nthr) pool.resize(nthr);
  g_bdim = block;
  g_gdim = grid;
  g_body = &body;
  size_t nwarps = (nthr + 31) / 32;
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        g_bid = uint3{bx, by, bz};
        warps.assign(nwarps, WarpState());
        for (size_t i = 0; i < nthr; ++i) {
          Fiber &f = pool[i];
          init_fiber(f);
          f.tid.x = (unsigned)(i % block.x);
          f.tid.y = (unsigned)((i / block.x) % block.y);
          f.tid.z = (unsigned)(i / ((size_t)block.x * block.y));
          f.lane = (unsigned)(i & 31);
          f.warp = (unsigned)(i >> 5);
          warps[f.warp].exists |= 1u << f.lane;
        }
        size_t remaining = nthr;
        while (remaining) {
          unsigned long long before = events;
          size_t waiting = 0;
          for (size_t i = 0; i < nthr; ++i) {
            Fiber &f = pool[i];
            if (f.done) continue;
            if (f.at_barrier) {
              waiting++;
              continue;
            }
            cur = &f;
            cuemu_switch(&sched_sp, f.sp);
            if (f.done) remaining--;
            else if (f.at_barrier) waiting++;
          }
          if (remaining && waiting == remaining) {
            for (size_t i = 0; i < nthr; ++i) pool[i].at_barrier = false;
            events++;
          }
          if (remaining && events == before) {
            fprintf(stderr, "cuda_emu: deadlock in block (%u,%u,%u): %zu threads alive, %zu at a barrier\n", bx, by, bz,
                    remaining, waiting);
            abort();
          }
        }
      }
  cur = nullptr;
}

template <class F>
static inline unsigned long long collective(unsigned mask, unsigned long long v, F f, unsigned long long aux = 0) {
  Fiber *me = cur;
  WarpState &w = warps[me->warp];
  unsigned bit = 1u << me->lane;
  mask &= w.exists;
  if (!(mask & bit)) {
    fprintf(stderr, "cuda_emu: lane %u not in its own collective mask %08x\n", me->lane, mask);
    abort();
  }
  w.val[me->lane] = v;
  w.aux[me->lane] = aux;
  w.arrived |= bit;
  if ((w.arrived & mask) == mask) {
    unsigned long long tmp[32];
    for (int l = 0; l < 32; ++l)
      if (mask >> l & 1) tmp[l] = f((unsigned)l, w.val, mask);
    for (int l = 0; l < 32; ++l)
      if (mask >> l & 1) w.res[l] = tmp[l];
    w.arrived &= ~mask;
    w.released |= mask;
    events++;
  }
  while (!(w.released & bit)) yield();
  w.released &= ~bit;
  events++;
  return w.res[me->lane];
}
template <class T>
static inline unsigned long long pack(T v) {
  static_assert(sizeof(T) <= 8, "collective value too wide");
  unsigned long long u = 0;
  memcpy(&u, &v, sizeof(T));
  return u;
}
template <class T>
static inline T unpack(unsigned long long u) {
  T v;
  memcpy(&v, &u, sizeof(T));
  return v;
}
}  // namespace cuemu

#define threadIdx (cuemu::cur->tid)
#define blockIdx (cuemu::g_bid)
#define blockDim (cuemu::g_bdim)
#define gridDim (cuemu::g_gdim)
static const int warpSize = 32;

static inline void __syncthreads() {
  cuemu::cur->at_barrier = true;
  cuemu::yield();
}
static inline void __syncwarp(unsigned mask = 0xffffffffu) {
  cuemu::collective(mask, 0, [](unsigned, const unsigned long long *, unsigned) { return 0ull; });
}
static inline void __threadfence() {}
static inline void __threadfence_block() {}

template <class T>
static inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
  return cuemu::unpack<T>(cuemu::collective(
      mask, cuemu::pack(v),
      [=](unsigned l, const unsigned long long *val, unsigned) {
        const unsigned long long *aux = val + 64;  // WarpState: val[32], res[32], aux[32]
        unsigned s = (l & ~(unsigned)(width - 1)) | ((unsigned)aux[l] & (unsigned)(width - 1));
        return val[s];
      },
      (unsigned long long)(unsigned)src));
}
template <class T>
static inline T __shfl_up_sync(unsigned mask, T v, unsigned d, int width = 32) {
  return cuemu::unpack<T>(cuemu::collective(
      mask, cuemu::pack(v),
      [=](unsigned l, const unsigned long long *val, unsigned) {
        unsigned base = l & ~(unsigned)(width - 1);
        unsigned dd = (unsigned)val[64 + l];
        return (l - base >= dd) ? val[l - dd] : val[l];
      },
      (unsigned long long)d));
}
template <class T>
static inline T __shfl_down_sync(unsigned mask, T v, unsigned d, int width = 32) {
  return cuemu::unpack<T>(cuemu::collective(
      mask, cuemu::pack(v),
      [=](unsigned l, const unsigned long long *val, unsigned) {
        unsigned base = l & ~(unsigned)(width - 1);
        unsigned dd = (unsigned)val[64 + l];
        return (l - base + dd < (unsigned)width) ? val[l + dd] : val[l];
      },
      (unsigned long long)d));
}
template <class T>
static inline T __shfl_xor_sync(unsigned mask, T v, int x, int width = 32) {
  (void)width;
  return cuemu::unpack<T>(cuemu::collective(
      mask, cuemu::pack(v),
      [=](unsigned l, const unsigned long long *val, unsigned) { return val[l ^ (unsigned)val[64 + l]]; },
      (unsigned long long)(unsigned)x));
}
static inline unsigned __ballot_sync(unsigned mask, int pred) {
  return (unsigned)cuemu::collective(mask, pred ? 1ull : 0ull, [](unsigned, const unsigned long long *val, unsigned m) {
    unsigned long long r = 0;
    for (int l = 0; l < 32; ++l)
      if ((m >> l & 1) && val[l]) r |= 1ull << l;
    return r;
  });
}
static inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
static inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, !pred) == 0; }
template <class T>
static inline unsigned __match_any_sync(unsigned mask, T v) {
  return (unsigned)cuemu::collective(mask, cuemu::pack(v), [](unsigned l, const unsigned long long *val, unsigned m) {
    unsigned long long r = 0;
    for (int k = 0; k < 32; ++k)
      if ((m >> k & 1) && val[k] == val[l]) r |= 1ull << k;
    return r;
  });
}
static inline unsigned __reduce_add_sync(unsigned mask, unsigned v) {
  return (unsigned)cuemu::collective(mask, v, [](unsigned, const unsigned long long *val, unsigned m) {
    unsigned long long r = 0;
    for (int k = 0; k < 32; ++k)
      if (m >> k & 1) r += (unsigned)val[k];
    return r & 0xffffffffull;
  });
}
static inline unsigned __reduce_or_sync(unsigned mask, unsigned v) {
  return (unsigned)cuemu::collective(mask, v, [](unsigned, const unsigned long long *val, unsigned m) {
    unsigned long long r = 0;
    for (int k = 0; k < 32; ++k)
      if (m >> k & 1) r |= (unsigned)val[k];
    return r;
  });
}
static inline unsigned __reduce_min_sync(unsigned mask, unsigned v) {
  return (unsigned)cuemu::collective(mask, v, [](unsigned, const unsigned long long *val, unsigned m) {
    unsigned long long r = 0xffffffffull;
    for (int k = 0; k < 32; ++k)
      if ((m >> k & 1) && (unsigned)val[k] < r) r = (unsigned)val[k];
    return r;
  });
}
static inline unsigned __reduce_max_sync(unsigned mask, unsigned v) {
  return (unsigned)cuemu::collective(mask, v, [](unsigned, const unsigned long long *val, unsigned m) {
    unsigned long long r = 0;
    for (int k = 0; k < 32; ++k)
      if ((m >> k & 1) && (unsigned)val[k] > r) r = (unsigned)val[k];
    return r;
  });
}

template <class T>
static inline T atomicAdd(T *p, T v) {
  T o = *p;
  *p = (T)(o + v);
  return o;
}
template <class T>
static inline T atomicOr(T *p, T v) {
  T o = *p;
  *p = o | v;
  return o;
}
template <class T>
static inline T atomicAnd(T *p, T v) {
  T o = *p;
  *p = o & v;
  return o;
}
template <class T>
static inline T atomicMax(T *p, T v) {
  T o = *p;
  if (v > o) *p = v;
  return o;
}
template <class T>
static inline T atomicMin(T *p, T v) {
  T o = *p;
  if (v < o) *p = v;
  return o;
}
template <class T>
static inline T atomicExch(T *p, T v) {
  T o = *p;
  *p = v;
  return o;
}
template <class T>
static inline T atomicCAS(T *p, T cmp, T v) {
  T o = *p;
  if (o == cmp) *p = v;
  return o;
}

static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
static inline int __clzll(long long v) { return v ? __builtin_clzll((unsigned long long)v) : 64; }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline unsigned __brev(unsigned v) {
  unsigned r = 0;
  for (int i = 0; i < 32; ++i) r |= ((v >> i) & 1u) << (31 - i);
  return r;
}
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned s) {
  unsigned long long src = ((unsigned long long)b << 32) | a;
  unsigned r = 0;
  for (int i = 0; i < 4; ++i) {
    unsigned sel = (s >> (4 * i)) & 0xf;
    unsigned byte = (unsigned)(src >> (8 * (sel & 7))) & 0xff;
    if (sel & 8) byte = (byte & 0x80) ? 0xff : 0;
    r |= byte << (8 * i);
  }
  return r;
}
static inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned sh) {
  sh &= 31;
  return sh ? (hi << sh) | (lo >> (32 - sh)) : hi;
}
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned sh) {
  sh &= 31;
  return sh ? (lo >> sh) | (hi << (32 - sh)) : lo;
}
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
template <class T>
static inline T __ldg(const T *p) {
  return *p;
}
// CUDA's global-namespace min / max overloads (device code calls them unqualified)
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned long min(unsigned long a, unsigned long b) { return a < b ? a : b; }
static inline unsigned long max(unsigned long a, unsigned long b) { return a > b ? a : b; }
static inline unsigned long long min(unsigned long long a, unsigned long long b) { return a < b ? a : b; }
static inline unsigned long long max(unsigned long long a, unsigned long long b) { return a > b ? a : b; }
static inline unsigned umin(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned umax(unsigned a, unsigned b) { return a > b ? a : b; }
static inline unsigned long long ullmin(unsigned long long a, unsigned long long b) { return a < b ? a : b; }
static inline unsigned long long ullmax(unsigned long long a, unsigned long long b) { return a > b ? a : b; }

namespace cuemu {
template <class K, class... A>
static inline void launch_k(dim3 grid, dim3 block, K kern, A... args) {
  auto tup = std::make_tuple(args...);
  std::function<void()> body = [&]() { std::apply(kern, tup); };
  launch(grid, block, body);
}
}  // namespace cuemu
#define B200Z_LAUNCH(kern, grid, block, smem, stream, ...) cuemu::launch_k(dim3(grid), dim3(block), kern, __VA_ARGS__)
