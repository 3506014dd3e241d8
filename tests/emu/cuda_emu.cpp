// Fiber scheduler behind cuda_emu.h (x86-64 SysV only) -- test infrastructure, see the header.
#include "cuda_emu.h"
#include <sys/mman.h>
#include <condition_variable>
#include <dlfcn.h>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#if !defined(__x86_64__)
#error "cuda_emu: the fiber switch is written for x86-64"
#endif

// void cuemu_switch(void **save_sp, void *load_sp): saves the callee-saved registers on the current stack, stores the
// stack pointer, loads another stack and restores its callee-saved registers.
extern "C" void cuemu_switch(void **save_sp, void *load_sp);
asm(R"(
.text
.globl cuemu_switch
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
.size cuemu_switch,.-cuemu_switch
)");

// AddressSanitizer build (build_emu.build(asan=True)): "device" allocations are ordinary heap blocks, so every
// out-of-bounds global load/store of a kernel is reported; the unused tail of the dynamic shared memory window is
// poisoned per launch; fiber switches are announced to the runtime.
#if defined(__SANITIZE_ADDRESS__)
#include <sanitizer/asan_interface.h>
#include <sanitizer/common_interface_defs.h>
#define CUEMU_ASAN 1
#else
#define CUEMU_ASAN 0
#endif

namespace cuemu {
thread_local uint3 t_threadIdx, t_blockIdx, t_blockDim, t_gridDim;
thread_local int t_lane;

namespace {
constexpr size_t STACK_BYTES = 256 * 1024;
constexpr int MAX_THREADS = 1024;
constexpr size_t SMEM_BYTES = 232448;     // 227 KB

// a rendezvous of part of a warp (collectives with a partial mask: lane groups that diverged from each other)
struct SubWarp {
  uint32_t mask;                       // 0: free
  uint64_t slot[2][32];
  uint32_t valid[2];
  int kind[2];
  int gen, arrived;
};
constexpr int MAX_SUB = 8;
struct Warp {
  uint64_t slot[2][32];
  uint32_t valid[2];
  int kind[2];
  int gen, arrived, live;
  uint32_t live_mask;
  SubWarp sub[MAX_SUB];
};
struct Fiber {
  void *sp;
  bool done;
  uint3 tid;
  int lane;
  Warp *warp;
  const volatile int *wait_ptr;    // non-null while parked on a barrier generation
  int wait_val;
};
struct BlockRt {
  std::vector<Fiber> fibers;
  std::vector<Warp> warps;
  int bgen, barrived, blive;
  void *main_sp;
  Fiber *cur;
  void (*tramp)(void *);
  void *closure;
  unsigned long long progress;
};
thread_local BlockRt R;
thread_local unsigned char *t_smem = nullptr;
thread_local char *t_stacks = nullptr;
#if CUEMU_ASAN
thread_local const void *t_main_bottom = nullptr;
thread_local size_t t_main_size = 0;
#endif

// fiber -> scheduler
inline void to_main(Fiber *f) {
#if CUEMU_ASAN
  void *fake = nullptr;
  __sanitizer_start_switch_fiber(&fake, t_main_bottom, t_main_size);
  cuemu_switch(&f->sp, R.main_sp);
  __sanitizer_finish_switch_fiber(fake, nullptr, nullptr);
#else
  cuemu_switch(&f->sp, R.main_sp);
#endif
}
// scheduler -> fiber t
inline void to_fiber(Fiber &f, int t) {
#if CUEMU_ASAN
  void *fake = nullptr;
  __sanitizer_start_switch_fiber(&fake, t_stacks + STACK_BYTES * (size_t)t, STACK_BYTES);
  cuemu_switch(&R.main_sp, f.sp);
  __sanitizer_finish_switch_fiber(fake, nullptr, nullptr);
#else
  (void)t;
  cuemu_switch(&R.main_sp, f.sp);
#endif
}

void park(const volatile int *gen_ptr, int gen) {
  Fiber *f = R.cur;
  f->wait_ptr = gen_ptr;
  f->wait_val = gen;
  while (*gen_ptr == gen) to_main(f);
  f->wait_ptr = nullptr;
}

void reorder(std::vector<int> &order, int sched, uint64_t &rng) {
  const int n = (int)order.size();
  if (sched == 1) { for (int t = 0; t < n; t++) order[t] = n - 1 - t; return; }
  for (int t = n - 1; t > 0; t--) {                    // Fisher-Yates with xorshift64
    rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17;
    std::swap(order[t], order[(int)(rng % (uint64_t)(t + 1))]);
  }
}

[[noreturn]] void fiber_main() {
#if CUEMU_ASAN
  __sanitizer_finish_switch_fiber(nullptr, &t_main_bottom, &t_main_size);    // first entry: learn the scheduler's stack
#endif
  R.tramp(R.closure);
  Fiber *f = R.cur;
  f->done = true;
  R.progress++;
  Warp &w = *f->warp;
  w.live--;
  w.live_mask &= ~(1u << f->lane);
  if (w.arrived > 0 && w.arrived == w.live) { w.arrived = 0; w.gen++; }        // the rest of the warp was waiting for us
  for (SubWarp &q : w.sub)
    if (q.mask && q.arrived > 0 && q.arrived == __builtin_popcount(q.mask & w.live_mask)) { q.arrived = 0; q.gen++; }
  R.blive--;
  if (R.barrived > 0 && R.barrived == R.blive) { R.barrived = 0; R.bgen++; }
  for (;;) to_main(f);
}
}  // namespace

const uint64_t *warp_publish(uint64_t v, uint32_t *valid, int kind) {
  Fiber *f = R.cur;
  Warp &w = *f->warp;
  const int g = w.gen, b = g & 1;
  if (w.arrived == 0) { w.valid[b] = 0; w.kind[b] = kind; }
  else if (w.kind[b] != kind) {
    std::fprintf(stderr, "cuda_emu: divergent warp collective (kind %d vs %d) in block %u thread %u\n", w.kind[b], kind, t_blockIdx.x, f->tid.x);
    std::abort();
  }
  w.slot[b][f->lane] = v;
  w.valid[b] |= 1u << f->lane;
  R.progress++;
  if (++w.arrived == w.live) { w.arrived = 0; w.gen++; }
  else park(&w.gen, g);
  *valid = w.valid[b];
  return w.slot[b];
}

// collective among the lanes of `mask` only (the named lanes must all reach it, like on hardware; lanes outside it are free to be
// anywhere else, e.g. in another group's collective).  Masks in flight at the same time must be equal or disjoint.
const uint64_t *warp_publish_masked(unsigned mask, uint64_t v, uint32_t *valid, int kind) {
  if (mask == 0xffffffffu) return warp_publish(v, valid, kind);
  Fiber *f = R.cur;
  Warp &w = *f->warp;
  if (!((mask >> f->lane) & 1u)) { std::fprintf(stderr, "cuda_emu: lane %d calls a collective whose mask %08x does not name it\n", f->lane, mask); std::abort(); }
  SubWarp *q = nullptr;
  for (SubWarp &c : w.sub) if (c.mask == mask) { q = &c; break; }
  if (!q) for (SubWarp &c : w.sub) {         // a slot belongs to one mask for the whole block (parked lanes keep pointers into it)
    if (c.mask && (c.mask & mask)) { std::fprintf(stderr, "cuda_emu: overlapping partial masks %08x / %08x\n", c.mask, mask); std::abort(); }
    if (!q && c.mask == 0) q = &c;
  }
  if (!q) { std::fprintf(stderr, "cuda_emu: more than %d partial-mask groups in one warp\n", MAX_SUB); std::abort(); }
  if (q->mask != mask) { q->mask = mask; q->gen = 0; q->arrived = 0; }
  const int g = q->gen, b = g & 1;
  if (q->arrived == 0) { q->valid[b] = 0; q->kind[b] = kind; }
  else if (q->kind[b] != kind) {
    std::fprintf(stderr, "cuda_emu: divergent group collective (kind %d vs %d, mask %08x) in block %u thread %u\n", q->kind[b], kind, mask, t_blockIdx.x, f->tid.x);
    std::abort();
  }
  q->slot[b][f->lane] = v;
  q->valid[b] |= 1u << f->lane;
  R.progress++;
  if (++q->arrived == __builtin_popcount(mask & w.live_mask)) { q->arrived = 0; q->gen++; }
  else park(&q->gen, g);
  *valid = q->valid[b];
  return q->slot[b];
}

void block_barrier() {
  const int g = R.bgen;
  R.progress++;
  if (++R.barrived == R.blive) { R.barrived = 0; R.bgen++; }
  else park(&R.bgen, g);
}

unsigned char *dyn_smem() { return t_smem; }

namespace {
std::mutex g_count_mu;
std::map<std::string, long> g_counts;
}  // namespace
void note_launch(const char *kernel) {
  std::lock_guard<std::mutex> lk(g_count_mu);
  g_counts[kernel]++;
}

void *dev_alloc(size_t bytes) {
  if (!bytes) return nullptr;
#if CUEMU_ASAN
  void *p = nullptr;                                   // exact size: the red zone starts right behind the last byte
  if (posix_memalign(&p, 256, bytes)) return nullptr;
  std::memset(p, 0xCD, bytes);
  return p;
#else
  const size_t n = (bytes + 255) / 256 * 256;
  void *p = std::aligned_alloc(256, n);
  if (p) std::memset(p, 0xCD, n);
  return p;
#endif
}
void dev_free(void *p) { std::free(p); }

int num_sms() {
  const char *e = std::getenv("CUEMU_SMS");
  const int n = e ? std::atoi(e) : 2;
  return n > 0 ? n : 2;
}

void launch_impl(dim3 grid, dim3 block, size_t smem, void (*tramp)(void *), void *closure) {
  const int nthreads = (int)(block.x * block.y * block.z);
  if (nthreads <= 0 || nthreads > MAX_THREADS || smem > SMEM_BYTES) {
    std::fprintf(stderr, "cuda_emu: bad launch configuration (%d threads, %zu bytes of shared memory)\n", nthreads, smem);
    std::abort();
  }
  if (R.cur) { std::fprintf(stderr, "cuda_emu: nested launch\n"); std::abort(); }
  if (!t_smem) t_smem = (unsigned char *)std::aligned_alloc(1024, (SMEM_BYTES + 1023) / 1024 * 1024);
  if (!t_stacks) {
    t_stacks = (char *)mmap(nullptr, STACK_BYTES * MAX_THREADS, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (t_stacks == (char *)MAP_FAILED) { std::perror("cuda_emu: mmap"); std::abort(); }
  }
  const int nwarps = (nthreads + 31) / 32;
  // CUEMU_SCHED: order in which the runnable threads of a block get the CPU between two collectives.
  //   unset/0 = thread order; 1 = reverse; N >= 2 = a fresh pseudo-random permutation (seed N) before every pass.
  // Code that is correct on the GPU cannot depend on it, so results must not change -- a cheap detector for missing
  // __syncwarp()/__syncthreads() and other order assumptions.
  static const int sched = [] { const char *e = std::getenv("CUEMU_SCHED"); return e ? std::atoi(e) : 0; }();
  static thread_local std::vector<int> order;
  static thread_local uint64_t rng = 0;
  if (sched != 0) { order.resize(nthreads); for (int t = 0; t < nthreads; t++) order[t] = t; if (!rng) rng = 0x9E3779B97F4A7C15ull * (uint64_t)sched; }
  R.fibers.resize(nthreads);
  R.warps.resize(nwarps);
  R.tramp = tramp;
  R.closure = closure;
  t_blockDim = {block.x, block.y, block.z};
  t_gridDim = {grid.x, grid.y, grid.z};
  for (unsigned bz = 0; bz < grid.z; bz++)
    for (unsigned by = 0; by < grid.y; by++)
      for (unsigned bx = 0; bx < grid.x; bx++) {
        t_blockIdx = {bx, by, bz};
#if CUEMU_ASAN
        __asan_unpoison_memory_region(t_smem, SMEM_BYTES);
#endif
        std::memset(t_smem, 0xCD, smem);
#if CUEMU_ASAN
        __asan_poison_memory_region(t_smem + smem, SMEM_BYTES - smem);               // beyond what this launch asked for
#endif
        for (int w = 0; w < nwarps; w++) { R.warps[w].gen = 0; R.warps[w].arrived = 0; R.warps[w].live = std::min(32, nthreads - 32 * w);
          R.warps[w].live_mask = R.warps[w].live >= 32 ? 0xffffffffu : ((1u << R.warps[w].live) - 1u);
          for (SubWarp &q : R.warps[w].sub) { q.mask = 0; q.arrived = 0; q.gen = 0; } }
        R.bgen = 0; R.barrived = 0; R.blive = nthreads; R.progress = 0;
        for (int t = 0; t < nthreads; t++) {
          Fiber &f = R.fibers[t];
          f.done = false;
          f.tid = {(unsigned)t % block.x, ((unsigned)t / block.x) % block.y, (unsigned)t / (block.x * block.y)};
          f.lane = t & 31;
          f.warp = &R.warps[t >> 5];
          f.wait_ptr = nullptr;
          void **sp = (void **)(t_stacks + STACK_BYTES * (size_t)(t + 1));
          *--sp = nullptr;                    // fake return address: keeps the stack 16-byte aligned as after a call
          *--sp = (void *)&fiber_main;
          for (int i = 0; i < 6; i++) *--sp = nullptr;
          f.sp = sp;
        }
        while (R.blive > 0) {
          const unsigned long long before = R.progress;
          if (sched != 0) reorder(order, sched, rng);
          for (int k = 0; k < nthreads; k++) {
            const int t = sched != 0 ? order[k] : k;
            Fiber &f = R.fibers[t];
            if (f.done || (f.wait_ptr && *f.wait_ptr == f.wait_val)) continue;
            R.cur = &f;
            t_threadIdx = f.tid;
            t_lane = f.lane;
            to_fiber(f, t);
          }
          if (R.progress == before && R.blive > 0) {
            std::fprintf(stderr, "cuda_emu: deadlock in block (%u,%u,%u): threads wait on a barrier the others never reach\n", bx, by, bz);
            std::abort();
          }
        }
        R.cur = nullptr;
      }
}
}  // namespace cuemu

// Launch counters for the tests: how often kernels whose name starts with `prefix` were launched since the last reset.
extern "C" long cuemu_launches(const char *prefix) {
  std::lock_guard<std::mutex> lk(cuemu::g_count_mu);
  long n = 0;
  const std::string p(prefix ? prefix : "");
  for (auto &kv : cuemu::g_counts) if (kv.first.compare(0, p.size(), p) == 0) n += kv.second;
  return n;
}
extern "C" void cuemu_reset_launches() {
  std::lock_guard<std::mutex> lk(cuemu::g_count_mu);
  cuemu::g_counts.clear();
}

// ---- in-process stand-in for NCCL: the ranks of a sharded run are OS threads of one test process ----
// build_emu.py points the driver's dlopen("libnccl.so.2") at this library, so the sharded code path (comm init, the
// per-round all-gather, the final all-reduces) runs unchanged.  Collectives rendezvous on a condition variable.
namespace {
struct FakeComm {
  std::mutex mu;
  std::condition_variable cv;
  int world = 0, joined = 0, arrived = 0, left = 0;
  unsigned long long gen = 0;
  std::vector<const void *> src;
  std::vector<void *> dst;
};
struct FakeRank { FakeComm *c; int rank; };
std::mutex g_comm_mu;
std::map<std::string, FakeComm *> g_comms;
unsigned long long g_next_id = 1;

// Runs `work` (on the last arriving rank, with every rank's buffers visible) between two rendezvous.
template <class W> void rendezvous(FakeRank *r, const void *send, void *recv, W work) {
  FakeComm &c = *r->c;
  std::unique_lock<std::mutex> lk(c.mu);
  c.src[r->rank] = send;
  c.dst[r->rank] = recv;
  const unsigned long long g = c.gen;
  if (++c.arrived == c.world) { work(c); c.arrived = 0; c.gen++; c.cv.notify_all(); }
  else c.cv.wait(lk, [&] { return c.gen != g; });
}
size_t dtype_bytes(int dt) { return dt <= 1 ? 1 : dt <= 3 ? 4 : dt <= 5 ? 8 : dt == 6 ? 2 : dt == 7 ? 4 : 8; }
}  // namespace

extern "C" const char *cuemu_self_path() {
  static std::string path;
  Dl_info info;
  if (path.empty() && dladdr((void *)&cuemu_self_path, &info) && info.dli_fname) path = info.dli_fname;
  return path.c_str();
}
extern "C" int ncclGetUniqueId(char *id128) {
  std::lock_guard<std::mutex> lk(g_comm_mu);
  std::memset(id128, 0, 128);
  std::snprintf(id128, 128, "cuemu-%llu", g_next_id++);
  return 0;
}
struct FakeId { char internal[128]; };
extern "C" int ncclCommInitRank(void **comm, int world, FakeId id, int rank) {
  FakeComm *c;
  {
    std::lock_guard<std::mutex> lk(g_comm_mu);
    FakeComm *&slot = g_comms[std::string(id.internal)];
    if (!slot) { slot = new FakeComm(); slot->world = world; slot->src.resize(world); slot->dst.resize(world); }
    c = slot;
  }
  if (c->world != world || rank < 0 || rank >= world) return 5;
  *comm = new FakeRank{c, rank};
  std::unique_lock<std::mutex> lk(c->mu);
  c->joined++;
  c->cv.notify_all();
  c->cv.wait(lk, [&] { return c->joined >= c->world; });       // like NCCL: returns once every rank has joined
  return 0;
}
extern "C" int ncclAllGather(const void *send, void *recv, size_t count, int dtype, void *comm, void *) {
  const size_t bytes = count * dtype_bytes(dtype);
  rendezvous((FakeRank *)comm, send, recv, [&](FakeComm &c) {
    std::vector<unsigned char> tmp(bytes * c.world);
    for (int r = 0; r < c.world; r++) std::memcpy(tmp.data() + bytes * r, c.src[r], bytes);     // staged: send may alias recv
    for (int r = 0; r < c.world; r++) std::memcpy(c.dst[r], tmp.data(), tmp.size());
  });
  return 0;
}
extern "C" int ncclAllReduce(const void *send, void *recv, size_t count, int dtype, int op, void *comm, void *) {
  if (op != 0 && !(op == 2 && dtype == 2)) return 4;              // ncclSum, and ncclMax on int32
  rendezvous((FakeRank *)comm, send, recv, [&](FakeComm &c) {
    const size_t eb = dtype_bytes(dtype);
    std::vector<unsigned char> acc(count * eb, 0);
    if (op == 2) std::memcpy(acc.data(), c.src[0], acc.size());
    for (int r = 0; r < c.world; r++)
      for (size_t i = 0; i < count; i++) {
        if (op == 2) { int32_t &a = ((int32_t *)acc.data())[i]; a = std::max(a, ((const int32_t *)c.src[r])[i]); }
        else if (dtype == 2 || dtype == 3) ((uint32_t *)acc.data())[i] += ((const uint32_t *)c.src[r])[i];
        else if (dtype == 4 || dtype == 5) ((uint64_t *)acc.data())[i] += ((const uint64_t *)c.src[r])[i];
        else if (dtype == 8) ((double *)acc.data())[i] += ((const double *)c.src[r])[i];
        else if (dtype <= 1) acc[i] += ((const unsigned char *)c.src[r])[i];
        else std::abort();
      }
    for (int r = 0; r < c.world; r++) std::memcpy(c.dst[r], acc.data(), acc.size());
  });
  return 0;
}
extern "C" int ncclCommDestroy(void *comm) { delete (FakeRank *)comm; return 0; }
extern "C" const char *ncclGetErrorString(int) { return "cuda_emu fake NCCL error"; }
