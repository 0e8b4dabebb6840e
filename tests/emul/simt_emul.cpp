// tests/emul/simt_emul.cpp -- TEST INFRASTRUCTURE ONLY (see simt_emul.h).
#include "simt_emul.h"

#include <sys/mman.h>
#include <ucontext.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace simt {

Dim3 threadIdx, blockIdx, blockDim, gridDim;
uint8_t *smem = nullptr;

namespace {

enum State { RUNNABLE, AT_BARRIER, AT_WARP, SPINNING, DONE };

struct Warp {
  unsigned arrived = 0;     // lanes that have deposited for the collective in flight
  unsigned gen = 0;         // completed collectives
  uint64_t in[32];
  uint64_t snap[2][32];     // deposits of the last completed collectives (by generation parity)
  unsigned live = 32;       // lanes that have not exited
};

struct Cta;
struct Fiber {
  ucontext_t ctx;
  void *stack = nullptr;
  State st = RUNNABLE;
  unsigned tid = 0;
  Cta *cta = nullptr;
  unsigned bar_gen = 0;   // barrier generation this fiber waits for
  unsigned warp_gen = 0;  // collective generation this fiber waits for
};

struct Cta {
  unsigned bid = 0;
  std::vector<Fiber> fibers;
  std::vector<Warp> warps;
  std::vector<uint8_t> smem;
  unsigned bar_arrived = 0, bar_gen = 0, live = 0;
  int bar_or = 0, bar_or_result[2] = {0, 0};
};

const size_t kStack = 256 << 10;
ucontext_t g_sched;
Fiber *g_cur = nullptr;
const std::function<void()> *g_body = nullptr;

void set_identity(Fiber *f) {
  threadIdx = Dim3{f->tid, 0, 0};
  blockIdx = Dim3{f->cta->bid, 0, 0};
  smem = f->cta->smem.data();
}

void yield_to_scheduler() {
  Fiber *me = g_cur;
  swapcontext(&me->ctx, &g_sched);
  set_identity(me);
}

void fiber_main() {
  (*g_body)();
  Fiber *me = g_cur;
  me->st = DONE;
  Cta *c = me->cta;
  c->live--;
  Warp &w = c->warps[me->tid >> 5];
  w.live--;
  // a collective that was only waiting for this lane completes without it
  if (w.live && w.arrived == w.live) {
    memcpy(w.snap[w.gen & 1], w.in, sizeof w.in);
    w.arrived = 0;
    w.gen++;
  }
  if (c->live && c->bar_arrived == c->live) {  // exited threads count as arrived
    c->bar_or_result[c->bar_gen & 1] = c->bar_or;
    c->bar_or = 0;
    c->bar_arrived = 0;
    c->bar_gen++;
  }
  swapcontext(&me->ctx, &g_sched);
}

// deposit v, wait for the whole warp, return the snapshot of all deposits
const uint64_t *collective(uint64_t v) {
  Fiber *me = g_cur;
  Warp &w = me->cta->warps[me->tid >> 5];
  const unsigned lane = me->tid & 31;
  const unsigned my_gen = w.gen;
  w.in[lane] = v;
  w.arrived++;
  if (w.arrived == w.live) {
    memcpy(w.snap[my_gen & 1], w.in, sizeof w.in);
    w.arrived = 0;
    w.gen++;
  } else {
    me->st = AT_WARP;
    me->warp_gen = my_gen + 1;
    yield_to_scheduler();
  }
  return w.snap[my_gen & 1];
}

}  // namespace

void syncthreads() { (void)syncthreads_or(0); }

int syncthreads_or(int pred) {
  Fiber *me = g_cur;
  Cta *c = me->cta;
  const unsigned my_gen = c->bar_gen;
  c->bar_or |= pred ? 1 : 0;
  c->bar_arrived++;
  if (c->bar_arrived == c->live) {
    c->bar_or_result[my_gen & 1] = c->bar_or;
    c->bar_or = 0;
    c->bar_arrived = 0;
    c->bar_gen++;
  } else {
    me->st = AT_BARRIER;
    me->bar_gen = my_gen + 1;
    yield_to_scheduler();
  }
  return c->bar_or_result[my_gen & 1];
}

void syncwarp() { (void)collective(0); }

void spin_yield() {
  g_cur->st = SPINNING;
  yield_to_scheduler();
}

uint64_t shfl_idx64(uint64_t v, int src) { return collective(v)[src & 31]; }
uint64_t shfl_up64(uint64_t v, unsigned d) {
  const unsigned lane = g_cur->tid & 31;
  const uint64_t *s = collective(v);
  return lane >= d ? s[lane - d] : v;
}
uint64_t shfl_down64(uint64_t v, unsigned d) {
  const unsigned lane = g_cur->tid & 31;
  const uint64_t *s = collective(v);
  return lane + d < 32 ? s[lane + d] : v;
}
uint64_t shfl_xor64(uint64_t v, unsigned m) {
  const unsigned lane = g_cur->tid & 31;
  return collective(v)[(lane ^ m) & 31];
}
uint32_t ballot(int pred) {
  const uint64_t *s = collective(pred ? 1 : 0);
  uint32_t r = 0;
  for (int i = 0; i < 32; ++i) r |= (uint32_t)(s[i] & 1u) << i;
  return r;
}
uint32_t reduce_add(uint32_t v) {
  const uint64_t *s = collective(v);
  uint32_t r = 0;
  for (int i = 0; i < 32; ++i) r += (uint32_t)s[i];
  return r;
}
uint32_t reduce_max(uint32_t v) {
  const uint64_t *s = collective(v);
  uint32_t r = 0;
  for (int i = 0; i < 32; ++i) r = (uint32_t)s[i] > r ? (uint32_t)s[i] : r;
  return r;
}
uint32_t reduce_min(uint32_t v) {
  const uint64_t *s = collective(v);
  uint32_t r = 0xffffffffu;
  for (int i = 0; i < 32; ++i) r = (uint32_t)s[i] < r ? (uint32_t)s[i] : r;
  return r;
}

void launch(unsigned grid, unsigned block, size_t smem_bytes, const std::function<void()> &body) {
  if (block == 0 || block % 32 != 0) { fprintf(stderr, "simt: block size must be a multiple of 32\n"); abort(); }
  if (g_cur) { fprintf(stderr, "simt: nested launch\n"); abort(); }
  gridDim = Dim3{grid, 1, 1};
  blockDim = Dim3{block, 1, 1};
  g_body = &body;
  std::vector<Cta> ctas(grid);
  for (unsigned b = 0; b < grid; ++b) {
    Cta &c = ctas[b];
    c.bid = b;
    c.fibers.resize(block);
    c.warps.resize(block / 32);
    c.smem.assign(smem_bytes + 64, 0xCD);  // shared memory is not zero on a GPU either
    c.live = block;
    for (unsigned t = 0; t < block; ++t) {
      Fiber &f = c.fibers[t];
      f.tid = t;
      f.cta = &c;
      f.stack = mmap(nullptr, kStack, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
      if (f.stack == MAP_FAILED) { perror("simt: mmap"); abort(); }
      getcontext(&f.ctx);
      f.ctx.uc_stack.ss_sp = f.stack;
      f.ctx.uc_stack.ss_size = kStack;
      f.ctx.uc_link = nullptr;
      makecontext(&f.ctx, fiber_main, 0);
    }
  }
  // scheduler: CTAs round-robin; inside a CTA warp after warp, every lane run until it blocks
  unsigned long long idle_rounds = 0;
  for (;;) {
    bool any_live = false, progress = false;
    for (unsigned b = 0; b < grid; ++b) {
      Cta &c = ctas[b];
      if (!c.live) continue;
      any_live = true;
      for (unsigned w = 0; w < block / 32; ++w) {
        bool again = true;
        while (again) {
          again = false;
          for (unsigned l = 0; l < 32; ++l) {
            Fiber &f = c.fibers[32 * w + l];
            if (f.st == DONE) continue;
            if (f.st == AT_BARRIER) { if (c.bar_gen < f.bar_gen) continue; f.st = RUNNABLE; }
            if (f.st == AT_WARP) { if (c.warps[w].gen < f.warp_gen) continue; f.st = RUNNABLE; }
            const bool was_spinning = f.st == SPINNING;
            f.st = RUNNABLE;
            g_cur = &f;
            set_identity(&f);
            swapcontext(&g_sched, &f.ctx);
            g_cur = nullptr;
            if (!(was_spinning && f.st == SPINNING)) progress = true;
            // a lane that stopped at a collective may have completed it for the others: sweep the warp again
            if (f.st != SPINNING && f.st != AT_BARRIER) again = true;
          }
        }
      }
    }
    if (!any_live) break;
    if (!progress && ++idle_rounds > 1000000ull) { fprintf(stderr, "simt: deadlock (no fiber can make progress)\n"); abort(); }
    if (progress) idle_rounds = 0;
  }
  for (Cta &c : ctas)
    for (Fiber &f : c.fibers) munmap(f.stack, kStack);
  g_body = nullptr;
}

}  // namespace simt
