// tests/emul/pg_emul.cpp -- TEST INFRASTRUCTURE ONLY.
//
// CPU lane emulator for pindel_b200/csrc/pg_warp.cuh: the 32 lanes of one warp run as ucontext
// fibers and meet at every warp collective (ballot / shfl / redux / syncwarp), so the *same* warp
// program that the CUDA kernel executes can be checked against the oracle on a machine without a
// GPU (pytest -m "not gpu").  Never linked into libpindel_pg.so.
#include <cstddef>
#include <assert.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <atomic>
#include <mutex>
#include <thread>
#include <vector>

// Fibers switch either through a 12-instruction stack swap (x86-64, the default: no system call) or through ucontext
// (-DPG_EMUL_UCONTEXT, and under AddressSanitizer, which knows swapcontext but not a hand-made switch; glibc's swapcontext
// saves and restores the signal mask with a system call per switch, a third of the emulator's time).
#if defined(__SANITIZE_ADDRESS__) && !defined(PG_EMUL_UCONTEXT)
#define PG_EMUL_UCONTEXT 1
#endif
#if !defined(__x86_64__) && !defined(PG_EMUL_UCONTEXT)
#define PG_EMUL_UCONTEXT 1
#endif
#ifdef PG_EMUL_UCONTEXT
#include <ucontext.h>
#else
extern "C" void pgemu_switch(void **save_sp, void *load_sp) __attribute__((visibility("hidden")));
asm(".text\n"
    ".hidden pgemu_switch\n"
    ".globl pgemu_switch\n"
    ".type pgemu_switch,@function\n"
    "pgemu_switch:\n"                      // callee-saved registers onto the old stack, stack pointers exchanged, popped from the new one
    "   pushq %rbp\n   pushq %rbx\n   pushq %r12\n   pushq %r13\n   pushq %r14\n   pushq %r15\n"
    "   movq %rsp, (%rdi)\n"
    "   movq %rsi, %rsp\n"
    "   popq %r15\n   popq %r14\n   popq %r13\n   popq %r12\n   popq %rbx\n   popq %rbp\n"
    "   ret\n"
    ".size pgemu_switch,.-pgemu_switch\n");
#endif

#define PG_EMULATE 1
#define PG_EMUL_TLS thread_local
#define PG_DEV static inline
#define PG_COLD static
#define PG_PIN(x) ((void)(x))
struct uint2 { uint32_t x, y; };

enum { OP_NONE, OP_BALLOT, OP_ANY, OP_SHFL, OP_SHFL_UP, OP_MAX, OP_ADD, OP_OR, OP_SYNC };
struct Emu {
#ifdef PG_EMUL_UCONTEXT
   ucontext_t sched;
   ucontext_t lane_ctx[32];
#else
   void *sched_sp;
   void *lane_sp[32];
#endif
   std::vector<char> stacks[32];
   int cur;
   bool done[32];
   int op[32];
   uint32_t in[32];
   int aux[32];
   uint32_t out[32];
   void (*body)(void *);
   void *arg;
};
// one emulated warp per OS thread (pgemu_search runs several at once): the running warp of this thread
static thread_local Emu *E;

#ifdef PG_EMUL_UCONTEXT
static inline void emu_to_sched(Emu *e, int lane) { swapcontext(&e->lane_ctx[lane], &e->sched); }
static inline void emu_to_lane(Emu *e, int lane) { swapcontext(&e->sched, &e->lane_ctx[lane]); }
#else
static inline void emu_to_sched(Emu *e, int lane) { pgemu_switch(&e->lane_sp[lane], e->sched_sp); }
static inline void emu_to_lane(Emu *e, int lane) { pgemu_switch(&e->sched_sp, e->lane_sp[lane]); }
#endif

static uint32_t emu_collective(int op, uint32_t v, int aux)
{
   Emu *e = E;
   int me = e->cur;
   e->op[me] = op; e->in[me] = v; e->aux[me] = aux;
   emu_to_sched(e, me);
   return e->out[me];
}
static void emu_trampoline()
{
   Emu *e = E;
   e->body(e->arg);
   e->done[e->cur] = true;
   emu_to_sched(e, e->cur);
   abort();                               // a finished lane is never resumed
}
static void emu_run_warp(void (*body)(void *), void *arg)
{
   Emu emu;
   E = &emu;
   emu.body = body; emu.arg = arg;
   for (int i = 0; i < 32; i++) {
      emu.stacks[i].resize(1 << 18);
#ifdef PG_EMUL_UCONTEXT
      getcontext(&emu.lane_ctx[i]);
      emu.lane_ctx[i].uc_stack.ss_sp = emu.stacks[i].data();
      emu.lane_ctx[i].uc_stack.ss_size = emu.stacks[i].size();
      emu.lane_ctx[i].uc_link = &emu.sched;
      makecontext(&emu.lane_ctx[i], emu_trampoline, 0);
#else
      // the frame pgemu_switch pops: six registers, then the trampoline as the return address; above it a null return
      // address, so that the trampoline starts with the stack alignment of a called function (rsp = 16 k + 8)
      uintptr_t top = ((uintptr_t)emu.stacks[i].data() + emu.stacks[i].size()) & ~(uintptr_t)15;
      void **sp = (void **)top;
      *--sp = NULL;
      *--sp = (void *)emu_trampoline;
      for (int r = 0; r < 6; r++) *--sp = NULL;
      emu.lane_sp[i] = (void *)sp;
#endif
      emu.done[i] = false; emu.op[i] = OP_NONE;
   }
   for (;;) {
      int ndone = 0;
      for (int i = 0; i < 32; i++) {
         if (emu.done[i]) { ndone++; continue; }
         emu.cur = i;
         emu_to_lane(&emu, i);
         if (emu.done[i]) ndone++;
      }
      if (ndone == 32) break;
      if (ndone != 0) { fprintf(stderr, "emulator: lanes diverged at exit (%d done)\n", ndone); abort(); }
      int op = emu.op[0];
      for (int i = 1; i < 32; i++) if (emu.op[i] != op) { fprintf(stderr, "emulator: divergent collective %d vs %d (lane %d)\n", op, emu.op[i], i); abort(); }
      uint32_t r = 0;
      switch (op) {
      case OP_BALLOT: for (int i = 0; i < 32; i++) if (emu.in[i]) r |= 1u << i; for (int i = 0; i < 32; i++) emu.out[i] = r; break;
      case OP_ANY: for (int i = 0; i < 32; i++) if (emu.in[i]) r = 1; for (int i = 0; i < 32; i++) emu.out[i] = r; break;
      case OP_SHFL: for (int i = 0; i < 32; i++) emu.out[i] = emu.in[emu.aux[i] & 31]; break;
      case OP_SHFL_UP: for (int i = 0; i < 32; i++) emu.out[i] = (i - emu.aux[i] >= 0) ? emu.in[i - emu.aux[i]] : emu.in[i]; break;
      case OP_MAX: for (int i = 0; i < 32; i++) if (emu.in[i] > r) r = emu.in[i]; for (int i = 0; i < 32; i++) emu.out[i] = r; break;
      case OP_ADD: for (int i = 0; i < 32; i++) r += emu.in[i]; for (int i = 0; i < 32; i++) emu.out[i] = r; break;
      case OP_OR: for (int i = 0; i < 32; i++) r |= emu.in[i]; for (int i = 0; i < 32; i++) emu.out[i] = r; break;
      case OP_SYNC: break;
      default: fprintf(stderr, "emulator: bad op\n"); abort();
      }
   }
   E = NULL;
}

PG_DEV int pg_lane() { return E->cur; }
PG_DEV uint32_t pg_ballot(bool p) { return emu_collective(OP_BALLOT, p ? 1u : 0u, 0); }
PG_DEV bool pg_any(bool p) { return emu_collective(OP_ANY, p ? 1u : 0u, 0) != 0; }
PG_DEV uint32_t pg_shfl(uint32_t v, int src) { return emu_collective(OP_SHFL, v, src); }
PG_DEV uint32_t pg_shfl_up(uint32_t v, int d) { return emu_collective(OP_SHFL_UP, v, d); }
PG_DEV uint32_t pg_redux_max(uint32_t v) { return emu_collective(OP_MAX, v, 0); }
PG_DEV uint32_t pg_redux_add(uint32_t v) { return emu_collective(OP_ADD, v, 0); }
PG_DEV uint32_t pg_redux_or(uint32_t v) { return emu_collective(OP_OR, v, 0); }
PG_DEV void pg_syncwarp() { emu_collective(OP_SYNC, 0, 0); }
PG_DEV int pg_popc(uint32_t v) { return __builtin_popcount(v); }
PG_DEV int pg_ffs(uint32_t v) { return __builtin_ffs((int)v); }
PG_DEV int pg_clz(uint32_t v) { return v ? __builtin_clz(v) : 32; }
PG_DEV uint32_t pg_bfind(uint32_t v) { return v ? 31u - (uint32_t)__builtin_clz(v) : 0xffffffffu; }
PG_DEV uint32_t pg_brev(uint32_t v)
{
   v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
   v = ((v >> 2) & 0x33333333u) | ((v & 0x33333333u) << 2);
   v = ((v >> 4) & 0x0F0F0F0Fu) | ((v & 0x0F0F0F0Fu) << 4);
   v = ((v >> 8) & 0x00FF00FFu) | ((v & 0x00FF00FFu) << 8);
   return (v >> 16) | (v << 16);
}
PG_DEV uint32_t pg_fsr(uint32_t lo, uint32_t hi, uint32_t s) { uint64_t v = ((uint64_t)hi << 32) | lo; return (uint32_t)(v >> (s & 31)); }
PG_DEV uint32_t pg_fsrc(uint32_t lo, uint32_t hi, uint32_t s) { uint64_t v = ((uint64_t)hi << 32) | lo; if (s > 32) s = 32; return (uint32_t)(v >> s); }
PG_DEV uint32_t pg_atomic_add(uint32_t *p, uint32_t v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
PG_DEV void pg_atomic_add64(unsigned long long *p, unsigned long long v) { __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
PG_DEV uint32_t pg_umax(uint32_t a, uint32_t b) { return a > b ? a : b; }
PG_DEV uint32_t pg_umin(uint32_t a, uint32_t b) { return a < b ? a : b; }
PG_DEV uint32_t pg_ldg32(const uint32_t *p) { return *p; }
PG_DEV uint2 pg_ldg64(const uint2 *p) { return *p; }
PG_DEV uint8_t pg_ldg8(const uint8_t *p) { return *p; }

#include "../../pindel_b200/csrc/pg_warp.cuh"
#include "../../pindel_b200/csrc/pg_pack.h"
#include "../../pindel_b200/csrc/pg_params.h"

template <int LW, int TL>
PG_DEV void pg_tile_load(PgWarpSmem<LW, TL> *ws, const PgChrom &c, int64_t w0, int nw, uint32_t &phase)
{
   (void)phase;
   pg_syncwarp();
   if (pg_lane() == 0) {
      assert(nw <= PG_TILE_WORDS);
      for (int i = 0; i < nw; i++) { ws->tile[i] = c.planes[w0 + i]; ws->tile_n[i] = c.nplane[w0 + i]; }
   }
   pg_syncwarp();
}

template <int LW, int TL>
struct EmuArgs {
   PgBatch bt;
   const PgParams *prm;
   const PgChrom *chroms;
   PgWarpSmem<LW, TL> *ws;
};

template <int LW, int TL>
static void emu_warp_body(void *p)
{
   EmuArgs<LW, TL> *a = (EmuArgs<LW, TL> *)p;
   PgWarpCtx<LW, TL> cx;
   cx.ws = a->ws; cx.prm = a->prm; cx.chroms = a->chroms;
   uint32_t phase = 0;
   for (;;) {
      uint32_t ri = 0;
      if (pg_lane() == 0) ri = pg_atomic_add(a->bt.counters, 1u);
      ri = pg_shfl(ri, 0);
      if ((int)ri >= a->bt.n_reads) break;
      pg_process_read<LW, TL>(cx, a->bt, (int)ri, a->bt.scratch, phase);
   }
}

// seeds that entered the far rings' filter, seeds that survived it, tiles filtered: the threads' counts, summed when they end
static unsigned long long g_filter_stats[3];
static std::mutex g_filter_stats_mutex;

static int emu_thread_count(int n_reads)
{
   int t = (int)std::thread::hardware_concurrency();
   if (const char *e = getenv("PG_EMUL_THREADS")) t = atoi(e);
   if (t > 16) t = 16;
   if (t > n_reads / 8) t = n_reads / 8;      // a handful of reads: one warp
   return t < 1 ? 1 : t;
}

// The batch on `threads` emulated warps at once, one OS thread each -- like the grid's warps they share nothing but the read
// ticket, the run pool's allocation counter and the statistics (atomics), and every warp has its own shared-memory image and
// scratch lists.  As on the device, the order of the runs in the pool then depends on the schedule; the records do not.
template <int LW, int TL>
static void run_variant(const PgBatch &bt, const PgParams *prm, const PgChrom *chroms)
{
   const int threads = emu_thread_count(bt.n_reads);
   auto one_warp = [&]() {
      PgWarpSmem<LW, TL> *ws = (PgWarpSmem<LW, TL> *)calloc(1, sizeof(PgWarpSmem<LW, TL>));
      std::vector<PgRun> scratch(3 * (size_t)bt.scratch_cap);
      EmuArgs<LW, TL> a;
      a.bt = bt; a.bt.scratch = scratch.data(); a.prm = prm; a.chroms = chroms; a.ws = ws;
      emu_run_warp(emu_warp_body<LW, TL>, &a);
      free(ws);
      std::lock_guard<std::mutex> g(g_filter_stats_mutex);
      for (int i = 0; i < 3; i++) { g_filter_stats[i] += pg_emul_filter_stats[i]; pg_emul_filter_stats[i] = 0; }
   };
   if (threads == 1) { one_warp(); return; }
   std::vector<std::thread> pool;
   for (int t = 0; t < threads; t++) pool.emplace_back(one_warp);
   for (std::thread &t : pool) t.join();
}

extern "C" {

// Same data contract as pg_search() in include/pindel_pg.h, minus the context: everything is passed in.
// chr_seq: padded ASCII chromosome strings.  params_blob: a PgParams.  Returns 0 / nonzero on overflow.
int pgemu_search(int n_chr, const char *const *chr_seq, const uint32_t *chr_len, const void *params_blob, int n_reads,
                 int mode, const uint8_t *seq_blob, const uint32_t *seq_off, const uint8_t *matched_d,
                 const uint32_t *matched_rel_pos, const int16_t *insert_size, const int32_t *chr_id,
                 const void *close_info, const void *regions, const uint32_t *region_off, void *out /* PgReadOut[n] */,
                 void *run_pool /* PgRun[cap] */, uint32_t run_pool_cap, uint32_t *n_runs_used)
{
   std::vector<PgPackedChrom> packed((size_t)n_chr);
   std::vector<PgChrom> chroms((size_t)n_chr);
   for (int i = 0; i < n_chr; i++) {
      pg_pack_chromosome(chr_seq[i], chr_len[i], PG_PADW, packed[i]);
      chroms[i].planes = (const uint2 *)packed[i].planes.data() + PG_PADW;
      chroms[i].nplane = packed[i].nplane.data() + PG_PADW;
      chroms[i].len = packed[i].len;
      chroms[i].nwords = packed[i].nwords;
   }
   alignas(8) uint32_t counters[8] = {0, 0, 0, 0, 0, 0, 0, 0};
   PgBatch bt;
   bt.n_reads = n_reads; bt.mode = mode; bt.seq_blob = seq_blob; bt.seq_off = seq_off;
   bt.matched_d = matched_d; bt.matched_rel_pos = matched_rel_pos; bt.insert_size = insert_size;
   bt.chr_id = chr_id; bt.close_info = (const PgCloseInfo *)close_info; bt.regions = (const PgRegion *)regions;
   bt.region_off = region_off; bt.out = (PgReadOut *)out; bt.run_pool = (PgRun *)run_pool;
   bt.run_pool_cap = run_pool_cap; bt.counters = counters; bt.scratch = NULL; bt.scratch_cap = 512;      // (scratch: one per warp, run_variant)
   const PgParams *prm = (const PgParams *)params_blob;
   // the same variant choice the library makes: levels needed = max g + ADD + 1, words = longest read / 32
   uint32_t maxlen = 0;
   for (int i = 0; i < n_reads; i++) if (seq_off[i + 1] - seq_off[i] > maxlen) maxlen = seq_off[i + 1] - seq_off[i];
   int need_t = 0;
   for (uint32_t l = 0; l <= maxlen && l < 512; l++) if ((int)prm->g[l] + prm->add_mm + 1 > need_t) need_t = (int)prm->g[l] + prm->add_mm + 1;
   if (maxlen <= 256 && need_t <= 5) run_variant<8, 5>(bt, prm, chroms.data());
   else if (maxlen <= 256 && need_t <= 6) run_variant<8, 6>(bt, prm, chroms.data());
   else if (maxlen <= 256 && need_t <= 7) run_variant<8, 7>(bt, prm, chroms.data());
   else if (maxlen <= 256 && need_t <= 8) run_variant<8, 8>(bt, prm, chroms.data());
   else if (maxlen <= 256 && need_t <= 12) run_variant<8, 12>(bt, prm, chroms.data());
   else if (maxlen <= 256 && need_t <= 16) run_variant<8, 16>(bt, prm, chroms.data());
   else if (need_t <= 8) run_variant<16, 8>(bt, prm, chroms.data());
   else if (need_t <= 16) run_variant<16, 16>(bt, prm, chroms.data());
   else run_variant<16, 32>(bt, prm, chroms.data());
   if (n_runs_used) *n_runs_used = counters[1];
   return (int)counters[2];
}

// seeds that entered the far rings' filter, seeds that survived it, tiles filtered -- since the last call
void pgemu_filter_stats(unsigned long long *out)
{
   std::lock_guard<std::mutex> g(g_filter_stats_mutex);
   for (int i = 0; i < 3; i++) { out[i] = g_filter_stats[i]; g_filter_stats[i] = 0; }
}

int pgemu_sizes(int which)
{
   switch (which) {
   case 0: return (int)sizeof(PgParams);
   case 1: return (int)sizeof(PgRun);
   case 2: return (int)sizeof(PgReadOut);
   case 3: return (int)sizeof(PgRegion);
   case 4: return (int)sizeof(PgCloseInfo);
   case 5: return (int)offsetof(PgParams, g);
   case 6: return (int)offsetof(PgParams, thr);
   default: return -1;
   }
}
}

extern "C" {
// Parameter block exactly as the library derives it (pg_params.h): Pindel's -e/-E/-u/-a/-H/-m/-x.
int pgemu_make_params(void *params_blob, double seq_error_rate, double sensitivity, double max_mismatch_rate, int add_mm,
                      int min_close, int min_perfect, int max_range_index)
{
   uint32_t tab[PG_TABLE];
   pg_prob_table(0.001 + seq_error_rate, sensitivity, tab);     // src/pindel.cpp:856
   return pg_fill_params((PgParams *)params_blob, add_mm, min_close, min_perfect, max_range_index, max_mismatch_rate, tab);
}
}
