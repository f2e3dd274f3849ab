// pindel_b200/csrc/pindel_pg.cu -- kernels + C ABI of libpindel_pg.so (sm_100a only).
//
//   pg_pack_kernel    : ASCII chromosome -> {lo,hi} bit-planes + N plane (2 bits/base + 1 bit/base)
//   pg_search_kernel  : persistent grid, one warp per read (pg_warp.cuh); the reference window of every search is
//                       staged into shared memory with TMA bulk copies (cp.async.bulk + mbarrier), results go to a
//                       run pool through one atomic allocation per read
//   extern "C" pg_*   : the ABI declared in include/pindel_pg.h
//
// There is no host implementation of the search in this library: without a CUDA device every call fails.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <string>
#include <thread>
#include <vector>

#include "pg_warp.cuh"
#include "pg_sv.cuh"
#include "pg_rp.cuh"
#include "pg_rp_host.h"
#include "pg_bgzf.cuh"
#include "pg_bgzf_host.h"
#include "pg_bam.cuh"
#include "pg_bam_host.h"
#include "pg_sr.cuh"
#include "pg_boxsort.cuh"
#include "pg_bd.cuh"
#include "pg_bd_host.h"
#include "pg_params.h"
#include "../../include/pindel_pg.h"
#include "pg_host_helpers.h"

static_assert(sizeof(pg_run) == sizeof(PgRun) && sizeof(pg_run) == 12, "run layout");
static_assert(sizeof(pg_read_out) == sizeof(PgReadOut) && sizeof(pg_read_out) == 12, "read_out layout");
static_assert(sizeof(pg_region) == sizeof(PgRegion) && sizeof(pg_region) == 12, "region layout");
static_assert(sizeof(pg_close_info) == sizeof(PgCloseInfo) && sizeof(pg_close_info) == 8, "close_info layout");
static_assert(PG_READ_FLIPPED == PG_F_FLIPPED && PG_READ_UNSUPPORTED == PG_F_TOOLONG && PG_READ_OVERFLOW == PG_F_OVERFLOW, "flags");

#define PG_PIPE_CHUNK_DEFAULT (1 << 19)   // reads per chunk of the pipelined path; PG_PIPE_CHUNK in the environment overrides it (tests)
#define PG_BLOCK_THREADS 32
#ifndef PG_MINB
#define PG_MINB 24         // resident CTAs (= warps) per SM the register allocator is asked to make room for (TL <= 8 variants):
                           // measured (scripts/kernel_ab.py, profiles/r2n_kernel_ab.txt) 32 warps at 64 registers 79.4 M reads/s,
                           // 24 at 80 registers 81.5, 20 at 96 75.2, 16 at 113 65.6 on the headline workload
#endif
#define PG_WARPS_PER_BLOCK (PG_BLOCK_THREADS / 32)

// ------------------------------------------------------------------------------------------------
// TMA tile staging: two bulk copies (bit-planes, N plane) completing on the warp's mbarrier
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pg_smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int LW, int TL>
__device__ __forceinline__ void pg_tile_load(PgWarpSmem<LW, TL> *ws, const PgChrom &c, int64_t w0, int nw, uint32_t &phase)
{
   __syncwarp();                                      // every lane is done with the previous tile
   uint32_t bar = pg_smem_addr(&ws->mbar);
   if (pg_lane() == 0) {
      uint32_t bytes_p = (uint32_t)nw * 8u, bytes_n = (uint32_t)nw * 4u;
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes_p + bytes_n) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                      pg_smem_addr(ws->tile)),
                   "l"(c.planes + w0), "r"(bytes_p), "r"(bar)
                   : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                      pg_smem_addr(ws->tile_n)),
                   "l"(c.nplane + w0), "r"(bytes_n), "r"(bar)
                   : "memory");
   }
   uint32_t done = 0;
   while (!done) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done)
                   : "r"(bar), "r"(phase)
                   : "memory");
   }
   phase ^= 1u;
}

// One warp per CTA: the warp's working set is a static __shared__ object, so every access to it is an LDS / STS with a
// link-time constant address (with several warps per CTA the per-warp base had to be re-derived from the shared window
// all over the hot loop, a few percent of all instructions), and the parameter block is read from the constant bank the
// kernel arguments live in.  The warps of an SM are its resident CTAs (PG_MINB of them).
template <int LW, int TL>
__global__ void __launch_bounds__(32, (TL <= 8) ? PG_MINB : 1) pg_search_kernel(const __grid_constant__ PgBatch bt, const __grid_constant__ PgParams prm,
                                                                                const PgChrom *__restrict__ chroms)
{
   __shared__ __align__(128) PgWarpSmem<LW, TL> ws_s;
   PgWarpSmem<LW, TL> *ws = &ws_s;
   if (pg_lane() == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(pg_smem_addr(&ws->mbar)) : "memory");
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
   }
   __syncwarp();
   PgWarpCtx<LW, TL> cx;
   cx.ws = ws; cx.prm = &prm; cx.chroms = chroms;
   uint32_t phase = 0;
   PgRun *scratch = bt.scratch + (size_t)blockIdx.x * 3u * bt.scratch_cap;
   for (;;) {
      uint32_t ri = 0;
      if (pg_lane() == 0) ri = atomicAdd(bt.counters, 1u);
      ri = __shfl_sync(0xffffffffu, ri, 0);
      if ((int)ri >= bt.n_reads) break;
      pg_process_read<LW, TL>(cx, bt, (int)ri, scratch, phase);
   }
}

// One warp per 32 bases: coalesced byte loads, three ballots, one 8-byte + one 4-byte store.
__global__ void pg_pack_kernel(const uint8_t *__restrict__ ascii, uint32_t n_bases, uint32_t first_word, uint2 *planes,
                               uint32_t *nplane)
{
   uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
   uint32_t lane = threadIdx.x & 31;
   uint32_t nw = (n_bases + 31u) >> 5;
   if (gw >= nw) return;
   uint32_t i = gw * 32u + lane;
   uint8_t c = (i < n_bases) ? ascii[i] : (uint8_t)'N';
   uint32_t code = (c == 'C') ? 1u : (c == 'G') ? 2u : (c == 'T') ? 3u : 0u;
   bool isn = !(c == 'A' || c == 'C' || c == 'G' || c == 'T');
   uint32_t lo = __ballot_sync(0xffffffffu, code & 1u), hi = __ballot_sync(0xffffffffu, code & 2u);
   uint32_t nn = __ballot_sync(0xffffffffu, isn);
   if (lane == 0) {
      planes[first_word + gw] = make_uint2(lo, hi);
      nplane[first_word + gw] = nn;
   }
}

// ------------------------------------------------------------------------------------------------
// SV classification: one thread per read runs the seven drivers (pg_sv.cuh).  The work per read is a few hundred
// integer operations over its own run lists, so the stage is bound by launch + record traffic, not by issue slots.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) pg_sv_kernel(PgSvBatch b, const PgSvParams *__restrict__ P, const PgParams *__restrict__ prm,
                                                    const PgChrom *__restrict__ chroms)
{
   const int t = blockIdx.x * blockDim.x + threadIdx.x;
   if (t >= b.n_reads) return;
   const int i = b.perm ? (int)b.perm[t] : t;
   PgSvEnv ev;
   ev.P = P;
   ev.chrom = chroms[P->chr];
   ev.records = b.records;
   ev.records_cap = b.records_cap;
   ev.counters = b.counters;
   ev.counts = b.counts;
   ev.n_reads = (uint32_t)b.n_reads;
   PgSvRead r;
   const uint32_t s0 = b.seq_off[i], s1 = b.seq_off[i + 1];
   r.seq = b.seq_blob + s0;
   r.rl = (int32_t)(s1 - s0);
   bool other_chr = false;
   if (b.res_out) {
      const PgReadOut o = b.res_out[i];
      if (o.flags & (PG_F_TOOLONG | PG_F_OVERFLOW)) return;
      r.rl = o.len_after;      // setUnmatchedSeq's strip and the flips are already in the resident sequence
      r.fr = b.res_pool + o.run_off + o.n_close_runs;
      r.nfr = o.n_far_runs;
      if (b.close_off) {
         r.cr = b.close_runs + b.close_off[i];
         r.ncr = (int32_t)(b.close_off[i + 1] - b.close_off[i]);
      } else {
         r.cr = b.res_pool + o.run_off;
         r.ncr = o.n_close_runs;
      }
      // FarFragName (UpdateFarFragName, src/pindel.cpp:1262-1270) is the chromosome of UP_Far[0]
      other_chr = r.nfr > 0 && (int32_t)r.fr[0].chr != (b.chr_id ? b.chr_id[i] : 0);
   } else {
      const uint32_t r0 = b.run_off[i], r1 = b.run_off[i + 1];
      r.cr = b.runs + r0;
      r.ncr = b.n_close_runs[i];
      r.fr = r.cr + r.ncr;
      r.nfr = (int32_t)(r1 - r0) - r.ncr;
   }
   r.max_e = prm->g[r.rl < 512 ? r.rl : 511];
   r.matched_d = b.matched_d[i];
   r.insert_size = b.insert_size[i];
   r.index = (uint32_t)i;
   r.left = r.right = 0;
   r.bp_left = r.bp_right = r.indel = 0;
   r.bp = 0;
   r.nt_size = r.nt_start = r.nt_len = r.nt_rot = 0;
   r.nt_src = PG_NT_EMPTY;
   const uint8_t fl = b.rflags ? b.rflags[i] : (uint8_t)0;
   r.used = (fl & PG_SVR_USED) != 0;
   r.dirty = 0;
   r.seq_no = 0;
   other_chr = other_chr || (fl & PG_SVR_OTHERCHR) != 0;
   pg_sv_classify_read(ev, r, other_chr);
}

// The final UnmatchedSeq of every read of a close+far launch, for the classification of the resident batch: what the host
// obtains by replaying the flips (PG_READ_NFLIPS) with setUnmatchedSeq(ReverseComplement(..)).  One warp per read.
//   strip:   L0 = length without trailing non-alphanumerics (setUnmatchedSeq, src/pindel.cpp:142-169)
//   1 flip:  k1 leading non-ACGTN characters become trailing NULs and are stripped: out[j] = comp(s[L0-1-j]), j < L0-k1
//   2 flips: out[j] = s[k1+j] with non-ACGTN -> NUL, length len_after
__global__ void __launch_bounds__(256) pg_sv_final_seq_kernel(int n_reads, const uint8_t *__restrict__ seq_blob, const uint32_t *__restrict__ seq_off,
                                                             const PgReadOut *__restrict__ out, uint8_t *__restrict__ final_blob)
{
   const int i = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
   if (i >= n_reads) return;
   const uint32_t s0 = seq_off[i], len = seq_off[i + 1] - s0;
   const uint8_t *s = seq_blob + s0;
   uint8_t *d = final_blob + s0;
   const PgReadOut o = out[i];
   const uint32_t nflips = (o.flags >> PG_F_NFLIPS_SHIFT) & 3u;
   uint32_t L0 = 0, k1 = len;
   bool found = false;
   for (uint32_t c0 = 0; c0 < len; c0 += 32) {
      const uint32_t j = c0 + lane;
      const uint8_t ch = j < len ? s[j] : (uint8_t)0;
      const bool alnum = (ch >= '0' && ch <= '9') || (ch >= 'A' && ch <= 'Z') || (ch >= 'a' && ch <= 'z');
      const bool acgtn = ch == 'A' || ch == 'C' || ch == 'G' || ch == 'T' || ch == 'N';
      const uint32_t ma = __ballot_sync(0xFFFFFFFFu, alnum), mk = __ballot_sync(0xFFFFFFFFu, acgtn);
      if (ma) L0 = c0 + 32 - __clz(ma);
      if (!found && mk) { k1 = c0 + __ffs(mk) - 1; found = true; }
   }
   if (k1 > L0) k1 = L0;
   const uint32_t Lf = o.len_after;
   for (uint32_t j = lane; j < len; j += 32) {
      uint8_t ch = 0;
      if (j < Lf) {
         if (nflips == 0) ch = s[j];
         else if (nflips == 1) {
            const uint8_t x = s[L0 - 1 - j];
            ch = x == 'A' ? 'T' : x == 'C' ? 'G' : x == 'G' ? 'C' : x == 'T' ? 'A' : x == 'N' ? 'N' : 0;
         } else {
            const uint8_t x = s[k1 + j];
            ch = (x == 'A' || x == 'C' || x == 'G' || x == 'T' || x == 'N') ? x : 0;
         }
      }
      d[j] = ch;
   }
}

// Thread-to-read assignment of pg_sv_kernel: reads with a '+' anchor first, then the others (stable), so that the lanes of a
// warp run the same arm of every driver.  flags -> (scan) -> perm; the number of '+' reads is the scan's total.
__global__ void pg_sv_strand_flag_kernel(int n, const uint8_t *__restrict__ matched_d, uint8_t *__restrict__ flags)
{
   const int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i < n) flags[i] = matched_d[i] == '+';
}
__global__ void pg_sv_perm_kernel(int n, const uint8_t *__restrict__ flags, const uint32_t *__restrict__ offsets, const uint32_t *__restrict__ n_plus,
                                  uint32_t *__restrict__ perm)
{
   const int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= n) return;
   const uint32_t before = offsets[i];      // '+' reads before i
   perm[flags[i] ? before : *n_plus + ((uint32_t)i - before)] = (uint32_t)i;
}

// Record order.  pg_sv_kernel hands out pool slots in completion order; the host replays records per driver with reads
// ascending, i.e. sorted by (driver, read, seq).  That position is an exclusive prefix sum over the per-(driver, read)
// record counts the kernel leaves in `counts` (driver-major), so the order is restored on the device by a scan and one
// gather instead of a host sort: pg_scan_reduce / pg_scan_blocks / pg_scan_apply (16 counts per thread, one 16-byte load
// each), then pg_sv_order_kernel moves every record to offsets[driver * n + read] + seq.
#define PG_SCAN_THREADS 256
#define PG_SCAN_PER_BLOCK (PG_SCAN_THREADS * 16)

__device__ __forceinline__ uint32_t pg_sum_bytes(uint32_t w) { return __dp4a(w, 0x01010101u, 0u); }

__global__ void __launch_bounds__(PG_SCAN_THREADS) pg_scan_reduce(const uint4 *__restrict__ counts16, uint32_t n16, uint32_t *__restrict__ block_sums)
{
   __shared__ uint32_t warp_sums[PG_SCAN_THREADS / 32];
   const uint32_t i = blockIdx.x * PG_SCAN_THREADS + threadIdx.x;
   uint32_t s = 0;
   if (i < n16) {
      const uint4 v = counts16[i];
      s = pg_sum_bytes(v.x) + pg_sum_bytes(v.y) + pg_sum_bytes(v.z) + pg_sum_bytes(v.w);
   }
   s = __reduce_add_sync(0xFFFFFFFFu, s);
   if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = s;
   __syncthreads();
   if (threadIdx.x == 0) {
      uint32_t t = 0;
      for (int w = 0; w < PG_SCAN_THREADS / 32; w++) t += warp_sums[w];
      block_sums[blockIdx.x] = t;
   }
}

// exclusive scan of the block sums in place, one block; total -> block_sums[n_blocks]
__global__ void __launch_bounds__(1024) pg_scan_blocks(uint32_t *block_sums, uint32_t n_blocks)
{
   __shared__ uint32_t warp_tot[32];
   __shared__ uint32_t carry_s;
   if (threadIdx.x == 0) carry_s = 0;
   __syncthreads();
   for (uint32_t base = 0; base < n_blocks; base += 1024) {
      const uint32_t i = base + threadIdx.x;
      const uint32_t v = i < n_blocks ? block_sums[i] : 0u;
      uint32_t x = v;
      for (int d = 1; d < 32; d <<= 1) {
         const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, d);
         if ((int)(threadIdx.x & 31) >= d) x += y;
      }
      if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = x;
      __syncthreads();
      uint32_t before = carry_s;
      for (uint32_t w = 0; w < (threadIdx.x >> 5); w++) before += warp_tot[w];
      if (i < n_blocks) block_sums[i] = before + x - v;
      __syncthreads();
      if (threadIdx.x == 1023) carry_s = before + x;
      __syncthreads();
   }
   if (threadIdx.x == 0) block_sums[n_blocks] = carry_s;
}

__global__ void __launch_bounds__(PG_SCAN_THREADS) pg_scan_apply(const uint4 *__restrict__ counts16, uint32_t n16, const uint32_t *__restrict__ block_sums,
                                                                 uint32_t *__restrict__ offsets)
{
   __shared__ uint32_t warp_tot[PG_SCAN_THREADS / 32];
   const uint32_t i = blockIdx.x * PG_SCAN_THREADS + threadIdx.x;
   uint4 v = make_uint4(0, 0, 0, 0);
   if (i < n16) v = counts16[i];
   const uint32_t mine = pg_sum_bytes(v.x) + pg_sum_bytes(v.y) + pg_sum_bytes(v.z) + pg_sum_bytes(v.w);
   uint32_t x = mine;
   for (int d = 1; d < 32; d <<= 1) {
      const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, d);
      if ((int)(threadIdx.x & 31) >= d) x += y;
   }
   if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = x;
   __syncthreads();
   uint32_t run = block_sums[blockIdx.x] + x - mine;
   for (uint32_t w = 0; w < (threadIdx.x >> 5); w++) run += warp_tot[w];
   if (i >= n16) return;
   const uint32_t words[4] = {v.x, v.y, v.z, v.w};
   uint4 *dst = reinterpret_cast<uint4 *>(offsets + (size_t)i * 16);
#pragma unroll
   for (int k = 0; k < 4; k++) {
      uint4 o;
      o.x = run; run += words[k] & 0xFFu;
      o.y = run; run += (words[k] >> 8) & 0xFFu;
      o.z = run; run += (words[k] >> 16) & 0xFFu;
      o.w = run; run += words[k] >> 24;
      dst[k] = o;
   }
}

__global__ void __launch_bounds__(256) pg_sv_order_kernel(const PgSvRecord *__restrict__ pool, uint32_t n_records, const uint32_t *__restrict__ offsets,
                                                          uint32_t n_reads, PgSvRecord *__restrict__ ordered)
{
   const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
   if (t >= n_records) return;
   const PgSvRecord rec = pool[t];
   ordered[offsets[(size_t)rec.driver * n_reads + rec.read] + rec.seq] = rec;
}

// ------------------------------------------------------------------------------------------------
// Context
// ------------------------------------------------------------------------------------------------
#define PG_RES_SLOTS 3

struct DevBuf {
   void *p = nullptr;
   size_t cap = 0;
};

struct pg_ctx {
   pg_settings st;
   int device = 0;
   int n_sm = 0;
   cudaStream_t own_stream = nullptr, stream = nullptr;
   std::string err;
   PgParams prm_h;
   uint32_t table[PG_TABLE];
   PgParams *prm_d = nullptr;
   // reference
   std::vector<void *> chrom_allocs;
   std::vector<PgChrom> chroms_h;
   PgChrom *chroms_d = nullptr;
   // resident batch
   int n = 0;
   bool have_anchor = false, have_chr = false, have_close = false, have_regions = false;
   uint32_t max_read_len = 0;
   size_t blob_bytes = 0;
   DevBuf seq_blob, seq_off, md, pos, isz, chr, closeinfo, regions, region_off, scratch;
   // Results rotate through three slots, so that the records of launch k can still be read (handed to a collective that only
   // gets SM time when a persistent search kernel drains, copied out) while launches k+1 and k+2 are queued: each slot has its
   // own header array, run pool, counters and completion event.
   struct ResSlot {
      DevBuf out, pool, counters;
      uint32_t pool_cap = 0;
      uint32_t *counters_h = nullptr;  // pinned copy of the counters, filled right behind the kernel
      cudaEvent_t done = nullptr;      // kernel + counters copy of the slot's launch
      int n = 0, mode = 0;
      bool launched = false;
      uint32_t runs = 0;
      uint64_t window_bases = 0, ups = 0;
   } rs[PG_RES_SLOTS];
   int cur = 0;                        // slot of the latest launch
   int last_mode = 0;
   bool launched = false;
   bool stats_valid = false;      // last_window_bases / last_ups / last_runs describe a completed call (resident or pipelined)
   uint64_t launches = 0;
   uint32_t last_runs = 0;
   uint64_t last_window_bases = 0, last_ups = 0;
   // SV classification (pg_classify)
   DevBuf sv_seq, sv_off, sv_md, sv_isz, sv_flags, sv_runoff, sv_nclose, sv_runs, sv_records, sv_params, sv_counters, sv_counts, sv_offsets,
      sv_blocksums, sv_ordered, sv_perm;
   PgSvParams sv_params_h;          // staging copy (the async upload reads it until the stream drains)
   DevBuf sv_close_off, sv_close_runs;
   uint64_t sv_aux_launches = 0;
   // BGZF inflation (pg_bgzf_inflate)
   DevBuf bz_file, bz_blocks, bz_out, bz_status, bm_info, bm_starts, bm_counts, bm_landing, bm_base, bm_recs;
   uint64_t bm_launches = 0, bm_n_recs = 0;
   bool bm_valid = false;       // bm_recs holds the table of the resident stream
   DevBuf bd_cls, bd_off, bd_sums, bd_rows, bd_src, bd_counters;
   DevBuf sr_info, sr_keys, sr_head, sr_slot, sr_next, sr_partner, sr_emit, sr_cands, sr_seqs, sr_refs;   // pg_bam_sr_candidates
   // the last split-read filter run (sr_filter_device): valid for stream generation sr_cache_gen and query sr_cache_q
   bool sr_cache_valid = false, sr_batch_ready = false;
   uint64_t bm_generation = 0, sr_cache_gen = 0;
   pg_sr_query sr_cache_q;
   int sr_cache_rc = 0, sr_batch_chr = -1;
   uint32_t sr_tot = 0, sr_tot_ref = 0, sr_stride = 1, sr_scanned = 0;
   uint32_t bm_last_segments = 0, bm_last_redone = 0;
   float bm_last_ms = 0.f;
   size_t bz_out_len = 0;
   uint64_t bz_launches = 0;
   float bz_last_ms = 0.f;
   // BreakDancer mask of the current window (pg_bd_load_region) and the clusters derived from it for the resident batch
   DevBuf bk_start, bk_drop, bk_sel, bk_counts, bk_offsets, bk_sums, bk_regions, bk_region_off, bk_pos, bk_over;
   PgBdRegionView bk_view;
   bool bk_loaded = false;
   uint64_t bk_launches = 0, bk_lookups = 0;
   // per-box read ordering (pg_box_sort)
   DevBuf bx_off, bx_keys, bx_order, bx_non, bx_event;
   uint64_t bx_launches = 0;
   float bx_last_ms = 0.f;
   // read-pair clustering (pg_rp_modify / pg_rp_summarize)
   DevBuf rp_in, rp_out, rp_rows, rp_off, rp_scratch, rp_merged, rp_nid, rp_counters, rp_keys, rp_flip, rp_chr;
   uint64_t rp_launches = 0;
   uint32_t rp_classes_small = 0, rp_classes_big = 0, rp_rows_in_classes = 0;   // of the last pg_rp_modify
   float rp_last_ms = 0.f;    // scan + order kernels
   uint64_t sv_launches = 0;
   cudaEvent_t sv_ev0 = nullptr, sv_ev1 = nullptr;   // around the last pg_sv_kernel launch
   float sv_last_ms = 0.f;
   struct PipeState *pipe = nullptr;    // streams, events and the two chunk slots of the pipelined path (lazy)
   int pipe_chunk = PG_PIPE_CHUNK_DEFAULT;
   bool pipe_chunk_forced = false;      // PG_PIPE_CHUNK given: exactly that chunk size, chunked from two chunks on (tests)
};

struct PipeState;
static void pipe_destroy(PipeState *ps);

static thread_local std::string g_create_error;

static int fail(pg_ctx *c, int code, const std::string &msg)
{
   if (c) c->err = msg;
   return code;
}
#define CU(call)                                                                                                      \
   do {                                                                                                               \
      cudaError_t e_ = (call);                                                                                        \
      if (e_ != cudaSuccess) return fail(ctx, PG_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));       \
   } while (0)

static int ensure(pg_ctx *ctx, DevBuf &b, size_t bytes)
{
   if (bytes <= b.cap) return PG_OK;
   if (b.p) CU(cudaFree(b.p));
   b.p = nullptr; b.cap = 0;
   size_t want = bytes + bytes / 4 + 256;
   CU(cudaMalloc(&b.p, want));
   b.cap = want;
   return PG_OK;
}

template <int LW, int TL>
static int launch_variant(pg_ctx *ctx, const PgBatch &bt_in, cudaStream_t stream, DevBuf &scratch)
{
   PgBatch bt = bt_in;
   size_t smem = 0;           // the warp's working set is static shared memory
   int per_sm = 0;
   CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pg_search_kernel<LW, TL>, PG_BLOCK_THREADS, smem));
   if (per_sm < 1) per_sm = 1;
   int grid = ctx->n_sm * per_sm;                       // a whole number of waves: one resident CTA set per SM
   long need_blocks = ((long)bt.n_reads + PG_WARPS_PER_BLOCK - 1) / PG_WARPS_PER_BLOCK;
   if (need_blocks < grid) grid = (int)(need_blocks > 0 ? need_blocks : 1);
   bt.scratch_cap = 32u * LW;
   int rc = ensure(ctx, scratch, sizeof(PgRun) * 3u * bt.scratch_cap * (size_t)grid * PG_WARPS_PER_BLOCK);
   if (rc != PG_OK) return rc;
   bt.scratch = (PgRun *)scratch.p;
   pg_search_kernel<LW, TL><<<grid, PG_BLOCK_THREADS, smem, stream>>>(bt, ctx->prm_h, ctx->chroms_d);
   CU(cudaGetLastError());
   ctx->launches++;
   return PG_OK;
}

// Picks the kernel variant: LW = 32-base words per read plane, TL = mismatch levels tracked per seed -- the smallest
// compiled value that covers the batch (exact for the common single-read-length batch, so the per-level loops carry no
// dead levels).
static int dispatch_variant(pg_ctx *ctx, const PgBatch &bt, uint32_t max_read_len, cudaStream_t stream, DevBuf &scratch)
{
   int need_t = 0;
   for (uint32_t l = 0; l <= max_read_len && l < 512; l++) {
      int t = (int)ctx->prm_h.g[l] + ctx->prm_h.add_mm + 1;
      if (t > need_t) need_t = t;
   }
   if (need_t > 32) return fail(ctx, PG_ERR_UNSUPPORTED, "more than 32 mismatch levels (g_maxMismatch + -a + 1)");
   if (max_read_len <= 256) {
      if (need_t <= 5) return launch_variant<8, 5>(ctx, bt, stream, scratch);
      if (need_t <= 6) return launch_variant<8, 6>(ctx, bt, stream, scratch);
      if (need_t <= 7) return launch_variant<8, 7>(ctx, bt, stream, scratch);
      if (need_t <= 8) return launch_variant<8, 8>(ctx, bt, stream, scratch);
      if (need_t <= 12) return launch_variant<8, 12>(ctx, bt, stream, scratch);
      if (need_t <= 16) return launch_variant<8, 16>(ctx, bt, stream, scratch);
   }
   if (need_t <= 8) return launch_variant<16, 8>(ctx, bt, stream, scratch);
   if (need_t <= 16) return launch_variant<16, 16>(ctx, bt, stream, scratch);
   return launch_variant<16, 32>(ctx, bt, stream, scratch);
}

extern "C" {

void pg_default_settings(pg_settings *s)
{
   memset(s, 0, sizeof(*s));
   s->seq_error_rate = 0.01;
   s->sensitivity = 0.95;
   s->max_mismatch_rate = 0.02;
   s->additional_mismatch = 1;
   s->min_close = 8;
   s->min_perfect_match = 3;
   s->max_range_index = 2;
   s->device = 0;
}

const char *pg_create_error(void) { return g_create_error.c_str(); }
const char *pg_last_error(pg_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

pg_ctx *pg_create(const pg_settings *s)
{
   g_create_error.clear();
   if (!s) { g_create_error = "null settings"; return nullptr; }
   if (s->additional_mismatch < 0 || s->min_close < 1 || s->min_perfect_match < 0 || s->max_range_index < 0 ||
       s->max_range_index > 9) {
      g_create_error = "settings out of range";
      return nullptr;
   }
   int ndev = 0;
   cudaError_t e = cudaGetDeviceCount(&ndev);
   if (e != cudaSuccess || ndev == 0) {
      g_create_error = std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                       " (libpindel_pg has no CPU path)";
      return nullptr;
   }
   if (s->device < 0 || s->device >= ndev) { g_create_error = "device ordinal out of range"; return nullptr; }
   pg_ctx *ctx = new pg_ctx();
   ctx->st = *s;
   ctx->device = s->device;
   pg_prob_table(0.001 + s->seq_error_rate, s->sensitivity, ctx->table);      // src/pindel.cpp:856
   if (pg_fill_params(&ctx->prm_h, s->additional_mismatch, s->min_close, s->min_perfect_match, s->max_range_index,
                      s->max_mismatch_rate, ctx->table) != 0) {
      g_create_error = "mismatch table does not fit";
      delete ctx;
      return nullptr;
   }
   if (const char *e = getenv("PG_FILTER_MIN_POS")) ctx->prm_h.filter_min_pos = (uint32_t)strtoul(e, nullptr, 10);   // tuning only
   bool ok = cudaSetDevice(ctx->device) == cudaSuccess &&
             cudaDeviceGetAttribute(&ctx->n_sm, cudaDevAttrMultiProcessorCount, ctx->device) == cudaSuccess &&
             cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) == cudaSuccess &&
             cudaMalloc(&ctx->prm_d, sizeof(PgParams)) == cudaSuccess &&
             cudaMemcpy(ctx->prm_d, &ctx->prm_h, sizeof(PgParams), cudaMemcpyHostToDevice) == cudaSuccess;
   for (int i = 0; ok && i < PG_RES_SLOTS; i++)
      ok = cudaHostAlloc((void **)&ctx->rs[i].counters_h, 32, cudaHostAllocDefault) == cudaSuccess &&
           cudaEventCreateWithFlags(&ctx->rs[i].done, cudaEventDisableTiming) == cudaSuccess;
   if (!ok) {
      g_create_error = std::string("CUDA initialisation failed: ") + cudaGetErrorString(cudaGetLastError());
      delete ctx;
      return nullptr;
   }
   ctx->stream = ctx->own_stream;
   if (const char *e = getenv("PG_PIPE_CHUNK")) { int v = atoi(e); if (v >= 256) { ctx->pipe_chunk = v; ctx->pipe_chunk_forced = true; } }
   return ctx;
}

void pg_destroy(pg_ctx *ctx)
{
   if (!ctx) return;
   cudaSetDevice(ctx->device);
   cudaDeviceSynchronize();
   for (void *p : ctx->chrom_allocs) cudaFree(p);
   DevBuf *bufs[] = {&ctx->seq_blob, &ctx->seq_off, &ctx->md, &ctx->pos, &ctx->isz, &ctx->chr, &ctx->closeinfo,
                     &ctx->regions, &ctx->region_off, &ctx->scratch};
   for (DevBuf *b : bufs) if (b->p) cudaFree(b->p);
   for (int i = 0; i < PG_RES_SLOTS; i++)
      for (DevBuf *b : {&ctx->rs[i].out, &ctx->rs[i].pool, &ctx->rs[i].counters}) if (b->p) cudaFree(b->p);
   DevBuf *svbufs[] = {&ctx->sv_seq, &ctx->sv_off, &ctx->sv_md, &ctx->sv_isz, &ctx->sv_flags, &ctx->sv_runoff, &ctx->sv_nclose,
                       &ctx->sv_runs, &ctx->sv_records, &ctx->sv_params, &ctx->sv_counters,
                       &ctx->sv_counts, &ctx->sv_offsets, &ctx->sv_blocksums, &ctx->sv_ordered, &ctx->sv_perm, &ctx->sv_close_off,
                       &ctx->sv_close_runs, &ctx->rp_in, &ctx->rp_out, &ctx->rp_rows, &ctx->rp_off, &ctx->rp_scratch, &ctx->rp_merged,
                       &ctx->rp_nid, &ctx->rp_counters, &ctx->rp_keys, &ctx->rp_flip, &ctx->rp_chr, &ctx->bz_file, &ctx->bz_blocks, &ctx->bz_out, &ctx->bz_status,
                       &ctx->bm_info, &ctx->bm_starts, &ctx->bm_counts, &ctx->bm_landing, &ctx->bm_base, &ctx->bm_recs,
                       &ctx->bd_cls, &ctx->bd_off, &ctx->bd_sums, &ctx->bd_rows, &ctx->bd_src, &ctx->bd_counters,
                       &ctx->sr_info, &ctx->sr_keys, &ctx->sr_head, &ctx->sr_slot, &ctx->sr_next, &ctx->sr_partner, &ctx->sr_emit, &ctx->sr_cands,
                       &ctx->sr_seqs, &ctx->sr_refs,
                       &ctx->bx_off, &ctx->bx_keys, &ctx->bx_order, &ctx->bx_non,
                       &ctx->bk_start, &ctx->bk_drop, &ctx->bk_sel, &ctx->bk_counts, &ctx->bk_offsets, &ctx->bk_sums, &ctx->bk_regions, &ctx->bk_region_off,
                       &ctx->bk_pos, &ctx->bk_over, &ctx->bx_event};
   for (DevBuf *b : svbufs) if (b->p) cudaFree(b->p);
   pipe_destroy(ctx->pipe);
   if (ctx->chroms_d) cudaFree(ctx->chroms_d);
   if (ctx->prm_d) cudaFree(ctx->prm_d);
   for (int i = 0; i < PG_RES_SLOTS; i++) {
      if (ctx->rs[i].counters_h) cudaFreeHost(ctx->rs[i].counters_h);
      if (ctx->rs[i].done) cudaEventDestroy(ctx->rs[i].done);
   }
   if (ctx->sv_ev0) { cudaEventDestroy(ctx->sv_ev0); cudaEventDestroy(ctx->sv_ev1); }
   if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
   delete ctx;
}

int pg_set_stream(pg_ctx *ctx, void *cuda_stream)
{
   if (!ctx) return PG_ERR_ARG;
   CU(cudaSetDevice(ctx->device));
   CU(cudaStreamSynchronize(ctx->stream));
   ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
   return PG_OK;
}

int pg_max_mismatch_table(pg_ctx *ctx, uint32_t *out)
{
   if (!ctx || !out) return PG_ERR_ARG;
   memcpy(out, ctx->table, sizeof(uint32_t) * PG_TABLE);
   return PG_OK;
}

int pg_set_reference(pg_ctx *ctx, int32_t n_chr, const char *const *seq, const uint32_t *len)
{
   if (!ctx || n_chr <= 0 || n_chr > 32767 || !seq || !len) return fail(ctx, PG_ERR_ARG, "pg_set_reference: bad arguments");
   CU(cudaSetDevice(ctx->device));
   CU(cudaStreamSynchronize(ctx->stream));
   for (void *p : ctx->chrom_allocs) cudaFree(p);
   ctx->chrom_allocs.clear();
   ctx->chroms_h.clear();
   if (ctx->chroms_d) { cudaFree(ctx->chroms_d); ctx->chroms_d = nullptr; }
   const size_t CHUNK = (size_t)64 << 20;        // bases per staging chunk (multiple of 32)
   uint8_t *stage = nullptr;
   // whatever way this function is left: the staging buffer goes, and a reference that was not completed leaves no
   // half-built chromosome list behind (the context then has no reference, as before the first call)
   struct Cleanup {
      pg_ctx *ctx; uint8_t *&stage; bool done;
      ~Cleanup()
      {
         if (stage) cudaFree(stage);
         if (done) return;
         for (void *p : ctx->chrom_allocs) cudaFree(p);
         ctx->chrom_allocs.clear();
         ctx->chroms_h.clear();
         if (ctx->chroms_d) { cudaFree(ctx->chroms_d); ctx->chroms_d = nullptr; }
      }
   } cleanup{ctx, stage, false};
   size_t maxlen = 0;
   for (int i = 0; i < n_chr; i++) if (len[i] > maxlen) maxlen = len[i];
   CU(cudaMalloc((void **)&stage, maxlen < CHUNK ? (maxlen ? maxlen : 1) : CHUNK));
   for (int i = 0; i < n_chr; i++) {
      uint32_t nwords = (len[i] + 31u) / 32u;
      size_t total = (size_t)nwords + 2u * PG_PADW;
      uint2 *planes = nullptr;
      uint32_t *nplane = nullptr;
      CU(cudaMalloc((void **)&planes, total * sizeof(uint2)));
      ctx->chrom_allocs.push_back(planes);
      CU(cudaMalloc((void **)&nplane, total * sizeof(uint32_t)));
      ctx->chrom_allocs.push_back(nplane);
      CU(cudaMemsetAsync(planes, 0, total * sizeof(uint2), ctx->stream));
      CU(cudaMemsetAsync(nplane, 0xff, total * sizeof(uint32_t), ctx->stream));   // guard band reads as N
      for (size_t b0 = 0; b0 < len[i]; b0 += CHUNK) {
         size_t nb = len[i] - b0 < CHUNK ? len[i] - b0 : CHUNK;
         CU(cudaMemcpyAsync(stage, seq[i] + b0, nb, cudaMemcpyHostToDevice, ctx->stream));
         uint32_t nw = (uint32_t)((nb + 31) / 32);
         uint32_t threads = 256, blocks = (nw * 32u + threads - 1) / threads;
         pg_pack_kernel<<<blocks, threads, 0, ctx->stream>>>(stage, (uint32_t)nb, (uint32_t)(PG_PADW + b0 / 32), planes, nplane);
         CU(cudaGetLastError());
         CU(cudaStreamSynchronize(ctx->stream));   // staging buffer and (possibly pageable) source are reused
      }
      PgChrom c;
      c.planes = planes + PG_PADW;
      c.nplane = nplane + PG_PADW;
      c.len = len[i];
      c.nwords = nwords;
      ctx->chroms_h.push_back(c);
   }
   CU(cudaMalloc((void **)&ctx->chroms_d, sizeof(PgChrom) * (size_t)n_chr));
   CU(cudaMemcpy(ctx->chroms_d, ctx->chroms_h.data(), sizeof(PgChrom) * (size_t)n_chr, cudaMemcpyHostToDevice));
   cleanup.done = true;
   return PG_OK;
}

int pg_upload(pg_ctx *ctx, const pg_reads *in)
{
   if (!ctx || !in || in->n < 0 || (in->n > 0 && (!in->seq_blob || !in->seq_off)))
      return fail(ctx, PG_ERR_ARG, "pg_upload: bad arguments");
   CU(cudaSetDevice(ctx->device));
   const int n = in->n;
   ctx->launched = false;
   ctx->stats_valid = false;
   ctx->sr_batch_ready = false;
   ctx->n = n;
   ctx->have_anchor = in->matched_d && in->matched_rel_pos && in->insert_size;
   ctx->have_chr = in->chr_id != nullptr;
   ctx->have_close = in->close_info != nullptr;
   ctx->have_regions = in->regions && in->region_off;
   if (n == 0) return PG_OK;
   // longest read and chromosome ids are validated here, on the host
   uint32_t maxlen = 0;
   for (int i = 0; i < n; i++) {
      if (in->seq_off[i + 1] < in->seq_off[i]) return fail(ctx, PG_ERR_ARG, "pg_upload: seq_off must be non-decreasing");
      uint32_t l = in->seq_off[i + 1] - in->seq_off[i];
      if (l > maxlen) maxlen = l;
   }
   ctx->max_read_len = maxlen;
   if (maxlen >= PG_TABLE)
      return fail(ctx, PG_ERR_UNSUPPORTED, "read longer than 499 bases (the reference's g_maxMismatch table has 500 entries)");
   if (ctx->have_chr) {
      for (int i = 0; i < n; i++)
         if (in->chr_id[i] < 0 || (size_t)in->chr_id[i] >= ctx->chroms_h.size())
            return fail(ctx, PG_ERR_ARG, "pg_upload: chr_id out of range");
   }
   size_t blob = in->seq_off[n];
   ctx->blob_bytes = blob;
   int rc;
#define UP(buf, src, bytes)                                                                              \
   if ((rc = ensure(ctx, buf, (bytes))) != PG_OK) return rc;                                             \
   CU(cudaMemcpyAsync(buf.p, (src), (bytes), cudaMemcpyHostToDevice, ctx->stream))
   UP(ctx->seq_blob, in->seq_blob, blob ? blob : 1);
   UP(ctx->seq_off, in->seq_off, sizeof(uint32_t) * ((size_t)n + 1));
   if (ctx->have_anchor) {
      UP(ctx->md, in->matched_d, (size_t)n);
      UP(ctx->pos, in->matched_rel_pos, sizeof(uint32_t) * (size_t)n);
      UP(ctx->isz, in->insert_size, sizeof(int16_t) * (size_t)n);
   }
   if (ctx->have_chr) { UP(ctx->chr, in->chr_id, sizeof(int32_t) * (size_t)n); }
   if (ctx->have_close) { UP(ctx->closeinfo, in->close_info, sizeof(pg_close_info) * (size_t)n); }
   if (ctx->have_regions) {
      size_t nr = in->region_off[n];
      for (size_t i = 0; i < nr; i++)
         if (in->regions[i].chr < 0 || (size_t)in->regions[i].chr >= ctx->chroms_h.size())
            return fail(ctx, PG_ERR_ARG, "pg_upload: region chromosome out of range");
      UP(ctx->regions, in->regions, sizeof(pg_region) * (nr ? nr : 1));
      UP(ctx->region_off, in->region_off, sizeof(uint32_t) * ((size_t)n + 1));
   }
#undef UP
   return PG_OK;
}

// The clusters of n reads from the loaded BreakDancer region, on the device: count pass, scan, fill pass.  Close-end positions
// from ci (the resident batch's close_info) or from pos.  Leaves regions / region_off in ctx->bk_regions / bk_region_off and the
// total in *total_h (after a stream synchronisation when total_h is given).
static int bd_lookup_device(pg_ctx *ctx, uint32_t n, const PgCloseInfo *ci, const uint32_t *pos, uint32_t *total_h)
{
   cudaStream_t st = ctx->stream;
   const uint32_t n16 = (n + 15u) / 16u, n_blocks = (n16 + PG_SCAN_THREADS - 1) / PG_SCAN_THREADS;
   int rc;
   if ((rc = ensure(ctx, ctx->bk_counts, (size_t)n16 * 16)) != PG_OK) return rc;
   if ((rc = ensure(ctx, ctx->bk_offsets, sizeof(uint32_t) * (size_t)n16 * 16)) != PG_OK) return rc;
   if ((rc = ensure(ctx, ctx->bk_sums, sizeof(uint32_t) * ((size_t)n_blocks + 1))) != PG_OK) return rc;
   if ((rc = ensure(ctx, ctx->bk_region_off, sizeof(uint32_t) * ((size_t)n + 1))) != PG_OK) return rc;
   if ((rc = ensure(ctx, ctx->bk_over, sizeof(uint32_t))) != PG_OK) return rc;
   CU(cudaMemsetAsync(ctx->bk_counts.p, 0, (size_t)n16 * 16, st));
   CU(cudaMemsetAsync(ctx->bk_over.p, 0, sizeof(uint32_t), st));
   const uint32_t grid = (n + 255u) / 256u;
   pg_bd_count_kernel<<<grid, 256, 0, st>>>(ctx->bk_view, n, ci, pos, (uint8_t *)ctx->bk_counts.p, (uint32_t *)ctx->bk_over.p);
   pg_scan_reduce<<<n_blocks, PG_SCAN_THREADS, 0, st>>>((const uint4 *)ctx->bk_counts.p, n16, (uint32_t *)ctx->bk_sums.p);
   pg_scan_blocks<<<1, 1024, 0, st>>>((uint32_t *)ctx->bk_sums.p, n_blocks);
   pg_scan_apply<<<n_blocks, PG_SCAN_THREADS, 0, st>>>((const uint4 *)ctx->bk_counts.p, n16, (const uint32_t *)ctx->bk_sums.p, (uint32_t *)ctx->bk_offsets.p);
   CU(cudaGetLastError());
   uint32_t tot = 0, over = 0;
   CU(cudaMemcpyAsync(&tot, (const uint32_t *)ctx->bk_sums.p + n_blocks, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
   CU(cudaMemcpyAsync(&over, ctx->bk_over.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
   CU(cudaStreamSynchronize(st));
   if (over) return fail(ctx, PG_ERR_UNSUPPORTED, "BreakDancer cluster of more than 255 windows for one read");
   if ((rc = ensure(ctx, ctx->bk_regions, sizeof(PgRegion) * ((size_t)tot + 1))) != PG_OK) return rc;
   pg_bd_fill_kernel<<<(n + 256u) / 256u, 256, 0, st>>>(ctx->bk_view, n, ci, pos, (const uint32_t *)ctx->bk_offsets.p, (const uint8_t *)ctx->bk_counts.p,
                                                       (PgRegion *)ctx->bk_regions.p, (uint32_t *)ctx->bk_region_off.p, (const uint32_t *)ctx->bk_sums.p + n_blocks);
   CU(cudaGetLastError());
   ctx->bk_launches += 5;
   ctx->bk_lookups += n;
   if (total_h) *total_h = tot;
   return PG_OK;
}

// The search kernel over the resident batch into result slot `slot`; the slot's counters follow the kernel to the host
// and its event marks both, so the slot can be waited for on its own while later work is already queued on the stream.
static int launch_into(pg_ctx *ctx, int32_t mode, int slot)
{
   pg_ctx::ResSlot &r = ctx->rs[slot];
   const int n = ctx->n;
   r.n = n; r.mode = mode; r.launched = true;
   if (n == 0) return PG_OK;
   int rc;
   if ((rc = ensure(ctx, r.out, sizeof(PgReadOut) * (size_t)n)) != PG_OK) return rc;
   if ((rc = ensure(ctx, r.counters, 32)) != PG_OK) return rc;
   if (r.pool_cap == 0 || r.pool.p == nullptr) {
      size_t want = (size_t)n * 24u + 4096u;
      if (want > 0xfffffff0u) want = 0xfffffff0u;
      if ((rc = ensure(ctx, r.pool, sizeof(PgRun) * want)) != PG_OK) return rc;
      r.pool_cap = (uint32_t)(r.pool.cap / sizeof(PgRun));
   }
   CU(cudaMemsetAsync(r.counters.p, 0, 32, ctx->stream));
   PgBatch bt;
   memset(&bt, 0, sizeof(bt));
   bt.n_reads = n;
   bt.mode = mode;
   bt.seq_blob = (const uint8_t *)ctx->seq_blob.p;
   bt.seq_off = (const uint32_t *)ctx->seq_off.p;
   bt.matched_d = ctx->have_anchor ? (const uint8_t *)ctx->md.p : nullptr;
   bt.matched_rel_pos = ctx->have_anchor ? (const uint32_t *)ctx->pos.p : nullptr;
   bt.insert_size = ctx->have_anchor ? (const int16_t *)ctx->isz.p : nullptr;
   bt.chr_id = ctx->have_chr ? (const int32_t *)ctx->chr.p : nullptr;
   bt.close_info = (mode == 2 && ctx->have_close) ? (const PgCloseInfo *)ctx->closeinfo.p : nullptr;
   bt.regions = ctx->have_regions ? (const PgRegion *)ctx->regions.p : nullptr;
   bt.region_off = ctx->have_regions ? (const uint32_t *)ctx->region_off.p : nullptr;
   if (mode == 2 && !ctx->have_regions && ctx->bk_loaded && ctx->have_close) {
      // SearchFarEnd's first stop (src/pindel.cpp:1006): the read's BreakDancer cluster, looked up on the device from the close ends
      if ((rc = bd_lookup_device(ctx, (uint32_t)n, (const PgCloseInfo *)ctx->closeinfo.p, nullptr, nullptr)) != PG_OK) return rc;
      bt.regions = (const PgRegion *)ctx->bk_regions.p;
      bt.region_off = (const uint32_t *)ctx->bk_region_off.p;
   }
   bt.out = (PgReadOut *)r.out.p;
   bt.run_pool = (PgRun *)r.pool.p;
   bt.run_pool_cap = r.pool_cap;
   bt.counters = (uint32_t *)r.counters.p;
   rc = dispatch_variant(ctx, bt, ctx->max_read_len, ctx->stream, ctx->scratch);
   if (rc != PG_OK) return rc;
   CU(cudaMemcpyAsync(r.counters_h, r.counters.p, 32, cudaMemcpyDeviceToHost, ctx->stream));
   CU(cudaEventRecord(r.done, ctx->stream));
   return PG_OK;
}

int pg_launch(pg_ctx *ctx, int32_t mode)
{
   if (!ctx || mode < 1 || mode > 3) return fail(ctx, PG_ERR_ARG, "pg_launch: mode must be 1, 2 or 3");
   if (ctx->chroms_h.empty()) return fail(ctx, PG_ERR_STATE, "pg_launch: no reference (call pg_set_reference)");
   if ((mode & 1) && !ctx->have_anchor) return fail(ctx, PG_ERR_ARG, "close-end search needs matched_d, matched_rel_pos and insert_size");
   if (mode == 2 && !ctx->have_close) return fail(ctx, PG_ERR_ARG, "far-end-only search needs close_info");
   CU(cudaSetDevice(ctx->device));
   ctx->last_mode = mode;
   ctx->launched = true;
   ctx->cur = (ctx->cur + 1) % PG_RES_SLOTS;
   return launch_into(ctx, mode, ctx->cur);
}

int pg_sync(pg_ctx *ctx)
{
   if (!ctx) return PG_ERR_ARG;
   CU(cudaSetDevice(ctx->device));
   CU(cudaStreamSynchronize(ctx->stream));
   return PG_OK;
}

// Waits for the launch of a slot (its event, not the whole stream); if the run pool was too small, grows it to the exact size
// and runs the slot again.
static int finish_slot(pg_ctx *ctx, int slot)
{
   pg_ctx::ResSlot &r = ctx->rs[slot];
   if (!r.launched) return fail(ctx, PG_ERR_STATE, "no launch to wait for");
   if (r.n == 0) { r.runs = 0; r.window_bases = r.ups = 0; return PG_OK; }
   for (int attempt = 0; attempt < 3; attempt++) {
      CU(cudaEventSynchronize(r.done));
      if (r.counters_h[2] == 0) {
         r.runs = r.counters_h[1];
         memcpy(&r.window_bases, r.counters_h + 4, 8);
         memcpy(&r.ups, r.counters_h + 6, 8);
         return PG_OK;
      }
      if (r.n != ctx->n) return fail(ctx, PG_ERR_STATE, "run pool overflow in a launch whose batch is no longer resident");
      uint32_t need = r.counters_h[1];
      CU(cudaStreamSynchronize(ctx->stream));
      CU(cudaFree(r.pool.p));
      r.pool.p = nullptr; r.pool.cap = 0; r.pool_cap = 0;
      int rc = ensure(ctx, r.pool, sizeof(PgRun) * ((size_t)need + 1024u));
      if (rc != PG_OK) return rc;
      r.pool_cap = (uint32_t)(r.pool.cap / sizeof(PgRun));
      rc = launch_into(ctx, r.mode, slot);
      if (rc != PG_OK) return rc;
   }
   return fail(ctx, PG_ERR_CUDA, "run pool kept overflowing");
}

static int finish_launch(pg_ctx *ctx)
{
   if (!ctx->launched) return fail(ctx, PG_ERR_STATE, "no launch to wait for");
   int rc = finish_slot(ctx, ctx->cur);
   if (rc != PG_OK) return rc;
   const pg_ctx::ResSlot &r = ctx->rs[ctx->cur];
   ctx->last_runs = r.runs; ctx->last_window_bases = r.window_bases; ctx->last_ups = r.ups;
   ctx->stats_valid = true;
   return PG_OK;
}

int pg_download(pg_ctx *ctx, pg_results *res)
{
   if (!ctx || !res) return PG_ERR_ARG;
   CU(cudaSetDevice(ctx->device));
   int rc = finish_launch(ctx);
   if (rc != PG_OK) return rc;
   res->n_runs = ctx->last_runs;
   if (ctx->n == 0) return PG_OK;
   if (!res->out) return fail(ctx, PG_ERR_ARG, "pg_download: out is NULL");
   if (res->n_runs > res->runs_cap || (res->n_runs && !res->runs))
      return fail(ctx, PG_ERR_RUNS_CAP, "run buffer too small");
   CU(cudaMemcpyAsync(res->out, ctx->rs[ctx->cur].out.p, sizeof(PgReadOut) * (size_t)ctx->n, cudaMemcpyDeviceToHost, ctx->stream));
   if (res->n_runs)
      CU(cudaMemcpyAsync(res->runs, ctx->rs[ctx->cur].pool.p, sizeof(PgRun) * (size_t)res->n_runs, cudaMemcpyDeviceToHost, ctx->stream));
   CU(cudaStreamSynchronize(ctx->stream));
   return PG_OK;
}


} // extern "C"

// ------------------------------------------------------------------------------------------------
// Large batches: the batch is uploaded in chunks of ctx->pipe_chunk reads on a copy stream while the search kernel of the
// previous chunk runs on a second stream and (pg_search & co.) the records of the chunk before that travel to the caller's
// host buffers on a third.  Every chunk's kernel works on its range of the ONE resident batch and allocates its runs from
// the ONE run pool of the current result slot, so a chunk's runs are a contiguous stretch of the pool, run offsets are
// global, and when the call returns the whole batch and its results are resident exactly as after pg_upload + pg_launch
// (pg_device_results, pg_classify_resident work on them).
// ------------------------------------------------------------------------------------------------
#define PG_PIPE_MAX_CHUNKS 4096

struct PipeState {
   cudaStream_t s_in = nullptr, s_k = nullptr, s_out = nullptr;
   std::vector<cudaEvent_t> in_done, k_done;    // per chunk
   cudaEvent_t out_done = nullptr;
   uint32_t *counters_h = nullptr;               // pinned, 8 words per chunk
   int counters_cap = 0;
   bool busy = false;                            // work of a call may still be queued (only after an error return)
};

static void pipe_destroy(PipeState *ps)
{
   if (!ps) return;
   for (cudaEvent_t e : ps->in_done) cudaEventDestroy(e);
   for (cudaEvent_t e : ps->k_done) cudaEventDestroy(e);
   if (ps->out_done) cudaEventDestroy(ps->out_done);
   if (ps->counters_h) cudaFreeHost(ps->counters_h);
   if (ps->s_in) cudaStreamDestroy(ps->s_in);
   if (ps->s_k) cudaStreamDestroy(ps->s_k);
   if (ps->s_out) cudaStreamDestroy(ps->s_out);
   delete ps;
}

static int pipe_init(pg_ctx *ctx, PipeState *&ps, int nchunks)
{
   if (!ps) ps = new PipeState();
   if (!ps->s_in) CU(cudaStreamCreateWithFlags(&ps->s_in, cudaStreamNonBlocking));
   if (!ps->s_k) CU(cudaStreamCreateWithFlags(&ps->s_k, cudaStreamNonBlocking));
   if (!ps->s_out) CU(cudaStreamCreateWithFlags(&ps->s_out, cudaStreamNonBlocking));
   if (!ps->out_done) CU(cudaEventCreateWithFlags(&ps->out_done, cudaEventDisableTiming));
   while ((int)ps->in_done.size() < nchunks) {
      cudaEvent_t e = nullptr;
      CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      ps->in_done.push_back(e);
   }
   while ((int)ps->k_done.size() < nchunks) {
      cudaEvent_t e = nullptr;
      CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      ps->k_done.push_back(e);
   }
   if (ps->counters_cap < nchunks) {
      if (ps->counters_h) { cudaFreeHost(ps->counters_h); ps->counters_h = nullptr; ps->counters_cap = 0; }
      CU(cudaHostAlloc((void **)&ps->counters_h, 32 * (size_t)nchunks, cudaHostAllocDefault));
      ps->counters_cap = nchunks;
   }
   return PG_OK;
}

// res == nullptr: nothing is copied back (pg_search_resident); the call then returns as soon as everything is queued.
static int run_pipelined(pg_ctx *ctx, PipeState *&ps, const pg_reads *in, pg_results *res, int mode)
{
   const int n = in->n;
   const bool have_anchor = in->matched_d && in->matched_rel_pos && in->insert_size;
   const bool have_regions = in->regions && in->region_off;
   if ((mode & 1) && !have_anchor) return fail(ctx, PG_ERR_ARG, "close-end search needs matched_d, matched_rel_pos and insert_size");
   if (mode == 2 && !in->close_info) return fail(ctx, PG_ERR_ARG, "far-end-only search needs close_info");
   if (ctx->chroms_h.empty()) return fail(ctx, PG_ERR_STATE, "no reference (call pg_set_reference)");
   if (res && !res->out) return fail(ctx, PG_ERR_ARG, "results buffers missing");
   // (the per-read checks of pg_upload run chunk by chunk below, behind the copies and kernels of the chunks before)
   uint32_t maxlen = 0;
   int chunk = ctx->pipe_chunk;
   // batches below four full chunks are cut in four, so that their copies hide behind their own kernels too
   if (!ctx->pipe_chunk_forced && n < 4 * chunk) chunk = ((n + 3) / 4 + 1023) & ~1023;
   if ((n + chunk - 1) / chunk > PG_PIPE_MAX_CHUNKS) chunk = (n + PG_PIPE_MAX_CHUNKS - 1) / PG_PIPE_MAX_CHUNKS;
   const int nchunks = (n + chunk - 1) / chunk;
   int rc = pipe_init(ctx, ps, nchunks);
   if (rc != PG_OK) return rc;
   // Whatever way this function is left, the three streams are drained before the next call touches the buffers again.
   struct PipeIdle {
      PipeState *ps;
      void settle() const
      {
         if (!ps->busy) return;
         cudaStreamSynchronize(ps->s_in); cudaStreamSynchronize(ps->s_k); cudaStreamSynchronize(ps->s_out);
         ps->busy = false;
      }
      ~PipeIdle() { settle(); }
   } idle{ps};
   idle.settle();
   CU(cudaStreamSynchronize(ctx->stream));         // everything queued by the caller on its stream comes first
   // the batch becomes the resident batch
   ctx->sr_batch_ready = false;
   ctx->launched = false;
   ctx->stats_valid = false;
   ctx->n = n;
   ctx->have_anchor = have_anchor;
   ctx->have_chr = in->chr_id != nullptr;
   ctx->have_close = in->close_info != nullptr;
   ctx->have_regions = have_regions;
   ctx->max_read_len = maxlen;
   ctx->blob_bytes = in->seq_off[n];
   const size_t n_regions = have_regions ? in->region_off[n] : 0;
   if ((rc = ensure(ctx, ctx->seq_blob, ctx->blob_bytes ? ctx->blob_bytes : 1)) != PG_OK) return rc;
   if ((rc = ensure(ctx, ctx->seq_off, sizeof(uint32_t) * ((size_t)n + 1))) != PG_OK) return rc;
   if (have_anchor) {
      if ((rc = ensure(ctx, ctx->md, (size_t)n)) != PG_OK) return rc;
      if ((rc = ensure(ctx, ctx->pos, sizeof(uint32_t) * (size_t)n)) != PG_OK) return rc;
      if ((rc = ensure(ctx, ctx->isz, sizeof(int16_t) * (size_t)n)) != PG_OK) return rc;
   }
   if (in->chr_id && (rc = ensure(ctx, ctx->chr, sizeof(int32_t) * (size_t)n)) != PG_OK) return rc;
   if (in->close_info && (rc = ensure(ctx, ctx->closeinfo, sizeof(pg_close_info) * (size_t)n)) != PG_OK) return rc;
   if (have_regions) {
      if ((rc = ensure(ctx, ctx->regions, sizeof(pg_region) * (n_regions ? n_regions : 1))) != PG_OK) return rc;
      if ((rc = ensure(ctx, ctx->region_off, sizeof(uint32_t) * ((size_t)n + 1))) != PG_OK) return rc;
   }
   ctx->cur = (ctx->cur + 1) % PG_RES_SLOTS;
   pg_ctx::ResSlot &r = ctx->rs[ctx->cur];
   r.n = n; r.mode = mode; r.launched = false;
   if ((rc = ensure(ctx, r.out, sizeof(PgReadOut) * (size_t)n)) != PG_OK) return rc;
   if ((rc = ensure(ctx, r.counters, 32)) != PG_OK) return rc;
   if (r.pool_cap == 0 || r.pool.p == nullptr) {
      size_t want = (size_t)n * 24u + 4096u;
      if (want > 0xfffffff0u) want = 0xfffffff0u;
      if ((rc = ensure(ctx, r.pool, sizeof(PgRun) * want)) != PG_OK) return rc;
      r.pool_cap = (uint32_t)(r.pool.cap / sizeof(PgRun));
   }
   ps->busy = true;
   CU(cudaMemsetAsync(r.counters.p, 0, 32, ps->s_k));
   uint32_t runs_before = 0;      // runs allocated by the chunks already drained
   bool cap_exceeded = false, overflow = false;
   auto drain = [&](int c) -> int {                 // chunk c's kernel: wait for it, send its records to the caller
      CU(cudaEventSynchronize(ps->k_done[c]));
      const uint32_t *cn = ps->counters_h + 8 * (size_t)c;
      if (cn[2] != 0) { overflow = true; return PG_OK; }
      const uint32_t runs_now = cn[1];
      const int c0 = c * chunk, c1 = (c0 + chunk < n) ? c0 + chunk : n;
      if (res && !overflow) {
         CU(cudaStreamWaitEvent(ps->s_out, ps->k_done[c], 0));
         CU(cudaMemcpyAsync(res->out + c0, (const PgReadOut *)r.out.p + c0, sizeof(PgReadOut) * (size_t)(c1 - c0), cudaMemcpyDeviceToHost, ps->s_out));
         if (runs_now <= res->runs_cap && res->runs) {
            if (runs_now > runs_before)
               CU(cudaMemcpyAsync(res->runs + runs_before, (const PgRun *)r.pool.p + runs_before, sizeof(PgRun) * (size_t)(runs_now - runs_before),
                                  cudaMemcpyDeviceToHost, ps->s_out));
         } else if (runs_now) {
            cap_exceeded = true;
         }
      }
      runs_before = runs_now;
      return PG_OK;
   };
   for (int c = 0; c < nchunks; c++) {
      const int c0 = c * chunk, c1 = (c0 + chunk < n) ? c0 + chunk : n, m = c1 - c0;
      uint32_t chunk_maxlen = 0;
      for (int i = c0; i < c1; i++) {
         if (in->seq_off[i + 1] < in->seq_off[i]) return fail(ctx, PG_ERR_ARG, "seq_off must be non-decreasing");
         const uint32_t l = in->seq_off[i + 1] - in->seq_off[i];
         if (l > chunk_maxlen) chunk_maxlen = l;
         if (in->chr_id && (in->chr_id[i] < 0 || (size_t)in->chr_id[i] >= ctx->chroms_h.size()))
            return fail(ctx, PG_ERR_ARG, "chr_id out of range");
      }
      if (chunk_maxlen >= PG_TABLE)
         return fail(ctx, PG_ERR_UNSUPPORTED, "read longer than 499 bases (the reference's g_maxMismatch table has 500 entries)");
      if (have_regions) {
         if (in->region_off[c1] < in->region_off[c0]) return fail(ctx, PG_ERR_ARG, "region_off must be non-decreasing");
         for (size_t i = in->region_off[c0], nr = in->region_off[c1]; i < nr; i++)
            if (in->regions[i].chr < 0 || (size_t)in->regions[i].chr >= ctx->chroms_h.size())
               return fail(ctx, PG_ERR_ARG, "region chromosome out of range");
      }
      if (chunk_maxlen > maxlen) { maxlen = chunk_maxlen; ctx->max_read_len = maxlen; }
#define UPS(buf, src, off, bytes) CU(cudaMemcpyAsync((char *)buf.p + (off), (const char *)(src) + (off), (bytes), cudaMemcpyHostToDevice, ps->s_in))
      const size_t b0 = in->seq_off[c0], b1 = in->seq_off[c1];
      if (b1 > b0) UPS(ctx->seq_blob, in->seq_blob, b0, b1 - b0);
      UPS(ctx->seq_off, in->seq_off, sizeof(uint32_t) * (size_t)c0, sizeof(uint32_t) * ((size_t)m + 1));
      if (have_anchor) {
         UPS(ctx->md, in->matched_d, (size_t)c0, (size_t)m);
         UPS(ctx->pos, in->matched_rel_pos, sizeof(uint32_t) * (size_t)c0, sizeof(uint32_t) * (size_t)m);
         UPS(ctx->isz, in->insert_size, sizeof(int16_t) * (size_t)c0, sizeof(int16_t) * (size_t)m);
      }
      if (in->chr_id) UPS(ctx->chr, in->chr_id, sizeof(int32_t) * (size_t)c0, sizeof(int32_t) * (size_t)m);
      if (in->close_info) UPS(ctx->closeinfo, in->close_info, sizeof(pg_close_info) * (size_t)c0, sizeof(pg_close_info) * (size_t)m);
      if (have_regions) {
         const size_t r0 = in->region_off[c0], r1 = in->region_off[c1];
         if (r1 > r0) UPS(ctx->regions, in->regions, sizeof(pg_region) * r0, sizeof(pg_region) * (r1 - r0));
         UPS(ctx->region_off, in->region_off, sizeof(uint32_t) * (size_t)c0, sizeof(uint32_t) * ((size_t)m + 1));
      }
#undef UPS
      CU(cudaEventRecord(ps->in_done[c], ps->s_in));
      CU(cudaStreamWaitEvent(ps->s_k, ps->in_done[c], 0));
      CU(cudaMemsetAsync(r.counters.p, 0, 4, ps->s_k));       // the chunk's own read counter; the run allocator and the statistics run on
      PgBatch bt;
      memset(&bt, 0, sizeof(bt));
      bt.n_reads = m;
      bt.mode = mode;
      bt.seq_blob = (const uint8_t *)ctx->seq_blob.p;           // offsets are absolute: every chunk sees the whole blob / region list
      bt.seq_off = (const uint32_t *)ctx->seq_off.p + c0;
      bt.matched_d = have_anchor ? (const uint8_t *)ctx->md.p + c0 : nullptr;
      bt.matched_rel_pos = have_anchor ? (const uint32_t *)ctx->pos.p + c0 : nullptr;
      bt.insert_size = have_anchor ? (const int16_t *)ctx->isz.p + c0 : nullptr;
      bt.chr_id = in->chr_id ? (const int32_t *)ctx->chr.p + c0 : nullptr;
      bt.close_info = (mode == 2 && in->close_info) ? (const PgCloseInfo *)ctx->closeinfo.p + c0 : nullptr;
      bt.regions = have_regions ? (const PgRegion *)ctx->regions.p : nullptr;
      bt.region_off = have_regions ? (const uint32_t *)ctx->region_off.p + c0 : nullptr;
      bt.out = (PgReadOut *)r.out.p + c0;
      bt.run_pool = (PgRun *)r.pool.p;
      bt.run_pool_cap = r.pool_cap;
      bt.counters = (uint32_t *)r.counters.p;
      rc = dispatch_variant(ctx, bt, chunk_maxlen, ps->s_k, ctx->scratch);
      if (rc != PG_OK) return rc;
      CU(cudaMemcpyAsync(ps->counters_h + 8 * (size_t)c, r.counters.p, 32, cudaMemcpyDeviceToHost, ps->s_k));
      CU(cudaEventRecord(ps->k_done[c], ps->s_k));
      if (res && c >= 1) { rc = drain(c - 1); if (rc != PG_OK) return rc; }
   }
   // the slot's own completion record: final counters + event, and the caller's stream waits for it
   CU(cudaMemcpyAsync(r.counters_h, r.counters.p, 32, cudaMemcpyDeviceToHost, ps->s_k));
   CU(cudaEventRecord(r.done, ps->s_k));
   CU(cudaStreamWaitEvent(ctx->stream, r.done, 0));
   r.launched = true;
   ctx->launched = true;
   ctx->last_mode = mode;
   if (!res) return PG_OK;          // resident only: nothing to wait for here (ps->busy stays set; the next call settles)
   rc = drain(nchunks - 1);
   if (rc != PG_OK) return rc;
   CU(cudaEventRecord(ps->out_done, ps->s_out));
   CU(cudaEventSynchronize(ps->out_done));
   if (overflow) {
      // the run pool was too small for some chunk: finish_slot grows it to the exact size and repeats the launch over the
      // (now resident) batch, then everything is copied in one piece
      rc = finish_launch(ctx);
      if (rc != PG_OK) return rc;
      res->n_runs = ctx->last_runs;
      if (res->n_runs > res->runs_cap || (res->n_runs && !res->runs))
         return fail(ctx, PG_ERR_RUNS_CAP, "run buffer too small (results->n_runs holds the required size; repeat the call)");
      CU(cudaMemcpyAsync(res->out, r.out.p, sizeof(PgReadOut) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
      if (res->n_runs) CU(cudaMemcpyAsync(res->runs, r.pool.p, sizeof(PgRun) * (size_t)res->n_runs, cudaMemcpyDeviceToHost, ctx->stream));
      CU(cudaStreamSynchronize(ctx->stream));
      return PG_OK;
   }
   rc = finish_launch(ctx);        // totals of the call (the event is long reached)
   if (rc != PG_OK) return rc;
   res->n_runs = ctx->last_runs;
   if (cap_exceeded) return fail(ctx, PG_ERR_RUNS_CAP, "run buffer too small (results->n_runs holds the required size; repeat the call)");
   return PG_OK;
}

// batches from this size on take the chunked path (copies of chunk k +- 1 under the kernel of chunk k)
#define PG_PIPE_MIN_READS (1 << 18)
static inline bool pipe_applies(const pg_ctx *ctx, int n)
{
   return ctx->pipe_chunk_forced ? n >= 2 * ctx->pipe_chunk : n >= PG_PIPE_MIN_READS;
}

// a far-end batch without clusters of its own takes them from the loaded BreakDancer region (pg_bd_load_region)
static inline bool bd_applies(const pg_ctx *ctx, const pg_reads *in, int mode)
{
   return mode == 2 && ctx->bk_loaded && in->close_info && !(in->regions && in->region_off);
}

extern "C" {

static int run_all(pg_ctx *ctx, const pg_reads *in, pg_results *res, int mode)
{
   // (a far-end batch that takes its clusters from the loaded BreakDancer region goes through upload + launch: the lookup
   // runs on the whole resident batch)
   if (ctx && in && pipe_applies(ctx, in->n) && in->seq_blob && in->seq_off && !bd_applies(ctx, in, mode)) {
      CU(cudaSetDevice(ctx->device));
      return run_pipelined(ctx, ctx->pipe, in, res, mode);
   }
   int rc = pg_upload(ctx, in);
   if (rc != PG_OK) return rc;
   rc = pg_launch(ctx, mode);
   if (rc != PG_OK) return rc;
   return pg_download(ctx, res);
}

int pg_close_end(pg_ctx *ctx, const pg_reads *in, pg_results *res) { return run_all(ctx, in, res, 1); }
int pg_far_end(pg_ctx *ctx, const pg_reads *in, pg_results *res) { return run_all(ctx, in, res, 2); }
int pg_search(pg_ctx *ctx, const pg_reads *in, pg_results *res) { return run_all(ctx, in, res, 3); }

int pg_search_resident(pg_ctx *ctx, const pg_reads *in, int32_t mode)
{
   if (!ctx || !in || in->n < 0 || mode < 1 || mode > 3) return fail(ctx, PG_ERR_ARG, "pg_search_resident: bad arguments");
   if (pipe_applies(ctx, in->n) && in->seq_blob && in->seq_off && !bd_applies(ctx, in, mode)) {
      CU(cudaSetDevice(ctx->device));
      return run_pipelined(ctx, ctx->pipe, in, nullptr, mode);
   }
   int rc = pg_upload(ctx, in);
   if (rc != PG_OK) return rc;
   return pg_launch(ctx, mode);
}

// ------------------------------------------------------------------------------------------------
// Several GPUs behind one handle ("device id(s) ... multi-GPU fan-out is internal").  A group owns one
// context per device, the reference is replicated, and a batch is dealt to the devices in contiguous ranges of reads: one
// host thread per device uploads its range chunk by chunk behind its kernel (pg_search_resident), the run counts are
// exchanged, and every device copies its records straight to their final place in the caller's buffers -- the only
// "collective" of a single process is that prefix sum.  Reads are independent units, so the result is that of one device.
// ------------------------------------------------------------------------------------------------
struct pg_group {
   std::vector<pg_ctx *> ctx;
   std::string err;
};

static int group_run(pg_group *g, const pg_reads *in, pg_results *res, int mode)
{
   if (!g || !in || !res || in->n < 0) return PG_ERR_ARG;
   const int nd = (int)g->ctx.size(), n = in->n;
   res->n_runs = 0;
   if (n == 0) return PG_OK;
   if (!in->seq_blob || !in->seq_off || !res->out) { g->err = "pg_group: null buffers"; return PG_ERR_ARG; }
   std::vector<int> c0(nd + 1);
   for (int k = 0; k <= nd; k++) c0[k] = (int)((long long)n * k / nd);
   std::vector<int> rc(nd, PG_OK);
   std::vector<uint32_t> nruns(nd, 0);
   std::vector<std::vector<uint32_t>> off(nd), roff(nd);
   std::vector<pg_reads> sub(nd);
   // phase 1: every device searches its range; batch and results stay on the device
   auto phase1 = [&](int k) {
      const int a = c0[k], b = c0[k + 1], m = b - a;
      if (m == 0) return;
      pg_reads &r = sub[k];
      r = *in;
      r.n = m;
      off[k].resize((size_t)m + 1);
      for (int i = 0; i <= m; i++) off[k][i] = in->seq_off[a + i] - in->seq_off[a];
      r.seq_blob = in->seq_blob + in->seq_off[a];
      r.seq_off = off[k].data();
      if (in->matched_d) r.matched_d = in->matched_d + a;
      if (in->matched_rel_pos) r.matched_rel_pos = in->matched_rel_pos + a;
      if (in->insert_size) r.insert_size = in->insert_size + a;
      if (in->chr_id) r.chr_id = in->chr_id + a;
      if (in->close_info) r.close_info = in->close_info + a;
      if (in->regions && in->region_off) {
         roff[k].resize((size_t)m + 1);
         for (int i = 0; i <= m; i++) roff[k][i] = in->region_off[a + i] - in->region_off[a];
         r.regions = in->regions + in->region_off[a];
         r.region_off = roff[k].data();
      }
      int e = pg_search_resident(g->ctx[k], &r, mode);
      if (e == PG_OK) e = finish_launch(g->ctx[k]);
      rc[k] = e;
      nruns[k] = g->ctx[k]->last_runs;
   };
   {
      std::vector<std::thread> th;
      for (int k = 1; k < nd; k++) th.emplace_back(phase1, k);
      phase1(0);
      for (auto &t : th) t.join();
   }
   for (int k = 0; k < nd; k++)
      if (rc[k] != PG_OK) { g->err = "device " + std::to_string(g->ctx[k]->device) + ": " + g->ctx[k]->err; return rc[k]; }
   std::vector<uint64_t> base(nd + 1, 0);
   for (int k = 0; k < nd; k++) base[k + 1] = base[k] + nruns[k];
   if (base[nd] > 0xFFFFFFFFull) { g->err = "pg_group: more than 2^32 runs in one batch"; return PG_ERR_UNSUPPORTED; }
   res->n_runs = (uint32_t)base[nd];
   if (res->n_runs > res->runs_cap || (res->n_runs && !res->runs)) { g->err = "run buffer too small (results->n_runs holds the required size; repeat the call)"; return PG_ERR_RUNS_CAP; }
   // phase 2: the records go to their final place, run offsets made global
   auto phase2 = [&](int k) {
      const int a = c0[k], m = c0[k + 1] - a;
      if (m == 0) return;
      pg_results r;
      r.out = res->out + a; r.runs = res->runs ? res->runs + base[k] : nullptr; r.runs_cap = nruns[k]; r.n_runs = 0;
      rc[k] = pg_download(g->ctx[k], &r);
      if (rc[k] == PG_OK && base[k])
         for (int i = 0; i < m; i++) res->out[a + i].run_off += (uint32_t)base[k];
   };
   {
      std::vector<std::thread> th;
      for (int k = 1; k < nd; k++) th.emplace_back(phase2, k);
      phase2(0);
      for (auto &t : th) t.join();
   }
   for (int k = 0; k < nd; k++)
      if (rc[k] != PG_OK) { g->err = "device " + std::to_string(g->ctx[k]->device) + ": " + g->ctx[k]->err; return rc[k]; }
   return PG_OK;
}

pg_group *pg_group_create(const pg_settings *s, int32_t n_devices, const int32_t *devices)
{
   g_create_error.clear();
   if (!s || n_devices < 1 || n_devices > 64) { g_create_error = "pg_group_create: bad arguments"; return nullptr; }
   pg_group *g = new pg_group();
   for (int k = 0; k < n_devices; k++) {
      pg_settings sk = *s;
      sk.device = devices ? devices[k] : k;
      pg_ctx *c = pg_create(&sk);
      if (!c) { for (pg_ctx *x : g->ctx) pg_destroy(x); delete g; return nullptr; }
      g->ctx.push_back(c);
   }
   return g;
}

void pg_group_destroy(pg_group *g)
{
   if (!g) return;
   for (pg_ctx *c : g->ctx) pg_destroy(c);
   delete g;
}

const char *pg_group_last_error(pg_group *g) { return g ? g->err.c_str() : "null group"; }
int32_t pg_group_size(pg_group *g) { return g ? (int32_t)g->ctx.size() : 0; }
pg_ctx *pg_group_context(pg_group *g, int32_t k) { return (g && k >= 0 && (size_t)k < g->ctx.size()) ? g->ctx[k] : nullptr; }

int pg_group_set_reference(pg_group *g, int32_t n_chr, const char *const *seq, const uint32_t *len)
{
   if (!g) return PG_ERR_ARG;
   std::vector<int> rc(g->ctx.size(), PG_OK);
   std::vector<std::thread> th;
   for (size_t k = 1; k < g->ctx.size(); k++) th.emplace_back([&, k] { rc[k] = pg_set_reference(g->ctx[k], n_chr, seq, len); });
   rc[0] = pg_set_reference(g->ctx[0], n_chr, seq, len);
   for (auto &t : th) t.join();
   for (size_t k = 0; k < g->ctx.size(); k++)
      if (rc[k] != PG_OK) { g->err = "device " + std::to_string(g->ctx[k]->device) + ": " + g->ctx[k]->err; return rc[k]; }
   return PG_OK;
}

int pg_group_close_end(pg_group *g, const pg_reads *in, pg_results *res) { return group_run(g, in, res, 1); }
int pg_group_far_end(pg_group *g, const pg_reads *in, pg_results *res) { return group_run(g, in, res, 2); }
int pg_group_search(pg_group *g, const pg_reads *in, pg_results *res) { return group_run(g, in, res, 3); }

uint64_t pg_kernel_launches(pg_ctx *ctx) { return ctx ? ctx->launches : 0; }
uint32_t pg_last_run_count(pg_ctx *ctx) { return ctx ? ctx->last_runs : 0; }

int pg_last_stats(pg_ctx *ctx, uint64_t *window_bases, uint64_t *unique_points)
{
   if (!ctx) return PG_ERR_ARG;
   CU(cudaSetDevice(ctx->device));
   int rc = finish_launch(ctx);
   if (rc != PG_OK) return rc;
   if (window_bases) *window_bases = ctx->last_window_bases;
   if (unique_points) *unique_points = ctx->last_ups;
   return PG_OK;
}

int pg_device_results(pg_ctx *ctx, void **out_dev, void **runs_dev, uint32_t *n_runs)
{
   if (!ctx) return PG_ERR_ARG;
   CU(cudaSetDevice(ctx->device));
   int rc = finish_launch(ctx);
   if (rc != PG_OK) return rc;
   if (out_dev) *out_dev = ctx->rs[ctx->cur].out.p;
   if (runs_dev) *runs_dev = ctx->rs[ctx->cur].pool.p;
   if (n_runs) *n_runs = ctx->last_runs;
   return PG_OK;
}

int pg_device_results_previous(pg_ctx *ctx, void **out_dev, void **runs_dev, uint32_t *n_runs, int32_t *n_reads)
{
   if (!ctx) return PG_ERR_ARG;
   CU(cudaSetDevice(ctx->device));
   const int slot = (ctx->cur + PG_RES_SLOTS - 1) % PG_RES_SLOTS;
   if (!ctx->rs[slot].launched) return fail(ctx, PG_ERR_STATE, "pg_device_results_previous: only one launch so far");
   int rc = finish_slot(ctx, slot);
   if (rc != PG_OK) return rc;
   if (out_dev) *out_dev = ctx->rs[slot].out.p;
   if (runs_dev) *runs_dev = ctx->rs[slot].pool.p;
   if (n_runs) *n_runs = ctx->rs[slot].runs;
   if (n_reads) *n_reads = ctx->rs[slot].n;
   return PG_OK;
}

void *pg_host_alloc(size_t bytes)
{
   void *p = nullptr;
   if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) return nullptr;
   return p;
}
void pg_host_free(void *p) { if (p) cudaFreeHost(p); }

uint32_t pg_expand_runs(const pg_run *runs, uint32_t n_runs, pg_up *out) { return pgh_expand_runs(runs, n_runs, out); }
uint32_t pg_clean_unique_points(pg_up *ups, uint32_t n) { return pgh_clean_unique_points(ups, n); }

// ---- SV classification ----------------------------------------------------------------------------------------
// The device part of the classification stage on inputs that are already resident: pg_sv_kernel into a record pool
// (grown and repeated on overflow), the record order restored by scan + gather, one copy back.
static int sv_classify_device(pg_ctx *ctx, PgSvBatch &b, pg_sv_results *res)
{
   cudaStream_t st = ctx->stream;
   const int n = b.n_reads;
   int rc;
   // records: most reads fire at most once per driver and stop at the first accepted pair
   size_t cap = (size_t)n + (size_t)n / 2 + 1024;
   const size_t n_counts = (size_t)PG_SV_NDRIVERS * (size_t)n;
   const uint32_t n16 = (uint32_t)((n_counts + 15) / 16);
   const uint32_t n_blocks = (n16 + PG_SCAN_THREADS - 1) / PG_SCAN_THREADS;
   if ((rc = ensure(ctx, ctx->sv_counters, sizeof(uint32_t) * 4)) != PG_OK) return rc;
   if ((rc = ensure(ctx, ctx->sv_counts, (size_t)n16 * 16)) != PG_OK) return rc;
   if ((rc = ensure(ctx, ctx->sv_offsets, sizeof(uint32_t) * (size_t)n16 * 16)) != PG_OK) return rc;
   if ((rc = ensure(ctx, ctx->sv_blocksums, sizeof(uint32_t) * ((size_t)n_blocks + 1))) != PG_OK) return rc;
   uint32_t counters[4];
   if (!ctx->sv_ev0) { CU(cudaEventCreate(&ctx->sv_ev0)); CU(cudaEventCreate(&ctx->sv_ev1)); }
   if ((rc = ensure(ctx, ctx->sv_perm, sizeof(uint32_t) * (size_t)n)) != PG_OK) return rc;
   const bool group_by_strand = getenv("PG_SV_NO_GROUPING") == nullptr;
   for (int attempt = 0;; attempt++) {
      if ((rc = ensure(ctx, ctx->sv_records, sizeof(PgSvRecord) * cap)) != PG_OK) return rc;
      CU(cudaEventRecord(ctx->sv_ev0, st));
      b.perm = nullptr;
      if (group_by_strand) {      // the scan buffers are free until the record-order pass below
         const uint32_t m16 = ((uint32_t)n + 15u) / 16u, m_blocks = (m16 + PG_SCAN_THREADS - 1) / PG_SCAN_THREADS;
         CU(cudaMemsetAsync(ctx->sv_counts.p, 0, (size_t)m16 * 16, st));
         pg_sv_strand_flag_kernel<<<(n + 255) / 256, 256, 0, st>>>(n, b.matched_d, (uint8_t *)ctx->sv_counts.p);
         pg_scan_reduce<<<m_blocks, PG_SCAN_THREADS, 0, st>>>((const uint4 *)ctx->sv_counts.p, m16, (uint32_t *)ctx->sv_blocksums.p);
         pg_scan_blocks<<<1, 1024, 0, st>>>((uint32_t *)ctx->sv_blocksums.p, m_blocks);
         pg_scan_apply<<<m_blocks, PG_SCAN_THREADS, 0, st>>>((const uint4 *)ctx->sv_counts.p, m16, (const uint32_t *)ctx->sv_blocksums.p,
                                                            (uint32_t *)ctx->sv_offsets.p);
         pg_sv_perm_kernel<<<(n + 255) / 256, 256, 0, st>>>(n, (const uint8_t *)ctx->sv_counts.p, (const uint32_t *)ctx->sv_offsets.p,
                                                         (const uint32_t *)ctx->sv_blocksums.p + m_blocks, (uint32_t *)ctx->sv_perm.p);
         CU(cudaGetLastError());
         ctx->sv_aux_launches += 5;
         b.perm = (const uint32_t *)ctx->sv_perm.p;
      }
      CU(cudaMemsetAsync(ctx->sv_counters.p, 0, sizeof(uint32_t) * 4, st));
      CU(cudaMemsetAsync(ctx->sv_counts.p, 0, (size_t)n16 * 16, st));
      b.records = (PgSvRecord *)ctx->sv_records.p;
      b.records_cap = (uint32_t)cap;
      b.counters = (uint32_t *)ctx->sv_counters.p;
      b.counts = (uint8_t *)ctx->sv_counts.p;
      pg_sv_kernel<<<(n + 127) / 128, 128, 0, st>>>(b, (const PgSvParams *)ctx->sv_params.p, ctx->prm_d, ctx->chroms_d);
      CU(cudaGetLastError());
      ctx->sv_launches++;
      pg_scan_reduce<<<n_blocks, PG_SCAN_THREADS, 0, st>>>((const uint4 *)ctx->sv_counts.p, n16, (uint32_t *)ctx->sv_blocksums.p);
      pg_scan_blocks<<<1, 1024, 0, st>>>((uint32_t *)ctx->sv_blocksums.p, n_blocks);
      pg_scan_apply<<<n_blocks, PG_SCAN_THREADS, 0, st>>>((const uint4 *)ctx->sv_counts.p, n16, (const uint32_t *)ctx->sv_blocksums.p,
                                                          (uint32_t *)ctx->sv_offsets.p);
      CU(cudaGetLastError());
      CU(cudaMemcpyAsync(counters, ctx->sv_counters.p, sizeof counters, cudaMemcpyDeviceToHost, st));
      CU(cudaStreamSynchronize(st));
      if (counters[0] <= cap) break;
      if (attempt >= 4) return fail(ctx, PG_ERR_CUDA, "pg_classify: record pool keeps overflowing");
      cap = (size_t)counters[0] + 1024;
   }
   res->n_records = counters[0];
   if (counters[0]) {
      if ((rc = ensure(ctx, ctx->sv_ordered, sizeof(PgSvRecord) * (size_t)counters[0])) != PG_OK) return rc;
      pg_sv_order_kernel<<<(counters[0] + 255) / 256, 256, 0, st>>>((const PgSvRecord *)ctx->sv_records.p, counters[0],
                                                                    (const uint32_t *)ctx->sv_offsets.p, (uint32_t)n,
                                                                    (PgSvRecord *)ctx->sv_ordered.p);
      CU(cudaGetLastError());
      ctx->sv_aux_launches += 4;
   }
   CU(cudaEventRecord(ctx->sv_ev1, st));
   CU(cudaStreamSynchronize(st));
   CU(cudaEventElapsedTime(&ctx->sv_last_ms, ctx->sv_ev0, ctx->sv_ev1));
   if (counters[0] > res->records_cap || (counters[0] && !res->records))
      return fail(ctx, PG_ERR_RUNS_CAP, "pg_classify: record buffer too small (n_records holds the required size)");
   if (counters[0]) {
      CU(cudaMemcpyAsync(res->records, ctx->sv_ordered.p, sizeof(PgSvRecord) * (size_t)counters[0], cudaMemcpyDeviceToHost, st));
      CU(cudaStreamSynchronize(st));
   }
   return PG_OK;
}

static_assert(sizeof(pg_sv_record) == sizeof(PgSvRecord) && sizeof(pg_sv_record) == 44, "sv record layout");

static int sv_upload_window(pg_ctx *ctx, const pg_sv_window *win)
{
   if (win->chr < 0 || win->chr >= (int32_t)ctx->chroms_h.size()) return fail(ctx, PG_ERR_ARG, "pg_classify: chromosome out of range");
   if (win->box_size == 0) return fail(ctx, PG_ERR_ARG, "pg_classify: box_size is 0");
   PgSvParams &P = ctx->sv_params_h;
   memset(&P, 0, sizeof P);
   P.driver_mask = win->driver_mask & PG_SV_ALL;
   P.chr = win->chr;
   P.window_end = win->window_end;
   P.num_boxes = win->num_boxes;
   P.box_size = win->box_size;
   P.region_start_defined = win->region_start_defined;
   P.region_start = win->region_start;
   P.region_end_defined = win->region_end_defined;
   P.region_end = win->region_end;
   P.min_inv = win->min_inversion_size;
   P.min_matched = win->min_matched_bases;
   for (int k = 0; k < 1032; k++) P.nt_thr[k] = (uint16_t)(short)(1 + ctx->st.seq_error_rate * k); // src/search_deletions_nt.cpp:50-51
   int rc;
   if ((rc = ensure(ctx, ctx->sv_params, sizeof(PgSvParams))) != PG_OK) return rc;
   CU(cudaMemcpyAsync(ctx->sv_params.p, &P, sizeof P, cudaMemcpyHostToDevice, ctx->stream));
   return PG_OK;
}

int pg_classify(pg_ctx *ctx, const pg_sv_reads *in, const pg_sv_window *win, pg_sv_results *res)
{
   if (!ctx || !in || !win || !res) return fail(ctx, PG_ERR_ARG, "pg_classify: null argument");
   if (ctx->chroms_h.empty()) return fail(ctx, PG_ERR_STATE, "pg_classify: pg_set_reference has not been called");
   if (win->chr < 0 || win->chr >= (int32_t)ctx->chroms_h.size()) return fail(ctx, PG_ERR_ARG, "pg_classify: chromosome out of range");
   if (win->box_size == 0) return fail(ctx, PG_ERR_ARG, "pg_classify: box_size is 0");
   res->n_records = 0;
   const int n = in->n_reads;
   if (n < 0) return fail(ctx, PG_ERR_ARG, "pg_classify: negative read count");
   if (n == 0) return PG_OK;
   if (!in->seq_blob || !in->seq_off || !in->matched_d || !in->insert_size || !in->rflags || !in->run_off || !in->n_close_runs)
      return fail(ctx, PG_ERR_ARG, "pg_classify: null input array");
   const uint32_t n_runs = in->run_off[n];
   if (n_runs && !in->runs) return fail(ctx, PG_ERR_ARG, "pg_classify: null run array");
   for (int i = 0; i < n; i++) {
      if (in->run_off[i + 1] < in->run_off[i] || in->n_close_runs[i] > in->run_off[i + 1] - in->run_off[i])
         return fail(ctx, PG_ERR_ARG, "pg_classify: inconsistent run offsets");
      if (in->seq_off[i + 1] < in->seq_off[i] || in->seq_off[i + 1] - in->seq_off[i] >= 512)
         return fail(ctx, PG_ERR_UNSUPPORTED, "pg_classify: read of 512 bases or more");
   }
   CU(cudaSetDevice(ctx->device));
   cudaStream_t st = ctx->stream;
   int rc;
   if ((rc = sv_upload_window(ctx, win)) != PG_OK) return rc;
   const size_t seq_bytes = in->seq_off[n];
   if ((rc = ensure(ctx, ctx->sv_seq, seq_bytes + 4)) != PG_OK) return rc;
   if ((rc = ensure(ctx, ctx->sv_off, sizeof(uint32_t) * (size_t)(n + 1))) != PG_OK) return rc;
   if ((rc = ensure(ctx, ctx->sv_md, (size_t)n)) != PG_OK) return rc;
   if ((rc = ensure(ctx, ctx->sv_isz, sizeof(int16_t) * (size_t)n)) != PG_OK) return rc;
   if ((rc = ensure(ctx, ctx->sv_flags, (size_t)n)) != PG_OK) return rc;
   if ((rc = ensure(ctx, ctx->sv_runoff, sizeof(uint32_t) * (size_t)(n + 1))) != PG_OK) return rc;
   if ((rc = ensure(ctx, ctx->sv_nclose, sizeof(uint16_t) * (size_t)n)) != PG_OK) return rc;
   if ((rc = ensure(ctx, ctx->sv_runs, sizeof(PgRun) * (size_t)(n_runs + 1))) != PG_OK) return rc;
   CU(cudaMemcpyAsync(ctx->sv_seq.p, in->seq_blob, seq_bytes, cudaMemcpyHostToDevice, st));
   CU(cudaMemcpyAsync(ctx->sv_off.p, in->seq_off, sizeof(uint32_t) * (size_t)(n + 1), cudaMemcpyHostToDevice, st));
   CU(cudaMemcpyAsync(ctx->sv_md.p, in->matched_d, (size_t)n, cudaMemcpyHostToDevice, st));
   CU(cudaMemcpyAsync(ctx->sv_isz.p, in->insert_size, sizeof(int16_t) * (size_t)n, cudaMemcpyHostToDevice, st));
   CU(cudaMemcpyAsync(ctx->sv_flags.p, in->rflags, (size_t)n, cudaMemcpyHostToDevice, st));
   CU(cudaMemcpyAsync(ctx->sv_runoff.p, in->run_off, sizeof(uint32_t) * (size_t)(n + 1), cudaMemcpyHostToDevice, st));
   CU(cudaMemcpyAsync(ctx->sv_nclose.p, in->n_close_runs, sizeof(uint16_t) * (size_t)n, cudaMemcpyHostToDevice, st));
   if (n_runs) CU(cudaMemcpyAsync(ctx->sv_runs.p, in->runs, sizeof(PgRun) * (size_t)n_runs, cudaMemcpyHostToDevice, st));
   PgSvBatch b;
   memset(&b, 0, sizeof b);
   b.n_reads = n;
   b.seq_blob = (const uint8_t *)ctx->sv_seq.p;
   b.seq_off = (const uint32_t *)ctx->sv_off.p;
   b.matched_d = (const uint8_t *)ctx->sv_md.p;
   b.insert_size = (const int16_t *)ctx->sv_isz.p;
   b.rflags = (const uint8_t *)ctx->sv_flags.p;
   b.run_off = (const uint32_t *)ctx->sv_runoff.p;
   b.n_close_runs = (const uint16_t *)ctx->sv_nclose.p;
   b.runs = (const PgRun *)ctx->sv_runs.p;
   return sv_classify_device(ctx, b, res);
}

int pg_classify_resident(pg_ctx *ctx, const pg_sv_window *win, const uint8_t *rflags, const uint32_t *close_run_off,
                         const pg_run *close_runs, pg_sv_results *res)
{
   if (!ctx || !win || !res) return fail(ctx, PG_ERR_ARG, "pg_classify_resident: null argument");
   if (!ctx->launched || !(ctx->last_mode & 2)) return fail(ctx, PG_ERR_STATE, "pg_classify_resident: no far-end launch is resident");
   if (!ctx->have_anchor) return fail(ctx, PG_ERR_ARG, "pg_classify_resident: the resident batch has no matched_d / insert_size");
   if (ctx->last_mode == 2 && (!close_run_off || (close_run_off[ctx->n] && !close_runs)))
      return fail(ctx, PG_ERR_ARG, "pg_classify_resident: a far-end-only launch needs the reads' close runs");
   res->n_records = 0;
   CU(cudaSetDevice(ctx->device));
   int rc = finish_launch(ctx);
   if (rc != PG_OK) return rc;
   const int n = ctx->n;
   if (n == 0) return PG_OK;
   cudaStream_t st = ctx->stream;
   if ((rc = sv_upload_window(ctx, win)) != PG_OK) return rc;
   PgSvBatch b;
   memset(&b, 0, sizeof b);
   b.n_reads = n;
   b.seq_off = (const uint32_t *)ctx->seq_off.p;
   b.matched_d = (const uint8_t *)ctx->md.p;
   b.insert_size = (const int16_t *)ctx->isz.p;
   b.chr_id = ctx->have_chr ? (const int32_t *)ctx->chr.p : nullptr;
   b.res_out = (const PgReadOut *)ctx->rs[ctx->cur].out.p;
   b.res_pool = (const PgRun *)ctx->rs[ctx->cur].pool.p;
   if (rflags) {
      if ((rc = ensure(ctx, ctx->sv_flags, (size_t)n)) != PG_OK) return rc;
      CU(cudaMemcpyAsync(ctx->sv_flags.p, rflags, (size_t)n, cudaMemcpyHostToDevice, st));
      b.rflags = (const uint8_t *)ctx->sv_flags.p;
   }
   if (ctx->last_mode == 2) {
      for (int i = 0; i < n; i++)
         if (close_run_off[i + 1] < close_run_off[i]) return fail(ctx, PG_ERR_ARG, "pg_classify_resident: inconsistent close run offsets");
      const size_t nr = close_run_off[n];
      if ((rc = ensure(ctx, ctx->sv_close_off, sizeof(uint32_t) * ((size_t)n + 1))) != PG_OK) return rc;
      if ((rc = ensure(ctx, ctx->sv_close_runs, sizeof(PgRun) * (nr + 1))) != PG_OK) return rc;
      CU(cudaMemcpyAsync(ctx->sv_close_off.p, close_run_off, sizeof(uint32_t) * ((size_t)n + 1), cudaMemcpyHostToDevice, st));
      if (nr) CU(cudaMemcpyAsync(ctx->sv_close_runs.p, close_runs, sizeof(PgRun) * nr, cudaMemcpyHostToDevice, st));
      b.close_off = (const uint32_t *)ctx->sv_close_off.p;
      b.close_runs = (const PgRun *)ctx->sv_close_runs.p;
      b.seq_blob = (const uint8_t *)ctx->seq_blob.p;     // pg_far_end takes the sequences as GetCloseEnd left them
   } else {
      // close + far launch: the flips happened on the device; materialise the final sequences next to the originals
      if ((rc = ensure(ctx, ctx->sv_seq, ctx->blob_bytes + 4)) != PG_OK) return rc;
      pg_sv_final_seq_kernel<<<(n + 7) / 8, 256, 0, st>>>(n, (const uint8_t *)ctx->seq_blob.p, (const uint32_t *)ctx->seq_off.p,
                                                        (const PgReadOut *)ctx->rs[ctx->cur].out.p, (uint8_t *)ctx->sv_seq.p);
      CU(cudaGetLastError());
      ctx->sv_aux_launches++;
      b.seq_blob = (const uint8_t *)ctx->sv_seq.p;
   }
   return sv_classify_device(ctx, b, res);
}

uint64_t pg_classify_launches(pg_ctx *ctx) { return ctx ? ctx->sv_launches : 0; }
double pg_classify_last_kernel_ms(pg_ctx *ctx) { return ctx ? (double)ctx->sv_last_ms : 0.0; }

// ---- BGZF inflation (row N3, first half) ---------------------------------------------------------------------------
static_assert(sizeof(PgzBlock) == 32, "bgzf block layout");

int64_t pg_bgzf_inflated_size(const uint8_t *file, size_t file_len)
{
   if (!file && file_len) return -1;
   std::vector<PgzBlock> blocks;
   if (pg_bgzf_walk(file, file_len, blocks, nullptr) < 0) return -1;
   return blocks.empty() ? 0 : (int64_t)(blocks.back().out_off + blocks.back().isize);
}

int pg_bgzf_inflate(pg_ctx *ctx, const uint8_t *file, size_t file_len, uint8_t *out, size_t out_cap, size_t *out_len, int32_t check_crc)
{
   if (!ctx || (!file && file_len) || !out_len) return fail(ctx, PG_ERR_ARG, "pg_bgzf_inflate: bad arguments");
   std::vector<PgzBlock> blocks;
   size_t bad = 0;
   if (pg_bgzf_walk(file, file_len, blocks, &bad) < 0)
      return fail(ctx, PG_ERR_DATA, "pg_bgzf_inflate: not a BGZF block header at offset " + std::to_string(bad));
   const size_t total = blocks.empty() ? 0 : (size_t)(blocks.back().out_off + blocks.back().isize);
   *out_len = total;
   ctx->bz_out_len = 0;
   ctx->bm_valid = false; ctx->bm_generation++;
   if (out && out_cap < total) return fail(ctx, PG_ERR_RUNS_CAP, "pg_bgzf_inflate: output buffer too small (out_len holds the required size)");
   if (blocks.empty()) return PG_OK;
   CU(cudaSetDevice(ctx->device));
   cudaStream_t st = ctx->stream;
   const uint32_t nb = (uint32_t)blocks.size();
   int rc;
   if ((rc = ensure(ctx, ctx->bz_file, file_len + 16)) != PG_OK) return rc;        // the bit reader may look a word past the end
   if ((rc = ensure(ctx, ctx->bz_blocks, sizeof(PgzBlock) * (size_t)nb)) != PG_OK) return rc;
   if ((rc = ensure(ctx, ctx->bz_out, total + 16)) != PG_OK) return rc;
   if ((rc = ensure(ctx, ctx->bz_status, sizeof(uint32_t) * (size_t)nb)) != PG_OK) return rc;
   CU(cudaMemcpyAsync(ctx->bz_file.p, file, file_len, cudaMemcpyHostToDevice, st));
   CU(cudaMemsetAsync((uint8_t *)ctx->bz_file.p + file_len, 0, 16, st));
   CU(cudaMemcpyAsync(ctx->bz_blocks.p, blocks.data(), sizeof(PgzBlock) * (size_t)nb, cudaMemcpyHostToDevice, st));
   if (!ctx->sv_ev0) { CU(cudaEventCreate(&ctx->sv_ev0)); CU(cudaEventCreate(&ctx->sv_ev1)); }
   CU(cudaEventRecord(ctx->sv_ev0, st));
   pg_bgzf_inflate_kernel<<<(nb + PGZ_WARPS - 1) / PGZ_WARPS, 32 * PGZ_WARPS, 0, st>>>((const uint8_t *)ctx->bz_file.p, (const PgzBlock *)ctx->bz_blocks.p,
                                                                                           nb, (uint8_t *)ctx->bz_out.p, (uint32_t *)ctx->bz_status.p,
                                                                                           check_crc ? 1 : 0);
   CU(cudaGetLastError());
   CU(cudaEventRecord(ctx->sv_ev1, st));
   ctx->bz_launches++;
   std::vector<uint32_t> status(nb);
   CU(cudaMemcpyAsync(status.data(), ctx->bz_status.p, sizeof(uint32_t) * (size_t)nb, cudaMemcpyDeviceToHost, st));
   if (out && total) CU(cudaMemcpyAsync(out, ctx->bz_out.p, total, cudaMemcpyDeviceToHost, st));
   CU(cudaStreamSynchronize(st));
   CU(cudaEventElapsedTime(&ctx->bz_last_ms, ctx->sv_ev0, ctx->sv_ev1));
   for (uint32_t i = 0; i < nb; i++)
      if (status[i] != PGZ_OK)
         return fail(ctx, PG_ERR_DATA, "pg_bgzf_inflate: block " + std::to_string(i) + " (file offset " + std::to_string(blocks[i].comp_off) +
                                           ") is corrupt: " + (status[i] == PGZ_ERR_CRC ? "CRC32 mismatch" : "invalid DEFLATE stream, code " + std::to_string(status[i])));
   ctx->bz_out_len = total;
   return PG_OK;
}

int pg_bgzf_device_data(pg_ctx *ctx, void **dev, size_t *len)
{
   if (!ctx || !dev || !len) return fail(ctx, PG_ERR_ARG, "pg_bgzf_device_data: null argument");
   *dev = ctx->bz_out_len ? ctx->bz_out.p : nullptr;
   *len = ctx->bz_out_len;
   return PG_OK;
}

uint64_t pg_bgzf_launches(pg_ctx *ctx) { return ctx ? ctx->bz_launches : 0; }
double pg_bgzf_last_kernel_ms(pg_ctx *ctx) { return ctx ? (double)ctx->bz_last_ms : 0.0; }

// ---- BAM record index (row N3, between inflate and filter) -----------------------------------------------------------
static_assert(sizeof(pg_bam_rec) == sizeof(PgBamRec) && sizeof(pg_bam_rec) == 24, "bam rec layout");

int pg_bam_index(pg_ctx *ctx, pg_bam_rec *recs, uint64_t cap, uint64_t *n_recs, uint64_t *first_record, int32_t *n_ref)
{
   if (!ctx || !n_recs) return fail(ctx, PG_ERR_ARG, "pg_bam_index: null argument");
   if (ctx->bz_out_len == 0) return fail(ctx, PG_ERR_STATE, "pg_bam_index: no inflated stream is resident (call pg_bgzf_inflate)");
   CU(cudaSetDevice(ctx->device));
   cudaStream_t st = ctx->stream;
   const uint8_t *d = (const uint8_t *)ctx->bz_out.p;
   const uint64_t n = ctx->bz_out_len;
   int rc;
   if ((rc = ensure(ctx, ctx->bm_info, sizeof(uint64_t) * 2)) != PG_OK) return rc;
   if (!ctx->sv_ev0) { CU(cudaEventCreate(&ctx->sv_ev0)); CU(cudaEventCreate(&ctx->sv_ev1)); }
   CU(cudaEventRecord(ctx->sv_ev0, st));
   pg_bam_header_kernel<<<1, 32, 0, st>>>(d, n, (uint64_t *)ctx->bm_info.p);
   uint64_t info[2];
   CU(cudaMemcpyAsync(info, ctx->bm_info.p, sizeof info, cudaMemcpyDeviceToHost, st));
   CU(cudaStreamSynchronize(st));
   if (info[0] == PGB_NONE) return fail(ctx, PG_ERR_DATA, "pg_bam_index: the inflated stream is not a BAM file (magic / header)");
   const uint64_t first = info[0], seg = PGB_SEGMENT;
   const int32_t nref = (int32_t)info[1];
   if (first_record) *first_record = first;
   if (n_ref) *n_ref = nref;
   const uint32_t n_seg = first >= n ? 1u : (uint32_t)((n - first + seg - 1) / seg);
   if ((rc = ensure(ctx, ctx->bm_starts, sizeof(uint64_t) * (size_t)n_seg)) != PG_OK) return rc;
   if ((rc = ensure(ctx, ctx->bm_counts, sizeof(uint64_t) * (size_t)n_seg)) != PG_OK) return rc;
   if ((rc = ensure(ctx, ctx->bm_landing, sizeof(uint64_t) * (size_t)n_seg)) != PG_OK) return rc;
   if ((rc = ensure(ctx, ctx->bm_base, sizeof(uint64_t) * (size_t)n_seg)) != PG_OK) return rc;
   pg_bam_start_kernel<<<(n_seg + 7) / 8, 256, 0, st>>>(d, n, first, seg, n_seg, nref, (uint64_t *)ctx->bm_starts.p);
   ctx->bm_launches += 2;
   std::vector<uint64_t> starts(n_seg), counts(n_seg), landing(n_seg), base(n_seg);
   uint64_t total = 0;
   uint32_t redone = 0;
   uint64_t *const starts_d = (uint64_t *)ctx->bm_starts.p, *const counts_d = (uint64_t *)ctx->bm_counts.p, *const landing_d = (uint64_t *)ctx->bm_landing.p;
   auto walk_all = [&]() -> int {
      pg_bam_walk_kernel<<<(n_seg + 127) / 128, 128, 0, st>>>(d, n, first, seg, n_seg, starts_d, counts_d, landing_d, nullptr, nullptr);
      CU(cudaGetLastError());
      ctx->bm_launches++;
      CU(cudaMemcpyAsync(starts.data(), starts_d, sizeof(uint64_t) * (size_t)n_seg, cudaMemcpyDeviceToHost, st));
      CU(cudaMemcpyAsync(counts.data(), counts_d, sizeof(uint64_t) * (size_t)n_seg, cudaMemcpyDeviceToHost, st));
      CU(cudaMemcpyAsync(landing.data(), landing_d, sizeof(uint64_t) * (size_t)n_seg, cudaMemcpyDeviceToHost, st));
      CU(cudaStreamSynchronize(st));
      return PG_OK;
   };
   // a guess locked on to bytes that only look like records: that segment alone is walked again from the chain's true offset
   // (the kernel over one segment: its `first` is the segment's own begin, its arrays start at the segment's entries)
   auto walk_one = [&](uint32_t s) -> int {
      CU(cudaMemcpyAsync(starts_d + s, &starts[s], sizeof(uint64_t), cudaMemcpyHostToDevice, st));
      pg_bam_walk_kernel<<<1, 32, 0, st>>>(d, n, first + (uint64_t)s * seg, seg, 1u, starts_d + s, counts_d + s, landing_d + s, nullptr, nullptr);
      CU(cudaGetLastError());
      ctx->bm_launches++;
      CU(cudaMemcpyAsync(&counts[s], counts_d + s, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
      CU(cudaMemcpyAsync(&landing[s], landing_d + s, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
      CU(cudaStreamSynchronize(st));
      return PG_OK;
   };
   auto validate = [&](uint64_t *forced) { return pgb_validate(n_seg, first, n, seg, starts.data(), counts.data(), landing.data(), forced, &total); };
   if ((rc = pgb_settle(starts.data(), validate, walk_all, walk_one, &redone)) != 0) {
      if (rc < 0) return fail(ctx, PG_ERR_DATA, "pg_bam_index: the record chain does not end at the end of the stream (truncated or corrupt BAM)");
      return rc;
   }
   ctx->bm_last_segments = n_seg;
   ctx->bm_last_redone = redone;
   *n_recs = total;
   if (total > cap || (total && !recs)) {
      CU(cudaEventRecord(ctx->sv_ev1, st));
      return fail(ctx, PG_ERR_RUNS_CAP, "pg_bam_index: record buffer too small (n_recs holds the required size)");
   }
   if (total) {
      uint64_t b = 0;
      for (uint32_t s = 0; s < n_seg; s++) { base[s] = b; b += counts[s]; }
      // segments the chain jumps over must not write: their guess is void
      for (uint32_t s = 0; s < n_seg; s++) if (counts[s] == 0) starts[s] = PGB_NONE;
      if ((rc = ensure(ctx, ctx->bm_recs, sizeof(PgBamRec) * (size_t)total)) != PG_OK) return rc;
      CU(cudaMemcpyAsync(ctx->bm_starts.p, starts.data(), sizeof(uint64_t) * (size_t)n_seg, cudaMemcpyHostToDevice, st));
      CU(cudaMemcpyAsync(ctx->bm_base.p, base.data(), sizeof(uint64_t) * (size_t)n_seg, cudaMemcpyHostToDevice, st));
      pg_bam_walk_kernel<<<(n_seg + 127) / 128, 128, 0, st>>>(d, n, first, seg, n_seg, (const uint64_t *)ctx->bm_starts.p, nullptr, nullptr,
                                                             (const uint64_t *)ctx->bm_base.p, (PgBamRec *)ctx->bm_recs.p);
      CU(cudaGetLastError());
      ctx->bm_launches++;
      CU(cudaEventRecord(ctx->sv_ev1, st));
      CU(cudaMemcpyAsync(recs, ctx->bm_recs.p, sizeof(PgBamRec) * (size_t)total, cudaMemcpyDeviceToHost, st));
   } else {
      CU(cudaEventRecord(ctx->sv_ev1, st));
   }
   CU(cudaStreamSynchronize(st));
   CU(cudaEventElapsedTime(&ctx->bm_last_ms, ctx->sv_ev0, ctx->sv_ev1));
   ctx->bm_n_recs = total;
   ctx->bm_valid = true; ctx->bm_generation++;
   return PG_OK;
}

// The record table of a run of BGZF blocks (a window's chunk between two .bai virtual offsets) instead of a whole file: no BAM
// header in front -- the caller says where the first record starts and how many references the file has -- and the last
// record may be cut off by the end of the run (*tail = where it starts; == the stream's length when nothing is cut off).
int pg_bam_index_run(pg_ctx *ctx, uint64_t first_record, int32_t n_ref, pg_bam_rec *recs, uint64_t cap, uint64_t *n_recs, uint64_t *tail)
{
   if (!ctx || !n_recs || !tail) return fail(ctx, PG_ERR_ARG, "pg_bam_index_run: null argument");
   if (ctx->bz_out_len == 0) return fail(ctx, PG_ERR_STATE, "pg_bam_index_run: no inflated stream is resident (call pg_bgzf_inflate)");
   const uint64_t n = ctx->bz_out_len;
   if (first_record > n || n_ref < 0) return fail(ctx, PG_ERR_ARG, "pg_bam_index_run: first_record beyond the stream or negative n_ref");
   CU(cudaSetDevice(ctx->device));
   cudaStream_t st = ctx->stream;
   const uint8_t *d = (const uint8_t *)ctx->bz_out.p;
   const uint64_t first = first_record, seg = PGB_SEGMENT;
   ctx->bm_valid = false; ctx->bm_generation++;
   int rc;
   if (!ctx->sv_ev0) { CU(cudaEventCreate(&ctx->sv_ev0)); CU(cudaEventCreate(&ctx->sv_ev1)); }
   CU(cudaEventRecord(ctx->sv_ev0, st));
   const uint32_t n_seg = first >= n ? 1u : (uint32_t)((n - first + seg - 1) / seg);
   if ((rc = ensure(ctx, ctx->bm_starts, sizeof(uint64_t) * (size_t)n_seg)) != PG_OK) return rc;
   if ((rc = ensure(ctx, ctx->bm_counts, sizeof(uint64_t) * (size_t)n_seg)) != PG_OK) return rc;
   if ((rc = ensure(ctx, ctx->bm_landing, sizeof(uint64_t) * (size_t)n_seg)) != PG_OK) return rc;
   if ((rc = ensure(ctx, ctx->bm_base, sizeof(uint64_t) * (size_t)n_seg)) != PG_OK) return rc;
   pg_bam_start_kernel<<<(n_seg + 7) / 8, 256, 0, st>>>(d, n, first, seg, n_seg, n_ref, (uint64_t *)ctx->bm_starts.p);
   ctx->bm_launches++;
   std::vector<uint64_t> starts(n_seg), counts(n_seg), landing(n_seg), base(n_seg);
   uint64_t total = 0, tail_off = n;
   uint32_t redone = 0;
   uint64_t *const starts_d = (uint64_t *)ctx->bm_starts.p, *const counts_d = (uint64_t *)ctx->bm_counts.p, *const landing_d = (uint64_t *)ctx->bm_landing.p;
   auto walk_all = [&]() -> int {
      pg_bam_walk_run_kernel<<<(n_seg + 127) / 128, 128, 0, st>>>(d, n, first, seg, n_seg, starts_d, counts_d, landing_d, nullptr, nullptr);
      CU(cudaGetLastError());
      ctx->bm_launches++;
      CU(cudaMemcpyAsync(starts.data(), starts_d, sizeof(uint64_t) * (size_t)n_seg, cudaMemcpyDeviceToHost, st));
      CU(cudaMemcpyAsync(counts.data(), counts_d, sizeof(uint64_t) * (size_t)n_seg, cudaMemcpyDeviceToHost, st));
      CU(cudaMemcpyAsync(landing.data(), landing_d, sizeof(uint64_t) * (size_t)n_seg, cudaMemcpyDeviceToHost, st));
      CU(cudaStreamSynchronize(st));
      return PG_OK;
   };
   auto walk_one = [&](uint32_t s) -> int {      // one segment again, from the chain's offset (see pg_bam_index)
      CU(cudaMemcpyAsync(starts_d + s, &starts[s], sizeof(uint64_t), cudaMemcpyHostToDevice, st));
      pg_bam_walk_run_kernel<<<1, 32, 0, st>>>(d, n, first + (uint64_t)s * seg, seg, 1u, starts_d + s, counts_d + s, landing_d + s, nullptr, nullptr);
      CU(cudaGetLastError());
      ctx->bm_launches++;
      CU(cudaMemcpyAsync(&counts[s], counts_d + s, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
      CU(cudaMemcpyAsync(&landing[s], landing_d + s, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
      CU(cudaStreamSynchronize(st));
      return PG_OK;
   };
   auto validate = [&](uint64_t *forced) {
      return pgb_validate_run(n_seg, first, n, seg, starts.data(), counts.data(), landing.data(), forced, &total, &tail_off);
   };
   if ((rc = pgb_settle(starts.data(), validate, walk_all, walk_one, &redone)) != 0) {
      if (rc < 0) return fail(ctx, PG_ERR_DATA, "pg_bam_index_run: a block_size in the record chain cannot be a record (wrong first_record or corrupt BAM)");
      return rc;
   }
   ctx->bm_last_segments = n_seg;
   ctx->bm_last_redone = redone;
   *n_recs = total;
   *tail = tail_off;
   if (total > cap || (total && !recs)) {
      CU(cudaEventRecord(ctx->sv_ev1, st));
      return fail(ctx, PG_ERR_RUNS_CAP, "pg_bam_index_run: record buffer too small (n_recs holds the required size)");
   }
   if (total) {
      uint64_t b = 0;
      for (uint32_t s = 0; s < n_seg; s++) { base[s] = b; b += counts[s]; }
      for (uint32_t s = 0; s < n_seg; s++) if (counts[s] == 0) starts[s] = PGB_NONE;
      if ((rc = ensure(ctx, ctx->bm_recs, sizeof(PgBamRec) * (size_t)total)) != PG_OK) return rc;
      CU(cudaMemcpyAsync(ctx->bm_starts.p, starts.data(), sizeof(uint64_t) * (size_t)n_seg, cudaMemcpyHostToDevice, st));
      CU(cudaMemcpyAsync(ctx->bm_base.p, base.data(), sizeof(uint64_t) * (size_t)n_seg, cudaMemcpyHostToDevice, st));
      pg_bam_walk_run_kernel<<<(n_seg + 127) / 128, 128, 0, st>>>(d, n, first, seg, n_seg, (const uint64_t *)ctx->bm_starts.p, nullptr, nullptr,
                                                                 (const uint64_t *)ctx->bm_base.p, (PgBamRec *)ctx->bm_recs.p);
      CU(cudaGetLastError());
      ctx->bm_launches++;
      CU(cudaEventRecord(ctx->sv_ev1, st));
      CU(cudaMemcpyAsync(recs, ctx->bm_recs.p, sizeof(PgBamRec) * (size_t)total, cudaMemcpyDeviceToHost, st));
   } else {
      CU(cudaEventRecord(ctx->sv_ev1, st));
   }
   CU(cudaStreamSynchronize(st));
   CU(cudaEventElapsedTime(&ctx->bm_last_ms, ctx->sv_ev0, ctx->sv_ev1));
   ctx->bm_n_recs = total;
   ctx->bm_valid = true; ctx->bm_generation++;
   return PG_OK;
}

uint64_t pg_bam_launches(pg_ctx *ctx) { return ctx ? ctx->bm_launches : 0; }
double pg_bam_last_ms(pg_ctx *ctx) { return ctx ? (double)ctx->bm_last_ms : 0.0; }
void pg_bam_last_segments(pg_ctx *ctx, uint32_t *segments, uint32_t *walked_again)
{
   if (!ctx) return;
   if (segments) *segments = ctx->bm_last_segments;
   if (walked_again) *walked_again = ctx->bm_last_redone;
}

// ---- discordant-pair discovery on the indexed stream (row N2, first half) ---------------------------------------------
static_assert(sizeof(pg_rp_query) == sizeof(PgRpQuery) && sizeof(PgRpRow) == sizeof(pg_rp), "rp query / row layout");

int pg_bam_rp_discovery(pg_ctx *ctx, const pg_rp_query *q, pg_rp *rows, uint32_t *src, uint64_t cap, uint64_t *n_rows, pg_rp *inter_rows,
                        uint32_t *inter_src, uint64_t inter_cap, uint64_t *n_inter, uint64_t *n_scanned)
{
   if (!ctx || !q || !n_rows || !n_inter) return fail(ctx, PG_ERR_ARG, "pg_bam_rp_discovery: null argument");
   if (!ctx->bm_valid || ctx->bz_out_len == 0) return fail(ctx, PG_ERR_STATE, "pg_bam_rp_discovery: no indexed stream is resident (pg_bgzf_inflate + pg_bam_index)");
   *n_rows = *n_inter = 0;
   if (n_scanned) *n_scanned = 0;
   const uint64_t n64 = ctx->bm_n_recs;
   if (n64 == 0) return PG_OK;
   if (n64 > 0xFFFFFFF0ull) return fail(ctx, PG_ERR_UNSUPPORTED, "pg_bam_rp_discovery: more than 2^32 records");
   const uint32_t n = (uint32_t)n64;
   CU(cudaSetDevice(ctx->device));
   cudaStream_t st = ctx->stream;
   PgRpQuery qq;
   memcpy(&qq, q, sizeof qq);
   const uint32_t n16 = (n + 15u) / 16u, n_blocks = (n16 + PG_SCAN_THREADS - 1) / PG_SCAN_THREADS;
   int rc;
   if ((rc = ensure(ctx, ctx->bd_cls, (size_t)n16 * 16 * 2)) != PG_OK) return rc;
   if ((rc = ensure(ctx, ctx->bd_off, sizeof(uint32_t) * (size_t)n16 * 16 * 2)) != PG_OK) return rc;
   if ((rc = ensure(ctx, ctx->bd_sums, sizeof(uint32_t) * ((size_t)n_blocks + 1) * 2)) != PG_OK) return rc;
   if ((rc = ensure(ctx, ctx->bd_counters, sizeof(uint32_t) * 2)) != PG_OK) return rc;
   uint8_t *cls_intra = (uint8_t *)ctx->bd_cls.p, *cls_inter = cls_intra + (size_t)n16 * 16;
   uint32_t *off_intra = (uint32_t *)ctx->bd_off.p, *off_inter = off_intra + (size_t)n16 * 16;
   uint32_t *sums_intra = (uint32_t *)ctx->bd_sums.p, *sums_inter = sums_intra + n_blocks + 1;
   const uint32_t init[2] = {n, 0u};
   CU(cudaMemcpyAsync(ctx->bd_counters.p, init, sizeof init, cudaMemcpyHostToDevice, st));
   CU(cudaMemsetAsync(ctx->bd_cls.p, 0, (size_t)n16 * 16 * 2, st));
   const uint8_t *d = (const uint8_t *)ctx->bz_out.p;
   const PgBamRec *recs = (const PgBamRec *)ctx->bm_recs.p;
   if (!ctx->sv_ev0) { CU(cudaEventCreate(&ctx->sv_ev0)); CU(cudaEventCreate(&ctx->sv_ev1)); }
   CU(cudaEventRecord(ctx->sv_ev0, st));
   const uint32_t grid = (n + 255u) / 256u;
   pg_bam_stop_kernel<<<grid, 256, 0, st>>>(recs, n, qq, (uint32_t *)ctx->bd_counters.p);
   pg_bam_rpd_flag_kernel<<<grid, 256, 0, st>>>(d, recs, n, qq, (uint32_t *)ctx->bd_counters.p, cls_intra, cls_inter);
   for (int list = 0; list < 2; list++) {
      const uint4 *c16 = (const uint4 *)(list ? cls_inter : cls_intra);
      uint32_t *sums = list ? sums_inter : sums_intra;
      pg_scan_reduce<<<n_blocks, PG_SCAN_THREADS, 0, st>>>(c16, n16, sums);
      pg_scan_blocks<<<1, 1024, 0, st>>>(sums, n_blocks);
      pg_scan_apply<<<n_blocks, PG_SCAN_THREADS, 0, st>>>(c16, n16, sums, list ? off_inter : off_intra);
   }
   CU(cudaGetLastError());
   ctx->bm_launches += 8;
   uint32_t counters[2], tot_intra = 0, tot_inter = 0;
   CU(cudaMemcpyAsync(counters, ctx->bd_counters.p, sizeof counters, cudaMemcpyDeviceToHost, st));
   CU(cudaMemcpyAsync(&tot_intra, sums_intra + n_blocks, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
   CU(cudaMemcpyAsync(&tot_inter, sums_inter + n_blocks, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
   CU(cudaStreamSynchronize(st));
   *n_rows = tot_intra;
   *n_inter = tot_inter;
   if (n_scanned) *n_scanned = counters[1];
   if (tot_intra > cap || tot_inter > inter_cap || (tot_intra && (!rows || !src)) || (tot_inter && (!inter_rows || !inter_src)))
      return fail(ctx, PG_ERR_RUNS_CAP, "pg_bam_rp_discovery: row buffers too small (n_rows / n_inter hold the required sizes)");
   const size_t tot = (size_t)tot_intra + tot_inter;
   if (tot) {
      if ((rc = ensure(ctx, ctx->bd_rows, sizeof(PgRpRow) * tot)) != PG_OK) return rc;
      if ((rc = ensure(ctx, ctx->bd_src, sizeof(uint32_t) * tot)) != PG_OK) return rc;
      PgRpRow *rows_d = (PgRpRow *)ctx->bd_rows.p;
      uint32_t *src_d = (uint32_t *)ctx->bd_src.p;
      pg_bam_rpd_scatter_kernel<<<grid, 256, 0, st>>>(d, recs, n, qq, cls_intra, cls_inter, off_intra, off_inter, rows_d, src_d, rows_d + tot_intra,
                                                      src_d + tot_intra);
      CU(cudaGetLastError());
      ctx->bm_launches++;
      CU(cudaEventRecord(ctx->sv_ev1, st));
      if (tot_intra) {
         CU(cudaMemcpyAsync(rows, rows_d, sizeof(PgRpRow) * (size_t)tot_intra, cudaMemcpyDeviceToHost, st));
         CU(cudaMemcpyAsync(src, src_d, sizeof(uint32_t) * (size_t)tot_intra, cudaMemcpyDeviceToHost, st));
      }
      if (tot_inter) {
         CU(cudaMemcpyAsync(inter_rows, rows_d + tot_intra, sizeof(PgRpRow) * (size_t)tot_inter, cudaMemcpyDeviceToHost, st));
         CU(cudaMemcpyAsync(inter_src, src_d + tot_intra, sizeof(uint32_t) * (size_t)tot_inter, cudaMemcpyDeviceToHost, st));
      }
   } else {
      CU(cudaEventRecord(ctx->sv_ev1, st));
   }
   CU(cudaStreamSynchronize(st));
   CU(cudaEventElapsedTime(&ctx->bm_last_ms, ctx->sv_ev0, ctx->sv_ev1));
   return PG_OK;
}

// ---- the reader's split-read filter on the indexed stream (rest of row N3) ---------------------------------------------
static_assert(sizeof(pg_sr_query) == sizeof(PgSrQuery) && sizeof(pg_sr_query) == 40, "sr query layout");
static_assert(sizeof(pg_sr_cand) == sizeof(PgSrCand) && sizeof(pg_sr_cand) == 20, "sr candidate layout");
static_assert(sizeof(pg_sr_ref) == sizeof(PgSrRef) && sizeof(pg_sr_ref) == 8, "sr reference read layout");

// The device part of the split-read filter: the four per-record kernels, the two prefix counts and the scatter.  Leaves the
// candidates (ctx->sr_cands), their sequences at a fixed stride (ctx->sr_seqs, NUL-filled behind every sequence) and the
// reference-supporting reads (ctx->sr_refs) on the device and remembers the query, so that the usual size-then-rows pair of
// calls -- and pg_bam_sr_close_end behind pg_bam_sr_candidates -- runs the kernels once per stream and query.
static int sr_filter_device(pg_ctx *ctx, const pg_sr_query *q)
{
   if (!ctx->bm_valid || ctx->bz_out_len == 0) return fail(ctx, PG_ERR_STATE, "no indexed stream is resident (pg_bgzf_inflate + pg_bam_index)");
   if (ctx->sr_cache_valid && ctx->sr_cache_gen == ctx->bm_generation && memcmp(&ctx->sr_cache_q, q, sizeof *q) == 0) return ctx->sr_cache_rc;
   ctx->sr_cache_valid = false;
   ctx->sr_tot = ctx->sr_tot_ref = 0; ctx->sr_stride = 1; ctx->sr_scanned = 0;
   const uint64_t n64 = ctx->bm_n_recs;
   auto remember = [&](int rc) { ctx->sr_cache_valid = true; ctx->sr_cache_gen = ctx->bm_generation; ctx->sr_cache_q = *q; ctx->sr_cache_rc = rc; ctx->sr_batch_ready = false; return rc; };
   if (n64 == 0) return remember(PG_OK);
   if (n64 > 0x3FFFFFF0ull) return fail(ctx, PG_ERR_UNSUPPORTED, "split-read filter: more than 2^30 records");
   const uint32_t n = (uint32_t)n64;
   CU(cudaSetDevice(ctx->device));
   cudaStream_t st = ctx->stream;
   PgSrQuery qq;
   memcpy(&qq, q, sizeof qq);
   const uint32_t n16 = (n + 15u) / 16u, n_blocks = (n16 + PG_SCAN_THREADS - 1) / PG_SCAN_THREADS;
   uint32_t table_bits = 4;
   while ((1ull << table_bits) < 2ull * n) table_bits++;
   const size_t slots = (size_t)1 << table_bits;
   int rc;
   if ((rc = ensure(ctx, ctx->bd_cls, (size_t)n16 * 16 * 2)) != PG_OK) return rc;                       // n_sr | n_ref byte counts
   if ((rc = ensure(ctx, ctx->bd_off, sizeof(uint32_t) * (size_t)n16 * 16 * 2)) != PG_OK) return rc;
   if ((rc = ensure(ctx, ctx->bd_sums, sizeof(uint32_t) * ((size_t)n_blocks + 1) * 2)) != PG_OK) return rc;
   if ((rc = ensure(ctx, ctx->bd_counters, sizeof(uint32_t) * 4)) != PG_OK) return rc;
   if ((rc = ensure(ctx, ctx->sr_info, sizeof(PgSrInfo) * (size_t)n)) != PG_OK) return rc;
   if ((rc = ensure(ctx, ctx->sr_keys, sizeof(uint64_t) * slots)) != PG_OK) return rc;
   if ((rc = ensure(ctx, ctx->sr_head, sizeof(uint32_t) * slots)) != PG_OK) return rc;
   if ((rc = ensure(ctx, ctx->sr_slot, sizeof(uint32_t) * (size_t)n)) != PG_OK) return rc;
   if ((rc = ensure(ctx, ctx->sr_next, sizeof(uint32_t) * (size_t)n)) != PG_OK) return rc;
   if ((rc = ensure(ctx, ctx->sr_partner, sizeof(uint32_t) * (size_t)n)) != PG_OK) return rc;
   if ((rc = ensure(ctx, ctx->sr_emit, (size_t)n)) != PG_OK) return rc;
   uint8_t *cnt_sr = (uint8_t *)ctx->bd_cls.p, *cnt_ref = cnt_sr + (size_t)n16 * 16;
   uint32_t *off_sr = (uint32_t *)ctx->bd_off.p, *off_ref = off_sr + (size_t)n16 * 16;
   uint32_t *sums_sr = (uint32_t *)ctx->bd_sums.p, *sums_ref = sums_sr + n_blocks + 1;
   uint32_t *counters_d = (uint32_t *)ctx->bd_counters.p;
   const uint32_t init[4] = {n, 0u, 0u, 0u};
   CU(cudaMemcpyAsync(counters_d, init, sizeof init, cudaMemcpyHostToDevice, st));
   CU(cudaMemsetAsync(ctx->bd_cls.p, 0, (size_t)n16 * 16 * 2, st));
   CU(cudaMemsetAsync(ctx->sr_keys.p, 0, sizeof(uint64_t) * slots, st));
   CU(cudaMemsetAsync(ctx->sr_head.p, 0xFF, sizeof(uint32_t) * slots, st));                              // PGSR_NONE
   const uint8_t *d = (const uint8_t *)ctx->bz_out.p;
   const PgBamRec *recs = (const PgBamRec *)ctx->bm_recs.p;
   PgSrInfo *info = (PgSrInfo *)ctx->sr_info.p;
   uint32_t *slot = (uint32_t *)ctx->sr_slot.p, *next = (uint32_t *)ctx->sr_next.p, *partner = (uint32_t *)ctx->sr_partner.p;
   uint8_t *emit = (uint8_t *)ctx->sr_emit.p;
   if (!ctx->sv_ev0) { CU(cudaEventCreate(&ctx->sv_ev0)); CU(cudaEventCreate(&ctx->sv_ev1)); }
   CU(cudaEventRecord(ctx->sv_ev0, st));
   const uint32_t grid = (n + 255u) / 256u;
   pg_bam_stop_kernel<<<grid, 256, 0, st>>>(recs, n, pgsr_iterator(qq), counters_d);
   pg_sr_info_kernel<<<grid, 256, 0, st>>>(d, recs, n, qq, counters_d, info);
   pg_sr_insert_kernel<<<grid, 256, 0, st>>>(info, n, (unsigned long long *)ctx->sr_keys.p, (uint32_t *)ctx->sr_head.p, (uint32_t)(slots - 1), slot, next);
   pg_sr_pair_kernel<<<grid, 256, 0, st>>>(d, recs, n, qq, info, (const uint32_t *)ctx->sr_head.p, slot, next, partner, emit, cnt_sr, cnt_ref, counters_d);
   for (int list = 0; list < 2; list++) {
      const uint4 *c16 = (const uint4 *)(list ? cnt_ref : cnt_sr);
      uint32_t *sums = list ? sums_ref : sums_sr;
      pg_scan_reduce<<<n_blocks, PG_SCAN_THREADS, 0, st>>>(c16, n16, sums);
      pg_scan_blocks<<<1, 1024, 0, st>>>(sums, n_blocks);
      pg_scan_apply<<<n_blocks, PG_SCAN_THREADS, 0, st>>>(c16, n16, sums, list ? off_ref : off_sr);
   }
   CU(cudaGetLastError());
   ctx->bm_launches += 10;
   uint32_t counters[4], tot_sr = 0, tot_ref = 0;
   CU(cudaMemcpyAsync(counters, counters_d, sizeof counters, cudaMemcpyDeviceToHost, st));
   CU(cudaMemcpyAsync(&tot_sr, sums_sr + n_blocks, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
   CU(cudaMemcpyAsync(&tot_ref, sums_ref + n_blocks, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
   CU(cudaStreamSynchronize(st));
   ctx->sr_scanned = counters[1];
   if (counters[3] & PGSR_ERR_MALFORMED)
      return remember(fail(ctx, PG_ERR_DATA, "split-read filter: an alignment record of the window is malformed (its fields do not fit its block_size)"));
   if (counters[3] & PGSR_ERR_INSERT)
      return remember(fail(ctx, PG_ERR_DATA, "split-read filter: the insert size is not larger than a candidate read (the reference stops here, src/reader.cpp:866)"));
   const uint32_t str = counters[2] ? counters[2] : 1u;
   ctx->sr_tot = tot_sr; ctx->sr_tot_ref = tot_ref; ctx->sr_stride = str;
   const uint64_t seq_bytes = (uint64_t)tot_sr * str;
   if (tot_sr || tot_ref) {
      if ((rc = ensure(ctx, ctx->sr_cands, sizeof(PgSrCand) * ((size_t)tot_sr + 1))) != PG_OK) return rc;
      if ((rc = ensure(ctx, ctx->sr_seqs, (size_t)seq_bytes + 16)) != PG_OK) return rc;
      if ((rc = ensure(ctx, ctx->sr_refs, sizeof(PgSrRef) * ((size_t)tot_ref + 1))) != PG_OK) return rc;
      CU(cudaMemsetAsync(ctx->sr_seqs.p, 0, (size_t)seq_bytes + 16, st));
      pg_sr_scatter_kernel<<<grid, 256, 0, st>>>(d, recs, n, qq, info, partner, emit, off_sr, off_ref, (PgSrCand *)ctx->sr_cands.p, (uint8_t *)ctx->sr_seqs.p,
                                                 str, (PgSrRef *)ctx->sr_refs.p);
      CU(cudaGetLastError());
      ctx->bm_launches++;
   }
   CU(cudaEventRecord(ctx->sv_ev1, st));
   CU(cudaStreamSynchronize(st));
   CU(cudaEventElapsedTime(&ctx->bm_last_ms, ctx->sv_ev0, ctx->sv_ev1));
   return remember(PG_OK);
}

int pg_bam_sr_candidates(pg_ctx *ctx, const pg_sr_query *q, pg_sr_cand *cands, uint64_t cand_cap, uint64_t *n_cands, char *seqs, uint64_t seq_cap,
                         uint32_t *stride, pg_sr_ref *refs, uint64_t ref_cap, uint64_t *n_refs, uint64_t *n_scanned)
{
   if (!ctx || !q || !n_cands || !n_refs || !stride) return fail(ctx, PG_ERR_ARG, "pg_bam_sr_candidates: null argument");
   *n_cands = *n_refs = 0;
   *stride = 0;
   if (n_scanned) *n_scanned = 0;
   int rc = sr_filter_device(ctx, q);
   if (n_scanned) *n_scanned = ctx->sr_scanned;
   if (rc != PG_OK) return rc;
   const uint32_t tot_sr = ctx->sr_tot, tot_ref = ctx->sr_tot_ref, str = ctx->sr_stride;
   *n_cands = tot_sr; *n_refs = tot_ref; *stride = str;
   const uint64_t seq_bytes = (uint64_t)tot_sr * str;
   if (tot_sr > cand_cap || tot_ref > ref_cap || seq_bytes > seq_cap || (tot_sr && (!cands || !seqs)) || (tot_ref && !refs))
      return fail(ctx, PG_ERR_RUNS_CAP, "pg_bam_sr_candidates: buffers too small (n_cands / n_refs / stride hold the required sizes)");
   cudaStream_t st = ctx->stream;
   if (tot_sr) {
      CU(cudaMemcpyAsync(cands, ctx->sr_cands.p, sizeof(PgSrCand) * (size_t)tot_sr, cudaMemcpyDeviceToHost, st));
      CU(cudaMemcpyAsync(seqs, ctx->sr_seqs.p, (size_t)seq_bytes, cudaMemcpyDeviceToHost, st));
   }
   if (tot_ref) CU(cudaMemcpyAsync(refs, ctx->sr_refs.p, sizeof(PgSrRef) * (size_t)tot_ref, cudaMemcpyDeviceToHost, st));
   CU(cudaStreamSynchronize(st));
   return PG_OK;
}

// The candidates become the resident batch of the search kernel where they lie: warp-free, one thread per candidate for the
// anchor arrays; the sequences are used at their stride (the kernel strips the NUL fill as setUnmatchedSeq strips trailing
// non-alphanumerics), one device-to-device copy into the batch's blob.
__global__ void pg_sr_batch_kernel(uint32_t n, const PgSrCand *__restrict__ cands, uint32_t stride, int16_t insert_size, int32_t chr_id,
                                   uint32_t *__restrict__ seq_off, uint8_t *__restrict__ md, uint32_t *__restrict__ pos, int16_t *__restrict__ isz,
                                   int32_t *__restrict__ chr)
{
   const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i > n) return;
   seq_off[i] = i * stride;
   if (i == n) return;
   const PgSrCand c = cands[i];
   md[i] = c.matched_d;
   pos[i] = c.matched_rel_pos;
   isz[i] = insert_size;
   chr[i] = chr_id;
}

int pg_bam_sr_close_end(pg_ctx *ctx, const pg_sr_query *q, int32_t chr_id, pg_sr_cand *cands, uint64_t cand_cap, uint64_t *n_cands, char *seqs,
                        uint64_t seq_cap, uint32_t *stride, pg_sr_ref *refs, uint64_t ref_cap, uint64_t *n_refs, uint64_t *n_scanned, pg_results *res)
{
   if (!ctx || !q || !n_cands || !n_refs || !stride || !res) return fail(ctx, PG_ERR_ARG, "pg_bam_sr_close_end: null argument");
   if (chr_id < 0 || (size_t)chr_id >= ctx->chroms_h.size()) return fail(ctx, PG_ERR_ARG, "pg_bam_sr_close_end: chr_id out of range (pg_set_reference first)");
   *n_cands = *n_refs = 0;
   *stride = 0;
   res->n_runs = 0;
   if (n_scanned) *n_scanned = 0;
   int rc = sr_filter_device(ctx, q);
   if (n_scanned) *n_scanned = ctx->sr_scanned;
   if (rc != PG_OK) return rc;
   const uint32_t n = ctx->sr_tot, tot_ref = ctx->sr_tot_ref, str = ctx->sr_stride;
   *n_cands = n; *n_refs = tot_ref; *stride = str;
   const uint64_t seq_bytes = (uint64_t)n * str;
   if (seq_bytes >= 0xFFFFFFF0ull) return fail(ctx, PG_ERR_UNSUPPORTED, "pg_bam_sr_close_end: more than 4 GiB of candidate sequence in one window");
   if (str >= PG_TABLE) return fail(ctx, PG_ERR_UNSUPPORTED, "pg_bam_sr_close_end: candidate longer than 499 bases");
   cudaStream_t st = ctx->stream;
   if (n && !(ctx->sr_batch_ready && ctx->sr_batch_chr == chr_id)) {
      ctx->launched = false;
      ctx->stats_valid = false;
      ctx->n = (int)n;
      ctx->have_anchor = true; ctx->have_chr = true; ctx->have_close = false; ctx->have_regions = false;
      ctx->max_read_len = str;
      ctx->blob_bytes = (size_t)seq_bytes;
      if ((rc = ensure(ctx, ctx->seq_blob, (size_t)seq_bytes + 16)) != PG_OK) return rc;
      if ((rc = ensure(ctx, ctx->seq_off, sizeof(uint32_t) * ((size_t)n + 1))) != PG_OK) return rc;
      if ((rc = ensure(ctx, ctx->md, (size_t)n)) != PG_OK) return rc;
      if ((rc = ensure(ctx, ctx->pos, sizeof(uint32_t) * (size_t)n)) != PG_OK) return rc;
      if ((rc = ensure(ctx, ctx->isz, sizeof(int16_t) * (size_t)n)) != PG_OK) return rc;
      if ((rc = ensure(ctx, ctx->chr, sizeof(int32_t) * (size_t)n)) != PG_OK) return rc;
      CU(cudaMemcpyAsync(ctx->seq_blob.p, ctx->sr_seqs.p, (size_t)seq_bytes, cudaMemcpyDeviceToDevice, st));
      pg_sr_batch_kernel<<<(n + 256u) / 256u, 256, 0, st>>>(n, (const PgSrCand *)ctx->sr_cands.p, str, (int16_t)q->insert_size, chr_id, (uint32_t *)ctx->seq_off.p,
                                                           (uint8_t *)ctx->md.p, (uint32_t *)ctx->pos.p, (int16_t *)ctx->isz.p, (int32_t *)ctx->chr.p);
      CU(cudaGetLastError());
      ctx->bm_launches++;
      rc = pg_launch(ctx, 1);
      if (rc != PG_OK) return rc;
      ctx->sr_batch_ready = true;
      ctx->sr_batch_chr = chr_id;
   }
   if (n) {
      rc = finish_launch(ctx);
      if (rc != PG_OK) return rc;
      res->n_runs = ctx->last_runs;
   }
   if (n > cand_cap || tot_ref > ref_cap || seq_bytes > seq_cap || (n && (!cands || !seqs || !res->out)) || (tot_ref && !refs) ||
       res->n_runs > res->runs_cap || (res->n_runs && !res->runs))
      return fail(ctx, PG_ERR_RUNS_CAP, "pg_bam_sr_close_end: buffers too small (n_cands / n_refs / stride / results->n_runs hold the required sizes)");
   if (n) {
      CU(cudaMemcpyAsync(cands, ctx->sr_cands.p, sizeof(PgSrCand) * (size_t)n, cudaMemcpyDeviceToHost, st));
      CU(cudaMemcpyAsync(seqs, ctx->sr_seqs.p, (size_t)seq_bytes, cudaMemcpyDeviceToHost, st));
      CU(cudaMemcpyAsync(res->out, ctx->rs[ctx->cur].out.p, sizeof(PgReadOut) * (size_t)n, cudaMemcpyDeviceToHost, st));
      if (res->n_runs) CU(cudaMemcpyAsync(res->runs, ctx->rs[ctx->cur].pool.p, sizeof(PgRun) * (size_t)res->n_runs, cudaMemcpyDeviceToHost, st));
   }
   if (tot_ref) CU(cudaMemcpyAsync(refs, ctx->sr_refs.p, sizeof(PgSrRef) * (size_t)tot_ref, cudaMemcpyDeviceToHost, st));
   CU(cudaStreamSynchronize(st));
   return PG_OK;
}

// ---- BreakDancer mask and cluster lookup (row N2, last part) ----------------------------------------------------------
static_assert(sizeof(pg_bd_event) == sizeof(PgBdEvent) && sizeof(pg_bd_event) == 24, "bd event layout");

int pg_bd_load_region(pg_ctx *ctx, const pg_bd_event *events, uint32_t n_events, const int32_t *chr_name_rank, int32_t chr, uint32_t start, uint32_t end)
{
   if (!ctx || (n_events && !events) || !chr_name_rank) return fail(ctx, PG_ERR_ARG, "pg_bd_load_region: null argument");
   const int32_t n_chr = (int32_t)ctx->chroms_h.size();
   if (n_chr == 0) return fail(ctx, PG_ERR_STATE, "pg_bd_load_region: no reference (call pg_set_reference)");
   if (chr < 0 || chr >= n_chr || end < start) return fail(ctx, PG_ERR_ARG, "pg_bd_load_region: window out of range");
   CU(cudaSetDevice(ctx->device));
   ctx->bk_loaded = false;
   PgBdPrepared P;
   if (pgbd_prepare(reinterpret_cast<const PgBdEvent *>(events), n_events, chr_name_rank, n_chr, chr, start, end, P) != 0)
      return fail(ctx, PG_ERR_ARG, "pg_bd_load_region: events must be in the reference's order (sortOnFirstBDCoordinate) with chromosome indices of pg_set_reference");
   const size_t m = P.start.size();
   int rc;
   if ((rc = ensure(ctx, ctx->bk_start, sizeof(uint32_t) * (m + 1))) != PG_OK) return rc;
   if ((rc = ensure(ctx, ctx->bk_drop, sizeof(uint32_t) * (m + 1))) != PG_OK) return rc;
   if ((rc = ensure(ctx, ctx->bk_sel, sizeof(PgBdSel) * (m + 1))) != PG_OK) return rc;
   CU(cudaStreamSynchronize(ctx->stream));          // a launch may still read the previous region
   if (m) {
      CU(cudaMemcpy(ctx->bk_start.p, P.start.data(), sizeof(uint32_t) * m, cudaMemcpyHostToDevice));
      CU(cudaMemcpy(ctx->bk_drop.p, P.drop.data(), sizeof(uint32_t) * m, cudaMemcpyHostToDevice));
      CU(cudaMemcpy(ctx->bk_sel.p, P.sel.data(), sizeof(PgBdSel) * m, cudaMemcpyHostToDevice));
   }
   ctx->bk_view.start = (const uint32_t *)ctx->bk_start.p;
   ctx->bk_view.drop = (const uint32_t *)ctx->bk_drop.p;
   ctx->bk_view.sel = (const PgBdSel *)ctx->bk_sel.p;
   ctx->bk_view.m = (uint32_t)m; ctx->bk_view.z = P.z; ctx->bk_view.ws = P.ws; ctx->bk_view.we = P.we;
   ctx->bk_loaded = true;
   return PG_OK;
}

int pg_bd_unload(pg_ctx *ctx)
{
   if (!ctx) return PG_ERR_ARG;
   ctx->bk_loaded = false;
   return PG_OK;
}

int pg_bd_clusters(pg_ctx *ctx, uint32_t n, const uint32_t *close_last_abs_loc, uint32_t *region_off, pg_region *regions, uint64_t cap, uint64_t *n_regions)
{
   if (!ctx || !n_regions || (n && (!close_last_abs_loc || !region_off))) return fail(ctx, PG_ERR_ARG, "pg_bd_clusters: null argument");
   if (!ctx->bk_loaded) return fail(ctx, PG_ERR_STATE, "pg_bd_clusters: no region loaded (pg_bd_load_region)");
   *n_regions = 0;
   if (n == 0) return PG_OK;
   CU(cudaSetDevice(ctx->device));
   int rc;
   if ((rc = ensure(ctx, ctx->bk_pos, sizeof(uint32_t) * (size_t)n)) != PG_OK) return rc;
   CU(cudaMemcpyAsync(ctx->bk_pos.p, close_last_abs_loc, sizeof(uint32_t) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
   uint32_t tot = 0;
   if ((rc = bd_lookup_device(ctx, n, nullptr, (const uint32_t *)ctx->bk_pos.p, &tot)) != PG_OK) return rc;
   *n_regions = tot;
   CU(cudaMemcpyAsync(region_off, ctx->bk_region_off.p, sizeof(uint32_t) * ((size_t)n + 1), cudaMemcpyDeviceToHost, ctx->stream));
   if (tot > cap || (tot && !regions)) {
      CU(cudaStreamSynchronize(ctx->stream));
      return fail(ctx, PG_ERR_RUNS_CAP, "pg_bd_clusters: region buffer too small (n_regions holds the required size)");
   }
   if (tot) CU(cudaMemcpyAsync(regions, ctx->bk_regions.p, sizeof(PgRegion) * (size_t)tot, cudaMemcpyDeviceToHost, ctx->stream));
   CU(cudaStreamSynchronize(ctx->stream));
   return PG_OK;
}

uint64_t pg_bd_launches(pg_ctx *ctx) { return ctx ? ctx->bk_launches : 0; }
uint64_t pg_bd_lookups(pg_ctx *ctx) { return ctx ? ctx->bk_lookups : 0; }

// ---- per-box read ordering and duplicate marking (row N4) -----------------------------------------------------------
static_assert(sizeof(pg_sort_key) == sizeof(PgSortKey) && sizeof(pg_sort_key) == 32, "sort key layout");
static_assert(offsetof(pg_sort_key, matched_d) == offsetof(PgSortKey, matched_d) && offsetof(pg_sort_key, read_length) == offsetof(PgSortKey, read_length), "sort key layout");
static_assert((int)PG_BOX_D == (int)PGS_D && (int)PG_BOX_SI == (int)PGS_SI && (int)PG_BOX_DI == (int)PGS_DI && (int)PG_BOX_INV == (int)PGS_INV &&
              (int)PG_BOX_INV_NT == (int)PGS_INV_NT, "box kinds");
static_assert(PG_BOX_EVENT_GOOD == PGS_EV_GOOD && PG_BOX_EVENT_START == PGS_EV_START, "event flags");

int pg_box_sort(pg_ctx *ctx, uint32_t n_boxes, const uint32_t *box_off, const pg_sort_key *keys, uint32_t *order, uint8_t *nonunique)
{
   return pg_box_sort_kind(ctx, PG_BOX_D, n_boxes, box_off, keys, order, nonunique, nullptr);
}

int pg_box_sort_kind(pg_ctx *ctx, int32_t kind, uint32_t n_boxes, const uint32_t *box_off, const pg_sort_key *keys, uint32_t *order, uint8_t *nonunique, uint8_t *event)
{
   if (!ctx || (n_boxes && (!box_off || !order || !nonunique))) return fail(ctx, PG_ERR_ARG, "pg_box_sort: null argument");
   if (kind < PG_BOX_D || kind > PG_BOX_INV_NT) return fail(ctx, PG_ERR_ARG, "pg_box_sort: unknown kind");
   if (n_boxes == 0) return PG_OK;
   for (uint32_t b = 0; b < n_boxes; b++)
      if (box_off[b + 1] < box_off[b]) return fail(ctx, PG_ERR_ARG, "pg_box_sort: box offsets must be non-decreasing");
   const size_t total = box_off[n_boxes];
   if (total && !keys) return fail(ctx, PG_ERR_ARG, "pg_box_sort: null keys");
   if (total == 0) return PG_OK;
   CU(cudaSetDevice(ctx->device));
   cudaStream_t st = ctx->stream;
   int rc;
   if ((rc = ensure(ctx, ctx->bx_off, sizeof(uint32_t) * ((size_t)n_boxes + 1))) != PG_OK) return rc;
   if ((rc = ensure(ctx, ctx->bx_keys, sizeof(PgSortKey) * total)) != PG_OK) return rc;
   if ((rc = ensure(ctx, ctx->bx_order, sizeof(uint32_t) * total)) != PG_OK) return rc;
   if ((rc = ensure(ctx, ctx->bx_non, total)) != PG_OK) return rc;
   if (event && (rc = ensure(ctx, ctx->bx_event, total)) != PG_OK) return rc;
   CU(cudaMemcpyAsync(ctx->bx_off.p, box_off, sizeof(uint32_t) * ((size_t)n_boxes + 1), cudaMemcpyHostToDevice, st));
   CU(cudaMemcpyAsync(ctx->bx_keys.p, keys, sizeof(PgSortKey) * total, cudaMemcpyHostToDevice, st));
   if (!ctx->sv_ev0) { CU(cudaEventCreate(&ctx->sv_ev0)); CU(cudaEventCreate(&ctx->sv_ev1)); }
   CU(cudaEventRecord(ctx->sv_ev0, st));
   pg_box_sort_warp_kernel<<<(n_boxes + 7) / 8, 256, 0, st>>>(n_boxes, (const uint32_t *)ctx->bx_off.p, (const PgSortKey *)ctx->bx_keys.p,
                                                             (uint32_t *)ctx->bx_order.p, (uint8_t *)ctx->bx_non.p, (int)kind,
                                                             event ? (uint8_t *)ctx->bx_event.p : nullptr);
   CU(cudaGetLastError());
   CU(cudaEventRecord(ctx->sv_ev1, st));
   ctx->bx_launches++;
   CU(cudaMemcpyAsync(order, ctx->bx_order.p, sizeof(uint32_t) * total, cudaMemcpyDeviceToHost, st));
   CU(cudaMemcpyAsync(nonunique, ctx->bx_non.p, total, cudaMemcpyDeviceToHost, st));
   if (event) CU(cudaMemcpyAsync(event, ctx->bx_event.p, total, cudaMemcpyDeviceToHost, st));
   CU(cudaStreamSynchronize(st));
   CU(cudaEventElapsedTime(&ctx->bx_last_ms, ctx->sv_ev0, ctx->sv_ev1));
   return PG_OK;
}

uint64_t pg_box_sort_launches(pg_ctx *ctx) { return ctx ? ctx->bx_launches : 0; }
double pg_box_sort_last_kernel_ms(pg_ctx *ctx) { return ctx ? (double)ctx->bx_last_ms : 0.0; }

// ---- read-pair clustering (row N2) --------------------------------------------------------------------------------
static_assert(sizeof(pg_rp) == sizeof(PgRp) && sizeof(pg_rp) == 16, "rp layout");
static_assert(sizeof(pg_rp_cluster) == sizeof(PgRpCluster) && sizeof(pg_rp_cluster) == 28, "rp cluster layout");
#define PG_RP_SMEM_ROWS 12800u   // 200 KB of intervals per CTA

int pg_rp_modify(pg_ctx *ctx, int32_t n, const pg_rp *in, pg_rp_cluster *out)
{
   if (!ctx || n < 0 || (n > 0 && (!in || !out))) return fail(ctx, PG_ERR_ARG, "pg_rp_modify: bad arguments");
   if (n == 0) return PG_OK;
   CU(cudaSetDevice(ctx->device));
   cudaStream_t st = ctx->stream;
   std::vector<uint32_t> rows_small, off_small, rows_big, off_big;
   // classes of up to 1024 rows take the warp kernel, larger ones the CTA kernel; PG_RP_WARP_MAX lowers the limit so that a
   // small input exercises the CTA kernel (compute-sanitizer racecheck slows its barrier loop by orders of magnitude)
   uint32_t warp_max = 1024u;
   if (const char *e = getenv("PG_RP_WARP_MAX")) { const unsigned long v = strtoul(e, nullptr, 10); if (v >= 1 && v < 1024) warp_max = (uint32_t)v; }
   pg_rp_classes(n, reinterpret_cast<const PgRp *>(in), warp_max, rows_small, off_small, rows_big, off_big);
   const uint32_t n_small = (uint32_t)off_small.size() - 1, n_big = (uint32_t)off_big.size() - 1;
   const size_t rs = rows_small.size(), rb = rows_big.size();
   // device layout: [small class rows | big class rows], [small offsets | big offsets]
   int rc;
   if ((rc = ensure(ctx, ctx->rp_in, sizeof(PgRp) * (size_t)n)) != PG_OK) return rc;
   if ((rc = ensure(ctx, ctx->rp_out, sizeof(PgRpCluster) * (size_t)n)) != PG_OK) return rc;
   if ((rc = ensure(ctx, ctx->rp_rows, sizeof(uint32_t) * (rs + rb + 1))) != PG_OK) return rc;
   if ((rc = ensure(ctx, ctx->rp_off, sizeof(uint32_t) * ((size_t)n_small + n_big + 2))) != PG_OK) return rc;
   if ((rc = ensure(ctx, ctx->rp_scratch, sizeof(uint4) * (rs + rb + 1))) != PG_OK) return rc;
   CU(cudaMemcpyAsync(ctx->rp_in.p, in, sizeof(PgRp) * (size_t)n, cudaMemcpyHostToDevice, st));
   uint32_t *rows_d = (uint32_t *)ctx->rp_rows.p, *off_d = (uint32_t *)ctx->rp_off.p;
   if (rs) CU(cudaMemcpyAsync(rows_d, rows_small.data(), sizeof(uint32_t) * rs, cudaMemcpyHostToDevice, st));
   if (rb) CU(cudaMemcpyAsync(rows_d + rs, rows_big.data(), sizeof(uint32_t) * rb, cudaMemcpyHostToDevice, st));
   CU(cudaMemcpyAsync(off_d, off_small.data(), sizeof(uint32_t) * ((size_t)n_small + 1), cudaMemcpyHostToDevice, st));
   CU(cudaMemcpyAsync(off_d + n_small + 1, off_big.data(), sizeof(uint32_t) * ((size_t)n_big + 1), cudaMemcpyHostToDevice, st));
   if (!ctx->sv_ev0) { CU(cudaEventCreate(&ctx->sv_ev0)); CU(cudaEventCreate(&ctx->sv_ev1)); }
   CU(cudaEventRecord(ctx->sv_ev0, st));
   pg_rp_single_kernel<<<(n + 255) / 256, 256, 0, st>>>(n, (const PgRp *)ctx->rp_in.p, (PgRpCluster *)ctx->rp_out.p);
   ctx->rp_launches++;
   if (n_small) {
      pg_rp_modify_warp_kernel<<<(n_small + 7) / 8, 256, 0, st>>>((const PgRp *)ctx->rp_in.p, rows_d, off_d, n_small, (uint4 *)ctx->rp_scratch.p,
                                                                 (PgRpCluster *)ctx->rp_out.p);
      ctx->rp_launches++;
   }
   if (n_big) {
      uint32_t biggest = 0;
      for (uint32_t g = 0; g < n_big; g++) biggest = off_big[g + 1] - off_big[g] > biggest ? off_big[g + 1] - off_big[g] : biggest;
      const uint32_t smem_rows = biggest < PG_RP_SMEM_ROWS ? biggest : PG_RP_SMEM_ROWS;
      CU(cudaFuncSetAttribute(pg_rp_modify_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(uint4) * PG_RP_SMEM_ROWS)));
      pg_rp_modify_kernel<<<n_big, PG_RP_THREADS, sizeof(uint4) * (size_t)smem_rows, st>>>((const PgRp *)ctx->rp_in.p, rows_d + rs, off_d + n_small + 1,
                                                                                         (uint4 *)ctx->rp_scratch.p + rs, smem_rows,
                                                                                         (PgRpCluster *)ctx->rp_out.p);
      ctx->rp_launches++;
   }
   CU(cudaGetLastError());
   CU(cudaEventRecord(ctx->sv_ev1, st));
   ctx->rp_classes_small = n_small;
   ctx->rp_classes_big = n_big;
   ctx->rp_rows_in_classes = (uint32_t)(rs + rb);
   CU(cudaMemcpyAsync(out, ctx->rp_out.p, sizeof(PgRpCluster) * (size_t)n, cudaMemcpyDeviceToHost, st));
   CU(cudaStreamSynchronize(st));
   CU(cudaEventElapsedTime(&ctx->rp_last_ms, ctx->sv_ev0, ctx->sv_ev1));
   return PG_OK;
}

int pg_rp_summarize(pg_ctx *ctx, int32_t n, const pg_rp *in, pg_rp_cluster *inout)
{
   if (!ctx || n < 0 || (n > 0 && (!in || !inout))) return fail(ctx, PG_ERR_ARG, "pg_rp_summarize: bad arguments");
   if (n < 5) {   // src/bddata.cpp:474-478
      for (int i = 0; i < n; i++) inout[i].report = 0;
      return PG_OK;
   }
   CU(cudaSetDevice(ctx->device));
   cudaStream_t st = ctx->stream;
   int rc;
   if ((rc = ensure(ctx, ctx->rp_in, sizeof(PgRp) * (size_t)n)) != PG_OK) return rc;
   if ((rc = ensure(ctx, ctx->rp_out, sizeof(PgRpCluster) * (size_t)n)) != PG_OK) return rc;
   if ((rc = ensure(ctx, ctx->rp_merged, sizeof(uint32_t) * (size_t)n)) != PG_OK) return rc;
   if ((rc = ensure(ctx, ctx->rp_nid, sizeof(uint32_t) * (size_t)n)) != PG_OK) return rc;
   if ((rc = ensure(ctx, ctx->rp_counters, sizeof(uint32_t) * 2)) != PG_OK) return rc;
   if ((rc = ensure(ctx, ctx->rp_keys, sizeof(uint64_t) * (size_t)n)) != PG_OK) return rc;
   CU(cudaMemcpyAsync(ctx->rp_in.p, in, sizeof(PgRp) * (size_t)n, cudaMemcpyHostToDevice, st));
   CU(cudaMemcpyAsync(ctx->rp_out.p, inout, sizeof(PgRpCluster) * (size_t)n, cudaMemcpyHostToDevice, st));
   CU(cudaMemsetAsync(ctx->rp_nid.p, 0, sizeof(uint32_t) * (size_t)n, st));
   CU(cudaMemsetAsync(ctx->rp_counters.p, 0, sizeof(uint32_t) * 2, st));
   if (!ctx->sv_ev0) { CU(cudaEventCreate(&ctx->sv_ev0)); CU(cudaEventCreate(&ctx->sv_ev1)); }
   CU(cudaEventRecord(ctx->sv_ev0, st));
   pg_rp_key_kernel<<<(n + 255) / 256, 256, 0, st>>>(n, (const PgRp *)ctx->rp_in.p, (const PgRpCluster *)ctx->rp_out.p, (uint64_t *)ctx->rp_keys.p);
   pg_rp_group_kernel<<<(n + 7) / 8, 256, 0, st>>>(n, (const PgRp *)ctx->rp_in.p, (const PgRpCluster *)ctx->rp_out.p,
                                                   (const uint64_t *)ctx->rp_keys.p, (uint32_t *)ctx->rp_merged.p, (uint32_t *)ctx->rp_nid.p, (uint32_t *)ctx->rp_counters.p);
   pg_rp_report_kernel<<<(n + 255) / 256, 256, 0, st>>>(n, (PgRpCluster *)ctx->rp_out.p, (const uint32_t *)ctx->rp_merged.p,
                                                       (const uint32_t *)ctx->rp_nid.p, (const uint32_t *)ctx->rp_counters.p);
   CU(cudaGetLastError());
   CU(cudaEventRecord(ctx->sv_ev1, st));
   ctx->rp_launches += 3;
   CU(cudaMemcpyAsync(inout, ctx->rp_out.p, sizeof(PgRpCluster) * (size_t)n, cudaMemcpyDeviceToHost, st));
   CU(cudaStreamSynchronize(st));
   CU(cudaEventElapsedTime(&ctx->rp_last_ms, ctx->sv_ev0, ctx->sv_ev1));
   return PG_OK;
}

int pg_rp_modify_interchr(pg_ctx *ctx, int32_t n, const pg_rp *in, const int32_t *chr_a, const int32_t *chr_b, pg_rp_cluster *out)
{
   if (!ctx || n < 0 || (n > 0 && (!in || !out || !chr_a || !chr_b))) return fail(ctx, PG_ERR_ARG, "pg_rp_modify_interchr: bad arguments");
   if (n == 0) return PG_OK;
   CU(cudaSetDevice(ctx->device));
   cudaStream_t st = ctx->stream;
   std::vector<uint8_t> flip;
   std::vector<uint32_t> rows, off;
   pg_rp_classes_interchr(n, reinterpret_cast<const PgRp *>(in), chr_a, chr_b, flip, rows, off);
   const uint32_t n_classes = (uint32_t)off.size() - 1;
   int rc;
   if ((rc = ensure(ctx, ctx->rp_in, sizeof(PgRp) * (size_t)n)) != PG_OK) return rc;
   if ((rc = ensure(ctx, ctx->rp_out, sizeof(PgRpCluster) * (size_t)n)) != PG_OK) return rc;
   if ((rc = ensure(ctx, ctx->rp_flip, (size_t)n)) != PG_OK) return rc;
   if ((rc = ensure(ctx, ctx->rp_rows, sizeof(uint32_t) * (rows.size() + 1))) != PG_OK) return rc;
   if ((rc = ensure(ctx, ctx->rp_off, sizeof(uint32_t) * ((size_t)n_classes + 1))) != PG_OK) return rc;
   if ((rc = ensure(ctx, ctx->rp_scratch, sizeof(uint4) * (rows.size() + 1))) != PG_OK) return rc;
   CU(cudaMemcpyAsync(ctx->rp_in.p, in, sizeof(PgRp) * (size_t)n, cudaMemcpyHostToDevice, st));
   CU(cudaMemcpyAsync(ctx->rp_flip.p, flip.data(), (size_t)n, cudaMemcpyHostToDevice, st));
   if (!rows.empty()) CU(cudaMemcpyAsync(ctx->rp_rows.p, rows.data(), sizeof(uint32_t) * rows.size(), cudaMemcpyHostToDevice, st));
   CU(cudaMemcpyAsync(ctx->rp_off.p, off.data(), sizeof(uint32_t) * ((size_t)n_classes + 1), cudaMemcpyHostToDevice, st));
   if (!ctx->sv_ev0) { CU(cudaEventCreate(&ctx->sv_ev0)); CU(cudaEventCreate(&ctx->sv_ev1)); }
   CU(cudaEventRecord(ctx->sv_ev0, st));
   pg_rp_inter_single_kernel<<<(n + 255) / 256, 256, 0, st>>>(n, (const PgRp *)ctx->rp_in.p, (PgRpCluster *)ctx->rp_out.p);
   ctx->rp_launches++;
   if (n_classes) {
      pg_rp_modify_inter_warp_kernel<<<(n_classes + 7) / 8, 256, 0, st>>>((const PgRp *)ctx->rp_in.p, (const uint8_t *)ctx->rp_flip.p,
                                                                         (const uint32_t *)ctx->rp_rows.p, (const uint32_t *)ctx->rp_off.p, n_classes,
                                                                         (uint32_t)(n - 1), (uint4 *)ctx->rp_scratch.p, (PgRpCluster *)ctx->rp_out.p);
      ctx->rp_launches++;
   }
   CU(cudaGetLastError());
   CU(cudaEventRecord(ctx->sv_ev1, st));
   ctx->rp_classes_small = n_classes;
   ctx->rp_classes_big = 0;
   ctx->rp_rows_in_classes = (uint32_t)rows.size();
   CU(cudaMemcpyAsync(out, ctx->rp_out.p, sizeof(PgRpCluster) * (size_t)n, cudaMemcpyDeviceToHost, st));
   CU(cudaStreamSynchronize(st));
   CU(cudaEventElapsedTime(&ctx->rp_last_ms, ctx->sv_ev0, ctx->sv_ev1));
   return PG_OK;
}

int pg_rp_summarize_interchr(pg_ctx *ctx, int32_t n, const pg_rp *in, const int32_t *chr_a, const int32_t *chr_b, pg_rp_cluster *inout)
{
   if (!ctx || n < 0 || (n > 0 && (!in || !inout || !chr_a || !chr_b))) return fail(ctx, PG_ERR_ARG, "pg_rp_summarize_interchr: bad arguments");
   if (n == 0) return PG_OK;
   CU(cudaSetDevice(ctx->device));
   cudaStream_t st = ctx->stream;
   int rc;
   if ((rc = ensure(ctx, ctx->rp_in, sizeof(PgRp) * (size_t)n)) != PG_OK) return rc;
   if ((rc = ensure(ctx, ctx->rp_out, sizeof(PgRpCluster) * (size_t)n)) != PG_OK) return rc;
   if ((rc = ensure(ctx, ctx->rp_chr, sizeof(int32_t) * 2 * (size_t)n)) != PG_OK) return rc;
   if ((rc = ensure(ctx, ctx->rp_merged, sizeof(uint32_t) * (size_t)n)) != PG_OK) return rc;
   if ((rc = ensure(ctx, ctx->rp_nid, sizeof(uint32_t) * (size_t)n)) != PG_OK) return rc;
   if ((rc = ensure(ctx, ctx->rp_keys, sizeof(uint64_t) * (size_t)n)) != PG_OK) return rc;
   int32_t *ca = (int32_t *)ctx->rp_chr.p, *cb = ca + n;
   CU(cudaMemcpyAsync(ctx->rp_in.p, in, sizeof(PgRp) * (size_t)n, cudaMemcpyHostToDevice, st));
   CU(cudaMemcpyAsync(ctx->rp_out.p, inout, sizeof(PgRpCluster) * (size_t)n, cudaMemcpyHostToDevice, st));
   CU(cudaMemcpyAsync(ca, chr_a, sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice, st));
   CU(cudaMemcpyAsync(cb, chr_b, sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice, st));
   CU(cudaMemsetAsync(ctx->rp_nid.p, 0, sizeof(uint32_t) * (size_t)n, st));
   if (!ctx->sv_ev0) { CU(cudaEventCreate(&ctx->sv_ev0)); CU(cudaEventCreate(&ctx->sv_ev1)); }
   CU(cudaEventRecord(ctx->sv_ev0, st));
   pg_rp_inter_key_kernel<<<(n + 255) / 256, 256, 0, st>>>(n, (const PgRp *)ctx->rp_in.p, ca, cb, (const PgRpCluster *)ctx->rp_out.p, (uint64_t *)ctx->rp_keys.p);
   pg_rp_inter_group_kernel<<<(n + 7) / 8, 256, 0, st>>>(n, (const PgRp *)ctx->rp_in.p, ca, cb, (const PgRpCluster *)ctx->rp_out.p,
                                                         (const uint64_t *)ctx->rp_keys.p, (uint32_t *)ctx->rp_merged.p, (uint32_t *)ctx->rp_nid.p);
   pg_rp_inter_report_kernel<<<(n + 255) / 256, 256, 0, st>>>(n, (PgRpCluster *)ctx->rp_out.p, (const uint32_t *)ctx->rp_merged.p,
                                                             (const uint32_t *)ctx->rp_nid.p);
   CU(cudaGetLastError());
   CU(cudaEventRecord(ctx->sv_ev1, st));
   ctx->rp_launches += 3;
   CU(cudaMemcpyAsync(inout, ctx->rp_out.p, sizeof(PgRpCluster) * (size_t)n, cudaMemcpyDeviceToHost, st));
   CU(cudaStreamSynchronize(st));
   CU(cudaEventElapsedTime(&ctx->rp_last_ms, ctx->sv_ev0, ctx->sv_ev1));
   return PG_OK;
}

uint64_t pg_rp_launches(pg_ctx *ctx) { return ctx ? ctx->rp_launches : 0; }
void pg_rp_last_classes(pg_ctx *ctx, uint32_t *warp_classes, uint32_t *cta_classes, uint32_t *rows_in_classes)
{
   if (!ctx) return;
   if (warp_classes) *warp_classes = ctx->rp_classes_small;
   if (cta_classes) *cta_classes = ctx->rp_classes_big;
   if (rows_in_classes) *rows_in_classes = ctx->rp_rows_in_classes;
}
double pg_rp_last_kernel_ms(pg_ctx *ctx) { return ctx ? (double)ctx->rp_last_ms : 0.0; }

uint32_t pg_compress_ups(const pg_up *ups, uint32_t n, pg_run *out) { return pgh_compress_ups(ups, n, out); }

} // extern "C"
